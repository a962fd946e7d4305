"""Koopman machines — forward/forecast path of /root/reference/model.py on the GPU.

Same class names and call conventions as the reference (``GenericKM``, ``SparseKM`` alias, ``LISTAKM``,
``make_model``; ``encode / step_latent / decode / kmatrix``; ``load_state_dict`` with the reference's key names,
SURVEY.md §8a), but inference only: training losses, eigenvalue metrics and ODE rollouts (model.py:337-693) are
out of the hot path and not provided.  All math runs in libkmpc (csrc/forecast.cu + the GEMM kernels);
parameters are float32 CUDA tensors.

Additional batched entry points used by the backtest path:
  ``forecast_series``  all rebalancing steps of many paths at once, reading the delay window in place
  ``forecast_env``     the same for the test split of a FinanceEnv (backtest.py:85-121 for every t)
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional

import numpy as np

from . import _capi

_ACT = {"relu": 0, "tanh": 1, "gelu": 2}
_NORM = {"id": 0, "ball": 1}


def _get(cfg, path, default=None):
    cur = cfg
    for p in path.split("."):
        if cur is None or not hasattr(cur, p):
            return default
        cur = getattr(cur, p)
    return cur


def shrink(x, threshold: float):
    """Soft thresholding (model.py:30-40) on a torch tensor."""
    import torch
    return torch.sign(x) * torch.maximum(torch.abs(x) - threshold, torch.zeros_like(x))


class KoopmanMachine:
    """Base class (model.py:216-270).  Sub-classes fill ``self._params`` (reference state_dict keys)."""

    def __init__(self, cfg, observation_size: int, device="cuda"):
        import torch
        self.cfg = cfg
        self.observation_size = int(observation_size)
        self.target_size = int(_get(cfg, "MODEL.TARGET_SIZE"))
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("koopman_mpc_portfolio_rebalancing_b200 models run on CUDA devices only (no CPU fallback)")
        self._params: Dict[str, "torch.Tensor"] = {}
        self._native = {}          # (n_assets, delay) -> kmpc_model*
        self.training = False

    # -- torch.nn.Module look-alikes used by callers of the reference (backtest.py:91-92, run_experiment.py) ----
    def eval(self):
        self.training = False
        return self

    def to(self, device):
        return self

    def parameters(self):
        return iter(self._params.values())

    def state_dict(self):
        return dict(self._params)

    def load_state_dict(self, sd, strict: bool = True):
        import torch
        missing = [k for k in self._params if k not in sd]
        unexpected = [k for k in sd if k not in self._params]
        if strict and (missing or unexpected):
            raise RuntimeError(f"Error(s) in loading state_dict: missing {missing}, unexpected {unexpected}")
        for k, v in sd.items():
            if k not in self._params:
                continue
            t = torch.as_tensor(np.ascontiguousarray(v) if isinstance(v, np.ndarray) else v)
            t = t.to(device=self.device, dtype=torch.float32).contiguous()
            if tuple(t.shape) != tuple(self._params[k].shape):
                raise RuntimeError(f"size mismatch for {k}: {tuple(t.shape)} vs {tuple(self._params[k].shape)}")
            self._params[k] = t
        self._drop_native()
        return self

    def _drop_native(self):
        for m in self._native.values():
            _capi.lib().kmpc_model_free(m)
        self._native = {}

    def __del__(self):
        try:
            self._drop_native()
        except Exception:
            pass

    # -- native model ---------------------------------------------------------------------------------------------
    def _desc(self, n_assets: int, delay: int):
        raise NotImplementedError

    def native(self, n_assets: Optional[int] = None, delay: Optional[int] = None, raw_step: bool = False):
        """kmpc_model* for a given split of observation_size into (n_assets, delay).  encode/decode do not depend
        on the split; the in-place window forecast does.  ``raw_step``: the same weights with the latent
        normalisation switched off (``z @ kmat`` of rollout_latent_discrete, model.py:527-556)."""
        if n_assets is None:
            n_assets, delay = self.observation_size, 1
        key = (int(n_assets), int(delay), bool(raw_step))
        if key not in self._native:
            if key[0] * key[1] != self.observation_size:
                raise ValueError(f"n_assets*delay = {key[0] * key[1]} != observation_size {self.observation_size}")
            desc, keep = self._desc(key[0], key[1])
            if raw_step:
                desc.norm_fn = 0
            out = C.c_void_p()
            h = _capi.Handle.get(self.device.index or 0)
            _capi.check(_capi.lib().kmpc_model_load(h.ptr, C.byref(desc), C.byref(out)))
            del keep
            self._native[key] = out
        return self._native[key]

    def _handle(self):
        return _capi.Handle.get(self.device.index or 0)

    def _rows(self, x, width):
        import torch
        x = torch.as_tensor(x).to(device=self.device, dtype=torch.float32)
        if x.shape[-1] != width:
            raise ValueError(f"last dimension {x.shape[-1]} != {width}")
        lead = x.shape[:-1]
        return x.reshape(-1, width).contiguous(), lead

    # -- reference API ----------------------------------------------------------------------------------------------
    def encode(self, x):
        """Observations [..., observation_size] -> latent [..., target_size] (model.py:239-248)."""
        import torch
        x2, lead = self._rows(x, self.observation_size)
        out = torch.empty((x2.shape[0], self.target_size), dtype=torch.float32, device=self.device)
        if x2.shape[0]:
            _capi.check(_capi.lib().kmpc_encode(self._handle().ptr, self.native(), _capi.ptr(x2), x2.shape[0], _capi.ptr(out),
                                                _capi.stream_ptr(self.device.index or 0)))
        return out.reshape(*lead, self.target_size)

    def step_latent(self, y):
        """z @ K (+ norm for GenericKM 'ball') (model.py:311-321, 787-797)."""
        import torch
        y2, lead = self._rows(y, self.target_size)
        out = torch.empty_like(y2)
        if y2.shape[0]:
            _capi.check(_capi.lib().kmpc_step_latent(self._handle().ptr, self.native(), _capi.ptr(y2), y2.shape[0],
                                                     _capi.ptr(out), _capi.stream_ptr(self.device.index or 0)))
        return out.reshape(*lead, self.target_size)

    def rollout_latent_discrete(self, z0, num_steps: int):
        """z_{t+k} = z_0 K^k WITHOUT the latent normalisation (model.py:527-556): [batch, num_steps+1, target_size],
        z0 at index 0."""
        import torch
        z, lead = self._rows(z0, self.target_size)
        traj = [z]
        for _ in range(int(num_steps)):
            nxt = torch.empty_like(z)
            if z.shape[0]:
                _capi.check(_capi.lib().kmpc_step_latent(self._handle().ptr, self.native(raw_step=True), _capi.ptr(z), z.shape[0],
                                                         _capi.ptr(nxt), _capi.stream_ptr(self.device.index or 0)))
            z = nxt
            traj.append(z)
        return torch.stack(traj, dim=1)

    def rollout_sequence(self, x0, num_steps: int):
        """encode once, roll the latent state, decode every state (model.py:558-585): [batch, num_steps+1, observation_size];
        index 0 is the reconstruction of x0."""
        z_traj = self.rollout_latent_discrete(self.encode(x0), num_steps)
        b, n, _ = z_traj.shape
        return self.decode(z_traj.reshape(b * n, self.target_size)).reshape(b, n, self.observation_size)

    def decode(self, y):
        """latent [..., target_size] -> observations [..., observation_size] (model.py:250-259)."""
        import torch
        y2, lead = self._rows(y, self.target_size)
        out = torch.empty((y2.shape[0], self.observation_size), dtype=torch.float32, device=self.device)
        if y2.shape[0]:
            _capi.check(_capi.lib().kmpc_decode(self._handle().ptr, self.native(), _capi.ptr(y2), y2.shape[0], _capi.ptr(out),
                                                _capi.stream_ptr(self.device.index or 0)))
        return out.reshape(*lead, self.observation_size)

    def kmatrix(self):
        return self._params["kmat"]

    def step_env(self, x):
        return self.decode(self.step_latent(self.encode(x)))

    def reconstruction(self, x):
        return self.decode(self.encode(x))

    def rollout(self, x0, horizon: int, n_cols: Optional[int] = None):
        """evaluation.rollout_no_reencode (evaluation.py:44-74): [horizon, batch, n_cols] standardised predictions."""
        import torch
        x2, lead = self._rows(x0, self.observation_size)
        n_cols = self.observation_size if n_cols is None else int(n_cols)
        out = torch.empty((x2.shape[0], horizon, n_cols), dtype=torch.float32, device=self.device)
        _capi.check(_capi.lib().kmpc_rollout(self._handle().ptr, self.native(), _capi.ptr(x2), x2.shape[0], horizon, n_cols,
                                             _capi.ptr(out), _capi.stream_ptr(self.device.index or 0)))
        return out.permute(1, 0, 2).contiguous()

    # -- batched forecast of the backtest path -------------------------------------------------------------------------
    def forecast_series(self, z, mean, std, n_assets: int, delay: int, row0: int, t0: int, t1: int, horizon: int,
                        out=None):
        """z [B,T,ld] float32 CUDA standardised series (ld = n_assets rounded up to 4), mean/std float64 CUDA
        [N] or [B,N].  Returns yhat [B, t1-t0, H, N] float32: de-standardised k-step-ahead log-return forecasts
        of embedded rows row0+t0 .. row0+t1-1 (the loop at backtest.py:99-121, for every row at once)."""
        import torch
        B, T, ld = z.shape
        per_path = mean.dim() == 2
        if out is None:
            out = torch.empty((B, t1 - t0, horizon, n_assets), dtype=torch.float32, device=self.device)
        _capi.check(_capi.lib().kmpc_forecast(
            self._handle().ptr, self.native(n_assets, delay), _capi.ptr(z), ld, _capi.ptr(mean), _capi.ptr(std),
            int(per_path), B, T, row0, t0, t1, horizon, _capi.ptr(out), _capi.stream_ptr(self.device.index or 0)))
        return out

    def forecast_env(self, env, t0: int, t1: int, horizon: int):
        """yhat [t1-t0, H, N] for test rows t0..t1-1 of a FinanceEnv of this package."""
        z, mean, std = env.series_device()
        y = self.forecast_series(z.unsqueeze(0), mean, std, env.n_assets, env.embedding_dim, env.test_row0, t0, t1, horizon)
        return y[0]


def _linear_init(gen, out_f, in_f, bias, device):
    import torch
    k = 1.0 / np.sqrt(in_f)
    w = (torch.rand((out_f, in_f), generator=gen, device="cpu") * 2 - 1) * k
    b = ((torch.rand((out_f,), generator=gen, device="cpu") * 2 - 1) * k) if bias else None
    return w.to(device), (b.to(device) if b is not None else None)


def _mlp_params(params, prefix, dims, bias, gen, device):
    for li in range(len(dims) - 1):
        w, b = _linear_init(gen, dims[li + 1], dims[li], bias, device)
        params[f"{prefix}.network.{2 * li}.weight"] = w
        if b is not None:
            params[f"{prefix}.network.{2 * li}.bias"] = b


def _mlp_desc_arrays(params, prefix, dims):
    n = len(dims) - 1
    wp = (C.c_void_p * n)(*[params[f"{prefix}.network.{2 * i}.weight"].data_ptr() for i in range(n)])
    bp = (C.c_void_p * n)(*[(params[f"{prefix}.network.{2 * i}.bias"].data_ptr()
                             if f"{prefix}.network.{2 * i}.bias" in params else None) for i in range(n)])
    dm = (C.c_int32 * (n + 1))(*dims)
    return n, dm, wp, bp


class GenericKM(KoopmanMachine):
    """Koopman autoencoder with MLP encoder/decoder (model.py:701-797).  SparseKM is the same class."""

    def __init__(self, cfg, observation_size: int, device="cuda", seed: int = 0):
        import torch
        super().__init__(cfg, observation_size, device)
        Z = self.target_size
        self.enc_layers: List[int] = list(_get(cfg, "MODEL.ENCODER.LAYERS", []))
        self.dec_layers: List[int] = list(_get(cfg, "MODEL.DECODER.LAYERS", []))
        self.enc_bias = bool(_get(cfg, "MODEL.ENCODER.USE_BIAS", False))
        self.dec_bias = bool(_get(cfg, "MODEL.DECODER.USE_BIAS", False))
        self.enc_act = _get(cfg, "MODEL.ENCODER.ACTIVATION", "relu")
        self.dec_act = _get(cfg, "MODEL.DECODER.ACTIVATION", "relu")
        self.last_relu = bool(_get(cfg, "MODEL.ENCODER.LAST_RELU", False))
        self.norm_fn_name = _get(cfg, "MODEL.NORM_FN", "id")
        for name in (self.enc_act, self.dec_act):
            if name not in _ACT:
                raise ValueError(f"Unknown activation '{name}'. Available: {list(_ACT)}")
        if self.norm_fn_name not in _NORM:
            raise ValueError(f"Unknown norm function '{self.norm_fn_name}'")
        gen = torch.Generator().manual_seed(seed)
        _mlp_params(self._params, "encoder", [self.observation_size] + self.enc_layers + [Z], self.enc_bias, gen, self.device)
        _mlp_params(self._params, "decoder", [Z] + self.dec_layers + [self.observation_size], self.dec_bias, gen, self.device)
        self._params["kmat"] = torch.eye(Z, dtype=torch.float32, device=self.device)     # model.py:736

    def _desc(self, n_assets, delay):
        d = _capi.ModelDesc()
        d.kind, d.obs, d.n_assets, d.delay, d.latent = 0, self.observation_size, n_assets, delay, self.target_size
        d.norm_fn = _NORM[self.norm_fn_name]
        ne, edm, ewp, ebp = _mlp_desc_arrays(self._params, "encoder", [self.observation_size] + self.enc_layers + [self.target_size])
        nd, ddm, dwp, dbp = _mlp_desc_arrays(self._params, "decoder", [self.target_size] + self.dec_layers + [self.observation_size])
        d.n_enc, d.enc_dims_host, d.enc_w_host, d.enc_b_host = ne, edm, ewp, ebp
        d.enc_act, d.enc_last_relu = _ACT[self.enc_act], int(self.last_relu)
        d.n_dec, d.dec_dims_host, d.dec_w_host, d.dec_b_host, d.dec_act = nd, ddm, dwp, dbp, _ACT[self.dec_act]
        d.kmat = self._params["kmat"].data_ptr()
        return d, (edm, ewp, ebp, ddm, dwp, dbp)


SparseKM = GenericKM


class LISTAKM(KoopmanMachine):
    """Koopman machine with LISTA sparse encoder and normalised-dictionary decoder (model.py:120-209, 801-870)."""

    def __init__(self, cfg, observation_size: int, device="cuda", seed: int = 0):
        import torch
        super().__init__(cfg, observation_size, device)
        Z = self.target_size
        self.num_loops = int(_get(cfg, "MODEL.ENCODER.LISTA.NUM_LOOPS", 10))
        self.alpha = float(_get(cfg, "MODEL.ENCODER.LISTA.ALPHA", 0.1))
        self.L = float(_get(cfg, "MODEL.ENCODER.LISTA.L", 1e3))
        self.use_linear_encode = bool(_get(cfg, "MODEL.ENCODER.LISTA.LINEAR_ENCODER", False))
        self.enc_layers = list(_get(cfg, "MODEL.ENCODER.LAYERS", []))
        self.enc_bias = bool(_get(cfg, "MODEL.ENCODER.USE_BIAS", False))
        self.enc_act = _get(cfg, "MODEL.ENCODER.ACTIVATION", "relu")
        self.last_relu = bool(_get(cfg, "MODEL.ENCODER.LAST_RELU", False))
        gen = torch.Generator().manual_seed(seed)
        Wd = (torch.randn((self.observation_size, Z), generator=gen) * 0.01).to(self.device)      # model.py:818
        self._params["dict"] = Wd.T.contiguous()
        self._params["dict_init"] = Wd.clone()
        if self.use_linear_encode:
            self._params["lista.We.weight"] = ((1.0 / self.L) * Wd.T).contiguous()                 # model.py:175
        else:
            _mlp_params(self._params, "lista.We", [self.observation_size] + self.enc_layers + [Z], self.enc_bias, gen, self.device)
        self._params["lista.S"] = (torch.eye(Z, device=self.device) - (1.0 / self.L) * (Wd.T @ Wd)).contiguous()
        self._params["kmat"] = torch.eye(Z, dtype=torch.float32, device=self.device)

    def _desc(self, n_assets, delay):
        d = _capi.ModelDesc()
        d.kind, d.obs, d.n_assets, d.delay, d.latent, d.norm_fn = 1, self.observation_size, n_assets, delay, self.target_size, 0
        keep = ()
        d.lista_linear_encoder = int(self.use_linear_encode)
        if self.use_linear_encode:
            d.lista_We = self._params["lista.We.weight"].data_ptr()
        else:
            ne, edm, ewp, ebp = _mlp_desc_arrays(self._params, "lista.We", [self.observation_size] + self.enc_layers + [self.target_size])
            d.n_enc, d.enc_dims_host, d.enc_w_host, d.enc_b_host = ne, edm, ewp, ebp
            d.enc_act, d.enc_last_relu = _ACT[self.enc_act], int(self.last_relu)
            keep = (edm, ewp, ebp)
        d.kmat = self._params["kmat"].data_ptr()
        d.lista_S = self._params["lista.S"].data_ptr()
        d.lista_dict = self._params["dict"].data_ptr()
        d.lista_loops = self.num_loops
        d.lista_threshold = self.alpha / self.L
        return d, keep


_MODEL_REGISTRY = {"GenericKM": GenericKM, "SparseKM": GenericKM, "LISTAKM": LISTAKM}


def make_model(cfg, observation_size: int, device="cuda") -> KoopmanMachine:
    """Factory (model.py:885-903): cfg.MODEL.MODEL_NAME in {GenericKM, SparseKM, LISTAKM}."""
    name = _get(cfg, "MODEL.MODEL_NAME")
    if name not in _MODEL_REGISTRY:
        raise ValueError(f"Unknown model '{name}'. Available: {list(_MODEL_REGISTRY.keys())}")
    return _MODEL_REGISTRY[name](cfg, observation_size, device=device)


def model_config(model_name="GenericKM", target_size=16, enc_layers=(16, 16), dec_layers=(), enc_bias=False, dec_bias=False,
                 enc_act="relu", dec_act="relu", last_relu=False, norm_fn="id", lista_loops=10, lista_L=1e3, lista_alpha=0.1,
                 lista_linear=False):
    """Minimal stand-in for the reference Config tree (config.py:218-291): only the fields the forward path reads."""
    from types import SimpleNamespace as NS
    return NS(MODEL=NS(MODEL_NAME=model_name, TARGET_SIZE=target_size, NORM_FN=norm_fn,
                       ENCODER=NS(LAYERS=list(enc_layers), LAST_RELU=last_relu, USE_BIAS=enc_bias, ACTIVATION=enc_act,
                                  LISTA=NS(NUM_LOOPS=lista_loops, L=lista_L, ALPHA=lista_alpha, LINEAR_ENCODER=lista_linear)),
                       DECODER=NS(LAYERS=list(dec_layers), USE_BIAS=dec_bias, ACTIVATION=dec_act)))


def config_from_dict(d):
    """Attribute-style view of the nested dict the reference stores under ``checkpoint['config']``
    (``Config.to_dict()``, config.py:292-294; rebuilt there by ``Config.from_dict``, config.py:301-336).  Only
    the MODEL sub-tree is read by this package; the rest is kept as found."""
    from types import SimpleNamespace as NS
    if isinstance(d, dict):
        return NS(**{k: config_from_dict(v) for k, v in d.items()})
    return d


def _observation_size_of(sd):
    for key in ("encoder.network.0.weight", "lista.We.weight", "lista.We.network.0.weight"):
        if key in sd:
            return int(sd[key].shape[1])
    if "dict" in sd:
        return int(sd["dict"].shape[1])
    raise ValueError("cannot infer observation_size: state_dict has no encoder / dictionary weight")


def read_checkpoint(path, trust: bool = False):
    """Host part of the checkpoint loader: the file train.py:475-487 writes (``checkpoint.pt`` / ``last.pt``:
    step, epoch, model_state_dict, optimizer_state_dict, config, metrics, finance_metadata).  Returns
    ``(cfg, state_dict, info)`` with ``info = {step, epoch, metrics, finance_metadata, observation_size}``; a bare
    state_dict file is accepted too (cfg None).

    The file is read with ``weights_only=True`` (tensors, containers and the numpy scalar / array types that
    ``finance_metadata`` holds are allow-listed); a checkpoint that needs anything else is refused unless the caller
    passes ``trust=True``, which unpickles arbitrary objects and must only be used on files of known origin."""
    import pickle
    import torch
    safe = []
    try:
        import numpy as _np
        safe = [_np.dtype, _np.ndarray, type(_np.dtype("float64")), type(_np.dtype("float32")), type(_np.dtype("int64"))]
        core = getattr(_np, "_core", None) or getattr(_np, "core")
        safe += [core.multiarray.scalar, core.multiarray._reconstruct]
    except Exception:
        pass
    try:
        with torch.serialization.safe_globals(safe):
            ck = torch.load(path, map_location="cpu", weights_only=True)
    except pickle.UnpicklingError as e:
        if not trust:
            raise ValueError(f"{path}: cannot be read with weights_only=True ({e}); pass trust=True only for a file "
                             "of known origin") from e
        ck = torch.load(path, map_location="cpu", weights_only=False)
    if not isinstance(ck, dict):
        raise ValueError(f"{path}: not a checkpoint dictionary")
    if "model_state_dict" not in ck:
        if all(hasattr(v, "shape") for v in ck.values()) and "kmat" in ck:
            return None, dict(ck), {"observation_size": _observation_size_of(ck)}
        raise ValueError(f"{path}: no 'model_state_dict' entry (keys {sorted(ck)})")
    sd = dict(ck["model_state_dict"])
    meta = ck.get("finance_metadata") or {}
    obs = int(meta["observation_size"]) if "observation_size" in meta else _observation_size_of(sd)
    if obs != _observation_size_of(sd):
        raise ValueError(f"{path}: finance_metadata.observation_size {obs} != weight width {_observation_size_of(sd)}")
    info = {"step": ck.get("step"), "epoch": ck.get("epoch"), "metrics": ck.get("metrics"),
            "finance_metadata": meta, "observation_size": obs}
    cfg = config_from_dict(ck["config"]) if "config" in ck else None
    return cfg, sd, info


def load_checkpoint(path, cfg=None, device="cuda", trust: bool = False):
    """``torch.load`` + ``Config.from_dict`` + ``make_model`` + ``load_state_dict`` + ``eval`` exactly as
    run_experiment.py:67-79 / evaluate_checkpoints.py:121-151 do, for a model of this package.  ``cfg`` overrides
    the stored config (needed for bare state_dict files).  Returns ``(model, info)``."""
    ck_cfg, sd, info = read_checkpoint(path, trust=trust)
    cfg = cfg if cfg is not None else ck_cfg
    if cfg is None:
        raise ValueError(f"{path}: holds no config; pass cfg=")
    model = make_model(cfg, info["observation_size"], device=device)
    model.load_state_dict(sd)
    return model.eval(), info
