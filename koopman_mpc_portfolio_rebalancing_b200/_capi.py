"""ctypes binding of libkmpc.so (include/kmpc.h).  The library is built in-tree by build.py; importing this
module when it is missing raises, and creating a handle without a B200 raises: there is no CPU fallback."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libkmpc.so")

c_double_p = C.POINTER(C.c_double)
c_float_p = C.POINTER(C.c_float)
c_int32_p = C.POINTER(C.c_int32)
c_int64_p = C.POINTER(C.c_int64)
vp = C.c_void_p


class KmpcError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libkmpc error {code}: {msg}")
        self.code = code


class ModelDesc(C.Structure):
    _fields_ = [
        ("kind", C.c_int), ("obs", C.c_int), ("n_assets", C.c_int), ("delay", C.c_int), ("latent", C.c_int),
        ("norm_fn", C.c_int),
        ("n_enc", C.c_int), ("enc_dims_host", c_int32_p), ("enc_w_host", C.POINTER(vp)), ("enc_b_host", C.POINTER(vp)),
        ("enc_act", C.c_int), ("enc_last_relu", C.c_int),
        ("n_dec", C.c_int), ("dec_dims_host", c_int32_p), ("dec_w_host", C.POINTER(vp)), ("dec_b_host", C.POINTER(vp)),
        ("dec_act", C.c_int),
        ("kmat", vp),
        ("lista_linear_encoder", C.c_int), ("lista_We", vp), ("lista_S", vp), ("lista_dict", vp),
        ("lista_loops", C.c_int), ("lista_threshold", C.c_float),
    ]


class BacktestDesc(C.Structure):
    _fields_ = [
        ("B", C.c_int), ("N", C.c_int), ("H", C.c_int), ("rows", C.c_int), ("n_steps", C.c_int),
        ("rebalance_freq", C.c_int), ("allow_short", C.c_int),
        ("yhat", vp), ("yhat_index", vp), ("realized", vp), ("realized_index", vp),
        ("lam", vp), ("tau", vp), ("cost_coeff", vp), ("capital", vp),
        ("lam0", C.c_double), ("tau0", C.c_double), ("cost_coeff0", C.c_double), ("capital0", C.c_double),
        ("history", vp), ("metrics", vp), ("solve_stats", vp), ("final_weights", vp),
    ]


# name -> (restype, argtypes); every symbol include/kmpc.h declares
SIGNATURES = {
    "kmpc_version": (C.c_int, []),
    "kmpc_last_error": (C.c_char_p, []),
    "kmpc_create": (C.c_int, [C.c_int, C.POINTER(vp)]),
    "kmpc_destroy": (C.c_int, [vp]),
    "kmpc_launch_count": (C.c_int64, [vp]),
    "kmpc_mpc_supported": (C.c_int, [C.c_int, C.c_int]),
    "kmpc_set_solver_param": (C.c_int, [vp, C.c_int, C.c_double]),
    "kmpc_standardize": (C.c_int, [vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, C.c_int, vp]),
    "kmpc_embed_gather": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp]),
    "kmpc_embed_index_host": (C.c_int, [C.c_int, C.c_int, C.c_int, vp]),
    "kmpc_current_returns": (C.c_int, [vp, vp, C.c_int, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                       C.c_int, vp, vp]),
    "kmpc_model_load": (C.c_int, [vp, C.POINTER(ModelDesc), C.POINTER(vp)]),
    "kmpc_model_free": (C.c_int, [vp]),
    "kmpc_forecast": (C.c_int, [vp, vp, vp, C.c_int, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                C.c_int, vp, vp]),
    "kmpc_encode": (C.c_int, [vp, vp, vp, C.c_int, vp, vp]),
    "kmpc_step_latent": (C.c_int, [vp, vp, vp, C.c_int, vp, vp]),
    "kmpc_decode": (C.c_int, [vp, vp, vp, C.c_int, vp, vp]),
    "kmpc_rollout": (C.c_int, [vp, vp, vp, C.c_int, C.c_int, C.c_int, vp, vp]),
    "kmpc_set_gemm_mode": (C.c_int, [C.c_int]),
    "kmpc_set_forecast_fold": (C.c_int, [C.c_int]),
    "kmpc_set_gemm_fp16_pairs": (C.c_int, [C.c_int]),
    "kmpc_set_forecast_chunk_rows": (C.c_int, [C.c_int]),
    "kmpc_set_forecast_embedding": (C.c_int, [C.c_int]),
    "kmpc_debug_gemm": (C.c_int, [vp, vp, vp, C.c_int, C.c_int, C.c_int, vp, C.c_int]),
    "kmpc_mpc_solve": (C.c_int, [vp, vp, vp, vp, vp, vp, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int,
                                 vp, vp, vp, vp, vp, vp]),
    "kmpc_mv_supported": (C.c_int, [C.c_int, C.c_int]),
    "kmpc_mpc_mean_variance": (C.c_int, [vp, vp, vp, C.c_int, vp, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int,
                                         vp, vp, vp, vp, vp, vp]),
    "kmpc_mpc_mean_variance_host": (C.c_int, [vp, vp, vp, vp, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int,
                                              vp, vp, vp, vp, vp]),
    "kmpc_mpc_solve_host": (C.c_int, [vp, vp, C.c_int, vp, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int,
                                      vp, vp, vp, vp, vp]),
    "kmpc_backtest_run": (C.c_int, [vp, C.POINTER(BacktestDesc), vp]),
}

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -m koopman_mpc_portfolio_rebalancing_b200.build` "
                "(nvcc, sm_100a).  This package has no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)          # AttributeError if the ABI and the header drift apart
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc: int):
    if rc != 0:
        raise KmpcError(rc, lib().kmpc_last_error().decode())


class Handle:
    """One kmpc_handle per device.  Raises KmpcError when no sm_100 device is present."""

    _cache: dict = {}

    def __init__(self, device: int = 0):
        self.device = device
        self.ptr = vp()
        check(lib().kmpc_create(device, C.byref(self.ptr)))

    @classmethod
    def get(cls, device: int = 0) -> "Handle":
        if device not in cls._cache:
            cls._cache[device] = cls(device)
        return cls._cache[device]

    @property
    def launches(self) -> int:
        return int(lib().kmpc_launch_count(self.ptr))

    def __del__(self):
        try:
            if self.ptr:
                lib().kmpc_destroy(self.ptr)
        except Exception:
            pass


def ptr(t):
    """device/host pointer of a torch tensor / numpy array / None"""
    if t is None:
        return None
    if hasattr(t, "data_ptr"):
        return vp(t.data_ptr())
    return vp(t.ctypes.data)


def stream_ptr(device: int = 0):
    import torch
    return vp(torch.cuda.current_stream(device).cuda_stream)
