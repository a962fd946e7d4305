"""Baseline strategies of the reference that reuse the hot path (/root/reference/baselines.py).

``DMDStrategy`` (baselines.py:109-187): a linear operator K fitted on the training split by a pseudo-inverse
(``x_{t+1} = K x_t`` on the standardised delay-embedded state), rolled out H steps, first N entries de-standardised,
fed to the same ``solve_mpc_log_utility``.  The fit stays on the host exactly as in the reference
(``X' @ pinv(X)``, scipy); the forecast of every rebalancing step is the forecast path of this library with an
identity encoder / decoder and ``kmat = K^T`` (row-vector convention ``z_{k+1} = z_k @ kmat``), i.e. the folded
read-out ``yhat_k = y_t . (K^{k+1})[:N]^T`` evaluated by the tcgen05 GEMM for all steps at once, and the
persistent MPC + portfolio kernel behind ``run_backtest``.

``MarkowitzStrategy`` / ``solve_mpc_mean_variance`` (baselines.py:24-106, mpc.py:119-184) use a quadratic stage cost
and are not built yet.
"""
from __future__ import annotations

import numpy as np

from .backtest import KoopmanMPCStrategy
from .model import make_model, model_config
from .mpc import MPCConfig


class DMDStrategy(KoopmanMPCStrategy):
    """Dynamic Mode Decomposition (linear Koopman) strategy, same constructor as baselines.py:117-126."""

    def __init__(self, train_data, mpc_config: MPCConfig, device: str = "cuda"):
        data = train_data.detach().cpu().numpy() if hasattr(train_data, "detach") else np.asarray(train_data)
        self.K = self._fit_dmd(data)
        self.n_assets = None
        obs = self.K.shape[0]
        # identity encoder / decoder around kmat = K^T: the library's forecast path then evaluates
        # yhat_k = first N entries of K^{k+1} y_t for every test row (folded into one GEMM)
        cfg = model_config("GenericKM", obs, enc_layers=(), dec_layers=(), enc_bias=False, dec_bias=False, norm_fn="id")
        model = make_model(cfg, obs, device=device)
        eye = np.eye(obs, dtype=np.float32)
        model.load_state_dict({"encoder.network.0.weight": eye, "decoder.network.0.weight": eye,
                               "kmat": np.ascontiguousarray(self.K.T.astype(np.float32))})
        super().__init__(model, mpc_config, device)

    @staticmethod
    def _fit_dmd(data: np.ndarray) -> np.ndarray:
        """K = X' pinv(X) with X = data[:-1]^T, X' = data[1:]^T (baselines.py:127-145), in the dtype of ``data``."""
        from scipy.linalg import pinv
        X = data[:-1].T
        X_prime = data[1:].T
        return X_prime @ pinv(X)
