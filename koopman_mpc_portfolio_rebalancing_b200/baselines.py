"""Baseline strategies of the reference that reuse the hot path (/root/reference/baselines.py).

``DMDStrategy`` (baselines.py:109-187): a linear operator K fitted on the training split by a pseudo-inverse
(``x_{t+1} = K x_t`` on the standardised delay-embedded state), rolled out H steps, first N entries de-standardised,
fed to the same ``solve_mpc_log_utility``.  The fit stays on the host exactly as in the reference
(``X' @ pinv(X)``, scipy); the forecast of every rebalancing step is the forecast path of this library with an
identity encoder / decoder and ``kmat = K^T`` (row-vector convention ``z_{k+1} = z_k @ kmat``), i.e. the folded
read-out ``yhat_k = y_t . (K^{k+1})[:N]^T`` evaluated by the tcgen05 GEMM for all steps at once, and the
persistent MPC + portfolio kernel behind ``run_backtest``.

``MarkowitzStrategy`` (baselines.py:24-106): rolling-window mean / covariance estimated on the host exactly as the
reference does (float32 ``np.mean``, ``np.cov`` + 1e-6 I), then ``solve_mpc_mean_variance`` with H = 1 — the fp64
dense interior-point kernel ``csrc/mpc_mv.cu``.  It decides one step at a time through the generic ``run_backtest``
loop, like in the reference.
"""
from __future__ import annotations

import numpy as np

from .backtest import KoopmanMPCStrategy
from .model import make_model, model_config
from .backtest import Strategy
from .mpc import MPCConfig, solve_mpc_mean_variance


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


class MarkowitzStrategy(Strategy):
    """Classic mean-variance optimisation on rolling-window estimates (baselines.py:24-106)."""

    def __init__(self, risk_aversion: float = 1.0, cost_coeff: float = 0.001, allow_short: bool = False):
        self.risk_aversion = risk_aversion
        self.cost_coeff = cost_coeff
        self.allow_short = allow_short
        self.mpc_config = MPCConfig(horizon=1, gamma=risk_aversion, cost_coeff=cost_coeff, allow_short=allow_short,
                                    solver="ECOS")

    def rebalance(self, t, current_weights, env, lookback_window: int = 60) -> np.ndarray:
        past_data = env.test_dataset.data[:t + 1]
        past_returns = env.destandardize_returns(env.extract_current_returns(past_data)).cpu().numpy()
        if len(past_returns) < 5:                      # baselines.py:72-74
            return current_weights
        window = past_returns[-lookback_window:]
        mu = np.mean(window, axis=0)
        sigma = np.cov(window, rowvar=False)
        sigma += np.eye(len(mu)) * 1e-6
        w_opt, _ = solve_mpc_mean_variance(current_weights, mu.reshape(1, -1), sigma, self.mpc_config)
        return w_opt[0]
