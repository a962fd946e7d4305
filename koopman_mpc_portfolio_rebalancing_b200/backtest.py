"""Backtesting module — same names and call conventions as /root/reference/backtest.py, with the step loop
batch-resident on the GPU.

* ``BacktestConfig``         backtest.py:22-30 (field for field)
* ``Strategy`` / ``BuyAndHoldStrategy`` / ``KoopmanMPCStrategy``   backtest.py:32-131
* ``run_backtest(strategy, env, config, verbose)`` -> DataFrame[date, portfolio_value, return, turnover, cost]
                             backtest.py:133-219
* ``calculate_metrics(df)``  backtest.py:221-249
* ``run_backtest_batched``   the data-parallel form: B independent backtests (scenario paths, lambda/tau sweeps,
                             weight sets) in one launch of the persistent backtest kernel (csrc/mpc_lane_kernels.cuh).

For a ``KoopmanMPCStrategy`` the forecast never depends on the weights (backtest.py:85-121 reads only
``data[t]``), so ``run_backtest`` computes the forecasts of all steps in one batched pass and runs the
sequential MPC + portfolio loop inside one kernel.  Any other ``Strategy`` object (Buy&Hold, user classes) is
driven by the reference's own host loop semantics, step by step.
"""
from __future__ import annotations

import ctypes as C
from abc import ABC, abstractmethod
from dataclasses import dataclass
from typing import Dict, Optional

import numpy as np

from . import _capi
from .mpc import MPCConfig, solve_mpc_log_utility  # noqa: F401  (run_experiment.py imports MPCConfig from backtest)

METRIC_KEYS = ("Sharpe Ratio", "Max Drawdown", "Avg Turnover", "Final Value", "Total Return")
HISTORY_COLS = ("portfolio_value", "return", "turnover", "cost")


@dataclass
class BacktestConfig:
    """Configuration for backtesting (backtest.py:22-30)."""
    initial_capital: float = 10000.0
    horizon: int = 5
    rebalance_freq: int = 1
    cost_coeff: float = 0.001
    risk_free_rate: float = 0.0   # unused by the reference loop
    allow_short: bool = False     # unused by the reference loop (the MPC's own flag is MPCConfig.allow_short)


class Strategy(ABC):
    """Abstract base class for trading strategies (backtest.py:32-55)."""

    @abstractmethod
    def rebalance(self, t: int, current_weights: np.ndarray, env, lookback_window: int = 60) -> np.ndarray:
        ...


class BuyAndHoldStrategy(Strategy):
    """Equal weight at t == 0, then hold (backtest.py:57-65)."""

    def rebalance(self, t, current_weights, env, lookback_window=60):
        if t == 0:
            n_assets = env.n_assets
            return np.ones(n_assets) / n_assets
        return current_weights


class KoopmanMPCStrategy(Strategy):
    """Koopman forecast + MPC (backtest.py:67-131).  ``model`` is a KoopmanMachine of this package
    (model.GenericKM / model.LISTAKM).  ``device`` keeps the reference's signature and default (backtest.py:66,
    ``device='cpu'``): this library has no CPU path, so ``'cpu'`` means "wherever the model lives" (a CUDA device)."""

    def __init__(self, model, mpc_config: MPCConfig, device: str = "cpu"):
        self.model = model
        self.mpc_config = mpc_config
        self.device = str(getattr(model, "device", "cuda")) if str(device) == "cpu" else device

    def forecast(self, env, t0: int, t1: int):
        """yhat [t1-t0, H, N] float32 CUDA tensor for test rows t0..t1-1 (backtest.py:85-121 for all t at once)."""
        return self.model.forecast_env(env, t0, t1, self.mpc_config.horizon)

    def rebalance(self, t, current_weights, env, lookback_window=60):
        yhat = self.forecast(env, t, t + 1)[0].cpu().numpy()              # [H, N] float32
        new_weights, _ = solve_mpc_log_utility(current_weights, yhat, self.mpc_config)
        return new_weights[0]                                            # backtest.py:131


def _as_cuda(x, dtype, device):
    import torch
    if x is None:
        return None
    if isinstance(x, torch.Tensor):
        return x.to(device=device, dtype=dtype).contiguous()
    return torch.as_tensor(np.ascontiguousarray(x), dtype=dtype, device=device)


def run_backtest_batched(yhat, realized, *, n_steps: int, horizon: int, lam=None, tau=None, cost_coeff=None,
                         capital=None, lam0: float = 1e-3, tau0: float = 0.2, cost_coeff0: float = 1e-3,
                         capital0: float = 1e4, rebalance_freq: int = 1, allow_short: bool = False,
                         yhat_index=None, realized_index=None, B: Optional[int] = None, want_history: bool = False,
                         want_stats: bool = True, out_metrics=None):
    """Run B independent backtests on the device.

    yhat      [S, n_steps, H, N] float32 CUDA: forecasts of every step for S forecast sets
    realized  [Q, rows, N] float32 CUDA: de-standardised log-returns of every test row for Q price paths
    backtest b uses forecast set yhat_index[b] (default b) and path realized_index[b] (default b);
    lam/tau/cost_coeff/capital: optional per-backtest [B] float64 arrays (else the scalars).
    Returns dict(metrics [B,5] f64 CUDA, history [B,n_hist,4] or None, stats [B,4] i64 or None).
    """
    import torch
    assert yhat.is_cuda and realized.is_cuda and yhat.dtype == torch.float32 and realized.dtype == torch.float32
    dev = yhat.device
    S, ns, H, N = yhat.shape
    assert ns == n_steps and H == horizon
    Q, rows, N2 = realized.shape
    assert N2 == N
    if B is None:
        B = len(yhat_index) if yhat_index is not None else S
    yhat = yhat.contiguous(); realized = realized.contiguous()
    n_hist = (n_steps + rebalance_freq - 1) // rebalance_freq
    metrics = out_metrics if out_metrics is not None else torch.empty((B, 5), dtype=torch.float64, device=dev)
    history = torch.empty((B, n_hist, 4), dtype=torch.float64, device=dev) if want_history else None
    stats = torch.zeros((B, 4), dtype=torch.int64, device=dev) if want_stats else None
    lam_t, tau_t = _as_cuda(lam, torch.float64, dev), _as_cuda(tau, torch.float64, dev)
    cc_t, cap_t = _as_cuda(cost_coeff, torch.float64, dev), _as_cuda(capital, torch.float64, dev)
    yi_t, ri_t = _as_cuda(yhat_index, torch.int32, dev), _as_cuda(realized_index, torch.int32, dev)
    if yi_t is None and S != B:
        raise ValueError("yhat_index is required when the number of forecast sets differs from B")
    if ri_t is None and Q != B:
        raise ValueError("realized_index is required when the number of price paths differs from B")
    d = _capi.BacktestDesc()
    d.B, d.N, d.H, d.rows, d.n_steps, d.rebalance_freq, d.allow_short = B, N, H, rows, n_steps, rebalance_freq, int(allow_short)
    d.yhat, d.yhat_index, d.realized, d.realized_index = _capi.ptr(yhat), _capi.ptr(yi_t), _capi.ptr(realized), _capi.ptr(ri_t)
    d.lam, d.tau, d.cost_coeff, d.capital = _capi.ptr(lam_t), _capi.ptr(tau_t), _capi.ptr(cc_t), _capi.ptr(cap_t)
    d.lam0, d.tau0, d.cost_coeff0, d.capital0 = float(lam0), float(tau0), float(cost_coeff0), float(capital0)
    d.history, d.metrics, d.solve_stats, d.final_weights = _capi.ptr(history), _capi.ptr(metrics), _capi.ptr(stats), None
    h = _capi.Handle.get(dev.index or 0)
    _capi.check(_capi.lib().kmpc_backtest_run(h.ptr, C.byref(d), _capi.stream_ptr(dev.index or 0)))
    return {"metrics": metrics, "history": history, "stats": stats}


def run_backtest(strategy: Strategy, env, config: BacktestConfig, verbose: bool = True):
    """Run backtest loop (backtest.py:133-219).  Returns a DataFrame with daily metrics."""
    import pandas as pd
    n_steps = len(env.test_dataset) - config.horizon
    n_assets = env.n_assets
    dates = env.test_dataset.dates
    # The fused path evaluates the stock KoopmanMPCStrategy.rebalance for every step at once; a subclass that overrides
    # rebalance() (DMDStrategy does not) decides step by step through the generic loop below.
    fused = isinstance(strategy, KoopmanMPCStrategy) and type(strategy).rebalance is KoopmanMPCStrategy.rebalance
    if fused and hasattr(env, "realized_test_returns_device"):
        if n_steps <= 0:
            return pd.DataFrame([])
        mc = strategy.mpc_config
        yhat = strategy.forecast(env, 0, n_steps).unsqueeze(0)                  # [1, n_steps, H, N]
        realized = env.realized_test_returns_device().unsqueeze(0)               # [1, rows, N]
        out = run_backtest_batched(yhat, realized, n_steps=n_steps, horizon=mc.horizon, lam0=mc.cost_coeff,
                                   tau0=mc.max_turnover, cost_coeff0=config.cost_coeff, capital0=config.initial_capital,
                                   rebalance_freq=config.rebalance_freq, allow_short=mc.allow_short, want_history=True)
        hist = out["history"][0].cpu().numpy()
        ts = list(range(0, n_steps, config.rebalance_freq))
        df = pd.DataFrame({"date": [dates[t] for t in ts], "portfolio_value": hist[:, 0], "return": hist[:, 1],
                           "turnover": hist[:, 2], "cost": hist[:, 3]})
        df.attrs["solve_stats"] = out["stats"][0].cpu().numpy()
        return df
    # Any other Strategy object decides on the host, one step at a time; the book-keeping of a step is the
    # same arithmetic as the kernel's (f64 value/weights, f32 realised simple returns, backtest.py:179-208).
    book = _HostBook(n_assets, config.initial_capital, config.cost_coeff)
    simple = np.exp(env.destandardize_returns(env.extract_current_returns(env.test_dataset.data)).cpu().numpy()) - 1.0
    rows = []
    for t in range(0, n_steps, config.rebalance_freq):
        target = strategy.rebalance(t, book.weights, env)
        rec = book.step(np.asarray(target, dtype=np.float64), simple[t + 1] if t + 1 < len(simple) else None)
        rows.append({"date": dates[t], **rec})
    return pd.DataFrame(rows)


class _HostBook:
    """Portfolio state of one backtest on the host: value, drifting weights (starts at 1/N, backtest.py:161)."""

    def __init__(self, n_assets: int, capital: float, cost_coeff: float):
        self.weights = np.full(n_assets, 1.0 / n_assets)
        self.value = capital
        self.cost_coeff = cost_coeff

    def step(self, target: np.ndarray, simple_next) -> dict:
        traded = float(np.abs(target - self.weights).sum())
        fee = self.cost_coeff * traded * self.value
        self.value -= fee
        self.weights = target
        gain = 0.0
        if simple_next is not None:
            gain = np.sum(target * simple_next)
            self.value *= 1.0 + gain
            scale = 1.0 + gain
            if abs(scale) < 1e-8:            # guard of backtest.py:205-206
                scale = 1e-8
            self.weights = target * (1.0 + simple_next) / scale
        return {"portfolio_value": self.value, "return": gain, "turnover": traded, "cost": fee}


def calculate_metrics(df) -> Dict:
    """Sharpe, Max Drawdown, Turnover (backtest.py:221-249)."""
    if len(df) == 0:
        return {}
    returns = df["return"].values
    sharpe = np.sqrt(252) * np.mean(returns) / (np.std(returns) + 1e-8)
    cum_returns = (1 + returns).cumprod()
    peak = np.maximum.accumulate(cum_returns)
    max_dd = np.min((cum_returns - peak) / peak)
    return {
        "Sharpe Ratio": sharpe,
        "Max Drawdown": max_dd,
        "Avg Turnover": df["turnover"].mean(),
        "Final Value": df["portfolio_value"].iloc[-1],
        "Total Return": (df["portfolio_value"].iloc[-1] / df["portfolio_value"].iloc[0]) - 1.0,
    }
