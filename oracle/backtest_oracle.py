"""CPU oracle for the backtest step loop and its metrics (TEST INFRASTRUCTURE, not product code).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
leg may import this module.

Restates ``/root/reference/backtest.py``:

* ``run_backtest``       run_backtest loop        backtest.py:133-219  (n_steps, 1/N start, cost, value,
                                                  f32 realised return, f64 drift with the 1e-8 clamp)
* ``calculate_metrics``  calculate_metrics        backtest.py:221-249

The strategy's decision is abstracted as ``decide(t, w_cur) -> w_new`` so that the same loop serves
KoopmanMPC (forecast -> MPC, backtest.py:80-131) and Buy&Hold (backtest.py:57-65).

Pinned against the reference itself: tests/golden/make_golden.py runs the *unmodified* reference
``run_backtest`` / ``calculate_metrics`` (with a stub matplotlib and a substitute ``mpc`` module that
calls oracle/mpc_oracle.py, because cvxpy cannot be installed) and stores history + metrics in
tests/golden/backtest_cfg1.npz; tests/test_oracle_backtest.py replays them through this file.
"""
from __future__ import annotations

import numpy as np

from . import mpc_oracle

HISTORY_COLS = ("portfolio_value", "return", "turnover", "cost")
METRIC_KEYS = ("Sharpe Ratio", "Max Drawdown", "Avg Turnover", "Final Value", "Total Return")


def realized_simple_returns_f32(all_returns_f32: np.ndarray, exp_mode: str = "cr32") -> np.ndarray:
    """np.exp(all_returns[t+1]) - 1.0 on a float32 array stays float32 (backtest.py:193).

    exp_mode "numpy32": numpy's own fp32 exp, bit-faithful to the reference on the same numpy build (its SIMD
    exp is within ~1 ulp but not correctly rounded, and which kernel runs depends on the CPU).
    exp_mode "cr32": the correctly rounded fp32 exp, round_f32(exp_f64(y)) -- the platform-independent
    convention shared by the oracle and the CUDA path (see mpc_oracle.gross_returns_f32).  The two differ by
    at most 1 ulp of ~1.0 (1.2e-7 absolute on a daily simple return)."""
    if exp_mode == "numpy32":
        e = np.exp(all_returns_f32.astype(np.float32))
    else:
        e = np.exp(all_returns_f32.astype(np.float64)).astype(np.float32)
    return (e - np.float32(1.0)).astype(np.float32)


def run_backtest(decide, all_returns_f32: np.ndarray, n_rows_dataset: int, horizon: int,
                 initial_capital: float = 10000.0, cost_coeff: float = 0.001, rebalance_freq: int = 1,
                 exp_mode: str = "cr32"):
    """all_returns_f32 [rows, N]: de-standardised current log-returns of every test row (backtest.py:169-171).
    n_rows_dataset = len(env.test_dataset) = rows - sequence_length (data_finance.py:389).
    Returns history [n, 4] float64 with columns HISTORY_COLS and the list of t."""
    n_steps = n_rows_dataset - horizon
    N = all_returns_f32.shape[1]
    V = float(initial_capital)
    w = np.ones(N) / N
    simple = realized_simple_returns_f32(all_returns_f32, exp_mode)
    hist, ts = [], []
    for t in range(0, n_steps, rebalance_freq):
        target = np.asarray(decide(t, w), dtype=np.float64)
        turnover = float(np.sum(np.abs(target - w)))
        cost = cost_coeff * turnover * V
        w = target
        V -= cost
        port_ret = 0.0
        if t + 1 < all_returns_f32.shape[0]:
            r = simple[t + 1]
            port_ret = float(np.sum(w * r))
            V *= (1.0 + port_ret)
            denom = 1.0 + port_ret
            if abs(denom) < 1e-8:
                denom = 1e-8
            w = w * (1.0 + r) / denom
        hist.append((V, port_ret, turnover, cost))
        ts.append(t)
    return np.asarray(hist, dtype=np.float64).reshape(-1, 4), ts


def calculate_metrics(history: np.ndarray) -> dict:
    if len(history) == 0:
        return {}
    returns = history[:, 1]
    sharpe = np.sqrt(252) * np.mean(returns) / (np.std(returns) + 1e-8)
    cum = (1 + returns).cumprod()
    peak = np.maximum.accumulate(cum)
    max_dd = np.min((cum - peak) / peak)
    return {
        "Sharpe Ratio": float(sharpe),
        "Max Drawdown": float(max_dd),
        "Avg Turnover": float(history[:, 2].mean()),
        "Final Value": float(history[-1, 0]),
        "Total Return": float(history[-1, 0] / history[0, 0] - 1.0),
    }


def koopman_mpc_decider(yhat: np.ndarray, lam: float, tau: float, allow_short: bool = False,
                        method: str = "structured", stats: list | None = None):
    """decide(t, w) for KoopmanMPCStrategy given the forecasts yhat [T_s, H, N] f32 of every step
    (the forecast never depends on the weights, backtest.py:85-121).  Applies new_weights[0]
    (backtest.py:131) and the hold-weights fallback (mpc.py:113-115)."""
    fn = mpc_oracle.solve_dense if method == "dense" else mpc_oracle.solve_structured

    def decide(t, w):
        r = fn(w, yhat[t], lam, tau, allow_short)
        if stats is not None:
            stats.append((r.status, r.iters, r.value))
        return r.w[0]

    return decide
