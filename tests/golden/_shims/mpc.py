"""Substitute for /root/reference/mpc.py, used ONLY by tests/golden/make_golden.py so that the
unmodified reference backtest.py / baselines.py can be imported without cvxpy (absent, uninstallable
offline).  Same names and signatures as mpc.py:17-31; the solve is delegated to the fp64 oracle."""
from dataclasses import dataclass

import numpy as np

from oracle import mpc_oracle


@dataclass
class MPCConfig:
    horizon: int = 5
    gamma: float = 0.0
    cost_coeff: float = 0.001
    max_turnover: float = 0.2
    allow_short: bool = False
    solver: str = "ECOS"


CALLS = []  # (w_cur, yhat, w_opt, value) of every call, harvested by make_golden.py


def solve_mpc_log_utility(current_weights, predicted_log_returns, config):
    w, info = mpc_oracle.solve_mpc_log_utility(current_weights, predicted_log_returns, config, method="structured")
    CALLS.append((np.array(current_weights, dtype=np.float64), np.array(predicted_log_returns), w.copy(), info["value"]))
    return w, {"status": info["status"], "value": info["value"]}


MV_CALLS = []  # (w_cur, mu, sigma, w_opt, value) of every mean-variance call


def solve_mpc_mean_variance(current_weights, predicted_log_returns, cov_matrix, config):
    r = mpc_oracle.solve_mv_dense(current_weights, predicted_log_returns, cov_matrix, config.gamma, config.cost_coeff,
                                  config.allow_short)
    MV_CALLS.append((np.array(current_weights, dtype=np.float64), np.array(predicted_log_returns), np.array(cov_matrix),
                     r.w.copy(), r.value))
    status = mpc_oracle.STATUS_NAMES[r.status]
    if r.value is None:
        return r.w, {"status": status}
    return r.w, {"status": status, "value": r.value}
