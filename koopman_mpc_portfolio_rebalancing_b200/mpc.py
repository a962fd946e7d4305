"""MPC module — same names, arguments and error behaviour as /root/reference/mpc.py, solved on the GPU.

``MPCConfig`` mirrors mpc.py:17-25 field for field; ``solve_mpc_log_utility`` mirrors mpc.py:27-117: it never
raises on solver failure, it falls back to ``np.tile(current_weights, (H, 1))`` with ``value=None`` and a
non-"optimal" status string.  The solve itself is the lane-per-asset fp64 interior-point kernel of
csrc/mpc_lane.cuh, reached through the C ABI (kmpc_mpc_solve_host / kmpc_mpc_solve).  No CPU fallback: a shape without a
compiled kernel variant raises ``KmpcError`` (it is a deployment error, not a solver failure).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Dict, Tuple

import numpy as np

from . import _capi

STATUS_STRINGS = {0: "optimal", 1: "optimal_inaccurate", 2: "solver_error", 3: "nonfinite_input"}


@dataclass
class MPCConfig:
    """Configuration for MPC solver (mpc.py:17-25)."""
    horizon: int = 5
    gamma: float = 0.0          # unused by the log-utility program (kept for signature parity)
    cost_coeff: float = 0.001   # lambda of the objective
    max_turnover: float = 0.2   # tau; <= 0 disables the cap (mpc.py:94)
    allow_short: bool = False
    solver: str = "ECOS"        # accepted and ignored: the CUDA interior-point solver is the only backend


def solve_mpc_log_utility(current_weights: np.ndarray, predicted_log_returns: np.ndarray, config: MPCConfig,
                          device: int = 0) -> Tuple[np.ndarray, Dict]:
    """maximize sum_t [ log(w_t . exp(y_t)) - cost * ||w_t - w_{t-1}||_1 ]  s.t. sum(w_t)=1, w_t>=0,
    ||w_t - w_{t-1}||_1 <= max_turnover.  Returns (optimal_weights [H,N] float64, {"status", "value"})."""
    y = np.asarray(predicted_log_returns)
    if y.ndim != 2:
        raise ValueError("predicted_log_returns must be [horizon, n_assets]")   # the reference fails at H, N = shape
    H, N = y.shape
    f64 = y.dtype != np.float32
    y = np.ascontiguousarray(y, dtype=np.float64 if f64 else np.float32)
    w0 = np.ascontiguousarray(current_weights, dtype=np.float64).reshape(N)
    w = np.empty((H, N), dtype=np.float64)
    obj = np.empty(1, dtype=np.float64)
    kkt = np.empty(3, dtype=np.float64)
    st = np.empty(1, dtype=np.int32)
    it = np.empty(1, dtype=np.int32)
    h = _capi.Handle.get(device)
    _capi.check(_capi.lib().kmpc_mpc_solve_host(
        h.ptr, _capi.ptr(y), int(f64), _capi.ptr(w0), float(config.cost_coeff), float(config.max_turnover),
        int(bool(config.allow_short)), 1, H, N, _capi.ptr(w), _capi.ptr(obj), _capi.ptr(kkt), _capi.ptr(st), _capi.ptr(it)))
    status = STATUS_STRINGS[int(st[0])]
    if status not in ("optimal", "optimal_inaccurate"):
        return np.tile(w0, (H, 1)), {"status": status, "value": None, "kkt": tuple(kkt), "iterations": int(it[0])}
    return w, {"status": status, "value": float(obj[0]), "kkt": tuple(kkt), "iterations": int(it[0])}


def solve_mpc_mean_variance(current_weights: np.ndarray, predicted_log_returns: np.ndarray, cov_matrix: np.ndarray,
                            config: MPCConfig, device: int = 0) -> Tuple[np.ndarray, Dict]:
    """maximize sum_t [ w_t . mu_t - gamma * w_t' Sigma w_t - cost * ||w_t - w_{t-1}||_1 ]  s.t. sum(w_t)=1, w_t>=0
    unless allow_short; no turnover cap (mpc.py:119-184).  Returns (optimal_weights [H,N], {"status", "value"}); a
    non-optimal status returns tile(current_weights) and {"status"} only, like the reference (mpc.py:179-180)."""
    mu = np.ascontiguousarray(predicted_log_returns, dtype=np.float64)
    if mu.ndim != 2:
        raise ValueError("predicted_log_returns must be [horizon, n_assets]")
    H, N = mu.shape
    sig = np.ascontiguousarray(cov_matrix, dtype=np.float64)
    if sig.shape != (N, N):
        raise ValueError("cov_matrix must be [n_assets, n_assets]")
    w0 = np.ascontiguousarray(current_weights, dtype=np.float64).reshape(N)
    w = np.empty((H, N), dtype=np.float64)
    obj = np.empty(1, dtype=np.float64); kkt = np.empty(3, dtype=np.float64)
    st = np.empty(1, dtype=np.int32); it = np.empty(1, dtype=np.int32)
    h = _capi.Handle.get(device)
    _capi.check(_capi.lib().kmpc_mpc_mean_variance_host(
        h.ptr, _capi.ptr(mu), _capi.ptr(sig), _capi.ptr(w0), float(config.gamma), float(config.cost_coeff),
        int(bool(config.allow_short)), H, N, _capi.ptr(w), _capi.ptr(obj), _capi.ptr(kkt), _capi.ptr(st), _capi.ptr(it)))
    status = STATUS_STRINGS[int(st[0])]
    if status not in ("optimal", "optimal_inaccurate"):
        return np.tile(w0, (H, 1)), {"status": status}
    return w, {"status": status, "value": float(obj[0]), "kkt": tuple(kkt), "iterations": int(it[0])}


def solve_mean_variance_batch(w_cur, mu, sigma, gamma: float, cost_coeff: float = 1e-3, allow_short: bool = False):
    """P mean-variance problems resident on the device: w_cur [P,N], mu [P,H,N], sigma [N,N] or [P,N,N] (fp64 CUDA)."""
    import torch
    P, H, N = mu.shape
    dev = mu.device.index or 0
    mu = mu.contiguous().to(torch.float64); sigma = sigma.contiguous().to(torch.float64)
    w_cur = w_cur.contiguous().to(torch.float64)
    out = {"w": torch.empty((P, H, N), dtype=torch.float64, device=mu.device),
           "value": torch.empty(P, dtype=torch.float64, device=mu.device),
           "kkt": torch.empty((P, 3), dtype=torch.float64, device=mu.device),
           "status": torch.empty(P, dtype=torch.int32, device=mu.device),
           "iterations": torch.empty(P, dtype=torch.int32, device=mu.device)}
    h = _capi.Handle.get(dev)
    _capi.check(_capi.lib().kmpc_mpc_mean_variance(
        h.ptr, _capi.ptr(mu), _capi.ptr(sigma), int(sigma.dim() == 3), _capi.ptr(w_cur), float(gamma), float(cost_coeff),
        int(bool(allow_short)), P, H, N, _capi.ptr(out["w"]), _capi.ptr(out["value"]), _capi.ptr(out["kkt"]),
        _capi.ptr(out["status"]), _capi.ptr(out["iterations"]), _capi.stream_ptr(dev)))
    return out


def solve_mpc_batch(w_cur, yhat, lam=None, tau=None, cost_coeff: float = 1e-3, max_turnover: float = 0.2,
                    allow_short: bool = False):
    """P problems resident on the device.  w_cur [P,N] f64, yhat [P,H,N] f32 (or f64) CUDA tensors; lam/tau
    optional [P] f64 CUDA tensors.  Returns dict of CUDA tensors: w [P,H,N], value [P], kkt [P,3], status [P],
    iterations [P]."""
    import torch
    assert yhat.is_cuda and w_cur.is_cuda
    P, H, N = yhat.shape
    dev = yhat.device.index or 0
    yhat = yhat.contiguous()
    w_cur = w_cur.contiguous().to(torch.float64)
    out = {
        "w": torch.empty((P, H, N), dtype=torch.float64, device=yhat.device),
        "value": torch.empty(P, dtype=torch.float64, device=yhat.device),
        "kkt": torch.empty((P, 3), dtype=torch.float64, device=yhat.device),
        "status": torch.empty(P, dtype=torch.int32, device=yhat.device),
        "iterations": torch.empty(P, dtype=torch.int32, device=yhat.device),
    }
    if lam is not None:
        lam = lam.contiguous().to(torch.float64)
    if tau is not None:
        tau = tau.contiguous().to(torch.float64)
    y32 = yhat if yhat.dtype == torch.float32 else None
    y64 = yhat if yhat.dtype == torch.float64 else None
    if y32 is None and y64 is None:
        raise TypeError("yhat must be float32 or float64")
    h = _capi.Handle.get(dev)
    _capi.check(_capi.lib().kmpc_mpc_solve(
        h.ptr, _capi.ptr(y32), _capi.ptr(y64), _capi.ptr(w_cur), _capi.ptr(lam), _capi.ptr(tau), float(cost_coeff),
        float(max_turnover), int(bool(allow_short)), P, H, N, _capi.ptr(out["w"]), _capi.ptr(out["value"]),
        _capi.ptr(out["kkt"]), _capi.ptr(out["status"]), _capi.ptr(out["iterations"]), _capi.stream_ptr(dev)))
    return out
