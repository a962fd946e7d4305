"""Solver-independent optimality certificate for the MPC program of /root/reference/mpc.py:27-117
(TEST INFRASTRUCTURE, not product code; same import rules as the rest of ``oracle/``).

Why it exists: the reference's optimiser (cvxpy 1.7.5 -> SCS 3.2.9, uv.lock) cannot be installed offline, so no
solver OUTPUT of the reference can be pinned.  The CUDA interior-point kernel, ``mpc_oracle.solve_structured`` and
``mpc_oracle.solve_dense`` were all written for this repository; a modelling error they share (how the cost or the cap
enters, a loose acceptance bar) would pass a kernel-vs-oracle comparison.  This module bounds the distance of ANY
candidate plan from the optimum of the program *as mpc.py states it*, without an interior-point method and without
code shared with the solvers: concavity + one linear programme solved by scipy's HiGHS.

The program (mpc.py line numbers):   maximise  f(w) = sum_t log(w_t . R_t)                       :55, :74-80
                                                        - lam * sum_t ||w_t - w_{t-1}||_1          :66-67, :89-92, :101
                                     over F:   sum(w_t) = 1 (:83),  w_t >= 0 unless allow_short (:85-86),
                                               ||w_t - w_{t-1}||_1 <= tau when tau > 0 (:94-95, :102-103),  w_{-1} = w_cur.

For a candidate w with rho_t = w_t . R_t > 0, concavity of the logarithm gives, for every w' in F,

    log(w'_t . R_t)  <=  log(rho_t) - 1 + (R_t / rho_t) . w'_t ,

so   p* = max_F f   <=   UB(w) := sum_t (log rho_t - 1) + max_{w' in F} [ sum_t (R_t/rho_t) . w'_t - lam * sum_t ||w'_t - w'_{t-1}||_1 ].

The inner maximum is a linear programme (epigraph variables for the 1-norms).  If w is feasible then
f(w) <= p* <= UB(w): the certified suboptimality ``UB(w) - f(w)`` is zero exactly at an optimum (first-order optimality
of a concave maximum over a convex set, with the non-smooth part kept exact) and bounds the objective error of w
against the optimum ANY exact solver of the reference program would return — cvxpy/ECOS included.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
from scipy.optimize import linprog


def feasibility(w, w_cur, tau, allow_short=False):
    """(budget residual, most negative weight as a positive number, largest cap excess) of a plan w[H,N]"""
    w = np.asarray(w, dtype=np.float64)
    d = np.diff(np.vstack([np.asarray(w_cur, dtype=np.float64)[None], w]), axis=0)
    budget = float(np.abs(w.sum(axis=1) - 1.0).max())
    neg = 0.0 if allow_short else float(max(0.0, -w.min()))
    cap = float(max(0.0, np.abs(d).sum(axis=1).max() - tau)) if tau > 0 else 0.0
    return budget, neg, cap


def value(w, w_cur, R, lam):
    """f(w): the maximised objective of mpc.py:104, written out independently of mpc_oracle.objective"""
    w = np.asarray(w, dtype=np.float64)
    d = np.diff(np.vstack([np.asarray(w_cur, dtype=np.float64)[None], w]), axis=0)
    return float(np.log((w * R).sum(axis=1)).sum() - lam * np.abs(d).sum())


def upper_bound(w, w_cur, R, lam, tau, allow_short=False):
    """UB(w) of the module docstring.  Returns (UB, w_lp) with w_lp the maximiser of the linearised program, or
    (inf, None) when that LP is unbounded (short selling without a turnover cap)."""
    w = np.asarray(w, dtype=np.float64)
    w_cur = np.asarray(w_cur, dtype=np.float64)
    R = np.asarray(R, dtype=np.float64)
    H, N = w.shape
    rho = (w * R).sum(axis=1)
    if not np.all(rho > 0):
        raise ValueError("candidate has a non-positive portfolio growth w_t . R_t")
    g = R / rho[:, None]                                    # gradient of sum_t log(w_t . R_t) at w
    n = H * N
    has_u = (lam > 0) or (tau > 0)
    # variables x = [w'(H*N), u(H*N)];  linprog minimises c . x
    c = np.concatenate([-g.ravel(), np.full(n, lam) if has_u else np.zeros(0)])
    nv = c.size
    # budget rows (mpc.py:83)
    A_eq = sp.lil_matrix((H, nv))
    for t in range(H):
        A_eq[t, t * N:(t + 1) * N] = 1.0
    b_eq = np.ones(H)
    rows, b_ub = [], []
    if has_u:
        # +-(w'_t - w'_{t-1}) - u_t <= (+-) w_cur for t = 0, 0 otherwise   (mpc.py:66, :90)
        I = sp.identity(N, format="csr")
        for sgn in (1.0, -1.0):
            for t in range(H):
                blk = sp.lil_matrix((N, nv))
                blk[:, t * N:(t + 1) * N] = sgn * I
                if t > 0:
                    blk[:, (t - 1) * N:t * N] = -sgn * I
                blk[:, n + t * N:n + (t + 1) * N] = -I
                rows.append(blk.tocsr())
                b_ub.append(sgn * w_cur if t == 0 else np.zeros(N))
        if tau > 0:                                         # sum_i u_t,i <= tau   (mpc.py:94-95, :102-103)
            cap = sp.lil_matrix((H, nv))
            for t in range(H):
                cap[t, n + t * N:n + (t + 1) * N] = 1.0
            rows.append(cap.tocsr())
            b_ub.append(np.full(H, tau))
    A_ub = sp.vstack(rows).tocsr() if rows else None
    b_ub = np.concatenate(b_ub) if rows else None
    lo_w = None if allow_short else 0.0
    bounds = [(lo_w, None)] * n + [(0.0, None)] * (nv - n)
    res = linprog(c, A_ub=A_ub, b_ub=b_ub, A_eq=A_eq.tocsr(), b_eq=b_eq, bounds=bounds, method="highs",
                  options={"primal_feasibility_tolerance": 1e-10, "dual_feasibility_tolerance": 1e-10, "presolve": True})
    if res.status == 3:
        return np.inf, None
    if res.status != 0:
        raise RuntimeError(f"linprog failed: {res.message}")
    ub = float((np.log(rho) - 1.0).sum() - res.fun)
    return ub, res.x[:n].reshape(H, N)


def certify(w, w_cur, R, lam, tau, allow_short=False):
    """dict(value=f(w), upper=UB(w), gap=UB-f, feas=(budget, negativity, cap excess)).
    For a feasible w:  value <= optimum of the reference program <= upper."""
    f = value(w, w_cur, R, lam)
    ub, _ = upper_bound(w, w_cur, R, lam, tau, allow_short)
    return dict(value=f, upper=ub, gap=ub - f, feas=feasibility(w, w_cur, tau, allow_short))
