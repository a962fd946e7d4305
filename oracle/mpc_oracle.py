"""CPU oracle for the MPC program of the reference (TEST INFRASTRUCTURE, not product code).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
leg may import this module.  The product path (``koopman_mpc_portfolio_rebalancing_b200``) never does.

What it restates
----------------
``/root/reference/mpc.py:27-117`` (``solve_mpc_log_utility``):

    maximise   sum_t log(w_t . exp(yhat_t))                      mpc.py:55, 74-80
               - lam * ||w_0 - w_cur||_1                          mpc.py:66-67, 101
               - lam * sum_{t>0} ||w_t - w_{t-1}||_1              mpc.py:89-92
    s.t.       sum(w_t) = 1                 for every t           mpc.py:83
               w_t >= 0                     unless allow_short    mpc.py:85-86
               ||w_t - w_{t-1}||_1 <= tau   for every t incl. 0,  mpc.py:94-95, 102-103
                                            only when tau > 0
    fallback   non-optimal status -> tile(w_cur, (H,1)), value None      mpc.py:113-115

The arithmetic of the reference lives in third-party code that is absent from ``/root/reference`` and
from this image: cvxpy 1.7.5 (uv.lock) canonicalises the program and hands it to SCS 3.2.9 (``ECOS`` is
requested at mpc.py:25/108 but not locked, so ``cp.SolverError`` -> SCS, mpc.py:109-111).  Neither can be
installed offline.  This file therefore restates the *program* and solves it to its exact optimum in
fp64 (KKT <= 1e-10) with two independent methods:

* ``solve_dense``  -- textbook primal-dual interior point on the epigraph form with an explicit
  constraint matrix G (scipy.sparse) and a dense KKT solve.  Obviously-correct, slow (O((HN)^3)).
* ``solve_structured`` -- same central path, Newton system solved through the problem structure
  (per-asset H x H tridiagonal + rank-3H border).  Used for large N and for the CPU baseline timing;
  validated against ``solve_dense`` in tests/test_oracle_mpc.py.

Parity pinning: the reference's own tests for this boundary (tests/test_mpc.py T1-T3) only pin status,
shape, feasibility and two inequalities.  They are all checked in tests/test_oracle_mpc.py together with
the exact optima derived in SURVEY.md §8c (T2: w=[0.6,0.4], obj 0.06119156775022542; T3: w=[1,0], obj 0)
and an LP cross-check (lam=0,H=1 -> scipy.optimize.linprog on the linearised program bounds).
Beyond those KATs the reference's solver output cannot be produced here: **parity unpinned** against
cvxpy/SCS bit patterns; parity is defined against the exact optimum of the same program.
"""
from __future__ import annotations

import numpy as np

# loose acceptance ("optimal_inaccurate") of an iterate that could not be pushed to the tolerances; same constants as
# csrc/mpc_ipm.cuh (kLoosePres / kLooseDres / kLooseGap), see the comment there
LOOSE_PRES, LOOSE_DRES, LOOSE_GAP = 1e-8, 1e-4, 1e-7
MAX_FACTOR_RETRIES = 4        # kMaxFactorRetries of csrc/mpc_common.cuh
# Second attempt of a solve whose first attempt did not reach "optimal" (kRobust* in csrc/mpc_common.cuh): restart from
# the cold starting point with textbook-robust parameters — one common primal/dual step length, a shorter fraction to
# the boundary, a centring floor, a weaker proximal term — the parameterisation of solve_dense, which converges on
# every such instance.  On the 52 decisions of a config-2 step (1.0 M decisions) that the aggressive first attempt leaves
# "optimal_inaccurate" (near-degenerate optima where the dual residual stalls at 1e-7..1e-5 while the gap collapses;
# 4 of them 2-4e-6 off the optimal objective), the second attempt reaches "optimal" on all 52 in 12.6 iterations on
# average, worst objective error 6e-9 relative.
ROBUST_STEP_FRAC, ROBUST_SIGMA_MIN, ROBUST_DELTA = 0.995, 0.05, 1e-7
CORRECTOR_FULL_STEP = 0.3   # affine step below which the corrector's second-order term is scaled down (kCorrFull)
STATUS_OPTIMAL = 0
STATUS_INACCURATE = 1
STATUS_MAXITER = 2
STATUS_NONFINITE = 3
STATUS_NAMES = {0: "optimal", 1: "optimal_inaccurate", 2: "solver_error", 3: "nonfinite_input"}


def gross_returns_f32(yhat: np.ndarray) -> np.ndarray:
    """R = exp(yhat) rounded to fp32, as mpc.py:55 does on the fp32 array the strategy passes
    (backtest.py:119-121).  numpy's fp32 exp is within 1 ulp but platform dependent; the oracle and the
    CUDA path both use the correctly rounded value round_f32(exp_f64(y)) so that they see identical R."""
    y = np.asarray(yhat)
    if y.dtype == np.float32:
        return np.exp(y.astype(np.float64)).astype(np.float32).astype(np.float64)
    return np.exp(y.astype(np.float64))


def objective(w: np.ndarray, w_cur: np.ndarray, R: np.ndarray, lam: float) -> float:
    """Maximised objective value of mpc.py:104 for a given plan w[H,N] (fp64)."""
    H = w.shape[0]
    val = 0.0
    prev = w_cur
    for t in range(H):
        val += np.log(float(w[t] @ R[t])) - lam * float(np.abs(w[t] - prev).sum())
        prev = w[t]
    return val


def _initial_point(w_cur, H, N, tau, has_u, allow_short):
    base = np.asarray(w_cur, dtype=np.float64)
    if not allow_short:
        base = np.maximum(base, 0.0)
    sb = base.sum()
    base = base / sb if sb > 0 else np.full(N, 1.0 / N)
    eps = 0.1 if tau <= 0 else min(0.1, tau / 8.0)
    w1 = (1.0 - eps) * base + eps / N
    w = np.tile(w1, (H, 1))
    u = None
    if has_u:
        d = w.copy()
        d[0] -= w_cur
        d[1:] -= w[:-1]
        absd = np.abs(d)
        if tau > 0:
            room = tau - absd.sum(axis=1)
            if room[0] <= 0:  # w_cur too far from the simplex for the cap: infeasible program
                return None
            delta = room / (2.0 * N)
        else:
            delta = np.full(H, 0.05 / N)
        u = absd + delta[:, None]
    return w, u


class _Result(dict):
    __getattr__ = dict.__getitem__


def _finish(w, w_cur, R, lam, status, it, res, H):
    if status in (STATUS_OPTIMAL, STATUS_INACCURATE):
        val = objective(w, w_cur, R, lam)
        return _Result(w=w, value=val, status=status, iters=it, kkt=res)
    return _Result(w=np.tile(np.asarray(w_cur, dtype=np.float64), (H, 1)), value=None, status=status,
                   iters=it, kkt=res)


# ----------------------------------------------------------------------------------------------
# dense reference IPM
# ----------------------------------------------------------------------------------------------

def _build_constraints(w_cur, H, N, tau, has_u, allow_short):
    """Explicit G x <= h, A x = b for x = [w (H*N), u (H*N)] following mpc.py:83-103 one line each."""
    import scipy.sparse as sp

    n = H * N * (2 if has_u else 1)
    rows, cols, vals, h = [], [], [], []
    r = 0

    def wi(t, i):
        return t * N + i

    def ui(t, i):
        return H * N + t * N + i

    if not allow_short:  # w_t >= 0                                              mpc.py:85-86
        for t in range(H):
            for i in range(N):
                rows.append(r); cols.append(wi(t, i)); vals.append(-1.0); h.append(0.0); r += 1
    if has_u:  # |w_t - w_{t-1}| <= u_t (epigraph of the 1-norms at mpc.py:66, 90)
        for t in range(H):
            for i in range(N):
                for sgn in (+1.0, -1.0):
                    rows.append(r); cols.append(wi(t, i)); vals.append(sgn)
                    if t > 0:
                        rows.append(r); cols.append(wi(t - 1, i)); vals.append(-sgn)
                    rows.append(r); cols.append(ui(t, i)); vals.append(-1.0)
                    h.append(sgn * w_cur[i] if t == 0 else 0.0)
                    r += 1
        if tau > 0:  # delta_t <= max_turnover                                    mpc.py:94-95, 102-103
            for t in range(H):
                for i in range(N):
                    rows.append(r); cols.append(ui(t, i)); vals.append(1.0)
                h.append(tau); r += 1
    G = sp.csr_matrix((vals, (rows, cols)), shape=(r, n))
    A = np.zeros((H, n))
    for t in range(H):  # sum(w_t) == 1                                            mpc.py:83
        A[t, t * N:(t + 1) * N] = 1.0
    return G, np.asarray(h), A, np.ones(H)


def solve_dense(w_cur, yhat, lam, tau, allow_short=False, *, R=None, tol=1e-10, tol_dual=1e-9, max_iter=120):
    """Generic primal-dual interior point (Mehrotra predictor-corrector) with dense KKT solves."""
    w_cur = np.asarray(w_cur, dtype=np.float64)
    R = gross_returns_f32(yhat) if R is None else np.asarray(R, dtype=np.float64)
    H, N = R.shape
    if not (np.all(np.isfinite(R)) and np.all(np.isfinite(w_cur)) and np.all(R > 0)):
        return _finish(None, w_cur, R, lam, STATUS_NONFINITE, 0, (np.nan,) * 3, H)
    has_u = (lam > 0) or (tau > 0)
    init = _initial_point(w_cur, H, N, tau, has_u, allow_short)
    if init is None:
        return _finish(None, w_cur, R, lam, STATUS_MAXITER, 0, (np.inf,) * 3, H)
    w, u = init
    G, h, A, b = _build_constraints(w_cur, H, N, tau, has_u, allow_short)
    GT = G.T.tocsr()
    m = G.shape[0]
    HN = H * N
    x = np.concatenate([w.ravel(), u.ravel()]) if has_u else w.ravel().copy()
    n = x.size
    cvec = np.zeros(n)
    if has_u:
        cvec[HN:] = lam
    nu = np.ones(H)
    s = h - G @ x
    mu0 = 1e-3
    z = mu0 / s if m else np.zeros(0)
    status, res, it = STATUS_MAXITER, (np.inf,) * 3, 0
    best = None
    for it in range(1, max_iter + 1):
        wv = x[:HN].reshape(H, N)
        rho = (wv * R).sum(axis=1)
        grad = cvec.copy()
        grad[:HN] -= (R / rho[:, None]).ravel()
        s = h - G @ x
        r_d = grad + GT @ z + A.T @ nu
        r_p = A @ x - b
        gap = float(s @ z) if m else 0.0
        res = (float(np.abs(r_p).max()), float(np.abs(r_d).max()), gap)
        if res[0] < tol and res[1] < tol_dual and gap < tol:
            status = STATUS_OPTIMAL
            break
        if max(res[0], res[1]) < 1e-7 and gap < 1e-7:
            best = (x.copy(), res)
        mu = gap / max(m, 1)
        Hm = np.zeros((n, n))
        for t in range(H):
            sl = slice(t * N, (t + 1) * N)
            Hm[sl, sl] += np.outer(R[t], R[t]) / rho[t] ** 2
        if m:
            Dg = z / s
            Hm += (GT @ (G.multiply(Dg[:, None]))).toarray()
        KKT = np.block([[Hm, A.T], [A, np.zeros((H, H))]])
        KKT[np.arange(n), np.arange(n)] += 1e-300

        def newton(cterm):
            rhs = np.concatenate([-grad - A.T @ nu - (GT @ (cterm / s) if m else 0.0), -r_p])
            sol = np.linalg.solve(KKT, rhs)
            dx, dnu = sol[:n], sol[n:]
            Gdx = G @ dx if m else np.zeros(0)
            dz = (cterm / s - z) + (z / s) * Gdx if m else np.zeros(0)
            return dx, dnu, dz, -Gdx

        def max_step(ds, dz, dx):
            a = 1.0
            if m:
                neg = ds < 0
                if neg.any():
                    a = min(a, float((-s[neg] / ds[neg]).min()))
                neg = dz < 0
                if neg.any():
                    a = min(a, float((-z[neg] / dz[neg]).min()))
            if allow_short:  # keep the log argument positive
                drho = (dx[:HN].reshape(H, N) * R).sum(axis=1)
                neg = drho < 0
                if neg.any():
                    a = min(a, float((-rho[neg] / drho[neg]).min()))
            return a

        if m:
            dxa, dnua, dza, dsa = newton(np.zeros(m))
            aa = max_step(dsa, dza, dxa)
            mu_aff = float((s + aa * dsa) @ (z + aa * dza)) / m
            # conservative centring floor: the generic solver favours robustness over iteration count
            sigma = max(0.05, min(1.0, (mu_aff / mu) ** 3)) if mu > 0 else 0.0
            cterm = sigma * mu - dsa * dza
        else:
            cterm = np.zeros(0)
        dx, dnu, dz, ds = newton(cterm)
        a = min(1.0, 0.995 * max_step(ds, dz, dx)) if (m or allow_short) else 1.0
        x = x + a * dx
        nu = nu + a * dnu
        if m:
            z = z + a * dz
    else:
        if best is not None:
            x, res = best
            status = STATUS_INACCURATE
    wv = x[:HN].reshape(H, N).copy()
    return _finish(wv, w_cur, R, lam, status, it, res, H)


# ----------------------------------------------------------------------------------------------
# structured IPM (same central path; Newton system solved through the structure)
# ----------------------------------------------------------------------------------------------

def _path_green(a, e):
    """Green's functions of N independent path networks  ground -e[0]- 1 -e[1]- 2 ... -e[H-1]- H  with
    node-to-ground conductances a[k] (all >= 0, shape [H,N]).  T = diag(a) + Laplacian(e) is the SPD
    tridiagonal matrix of the reduced Newton system of one asset.  Returns

      G [l,j]  potential of node l            for a unit current injected at node j   (= T^{-1}[l,j])
      D [l,j]  potential drop v_l - v_{l-1}   for a unit current injected at node j
      DD[l,k]  potential drop v_l - v_{l-1}   for a unit dipole (+1 at node k, -1 at node k-1)

    Everything is built from products/sums of non-negative numbers (series/parallel conductances and
    multiplicative decay factors), so all three are componentwise accurate even when e ~ 1e+12 and
    a ~ 1e-12, which a Thomas solve followed by differencing is not."""
    H, N = a.shape
    inf = np.inf
    hL = np.zeros((H + 1, N)); hL[0] = inf            # hL[k]: conductance to ground seen at node k (k=1..H) leftwards incl. a_k ; hL[0]=inf (ground)
    qL = np.zeros((H + 1, N)); tL = np.zeros((H + 1, N))   # indexed by edge l=1..H (edge l joins node l-1 and l)
    for l in range(1, H + 1):
        el = e[l - 1]
        if l == 1:
            qL[l] = 1.0; tL[l] = 0.0
        else:
            den = el + hL[l - 1]
            qL[l] = hL[l - 1] / den; tL[l] = el / den
        hL[l] = a[l - 1] + el * qL[l]
    hR = np.zeros((H + 2, N))                           # hR[k]: conductance to ground seen at node k rightwards incl. a_k
    qR = np.zeros((H + 2, N)); tR = np.zeros((H + 2, N))  # indexed by edge l (joins node l-1 and l), used when propagating right
    hR[H] = a[H - 1]
    for l in range(H, 1, -1):
        el = e[l - 1]
        den = el + hR[l]
        qR[l] = hR[l] / den; tR[l] = el / den
        hR[l - 1] = a[l - 2] + el * qR[l]
    G = np.zeros((H, H, N)); D = np.zeros((H, H, N)); DD = np.zeros((H, H, N))
    for j in range(1, H + 1):
        gR = e[j] * qR[j + 1] if j < H else 0.0
        gjj = 1.0 / (hL[j] + gR)
        G[j - 1, j - 1] = gjj
        v = gjj
        for l in range(j + 1, H + 1):                   # propagate right
            D[l - 1, j - 1] = -v * qR[l]
            v = v * tR[l]
            G[l - 1, j - 1] = v
        v = gjj
        for l in range(j, 0, -1):                       # propagate left
            D[l - 1, j - 1] = v * qL[l]
            v = v * tL[l]
            if l >= 2:
                G[l - 2, j - 1] = v
    for k in range(1, H + 1):
        ek = e[k - 1]
        if k == 1:
            fL = np.ones(N); fR = np.zeros(N)
        else:
            den = hL[k - 1] + hR[k]
            fL = hL[k - 1] / den; fR = hR[k] / den
        V = 1.0 / (ek + hR[k] * fL)
        DD[k - 1, k - 1] = V
        v = V * fL                                       # potential of node k
        for l in range(k + 1, H + 1):
            DD[l - 1, k - 1] = -v * qR[l]
            v = v * tR[l]
        v = -V * fR                                      # potential of node k-1
        for l in range(k - 1, 0, -1):
            DD[l - 1, k - 1] = v * qL[l]
            v = v * tL[l]
    return G, D, DD


def _path_factors(a, e):
    """Sweep factors of the path networks of _path_green (same series/parallel recurrences), as arrays [H,N]:
    qL,tL (left sweep), qR,tR (right sweep), gjj = T^{-1}[j,j], fL,fR,vd (two-sided split at an edge)."""
    H, N = a.shape
    hL = np.zeros((H, N)); qL = np.ones((H, N)); tL = np.zeros((H, N))
    for k in range(H):
        if k > 0:
            inv = 1.0 / (e[k] + hL[k - 1]); qL[k] = hL[k - 1] * inv; tL[k] = e[k] * inv
        hL[k] = a[k] + e[k] * qL[k]
    hR = np.zeros((H, N)); qR = np.zeros((H, N)); tR = np.zeros((H, N))
    hR[H - 1] = a[H - 1]
    for k in range(H - 1, 0, -1):
        inv = 1.0 / (e[k] + hR[k]); qR[k] = hR[k] * inv; tR[k] = e[k] * inv
        hR[k - 1] = a[k - 1] + e[k] * qR[k]
    gjj = np.zeros((H, N)); fL = np.ones((H, N)); fR = np.zeros((H, N)); vd = np.zeros((H, N))
    for k in range(H):
        gR = e[k + 1] * qR[k + 1] if k + 1 < H else 0.0
        gjj[k] = 1.0 / (hL[k] + gR)
        if k > 0:
            inv = 1.0 / (hL[k - 1] + hR[k]); fL[k] = hL[k - 1] * inv; fR[k] = hR[k] * inv
        vd[k] = 1.0 / (e[k] + hR[k] * fL[k])
    return dict(qL=qL, tL=tL, qR=qR, tR=tR, gjj=gjj, fL=fL, fR=fR, vd=vd)


def _path_sweep(F, gw, pg):
    """(dw, dd) = (G gw - D^T pg, D gw - DD pg) in O(H) per asset instead of O(H^2): Norton equivalents of the
    sources left (JL) and right (JR) of every node, swept once in each direction, then the node potential and the
    drop across every edge from the two-sided split.  Same building blocks (factors in [0,1], no differencing of
    potentials) as the Green's functions; this is what csrc/mpc_lane.cuh applies."""
    H, N = gw.shape
    JL = np.zeros((H, N)); JR = np.zeros((H, N)); inc = np.zeros((H, N))
    for k in range(H):
        JL[k] = gw[k] - F["qL"][k] * pg[k] + (F["tL"][k] * JL[k - 1] if k > 0 else 0.0)
    JR[H - 1] = gw[H - 1]
    for k in range(H - 1, 0, -1):
        inc[k - 1] = F["tR"][k] * JR[k] + F["qR"][k] * pg[k]
        JR[k - 1] = gw[k - 1] + inc[k - 1]
    dw = F["gjj"] * (JL + inc)
    t = F["fL"] * JR - pg
    t[1:] -= F["fR"][1:] * JL[:-1]
    return dw, F["vd"] * t


def solve_structured(w_cur, yhat, lam, tau, allow_short=False, *, second_attempt=True, trace=None, **kw):
    """The solve as the CUDA kernel runs it (csrc/mpc_lane.cuh): the aggressive first attempt and, when that does not end
    "optimal", a second attempt from the cold start with the robust parameters (see ROBUST_* above).  `iters` counts
    the Newton steps of both attempts."""
    r = _solve_structured_once(w_cur, yhat, lam, tau, allow_short, trace=trace, **kw)
    if r.status == STATUS_OPTIMAL or r.status == STATUS_NONFINITE or not second_attempt:
        return r
    kw2 = dict(kw)
    kw2.update(delta=ROBUST_DELTA, step_frac=ROBUST_STEP_FRAC, split_steps=False, sigma_min=ROBUST_SIGMA_MIN)
    r2 = _solve_structured_once(w_cur, yhat, lam, tau, allow_short, trace=trace, **kw2)
    r2["iters"] = r.iters + r2.iters
    return r2


def _solve_structured_once(w_cur, yhat, lam, tau, allow_short=False, *, R=None, tol=1e-10, tol_dual=1e-8,
                           max_iter=100, trace=None, delta=1e-5, split_steps=True, step_frac=0.9999, mu0=1e-3,
                           dual_init=1e-3, apply="green", sigma_min=0.0):
    """Primal-dual IPM (Mehrotra); Newton step = per-asset path-network Green's functions + a (<=3H)
    dense border system.  This is the algorithm the CUDA kernel implements (csrc/mpc_lane.cuh).

    Unknowns per (stage k, asset i): w, u and duals zw (w>=0), zp (sp=u-d>=0), zq (sq=u+d>=0); per
    stage: zc (sc = tau - sum u >= 0) and nu (budget).  d_k = w_k - w_{k-1}, w_0 = w_cur.
    """
    w_cur = np.asarray(w_cur, dtype=np.float64)
    R = gross_returns_f32(yhat) if R is None else np.asarray(R, dtype=np.float64)
    H, N = R.shape
    if not (np.all(np.isfinite(R)) and np.all(np.isfinite(w_cur)) and np.all(R > 0)):
        return _finish(None, w_cur, R, lam, STATUS_NONFINITE, 0, (np.nan,) * 3, H)
    has_u = (lam > 0) or (tau > 0)
    has_c = has_u and tau > 0
    has_w = not allow_short
    init = _initial_point(w_cur, H, N, tau, has_u, allow_short)
    if init is None:
        return _finish(None, w_cur, R, lam, STATUS_MAXITER, 0, (np.inf,) * 3, H)
    w, u = init
    m = (H * N if has_w else 0) + (2 * H * N if has_u else 0) + (H if has_c else 0)
    nu = np.ones(H)
    zHN = np.zeros((H, N))

    def diff(x, first=None):
        dd = x.copy()
        dd[1:] -= x[:-1]
        if first is not None:
            dd[0] -= first
        return dd

    if has_u:
        d = diff(w, w_cur)
        sp_, sq_ = u - d, u + d
        zp, zq = mu0 / sp_, mu0 / sq_
    else:
        sp_ = sq_ = zp = zq = zHN
    if has_c:
        sc_ = tau - u.sum(axis=1)
        zc = mu0 / sc_
    else:
        sc_ = zc = np.zeros(H)
    zw = mu0 / w if has_w else zHN
    if dual_init and has_w:
        # dual-feasible start: zc = zeta0, zp = zq = (lam + zc)/2  =>  r_du = 0;
        # nu_k = max_i R/rho + kappa, zw = grad_w f + nu > 0      =>  r_dw = 0
        zeta0 = dual_init if has_c else 0.0
        if has_c: zc = np.full(H, zeta0)
        if has_u:
            # without a cap (zeta0 = 0) and a tiny lam the u-duals would start 1e4 x off-centre and collapse to zero
            # (dual residual stuck at lam); floor them at the level the cap dual gives otherwise
            zp = np.full((H, N), 0.5 * max(lam + zeta0, dual_init))
            zq = zp.copy()
        rho0 = (w * R).sum(axis=1)
        gw0 = -R / rho0[:, None]
        nu = (-gw0).max(axis=1) + dual_init
        zw = gw0 + nu[:, None]
    nb = (3 if has_c else 2) * H
    status, res, it = STATUS_MAXITER, (np.inf,) * 3, 0
    n_retry, delta0 = 0, delta
    for it in range(1, max_iter + 2):
        rho = (w * R).sum(axis=1)
        gw = -R / rho[:, None]
        y = zp - zq
        rdw = gw - zw + y + nu[:, None]
        rdw[:-1] -= y[1:]
        rdu = (lam - zp - zq + (zc[:, None] if has_c else 0.0)) if has_u else zHN
        rp = w.sum(axis=1) - 1.0
        gap = float((w * zw).sum()) if has_w else 0.0
        if has_u: gap += float((sp_ * zp).sum() + (sq_ * zq).sum())
        if has_c: gap += float((sc_ * zc).sum())
        res = (float(np.abs(rp).max()), max(float(np.abs(rdw).max()), float(np.abs(rdu).max())), gap)
        if trace is not None:
            trace.append(res)
        if not np.isfinite(res[1] + res[2]):
            break
        if res[0] < tol and res[1] < tol_dual and gap < tol:
            status = STATUS_OPTIMAL
            break
        # flat directions (curvature << delta): the proximal term makes the dual residual crawl at ~delta*|dx|
        # while the gap has long collapsed; the objective is converged, so stop as "optimal_inaccurate" instead
        # of iterating into round-off (mpc.py:113 accepts that status)
        if res[0] < tol and gap < 1e-6 * tol and res[1] < 1e-6:
            status = STATUS_INACCURATE
            break
        if it == max_iter + 1:
            break
        mu = gap / max(m, 1)
        if res[0] < tol and gap < tol:
            # endgame: primal residual and gap have converged, only the dual residual along flat directions is
            # left; shrink the proximal term so that the Newton step is no longer damped there
            delta = max(0.3 * delta, 1e-9)

        Dw0 = zw / w if has_w else zHN
        beta = 1.0 / rho ** 2
        if has_u:
            Dp, Dq = zp / sp_, zq / sq_
            E = Dp + Dq + delta
            F = Dq - Dp
            phi = F / E
            e = (4.0 * Dp * Dq + 2.0 * delta * (Dp + Dq) + delta * delta) / E
        else:
            Dp = Dq = F = phi = zHN
            E = np.ones((H, N))
            e = zHN
        G, D, DD = _path_green(Dw0 + delta, e)
        PF = _path_factors(Dw0 + delta, e) if apply == "sweep" else None
        # border matrix K = V^T M0^{-1} V + diag(1/beta, 0, sc/zc),  V = [Rt_j | 1t_j | et_j]
        K = np.zeros((nb, nb))
        for l in range(H):
            for j in range(H):
                g = G[l, j]
                K[l, j] = (R[l] * R[j] * g).sum()
                K[l, H + j] = K[H + j, l] = (R[l] * g).sum()
                K[H + l, H + j] = g.sum()
                if has_c:
                    dm = -phi[l] * D[l, j]
                    K[2 * H + l, j] = K[j, 2 * H + l] = (dm * R[j]).sum()
                    K[2 * H + l, H + j] = K[H + j, 2 * H + l] = dm.sum()
                    K[2 * H + l, 2 * H + j] = (phi[l] * phi[j] * DD[l, j]).sum() + \
                        ((1.0 / E[l]).sum() if l == j else 0.0)
        for k in range(H):
            K[k, k] += 1.0 / beta[k]
            if has_c:
                K[2 * H + k, 2 * H + k] += sc_[k] / zc[k]
        try:
            Kc = np.linalg.cholesky(K)
        except np.linalg.LinAlgError:
            # breakdown of the border factorisation (barrier weights spanning > 20 decades): the iterate is untouched;
            # retry with a stronger proximal term; from the second breakdown on stop if the loose bar is met (same rule as
            # LaneIpm::check in csrc/mpc_lane.cuh) before giving up
            loose = np.isfinite(res[1] + res[2]) and res[0] < LOOSE_PRES and res[1] < LOOSE_DRES and res[2] < LOOSE_GAP
            if n_retry >= MAX_FACTOR_RETRIES or (n_retry > 0 and loose):      # the first breakdown is always retried
                break
            n_retry += 1
            delta = min(max(delta, delta0) * 30.0, 1e-2)
            continue

        def ksolve(t):
            return np.linalg.solve(Kc.T, np.linalg.solve(Kc, t))

        def m0_solve(g_w, g_u):
            """(dw, dd, du) = M0^{-1} (g_w, g_u) through the Green's functions."""
            pg = phi * g_u                                   # dipole strengths (negated)
            if PF is not None:
                dw, dd = _path_sweep(PF, g_w, pg)
                du = (g_u - F * dd) / E if has_u else zHN
                return dw, dd, du
            dw = np.einsum('ljn,jn->ln', G, g_w) - np.einsum('kln,kn->ln', D, pg)
            dd = np.einsum('ljn,jn->ln', D, g_w) - np.einsum('lkn,kn->ln', DD, pg)
            du = (g_u - F * dd) / E if has_u else zHN
            return dw, dd, du

        def kkt_solve(g_w, g_u, q):
            dw0, dd0, du0 = m0_solve(g_w, g_u)
            t = np.zeros(nb)
            t[:H] = (R * dw0).sum(axis=1)
            t[H:2 * H] = dw0.sum(axis=1) - q
            if has_c:
                t[2 * H:] = du0.sum(axis=1)
            yv = ksolve(t)
            # dx = M0^{-1}(g - V y): fold the border columns into the right-hand side
            g_w2 = g_w - yv[:H, None] * R - yv[H:2 * H, None]
            g_u2 = g_u - yv[2 * H:, None] if has_c else g_u
            dw, dd, du = m0_solve(g_w2, g_u2)
            return dw, dd, du, yv[H:2 * H].copy(), g_u2, (yv[2 * H:].copy() if has_c else None)

        def newton(cw, cp_, cq_, cc_):
            # rhs_x = -grad f - A^T nu - G^T (c/s)
            g_w = -gw - nu[:, None]
            if has_w:
                g_w = g_w + cw / w
            g_u = zHN
            if has_u:
                tq = cp_ / sp_ - cq_ / sq_
                g_w = g_w - tq
                g_w[:-1] += tq[1:]
                g_u = -lam + cp_ / sp_ + cq_ / sq_
                if has_c:
                    g_u = g_u - (cc_ / sc_)[:, None]
            dw, dd, du, dnu, geff, yc = kkt_solve(g_w, g_u, -rp)
            # slack steps without cancellation: dsp = du - dd, dsq = du + dd
            if has_u:
                # E du + F dd = g_u - (zc/sc) sum du  (row of the full system) => use the effective g_u
                dsp = (geff - (2.0 * Dq + delta) * dd) / E
                dsq = (geff + (2.0 * Dp + delta) * dd) / E
                dzp = (cp_ / sp_ - zp) - Dp * dsp
                dzq = (cq_ / sq_ - zq) - Dq * dsq
            else:
                dsp = dsq = dzp = dzq = zHN
            dzw = (cw / w - zw) - Dw0 * dw if has_w else zHN
            if has_c:
                # border multiplier yc = (zc/sc) * sum(du) comes straight from the K solve: never rebuild
                # it as (huge) * (tiny cancelling sum)
                dsc = -yc * sc_ / zc
                dzc = (cc_ / sc_ - zc) + yc
            else:
                dsc = dzc = np.zeros(H)
            return dw, du, dnu, dsp, dsq, dsc, dzw, dzp, dzq, dzc

        def max_step(dw, dsp, dsq, dsc, dzw, dzp, dzq, dzc):
            ap = ad = 1.0

            def lim(v, dv, a0):
                neg = dv < 0
                if np.any(neg):
                    a0 = min(a0, float((-v[neg] / dv[neg]).min()))
                return a0
            if has_w:
                ap = lim(w, dw, ap); ad = lim(zw, dzw, ad)
            if has_u:
                ap = lim(sp_, dsp, ap); ap = lim(sq_, dsq, ap); ad = lim(zp, dzp, ad); ad = lim(zq, dzq, ad)
            if has_c:
                ap = lim(sc_, dsc, ap); ad = lim(zc, dzc, ad)
            if allow_short:
                ap = lim(rho, (dw * R).sum(axis=1), ap)
            if not split_steps:
                ap = ad = min(ap, ad)
            return ap, ad

        zH = np.zeros(H)
        if m:
            dw, du, dnu, dsp, dsq, dsc, dzw, dzp, dzq, dzc = newton(zHN, zHN, zHN, zH)
            aa, ab = max_step(dw, dsp, dsq, dsc, dzw, dzp, dzq, dzc)
            g2 = float(((w + aa * dw) * (zw + ab * dzw)).sum()) if has_w else 0.0
            if has_u:
                g2 += float(((sp_ + aa * dsp) * (zp + ab * dzp)).sum())
                g2 += float(((sq_ + aa * dsq) * (zq + ab * dzq)).sum())
            if has_c:
                g2 += float(((sc_ + aa * dsc) * (zc + ab * dzc)).sum())
            sigma = min(1.0, max(g2 / gap, 0.0)) ** 3 if gap > 0 else 0.0
            sigma = max(sigma, sigma_min)
            sm = sigma * mu
            # Mehrotra's second-order term is only trustworthy when the affine step is long; after a short affine
            # step (< 0.3) it is scaled down in proportion (on 500-asset, 10-stage instances the undamped corrector
            # produced steps of 0.01-0.05 for dozens of iterations: 52-64 -> 34-46 iterations; no effect on the rest)
            dmp = min(1.0, min(aa, ab) / CORRECTOR_FULL_STEP)
            cw = sm - dmp * dw * dzw if has_w else zHN
            cp_ = sm - dmp * dsp * dzp if has_u else zHN
            cq_ = sm - dmp * dsq * dzq if has_u else zHN
            cc_ = sm - dmp * dsc * dzc if has_c else zH
        else:
            cw = cp_ = cq_ = zHN; cc_ = zH
        dw, du, dnu, dsp, dsq, dsc, dzw, dzp, dzq, dzc = newton(cw, cp_, cq_, cc_)
        a_, b_ = max_step(dw, dsp, dsq, dsc, dzw, dzp, dzq, dzc) if (m or allow_short) else (1.0, 1.0)
        a_ = min(1.0, step_frac * a_); b_ = min(1.0, step_frac * b_)
        w = w + a_ * dw
        nu = nu + b_ * dnu
        if has_w: zw = zw + b_ * dzw
        if has_u:
            sp_ = sp_ + a_ * dsp; sq_ = sq_ + a_ * dsq
            zp = zp + b_ * dzp; zq = zq + b_ * dzq
        if has_c:
            sc_ = sc_ + a_ * dsc; zc = zc + b_ * dzc
    # not converged (iteration cap, numerical breakdown): accept the *current* iterate if it meets the loose
    # tolerances ("optimal_inaccurate", mpc.py:113), else fall back to holding the weights (mpc.py:113-115)
    if status != STATUS_OPTIMAL and np.isfinite(res[1] + res[2]) and res[0] < LOOSE_PRES and res[1] < LOOSE_DRES \
            and res[2] < LOOSE_GAP:
        status = STATUS_INACCURATE
    out = _finish(w.copy(), w_cur, R, lam, status, it, res, H)
    out["nu"] = np.asarray(nu, dtype=np.float64).copy()                 # budget duals, cap duals and portfolio returns of the last
    out["zc"] = np.asarray(zc, dtype=np.float64).copy()                 # iterate: what the active-set verification reads
    out["rho"] = (w * R).sum(axis=1)
    return out


# ----------------------------------------------------------------------------------------------
# active-set solve (csrc/mpc_lane_kernels.cuh, backtest_active_kernel)
# ----------------------------------------------------------------------------------------------
HELD_THRESHOLD = 1e-9
VERIFY_TOL = 1e-8


def excluded_asset_violations(yhat, members, lam, nu, zc, rho, tol=VERIFY_TOL):
    """Assets outside `members` whose optimality conditions fail at w_ik = 0 for all k, given the duals of a solution
    of the program restricted to `members`.  Asset i may stay at zero iff there are y_k in [-c_k, c_k], c_k = lam + zc_k
    (subgradients of |w_ik - w_i,k-1| at 0 under the cost and the cap's dual), with
        nu_k - R_ik / rho_k + y_k - y_{k+1} >= 0  for every stage   (stationarity with a non-negative bound dual);
    backwards from y_{H+1} = 0 the smallest admissible y_k = max(-c_k, y_{k+1} - g_k) is optimal for stage k - 1 too."""
    R = gross_returns_f32(yhat)
    H, N = R.shape
    c = lam + np.asarray(zc, dtype=np.float64)
    out = []
    inside = set(int(i) for i in members)
    for i in range(N):
        if i in inside:
            continue
        g = nu - R[:, i] / rho
        yk, bad = 0.0, False
        for k in range(H - 1, -1, -1):
            yk = max(-c[k], yk - g[k])
            if not (yk <= c[k] + tol):
                bad = True
        if bad:
            out.append(i)
    return out


def solve_active_set(w_cur, yhat, lam, tau, *, candidates_per_stage=2, max_active=32, **kw):
    """The reduced solve of the active-set backtest kernel, restated: S = held assets + the best forecasts of each stage,
    the structured interior point on S, the optimality conditions of the excluded assets, repair and re-solve.  Returns
    the result of the last solve with `w` padded to all N assets, `members`, `rounds` (solves) and `iters` summed, or None
    when S outgrows `max_active` (the kernel then hands the backtest to the full-width solver).  Long-only, as the kernel."""
    w_cur = np.asarray(w_cur, dtype=np.float64)
    yhat = np.asarray(yhat)
    H, N = yhat.shape
    S = set(np.nonzero(w_cur > HELD_THRESHOLD)[0].tolist())
    for k in range(H):
        S.update(np.argsort(yhat[k], kind="stable")[N - candidates_per_stage:].tolist() if candidates_per_stage else [])
    iters, rounds = 0, 0
    while True:
        if len(S) > max_active:
            return None
        idx = np.array(sorted(S))
        r = solve_structured(w_cur[idx], yhat[:, idx], lam, tau, False, apply="sweep", **kw)
        iters += r.iters; rounds += 1
        if r.status not in (STATUS_OPTIMAL, STATUS_INACCURATE):
            break
        viol = excluded_asset_violations(yhat, idx, lam if (lam > 0 or tau > 0) else 0.0, r["nu"],
                                         r["zc"] if tau > 0 else np.zeros(H), r["rho"])
        if not viol:
            break
        S.update(viol)
    w = np.zeros((H, N))
    w[:, idx] = r.w
    out = _Result(r)
    out["w"] = w; out["members"] = idx; out["rounds"] = rounds; out["iters"] = iters
    if r.status in (STATUS_OPTIMAL, STATUS_INACCURATE):
        w0 = np.where(w_cur > HELD_THRESHOLD, w_cur, 0.0)
        out["value"] = objective(w, w0, gross_returns_f32(yhat), lam)
    return out


def solve_mpc_log_utility(current_weights, predicted_log_returns, config, method="auto"):
    """Reference-signature wrapper (mpc.py:27-31) around the oracle solvers."""
    yhat = np.asarray(predicted_log_returns)
    H, N = yhat.shape
    lam, tau = float(config.cost_coeff), float(config.max_turnover)
    if method == "auto":
        method = "dense" if H * N <= 60 else "structured"
    fn = solve_dense if method == "dense" else solve_structured
    r = fn(current_weights, yhat, lam, tau, bool(config.allow_short))
    return r.w, {"status": STATUS_NAMES[r.status], "value": r.value, "iters": r.iters, "kkt": r.kkt}


# ----------------------------------------------------------------------------------------------
# mean-variance MPC (mpc.py:119-184), SURVEY section 8f
# ----------------------------------------------------------------------------------------------

def mv_objective(w, w_cur, mu, Sigma, gamma, lam):
    """Maximised objective of mpc.py:176 for a plan w[H,N] (fp64)."""
    w = np.asarray(w, dtype=np.float64)
    prev = np.vstack([np.asarray(w_cur, dtype=np.float64)[None, :], w[:-1]])
    quad = np.einsum('ti,ij,tj->', w, Sigma, w)
    return float((w * mu).sum() - gamma * quad - lam * np.abs(w - prev).sum())


def solve_mv_dense(w_cur, mu, Sigma, gamma, lam, allow_short=False, *, tol=1e-10, tol_dual=1e-9, max_iter=120):
    """maximise sum_t [ w_t.mu_t - gamma w_t' Sigma w_t ] - lam sum_t ||w_t - w_{t-1}||_1   (w_0 = w_cur)
    s.t. sum(w_t) = 1, w_t >= 0 unless allow_short; NO turnover cap (mpc.py:139-176).

    Same generic primal-dual interior point as solve_dense (explicit constraint matrix, dense KKT solves) with the
    quadratic stage cost instead of the logarithm.  Ground truth for csrc/mpc_mv.cuh."""
    w_cur = np.asarray(w_cur, dtype=np.float64)
    mu = np.asarray(mu, dtype=np.float64)
    Sigma = np.asarray(Sigma, dtype=np.float64)
    H, N = mu.shape
    if not (np.all(np.isfinite(mu)) and np.all(np.isfinite(w_cur)) and np.all(np.isfinite(Sigma))):
        return _Result(w=np.tile(w_cur, (H, 1)), value=None, status=STATUS_NONFINITE, iters=0, kkt=(np.nan,) * 3)
    has_u = lam > 0
    w, u = _initial_point(w_cur, H, N, 0.0, has_u, allow_short)
    G, h, A, b = _build_constraints(w_cur, H, N, 0.0, has_u, allow_short)
    GT = G.T.tocsr()
    m = G.shape[0]
    HN = H * N
    x = np.concatenate([w.ravel(), u.ravel()]) if has_u else w.ravel().copy()
    n = x.size
    cvec = np.zeros(n)
    cvec[:HN] = -mu.ravel()
    if has_u:
        cvec[HN:] = lam
    Q = np.zeros((n, n))
    for t in range(H):
        sl = slice(t * N, (t + 1) * N)
        Q[sl, sl] = 2.0 * gamma * Sigma
    nu = np.zeros(H)
    s = h - G @ x
    z = 1e-3 / s if m else np.zeros(0)
    status, res, it = STATUS_MAXITER, (np.inf,) * 3, 0
    for it in range(1, max_iter + 1):
        grad = cvec + Q @ x
        s = h - G @ x
        r_d = grad + (GT @ z if m else 0.0) + A.T @ nu
        r_p = A @ x - b
        gap = float(s @ z) if m else 0.0
        res = (float(np.abs(r_p).max()), float(np.abs(r_d).max()), gap)
        if res[0] < tol and res[1] < tol_dual and gap < tol:
            status = STATUS_OPTIMAL
            break
        mu_c = gap / max(m, 1)
        Hm = Q.copy()
        if m:
            Hm += (GT @ (G.multiply((z / s)[:, None]))).toarray()
        KKT = np.block([[Hm, A.T], [A, np.zeros((H, H))]])
        KKT[np.arange(n), np.arange(n)] += 1e-300

        def newton(cterm):
            rhs = np.concatenate([-grad - A.T @ nu - (GT @ (cterm / s) if m else 0.0), -r_p])
            sol = np.linalg.solve(KKT, rhs)
            dx, dnu = sol[:n], sol[n:]
            Gdx = G @ dx if m else np.zeros(0)
            dz = (cterm / s - z) + (z / s) * Gdx if m else np.zeros(0)
            return dx, dnu, dz, -Gdx

        def max_step(ds, dz):
            a = 1.0
            for v, dv in ((s, ds), (z, dz)):
                neg = dv < 0
                if m and neg.any():
                    a = min(a, float((-v[neg] / dv[neg]).min()))
            return a

        if m:
            dxa, dnua, dza, dsa = newton(np.zeros(m))
            aa = max_step(dsa, dza)
            mu_aff = float((s + aa * dsa) @ (z + aa * dza)) / m
            sigma = max(0.05, min(1.0, (mu_aff / mu_c) ** 3)) if mu_c > 0 else 0.0
            cterm = sigma * mu_c - dsa * dza
        else:
            cterm = np.zeros(0)
        dx, dnu, dz, ds = newton(cterm)
        a = min(1.0, 0.995 * max_step(ds, dz)) if m else 1.0
        x = x + a * dx
        nu = nu + a * dnu
        if m:
            z = z + a * dz
    wv = x[:HN].reshape(H, N).copy()
    if status != STATUS_OPTIMAL and np.isfinite(res[1] + res[2]) and res[0] < LOOSE_PRES and res[1] < LOOSE_DRES \
            and res[2] < LOOSE_GAP:
        status = STATUS_INACCURATE
    if status in (STATUS_OPTIMAL, STATUS_INACCURATE):
        return _Result(w=wv, value=mv_objective(wv, w_cur, mu, Sigma, gamma, lam), status=status, iters=it, kkt=res)
    return _Result(w=np.tile(w_cur, (H, 1)), value=None, status=status, iters=it, kkt=res)
