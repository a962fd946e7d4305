"""MPC oracle: the reference's own tests (reference tests/test_mpc.py T1-T3), the exact optima derived in
SURVEY.md §8c, an LP cross-check, and dense-vs-structured agreement."""
import types

import numpy as np
import pytest
from scipy.optimize import linprog

from oracle import mpc_oracle as mo


def cfg(**kw):
    d = dict(horizon=5, gamma=0.0, cost_coeff=0.001, max_turnover=0.2, allow_short=False, solver="ECOS")
    d.update(kw)
    return types.SimpleNamespace(**d)


@pytest.mark.parametrize("method", ["dense", "structured"])
def test_T1_feasibility(method):
    N, H = 5, 3
    w, info = mo.solve_mpc_log_utility(np.ones(N) / N, np.zeros((H, N)), cfg(horizon=H, cost_coeff=0.0), method)
    assert info["status"] == "optimal" and w.shape == (H, N)
    for t in range(H):
        assert np.isclose(w[t].sum(), 1.0) and np.all(w[t] >= -1e-5)
    assert abs(info["value"]) < 1e-12


@pytest.mark.parametrize("method", ["dense", "structured"])
def test_T2_preference_exact(method):
    R = np.exp(np.array([[0.1, 0.0]]))
    r = (mo.solve_dense if method == "dense" else mo.solve_structured)(np.array([0.5, 0.5]), None, 0.0, 0.2, R=R)
    assert r.status == mo.STATUS_OPTIMAL
    assert r.w[0, 0] > 0.5 > r.w[0, 1]
    assert np.allclose(r.w[0], [0.6, 0.4], atol=1e-8)
    assert abs(r.value - 0.06119156775022542) < 1e-10


@pytest.mark.parametrize("method", ["dense", "structured"])
def test_T3_transaction_costs(method):
    w, info = mo.solve_mpc_log_utility(np.array([1.0, 0.0]), np.array([[0.0, 0.01]]), cfg(horizon=1, cost_coeff=10.0), method)
    assert np.allclose(w[0], [1.0, 0.0], atol=1e-2)
    assert np.allclose(w[0], [1.0, 0.0], atol=1e-8) and abs(info["value"]) < 1e-9


def test_lp_crosscheck_lam0_H1():
    """lam=0, H=1: log is monotone, so the optimiser maximises R.w over simplex ∩ L1-ball -> an LP."""
    rng = np.random.default_rng(7)
    for N in (4, 9, 20):
        w0 = rng.dirichlet(np.ones(N))
        R = np.exp(rng.standard_normal((1, N)) * 0.02)
        tau = 0.3
        # variables [w, u]: max R.w  s.t. sum w = 1, |w - w0| <= u, sum u <= tau, w >= 0
        c = np.concatenate([-R[0], np.zeros(N)])
        A_ub = np.block([[np.eye(N), -np.eye(N)], [-np.eye(N), -np.eye(N)], [np.zeros((1, N)), np.ones((1, N))]])
        b_ub = np.concatenate([w0, -w0, [tau]])
        lp = linprog(c, A_ub=A_ub, b_ub=b_ub, A_eq=np.concatenate([np.ones(N), np.zeros(N)])[None], b_eq=[1.0],
                     bounds=[(0, None)] * (2 * N), method="highs")
        for fn in (mo.solve_dense, mo.solve_structured):
            r = fn(w0, None, 0.0, tau, R=R)
            assert r.status == mo.STATUS_OPTIMAL
            assert abs(np.log(-lp.fun) - r.value) < 1e-9


def test_dense_vs_structured_random():
    rng = np.random.default_rng(1)
    worst_obj = worst_w = 0.0
    for (N, H) in [(3, 2), (5, 3), (10, 5), (20, 5)]:
        for trial in range(6):
            w0 = rng.dirichlet(np.ones(N) * rng.choice([0.3, 1, 5]))
            y = (rng.standard_normal((H, N)) * rng.choice([0.003, 0.01, 0.05])).astype(np.float32)
            lam, tau = (1e-3, 0.2) if trial < 3 else (rng.choice([0, 1e-4, 1e-2]), rng.choice([0.05, 1.0, 0.0]))
            a, b = mo.solve_dense(w0, y, lam, tau), mo.solve_structured(w0, y, lam, tau)
            assert a.status == mo.STATUS_OPTIMAL and b.status == mo.STATUS_OPTIMAL
            worst_obj = max(worst_obj, abs(a.value - b.value) / max(abs(a.value), 1e-3))
            worst_w = max(worst_w, np.abs(a.w - b.w).max())
            # feasibility of the structured solution
            assert np.allclose(b.w.sum(axis=1), 1.0, atol=1e-9) and b.w.min() > -1e-12
            if tau > 0:
                d = np.abs(np.diff(np.vstack([w0, b.w]), axis=0)).sum(axis=1)
                assert d.max() <= tau + 1e-8
    assert worst_obj < 1e-6, worst_obj
    assert worst_w < 1e-4, worst_w


def test_fallback_on_nonfinite():
    y = np.zeros((2, 3), np.float32); y[0, 1] = np.nan
    w0 = np.array([0.2, 0.3, 0.5])
    w, info = mo.solve_mpc_log_utility(w0, y, cfg(horizon=2))
    assert info["value"] is None and info["status"] not in ("optimal", "optimal_inaccurate")
    assert np.array_equal(w, np.tile(w0, (2, 1)))


def test_allow_short_and_no_cap():
    rng = np.random.default_rng(5)
    w0 = rng.dirichlet(np.ones(4)); y = (rng.standard_normal((2, 4)) * 0.01).astype(np.float32)
    a = mo.solve_dense(w0, y, 5e-3, 0.3, allow_short=True)
    b = mo.solve_structured(w0, y, 5e-3, 0.3, allow_short=True)
    assert a.status == 0 and b.status == 0 and abs(a.value - b.value) < 1e-8
    a = mo.solve_dense(w0, y, 2e-2, 0.0); b = mo.solve_structured(w0, y, 2e-2, 0.0)
    assert a.status == 0 and b.status == 0 and abs(a.value - b.value) < 1e-8


def test_sweep_apply_matches_green_apply():
    """solve_structured(apply="sweep") — the O(H) two-sided Norton sweep that csrc/mpc_lane.cuh uses for M0^{-1} —
    walks the same central path as the explicit Green's functions: same status, iteration count, objective, plan."""
    rng = np.random.default_rng(5)
    for p in range(40):
        N = int(rng.choice([5, 10, 50])); H = int(rng.choice([1, 3, 5]))
        w0 = rng.dirichlet(np.ones(N) * 0.5)
        y = (3e-4 + rng.standard_normal((H, N)) * 0.01).astype(np.float32)
        lam = float(10 ** rng.uniform(-5, -1)); tau = float(rng.choice([0.0, 0.05, 0.2, 1.0]))
        a = mo.solve_structured(w0, y, lam, tau)
        b = mo.solve_structured(w0, y, lam, tau, apply="sweep")
        assert a.status == b.status and a.iters == b.iters
        if a.value is not None:
            assert abs(a.value - b.value) < 1e-10
            assert np.abs(a.w - b.w).max() < 1e-8


def test_path_sweep_componentwise_accuracy():
    """_path_sweep against the explicit Green's functions in extended precision, conductances spanning 24 decades:
    errors stay at round-off relative to the sum of the absolute terms (no cancellation from differencing)."""
    rng = np.random.default_rng(0)
    H, N = 5, 4000
    a = 10.0 ** rng.uniform(-12, 12, (H, N)); e = 10.0 ** rng.uniform(-12, 12, (H, N))
    gw = rng.standard_normal((H, N)); pg = rng.standard_normal((H, N))
    ld = np.longdouble
    G, D, DD = mo._path_green(a.astype(ld), e.astype(ld))
    gwl, pgl = gw.astype(ld), pg.astype(ld)
    dw_ref = np.einsum('ljn,jn->ln', G, gwl) - np.einsum('kln,kn->ln', D, pgl)
    dd_ref = np.einsum('ljn,jn->ln', D, gwl) - np.einsum('lkn,kn->ln', DD, pgl)
    sw = np.einsum('ljn,jn->ln', np.abs(G), np.abs(gwl)) + np.einsum('kln,kn->ln', np.abs(D), np.abs(pgl))
    sd = np.einsum('ljn,jn->ln', np.abs(D), np.abs(gwl)) + np.einsum('lkn,kn->ln', np.abs(DD), np.abs(pgl))
    dw, dd = mo._path_sweep(mo._path_factors(a, e), gw, pg)
    assert float(np.max(np.abs(dw - dw_ref) / sw)) < 5e-15
    assert float(np.max(np.abs(dd - dd_ref) / sd)) < 5e-15


def test_mean_variance_oracle_known_answers():
    """solve_mv_dense (restatement of mpc.py:119-184) pinned by a closed form (lam = 0, shorting allowed, H = 1:
    w = Sigma^{-1}(mu - nu 1) / (2 gamma)) and by SLSQP on the smooth epigraph form of small instances."""
    from scipy.optimize import minimize
    rng = np.random.default_rng(1)
    N = 6
    X = rng.standard_normal((60, N)) * 0.01
    S = np.cov(X, rowvar=False) + 1e-6 * np.eye(N); mu = X.mean(0); g = 5.0
    r = mo.solve_mv_dense(np.ones(N) / N, mu[None], S, g, 0.0, allow_short=True)
    Si = np.linalg.inv(S); one = np.ones(N)
    nu = (one @ Si @ mu - 2 * g) / (one @ Si @ one)
    assert r.status == 0 and np.abs(r.w[0] - Si @ (mu - nu) / (2 * g)).max() < 1e-12
    for trial in range(2):
        N, H = 4, 2
        X = rng.standard_normal((30, N)) * 0.01
        S = np.cov(X, rowvar=False) + 1e-6 * np.eye(N); mu = rng.standard_normal((H, N)) * 1e-3
        w0 = rng.dirichlet(np.ones(N)); lam = 1e-3; g = 2.0
        r = mo.solve_mv_dense(w0, mu, S, g, lam)

        def f(xv):
            w = xv[:H * N].reshape(H, N); u = xv[H * N:]
            return -((w * mu).sum() - g * np.einsum('ti,ij,tj->', w, S, w) - lam * u.sum())

        def ineq(xv):
            w = xv[:H * N].reshape(H, N); u = xv[H * N:].reshape(H, N)
            d = w - np.vstack([w0[None], w[:-1]])
            return np.concatenate([w.ravel(), (u - d).ravel(), (u + d).ravel()])
        cons = [{'type': 'eq', 'fun': (lambda xv, t=t: xv[t * N:(t + 1) * N].sum() - 1)} for t in range(H)]
        cons.append({'type': 'ineq', 'fun': ineq})
        sol = minimize(f, np.concatenate([np.tile(w0, H), np.full(H * N, 0.1)]), constraints=cons, method='SLSQP',
                       options={'ftol': 1e-14, 'maxiter': 500})
        assert r.status == 0 and abs(r.value + sol.fun) < 1e-9
        assert np.abs(r.w - sol.x[:H * N].reshape(H, N)).max() < 1e-6
        assert abs(mo.mv_objective(r.w, w0, mu, S, g, lam) - r.value) < 1e-15


# ---- solver-independent optimality certificate (oracle/mpc_certificate.py) ------------------------------------------

def test_certificate_on_reference_known_answers():
    """T2 / T3 of the reference's tests/test_mpc.py at their exact optima: zero certified gap; a perturbed plan: a
    positive gap of the size of its objective loss."""
    from oracle import mpc_certificate as mc
    R = np.exp(np.array([[0.1, 0.0]]))
    c = mc.certify(np.array([[0.6, 0.4]]), np.array([0.5, 0.5]), R, 0.0, 0.2)
    assert abs(c["value"] - 0.06119156775022542) < 1e-12 and abs(c["gap"]) < 1e-9 and max(c["feas"]) < 1e-12
    c2 = mc.certify(np.array([[0.55, 0.45]]), np.array([0.5, 0.5]), R, 0.0, 0.2)
    assert c2["gap"] > 0.9 * (c["value"] - c2["value"]) > 0
    R3 = np.exp(np.array([[0.0, 0.01]]))
    c3 = mc.certify(np.array([[1.0, 0.0]]), np.array([1.0, 0.0]), R3, 10.0, 0.2)
    assert abs(c3["value"]) < 1e-12 and abs(c3["gap"]) < 1e-9


@pytest.mark.parametrize("N,H", [(3, 2), (10, 5), (50, 5)])
def test_certificate_bounds_both_oracles(N, H):
    """value <= optimum <= upper for the plans of BOTH oracle solvers, within the parity bar (1e-6 relative), over mixed
    lambda / tau including the uncapped and the cost-free cases; and the certificate rejects a plan that solves a
    DIFFERENT program (cap ignored / cost ignored), which a twin-vs-twin comparison would not."""
    from oracle import mpc_certificate as mc
    rng = np.random.default_rng(31 * N + H)
    for p in range(6):
        w0 = rng.dirichlet(np.ones(N) * rng.choice([0.3, 1.0, 5.0]))
        y = (3e-4 + rng.standard_normal((H, N)) * rng.choice([0.003, 0.01, 0.03])).astype(np.float32)
        lam = float(rng.choice([1e-3, 0.0, 1e-4, 1e-2])); tau = float(rng.choice([0.2, 0.05, 1.0, 0.0]))
        R = mo.gross_returns_f32(y)
        for fn in (mo.solve_structured, mo.solve_dense):
            r = fn(w0, y, lam, tau)
            assert r.status == mo.STATUS_OPTIMAL
            c = mc.certify(r.w, w0, R, lam, tau)
            assert abs(c["value"] - r.value) < 1e-12
            assert -1e-9 < c["gap"] < 1e-6 * max(abs(c["value"]), 1e-3), (fn.__name__, p, c)
            assert c["feas"][0] < 1e-9 and c["feas"][1] < 1e-10 and c["feas"][2] < 1e-9
    # a plan optimal for a different program is NOT certified for this one
    w0 = rng.dirichlet(np.ones(N)); y = (rng.standard_normal((H, N)) * 0.02).astype(np.float32)
    R = mo.gross_returns_f32(y)
    no_cost = mo.solve_structured(w0, y, 0.0, 1.0)
    c = mc.certify(no_cost.w, w0, R, 2e-2, 1.0)                   # judged with a cost it ignored
    assert c["gap"] > 1e-4
    uncapped = mo.solve_structured(w0, y, 1e-4, 0.0)
    assert mc.feasibility(uncapped.w, w0, 0.05)[2] > 1e-3         # judged against a cap it ignored


@pytest.mark.parametrize("cands", [2, 0])
def test_active_set_restatement_reaches_the_full_optimum(cands):
    """oracle.solve_active_set (the reduced solve of csrc/mpc_lane_kernels.cuh::backtest_active_kernel, restated): the
    program restricted to the held assets (+ the best forecasts of each stage) followed by the optimality check of the
    excluded assets and repair reaches the optimum of the FULL program — against the full structured solve (objective
    1e-8) and, solver-independently, against the certificate (1e-6 relative, the parity bar) — over concentrated
    portfolios, mixed costs and caps including the uncapped and the cost-free cases.  cands = 0 starts from the held
    assets alone, so that every asset of the plan has to come in through the check."""
    from oracle import mpc_certificate as mc
    rng = np.random.default_rng(97 + cands)
    rounds = 0
    for p in range(14):
        N = int(rng.choice([40, 50, 64])); H = int(rng.choice([3, 5]))
        k = int(rng.integers(1, 7))
        w0 = np.zeros(N); w0[rng.choice(N, k, replace=False)] = rng.dirichlet(np.ones(k))
        drift = rng.standard_normal(N) * rng.choice([0.001, 0.004])
        y = (3e-4 + drift + rng.standard_normal((H, N)) * rng.choice([2e-4, 2e-3])).astype(np.float32)
        lam = float(rng.choice([1e-3, 1e-3, 1e-4, 0.0, 1e-2])); tau = float(rng.choice([0.2, 0.2, 0.05, 1.0, 0.0]))
        full = mo.solve_structured(w0, y, lam, tau, apply="sweep")
        act = mo.solve_active_set(w0, y, lam, tau, candidates_per_stage=cands, max_active=32 if cands else N)   # (the 32 lanes of the kernel's warp are not the point of the held-only variant)
        assert full.status == mo.STATUS_OPTIMAL and act is not None and act.status == mo.STATUS_OPTIMAL
        assert abs(act.value - full.value) < 1e-8 * max(abs(full.value), 1e-3), (p, act.value, full.value)
        c = mc.certify(act.w, w0, mo.gross_returns_f32(y), lam, tau)
        assert -1e-9 < c["gap"] < 1e-6 * max(abs(c["value"]), 1e-3), (p, c)
        assert c["feas"][0] < 1e-9 and c["feas"][1] < 1e-10 and c["feas"][2] < 1e-9
        assert np.all(act.w[:, np.setdiff1d(np.arange(N), act.members)] == 0.0)
        rounds += act.rounds - 1
    if cands == 0:
        assert rounds > 0                                        # the check-and-repair path really ran


def test_active_set_restatement_wide_universes():
    """The same restatement at the shape of the WIDE reduced-solve kernel (universes of 129..512 assets, up to 256 active
    assets: backtest_active_kernel<H, 1, FIX, 2, 8>): portfolios that still hold 60-220 of 300 / 500 assets, H = 5 and 10,
    against the full structured solve and the certificate."""
    from oracle import mpc_certificate as mc
    rng = np.random.default_rng(311)
    for p, (N, H) in enumerate([(300, 5), (300, 5), (500, 10), (200, 10)]):
        k = int(rng.integers(60, min(220, N - 20)))
        w0 = np.zeros(N); w0[rng.choice(N, k, replace=False)] = rng.dirichlet(np.ones(k))
        drift = rng.standard_normal(N) * 0.002
        y = (3e-4 + drift + rng.standard_normal((H, N)) * 1e-3).astype(np.float32)
        lam, tau = (1e-3, 0.2) if p % 2 == 0 else (1e-4, 0.5)
        full = mo.solve_structured(w0, y, lam, tau, apply="sweep")
        act = mo.solve_active_set(w0, y, lam, tau, max_active=256)
        assert full.status == mo.STATUS_OPTIMAL and act is not None and act.status == mo.STATUS_OPTIMAL
        assert len(act.members) <= 256 and len(act.members) < N
        assert abs(act.value - full.value) < 1e-8 * max(abs(full.value), 1e-3), (p, act.value, full.value)
        c = mc.certify(act.w, w0, mo.gross_returns_f32(y), lam, tau)
        assert -1e-9 < c["gap"] < 1e-6 * max(abs(c["value"]), 1e-3), (p, c)
        assert np.all(act.w[:, np.setdiff1d(np.arange(N), act.members)] == 0.0)


def test_second_attempt_solves_the_stalling_decisions(golden):
    """The 52 decisions of a config-2 step (1.0 M decisions, collected on the GPU with scripts/find_failures.py) that the
    aggressive first attempt leaves `optimal_inaccurate`: near-degenerate optima, the dual residual stalls at 1e-7..1e-5
    while the gap collapses, a few plans 2-5e-6 off the optimal objective.  The second attempt (robust parameters, see
    ROBUST_* in mpc_oracle.py) must reach `optimal` on every one, within the objective bar of the dense oracle (values
    stored in the fixture) and certified by the LP certificate."""
    from oracle import mpc_certificate as mc
    g = golden("stall_instances.npz")
    n_first = 0
    for p in range(0, len(g["w"]), 3):
        first = mo.solve_structured(g["w"][p], g["y"][p], 1e-3, 0.2, apply="sweep", second_attempt=False)
        n_first += first.status != mo.STATUS_OPTIMAL
        r = mo.solve_structured(g["w"][p], g["y"][p], 1e-3, 0.2, apply="sweep")
        assert r.status == mo.STATUS_OPTIMAL, (p, r.status, r.kkt)
        assert abs(r.value - g["value"][p]) <= 1e-6 * max(abs(g["value"][p]), 1e-3)
        c = mc.certify(r.w, g["w"][p], mo.gross_returns_f32(g["y"][p]), 1e-3, 0.2)
        assert c["gap"] < 1e-6 * max(abs(c["value"]), 1e-3)
    assert n_first >= 5          # the fixture still exercises the path: the first attempt alone stalls on many of them
