"""CUDA MPC solver (through the C ABI) vs the fp64 oracle.  Parity bar (BASELINE.json north_star): objective
within 1e-6 relative, weights within 1e-4 L-inf (where the optimiser is well conditioned), KKT residuals reported."""
import types

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


OBJ_RTOL = 1e-6      # |obj_gpu - obj_oracle| <= OBJ_RTOL * max(|obj_oracle|, OBJ_FLOOR)
OBJ_FLOOR = 1e-3     # objectives are sums of daily log-growth; below 1e-3 the bar is absolute 1e-9
W_ATOL = 1e-4


def _mods():
    import torch
    from koopman_mpc_portfolio_rebalancing_b200 import mpc
    from oracle import mpc_oracle as mo
    return torch, mpc, mo


def _certified_gap(W, w0, y, lam, tau, allow_short=False):
    """solver-independent bound on the objective error of plan W against the optimum of mpc.py's program
    (oracle/mpc_certificate.py: concavity + one HiGHS LP), relative like the parity bar; also asserts feasibility"""
    from oracle import mpc_certificate as mc, mpc_oracle as mo
    c = mc.certify(W, w0, mo.gross_returns_f32(y), float(lam), float(tau), allow_short)
    assert c["feas"][0] < 1e-9 and c["feas"][1] < 1e-10, c["feas"]
    assert c["gap"] > -1e-9, c                                   # value <= optimum <= upper
    return c["gap"] / max(abs(c["value"]), OBJ_FLOOR)


def test_reference_tests_T1_T2_T3():
    """reference tests/test_mpc.py, verbatim semantics, through the drop-in signature"""
    torch, mpc, mo = _mods()
    N, H = 5, 3
    w, info = mpc.solve_mpc_log_utility(np.ones(N) / N, np.zeros((H, N)), mpc.MPCConfig(horizon=H, cost_coeff=0.0))
    assert info["status"] == "optimal" and w.shape == (H, N)
    for t in range(H):
        assert np.isclose(np.sum(w[t]), 1.0) and np.all(w[t] >= -1e-5)
    w, info = mpc.solve_mpc_log_utility(np.array([0.5, 0.5]), np.array([[0.1, 0.0]]), mpc.MPCConfig(horizon=1, cost_coeff=0.0))
    assert w[0, 0] > 0.5 and w[0, 1] < 0.5
    assert np.allclose(w[0], [0.6, 0.4], atol=1e-7) and abs(info["value"] - 0.06119156775022542) < 1e-9
    w, info = mpc.solve_mpc_log_utility(np.array([1.0, 0.0]), np.array([[0.0, 0.01]]), mpc.MPCConfig(horizon=1, cost_coeff=10.0))
    assert np.allclose(w[0], [1.0, 0.0], atol=1e-2)
    assert np.allclose(w[0], [1.0, 0.0], atol=1e-7) and abs(info["value"]) < 1e-8


def test_fallback_never_raises():
    torch, mpc, mo = _mods()
    y = np.zeros((2, 3), np.float32); y[0, 1] = np.nan
    w0 = np.array([0.2, 0.3, 0.5])
    w, info = mpc.solve_mpc_log_utility(w0, y, mpc.MPCConfig(horizon=2))
    assert info["value"] is None and info["status"] not in ("optimal", "optimal_inaccurate")
    assert np.array_equal(w, np.tile(w0, (2, 1)))


@pytest.mark.parametrize("N,H", [(2, 1), (5, 3), (10, 5), (33, 2), (50, 5), (64, 4)])
def test_random_instances_vs_oracle(N, H):
    """48 instances per shape with mixed per-problem lambda / tau.  Three arbiters: the numpy twin of the kernel
    (solve_structured: objective AND weights), on every third instance the independent dense interior-point oracle
    (solve_dense: explicit constraint matrix of mpc.py), and on EVERY instance the solver-independent optimality
    certificate (certified objective error against the optimum of mpc.py's program, no IPM involved)."""
    torch, mpc, mo = _mods()
    rng = np.random.default_rng(100 * N + H)
    P = 48
    w0 = np.stack([rng.dirichlet(np.ones(N) * rng.choice([0.3, 1.0, 5.0])) for _ in range(P)])
    y = np.stack([(3e-4 + rng.standard_normal((H, N)) * rng.choice([0.003, 0.01, 0.03])) for _ in range(P)]).astype(np.float32)
    lam = np.where(rng.random(P) < 0.5, 1e-3, rng.choice([0.0, 1e-4, 1e-2], P))
    tau = np.where(rng.random(P) < 0.5, 0.2, rng.choice([0.05, 1.0, 0.0], P))
    out = mpc.solve_mpc_batch(torch.from_numpy(w0).cuda(), torch.from_numpy(y).cuda(),
                              lam=torch.from_numpy(lam).cuda(), tau=torch.from_numpy(tau).cuda())
    W = out["w"].cpu().numpy(); val = out["value"].cpu().numpy(); st = out["status"].cpu().numpy()
    kkt = out["kkt"].cpu().numpy(); its = out["iterations"].cpu().numpy()
    worst_obj = worst_w = worst_dense = worst_cert = 0.0
    for p in range(P):
        ref = mo.solve_structured(w0[p], y[p], float(lam[p]), float(tau[p]))
        assert ref.status == mo.STATUS_OPTIMAL
        assert st[p] == 0, (p, st[p], kkt[p], its[p])
        worst_obj = max(worst_obj, abs(val[p] - ref.value) / max(abs(ref.value), OBJ_FLOOR))
        worst_w = max(worst_w, np.abs(W[p] - ref.w).max())
        if p % 3 == 0:
            dense = mo.solve_dense(w0[p], y[p], float(lam[p]), float(tau[p]))
            assert dense.status == mo.STATUS_OPTIMAL
            worst_dense = max(worst_dense, abs(val[p] - dense.value) / max(abs(dense.value), OBJ_FLOOR))
        worst_cert = max(worst_cert, _certified_gap(W[p], w0[p], y[p], lam[p], tau[p]))
        assert np.allclose(W[p].sum(axis=1), 1.0, atol=1e-8) and W[p].min() > -1e-10
        if tau[p] > 0:
            turn = np.abs(np.diff(np.vstack([w0[p], W[p]]), axis=0)).sum(axis=1)
            assert turn.max() <= tau[p] + 1e-7
        assert kkt[p, 0] < 1e-8 and kkt[p, 1] < 1e-6 and kkt[p, 2] < 1e-8       # reported KKT residuals
    print(f"N={N} H={H}: worst rel obj gap {worst_obj:.2e} (twin) {worst_dense:.2e} (dense), certified {worst_cert:.2e}, "
          f"worst |dw|_inf {worst_w:.2e}, mean iters {its.mean():.1f}")
    assert worst_obj < OBJ_RTOL, worst_obj
    assert worst_dense < OBJ_RTOL, worst_dense
    assert worst_cert < OBJ_RTOL, worst_cert
    assert worst_w < W_ATOL, worst_w


def test_wide_parity_sweep_dense_and_certificate():
    """The wide random mix (round 1 ran it as a script, now a test): shapes N 2..64, H 1..5, lambda in {0, 1e-5..1e-1}, tau in
    {0, 0.01..1}, concentrated and diffuse current weights, calm and wild forecasts; 24 instances per shape.  Every
    accepted plan is certified against the optimum of mpc.py's program (certificate) and every fourth one is compared
    with the independent dense oracle; no instance may fall back."""
    torch, mpc, mo = _mods()
    rng = np.random.default_rng(2024)
    shapes = [(2, 1), (3, 5), (7, 2), (10, 5), (17, 3), (32, 5), (33, 4), (50, 5), (64, 5), (64, 1)]
    P = 24
    n_inacc = 0
    for (N, H) in shapes:
        w0 = np.stack([rng.dirichlet(np.ones(N) * rng.choice([0.05, 0.3, 1.0, 5.0])) for _ in range(P)])
        y = np.stack([(3e-4 + rng.standard_normal((H, N)) * rng.choice([0.001, 0.003, 0.01, 0.03, 0.1])) for _ in range(P)]).astype(np.float32)
        lam = rng.choice([0.0, 1e-5, 1e-4, 1e-3, 1e-2, 1e-1], P)
        tau = rng.choice([0.0, 0.01, 0.05, 0.2, 0.5, 1.0], P)
        out = mpc.solve_mpc_batch(torch.from_numpy(w0).cuda(), torch.from_numpy(y).cuda(), lam=torch.from_numpy(lam).cuda(),
                                  tau=torch.from_numpy(tau).cuda())
        W = out["w"].cpu().numpy(); val = out["value"].cpu().numpy(); st = out["status"].cpu().numpy()
        assert (st <= 1).all(), (N, H, st)
        n_inacc += int((st == 1).sum())
        for p in range(P):
            cert = _certified_gap(W[p], w0[p], y[p], lam[p], tau[p])
            assert cert < OBJ_RTOL, (N, H, p, lam[p], tau[p], st[p], cert)
            if tau[p] > 0:
                assert np.abs(W[p][0] - w0[p]).sum() <= tau[p] + 1e-9
            if p % 4 == 0:
                dense = mo.solve_dense(w0[p], y[p], float(lam[p]), float(tau[p]))
                if dense.status == mo.STATUS_OPTIMAL:
                    assert abs(val[p] - dense.value) <= OBJ_RTOL * max(abs(dense.value), OBJ_FLOOR), (N, H, p)
    assert n_inacc <= 4, n_inacc


def test_dense_oracle_crosscheck_small():
    """independent algorithm (generic dense IPM) on a few instances: objective parity"""
    torch, mpc, mo = _mods()
    rng = np.random.default_rng(9)
    for (N, H) in [(4, 2), (10, 5), (20, 5)]:
        w0 = rng.dirichlet(np.ones(N)); y = (rng.standard_normal((H, N)) * 0.01).astype(np.float32)
        w, info = mpc.solve_mpc_log_utility(w0, y, mpc.MPCConfig(horizon=H))
        ref = mo.solve_dense(w0, y, 1e-3, 0.2)
        assert info["status"] == "optimal" and ref.status == 0
        assert abs(info["value"] - ref.value) <= OBJ_RTOL * max(abs(ref.value), OBJ_FLOOR)


def test_allow_short_and_uncapped():
    torch, mpc, mo = _mods()
    rng = np.random.default_rng(5)
    w0 = rng.dirichlet(np.ones(6)); y = (rng.standard_normal((3, 6)) * 0.01).astype(np.float32)
    for kw in (dict(cost_coeff=5e-3, max_turnover=0.3, allow_short=True), dict(cost_coeff=2e-2, max_turnover=0.0)):
        w, info = mpc.solve_mpc_log_utility(w0, y, mpc.MPCConfig(horizon=3, **kw))
        ref = mo.solve_structured(w0, y, kw["cost_coeff"], kw["max_turnover"], kw.get("allow_short", False))
        assert info["status"] == "optimal"
        assert abs(info["value"] - ref.value) <= OBJ_RTOL * max(abs(ref.value), OBJ_FLOOR)


def test_unsupported_shape_fails_loudly():
    torch, mpc, mo = _mods()
    from koopman_mpc_portfolio_rebalancing_b200 import _capi
    with pytest.raises(_capi.KmpcError):
        mpc.solve_mpc_log_utility(np.ones(700) / 700, np.zeros((5, 700), np.float32), mpc.MPCConfig())


def test_config5_shape_100_assets():
    """N = 100, H = 5 (Monte-Carlo stress-test shape, BASELINE config 5): CTA layout with four lane groups"""
    torch, mpc, mo = _mods()
    from koopman_mpc_portfolio_rebalancing_b200 import _capi
    if _capi.lib().kmpc_mpc_supported(5, 100) != 1:
        pytest.skip("no kernel variant for N=100 in this layout")
    rng = np.random.default_rng(77)
    P, N, H = 12, 100, 5
    w0 = np.stack([rng.dirichlet(np.ones(N) * 0.5) for _ in range(P)])
    y = (3e-4 + rng.standard_normal((P, H, N)) * 0.01).astype(np.float32)
    out = mpc.solve_mpc_batch(torch.from_numpy(w0).cuda(), torch.from_numpy(y).cuda())
    W = out["w"].cpu().numpy(); val = out["value"].cpu().numpy(); st = out["status"].cpu().numpy()
    for p in range(P):
        ref = mo.solve_structured(w0[p], y[p], 1e-3, 0.2)
        assert st[p] == 0 and ref.status == 0
        assert abs(val[p] - ref.value) <= OBJ_RTOL * max(abs(ref.value), OBJ_FLOOR)
        assert np.abs(W[p] - ref.w).max() < W_ATOL


@pytest.mark.parametrize("N,H", [(12, 10), (50, 10), (100, 10), (500, 10), (300, 5)])
def test_large_shapes_lane_layout(N, H):
    """BASELINE config 3 (LISTAKM, 500 assets, H = 10, turnover cap) and the other shapes only the lane layout
    covers (H up to 10, N up to 512; factors and targets thread-private, see LaneIpm LOC)."""
    torch, mpc, mo = _mods()
    from koopman_mpc_portfolio_rebalancing_b200 import _capi
    assert _capi.lib().kmpc_mpc_supported(H, N) == 1
    rng = np.random.default_rng(1000 * H + N)
    P = 6 if N >= 300 else 16
    w0 = np.stack([rng.dirichlet(np.ones(N) * rng.choice([0.3, 1.0])) for _ in range(P)])
    y = np.stack([(3e-4 + rng.standard_normal((H, N)) * rng.choice([0.003, 0.01])) for _ in range(P)]).astype(np.float32)
    out = mpc.solve_mpc_batch(torch.from_numpy(w0).cuda(), torch.from_numpy(y).cuda())
    W = out["w"].cpu().numpy(); val = out["value"].cpu().numpy(); st = out["status"].cpu().numpy()
    kkt = out["kkt"].cpu().numpy()
    for p in range(P):
        ref = mo.solve_structured(w0[p], y[p], 1e-3, 0.2, apply="sweep")
        assert ref.status == mo.STATUS_OPTIMAL and st[p] == 0, (p, st[p], kkt[p])
        assert abs(val[p] - ref.value) <= OBJ_RTOL * max(abs(ref.value), OBJ_FLOOR)
        assert np.abs(W[p] - ref.w).max() < W_ATOL
        turn = np.abs(np.diff(np.vstack([w0[p], W[p]]), axis=0)).sum(axis=1)
        assert turn.max() <= 0.2 + 1e-7


def test_uncapped_small_lambda_converges():
    """max_turnover <= 0 means no cap (mpc.py:94): with a tiny cost_coeff the u-duals of the dual-feasible start
    used to collapse (dual residual stuck at lam, 12 % of such instances failed).  All must reach the oracle optimum."""
    torch, mpc, mo = _mods()
    rng = np.random.default_rng(404)
    P, N, H = 40, 50, 3
    w0 = np.stack([rng.dirichlet(np.ones(N) * 0.5) for _ in range(P)])
    y = (3e-4 + rng.standard_normal((P, H, N)) * 0.01).astype(np.float32)
    lam = 10 ** rng.uniform(-5, -3, P); tau = np.zeros(P)
    out = mpc.solve_mpc_batch(torch.from_numpy(w0).cuda(), torch.from_numpy(y).cuda(),
                              lam=torch.from_numpy(lam).cuda(), tau=torch.from_numpy(tau).cuda())
    st = out["status"].cpu().numpy(); val = out["value"].cpu().numpy(); its = out["iterations"].cpu().numpy()
    assert (st == 0).all(), (st, its)
    for p in range(0, P, 4):
        ref = mo.solve_dense(w0[p], y[p], float(lam[p]), 0.0)
        assert ref.status == 0
        assert abs(val[p] - ref.value) <= OBJ_RTOL * max(abs(ref.value), OBJ_FLOOR)
    assert its.mean() < 16


def test_mu0_start_runs_on_generic_kernel():
    """KMPC_PARAM_DUAL_INIT = 0 selects the mu0-based starting point, which only the generic lane kernel instantiation
    contains: defaults (lam > 0, tau > 0, long-only) must then be routed away from the FIX instantiation and still
    reach the oracle's optimum."""
    torch, mpc, mo = _mods()
    from koopman_mpc_portfolio_rebalancing_b200 import _capi
    h = _capi.Handle.get(0)
    rng = np.random.default_rng(5)
    N, H, P = 50, 5, 16
    w0 = np.stack([rng.dirichlet(np.ones(N)) for _ in range(P)])
    y = (3e-4 + rng.standard_normal((P, H, N)) * 0.01).astype(np.float32)
    _capi.check(_capi.lib().kmpc_set_solver_param(h.ptr, 2, 0.0))
    try:
        out = mpc.solve_mpc_batch(torch.from_numpy(w0).cuda(), torch.from_numpy(y).cuda())    # lam 1e-3, tau 0.2
        val = out["value"].cpu().numpy(); st = out["status"].cpu().numpy(); W = out["w"].cpu().numpy()
    finally:
        _capi.check(_capi.lib().kmpc_set_solver_param(h.ptr, 0, 0.0))
    for p in range(P):
        ref = mo.solve_structured(w0[p], y[p], 1e-3, 0.2)
        assert st[p] == 0 and abs(val[p] - ref.value) <= OBJ_RTOL * max(abs(ref.value), OBJ_FLOOR), (p, st[p], val[p], ref.value)
        assert np.abs(W[p] - ref.w).max() < W_ATOL
    with pytest.raises(_capi.KmpcError):
        _capi.check(_capi.lib().kmpc_set_solver_param(h.ptr, 99, 1.0))


def test_hard_instances_no_fallback(golden):
    """40 decisions of the config-2 replay on which an earlier version of the solver ended `optimal_inaccurate` or fell
    back to holding the weights (collected with scripts/find_failures.py: barrier weights spanning > 20 decades next to a
    flat optimum; optimal values from oracle.solve_dense, the generic dense interior-point method).  With the retried
    factorisation every one of them must be solved (no fallback) to the objective bar; the optimum is flat there, so
    weights are NOT compared (the two oracle methods themselves differ by up to 0.05 on them)."""
    torch, mpc, mo = _mods()
    g = golden("hard_instances.npz")
    out = mpc.solve_mpc_batch(torch.from_numpy(g["w"]).cuda(), torch.from_numpy(g["y"]).cuda())
    st = out["status"].cpu().numpy(); val = out["value"].cpu().numpy(); W = out["w"].cpu().numpy()
    assert (st <= 1).all(), st
    rel = np.abs(val - g["value"]) / np.maximum(np.abs(g["value"]), OBJ_FLOOR)
    assert rel.max() < OBJ_RTOL, (rel.max(), st[np.argmax(rel)])
    assert np.allclose(W.sum(axis=2), 1.0, atol=1e-8) and W.min() > -1e-10
    assert np.abs(W[:, 0] - g["w"]).sum(axis=1).max() <= 0.2 + 1e-9            # first trade inside the cap
    assert (st == 0).sum() >= 30                                               # most of them now reach the tolerances
    for p in range(len(st)):                                                   # and each plan is certified optimal
        assert _certified_gap(W[p], g["w"][p], g["y"][p], 1e-3, 0.2) < OBJ_RTOL, p


def test_second_attempt_solves_the_stalling_decisions(golden):
    """tests/golden/stall_instances.npz: the 52 decisions of a config-2 step that the aggressive first attempt of the
    solver leaves `optimal_inaccurate` (a few of them 2-5e-6 off the optimal objective).  With the second attempt
    (default) every one must end `optimal`, within the objective bar of the dense oracle and certified by the LP
    certificate; with it switched off the first attempt's behaviour is still there (the fixture exercises the path)."""
    torch, mpc, mo = _mods()
    from koopman_mpc_portfolio_rebalancing_b200 import _capi
    g = golden("stall_instances.npz")
    w0 = torch.from_numpy(g["w"]).cuda(); y = torch.from_numpy(g["y"]).cuda()
    out = mpc.solve_mpc_batch(w0, y)
    st = out["status"].cpu().numpy(); val = out["value"].cpu().numpy(); W = out["w"].cpu().numpy()
    its = out["iterations"].cpu().numpy()
    assert (st == 0).all(), (st, its)
    rel = np.abs(val - g["value"]) / np.maximum(np.abs(g["value"]), OBJ_FLOOR)
    assert rel.max() < OBJ_RTOL, rel.max()
    for p in range(len(st)):
        assert _certified_gap(W[p], g["w"][p], g["y"][p], 1e-3, 0.2) < OBJ_RTOL, p
    h = _capi.Handle.get(0)
    _capi.check(_capi.lib().kmpc_set_solver_param(h.ptr, 5, 0.0))
    try:
        st1 = mpc.solve_mpc_batch(w0, y)["status"].cpu().numpy()
    finally:
        _capi.check(_capi.lib().kmpc_set_solver_param(h.ptr, 0, 0.0))
    print(f"first attempt only: {(st1 != 0).sum()} of {len(st1)} not optimal; with the second attempt: 0; "
          f"iterations {its.mean():.1f} on average, worst objective gap {rel.max():.1e}")
    assert (st1 != 0).sum() >= 10 and (st1 <= 1).all()


def test_inaccurate_decisions_of_a_full_config2_step():
    """Every decision of a full config-2 step (4096 backtests x 246 decisions, replayed day by day through the batch
    solver) that ends `optimal_inaccurate`: the plan must meet the 1e-6 objective bar against the independent dense
    oracle and the certificate (its loose acceptance bar is the builder's choice, mpc_common.cuh kLoose*), nothing may
    fall back, and pulling the first trade back onto the turnover cap (clip_first_trade; the reference never does
    that) must not have moved any weight by more than 2e-5."""
    import bench
    torch, mpc, mo = _mods()
    from koopman_mpc_portfolio_rebalancing_b200 import _capi, engine, model as km, synthetic, backtest as bt
    w = bench.WORKLOADS["cfg2"]
    B, N, d, H, Z, rows = w["B"], w["N"], w["d"], w["H"], w["Z"], w["rows"]
    m = km.make_model(km.model_config("GenericKM", Z, w["enc"], enc_bias=True), N * d)
    m.load_state_dict(synthetic.generic_km_weights(0, N * d, w["enc"], Z))
    eng = engine.BatchedBacktester(m, N, d, bt.MPCConfig(horizon=H), bt.BacktestConfig(horizon=H))
    lr, mean, std, T = bench.make_inputs(w, B, 10_000)
    out = eng.run_device(torch.from_numpy(lr).cuda(), torch.from_numpy(mean).cuda(), torch.from_numpy(std).cuda(), 0, rows)
    yhat, realized = out["yhat"], out["realized"]
    ns = yhat.shape[1]
    wc = torch.full((B, N), 1.0 / N, dtype=torch.float64, device="cuda")
    bad_w, bad_y, bad_plan, bad_val = [], [], [], []
    n_fail = 0
    for t in range(ns):
        r = mpc.solve_mpc_batch(wc, yhat[:, t].contiguous())
        st = r["status"]
        n_fail += int((st >= 2).sum())
        for i in torch.nonzero(st == 1).flatten().tolist():
            bad_w.append(wc[i].cpu().numpy()); bad_y.append(yhat[i, t].cpu().numpy())
            bad_plan.append(r["w"][i].cpu().numpy()); bad_val.append(float(r["value"][i]))
        wn = r["w"][:, 0, :]
        rr = torch.exp(realized[:, t + 1].double()).float() - 1.0
        pr = (wn * rr.double()).sum(dim=1, keepdim=True)
        wc = wn * (1.0 + rr).double() / (1.0 + pr)
    print(f"{B * ns} decisions: {len(bad_w)} optimal_inaccurate, {n_fail} fallbacks")
    assert n_fail == 0
    assert len(bad_w) <= 20                                   # ~50 per million before the second attempt, ~0 with it
    if not bad_w:
        return
    bw, by, bp = np.array(bad_w), np.array(bad_y), np.array(bad_plan)
    for p in range(len(bw)):                                  # certificate: every one of them
        assert _certified_gap(bp[p], bw[p], by[p], 1e-3, 0.2) < OBJ_RTOL, p
        assert np.abs(bp[p][0] - bw[p]).sum() <= 0.2 + 1e-9
    for p in range(min(len(bw), 48)):                         # dense oracle: a sample of >= 40 when there are that many
        dense = mo.solve_dense(bw[p], by[p], 1e-3, 0.2)
        assert dense.status <= 1
        assert abs(bad_val[p] - dense.value) <= OBJ_RTOL * max(abs(dense.value), OBJ_FLOOR), (p, bad_val[p], dense.value)
    h = _capi.Handle.get(0)
    _capi.check(_capi.lib().kmpc_set_solver_param(h.ptr, 4, 0.0))            # same instances without the clip
    try:
        raw = mpc.solve_mpc_batch(torch.from_numpy(bw).cuda(), torch.from_numpy(by).cuda())["w"].cpu().numpy()
    finally:
        _capi.check(_capi.lib().kmpc_set_solver_param(h.ptr, 0, 0.0))
    moved = np.abs(raw - bp).max()
    excess = (np.abs(raw[:, 0] - bw).sum(axis=1) - 0.2).max()
    print(f"clip_first_trade: largest weight change {moved:.2e}, largest cap excess before the clip {excess:.2e}")
    assert moved <= 2e-5 and excess <= 2e-5
