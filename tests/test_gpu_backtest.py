"""Persistent backtest kernel (MPC + portfolio step + metrics fused) vs the oracle loop and the golden history of
the unmodified reference run_backtest (tests/golden/backtest_cfg1.npz)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _mods():
    import torch
    from koopman_mpc_portfolio_rebalancing_b200 import backtest as bt
    from oracle import backtest_oracle as bo, data_oracle as do
    return torch, bt, bo, do


def test_cfg1_history_vs_reference_golden(golden):
    """Same forecasts as the reference run -> same decisions; history within the fp32-exp ulp noise documented
    in oracle/backtest_oracle.py (return <= 1.2e-7 abs per day, value 2e-6 rel)."""
    torch, bt, bo, do = _mods()
    from koopman_mpc_portfolio_rebalancing_b200 import synthetic
    g = golden("backtest_cfg1.npz")
    T, d, N, H = int(g["T"]), 20, 10, 5
    lr = synthetic.gbm_log_returns(int(g["log_returns_seed"]), T, N)
    z = do.standardize(lr, g["mean"], g["std"])
    emb = do.time_delay_embedding(z, d)
    _, _, (c0, c1) = do.split_rows(T, int(g["n_train_days"]), int(g["n_val_days"]), d)
    all_ret = do.destandardize(do.extract_current_returns(emb[c0:c1], N), g["mean"], g["std"])
    yhat = torch.from_numpy(g["yhat"]).cuda().unsqueeze(0)
    realized = torch.from_numpy(all_ret).cuda().unsqueeze(0)
    out = bt.run_backtest_batched(yhat, realized, n_steps=246, horizon=H, want_history=True)
    hist = out["history"][0].cpu().numpy(); met = out["metrics"][0].cpu().numpy(); stats = out["stats"][0].cpu().numpy()
    assert stats[0] == 246 and stats[1] == 0 and stats[2] == 0, stats
    ref_hist, _ = bo.run_backtest(bo.koopman_mpc_decider(g["yhat"], 1e-3, 0.2), all_ret, int(g["test_len"]), H)
    # vs the oracle loop with the same (platform-independent) exp convention: tight
    assert np.allclose(hist[:, 1], ref_hist[:, 1], atol=2e-9), np.abs(hist[:, 1] - ref_hist[:, 1]).max()
    assert np.allclose(hist[:, 0], ref_hist[:, 0], rtol=1e-7)
    assert np.allclose(hist[:, 2], ref_hist[:, 2], atol=2e-6)
    m = bo.calculate_metrics(ref_hist)
    assert np.allclose(met, [m[k] for k in bo.METRIC_KEYS], rtol=1e-6, atol=1e-7)
    # vs the unmodified reference run: numpy's fp32 exp is not correctly rounded (<= 1 ulp of ~1.0 per day)
    assert np.abs(hist[:, 1] - g["history"][:, 1]).max() < 2.5e-7
    assert np.allclose(hist[:, 0], g["history"][:, 0], rtol=5e-6)
    assert np.allclose(met, g["metrics"], rtol=2e-4, atol=2e-5)
    # device metrics == calculate_metrics on the device history
    import pandas as pd
    df = pd.DataFrame(hist, columns=list(bt.HISTORY_COLS))
    m2 = bt.calculate_metrics(df)
    assert np.allclose(met, [m2[k] for k in bt.METRIC_KEYS], rtol=1e-10, atol=1e-12)


def test_batched_sweep_matches_individual_runs():
    """lambda/tau sweep over shared forecasts + several price paths (config-4 shape, tiny): every backtest equals its
    own oracle run; sharing indices must not mix paths."""
    torch, bt, bo, do = _mods()
    rng = np.random.default_rng(3)
    N, H, rows = 12, 3, 40
    n_steps = rows - 1 - H
    S, Q = 2, 3
    yhat = (3e-4 + 0.01 * rng.standard_normal((S, n_steps, H, N))).astype(np.float32)
    realized = (0.012 * rng.standard_normal((Q, rows, N))).astype(np.float32)
    lam = np.array([1e-3, 1e-2, 0.0, 1e-4, 1e-3, 5e-3]); tau = np.array([0.2, 0.05, 0.5, 0.0, 1.0, 0.2])
    yi = np.array([0, 1, 0, 1, 0, 1], np.int32); ri = np.array([0, 1, 2, 0, 1, 2], np.int32)
    cc = np.array([1e-3, 2e-3, 0.0, 1e-3, 1e-3, 5e-4]); cap = np.array([1e4, 5e3, 1.0, 1e4, 2e4, 1e4])
    out = bt.run_backtest_batched(torch.from_numpy(yhat).cuda(), torch.from_numpy(realized).cuda(), n_steps=n_steps,
                                  horizon=H, lam=lam, tau=tau, cost_coeff=cc, capital=cap, yhat_index=yi,
                                  realized_index=ri, want_history=True)
    hist = out["history"].cpu().numpy(); met = out["metrics"].cpu().numpy()
    for b in range(6):
        rh, _ = bo.run_backtest(bo.koopman_mpc_decider(yhat[yi[b]], lam[b], tau[b]), realized[ri[b]], rows - 1, H,
                                initial_capital=cap[b], cost_coeff=cc[b])
        assert np.allclose(hist[b][:, 0], rh[:, 0], rtol=1e-6), b
        assert np.allclose(hist[b][:, 1], rh[:, 1], atol=1e-7), b
        m = bo.calculate_metrics(rh)
        assert np.allclose(met[b], [m[k] for k in bo.METRIC_KEYS], rtol=1e-5, atol=1e-6), b


def test_rebalance_freq_and_short_horizon():
    torch, bt, bo, do = _mods()
    rng = np.random.default_rng(4)
    N, H, rows = 5, 2, 30
    n_steps = rows - 1 - H
    yhat = (0.01 * rng.standard_normal((1, n_steps, H, N))).astype(np.float32)
    realized = (0.012 * rng.standard_normal((1, rows, N))).astype(np.float32)
    out = bt.run_backtest_batched(torch.from_numpy(yhat).cuda(), torch.from_numpy(realized).cuda(), n_steps=n_steps,
                                  horizon=H, rebalance_freq=3, want_history=True)
    rh, ts = bo.run_backtest(bo.koopman_mpc_decider(yhat[0], 1e-3, 0.2), realized[0], rows - 1, H, rebalance_freq=3)
    hist = out["history"][0].cpu().numpy()
    assert hist.shape == rh.shape == (len(range(0, n_steps, 3)), 4)
    assert np.allclose(hist[:, 0], rh[:, 0], rtol=1e-6)


def test_config3_shape_backtest_vs_oracle_loop():
    """500 assets, H = 10 (BASELINE config 3 shape): persistent backtest kernel vs the oracle loop on a few steps."""
    torch, bt, bo, do = _mods()
    rng = np.random.default_rng(33)
    B, N, H, rows = 2, 500, 10, 17
    ns = rows - 1 - H
    yhat = (3e-4 + rng.standard_normal((B, ns, H, N)) * 0.006).astype(np.float32)
    realized = (3e-4 + rng.standard_normal((B, rows, N)) * 0.012).astype(np.float32)
    out = bt.run_backtest_batched(torch.from_numpy(yhat).cuda(), torch.from_numpy(realized).cuda(), n_steps=ns, horizon=H,
                                  lam0=1e-3, tau0=0.2, cost_coeff0=1e-3, capital0=1e4, want_history=True)
    hist = out["history"].cpu().numpy()
    for b in range(B):
        ref, _ = bo.run_backtest(bo.koopman_mpc_decider(yhat[b], 1e-3, 0.2), realized[b], rows - 1, H)
        ref = np.asarray(ref)
        assert hist[b].shape == ref.shape
        assert np.allclose(hist[b][:, 0], ref[:, 0], rtol=1e-7)           # portfolio value
        assert np.allclose(hist[b][:, 1:], ref[:, 1:], atol=2e-7)         # return, turnover, cost


@pytest.mark.parametrize("N,H,mixed", [(10, 5, False), (50, 5, False), (50, 5, True), (100, 3, True), (40, 10, False)])
def test_persistent_kernel_is_schedule_independent(N, H, mixed):
    """Race hunting without a sanitizer (compute-sanitizer is closed on this GPU pool): the persistent kernel runs
    several backtests per block in lockstep with named barriers, a barrier-as-exit-vote every 4th trip, shared-memory
    tiles reused between phases and a DMMA contraction whose partial products cross warps.  A race or a missed barrier
    shows as a dependence on WHO shares the block: every backtest's full history must be bit-identical whether it runs
    alone, with 2, 5, 9 or 37 neighbours (more backtests than slots of a block, so slots re-fetch work), in any position
    of the batch, and from run to run.  G = 1, 2, 4 kernels and the H = 10 (thread-private factors) variant; `mixed`:
    per-backtest lambda / tau with zeros (the generic instantiation; backtest 0 carries a zero so that every subset
    takes that instantiation), else the reference defaults (the FIX instantiation)."""
    torch, bt, bo, do = _mods()
    rng = np.random.default_rng(500 + N)
    Bmax, rows = 37, 12 + H
    ns = rows - 1 - H
    yhat = (3e-4 + rng.standard_normal((Bmax, ns, H, N)) * 0.01).astype(np.float32)
    realized = (3e-4 + rng.standard_normal((Bmax, rows, N)) * 0.012).astype(np.float32)
    lam = np.where(rng.random(Bmax) < 0.7, 1e-3, rng.choice([0.0, 1e-2], Bmax)) if mixed else np.full(Bmax, 1e-3)
    tau = np.where(rng.random(Bmax) < 0.7, 0.2, rng.choice([0.0, 0.05], Bmax)) if mixed else np.full(Bmax, 0.2)
    if mixed:
        lam[0] = 0.0
    yd, rd = torch.from_numpy(yhat).cuda(), torch.from_numpy(realized).cuda()

    def run(idx):
        idx = np.asarray(idx)
        out = bt.run_backtest_batched(yd, rd, n_steps=ns, horizon=H, lam=lam[idx], tau=tau[idx],
                                      yhat_index=idx.astype(np.int32), realized_index=idx.astype(np.int32), B=len(idx),
                                      want_history=True)
        return out["history"].cpu().numpy(), out["metrics"].cpu().numpy()

    full_h, full_m = run(np.arange(Bmax))
    assert np.isfinite(full_h).all()
    for rep in range(3):                                       # run to run
        h2, m2 = run(np.arange(Bmax))
        assert np.array_equal(h2, full_h) and np.array_equal(m2, full_m)
    for k in (1, 2, 5, 9):                                     # alone / few neighbours
        h, m = run(np.arange(k))
        assert np.array_equal(h, full_h[:k]) and np.array_equal(m, full_m[:k]), k
    perm = np.concatenate([[0], 1 + rng.permutation(Bmax - 1)])   # any position in the batch (0 stays: see `mixed`)
    perm[[0, 17]] = perm[[17, 0]]
    h, m = run(perm)
    assert np.array_equal(h, full_h[perm]) and np.array_equal(m, full_m[perm])


@pytest.mark.parametrize("N,H,B", [(50, 5, 96), (100, 5, 96), (40, 3, 96), (140, 10, 96), (300, 5, 96), (160, 5, 500)])
def test_active_set_pipeline_matches_full_solver(N, H, B):
    """KMPC_PARAM_ACTIVE_SET (default on): once a backtest's portfolio has concentrated, the persistent kernel hands it to
    backtest_active_kernel, which solves every decision on the held assets + the best forecasts of each stage (one warp per
    problem) and then checks the optimality conditions of all excluded assets against the duals of the reduced solution
    (mpc_lane_kernels.cuh).  Against the kernel that solves all N assets at every decision, on forecasts with persistent
    per-asset drifts (so that portfolios do concentrate), mixed per-backtest costs and caps: every decision optimal in
    both, histories equal at the end-to-end bar, fewer Newton steps.  Mode 2 starts every set from the held assets alone,
    so that the assets of the plan have to come in through the check-and-repair path: same histories, more solves.
    (140, 10), (300, 5): the config-3 route — full-width kernel of 16 warps for the dense start, then the WIDE reduced-solve
    kernel (eight warps per problem, up to 256 active assets, hand-over at 228 / 238 held assets), forecasts read from global
    memory instead of the shared-memory stage.  (160, 5, 500): more backtests than the wide kernel has slots (148 on a
    B200), so that backtests travel between slots through the ready queue and their saved state."""
    import torch
    from koopman_mpc_portfolio_rebalancing_b200 import _capi, backtest as bt
    rows = 70 if B <= 96 else 50
    ns = rows - 1 - H
    g = torch.Generator(device="cuda").manual_seed(100 * N + H)
    drift = 2e-3 * torch.randn((B, 1, 1, N), device="cuda", generator=g)
    regime = torch.where(torch.arange(ns, device="cuda").view(1, ns, 1, 1) >= ns // 2, -1.0, 1.0)   # drifts flip half-way
    yhat = (3e-4 + drift * regime + 3e-4 * torch.randn((B, ns, H, N), device="cuda", generator=g)).float()
    realized = (3e-4 + 1.2e-2 * torch.randn((B, rows, N), device="cuda", generator=g)).float()
    rng = np.random.default_rng(N)
    lam = rng.choice([1e-3, 1e-4, 3e-3], B); tau = rng.choice([0.2, 0.1, 0.5], B)
    h = _capi.Handle.get(0)
    res = {}
    try:
        for mode in (0, 1, 2):
            _capi.check(_capi.lib().kmpc_set_solver_param(h.ptr, 7, float(mode)))
            out = bt.run_backtest_batched(yhat, realized, n_steps=ns, horizon=H, lam=lam, tau=tau, want_history=True)
            torch.cuda.synchronize()
            res[mode] = (out["history"].cpu().numpy(), out["stats"].cpu().numpy(), out["metrics"].cpu().numpy())
    finally:
        _capi.check(_capi.lib().kmpc_set_solver_param(h.ptr, 7, 1.0))
    for mode in (0, 1, 2):
        st = res[mode][1].sum(axis=0)
        assert st[0] == B * ns and st[1] == 0 and st[2] == 0, (mode, st)            # every decision optimal
    v0 = res[0][0][..., 0]
    for mode in (1, 2):
        dv = np.abs(res[mode][0][..., 0] / v0 - 1).max(axis=1)                        # per backtest
        dturn = np.abs(res[mode][0][..., 2] - res[0][0][..., 2]).max(axis=1)          # turnover per day
        if B <= 96:
            assert dv.max() < 1e-4, (mode, dv.max())
            assert dturn.max() < 1e-3
        else:
            # among hundreds of backtests a decision with a FLAT optimum turns up (backtest 463 here: on day 5 the full and
            # the reduced plan agree to 3e-12 in the objective, certified gaps 1e-12, and differ by 2.7e-3 in the first trade;
            # replayed on the CPU oracle, profiles/README.md): both are optima, the trajectories part by ~1e-4
            ok = dv < 1e-4
            assert ok.mean() >= 0.99 and dv.max() < 2e-3, (mode, dv.max(), int((~ok).sum()))
            assert dturn[ok].max() < 1e-3
    it0, it1, it2 = (int(res[m][1][:, 3].sum()) for m in (0, 1, 2))
    assert it1 < it0                                    # reduced problems take fewer Newton steps
    assert it2 > it1                                    # mode 2 had to re-solve: the repair path ran
