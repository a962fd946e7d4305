"""Mean-variance MPC (mpc.py:119-184) and MarkowitzStrategy (baselines.py:24-106) on the device (csrc/mpc_mv.cu) vs
the fp64 oracle oracle/mpc_oracle.py::solve_mv_dense and the golden run of the unmodified reference strategy."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
OBJ_RTOL, OBJ_FLOOR, W_ATOL = 1e-6, 1e-3, 1e-4


def _instance(rng, N, H):
    X = rng.standard_normal((60, N)) * rng.uniform(0.005, 0.02, N)
    S = np.cov(X, rowvar=False) + 1e-6 * np.eye(N)
    mu = (3e-4 + rng.standard_normal((H, N)) * 2e-3)
    w0 = rng.dirichlet(np.ones(N) * rng.choice([0.3, 1.0]))
    return mu, S, w0


@pytest.mark.parametrize("N,H", [(6, 1), (10, 3), (50, 1), (50, 3), (100, 1), (32, 5)])
def test_random_instances_vs_oracle(N, H):
    import torch
    from koopman_mpc_portfolio_rebalancing_b200 import mpc
    from oracle import mpc_oracle as mo
    rng = np.random.default_rng(17 * N + H)
    P = 8
    insts = [_instance(rng, N, H) for _ in range(P)]
    for (gamma, lam, short) in [(2.0, 1e-3, False), (0.5, 0.0, False), (5.0, 1e-4, True)]:
        mu = torch.from_numpy(np.stack([i[0] for i in insts])).cuda()
        sig = torch.from_numpy(np.stack([i[1] for i in insts])).cuda()
        wc = torch.from_numpy(np.stack([i[2] for i in insts])).cuda()
        out = mpc.solve_mean_variance_batch(wc, mu, sig, gamma, cost_coeff=lam, allow_short=short)
        W = out["w"].cpu().numpy(); val = out["value"].cpu().numpy(); st = out["status"].cpu().numpy()
        kkt = out["kkt"].cpu().numpy()
        for p in range(P):
            ref = mo.solve_mv_dense(insts[p][2], insts[p][0], insts[p][1], gamma, lam, short)
            assert ref.status == 0
            assert st[p] in (0, 1), (p, st[p], kkt[p])
            assert abs(val[p] - ref.value) <= OBJ_RTOL * max(abs(ref.value), OBJ_FLOOR), (p, val[p], ref.value, kkt[p])
            if st[p] == 0:
                # shorting allowed + a near-singular 60-sample covariance gives leveraged weights of O(50): relative bar
                assert np.abs(W[p][0] - ref.w[0]).max() < W_ATOL * max(1.0, np.abs(ref.w[0]).max())
            assert np.allclose(W[p].sum(axis=1), 1.0, atol=1e-8)
            if not short:
                assert W[p].min() > -1e-10


def test_drop_in_signature_and_unsupported_shape():
    from koopman_mpc_portfolio_rebalancing_b200 import mpc, _capi
    from oracle import mpc_oracle as mo
    rng = np.random.default_rng(3)
    mu, S, w0 = _instance(rng, 8, 1)
    cfg = mpc.MPCConfig(horizon=1, gamma=1.0, cost_coeff=1e-3)
    w, info = mpc.solve_mpc_mean_variance(w0, mu.astype(np.float32), S, cfg)       # float32 mu, as the strategy passes it
    ref = mo.solve_mv_dense(w0, mu.astype(np.float32), S, 1.0, 1e-3)
    assert info["status"] == "optimal" and w.shape == (1, 8)
    assert abs(info["value"] - ref.value) <= OBJ_RTOL * max(abs(ref.value), OBJ_FLOOR)
    with pytest.raises(_capi.KmpcError):
        mpc.solve_mpc_mean_variance(np.ones(700) / 700, np.zeros((2, 700)), np.eye(700), cfg)   # H*N > 1280: fails loudly


@pytest.mark.parametrize("N,H", [(200, 1), (100, 3), (320, 2), (500, 1)])
def test_large_shapes_block_kernel_vs_oracle(N, H):
    """H*N > 160 (the reference has no cap, mpc.py:119-184): one block of 256 threads per problem, the dense Newton matrix
    in a global workspace.  Same bars as the warp kernel; long-only and shorting, with and without a cost."""
    import torch
    from koopman_mpc_portfolio_rebalancing_b200 import _capi, mpc
    from oracle import mpc_oracle as mo
    assert _capi.lib().kmpc_mv_supported(H, N) == 1
    rng = np.random.default_rng(7 * N + H)
    P = 3 if N >= 320 else 4
    insts = [_instance(rng, N, H) for _ in range(P)]
    for (gamma, lam, short) in [(2.0, 1e-3, False), (1.0, 0.0, True)]:
        mu = torch.from_numpy(np.stack([i[0] for i in insts])).cuda()
        sig = torch.from_numpy(np.stack([i[1] for i in insts])).cuda()
        wc = torch.from_numpy(np.stack([i[2] for i in insts])).cuda()
        out = mpc.solve_mean_variance_batch(wc, mu, sig, gamma, cost_coeff=lam, allow_short=short)
        W = out["w"].cpu().numpy(); val = out["value"].cpu().numpy(); st = out["status"].cpu().numpy()
        kkt = out["kkt"].cpu().numpy()
        for p in range(P if N < 500 else 1):
            ref = mo.solve_mv_dense(insts[p][2], insts[p][0], insts[p][1], gamma, lam, short)
            assert ref.status == 0
            assert st[p] in (0, 1), (p, st[p], kkt[p])
            assert abs(val[p] - ref.value) <= OBJ_RTOL * max(abs(ref.value), OBJ_FLOOR), (p, val[p], ref.value, kkt[p])
            if st[p] == 0:
                assert np.abs(W[p][0] - ref.w[0]).max() < W_ATOL * max(1.0, np.abs(ref.w[0]).max())
            assert np.allclose(W[p].sum(axis=1), 1.0, atol=1e-8)
            if not short:
                assert W[p].min() > -1e-10
        assert (st <= 1).all()


def test_batched_markowitz_backtest_vs_strategy_loop():
    """engine.run_markowitz_batched: B Markowitz backtests advanced together (rolling mean / covariance on the device,
    one batched mean-variance solve per step, portfolio step on the device) must reproduce, path by path, the
    MarkowitzStrategy + run_backtest loop (itself pinned to the golden run of the unmodified reference strategy)."""
    import torch
    from koopman_mpc_portfolio_rebalancing_b200 import backtest as bt, baselines, data_finance as df, engine, synthetic
    B, T, N, d = 3, 150, 12, 4
    lrs = [synthetic.gbm_log_returns(40 + b, T, N) for b in range(B)]
    envs = [df.create_finance_env_from_returns(lr, embedding_dim=d, n_train_days=60, n_val_days=20) for lr in lrs]
    cfg = bt.BacktestConfig(initial_capital=1e4, horizon=1, cost_coeff=1e-3)
    realized = torch.stack([e.realized_test_returns_device() for e in envs])          # [B, rows, N] float32
    out = engine.run_markowitz_batched(realized, risk_aversion=1.5, cost_coeff=1e-3, bt_config=cfg, want_history=True)
    hist = out["history"].cpu().numpy(); met = out["metrics"].cpu().numpy()
    for b in range(B):
        strat = baselines.MarkowitzStrategy(risk_aversion=1.5, cost_coeff=1e-3)
        ref = bt.run_backtest(strat, envs[b], cfg, verbose=False)
        assert hist[b].shape[0] == len(ref)
        assert np.allclose(hist[b][:, 0], ref["portfolio_value"].values, rtol=1e-6)
        assert np.allclose(hist[b][:, 2], ref["turnover"].values, atol=1e-5)
        m = bt.calculate_metrics(ref)
        assert np.allclose(met[b], [m[k] for k in bt.METRIC_KEYS], rtol=1e-4, atol=1e-5)


def test_markowitz_strategy_vs_reference_golden(golden):
    """MarkowitzStrategy through run_backtest (host loop, device solve per step) vs the golden run of the UNMODIFIED
    reference strategy: same (mu, Sigma) estimates at every step, same weights, history and metrics."""
    from koopman_mpc_portfolio_rebalancing_b200 import backtest as bt, baselines, data_finance as df, mpc, synthetic
    g = golden("markowitz_small.npz")
    T, N, d = int(g["T"]), int(g["N"]), int(g["d"])
    lr = synthetic.gbm_log_returns(int(g["log_returns_seed"]), T, N)
    env = df.create_finance_env_from_returns(lr, embedding_dim=d, n_train_days=int(g["n_train_days"]),
                                             n_val_days=int(g["n_val_days"]))
    # stage-wise: the reference's own (w_cur, mu, Sigma) of a few steps -> same optimum
    for k in (0, 10, 30, 53):
        w, info = mpc.solve_mpc_mean_variance(g["w_cur"][k], g["mu"][k], g["sigma"][k],
                                              mpc.MPCConfig(horizon=1, gamma=float(g["gamma"]), cost_coeff=1e-3))
        assert info["status"] in ("optimal", "optimal_inaccurate")
        assert abs(info["value"] - g["value"][k]) <= OBJ_RTOL * max(abs(g["value"][k]), OBJ_FLOOR)
        assert np.abs(w - g["w_opt"][k]).max() < W_ATOL
    strat = baselines.MarkowitzStrategy(risk_aversion=float(g["gamma"]), cost_coeff=1e-3)
    hist = bt.run_backtest(strat, env, bt.BacktestConfig(initial_capital=1e4, horizon=1, cost_coeff=1e-3), verbose=False)
    assert len(hist) == len(g["history"])
    assert np.allclose(hist["portfolio_value"].values, g["history"][:, 0], rtol=1e-4)
    met = bt.calculate_metrics(hist)
    assert np.allclose([met[k] for k in bt.METRIC_KEYS], g["metrics"], rtol=2e-3, atol=2e-4)


def test_markowitz_n50_vs_reference_golden(golden):
    """50 assets, H = 1 (tests/golden/markowitz_n50.npz: the UNMODIFIED reference MarkowitzStrategy + run_backtest, the
    mean-variance solve substituted by the dense fp64 oracle): stage-wise the reference's own (w_cur, mu, Sigma) give the
    same optimum; the strategy loop and the BATCHED backtest (engine.run_markowitz_batched) reproduce its history and
    metrics."""
    import torch
    from koopman_mpc_portfolio_rebalancing_b200 import backtest as bt, baselines, data_finance as df, engine, mpc, synthetic
    g = golden("markowitz_n50.npz")
    T, N, d = int(g["T"]), int(g["N"]), int(g["d"])
    assert N == 50
    for k in range(len(g["value"])):
        w, info = mpc.solve_mpc_mean_variance(g["w_cur"][k], g["mu"][k], g["sigma"][k],
                                              mpc.MPCConfig(horizon=1, gamma=float(g["gamma"]), cost_coeff=1e-3))
        assert info["status"] in ("optimal", "optimal_inaccurate")
        assert abs(info["value"] - g["value"][k]) <= OBJ_RTOL * max(abs(g["value"][k]), OBJ_FLOOR)
        assert np.abs(w - g["w_opt"][k]).max() < W_ATOL
    lr = synthetic.gbm_log_returns(int(g["log_returns_seed"]), T, N)
    env = df.create_finance_env_from_returns(lr, embedding_dim=d, n_train_days=int(g["n_train_days"]),
                                             n_val_days=int(g["n_val_days"]))
    cfg = bt.BacktestConfig(initial_capital=1e4, horizon=1, cost_coeff=1e-3)
    hist = bt.run_backtest(baselines.MarkowitzStrategy(risk_aversion=float(g["gamma"]), cost_coeff=1e-3), env, cfg, verbose=False)
    assert len(hist) == len(g["history"])
    assert np.allclose(hist["portfolio_value"].values, g["history"][:, 0], rtol=1e-4)
    out = engine.run_markowitz_batched(env.realized_test_returns_device().unsqueeze(0), risk_aversion=float(g["gamma"]),
                                       cost_coeff=1e-3, bt_config=cfg, want_history=True)
    hb = out["history"][0].cpu().numpy()
    assert hb.shape == g["history"].shape
    assert np.allclose(hb[:, 0], g["history"][:, 0], rtol=1e-4)
    assert np.allclose(out["metrics"][0].cpu().numpy(), g["metrics"], rtol=2e-3, atol=2e-4)
    assert int(out["status_counts"][2]) == 0
