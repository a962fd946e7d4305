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
        mpc.solve_mpc_mean_variance(np.ones(100) / 100, np.zeros((5, 100)), np.eye(100), cfg)   # H*N > 160: fails loudly


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
