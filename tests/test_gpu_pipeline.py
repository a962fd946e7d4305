"""End-to-end through the reference-shaped API: KoopmanMPCStrategy + run_backtest + calculate_metrics on config 1
vs the golden run of the unmodified reference; the batched engine vs per-path oracle runs."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def cfg1(golden):
    from koopman_mpc_portfolio_rebalancing_b200 import data_finance as df, model as km, synthetic
    g = golden("backtest_cfg1.npz")
    lr = synthetic.gbm_log_returns(int(g["log_returns_seed"]), int(g["T"]), 10)
    env = df.create_finance_env_from_returns(lr, embedding_dim=20, n_train_days=int(g["n_train_days"]),
                                             n_val_days=int(g["n_val_days"]))
    m = km.make_model(km.model_config("GenericKM", 128, [1024, 1024], enc_bias=True), 200)
    m.load_state_dict(synthetic.generic_km_weights(0, 200, [1024, 1024], 128))
    return g, env, m


def test_config1_drop_in_run_backtest(golden):
    """README.md:45-68 call pattern.  Forecasts match the reference's to 1e-5; the LP-like MPC can amplify a 1e-6
    forecast difference into a different vertex on near-ties, so history/metrics get the looser bar documented in
    tests/test_oracle_backtest.py::test_full_oracle_path_reproduces_reference_backtest."""
    from koopman_mpc_portfolio_rebalancing_b200 import backtest as bt
    g, env, m = cfg1(golden)
    assert np.array_equal(env.test_dataset.data[:3].cpu().numpy(), g["test_first_rows"])
    strat = bt.KoopmanMPCStrategy(m, bt.MPCConfig(horizon=5, cost_coeff=1e-3, max_turnover=0.2))
    yhat = strat.forecast(env, 0, 246).cpu().numpy()
    rel = np.abs(yhat - g["yhat"]).reshape(246, -1).max(1) / np.abs(g["yhat"]).reshape(246, -1).max(1)
    assert rel.max() < 1e-5, rel.max()
    df_ = bt.run_backtest(strat, env, bt.BacktestConfig(initial_capital=1e4, horizon=5, cost_coeff=1e-3), verbose=False)
    assert list(df_.columns) == ["date", "portfolio_value", "return", "turnover", "cost"] and len(df_) == 246
    met = bt.calculate_metrics(df_)
    assert np.allclose([met[k] for k in bt.METRIC_KEYS], g["metrics"], rtol=2e-4, atol=2e-4)
    assert np.allclose(df_["portfolio_value"].values, g["history"][:, 0], rtol=1e-4)
    assert df_.attrs["solve_stats"][2] == 0          # no fallback decisions
    # single-step drop-in: Strategy.rebalance(t, w, env) (backtest.py:80-131)
    w = strat.rebalance(0, np.ones(10) / 10, env)
    assert np.abs(w - g["w_opt"][0, 0]).max() < 1e-4
    # Buy & Hold through the generic host loop
    bh = bt.run_backtest(bt.BuyAndHoldStrategy(), env, bt.BacktestConfig(initial_capital=1e4, horizon=5, cost_coeff=1e-3), verbose=False)
    assert np.allclose(bh["portfolio_value"].values, g["bh_history"][:, 0], rtol=1e-6)


def test_batched_engine_vs_oracle_paths():
    """scenario batch (config-2 shape, tiny): every path equals its own oracle backtest"""
    import torch
    from koopman_mpc_portfolio_rebalancing_b200 import engine, model as km, synthetic, backtest as bt
    from oracle import backtest_oracle as bo, data_oracle as do, forecast_oracle as fo
    B, N, d, H, rows, Z = 5, 16, 8, 5, 40, 64
    T = rows + d - 1
    lr = synthetic.gbm_log_returns_batch(100, B, T, N)
    mean = lr.mean(axis=1); std = lr.std(axis=1, ddof=1)
    sd = synthetic.generic_km_weights(1, N * d, [96, 96], Z)
    m = km.make_model(km.model_config("GenericKM", Z, [96, 96], enc_bias=True), N * d)
    m.load_state_dict(sd)
    eng = engine.BatchedBacktester(m, N, d, bt.MPCConfig(horizon=H), bt.BacktestConfig(horizon=H))
    res = eng.run(engine.PathBatch(lr, mean, std, 0, rows), want_history=True)
    spec = fo.ModelSpec(kind="generic", act="relu", last_relu=False, norm_fn="id", dec_act="relu")
    ns = rows - 1 - H
    worst = 0.0
    for b in range(B):
        emb = do.time_delay_embedding(do.standardize(lr[b], mean[b], std[b]), d)
        yhat = fo.forecast(emb[:ns], sd, spec, H, N, mean[b], std[b])
        allr = do.destandardize(do.extract_current_returns(emb, N), mean[b], std[b])
        rh, _ = bo.run_backtest(bo.koopman_mpc_decider(yhat, 1e-3, 0.2), allr, rows - 1, H)
        worst = max(worst, np.abs(res["history"][b][:, 0] / rh[:, 0] - 1).max())
        mo_ = bo.calculate_metrics(rh)
        assert np.allclose(res["metrics"][b], [mo_[k] for k in bo.METRIC_KEYS], rtol=5e-4, atol=5e-4), b
    assert worst < 2e-4, worst
    assert res["stats"][:, 2].sum() == 0


def test_config3_shape_lista_pipeline_vs_oracle():
    """BASELINE config 3 at full model size (LISTAKM linear encoder, 500 assets, d = 10 -> obs 5000, Z = 2048, 10
    LISTA loops, H = 10, turnover cap 0.2), a few short backtests: forecast within 1e-5 (norm-wise) of the oracle,
    every path's history equal to the oracle backtest run on the SAME forecasts (stage-wise parity: with 500 assets
    whose forecasts differ by ~1e-5 the LP-like program is near-degenerate, and a 1e-6 forecast difference moves the
    optimal vertex)."""
    import torch
    from koopman_mpc_portfolio_rebalancing_b200 import engine, model as km, synthetic, backtest as bt
    from oracle import backtest_oracle as bo, data_oracle as do, forecast_oracle as fo
    B, N, d, H, rows, Z = 2, 500, 10, 10, 16, 2048
    T = rows + d - 1
    lr = synthetic.gbm_log_returns_batch(300, B, T, N)
    mean = np.full((B, N), 3e-4); std = np.full((B, N), 0.014)
    sd, L = synthetic.lista_km_weights(7, N * d, Z)
    m = km.make_model(km.model_config("LISTAKM", Z, lista_loops=10, lista_L=L, lista_alpha=5e-3, lista_linear=True), N * d)
    m.load_state_dict(sd)
    eng = engine.BatchedBacktester(m, N, d, bt.MPCConfig(horizon=H, cost_coeff=1e-3, max_turnover=0.2),
                                   bt.BacktestConfig(horizon=H, cost_coeff=1e-3))
    out = eng.run_device(torch.from_numpy(lr).cuda(), torch.from_numpy(mean).cuda(), torch.from_numpy(std).cuda(), 0, rows,
                         want_history=True)
    yhat_gpu = out["yhat"].cpu().numpy(); hist = out["history"].cpu().numpy()
    spec = fo.ModelSpec(kind="lista", linear_encoder=True, alpha=5e-3, L=L, loops=10, act="relu", last_relu=False)
    ns = rows - 1 - H
    for b in range(B):
        emb = do.time_delay_embedding(do.standardize(lr[b], mean[b], std[b]), d)
        yhat = fo.forecast(emb[:ns], sd, spec, H, N, mean[b], std[b])
        rel = np.abs(yhat_gpu[b] - yhat).reshape(ns, -1).max(1) / np.abs(yhat).reshape(ns, -1).max(1)
        assert rel.max() < 1e-5, rel.max()
        allr = do.destandardize(do.extract_current_returns(emb, N), mean[b], std[b])
        rh, _ = bo.run_backtest(bo.koopman_mpc_decider(yhat_gpu[b], 1e-3, 0.2), allr, rows - 1, H)
        rh = np.asarray(rh)
        assert np.allclose(hist[b][:, 0], rh[:, 0], rtol=1e-6), np.abs(hist[b][:, 0] / rh[:, 0] - 1).max()
        assert np.allclose(hist[b][:, 1:], rh[:, 1:], atol=1e-6)
    assert out["stats"].cpu().numpy()[:, 2].sum() == 0


def test_dmd_strategy_vs_reference_golden(golden):
    """DMDStrategy (baselines.py:109-187) through the library: K fitted on the host like the reference, forecasts of
    all steps by the folded read-out GEMM, backtest by the persistent MPC kernel — against the golden run of the
    UNMODIFIED reference DMDStrategy + run_backtest (dmd_small.npz)."""
    from koopman_mpc_portfolio_rebalancing_b200 import backtest as bt, baselines, data_finance as df, synthetic
    g = golden("dmd_small.npz")
    T, N, d, H = int(g["T"]), int(g["N"]), int(g["d"]), int(g["H"])
    lr = synthetic.gbm_log_returns(int(g["log_returns_seed"]), T, N)
    env = df.create_finance_env_from_returns(lr, embedding_dim=d, n_train_days=int(g["n_train_days"]),
                                             n_val_days=int(g["n_val_days"]))
    strat = baselines.DMDStrategy(env.train_dataset.data, bt.MPCConfig(horizon=H, cost_coeff=1e-3, max_turnover=0.2))
    assert np.abs(strat.K - g["K"]).max() <= 2e-5 * np.abs(g["K"]).max()
    ns = len(env.test_dataset) - H
    yhat = strat.forecast(env, 0, ns).cpu().numpy()
    rel = np.abs(yhat - g["yhat"]).reshape(ns, -1).max(1) / np.abs(g["yhat"]).reshape(ns, -1).max(1)
    assert rel.max() < 1e-5, rel.max()
    hist = bt.run_backtest(strat, env, bt.BacktestConfig(initial_capital=1e4, horizon=H, cost_coeff=1e-3), verbose=False)
    assert len(hist) == ns
    assert np.allclose(hist["portfolio_value"].values, g["history"][:, 0], rtol=1e-4)
    met = bt.calculate_metrics(hist)
    assert np.allclose([met[k] for k in bt.METRIC_KEYS], g["metrics"], rtol=2e-3, atol=2e-4)
    w = strat.rebalance(0, np.ones(N) / N, env)                         # single-step drop-in (baselines.py:147-187)
    assert np.abs(w - g["w_opt"][0, 0]).max() < 1e-4


def test_sweep_grid_shares_forecasts_config4_shape():
    """BASELINE config 4 in small: 2 weight sets x 3 lambdas x 3 taus on one price path; forecasts computed once per
    weight set and shared through yhat_index; every grid cell equals its own oracle backtest; two shards cover the grid."""
    import torch
    from koopman_mpc_portfolio_rebalancing_b200 import engine, model as km, synthetic, backtest as bt
    from oracle import backtest_oracle as bo, data_oracle as do, forecast_oracle as fo
    N, d, H, rows, Z = 12, 6, 5, 30, 64
    T = rows + d - 1
    lr = synthetic.gbm_log_returns(77, T, N)
    mean = lr.mean(axis=0); std = lr.std(axis=0, ddof=1)
    sds, models = [], []
    for s in range(2):
        sd = synthetic.generic_km_weights(40 + s, N * d, [64, 64], Z)
        m = km.make_model(km.model_config("GenericKM", Z, [64, 64], enc_bias=True), N * d)
        m.load_state_dict(sd); sds.append(sd); models.append(m)
    lam_grid = [1e-5, 1e-3, 1e-1]; tau_grid = [0.01, 0.2, 1.0]
    out = engine.run_grid(models, N, d, lr, mean, std, lam_grid, tau_grid, rows=rows, horizon=H)
    met = out["metrics"].cpu().numpy()
    assert met.shape == (18, 5) and out["stats"].cpu().numpy()[:, 2].sum() == 0
    spec = fo.ModelSpec(kind="generic", act="relu", last_relu=False, norm_fn="id", dec_act="relu")
    emb = do.time_delay_embedding(do.standardize(lr, mean, std), d)
    allr = do.destandardize(do.extract_current_returns(emb, N), mean, std)
    ns = rows - 1 - H
    yhat_gpu = out["yhat"].cpu().numpy()
    for s in range(2):
        want = fo.forecast(emb[:ns], sds[s], spec, H, N, mean, std)
        rel = np.abs(yhat_gpu[s] - want).reshape(ns, -1).max(1) / np.abs(want).reshape(ns, -1).max(1)
        assert rel.max() < 1e-5
        for li in (0, 2):
            for ti in (0, 1, 2):
                rh, _ = bo.run_backtest(bo.koopman_mpc_decider(yhat_gpu[s], lam_grid[li], tau_grid[ti]), allr, rows - 1, H)
                mo_ = bo.calculate_metrics(np.asarray(rh))
                b = (s * 3 + li) * 3 + ti
                assert np.allclose(met[b], [mo_[k] for k in bo.METRIC_KEYS], rtol=1e-6, atol=1e-7), (s, li, ti)
    # two shards of the grid reproduce the full run
    a = engine.run_grid(models, N, d, lr, mean, std, lam_grid, tau_grid, rows=rows, horizon=H, shard=(0, 2))
    b2 = engine.run_grid(models, N, d, lr, mean, std, lam_grid, tau_grid, rows=rows, horizon=H, shard=(1, 2))
    assert a["ids"] == (0, 9) and b2["ids"] == (9, 18)
    assert np.array_equal(np.vstack([a["metrics"].cpu().numpy(), b2["metrics"].cpu().numpy()]), met)


def test_bootstrap_paths_config5_shape():
    """BASELINE config 5 in small: 100 assets, bootstrap paths of one historical block; a rank regenerates exactly its
    shard; a sampled path equals its own oracle backtest."""
    import torch
    from koopman_mpc_portfolio_rebalancing_b200 import engine, model as km, synthetic, backtest as bt
    from oracle import backtest_oracle as bo, data_oracle as do, forecast_oracle as fo
    N, d, H, rows, Z = 100, 5, 5, 24, 128
    T = rows + d - 1
    hist = synthetic.gbm_log_returns(5, 400, N)
    mean = hist.mean(axis=0); std = hist.std(axis=0, ddof=1)
    paths, idx = engine.bootstrap_paths(hist, 6, T, seed=1234)
    p2, idx2 = engine.bootstrap_paths(hist, 3, T, seed=1234, offset=3)
    assert torch.equal(idx[3:], idx2) and torch.equal(paths[3:], p2)
    assert np.array_equal(paths.cpu().numpy(), hist[idx.cpu().numpy()])
    sd = synthetic.generic_km_weights(50, N * d, [128, 128], Z)
    m = km.make_model(km.model_config("GenericKM", Z, [128, 128], enc_bias=True), N * d)
    m.load_state_dict(sd)
    eng = engine.BatchedBacktester(m, N, d, bt.MPCConfig(horizon=H), bt.BacktestConfig(horizon=H))
    out = eng.run_device(paths, torch.from_numpy(mean).cuda(), torch.from_numpy(std).cuda(), 0, rows, want_history=True)
    assert out["stats"].cpu().numpy()[:, 2].sum() == 0
    yhat_gpu = out["yhat"].cpu().numpy(); hist_gpu = out["history"].cpu().numpy()
    spec = fo.ModelSpec(kind="generic", act="relu", last_relu=False, norm_fn="id", dec_act="relu")
    ns = rows - 1 - H
    for b in (0, 5):
        lrb = paths[b].cpu().numpy()
        emb = do.time_delay_embedding(do.standardize(lrb, mean, std), d)
        want = fo.forecast(emb[:ns], sd, spec, H, N, mean, std)
        rel = np.abs(yhat_gpu[b] - want).reshape(ns, -1).max(1) / np.abs(want).reshape(ns, -1).max(1)
        assert rel.max() < 1e-5
        allr = do.destandardize(do.extract_current_returns(emb, N), mean, std)
        rh, _ = bo.run_backtest(bo.koopman_mpc_decider(yhat_gpu[b], 1e-3, 0.2), allr, rows - 1, H)
        assert np.allclose(hist_gpu[b][:, 0], np.asarray(rh)[:, 0], rtol=1e-6)


def test_pipelined_host_copy_equals_device_run():
    """engine.run with host inputs copies the paths in slices on a side stream while earlier slices are standardised
    and forecast; the result must be bit-identical to the one-shot device run."""
    import torch
    from koopman_mpc_portfolio_rebalancing_b200 import engine, model as km, synthetic, backtest as bt
    B, N, d, H, rows, Z = 37, 12, 6, 5, 40, 64
    T = rows + d - 1
    lr = synthetic.gbm_log_returns_batch(900, B, T, N)
    mean = lr.mean(axis=1); std = lr.std(axis=1, ddof=1)
    m = km.make_model(km.model_config("GenericKM", Z, [64, 64], enc_bias=True), N * d)
    m.load_state_dict(synthetic.generic_km_weights(2, N * d, [64, 64], Z))
    eng = engine.BatchedBacktester(m, N, d, bt.MPCConfig(horizon=H), bt.BacktestConfig(horizon=H))
    lr_pinned = torch.from_numpy(lr).pin_memory()
    a = eng.run(engine.PathBatch(lr_pinned, mean, std, 0, rows), want_history=True, copy_chunks=4)
    b = eng.run(engine.PathBatch(lr_pinned, mean, std, 0, rows), want_history=True, copy_chunks=1)
    dev = eng.run_device(torch.from_numpy(lr).cuda(), torch.from_numpy(mean).cuda(), torch.from_numpy(std).cuda(), 0, rows,
                         want_history=True)
    assert np.array_equal(a["metrics"], b["metrics"]) and np.array_equal(a["history"], b["history"])
    assert np.array_equal(a["metrics"], dev["metrics"].cpu().numpy())
    # per-path statistics sliced consistently: shared statistics as well
    a2 = eng.run(engine.PathBatch(lr_pinned, mean[0], std[0], 0, rows), copy_chunks=4)
    b2 = eng.run(engine.PathBatch(lr_pinned, mean[0], std[0], 0, rows), copy_chunks=1)
    assert np.array_equal(a2["metrics"], b2["metrics"])


def test_full_size_config2_properties():
    """BASELINE config 2 at FULL size (4096 scenario backtests x 246 decisions, 50 assets, d = 20, H = 5, the
    1000->1024->1024->1024 encoder) through size-independent properties: bitwise determinism, invariance under
    sharding the batch, the turnover cap and positivity on every one of the 1.0 M decisions, device metrics equal
    to calculate_metrics of the device history, and a sampled path equal to its own oracle backtest."""
    import pandas as pd
    import torch
    import bench
    from koopman_mpc_portfolio_rebalancing_b200 import engine, model as km, synthetic, backtest as bt
    from oracle import backtest_oracle as bo, data_oracle as do
    w = bench.WORKLOADS["cfg2"]
    B, N, d, H, Z, rows = w["B"], w["N"], w["d"], w["H"], w["Z"], w["rows"]
    ns = rows - 1 - H
    lr, mean, std, T = bench.make_inputs(w, B, 10_000)
    m = km.make_model(km.model_config("GenericKM", Z, w["enc"], enc_bias=True), N * d)
    m.load_state_dict(synthetic.generic_km_weights(0, N * d, w["enc"], Z))
    eng = engine.BatchedBacktester(m, N, d, bt.MPCConfig(horizon=H, cost_coeff=1e-3, max_turnover=0.2),
                                   bt.BacktestConfig(initial_capital=1e4, horizon=H, cost_coeff=1e-3))
    lr_d, mean_d, std_d = torch.from_numpy(lr).cuda(), torch.from_numpy(mean).cuda(), torch.from_numpy(std).cuda()
    out = eng.run_device(lr_d, mean_d, std_d, 0, rows, want_history=True)
    met = out["metrics"].cpu().numpy().copy(); hist = out["history"].cpu().numpy().copy()
    stats = out["stats"].cpu().numpy().copy(); yhat = out["yhat"][[7, 3000]].cpu().numpy().copy()
    assert met.shape == (B, 5) and hist.shape == (B, ns, 4) and np.isfinite(met).all() and np.isfinite(hist).all()
    assert stats[:, :3].sum() == B * ns and stats[:, 2].sum() <= 10            # fallbacks: a few per million at most
    # every decision respects the cap (first-stage turnover <= tau) and keeps the value positive
    assert hist[:, :, 2].max() <= 0.2 + 1e-7 and hist[:, :, 0].min() > 0
    # determinism and shard invariance (no cross-backtest state)
    out2 = eng.run_device(lr_d, mean_d, std_d, 0, rows, want_history=True)
    assert np.array_equal(out2["metrics"].cpu().numpy(), met) and np.array_equal(out2["history"].cpu().numpy(), hist)
    sl = slice(1000, 1777)
    out3 = eng.run_device(lr_d[sl].contiguous(), mean_d[sl].contiguous(), std_d[sl].contiguous(), 0, rows, want_history=True)
    assert np.array_equal(out3["metrics"].cpu().numpy(), met[sl])
    # device metrics == calculate_metrics(history) (backtest.py:221-249)
    for b in (0, 1234, 4095):
        df = pd.DataFrame(hist[b], columns=list(bt.HISTORY_COLS))
        m2 = bt.calculate_metrics(df)
        assert np.allclose(met[b], [m2[k] for k in bt.METRIC_KEYS], rtol=1e-10, atol=1e-12)
    # sampled paths vs the oracle loop on the same forecasts
    for j, b in enumerate((7, 3000)):
        emb = do.time_delay_embedding(do.standardize(lr[b], mean[b], std[b]), d)
        allr = do.destandardize(do.extract_current_returns(emb, N), mean[b], std[b])
        rh, _ = bo.run_backtest(bo.koopman_mpc_decider(yhat[j], 1e-3, 0.2), allr, rows - 1, H)
        assert np.allclose(hist[b][:, 0], np.asarray(rh)[:, 0], rtol=1e-6)


def test_compare_strategies_table(golden, tmp_path):
    """The four-strategy comparison of run_experiment.py:85-137 (Buy & Hold / Markowitz / DMD-MPC / Koopman-MPC) and
    its full_comparison_metrics.csv: the DMD row must reproduce the golden run of the reference DMDStrategy."""
    import pandas as pd
    from koopman_mpc_portfolio_rebalancing_b200 import backtest as bt, data_finance as df, engine, model as km, synthetic
    g = golden("dmd_small.npz")
    T, N, d, H = int(g["T"]), int(g["N"]), int(g["d"]), int(g["H"])
    lr = synthetic.gbm_log_returns(int(g["log_returns_seed"]), T, N)
    env = df.create_finance_env_from_returns(lr, embedding_dim=d, n_train_days=int(g["n_train_days"]),
                                             n_val_days=int(g["n_val_days"]))
    m = km.make_model(km.model_config("GenericKM", 16, [32, 32], enc_bias=True), N * d)
    m.load_state_dict(synthetic.generic_km_weights(7, N * d, [32, 32], 16))
    results, table = engine.compare_strategies(
        m, env, bt.BacktestConfig(initial_capital=1e4, horizon=H, cost_coeff=1e-3),
        bt.MPCConfig(horizon=H, cost_coeff=1e-3, max_turnover=0.2), out_dir=str(tmp_path))
    assert list(table.index) == ["Buy & Hold", "Markowitz", "DMD-MPC", "Koopman-MPC"]
    assert list(table.columns) == list(bt.METRIC_KEYS)
    assert np.allclose(table.loc["DMD-MPC"].values.astype(float), g["metrics"], rtol=2e-3, atol=2e-4)
    ns = len(env.test_dataset) - H
    assert all(len(results[k]) == ns for k in results)
    assert results["Buy & Hold"]["turnover"].abs().max() < 1e-12          # drifted weights are returned as the target
    assert np.isfinite(table.values.astype(float)).all()
    back = pd.read_csv(tmp_path / "full_comparison_metrics.csv", index_col=0)
    assert np.allclose(back.values, table.values.astype(float), rtol=1e-12)
    with pytest.raises(ValueError):
        engine.compare_strategies(m, env, strategies=("nope",))


def test_two_devices_from_one_process():
    """One process driving two GPUs through one handle each (kmpc.h: "one handle per device"): the once-only kernel setup
    (dynamic shared-memory attribute, occupancy, SM count) and the handle's scratch buffers are per DEVICE, so the same
    small batch gives bit-identical metrics on cuda:0 and cuda:1.  Needs two visible GPUs (skipped otherwise; run with
    `gpurun --gpus 2`)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two visible GPUs")
    from koopman_mpc_portfolio_rebalancing_b200 import engine, model as km, synthetic, backtest as bt
    B, N, d, H, rows, Z = 6, 50, 8, 5, 40, 128
    T = rows + d - 1
    lr = synthetic.gbm_log_returns_batch(3, B, T, N)
    mean, std = lr.mean(axis=1), lr.std(axis=1, ddof=1)
    sd = synthetic.generic_km_weights(5, N * d, [128, 128], Z)
    res = []
    for dev in ("cuda:1", "cuda:0", "cuda:1"):                       # device 1 FIRST: nothing may be cached for device 0 only
        m = km.make_model(km.model_config("GenericKM", Z, [128, 128], enc_bias=True), N * d, device=dev)
        m.load_state_dict(sd)
        eng = engine.BatchedBacktester(m, N, d, bt.MPCConfig(horizon=H), bt.BacktestConfig(horizon=H), device=dev)
        with torch.cuda.device(dev):
            out = eng.run(engine.PathBatch(lr, mean, std, 0, rows), want_history=True)
        res.append(out)
    for r in res[1:]:
        assert np.array_equal(r["metrics"], res[0]["metrics"]) and np.array_equal(r["history"], res[0]["history"])
    assert res[0]["stats"][:, 2].sum() == 0
