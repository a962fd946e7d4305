"""Generate the golden fixtures in this directory from the REFERENCE ITSELF (/root/reference, Python,
importable in the build container only).  Run:  python tests/golden/make_golden.py

Nothing under tests/ reads /root/reference at test time: the GPU box does not have it.  The fixtures are
small .npz files; large weights are regenerated from seeds by koopman_mpc_portfolio_rebalancing_b200.synthetic.

  test_sequences_small.npz  FinanceEnv.get_test_sequences / verify_embedding_shift of the data_small env
  data_small.npz       compute_standardization_stats / create_finance_splits / time_delay_embedding
                       (data_finance.py:211-353) on a synthetic business-day frame
  forecast_*.npz       model.encode / step_latent / decode + extract/destandardize rollouts
                       (model.py, backtest.py:85-121) for GenericKM (relu/id, tanh/ball, gelu + MLP decoder)
                       and LISTAKM (linear and MLP encoder); weights stored (tiny models)
  forecast_cfg1.npz    finance_sparse preset, TARGET_SIZE=128, N=10, d=20 (BASELINE config 1); weights from seed
  prices_small.npz     clean_price_data -> compute_log_returns -> stats -> splits from a price frame with gaps
  sequences_small.npz  model.rollout_latent_discrete / rollout_sequence (model.py:527-585) on three tiny models
  rollouts_small.npz   evaluation.rollout_{no,every_step,periodic}_reencode (evaluation.py:44-134) on two tiny models
  markowitz_small.npz  UNMODIFIED reference MarkowitzStrategy + run_backtest (substitute mean-variance solve = fp64 oracle)
  dmd_small.npz        UNMODIFIED reference DMDStrategy (baselines.py:109-187) + run_backtest on a small env: fitted K,
                       the forecasts handed to the MPC, history, metrics
  checkpoint_*.pt      files in the layout train.py:475-487 saves (reference model + Config.to_dict() + Adam state), for two
                       tiny models whose forecasts are in forecast_generic_small.npz / forecast_lista_linear.npz
  backtest_cfg1.npz    UNMODIFIED reference run_backtest + KoopmanMPCStrategy + calculate_metrics on config 1
                       with the substitute mpc module (tests/golden/_shims/mpc.py): history, metrics, every
                       MPC call's (w_cur, yhat, w_opt, value)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
sys.path.insert(0, os.path.join(HERE, "_shims"))  # must precede /root/reference: mpc, matplotlib

import pandas as pd  # noqa: E402
import torch  # noqa: E402

import config as ref_config  # noqa: E402
import data_finance as ref_data  # noqa: E402
import model as ref_model  # noqa: E402
import backtest as ref_backtest  # noqa: E402
import mpc as shim_mpc  # noqa: E402

from koopman_mpc_portfolio_rebalancing_b200 import synthetic  # noqa: E402

torch.set_num_threads(1)


def make_frame(seed, T, N, start="2012-01-02"):
    lr = synthetic.gbm_log_returns(seed, T, N)
    dates = pd.bdate_range(start, periods=T)
    return pd.DataFrame(lr, index=dates, columns=[f"A{i}" for i in range(N)])


def build_env(frame, train_end, val_end, d, seq_len=1):
    stats = ref_data.compute_standardization_stats(frame, train_end)
    tr, trd, va, vad, te, ted = ref_data.create_finance_splits(frame, stats, train_end, val_end, d)
    env = ref_data.FinanceEnv(
        ref_data.FinanceDataset(tr, trd, seq_len), ref_data.FinanceDataset(va, vad, seq_len),
        ref_data.FinanceDataset(te, ted, seq_len), stats,
        {"n_assets": frame.shape[1], "embedding_dim": d})
    return env, stats, (tr, va, te)


def gen_data_small():
    frame = make_frame(11, 80, 3)
    d = 4
    train_end, val_end = str(frame.index[39].date()), str(frame.index[59].date())
    env, stats, (tr, va, te) = build_env(frame, train_end, val_end, d)
    std_all = ref_data.standardize_returns(frame, stats).values.astype(np.float32)
    emb = ref_data.time_delay_embedding(std_all, d)
    np.savez(os.path.join(HERE, "data_small.npz"), log_returns=frame.values, n_train_days=40, n_val_days=20,
             d=d, mean=stats.mean, std=stats.std, standardized=std_all, embedded=emb, train=tr, val=va, test=te,
             test_len=len(env.test_dataset))
    # FinanceEnv.get_test_sequences (data_finance.py:672-715) and verify_embedding_shift (:515-540) of the same env
    i1, f1 = env.get_test_sequences(num_sequences=5, max_length=7)
    i2, f2 = env.get_test_sequences()                       # defaults: clamped by the 20-row test split
    np.savez(os.path.join(HERE, "test_sequences_small.npz"), init_5_7=i1.numpy(), future_5_7=f1.numpy(),
             init_default=i2.numpy(), future_default=f2.numpy(),
             shift_ok=ref_data.verify_embedding_shift(emb, 3, d),
             shift_broken=ref_data.verify_embedding_shift(emb[::2], 3, d))


def ref_forecast(model, env_like, obs, H):
    """The forecast loop of KoopmanMPCStrategy.rebalance (backtest.py:85-121), batched over rows."""
    out = []
    with torch.no_grad():
        model.eval()
        z = model.encode(torch.from_numpy(obs))
        for _ in range(H):
            z = model.step_latent(z)
            x = model.decode(z)
            y = env_like.destandardize_returns(env_like.extract_current_returns(x))
            out.append(y.numpy())
    return np.stack(out, axis=1)  # [M, H, N]


class EnvLike:
    def __init__(self, N, mean, std):
        self.n_assets = N
        self.stats = ref_data.FinanceStats(mean=mean, std=std, tickers=[])

    extract_current_returns = ref_data.FinanceEnv.extract_current_returns
    destandardize_returns = ref_data.FinanceEnv.destandardize_returns


def gen_forecast(name, cfg, N, d, sd_np, H, M, seed, store_weights=True, lista_L=None):
    obs_size = N * d
    model = ref_model.make_model(cfg, obs_size)
    missing = model.load_state_dict({k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in sd_np.items()}, strict=True)
    rng = np.random.default_rng(seed)
    obs = rng.standard_normal((M, obs_size)).astype(np.float32)
    mean = rng.normal(3e-4, 1e-4, N)
    std = rng.uniform(0.008, 0.02, N)
    yhat = ref_forecast(model, EnvLike(N, mean, std), obs, H)
    with torch.no_grad():
        z0 = model.encode(torch.from_numpy(obs)).numpy()
    out = dict(obs=obs, mean=mean, std=std, yhat=yhat, z0=z0, H=H, N=N, d=d)
    if store_weights:
        out.update({"sd::" + k: v for k, v in sd_np.items()})
    np.savez(os.path.join(HERE, f"forecast_{name}.npz"), **out)
    return model


def gen_forecasts():
    # 1. GenericKM relu / id / bias / linear decoder (shape of the finance_sparse preset, tiny)
    cfg = ref_config.get_config("finance_sparse")
    cfg.MODEL.TARGET_SIZE = 8
    cfg.MODEL.ENCODER.LAYERS = [16, 16]
    sd = synthetic.generic_km_weights(1, 12, [16, 16], 8)
    gen_forecast("generic_small", cfg, 3, 4, sd, 5, 7, 101)
    # 2. tanh, ball norm, last_relu False, no bias
    cfg = ref_config.get_config("generic")
    cfg.MODEL.TARGET_SIZE = 8
    cfg.MODEL.ENCODER.LAYERS = [16]
    cfg.MODEL.ENCODER.ACTIVATION = "tanh"
    cfg.MODEL.NORM_FN = "ball"
    sd = synthetic.generic_km_weights(2, 12, [16], 8, enc_bias=False)
    gen_forecast("generic_tanh_ball", cfg, 3, 4, sd, 3, 5, 102)
    # 3. gelu encoder with last_relu, MLP decoder (relu) with bias
    cfg = ref_config.get_config("generic_sparse")
    cfg.MODEL.TARGET_SIZE = 8
    cfg.MODEL.ENCODER.LAYERS = [16, 12]
    cfg.MODEL.ENCODER.ACTIVATION = "gelu"
    cfg.MODEL.DECODER.LAYERS = [10]
    cfg.MODEL.DECODER.USE_BIAS = True
    sd = synthetic.generic_km_weights(3, 12, [16, 12], 8, dec_layers=[10], dec_bias=True)
    gen_forecast("generic_gelu_mlpdec", cfg, 3, 4, sd, 4, 6, 103)
    # 4. LISTAKM, linear encoder
    cfg = ref_config.get_config("lista")
    cfg.MODEL.TARGET_SIZE = 16
    sd, L = synthetic.lista_km_weights(4, 12, 16)
    cfg.MODEL.ENCODER.LISTA.L = L
    gen_forecast("lista_linear", cfg, 3, 4, sd, 5, 6, 104)
    np.savez(os.path.join(HERE, "forecast_lista_linear_meta.npz"), L=L, alpha=cfg.MODEL.ENCODER.LISTA.ALPHA,
             loops=cfg.MODEL.ENCODER.LISTA.NUM_LOOPS)
    # 5. LISTAKM, MLP encoder (lista_nonlinear preset: relu, bias, last_relu True)
    cfg = ref_config.get_config("lista_nonlinear")
    cfg.MODEL.TARGET_SIZE = 16
    cfg.MODEL.ENCODER.LAYERS = [12, 12]
    sd, L = synthetic.lista_km_weights(5, 12, 16, enc_layers=[12, 12])
    cfg.MODEL.ENCODER.LISTA.L = 50.0
    cfg.MODEL.ENCODER.LISTA.ALPHA = 0.5
    gen_forecast("lista_mlp", cfg, 3, 4, sd, 5, 6, 105)
    np.savez(os.path.join(HERE, "forecast_lista_mlp_meta.npz"), L=50.0, alpha=0.5, loops=cfg.MODEL.ENCODER.LISTA.NUM_LOOPS)


def gen_forecast_cfg2():
    """BASELINE config 2 at its FULL architecture (finance_sparse code defaults: 1000 -> 1024 -> 1024 -> 1024, linear
    decoder, the weights bench.py uses) through the reference model.py (torch CPU fp32), the forecast loop of
    KoopmanMPCStrategy.rebalance (backtest.py:85-121) on 2 scenario paths x 246 rows of the bench inputs
    (bench.make_inputs(cfg2, 2, 10_000): GBM paths 10 000 and 10 001).  Also the same forecasts from the model in
    float64 (model.double()) to bound the reference's own fp32 rounding."""
    import bench
    w = bench.WORKLOADS["cfg2"]
    N, d, H, Z, rows = w["N"], w["d"], w["H"], w["Z"], w["rows"]
    ns = rows - 1 - H
    cfg = ref_config.get_config("finance_sparse")
    assert cfg.MODEL.TARGET_SIZE == Z and list(cfg.MODEL.ENCODER.LAYERS) == w["enc"]
    sd = synthetic.generic_km_weights(0, N * d, w["enc"], Z)
    model = ref_model.make_model(cfg, N * d)
    model.load_state_dict({k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in sd.items()}, strict=True)
    lr, mean, std, T = bench.make_inputs(w, 2, 10_000)
    yh32, yh64 = [], []
    torch.set_num_threads(8)
    for b in range(2):
        frame = pd.DataFrame(lr[b])
        z = ((frame - mean[b]) / std[b]).values.astype(np.float32)       # standardize_returns + the f32 cast (data_finance.py:243-259, :331)
        emb = ref_data.time_delay_embedding(z, d)
        env = EnvLike(N, mean[b], std[b])
        yh32.append(ref_forecast(model, env, emb[:ns], H))
        m64 = ref_model.make_model(cfg, N * d).double()
        m64.load_state_dict({k: torch.from_numpy(np.ascontiguousarray(v)).double() for k, v in sd.items()}, strict=True)
        out = []
        with torch.no_grad():
            m64.eval()
            zz = m64.encode(torch.from_numpy(emb[:ns]).double())
            for _ in range(H):
                zz = m64.step_latent(zz)
                x = m64.decode(zz)[..., :N]
                out.append((x * torch.from_numpy(std[b]) + torch.from_numpy(mean[b])).numpy())
        yh64.append(np.stack(out, axis=1))
    torch.set_num_threads(1)
    yh32 = np.stack(yh32).astype(np.float32); yh64 = np.stack(yh64)
    rel = np.abs(yh32 - yh64).reshape(2 * ns, -1).max(1) / np.abs(yh64).reshape(2 * ns, -1).max(1)
    print("cfg2 forecast golden: reference fp32 vs its own fp64 model, worst row-wise rel", rel.max())
    np.savez_compressed(os.path.join(HERE, "forecast_cfg2.npz"), yhat=yh32, yhat_f64model=yh64.astype(np.float32),
                        seed=10_000, paths=2, ns=ns)


def cfg1_model():
    cfg = ref_config.get_config("finance_sparse")
    cfg.MODEL.TARGET_SIZE = 128
    sd = synthetic.generic_km_weights(0, 200, [1024, 1024], 128)
    return cfg, sd


def gen_cfg1():
    cfg, sd = cfg1_model()
    model = gen_forecast("cfg1", cfg, 10, 20, sd, 5, 16, 106, store_weights=False)
    # --- the reference backtest, unmodified, on config 1 ---
    N, d, H = 10, 20, 5
    T = 1200
    frame = make_frame(0, T, N)
    # test split = exactly 252 embedded rows
    val_end = str(frame.index[T - 253].date())
    train_end = str(frame.index[T - 253 - 200].date())
    env, stats, (tr, va, te) = build_env(frame, train_end, val_end, d)
    assert te.shape == (252, 200), te.shape
    mpc_cfg = shim_mpc.MPCConfig(horizon=H, cost_coeff=1e-3, max_turnover=0.2)
    bt_cfg = ref_backtest.BacktestConfig(initial_capital=1e4, horizon=H, cost_coeff=1e-3)
    shim_mpc.CALLS.clear()
    strat = ref_backtest.KoopmanMPCStrategy(model, mpc_cfg, device="cpu")
    df = ref_backtest.run_backtest(strat, env, bt_cfg, verbose=False)
    metrics = ref_backtest.calculate_metrics(df)
    calls = shim_mpc.CALLS
    assert len(df) == 246 == len(calls)
    bh = ref_backtest.run_backtest(ref_backtest.BuyAndHoldStrategy(), env, bt_cfg, verbose=False)
    bh_metrics = ref_backtest.calculate_metrics(bh)
    np.savez(
        os.path.join(HERE, "backtest_cfg1.npz"),
        T=T, n_train_days=T - 253 - 200 + 1, n_val_days=200, log_returns_seed=0,
        mean=stats.mean, std=stats.std,
        history=df[["portfolio_value", "return", "turnover", "cost"]].values.astype(np.float64),
        metrics=np.array([metrics[k] for k in ("Sharpe Ratio", "Max Drawdown", "Avg Turnover", "Final Value", "Total Return")]),
        bh_history=bh[["portfolio_value", "return", "turnover", "cost"]].values.astype(np.float64),
        bh_metrics=np.array([bh_metrics[k] for k in ("Sharpe Ratio", "Max Drawdown", "Avg Turnover", "Final Value", "Total Return")]),
        yhat=np.stack([c[1] for c in calls]).astype(np.float32),
        w_cur=np.stack([c[0] for c in calls]), w_opt=np.stack([c[2] for c in calls]),
        value=np.array([np.nan if c[3] is None else c[3] for c in calls]),
        test_first_rows=te[:3], test_len=len(env.test_dataset),
    )
    print("cfg1 metrics", metrics)
    print("buy&hold    ", bh_metrics)


def gen_markowitz(N=6, name="markowitz_small.npz", keep_every=1):
    """UNMODIFIED reference MarkowitzStrategy (baselines.py:24-106) + run_backtest on the small env of gen_dmd: the
    (mu, Sigma) it estimates at every step, the weights the (substitute, fp64 oracle) mean-variance solve returns,
    history and metrics.  N = 50 (markowitz_n50.npz): the 50-asset case of VERDICT item 9; only every 8th solver call
    is stored (Sigma is 50 x 50)."""
    import baselines as ref_baselines
    d, H = 4, 1
    T = 700
    frame = make_frame(21, T, N)
    val_end = str(frame.index[T - 61].date())
    train_end = str(frame.index[T - 61 - 100].date())
    env, stats, _ = build_env(frame, train_end, val_end, d)
    bt_cfg = ref_backtest.BacktestConfig(initial_capital=1e4, horizon=H, cost_coeff=1e-3)
    strat = ref_baselines.MarkowitzStrategy(risk_aversion=2.0, cost_coeff=1e-3)
    shim_mpc.MV_CALLS.clear()
    df = ref_backtest.run_backtest(strat, env, bt_cfg, verbose=False)
    metrics = ref_backtest.calculate_metrics(df)
    calls = shim_mpc.MV_CALLS[::keep_every]
    np.savez_compressed(
        os.path.join(HERE, name), call_stride=keep_every,
        T=T, N=N, d=d, log_returns_seed=21, n_train_days=T - 61 - 100 + 1, n_val_days=100, gamma=2.0,
        mean=stats.mean, std=stats.std,
        mu=np.stack([c[1] for c in calls]), sigma=np.stack([c[2] for c in calls]),
        w_cur=np.stack([c[0] for c in calls]), w_opt=np.stack([c[3] for c in calls]),
        value=np.array([np.nan if c[4] is None else c[4] for c in calls]),
        history=df[["portfolio_value", "return", "turnover", "cost"]].values.astype(np.float64),
        metrics=np.array([metrics[k] for k in ("Sharpe Ratio", "Max Drawdown", "Avg Turnover", "Final Value", "Total Return")]),
    )
    print("markowitz metrics", metrics, "calls", len(calls), "steps", len(df))


def gen_rollouts():
    """reference evaluation.rollout_no_reencode / rollout_every_step_reencode / rollout_periodic_reencode
    (evaluation.py:44-134) on two of the tiny forecast models (weights already stored in forecast_*.npz)."""
    import evaluation as ref_eval
    out = {}
    for name, preset in (("generic_small", "finance_sparse"), ("lista_linear", "lista")):
        g = np.load(os.path.join(HERE, f"forecast_{name}.npz"))
        sd = {k[4:]: g[k] for k in g.files if k.startswith("sd::")}
        cfg = ref_config.get_config(preset)
        if name == "generic_small":
            cfg.MODEL.TARGET_SIZE = 8; cfg.MODEL.ENCODER.LAYERS = [16, 16]
        else:
            cfg.MODEL.TARGET_SIZE = 16
            cfg.MODEL.ENCODER.LISTA.L = float(np.load(os.path.join(HERE, "forecast_lista_linear_meta.npz"))["L"])
        model = ref_model.make_model(cfg, g["obs"].shape[1])
        model.load_state_dict({k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in sd.items()}, strict=True)
        x0 = torch.from_numpy(g["obs"])
        out[f"{name}::no_reencode"] = ref_eval.rollout_no_reencode(model, x0, 6).numpy()
        out[f"{name}::every_step"] = ref_eval.rollout_every_step_reencode(model, x0, 6).numpy()
        out[f"{name}::periodic2"] = ref_eval.rollout_periodic_reencode(model, x0, 6, 2).numpy()
        # train.evaluate_finance (train.py:221-300) against a seeded "future": curves of every mode, best mode
        import train as ref_train
        fut = np.random.default_rng(77).standard_normal((8, x0.shape[0], x0.shape[1])).astype(np.float32) * 0.5
        ev = ref_train.evaluate_finance(model, x0, torch.from_numpy(fut), max_horizon=6, periodic_reencode_periods=[2, 3])
        out[f"{name}::eval_future"] = fut
        for mode, curve in ev["mse_curves"].items():
            out[f"{name}::eval_mse::{mode}"] = curve.numpy()
            out[f"{name}::eval_l2::{mode}"] = ev["l2_curves"][mode].numpy()
        out[f"{name}::eval_best"] = np.array(list(ev["mse_curves"]).index(ev["best_mode"]))
        out[f"{name}::eval_scalars"] = np.array([ev["mean_mse_reencode"], ev["mean_mse_no_reencode"],
                                                 ev["final_mse_reencode"], ev["final_mse_no_reencode"], ev["best_mse"]])
    np.savez(os.path.join(HERE, "rollouts_small.npz"), **out)
    print("rollouts", {k: v.shape for k, v in out.items()})


def gen_prices():
    """reference clean_price_data -> compute_log_returns -> compute_standardization_stats -> create_finance_splits
    (data_finance.py:147-353) on a synthetic price frame with gaps (short gaps are filled, a 7-day gap and a leading
    NaN drop rows, one asset with 20 % missing is dropped)."""
    rng = np.random.default_rng(21)
    T, N, d = 90, 4, 3
    lr = synthetic.gbm_log_returns(21, T, N)
    prices = pd.DataFrame(100.0 * np.exp(np.cumsum(lr, axis=0)), index=pd.bdate_range("2015-01-05", periods=T),
                          columns=[f"P{i}" for i in range(N)])
    prices.iloc[10:12, 0] = np.nan            # 2-day gap: forward-filled
    prices.iloc[30:37, 1] = np.nan            # 7-day gap: 5 filled, 2 rows dropped
    prices.iloc[0, 2] = np.nan                # nothing to fill from: row dropped
    prices.iloc[rng.choice(T, 18, replace=False), 3] = np.nan     # 20 % missing: asset dropped
    clean = ref_data.clean_price_data(prices)
    logret = ref_data.compute_log_returns(clean)
    train_end, val_end = str(logret.index[49].date()), str(logret.index[69].date())
    stats = ref_data.compute_standardization_stats(logret, train_end)
    tr, trd, va, vad, te, ted = ref_data.create_finance_splits(logret, stats, train_end, val_end, d)
    np.savez(os.path.join(HERE, "prices_small.npz"), prices=prices.values, clean=clean.values,
             clean_rows=np.array([prices.index.get_loc(i) for i in clean.index]),
             clean_cols=np.array([list(prices.columns).index(c) for c in clean.columns]),
             log_returns=logret.values, train_end=train_end, val_end=val_end, d=d, mean=stats.mean, std=stats.std,
             train=tr, val=va, test=te)
    print("prices", prices.shape, "->", clean.shape, tr.shape, va.shape, te.shape)


def gen_sequences():
    """reference rollout_latent_discrete / rollout_sequence (model.py:527-585; NO latent normalisation in the unroll,
    unlike step_latent) on three of the tiny forecast models, the ball-normalised one included."""
    out = {}
    for name in ("generic_small", "generic_tanh_ball", "lista_linear"):
        g = np.load(os.path.join(HERE, f"forecast_{name}.npz"))
        sd = {k[4:]: g[k] for k in g.files if k.startswith("sd::")}
        if name == "generic_small":
            cfg = ref_config.get_config("finance_sparse"); cfg.MODEL.TARGET_SIZE = 8; cfg.MODEL.ENCODER.LAYERS = [16, 16]
        elif name == "generic_tanh_ball":
            cfg = ref_config.get_config("generic"); cfg.MODEL.TARGET_SIZE = 8; cfg.MODEL.ENCODER.LAYERS = [16]
            cfg.MODEL.ENCODER.ACTIVATION = "tanh"; cfg.MODEL.NORM_FN = "ball"
        else:
            cfg = ref_config.get_config("lista"); cfg.MODEL.TARGET_SIZE = 16
            cfg.MODEL.ENCODER.LISTA.L = float(np.load(os.path.join(HERE, "forecast_lista_linear_meta.npz"))["L"])
        model = ref_model.make_model(cfg, g["obs"].shape[1])
        model.load_state_dict({k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in sd.items()}, strict=True)
        model.eval()
        with torch.no_grad():
            x0 = torch.from_numpy(g["obs"])
            out[f"{name}::latent"] = model.rollout_latent_discrete(model.encode(x0), 4).numpy()
            out[f"{name}::sequence"] = model.rollout_sequence(x0, 4).numpy()
    np.savez(os.path.join(HERE, "sequences_small.npz"), **out)
    print("sequences", {k: v.shape for k, v in out.items()})


def gen_dmd():
    """UNMODIFIED reference DMDStrategy (baselines.py:109-187) + run_backtest on a small synthetic env: the fitted K,
    the forecasts it hands to the MPC at every step, the history and the metrics."""
    import baselines as ref_baselines
    N, d, H = 6, 4, 5
    T = 700
    frame = make_frame(21, T, N)
    val_end = str(frame.index[T - 61].date())
    train_end = str(frame.index[T - 61 - 100].date())
    env, stats, (tr, va, te) = build_env(frame, train_end, val_end, d)
    mpc_cfg = shim_mpc.MPCConfig(horizon=H, cost_coeff=1e-3, max_turnover=0.2)
    bt_cfg = ref_backtest.BacktestConfig(initial_capital=1e4, horizon=H, cost_coeff=1e-3)
    strat = ref_baselines.DMDStrategy(env.train_dataset.data, mpc_cfg)
    shim_mpc.CALLS.clear()
    df = ref_backtest.run_backtest(strat, env, bt_cfg, verbose=False)
    metrics = ref_backtest.calculate_metrics(df)
    calls = shim_mpc.CALLS
    assert len(df) == len(calls) == len(env.test_dataset) - H
    np.savez(
        os.path.join(HERE, "dmd_small.npz"),
        T=T, N=N, d=d, H=H, log_returns_seed=21, n_train_days=T - 61 - 100 + 1, n_val_days=100,
        mean=stats.mean, std=stats.std, K=strat.K,
        yhat=np.stack([c[1] for c in calls]).astype(np.float32),
        w_cur=np.stack([c[0] for c in calls]), w_opt=np.stack([c[2] for c in calls]),
        history=df[["portfolio_value", "return", "turnover", "cost"]].values.astype(np.float64),
        metrics=np.array([metrics[k] for k in ("Sharpe Ratio", "Max Drawdown", "Avg Turnover", "Final Value", "Total Return")]),
    )
    print("dmd metrics", metrics, "K dtype", strat.K.dtype)


def gen_checkpoints():
    """The dictionary train.py:475-483 saves, written by the reference's own classes (make_model state_dict,
    Config.to_dict, torch Adam) for the tiny models of forecast_generic_small.npz and forecast_lista_linear.npz."""
    def one(name, cfg, N, d, sd_np):
        model = ref_model.make_model(cfg, N * d)
        model.load_state_dict({k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in sd_np.items()}, strict=True)
        opt = torch.optim.Adam(model.parameters(), lr=1e-3)
        meta = {"tickers": [f"A{i}" for i in range(N)], "n_assets": N, "embedding_dim": d, "observation_size": N * d,
                "train_samples": 40, "val_samples": 10, "test_samples": 10,
                "train_date_range": ("2012-01-05", "2012-03-01"), "prices_shape": (80, N)}
        torch.save({"step": 123, "epoch": 4, "model_state_dict": model.state_dict(), "optimizer_state_dict": opt.state_dict(),
                    "config": cfg.to_dict(), "metrics": {"loss": 0.5}, "finance_metadata": meta},
                   os.path.join(HERE, f"checkpoint_{name}.pt"))
    cfg = ref_config.get_config("finance_sparse")
    cfg.MODEL.TARGET_SIZE = 8
    cfg.MODEL.ENCODER.LAYERS = [16, 16]
    one("generic_small", cfg, 3, 4, synthetic.generic_km_weights(1, 12, [16, 16], 8))
    cfg = ref_config.get_config("lista")
    cfg.MODEL.TARGET_SIZE = 16
    sd, L = synthetic.lista_km_weights(4, 12, 16)
    cfg.MODEL.ENCODER.LISTA.L = L
    one("lista_linear", cfg, 3, 4, sd)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "prices":
        gen_prices()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "sequences":
        gen_sequences()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "data":
        gen_data_small()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "forecast_cfg2":
        gen_forecast_cfg2()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "checkpoints":
        gen_checkpoints()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "dmd":
        gen_dmd()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "markowitz":
        gen_markowitz()
        gen_markowitz(50, "markowitz_n50.npz", 8)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "rollouts":
        gen_rollouts()
        sys.exit(0)
    gen_data_small()
    gen_forecasts()
    gen_forecast_cfg2()
    gen_cfg1()
    gen_dmd()
    gen_rollouts()
    gen_markowitz()
    gen_markowitz(50, "markowitz_n50.npz", 8)
    gen_checkpoints()
    gen_sequences()
    gen_prices()
    print("golden fixtures written to", HERE)
