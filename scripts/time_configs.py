"""Ad-hoc timing of the BASELINE configs other than config 2 (parity-test cases, not bench lines): config 3 (LISTAKM,
500 assets, d = 10, H = 10) and config 5 (100 assets, bootstrap paths) on one GPU, CUDA events per stage.
Usage: python scripts/time_configs.py [cfg3_backtests] [cfg5_backtests]"""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from koopman_mpc_portfolio_rebalancing_b200 import engine, model as km, synthetic, backtest as bt


def run(name, eng, lr, mean, std, rows):
    lr_d = torch.from_numpy(lr).cuda() if isinstance(lr, np.ndarray) else lr
    mean_d, std_d = torch.from_numpy(mean).cuda(), torch.from_numpy(std).cuda()
    for _ in range(2):
        tm = {}
        out = eng.run_device(lr_d, mean_d, std_d, 0, rows, timings=tm)
    torch.cuda.synchronize()
    ev = tm["_events"]
    B = lr_d.shape[0]; ns = eng.n_steps(rows)
    st = out["stats"].cpu().numpy()
    ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(3)]
    print(f"{name}: {B} backtests x {ns} decisions: data {ms[0]:.1f} ms, forecast {ms[1]:.1f} ms, mpc+portfolio {ms[2]:.1f} ms "
          f"-> {B * ns / (sum(ms) * 1e-3):.3e} decisions/s; iterations/decision {st[:, 3].sum() / (B * ns):.1f}, "
          f"optimal {st[:, 0].sum()}, inaccurate {st[:, 1].sum()}, fallback {st[:, 2].sum()}")


b3 = int(sys.argv[1]) if len(sys.argv) > 1 else 148
b5 = int(sys.argv[2]) if len(sys.argv) > 2 else 2368
# config 3: LISTAKM linear encoder, 500 assets, d = 10, Z = 2048, 10 loops, H = 10, turnover cap 0.2
N, d, H, Z, rows = 500, 10, 10, 2048, 252
T = rows + d - 1
lr = synthetic.gbm_log_returns_batch(1000, b3, T, N)
mean = lr.mean(axis=1); std = np.maximum(lr.std(axis=1, ddof=1), 1e-8)
sd, L = synthetic.lista_km_weights(0, N * d, Z)
m = km.make_model(km.model_config("LISTAKM", Z, lista_loops=10, lista_L=L, lista_alpha=5e-3, lista_linear=True), N * d)
m.load_state_dict(sd)
eng = engine.BatchedBacktester(m, N, d, bt.MPCConfig(horizon=H, cost_coeff=1e-3, max_turnover=0.2), bt.BacktestConfig(horizon=H))
run("cfg3 (LISTAKM 500 assets, H=10)", eng, lr, mean, std, rows)
del eng, m
# config 5: GenericKM, 100 assets, bootstrap paths of one 3000-day block
N, d, H, Z, rows = 100, 20, 5, 1024, 252
T = rows + d - 1
hist = synthetic.gbm_log_returns(0, 3000, N)
paths, _ = engine.bootstrap_paths(hist, b5, T, seed=1234)
mean = np.tile(hist.mean(axis=0), (b5, 1)); std = np.tile(hist.std(axis=0, ddof=1), (b5, 1))
m = km.make_model(km.model_config("GenericKM", Z, [1024, 1024], enc_bias=True), N * d)
m.load_state_dict(synthetic.generic_km_weights(0, N * d, [1024, 1024], Z))
eng = engine.BatchedBacktester(m, N, d, bt.MPCConfig(horizon=H), bt.BacktestConfig(horizon=H))
run("cfg5 (GenericKM 100 assets, bootstrap paths)", eng, paths, mean, std, rows)
del eng, m, paths
torch.cuda.empty_cache()
# config 4: sweep grid on ONE price path: 16 weight sets x 64 lambda x 64 tau = 65 536 backtests sharing 16 forecast sets
b4 = int(sys.argv[3]) if len(sys.argv) > 3 else 64
N, d, H, Z, rows = 50, 20, 5, 1024, 252
T = rows + d - 1
lr1 = synthetic.gbm_log_returns(0, T, N)
models = []
for s_ in range(16):
    mm = km.make_model(km.model_config("GenericKM", Z, [1024, 1024], enc_bias=True), N * d)
    mm.load_state_dict(synthetic.generic_km_weights(s_, N * d, [1024, 1024], Z))
    models.append(mm)
lam_grid = np.logspace(-5, -1, b4); tau_grid = np.linspace(0.01, 1.0, b4)
for _ in range(2):
    torch.cuda.synchronize(); t0 = time.time()
    out = engine.run_grid(models, N, d, lr1, lr1.mean(axis=0), np.maximum(lr1.std(axis=0, ddof=1), 1e-8), lam_grid, tau_grid, rows=rows, horizon=H)
    torch.cuda.synchronize(); dt = time.time() - t0
st = out["stats"].cpu().numpy(); nb = st.shape[0]; ns = rows - 1 - H
print(f"cfg4 (grid 16 x {b4} x {b4} on one path): {nb} backtests x {ns} decisions in {dt * 1e3:.0f} ms wall -> {nb * ns / dt:.3e} decisions/s; "
      f"iterations/decision {st[:, 3].sum() / (nb * ns):.1f}, optimal {st[:, 0].sum()}, inaccurate {st[:, 1].sum()}, fallback {st[:, 2].sum()}")
