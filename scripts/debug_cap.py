"""Diagnostic: run BASELINE config 2 at full size on the GPU and list every decision whose first-stage turnover exceeds the
cap (should print `violations 0`) together with the solver outcome counts.  Usage: python scripts/debug_cap.py"""
import numpy as np, torch, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from koopman_mpc_portfolio_rebalancing_b200 import engine, model as km, synthetic, backtest as bt
w = bench.WORKLOADS["cfg2"]
B, N, d, H, Z, rows = w["B"], w["N"], w["d"], w["H"], w["Z"], w["rows"]
lr, mean, std, T = bench.make_inputs(w, B, 10_000)
m = km.make_model(km.model_config("GenericKM", Z, w["enc"], enc_bias=True), N * d)
m.load_state_dict(synthetic.generic_km_weights(0, N * d, w["enc"], Z))
eng = engine.BatchedBacktester(m, N, d, bt.MPCConfig(horizon=H, cost_coeff=1e-3, max_turnover=0.2),
                               bt.BacktestConfig(initial_capital=1e4, horizon=H, cost_coeff=1e-3))
out = eng.run_device(torch.from_numpy(lr).cuda(), torch.from_numpy(mean).cuda(), torch.from_numpy(std).cuda(), 0, rows, want_history=True)
hist = out["history"].cpu().numpy(); stats = out["stats"].cpu().numpy()
print("stats", stats[:, :4].sum(0), "max turn", hist[:, :, 2].max(), "min V", hist[:, :, 0].min())
idx = np.argwhere(hist[:, :, 2] > 0.2 + 1e-7)
print("violations", len(idx), idx[:10])
for b, t in idx[:5]:
    print(b, t, hist[b, max(0, t - 1):t + 2])
