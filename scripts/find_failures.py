"""Replay config-2 backtests step by step with the batched MPC API and collect the instances whose solve does not
report 'optimal' (diagnostics for the interior-point solver).  Runs on a GPU box; writes gpurun_out/failures.npz."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from koopman_mpc_portfolio_rebalancing_b200 import engine, model as km, synthetic, backtest as bt, mpc
import bench

w = bench.WORKLOADS["cfg2"]
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
N, d, H, Z, rows = w["N"], w["d"], w["H"], w["Z"], w["rows"]
m = km.make_model(km.model_config("GenericKM", Z, w["enc"], enc_bias=True), N * d)
m.load_state_dict(synthetic.generic_km_weights(0, N * d, w["enc"], Z))
eng = engine.BatchedBacktester(m, N, d, bt.MPCConfig(horizon=H), bt.BacktestConfig(horizon=H))
lr, mean, std, T = bench.make_inputs(w, B, 10_000)
out = eng.run_device(torch.from_numpy(lr).cuda(), torch.from_numpy(mean).cuda(), torch.from_numpy(std).cuda(), 0, rows)
yhat, realized = out["yhat"], out["realized"]
ns = yhat.shape[1]
wc = torch.full((B, N), 1.0 / N, dtype=torch.float64, device="cuda")
bad = {"w": [], "y": [], "st": [], "kkt": [], "it": []}
for t in range(ns):
    r = mpc.solve_mpc_batch(wc, yhat[:, t].contiguous())
    st = r["status"]
    idx = torch.nonzero(st != 0).flatten()
    for i in idx.tolist():
        bad["w"].append(wc[i].cpu().numpy()); bad["y"].append(yhat[i, t].cpu().numpy()); bad["st"].append(int(st[i]))
        bad["kkt"].append(r["kkt"][i].cpu().numpy()); bad["it"].append(int(r["iterations"][i]))
    wn = r["w"][:, 0, :]
    rr = torch.exp(realized[:, t + 1].double()).float() - 1.0
    pr = (wn * rr.double()).sum(dim=1, keepdim=True)
    wc = wn * (1.0 + rr).double() / (1.0 + pr)
print("non-optimal:", len(bad["st"]), "of", B * ns, "statuses", np.bincount(bad["st"], minlength=4))
os.makedirs("gpurun_out", exist_ok=True)
np.savez("gpurun_out/failures.npz", **{k: np.array(v) for k, v in bad.items()})
if bad["st"]:
    k = np.array(bad["kkt"]); print("kkt of failures (first 10):\n", k[:10]); print("iters", bad["it"][:20])
