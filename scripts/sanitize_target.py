"""Small run of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck), one tool per gpurun call:
  timeout 600 compute-sanitizer --tool memcheck python scripts/sanitize_target.py [part]
parts: lane (persistent backtest kernels G = 1, 2, 4 + mpc_solve), gemm (tcgen05 fp16-pair and 3xTF32 chains + SIMT),
       mv (mean-variance kernel), all (default)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from koopman_mpc_portfolio_rebalancing_b200 import _capi, backtest as bt, engine, model as km, mpc, synthetic

part = sys.argv[1] if len(sys.argv) > 1 else "all"


def run_engine(B, N, d, H, rows, enc, Z, fp16=True):
    T = rows + d - 1
    lr = synthetic.gbm_log_returns_batch(7, B, T, N)
    mean, std = lr.mean(axis=1), lr.std(axis=1, ddof=1)
    m = km.make_model(km.model_config("GenericKM", Z, enc, enc_bias=True), N * d)
    m.load_state_dict(synthetic.generic_km_weights(3, N * d, enc, Z))
    eng = engine.BatchedBacktester(m, N, d, bt.MPCConfig(horizon=H), bt.BacktestConfig(horizon=H))
    _capi.lib().kmpc_set_gemm_fp16_pairs(1 if fp16 else 0)
    out = eng.run_device(torch.from_numpy(lr).cuda(), torch.from_numpy(mean).cuda(), torch.from_numpy(std).cuda(), 0, rows,
                         want_history=True)
    torch.cuda.synchronize()
    _capi.lib().kmpc_set_gemm_fp16_pairs(1)
    st = out["stats"].sum(dim=0).cpu().numpy()
    print(f"  engine B={B} N={N} H={H} rows={rows} fp16={fp16}: statuses {st[:3]}, final value {out['metrics'][:, 3].mean().item():.2f}")


if part in ("gemm", "all"):
    print("gemm: tcgen05 fp16-pair chain, 3xTF32 chain, SIMT")
    run_engine(4, 12, 8, 5, 72, [128, 128], 128, fp16=True)      # gemm_tc16_kernel (+ gated gemm_tc launches)
    run_engine(4, 12, 8, 5, 72, [128, 128], 128, fp16=False)     # gemm_tc_kernel
    run_engine(2, 10, 6, 5, 20, [64, 64], 32)                    # gemm_simt_kernel
if part in ("lane", "all"):
    print("lane: persistent backtest kernels")
    run_engine(9, 10, 6, 5, 14, [64, 64], 32)                    # G = 1, 8 slots per block, more backtests than slots in one block
    run_engine(6, 50, 4, 5, 12, [64, 64], 32)                    # G = 2 (config-2 kernel), DMMA border assembly
    run_engine(3, 100, 4, 3, 10, [64, 64], 32)                   # G = 4
    rng = np.random.default_rng(0)
    for (N, H, P) in [(50, 5, 6), (7, 2, 5), (40, 10, 3)]:
        w0 = np.stack([rng.dirichlet(np.ones(N)) for _ in range(P)])
        y = (rng.standard_normal((P, H, N)) * 0.01).astype(np.float32)
        lam = np.array([1e-3, 0.0, 1e-2, 1e-3, 0.0, 1e-3][:P]); tau = np.array([0.2, 0.0, 0.5, 0.0, 0.2, 0.05][:P])
        out = mpc.solve_mpc_batch(torch.from_numpy(w0).cuda(), torch.from_numpy(y).cuda(), lam=torch.from_numpy(lam).cuda(),
                                  tau=torch.from_numpy(tau).cuda())
        torch.cuda.synchronize()
        print(f"  mpc_solve N={N} H={H}: statuses {out['status'].cpu().numpy()}")
if part in ("mv", "all"):
    print("mv: mean-variance kernel")
    rng = np.random.default_rng(1)
    N, H = 12, 2
    A = rng.standard_normal((N, N)) * 0.01
    w, info = mpc.solve_mpc_mean_variance(np.ones(N) / N, rng.standard_normal((H, N)) * 1e-3, A @ A.T + 1e-4 * np.eye(N),
                                          mpc.MPCConfig(horizon=H, gamma=1.0))
    print("  mv status", info["status"])
print("sanitize target done")
