"""Throughput of the persistent MPC kernel by problem width (development tool): B synthetic backtests x 246 decisions,
N assets (G = warps per problem follows from N).   python scripts/lane_shapes.py N [N ...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import numpy as np
    import torch
    from koopman_mpc_portfolio_rebalancing_b200 import backtest as bt
    B, rows, H = 4096, 252, 5
    ns = rows - 1 - H
    for N in [int(a) for a in sys.argv[1:]]:
        g = torch.Generator(device="cuda").manual_seed(N)
        yhat = (3e-4 + 2e-3 * torch.randn((B, ns, H, N), device="cuda", generator=g)).float()
        realized = (3e-4 + 1.2e-2 * torch.randn((B, rows, N), device="cuda", generator=g)).float()
        for _ in range(2):
            out = bt.run_backtest_batched(yhat, realized, n_steps=ns, horizon=H)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = bt.run_backtest_batched(yhat, realized, n_steps=ns, horizon=H)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        st = out["stats"].sum(dim=0).cpu().numpy()
        its = st[3] / (B * ns)
        print(f"N={N}: {ms:8.2f} ms  {B * ns / ms / 1e3:7.2f} M decisions/s  iterations {its:5.2f}  "
              f"ns per iteration (whole GPU) {ms * 1e6 / st[3]:6.2f}  optimal {int(st[0])} / {B * ns}")


if __name__ == "__main__":
    main()
