"""Latency of one Newton iteration per problem width at H = 10 with ONE problem per SM (development tool): 148 synthetic
backtests x 40 decisions, full-width solves only (active-set route off).   python scripts/lane_latency.py N [N ...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from koopman_mpc_portfolio_rebalancing_b200 import _capi, backtest as bt
    _capi.check(_capi.lib().kmpc_set_solver_param(_capi.Handle.get(0).ptr, 7, 0.0))     # KMPC_PARAM_ACTIVE_SET off
    B, H, ns = 148, 10, 40
    rows = ns + 1 + H
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
        print(f"N={N}: {ms:8.2f} ms  iterations/decision {its:5.2f}  us per iteration and problem {ms * 1e3 / (ns * its):7.2f}  "
              f"optimal {int(st[0])} / {B * ns}", flush=True)


if __name__ == "__main__":
    main()
