"""Small run of the active-set backtest pipeline (development tool).  python scripts/as_debug.py [lib.so] [B] [rows]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
def main():
    import numpy as np, torch
    from koopman_mpc_portfolio_rebalancing_b200 import _capi
    args = sys.argv[1:]
    if args and args[0].endswith(".so"):
        _capi.LIB_PATH = args.pop(0)
    B = int(args[0]) if args else 8
    rows = int(args[1]) if len(args) > 1 else 40
    from koopman_mpc_portfolio_rebalancing_b200 import backtest as bt
    N, H = 50, 5
    ns = rows - 1 - H
    g = torch.Generator(device="cuda").manual_seed(1)
    # persistent forecasts (as a trained model gives): a per-asset drift plus small daily noise
    drift = 2e-3 * torch.randn((B, 1, 1, N), device="cuda", generator=g)
    yhat = (3e-4 + drift + 2e-4 * torch.randn((B, ns, H, N), device="cuda", generator=g)).float()
    realized = (3e-4 + 1.2e-2 * torch.randn((B, rows, N), device="cuda", generator=g)).float()
    res = {}
    modes = [int(x) for x in args[2].split(",")] if len(args) > 2 else [0, 1]
    for mode in modes:
        _capi.check(_capi.lib().kmpc_set_solver_param(_capi.Handle.get(0).ptr, 7, float(mode)))
        out = bt.run_backtest_batched(yhat, realized, n_steps=ns, horizon=H, want_history=True)
        torch.cuda.synchronize()
        res[mode] = (out["metrics"].cpu().numpy(), out["history"].cpu().numpy(), out["stats"].cpu().numpy())
        print("mode", mode, "stats", res[mode][2].sum(axis=0), "final", res[mode][0][:4, 3])
    if 0 in res and 1 in res: print("max |d metrics|", np.abs(res[0][0] - res[1][0]).max(), "max rel d value", np.abs(res[0][1][..., 0] / res[1][1][..., 0] - 1).max())
if __name__ == "__main__":
    main()
