"""A/B timing of tuning builds of the lane MPC kernel (development tool, not product code).

  python scripts/ab_bench.py build <name> "<-D flags>"     # here (no GPU): build_ab/libkmpc_<name>.so, (H,G) = (5,1) + (5,2) only
  python scripts/ab_bench.py run [steps]                    # on the GPU box: time every build_ab/*.so + the product lib
  python scripts/ab_bench.py one <lib.so> [steps]           # one library (child process of `run`)

Each run is the config-2 device step of bench.py (4096 backtests x 246 decisions): ms of the MPC + portfolio stage
(CUDA events), iterations per decision and status counts."""
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
AB = os.path.join(ROOT, "build_ab")


def one(lib, steps, params=()):
    import numpy as np
    import torch
    from koopman_mpc_portfolio_rebalancing_b200 import _capi
    _capi.LIB_PATH = lib
    for kv in params:                                # e.g. 6=3e-10  (KMPC_PARAM_TOL), chunk=9102, gemm=2
        k, v = kv.split("=")
        if k == "chunk":
            _capi.lib().kmpc_set_forecast_chunk_rows(int(v)); continue
        if k == "gemm":
            _capi.lib().kmpc_set_gemm_fp16_pairs(int(v)); continue
        _capi.check(_capi.lib().kmpc_set_solver_param(_capi.Handle.get(0).ptr, int(k), float(v)))
    from koopman_mpc_portfolio_rebalancing_b200 import backtest as bt, engine, model as km, synthetic
    import bench
    w = bench.WORKLOADS["cfg2"]
    B, N, d, H, Z, rows = w["B"], w["N"], w["d"], w["H"], w["Z"], w["rows"]
    m = km.make_model(km.model_config("GenericKM", Z, w["enc"], enc_bias=True), N * d)
    m.load_state_dict(synthetic.generic_km_weights(0, N * d, w["enc"], Z))
    eng = engine.BatchedBacktester(m, N, d, bt.MPCConfig(horizon=H), bt.BacktestConfig(horizon=H))
    lr, mean, std, T = bench.make_inputs(w, B, 10_000)
    lr_d, mean_d, std_d = torch.from_numpy(lr).cuda(), torch.from_numpy(mean).cuda(), torch.from_numpy(std).cuda()
    for _ in range(2):
        out = eng.run_device(lr_d, mean_d, std_d, 0, rows)
    torch.cuda.synchronize()
    ms = []; fc = []
    for _ in range(steps):
        tm = {}
        out = eng.run_device(lr_d, mean_d, std_d, 0, rows, timings=tm)
        torch.cuda.synchronize()
        ev = tm["_events"]
        ms.append(ev[2].elapsed_time(ev[3])); fc.append(ev[1].elapsed_time(ev[2]))
    st = out["stats"].sum(dim=0).cpu().numpy()
    met = out["metrics"].cpu().numpy()
    print(json.dumps({"lib": os.path.basename(lib), "params": list(params), "forecast_ms": float(np.median(fc)), "mpc_ms": float(np.median(ms)), "mpc_ms_all": [round(x, 2) for x in ms],
                      "iters": float(st[3]) / (B * (rows - 1 - H)), "optimal": int(st[0]), "inaccurate": int(st[1]),
                      "fallback": int(st[2]), "mean_final_value": float(met[:, 3].mean())}))


def main():
    mode = sys.argv[1]
    if mode == "build":
        from koopman_mpc_portfolio_rebalancing_b200 import build as kb
        os.makedirs(AB, exist_ok=True)
        name, defs = sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
        gdefs = sys.argv[4] if len(sys.argv) > 4 else ""          # -D flags for the tcgen05 GEMM units
        kb.build(verbose=False, lib=os.path.join(AB, f"libkmpc_{name}.so"), lane_variants=[(5, 1), (5, 2)], lane_defs=defs, gemm_defs=gdefs)
        kb.build(verbose=False)                     # restore the generated variant list of the product build
        print("built", name, defs)
    elif mode == "run":
        steps = sys.argv[2] if len(sys.argv) > 2 else "3"
        libs = sorted(glob.glob(os.path.join(AB, "*.so")))
        libs.append(os.path.join(ROOT, "koopman_mpc_portfolio_rebalancing_b200", "libkmpc.so"))
        for lib in libs:
            p = subprocess.run([sys.executable, os.path.abspath(__file__), "one", lib, steps], capture_output=True, text=True)
            line = [l for l in p.stdout.splitlines() if l.startswith("{")]
            print(line[-1] if line else f"{os.path.basename(lib)} FAILED: {p.stderr[-400:]}")
    elif mode == "one":
        one(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 3, sys.argv[4:])


if __name__ == "__main__":
    main()
