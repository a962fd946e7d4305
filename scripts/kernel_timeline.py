"""Per-kernel durations and stream gaps of one config-2 device step (development tool): CUPTI activity records through
torch.profiler, i.e. the kernels run concurrently / back to back as in a normal run (ncu serialises and flushes).

  python scripts/kernel_timeline.py [lib.so] [k=v ...]      # same parameter syntax as ab_bench.py one
"""
import collections
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import numpy as np
    import torch
    from torch.profiler import profile, ProfilerActivity
    from koopman_mpc_portfolio_rebalancing_b200 import _capi
    args = sys.argv[1:]
    if args and args[0].endswith(".so"):
        _capi.LIB_PATH = args.pop(0)
    for kv in args:
        k, v = kv.split("=")
        if k == "chunk":
            _capi.lib().kmpc_set_forecast_chunk_rows(int(v))
        elif k == "gemm":
            _capi.lib().kmpc_set_gemm_fp16_pairs(int(v))
        elif k == "embed":
            _capi.lib().kmpc_set_forecast_embedding(int(v))
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
        eng.run_device(lr_d, mean_d, std_d, 0, rows)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        eng.run_device(lr_d, mean_d, std_d, 0, rows)
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    agg = collections.defaultdict(list)
    for e in evs:
        agg[e.name[:70]].append(e.time_range.end - e.time_range.start)
    print("kernel                                                                  n     total_ms   mean_us   min_us   max_us")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f"{k:70s} {len(v):5d} {sum(v) / 1e3:10.3f} {np.mean(v):9.1f} {min(v):8.1f} {max(v):8.1f}")
    # stream gaps inside the forecast stage: from the first to the last gemm_tc16 launch
    g = [i for i, e in enumerate(evs) if "gemm_tc16" in e.name]
    if g:
        seg = evs[g[0]:g[-1] + 1]
        span = seg[-1].time_range.end - seg[0].time_range.start
        busy = sum(e.time_range.end - e.time_range.start for e in seg)
        gaps = [seg[i + 1].time_range.start - seg[i].time_range.end for i in range(len(seg) - 1)]
        print(json.dumps({"forecast_span_ms": span / 1e3, "kernel_busy_ms": busy / 1e3, "gap_total_ms": sum(x for x in gaps if x > 0) / 1e3,
                          "n_kernels": len(seg), "max_gap_us": max(gaps)}))
        # the first chunk's launches in order
        for e in seg[:44]:
            print(f"  {e.time_range.start - seg[0].time_range.start:10.1f} us  +{e.time_range.end - e.time_range.start:8.1f} us  {e.name[:60]}")


if __name__ == "__main__":
    main()
