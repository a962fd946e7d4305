"""One timed pass of BASELINE config 3 (LISTAKM, 500 assets, H = 10), 4 or 5 (development tool); add cfg4 / cfg5 to choose.
  python scripts/cfg3_run.py [lib.so] [param=value ...]      # e.g. 7=0: KMPC_PARAM_CLUSTER off"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from koopman_mpc_portfolio_rebalancing_b200 import _capi
    args = sys.argv[1:]
    if args and args[0].endswith(".so"):
        _capi.LIB_PATH = args.pop(0)
    import bench
    for kv in [a for a in args if "=" in a]:          # e.g. 7=0  (KMPC_PARAM_ACTIVE_SET off)
        k, v = kv.split("=")
        _capi.check(_capi.lib().kmpc_set_solver_param(_capi.Handle.get(0).ptr, int(k), float(v)))
    dev = torch.device("cuda:0")
    prof = "prof" in args
    if prof:
        args.remove("prof")
    which = "cfg3"
    for a in list(args):
        if a in ("cfg3", "cfg4", "cfg5"):
            which = a; args.remove(a)
    if prof:                                          # CUPTI durations of the kernels of the pass (torch.profiler)
        from torch.profiler import profile, ProfilerActivity
        import collections
        with profile(activities=[ProfilerActivity.CUDA]) as pr:
            res = bench.other_configs(dev, 0, 1, torch.cuda.synchronize, which=(which,))
            torch.cuda.synchronize()
        agg = collections.defaultdict(list)
        for e in pr.events():
            if e.device_type == torch.autograd.DeviceType.CUDA:
                agg[e.name[:80]].append(e.time_range.end - e.time_range.start)
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1]))[:8]:
            print(f"{k:80s} n={len(v):4d} total={sum(v) / 1e3:10.2f} ms  each: {[round(x / 1e3, 1) for x in v[:6]]}")
    else:
        res = bench.other_configs(dev, 0, 1, torch.cuda.synchronize, which=(which,))
        torch.cuda.synchronize()
    print(json.dumps(res[which]))


if __name__ == "__main__":
    main()
