"""One timed pass of BASELINE config 3 (LISTAKM, 500 assets, H = 10; development tool).
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
    for kv in args:                                   # e.g. 7=0  (KMPC_PARAM_CLUSTER off)
        k, v = kv.split("=")
        _capi.check(_capi.lib().kmpc_set_solver_param(_capi.Handle.get(0).ptr, int(k), float(v)))
    dev = torch.device("cuda:0")
    res = bench.other_configs(dev, 0, 1, torch.cuda.synchronize, which=("cfg3",))
    torch.cuda.synchronize()
    print(json.dumps(res["cfg3"]))


if __name__ == "__main__":
    main()
