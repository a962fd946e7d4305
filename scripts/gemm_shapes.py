"""Kernel time of the fp16-pair tcgen05 GEMM (fp32 output path) for a list of shapes (development tool): CUPTI durations
through torch.profiler around kmpc_debug_gemm(mode 2).   python scripts/gemm_shapes.py M,N,K [M,N,K ...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from torch.profiler import profile, ProfilerActivity
    from koopman_mpc_portfolio_rebalancing_b200 import _capi
    h = _capi.Handle.get(0)
    data = "randn"
    args = sys.argv[1:]
    if args and "," not in args[0]:
        data = args.pop(0)                      # randn | relu (half of A zero) | zeros: tensor-core power depends on the operands
    shapes = [tuple(int(x) for x in a.split(",")) for a in args]
    for (M, N, K) in shapes:
        A = torch.randn(M, K, device="cuda"); W = torch.randn(N, K, device="cuda") / K ** 0.5
        if data == "relu":
            A = torch.relu(A)
        elif data == "zeros":
            A = torch.zeros_like(A)
        out = torch.empty(M, N, device="cuda")
        for _ in range(2):
            _capi.check(_capi.lib().kmpc_debug_gemm(h.ptr, _capi.ptr(A), _capi.ptr(W), M, N, K, _capi.ptr(out), 2))
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(3):
                _capi.check(_capi.lib().kmpc_debug_gemm(h.ptr, _capi.ptr(A), _capi.ptr(W), M, N, K, _capi.ptr(out), 2))
            torch.cuda.synchronize()
        d = [e.time_range.end - e.time_range.start for e in prof.events() if "gemm_tc16" in e.name]
        tiles = ((M + 127) // 128) * ((N + 127) // 128)
        waves = -(-tiles // 148)
        print(f"{data} M={M} N={N} K={K}: {min(d):8.1f} us  tiles={tiles} rounds={waves}  us/round={min(d) / waves:6.1f}")


if __name__ == "__main__":
    main()
