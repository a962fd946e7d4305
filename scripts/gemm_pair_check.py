"""Correctness and speed of the CTA-pair fp16-pair GEMM (kmpc_set_gemm_fp16_pairs(2)) against the single-CTA kernel
and a float64 product.  Development tool: python scripts/gemm_pair_check.py"""
import ctypes as C
import os
import sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from koopman_mpc_portfolio_rebalancing_b200 import _capi

L = _capi.lib(); h = _capi.Handle.get(0)
rng = np.random.default_rng(0)
for (M, N, K) in [(256, 128, 64), (512, 256, 256), (1024, 1024, 1024), (640, 250, 1024), (32768, 1024, 1024)]:
    A = torch.from_numpy(rng.standard_normal((M, K)).astype(np.float32)).cuda()
    W = torch.from_numpy((rng.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32)).cuda()
    ref = (A.double() @ W.double().T)
    out = {}
    for mode in (1, 2):
        L.kmpc_set_gemm_fp16_pairs(mode)
        Cm = torch.zeros((M, N), dtype=torch.float32, device="cuda")
        _capi.check(L.kmpc_debug_gemm(h.ptr, _capi.ptr(A), _capi.ptr(W), M, N, K, _capi.ptr(Cm), 2))
        torch.cuda.synchronize()
        err = ((Cm.double() - ref).abs().amax(dim=1) / ref.abs().amax(dim=1)).max().item()
        out[mode] = (Cm, err)
    same = torch.equal(out[1][0], out[2][0])
    print(f"M={M} N={N} K={K}: row-wise rel err single {out[1][1]:.2e} pair {out[2][1]:.2e} bit-identical {same}", flush=True)
    assert out[2][1] < 5e-6
L.kmpc_set_gemm_fp16_pairs(1)
print("pair kernel ok")
