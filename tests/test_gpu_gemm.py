"""GEMM kernels of the forecast path against an fp64 reference on the device: the tcgen05 3xTF32 kernel must be
fp32-accurate (the forecast bar is 1e-5 norm-wise; a plain TF32 GEMM would sit at ~1e-3)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def run(A, W, mode):
    import torch
    from koopman_mpc_portfolio_rebalancing_b200 import _capi
    M, K = A.shape
    N = W.shape[0]
    out = torch.empty((M, N), dtype=torch.float32, device="cuda")
    h = _capi.Handle.get(0)
    _capi.check(_capi.lib().kmpc_debug_gemm(h.ptr, _capi.ptr(A), _capi.ptr(W), M, N, K, _capi.ptr(out), mode))
    return out


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (256, 128, 64), (384, 256, 1024), (1000, 1024, 1040), (4096, 1024, 1024),
                                   (130, 192, 72), (777, 320, 200)])
def test_tcgen05_gemm_is_fp32_accurate(M, N, K):
    import torch
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K)
    A = torch.randn((M, K), generator=g, device="cuda", dtype=torch.float32)
    W = torch.randn((N, K), generator=g, device="cuda", dtype=torch.float32) / K ** 0.5
    ref = (A.double() @ W.double().T)
    simt = run(A, W, 0)
    tc = run(A, W, 1)
    scale = ref.abs().max(dim=1, keepdim=True).values
    e_simt = ((simt.double() - ref).abs() / scale).max().item()
    e_tc = ((tc.double() - ref).abs() / scale).max().item()
    tc16 = run(A, W, 2)
    e_tc16 = ((tc16.double() - ref).abs() / scale).max().item()
    print(f"M={M} N={N} K={K}: row-wise rel err  simt {e_simt:.2e}  tcgen05-3xTF32 {e_tc:.2e}  tcgen05-fp16-pairs {e_tc16:.2e}")
    assert e_simt < 4e-6
    assert e_tc < 4e-6
    assert e_tc16 < 4e-6


def test_small_shapes_are_not_eligible_for_tcgen05():
    import torch
    from koopman_mpc_portfolio_rebalancing_b200 import _capi
    A = torch.randn((16, 64), device="cuda"); W = torch.randn((32, 64), device="cuda")
    with pytest.raises(_capi.KmpcError):
        run(A, W, 1)
    out = run(A, W, 0)
    assert torch.allclose(out, A @ W.T, atol=1e-4)


@pytest.mark.parametrize("M,N,K", [(512, 256, 256), (640, 250, 1024), (1000, 1024, 1040), (4096, 1024, 1024), (777, 320, 200)])
def test_cta_pair_variant_is_bit_identical(M, N, K):
    """kmpc_set_gemm_fp16_pairs(2): the same fp16-pair GEMM on CTA pairs (tcgen05 cta_group::2, 256 x 128 tiles, each CTA
    staging half of the weight tile, barriers across the pair).  Same MMA sequence per output row, so the result must be
    bit-identical to the single-CTA kernel — including ragged M (rows beyond the matrix in the second CTA of a pair) and a
    ragged N tile."""
    import torch
    from koopman_mpc_portfolio_rebalancing_b200 import _capi
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    A = torch.randn((M, K), generator=g, device="cuda", dtype=torch.float32)
    W = torch.randn((N, K), generator=g, device="cuda", dtype=torch.float32) / K ** 0.5
    single = run(A, W, 2)
    try:
        _capi.lib().kmpc_set_gemm_fp16_pairs(2)
        pair = run(A, W, 2)
        pair2 = run(A, W, 2)
    finally:
        _capi.lib().kmpc_set_gemm_fp16_pairs(1)
    assert torch.equal(single, pair) and torch.equal(pair, pair2)
