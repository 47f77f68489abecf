"""Decode-path kernels checked one by one against plain PyTorch fp32 evaluations of the same op:
the decode-attention kernel (every implementation / tile shape, bf16 and fp32 head-major caches, ragged key counts),
every VQ-gather implementation (bit-exact vs index_select) and the GEMV kernel of the single-clip path."""
import ctypes as C
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _attn_fn():
    from dim_b200 import _lib
    lib = _lib.load()
    fn = lib.dim_debug_attn_decode
    fn.restype = C.c_int
    fn.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
    return lib, fn


@pytest.mark.parametrize("impl", [0, 1, 2, 3, 4])
@pytest.mark.parametrize("bf16", [1, 0])
@pytest.mark.parametrize("B,Tk", [(2, 1), (3, 7), (2, 63), (2, 64), (5, 65), (3, 129), (2, 300), (1, 1024)])
def test_decode_attention_matches_torch(impl, bf16, B, Tk):
    """One query per (clip, head) over a head-major cache [B,H,Tk,64]: softmax(q.k * 0.125) @ v in fp32.
    Tolerance: fp32 accumulation order only (the cache values are exactly representable in their dtype): 2e-5 abs."""
    lib, fn = _attn_fn()
    H = 12
    g = torch.Generator().manual_seed(B * 1000 + Tk + impl)
    dt = torch.bfloat16 if bf16 else torch.float32
    k = torch.randn(B, H, Tk, 64, generator=g).to(dt)
    v = torch.randn(B, H, Tk, 64, generator=g).to(dt)
    q = torch.randn(B, H * 64, generator=g)
    att = torch.softmax(torch.einsum("bhd,bhkd->bhk", q.view(B, H, 64), k.float()) * 0.125, -1)
    ref = torch.einsum("bhk,bhkd->bhd", att, v.float()).reshape(B, H * 64)
    kc, vc, qc = k.cuda(), v.cuda(), q.cuda()
    out = torch.empty(B, H * 64, device="cuda")
    scratch = torch.empty(H * 64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    try:
        rc = fn(impl, kc.data_ptr(), vc.data_ptr(), qc.data_ptr(), out.data_ptr(), B, H, Tk, bf16, st)
        assert rc == 0, lib.dim_last_error()
        torch.cuda.synchronize()
    finally:                                            # leave the library on its default implementation
        fn(0, kc.data_ptr(), vc.data_ptr(), qc.data_ptr(), scratch.data_ptr(), 1, H, 1, bf16, st)
        torch.cuda.synchronize()
    err = float((out.cpu() - ref).abs().max())
    assert err < 2e-5, err


@pytest.mark.parametrize("mode", [0, 1, 2, 3, 4, 5, 6, -1])
@pytest.mark.parametrize("N", [1, 31, 256, 1000, 76544])
def test_vq_gather_every_implementation_is_bit_exact(mode, N):
    from dim_b200 import _lib, ops
    lib = _lib.load()
    g = torch.Generator().manual_seed(N + 17)
    E = torch.randn(512, 128, generator=g)
    idx = torch.randint(0, 512, (N,), generator=g)
    try:
        lib.dim_debug_vq_gather_mode(mode)
        out = ops.vq_gather(idx.cuda(), E.cuda())
        torch.cuda.synchronize()
    finally:
        lib.dim_debug_vq_gather_mode(-1)
    assert torch.equal(out.cpu(), E[idx])


@pytest.mark.parametrize("mode", [0, 2, 6])
def test_vq_gather_bad_indices_counted_in_every_mode(mode):
    from dim_b200 import _lib, ops
    lib = _lib.load()
    E = torch.randn(512, 128, generator=torch.Generator().manual_seed(2))
    idx = torch.randint(0, 512, (5000,), generator=torch.Generator().manual_seed(3))
    idx[7], idx[4321], idx[4999] = -1, 512, 99999
    try:
        lib.dim_debug_vq_gather_mode(mode)
        out, bad = ops.vq_gather(idx.cuda(), E.cuda(), count_bad=True)
        torch.cuda.synchronize()
    finally:
        lib.dim_debug_vq_gather_mode(-1)
    assert int(bad.item()) == 3
    assert torch.equal(out.cpu(), E[idx.clamp(0, 511)])


@pytest.mark.parametrize("M", [1, 2, 3, 4, 5, 8])
@pytest.mark.parametrize("N,K", [(2304, 1152), (1152, 768), (1152, 4608), (512, 1152), (384, 56), (4, 8)])
def test_gemv_small_batch(M, N, K):
    """The single-clip decode path (M <= 8): one CTA per output column, the weight row split over 128 lanes."""
    from dim_b200 import ops
    g = torch.Generator().manual_seed(M * 31 + N + K)
    a = torch.randn(M, K, generator=g)
    w = torch.randn(N, K, generator=g) / math.sqrt(K)
    b = torch.randn(N, generator=g) * 0.1
    r = torch.randn(M, N, generator=g)
    ref = F.gelu(F.linear(a, w, b)) + r
    out = ops.linear(a.cuda(), w.cuda(), b.cuda(), r.cuda(), act=3).cpu()
    assert torch.allclose(out, ref, atol=2e-5, rtol=1e-5), float((out - ref).abs().max())
