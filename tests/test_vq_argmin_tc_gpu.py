"""Tensor-core codebook argmin (csrc/vq_tc.cu: bf16 tcgen05 shortlist + exact fp32 re-rank) against the exact FFMA kernel
(csrc/vq.cu: vq_argmin_f32), which test_ops_gpu.py / test_vqvae_gpu.py pin to the reference formula and to goldens minted from
the real reference (models/lib/quantizer.py:38-45).  The two must return IDENTICAL indices on every input: the shortlist bound
(2 eps, vq_tc.cu header) makes the exact argmin a member of the shortlist, and the re-rank repeats the exact kernel's fp32
arithmetic operation for operation."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu

from dim_b200 import _lib, ops  # noqa: E402


def _g(seed):
    return torch.Generator().manual_seed(seed)


def _both(z, E):
    lib = _lib.load()
    lib.dim_debug_vq_argmin_impl.argtypes = [C.c_int]
    try:
        lib.dim_debug_vq_argmin_impl(1)
        exact = ops.vq_argmin(z, E)
        lib.dim_debug_vq_argmin_impl(0)
        tc = ops.vq_argmin(z, E)
    finally:
        lib.dim_debug_vq_argmin_impl(0)
    return exact, tc


def _stats(z, E):
    lib = _lib.load()
    fn = lib.dim_debug_vq_argmin_tc_stats
    fn.restype = C.c_int
    fn.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    idx = torch.empty(z.shape[0], dtype=torch.int64, device=z.device)
    st = torch.zeros(2, dtype=torch.int32, device=z.device)
    assert fn(z.data_ptr(), E.data_ptr(), idx.data_ptr(), z.shape[0], st.data_ptr(), torch.cuda.current_stream().cuda_stream) == 0
    multi, cand = st.tolist()
    return idx, multi, cand


@pytest.mark.parametrize("N", [512, 513, 5000, 76800, 1 << 20])
@pytest.mark.parametrize("kind", ["scaled", "default_init", "wide", "clustered"])
def test_identical_to_exact_kernel(N, kind):
    g = _g(N % 1000 + len(kind))
    z = torch.randn(N, 128, generator=g) * 0.7
    if kind == "scaled":                   # the benchmark codebook: N(0,1) * 0.5
        E = torch.randn(512, 128, generator=g) * 0.5
    elif kind == "default_init":           # nn.Embedding.uniform_(-1/512, 1/512): all codes ~ 0, distances nearly equal
        E = (torch.rand(512, 128, generator=g) * 2 - 1) / 512
    elif kind == "wide":                   # large dynamic range in both operands
        E = torch.randn(512, 128, generator=g) * torch.logspace(-3, 2, 512)[:, None]
        z = z * torch.logspace(-2, 2, N)[torch.randperm(N, generator=g)][:, None]
    else:                                  # tokens sitting next to codes, codes in tight clusters: many near-ties
        C0 = torch.randn(16, 128, generator=g)
        E = C0.repeat_interleave(32, 0) + torch.randn(512, 128, generator=g) * 1e-3
        z = E[torch.randint(0, 512, (N,), generator=g)] + torch.randn(N, 128, generator=g) * 1e-3
    exact, tc = _both(z.cuda(), E.cuda())
    assert torch.equal(exact, tc), f"{int((exact != tc).sum())} of {N} indices differ"


def test_ties_resolve_to_the_first_index():
    E = torch.randn(512, 128, generator=_g(5))
    E[300] = E[17]
    E[400] = E[17]
    z = torch.cat([E[[17, 300, 400, 5]], torch.randn(1020, 128, generator=_g(6))])
    exact, tc = _both(z.cuda(), E.cuda())
    assert tc[:4].tolist() == [17, 17, 17, 5] and torch.equal(exact, tc)


def test_shortlists_are_short():
    """The exact re-rank is the slow path: on the benchmark distribution most tokens must shortlist few codes."""
    g = _g(9)
    N = 200000
    z = (torch.randn(N, 128, generator=g) * 0.7).cuda()
    E = (torch.randn(512, 128, generator=g) * 0.5).cuda()
    idx, multi, cand = _stats(z, E)
    print(f"tokens with more than one shortlisted code: {multi}/{N}; shortlisted codes per token: {cand / N:.2f}")
    assert cand / N < 1.5
    assert torch.equal(idx, ops.vq_argmin(z, E))


def test_encode_goldens_unchanged(golden, vq_sd):
    """VQAutoEncoder.encode through the tensor-core argmin: the code indices of the goldens minted from the real reference
    stay bit-exact (the T=300 clip of BASELINE configs[0] repeated to reach the tensor-core kernel's minimum size)."""
    from dim_b200.engine import PREC_FP32_TC, Handle, VQEngine
    from dim_b200.schema import VQConfig
    h = Handle()
    h.register(vq_sd)
    vq = VQEngine(h, VQConfig(), precision=PREC_FP32_TC)
    case = golden["cases"]["c1_T300_B1"]
    g = torch.Generator().manual_seed(case["x_seed"])
    x = (torch.randn(case["B"], case["T"], 56, generator=g) * case["x_scale"]).cuda()
    B = 4                                                             # 1200 tokens >= 512: tensor-core argmin
    zero = torch.zeros(B, dtype=torch.int32).cuda()
    idx, _, _ = vq.encode(x.repeat(B, 1, 1), batch_index=zero)
    assert torch.equal(idx.cpu(), case["idx"].repeat(B, 1))
