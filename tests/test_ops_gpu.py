"""Operator-level parity: each C-ABI kernel against a plain PyTorch fp32 CPU evaluation of the same reference op."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from oracle import vqvae as OV  # noqa: E402
from oracle import xt as OX  # noqa: E402


def _g(seed):
    return torch.Generator().manual_seed(seed)


@pytest.mark.parametrize("M,N,K", [(300, 384, 56), (300, 1152, 384), (1, 2304, 1152), (3, 512, 1152), (8, 1152, 4608),
                                    (17, 768, 1152), (600, 56, 384), (4096, 1536, 384), (33, 128, 384)])
@pytest.mark.parametrize("act", [0, 1, 2, 3])
def test_linear(M, N, K, act):
    from dim_b200 import ops
    g = _g(M * 7 + N + K + act)
    a = torch.randn(M, K, generator=g)
    w = torch.randn(N, K, generator=g) / math.sqrt(K)
    b = torch.randn(N, generator=g) * 0.1
    r = torch.randn(M, N, generator=g)
    ref = F.linear(a, w, b)
    ref = [ref, F.leaky_relu(ref, 0.2), OV.gelu_tanh(ref), F.gelu(ref)][act] + r
    out = ops.linear(a.cuda(), w.cuda(), b.cuda(), r.cuda(), act=act, slope=0.2).cpu()
    # fp32 accumulation in a different order: tolerance 2e-5 abs on O(1) values (observed ~2e-6)
    assert torch.allclose(out, ref, atol=2e-5, rtol=1e-5), float((out - ref).abs().max())


def test_linear_no_bias_no_residual():
    from dim_b200 import ops
    g = _g(3)
    a, w = torch.randn(70, 384, generator=g), torch.randn(56, 384, generator=g) / 20
    out = ops.linear(a.cuda(), w.cuda()).cpu()
    assert torch.allclose(out, F.linear(a, w), atol=2e-5, rtol=1e-5)


@pytest.mark.parametrize("B,T", [(1, 300), (3, 64), (2, 5), (2, 1)])
def test_conv5_instnorm(B, T, vq_sd):
    from dim_b200 import ops
    g = _g(B * 100 + T)
    x = torch.randn(B, T, 384, generator=g)
    w, b = vq_sd["encoder.squasher.0.0.weight"], vq_sd["encoder.squasher.0.0.bias"]
    y_ref = F.leaky_relu(F.conv1d(F.pad(x.permute(0, 2, 1), (2, 2), mode="replicate"), w, b), 0.2)
    wr = ops.repack_conv_weight(w.cuda())
    assert torch.equal(wr.cpu(), w.permute(0, 2, 1).contiguous())
    y = ops.conv5_leaky(x.cuda(), wr, b.cuda(), 0.2)
    assert torch.allclose(y.cpu(), y_ref.permute(0, 2, 1), atol=3e-5, rtol=1e-5)
    if T > 1:
        n_ref = F.instance_norm(y_ref, eps=1e-5).permute(0, 2, 1)
        n = ops.instance_norm_(y.clone())
        assert torch.allclose(n.cpu(), n_ref, atol=5e-5, rtol=1e-4), float((n.cpu() - n_ref).abs().max())


def test_conv5_ragged_lens(vq_sd):
    """lens[b] < T must equal running the valid prefix alone (replicate padding at lens-1)."""
    from dim_b200 import ops
    g = _g(11)
    B, T = 3, 40
    lens = torch.tensor([40, 17, 5], dtype=torch.int32)
    x = torch.randn(B, T, 384, generator=g)
    w, b = vq_sd["decoder.expander.0.0.weight"], vq_sd["decoder.expander.0.0.bias"]
    wr = ops.repack_conv_weight(w.cuda())
    y = ops.conv5_leaky(x.cuda(), wr, b.cuda(), 0.2, lens=lens.cuda())
    n = ops.instance_norm_(y.clone(), lens=lens.cuda()).cpu()
    for i in range(B):
        L = int(lens[i])
        yr = F.leaky_relu(F.conv1d(F.pad(x[i:i + 1, :L].permute(0, 2, 1), (2, 2), mode="replicate"), w, b), 0.2)
        assert torch.allclose(y[i, :L].cpu(), yr[0].t(), atol=3e-5, rtol=1e-5)
        nr = F.instance_norm(yr, eps=1e-5)[0].t()
        assert torch.allclose(n[i, :L], nr, atol=5e-5, rtol=1e-4)


@pytest.mark.parametrize("rows,dim,bias", [(300, 384, True), (7, 1152, False), (1, 1152, False), (1000, 384, False)])
def test_layer_norm(rows, dim, bias):
    from dim_b200 import ops
    g = _g(rows + dim)
    x = torch.randn(rows, dim, generator=g) * 3 + 1
    gain = 1 + 0.1 * torch.randn(dim, generator=g)
    b = 0.1 * torch.randn(dim, generator=g) if bias else None
    ref = F.layer_norm(x, (dim,), gain, b, 1e-5)
    out = ops.layer_norm(x.cuda(), gain.cuda(), None if b is None else b.cuda()).cpu()
    assert torch.allclose(out, ref, atol=5e-6, rtol=1e-5), float((out - ref).abs().max())


@pytest.mark.parametrize("B,T,H,Dh,causal,masked", [(1, 300, 8, 48, False, False), (2, 64, 8, 48, False, False),
                                                    (2, 5, 8, 48, False, False), (1, 300, 12, 64, True, True),
                                                    (3, 130, 12, 64, True, True), (2, 64, 12, 64, False, True),
                                                    (1, 1, 12, 64, True, False)])
def test_attention(B, T, H, Dh, causal, masked):
    from dim_b200 import ops
    g = _g(B + T + H + Dh)
    inner = H * Dh
    qkv = torch.randn(B, T, 3 * inner, generator=g)
    scale = (inner ** -0.5) if Dh == 48 else Dh ** -0.5
    q, k, v = [t.view(B, T, H, Dh).permute(0, 2, 1, 3) for t in qkv.split(inner, dim=-1)]
    mask = None
    if masked:
        lens = torch.randint(max(1, T // 2), T + 1, (B,), generator=g)
        lens[0] = T
        mask = torch.arange(T)[None] < lens[:, None]
    am = ~torch.triu(torch.ones(T, T), diagonal=1).bool() if causal else None
    ref = OX.attend(q * (scale / Dh ** -0.5), k, v, key_mask=mask, attn_mask=am)   # attend() scales by Dh**-0.5
    out = ops.attention(qkv.cuda(), H, Dh, scale, key_mask=None if mask is None else mask.to(torch.uint8).cuda(),
                        causal=causal).cpu()
    assert torch.allclose(out, ref, atol=2e-5, rtol=1e-4), float((out - ref).abs().max())


@pytest.mark.parametrize("N", [1, 63, 300, 5000])
@pytest.mark.parametrize("scale", [0.5, 1.0 / 512])
def test_vq_argmin(N, scale):
    """Bit-exact indices vs the reference formula; tokens whose reference top-2 gap is within 4 ulp of d are
    tie-ambiguous (accumulation order decides) and are only required to pick one of the two."""
    from dim_b200 import ops
    g = _g(N)
    z = torch.randn(N, 128, generator=g) * 0.7
    E = torch.randn(512, 128, generator=g) * scale if scale > 0.01 else (torch.rand(512, 128, generator=g) * 2 - 1) * scale
    d = OV.distances(z, E)
    ref = torch.argmin(d, dim=1)
    d64 = (z.double()[:, None, :] - E.double()[None]).pow(2).sum(-1) if N <= 300 else None
    idx = ops.vq_argmin(z.cuda(), E.cuda()).cpu()
    top2 = torch.topk(d, 2, dim=1, largest=False)
    gap = top2.values[:, 1] - top2.values[:, 0]
    ulp = torch.finfo(torch.float32).eps * top2.values[:, 0].abs()
    ambiguous = gap <= 4 * ulp
    bad = (idx != ref) & ~ambiguous
    assert not bad.any(), f"{int(bad.sum())} non-ambiguous mismatches"
    mism = idx != ref
    if mism.any():   # ambiguous ones must still be the runner-up
        assert torch.equal(idx[mism], top2.indices[mism, 1])
    if d64 is not None and scale == 0.5:
        assert torch.equal(idx, d64.argmin(1))


def test_vq_argmin_ties_first_index():
    from dim_b200 import ops
    E = torch.randn(512, 128, generator=_g(5))
    E[300] = E[17]
    E[400] = E[17]
    z = E[[17, 300, 400, 5]].clone()
    idx = ops.vq_argmin(z.cuda(), E.cuda()).cpu()
    assert idx.tolist() == [17, 17, 17, 5]


@pytest.mark.parametrize("N", [1, 5, 299, 76544])
def test_vq_gather(N):
    from dim_b200 import ops
    g = _g(N)
    E = torch.randn(512, 128, generator=g)
    idx = torch.randint(0, 512, (N,), generator=g)
    out = ops.vq_gather(idx.cuda(), E.cuda()).cpu()
    assert torch.equal(out, E[idx])                                  # bit-exact rows
    if N <= 299:
        sd = {"quantize.embedding.weight": E}
        assert torch.equal(out, OV.codebook_entry(sd, idx))          # == the reference's one-hot matmul


def test_vq_gather_out_of_range_counted():
    from dim_b200 import ops
    E = torch.randn(512, 128, generator=_g(1))
    idx = torch.tensor([0, 511, 512, -100, 7])
    out, bad = ops.vq_gather(idx.cuda(), E.cuda(), count_bad=True)
    assert int(bad.item()) == 2
    assert torch.equal(out.cpu()[[0, 1, 4]], E[[0, 511, 7]])


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (256, 2304, 1152), (300, 384, 56), (100, 56, 384), (256, 1152, 4608),
                                    (1000, 1536, 384), (4096, 384, 1920), (257, 512, 1152)])
@pytest.mark.parametrize("planes", [1, 2, 3])
def test_linear_tcgen05(M, N, K, planes):
    """tcgen05 GEMM on bf16 planes.  planes=1 must equal an fp32 matmul of the bf16-rounded operands; planes=3 (6 products)
    must match the fp32 reference like the FFMA kernel does (2e-5); planes=2 (3 products) within 2e-4."""
    from dim_b200 import ops
    g = _g(M + N + K)
    a = torch.randn(M, K, generator=g)
    w = torch.randn(N, K, generator=g) / math.sqrt(K)
    b = torch.randn(N, generator=g) * 0.1
    r = torch.randn(M, N, generator=g)
    out = ops.linear_tc(a.cuda(), w.cuda(), b.cuda(), r.cuda(), act=2, planes=planes).cpu()
    if planes == 1:
        ref = OV.gelu_tanh(F.linear(a.bfloat16().float(), w.bfloat16().float(), b)) + r
        tol = 2e-5
    else:
        ref = OV.gelu_tanh(F.linear(a, w, b)) + r
        tol = 2e-5 if planes == 3 else 2e-4
    tol *= max(1.0, K / 1152)          # the tensor core's fp32 accumulator truncates: error grows ~linearly with K
    err = float((out - ref).abs().max())
    assert err <= tol, err


@pytest.mark.parametrize("M,N,K", [(4096, 1152, 768), (2500, 768, 1152), (2048, 1000, 384), (3000, 4608, 1152), (2048, 128, 384),
                                    (5000, 2304, 1152), (2304, 192, 64), (19072, 1152, 1152)])
@pytest.mark.parametrize("planes", [1, 3])
@pytest.mark.parametrize("act", [0, 3])
def test_linear_tcgen05_persistent(M, N, K, planes, act):
    """Large-M GEMMs run on the persistent kernel (gemm_bf16_tcgen05_persist: 128 x 256/192/128 tiles, two TMEM accumulators).  It
    must agree with the fp32 reference like the one-tile-per-CTA kernel does AND be bit-identical to it (same accumulation order
    over K, same epilogue arithmetic): forcing a tile width through the tuning hook selects the old kernel."""
    from dim_b200 import ops, _lib
    g = _g(M + N + K + planes)
    a = torch.randn(M, K, generator=g)
    w = torch.randn(N, K, generator=g) / math.sqrt(K)
    b = torch.randn(N, generator=g) * 0.1
    r = torch.randn(M, N, generator=g)
    ac, wc, bc, rc = a.cuda(), w.cuda(), b.cuda(), r.cuda()
    out = ops.linear_tc(ac, wc, bc, rc, act=act, planes=planes)
    lib = _lib.load()
    lib.dim_debug_tc_bn(128)
    try:
        old = ops.linear_tc(ac, wc, bc, rc, act=act, planes=planes)
    finally:
        lib.dim_debug_tc_bn(0)
    assert torch.equal(out, old)
    f = (lambda x: 0.5 * x * (1.0 + torch.erf(x * 0.7071067811865476))) if act == 3 else (lambda x: x)
    if planes == 1:
        ref = f(F.linear(a.bfloat16().float(), w.bfloat16().float(), b)) + r
    else:
        ref = f(F.linear(a, w, b)) + r
    tol = 2e-5 * max(1.0, K / 1152)
    err = float((out.cpu() - ref).abs().max())
    assert err <= tol, err


def test_split_planes_is_exact():
    from dim_b200 import ops
    x = torch.randn(37, 56, generator=_g(2)) * 3
    p = ops.split_planes(x.cuda(), 3).cpu().float().view(37, 3, 64)
    assert torch.equal(p[:, :, 56:], torch.zeros(37, 3, 8))
    assert torch.equal(p[:, 0, :56], x.bfloat16().float())
    rec = p[:, 0, :56] + p[:, 1, :56] + p[:, 2, :56]
    assert float((rec - x).abs().max()) <= 2e-7 * float(x.abs().max())


@pytest.mark.parametrize("case", ["vico_120x64", "short_7x8", "odd_333x16"])
def test_feature_resampling_matches_reference_golden(case):
    """SURVEY 8(f).3 -- dim_resample_features vs outputs of the reference's own functions (tests/golden/resample_reference.pt, made
    by tests/golden/make_resample_golden.py): the window mean (code/vico_preprocessing.py:7-19) is exact in fp32 for a window of 1
    and within fp32 rounding of the reference's float64 mean otherwise; the linear resampler (code/dataset/l2l.py:23-29, ATen
    align_corners=True index arithmetic restated in the kernel) within 1e-6 abs."""
    import os
    from dim_b200 import ops
    g = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "resample_reference.pt"), weights_only=False)
    c = g["cases"][case]
    x = c["x"].cuda()
    out = ops.resample_window_mean(x, 0.6).cpu()
    assert torch.equal(out.double(), c["window_mean_0.6"])                       # window of 1: a copy of the first 60 %
    out = ops.resample_window_mean(x, 0.25).cpu()
    assert out.shape == c["window_mean_0.25"].shape
    assert torch.allclose(out.double(), c["window_mean_0.25"], atol=1e-6)
    for n, ref in c["linear_new_t"].items():
        out = ops.resample_linear(x, n).cpu()
        assert out.shape == ref.shape and torch.allclose(out, ref, atol=1e-6), (n, float((out - ref).abs().max()))


def test_feature_resampling_full_size():
    """HuBERT-shape input (50 fps x 20 s, 768-d) -> 30 fps: properties at a size the goldens do not cover."""
    from dim_b200 import ops
    x = torch.randn(1000, 768, generator=_g(4)).cuda()
    y = ops.resample_linear(x, 600)
    assert torch.equal(y[0], x[0]) and torch.allclose(y[-1], x[-1], atol=1e-6)   # align_corners: end points map to end points
    assert torch.equal(ops.resample_linear(x, 1000), x)                          # identity when new_t == t
    ref = torch.nn.functional.interpolate(x.t()[None], size=600, mode="linear", align_corners=True)[0].t()
    assert torch.allclose(y, ref, atol=1e-6)
    assert torch.equal(ops.resample_window_mean(x, 0.6), x[:600])
