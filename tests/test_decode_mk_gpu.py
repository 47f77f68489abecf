"""The persistent decode kernel (csrc/decode_mk.cu: every step of decoder_joint.generate in one cooperative launch) against the
per-kernel CUDA-graph decode path it replaces (seq2seq_pretrain.py:450; x-transformers generate).

The two implementations share no GEMM / LayerNorm / sampling code, only the arithmetic contract, so they cross-check each other:
fp32-grade mode: logits within 2e-4 of each other at every step two greedy decodes share, identical codes wherever the top-2
margin exceeds the tolerance.  The oracle comparisons proper live in test_slmft_gpu.py (which now runs through this kernel in the
tensor-core modes)."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

import dim_b200  # noqa: E402
from dim_b200 import _lib  # noqa: E402
from dim_b200.schema import S2SConfig  # noqa: E402

S2S = S2SConfig()


@pytest.fixture(scope="module")
def engine_tc(slmft_sd):
    from dim_b200.engine import PREC_FP32_TC, Handle, SLMFTEngine
    h = Handle()
    h.register(slmft_sd)
    return SLMFTEngine(h, S2S, precision=PREC_FP32_TC)


@pytest.fixture(scope="module")
def engine_bf16(slmft_sd):
    from dim_b200.engine import PREC_BF16, Handle, SLMFTEngine
    h = Handle()
    h.register(slmft_sd)
    return SLMFTEngine(h, S2S, precision=PREC_BF16)


def _both(fn):
    """Run fn() with the persistent kernel, then with the per-kernel graph path."""
    try:
        _lib.decode_set_impl(0)
        a = fn()
        _lib.decode_set_impl(1)
        b = fn()
    finally:
        _lib.decode_set_impl(0)
    return a, b


def _check_against(codes, logits, ref_codes, ref_logits, tol):
    B = codes.shape[0]
    for b in range(B):
        neq = (codes[b] != ref_codes[b]).nonzero()
        t_end = codes.shape[1] if len(neq) == 0 else int(neq[0]) + 1
        d = float((logits[b, :t_end] - ref_logits[b, :t_end]).abs().max())
        assert d < tol, f"row {b}: logits differ by {d} before the first code difference (step {t_end - 1})"
        if len(neq):
            t = int(neq[0])
            top2 = torch.topk(ref_logits[b, t], 2).values
            assert float(top2[0] - top2[1]) < 2 * tol, f"row {b} diverged at step {t} with margin {float(top2[0] - top2[1])}"


@pytest.mark.parametrize("B,T,ragged", [(9, 24, True), (33, 33, True), (70, 14, True), (130, 10, True), (256, 8, True), (300, 6, True)])
def test_greedy_matches_graph_path_fp32_grade(engine_tc, B, T, ragged):
    s2s = engine_tc
    c = dim_b200.synth.make_clips(B, T, seed=300 + B, ragged=ragged)
    ctx = s2s.context(c["v_speaker"].cuda(), c["v_audio"].cuda(), c["mask"].cuda())
    m = c["mask"].cuda()
    prompt = torch.randint(0, 512, (B,), generator=torch.Generator().manual_seed(B)).cuda()
    (codes, logits), (rcodes, rlogits) = _both(lambda: s2s.generate(ctx, m, prompt, T - 1, return_logits=True))
    _check_against(codes.cpu(), logits.cpu(), rcodes.cpu(), rlogits.cpu(), 2e-4)


def test_sampling_matches_graph_path(engine_tc):
    """top-k(52) / softmax / inverse-CDF with supplied uniforms: the warp-per-row sampler of the persistent kernel takes the same
    decisions as the block sampler (rowops.cu) given the same logits; logits differ by accumulation order only, so the drawn
    sequences agree until a uniform lands within that noise of a CDF boundary."""
    s2s = engine_tc
    B, T = 40, 30
    c = dim_b200.synth.make_clips(B, T, seed=11, ragged=True)
    ctx = s2s.context(c["v_speaker"].cuda(), c["v_audio"].cuda(), c["mask"].cuda())
    m = c["mask"].cuda()
    prompt = torch.randint(0, 512, (B,), generator=torch.Generator().manual_seed(5)).cuda()
    u = torch.rand(B, T - 1, generator=torch.Generator().manual_seed(6)).cuda()
    k = math.ceil(0.1 * 512)
    (codes, logits), (rcodes, rlogits) = _both(
        lambda: s2s.generate(ctx, m, prompt, T - 1, temperature=1.0, uniforms=u, top_k=k, return_logits=True))
    codes, logits, rcodes, rlogits = codes.cpu(), logits.cpu(), rcodes.cpu(), rlogits.cpu()
    same_rows = 0
    for b in range(B):
        neq = (codes[b] != rcodes[b]).nonzero()
        t_end = T - 1 if len(neq) == 0 else int(neq[0]) + 1
        assert float((logits[b, :t_end] - rlogits[b, :t_end]).abs().max()) < 2e-4
        if len(neq) == 0:
            same_rows += 1
            continue
        # the first differing draw must sit at a CDF boundary of the reference distribution (within the logit noise)
        t = int(neq[0])
        lr = rlogits[b, t].double()
        kth = torch.topk(lr, k).values[-1]
        p = torch.where(lr >= kth, (lr - lr.max()).exp(), torch.zeros_like(lr))
        cdf = (p / p.sum()).cumsum(0)
        target = float(u[b, t])
        assert float((cdf - target).abs().min()) < 1e-3, f"row {b} step {t}: draw {target} is not near a CDF boundary"
    assert same_rows >= B - 4, same_rows
    assert len(codes.unique()) > 8


def test_rows_do_not_depend_on_the_batch(engine_tc):
    """No arithmetic crosses rows and the split-K factors depend on (N, K) only: a clip decoded in a batch of 9, of 14 or of 200 yields
    the same bits (this is what makes sharded == unsharded).  (<= 8 rows run the GEMV chain: another kernel family.)"""
    s2s = engine_tc
    B, T = 200, 9
    c = dim_b200.synth.make_clips(B, T, seed=21, ragged=True)
    ctx = s2s.context(c["v_speaker"].cuda(), c["v_audio"].cuda(), c["mask"].cuda())
    m = c["mask"].cuda()
    prompt = torch.randint(0, 512, (B,), generator=torch.Generator().manual_seed(7)).cuda()
    u = torch.rand(B, T - 1, generator=torch.Generator().manual_seed(8)).cuda()
    full, full_logits = s2s.generate(ctx, m, prompt, T - 1, temperature=1.0, uniforms=u, return_logits=True)
    for sl in (slice(0, 9), slice(120, 134), slice(187, 200)):       # > 8 rows: the persistent kernel on both sides
        part, part_logits = s2s.generate(ctx[sl].contiguous(), m[sl].contiguous(), prompt[sl].contiguous(), T - 1, temperature=1.0,
                                         uniforms=u[sl].contiguous(), return_logits=True)
        assert torch.equal(part, full[sl]) and torch.equal(part_logits, full_logits[sl])


def test_bf16_mode_matches_graph_path(engine_bf16):
    """bf16 operands + bf16 KV cache: both implementations round the same tensors to bf16 (GEMM operands, cache rows), so their
    first-step logits agree to accumulation order and whole sequences mostly coincide."""
    s2s = engine_bf16
    B, T = 64, 16
    c = dim_b200.synth.make_clips(B, T, seed=31, ragged=True)
    ctx = s2s.context(c["v_speaker"].cuda(), c["v_audio"].cuda(), c["mask"].cuda())
    m = c["mask"].cuda()
    prompt = torch.randint(0, 512, (B,), generator=torch.Generator().manual_seed(9)).cuda()
    (codes, logits), (rcodes, rlogits) = _both(lambda: s2s.generate(ctx, m, prompt, T - 1, return_logits=True))
    assert float((logits[:, 0] - rlogits[:, 0]).abs().max()) < 5e-2      # bf16 roundings of intermediates flip with accumulation order
    assert float((codes == rcodes).float().mean()) > 0.8


def test_bf16_mode_arbitrary_key_mask(engine_bf16):
    """Key-padding masks that are NOT prefix masks (holes) take the byte-mask path of the mma attention items; prefix masks take the
    per-clip length (mask_prefix_kernel).  Half of the clips get holes; both kinds must agree with the per-kernel graph path."""
    s2s = engine_bf16
    B, T = 48, 20
    c = dim_b200.synth.make_clips(B, T, seed=77, ragged=True)
    m = c["mask"].clone()
    g = torch.Generator().manual_seed(5)
    holes = torch.rand(B, T, generator=g) < 0.25
    holes[:, 0] = False
    holes[::2] = False                                   # even clips keep their prefix mask
    m = (m.bool() & ~holes).to(c["mask"].dtype)
    assert bool((m[1::2].bool() != c["mask"][1::2].bool()).any())
    mc = m.cuda()
    ctx = s2s.context(c["v_speaker"].cuda(), c["v_audio"].cuda(), mc)
    prompt = torch.randint(0, 512, (B,), generator=torch.Generator().manual_seed(10)).cuda()
    (codes, logits), (rcodes, rlogits) = _both(lambda: s2s.generate(ctx, mc, prompt, T - 1, return_logits=True))
    assert float((logits[:, 0] - rlogits[:, 0]).abs().max()) < 5e-2
    assert float((codes == rcodes).float().mean()) > 0.8
    # the mask matters: dropping the holes changes the first-step logits of the clips that had them
    (codes2, logits2) = s2s.generate(ctx, c["mask"].cuda(), prompt, T - 1, return_logits=True)
    assert float((logits2[1::2, 0] - logits[1::2, 0]).abs().max()) > 1e-3
    assert float((logits2[::2, 0] - logits[::2, 0]).abs().max()) == 0.0


def test_samples_share_the_context(engine_tc):
    s2s = engine_tc
    B, S, T = 11, 3, 12
    c = dim_b200.synth.make_clips(B, T, seed=41, ragged=True)
    ctx = s2s.context(c["v_speaker"].cuda(), c["v_audio"].cuda(), c["mask"].cuda())
    m = c["mask"].cuda()
    prompt = torch.randint(0, 512, (B,), generator=torch.Generator().manual_seed(10)).cuda()
    u = torch.rand(B, S, T - 1, generator=torch.Generator().manual_seed(11)).cuda()
    codes, logits = s2s.generate_samples(ctx, m, prompt, T - 1, S, u, return_logits=True)
    for j in range(S):
        cj, lj = s2s.generate(ctx, m, prompt, T - 1, temperature=1.0, uniforms=u[:, j].contiguous(), return_logits=True)
        assert torch.equal(codes[:, j], cj) and torch.equal(logits[:, j], lj)


def test_phase_trace(engine_tc):
    s2s = engine_tc
    B, T = 16, 10
    c = dim_b200.synth.make_clips(B, T, seed=51)
    ctx = s2s.context(c["v_speaker"].cuda(), c["v_audio"].cuda(), c["mask"].cuda())
    prompt = torch.zeros(B, dtype=torch.int64).cuda()
    _lib.decode_trace_enable(True)
    try:
        s2s.generate(ctx, c["mask"].cuda(), prompt, T - 1)
        tr = _lib.decode_trace_collect()
    finally:
        _lib.decode_trace_enable(False)
    assert len(tr) == 12 * S2S.depth + 2
    kinds = [k for k, _ in tr]
    assert kinds.count("gemm") == 6 * S2S.depth + 1 and kinds.count("attention") == 2 * S2S.depth
    assert all(ms > 0 for _, ms in tr)



def test_small_batch_gemv_phases_match_graph_path(slmft_sd):
    """<= 8 decode rows through the persistent kernel's GEMV phases (opt-in, DIM_SMALL_BATCH_MK=1: slower than the per-kernel chain
    today) against the per-kernel chain: same codes (or a near-tie at the first difference), logits within 2e-4.  Runs in a
    subprocess because the switch is read once per process."""
    import os
    import subprocess
    import sys
    code = r'''
import sys, torch
sys.path.insert(0, %r)
import dim_b200
from dim_b200.engine import PREC_FP32_TC, Handle, SLMFTEngine
from dim_b200.schema import S2SConfig
sd = dim_b200.synth.make_slmft_state_dict(131)
h = Handle(); h.register(sd)
e = SLMFTEngine(h, S2SConfig(), precision=PREC_FP32_TC)
c = dim_b200.synth.make_clips(3, 40, seed=9, ragged=True)
ctx = e.context(c["v_speaker"].cuda(), c["v_audio"].cuda(), c["mask"].cuda())
codes, logits = e.generate(ctx, c["mask"].cuda(), torch.tensor([3, 5, 7]).cuda(), 39, return_logits=True)
torch.save((codes.cpu(), logits.cpu()), sys.argv[1])
''' % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    import tempfile
    outs = []
    for env in ({}, {"DIM_SMALL_BATCH_MK": "1"}):
        with tempfile.NamedTemporaryFile(suffix=".pt") as f:
            r = subprocess.run([sys.executable, "-c", code, f.name], env=dict(os.environ, **env), capture_output=True, text=True, timeout=600)
            assert r.returncode == 0, r.stderr[-2000:]
            outs.append(torch.load(f.name))
    (c0, l0), (c1, l1) = outs
    for b in range(3):
        neq = (c0[b] != c1[b]).nonzero()
        n = len(c0[b]) if len(neq) == 0 else int(neq[0])
        upto = min(n + 1, len(c0[b]))
        assert float((l0[b, :upto] - l1[b, :upto]).abs().max()) < 2e-4
        if n < len(c0[b]):
            top2 = torch.topk(l0[b, n].double(), 2).values
            assert float(top2[0] - top2[1]) < 4e-4
