"""SLMFT (seq2seq_pretrain.py:431-514) on the GPU vs the restated oracle (x-transformers half: parity UNPINNED, see
oracle/xt.py).  Greedy decoding is compared token by token; at the first divergence, if any, the oracle's top-2 logit
margin must be below 1e-4 (tie-ambiguous), otherwise the test fails."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

import dim_b200  # noqa: E402
from dim_b200.schema import S2SConfig, VQConfig  # noqa: E402
from oracle import slmft as OS  # noqa: E402
from oracle import xt as OX  # noqa: E402
from parity_util import explain_first_difference  # noqa: E402

S2S, VQ = S2SConfig(), VQConfig()
TOL = 1e-4


@pytest.fixture(scope="module", params=["fp32_ffma", "fp32_tcgen05"])
def engines(slmft_sd, request):
    from dim_b200.engine import PREC_FP32, PREC_FP32_TC, Handle, SLMFTEngine, VQEngine
    prec = PREC_FP32 if request.param == "fp32_ffma" else PREC_FP32_TC
    h = Handle()
    h.register(slmft_sd)
    return SLMFTEngine(h, S2S, precision=prec), VQEngine(h, VQ, prefix="listener_vq.", precision=prec)


@pytest.mark.parametrize("B,T,ragged", [(1, 40, False), (3, 70, True), (2, 130, True)])
def test_context(engines, slmft_sd, B, T, ragged):
    s2s, _ = engines
    c = dim_b200.synth.make_clips(B, T, seed=B + T, ragged=ragged)
    ctx = s2s.context(c["v_speaker"].cuda(), c["v_audio"].cuda(), c["mask"].cuda()).cpu()
    x_s = OS.forward_encoder(slmft_sd, c["v_speaker"], c["mask"], S2S)
    ref = OS.decoder_context(slmft_sd, x_s, c["v_audio"])
    m = c["mask"]
    # padded query rows are "discarded downstream" (masked as cross-attention keys); compare valid rows
    assert torch.allclose(ctx[m], ref[m], atol=TOL), float((ctx[m] - ref[m]).abs().max())


def _compare_codes(codes, logits, ref_codes, ref_logits):
    B, S = ref_codes.shape
    for b in range(B):
        neq = (codes[b] != ref_codes[b]).nonzero()
        if len(neq) == 0:
            assert torch.allclose(logits[b], ref_logits[b], atol=5e-4), float((logits[b] - ref_logits[b]).abs().max())
            continue
        t = int(neq[0])
        assert torch.allclose(logits[b, :t + 1], ref_logits[b, :t + 1], atol=5e-4)
        top2 = torch.topk(ref_logits[b, t], 2).values
        assert float(top2[0] - top2[1]) < 1e-4, f"sample {b} diverged at step {t} with margin {float(top2[0]-top2[1])}"


@pytest.mark.parametrize("B,T,ragged", [(1, 48, False), (3, 40, True), (9, 24, False)])
def test_generate_greedy(engines, slmft_sd, B, T, ragged):
    s2s, _ = engines
    c = dim_b200.synth.make_clips(B, T, seed=100 + B, ragged=ragged)
    x_s = OS.forward_encoder(slmft_sd, c["v_speaker"], c["mask"], S2S)
    ctx = OS.decoder_context(slmft_sd, x_s, c["v_audio"])
    prompt = torch.randint(0, 512, (B, 1), generator=torch.Generator().manual_seed(B))
    ref_codes, ref_logits = OX.generate(slmft_sd, "decoder_joint.net", prompt, T - 1, S2S.depth, ctx, c["mask"],
                                        return_logits=True)
    codes, logits = s2s.generate(ctx.cuda(), c["mask"].cuda(), prompt.cuda(), T - 1, return_logits=True)
    _compare_codes(codes.cpu(), logits.cpu(), ref_codes, ref_logits)


def test_generate_sampled_with_uniforms(engines, slmft_sd):
    """temperature 1, top-k 52 (ceil(0.1*512)), draws from supplied uniforms."""
    s2s, _ = engines
    B, T = 2, 36
    c = dim_b200.synth.make_clips(B, T, seed=5)
    x_s = OS.forward_encoder(slmft_sd, c["v_speaker"], c["mask"], S2S)
    ctx = OS.decoder_context(slmft_sd, x_s, c["v_audio"])
    prompt = torch.tensor([[3], [400]])
    u = torch.rand(B, T - 1, generator=torch.Generator().manual_seed(77))
    ref, ref_logits = OX.generate(slmft_sd, "decoder_joint.net", prompt, T - 1, S2S.depth, ctx, c["mask"], temperature=1.0, uniforms=u,
                                  return_logits=True)
    out, logits = s2s.generate(ctx.cuda(), c["mask"].cuda(), prompt.cuda(), T - 1, temperature=1.0, uniforms=u.cuda(),
                               top_k=math.ceil(0.1 * 512), return_logits=True)
    out, logits = out.cpu(), logits.cpu()
    # a draw may legitimately flip only when it sits at a CDF boundary of the oracle's distribution (within the logit tolerance):
    # explain_first_difference proves that, and that the logits agree up to the flip
    for b in range(B):
        explain_first_difference(out[b], logits[b], ref[b], ref_logits[b], u[b])
    assert len(out.unique()) > 8                                    # sampling really explores the codebook


def test_forward_val_end_to_end(engines, slmft_sd):
    """SLMFT.forward(mode='val') = listener VQ encode -> encoders -> generate -> gather -> VQ decode."""
    from dim_b200.compat_api import slmft_forward_val
    s2s, vq = engines
    B, T = 2, 32
    c = dim_b200.synth.make_clips(B, T, seed=8, ragged=True)
    ref_loss, _, ref_pred, inter = OS.forward_val(slmft_sd, c["v_speaker"], c["v_listener"], c["v_audio"], c["mask"], S2S, VQ,
                                                  return_intermediates=True)
    loss, d, pred, codes = slmft_forward_val(s2s, vq, c["v_speaker"].cuda(), c["v_listener"].cuda(), c["v_audio"].cuda(),
                                             c["mask"].cuda(), return_codes=True)
    assert torch.equal(codes.cpu(), inter["codes"])
    assert torch.allclose(pred.cpu(), ref_pred, atol=TOL), float((pred.cpu() - ref_pred).abs().max())
    assert abs(float(loss) - float(ref_loss)) < 1e-4


def test_generate_greedy_batch96_hits_tensor_core_decode(engines, slmft_sd):
    """B = 96 >= 64 rows: in the tcgen05 mode the per-step decode GEMMs run on the tensor cores too."""
    s2s, _ = engines
    B, T = 96, 12
    c = dim_b200.synth.make_clips(B, T, seed=55)
    x_s = OS.forward_encoder(slmft_sd, c["v_speaker"], c["mask"], S2S)
    ctx = OS.decoder_context(slmft_sd, x_s, c["v_audio"])
    prompt = torch.randint(0, 512, (B, 1), generator=torch.Generator().manual_seed(1))
    ref_codes, ref_logits = OX.generate(slmft_sd, "decoder_joint.net", prompt, T - 1, S2S.depth, ctx, c["mask"],
                                        return_logits=True)
    codes, logits = s2s.generate(ctx.cuda(), c["mask"].cuda(), prompt.cuda(), T - 1, return_logits=True)
    _compare_codes(codes.cpu(), logits.cpu(), ref_codes, ref_logits)


def test_bf16_mode_is_close(slmft_sd):
    """bf16 GEMM operands (BASELINE configs[2]): not a parity mode; logits of the first decode step must stay within
    bf16 rounding noise of the fp32 oracle and the path must run end to end."""
    from dim_b200.engine import PREC_BF16, Handle, SLMFTEngine
    h = Handle()
    h.register(slmft_sd)
    s2s = SLMFTEngine(h, S2S, precision=PREC_BF16)
    B, T = 64, 20
    c = dim_b200.synth.make_clips(B, T, seed=56)
    ctx_ref = OS.decoder_context(slmft_sd, OS.forward_encoder(slmft_sd, c["v_speaker"], c["mask"], S2S), c["v_audio"])
    ctx = s2s.context(c["v_speaker"].cuda(), c["v_audio"].cuda(), c["mask"].cuda()).cpu()
    assert float((ctx - ctx_ref).abs().max()) < 0.15
    prompt = torch.randint(0, 512, (B, 1), generator=torch.Generator().manual_seed(2))
    ref_codes, ref_logits = OX.generate(slmft_sd, "decoder_joint.net", prompt, 1, S2S.depth, ctx_ref, c["mask"], return_logits=True)
    codes, logits = s2s.generate(ctx_ref.cuda(), c["mask"].cuda(), prompt.cuda(), T - 1, return_logits=True)
    assert float((logits[:, 0].cpu() - ref_logits[:, 0]).abs().max()) < 0.1
    assert codes.shape == (B, T - 1) and int(codes.min()) >= 0 and int(codes.max()) < 512


def test_concurrent_group_decoding_is_bit_identical(slmft_sd):
    """B = 256 is decoded as two concurrent 128-clip groups on side streams (CUDA graphs); every clip's codes and logits
    must equal those of the same clip decoded in a small batch (no cross-row arithmetic; split-K depends on K only)."""
    from dim_b200.engine import PREC_FP32_TC, Handle, SLMFTEngine
    h = Handle()
    h.register(slmft_sd)
    s2s = SLMFTEngine(h, S2S, precision=PREC_FP32_TC)
    B, T = 256, 10
    c = dim_b200.synth.make_clips(B, T, seed=77, ragged=True)
    ctx = s2s.context(c["v_speaker"].cuda(), c["v_audio"].cuda(), c["mask"].cuda())
    prompt = torch.randint(0, 512, (B,), generator=torch.Generator().manual_seed(3)).cuda()
    u = torch.rand(B, T - 1, generator=torch.Generator().manual_seed(4)).cuda()
    m = c["mask"].cuda()
    full, full_logits = s2s.generate(ctx, m, prompt, T - 1, temperature=1.0, uniforms=u, return_logits=True)
    again = s2s.generate(ctx, m, prompt, T - 1, temperature=1.0, uniforms=u)          # graph replay path
    assert torch.equal(full, again)
    sl = slice(120, 200)                                                               # straddles the group boundary
    part, part_logits = s2s.generate(ctx[sl].contiguous(), m[sl].contiguous(), prompt[sl].contiguous(), T - 1, temperature=1.0,
                                     uniforms=u[sl].contiguous(), return_logits=True)
    assert torch.equal(part, full[sl]) and torch.equal(part_logits, full_logits[sl])
    assert len(full.unique()) > 8


def test_host_buffer_entry_point_equals_device_call(engines):
    """compat_api.slmft_forward_val_host (pinned host inputs, H2D copies on a side stream overlapping the VQ encode, frames copied
    back to pinned memory) must return exactly what slmft_forward_val returns on resident inputs."""
    from dim_b200.compat_api import slmft_forward_val, slmft_forward_val_host
    s2s, vq = engines
    B, T = 3, 24
    c = dim_b200.synth.make_clips(B, T, seed=12, ragged=True)
    u = torch.rand(B, T - 1, generator=torch.Generator().manual_seed(5)).cuda()
    host = {k: c[k].pin_memory() for k in ("v_speaker", "v_listener", "v_audio", "mask")}
    out_host = torch.empty(B, T - 1, 56).pin_memory()
    for _ in range(2):                                   # twice: the second call reuses the side stream and buffers
        l1, _, p1, c1 = slmft_forward_val_host(s2s, vq, host, torch.device("cuda", torch.cuda.current_device()), uniforms=u,
                                               out_host=out_host)
    torch.cuda.synchronize()
    l2, _, p2, c2 = slmft_forward_val(s2s, vq, c["v_speaker"].cuda(), c["v_listener"].cuda(), c["v_audio"].cuda(), c["mask"].cuda(),
                                      uniforms=u, return_codes=True)
    assert torch.equal(c1, c2) and torch.equal(p1, p2) and torch.equal(out_host, p2.cpu())
    assert float(l1) == float(l2)


@pytest.mark.parametrize("B,S", [(70, 2), (3, 4)])
def test_generate_samples_equals_separate_calls(engines, slmft_sd, B, S):
    """`samples` draws per clip over ONE context projection (shared cross-attention K/V) must equal `samples`
    separate generate calls with the same uniforms.  Bit for bit whenever both run the same GEMM kernel family (tensor-core
    engine with >= 64 rows on both sides: no arithmetic crosses rows and the split-K factor depends on the call site only);
    when the row counts straddle a kernel-family boundary (FFMA GEMV for <= 8 rows, tiled FFMA / tcgen05 above) the
    accumulation order differs and the logits must agree to 1e-4."""
    from dim_b200.engine import PREC_FP32_TC
    s2s, _ = engines
    T = 12
    c = dim_b200.synth.make_clips(B, T, seed=31, ragged=True)
    ctx = s2s.context(c["v_speaker"].cuda(), c["v_audio"].cuda(), c["mask"].cuda())
    m = c["mask"].cuda()
    prompt = torch.randint(0, 512, (B,), generator=torch.Generator().manual_seed(8)).cuda()
    u = torch.rand(B, S, T - 1, generator=torch.Generator().manual_seed(9)).cuda()
    codes, logits = s2s.generate_samples(ctx, m, prompt, T - 1, S, u, return_logits=True)
    again = s2s.generate_samples(ctx, m, prompt, T - 1, S, u)                       # graph replay path
    assert torch.equal(codes, again)
    exact = s2s.precision == PREC_FP32_TC and B >= 64
    for j in range(S):
        cj, lj = s2s.generate(ctx, m, prompt, T - 1, temperature=1.0, uniforms=u[:, j].contiguous(), return_logits=True)
        if exact:
            assert torch.equal(codes[:, j], cj), f"sample {j}: codes differ"
            assert torch.equal(logits[:, j], lj), f"sample {j}: logits differ"
        else:
            assert torch.allclose(logits[:, j, 0], lj[:, 0], atol=1e-4)             # first step: same inputs on both sides
            agree = float((codes[:, j] == cj).float().mean())
            assert agree > 0.9, agree                                               # a draw at a CDF boundary may flip a tail
    assert len(codes.unique()) > 8


def test_forward_val_samples_equals_forward_val(engines):
    from dim_b200.compat_api import best_of_n, frechet_distance_torch, slmft_forward_val, slmft_forward_val_samples
    s2s, vq = engines
    B, T, S = 2, 24, 3
    c = dim_b200.synth.make_clips(B, T, seed=41, ragged=True)
    d = {k: c[k].cuda() for k in ("v_speaker", "v_listener", "v_audio", "mask")}
    u = torch.rand(B, S, T - 1, generator=torch.Generator().manual_seed(2)).cuda()
    pred, codes = slmft_forward_val_samples(s2s, vq, d["v_speaker"], d["v_listener"], d["v_audio"], d["mask"], S, uniforms=u)
    assert pred.shape == (B, S, T - 1, 56)
    for j in range(S):
        _, _, pj, cj = slmft_forward_val(s2s, vq, d["v_speaker"], d["v_listener"], d["v_audio"], d["mask"],
                                         uniforms=u[:, j].contiguous(), return_codes=True)
        # codes: same GEMV kernel on both sides -> identical; frames: the VQ decode of 3x the rows may cross the FFMA / tensor-core
        # boundary (64 rows), so the decoded coefficients agree to fp32 accumulation order
        assert torch.equal(codes[:, j], cj) and torch.allclose(pred[:, j], pj, atol=2e-5)
    # device-side best-of-N == the reference's host-side selection (numpy cov + scipy sqrtm, x_engine_pt.py:260-268)
    import numpy as np
    from scipy import linalg
    lengths = [int(v) - 1 for v in c["lengths"]]
    picked, chosen, fd = best_of_n(pred, d["v_listener"][:, 1:], lengths)
    for b in range(B):
        n = lengths[b]
        t = c["v_listener"][b, 1:n + 1].numpy().astype(np.float64)
        ref = []
        for j in range(S):
            p = pred[b, j, :n].cpu().numpy().astype(np.float64)
            mu1, s1, mu2, s2 = t.mean(0), np.cov(t, rowvar=False), p.mean(0), np.cov(p, rowvar=False)
            cm = linalg.sqrtm(s1.dot(s2))
            cm = cm.real if np.iscomplexobj(cm) else cm
            ref.append(float((mu1 - mu2).dot(mu1 - mu2) + np.trace(s1) + np.trace(s2) - 2 * np.trace(cm)))
        assert int(np.argmin(ref)) == int(chosen[b])
        # n < 56 frames: rank-deficient covariances, where scipy's sqrtm itself carries ~1e-4 of round-off
        assert np.allclose(fd[b].cpu().numpy(), ref, rtol=1e-3, atol=1e-3), (fd[b], ref)
        assert torch.equal(picked[b], pred[b, int(chosen[b]), :n])


@pytest.mark.parametrize("B,T,ragged,with_kv_mask", [(2, 30, True, True), (3, 70, True, False), (70, 12, False, True)])
def test_teacher_forced_forward(engines, slmft_sd, B, T, ragged, with_kv_mask):
    """SURVEY 8(f).2 -- SLMFT.forward(mode='train') forward pass (seq2seq_pretrain.py:447-448, x-transformers
    AutoregressiveWrapper.forward with the random self_attn_kv_mask supplied explicitly): logits within 5e-4 of the restated
    oracle on the valid positions, cross-entropy within 1e-4, argmax codes equal wherever the oracle's top-2 margin > 1e-3."""
    import torch.nn.functional as F
    from dim_b200.compat_api import draw_kv_mask, slmft_forward_train
    s2s, vq = engines
    c = dim_b200.synth.make_clips(B, T, seed=200 + B, ragged=ragged)
    kv = draw_kv_mask((B, T - 1), 0.15, "cpu", generator=torch.Generator().manual_seed(3)) if with_kv_mask else None
    total, d, pred, logits = slmft_forward_train(s2s, vq, c["v_speaker"].cuda(), c["v_listener"].cuda(), c["v_audio"].cuda(),
                                                 c["mask"].cuda(), kv_mask=None if kv is None else kv.cuda(), mask_prob=0.0,
                                                 return_logits=True)
    z_l = OS.forward_vq_listener(slmft_sd, c["v_listener"], c["mask"], VQ)
    x_s = OS.forward_encoder(slmft_sd, c["v_speaker"], c["mask"], S2S)
    ctx = OS.decoder_context(slmft_sd, x_s, c["v_audio"])
    ref_loss, ref_logits = OX.teacher_forced(slmft_sd, "decoder_joint.net", z_l, S2S.depth, ctx, c["mask"], kv_mask=kv)
    valid = z_l[:, 1:] != -100                                           # positions that enter the loss
    diff = (logits.cpu() - ref_logits)[valid].abs().max()
    assert float(diff) < 5e-4, float(diff)
    assert abs(float(d["l_ce_l"]) - float(ref_loss)) < 1e-4
    top2 = torch.topk(ref_logits, 2, dim=-1).values
    clear = valid & ((top2[..., 0] - top2[..., 1]) > 1e-3)
    assert torch.equal(logits.argmax(-1).cpu()[clear], ref_logits.argmax(-1)[clear])
    assert pred.shape == (B, T - 1, 56) and torch.isfinite(pred).all()
    assert abs(float(total) - float(d["l_ce_l"]) - float(d["l_cont_l"])) < 1e-6


def test_draw_kv_mask_follows_upstream_recipe():
    from dim_b200.compat_api import draw_kv_mask
    m = draw_kv_mask((5, 40), 0.15, "cuda")
    assert m.shape == (5, 40) and m.dtype == torch.bool
    assert bool(m[:, 0].all())                                          # the first key is never masked
    assert (~m).sum(1).tolist() == [min(int(41 * 0.15), 40)] * 5     # T = 41 before the shift


def test_long_clip_T1024_lm_listener_shape(engines, slmft_sd):
    """BASELINE configs[4] shape (LM-Listener chunks are capped at 1024 frames, dataset/data_loader.py:215-227): one 1024-frame
    clip through the whole val forward, greedy, against the oracle -- codes token for token (or an oracle near-tie at the first
    difference), decoded coefficients within 1e-4 where the codes agree."""
    from dim_b200.compat_api import slmft_forward_val
    s2s, vq = engines
    B, T = 1, 1024
    c = dim_b200.synth.make_clips(B, T, seed=77)
    ref_loss, _, ref_pred, inter = OS.forward_val(slmft_sd, c["v_speaker"], c["v_listener"], c["v_audio"], c["mask"], S2S, VQ,
                                                  return_intermediates=True)
    loss, d, pred, codes = slmft_forward_val(s2s, vq, c["v_speaker"].cuda(), c["v_listener"].cuda(), c["v_audio"].cuda(),
                                             c["mask"].cuda(), return_codes=True)
    z_l = OS.forward_vq_listener(slmft_sd, c["v_listener"], c["mask"], VQ)
    from dim_b200.compat_api import listener_codes
    assert torch.equal(listener_codes(vq, c["v_listener"].cuda(), c["mask"].cuda()).cpu(), z_l)      # 1024 VQ codes bit-exact
    got, ref = codes.cpu()[0], inter["codes"][0]
    if torch.equal(got, ref):
        assert torch.allclose(pred.cpu(), ref_pred, atol=TOL), float((pred.cpu() - ref_pred).abs().max())
    else:
        # 1023 dependent steps: a difference is accepted only as a proven oracle near-tie (top-2 margin below the logit tolerance,
        # logits equal within tolerance up to that step) -- never because it happens late
        ctx = s2s.context(c["v_speaker"].cuda(), c["v_audio"].cuda(), c["mask"].cuda())
        codes2, logits = s2s.generate(ctx, c["mask"].cuda(), z_l[:, 0].cuda(), T - 1, return_logits=True)
        assert torch.equal(codes2.cpu()[0], got)
        _, ref_logits = OX.generate(slmft_sd, "decoder_joint.net", z_l[:, 0:1], T - 1, S2S.depth, inter["ctx"], c["mask"], return_logits=True)
        explain_first_difference(got, logits[0].cpu(), ref, ref_logits[0])
