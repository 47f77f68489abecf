"""Oracle parity on BASELINE.json's own configurations (the transformer half is checked against the RESTATED oracle oracle/xt.py:
x-transformers 1.30.16 is not obtainable offline, "parity unpinned"; the VQ half of the oracle is pinned to the real reference).

  configs[1]  single-clip seq2seq generate, T=300, fp32, KV cache on: greedy and sampled-with-uniforms, both fp32-grade engines
              (FFMA and tcgen05 3-plane split)                                           seq2seq_pretrain.py:496-514
  configs[2]  batch=256 ViCo-shape clips at full size: rows 0, 127, 128, 255 of the batch (global batch index for the VQ
              decoder's positional-encoding quirk) against the oracle run on those four clips.

A divergence of the decoded sequence is never excused by its position: at the first differing token the test PROVES the tie --
greedy: the oracle's top-2 logit margin is below the logit tolerance; sampled: the uniform sits within the probability tolerance of
a CDF boundary of the oracle's distribution -- and the logits up to that step agree within tolerance.
"""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

import dim_b200  # noqa: E402
from dim_b200.schema import S2SConfig, VQConfig  # noqa: E402
from oracle import slmft as OS  # noqa: E402
from oracle import vqvae as OV  # noqa: E402
from oracle import xt as OX  # noqa: E402

S2S, VQ = S2SConfig(), VQConfig()
from parity_util import K, LOGIT_TOL, explain_first_difference  # noqa: E402


@pytest.fixture(scope="module", params=["fp32_ffma", "fp32_tcgen05"])
def engines(slmft_sd, request):
    from dim_b200.engine import PREC_FP32, PREC_FP32_TC, Handle, SLMFTEngine, VQEngine
    prec = PREC_FP32 if request.param == "fp32_ffma" else PREC_FP32_TC
    h = Handle()
    h.register(slmft_sd)
    return SLMFTEngine(h, S2S, precision=prec), VQEngine(h, VQ, prefix="listener_vq.", precision=prec)


@pytest.fixture(scope="module")
def config1_oracle(slmft_sd):
    """BASELINE configs[1] inputs: one clip, T=300, speaker motion = ones (ViCo loader behaviour), audio randn,
    listener randn * 0.3 via synth.make_clips; the oracle's greedy and sampled decodes with logits."""
    T = 300
    c = dim_b200.synth.make_clips(1, T, seed=2024, speaker="ones")
    u = torch.rand(1, T - 1, generator=torch.Generator().manual_seed(2025))
    z_l = OS.forward_vq_listener(slmft_sd, c["v_listener"], c["mask"], VQ)
    x_s = OS.forward_encoder(slmft_sd, c["v_speaker"], c["mask"], S2S)
    ctx = OS.decoder_context(slmft_sd, x_s, c["v_audio"])
    g_codes, g_logits = OX.generate(slmft_sd, "decoder_joint.net", z_l[:, 0:1], T - 1, S2S.depth, ctx, c["mask"], return_logits=True)
    s_codes, s_logits = OX.generate(slmft_sd, "decoder_joint.net", z_l[:, 0:1], T - 1, S2S.depth, ctx, c["mask"], temperature=1.0,
                                    uniforms=u, return_logits=True)
    return dict(c=c, u=u, z_l=z_l, ctx=ctx, greedy=(g_codes, g_logits), sampled=(s_codes, s_logits))


@pytest.mark.parametrize("mode", ["greedy", "sampled"])
def test_config1_single_clip_T300(engines, slmft_sd, config1_oracle, mode):
    from dim_b200.compat_api import listener_codes
    s2s, vq = engines
    o = config1_oracle
    c, T = o["c"], 300
    d = {k: c[k].cuda() for k in ("v_speaker", "v_listener", "v_audio", "mask")}
    z_l = listener_codes(vq, d["v_listener"], d["mask"])
    assert torch.equal(z_l.cpu(), o["z_l"])                                           # 300 VQ code indices bit-exact
    ctx = s2s.context(d["v_speaker"], d["v_audio"], d["mask"])
    assert float((ctx.cpu() - o["ctx"]).abs().max()) < 1e-4
    ref_codes, ref_logits = o[mode]
    if mode == "greedy":
        codes, logits = s2s.generate(ctx, d["mask"], z_l[:, 0], T - 1, return_logits=True)
        n = explain_first_difference(codes[0].cpu(), logits[0].cpu(), ref_codes[0], ref_logits[0])
    else:
        codes, logits = s2s.generate(ctx, d["mask"], z_l[:, 0], T - 1, temperature=1.0, uniforms=o["u"].cuda(), top_k=K, return_logits=True)
        n = explain_first_difference(codes[0].cpu(), logits[0].cpu(), ref_codes[0], ref_logits[0], o["u"][0])
    if n == T - 1:      # same code sequence: decoded FLAME coefficients within 1e-4 of the reference decode
        pred = vq.decode(codes=codes)
        ref_pred = OV.decode_indices(slmft_sd, ref_codes, VQ, None, prefix="listener_vq.")
        assert float((pred.cpu() - ref_pred).abs().max()) < 1e-4
    if mode == "sampled":
        assert len(codes.unique()) > 8                                                # sampling really explores the codebook


def test_config2_rows_of_the_full_batch(slmft_sd):
    """B=256 x T=300 through the fp32-grade tensor-core engine (the mode whose results meet the parity bar); rows 0, 127, 128
    and 255 -- both sides of the 128-row MMA tile boundary, first and last row -- against the oracle on those clips."""
    from dim_b200.compat_api import listener_codes
    from dim_b200.engine import PREC_FP32_TC, Handle, SLMFTEngine, VQEngine
    h = Handle()
    h.register(slmft_sd)
    s2s, vq = SLMFTEngine(h, S2S, precision=PREC_FP32_TC), VQEngine(h, VQ, prefix="listener_vq.", precision=PREC_FP32_TC)
    B, T = 256, 300
    rows = [0, 127, 128, 255]
    c = dim_b200.synth.make_clips(B, T, seed=4242, ragged=True)
    u = torch.rand(B, T - 1, generator=torch.Generator().manual_seed(4243))
    d = {k: c[k].cuda() for k in ("v_speaker", "v_listener", "v_audio", "mask")}
    bi = torch.arange(B, dtype=torch.int32).cuda()
    z_l = listener_codes(vq, d["v_listener"], d["mask"])
    ctx = s2s.context(d["v_speaker"], d["v_audio"], d["mask"])
    codes, logits = s2s.generate(ctx, d["mask"], z_l[:, 0], T - 1, temperature=1.0, uniforms=u.cuda(), top_k=K, return_logits=True)
    pred = vq.decode(codes=codes, batch_index=bi)
    sub = {k: c[k][rows] for k in ("v_speaker", "v_listener", "v_audio", "mask")}
    o_zl = OS.forward_vq_listener(slmft_sd, sub["v_listener"], sub["mask"], VQ)
    assert torch.equal(z_l[rows].cpu(), o_zl)
    o_ctx = OS.decoder_context(slmft_sd, OS.forward_encoder(slmft_sd, sub["v_speaker"], sub["mask"], S2S), sub["v_audio"])
    m = sub["mask"]
    assert float((ctx[rows].cpu()[m] - o_ctx[m]).abs().max()) < 1e-4
    o_codes, o_logits = OX.generate(slmft_sd, "decoder_joint.net", o_zl[:, 0:1], T - 1, S2S.depth, o_ctx, sub["mask"], temperature=1.0,
                                    uniforms=u[rows], return_logits=True)
    o_pred = OV.decode_indices(slmft_sd, o_codes, VQ, torch.tensor(rows), prefix="listener_vq.")
    for i, r in enumerate(rows):
        # only the steps inside the clip matter: frames past the clip's length are padding (masked out of the loss and metrics)
        n_valid = int(c["mask"][r].sum()) - 1
        n = explain_first_difference(codes[r, :n_valid].cpu(), logits[r, :n_valid].cpu(), o_codes[i, :n_valid], o_logits[i, :n_valid], u[r])
        if n == n_valid and torch.equal(codes[r].cpu(), o_codes[i]):
            assert float((pred[r].cpu() - o_pred[i]).abs().max()) < 1e-4, f"row {r}"
