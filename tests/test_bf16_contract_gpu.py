"""What DIM_PREC_BF16 (the headline mode of bench.py: BASELINE.json configs[2], "bf16 fused transformer + VQ decode") promises.

bf16 is not a parity mode -- GEMM operands, the K/V caches and the attention's q / p operands are rounded to bf16, accumulation,
softmax, LayerNorm and the residual stream stay fp32 -- so instead of equality with the oracle it is held to a CONTRACT against the
fp32-grade engine (DIM_PREC_FP32_TC, which meets the parity bars of test_slmft_gpu.py / test_baseline_configs_gpu.py) on ViCo-shape
clips at the full length T = 300:

  (a) teacher-forced decoder logits (same tokens, same context, 299 positions): max |difference| <= LOGIT_MAX, mean <= LOGIT_MEAN;
  (b) greedy decoding: a bf16 sequence leaves the fp32-grade sequence only where the fp32-grade top-2 logit margin is smaller than
      twice the logit error observed on the common prefix (a near-tie at bf16 resolution), and most of the prompt-adjacent steps agree;
  (c) VQ decode of the SAME codes: decoded FLAME coefficients within COEFF_MAX of the fp32-grade decode.

The bounds are 2-3x the values measured on B200 (profiles/r02_notes.md), i.e. they detect a precision regression, not noise.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu

import dim_b200  # noqa: E402
from dim_b200.schema import S2SConfig, VQConfig  # noqa: E402

S2S, VQ = S2SConfig(), VQConfig()
LOGIT_MAX, LOGIT_MEAN, COEFF_MAX = 0.05, 0.008, 0.04          # measured: 0.0152 / 0.0024 (0.0209 / 0.0031 with bf16 encoders), 0.0133
B, T = 16, 300


@pytest.fixture(scope="module")
def setup(slmft_sd):
    from dim_b200.compat_api import listener_codes
    from dim_b200.engine import PREC_BF16, PREC_FP32_TC, Handle, SLMFTEngine, VQEngine
    h = Handle()
    h.register(slmft_sd)
    hi = SLMFTEngine(h, S2S, precision=PREC_FP32_TC)
    lo = SLMFTEngine(h, S2S, precision=PREC_BF16)
    vq_hi = VQEngine(h, VQ, prefix="listener_vq.", precision=PREC_FP32_TC)
    vq_lo = VQEngine(h, VQ, prefix="listener_vq.", precision=PREC_BF16)
    c = dim_b200.synth.make_clips(B, T, seed=909, ragged=True)
    d = {k: c[k].cuda() for k in ("v_speaker", "v_listener", "v_audio", "mask")}
    z_l = listener_codes(vq_hi, d["v_listener"], d["mask"])
    ctx = hi.context(d["v_speaker"], d["v_audio"], d["mask"])                # one context for both: (a) and (b) isolate the decoder
    return dict(hi=hi, lo=lo, vq_hi=vq_hi, vq_lo=vq_lo, d=d, z_l=z_l, ctx=ctx, lengths=c["lengths"])


def test_teacher_forced_logit_error(setup):
    s = setup
    inp = s["z_l"][:, :-1].clone()
    inp[inp == -100] = 0
    ref = s["hi"].teacher_forced(s["ctx"], s["d"]["mask"], inp)
    got = s["lo"].teacher_forced(s["ctx"], s["d"]["mask"], inp)
    valid = s["z_l"][:, 1:] != -100
    diff = (got - ref).abs()[valid]
    print(f"bf16 teacher-forced logits at T={T}: max |diff| {float(diff.max()):.4f}, mean {float(diff.mean()):.5f}")
    assert float(diff.max()) < LOGIT_MAX and float(diff.mean()) < LOGIT_MEAN
    # the bf16 context (speaker encoders in bf16) on top: still within the same bounds at the logits
    ctx_lo = s["lo"].context(s["d"]["v_speaker"], s["d"]["v_audio"], s["d"]["mask"])
    got2 = s["lo"].teacher_forced(ctx_lo, s["d"]["mask"], inp)
    diff2 = (got2 - ref).abs()[valid]
    print(f"  with the bf16 speaker encoders as well: max {float(diff2.max()):.4f}, mean {float(diff2.mean()):.5f}")
    assert float(diff2.max()) < 2 * LOGIT_MAX and float(diff2.mean()) < 2 * LOGIT_MEAN


def test_greedy_divergence_only_at_near_ties(setup):
    s = setup
    prompt = s["z_l"][:, 0]
    rc, rl = s["hi"].generate(s["ctx"], s["d"]["mask"], prompt, T - 1, return_logits=True)
    gc, gl = s["lo"].generate(s["ctx"], s["d"]["mask"], prompt, T - 1, return_logits=True)
    rc, rl, gc, gl = rc.cpu(), rl.cpu(), gc.cpu(), gl.cpu()
    agree_len = []
    for b in range(B):
        neq = (gc[b] != rc[b]).nonzero()
        n = T - 1 if len(neq) == 0 else int(neq[0])
        agree_len.append(n)
        upto = min(n + 1, T - 1)
        e = float((gl[b, :upto] - rl[b, :upto]).abs().max())
        assert e < LOGIT_MAX, f"row {b}: logit error {e} on the common prefix"
        if n < T - 1:
            top2 = torch.topk(rl[b, n], 2).values
            margin = float(top2[0] - top2[1])
            assert margin < 2 * e + 1e-6, f"row {b}: bf16 left the fp32-grade sequence at step {n} where the margin {margin} exceeds twice the logit error {e}"
    print(f"bf16 greedy decode at T={T}: steps in common with the fp32-grade decode per clip: min {min(agree_len)}, median {sorted(agree_len)[B // 2]}, "
          f"max {max(agree_len)}")
    assert sorted(agree_len)[B // 2] >= 3


def test_vq_decode_of_same_codes(setup):
    s = setup
    codes = torch.randint(0, 512, (B, T - 1), generator=torch.Generator().manual_seed(5)).cuda()
    bi = torch.arange(B, dtype=torch.int32).cuda()
    ref = s["vq_hi"].decode(codes=codes, batch_index=bi)
    got = s["vq_lo"].decode(codes=codes, batch_index=bi)
    err = float((got - ref).abs().max())
    rel = float((got - ref).abs().mean() / ref.abs().mean())
    print(f"bf16 VQ decode of the same codes: max |coefficient diff| {err:.4f} (mean relative {rel:.4f})")
    assert err < COEFF_MAX
