"""CPU checks of the oracle itself: against the golden vectors minted from the real reference (VQ half) and
self-consistency of the restated x-transformers half (cache vs no cache, cross-KV once vs every step)."""
import hashlib
import math
import os

import torch

import dim_b200
from dim_b200.schema import S2SConfig, VQConfig, slmft_schema, vqvae_schema
from oracle import slmft as OS
from oracle import vqvae as OV
from oracle import xt as OX

VQ, S2S = VQConfig(), S2SConfig()


def _x(case):
    g = torch.Generator().manual_seed(case["x_seed"])
    return torch.randn(case["B"], case["T"], 56, generator=g) * case["x_scale"]


def test_schema_parameter_counts():
    n = sum(math.prod(s) for k, s in vqvae_schema().items() if not k.endswith(".pe"))
    assert n == 23258496                                    # SURVEY Appendix B (probe of the real reference)
    own = sum(math.prod(s) for k, s in slmft_schema().items()) - 2 * sum(math.prod(s) for s in vqvae_schema().values())
    assert 100e6 < own < 110e6


def test_synthetic_weights_are_reproducible(vq_sd, golden):
    h = hashlib.sha256()
    for k in sorted(vq_sd):
        h.update(k.encode())
        h.update(vq_sd[k].contiguous().numpy().tobytes())
    assert h.hexdigest() == golden["weights_sha256"]


def test_oracle_matches_reference_golden(vq_sd, golden):
    """The restatement reproduces the REAL reference's outputs stored in tests/golden/vq_reference.pt."""
    for name, case in golden["cases"].items():
        x = _x(case)
        quant, loss, (ppl, onehot, idx) = OV.encode(vq_sd, x, VQ)
        assert torch.equal(idx.view(case["B"], case["T"]), case["idx"]), name
        assert torch.allclose(OV.encoder(vq_sd, x, VQ), case["z"], atol=1e-5), name
        assert torch.allclose(OV.decode(vq_sd, quant, VQ), case["dec"], atol=1e-5), name
        assert torch.allclose(OV.decode_indices(vq_sd, case["idx"], VQ), case["dec_idx"], atol=1e-5), name
        assert torch.allclose(loss, case["loss"], atol=1e-6) and torch.allclose(ppl, case["perplexity"], atol=1e-4)
        assert onehot.sum() == case["B"] * case["T"]


def test_f4_batch_index_quirk(vq_sd, golden):
    assert golden["f4_fraction_changed_b1_vs_b0"] > 0.05
    x = torch.randn(1, 32, 56, generator=torch.Generator().manual_seed(4)) * 0.3
    a = OV.encoder(vq_sd, x.repeat(2, 1, 1), VQ)
    assert not torch.allclose(a[0], a[1])
    b = OV.encoder(vq_sd, x, VQ, batch_index=torch.tensor([1]))
    assert torch.allclose(a[1], b[0], atol=1e-6)


def test_codebook_entry_is_a_gather(vq_sd):
    idx = torch.randint(0, 512, (77,), generator=torch.Generator().manual_seed(1))
    assert torch.equal(OV.codebook_entry(vq_sd, idx), vq_sd["quantize.embedding.weight"][idx])


def _ctx(slmft_sd, B, T, seed, ragged=False):
    c = dim_b200.synth.make_clips(B, T, seed=seed, ragged=ragged)
    x_s = OS.forward_encoder(slmft_sd, c["v_speaker"], c["mask"], S2S)
    return c, OS.decoder_context(slmft_sd, x_s, c["v_audio"])


def test_generate_variants_agree(slmft_sd):
    c, ctx = _ctx(slmft_sd, 2, 12, 3, ragged=True)
    prompt = torch.tensor([[5], [77]])
    base = OX.generate(slmft_sd, "decoder_joint.net", prompt, 11, 4, ctx, c["mask"])
    assert base.shape == (2, 11) and base.dtype == torch.int64
    assert torch.equal(base, OX.generate(slmft_sd, "decoder_joint.net", prompt, 11, 4, ctx, c["mask"], use_cache=False))
    assert torch.equal(base, OX.generate(slmft_sd, "decoder_joint.net", prompt, 11, 4, ctx, c["mask"], cross_kv_once=False))


def test_teacher_forced_logits_match_generate(slmft_sd):
    c, ctx = _ctx(slmft_sd, 1, 10, 4)
    prompt = torch.tensor([[9]])
    codes, logits = OX.generate(slmft_sd, "decoder_joint.net", prompt, 9, 4, ctx, c["mask"], return_logits=True)
    seq = torch.cat([prompt, codes], 1)
    _, tf = OX.teacher_forced(slmft_sd, "decoder_joint.net", seq, 4, ctx, c["mask"])
    assert torch.allclose(tf, logits, atol=2e-5)


def test_context_mask_blocks_padding(slmft_sd):
    """Frames beyond a clip's length must not influence its generated codes (cross-attention key mask)."""
    c, ctx = _ctx(slmft_sd, 1, 16, 5)
    mask = c["mask"].clone()
    mask[0, 10:] = False
    ctx2 = ctx.clone()
    ctx2[0, 10:] = 123.0
    p = torch.tensor([[1]])
    a = OX.generate(slmft_sd, "decoder_joint.net", p, 8, 4, ctx, mask)
    b = OX.generate(slmft_sd, "decoder_joint.net", p, 8, 4, ctx2, mask)
    assert torch.equal(a, b)


def test_top_k_and_inverse_cdf():
    logits = torch.tensor([[0.0, 3.0, 1.0, 2.0, -1.0]])
    f = OX.top_k_filter(logits, k=2)
    assert torch.isinf(f[0, [0, 2, 4]]).all() and f[0, 1] == 3 and f[0, 3] == 2
    assert math.ceil(0.1 * 512) == 52
    probs = torch.tensor([[0.1, 0.0, 0.6, 0.3]])
    assert OX.sample_from_uniform(probs, torch.tensor([0.05])).item() == 0
    assert OX.sample_from_uniform(probs, torch.tensor([0.1000001])).item() == 2
    assert OX.sample_from_uniform(probs, torch.tensor([0.95])).item() == 3
    assert OX.sample_from_uniform(probs, torch.tensor([1.0])).item() == 3


def test_forward_val_shapes_and_as_reference_equivalence(slmft_sd):
    c = dim_b200.synth.make_clips(2, 10, seed=6, ragged=True)
    a = OS.forward_val(slmft_sd, c["v_speaker"], c["v_listener"], c["v_audio"], c["mask"], S2S, VQ)
    b = OS.forward_val(slmft_sd, c["v_speaker"], c["v_listener"], c["v_audio"], c["mask"], S2S, VQ, as_reference=True,
                       cross_kv_once=False)
    assert a[2].shape == (2, 9, 56) and torch.equal(a[2], b[2]) and set(a[1]) == {"l_ce_s", "l_ce_l", "l_cont_s", "l_cont_l", "nce", "c_acc"}


def test_feature_resamplers_match_reference_golden():
    """oracle/features.py vs the outputs of the reference's own functions (tests/golden/make_resample_golden.py)."""
    import numpy as np
    from oracle import features as OF
    g = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "resample_reference.pt"), weights_only=False)
    for name, c in g["cases"].items():
        x = c["x"].numpy()
        assert np.array_equal(OF.window_mean(x), c["window_mean_0.6"].numpy()), name
        assert np.array_equal(OF.window_mean(x, factor=0.25), c["window_mean_0.25"].numpy()), name
        assert np.array_equal(OF.window_mean(x), x[: int(len(x) * 0.6)].astype(np.float64))        # the window-of-1 quirk
        for n, ref in c["linear_new_t"].items():
            assert np.array_equal(OF.linear(x, n), ref.numpy()), (name, n)


def test_speaker_oracle_matches_reference_golden():
    """VQSpeakerAutoEncoder (stage1_BIWI.py:140-251): oracle/vqvae.py against the goldens minted from the real reference class
    (tests/golden/make_speaker_golden.py): 8 codes per frame, two decoders concatenated."""
    import os
    from dim_b200.schema import SPEAKER_VQ
    g = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "vq_speaker_reference.pt"), weights_only=False)
    sd = dim_b200.synth.make_vqspeaker_state_dict(g["weights_seed"])
    case = g["cases"]["sp_T24_B3"]
    x = torch.randn(case["B"], case["T"], 824, generator=torch.Generator().manual_seed(case["x_seed"])) * case["x_scale"]
    with torch.no_grad():
        quant, loss, (_, _, idx) = OV.encode(sd, x, SPEAKER_VQ)
        dec = OV.speaker_decode(sd, quant, SPEAKER_VQ)
    assert torch.equal(idx.view(case["B"], -1), case["idx"])
    assert float((dec - case["dec"]).abs().max()) < 1e-5 and abs(float(loss) - float(case["loss"])) < 1e-6


def test_emoca_converter_oracle_matches_reference_golden():
    """EmocaConverter (seq2seq_pretrain.py:759-832): oracle/speaker_mesh.py against the golden minted from the real reference class
    (tests/golden/make_emoca_golden.py) at the real 70110-d mesh size; the state_dict key set is the reference's."""
    import os
    from oracle import speaker_mesh as OM
    g = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "emoca_reference.pt"), weights_only=False)
    sd = dim_b200.synth.make_emoca_converter_state_dict(g["weights_seed"])
    gen = torch.Generator().manual_seed(g["x_seed"])
    v = torch.randn(g["B"], g["T"], 56, generator=gen) * 0.3
    tpl = torch.randn(g["B"], 70110, generator=gen) * 0.1
    out, dec = OM.emoca_converter_forward(sd, tpl, v, VQConfig())
    assert float((dec - g["dec"]).abs().max()) <= 1e-5
    assert float((out[..., ::g["stride"]] - g["out_strided"]).abs().max()) <= 1e-5
    assert abs(float(out.double().sum()) - g["out_sum"]) <= 1e-5 * g["out_abs_sum"]
    assert float((OM.mesh_to_motion(sd, out, tpl) - g["motion"]).abs().max()) <= 1e-4
