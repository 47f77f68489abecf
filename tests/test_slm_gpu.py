"""SLM pre-training forward (reference: code/seq2seq_pretrain.py:72-323) on the GPU against the restated oracle
(oracle/slmft.py: slm_forward; its x-transformers half is the unpinned restatement of oracle/xt.py, so this is a consistency check
vs the restated reference, like every seq2seq test).  The two random masks are passed to both sides."""
import pytest
import torch

pytestmark = pytest.mark.gpu

import dim_b200  # noqa: E402
from dim_b200 import compat_api  # noqa: E402
from dim_b200.schema import S2SConfig, VQConfig  # noqa: E402
from oracle import slmft as OS  # noqa: E402

S2S, VQ = S2SConfig(), VQConfig()


@pytest.fixture(scope="module")
def slm_sd():
    return dim_b200.synth.make_slm_state_dict(131)


@pytest.fixture(scope="module", params=["fp32_ffma", "fp32_tcgen05"])
def engines(slm_sd, request):
    from dim_b200.engine import PREC_FP32, PREC_FP32_TC, Handle, SLMFTEngine, VQEngine
    prec = PREC_FP32 if request.param == "fp32_ffma" else PREC_FP32_TC
    h = Handle()
    h.register(slm_sd)
    return (SLMFTEngine(h, S2S, precision=prec), VQEngine(h, VQ, prefix="speaker_vq.", precision=prec),
            VQEngine(h, VQ, prefix="listener_vq.", precision=prec))


def _case(B, T, seed):
    c = dim_b200.synth.make_clips(B, T, seed=seed, ragged=True)
    torch.manual_seed(seed)
    ms = compat_api.random_masking_unstructured(c["mask"], 0.15)
    ml = compat_api.random_masking_unstructured(c["mask"], 0.15)
    return c, ms, ml


def test_encoder_call_matches_oracle(engines, slm_sd):
    s2s, _, _ = engines
    c, _, _ = _case(3, 20, 5)
    x = c["v_listener"]
    from oracle import xt as OX
    ref = OX.continuous_wrapper(slm_sd, "encoder_l", x, S2S.depth, c["mask"])
    got = s2s.encode("encoder_l", x.cuda(), c["mask"].cuda())
    valid = c["mask"]
    assert float((got.cpu() - ref)[valid].abs().max()) < 1e-4
    ref2 = torch.nn.functional.layer_norm(OX.continuous_wrapper(slm_sd, "encoder_joint", ref, S2S.depth, c["mask"]), (S2S.dim,),
                                          slm_sd["norm.weight"], slm_sd["norm.bias"], 1e-5)
    got2 = s2s.encode("encoder_joint", ref.cuda(), c["mask"].cuda(), norm="norm")
    assert float((got2.cpu() - ref2)[valid].abs().max()) < 1e-4


def test_slm_forward_matches_oracle(engines, slm_sd):
    s2s, vq_s, vq_l = engines
    c, ms, ml = _case(4, 24, 7)
    ref_total, ref_d, ref_p = OS.slm_forward(slm_sd, c["v_speaker"], c["v_listener"], c["v_audio"], c["mask"], ms, ml, S2S, VQ)
    g = lambda k: slm_sd[k].cuda()
    total, d, p = compat_api.slm_forward(s2s, vq_s, vq_l, c["v_speaker"].cuda(), c["v_listener"].cuda(), c["v_audio"].cuda(), c["mask"].cuda(),
                                         g("patch_embed_s"), g("patch_embed_l"), g("patch_embed_dec_s"), g("patch_embed_dec_l"),
                                         mask_speaker=ms.cuda(), mask_listener=ml.cuda(), return_parts=True)
    m = c["mask"]
    assert torch.equal(p["z_s"].cpu(), ref_p["z_s"]) and torch.equal(p["z_l"].cpu(), ref_p["z_l"])        # VQ codes: bit-exact
    for k in ("x_s", "x_l"):
        assert float((p[k].cpu() - ref_p[k])[m].abs().max()) < 1e-4, k
    mj = torch.cat([m, m], dim=-1)
    assert float((p["x_joint"].cpu() - ref_p["x_joint"])[mj].abs().max()) < 1e-4
    # logits at the positions that enter the loss (targets != -100)
    for k, z in (("px_s", ref_p["z_s"]), ("px_l", ref_p["z_l"])):
        sel = z[:, 1:] != -100
        assert float((p[k].cpu() - ref_p[k])[sel].abs().max()) < 5e-4, k
    for k in ("l_ce_s", "l_ce_l", "nce"):
        assert abs(float(d[k]) - float(ref_d[k])) < 1e-4, (k, float(d[k]), float(ref_d[k]))
    assert float(d["c_acc"]) == float(ref_d["c_acc"])
    # decoded frames agree wherever the argmax codes agree (an argmax can flip only at a numerical tie)
    for k, pk in (("pred_s", "px_s"), ("pred_l", "px_l")):
        same = (p[pk].cpu().argmax(-1) == ref_p[pk].argmax(-1)).all(dim=1)
        assert int(same.sum()) >= 3
        assert float((p[k].cpu()[same] - ref_p[k][same]).abs().max()) < 1e-4
    if bool(((p["px_s"].cpu().argmax(-1) == ref_p["px_s"].argmax(-1)).all()) and ((p["px_l"].cpu().argmax(-1) == ref_p["px_l"].argmax(-1)).all())):
        assert abs(float(total) - float(ref_total)) < 5e-4
