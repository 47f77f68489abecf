"""VQSpeakerAutoEncoder (reference: code/models/stage1_BIWI.py:140-251, config_speaker_old.yaml: in_dim 824, hidden 768,
8 heads of 96, 8 codes per frame, decoder_v -> 56 and decoder_a -> 768 channels) on the GPU vs the golden vectors minted from the REAL
reference class (tests/golden/make_speaker_golden.py).  Same bar as the listener VQ-VAE: bit-exact code indices, floats within 1e-4."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

import dim_b200  # noqa: E402
from dim_b200.schema import SPEAKER_OUT_DIMS, SPEAKER_VQ  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOL = 1e-4


@pytest.fixture(scope="module")
def sp_golden():
    return torch.load(os.path.join(ROOT, "tests", "golden", "vq_speaker_reference.pt"), weights_only=False)


@pytest.fixture(scope="module")
def sp_sd(sp_golden):
    return dim_b200.synth.make_vqspeaker_state_dict(sp_golden["weights_seed"])


@pytest.fixture(scope="module", params=["fp32_ffma", "fp32_tcgen05"])
def engines(sp_sd, request):
    from dim_b200.engine import PREC_FP32, PREC_FP32_TC, Handle, VQEngine
    prec = PREC_FP32 if request.param == "fp32_ffma" else PREC_FP32_TC
    h = Handle()
    h.register(sp_sd)
    enc = VQEngine(h, SPEAKER_VQ, precision=prec, encoder="encoder", decoder=None)
    decs = [VQEngine(h, SPEAKER_VQ, precision=prec, encoder=None, decoder=n, out_dim=od) for n, od in SPEAKER_OUT_DIMS]
    return enc, decs


def _x(case):
    g = torch.Generator().manual_seed(case["x_seed"])
    return torch.randn(case["B"], case["T"], SPEAKER_VQ.in_dim, generator=g) * case["x_scale"]


@pytest.mark.parametrize("name", ["sp_T40_B2", "sp_T24_B3"])
def test_speaker_golden_roundtrip(engines, sp_golden, name):
    enc, decs = engines
    case = sp_golden["cases"][name]
    x = _x(case).cuda()
    idx, z, quant = enc.encode(x, want_z=True, want_quant=True)
    assert z.shape == (case["B"], case["T"], 1024) and idx.shape == (case["B"], case["T"] * 8)
    assert torch.allclose(z.cpu(), case["z"], atol=TOL), float((z.cpu() - case["z"]).abs().max())
    assert torch.equal(idx.cpu(), case["idx"]), f"{int((idx.cpu() != case['idx']).sum())} code mismatches"
    dec = torch.cat([d.decode(quant=quant) for d in decs], dim=-1)
    assert dec.shape == (case["B"], case["T"], 824)
    assert torch.allclose(dec.cpu(), case["dec_idx"], atol=TOL), float((dec.cpu() - case["dec_idx"]).abs().max())
    dec2 = torch.cat([d.decode(codes=idx) for d in decs], dim=-1)
    assert torch.equal(dec2, dec)


def test_halves_refuse_the_other_call(engines):
    enc, decs = engines
    with pytest.raises(RuntimeError, match="without a decoder"):
        enc.decode(codes=torch.zeros(1, 8, dtype=torch.int64, device="cuda"))
    with pytest.raises(RuntimeError, match="without an encoder"):
        decs[0].encode(torch.zeros(1, 4, 824, device="cuda"))


def test_compat_class_matches_golden(sp_sd, sp_golden):
    """models.get_model(config_speaker_old.yaml-shaped cfg) -> VQSpeakerAutoEncoder with the reference's API and state_dict keys."""
    compat = os.path.join(ROOT, "dyadic-interaction-modeling_b200", "compat")
    sys.path.insert(0, compat)
    try:
        from types import SimpleNamespace
        from models import get_model
        cfg = SimpleNamespace(arch="stage1_BIWI_speaker", in_dim=824, hidden_size=768, num_hidden_layers=6, num_attention_heads=8,
                              intermediate_size=1536, quant_factor=0, face_quan_num=8, neg=0.2, INaffine=False, n_embed=512, zquant_dim=128)
        m = get_model(cfg)
        assert type(m).__name__ == "VQSpeakerAutoEncoder"
        assert set(m.state_dict().keys()) == set(sp_sd.keys())
        m.load_state_dict(sp_sd, strict=True)
        m = m.cuda().eval()
        case = sp_golden["cases"]["sp_T24_B3"]
        x = _x(case).cuda()
        dec, loss, (ppl, onehot, idx) = m(x)
        assert torch.equal(idx.view(case["B"], -1).cpu(), case["idx"])
        assert abs(float(loss) - float(case["loss"])) < 1e-5
        assert float((dec.cpu() - case["dec"]).abs().max()) < TOL
        quant, _ = m.get_quant(x)
        assert quant.shape == (case["B"], 128, case["T"] * 8)
    finally:
        sys.path.remove(compat)
        for k in [k for k in sys.modules if k == "models" or k.startswith("models.") or k == "base" or k.startswith("base.")]:
            del sys.modules[k]
