"""SURVEY 8(f).4, last sibling: EmocaConverter / SpeakerSLMFT (reference code/seq2seq_pretrain.py:516-832) on the B200 kernels.

Kernel level: dim_lstm_layer_f32 against torch.nn.LSTM on the CPU (the reference's own dependency), dim_linear_ragged_f32 against
float64 matmuls, the 56-channel squasher.  Model level: the compat classes against oracle/speaker_mesh.py and against the golden
outputs minted from the REAL EmocaConverter class (tests/golden/make_emoca_golden.py).
Tolerances (floating point; stated per assert): LSTM outputs live in (-1, 1): 2e-5 absolute; Linear over K = 70110: 2e-4 relative
to the output scale (fp32 accumulation order differs from MKL's)."""
import os
import sys

import pytest
import torch
import torch.nn as nn

pytestmark = pytest.mark.gpu

import dim_b200  # noqa: E402
from dim_b200 import compat_api, ops  # noqa: E402
from dim_b200.schema import S2SConfig, VQConfig  # noqa: E402
from oracle import speaker_mesh as OM  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COMPAT = os.path.join(ROOT, "dyadic-interaction-modeling_b200", "compat")
GOLDEN = os.path.join(ROOT, "tests", "golden", "emoca_reference.pt")


@pytest.fixture(scope="module")
def compat():
    sys.path.insert(0, COMPAT)
    import seq2seq_pretrain
    yield dict(s2s=seq2seq_pretrain)
    sys.path.remove(COMPAT)


def _build(cls, tmp_path, **kw):
    import shutil
    shutil.copy(os.path.join(COMPAT, "config.yaml"), tmp_path / "config.yaml")
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        return cls(**kw)
    finally:
        os.chdir(cwd)


@pytest.mark.parametrize("B,T,in_dim,H,layers,bidir", [
    (1, 27, 56, 384, 2, True),        # the reference's __main__ shape at B = 1 (:834-841)
    (8, 138, 56, 384, 2, True),       # EmocaConverter example (:846)
    (33, 50, 56, 384, 2, True),       # several lanes per row, ragged last pass
    (256, 12, 56, 384, 2, True),      # one row per thread
    (300, 6, 56, 384, 1, True),       # two passes over the batch
    (5, 40, 64, 128, 1, False),       # unidirectional, other hidden size (8 units per CTA)
    (3, 1, 56, 384, 2, True),         # a single step: no recurrence at all
])
def test_lstm_matches_torch(B, T, in_dim, H, layers, bidir):
    torch.manual_seed(B * 1000 + T)
    ref = nn.LSTM(input_size=in_dim, hidden_size=H, num_layers=layers, batch_first=True, bidirectional=bidir).eval()
    x = torch.randn(B, T, in_dim)
    with torch.no_grad():
        want, _ = ref(x)
    params = {k: v.detach().cuda() for k, v in ref.named_parameters()}
    got = ops.lstm(x.cuda(), params, hidden=H, layers=layers, bidirectional=bidir)
    assert got.shape == want.shape
    err = float((got.cpu() - want).abs().max())
    assert err <= 2e-5, err
    again = ops.lstm(x.cuda(), params, hidden=H, layers=layers, bidirectional=bidir)
    assert torch.equal(got, again)                                   # deterministic for a given batch size


@pytest.mark.parametrize("M,N,K,act", [
    (54, 56, 70110, 1),       # vertice_mapping (:777): split over K
    (26, 70110, 768, 0),      # vertice_map_reverse[2] (:806): unaligned N and ldc
    (5, 7, 13, 1), (64, 64, 16, 0), (65, 3, 2049, 0), (1, 70110, 768, 0),
])
def test_linear_ragged(M, N, K, act):
    g = torch.Generator().manual_seed(M + N + K)
    a = torch.randn(M, K, generator=g)
    w = torch.randn(N, K, generator=g) / K ** 0.5
    b = torch.randn(N, generator=g)
    want = a.double() @ w.double().t() + b.double()
    if act:
        want = torch.where(want > 0, want, want * 0.2)
    got = ops.linear_ragged(a.cuda(), w.cuda(), b.cuda(), act=act, slope=0.2).cpu()
    err = float((got.double() - want).abs().max())
    assert err <= 2e-4 * max(1.0, float(want.abs().max())), err
    assert torch.equal(got, ops.linear_ragged(a.cuda(), w.cuda(), b.cuda(), act=act, slope=0.2).cpu())


def test_mesh_to_motion_and_back_small():
    """size = 150 (50 vertices): every kernel of the two mesh paths against the torch modules."""
    sd = dim_b200.synth.make_emoca_converter_state_dict(7, size=150)
    g = torch.Generator().manual_seed(5)
    v, tpl = torch.randn(3, 20, 150, generator=g), torch.randn(3, 150, generator=g)
    want = OM.mesh_to_motion(sd, v, tpl)
    c = {k: t.cuda() for k, t in sd.items()}
    got = compat_api.mesh_to_motion(v.cuda(), tpl.cuda(), c["vertice_mapping.0.weight"], c["vertice_mapping.0.bias"],
                                    c["squasher.0.0.weight"], c["squasher.0.0.bias"])
    assert float((got.cpu() - want).abs().max()) <= 2e-4              # InstanceNorm output, O(1) values
    dec = torch.randn(3, 19, 56, generator=g) * 0.5
    for which, lstm, head in (("emoca", "vertice_map_reverse_lstm", "vertice_map_reverse"),
                              ("mesh", "vertice_map_reverse_lstm_2", "vertice_map_reverse2")):
        want = OM.motion_to_mesh(sd, dec, which, tpl)
        lp = {k[len(lstm) + 1:]: t for k, t in c.items() if k.startswith(lstm + ".")}
        got = compat_api.motion_to_mesh(dec.cuda(), lp, c[head + ".0.weight"], c[head + ".0.bias"], c[head + ".2.weight"],
                                        c[head + ".2.bias"], template=tpl.cuda())
        assert float((got.cpu() - want).abs().max()) <= 5e-5, which


def test_emoca_converter_class_full_size(compat, tmp_path):
    """EmocaConverter (:759-832) at the real 70110-d mesh size: reference state_dict keys, forward == oracle, and == the golden
    outputs of the REAL reference class when the fixture is present."""
    sd = dim_b200.synth.make_emoca_converter_state_dict(131)
    model = _build(compat["s2s"].EmocaConverter, tmp_path, load_vq_checkpoints=False)
    assert set(model.state_dict().keys()) == set(sd.keys())
    model.load_state_dict(sd, strict=True)
    model = model.cuda().eval()
    g = torch.Generator().manual_seed(17)
    B, T = 2, 12
    v = torch.randn(B, T, 56, generator=g) * 0.3
    tpl = torch.randn(B, 70110, generator=g) * 0.1
    out, none = model(None, tpl.cuda(), v.cuda())
    assert none is None and out.shape == (B, T, 70110)
    want, dec = OM.emoca_converter_forward(sd, tpl, v, VQConfig())
    assert float((out.cpu() - want).abs().max()) <= 1e-4
    motion = model.forward_motion(want.cuda(), tpl.cuda())            # the commented-out input path (:817-819), (B,T,56)
    assert float((motion.cpu() - OM.mesh_to_motion(sd, want, tpl)).abs().max()) <= 5e-4
    if os.path.exists(GOLDEN):
        gold = torch.load(GOLDEN, weights_only=False)
        assert gold["weights_seed"] == 131 and gold["x_seed"] == 17
        assert float((out.cpu()[..., ::gold["stride"]] - gold["out_strided"]).abs().max()) <= 1e-4
        assert abs(float(out.double().sum()) - gold["out_sum"]) <= 1e-4 * gold["out_abs_sum"]
        assert float((motion.cpu() - gold["motion"]).abs().max()) <= 5e-4


def _speaker_case(B, T, size, seed):
    g = torch.Generator().manual_seed(seed)
    c = dim_b200.synth.make_clips(B, T, seed=seed, ragged=B > 1)
    v_mesh = torch.randn(B, T, size, generator=g) * 0.2
    tpl = torch.randn(B, size, generator=g) * 0.1
    return c, v_mesh, tpl


@pytest.mark.parametrize("mode,ids", [("train", None), ("val", None), ("val", [3])])
def test_speaker_slmft_forward_small(compat, tmp_path, mode, ids):
    """SpeakerSLMFT.forward (:707-757) at B = 1 (the only batch size its mouth loss admits, :738-739), size = 150."""
    size, mouth = 150, [1, 4, 7, 20, 33, 49]
    sd = dim_b200.synth.make_speaker_slmft_state_dict(131, size=size)
    model = _build(compat["s2s"].SpeakerSLMFT, tmp_path, load_checkpoints=False, size=size, mouth_map=mouth)
    assert set(model.state_dict().keys()) == set(sd.keys())
    model.load_state_dict(sd, strict=True)
    model = model.cuda().eval()
    model.greedy = True
    c, v_mesh, tpl = _speaker_case(1, 24, size, 9)
    sid = None if ids is None else torch.tensor(ids)
    total, d, pred = model(v_mesh.cuda(), c["v_listener"].cuda(), c["v_audio"].cuda(), c["mask"].cuda(), tpl.cuda(), mode=mode,
                           speaker_ids=None if sid is None else sid.cuda())
    rt, rd, rp, parts = OM.speaker_slmft_forward(sd, v_mesh, c["v_listener"], c["v_audio"], c["mask"], tpl, mouth, S2SConfig(),
                                                 VQConfig(), mode=mode, speaker_ids=sid, temperature=0.0)
    assert set(d) == {"l_ce_s", "l_ce_l", "l_cont_s", "l_cont_l", "nce", "c_acc"}
    assert torch.equal(model.last_parts["z"].cpu(), parts["z"])
    if torch.equal(model.last_parts["codes"].cpu(), parts["codes"]):
        assert float((pred.cpu() - rp).abs().max()) <= 1e-4
        assert float((model.last_parts["pred_mesh"].cpu() - parts["pred_mesh"]).abs().max()) <= 1e-4
        for k in ("l_ce_l", "l_cont_s", "l_cont_l"):
            assert abs(float(d[k]) - float(rd[k])) <= 1e-4 * max(1.0, abs(float(rd[k]))), k
        assert abs(float(total) - float(rt)) <= 1e-4 * max(1.0, abs(float(rt)))
    else:   # a code flipped: only legitimate at a logit near-tie (fp32 accumulation order); prove it
        assert mode == "val" or parts["logits"] is not None
        a, b = model.last_parts["codes"].cpu(), parts["codes"]
        first = int((a != b).nonzero()[0, 1])
        if mode == "train":
            lg = parts["logits"][0, first]
            assert abs(float(lg[a[0, first]] - lg[b[0, first]])) <= 1e-4
        else:
            pytest.fail(f"greedy decode diverged at step {first} without a tie proof")


def test_speaker_slmft_rejects_batches_like_the_reference(compat, tmp_path):
    """:738-739 compares B*(T-1) predicted rows with B*T-1 target rows: for B > 1 torch raises, here as there."""
    size, mouth = 150, [1, 4, 7]
    sd = dim_b200.synth.make_speaker_slmft_state_dict(131, size=size)
    model = _build(compat["s2s"].SpeakerSLMFT, tmp_path, load_checkpoints=False, size=size, mouth_map=mouth)
    model.load_state_dict(sd, strict=True)
    model = model.cuda().eval()
    model.greedy = True
    c, v_mesh, tpl = _speaker_case(2, 10, size, 4)
    with pytest.raises(RuntimeError):
        model(v_mesh.cuda(), c["v_listener"].cuda(), c["v_audio"].cuda(), c["mask"].cuda(), tpl.cuda(), mode="train")


def test_cpu_tensors_are_refused():
    with pytest.raises(ValueError):
        ops.linear_ragged(torch.zeros(2, 3), torch.zeros(4, 3))
