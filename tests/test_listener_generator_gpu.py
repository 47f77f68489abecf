"""seq2seq.ListenerGenerator (reference: code/seq2seq.py:13-75,138-290) on the GPU against the restated oracle
(oracle/slmft.listener_generator: the real-reference-pinned VQ-VAE restatements + the unpinned x-transformers restatement), including
the reference's un-permuted view of the speaker latents and the decoder's absolute positional table inside `generate`."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

import dim_b200  # noqa: E402
from dim_b200.schema import SPEAKER_VQ, VQConfig  # noqa: E402
from oracle import slmft as OS  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COMPAT = os.path.join(ROOT, "dyadic-interaction-modeling_b200", "compat")


@pytest.fixture(scope="module")
def lg(tmp_path_factory):
    sd = dim_b200.synth.make_listener_generator_state_dict(131)
    tmp = tmp_path_factory.mktemp("lg")
    import shutil
    shutil.copy(os.path.join(COMPAT, "config.yaml"), tmp / "config.yaml")
    txt = open(os.path.join(COMPAT, "config.yaml")).read()
    txt = txt.replace("arch: stage1_BIWI", "arch: stage1_BIWI_speaker").replace("in_dim: 56", "in_dim: 824")
    txt = txt.replace("hidden_size: 384", "hidden_size: 768").replace("face_quan_num: 1", "face_quan_num: 8")
    (tmp / "config_speaker_old.yaml").write_text(txt)
    sys.path.insert(0, COMPAT)
    cwd = os.getcwd()
    os.chdir(tmp)
    try:
        import seq2seq
        model = seq2seq.ListenerGenerator(load_vq_checkpoints=False)
    finally:
        os.chdir(cwd)
    assert set(model.state_dict().keys()) == set(sd.keys()), set(model.state_dict().keys()) ^ set(sd.keys())
    model.load_state_dict(sd, strict=True)
    yield model.cuda().eval(), sd
    sys.path.remove(COMPAT)


def _clips(B, T, seed):
    c = dim_b200.synth.make_clips(B, T, seed=seed, ragged=True)
    g = torch.Generator().manual_seed(seed + 1)
    v_speaker = torch.cat([c["v_speaker"], torch.randn(B, T, 768, generator=g) * 0.3], dim=-1)       # 824-d speaker frames
    return v_speaker, c["v_listener"], c["mask"]


def test_forward_matches_oracle(lg):
    model, sd = lg
    v_s, v_l, mask = _clips(3, 16, 5)
    ref_loss, ref_pred, ref_logits, ref_x, ref_z = OS.listener_generator(sd, v_s, v_l, mask, SPEAKER_VQ, VQConfig())
    x_sp, z_l = model._inputs(v_s.cuda(), v_l.cuda(), mask.cuda())
    assert torch.equal(z_l.cpu(), ref_z)
    assert float((x_sp.cpu() - ref_x).abs().max()) < 1e-6                     # exact codebook rows vs the straight-through value
    loss, pred = model(v_s.cuda(), v_l.cuda(), mask.cuda())
    sel = ref_z[:, 1:] != -100
    s2s = model.engines()[0]
    enc = s2s.encode("encoder_s", x_sp, mask.cuda())
    inp = z_l[:, :-1].clone()
    inp[inp == -100] = 0
    logits = s2s.teacher_forced(enc, mask.cuda(), inp, None)
    assert float((logits.cpu() - ref_logits)[sel].abs().max()) < 5e-4
    same = (logits.cpu().argmax(-1) == ref_logits.argmax(-1)).all(dim=1)
    assert int(same.sum()) >= 2
    assert float((pred.cpu()[same] - ref_pred[same]).abs().max()) < 1e-4
    if bool(same.all()):
        assert abs(float(loss) - float(ref_loss)) < 5e-4


def test_generate_matches_oracle(lg):
    """generate(): KV-cached decode with the positional table added per step == the oracle's uncached full recompute."""
    model, sd = lg
    v_s, v_l, mask = _clips(2, 12, 9)
    out = OS.listener_generator(sd, v_s, v_l, mask, SPEAKER_VQ, VQConfig(), generate_steps=12)
    ref_gen = out[-1]
    pred, z_l = model.generate(v_s.cuda(), v_l.cuda(), mask.cuda())
    assert pred.shape == (2, 12) and torch.equal(z_l.cpu(), out[4])
    agree = (pred.cpu() == ref_gen)
    # greedy: sequences identical, or the first difference sits at an oracle near-tie (checked through the logits of that step)
    for b in range(2):
        if not bool(agree[b].all()):
            from parity_util import explain_first_difference
            s2s = model.engines()[0]
            x_sp, _ = model._inputs(v_s.cuda(), v_l.cuda(), mask.cuda())
            enc = s2s.encode("encoder_s", x_sp, mask.cuda())
            c2, logits = s2s.generate(enc, mask.cuda(), z_l[:, 0], 12, return_logits=True)
            from oracle import xt as OX
            enc_ref = OX.continuous_wrapper(sd, "generator.encoder", out[3], 6, mask)
            _, ref_logits = OX.generate(sd, "generator.decoder.net", out[4][:, 0:1], 12, 6, enc_ref, mask, use_cache=False, return_logits=True)
            explain_first_difference(c2.cpu()[b], logits[b].cpu(), ref_gen[b], ref_logits[b])


@pytest.mark.parametrize("sp,li", [(True, False), (False, True), (True, True)])
def test_identity_tokens_match_oracle(lg, sp, li):
    """speaker_ids / listener_ids (seq2seq.py:240-250; Transformer.forward :47-67): the speaker token enters the encoder input, the
    listener token the decoder context; the first logit row is dropped when the listener token is present."""
    model, sd = lg
    v_s, v_l, mask = _clips(3, 14, 11)
    sid = torch.tensor([5, 0, 99]) if sp else None
    lid = torch.tensor([7, 42, 1]) if li else None
    ref_loss, ref_pred, ref_logits, _, ref_z = OS.listener_generator(sd, v_s, v_l, mask, SPEAKER_VQ, VQConfig(), speaker_ids=sid, listener_ids=lid)
    loss, pred = model(v_s.cuda(), v_l.cuda(), mask.cuda(), speaker_ids=None if sid is None else sid.cuda(),
                       listener_ids=None if lid is None else lid.cuda())
    logits = model.last_logits.cpu()
    assert logits.shape == ref_logits.shape == (3, 13, 512)
    sel = ref_z[:, 1:] != -100
    assert float((logits - ref_logits)[sel].abs().max()) < 5e-4
    same = (logits.argmax(-1) == ref_logits.argmax(-1)).all(dim=1)
    assert int(same.sum()) >= 2
    assert float((pred.cpu()[same] - ref_pred[same]).abs().max()) < 1e-4
    if bool(same.all()):
        assert abs(float(loss) - float(ref_loss)) < 5e-4
