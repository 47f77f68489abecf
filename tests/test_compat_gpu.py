"""Drop-in surface on the GPU: the flow of the reference's test_s2s_pretrain.py (SLMFT() from files in the working directory,
load_state_dict('best_vico_causal.pt'), evaluate_test_epoch) against the CPU oracle."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

import dim_b200  # noqa: E402
from dim_b200.schema import S2SConfig, VQConfig  # noqa: E402
from oracle import slmft as OS  # noqa: E402
from oracle import vqvae as OV  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COMPAT = os.path.join(ROOT, "dyadic-interaction-modeling_b200", "compat")


@pytest.fixture(scope="module")
def compat():
    sys.path.insert(0, COMPAT)
    import base.config as config
    import models
    import seq2seq_pretrain
    import x_engine_pt
    yield dict(config=config, models=models, s2s=seq2seq_pretrain, engine=x_engine_pt)
    sys.path.remove(COMPAT)


def test_vqautoencoder_api(compat, vq_sd):
    cfg = compat["config"].load_cfg_from_cfg_file(os.path.join(COMPAT, "config.yaml"))
    m = compat["models"].get_model(cfg)
    m.load_state_dict(vq_sd)
    m = m.cuda().eval()
    x = torch.randn(2, 40, 56, generator=torch.Generator().manual_seed(1)) * 0.3
    quant, loss, (ppl, onehot, idx) = m.encode(x.cuda())
    rq, rl, (rp, ro, ri) = OV.encode(vq_sd, x, VQConfig())
    assert quant.shape == (2, 128, 40) and onehot.shape == (80, 512) and idx.shape == (80, 1) and idx.dtype == torch.int64
    assert torch.equal(idx.cpu(), ri) and torch.equal(onehot.cpu(), ro)
    assert torch.allclose(quant.cpu(), rq, atol=1e-4) and abs(float(loss) - float(rl)) < 1e-4 and abs(float(ppl) - float(rp)) < 1e-3
    dec, _, _ = m(x.cuda())
    assert torch.allclose(dec.cpu(), OV.decode(vq_sd, rq, VQConfig()), atol=1e-4)
    img = m.decode_to_img(idx, (2, 40, 128))
    assert torch.allclose(img.cpu(), OV.decode_indices(vq_sd, ri.view(2, 40), VQConfig()), atol=1e-4)
    feat = m.entry_to_feature(idx, (2, 40, 128))
    assert torch.equal(feat.cpu(), vq_sd["quantize.embedding.weight"][ri.view(-1)].view(2, 40, 128))
    q2, i2 = m.get_quant(x.cuda())
    assert torch.equal(i2, idx)


def test_reference_eval_script_flow(compat, slmft_sd, tmp_path):
    """test_s2s_pretrain.py:41-75 with synthetic files: zero-arg SLMFT() in a directory shaped like the reference's code/."""
    (tmp_path / "runs_speaker_new" / "_RANK0" / "model").mkdir(parents=True)
    (tmp_path / "runs" / "listener_exp" / "model").mkdir(parents=True)
    import shutil
    shutil.copy(os.path.join(COMPAT, "config.yaml"), tmp_path / "config.yaml")
    sp = {k[len("speaker_vq."):]: v for k, v in slmft_sd.items() if k.startswith("speaker_vq.")}
    li = {k[len("listener_vq."):]: v for k, v in slmft_sd.items() if k.startswith("listener_vq.")}
    torch.save({"state_dict": sp}, tmp_path / "runs_speaker_new" / "_RANK0" / "model" / "model.pth.tar")
    torch.save({"state_dict": li}, tmp_path / "runs" / "listener_exp" / "model" / "model.pth.tar")
    torch.save(slmft_sd, tmp_path / "best_vico_causal.pt")
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        device = torch.device("cuda:0")
        model = compat["s2s"].SLMFT().to(device)                                  # test_s2s_pretrain.py:43
        model.load_state_dict(torch.load("best_vico_causal.pt"))                  # :46-47
    finally:
        os.chdir(cwd)
    model.eval()
    B, T = 2, 28
    c = dim_b200.synth.make_clips(B, T, seed=12, ragged=True)
    model.greedy = True
    loss, d, pred = model(c["v_speaker"].cuda(), c["v_listener"].cuda(), c["v_audio"].cuda(), c["mask"].cuda(), mode="val")
    ref_loss, ref_d, ref_pred, inter = OS.forward_val(slmft_sd, c["v_speaker"], c["v_listener"], c["v_audio"], c["mask"],
                                                      S2SConfig(), VQConfig(), return_intermediates=True)
    assert set(d) == set(ref_d) and pred.shape == (B, T - 1, 56)
    assert torch.equal(model.last_codes.cpu(), inter["codes"])
    assert torch.allclose(pred.cpu(), ref_pred, atol=1e-4) and abs(float(loss) - float(ref_loss)) < 1e-4
    # pieces of the reference API
    x_s = model.forward_encoder(c["v_speaker"].cuda(), c["mask"].cuda())
    m = c["mask"]
    assert torch.allclose(x_s.cpu()[m], inter["x_s"][m], atol=1e-4)
    z_s, z_l = model.forward_vq(c["v_speaker"].cuda(), c["v_listener"].cuda(), c["mask"].cuda())
    assert torch.equal(z_l.cpu(), inter["z_l"])

    # x_engine_pt.evaluate_test_epoch (x_engine_pt.py:232): loader yields (src (B,T,824), tgt, src_len, _, data_ids)
    model.greedy = False
    src = torch.cat([c["v_speaker"], c["v_audio"]], dim=2)
    loader = [(src, c["v_listener"], [int(v) for v in c["lengths"]], None, ["clip_a", "clip_b"])]
    y_true, y_pred, x_all, ids = compat["engine"].evaluate_test_epoch(model, loader, device, beam_size=3)
    assert ids == ["clip_a", "clip_b"] and len(y_pred) == 2
    for j in range(B):
        n = int(c["lengths"][j]) - 1
        assert y_true[j].shape == (n, 56) and y_pred[j].shape == (n, 56) and x_all[j].shape == (n, 56)


def test_slm_pretraining_class_forward(compat, tmp_path):
    """seq2seq_pretrain.SLM (the pre-training model, reference seq2seq_pretrain.py:72-323; what train_s2s_pretrain.py builds):
    zero-arg-style constructor from files, load_state_dict of an SLM checkpoint (SLMFT's keys + the decoder's positional table),
    forward() -> (total_loss, dict, None), equal to the restated oracle with the same random masks."""
    from dim_b200 import compat_api
    sd = dim_b200.synth.make_slm_state_dict(131)
    import shutil
    shutil.copy(os.path.join(COMPAT, "config.yaml"), tmp_path / "config.yaml")
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        model = compat["s2s"].SLM(load_vq_checkpoints=False).to("cuda:0")
    finally:
        os.chdir(cwd)
    assert set(model.state_dict().keys()) == set(sd.keys())
    model.load_state_dict(sd, strict=True)
    model.eval()
    c = dim_b200.synth.make_clips(3, 22, seed=21, ragged=True)
    torch.manual_seed(3)
    model.mask_speaker = compat_api.random_masking_unstructured(c["mask"], 0.15).cuda()
    model.mask_listener = compat_api.random_masking_unstructured(c["mask"], 0.15).cuda()
    total, d, none = model(c["v_speaker"].cuda(), c["v_listener"].cuda(), c["v_audio"].cuda(), c["mask"].cuda())
    assert none is None and set(d) == {"l_ce_s", "l_ce_l", "l_cont_s", "l_cont_l", "nce", "c_acc"}
    ref_total, ref_d, _ = OS.slm_forward(sd, c["v_speaker"], c["v_listener"], c["v_audio"], c["mask"], model.mask_speaker.cpu(),
                                         model.mask_listener.cpu(), S2SConfig(), VQConfig())
    for k in ("l_ce_s", "l_ce_l", "nce"):
        assert abs(float(d[k]) - float(ref_d[k])) < 1e-4, k
    assert torch.isfinite(total)


def test_autoregressive_wrapper_forward_delegates(compat, slmft_sd, tmp_path):
    """decoder_joint(z, context=, context_mask=, return_outputs=True) (seq2seq_pretrain.py:448) on the stand-in wrapper: the teacher-
    forced loss / logits of the owning model's engine, equal to forward_decoder(mode='train') with the same key mask."""
    import shutil
    shutil.copy(os.path.join(COMPAT, "config.yaml"), tmp_path / "config.yaml")
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        model = compat["s2s"].SLMFT(load_vq_checkpoints=False).to("cuda:0")
    finally:
        os.chdir(cwd)
    model.load_state_dict(slmft_sd)
    model.eval()
    c = dim_b200.synth.make_clips(2, 20, seed=3, ragged=True)
    m = c["mask"].cuda()
    x_s = model.forward_encoder(c["v_speaker"].cuda(), m)
    ctx = torch.cat([x_s + model.patch_embed_dec_s, c["v_audio"].cuda()], dim=-1)
    _, z_l = model.forward_vq(c["v_speaker"].cuda(), c["v_listener"].cuda(), m)
    kv = torch.ones(2, 19, dtype=torch.bool, device="cuda")
    loss, (logits, cache) = model.decoder_joint(z_l, context=ctx, context_mask=m, return_outputs=True, self_attn_kv_mask=kv)
    model.train_kv_mask = kv
    ref_loss, ref_logits = model.forward_decoder(x_s, z_l, c["v_audio"].cuda(), m, "train")
    assert cache is None and torch.equal(logits, ref_logits) and float(loss) == float(ref_loss)
