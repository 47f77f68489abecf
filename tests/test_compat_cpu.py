"""The drop-in module surface (dyadic-interaction-modeling_b200/compat): names, constructor contracts and state_dict keys
match the reference, checkpoints load strictly, and nothing computes without the CUDA library."""
import json
import os
import subprocess
import sys

import pytest
import torch

import dim_b200
from dim_b200.schema import slmft_schema, vqvae_schema

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COMPAT = os.path.join(ROOT, "dyadic-interaction-modeling_b200", "compat")
CFG = os.path.join(COMPAT, "config.yaml")


@pytest.fixture(scope="module")
def compat():
    sys.path.insert(0, COMPAT)
    import base.config as config
    import models
    import seq2seq_pretrain
    import x_engine_pt
    yield dict(config=config, models=models, s2s=seq2seq_pretrain, engine=x_engine_pt)
    sys.path.remove(COMPAT)


def test_config_loader_flattens_and_coerces(compat):
    cfg = compat["config"].load_cfg_from_cfg_file(CFG)
    assert cfg.arch == "stage1_BIWI" and cfg.hidden_size == 384 and cfg.n_embed == 512 and cfg.neg == 0.2
    cfg2 = compat["config"].merge_cfg_from_list(cfg, ["batch_size", "4", "TRAIN.base_lr", "0.5", "neg", "1"])
    assert cfg2.batch_size == 4 and cfg2.base_lr == 0.5 and cfg2.neg == 1.0 and isinstance(cfg2.neg, float)
    with pytest.raises(AssertionError):
        compat["config"].merge_cfg_from_list(cfg, ["no_such_key", "1"])


def test_vq_state_dict_matches_schema_and_loads_strict(compat, vq_sd):
    cfg = compat["config"].load_cfg_from_cfg_file(CFG)
    m = compat["models"].get_model(cfg)
    sd = m.state_dict()
    assert list(sd.keys()) == list(vqvae_schema().keys()) or set(sd.keys()) == set(vqvae_schema().keys())
    assert {k: tuple(v.shape) for k, v in sd.items()} == dict(vqvae_schema())
    m.load_state_dict(vq_sd, strict=True)
    assert torch.equal(m.quantize.embedding.weight, vq_sd["quantize.embedding.weight"])
    assert abs(float(m.__class__(cfg).quantize.embedding.weight.abs().max())) <= 1.0 / 512 + 1e-9     # reference init U(+-1/n_e)
    with pytest.raises(Exception):
        cfg_bad = compat["config"].merge_cfg_from_list(cfg, ["arch", "stage2"])
        compat["models"].get_model(cfg_bad)


@pytest.mark.skipif(not os.path.isdir("/root/reference/code"), reason="reference tree not present on this box")
def test_vq_keys_equal_the_real_reference():
    code = ("import sys, json; sys.path.insert(0, '/root/reference/code');"
            "from base import config; from models import get_model;"
            "m = get_model(config.load_cfg_from_cfg_file('/root/reference/code/config.yaml'));"
            "print(json.dumps({k: list(v.shape) for k, v in m.state_dict().items()}))")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, check=True).stdout
    ref = {k: tuple(v) for k, v in json.loads(out.strip().splitlines()[-1]).items()}
    assert ref == dict(vqvae_schema())


def test_slmft_state_dict_matches_schema_and_loads_strict(compat, slmft_sd):
    m = compat["s2s"].SLMFT(config_path=CFG, load_vq_checkpoints=False)
    sd = m.state_dict()
    want = dict(slmft_schema())
    assert {k: tuple(v.shape) for k, v in sd.items()} == want
    m.load_state_dict(slmft_sd, strict=True)
    for name in ("speaker_vq", "listener_vq", "encoder_s", "encoder_l", "encoder_joint", "decoder_joint", "patch_embed_s",
                 "patch_embed_l", "patch_embed_dec_s", "patch_embed_dec_l", "norm_s", "norm_l", "norm"):
        assert hasattr(m, name), name
    assert hasattr(m.decoder_joint, "net") and hasattr(m.decoder_joint, "generate")
    assert not any(p.requires_grad for p in m.listener_vq.parameters())            # frozen like the reference (:351-364)
    assert 149e6 < sum(p.numel() for p in m.parameters()) < 151e6          # SLMFT ~ 150 M parameters


def test_no_cpu_fallback(compat, slmft_sd):
    m = compat["s2s"].SLMFT(config_path=CFG, load_vq_checkpoints=False)
    c = dim_b200.synth.make_clips(1, 8)
    with pytest.raises(RuntimeError):
        m(c["v_speaker"], c["v_listener"], c["v_audio"], c["mask"], mode="val")
    with pytest.raises(RuntimeError):                      # the teacher-forced forward runs on the same CUDA-only engines
        m(c["v_speaker"], c["v_listener"], c["v_audio"], c["mask"], mode="train")
    cfg = compat["config"].load_cfg_from_cfg_file(CFG)
    with pytest.raises(RuntimeError):
        compat["models"].get_model(cfg).encode(c["v_listener"])


def test_pos_embed_and_x_utils_importable(compat):
    import pos_embed
    import x_utils
    e = pos_embed.get_1d_sincos_pos_embed_from_grid(8, [0, 1, 2])
    assert e.shape == (3, 8) and abs(e[0, 4] - 1.0) < 1e-12
    d = {"num_tokens": 512, "max_seq_len": 2048, "depth": 4}
    assert x_utils.pick_and_pop(["num_tokens", "max_seq_len"], d) == {"num_tokens": 512, "max_seq_len": 2048} and d == {"depth": 4}
    assert x_utils.groupby_prefix_and_trim("attn_", {"attn_heads": 2, "depth": 1}) == ({"heads": 2}, {"depth": 1})


def test_frechet_distance_torch_matches_scipy_formula():
    """compat_api.frechet_distance_torch (fp64, two eigh) == metrics/eval_utils.py's numpy-cov + scipy-sqrtm formula, on
    full-rank (n > d) and rank-deficient (n < d) samples; batched over the leading dimension of y."""
    import numpy as np
    from scipy import linalg
    from dim_b200.compat_api import frechet_distance_torch
    g = torch.Generator().manual_seed(0)
    for n, d in ((200, 56), (40, 56), (12, 8)):
        x = torch.randn(n, d, generator=g)
        y = torch.randn(3, n, d, generator=g) * 1.3 + 0.2
        got = frechet_distance_torch(x, y).numpy()
        for j in range(3):
            a, b = x.numpy().astype(np.float64), y[j].numpy().astype(np.float64)
            mu1, s1, mu2, s2 = a.mean(0), np.cov(a, rowvar=False), b.mean(0), np.cov(b, rowvar=False)
            cm = linalg.sqrtm(s1.dot(s2))
            cm = cm.real if np.iscomplexobj(cm) else cm
            ref = (mu1 - mu2).dot(mu1 - mu2) + np.trace(s1) + np.trace(s2) - 2 * np.trace(cm)
            assert abs(got[j] - ref) <= 1e-6 * max(1.0, abs(ref)) + (1e-3 if n < d else 0.0), (n, d, got[j], ref)


def test_mymetrics_print_metrics_matches_reference_golden(compat, capsys):
    """compat/mymetrics.print_metrics prints the reference's lines (mymetrics.py:7-88) with the golden values of
    tests/golden/make_metrics_golden.py and returns (fid_pose, fid_exp) like the reference."""
    import mymetrics
    g = torch.load(os.path.join(ROOT, "tests", "golden", "metrics_reference.pt"), weights_only=False)
    to_np = lambda seq: [t.numpy() for t in seq]
    if torch.cuda.is_available():
        pytest.skip("device placement is covered by tests/test_metrics.py on the GPU box")
    fp, fe = mymetrics.print_metrics(to_np(g["gt"]), to_np(g["pred"]), to_np(g["x"]))
    assert abs(fp - g["ref"]["fid_pose"]) < 1e-5 * g["ref"]["fid_pose"] and abs(fe - g["ref"]["fid_exp"]) < 1e-5 * g["ref"]["fid_exp"]
    lines = capsys.readouterr().out.strip().splitlines()
    assert [l.split(":")[0] for l in lines] == ["fid_pose", "fid_exp", "pfid_pose", "pfid_exp", "mse_pose", "mse_exp", "sid_pose",
                                                "sid_exp", "var_pose", "var_exp", "rpcc pose", "rpcc exp", "sts pose", "sts exp"]
    mymetrics.print_metrics_full(to_np(g["gt"]), to_np(g["pred"]), to_np(g["x"]))
    assert [l.split(":")[0] for l in capsys.readouterr().out.strip().splitlines()] == ["fid", "pfid", "mse", "var"]


def test_compat_eval_utils_numpy_api_matches_reference_golden(compat):
    """compat/metrics/eval_utils.py keeps the reference's numpy-in / float-out helpers (sts, calcuate_sid, calculate_variance,
    FD) that test_l2l.py and mymetrics import; values against tests/golden/metrics_reference.pt."""
    import numpy as np
    from metrics import eval_utils as EU
    g = torch.load(os.path.join(ROOT, "tests", "golden", "metrics_reference.pt"), weights_only=False)
    gt, pred = [t.numpy() for t in g["gt"]], [t.numpy() for t in g["pred"]]
    G, P = np.concatenate(gt, 0), np.concatenate(pred, 0)
    assert abs(EU.sts(G[:, 0:6], P[:, 0:6]) - g["ref"]["sts_pose"]) < 1e-5 * g["ref"]["sts_pose"]
    assert abs(EU.sts(G[:, 6:], P[:, 6:]) - g["ref"]["sts_exp"]) < 1e-5 * g["ref"]["sts_exp"]
    if not torch.cuda.is_available():
        assert abs(EU.calcuate_sid(gt, pred, type="pose") - g["ref"]["sid_pose"][0]) < 1e-9
        assert abs(EU.calcuate_sid(gt, gt, type="exp") - g["ref"]["sid_exp"][1]) < 1e-9
    assert abs(EU.calculate_variance(G) - float(np.sum(np.var(G, axis=0)))) < 1e-12
    mu1, s1 = EU.calculate_activation_statistics(gt[0][:, 0:6])
    mu2, s2 = EU.calculate_activation_statistics(pred[0][:, 0:6])
    assert EU.calculate_frechet_distance(mu1, s1, mu2, s2) > 0
