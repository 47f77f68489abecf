"""dim_b200.metrics (the reference's metric suite as device tensor ops) against golden values produced by the
reference's own functions (tests/golden/make_metrics_golden.py imports code/metrics/eval_utils.py unmodified).

Tolerances: FD / P-FD 1e-5 relative (scipy's Schur sqrtm vs two eigh); MSE / var / STS 1e-5 relative (the reference accumulates
some of them in fp32, here everything is fp64); rPCC 1e-9 abs; SID exact cluster assignments -> 1e-9."""
import os

import pytest
import torch

import dim_b200  # noqa: F401
from dim_b200 import metrics as M

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "metrics_reference.pt")


def _check(out, ref):
    for k in ("fid_pose", "fid_exp", "pfid_pose", "pfid_exp", "mse_pose", "mse_exp", "sts_pose", "sts_exp"):
        assert abs(out[k] - ref[k]) <= 1e-5 * abs(ref[k]) + 1e-9, (k, out[k], ref[k])
    for k in ("var_pose", "var_exp"):
        assert all(abs(a - b) <= 1e-5 * abs(b) for a, b in zip(out[k], ref[k])), (k, out[k], ref[k])
    for k in ("rpcc_pose", "rpcc_exp"):
        assert abs(out[k] - ref[k]) < 1e-7, (k, out[k], ref[k])          # corrcoef of fp32 inputs: fp32 -> fp64 promotion order
    for k in ("sid_pose", "sid_exp"):
        assert all(abs(a - b) < 1e-9 for a, b in zip(out[k], ref[k])), (k, out[k], ref[k])


def test_metric_suite_matches_reference_golden_cpu():
    g = torch.load(GOLD, weights_only=False)
    _check(M.metrics_suite(g["gt"], g["pred"], g["x"]), g["ref"])


def test_sts_counts_clip_boundaries_like_the_reference():
    """The reference differences the CONCATENATED sequence (mymetrics.py:66-86), so the jump between two clips contributes."""
    a = torch.zeros(4, 6)
    b = torch.zeros(4, 6)
    b[2:] = 1.0                                                          # one jump of 1 in every dim at frame 2
    assert abs(float(M.sts(a, b)) - (6 * 1.0 / 0.1) ** 0.5) < 1e-12


@pytest.mark.gpu
def test_metric_suite_on_the_gpu():
    """Same values with every tensor on the device; the SID assignment runs on the codebook-argmin kernel."""
    g = torch.load(GOLD, weights_only=False)
    dev = lambda seq: [t.cuda() for t in seq]
    _check(M.metrics_suite(dev(g["gt"]), dev(g["pred"]), dev(g["x"])), g["ref"])


def _biwi_case():
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_biwi_metrics_golden", os.path.join(os.path.dirname(GOLD), "make_biwi_metrics_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)           # only its seeded input generator is used: nothing under /root/reference is read
    return mod.biwi_case(11)


def test_biwi_vertex_metrics_match_reference_golden(tmp_path, capsys):
    """mymetrics.print_biwi_metrics (code/mymetrics.py:122-182; the SpeakerSLMFT / BIWI eval): lip vertex error and FDD against the
    values the REAL reference function printed for the same seeded vertices (tests/golden/make_biwi_metrics_golden.py); the compat
    function reads the same two data files and prints the same two lines.  Tolerance 1e-5 relative (the reference works in fp32)."""
    import pickle
    import sys
    g = torch.load(os.path.join(os.path.dirname(GOLD), "biwi_metrics_reference.pt"), weights_only=False)
    names, templates, gt, pred, mouth, upper = _biwi_case()
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    t = lambda a: torch.from_numpy(a).to(dev)
    lve, fdd = M.biwi_metrics([t(a) for a in gt], [t(a) for a in pred], [t(templates[n.split("_")[0]]) for n in names], mouth, upper)
    assert abs(lve - g["lve"]) <= 1e-5 * abs(g["lve"]) and abs(fdd - g["fdd"]) <= 1e-5 * abs(g["fdd"]) + 1e-9
    (tmp_path / "regions").mkdir()
    pickle.dump(templates, open(tmp_path / "templates.pkl", "wb"))
    (tmp_path / "regions" / "lve.txt").write_text(", ".join(map(str, mouth)))
    (tmp_path / "regions" / "fdd.txt").write_text(", ".join(map(str, upper)))
    compat = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "dyadic-interaction-modeling_b200", "compat")
    sys.path.insert(0, compat)
    try:
        import mymetrics
        lve2, fdd2 = mymetrics.print_biwi_metrics(gt, pred, names, templates_path=str(tmp_path / "templates.pkl"), region_path=str(tmp_path / "regions"))
    finally:
        sys.path.remove(compat)
    out = capsys.readouterr().out
    assert (lve2, fdd2) == (lve, fdd)
    assert "Lip Vertex Error: {:.4e}".format(g["lve"]) in out and "FDD: {:.4e}".format(g["fdd"]) in out
