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
