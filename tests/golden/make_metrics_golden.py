"""Golden values of the reference's metric suite on a small synthetic eval result.

Imports the REAL reference (code/metrics/eval_utils.py; code/mymetrics.py's formulas are evaluated with those functions and the
numpy expressions of mymetrics.py:7-88 copied as calls, since print_metrics only prints).  Run in the build container:
    python tests/golden/make_metrics_golden.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, "/root/reference/code")
# The reference calls scipy.linalg.sqrtm(..., disp=False) (eval_utils.py:28), a keyword removed from current scipy: give it back
# the old calling convention ((sqrtm, error-estimate) tuple) so that the UNMODIFIED reference function runs.
import scipy.linalg as _sl  # noqa: E402
_sqrtm = _sl.sqrtm
_sl.sqrtm = lambda a, disp=True, **kw: _sqrtm(a) if disp else (_sqrtm(a), 0.0)
from metrics.eval_utils import calculate_activation_statistics, calculate_frechet_distance, calcuate_sid, sts  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "metrics_reference.pt")


def main():
    g = np.random.default_rng(3)
    lens = [90, 120, 75, 140, 101]
    gt = [g.standard_normal((n, 56)).astype(np.float32).cumsum(0) * 0.05 for n in lens]
    pred = [a + g.standard_normal(a.shape).astype(np.float32) * 0.1 for a in gt]
    x = [g.standard_normal((n, 56)).astype(np.float32).cumsum(0) * 0.05 for n in lens]

    def fd(a, b):
        m1, s1 = calculate_activation_statistics(a)
        m2, s2 = calculate_activation_statistics(b)
        return calculate_frechet_distance(m1, s1, m2, s2)

    P, E = slice(0, 6), slice(6, 56)
    ref = {}
    for name, c in (("pose", P), ("exp", E)):
        ref[f"fid_{name}"] = float(np.mean([fd(a[:, c], b[:, c]) for a, b in zip(gt, pred)]))
        ref[f"pfid_{name}"] = float(np.mean([fd(np.concatenate([s[:, c], a[:, c]], -1), np.concatenate([s[:, c], b[:, c]], -1))
                                             for a, b, s in zip(gt, pred, x)]))
        ref[f"mse_{name}"] = float(np.mean([np.mean((a[:, c] - b[:, c]) ** 2) for a, b in zip(gt, pred)]))
        ref[f"sid_{name}"] = (float(calcuate_sid(gt, pred, type=name)), float(calcuate_sid(gt, gt, type=name)))
    G, Pr, X = np.concatenate(gt, 0), np.concatenate(pred, 0), np.concatenate(x, 0)
    for name, c in (("pose", P), ("exp", E)):
        ref[f"var_{name}"] = (float(np.var(G[:, c].reshape(-1))), float(np.var(Pr[:, c].reshape(-1))))
        pcc_xy = np.corrcoef(G[:, c].reshape(-1), X[:, c].reshape(-1))[0, 1]
        pcc_xp = np.corrcoef(Pr[:, c].reshape(-1), X[:, c].reshape(-1))[0, 1]
        ref[f"rpcc_{name}"] = float(abs(pcc_xy - pcc_xp))
        ref[f"sts_{name}"] = float(sts(G[:, c], Pr[:, c]))
    torch.save({"gt": [torch.from_numpy(a) for a in gt], "pred": [torch.from_numpy(a) for a in pred],
                "x": [torch.from_numpy(a) for a in x], "ref": ref}, OUT)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")
    for k, v in ref.items():
        print(k, v)


if __name__ == "__main__":
    main()
