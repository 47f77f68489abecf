"""Golden values of the reference's BIWI vertex metrics (code/mymetrics.py:122-182, print_biwi_metrics: lip vertex error + FDD).

Runs the REAL reference function in a scratch directory laid out like the paths it hard-codes (../data/BIWI_data/templates.pkl,
../data/CodeTalker/BIWI/regions/{lve,fdd}.txt) on seeded synthetic vertices (23370 x 3 per frame, the size it hard-codes):
    python tests/golden/make_biwi_metrics_golden.py
Inputs are regenerated from the stored seed by tests/test_metrics.py::biwi_case."""
import os
import pickle
import shutil
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference/code")


def biwi_case(seed=11):
    g = np.random.default_rng(seed)
    V3 = 23370 * 3
    names = ["F1_e01", "M2_e07", "F1_e03"]
    templates = {"F1": g.standard_normal(V3).astype(np.float32) * 0.1, "M2": g.standard_normal(V3).astype(np.float32) * 0.1}
    lens = [7, 5, 9]
    gt = [templates[n.split("_")[0]][None] + g.standard_normal((L, V3)).astype(np.float32).cumsum(0) * 0.01 for n, L in zip(names, lens)]
    pred = [np.concatenate([a + g.standard_normal(a.shape).astype(np.float32) * 0.02, a[-2:]], 0) for a in gt]   # 2 surplus frames (:145)
    mouth = sorted(g.choice(23370, 40, replace=False).tolist())
    upper = sorted(g.choice(23370, 55, replace=False).tolist())
    return names, templates, gt, pred, mouth, upper


def main():
    import mymetrics as R
    names, templates, gt, pred, mouth, upper = biwi_case()
    work = tempfile.mkdtemp(prefix="biwi_golden_")
    os.makedirs(os.path.join(work, "code"))
    os.makedirs(os.path.join(work, "data", "BIWI_data"))
    os.makedirs(os.path.join(work, "data", "CodeTalker", "BIWI", "regions"))
    pickle.dump(templates, open(os.path.join(work, "data", "BIWI_data", "templates.pkl"), "wb"))
    open(os.path.join(work, "data", "CodeTalker", "BIWI", "regions", "lve.txt"), "w").write(", ".join(map(str, mouth)))
    open(os.path.join(work, "data", "CodeTalker", "BIWI", "regions", "fdd.txt"), "w").write(", ".join(map(str, upper)))
    cwd = os.getcwd()
    os.chdir(os.path.join(work, "code"))
    try:
        lve, fdd = R.print_biwi_metrics(gt, pred, names)
    finally:
        os.chdir(cwd)
        shutil.rmtree(work)
    out = os.path.join(HERE, "biwi_metrics_reference.pt")
    torch.save({"seed": 11, "lve": float(lve), "fdd": float(fdd)}, out)
    print("lve", lve, "fdd", fdd, "->", out)


if __name__ == "__main__":
    main()
