"""BASELINE configs[4] plumbing: the reference's test_l2l.py, UNCHANGED, run on predictions written by
dim_b200.l2l_artifacts, and its printed metrics compared with dim_b200.metrics on the same arrays.

Needs the reference tree (/root/reference: build container only) -- the script is executed from there, not copied.  Two shims,
both outside the script: a `pickle5` module (alias of pickle, the reference imports it) and scipy.linalg.sqrtm's removed `disp`
keyword (eval_utils.py:28), installed through a sitecustomize on PYTHONPATH."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

import dim_b200  # noqa: F401
from dim_b200 import l2l_artifacts as A
from dim_b200 import metrics as M

REF = "/root/reference/code"


@pytest.mark.skipif(not os.path.isfile(os.path.join(REF, "test_l2l.py")), reason="reference tree not present on this box")
def test_reference_test_l2l_runs_unchanged_on_our_artefacts(tmp_path):
    g = np.random.default_rng(11)
    ids = [f"clip{i:02d}" for i in range(4)]
    lens = [70, 64, 90, 81]
    gt = [g.standard_normal((n, 56)).astype(np.float32).cumsum(0) * 0.05 for n in lens]
    pred = [a + g.standard_normal(a.shape).astype(np.float32) * 0.1 for a in gt]
    x = [g.standard_normal((n, 56)).astype(np.float32).cumsum(0) * 0.05 for n in lens]
    data = tmp_path / "data"
    A.write_l2l_fixtures(str(data), ids, gt, x)
    A.write_l2l_predictions(str(data / "l2l_vico_predictions.pkl"), ids, pred)
    shim = tmp_path / "shim"
    shim.mkdir()
    (shim / "pickle5.py").write_text("from pickle import *\nfrom pickle import load, dump, loads, dumps\n")
    (shim / "sitecustomize.py").write_text(
        "import scipy.linalg as _sl\n_f = _sl.sqrtm\n_sl.sqrtm = lambda a, disp=True, **kw: _f(a) if disp else (_f(a), 0.0)\n")
    cwd = tmp_path / "code"
    cwd.mkdir()
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([str(shim), REF]), OMP_NUM_THREADS="4")
    r = subprocess.run([sys.executable, os.path.join(REF, "test_l2l.py")], cwd=str(cwd), env=env, capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    printed = {}
    for line in r.stdout.splitlines():
        m = re.match(r"^([a-z_ ]+):\s+(.*)$", line.strip())
        if m:
            printed[m.group(1).strip()] = [float(v) for v in m.group(2).split()]
    ours = M.metrics_suite([torch.from_numpy(a) for a in gt], [torch.from_numpy(a) for a in pred], [torch.from_numpy(a) for a in x])
    pairs = {"fid_pose": "fid_pose", "fid_exp": "fid_exp", "pfid_pose": "pfid_pose", "pfid_exp": "pfid_exp", "mse_pose": "mse_pose",
             "mse_exp": "mse_exp", "rpcc pose": "rpcc_pose", "rpcc exp": "rpcc_exp", "sts pose": "sts_pose"}
    for k_ref, k in pairs.items():
        assert k_ref in printed, (k_ref, sorted(printed))
        assert abs(printed[k_ref][0] - ours[k]) <= 1e-4 * abs(ours[k]) + 1e-7, (k_ref, printed[k_ref], ours[k])
    for k_ref, k in (("sid_pose", "sid_pose"), ("sid_exp", "sid_exp"), ("var_pose", "var_pose"), ("var_exp", "var_exp")):
        assert np.allclose(printed[k_ref], ours[k], rtol=1e-4), (k_ref, printed[k_ref], ours[k])
