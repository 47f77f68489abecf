"""The import blocks of the reference's entry scripts resolve against compat/ (+ the reference's own host-side modules + the
import-time shims): what a maintainer needs before running test_s2s_pretrain.py / test_l2l.py unchanged.

Build container only (the scripts are parsed from /root/reference; nothing is executed beyond their import statements)."""
import ast
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COMPAT = os.path.join(ROOT, "dyadic-interaction-modeling_b200", "compat")
REF = "/root/reference/code"


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present on this box")
@pytest.mark.parametrize("script", ["test_s2s_pretrain.py", "test_l2l.py"])
def test_import_block_resolves(script):
    tree = ast.parse(open(os.path.join(REF, script)).read())
    imports = [n for n in tree.body if isinstance(n, (ast.Import, ast.ImportFrom))]
    assert imports
    src = "\n".join(ast.unparse(n) for n in imports)
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([ROOT, COMPAT, REF, os.path.join(COMPAT, "_shims")]))
    tail = ("\nimport seq2seq_pretrain, x_engine_pt, mymetrics\n"
            "assert all('dyadic-interaction-modeling_b200' in m.__file__ for m in (seq2seq_pretrain, x_engine_pt, mymetrics))\n"
            "print('imports ok')") if script == "test_s2s_pretrain.py" else "\nprint('imports ok')"
    r = subprocess.run([sys.executable, "-c", src + tail], env=env, capture_output=True, text=True, timeout=600, cwd="/tmp")
    assert r.returncode == 0 and "imports ok" in r.stdout, r.stderr[-3000:]


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present on this box")
def test_reference_vico_loader_reads_our_fixtures(tmp_path):
    """the REAL loader (dataset/data_loader.get_vico_dataloaders + pad_collate) on fixtures written by
    l2l_artifacts.write_vico_fixtures yields the batches evaluate_test_epoch consumes: src (1,T,824) = ones(56) | audio(768)
    (the loader replaces the speaker motion by ones, data_loader.py:147), tgt (1,T,56), src_len, ids."""
    import dim_b200
    from dim_b200 import l2l_artifacts as A
    clips = dim_b200.synth.make_clips(4, 40, seed=2, ragged=True)                 # clip 3 -> the (unused) train split
    clips = {k: (v.numpy() if hasattr(v, "numpy") else v) for k, v in clips.items()}
    A.write_vico_fixtures(str(tmp_path / "data"), clips, split=["test", "test", "test", "train"])
    cwd = tmp_path / "code"
    cwd.mkdir()
    code = ("from dataset.data_loader import get_vico_dataloaders\n"
            "import torch\n"
            "ds = get_vico_dataloaders(batch_size=1)\n"
            "out = []\n"
            "for src, tgt, src_len, ids, names in ds['valid']:\n"
            "    assert src.shape[2] == 824 and tgt.shape[2] == 56 and src.shape[1] == tgt.shape[1] == src_len[0]\n"
            "    assert bool((src[..., :56] == 1).all())\n"
            "    out.append((src_len[0], float(src[..., 56:].sum()), float(tgt.sum())))\n"
            "print('BATCHES', out)\n")
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([ROOT, COMPAT, REF, os.path.join(COMPAT, "_shims")]))
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600, cwd=str(cwd))
    assert r.returncode == 0, r.stderr[-3000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("BATCHES")][0]
    got = eval(line[len("BATCHES"):])
    assert len(got) == 3
    for i, (n, a_sum, t_sum) in enumerate(got):
        L = int(clips["lengths"][i])
        assert n == L
        assert abs(a_sum - float(clips["v_audio"][i][:L].sum())) < 1e-2 and abs(t_sum - float(clips["v_listener"][i][:L].sum())) < 1e-3
