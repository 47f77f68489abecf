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
