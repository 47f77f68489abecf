"""The C-ABI shared library builds for sm_100a, loads without a GPU, exports every symbol include/dimb200.h declares,
and refuses to compute (loudly) when no device is present."""
import ctypes
import os
import re

import pytest
import torch

from dim_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "dimb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dim_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_what_python_binds():
    assert declared_symbols() == sorted(_lib.SIGNATURES)


def test_library_builds_loads_and_exports_every_symbol():
    _lib.build()
    lib = _lib.load()
    for name in declared_symbols():
        assert hasattr(lib, name), name
    assert lib.dim_version() >= 100


def test_sass_is_sm100():
    import shutil
    import subprocess
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not on PATH")
    out = subprocess.run(["cuobjdump", "--list-elf", _lib.SO_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    lib = _lib.load()
    h = ctypes.c_void_p()
    assert lib.dim_create(ctypes.byref(h), 0) != 0
    assert lib.dim_vq_argmin(None, None, None, 4, 128, 512, None) == 2          # DIM_ENODEVICE
    assert b"no CPU fallback" in lib.dim_last_error() or b"CUDA" in lib.dim_last_error()
    from dim_b200 import ops
    with pytest.raises((RuntimeError, ValueError)):
        ops.vq_argmin(torch.zeros(4, 128), torch.zeros(512, 128))
    from dim_b200.engine import Handle
    with pytest.raises(RuntimeError):
        Handle()


def test_committed_ncu_captures_parse():
    """profiles/*_ncu/*.raw.csv.gz (raw pages of the `ncu --set full` captures taken on the B200 box) stay readable by the
    summary script, and profiles/ncu_traffic.json carries the DRAM bytes per launch bench.py reports as `roofline.traffic`."""
    import glob
    import json
    import sys
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    import ncu_csv_summary as S
    paths = glob.glob(os.path.join(ROOT, "profiles", "r01z_ncu", "*.raw.csv.gz"))
    assert len(paths) >= 5
    names = set()
    for p in paths:
        hdr, units, rows = S.read(p)
        assert hdr and rows and "gpu__time_duration.sum" in hdr and "dram__bytes_read.sum" in hdr
        names.update(r[hdr.index("Kernel Name")] for r in rows)
    joined = " ".join(names)
    for kernel in ("attn_decode_kernel", "gemm_bf16_tcgen05", "attn_prefill_mma", "vq_gather_pf_kernel", "vq_argmin_f32", "layer_norm_kernel"):
        assert kernel in joined, kernel
    t = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
    assert 1.5e8 < t["attn_decode"] < 3e8 and t["vq_gather"] > 2e9
