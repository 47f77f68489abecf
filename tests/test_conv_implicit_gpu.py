"""The 5-tap replicate-padded Conv1d of the VQ-VAE (stage1_BIWI.py:265-267 squasher, :331-333 expander) as an IMPLICIT tcgen05 GEMM:
five row-shifted TMA boxes of one padded frame matrix instead of im2col plane rows.  Same products in the same order,
so the results must be BIT-identical to the explicit path (DIM_CONV_IM2COL=1; the switch is read once per process -> subprocesses),
for ragged clips (replicate padding at lens[b]-1), clips shorter than the kernel, both precisions and both VQ-VAE shapes."""
import os
import subprocess
import sys
import tempfile

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CODE = r'''
import sys, torch
sys.path.insert(0, %r)
import dim_b200
from dim_b200.engine import PREC_BF16, PREC_FP32_TC, Handle, VQEngine
from dim_b200.schema import SPEAKER_VQ, SPEAKER_OUT_DIMS, VQConfig
out = {}
sd = dim_b200.synth.make_vqvae_state_dict(131)
h = Handle(); h.register(sd)
for prec, name in ((PREC_FP32_TC, "tc"), (PREC_BF16, "bf16")):
    e = VQEngine(h, VQConfig(), precision=prec)
    for B, T in ((3, 37), (2, 3), (16, 300), (1, 1)):
        g = torch.Generator().manual_seed(B * 100 + T)
        x = (torch.randn(B, T, 56, generator=g) * 0.3).cuda()
        lens = torch.tensor([max(1, T - 5 * i) for i in range(B)], dtype=torch.int32).cuda()
        idx, z, _ = e.encode(x, lens=lens, want_z=True)
        dec = e.decode(codes=idx.view(B, T))
        out[(name, B, T)] = (idx.cpu(), z.cpu(), dec.cpu())
ssd = dim_b200.synth.make_vqspeaker_state_dict(137)
hs = Handle(); hs.register(ssd)
enc = VQEngine(hs, SPEAKER_VQ, precision=PREC_FP32_TC, encoder="encoder", decoder=None)
x = (torch.randn(2, 21, 824, generator=torch.Generator().manual_seed(5)) * 0.3).cuda()
idx, z, _ = enc.encode(x, want_z=True)
out[("speaker", 2, 21)] = (idx.cpu(), z.cpu())
torch.save(out, sys.argv[1])
''' % ROOT


def _run(env):
    with tempfile.NamedTemporaryFile(suffix=".pt") as f:
        r = subprocess.run([sys.executable, "-c", CODE, f.name], env=dict(os.environ, **env), capture_output=True, text=True, timeout=900)
        assert r.returncode == 0, r.stderr[-3000:]
        return torch.load(f.name, weights_only=False)


def test_implicit_conv_is_bit_identical_to_im2col():
    implicit, explicit = _run({}), _run({"DIM_CONV_IM2COL": "1"})
    assert implicit.keys() == explicit.keys() and len(implicit) == 9
    for k in implicit:
        for a, b in zip(implicit[k], explicit[k]):
            assert torch.equal(a, b), k
