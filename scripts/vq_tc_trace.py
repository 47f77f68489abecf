"""clock64 timeline of CTA 0 of the tensor-core codebook argmin (csrc/vq_tc.cu): per role and tile, cycles relative to the MMA warp's
first event of the first printed tile."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import dim_b200  # noqa: F401
from dim_b200 import _lib, ops  # noqa: F401

lib = _lib.load()
fn = lib.dim_debug_vq_argmin_tc_trace
fn.restype = C.c_int
fn.argtypes = [C.c_void_p] * 3 + [C.c_int, C.c_void_p, C.c_void_p]
g = torch.Generator().manual_seed(0)
N = 1 << 20
E = (torch.randn(512, 128, generator=g) * 0.5).cuda()
z = (torch.randn(N, 128, generator=g) * 0.7).cuda()
o = torch.empty(N, dtype=torch.int64, device="cuda")
tr = torch.zeros(4 * 16 * 12, dtype=torch.int64, device="cuda")
s = torch.cuda.current_stream().cuda_stream
lib.dim_debug_vq_argmin_impl.argtypes = [C.c_int]
lib.dim_debug_vq_argmin_impl(int(os.environ.get("VQ_TC_DBG", "0")) << 8)
for _ in range(3):
    assert fn(z.data_ptr(), E.data_ptr(), o.data_ptr(), N, tr.data_ptr(), s) == 0
torch.cuda.synchronize()
t = tr.cpu().view(4, 16, 12)
names = {0: ("mma", ["zfull", "accfree0", "accfree1"]), 1: ("rerank", ["listfull", "evald", "bar"]), 2: ("conv", ["zfree", "conv_a", "conv_b"]),
         3: ("epi", ["full0", "p1h0", "full1", "p1h1", "bar", "listfree", "p2h0", "p2h1", "pushed"])}
t0 = int(t[0, 6, 0])
for it in range(6, 11):
    for role in (2, 0, 3, 1):
        nm, ev = names[role]
        print(f"tile {it:2d} {nm:7s} " + "  ".join(f"{e}={int(t[role, it, i]) - t0}" for i, e in enumerate(ev)))
