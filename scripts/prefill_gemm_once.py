"""Three launches of the persistent prefill GEMM for an ncu capture (GPU box only):
  ncu --set full --import-source on --clock-control none -k regex:gemm_bf16_tcgen05_persist -c 3 -o out python scripts/prefill_gemm_once.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dim_b200
from dim_b200 import _lib, ops

lib = _lib.load()
M = 76800
for (N, K, planes, act) in ((1536, 1152, 3, 0), (2304, 384, 1, 0), (1536, 384, 1, 3)):
    a = torch.randn(M, K, device="cuda"); w = torch.randn(N, K, device="cuda") / K ** 0.5
    b = torch.randn(N, device="cuda") * 0.1
    ap, wp = ops.split_planes(a, planes), ops.split_planes(w, planes)
    out = torch.empty(M, N, device="cuda")
    _lib.check(lib.dim_linear_bf16_planes(ap.data_ptr(), wp.data_ptr(), K, planes, b.data_ptr(), None, N, out.data_ptr(), N, M, N, act, 0.0,
                                          torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    print("done", N, K, planes, act, float(out[0, 0]))
