"""Per-shape timing of the large-M (prefill) tcgen05 GEMMs: persistent kernel vs the one-tile-per-CTA kernel (GPU box only).
usage: python scripts/prefill_gemm_bench.py [M]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dim_b200
from dim_b200 import _lib, ops

lib = _lib.load()
M = int(sys.argv[1]) if len(sys.argv) > 1 else 76800
SHAPES = [  # (N, K, act, residual, what)
    (2304, 384, 0, 0, "s2s encoder QKV"), (384, 768, 0, 1, "s2s encoder out-proj + res"), (1536, 384, 3, 0, "s2s encoder FF1 (GELU)"),
    (384, 1536, 0, 1, "s2s encoder FF2 + res"), (1536, 1152, 0, 0, "cross K/V projection"), (1152, 384, 0, 0, "VQ QKV"),
    (384, 384, 0, 1, "VQ out-proj + res"), (1536, 384, 2, 0, "VQ FF1 (GELU tanh)"), (4608, 1152, 3, 0, "decoder-width FF1"),
    (1152, 4608, 0, 1, "decoder-width FF2 + res"),
]


def time_call(call, iters=10):
    for _ in range(3): call()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(iters): call()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / iters


if len(sys.argv) > 2 and sys.argv[2] == "bn":
    # tile-width sweep through the tuning hook dim_debug_tc_bn(-bn)
    for planes in (1, 3):
        for (N, K) in ((1152, 384), (384, 384), (384, 1536), (1152, 4608), (1536, 384), (2304, 384), (384, 768)):
            a = torch.randn(M, K, device="cuda"); w = torch.randn(N, K, device="cuda") / K ** 0.5
            ap, wp = ops.split_planes(a, planes), ops.split_planes(w, planes)
            out = torch.empty(M, N, device="cuda")
            def call():
                _lib.check(lib.dim_linear_bf16_planes(ap.data_ptr(), wp.data_ptr(), K, planes, None, None, N, out.data_ptr(), N, M, N, 0,
                                                      0.0, torch.cuda.current_stream().cuda_stream))
            fl = 2.0 * M * N * K * {1: 1, 3: 6}[planes]
            row = []
            for bn in (64, 96, 128, 192, 256):
                lib.dim_debug_tc_bn(-bn)
                t = time_call(call)
                row.append(f"bn={bn}: {t:7.1f} us {fl / t / 1e6:6.0f} TF")
            lib.dim_debug_tc_bn(0)
            print(f"planes={planes} N={N} K={K}: " + " | ".join(row), flush=True)
    sys.exit(0)

for planes in (1, 3):
    for (N, K, act, res, what) in SHAPES:
        a = torch.randn(M, K, device="cuda"); w = torch.randn(N, K, device="cuda") / K ** 0.5
        b = torch.randn(N, device="cuda") * 0.1
        r = torch.randn(M, N, device="cuda") if res else None
        ap, wp = ops.split_planes(a, planes), ops.split_planes(w, planes)
        out = torch.empty(M, N, device="cuda")
        def call():
            _lib.check(lib.dim_linear_bf16_planes(ap.data_ptr(), wp.data_ptr(), K, planes, b.data_ptr(), r.data_ptr() if res else None, N, out.data_ptr(), N, M, N, act,
                                                  0.0, torch.cuda.current_stream().cuda_stream))
        npairs = {1: 1, 2: 3, 3: 6}[planes]
        fl = 2.0 * M * N * K * npairs
        by = 2.0 * (M + N) * K * planes + 4.0 * M * N * (2 if res else 1)
        t_new = time_call(call)
        lib.dim_debug_tc_bn(128)
        t_old = time_call(call)
        lib.dim_debug_tc_bn(0)
        print(f"planes={planes} N={N:5d} K={K:5d} act={act} {what:28s}: persistent {t_new:8.1f} us = {fl / t_new / 1e6:7.1f} TFLOP/s ({by / t_new / 1e3:6.0f} GB/s)"
              f" | one-tile {t_old:8.1f} us = {fl / t_old / 1e6:7.1f} TFLOP/s", flush=True)
        del a, w, ap, wp, out, r
