"""Tuning aid: per-phase clock64 timeline of one CTA of the tcgen05 GEMM + event timings over shapes (GPU box only)."""
import ctypes as C
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dim_b200
from dim_b200 import _lib, ops

lib = _lib.load()
lib.dim_debug_tc.argtypes = [C.c_void_p, C.c_int]
dbg = torch.zeros(64, dtype=torch.int64, device="cuda")


def run(M, N, K, planes=1, splits=0, iters=50, quiet=None):
    a = torch.randn(M, K, device="cuda"); w = torch.randn(N, K, device="cuda") / K ** 0.5
    ap, wp = ops.split_planes(a, planes), ops.split_planes(w, planes)
    out = torch.empty(M, N, device="cuda")
    def call():
        _lib.check(lib.dim_linear_bf16_planes(ap.data_ptr(), wp.data_ptr(), K, planes, None, None, N, out.data_ptr(), N, M, N, 0, 0.0,
                                              torch.cuda.current_stream().cuda_stream))
    lib.dim_debug_tc(dbg.data_ptr(), splits)
    dbg.zero_(); call(); torch.cuda.synchronize()
    d = dbg.cpu().tolist()
    lib.dim_debug_tc(None, splits)
    for _ in range(5): call()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(iters): call()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / iters
    t0 = d[0]
    tma = [x - t0 for x in d[8:24] if x]
    mma = [x - t0 for x in d[24:40] if x]
    if quiet is not None:
        print(f"M={M} N={N} K={K} {quiet} splits={splits}: {us:.2f} us/launch | kernel cycles={d[4]-t0} accum_ready={d[2]-t0} synced={d[6]-t0} epi_done={d[3]-t0}")
        return
    print(f"M={M} N={N} K={K} planes={planes} splits={splits}: {us:.1f} us/launch | cycles: setup={d[1]-t0} accum_ready={d[2]-t0} first_chunk={d[40]-t0} staged={d[5]-t0} synced={d[6]-t0} epi_done={d[3]-t0} end={d[4]-t0}")
    print("   tma issue:", tma)
    print("   mma start:", mma)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "sweep3":
        # same sweep for the fp32-grade mode (3 bf16 planes, 6 plane products): M = 128 rows of one decode group
        for (N, K) in ((2304, 1152), (1152, 768), (768, 1152), (4608, 1152), (1152, 4608), (512, 1152)):
            for bn in (128, 64, 32):
                for sp in (1, 2, 4, 8):
                    lib.dim_debug_tc_bn(bn)
                    run(int(sys.argv[2]) if len(sys.argv) > 2 else 128, N, K, planes=3, splits=sp, quiet=f"planes=3 bn={bn}")
        lib.dim_debug_tc_bn(0)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "sweep":
        # tile width x split-K sweep for the decode-step GEMM shapes (M = 128 rows of one decode group)
        for (N, K) in ((2304, 1152), (1152, 768), (768, 1152), (4608, 1152), (1152, 4608), (512, 1152)):
            for bn in (128, 64, 32):
                for sp in (1, 2, 4, 8):
                    if sp > K // 64 // 2:
                        continue
                    lib.dim_debug_tc_bn(bn)
                    run(int(sys.argv[2]) if len(sys.argv) > 2 else 128, N, K, planes=1, splits=sp, quiet=f"bn={bn}")
        lib.dim_debug_tc_bn(0)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "decode":
        # the seven GEMM shapes of one decode step (M = rows of one decode group), bf16 and 3-plane, automatic split-K
        for M in (128, 256):
            for (N, K) in ((2304, 1152), (1152, 768), (768, 1152), (4608, 1152), (1152, 4608), (512, 1152)):
                run(M, N, K, planes=1)
        for (N, K) in ((2304, 1152), (1152, 768), (1152, 4608)):
            run(128, N, K, planes=3)
            run(128, N, K, planes=1, splits=1)
        sys.exit(0)
    run(256, 2304, 1152); run(256, 2304, 1152, splits=1); run(128, 128, 384, splits=1); run(128, 32, 384, splits=1)
    run(8192, 1152, 1152, splits=1)
