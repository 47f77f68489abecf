"""Tuning aid: per-kernel cost INSIDE a CUDA graph (kernel time + dependent-launch gap) for the decode-step kernel types.
Captures N back-to-back launches of one kernel in a torch CUDA graph (our C-ABI launches on the capture stream), replays
it, and prints microseconds per kernel node.  Compare with the in-kernel cycle counts of scripts/tc_timeline.py."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import dim_b200  # noqa: F401
from dim_b200 import _lib, ops

lib = _lib.load()
N = 64


def graph_time(fn, label):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(N):
            fn()
    for _ in range(3):
        g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(10):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    print(f"{label}: {e0.elapsed_time(e1) * 1e3 / (10 * N):.2f} us per kernel node in a graph chain")


def main():
    M = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    x = torch.randn(M, 1152, device="cuda")
    g = torch.ones(1152, device="cuda")
    graph_time(lambda: ops.layer_norm(x, g), f"layer_norm {M}x1152 (fp32 out)")
    for (Nn, K) in ((2304, 1152), (1152, 768), (768, 1152), (4608, 1152), (1152, 4608), (512, 1152)):
        a = torch.randn(M, K, device="cuda")
        w = torch.randn(Nn, K, device="cuda") / K ** 0.5
        ap, wp = ops.split_planes(a, 1), ops.split_planes(w, 1)
        out = torch.empty(M, Nn, device="cuda")

        def call():
            _lib.check(lib.dim_linear_bf16_planes(ap.data_ptr(), wp.data_ptr(), K, 1, None, None, Nn, out.data_ptr(), Nn, M, Nn, 0, 0.0,
                                                  torch.cuda.current_stream().cuda_stream))
        graph_time(call, f"gemm bf16 M={M} N={Nn} K={K}")
    fn = lib.dim_debug_attn_decode
    fn.restype = C.c_int
    fn.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
    for Tk in (8, 150, 300):
        k = torch.randn(M, 12, Tk, 64, device="cuda").bfloat16()
        v = torch.randn(M, 12, Tk, 64, device="cuda").bfloat16()
        q = torch.randn(M, 768, device="cuda")
        o = torch.empty(M, 768, device="cuda")
        graph_time(lambda: _lib.check(fn(0, k.data_ptr(), v.data_ptr(), q.data_ptr(), o.data_ptr(), M, 12, Tk, 1,
                                         torch.cuda.current_stream().cuda_stream)), f"attn_decode bf16 {M} clips x {Tk} keys (L2-resident when small)")


if __name__ == "__main__":
    main()
