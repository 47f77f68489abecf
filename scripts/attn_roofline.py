"""Decode-attention roofline: one cross-attention-shaped launch (one query per (clip, head), Tk keys) timed alone with CUDA
events, cycling over several K/V buffers so that every launch streams from HBM (footprint >> 126 MB L2).
Algorithmic bytes = K and V rows read once: B*H*Tk*64*2*elem_size.  One JSON line per configuration."""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import dim_b200  # noqa: F401
from dim_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0


def main():
    lib = _lib.load()
    fn = lib.dim_debug_attn_decode
    fn.restype = C.c_int
    fn.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
    s = torch.cuda.current_stream().cuda_stream
    H, NB = 12, 6
    for bf16 in (1, 0):
        dt = torch.bfloat16 if bf16 else torch.float32
        for B, Tk in ((256, 300), (256, 150), (256, 40), (256, 1024)):
            ks = [torch.randn(B, H, Tk, 64, device="cuda").to(dt) for _ in range(NB)]
            vs = [torch.randn(B, H, Tk, 64, device="cuda").to(dt) for _ in range(NB)]
            q = torch.randn(B, H * 64, device="cuda")
            out = torch.empty(B, H * 64, device="cuda")
            ref = None
            for impl, name in ((0, "ring (default: bf16 nt64 tile64; fp32 nt128 tile64)"), (4, "ring nt128 tile128/64"), (2, "ring nt64 tile64/32"), (3, "ring nt128 tile64/32"), (1, "lanes")):
                def call(i):
                    _lib.check(fn(impl, ks[i % NB].data_ptr(), vs[i % NB].data_ptr(), q.data_ptr(), out.data_ptr(), B, H, Tk, bf16, s))
                for i in range(3):
                    call(i)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                iters = 30
                e0.record()
                for i in range(iters):
                    call(i)
                e1.record()
                torch.cuda.synchronize()
                t = e0.elapsed_time(e1) / iters * 1e-3
                call(0)
                o = out.clone()
                if ref is None:
                    att = torch.softmax(torch.einsum("bhd,bhkd->bhk", q.view(B, H, 64), ks[0].float()) * 0.125, -1)
                    ref = torch.einsum("bhk,bhkd->bhd", att, vs[0].float()).reshape(B, H * 64)
                nbytes = B * H * Tk * 64 * 2 * (2 if bf16 else 4)
                print(json.dumps({"kernel": "attn_decode", "impl": name, "kv": "bf16" if bf16 else "fp32", "clips": B, "keys": Tk,
                                  "us": t * 1e6, "GB/s": nbytes / t / 1e9, "frac_of_hbm_peak": nbytes / t / 1e9 / PEAK, "peak": PEAK,
                                  "max_abs_err_vs_torch": float((o - ref).abs().max())}), flush=True)
            del ks, vs


if __name__ == "__main__":
    main()
