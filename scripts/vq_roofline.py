"""VQ-lookup roofline (BASELINE metric "VQ-lookup HBM GB/s vs peak"): times dim_vq_gather and dim_vq_argmin alone with CUDA
events on sizes larger than L2, algorithmic bytes = 520 B per code/token.  Prints one JSON line per kernel."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import dim_b200  # noqa: F401
from dim_b200 import ops

PEAK = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"] \
    if os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) else 6650.0


NCU = os.environ.get("VQ_NCU") == "1"        # under ncu: largest size only, one launch per implementation


def timeit(fn, iters=20, warm=3):
    if NCU:
        iters, warm = 1, 0
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3


def main():
    g = torch.Generator().manual_seed(0)
    E = (torch.randn(512, 128, generator=g) * 0.5).cuda()
    lib = dim_b200._lib.load()
    s = torch.cuda.current_stream().cuda_stream
    for N in ((1 << 22,) if NCU else (76544, 1 << 20, 1 << 22)):                     # B=256 x 299 codes; 0.5 GB; 2 GB of rows (> 126 MB L2)
        idx = torch.randint(0, 512, (N,), generator=g).cuda()
        out = torch.empty(N, 128, device="cuda")
        ref = E[idx]
        t = timeit(lambda: out.zero_())                       # write-only ceiling of the same footprint (cudaMemset-class)
        print(json.dumps({"kernel": "write_only_ceiling(torch zero_)", "codes": N, "us": t * 1e6, "GB/s": N * 512 / t / 1e9,
                          "frac_of_hbm_peak": N * 512 / t / 1e9 / PEAK}))
        for mode, name in ((0, "warp-per-row st.cs"), (1, "smem-staged TMA bulk store"), (2, "prefetched idx + st.cs"),
                           (3, "prefetched idx + plain st"), (6, "prefetched idx, 256-code chunks"), (4, "prefetched idx, 128-code chunks"),
                           (5, "prefetched idx, 64-code chunks"), (-1, "default")):
            lib.dim_debug_vq_gather_mode(mode)
            out.zero_()
            t = timeit(lambda: lib.dim_vq_gather(idx.data_ptr(), E.data_ptr(), out.data_ptr(), N, 128, 512, None, s))
            gbs = N * 520 / t / 1e9
            print(json.dumps({"kernel": "vq_gather", "impl": name, "codes": N, "us": t * 1e6, "GB/s": gbs,
                              "frac_of_hbm_peak": gbs / PEAK, "peak": PEAK, "algorithmic_bytes_per_code": 520,
                              "bit_exact_vs_index_select": bool(torch.equal(out, ref))}))
        lib.dim_debug_vq_gather_mode(-1)
        del out, ref
    import ctypes as C
    lib.dim_debug_vq_argmin_impl.argtypes = [C.c_int]
    tc_peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json"))).get("bf16_tflops", 1675.8)
    if os.environ.get("VQ_TC_DBG"):                # timing ablations of the tensor-core kernel (results wrong by construction)
        N = 1 << 20
        z = (torch.randn(N, 128, generator=g) * 0.7).cuda()
        o = torch.empty(N, dtype=torch.int64, device="cuda")
        for dbg in [int(x) for x in os.environ["VQ_TC_DBG"].split(",")]:
            lib.dim_debug_vq_argmin_impl(dbg << 8)
            t = timeit(lambda: lib.dim_vq_argmin(z.data_ptr(), E.data_ptr(), o.data_ptr(), N, 128, 512, s))
            print(json.dumps({"ablation": dbg, "us": t * 1e6}))
        lib.dim_debug_vq_argmin_impl(0)
        return
    for N in ((1 << 20,) if NCU else (76800, 1 << 20, 1 << 22)):
        z = (torch.randn(N, 128, generator=g) * 0.7).cuda()
        o = torch.empty(N, dtype=torch.int64, device="cuda")
        res = {}
        for impl, name in ((1, "vq_argmin_f32 (exact FFMA)"), (0, "vq_argmin_tc (fp16 tcgen05 shortlist + exact fp32 re-rank)")):
            lib.dim_debug_vq_argmin_impl(impl)
            t = timeit(lambda: lib.dim_vq_argmin(z.data_ptr(), E.data_ptr(), o.data_ptr(), N, 128, 512, s))
            res[impl] = o.clone()
            gbs = N * 520 / t / 1e9
            print(json.dumps({"kernel": name, "tokens": N, "us": t * 1e6, "GB/s": gbs, "frac_of_hbm_peak": gbs / PEAK,
                              "TFLOP/s": N * 131072 / t / 1e12, "frac_of_tensor_peak": N * 131072 / t / 1e12 / tc_peak if impl == 0 else None,
                              "algorithmic_bytes_per_token": 520,
                              "note": "time includes the 2 small codebook-preparation launches" if impl == 0 else ""}))
        lib.dim_debug_vq_argmin_impl(0)
        print(json.dumps({"check": "tensor-core indices == exact-kernel indices", "tokens": N, "equal": bool(torch.equal(res[0], res[1]))}))


if __name__ == "__main__":
    main()
