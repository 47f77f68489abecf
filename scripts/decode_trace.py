"""Per-phase time of the persistent decode kernel (CTA 0's view, summed over the steps of one generate call).
    python scripts/decode_trace.py [B] [T] [precision: bf16|fp32_tc]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import dim_b200  # noqa: E402
from dim_b200 import _lib  # noqa: E402
from dim_b200.engine import PREC_BF16, PREC_FP32_TC, Handle, SLMFTEngine  # noqa: E402
from dim_b200.schema import S2SConfig  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
T = int(sys.argv[2]) if len(sys.argv) > 2 else 300
precs = [sys.argv[3]] if len(sys.argv) > 3 else ["bf16", "fp32_tc"]
h = Handle()
h.register(dim_b200.synth.make_slmft_state_dict(131))
for prec in precs:
    s2s = SLMFTEngine(h, S2SConfig(), precision=PREC_BF16 if prec == "bf16" else PREC_FP32_TC)
    c = dim_b200.synth.make_clips(B, T, seed=1)
    ctx = s2s.context(c["v_speaker"].cuda(), c["v_audio"].cuda(), c["mask"].cuda())
    m = c["mask"].cuda()
    prompt = torch.zeros(B, dtype=torch.int64).cuda()
    u = torch.rand(B, T - 1, generator=torch.Generator().manual_seed(1)).cuda()
    for _ in range(2):
        s2s.generate(ctx, m, prompt, T - 1, temperature=1.0, uniforms=u)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    s2s.generate(ctx, m, prompt, T - 1, temperature=1.0, uniforms=u)
    e1.record()
    torch.cuda.synchronize()
    total = e0.elapsed_time(e1)
    _lib.decode_trace_enable(True)
    s2s.generate(ctx, m, prompt, T - 1, temperature=1.0, uniforms=u)
    tr = _lib.decode_trace_collect()
    _lib.decode_trace_enable(False)
    steps = T - 1
    print(f"== {prec}  B={B} T={T}: generate {total:.1f} ms = {1e3 * total / steps:.1f} us/step; traced sum {sum(ms for _, ms in tr):.1f} ms")
    agg = {}
    for i, (k, ms) in enumerate(tr):
        print(f"  phase {i:2d} {k:24s} {1e3 * ms / steps:8.2f} us/step")
        agg[k] = agg.get(k, 0.0) + ms
    for k, ms in sorted(agg.items(), key=lambda kv: -kv[1]):
        print(f"  {k:24s} {1e3 * ms / steps:8.1f} us/step  {100 * ms / sum(agg.values()):5.1f} %")
    del s2s
    torch.cuda.empty_cache()
