"""Tuning aid: CPU enqueue time vs GPU time of dim_slmft_generate (B=256, T=300) under DIM_NO_GROUPS / DIM_NO_GRAPH."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dim_b200
from dim_b200.engine import PREC_BF16, PREC_FP32_TC, Handle, SLMFTEngine
from dim_b200.schema import S2SConfig

prec = PREC_BF16 if (len(sys.argv) > 1 and sys.argv[1] == "bf16") else PREC_FP32_TC
B, T = int(os.environ.get("B", 256)), int(os.environ.get("T", 300))
h = Handle(); h.register(dim_b200.synth.make_slmft_state_dict(131))
s2s = SLMFTEngine(h, S2SConfig(), precision=prec)
c = dim_b200.synth.make_clips(B, T, seed=1)
ctx = s2s.context(c["v_speaker"].cuda(), c["v_audio"].cuda(), c["mask"].cuda())
m = c["mask"].cuda(); prompt = torch.zeros(B, dtype=torch.int64, device="cuda"); u = torch.rand(B, T - 1, device="cuda")
for _ in range(2):
    s2s.generate(ctx, m, prompt, T - 1, temperature=1.0, uniforms=u)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter(); e0.record()
s2s.generate(ctx, m, prompt, T - 1, temperature=1.0, uniforms=u)
e1.record(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"groups={'off' if os.environ.get('DIM_NO_GROUPS') else 'on'} graph={'off' if os.environ.get('DIM_NO_GRAPH') else 'on'} "
      f"B={B} T={T}: cpu enqueue {1e3*(t1-t0):.1f} ms, gpu {e0.elapsed_time(e1):.1f} ms, wall {1e3*(t2-t0):.1f} ms")
