"""Persistent decode kernel vs per-kernel path: which rows / steps differ?  python scripts/mk_check.py B T [bf16|fp32_tc]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import dim_b200  # noqa: E402
from dim_b200 import _lib  # noqa: E402
from dim_b200.engine import PREC_BF16, PREC_FP32_TC, Handle, SLMFTEngine  # noqa: E402
from dim_b200.schema import S2SConfig  # noqa: E402

B, T = int(sys.argv[1]), int(sys.argv[2])
prec = sys.argv[3] if len(sys.argv) > 3 else "fp32_tc"
h = Handle()
h.register(dim_b200.synth.make_slmft_state_dict(131))
s2s = SLMFTEngine(h, S2SConfig(), precision=PREC_BF16 if prec == "bf16" else PREC_FP32_TC)
c = dim_b200.synth.make_clips(B, T, seed=300 + B, ragged=True)
ctx = s2s.context(c["v_speaker"].cuda(), c["v_audio"].cuda(), c["mask"].cuda())
m = c["mask"].cuda()
prompt = torch.randint(0, 512, (B,), generator=torch.Generator().manual_seed(B)).cuda()
_lib.decode_set_impl(1)
rc, rl = s2s.generate(ctx, m, prompt, T - 1, return_logits=True)
_lib.decode_set_impl(0)
for rep in range(3):
    cc, ll = s2s.generate(ctx, m, prompt, T - 1, return_logits=True)
    d = (ll - rl).abs().amax(dim=2)                      # (B, steps)
    bad = (d > (0.05 if prec == "bf16" else 2e-4))
    first_bad_step = torch.where(bad.any(1), bad.float().argmax(1), torch.full((B,), -1, device=d.device))
    rows = bad.any(1).nonzero().flatten().tolist()
    print(f"rep {rep}: B={B} T={T} {prec}: rows with a logit difference: {len(rows)} {rows[:24]}; first bad step per such row: "
          f"{[int(first_bad_step[r]) for r in rows[:24]]}; max diff step0 {float(d[:, 0].max()):.2e}")
