"""SURVEY 8(f).2 -- SLMFT.forward(mode='train') forward pass (teacher forcing) on B ViCo-shape clips: listener VQ encode, speaker
encoders, teacher-forced decoder (self + cross attention over the whole sequence), logits, CE, argmax, VQ decode.
One JSON line per arithmetic mode: frames/s = B * (T-1) / time."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import dim_b200
from dim_b200.compat_api import draw_kv_mask, slmft_forward_train
from dim_b200.engine import PREC_BF16, PREC_FP32_TC, Handle, SLMFTEngine, VQEngine
from dim_b200.schema import S2SConfig, VQConfig

B, T = int(os.environ.get("B", 256)), int(os.environ.get("T", 300))
h = Handle()
h.register(dim_b200.synth.make_slmft_state_dict(131))
vq = VQEngine(h, VQConfig(), prefix="listener_vq.", precision=PREC_FP32_TC)
c = dim_b200.synth.make_clips(B, T, seed=5)
d = {k: c[k].cuda() for k in ("v_speaker", "v_listener", "v_audio", "mask")}
kv = draw_kv_mask((B, T - 1), 0.15, "cuda")
for name, prec in (("bf16", PREC_BF16), ("fp32_tc", PREC_FP32_TC)):
    s2s = SLMFTEngine(h, S2SConfig(), precision=prec)
    step = lambda: slmft_forward_train(s2s, vq, d["v_speaker"], d["v_listener"], d["v_audio"], d["mask"], kv_mask=kv)
    for _ in range(3):
        step()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(5):
        total, dd, pred = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(json.dumps({"path": "SLMFT.forward(mode='train') forward pass", "precision": name, "clips": B, "frames_per_clip": T,
                      "ms_per_batch": ms, "frames_per_s": B * (T - 1) / (ms * 1e-3), "l_ce": float(dd["l_ce_l"]),
                      "l_cont": float(dd["l_cont_l"])}), flush=True)
    del s2s
