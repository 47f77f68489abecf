"""SURVEY 8(f).1 -- the 10-sample best-of-N loop of x_engine_pt.evaluate_test_epoch (code/x_engine_pt.py:255-270) on one
batch of ViCo-shape clips: (a) the reference's structure, 10 model calls + host-side numpy/scipy selection;
(b) one multi-sample pass (encoders and cross-attention K/V once per clip, 10 x B decode rows sharing each clip's K/V)
+ device-side selection.  Prints one JSON line per arm: selected listener frames per second (B * (T-1) frames per batch)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from scipy import linalg

import dim_b200
from dim_b200.compat_api import best_of_n, slmft_forward_val, slmft_forward_val_samples
from dim_b200.engine import PREC_BF16, PREC_FP32_TC, Handle, SLMFTEngine, VQEngine
from dim_b200.schema import S2SConfig, VQConfig


def host_select(preds, tgt, lengths):
    keep = []
    for b in range(tgt.shape[0]):
        n = lengths[b]
        t = tgt[b, :n].astype(np.float64)
        mu1, s1 = t.mean(0), np.cov(t, rowvar=False)
        best, arg = float("inf"), None
        for p in preds:
            x = p[b, :n].astype(np.float64)
            mu2, s2 = x.mean(0), np.cov(x, rowvar=False)
            cm = linalg.sqrtm(s1.dot(s2))
            cm = cm.real if np.iscomplexobj(cm) else cm
            fd = (mu1 - mu2).dot(mu1 - mu2) + np.trace(s1) + np.trace(s2) - 2 * np.trace(cm)
            if fd < best:
                best, arg = fd, x
        keep.append(arg)
    return keep


def main():
    B, T, S = int(os.environ.get("B", 64)), int(os.environ.get("T", 300)), 10      # >= 64 clips: both arms on the tensor cores
    bf16 = os.environ.get("PREC", "bf16") == "bf16"
    h = Handle()
    h.register(dim_b200.synth.make_slmft_state_dict(131))
    s2s = SLMFTEngine(h, S2SConfig(), precision=PREC_BF16 if bf16 else PREC_FP32_TC)
    vq = VQEngine(h, VQConfig(), prefix="listener_vq.", precision=PREC_FP32_TC)
    c = dim_b200.synth.make_clips(B, T, seed=5)
    d = {k: c[k].cuda() for k in ("v_speaker", "v_listener", "v_audio", "mask")}
    u = torch.rand(B, S, T - 1, generator=torch.Generator().manual_seed(1)).cuda()
    lengths = [T - 1] * B
    tgt_np = c["v_listener"][:, 1:].numpy()

    def arm_reference_structure():
        preds = []
        for j in range(S):
            _, _, p = slmft_forward_val(s2s, vq, d["v_speaker"], d["v_listener"], d["v_audio"], d["mask"], uniforms=u[:, j].contiguous())
            preds.append(p.cpu().numpy())
        return host_select(preds, tgt_np, lengths)

    def arm_one_pass():
        pred, _ = slmft_forward_val_samples(s2s, vq, d["v_speaker"], d["v_listener"], d["v_audio"], d["mask"], S, uniforms=u)
        picked, chosen, _ = best_of_n(pred, d["v_listener"][:, 1:], lengths)
        return [p.cpu().numpy() for p in picked]

    results = {}
    for name, fn in (("10 model calls + host numpy/scipy selection (reference structure)", arm_reference_structure),
                     ("one multi-sample pass + device-side selection", arm_one_pass)):
        fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n = 3
        for _ in range(n):
            out = fn()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / n
        results[name] = out
        print(json.dumps({"arm": name, "clips": B, "frames_per_clip": T, "samples_per_clip": S, "precision": "bf16" if bf16 else "fp32_tc",
                          "ms_per_batch": 1e3 * dt, "selected_frames_per_s": B * (T - 1) / dt,
                          "generated_frames_per_s": B * S * (T - 1) / dt}), flush=True)
    a, b = list(results.values())
    same = all(np.array_equal(x, y) for x, y in zip(a, b))
    print(json.dumps({"both_arms_select_identical_frames": bool(same)}))


if __name__ == "__main__":
    main()
