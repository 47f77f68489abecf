"""Run the prefill half of one bench step (VQ encode of 256 x 300 listener frames + SLMFT context) twice in plain-bf16 mode: a target
for `ncu -k regex:attn_prefill_mma` / `-k regex:gemm_bf16_tcgen05` captures of the prefill kernels at BASELINE shapes."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dim_b200  # noqa: E402
from dim_b200.engine import PREC_BF16, PREC_FP32_TC, Handle, SLMFTEngine, VQEngine  # noqa: E402
from dim_b200.schema import S2SConfig, VQConfig  # noqa: E402


def main():
    B, T = int(sys.argv[1]) if len(sys.argv) > 1 else 256, int(sys.argv[2]) if len(sys.argv) > 2 else 300
    sd = dim_b200.synth.make_slmft_state_dict(131)
    h = Handle(); h.register(sd)
    s2s = SLMFTEngine(h, S2SConfig(), precision=PREC_BF16)
    vq = VQEngine(h, VQConfig(), prefix="listener_vq.", precision=PREC_FP32_TC)
    c = dim_b200.synth.make_clips(B, T, seed=1)
    vs, va, vl, m = (c[k].cuda() for k in ("v_speaker", "v_audio", "v_listener", "mask"))
    for _ in range(2):
        vq.encode(vl)
        s2s.context(vs, va, m)
    torch.cuda.synchronize()
    print("done")


if __name__ == "__main__":
    main()
