"""Timing of the LSTM + vertex head (EmocaConverter / SpeakerSLMFT, seq2seq_pretrain.py:789-807) on the GPU: CUDA events around
the calls, after warm-up; torch.nn.LSTM (cuDNN) timed beside it as the library baseline."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dim_b200  # noqa: E402
from dim_b200 import compat_api, ops  # noqa: E402


def cuda_ms(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def main():
    torch.backends.cudnn.allow_tf32 = False      # cuDNN RNNs default to TF32 products; compare fp32 with fp32
    ref = torch.nn.LSTM(56, 384, 2, batch_first=True, bidirectional=True).cuda().eval()
    params = {k: v.detach() for k, v in ref.named_parameters()}
    for B, T in ((1, 27), (1, 300), (8, 138), (64, 300), (256, 300)):
        x = torch.randn(B, T, 56, device="cuda")
        with torch.no_grad():
            t_ref = cuda_ms(lambda: ref(x))
        t = cuda_ms(lambda: ops.lstm(x, params, hidden=384))
        print(f"lstm B={B:4d} T={T:4d}: {t:8.3f} ms  ({t / (2 * T) * 1e3:6.2f} us per step and layer)   cuDNN {t_ref:8.3f} ms")
        if B >= 64:
            from dim_b200 import _lib
            _lib.profile_enable(True)
            ops.lstm(x, params, hidden=384)
            torch.cuda.synchronize()
            for e in _lib.profile_collect():
                print(f"      {e['category']:<14s} launches {e['launches']:3d}  {e['ms']:8.3f} ms")
            _lib.profile_enable(False)
    sd = {k: v.cuda() for k, v in dim_b200.synth.make_emoca_converter_state_dict(1).items()}
    for M in (26, 299):
        h = torch.randn(M, 768, device="cuda")
        t = cuda_ms(lambda: ops.linear_ragged(h, sd["vertice_map_reverse.2.weight"], sd["vertice_map_reverse.2.bias"]))
        gb = (70110 * 768 + M * 70110 + M * 768) * 4 / 1e9
        print(f"linear 768->70110 M={M}: {t:7.3f} ms  {gb / t * 1e3:7.1f} GB/s  {2 * M * 70110 * 768 / t / 1e9:7.2f} TFLOP/s")
        v = torch.randn(M, 70110, device="cuda")
        t = cuda_ms(lambda: ops.linear_ragged(v, sd["vertice_mapping.0.weight"], sd["vertice_mapping.0.bias"], act=1, slope=0.2))
        gb = (70110 * 56 + M * 70110 + M * 56) * 4 / 1e9
        print(f"linear 70110->56 M={M}: {t:7.3f} ms  {gb / t * 1e3:7.1f} GB/s")


if __name__ == "__main__":
    main()
