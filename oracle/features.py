"""Oracle: CPU restatement of the reference's two audio-feature resamplers (test infrastructure).

PARITY PINNED: tests/golden/resample_reference.pt holds outputs of the reference functions themselves, executed from
/root/reference/code via tests/golden/make_resample_golden.py; tests/test_oracle.py checks this restatement against them.

  window_mean   code/vico_preprocessing.py:7-19   downsample_mean(array, factor=0.6): new_t = int(t*factor), window = int(t/new_t),
                out[i] = mean(array[i*window : i*window+window]) in float64  (window == 1 at factor 0.6: the "50 -> 30 fps" step
                of the ViCo preprocessing keeps the first 60 % of the frames unchanged)
  linear        code/dataset/l2l.py:23-29         F.interpolate(size=new_t, mode='linear', align_corners=True) over time
"""
import numpy as np
import torch


def window_mean(array, factor=0.6):
    t, d = array.shape
    new_t = int(t * factor)
    window = int(t / new_t)
    out = np.zeros((new_t, d))
    for i in range(new_t):
        out[i] = np.mean(array[i * window:i * window + window], axis=0)
    return out


def linear(array, new_t):
    x = torch.from_numpy(np.asarray(array)).unsqueeze(0).permute(0, 2, 1)
    y = torch.nn.functional.interpolate(x, size=(new_t), mode="linear", align_corners=True)
    return y.permute(0, 2, 1).squeeze(0).numpy()
