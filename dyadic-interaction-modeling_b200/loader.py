"""GPU-side loaders for the reference's on-disk formats.

The files are read on the host (pickle / torch.load is host code by nature), but a batch then costs ONE packed pinned-memory
staging buffer, ONE host->device copy and ONE kernel (dim_assemble_batch) that writes the padded `src`, `tgt` and `mask` tensors --
instead of the reference's per-clip FloatTensor conversions, torch.cat, pad_sequence and a pageable copy of the padded batch
(dataset/data_loader.py:138-152, :429-439).  Outputs equal the reference loader + collate bit for bit.

  ViCoClips            dataset/data_loader.py:108-152 (ViCoDataset): RLD_data.csv + vico_processed_30fps/<id>.pkl
  LmListenerSegments   dataset/data_loader.py:210-245 (zeros audio) and dataset/l2l.py:31-76 (HuBERT features, linearly resampled to
                       the motion length on the device with dim_resample_features): segments_<mode>.pth, 1024-frame chunks
"""
from __future__ import annotations

import os
import pickle

import numpy as np
import torch

from . import _lib, ops

SENTIMENT = {"neutral": 0, "positive": 1, "negative": 2}


def _ptr(t):
    return None if t is None else t.data_ptr()


def assemble_batch(speaker, audio, listener, lengths, T=None, device="cuda", motion_dim=56, audio_dim=768):
    """Packed host clips -> padded device batch.  speaker / audio: list of (len_b, D) float32 arrays or None (None: ones / zeros,
    like the two reference loaders); listener: list of (len_b, motion_dim).  Returns (src (B,T,motion+audio), tgt (B,T,motion),
    mask (B,T) bool) on `device`."""
    B = len(listener)
    lengths = [int(n) for n in lengths]
    T = T or max(lengths)
    total = sum(lengths)
    widths = [(speaker, motion_dim), (audio, audio_dim), (listener, motion_dim)]
    cols = sum(w for a, w in widths if a is not None)
    stage = torch.empty(total * cols, dtype=torch.float32).pin_memory()        # one staging buffer: [speaker | audio | listener] blocks
    views, o = [], 0
    for arrs, w in widths:
        if arrs is None:
            views.append(None)
            continue
        v = stage[o:o + total * w].view(total, w)
        r = 0
        for a, n in zip(arrs, lengths):
            v[r:r + n] = torch.as_tensor(np.asarray(a[:n]), dtype=torch.float32)
            r += n
        views.append((o, total * w, w))
        o += total * w
    dev = torch.device(device)
    dstage = stage.to(dev, non_blocking=True)
    parts = [None if v is None else dstage[v[0]:v[0] + v[1]].view(total, v[2]) for v in views]
    offsets = torch.tensor([0] + list(np.cumsum(lengths)), dtype=torch.int64).to(dev, non_blocking=True)
    src = torch.empty(B, T, motion_dim + audio_dim, dtype=torch.float32, device=dev)
    tgt = torch.empty(B, T, motion_dim, dtype=torch.float32, device=dev)
    mask = torch.empty(B, T, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().dim_assemble_batch(_ptr(parts[0]), _ptr(parts[1]), _ptr(parts[2]), offsets.data_ptr(), B, T, motion_dim,
                                                  audio_dim, src.data_ptr(), tgt.data_ptr(), mask.data_ptr(),
                                                  torch.cuda.current_stream().cuda_stream), "dim_assemble_batch")
    return src, tgt, mask.bool()


class ViCoClips:
    """ViCoDataset (data_loader.py:108-152) + pad_collate (:429-439) + the batch split of x_engine_pt.py:246-249."""

    def __init__(self, data_path, meta_data_path, mode="train"):
        import pandas as pd
        meta = pd.read_csv(meta_data_path).values
        ids = [meta[i, 1] for i in range(len(meta)) if meta[i, 6] == mode]
        self.paths, self.clips = [], []
        for cid in ids:
            p = os.path.join(data_path, cid + ".pkl")
            if not os.path.exists(p):
                continue
            with open(p, "rb") as f:
                d = pickle.load(f)
            n = len(d["video_speaker"])
            if n == len(d["audio"]) == len(d["video_listener"]) and 1024 >= n >= 5:
                self.paths.append(p)
                self.clips.append(d)
        self.id2speaker = {meta[i, 1]: meta[i, 5] for i in range(len(meta))}
        self.id2listener = {meta[i, 1]: meta[i, 4] for i in range(len(meta))}
        self.id2sentiment = {meta[i, 1]: SENTIMENT[meta[i, 0]] for i in range(len(meta))}

    def __len__(self):
        return len(self.clips)

    def batches(self, batch_size, device="cuda", speaker_ones=True):
        """Yields (src, tgt, src_len, (speaker_ids, listener_ids), names, mask): the first five are what the reference DataLoader
        yields (shuffle=False), on the device; speaker_ones mirrors `torch.ones_like(video_feats_speaker)` (:147)."""
        for b0 in range(0, len(self.clips), batch_size):
            cl = self.clips[b0:b0 + batch_size]
            names = self.paths[b0:b0 + batch_size]
            lens = [len(d["video_listener"]) for d in cl]
            src, tgt, mask = assemble_batch(None if speaker_ones else [d["video_speaker"] for d in cl], [d["audio"] for d in cl],
                                            [d["video_listener"] for d in cl], lens, device=device)
            uid = [os.path.basename(p).split(".")[0] for p in names]
            ids = (torch.LongTensor([self.id2speaker[u] for u in uid]), torch.LongTensor([self.id2listener[u] for u in uid]))
            yield src, tgt, lens, ids, names, mask


class LmListenerSegments:
    """LmListenerDataset: `hubert=False` is dataset/data_loader.py:210-245 (audio = zeros), `hubert=True` is dataset/l2l.py:31-76
    (segments with 'hubert_feat', resampled to the motion length by linear interpolation, align_corners=True -- here on the device).
    Segments of >= 1024 frames are cut into 1024-frame chunks, shorter than 24 frames or with unequal p0/p1 lengths dropped."""

    def __init__(self, data_path, mode="train", hubert=False, device="cuda"):
        cur = torch.load(os.path.join(data_path, f"segments_{mode}.pth"), weights_only=False)
        self.items, self.hubert = [], hubert
        for i in range(len(cur)):
            it = cur[i]
            if hubert:
                if "hubert_feat" not in it or it["split_start_time"] == it["split_end_time"]:
                    continue
            n = len(it["p0_exp"])
            if not (n == len(it["p1_exp"]) and n >= 24):
                continue
            feat = None
            if hubert:                                        # l2l.py:23-29,45: (t,768) -> (n,768) on the device, back to the host list
                x = torch.as_tensor(np.asarray(it["hubert_feat"]), dtype=torch.float32).to(device)
                feat = ops.resample_linear(x.contiguous(), n).cpu().numpy()
            if n < 1024:
                self.items.append(self._item(it, slice(0, n), feat, it["fname"]))
            else:
                for j in range(n // 1024):
                    name = (str(i) + "**" + str(it["split_start_time"]) + "**" + str(it["split_end_time"]) + "**" + it["fname"]) if hubert \
                        else it["fname"]
                    self.items.append(self._item(it, slice(j * 1024, (j + 1) * 1024), feat, name))

    @staticmethod
    def _item(it, sl, feat, name):
        sp = np.concatenate([np.asarray(it["p1_pose"])[sl], np.asarray(it["p1_exp"])[sl]], axis=1).astype(np.float32)
        li = np.concatenate([np.asarray(it["p0_pose"])[sl], np.asarray(it["p0_exp"])[sl]], axis=1).astype(np.float32)
        return dict(speaker=sp, listener=li, audio=None if feat is None else feat[sl], fname=name)

    def __len__(self):
        return len(self.items)

    def batches(self, batch_size, device="cuda"):
        """Yields (src, tgt, x_lens, y_lens, names, mask) like pad_collate_lm (data_loader.py:441-449 / l2l.py:78-85), on the device."""
        for b0 in range(0, len(self.items), batch_size):
            it = self.items[b0:b0 + batch_size]
            lens = [len(d["listener"]) for d in it]
            audio = [d["audio"] for d in it] if self.hubert else None
            src, tgt, mask = assemble_batch([d["speaker"] for d in it], audio, [d["listener"] for d in it], lens, device=device)
            yield src, tgt, lens, list(lens), [d["fname"] for d in it], mask
