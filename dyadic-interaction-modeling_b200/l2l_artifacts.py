"""Artefacts that connect the eval loop to the reference's offline metrics script (last bullet; BASELINE configs[4]).

`test_s2s_pretrain.py:78-84` pickles `{'y_true','y_pred','data_ids'}`, but `test_l2l.py` (a pure metrics script, no model code)
reads something else: `../data/l2l_vico_predictions.pkl` = dict clip id -> (L,56) array whose columns are ordered expression(50)
then pose(6) (it rotates 50:56 to the front, test_l2l.py:80-82), `../data/RLD_data.csv` (positional columns: [1] clip id,
[2] listener file, [3] speaker file, test_l2l.py:20-26) and per-frame EMOCA directories
`../data/vico_dataset/emoca/<file>/EMOCA_v2_lr_mse_20/0*/{exp,pose,detail}.npy` (:36-56).  The conversion is not in the reference.
This module writes both sides from in-memory arrays so that the UNCHANGED script can be run on the B200 path's output
(tests/test_l2l_script.py does exactly that and compares its printed metrics with dim_b200.metrics)."""
from __future__ import annotations

import os
import pickle

import numpy as np


def write_l2l_predictions(path, data_ids, y_preds):
    """evaluate_test_epoch's (data_ids, y_preds) -> the pickle test_l2l.py loads.  y_preds[i]: (L,56) with the model's column
    order pose(0:6) | expression(6:56); stored as expression | pose, which the script rotates back."""
    d = {}
    for cid, p in zip(data_ids, y_preds):
        p = np.asarray(p)
        d[cid] = np.concatenate([p[:, 6:56], p[:, 0:6]], axis=1)
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    with open(path, "wb") as f:
        pickle.dump(d, f)
    return d


def write_emoca_dir(root, file_name, frames):
    """frames (L,56) pose|expression -> <root>/<file_name>/EMOCA_v2_lr_mse_20/NNNNNN_000/{pose,exp,detail}.npy, one directory per
    frame (the layout test_l2l.py:36-56 walks; entries must start with '0')."""
    base = os.path.join(root, file_name, "EMOCA_v2_lr_mse_20")
    for i, fr in enumerate(np.asarray(frames)):
        d = os.path.join(base, f"{i:06d}_000")
        os.makedirs(d, exist_ok=True)
        np.save(os.path.join(d, "pose.npy"), fr[0:6])
        np.save(os.path.join(d, "exp.npy"), fr[6:56])
        np.save(os.path.join(d, "detail.npy"), np.zeros(128, dtype=fr.dtype))


def write_l2l_fixtures(data_root, data_ids, first, second, split="test"):
    """Synthetic `../data` tree for test_l2l.py: RLD_data.csv + EMOCA directories.  `first[i]` is stored under the file named in
    CSV column [3] -- the script reads it as `video_feats` and uses it as the ground truth -- and `second[i]` under column [2]'s
    file, which the script uses as `x` (its variable names are swapped with respect to the CSV's; kept as is)."""
    import pandas as pd
    rows = []
    emoca = os.path.join(data_root, "vico_dataset", "emoca")
    for i, cid in enumerate(data_ids):
        f3, f2 = f"{cid}_a", f"{cid}_b"
        write_emoca_dir(emoca, f3, first[i])
        write_emoca_dir(emoca, f2, second[i])
        rows.append(["neutral", cid, f2, f3, f"L{i}", f"S{i}", split])
    os.makedirs(data_root, exist_ok=True)
    pd.DataFrame(rows, columns=["sentiment", "uuid", "listener", "speaker", "listener_id", "speaker_id", "split"]).to_csv(
        os.path.join(data_root, "RLD_data.csv"), index=False)


def write_vico_fixtures(data_root, clips, ids=None, split="test"):
    """Synthetic ViCo inputs in the layout the reference's loader reads (dataset/data_loader.py:110-152, get_vico_dataloaders
    :461-475): `<data_root>/RLD_data.csv` (positional columns [0] sentiment, [1] clip id, [4] listener id, [5] speaker id,
    [6] split) and `<data_root>/vico_processed_30fps/<clip id>.pkl` = {'video_speaker' (T,56), 'audio' (T,768),
    'video_listener' (T,56)}.  clips: dict with v_speaker / v_audio / v_listener (B,T,.) tensors or arrays and `lengths` (B).
    split: one name for all clips or a list per clip (get_vico_dataloaders builds a 'train' AND a 'test' dataset and fails on an
    empty one)."""
    import pandas as pd
    B = len(clips["v_listener"])
    ids = ids or [f"vico{i:03d}" for i in range(B)]
    d = os.path.join(data_root, "vico_processed_30fps")
    os.makedirs(d, exist_ok=True)
    rows = []
    for i, cid in enumerate(ids):
        n = int(clips["lengths"][i]) if "lengths" in clips else len(clips["v_listener"][i])
        rec = {"video_speaker": np.asarray(clips["v_speaker"][i][:n], dtype=np.float32),
               "audio": np.asarray(clips["v_audio"][i][:n], dtype=np.float32),
               "video_listener": np.asarray(clips["v_listener"][i][:n], dtype=np.float32)}
        with open(os.path.join(d, cid + ".pkl"), "wb") as f:
            pickle.dump(rec, f)
        rows.append(["neutral", cid, f"{cid}_l", f"{cid}_s", i, 100 + i, split if isinstance(split, str) else split[i]])
    pd.DataFrame(rows, columns=["sentiment", "uuid", "listener", "speaker", "listener_id", "speaker_id", "split"]).to_csv(
        os.path.join(data_root, "RLD_data.csv"), index=False)
    return ids


def make_lm_listener_segments(seed=0, lengths=(30, 10, 1100, 64), hubert=True):
    """Synthetic `segments_<mode>.pth` content for the LM-Listener loaders (dataset/data_loader.py:210-227, dataset/l2l.py:31-60): a list
    of dicts with p0/p1 expression (n,50) and pose (n,6) arrays, `fname`, split times and (hubert) a (t,768) feature array at ~50 fps
    for n frames at 30 fps.  Lengths < 24 are dropped by the loaders, >= 1024 are cut into 1024-frame chunks."""
    g = np.random.default_rng(seed)
    segs = []
    for i, n in enumerate(lengths):
        d = {"p0_exp": g.standard_normal((n, 50)).astype(np.float32) * 0.3, "p1_exp": g.standard_normal((n, 50)).astype(np.float32) * 0.3,
             "p0_pose": g.standard_normal((n, 6)).astype(np.float32) * 0.1, "p1_pose": g.standard_normal((n, 6)).astype(np.float32) * 0.1,
             "fname": f"seg_{i:03d}", "split_start_time": float(i), "split_end_time": float(i) + n / 30.0}
        if hubert:
            d["hubert_feat"] = g.standard_normal((max(2, int(n * 50 / 30)), 768)).astype(np.float32)
        segs.append(d)
    return segs


def write_lm_listener_fixtures(data_path, mode="test", **kw):
    import torch
    os.makedirs(data_path, exist_ok=True)
    segs = make_lm_listener_segments(**kw)
    torch.save(segs, os.path.join(data_path, f"segments_{mode}.pth"))
    return segs
