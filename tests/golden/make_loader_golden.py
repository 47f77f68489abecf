"""Generate tests/golden/loader_reference.pt by running the REAL reference loaders (/root/reference/code/dataset/data_loader.py:
ViCoDataset + pad_collate, LmListenerDataset + pad_collate_lm; /root/reference/code/dataset/l2l.py: LmListenerDataset with HuBERT
features) on synthetic fixtures written by dim_b200.l2l_artifacts (seeds stored).  Run in the build container only:
    python tests/golden/make_loader_golden.py
Each batch is stored as shapes, lengths, names, float64 sums and a strided sample of the padded tensors (the full 1024-frame chunks
would be tens of MB)."""
import os
import subprocess
import sys
import tempfile

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
COMPAT = os.path.join(ROOT, "dyadic-interaction-modeling_b200", "compat")
REF = "/root/reference/code"
sys.path.insert(0, ROOT)

CODE = r'''
import sys, torch
_load = torch.load                      # the reference predates torch.load's weights_only=True default (its .pth files hold numpy arrays)
torch.load = lambda *a, **k: _load(*a, **{**k, "weights_only": False})
from torch.utils import data
out = {}
def digest(t):
    f = t.reshape(-1)
    return dict(shape=tuple(t.shape), sum=float(f.double().sum()), sample=f[::97].clone())
import dataset.data_loader as DL
ds = DL.ViCoDataset("../data/vico_processed_30fps", "../data/RLD_data.csv", mode="test")
b = []
for src, tgt, lens, ids, names in data.DataLoader(ds, batch_size=3, shuffle=False, collate_fn=DL.pad_collate):
    b.append(dict(src=digest(src), tgt=digest(tgt), lens=list(lens), speaker_ids=ids[0].tolist(), listener_ids=ids[1].tolist(),
                  names=[n.split("/")[-1] for n in names]))
out["vico"] = b
ds = DL.LmListenerDataset("../data/lm", mode="test")
b = []
for src, tgt, xl, yl, names in data.DataLoader(ds, batch_size=2, shuffle=False, collate_fn=DL.pad_collate_lm):
    b.append(dict(src=digest(src), tgt=digest(tgt), lens=list(xl), names=list(names)))
out["lm_zeros"] = b
import dataset.l2l as L2L
ds = L2L.LmListenerDataset("../data/lm", mode="test")
b = []
for src, tgt, xl, yl, names in data.DataLoader(ds, batch_size=2, shuffle=False, collate_fn=L2L.pad_collate_lm):
    b.append(dict(src=digest(src), tgt=digest(tgt), lens=list(xl), names=list(names)))
out["lm_hubert"] = b
torch.save(out, sys.argv[1])
print("ok", {k: len(v) for k, v in out.items()})
'''


def write_fixtures(root):
    import dim_b200
    from dim_b200 import l2l_artifacts as A
    clips = dim_b200.synth.make_clips(5, 48, seed=11, ragged=True)
    clips = {k: (v.numpy() if hasattr(v, "numpy") else v) for k, v in clips.items()}
    A.write_vico_fixtures(os.path.join(root, "data"), clips, split=["test", "test", "train", "test", "test"])
    A.write_lm_listener_fixtures(os.path.join(root, "data", "lm"), mode="test", seed=5, lengths=(30, 10, 1100, 64))


def main():
    with tempfile.TemporaryDirectory() as tmp:
        write_fixtures(tmp)
        cwd = os.path.join(tmp, "code")
        os.makedirs(cwd)
        path = os.path.join(HERE, "loader_reference.pt")
        env = dict(os.environ, PYTHONPATH=os.pathsep.join([ROOT, COMPAT, REF, os.path.join(COMPAT, "_shims")]))
        r = subprocess.run([sys.executable, "-c", CODE, path], env=env, capture_output=True, text=True, cwd=cwd)
        print(r.stdout[-2000:], r.stderr[-3000:])
        assert r.returncode == 0
    out = torch.load(path, weights_only=False)
    out["fixtures"] = dict(vico=dict(batch=5, frames=48, seed=11, split=["test", "test", "train", "test", "test"]),
                           lm=dict(seed=5, lengths=(30, 10, 1100, 64)))
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
