"""Golden vectors for the feature resamplers, produced by the REFERENCE functions themselves.

The two modules cannot be imported offline (pickle5 / s3prl / torchaudio side imports, and module-level code that reads the
datasets), so each function's source is cut out of the reference file with `ast` and executed unmodified:
  code/vico_preprocessing.py:7-19   downsample_mean(array, factor=0.6)   (window mean)
  code/dataset/l2l.py:23-29         downsample_mean(array, new_t)        (linear, align_corners=True)
Run in the build container (needs /root/reference):  python tests/golden/make_resample_golden.py"""
import ast
import os

import numpy as np
import torch

REF = "/root/reference/code"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "resample_reference.pt")


def cut(path, name):
    src = open(path).read()
    node = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == name)
    ns = {"np": np, "torch": torch}
    exec(compile(ast.Module(body=[node], type_ignores=[]), path, "exec"), ns)
    return ns[name]


def main():
    win = cut(os.path.join(REF, "vico_preprocessing.py"), "downsample_mean")
    lin = cut(os.path.join(REF, "dataset", "l2l.py"), "downsample_mean")
    g = np.random.default_rng(0)
    cases = {}
    for name, t, d in (("vico_120x64", 120, 64), ("short_7x8", 7, 8), ("odd_333x16", 333, 16)):
        x = g.standard_normal((t, d)).astype(np.float32)
        cases[name] = {"x": torch.from_numpy(x), "window_mean_0.6": torch.from_numpy(win(x)),           # float64, like the reference
                       "window_mean_0.25": torch.from_numpy(win(x, factor=0.25)),
                       "linear_new_t": {n: torch.from_numpy(lin(x, n)) for n in (int(t * 0.6), t, 2 * t + 1, 3)}}
    torch.save({"cases": cases, "source": "reference functions executed via ast (see this script)"}, OUT)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
