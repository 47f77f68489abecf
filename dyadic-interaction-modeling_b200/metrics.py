"""The listener-motion metric suite of the reference's eval scripts, on the tensors' device.

Reference: code/mymetrics.py:7-88 (`print_metrics`) over code/metrics/eval_utils.py -- FD and paired FD (P-FD) per clip for the
pose (columns 0:6) and expression (6:56) halves, MSE, SID (k-means histogram entropy), variance, residual Pearson correlation
(rPCC) and STS.  The reference computes them on the host with numpy / scipy / sklearn (STS as a pure-Python double loop); here
everything after the k-means FIT runs as tensor ops on the device the predictions already live on, so an eval epoch copies back
a dozen scalars instead of every prediction.  Values equal the reference's to fp64 round-off (tests/test_metrics.py).

The k-means fit itself (`KMeans(n_clusters=k, random_state=0, n_init='auto')`, eval_utils.py:63) stays sklearn on the host -- it
is a seeded iterative fit on the ground truth only; the assignment of the predictions to the fitted centroids is the same
nearest-neighbour search as the VQ codebook lookup and runs on the codebook-argmin kernel when the tensors are on a GPU.
"""
from __future__ import annotations

import math

import torch

from .compat_api import frechet_distance_torch

POSE, EXP = slice(0, 6), slice(6, 56)


def _cat(seq):
    return torch.cat([t.double() for t in seq], dim=0)


def fd_per_clip(gt, pred, cols):
    """mean over clips of FD(gt[i][:, cols], pred[i][:, cols])   (mymetrics.py:11-27)"""
    return torch.stack([frechet_distance_torch(g[:, cols], p[:, cols]) for g, p in zip(gt, pred)]).mean()


def pfd_per_clip(gt, pred, x, cols):
    """paired FD: the speaker's columns are concatenated in front of the listener's   (mymetrics.py:29-43)"""
    vals = [frechet_distance_torch(torch.cat([s[:, cols], g[:, cols]], -1), torch.cat([s[:, cols], p[:, cols]], -1))
            for g, p, s in zip(gt, pred, x)]
    return torch.stack(vals).mean()


def mse_per_clip(gt, pred, cols):
    return torch.stack([((g[:, cols].double() - p[:, cols].double()) ** 2).mean() for g, p in zip(gt, pred)]).mean()


def pearson(a, b):
    a, b = a.double().reshape(-1), b.double().reshape(-1)
    a, b = a - a.mean(), b - b.mean()
    return (a * b).sum() / torch.sqrt((a * a).sum() * (b * b).sum())


def sts(x, y, timestep=0.1):
    """eval_utils.sts (eval_utils.py:85-91): sqrt(sum over dims and frames of ((dx - dy)^2 / timestep)), differences taken along
    the CONCATENATED sequence (clip boundaries included, like the reference)."""
    dx, dy = x.double()[1:] - x.double()[:-1], y.double()[1:] - y.double()[:-1]
    return torch.sqrt((((dx - dy) ** 2) / timestep).sum())


def nearest_centroid(points, centroids):
    """KMeans.predict: index of the nearest centroid (squared Euclidean distance).  On a GPU this is the VQ codebook-argmin kernel
    (features zero-padded to its 64-wide rows: padding changes no distance)."""
    if points.is_cuda and points.shape[1] <= 64:
        from . import ops
        D = 64
        z = torch.zeros(points.shape[0], D, dtype=torch.float32, device=points.device)
        z[:, :points.shape[1]] = points.float()
        E = torch.zeros(centroids.shape[0], D, dtype=torch.float32, device=points.device)
        E[:, :centroids.shape[1]] = centroids.float()
        return ops.vq_argmin(z, E)
    return torch.cdist(points.double(), centroids.double()).argmin(dim=1)


def sid(gt, pred, kind="exp", centroids=None):
    """eval_utils.calcuate_sid (eval_utils.py:49-83): entropy (bits) of the histogram of the predictions over k-means clusters
    fitted on the ground truth; k = 40 (exp) / 20 (pose).  `centroids` may carry a fit made elsewhere."""
    cols, k = (EXP, 40) if kind == "exp" else (POSE, 20)
    g, p = _cat(gt)[:, cols], _cat(pred)[:, cols]
    if centroids is None:
        from sklearn.cluster import KMeans
        km = KMeans(n_clusters=k, random_state=0, n_init="auto").fit(g.cpu().numpy())
        centroids = torch.from_numpy(km.cluster_centers_).to(p.device)
    lab = nearest_centroid(p, centroids.to(p.device))
    hist = torch.bincount(lab, minlength=k).double()
    hist = hist / hist.sum()
    return -(hist * torch.log2(hist + 1e-6)).sum()


@torch.no_grad()
def metrics_suite(y_true, y_pred, x, with_sid=True):
    """Everything `print_metrics(y_true, y_pred, x)` prints, as a dict of Python floats (sid_* as (pred, gt) pairs, var_* as
    (gt, pred) pairs, like the printed lines).  y_true / y_pred / x: lists of (n_i, 56) tensors on one device."""
    out = {"fid_pose": fd_per_clip(y_true, y_pred, POSE), "fid_exp": fd_per_clip(y_true, y_pred, EXP),
           "pfid_pose": pfd_per_clip(y_true, y_pred, x, POSE), "pfid_exp": pfd_per_clip(y_true, y_pred, x, EXP),
           "mse_pose": mse_per_clip(y_true, y_pred, POSE), "mse_exp": mse_per_clip(y_true, y_pred, EXP)}
    g, p, s = _cat(y_true), _cat(y_pred), _cat(x)[:, :56]
    out["var_pose"] = (g[:, POSE].var(unbiased=False), p[:, POSE].var(unbiased=False))
    out["var_exp"] = (g[:, EXP].var(unbiased=False), p[:, EXP].var(unbiased=False))
    out["rpcc_pose"] = (pearson(g[:, POSE], s[:, POSE]) - pearson(p[:, POSE], s[:, POSE])).abs()
    out["rpcc_exp"] = (pearson(g[:, EXP], s[:, EXP]) - pearson(p[:, EXP], s[:, EXP])).abs()
    out["sts_pose"], out["sts_exp"] = sts(g[:, POSE], p[:, POSE]), sts(g[:, EXP], p[:, EXP])
    if with_sid:
        out["sid_pose"] = (sid(y_true, y_pred, "pose"), sid(y_true, y_true, "pose"))
        out["sid_exp"] = (sid(y_true, y_pred, "exp"), sid(y_true, y_true, "exp"))
    f = lambda v: tuple(float(t) for t in v) if isinstance(v, tuple) else float(v)
    return {k: f(v) for k, v in out.items()}


def biwi_metrics(y_true, y_pred, templates, mouth_map, upper_map):
    """mymetrics.print_biwi_metrics (code/mymetrics.py:122-182) on the device: lip vertex error (LVE) = mean over ALL frames of the
    max over the mouth vertices of the squared distance between ground-truth and predicted vertex, and the upper-face dynamics
    deviation (FDD) = mean over clips of (mean over upper-face vertices of the temporal std of the squared motion norm, ground truth
    minus prediction).  y_true[i] (n_i, V*3), y_pred[i] (>= n_i, V*3: cut to n_i frames, :145), templates[i] (V*3).  -> (lve, fdd)."""
    mouth = torch.as_tensor(mouth_map, dtype=torch.long, device=y_true[0].device)
    upper = torch.as_tensor(upper_map, dtype=torch.long, device=y_true[0].device)
    err, fdd = [], []
    for g, p, tpl in zip(y_true, y_pred, templates):
        g = g.double().reshape(g.shape[0], -1, 3)
        p = p.double().reshape(p.shape[0], -1, 3)[: g.shape[0]]
        t = tpl.double().reshape(1, -1, 3)
        err.append(((g[:, mouth] - p[:, mouth]) ** 2).sum(dim=2).max(dim=1).values)
        std = lambda m: (m[:, upper] ** 2).sum(dim=2).std(dim=0, unbiased=False).mean()      # np.std: population std over time
        fdd.append(std(g - t) - std(p - t))
    return float(torch.cat(err).mean()), float(torch.stack(fdd).mean())
