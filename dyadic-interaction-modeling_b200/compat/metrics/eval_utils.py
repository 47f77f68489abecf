"""Frechet distance helpers used by the eval loop's best-of-N selection (reference: code/metrics/eval_utils.py,
called from code/x_engine_pt.py:260-268).  Host-side numpy/scipy post-processing, as in the reference."""
import numpy as np
from scipy import linalg


def calculate_activation_statistics(activations):
    return np.mean(activations, axis=0), np.cov(activations, rowvar=False)


def calculate_frechet_distance(mu1, sigma1, mu2, sigma2, eps=1e-6):
    mu1, mu2 = np.atleast_1d(mu1), np.atleast_1d(mu2)
    sigma1, sigma2 = np.atleast_2d(sigma1), np.atleast_2d(sigma2)
    diff = mu1 - mu2
    covmean = linalg.sqrtm(sigma1.dot(sigma2))
    if not np.isfinite(covmean).all():
        off = np.eye(sigma1.shape[0]) * eps
        covmean = linalg.sqrtm((sigma1 + off).dot(sigma2 + off))
    if np.iscomplexobj(covmean):
        covmean = covmean.real
    return diff.dot(diff) + np.trace(sigma1) + np.trace(sigma2) - 2 * np.trace(covmean)


def calculate_variance(activations):
    """eval_utils.py:46-47"""
    return np.sum(np.var(activations, axis=0))


def calcuate_sid(gt, pred, type='exp'):
    """eval_utils.py:49-83 (name spelled as in the reference): entropy of the predictions' histogram over k-means clusters fitted
    on the ground truth (k = 40 expression / 20 pose, KMeans(random_state=0, n_init='auto')).  The assignment and the entropy run
    as tensor ops (dim_b200.metrics.sid); lists of (n_i, 56) numpy arrays in, float out."""
    import torch
    from dim_b200 import metrics as M
    dev = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
    to = lambda seq: [torch.as_tensor(np.asarray(a)).to(dev) for a in seq]
    return float(M.sid(to(gt), to(pred), "exp" if type == "exp" else "pose"))


def sts(x, y, timestep=0.1):
    """eval_utils.py:85-91, vectorised: sqrt(sum(((dx - dy)^2) / timestep)) over the concatenated sequence."""
    x, y = np.asarray(x, dtype=np.float64), np.asarray(y, dtype=np.float64)
    return float(np.sqrt(((((x[1:] - x[:-1]) - (y[1:] - y[:-1])) ** 2) / timestep).sum()))
