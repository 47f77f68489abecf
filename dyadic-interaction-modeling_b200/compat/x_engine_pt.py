"""x_engine_pt.evaluate_test_epoch on the B200 model (reference: code/x_engine_pt.py:232-277).

Same loop contract: for every batch draw `beam_size` = 10 stochastic generations, and per clip keep the sample whose
Frechet distance to the ground truth is smallest.  Returns (y_trues, y_preds, x_all, data_ids) as lists of numpy arrays
of shape (src_len-1, 56).

When the model offers `forward_val_samples` (the B200 SLMFT does) the 10 generations of a batch are ONE pass -- the
encoders and the cross-attention K/V projection run once per clip, the decode runs 10 x B rows sharing each clip's K/V --
and the Frechet-distance selection happens on the device (dim_b200.compat_api.best_of_n: fp64 covariances + two `eigh`
instead of scipy's sqrtm), so only the winning sample of each clip is copied to the host.  Any other model object falls
back to the reference's loop (10 model calls, host-side numpy/scipy selection)."""
import numpy as np
import torch

from metrics.eval_utils import calculate_activation_statistics, calculate_frechet_distance

try:  # tqdm is optional
    from tqdm import tqdm
except Exception:  # pragma: no cover
    tqdm = lambda x: x


def evaluate_test_epoch(model, loader, device, beam_size=10):
    y_trues_all, y_preds_all, x_all, data_ids_all = [], [], [], []
    model.eval()
    with torch.no_grad():
        for batch in tqdm(loader):
            src, tgt, src_len, _, data_ids = batch
            src, tgt = src.to(device), tgt.to(device)
            src_s_v, src_s_a = torch.split(src, [56, 768], dim=2)
            lens = torch.as_tensor(list(src_len), device=device)
            mask = torch.arange(src.shape[1], device=device)[None, :] < lens[:, None]
            y_true = tgt[:, 1:, :].cpu().numpy()
            x_np = src_s_v.cpu().numpy()
            B = src.shape[0]
            truth_stats = []
            for j in range(B):
                n = int(src_len[j]) - 1
                y_trues_all.append(y_true[j][:n])
                data_ids_all.append(data_ids[j])
                x_all.append(x_np[j, :n])
                truth_stats.append(calculate_activation_statistics(y_true[j][:n]))
            if hasattr(model, "forward_val_samples"):
                from dim_b200.compat_api import best_of_n
                preds, _ = model.forward_val_samples(src_s_v.contiguous(), tgt, src_s_a.contiguous(), mask, beam_size)
                picked, _, _ = best_of_n(preds, tgt[:, 1:, :], [int(n) - 1 for n in src_len])
                y_preds_all.extend(p.cpu().numpy() for p in picked)
                continue
            best, keep = [float("inf")] * B, [None] * B
            for _ in range(beam_size):
                _, _, y_preds = model(src_s_v.contiguous(), tgt, src_s_a.contiguous(), mask, mode="val")
                yp = y_preds.cpu().numpy()
                for j in range(B):
                    n = int(src_len[j]) - 1
                    mu2, sigma2 = calculate_activation_statistics(yp[j][:n])
                    fd = calculate_frechet_distance(truth_stats[j][0], truth_stats[j][1], mu2, sigma2)
                    if fd < best[j]:
                        best[j], keep[j] = fd, yp[j][:n].copy()
            y_preds_all.extend(keep)
    return y_trues_all, y_preds_all, x_all, data_ids_all


def evaluate_finetune_epoch(model, loader, device):
    """Teacher-forced evaluation (reference: code/x_engine_pt.py:201-230): one mode='train' forward per batch, predictions cut to
    src_len - 1 frames.  Returns (y_trues, y_preds, x_all, data_ids)."""
    y_trues_all, y_preds_all, x_all, data_ids_all = [], [], [], []
    model.eval()
    with torch.no_grad():
        for batch in tqdm(loader):
            src, tgt, src_len, _, data_ids = batch
            src, tgt = src.to(device), tgt.to(device)
            src_s_v, src_s_a = torch.split(src, [56, 768], dim=2)
            lens = torch.as_tensor(list(src_len), device=device)
            mask = torch.arange(src.shape[1], device=device)[None, :] < lens[:, None]
            _, _, y_preds = model(src_s_v.contiguous(), tgt, src_s_a.contiguous(), mask, mode="train")
            yp, yt, xs = y_preds.cpu().numpy(), tgt[:, 1:, :].cpu().numpy(), src_s_v.cpu().numpy()
            for j in range(len(yp)):
                n = int(src_len[j]) - 1
                y_preds_all.append(yp[j][:n])
                y_trues_all.append(yt[j][:n])
                x_all.append(xs[j, :n])
                data_ids_all.append(data_ids[j])
    return y_trues_all, y_preds_all, x_all, data_ids_all


def train_epoch(model, loader, optimizer, device, scheduler=None, clip=1.0, print_freq=10, epoch=0):
    """Name kept for `from x_engine_pt import train_epoch, ...` (test_s2s_pretrain.py:6; the call is commented out there).
    Training needs the backward pass, which is out of scope."""
    raise NotImplementedError("train_epoch: fine-tuning (backward pass) is outside the inference path built here")
