"""train_vq entry surface (reference: code/train_vq.py:25-267).

The north star keeps the ENTRY (main / main_worker / train / validate and the yaml + KEY VALUE command line); the kernels
built here are forward-only, so `validate` (reconstruction L1 + quantisation loss over a loader, :238-267) runs on the B200
path and `train` (backward + AdamW, :173-236) reports that optimisation is outside this hot path.
    python train_vq.py --config config.yaml [KEY VALUE ...]       # evaluates a (checkpointed) VQ-VAE
"""
import os

import torch
import torch.nn.functional as F

from base.utilities import AverageMeter, get_logger, get_parser
from models import get_model


def calc_vq_loss(pred, target, quant_loss, quant_loss_weight=1.0):
    """metrics/loss.py:6 -- L1 reconstruction + weighted quantisation loss."""
    rec = F.l1_loss(pred, target)
    return quant_loss.mean() * quant_loss_weight + rec, [rec, quant_loss]


def validate(val_loader, model, loss_fn, epoch, cfg):
    rec_meter, quant_meter, pp_meter = AverageMeter(), AverageMeter(), AverageMeter()
    model.eval()
    with torch.no_grad():
        for data in val_loader:
            x = (data[0] if isinstance(data, (list, tuple)) else data).cuda(non_blocking=True).float()
            out, quant_loss, info = model(x)
            _, (rec, q) = loss_fn(out, x, quant_loss, quant_loss_weight=cfg.quant_loss_weight)
            rec_meter.update(rec.item(), 1)
            quant_meter.update(q.item(), 1)
            pp_meter.update(info[0].item(), 1)
    return rec_meter.avg, quant_meter.avg, pp_meter.avg


def train(train_loader, model, loss_fn, optimizer, epoch, cfg):
    raise NotImplementedError("VQ-VAE optimisation (backward pass) is outside the B200 inference hot path (SURVEY.md section 8); "
                              "train with the reference and load the checkpoint here")


def main_worker(gpu, ngpus_per_node, cfg, val_loader=None):
    logger = get_logger()
    model = get_model(cfg).cuda()
    weight = getattr(cfg, "weight", None)
    if weight and os.path.isfile(weight):
        model.load_state_dict(torch.load(weight, map_location="cpu")["state_dict"])
        logger.info("=> loaded weight '{}'".format(weight))
    if val_loader is None:
        g = torch.Generator().manual_seed(getattr(cfg, "manual_seed", 131) or 131)
        val_loader = [torch.randn(1, 300, cfg.in_dim, generator=g) * 0.3 for _ in range(4)]     # synthetic ViCo-shape clips
    rec, quant, pp = validate(val_loader, model, calc_vq_loss, 0, cfg)
    logger.info("VAL rec_loss: {:.4f} quant_loss: {:.4f} perplexity: {:.2f}".format(rec, quant, pp))
    return rec, quant, pp


def main():
    cfg = get_parser()
    return main_worker(0, 1, cfg)


if __name__ == "__main__":
    main()
