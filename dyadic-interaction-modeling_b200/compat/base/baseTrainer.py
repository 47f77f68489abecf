"""Checkpoint helpers (reference: code/base/baseTrainer.py:26-66): `{'state_dict': ...}` containers."""
import os
from collections import OrderedDict

import torch
import torch.distributed as dist
from torch.nn.parallel import DataParallel, DistributedDataParallel


def save_checkpoint(model, other_state=None, sav_path="", filename="model.pth.tar", stage=1):
    other_state = {} if other_state is None else other_state
    inner = model.module if isinstance(model, (DistributedDataParallel, DataParallel)) else model
    if not isinstance(inner, torch.nn.Module):
        raise ValueError("model must be nn.Module or nn.DataParallel!")
    weight = inner.state_dict()
    if stage == 2:
        weight = OrderedDict((k, v) for k, v in weight.items() if "autoencoder" not in k)
    os.makedirs(sav_path, exist_ok=True)
    other_state["state_dict"] = weight
    torch.save(other_state, os.path.join(sav_path, filename))


def load_state_dict(model, state_dict, strict=True):
    inner = model.module if isinstance(model, (DistributedDataParallel, DataParallel)) else model
    inner.load_state_dict(state_dict, strict=strict)


def state_dict_remove_module(state_dict):
    return OrderedDict((k.replace("module.", ""), v) for k, v in state_dict.items())


def reduce_tensor(tensor, args):
    rt = tensor.clone()
    dist.all_reduce(rt, op=dist.ReduceOp.SUM)
    rt /= args.world_size
    return rt
