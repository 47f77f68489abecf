"""Kwarg helpers the reference imports with `from x_utils import *` (reference: code/x_utils.py:5-37)."""
from functools import partial

import torch.nn.functional as F


def exists(val):
    return val is not None


def max_neg_value(tensor):
    import torch
    return -torch.finfo(tensor.dtype).max


def pad_at_dim(t, pad, dim=-1, value=0.0):
    right = (-dim - 1) if dim < 0 else (t.ndim - dim - 1)
    return F.pad(t, ((0, 0) * right) + tuple(pad), value=value)


def pick_and_pop(keys, d):
    return {k: d.pop(k) for k in keys}


def group_dict_by_key(cond, d):
    yes, no = {}, {}
    for k, v in d.items():
        (yes if cond(k) else no)[k] = v
    return yes, no


def string_begins_with(prefix, s):
    return s.startswith(prefix)


def group_by_key_prefix(prefix, d):
    return group_dict_by_key(partial(string_begins_with, prefix), d)


def groupby_prefix_and_trim(prefix, d):
    with_p, rest = group_by_key_prefix(prefix, d)
    return {k[len(prefix):]: v for k, v in with_p.items()}, rest
