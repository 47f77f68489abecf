"""Parameter containers whose state_dict keys equal the reference's.

The reference models are ordinary nn.Module trees; a checkpoint is the flat dict of their dotted parameter names.  Here the
arithmetic lives in the CUDA library, so the Python side only needs modules that OWN tensors under the same dotted names
(so that `load_state_dict(strict=True)` of a real checkpoint works and `state_dict()` round-trips).  `ParamTree` builds
that hierarchy from a schema (key -> shape, dim_b200.schema) instead of re-declaring every layer class.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import torch
import torch.nn as nn

from .synth import sinusoid_table


class ParamTree(nn.Module):
    """Nested container built from dotted keys.  Leaves named `pe` are buffers (sinusoid tables), others nn.Parameter."""

    def __init__(self, schema: "OrderedDict[str, tuple]", init=None):
        super().__init__()
        for key, shape in schema.items():
            parts = key.split(".")
            node = self
            for p in parts[:-1]:
                if p not in node._modules:
                    node.add_module(p, nn.Module())
                node = node._modules[p]
            leaf = parts[-1]
            if leaf == "pe":
                node.register_buffer("pe", sinusoid_table(shape[0], shape[2]))
            else:
                node.register_parameter(leaf, nn.Parameter((init or default_init)(key, shape)))

    def forward(self, *a, **k):  # pragma: no cover - containers hold weights only
        raise RuntimeError("ParamTree holds weights; the computation is done by the owning model through dim_b200.engine")


def default_init(key: str, shape) -> torch.Tensor:
    """torch's default initialisers for the layer kinds that occur (Linear/Conv1d: kaiming-uniform a=sqrt(5);
    LayerNorm: ones/zeros; Embedding: N(0,1))."""
    leaf = key.rsplit(".", 1)[-1]
    if "norm" in key or (len(shape) == 1 and leaf == "weight"):
        return torch.ones(shape) if leaf == "weight" else torch.zeros(shape)
    if "emb.weight" in key:
        t = torch.empty(shape)
        if "token_emb" in key:
            nn.init.kaiming_normal_(t)
        else:
            nn.init.normal_(t)
        return t
    if key.startswith("patch_embed"):
        return torch.zeros(shape)
    if leaf == "weight":
        t = torch.empty(shape)
        nn.init.kaiming_uniform_(t, a=math.sqrt(5))
        return t
    fan_in = 1
    return torch.zeros(shape).uniform_(-0.05, 0.05) if leaf == "bias" else torch.zeros(shape)


def strip_prefix(schema, prefix):
    return OrderedDict((k[len(prefix):], v) for k, v in schema.items() if k.startswith(prefix))


def fingerprint(module: nn.Module):
    """Cheap identity of a module's storage: changes when parameters are moved, reloaded or updated in place."""
    fp = []
    for t in list(module.parameters()) + list(module.buffers()):
        fp.append((t.data_ptr(), t._version, t.device.index if t.is_cuda else -1))
    return hash(tuple(fp))
