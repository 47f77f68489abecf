"""Stand-in for the un-vendored dependency `x-transformers==1.30.16` (reference: code/requirements.txt:99, imported at
code/seq2seq_pretrain.py:10).

DIM uses five names: ContinuousTransformerWrapper, TransformerWrapper, Encoder, Decoder, AutoregressiveWrapper
(+ ContinuousAutoregressiveWrapper, imported but unused).  Here they are PARAMETER CONTAINERS with upstream's constructor
arguments and state_dict keys -- so `best_vico_causal.pt` loads with strict=True -- while the
arithmetic (pre-norm blocks, 12x64-head attention, erf-GELU feed-forward, KV-cached generate with top-k sampling) runs in
libdimb200 through the owning SLMFT model (dim_slmft_context / dim_slmft_generate).  Calling a wrapper on its own is not
supported: the C-ABI exposes the fused SLMFT-level entry points only.
"""
from collections import OrderedDict

import torch
import torch.nn as nn

from dim_b200.paramtree import ParamTree
from dim_b200.schema import xt_layers_schema

DEFAULT_DIM_HEAD = 64


class AttentionLayers(nn.Module):
    def __init__(self, dim, depth, heads=8, causal=False, cross_attend=False, dim_head=DEFAULT_DIM_HEAD, ff_mult=4, **kwargs):
        super().__init__()
        # upstream swallows unknown kwargs (DIM passes a stray max_seq_len: seq2seq_pretrain.py:372-380)
        self.dim, self.depth, self.heads, self.causal, self.cross_attend = dim, depth, heads, causal, cross_attend
        self.dim_head, self.ff_mult = dim_head, ff_mult
        tree = ParamTree(xt_layers_schema("x", dim, depth, heads * dim_head, cross_attend, ff_mult))
        self.layers = tree.x.layers
        self.final_norm = tree.x.final_norm

    def forward(self, *a, **k):
        raise RuntimeError("x_transformers stand-in: layers are evaluated by libdimb200 through SLMFT, not stand-alone")


class Encoder(AttentionLayers):
    def __init__(self, **kwargs):
        assert "causal" not in kwargs
        super().__init__(causal=False, **kwargs)


class Decoder(AttentionLayers):
    def __init__(self, **kwargs):
        assert "causal" not in kwargs
        super().__init__(causal=True, **kwargs)


class _Emb(nn.Module):
    def __init__(self, n, dim, kaiming=False):
        super().__init__()
        self.emb = nn.Embedding(n, dim)
        if kaiming:
            nn.init.kaiming_normal_(self.emb.weight)


class ContinuousTransformerWrapper(nn.Module):
    def __init__(self, *, max_seq_len, attn_layers, dim_in=None, dim_out=None, emb_dim=None, emb_dropout=0.0,
                 use_abs_pos_emb=True, **kwargs):
        super().__init__()
        dim = attn_layers.dim
        self.max_seq_len = max_seq_len
        self.pos_emb = _Emb(max_seq_len, dim) if use_abs_pos_emb else None
        self.project_in = nn.Linear(dim_in, dim, bias=False) if dim_in is not None else nn.Identity()
        self.attn_layers = attn_layers
        self.project_out = nn.Linear(dim, dim_out, bias=False) if dim_out is not None else nn.Identity()

    def forward(self, *a, **k):
        raise RuntimeError("x_transformers stand-in: evaluated by libdimb200 through SLMFT.forward_encoder")


class TransformerWrapper(nn.Module):
    def __init__(self, *, num_tokens, max_seq_len, attn_layers, emb_dim=None, emb_dropout=0.0, use_abs_pos_emb=True,
                 scaled_sinu_pos_emb=False, tie_embedding=False, **kwargs):
        super().__init__()
        dim = attn_layers.dim
        self.num_tokens, self.max_seq_len = num_tokens, max_seq_len
        self.token_emb = _Emb(num_tokens, dim, kaiming=True)
        self.pos_emb = _Emb(max_seq_len, dim) if use_abs_pos_emb else None     # SLMFT: use_abs_pos_emb=False
        self.attn_layers = attn_layers
        self.to_logits = nn.Linear(dim, num_tokens, bias=False)

    def forward(self, *a, **k):
        raise RuntimeError("x_transformers stand-in: evaluated by libdimb200 through SLMFT.forward_decoder")


class AutoregressiveWrapper(nn.Module):
    """`.generate(prompts, seq_len, context=, context_mask=)` is served by the owning SLMFT's engine (bound lazily)."""

    def __init__(self, net, ignore_index=-100, pad_value=0, mask_prob=0.0, add_attn_z_loss=False):
        super().__init__()
        self.net, self.ignore_index, self.pad_value, self.mask_prob = net, ignore_index, pad_value, mask_prob
        self.max_seq_len = net.max_seq_len
        self._generate_impl = None
        self._forward_impl = None

    def bind(self, fn, forward_fn=None):
        self._generate_impl = fn
        if forward_fn is not None:
            self._forward_impl = forward_fn

    @torch.no_grad()
    def generate(self, prompts, seq_len, eos_token=None, temperature=1.0, filter_logits_fn=None, filter_thres=0.9,
                 context=None, context_mask=None, uniforms=None, **kwargs):
        if self._generate_impl is None:
            raise RuntimeError("AutoregressiveWrapper.generate needs the owning SLMFT (engine binding)")
        return self._generate_impl(prompts, seq_len, temperature=temperature, context=context, context_mask=context_mask,
                                   uniforms=uniforms)

    @torch.no_grad()
    def forward(self, x, return_outputs=False, context=None, context_mask=None, self_attn_kv_mask=None, **kwargs):
        """AutoregressiveWrapper.forward(x, context=, context_mask=, return_outputs=) -> loss [, (logits, cache=None)]: teacher forcing
        through the owning model's engine (dim_slmft_teacher_forced).  inp = x[:, :-1] with ignore_index -> pad_value, target =
        x[:, 1:], cross-entropy with ignore_index; the random key mask of mask_prob > 0 is drawn like upstream unless
        `self_attn_kv_mask` (B, L-1) is given.  Forward only: the loss carries no autograd graph."""
        if self._forward_impl is None:
            raise RuntimeError("AutoregressiveWrapper.forward needs the owning model (engine binding)")
        inp, target = x[:, :-1].clone(), x[:, 1:]
        inp[inp == self.ignore_index] = self.pad_value
        kv = self_attn_kv_mask
        if kv is None and self.mask_prob > 0:
            from dim_b200.compat_api import draw_kv_mask
            kv = draw_kv_mask(inp.shape, self.mask_prob, inp.device)
        logits = self._forward_impl(context, context_mask, inp, kv)
        loss = torch.nn.functional.cross_entropy(logits.transpose(1, 2), target, ignore_index=self.ignore_index)
        return (loss, (logits, None)) if return_outputs else loss


class ContinuousAutoregressiveWrapper(nn.Module):      # imported by the reference, never instantiated by SLMFT
    def __init__(self, *a, **k):
        super().__init__()
        raise NotImplementedError("not used on the DIM inference path")
