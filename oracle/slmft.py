"""Oracle: SLMFT.forward(mode='val') composed from oracle.vqvae + oracle.xt (test infrastructure).

Follows /root/reference/code/seq2seq_pretrain.py:
  forward_vq          :480-494   per-sample B=1 VQ encodes of the valid frames (so every sample gets pe[0])
  forward_encoder     :431-442   v + patch_embed_s -> encoder_s -> encoder_joint -> norm_s, causal attn_mask + key mask
  forward_decoder     :444-452   context = cat(x_s + patch_embed_dec_s, audio); generate from z_l[:,0:1] for T-1 steps
  forward_vq_decoder  :454-464   exact codebook rows -> listener_vq.decode (batched: sample b gets pe[b])
  forward_continuous_loss :466-478
  forward             :496-514
The reference hard-codes .cuda() (:437) and cannot run on CPU; this restatement is device-free.
PARITY: the VQ half is pinned (oracle/vqvae.py); the transformer half is UNPINNED (oracle/xt.py).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import vqvae as V
from . import xt as X


def forward_vq_listener(sd, v_listener, mask, vq_cfg):
    """z_listener (B,T) int64, padded with -100 (seq2seq_pretrain.py:489-491)."""
    B, T, _ = v_listener.shape
    rows = []
    for i in range(B):
        idx = V.encode(sd, v_listener[i][mask[i]].unsqueeze(0), vq_cfg, prefix="listener_vq.")[2][2].reshape(-1)
        rows.append(F.pad(idx, (0, T - idx.shape[-1]), value=-100))
    return torch.stack(rows, 0)


def forward_vq_speaker(sd, v_speaker, mask, vq_cfg):
    """z_speaker (B,T) int64 padded with 0 (:486-488).  Computed by the reference but never consumed in val mode."""
    B, T, _ = v_speaker.shape
    rows = []
    for i in range(B):
        idx = V.encode(sd, v_speaker[i][mask[i]].unsqueeze(0), vq_cfg, prefix="speaker_vq.")[2][2].reshape(-1)
        rows.append(F.pad(idx, (0, T - idx.shape[-1]), value=0))
    return torch.stack(rows, 0)


def forward_encoder(sd, v_speaker, mask, s2s_cfg):
    x = v_speaker.clone() + sd["patch_embed_s"]
    n = x.shape[1]
    attn_mask = ~torch.triu(torch.ones(n, n), diagonal=1).bool()
    x = X.continuous_wrapper(sd, "encoder_s", x, s2s_cfg.depth, mask, attn_mask)
    x = X.continuous_wrapper(sd, "encoder_joint", x, s2s_cfg.depth, mask, attn_mask)
    return F.layer_norm(x, (x.shape[-1],), sd["norm_s.weight"], sd["norm_s.bias"], 1e-5)


def decoder_context(sd, x_s, v_audio):
    return torch.cat([x_s + sd["patch_embed_dec_s"], v_audio], dim=-1)


def continuous_loss(pred, target, mask):
    target, mask = target[:, 1:, :], mask[:, 1:]
    B = len(target)
    target = target.reshape(B * target.shape[1], -1)
    pred = pred.reshape(B * pred.shape[1], -1)
    mask = mask.reshape(-1)
    p, t = pred[mask], target[mask]
    return torch.mean(F.pairwise_distance(p[:, 6:], t[:, 6:])) + torch.mean(F.pairwise_distance(p[:, 0:6], t[:, 0:6]))


@torch.no_grad()
def forward_val(sd, v_speaker, v_listener, v_audio, mask, s2s_cfg, vq_cfg, temperature=0.0, uniforms=None,
                batch_index=None, cross_kv_once=True, as_reference=False, return_intermediates=False):
    """-> (total_loss, dict, pred_cont_seq_l (B,T-1,56)) like SLMFT.forward(..., mode='val').

    temperature=0 selects upstream's greedy branch (deterministic parity); temperature>0 uses `uniforms`.
    as_reference=True also executes the work the reference does and discards (speaker encodes, the duplicated
    forward_vq call: SURVEY F10) - for CPU-baseline timing only.
    batch_index: global batch positions for the decode-side positional-encoding quirk when sharded."""
    if as_reference:
        forward_vq_speaker(sd, v_speaker, mask, vq_cfg)
        forward_vq_listener(sd, v_listener, mask, vq_cfg)
        forward_vq_speaker(sd, v_speaker, mask, vq_cfg)
    z_l = forward_vq_listener(sd, v_listener, mask, vq_cfg)
    x_s = forward_encoder(sd, v_speaker, mask, s2s_cfg)
    ctx = decoder_context(sd, x_s, v_audio)
    codes = X.generate(sd, "decoder_joint.net", z_l[:, 0:1], z_l.shape[1] - 1, s2s_cfg.depth, ctx, mask,
                       temperature=temperature, uniforms=uniforms, top_k_frac=s2s_cfg.top_k_frac,
                       cross_kv_once=cross_kv_once)
    pred = V.decode_indices(sd, codes, vq_cfg, batch_index, prefix="listener_vq.")
    l_cont = continuous_loss(pred, v_listener, mask)
    d = {"l_ce_s": 0, "l_ce_l": 0.0, "l_cont_s": 0, "l_cont_l": l_cont, "nce": 0, "c_acc": 0}
    if return_intermediates:
        return l_cont, d, pred, dict(z_l=z_l, x_s=x_s, ctx=ctx, codes=codes)
    return l_cont, d, pred
