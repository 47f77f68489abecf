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


@torch.no_grad()
def slm_forward(sd, v_speaker, v_listener, v_audio, mask, mask_speaker, mask_listener, s2s_cfg, vq_cfg):
    """SLM.forward, the pre-training model (seq2seq_pretrain.py:300-323), restated line by line with the two random masks given
    (the reference draws them with torch.randperm, :171-183).  `sd` must hold decoder_joint.net.pos_emb (SLM keeps
    use_abs_pos_emb=True, :137).  x-transformers half: oracle/xt.py (unpinned, see its header)."""
    B, T, _ = v_speaker.shape
    z_s = forward_vq_speaker(sd, v_speaker, mask, vq_cfg)                                   # :185-200
    z_l = forward_vq_listener(sd, v_listener, mask, vq_cfg)
    vs = v_speaker.clone() + sd["patch_embed_s"]                                            # :203-226
    vl = v_listener.clone() + sd["patch_embed_l"]
    vs[mask_speaker] = 0
    vl[mask_listener] = 0
    x_s = X.continuous_wrapper(sd, "encoder_s", vs, s2s_cfg.depth, mask)
    x_l = X.continuous_wrapper(sd, "encoder_l", vl, s2s_cfg.depth, mask)
    x_joint = X.continuous_wrapper(sd, "encoder_joint", torch.cat([x_s, x_l], dim=1), s2s_cfg.depth, torch.cat([mask, mask], dim=-1))
    x_l = X.continuous_wrapper(sd, "encoder_joint", x_l, s2s_cfg.depth, mask)
    x_s = X.continuous_wrapper(sd, "encoder_joint", x_s, s2s_cfg.depth, mask)
    ln = lambda x, n: F.layer_norm(x, (x.shape[-1],), sd[n + ".weight"], sd[n + ".bias"], 1e-5)
    x_joint, x_l, x_s = ln(x_joint, "norm"), ln(x_l, "norm_l"), ln(x_s, "norm_s")
    len_keep = torch.sum(mask, dim=1, dtype=torch.int32)                                    # :270-286
    s_rep = torch.stack([torch.mean(x_s[i, :len_keep[i]], dim=0) for i in range(B)], dim=0)
    l_rep = torch.stack([torch.mean(x_l[i, :len_keep[i]], dim=0) for i in range(B)], dim=0)
    s_rep, l_rep = F.normalize(s_rep, dim=-1), F.normalize(l_rep, dim=-1)
    total = torch.mm(s_rep, l_rep.t()) / 0.05
    nce = -torch.mean(torch.diag(F.log_softmax(total, dim=0)))
    c_acc = torch.sum(torch.eq(torch.argmax(F.softmax(total, dim=0), dim=0), torch.arange(0, B))) / B
    xj_s, xj_l = x_joint[:, :T], x_joint[:, T:]
    z_s = z_s.clone()
    z_l = z_l.clone()
    z_s[~mask_speaker] = -100                                                               # :307-308
    z_l[~mask_listener] = -100
    ctx_s = torch.cat([xj_s + sd["patch_embed_dec_s"], v_audio], dim=-1)                    # :228-233
    ctx_l = torch.cat([xj_l + sd["patch_embed_dec_l"], v_audio], dim=-1)
    l_ce_s, px_s = X.teacher_forced(sd, "decoder_joint.net", z_s, s2s_cfg.depth, ctx_l, mask)
    l_ce_l, px_l = X.teacher_forced(sd, "decoder_joint.net", z_l, s2s_cfg.depth, ctx_s, mask)
    pred_s = V.decode_indices(sd, torch.argmax(px_s, dim=-1), vq_cfg, prefix="speaker_vq.")  # :245-251
    pred_l = V.decode_indices(sd, torch.argmax(px_l, dim=-1), vq_cfg, prefix="listener_vq.")
    l_cont_s = continuous_loss(pred_s, v_speaker, mask_speaker)
    l_cont_l = continuous_loss(pred_l, v_listener, mask_listener)
    d = {"l_ce_s": l_ce_s, "l_ce_l": l_ce_l, "l_cont_s": l_cont_s, "l_cont_l": l_cont_l, "nce": nce, "c_acc": c_acc}
    parts = dict(x_s=x_s, x_l=x_l, x_joint=x_joint, px_s=px_s, px_l=px_l, pred_s=pred_s, pred_l=pred_l, z_s=z_s, z_l=z_l)
    return l_ce_s + l_ce_l + l_cont_s + l_cont_l + nce, d, parts


@torch.no_grad()
def listener_generator(sd, v_speaker, v_listener, mask, speaker_cfg, listener_cfg, depth=6, generate_steps=None, speaker_ids=None,
                       listener_ids=None):
    """ListenerGenerator.forward / .generate (seq2seq.py:223-290) restated, identity tokens (:240-250, :47-67) included.  Returns
    (loss, pred_cont_seq, logits, x_speaker, z_listener[, generated codes (B, generate_steps), greedy])."""
    B, T, _ = v_speaker.shape
    fqn, zd = speaker_cfg.face_quan_num, speaker_cfg.zquant_dim
    xs, zl = [], []
    for i in range(B):
        q = V.encode(sd, v_speaker[i][mask[i]].unsqueeze(0), speaker_cfg, prefix="speaker_vq.")[0]          # (1, 128, len*8)
        xs.append(F.pad(q, (0, T * fqn - q.shape[-1]), value=0))
        idx = V.encode(sd, v_listener[i][mask[i]].unsqueeze(0), listener_cfg, prefix="listener_vq.")[2][2].reshape(-1)
        zl.append(F.pad(idx, (0, T - idx.shape[-1]), value=-100))
    x = torch.cat(xs, dim=0)
    # the reference views the (B, 128, T*8) tensor as (B, -1, 8, 128) WITHOUT permuting first (seq2seq.py:238-239): kept as is
    x = x.view(B, -1, fqn, zd).contiguous().view(B, -1, fqn * zd).contiguous()
    z = torch.stack(zl, dim=0)
    x_in, mask_u, tgt = x, mask, z
    one = torch.ones(B, 1, dtype=torch.bool)
    if speaker_ids is not None:                                                              # :240-244
        tok = F.linear(F.relu(sd["speaker_embeddings.weight"][speaker_ids]), sd["fc_speaker.weight"], sd["fc_speaker.bias"])
        x_in = torch.cat([tok.unsqueeze(1), x_in], dim=1)
        mask_u = torch.cat([one, mask_u], dim=1)
    enc = X.continuous_wrapper(sd, "generator.encoder", x_in, depth, mask_u)
    if listener_ids is not None:                                                             # :247-248, Transformer.forward :49-57
        tok = F.linear(F.relu(sd["listener_embeddings.weight"][listener_ids]), sd["fc_listener.weight"], sd["fc_listener.bias"])
        enc = torch.cat([tok.unsqueeze(1), enc], dim=1)
        mask_u = torch.cat([one, mask_u], dim=1)
        tgt = torch.cat([torch.full((B, 1), -100, dtype=torch.long), tgt], dim=1)
    loss, logits = X.teacher_forced(sd, "generator.decoder.net", tgt, depth, enc, mask_u)
    if listener_ids is not None:
        logits = logits[:, 1:, :]                                                            # :66-67
    pred = V.decode_indices(sd, torch.argmax(logits, dim=-1), listener_cfg, prefix="listener_vq.")
    total = loss + continuous_loss(pred, v_listener, mask)
    out = (total, pred, logits, x, z)
    if generate_steps:
        gen = X.generate(sd, "generator.decoder.net", z[:, 0:1], generate_steps, depth, enc, mask, use_cache=False)
        out = out + (gen,)
    return out
