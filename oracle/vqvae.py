"""Oracle: op-exact CPU restatement of the reference VQ-VAE (test infrastructure, see oracle/__init__.py).

Functions take `sd`, a state_dict with the reference's key names (dyadic-interaction-modeling_b200/schema.py),
and reproduce, op for op and in the same association order, what these reference lines compute:

  encoder        /root/reference/code/models/stage1_BIWI.py:307-317 (ctor :257-305)
  decoder        stage1_BIWI.py:376-393 (ctor :323-374)
  layer          models/lib/base_models.py:17-24 (Norm), :33-40 (Residual), :52-59 (MLP), :125-146 (Attention)
  gelu (tanh)    utils/base_model_util.py:81-94
  pos. encoding  base_models.py:271-273  -- indexes the table by BATCH position
  quantize       models/lib/quantizer.py:35-66 ; get_codebook_entry :79-90
  encode/decode  stage1_BIWI.py:22-37

`batch_index` generalises the F4 quirk for sharded execution: sample b receives pe[batch_index[b]]; the reference
is batch_index = arange(B).  Pinned against the real reference by tests/golden/make_golden.py.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


def gelu_tanh(x):
    cdf = 0.5 * (1.0 + torch.tanh((np.sqrt(2 / np.pi) * (x + 0.044715 * torch.pow(x, 3)))))
    return x * cdf


def _layer(sd, p, l, x, heads):
    a, m = f"{p}.net.{2*l}.fn", f"{p}.net.{2*l+1}.fn"
    H = x.shape[-1]
    scale = H ** -0.5                                           # hidden_size**-0.5, NOT head_dim (F5)
    h = F.layer_norm(x, (H,), sd[f"{a}.norm.weight"], sd[f"{a}.norm.bias"], 1e-5)
    qkv = F.linear(h, sd[f"{a}.fn.to_qkv.weight"])
    B, N, _ = qkv.shape
    qkv = qkv.view(B, N, 3, heads, H // heads).permute(2, 0, 3, 1, 4)   # 'b n (qkv h d) -> qkv b h n d'
    q, k, v = qkv[0], qkv[1], qkv[2]
    dots = torch.einsum("bhid,bhjd->bhij", q, k) * scale
    attn = F.softmax(dots, dim=-1)
    out = torch.einsum("bhij,bhjd->bhid", attn, v)
    out = out.permute(0, 2, 1, 3).reshape(B, N, H)
    out = F.linear(out, sd[f"{a}.fn.to_out.weight"], sd[f"{a}.fn.to_out.bias"])
    x = out + x
    h = F.layer_norm(x, (H,), sd[f"{m}.norm.weight"], sd[f"{m}.norm.bias"], 1e-5)
    h = F.linear(gelu_tanh(F.linear(h, sd[f"{m}.fn.l1.weight"], sd[f"{m}.fn.l1.bias"])),
                 sd[f"{m}.fn.l2.weight"], sd[f"{m}.fn.l2.bias"])
    return h + x


def _conv_block(x_bcl, w, b, neg):
    """Conv1d(k5, stride1, replicate pad 2) -> LeakyReLU -> InstanceNorm1d(affine=False)."""
    y = F.conv1d(F.pad(x_bcl, (2, 2), mode="replicate"), w, b)
    y = F.leaky_relu(y, neg)
    return F.instance_norm(y, eps=1e-5)


def _add_pe(x, pe, batch_index):
    B = x.shape[0]
    if batch_index is None:
        return x + pe[:B, :]                                    # (B,1,H) broadcast over time: the F4 quirk
    return x + pe[batch_index.long()]


def encoder(sd, x, cfg, batch_index=None, prefix=""):
    """x (B,T,in_dim) -> (B,T,face_quan_num*zquant_dim)."""
    p = prefix + "encoder"
    h = F.leaky_relu(F.linear(x, sd[f"{p}.vertice_mapping.0.weight"], sd[f"{p}.vertice_mapping.0.bias"]), cfg.neg)
    h = _conv_block(h.permute(0, 2, 1), sd[f"{p}.squasher.0.0.weight"], sd[f"{p}.squasher.0.0.bias"],
                    cfg.neg).permute(0, 2, 1)
    h = F.linear(h, sd[f"{p}.encoder_linear_embedding.net.weight"], sd[f"{p}.encoder_linear_embedding.net.bias"])
    h = _add_pe(h, sd[f"{p}.encoder_pos_embedding.pe"], batch_index)
    for l in range(cfg.num_hidden_layers):
        h = _layer(sd, f"{p}.encoder_transformer", l, h, cfg.num_attention_heads)
    return F.linear(h, sd[f"{p}.encoder_linear_embedding_post.net.weight"],
                    sd[f"{p}.encoder_linear_embedding_post.net.bias"])


def distances(z_flat, E):
    """d = sum(z^2) + sum(e^2) - 2 z E^T in this association order (quantizer.py:38-40)."""
    return torch.sum(z_flat ** 2, dim=1, keepdim=True) + torch.sum(E ** 2, dim=1) - 2 * torch.matmul(z_flat, E.t())


def quantize(sd, z, prefix=""):
    """z (B,L,D) -> (z_q (B,D,L) straight-through value, loss, (perplexity, one-hot (N,K), idx (N,1) int64))."""
    E = sd[f"{prefix}quantize.embedding.weight"]
    zf = z.reshape(-1, E.shape[1])
    d = distances(zf, E)
    idx = torch.argmin(d, dim=1).unsqueeze(1)
    onehot = torch.zeros(idx.shape[0], E.shape[0]).to(z)
    onehot.scatter_(1, idx, 1)
    z_q = torch.matmul(onehot, E).view(z.shape)
    loss = 0.25 * torch.mean((z_q - z) ** 2) + torch.mean((z_q - z) ** 2)
    z_q = z + (z_q - z)                                          # straight-through; not bit-equal to E[idx] (F14)
    e_mean = torch.mean(onehot, dim=0)
    perplexity = torch.exp(-torch.sum(e_mean * torch.log(e_mean + 1e-10)))
    return z_q.permute(0, 2, 1).contiguous(), loss, (perplexity, onehot, idx)


def codebook_entry(sd, indices, prefix=""):
    """indices (N,) int64 -> rows (N,D) via the reference's one-hot matmul (quantizer.py:79-90)."""
    E = sd[f"{prefix}quantize.embedding.weight"]
    onehot = torch.zeros(indices.shape[0], E.shape[0]).to(indices)
    onehot.scatter_(1, indices[:, None], 1)
    return torch.matmul(onehot.float(), E)


def encode(sd, x, cfg, batch_index=None, prefix=""):
    h = encoder(sd, x, cfg, batch_index, prefix)
    h = h.view(x.shape[0], -1, cfg.face_quan_num, cfg.zquant_dim).view(x.shape[0], -1, cfg.zquant_dim)
    return quantize(sd, h, prefix)


def decode(sd, quant_bcl, cfg, batch_index=None, prefix="", name="decoder"):
    """quant (B,C,L*fqn) -> (B,L,out_dim) through the TransformerDecoder registered as `name`."""
    p = prefix + name
    q = quant_bcl.permute(0, 2, 1)
    q = q.reshape(q.shape[0], -1, cfg.face_quan_num, cfg.zquant_dim)
    q = q.reshape(q.shape[0], -1, cfg.face_quan_num * cfg.zquant_dim)
    h = F.linear(q, sd[f"{p}.decoder_linear_embedding_pre.net.weight"], sd[f"{p}.decoder_linear_embedding_pre.net.bias"])
    h = _conv_block(h.permute(0, 2, 1), sd[f"{p}.expander.0.0.weight"], sd[f"{p}.expander.0.0.bias"],
                    cfg.neg).permute(0, 2, 1)
    h = F.linear(h, sd[f"{p}.decoder_linear_embedding.net.weight"], sd[f"{p}.decoder_linear_embedding.net.bias"])
    h = _add_pe(h, sd[f"{p}.decoder_pos_embedding.pe"], batch_index)
    for l in range(cfg.num_hidden_layers):
        h = _layer(sd, f"{p}.decoder_transformer", l, h, cfg.num_attention_heads)
    return F.linear(h, sd[f"{p}.vertice_map_reverse.weight"])


def decode_indices(sd, indices_bl, cfg, batch_index=None, prefix=""):
    """codes (B,L) int64 -> frames (B,L,in_dim): exact codebook rows then decode
    (seq2seq_pretrain.py:454-464 / stage1_BIWI.py:99-105)."""
    B, L = indices_bl.shape
    zq = codebook_entry(sd, indices_bl.reshape(-1), prefix).view(B, L, -1).permute(0, 2, 1)
    return decode(sd, zq, cfg, batch_index, prefix)


def speaker_decode(sd, quant_bcl, cfg, batch_index=None, prefix=""):
    """VQSpeakerAutoEncoder.decode (stage1_BIWI.py:156-165): cat([decoder_v(q) (..,56), decoder_a(q) (..,768)], -1)."""
    return torch.cat([decode(sd, quant_bcl, cfg, batch_index, prefix, "decoder_v"),
                      decode(sd, quant_bcl, cfg, batch_index, prefix, "decoder_a")], dim=-1)


def roundtrip(sd, x, cfg, batch_index=None, prefix=""):
    """VQAutoEncoder.forward (stage1_BIWI.py:49-55)."""
    quant, loss, info = encode(sd, x, cfg, batch_index, prefix)
    return decode(sd, quant, cfg, batch_index, prefix), loss, info
