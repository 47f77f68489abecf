"""Oracle: CPU restatement of the x-transformers==1.30.16 subset used by DIM (test infrastructure).

PARITY UNPINNED.  x-transformers is a third-party dependency pinned in /root/reference/code/requirements.txt:99,
absent from /root/reference and not installable offline.  This file restates its published algorithm as
recorded in SURVEY.md Appendix A (A.1-A.7), anchored on the reference's own call sites:

  construction   seq2seq_pretrain.py:369-418  (Encoder depth 4 heads 12; Decoder dim 1152 cross_attend)
  encoder call   seq2seq_pretrain.py:439-440  (mask=, attn_mask=, return_embeddings=True)
  generate call  seq2seq_pretrain.py:450      (prompt (B,1), seq_len=T-1, context=, context_mask=)
  teacher forced seq2seq_pretrain.py:448

Functions operate on a state_dict `sd` with upstream key names.  Optional biases
(`project_in.bias`, `to_logits.bias`, `final_norm.bias`, pre-norm `.bias`) are honoured when the key exists, so a
checkpoint written by a build with those flags still evaluates correctly.

Semantics restated (each item: SURVEY Appendix A):
  * pre-norm residual blocks, final norm; LayerNorm eps 1e-5, gain, no bias                      (A.2)
  * attention: bias-free to_q/k/v/out, 12 heads x 64 regardless of dim, scale 64**-0.5,
    masked logits filled with -finfo.max, key mask & attn_mask combined, then causal mask,
    softmax in fp32                                                                            (A.3)
  * feed-forward Linear(+b) -> exact erf GELU -> Linear(+b)                                    (A.4)
  * continuous wrapper: bias-free project_in, + pos_emb(arange) * dim**-0.5; project_out skipped (A.5)
  * token wrapper: token_emb, no positional embedding for SLMFT, bias-free to_logits            (A.5)
  * generate: KV-cached greedy / top-k(ceil(0.1*V)) sampling, returns tokens without prompt     (A.6, A.7)
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

HEADS, DIM_HEAD = 12, 64


def _norm(sd, key, x):
    return F.layer_norm(x, (x.shape[-1],), sd[f"{key}.weight"], sd.get(f"{key}.bias"), 1e-5)


def _split_heads(t, heads):
    b, n, _ = t.shape
    return t.view(b, n, heads, -1).permute(0, 2, 1, 3)                       # 'b n (h d) -> b h n d'


def attend(q, k, v, key_mask=None, attn_mask=None, causal=False):
    """q (b,h,i,d), k/v (b,h,j,d); key_mask (b,j) True=keep; attn_mask (i,j) True=may attend."""
    scale = q.shape[-1] ** -0.5
    dots = torch.einsum("bhid,bhjd->bhij", q, k) * scale
    i, j = dots.shape[-2:]
    neg = -torch.finfo(dots.dtype).max
    final = None
    if key_mask is not None:
        final = key_mask[:, None, None, :]
    if attn_mask is not None:
        am = attn_mask[None, None, :, :]
        final = am if final is None else (final & am)                    # ~(~mask | ~attn_mask)
    if final is not None:
        dots = dots.masked_fill(~final, neg)
    if causal:
        cm = torch.ones((i, j), dtype=torch.bool).triu(j - i + 1)
        dots = dots.masked_fill(cm, neg)
    attn = F.softmax(dots, dim=-1, dtype=torch.float32).type(dots.dtype)
    out = torch.einsum("bhij,bhjd->bhid", attn, v)
    b, h, n, d = out.shape
    return out.permute(0, 2, 1, 3).reshape(b, n, h * d)


def attention_block(sd, p, x, context=None, key_mask=None, attn_mask=None, causal=False, cache=None, heads=None):
    """One Attention module.  `cache` = (k,v) of earlier positions for cached causal self-attention;
    returns (out, (k,v)).  heads: inner dim / 64 (x-transformers' dim_head default) unless given: 12 for DIM's SLM / SLMFT,
    8 for the older ListenerGenerator (seq2seq.py:172-182)."""
    heads = heads or sd[f"{p}.to_q.weight"].shape[0] // DIM_HEAD
    kv_in = x if context is None else context
    q = _split_heads(F.linear(x, sd[f"{p}.to_q.weight"]), heads)
    k = _split_heads(F.linear(kv_in, sd[f"{p}.to_k.weight"]), heads)
    v = _split_heads(F.linear(kv_in, sd[f"{p}.to_v.weight"]), heads)
    if cache is not None:
        k = torch.cat((cache[0], k), dim=-2)
        v = torch.cat((cache[1], v), dim=-2)
    out = attend(q, k, v, key_mask, attn_mask, causal)
    return F.linear(out, sd[f"{p}.to_out.weight"]), (k, v)


def feed_forward(sd, p, x):
    h = F.gelu(F.linear(x, sd[f"{p}.ff.0.0.weight"], sd[f"{p}.ff.0.0.bias"]))      # exact erf GELU
    return F.linear(h, sd[f"{p}.ff.2.weight"], sd[f"{p}.ff.2.bias"])


def encoder_layers(sd, p, x, depth, mask=None, attn_mask=None):
    """x-transformers Encoder (non-causal; causality comes from attn_mask as DIM passes it)."""
    for i in range(2 * depth):
        lp = f"{p}.layers.{i}"
        h = _norm(sd, f"{lp}.0.0", x)
        if i % 2 == 0:
            out, _ = attention_block(sd, f"{lp}.1", h, key_mask=mask, attn_mask=attn_mask)
        else:
            out = feed_forward(sd, f"{lp}.1", h)
        x = out + x
    return _norm(sd, f"{p}.final_norm", x)


def continuous_wrapper(sd, name, x, depth, mask=None, attn_mask=None):
    """ContinuousTransformerWrapper(...)(x, mask=, attn_mask=, return_embeddings=True)."""
    dim = sd[f"{name}.project_in.weight"].shape[0]
    h = F.linear(x, sd[f"{name}.project_in.weight"], sd.get(f"{name}.project_in.bias"))
    n = x.shape[1]
    h = h + sd[f"{name}.pos_emb.emb.weight"][:n] * (dim ** -0.5)
    return encoder_layers(sd, f"{name}.attn_layers", h, depth, mask, attn_mask)


def decoder_layers(sd, p, x, depth, context, context_mask, caches=None, cross_kv=None, self_kv_mask=None):
    """x-transformers Decoder with cross attention, layer order (a, c, f) * depth.

    caches: list (one per self-attn layer) of (k,v) or None -> cached single-step mode when given.
    cross_kv: optional list of precomputed cross (k,v) per cross layer (the "project once" variant; numerically
    identical to re-projecting the context every step, which is what upstream does - SURVEY F9).
    Returns (x, new_caches, cross_kv_used)."""
    new_caches, used_cross = [], []
    ai = ci = 0
    for i in range(3 * depth):
        lp = f"{p}.layers.{i}"
        h = _norm(sd, f"{lp}.0.0", x)
        kind = "acf"[i % 3]
        if kind == "a":
            cache = caches[ai] if caches is not None else None
            out, kv = attention_block(sd, f"{lp}.1", h, causal=True, cache=cache,
                                      key_mask=self_kv_mask if cache is None else None)
            new_caches.append(kv)
            ai += 1
        elif kind == "c":
            if cross_kv is not None:
                q = _split_heads(F.linear(h, sd[f"{lp}.1.to_q.weight"]), sd[f"{lp}.1.to_q.weight"].shape[0] // DIM_HEAD)
                k, v = cross_kv[ci]
                out = F.linear(attend(q, k, v, key_mask=context_mask), sd[f"{lp}.1.to_out.weight"])
                used_cross.append((k, v))
            else:
                out, kv = attention_block(sd, f"{lp}.1", h, context=context, key_mask=context_mask)
                used_cross.append(kv)
            ci += 1
        else:
            out = feed_forward(sd, f"{lp}.1", h)
        x = out + x
    return _norm(sd, f"{p}.final_norm", x), new_caches, used_cross


def token_wrapper_logits(sd, name, tokens, depth, context, context_mask, caches=None, cross_kv=None,
                         self_kv_mask=None):
    """TransformerWrapper.forward -> logits.  SLMFT: use_abs_pos_emb=False, so nothing is added to the
    token embedding (seq2seq_pretrain.py:386)."""
    x = F.embedding(tokens, sd[f"{name}.token_emb.emb.weight"])
    pe = sd.get(f"{name}.pos_emb.emb.weight")
    if pe is not None:                                                     # SLM / SpeakerSLMFT flavour
        assert caches is None, "abs-pos-emb + cache not restated"
        x = x + pe[: x.shape[1]] * (x.shape[-1] ** -0.5)
    x, new_caches, used_cross = decoder_layers(sd, f"{name}.attn_layers", x, depth, context, context_mask,
                                               caches, cross_kv, self_kv_mask)
    return F.linear(x, sd[f"{name}.to_logits.weight"], sd.get(f"{name}.to_logits.bias")), new_caches, used_cross


def top_k_filter(logits, frac=0.1, k=None):
    k = math.ceil(frac * logits.shape[-1]) if k is None else k
    val, ind = torch.topk(logits, k, dim=-1)
    out = torch.full_like(logits, float("-inf"))
    return out.scatter_(-1, ind, val)


def sample_from_uniform(probs, u):
    """Inverse-CDF draw in index order: smallest i with cumsum(probs)[i] > u*sum(probs).
    (torch.multinomial consumes the RNG differently; the distribution is the same.)"""
    c = torch.cumsum(probs.double(), dim=-1)
    t = u.double().unsqueeze(-1) * c[..., -1:]
    idx = (c <= t).sum(dim=-1)
    return idx.clamp_(max=probs.shape[-1] - 1)


@torch.no_grad()
def generate(sd, name, prompt, seq_len, depth, context, context_mask, temperature=0.0, uniforms=None,
             top_k_frac=0.1, top_k=None, cross_kv_once=True, use_cache=True, return_logits=False):
    """AutoregressiveWrapper.generate restated.  prompt (B,1) int64 -> (B,seq_len) int64 (prompt stripped).

    temperature == 0 -> argmax (upstream's greedy branch).  temperature > 0 -> top-k filter, softmax(logits/T),
    draw with the supplied `uniforms` (B,seq_len) (one per step) by inverse CDF.
    cross_kv_once=False re-projects the context in every step like upstream; use_cache=False recomputes the whole
    prefix every step (reference semantics without the cache).  All variants give identical tokens."""
    out = prompt.clone()
    caches, cross = None, None
    all_logits = []
    for t in range(seq_len):
        if use_cache:
            x = out[:, -1:] if caches is not None else out
            logits, caches, used = token_wrapper_logits(sd, name, x, depth, context, context_mask, caches,
                                                        cross if cross_kv_once else None)
            if cross_kv_once:
                cross = used
        else:
            logits, _, _ = token_wrapper_logits(sd, name, out, depth, context, context_mask)
        logits = logits[:, -1]
        if return_logits:
            all_logits.append(logits.clone())
        if temperature == 0.0:
            sample = logits.argmax(dim=-1, keepdim=True)
        else:
            filt = top_k_filter(logits, top_k_frac, top_k)
            probs = F.softmax(filt / temperature, dim=-1)
            sample = sample_from_uniform(probs, uniforms[:, t]).unsqueeze(-1)
        out = torch.cat((out, sample), dim=-1)
    res = out[:, prompt.shape[1]:]
    return (res, torch.stack(all_logits, 1)) if return_logits else res


def teacher_forced(sd, name, tokens, depth, context, context_mask, ignore_index=-100, pad_value=0, kv_mask=None):
    """AutoregressiveWrapper.forward(x, return_outputs=True) -> (loss, logits (B,T-1,V)).
    `kv_mask` is the self_attn_kv_mask upstream draws at random when mask_prob>0; pass it explicitly (or None)."""
    inp, target = tokens[:, :-1].clone(), tokens[:, 1:]
    inp[inp == ignore_index] = pad_value
    logits, _, _ = token_wrapper_logits(sd, name, inp, depth, context, context_mask, self_kv_mask=kv_mask)
    loss = F.cross_entropy(logits.transpose(1, 2), target, ignore_index=ignore_index)
    return loss, logits
