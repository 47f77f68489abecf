"""Deterministic synthetic checkpoints and clips (there are no datasets or trained weights offline).

Every tensor is drawn from its own generator seeded by (seed, crc32(key)), so the same state_dict comes out
on any box and independent of key order.  Shapes follow schema.py; magnitudes follow the modules' default
initialisers (nn.Linear: U(+-1/sqrt(fan_in)); nn.Embedding: N(0,1); x-transformers token_emb: kaiming normal),
except the codebook, which is drawn at O(1) scale: the reference's U(+-1/512) init makes all 512 codes
near-identical and produces artificial near-ties.

Clip shapes follow the loaders: ViCo 30 fps T=300 (5..1024 allowed), CANDOR T<=250,
LM-Listener T in {64, 1024}; speaker/listener motion 56-d (50 exp + 6 pose), audio 768-d HuBERT features.
"""
from __future__ import annotations

import math
import re
import zlib
from collections import OrderedDict

import torch

from .schema import S2SConfig, VQConfig, slmft_own_schema, vqvae_schema


def _gen(seed: int, key: str) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed(((seed + 1) * 1000003 + zlib.crc32(key.encode())) & 0x7FFFFFFF)
    return g


def sinusoid_table(max_len: int, d_model: int) -> torch.Tensor:
    """The `pe` buffer of PositionalEncoding (models/lib/base_models.py:263-269), shape (max_len,1,d)."""
    pe = torch.zeros(max_len, d_model)
    position = torch.arange(0, max_len, dtype=torch.float).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, d_model, 2).float() * (-math.log(10000.0) / d_model))
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return pe.unsqueeze(0).transpose(0, 1).contiguous()


def _draw(key: str, shape, seed: int) -> torch.Tensor:
    g = _gen(seed, key)
    leaf = key.rsplit(".", 1)[-1]
    if key.endswith(".pe"):
        return sinusoid_table(shape[0], shape[2])
    if key.endswith("quantize.embedding.weight"):
        return torch.randn(shape, generator=g) * 0.5
    if "token_emb" in key:
        return torch.randn(shape, generator=g) * math.sqrt(2.0 / shape[1])
    if "pos_emb" in key:
        return torch.randn(shape, generator=g)
    if key.startswith("patch_embed"):
        return torch.randn(shape, generator=g) * 0.02
    is_norm = (".norm." in key or key.startswith("norm") or "final_norm" in key
               or re.search(r"layers\.\d+\.0\.0\.weight$", key) is not None)
    if is_norm:
        if leaf == "weight":
            return 1.0 + 0.1 * torch.randn(shape, generator=g)
        return 0.05 * torch.randn(shape, generator=g)
    if leaf == "weight":
        fan_in = 1
        for s in shape[1:]:
            fan_in *= s
        b = 1.0 / math.sqrt(fan_in)
        return (torch.rand(shape, generator=g) * 2 - 1) * b
    if leaf == "bias":
        return (torch.rand(shape, generator=g) * 2 - 1) * 0.05
    if leaf.startswith(("weight_ih_l", "weight_hh_l", "bias_ih_l", "bias_hh_l")):       # nn.LSTM: U(+-1/sqrt(hidden)) everywhere
        return (torch.rand(shape, generator=g) * 2 - 1) / math.sqrt(shape[0] // 4)
    if key == "W":
        return torch.randn(shape, generator=g)
    raise KeyError(key)


def make_vqvae_state_dict(seed: int = 131, cfg: VQConfig = VQConfig()) -> "OrderedDict[str, torch.Tensor]":
    return OrderedDict((k, _draw(k, s, seed)) for k, s in vqvae_schema(cfg).items())


def make_vqspeaker_state_dict(seed: int = 131, cfg: VQConfig = None) -> "OrderedDict[str, torch.Tensor]":
    """VQSpeakerAutoEncoder (stage1_BIWI.py:140-151) with synthetic weights, reference key names."""
    from .schema import SPEAKER_VQ, vqspeaker_schema
    return OrderedDict((k, _draw(k, s, seed)) for k, s in vqspeaker_schema(cfg or SPEAKER_VQ).items())


def make_slmft_state_dict(seed: int = 131, cfg: S2SConfig = S2SConfig(), vq: VQConfig = VQConfig()):
    sd = OrderedDict()
    for pre, off in (("speaker_vq", 1), ("listener_vq", 2)):
        for k, v in make_vqvae_state_dict(seed + off, vq).items():
            sd[f"{pre}.{k}"] = v
    for k, s in slmft_own_schema(cfg).items():
        sd[k] = _draw(k, s, seed)
    return sd


def make_listener_generator_state_dict(seed: int = 131):
    """ListenerGenerator (seq2seq.py:138-199): speaker_vq (VQSpeakerAutoEncoder) + listener_vq + generator (encoder / decoder of dim 512,
    depth 6, 8 heads, dim_in 1024, decoder with a positional table) + the identity-embedding heads, reference key names."""
    from .schema import SPEAKER_VQ, vqspeaker_schema, xt_encoder_wrapper_schema, xt_layers_schema
    cfg = S2SConfig(dim_in=1024, dim=512, dim_audio=0, depth=6, heads=8, dim_head=64, max_seq_len=1024, num_tokens=512, ff_mult=4)
    sd = OrderedDict()
    for k, sh in vqspeaker_schema(SPEAKER_VQ).items():
        sd["speaker_vq." + k] = _draw("speaker_vq." + k, sh, seed + 1)
    for k, v in make_vqvae_state_dict(seed + 2).items():
        sd["listener_vq." + k] = v
    for k, sh in xt_encoder_wrapper_schema("generator.encoder", cfg.dim_in, cfg).items():
        sd[k] = _draw(k, sh, seed)
    d = OrderedDict()
    d["generator.decoder.net.token_emb.emb.weight"] = (cfg.num_tokens, cfg.dim)
    d["generator.decoder.net.pos_emb.emb.weight"] = (cfg.max_seq_len, cfg.dim)
    d.update(xt_layers_schema("generator.decoder.net.attn_layers", cfg.dim, cfg.depth, cfg.inner, True, cfg.ff_mult))
    d["generator.decoder.net.to_logits.weight"] = (cfg.num_tokens, cfg.dim)
    for k, sh in d.items():
        sd[k] = _draw(k, sh, seed)
    g = _gen(seed, "identity_heads")
    sd["speaker_embeddings.weight"] = torch.randn(100, 256, generator=g)
    sd["listener_embeddings.weight"] = torch.randn(100, 256, generator=g)
    sd["fc_speaker.weight"] = torch.randn(1024, 256, generator=g) * 0.05
    sd["fc_speaker.bias"] = torch.zeros(1024)
    sd["fc_listener.weight"] = torch.randn(512, 256, generator=g) * 0.05
    sd["fc_listener.bias"] = torch.zeros(512)
    return sd


def make_slm_state_dict(seed: int = 131, cfg: S2SConfig = S2SConfig(), vq: VQConfig = VQConfig()):
    """SLM (pre-training model, seq2seq_pretrain.py:72-169): SLMFT's keys + the decoder's absolute positional table (:137)."""
    sd = make_slmft_state_dict(seed, cfg, vq)
    key = "decoder_joint.net.pos_emb.emb.weight"
    sd[key] = _draw(key, (cfg.max_seq_len, cfg.dec_dim), seed)
    return sd


def make_clips(batch: int, frames: int, seed: int = 0, speaker: str = "randn", ragged: bool = False):
    """Synthetic dyadic clips.  Returns dict(v_speaker (B,T,56), v_listener (B,T,56), v_audio (B,T,768),
    lengths (B,) int64, mask (B,T) bool) on CPU in fp32.

    speaker="ones" reproduces the ViCo loader, which replaces speaker motion by ones (dataset/data_loader.py:147).
    """
    g = _gen(seed, f"clips/{batch}/{frames}")
    v_l = torch.randn(batch, frames, 56, generator=g) * 0.3
    v_a = torch.randn(batch, frames, 768, generator=g)
    v_s = torch.ones(batch, frames, 56) if speaker == "ones" else torch.randn(batch, frames, 56, generator=g) * 0.3
    if ragged:
        lengths = torch.randint(max(5, frames // 2), frames + 1, (batch,), generator=g)
        lengths[0] = frames
    else:
        lengths = torch.full((batch,), frames, dtype=torch.int64)
    mask = torch.arange(frames)[None, :] < lengths[:, None]
    return dict(v_speaker=v_s, v_listener=v_l, v_audio=v_a, lengths=lengths, mask=mask)


CLIP_SHAPES = {          # name -> (frames, note)
    "vico": 300,         # 30 fps x 10 s (BASELINE.json configs[1], [2])
    "candor": 250,       # loader keeps 5..250 frames (dataset/data_loader.py:379)
    "lm_listener_64": 64,
    "lm_listener_1024": 1024,
}


def make_emoca_converter_state_dict(seed: int = 131, size: int = None, vq: VQConfig = VQConfig()):
    """EmocaConverter (seq2seq_pretrain.py:759-813): speaker_vq (a 56-d VQAutoEncoder, ./config.yaml) + the mesh modules."""
    from .schema import MESH_SIZE, emoca_converter_own_schema
    sd = OrderedDict()
    for k, v in make_vqvae_state_dict(seed + 1, vq).items():
        sd["speaker_vq." + k] = v
    for k, sh in emoca_converter_own_schema(size or MESH_SIZE, vq.in_dim).items():
        sd[k] = _draw(k, sh, seed)
    return sd


def make_speaker_slmft_state_dict(seed: int = 131, size: int = None, cfg: S2SConfig = S2SConfig(), vq: VQConfig = VQConfig()):
    """SpeakerSLMFT (seq2seq_pretrain.py:516-637) with synthetic weights, reference key names."""
    from .schema import MESH_SIZE, speaker_slmft_own_schema
    sd = OrderedDict()
    for pre, off in (("speaker_vq", 1), ("listener_vq", 2)):
        for k, v in make_vqvae_state_dict(seed + off, vq).items():
            sd[f"{pre}.{k}"] = v
    for k, sh in speaker_slmft_own_schema(cfg, size or MESH_SIZE).items():
        sd[k] = torch.randn(sh, generator=_gen(seed, k)) if k == "speaker_embed.weight" else _draw(k, sh, seed)   # nn.Embedding: N(0,1)
    return sd
