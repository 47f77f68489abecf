"""state_dict schemas (key -> shape) of the two model families on the DIM hot path.

The key names are the reference's own so that real checkpoints load unchanged:
  * VQ-VAE  : /root/reference/code/models/stage1_BIWI.py:10-37,254-393 (+ lib/base_models.py, lib/quantizer.py)
  * SLMFT   : /root/reference/code/seq2seq_pretrain.py:325-429 (x-transformers 1.30.16 module tree,
              SURVEY.md Appendix A.8)
Only shapes live here; no arithmetic.
"""
from __future__ import annotations

from collections import OrderedDict
from dataclasses import dataclass


@dataclass(frozen=True)
class VQConfig:
    """Mirror of the yaml keys the VQ-VAE reads (code/config.yaml:15-30)."""
    in_dim: int = 56
    hidden_size: int = 384
    num_hidden_layers: int = 6
    num_attention_heads: int = 8
    intermediate_size: int = 1536
    quant_factor: int = 0
    face_quan_num: int = 1
    neg: float = 0.2
    INaffine: bool = False
    n_embed: int = 512
    zquant_dim: int = 128
    pe_max_len: int = 5000

    @staticmethod
    def from_cfg(cfg) -> "VQConfig":
        g = (lambda k, d: cfg[k] if k in cfg else d) if isinstance(cfg, dict) else (lambda k, d: getattr(cfg, k, d))
        return VQConfig(
            in_dim=int(g("in_dim", 56)), hidden_size=int(g("hidden_size", 384)),
            num_hidden_layers=int(g("num_hidden_layers", 6)),
            num_attention_heads=int(g("num_attention_heads", 8)),
            intermediate_size=int(g("intermediate_size", 1536)),
            quant_factor=int(g("quant_factor", 0)), face_quan_num=int(g("face_quan_num", 1)),
            neg=float(g("neg", 0.2)), INaffine=bool(g("INaffine", False)),
            n_embed=int(g("n_embed", 512)), zquant_dim=int(g("zquant_dim", 128)))


@dataclass(frozen=True)
class S2SConfig:
    """Hyper-parameters hard-coded in SLMFT.__init__ (seq2seq_pretrain.py:369-418)."""
    dim_in: int = 56
    dim: int = 384
    dim_audio: int = 768
    depth: int = 4
    heads: int = 12
    dim_head: int = 64          # x-transformers default, independent of dim
    max_seq_len: int = 2048
    num_tokens: int = 512
    ff_mult: int = 4
    top_k_frac: float = 0.1     # AutoregressiveWrapper.generate default filter

    @property
    def dec_dim(self) -> int:
        return self.dim + self.dim_audio

    @property
    def inner(self) -> int:
        return self.heads * self.dim_head


def vq_transformer_keys(prefix: str, c: VQConfig) -> "OrderedDict[str, tuple]":
    H, F = c.hidden_size, c.intermediate_size
    d = OrderedDict()
    for l in range(c.num_hidden_layers):
        a, m = f"{prefix}.net.{2*l}.fn", f"{prefix}.net.{2*l+1}.fn"
        d[f"{a}.norm.weight"] = (H,)
        d[f"{a}.norm.bias"] = (H,)
        d[f"{a}.fn.to_qkv.weight"] = (3 * H, H)
        d[f"{a}.fn.to_out.weight"] = (H, H)
        d[f"{a}.fn.to_out.bias"] = (H,)
        d[f"{m}.norm.weight"] = (H,)
        d[f"{m}.norm.bias"] = (H,)
        d[f"{m}.fn.l1.weight"] = (F, H)
        d[f"{m}.fn.l1.bias"] = (F,)
        d[f"{m}.fn.l2.weight"] = (H, F)
        d[f"{m}.fn.l2.bias"] = (H,)
    return d


def vqvae_schema(c: VQConfig = VQConfig()) -> "OrderedDict[str, tuple]":
    """Keys of VQAutoEncoder.state_dict() for quant_factor == 0 (the DIM configuration)."""
    assert c.quant_factor == 0, "only quant_factor 0 (DIM's config.yaml) is on the hot path"
    H, Z = c.hidden_size, c.face_quan_num * c.zquant_dim
    d = OrderedDict()
    d["encoder.vertice_mapping.0.weight"] = (H, c.in_dim)
    d["encoder.vertice_mapping.0.bias"] = (H,)
    d["encoder.squasher.0.0.weight"] = (H, H, 5)
    d["encoder.squasher.0.0.bias"] = (H,)
    d.update(vq_transformer_keys("encoder.encoder_transformer", c))
    d["encoder.encoder_pos_embedding.pe"] = (c.pe_max_len, 1, H)
    d["encoder.encoder_linear_embedding.net.weight"] = (H, H)
    d["encoder.encoder_linear_embedding.net.bias"] = (H,)
    d["encoder.encoder_linear_embedding_post.net.weight"] = (Z, H)
    d["encoder.encoder_linear_embedding_post.net.bias"] = (Z,)
    d["decoder.expander.0.0.weight"] = (H, H, 5)
    d["decoder.expander.0.0.bias"] = (H,)
    d.update(vq_transformer_keys("decoder.decoder_transformer", c))
    d["decoder.decoder_pos_embedding.pe"] = (c.pe_max_len, 1, H)
    d["decoder.decoder_linear_embedding.net.weight"] = (H, H)
    d["decoder.decoder_linear_embedding.net.bias"] = (H,)
    d["decoder.decoder_linear_embedding_pre.net.weight"] = (H, Z)
    d["decoder.decoder_linear_embedding_pre.net.bias"] = (H,)
    d["decoder.vertice_map_reverse.weight"] = (c.in_dim, H)
    d["quantize.embedding.weight"] = (c.n_embed, c.zquant_dim)
    return d


def _vq_decoder_keys(name: str, c: VQConfig, out_dim: int) -> "OrderedDict[str, tuple]":
    H, Z = c.hidden_size, c.face_quan_num * c.zquant_dim
    d = OrderedDict()
    d[f"{name}.expander.0.0.weight"] = (H, H, 5)
    d[f"{name}.expander.0.0.bias"] = (H,)
    d.update(vq_transformer_keys(f"{name}.decoder_transformer", c))
    d[f"{name}.decoder_pos_embedding.pe"] = (c.pe_max_len, 1, H)
    d[f"{name}.decoder_linear_embedding.net.weight"] = (H, H)
    d[f"{name}.decoder_linear_embedding.net.bias"] = (H,)
    d[f"{name}.decoder_linear_embedding_pre.net.weight"] = (H, Z)
    d[f"{name}.decoder_linear_embedding_pre.net.bias"] = (H,)
    d[f"{name}.vertice_map_reverse.weight"] = (out_dim, H)
    return d


# code/config_speaker_old.yaml:15-30 (arch stage1_BIWI_speaker): motion (56) + audio (768) frames, 8 codes per frame
SPEAKER_VQ = VQConfig(in_dim=824, hidden_size=768, face_quan_num=8)
SPEAKER_OUT_DIMS = (("decoder_v", 56), ("decoder_a", 768))           # stage1_BIWI.py:146-147


def vqspeaker_schema(c: VQConfig = SPEAKER_VQ) -> "OrderedDict[str, tuple]":
    """Keys of VQSpeakerAutoEncoder.state_dict() (stage1_BIWI.py:140-151): one encoder, two decoders, one codebook."""
    assert c.quant_factor == 0
    full = vqvae_schema(c)
    d = OrderedDict((k, v) for k, v in full.items() if k.startswith("encoder."))
    for name, out_dim in SPEAKER_OUT_DIMS:
        d.update(_vq_decoder_keys(name, c, out_dim))
    d["quantize.embedding.weight"] = full["quantize.embedding.weight"]
    return d


def xt_layers_schema(prefix: str, dim: int, depth: int, inner: int, cross: bool, ff_mult: int = 4):
    """x-transformers AttentionLayers keys. Layer order ('a','f')*depth or ('a','c','f')*depth."""
    d = OrderedDict()
    kinds = (("a", "c", "f") if cross else ("a", "f")) * depth
    for i, kind in enumerate(kinds):
        p = f"{prefix}.layers.{i}"
        d[f"{p}.0.0.weight"] = (dim,)                     # bias-free pre-norm gain
        if kind in ("a", "c"):
            d[f"{p}.1.to_q.weight"] = (inner, dim)
            d[f"{p}.1.to_k.weight"] = (inner, dim)
            d[f"{p}.1.to_v.weight"] = (inner, dim)
            d[f"{p}.1.to_out.weight"] = (dim, inner)
        else:
            d[f"{p}.1.ff.0.0.weight"] = (ff_mult * dim, dim)
            d[f"{p}.1.ff.0.0.bias"] = (ff_mult * dim,)
            d[f"{p}.1.ff.2.weight"] = (dim, ff_mult * dim)
            d[f"{p}.1.ff.2.bias"] = (dim,)
    d[f"{prefix}.final_norm.weight"] = (dim,)
    return d


def xt_encoder_wrapper_schema(name: str, dim_in: int, c: S2SConfig):
    d = OrderedDict()
    d[f"{name}.project_in.weight"] = (c.dim, dim_in)
    d[f"{name}.pos_emb.emb.weight"] = (c.max_seq_len, c.dim)
    d.update(xt_layers_schema(f"{name}.attn_layers", c.dim, c.depth, c.inner, False, c.ff_mult))
    d[f"{name}.project_out.weight"] = (c.dim, c.dim)      # present in the ckpt, unused with return_embeddings
    return d


def slmft_own_schema(c: S2SConfig = S2SConfig()):
    """SLMFT keys excluding the two VQ-VAEs."""
    d = OrderedDict()
    d["patch_embed_s"] = (1, 1, c.dim_in)
    d["patch_embed_l"] = (1, 1, c.dim_in)
    d["patch_embed_dec_s"] = (1, 1, c.dim)
    d["patch_embed_dec_l"] = (1, 1, c.dim)
    d.update(xt_encoder_wrapper_schema("encoder_s", c.dim_in, c))
    d.update(xt_encoder_wrapper_schema("encoder_l", c.dim_in, c))
    d.update(xt_encoder_wrapper_schema("encoder_joint", c.dim, c))
    for n in ("norm_s", "norm_l", "norm"):
        d[f"{n}.weight"] = (c.dim,)
        d[f"{n}.bias"] = (c.dim,)
    D = c.dec_dim
    d["decoder_joint.net.token_emb.emb.weight"] = (c.num_tokens, D)
    d.update(xt_layers_schema("decoder_joint.net.attn_layers", D, c.depth, c.inner, True, c.ff_mult))
    d["decoder_joint.net.to_logits.weight"] = (c.num_tokens, D)
    return d


def slmft_schema(c: S2SConfig = S2SConfig(), vq: VQConfig = VQConfig()):
    d = OrderedDict()
    for pre in ("speaker_vq", "listener_vq"):
        for k, s in vqvae_schema(vq).items():
            d[f"{pre}.{k}"] = s
    d.update(slmft_own_schema(c))
    return d


MESH_SIZE = 70110       # BIWI vertex vector (23370 vertices x 3), hard-coded at seq2seq_pretrain.py:775
LSTM_HIDDEN = 384


def lstm_schema(name: str, in_dim: int, hidden: int, layers: int = 2, bidirectional: bool = True):
    """torch.nn.LSTM parameter names and shapes."""
    d = OrderedDict()
    nd = 2 if bidirectional else 1
    for k in range(layers):
        for sfx in ("", "_reverse")[:nd]:
            d[f"{name}.weight_ih_l{k}{sfx}"] = (4 * hidden, in_dim if k == 0 else nd * hidden)
            d[f"{name}.weight_hh_l{k}{sfx}"] = (4 * hidden, hidden)
            d[f"{name}.bias_ih_l{k}{sfx}"] = (4 * hidden,)
            d[f"{name}.bias_hh_l{k}{sfx}"] = (4 * hidden,)
    return d


def emoca_converter_own_schema(size: int = MESH_SIZE, dim: int = 56):
    """EmocaConverter's own modules (seq2seq_pretrain.py:775-813), i.e. without speaker_vq."""
    d = OrderedDict()
    d["vertice_mapping.0.weight"] = (dim, size)
    d["vertice_mapping.0.bias"] = (dim,)
    d["squasher.0.0.weight"] = (dim, dim, 5)
    d["squasher.0.0.bias"] = (dim,)
    d.update(lstm_schema("vertice_map_reverse_lstm", dim, LSTM_HIDDEN))
    d.update(lstm_schema("vertice_map_reverse_lstm_2", dim, LSTM_HIDDEN))
    for n in ("vertice_map_reverse", "vertice_map_reverse2"):
        d[f"{n}.0.weight"] = (2 * LSTM_HIDDEN, 2 * LSTM_HIDDEN)
        d[f"{n}.0.bias"] = (2 * LSTM_HIDDEN,)
        d[f"{n}.2.weight"] = (size, 2 * LSTM_HIDDEN)
        d[f"{n}.2.bias"] = (size,)
    return d


def speaker_slmft_own_schema(c: S2SConfig = S2SConfig(), size: int = MESH_SIZE):
    """SpeakerSLMFT keys excluding the two VQ-VAEs (seq2seq_pretrain.py:516-637): SLMFT's transformer stack with the decoder's
    absolute positional table, the converter's mesh modules (squasher and both LSTM heads are attributes of the model, :563-568),
    W (2,) and speaker_embed (15, 384)."""
    d = slmft_own_schema(c)
    d["decoder_joint.net.pos_emb.emb.weight"] = (c.max_seq_len, c.dec_dim)
    d.update(emoca_converter_own_schema(size, c.dim_in))
    d["W"] = (2,)
    d["speaker_embed.weight"] = (15, c.dim)
    return d
