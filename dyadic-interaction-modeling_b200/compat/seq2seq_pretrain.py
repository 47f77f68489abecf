"""seq2seq_pretrain.SLMFT on the B200 kernels (reference: code/seq2seq_pretrain.py:325-514).

Same zero-argument constructor (reads ./config.yaml and the two VQ checkpoints from the working directory), same
sub-module names and state_dict keys, same `forward(v_speaker, v_listener, v_audio, mask, mode, ...)` contract returning
`(total_loss, dict, pred_cont_seq_l)`.  mode='val' is the inference hot path and runs entirely in libdimb200:

    forward_vq           listener VQ encode of each sample's valid prefix (batched, lens + batch slot 0)   :480-494
    forward_encoder      encoder_s -> encoder_joint -> norm_s  (+ context concat)                          :431-446
    forward_decoder      decoder_joint.generate: KV-cached AR decoding, top-k(52) sampling, T-1 steps      :444-452
    forward_vq_decoder   codebook gather fused into listener_vq.decode (batched: sample b gets pe[b], F4)  :454-464

Differences in WORK, not in results: the speaker VQ encodes and the duplicated forward_vq call, whose outputs the reference
discards, are skipped; cross-attention K/V are projected once (F9).  Sampling draws come from torch.rand on
the device (one uniform per step and clip) instead of torch.multinomial's stream: same distribution, different stream;
set `self.decode_uniforms` (B,T-1) or `self.greedy = True` for reproducible decoding.
mode='train' runs the teacher-forced FORWARD pass (logits, CE + continuous loss, argmax decode) like
x_engine_pt.evaluate_finetune_epoch needs; no autograd graph is built (fine-tuning itself is out of scope).
"""
import os

import torch
import torch.nn as nn

from base import config
from base.baseTrainer import load_state_dict
from models import get_model
from x_transformers import (AutoregressiveWrapper, ContinuousAutoregressiveWrapper, ContinuousTransformerWrapper,  # noqa: F401
                            Decoder, Encoder, TransformerWrapper)
from x_utils import *  # noqa: F401,F403

from dim_b200 import compat_api
from dim_b200.engine import PREC_BF16, PREC_FP32, PREC_FP32_TC, Handle, SLMFTEngine, VQEngine  # noqa: F401
from dim_b200.paramtree import fingerprint
from dim_b200.schema import S2SConfig, VQConfig


class SLMFT(nn.Module):
    precision = PREC_FP32_TC          # class-level switch: PREC_FP32 (FFMA), PREC_FP32_TC (parity, tensor cores), PREC_BF16

    def __init__(self, config_path="./config.yaml", model_speaker_pth="./runs_speaker_new/_RANK0/model/model.pth.tar",
                 model_listener_pth="./runs/listener_exp/model/model.pth.tar", load_vq_checkpoints=True):
        super().__init__()
        config_speaker = config.load_cfg_from_cfg_file(config_path)
        config_listener = config.load_cfg_from_cfg_file(config_path)
        model_speaker, model_listener = get_model(config_speaker), get_model(config_listener)
        if load_vq_checkpoints:
            to_cpu = lambda storage, loc: storage.cpu()
            load_state_dict(model_speaker, torch.load(model_speaker_pth, map_location=to_cpu)["state_dict"])
            load_state_dict(model_listener, torch.load(model_listener_pth, map_location=to_cpu)["state_dict"])
            print("Load models successfully")
        self.speaker_face_quan_num = config_speaker.face_quan_num
        self.speaker_zquant_dim = config_speaker.zquant_dim
        self.speaker_vq, self.listener_vq = model_speaker.eval(), model_listener.eval()
        for p in list(self.speaker_vq.parameters()) + list(self.listener_vq.parameters()):
            p.requires_grad = False

        dim_in, dim, enc_max_seq_len, dim_a = 56, 384, 2048, 768
        enc_kwargs = {"depth": 4, "heads": 12, "max_seq_len": 2048}
        dec_kwargs = {"depth": 4, "heads": 12, "max_seq_len": 2048, "num_tokens": 512}
        dec_transformer_kwargs = pick_and_pop(["num_tokens", "max_seq_len"], dec_kwargs)  # noqa: F405
        dec_transformer_kwargs["emb_dropout"] = dec_kwargs.pop("emb_dropout", 0)
        dec_transformer_kwargs["scaled_sinu_pos_emb"] = dec_kwargs.pop("scaled_sinu_pos_emb", False)
        dec_transformer_kwargs["use_abs_pos_emb"] = dec_kwargs.pop("use_abs_pos_emb", False)
        mk_enc = lambda d_in: ContinuousTransformerWrapper(dim_in=d_in, dim_out=dim, max_seq_len=enc_max_seq_len,
                                                           attn_layers=Encoder(dim=dim, **enc_kwargs))
        self.encoder_s, self.encoder_l, self.encoder_joint = mk_enc(dim_in), mk_enc(dim_in), mk_enc(dim)
        self.patch_embed_s = nn.Parameter(torch.zeros(1, 1, dim_in))
        self.patch_embed_l = nn.Parameter(torch.zeros(1, 1, dim_in))
        self.patch_embed_dec_s = nn.Parameter(torch.zeros(1, 1, dim))
        self.patch_embed_dec_l = nn.Parameter(torch.zeros(1, 1, dim))
        self.norm_s, self.norm_l, self.norm = nn.LayerNorm(dim), nn.LayerNorm(dim), nn.LayerNorm(dim)
        self.decoder_joint = TransformerWrapper(**dec_transformer_kwargs,
                                                attn_layers=Decoder(dim=dim + dim_a, cross_attend=True, **dec_kwargs))
        self.decoder_joint = AutoregressiveWrapper(self.decoder_joint, ignore_index=-100, pad_value=0, mask_prob=0.15)
        self.decoder_joint.bind(self._generate, self._teacher_forced)

        self._cfg, self._vq_cfg = S2SConfig(), VQConfig.from_cfg(config_listener)
        self._engines, self._fp = None, None
        self.greedy = False             # True: argmax decoding (deterministic parity mode)
        self.decode_uniforms = None     # optional (B, T-1) uniforms for reproducible sampling
        self.last_codes = None          # generated code sequences of the last val forward (B, T-1) int64
        self.train_kv_mask = None       # optional (B, T-1) bool key mask for mode='train' (None: drawn at random, mask_prob 0.15)

    # ---- engine binding ----
    def engines(self):
        fp = fingerprint(self)
        if self._engines is None or fp != self._fp:
            dev = next(self.parameters()).device
            if dev.type != "cuda":
                raise RuntimeError("SLMFT runs on CUDA only (sm_100a kernels, no CPU fallback): call .to('cuda') first")
            h = Handle(dev.index)
            h.register(self.state_dict())
            vq_prec = PREC_FP32 if self.precision == PREC_FP32 else PREC_FP32_TC
            self._engines = (SLMFTEngine(h, self._cfg, precision=self.precision),
                             VQEngine(h, self._vq_cfg, prefix="listener_vq.", precision=vq_prec))
            self._fp = fp
        return self._engines

    def _teacher_forced(self, context, context_mask, inp, kv_mask):
        return self.engines()[0].teacher_forced(context.contiguous(), context_mask, inp, kv_mask)

    def _generate(self, prompts, seq_len, temperature=1.0, context=None, context_mask=None, uniforms=None):
        s2s, _ = self.engines()
        if self.greedy or temperature == 0.0:
            return s2s.generate(context, context_mask, prompts, seq_len, temperature=0.0)
        if uniforms is None:
            uniforms = self.decode_uniforms if self.decode_uniforms is not None else \
                torch.rand(prompts.shape[0], seq_len, device=context.device)
        return s2s.generate(context, context_mask, prompts, seq_len, temperature=temperature, uniforms=uniforms)

    # ---- reference methods ----
    def forward_encoder(self, v_speaker, mask):
        """-> x_s (B,T,384) = norm_s(encoder_joint(encoder_s(v + patch_embed_s)))."""
        s2s, _ = self.engines()
        return s2s.context(v_speaker.float(), None, mask, want="x_s")

    def forward_decoder(self, x_s, z_l, x_a, mask, mode):
        x_s = torch.cat([x_s + self.patch_embed_dec_s, x_a], dim=-1)
        if mode == "train":                                 # :447-448, forward only
            s2s, _ = self.engines()
            inp, target = z_l[:, :-1].clone(), z_l[:, 1:]
            inp[inp == -100] = 0
            kv = self.train_kv_mask if self.train_kv_mask is not None else compat_api.draw_kv_mask(inp.shape, 0.15, inp.device)
            logits = s2s.teacher_forced(x_s.contiguous(), mask, inp, kv)
            return torch.nn.functional.cross_entropy(logits.transpose(1, 2), target, ignore_index=-100), logits
        px_l = self.decoder_joint.generate(z_l[:, 0].unsqueeze(1), seq_len=z_l.shape[1] - 1, context=x_s, context_mask=mask)
        return 0.0, px_l

    def forward_vq_decoder(self, logits_l, mode="train", batch_index=None):
        pred_seq_l = torch.argmax(logits_l, dim=-1) if mode == "train" else logits_l
        _, vq = self.engines()
        return vq.decode(codes=pred_seq_l.contiguous(), batch_index=batch_index)

    def forward_continuous_loss(self, pred, target, mask):
        return compat_api.continuous_loss(pred, target, mask)

    def forward_vq(self, v_speaker, v_listener, mask):
        """-> (z_speaker, z_listener) (B,T) int64.  z_speaker is a zero placeholder: the reference computes and never
        reads it in either mode; z_listener is padded with -100 like :490."""
        _, vq = self.engines()
        z_l = compat_api.listener_codes(vq, v_listener.float(), mask)
        return torch.zeros_like(z_l), z_l

    def forward_val_samples(self, v_speaker, v_listener, v_audio, mask, samples, uniforms=None, batch_index=None):
        """`samples` stochastic mode='val' generations per clip in one pass -> (pred (B,samples,T-1,56), codes (B,samples,T-1)).
        Equal, sample by sample, to calling forward(mode='val') `samples` times (x_engine_pt.py:257-258) with the same draws;
        the encoders and the cross-attention K/V projection run once per clip instead of once per call."""
        s2s, vq = self.engines()
        return compat_api.slmft_forward_val_samples(s2s, vq, v_speaker.float(), v_listener.float(), v_audio.float(), mask, samples,
                                                    uniforms=uniforms, batch_index=batch_index)

    def forward(self, v_speaker, v_listener, v_audio, mask, mode="train", speaker_ids=None, listener_ids=None,
                batch_index=None):
        s2s, vq = self.engines()
        if mode == "train":
            # teacher forcing (what x_engine_pt.evaluate_finetune_epoch:217 runs): FORWARD ONLY -- the returned loss carries no
            # autograd graph, fine-tuning (backward) is outside the path built here.  self.train_kv_mask pins the random key mask.
            return compat_api.slmft_forward_train(s2s, vq, v_speaker.float(), v_listener.float(), v_audio.float(), mask,
                                                  kv_mask=self.train_kv_mask, mask_prob=0.15, batch_index=batch_index)
        uniforms = None if self.greedy else (self.decode_uniforms if self.decode_uniforms is not None else
                                             torch.rand(v_speaker.shape[0], v_speaker.shape[1] - 1, device=v_speaker.device))
        loss, d, pred, codes = compat_api.slmft_forward_val(s2s, vq, v_speaker.float(), v_listener.float(), v_audio.float(),
                                                            mask, temperature=1.0, uniforms=uniforms, batch_index=batch_index,
                                                            return_codes=True, greedy=self.greedy)
        self.last_codes = codes
        return loss, d, pred


class SLM(SLMFT):
    """seq2seq_pretrain.SLM, the DIM PRE-TRAINING model (reference: code/seq2seq_pretrain.py:72-323), forward only: three
    encoders (speaker, listener, joint -- no causal mask here), 15 % random masking of both streams, the speaker<->listener
    contrastive NCE, and two teacher-forced passes of decoder_joint (which keeps its absolute positional table in this model,
    :137) predicting each side's masked VQ codes from the OTHER side's joint representation + audio.  Runs on the same kernels as
    SLMFT: dim_slmft_encode (libdimb200) per encoder call, dim_slmft_teacher_forced per decoder pass, the two VQ-VAEs.
    `self.mask_speaker` / `self.mask_listener` (B,T) bool pin the random masks (None: drawn with torch.randperm like :171-183).
    Returns (total_loss, dict, None) like :323.  No autograd graph is built (pre-training itself is out of scope)."""

    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        dim, dim_a = 384, 768
        dec_kwargs = {"depth": 4, "heads": 12, "max_seq_len": 2048, "num_tokens": 512}
        kw = pick_and_pop(["num_tokens", "max_seq_len"], dec_kwargs)  # noqa: F405
        kw.update(emb_dropout=0, scaled_sinu_pos_emb=False, use_abs_pos_emb=True)                      # :135-137
        self.decoder_joint = AutoregressiveWrapper(TransformerWrapper(**kw, attn_layers=Decoder(dim=dim + dim_a, cross_attend=True, **dec_kwargs)),
                                                   ignore_index=-100, pad_value=0)                     # :163-165 (mask_prob default 0)
        self.decoder_joint.bind(self._generate, self._teacher_forced)
        self.mask_speaker = self.mask_listener = None
        self.last_parts = None

    def engines(self):
        fp = fingerprint(self)
        if self._engines is None or fp != self._fp:
            dev = next(self.parameters()).device
            if dev.type != "cuda":
                raise RuntimeError("SLM runs on CUDA only (sm_100a kernels, no CPU fallback): call .to('cuda') first")
            h = Handle(dev.index)
            h.register(self.state_dict())
            vq_prec = PREC_FP32 if self.precision == PREC_FP32 else PREC_FP32_TC
            self._engines = (SLMFTEngine(h, self._cfg, precision=self.precision),
                             VQEngine(h, self._vq_cfg, prefix="listener_vq.", precision=vq_prec),
                             VQEngine(h, self._vq_cfg, prefix="speaker_vq.", precision=vq_prec))
            self._fp = fp
        return self._engines

    def random_masking_unstructured(self, x, mask, mask_ratio):
        return compat_api.random_masking_unstructured(mask, mask_ratio)

    def forward_contrastive(self, s_rep, l_rep, mask, bidirect_contrast=False):
        nce, acc = compat_api.contrastive(s_rep, l_rep, mask)
        if bidirect_contrast:
            nce2, acc2 = compat_api.contrastive(l_rep, s_rep, mask)
            return (nce + nce2) / 2, (acc + acc2) / 2
        return nce, acc

    def forward(self, v_speaker, v_listener, v_audio, mask, speaker_ids=None, listener_ids=None, mode="train"):
        s2s, vq_l, vq_s = self.engines()
        total, d, parts = compat_api.slm_forward(s2s, vq_s, vq_l, v_speaker.float(), v_listener.float(), v_audio.float(), mask,
                                                 self.patch_embed_s.detach(), self.patch_embed_l.detach(),
                                                 self.patch_embed_dec_s.detach(), self.patch_embed_dec_l.detach(),
                                                 mask_speaker=self.mask_speaker, mask_listener=self.mask_listener, return_parts=True)
        self.last_parts = parts
        return total, d, None


def _mesh_modules(owner, size, dim=56):
    """The converter's modules with the reference's names (seq2seq_pretrain.py:777-813); parameter containers only."""
    owner.vertice_mapping = nn.Sequential(nn.Linear(size, dim), nn.LeakyReLU(0.2, True))
    owner.squasher = nn.Sequential(nn.Sequential(nn.Conv1d(dim, dim, 5, stride=1, padding=2, padding_mode="replicate"),
                                                 nn.LeakyReLU(0.2, True), nn.InstanceNorm1d(dim, affine=False)))
    lstm = lambda: nn.LSTM(input_size=dim, hidden_size=384, num_layers=2, batch_first=True, bidirectional=True)
    head = lambda: nn.Sequential(nn.Linear(768, 768), nn.LeakyReLU(0.2, True), nn.Linear(768, size))
    owner.vertice_map_reverse_lstm, owner.vertice_map_reverse_lstm_2 = lstm(), lstm()
    owner.vertice_map_reverse, owner.vertice_map_reverse2 = head(), head()


def _mesh_head(owner, which="emoca"):
    lstm, head = (owner.vertice_map_reverse_lstm, owner.vertice_map_reverse) if which == "emoca" else \
        (owner.vertice_map_reverse_lstm_2, owner.vertice_map_reverse2)
    return dict(lstm.named_parameters()), head[0].weight, head[0].bias, head[2].weight, head[2].bias


class EmocaConverter(nn.Module):
    """seq2seq_pretrain.EmocaConverter (reference: code/seq2seq_pretrain.py:759-825) on the B200 kernels: the frozen 56-d speaker
    VQ-VAE followed by the 2-layer bidirectional LSTM(384) and the 768 -> 768 -> 70110 vertex head.  Same module names and
    state_dict keys; forward(inputs, template, v_speaker) -> (outputs (B,T,size), None) like :815-832 (`inputs` is unused there too).
    `size` exposes the hard-coded 70110 (:775) for small tests."""
    precision = PREC_FP32_TC

    def __init__(self, config_path="./config.yaml", model_speaker_pth="./runs_speaker_new/_RANK0/model/model.pth.tar",
                 load_vq_checkpoints=True, size=70110):
        super().__init__()
        config_speaker = config.load_cfg_from_cfg_file(config_path)
        model_speaker = get_model(config_speaker)
        if load_vq_checkpoints:
            ckpt = torch.load(model_speaker_pth, map_location=lambda storage, loc: storage.cpu())
            load_state_dict(model_speaker, ckpt["state_dict"])
            print("Load models successfully")
        self.speaker_face_quan_num, self.speaker_zquant_dim = config_speaker.face_quan_num, config_speaker.zquant_dim
        self.speaker_vq = model_speaker.eval()
        for p in self.speaker_vq.parameters():
            p.requires_grad = False
        _mesh_modules(self, size)
        self.criterion = nn.MSELoss()

    def forward_motion(self, inputs, template):
        """vertice_mapping + squasher on (inputs - template) (the lines commented out at :817-819; live in SpeakerSLMFT :710-713)."""
        return compat_api.mesh_to_motion(inputs.float(), template.float(), self.vertice_mapping[0].weight, self.vertice_mapping[0].bias,
                                         self.squasher[0][0].weight, self.squasher[0][0].bias)

    @torch.no_grad()
    def forward(self, inputs, template, v_speaker):
        self.speaker_vq.precision = PREC_FP32 if self.precision == PREC_FP32 else PREC_FP32_TC
        dec, _, _ = self.speaker_vq(v_speaker)                                           # :820
        return compat_api.motion_to_mesh(dec, *_mesh_head(self), template=template.float()), None


class SpeakerSLMFT(nn.Module):
    """seq2seq_pretrain.SpeakerSLMFT (reference: code/seq2seq_pretrain.py:516-757), forward only, on the B200 kernels: listener-VQ
    codes of the speaker's EMOCA coefficients -> decoder_joint (teacher forced in mode='train', generate otherwise) over
    cat(speaker embedding | zeros, audio) -> speaker-VQ decode -> LSTM + vertex head -> losses.  Same module names / state_dict
    keys as the reference (its three encoders and patch embeddings exist and are never called in forward, like there).
    Paths the reference hard-codes are constructor arguments with its values as defaults; `mouth_map` (list of vertex ids) replaces
    reading ../data/CodeTalker/BIWI/regions/lve.txt when given.  NOTE (:738-739): the mouth loss drops one row of the FLATTENED
    batch, so like the reference the call only works for B == 1 (or when torch can broadcast the two row counts)."""
    precision = PREC_FP32_TC

    def __init__(self, config_path="./config.yaml", model_speaker_pth="./runs_speaker_new/_RANK0/model/model.pth.tar",
                 model_listener_pth="./runs/listener_exp/model/model.pth.tar", model_converter_path="./best_converter.pt",
                 mouth_map_path="../data/CodeTalker/BIWI/regions/lve.txt", load_checkpoints=True, size=70110, mouth_map=None):
        super().__init__()
        cfg_s, cfg_l = config.load_cfg_from_cfg_file(config_path), config.load_cfg_from_cfg_file(config_path)
        model_speaker, model_listener = get_model(cfg_s), get_model(cfg_l)
        _mesh_modules(self, size)
        if load_checkpoints:
            to_cpu = lambda storage, loc: storage.cpu()
            load_state_dict(model_speaker, torch.load(model_speaker_pth, map_location=to_cpu)["state_dict"])
            load_state_dict(model_listener, torch.load(model_listener_pth, map_location=to_cpu)["state_dict"])
            print("Load models successfully")
            conv = torch.load(model_converter_path, map_location=to_cpu)                  # EmocaConverter state_dict (:549-552)
            own = {k: v for k, v in conv.items() if not k.startswith("speaker_vq.")}
            self.load_state_dict(own, strict=False)
        self.speaker_face_quan_num, self.speaker_zquant_dim = cfg_s.face_quan_num, cfg_s.zquant_dim
        self.listener_vq, self.speaker_vq = model_listener.eval(), model_speaker.eval()
        for p in list(self.listener_vq.parameters()) + list(self.speaker_vq.parameters()):
            p.requires_grad = False

        dim_in, dim, enc_max_seq_len, dim_a = 56, 384, 2048, 768
        enc_kwargs = {"depth": 4, "heads": 12, "max_seq_len": 2048}
        dec_kwargs = {"depth": 4, "heads": 12, "max_seq_len": 2048, "num_tokens": 512}
        kw = pick_and_pop(["num_tokens", "max_seq_len"], dec_kwargs)  # noqa: F405
        kw.update(emb_dropout=0, scaled_sinu_pos_emb=False, use_abs_pos_emb=True)          # :587-589
        mk_enc = lambda d_in: ContinuousTransformerWrapper(dim_in=d_in, dim_out=dim, max_seq_len=enc_max_seq_len,
                                                           attn_layers=Encoder(dim=dim, **enc_kwargs))
        self.encoder_s, self.encoder_l, self.encoder_joint = mk_enc(dim_in), mk_enc(dim_in), mk_enc(dim)
        self.patch_embed_s = nn.Parameter(torch.zeros(1, 1, dim_in))
        self.patch_embed_l = nn.Parameter(torch.zeros(1, 1, dim_in))
        self.patch_embed_dec_s = nn.Parameter(torch.zeros(1, 1, dim))
        self.patch_embed_dec_l = nn.Parameter(torch.zeros(1, 1, dim))
        self.norm_s, self.norm_l, self.norm = nn.LayerNorm(dim), nn.LayerNorm(dim), nn.LayerNorm(dim)
        self.decoder_joint = AutoregressiveWrapper(TransformerWrapper(**kw, attn_layers=Decoder(dim=dim + dim_a, cross_attend=True, **dec_kwargs)),
                                                   ignore_index=-100, pad_value=0)
        self.mse_loss = nn.MSELoss()
        if mouth_map is None:
            with open(mouth_map_path) as f:
                mouth_map = [int(i) for i in f.read().split(", ")]
        self.mouth_map = list(mouth_map)
        self.W = nn.Parameter(torch.randn(2))
        self.speaker_embed = nn.Embedding(15, 384)

        self._cfg, self._vq_cfg = S2SConfig(), VQConfig.from_cfg(cfg_l)
        self._engines, self._fp = None, None
        self.greedy, self.decode_uniforms, self.last_parts = False, None, None

    def engines(self):
        fp = fingerprint(self)
        if self._engines is None or fp != self._fp:
            dev = next(self.parameters()).device
            if dev.type != "cuda":
                raise RuntimeError("SpeakerSLMFT runs on CUDA only (sm_100a kernels, no CPU fallback): call .to('cuda') first")
            h = Handle(dev.index)
            h.register(self.state_dict())
            vq_prec = PREC_FP32 if self.precision == PREC_FP32 else PREC_FP32_TC
            self._engines = (SLMFTEngine(h, self._cfg, precision=self.precision),
                             VQEngine(h, self._vq_cfg, prefix="speaker_vq.", precision=vq_prec),
                             VQEngine(h, self._vq_cfg, prefix="listener_vq.", precision=vq_prec))
            self._mouth = torch.as_tensor(self.mouth_map, dtype=torch.long, device=dev)
            self._fp = fp
        return self._engines

    def forward_motion(self, v_speaker, template):
        """:710-713: vertice_mapping + squasher (their result only feeds the dead z_s in forward)."""
        return compat_api.mesh_to_motion(v_speaker.float(), template.float(), self.vertice_mapping[0].weight, self.vertice_mapping[0].bias,
                                         self.squasher[0][0].weight, self.squasher[0][0].bias)

    def forward_vq_decoder(self, logits_s, type="emoca", mode="train"):
        _, vq_s, _ = self.engines()
        codes = torch.argmax(logits_s, dim=-1) if mode == "train" else logits_s
        pred_emoca = vq_s.decode(codes=codes.contiguous())
        return compat_api.motion_to_mesh(pred_emoca, *_mesh_head(self, type)), pred_emoca

    def forward(self, v_speaker, v_speaker_emoca, v_audio, mask, template, mode="train", speaker_ids=None):
        s2s, vq_s, vq_l = self.engines()
        rows = None if speaker_ids is None else self.speaker_embed.weight.detach()[speaker_ids]
        uniforms = None if self.greedy else self.decode_uniforms
        total, d, pred_emoca, parts = compat_api.speaker_slmft_forward(
            s2s, vq_s, vq_l, _mesh_head(self), v_speaker.float(), v_speaker_emoca.float(), v_audio.float(), mask, template.float(),
            self.patch_embed_dec_l.detach(), self._mouth, mode=mode, speaker_rows=rows, uniforms=uniforms, greedy=self.greedy)
        self.last_parts = parts
        return total, d, pred_emoca
