"""seq2seq.ListenerGenerator on the B200 kernels (reference: code/seq2seq.py:13-75 `Transformer`, :138-290 `ListenerGenerator`), the
older DIM listener model: the speaker VQ-VAE (VQSpeakerAutoEncoder, 8 codes per frame) turns the 824-d speaker frames
into 1024-d quantised latents, an x-transformers encoder (dim 512, depth 6, 8 heads) reads them, and a decoder (dim 512, depth 6,
8 heads, cross-attention, absolute positional table) predicts the listener's VQ codes, which listener_vq decodes to 56-d frames.

Same constructor files, sub-module names and state_dict keys as the reference.  The arithmetic runs in libdimb200 through the same
entry points as SLMFT: the generator's tensors are registered as `encoder_s.*` / `decoder_joint.net.*` of a single-encoder model
(dim_slmft_build accepts a model without encoder_joint / patch embeddings / norm_s), one dim_slmft_encode call gives the context,
dim_slmft_teacher_forced the logits of `forward`, dim_slmft_generate (persistent decode kernel, positional table added per step)
the codes of `generate`.  speaker_ids / listener_ids (identity tokens: :240-250) are supported in forward, like the reference (its generate takes none).  Forward only.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from base import config
from base.baseTrainer import load_state_dict
from models import get_model
from x_transformers import (AutoregressiveWrapper, ContinuousAutoregressiveWrapper, ContinuousTransformerWrapper,  # noqa: F401
                            Decoder, Encoder, TransformerWrapper)
from x_utils import *  # noqa: F401,F403

from dim_b200 import compat_api, ops
from dim_b200.engine import PREC_FP32, PREC_FP32_TC, Handle, SLMFTEngine, VQEngine
from dim_b200.paramtree import fingerprint
from dim_b200.schema import S2SConfig, VQConfig


class Transformer(nn.Module):
    """code/seq2seq.py:13-75: parameter containers with upstream's keys (`encoder.*`, `decoder.net.*`)."""

    def __init__(self, dim_in, dim, enc_max_seq_len, cross_attn_tokens_dropout=0., **kwargs):
        super().__init__()
        enc_kwargs, kwargs = groupby_prefix_and_trim("enc_", kwargs)  # noqa: F405
        dec_kwargs, kwargs = groupby_prefix_and_trim("dec_", kwargs)  # noqa: F405
        self.dims = dict(dim_in=dim_in, dim=dim, depth=enc_kwargs["depth"], heads=enc_kwargs["heads"], max_seq_len=enc_max_seq_len,
                         num_tokens=dec_kwargs["num_tokens"])
        assert dec_kwargs["depth"] == enc_kwargs["depth"] and dec_kwargs["heads"] == enc_kwargs["heads"], \
            "the engine shares depth / heads between encoder and decoder (true for ListenerGenerator: 6 / 8)"
        self.encoder = ContinuousTransformerWrapper(dim_in=dim_in, dim_out=dim, max_seq_len=enc_max_seq_len,
                                                    attn_layers=Encoder(dim=dim, **enc_kwargs))
        tk = pick_and_pop(["num_tokens", "max_seq_len"], dec_kwargs)  # noqa: F405
        tk.update(emb_dropout=dec_kwargs.pop("emb_dropout", 0), scaled_sinu_pos_emb=dec_kwargs.pop("scaled_sinu_pos_emb", False),
                  use_abs_pos_emb=dec_kwargs.pop("use_abs_pos_emb", True))
        self.decoder = AutoregressiveWrapper(TransformerWrapper(**tk, attn_layers=Decoder(dim=dim, cross_attend=True, **dec_kwargs)),
                                             ignore_index=-100, pad_value=0)


class ListenerGenerator(nn.Module):
    precision = PREC_FP32_TC

    def __init__(self, config_speaker_pth="./config_speaker_old.yaml", config_listener_pth="./config.yaml",
                 model_speaker_pth="./runs/speaker_exp/model/model.pth.tar", model_listener_pth="./runs/listener_exp/model/model.pth.tar",
                 load_vq_checkpoints=True):
        super().__init__()
        config_speaker = config.load_cfg_from_cfg_file(config_speaker_pth)
        config_listener = config.load_cfg_from_cfg_file(config_listener_pth)
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
        self.generator = Transformer(dim_in=self.speaker_face_quan_num * self.speaker_zquant_dim, dim=512, enc_depth=6, enc_heads=8,
                                     enc_max_seq_len=1024, dec_num_tokens=512, dec_depth=6, dec_heads=8, dec_max_seq_len=1024)
        self.speaker_embeddings = nn.Embedding(100, 256)
        self.listener_embeddings = nn.Embedding(100, 256)
        self.fc_speaker = nn.Linear(256, 1024)
        self.fc_listener = nn.Linear(256, 512)
        d = self.generator.dims
        self._cfg = S2SConfig(dim_in=d["dim_in"], dim=d["dim"], dim_audio=0, depth=d["depth"], heads=d["heads"], dim_head=64,
                              max_seq_len=d["max_seq_len"], num_tokens=d["num_tokens"], ff_mult=4)
        self._sp_cfg, self._li_cfg = VQConfig.from_cfg(config_speaker), VQConfig.from_cfg(config_listener)
        self._engines, self._fp = None, None
        self.greedy = True                  # generate(): argmax decoding unless decode_uniforms (B, T) is set
        self.decode_uniforms = None

    def engines(self):
        fp = fingerprint(self)
        if self._engines is None or fp != self._fp:
            dev = next(self.parameters()).device
            if dev.type != "cuda":
                raise RuntimeError("ListenerGenerator runs on CUDA only (sm_100a kernels, no CPU fallback): call .to('cuda') first")
            h = Handle(dev.index)
            sd = self.state_dict()
            h.register({k: v for k, v in sd.items() if k.startswith(("speaker_vq.", "listener_vq."))})
            # the generator under the names the single-encoder SLMFT model reads
            h.register({"encoder_s." + k[len("generator.encoder."):]: v for k, v in sd.items() if k.startswith("generator.encoder.")})
            h.register({"decoder_joint." + k[len("generator.decoder."):]: v for k, v in sd.items() if k.startswith("generator.decoder.")})
            vq_prec = PREC_FP32 if self.precision == PREC_FP32 else PREC_FP32_TC
            self._engines = (SLMFTEngine(h, self._cfg, precision=self.precision),
                             VQEngine(h, self._sp_cfg, prefix="speaker_vq.", precision=vq_prec, encoder="encoder", decoder=None),
                             VQEngine(h, self._li_cfg, prefix="listener_vq.", precision=vq_prec))
            self._fp = fp
        return self._engines

    @torch.no_grad()
    def _inputs(self, v_speaker, v_listener, mask):
        """The per-sample VQ loop of :225-236 as two batched encodes of each clip's valid prefix.  x_speaker follows the reference to
        the letter: the speaker `quant` tensors (1, 128, len*8) are zero-padded to (128, T*8), stacked, and VIEWED as (B, T, 8, 128) ->
        (B, T, 1024) without a permute (:238-239), i.e. row t is 1024 consecutive elements of the channel-major array, not frame t's
        latents.  z_listener (B,T) int64 with -100 on padding."""
        s2s, vq_s, vq_l = self.engines()
        B, T, _ = v_speaker.shape
        fqn = self.speaker_face_quan_num
        lens = mask.sum(dim=1).to(torch.int32)
        zero = torch.zeros(B, dtype=torch.int32, device=v_speaker.device)
        _, _, quant = vq_s.encode(v_speaker.float(), lens=lens, batch_index=zero, want_quant=True)    # (B, 128, T*8) = E[idx], channel-major
        keep = (torch.arange(T * fqn, device=quant.device)[None, :] < (lens.long() * fqn)[:, None]).unsqueeze(1)
        x_speaker = (quant * keep).contiguous().view(B, -1, fqn, self.speaker_zquant_dim).view(B, -1, fqn * self.speaker_zquant_dim)
        z_listener = compat_api.listener_codes(vq_l, v_listener.float(), mask)
        return x_speaker, z_listener

    def _identity_row(self, emb, fc, ids):
        """fc(relu(embedding(ids))) (:242, :248) -> (B, fc.out_features); the Linear runs in libdimb200 (dim_linear_f32)."""
        from dim_b200 import ops
        rows = torch.relu(emb.weight.detach()[ids]).contiguous()
        return ops.linear(rows, fc.weight.detach().contiguous(), fc.bias.detach())

    @torch.no_grad()
    def forward(self, v_speaker, v_listener, mask, speaker_ids=None, listener_ids=None):
        """:223-261.  Identity tokens (:240-250, Transformer.forward :47-67): the speaker token is prepended to the ENCODER INPUT (mask
        grows by one True), the listener token to the ENCODER OUTPUT (the decoder's context; mask grows again, the target gets a leading
        -100 and the first logit row is dropped)."""
        s2s, _, vq_l = self.engines()
        x_speaker, z_listener = self._inputs(v_speaker, v_listener, mask)
        B = mask.shape[0]
        one = torch.ones(B, 1, dtype=torch.bool, device=mask.device)
        mask_u, tgt = mask, z_listener
        if speaker_ids is not None:
            x_speaker = torch.cat([self._identity_row(self.speaker_embeddings, self.fc_speaker, speaker_ids).unsqueeze(1), x_speaker], dim=1)
            mask_u = torch.cat([one, mask_u], dim=1)
        enc = s2s.encode("encoder_s", x_speaker.contiguous(), mask_u)                               # generator.encoder(..., mask=mask)
        if listener_ids is not None:
            enc = torch.cat([self._identity_row(self.listener_embeddings, self.fc_listener, listener_ids).unsqueeze(1), enc], dim=1)
            mask_u = torch.cat([one, mask_u], dim=1)
            tgt = torch.cat([torch.full((B, 1), -100, dtype=tgt.dtype, device=tgt.device), tgt], dim=1)
        inp, target = tgt[:, :-1].clone(), tgt[:, 1:]
        inp[inp == -100] = 0
        logits = s2s.teacher_forced(enc.contiguous(), mask_u, inp.contiguous(), None)
        loss = F.cross_entropy(logits.transpose(1, 2), target, ignore_index=-100)
        if listener_ids is not None:
            logits = logits[:, 1:, :]
        pred_cont_seq = vq_l.decode(codes=torch.argmax(logits, dim=-1).contiguous())
        loss_cont = compat_api.continuous_loss(pred_cont_seq, v_listener.float(), mask)
        self.last_logits = logits
        return loss + loss_cont, pred_cont_seq

    @torch.no_grad()
    def generate(self, v_speaker, v_listener, mask):
        """-> (z_listener_pred (B, T) int64, z_listener (B, T)) like :263-290 (seq_len = T generated codes after the prompt code)."""
        s2s, _, _ = self.engines()
        x_speaker, z_listener = self._inputs(v_speaker, v_listener, mask)
        enc = s2s.encode("encoder_s", x_speaker, mask)
        T = v_speaker.shape[1]
        if self.greedy and self.decode_uniforms is None:
            pred = s2s.generate(enc, mask, z_listener[:, 0], T, temperature=0.0)
        else:
            u = self.decode_uniforms if self.decode_uniforms is not None else torch.rand(v_speaker.shape[0], T, device=enc.device)
            pred = s2s.generate(enc, mask, z_listener[:, 0], T, temperature=1.0, uniforms=u)
        return pred, z_listener
