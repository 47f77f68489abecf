"""Oracle: EmocaConverter / SpeakerSLMFT (test infrastructure; never imported by the product path).

Follows /root/reference/code/seq2seq_pretrain.py:
  EmocaConverter modules   :775-813   Linear(70110,56)+LeakyReLU, squasher (Conv1d k5 replicate + LeakyReLU + InstanceNorm1d),
                                      nn.LSTM(56, 384, 2 layers, bidirectional), Linear(768,768)+LeakyReLU+Linear(768,70110)
  EmocaConverter.forward   :815-832
  SpeakerSLMFT.forward     :707-757   (forward_vq :693-705, forward_decoder :639-647, forward_vq_decoder :649-663)
PARITY: the mesh modules ARE torch.nn modules (nn.LSTM, nn.Linear, nn.Conv1d, nn.InstanceNorm1d: the reference's own dependency,
run here on the CPU) loaded with the state_dict -- pinned, and checked against the real EmocaConverter class by
tests/golden/make_emoca_golden.py; the VQ half is pinned (oracle/vqvae.py); the decoder_joint half is UNPINNED (oracle/xt.py).
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import slmft as S
from . import vqvae as V
from . import xt as X


def _lstm(sd, name, in_dim, hidden=384):
    m = nn.LSTM(input_size=in_dim, hidden_size=hidden, num_layers=2, batch_first=True, bidirectional=True)
    m.load_state_dict({k[len(name) + 1:]: v for k, v in sd.items() if k.startswith(name + ".")}, strict=True)
    return m.eval()


@torch.no_grad()
def mesh_to_motion(sd, v, template):
    """:710-713."""
    x = v - template.unsqueeze(1)
    x = F.leaky_relu(F.linear(x, sd["vertice_mapping.0.weight"], sd["vertice_mapping.0.bias"]), 0.2)
    x = x.permute(0, 2, 1)
    x = F.conv1d(F.pad(x, (2, 2), mode="replicate"), sd["squasher.0.0.weight"], sd["squasher.0.0.bias"])
    x = F.instance_norm(F.leaky_relu(x, 0.2), eps=1e-5)
    return x.permute(0, 2, 1)


@torch.no_grad()
def motion_to_mesh(sd, dec, which="emoca", template=None):
    """:657-658 / :823-825."""
    lstm, head = ("vertice_map_reverse_lstm", "vertice_map_reverse") if which == "emoca" else \
        ("vertice_map_reverse_lstm_2", "vertice_map_reverse2")
    h, _ = _lstm(sd, lstm, dec.shape[-1])(dec)
    h = F.leaky_relu(F.linear(h, sd[head + ".0.weight"], sd[head + ".0.bias"]), 0.2)
    out = F.linear(h, sd[head + ".2.weight"], sd[head + ".2.bias"])
    return out if template is None else out + template.unsqueeze(1)


@torch.no_grad()
def emoca_converter_forward(sd, template, v_speaker, vq_cfg):
    """EmocaConverter.forward (:815-832): dec = speaker_vq(v_speaker)[0] -> LSTM -> head -> + template."""
    dec = V.roundtrip(sd, v_speaker, vq_cfg, prefix="speaker_vq.")[0]
    return motion_to_mesh(sd, dec, "emoca", template), dec


@torch.no_grad()
def speaker_slmft_forward(sd, v_speaker, v_speaker_emoca, v_audio, mask, template, mouth_map, s2s_cfg, vq_cfg, mode="train",
                          speaker_ids=None, temperature=0.0, uniforms=None):
    """SpeakerSLMFT.forward restated line by line (dead values -- the squashed v_speaker and its speaker-VQ codes -- left out).
    temperature=0 selects upstream's greedy branch in generate; temperature>0 needs `uniforms`."""
    B, T, size = v_speaker.shape
    z = S.forward_vq_listener(sd, v_speaker_emoca, mask, vq_cfg)                              # :701-703
    if speaker_ids is None:
        x_l = torch.zeros(B, T, 384)
    else:
        x_l = sd["speaker_embed.weight"][speaker_ids].unsqueeze(1).repeat(1, T, 1)
    ctx = torch.cat([x_l + sd["patch_embed_dec_l"], v_audio], dim=-1)
    if mode == "train":
        l_ce, logits = X.teacher_forced(sd, "decoder_joint.net", z, s2s_cfg.depth, ctx, mask)
        codes = torch.argmax(logits, dim=-1)
    else:
        codes = X.generate(sd, "decoder_joint.net", z[:, 0:1], T - 1, s2s_cfg.depth, ctx, mask, temperature=temperature,
                           uniforms=uniforms, top_k_frac=s2s_cfg.top_k_frac, use_cache=False)      # positional table: prefix recomputed
        l_ce, logits = 0.0, None
    pred_emoca = V.decode_indices(sd, codes, vq_cfg, prefix="speaker_vq.")                    # :652-656
    pred_mesh = motion_to_mesh(sd, pred_emoca, "emoca", template)                             # :657-658, :731
    nv = size // 3
    orig_mouth = v_speaker.reshape(-1, nv, 3)[:, mouth_map, :].reshape(-1, len(mouth_map) * 3)
    pred_mouth = pred_mesh.reshape(-1, nv, 3)[:, mouth_map, :].reshape(-1, len(mouth_map) * 3)
    l_mouth = F.mse_loss(pred_mouth, orig_mouth[1:, :])                                       # :739
    l_emoca = F.mse_loss(pred_emoca, v_speaker_emoca[:, 1:, :])
    d = {"l_ce_s": 0, "l_ce_l": l_ce, "l_cont_s": l_mouth, "l_cont_l": l_emoca, "nce": 0, "c_acc": 0}
    return l_ce + l_emoca, d, pred_emoca, dict(z=z, codes=codes, logits=logits, pred_mesh=pred_mesh)
