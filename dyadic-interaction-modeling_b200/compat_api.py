"""SLMFT.forward(mode='val') on the engines (/root/reference/code/seq2seq_pretrain.py:496-514).

Same results as the reference call; the work the reference computes and throws away is skipped:
the speaker VQ encodes (z_speaker is never read) and the duplicated forward_vq call.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def continuous_loss(pred, target, mask):
    """forward_continuous_loss (seq2seq_pretrain.py:466-478): reporting only, stays in PyTorch."""
    target, mask = target[:, 1:, :], mask[:, 1:]
    B = len(target)
    target = target.reshape(B * target.shape[1], -1)
    pred = pred.reshape(B * pred.shape[1], -1)
    mask = mask.reshape(-1)
    p, t = pred[mask], target[mask]
    return torch.mean(F.pairwise_distance(p[:, 6:], t[:, 6:])) + torch.mean(F.pairwise_distance(p[:, 0:6], t[:, 0:6]))


@torch.no_grad()
def listener_codes(vq_engine, v_listener, mask):
    """forward_vq, listener half (:489-491): per-sample encode of the valid prefix == batched encode with
    lens and batch_index 0 for everyone.  Returns z_l (B,T) int64 with -100 on padding."""
    B, T, _ = v_listener.shape
    lens = mask.sum(dim=1).to(torch.int32)
    zero = torch.zeros(B, dtype=torch.int32, device=v_listener.device)
    idx, _, _ = vq_engine.encode(v_listener, lens=lens, batch_index=zero)
    return idx.masked_fill(~mask, -100)


@torch.no_grad()
def slmft_forward_val(s2s_engine, vq_engine, v_speaker, v_listener, v_audio, mask, temperature=1.0, uniforms=None,
                      batch_index=None, return_codes=False, greedy=None, vq_decode_engine=None):
    """-> (total_loss, dict, pred_cont_seq_l (B,T-1,56)) like SLMFT.forward(..., mode='val').

    The reference samples (temperature 1, top-k 52, torch.multinomial).  Here: `uniforms` (B,T-1) given -> inverse-CDF
    sampling with them; uniforms None and greedy is not False -> argmax decoding (the deterministic parity mode).
    vq_decode_engine: optional second VQEngine over the same weights used for the codes -> frames decode only (the bf16
    mode decodes with bf16 GEMM operands; the ENCODE side always stays fp32-grade so code indices are exact)."""
    B, T, _ = v_speaker.shape
    z_l = listener_codes(vq_engine, v_listener, mask)
    ctx = s2s_engine.context(v_speaker, v_audio, mask)
    if uniforms is None and greedy is not False:
        codes = s2s_engine.generate(ctx, mask, z_l[:, 0], T - 1, temperature=0.0)
    else:
        if uniforms is None:
            uniforms = torch.rand(B, T - 1, device=v_speaker.device)
        codes = s2s_engine.generate(ctx, mask, z_l[:, 0], T - 1, temperature=temperature, uniforms=uniforms)
    pred = (vq_decode_engine or vq_engine).decode(codes=codes, batch_index=batch_index)
    l_cont = continuous_loss(pred, v_listener, mask)
    d = {"l_ce_s": 0, "l_ce_l": 0.0, "l_cont_s": 0, "l_cont_l": l_cont, "nce": 0, "c_acc": 0}
    if return_codes:
        return l_cont, d, pred, codes
    return l_cont, d, pred


_copy_streams = {}


@torch.no_grad()
def slmft_forward_val_host(s2s_engine, vq_engine, host, device, temperature=1.0, uniforms=None, batch_index=None,
                           vq_decode_engine=None, out_host=None):
    """Same forward as slmft_forward_val, fed from (pinned) HOST tensors: host = dict(v_speaker, v_listener, v_audio, mask).

    The host->device copies run on a side stream in the order the stages need them (listener motion, speaker motion + mask,
    then the 768-d audio features, by far the largest), so the listener VQ encode and the speaker encoders overlap the audio
    copy; the decoded frames are copied back into `out_host` (pinned) when given.  Returns (loss, dict, pred[, codes])
    like slmft_forward_val(return_codes=True)."""
    dev = torch.device(device)
    main = torch.cuda.current_stream(dev)
    cs = _copy_streams.setdefault(dev.index, torch.cuda.Stream(dev))
    cs.wait_stream(main)                                  # buffers of the previous call are free
    dv, ev = {}, {}
    with torch.cuda.stream(cs):
        for k in ("v_listener", "mask", "v_speaker", "v_audio"):
            dv[k] = host[k].to(dev, non_blocking=True)
            ev[k] = torch.cuda.Event()
            ev[k].record(cs)
    for k in ("v_listener", "mask"):
        main.wait_event(ev[k])
    B, T, _ = dv["v_listener"].shape
    z_l = listener_codes(vq_engine, dv["v_listener"], dv["mask"])
    main.wait_event(ev["v_speaker"])
    main.wait_event(ev["v_audio"])
    ctx = s2s_engine.context(dv["v_speaker"], dv["v_audio"], dv["mask"])
    if uniforms is None:
        uniforms = torch.rand(B, T - 1, device=dev)
    codes = s2s_engine.generate(ctx, dv["mask"], z_l[:, 0], T - 1, temperature=temperature, uniforms=uniforms)
    pred = (vq_decode_engine or vq_engine).decode(codes=codes, batch_index=batch_index)
    if out_host is not None:
        out_host.copy_(pred, non_blocking=True)
    l_cont = continuous_loss(pred, dv["v_listener"], dv["mask"])
    for t in dv.values():
        t.record_stream(main)
    d = {"l_ce_s": 0, "l_ce_l": 0.0, "l_cont_s": 0, "l_cont_l": l_cont, "nce": 0, "c_acc": 0}
    return l_cont, d, pred, codes


@torch.no_grad()
def slmft_forward_val_samples(s2s_engine, vq_engine, v_speaker, v_listener, v_audio, mask, samples, uniforms=None,
                              temperature=1.0, batch_index=None, vq_decode_engine=None):
    """`samples` stochastic generations per clip in ONE pass (what x_engine_pt.py:255-270 obtains with `samples` model calls):
    the listener VQ encode, the speaker encoders and the cross-attention K/V projection run once per clip, only the decode and
    the VQ decode run per sample.  uniforms (B,samples,T-1) or None.  Returns (pred (B,samples,T-1,56), codes (B,samples,T-1));
    pred[:, j] is bit-identical to slmft_forward_val(..., uniforms=uniforms[:, j])'s prediction."""
    B, T, _ = v_speaker.shape
    z_l = listener_codes(vq_engine, v_listener, mask)
    ctx = s2s_engine.context(v_speaker, v_audio, mask)
    if uniforms is None:
        uniforms = torch.rand(B, samples, T - 1, device=v_speaker.device)
    codes = s2s_engine.generate_samples(ctx, mask, z_l[:, 0], T - 1, samples, uniforms, temperature=temperature)
    if batch_index is None:
        batch_index = torch.arange(B, dtype=torch.int32, device=v_speaker.device)
    # sample j of clip b sits in batch slot b of "its" model call: the VQ decoder's batch-index positional encoding
    bi = batch_index.to(torch.int32).repeat_interleave(samples)
    pred = (vq_decode_engine or vq_engine).decode(codes=codes.reshape(B * samples, T - 1), batch_index=bi)
    return pred.reshape(B, samples, T - 1, -1), codes


def frechet_distance_torch(x, y):
    """Frechet distance between the Gaussians fitted to the rows of x (n,d) and of y (..., n, d), on the tensors' device in fp64:
    |mu1-mu2|^2 + tr(S1) + tr(S2) - 2 tr(sqrtm(S1 S2))  with np.cov's n-1 normalisation (metrics/eval_utils.py:6-44).
    tr(sqrtm(S1 S2)) = sum of the square roots of the eigenvalues of S1^(1/2) S2 S1^(1/2), a symmetric PSD matrix: two `eigh`
    calls instead of scipy's Schur-based sqrtm of a non-symmetric product (identical value, no complex round-off)."""
    x, y = x.double(), y.double()
    n = x.shape[-2]
    mu1, mu2 = x.mean(-2), y.mean(-2)
    xc, yc = x - mu1.unsqueeze(-2), y - mu2.unsqueeze(-2)
    s1 = xc.transpose(-1, -2) @ xc / (n - 1)
    s2 = yc.transpose(-1, -2) @ yc / (y.shape[-2] - 1)
    w, v = torch.linalg.eigh(s1)
    r1 = (v * w.clamp_min(0).sqrt().unsqueeze(-2)) @ v.transpose(-1, -2)          # S1^(1/2)
    ev = torch.linalg.eigvalsh(r1 @ s2 @ r1)
    tr = ev.clamp_min(0).sqrt().sum(-1)
    d = mu1 - mu2
    return (d * d).sum(-1) + torch.diagonal(s1, dim1=-2, dim2=-1).sum(-1) + torch.diagonal(s2, dim1=-2, dim2=-1).sum(-1) - 2 * tr


@torch.no_grad()
def best_of_n(pred, target, lengths):
    """Per clip, the sample whose Frechet distance to the ground truth is smallest (x_engine_pt.py:260-268), selected on the
    device: pred (B,S,L,56), target (B,L,56), lengths[b] = valid frames.  Returns (list of (n_b,56) tensors, chosen (B,), fd (B,S))."""
    B, S = pred.shape[:2]
    keep, chosen, fds = [], [], []
    for b in range(B):
        n = int(lengths[b])
        fd = frechet_distance_torch(target[b, :n], pred[b, :, :n])
        j = int(torch.argmin(fd))                       # first minimum, like the reference's strict `<` scan
        keep.append(pred[b, j, :n])
        chosen.append(j)
        fds.append(fd)
    return keep, torch.tensor(chosen), torch.stack(fds)


def draw_kv_mask(inp_shape, mask_prob, device, generator=None):
    """AutoregressiveWrapper.forward's random `self_attn_kv_mask` (x-transformers 1.30.16; drawn even in eval mode):
    rand = randn(inp.shape); rand[:, 0] = -max; with T = inp.shape[1] + 1 (the length before the shift) the
    num_mask = min(int(T * mask_prob), T - 1) largest draws of each row are masked.  Returns (B, seq) bool, True = key kept."""
    B, seq = inp_shape
    rand = torch.randn(B, seq, device=device, generator=generator)
    rand[:, 0] = -torch.finfo(rand.dtype).max
    # upstream reads `seq` from x BEFORE the shift (x.shape[1] = len(inp) + 1): num_mask = min(int(T * mask_prob), T - 1)
    num_mask = min(int((seq + 1) * mask_prob), seq)
    indices = rand.topk(num_mask, dim=-1).indices
    return ~torch.zeros(B, seq, device=device).scatter(1, indices, 1.0).bool()


@torch.no_grad()
def slmft_forward_train(s2s_engine, vq_engine, v_speaker, v_listener, v_audio, mask, kv_mask=None, mask_prob=0.15,
                        batch_index=None, return_logits=False):
    """SLMFT.forward(mode='train') forward pass (seq2seq_pretrain.py:496-514 with :447-448, :456; what
    x_engine_pt.evaluate_finetune_epoch:217 runs): teacher-forced decoder logits over the listener codes, cross-entropy against
    the shifted codes (ignore_index -100), argmax codes -> VQ decode -> continuous loss.  kv_mask: the self-attention key mask
    upstream draws at random (mask_prob 0.15); None draws one with draw_kv_mask.  Forward only: no autograd graph is built."""
    B, T, _ = v_speaker.shape
    z_l = listener_codes(vq_engine, v_listener, mask)                    # (B,T) with -100 on padding
    ctx = s2s_engine.context(v_speaker, v_audio, mask)
    inp, target = z_l[:, :-1].clone(), z_l[:, 1:]
    inp[inp == -100] = 0                                                 # ignore_index -> pad_value
    if kv_mask is None and mask_prob > 0:
        kv_mask = draw_kv_mask(inp.shape, mask_prob, inp.device)
    logits = s2s_engine.teacher_forced(ctx, mask, inp, kv_mask)
    l_ce = F.cross_entropy(logits.transpose(1, 2), target, ignore_index=-100)
    codes = torch.argmax(logits, dim=-1)
    pred = vq_engine.decode(codes=codes, batch_index=batch_index)
    l_cont = continuous_loss(pred, v_listener, mask)
    d = {"l_ce_s": 0, "l_ce_l": l_ce, "l_cont_s": 0, "l_cont_l": l_cont, "nce": 0, "c_acc": 0}
    out = (l_ce + l_cont, d, pred)
    return out + (logits,) if return_logits else out


# ---- SLM pre-training forward (seq2seq_pretrain.py:72-323) -----------------------------------------------------------------------
def random_masking_unstructured(mask, mask_ratio):
    """SLM.random_masking_unstructured (seq2seq_pretrain.py:171-183): per clip, int(len * ratio) positions of its valid prefix drawn
    with torch.randperm (host RNG, like the reference).  True = masked."""
    N, L = mask.shape
    lens = mask.sum(dim=1).to(torch.int32).tolist()
    final = torch.zeros(N, L, dtype=torch.bool)
    for i, n in enumerate(lens):
        final[i, :n][torch.randperm(n)[: int(n * mask_ratio)]] = True
    return final.to(mask.device)


def contrastive(s_rep, l_rep, mask, temperature=0.05):
    """SLM.forward_contrastive (seq2seq_pretrain.py:270-298), single direction: mean over each clip's valid frames, L2 normalise,
    NCE over the batch.  Returns (nce, c_acc)."""
    m = mask.unsqueeze(-1).to(s_rep.dtype)
    n = m.sum(dim=1)
    s = F.normalize((s_rep * m).sum(dim=1) / n, dim=-1)
    l = F.normalize((l_rep * m).sum(dim=1) / n, dim=-1)
    total = torch.mm(s, l.t()) / temperature
    nce = -torch.mean(torch.diag(F.log_softmax(total, dim=0)))
    c_acc = torch.sum(torch.eq(torch.argmax(F.softmax(total, dim=0), dim=0), torch.arange(total.shape[0], device=s.device))) / total.shape[0]
    return nce, c_acc


def _masked_cont_loss(pred, target, sel):
    """SLM.forward_continuous_loss (seq2seq_pretrain.py:253-268) with the RANDOM-masking selection `sel` (B,T) in place of the
    padding mask (the reference passes mask_speaker / mask_listener here, :311-312)."""
    return continuous_loss(pred, target, sel)


@torch.no_grad()
def slm_forward(s2s_engine, speaker_vq_engine, listener_vq_engine, v_speaker, v_listener, v_audio, mask, patch_s, patch_l,
                patch_dec_s, patch_dec_l, mask_speaker=None, mask_listener=None, mask_ratio=0.15, return_parts=False):
    """SLM.forward (pre-training model, seq2seq_pretrain.py:300-323), forward only:
      forward_vq (:185-200)  both VQ encoders on each clip's valid prefix, speaker codes padded with 0, listener with -100;
      forward_encoder (:203-226)  random masking, encoder_s / encoder_l, encoder_joint on cat([x_s, x_l]) and on each alone,
                                  norm / norm_l / norm_s;
      forward_contrastive (:270-298), forward_decoder (:228-243: two teacher-forced decoder passes over the other side's joint
      representation + audio), forward_vq_decoder (:245-251), forward_continuous_loss on the masked positions.
    mask_speaker / mask_listener: the random masks (None: drawn like the reference does).  Returns (total_loss, dict, None)."""
    B, T, _ = v_speaker.shape
    z_s = listener_codes(speaker_vq_engine, v_speaker, mask)             # same per-clip prefix encode; padding rewritten below
    z_s = torch.where(mask, z_s, torch.zeros_like(z_s))                  # speaker: pad value 0 (:194)
    z_l = listener_codes(listener_vq_engine, v_listener, mask)           # listener: pad value -100 (:197)
    if mask_speaker is None:
        mask_speaker = random_masking_unstructured(mask, mask_ratio)
    if mask_listener is None:
        mask_listener = random_masking_unstructured(mask, mask_ratio)
    vs = v_speaker + patch_s.reshape(1, 1, -1)
    vl = v_listener + patch_l.reshape(1, 1, -1)
    vs[mask_speaker] = 0
    vl[mask_listener] = 0
    e = s2s_engine.encode
    x_s = e("encoder_s", vs, mask)
    x_l = e("encoder_l", vl, mask)
    x_joint = e("encoder_joint", torch.cat([x_s, x_l], dim=1), torch.cat([mask, mask], dim=-1), norm="norm")
    x_l = e("encoder_joint", x_l, mask, norm="norm_l")
    x_s = e("encoder_joint", x_s, mask, norm="norm_s")
    nce, c_acc = contrastive(x_s, x_l, mask)
    xj_s, xj_l = x_joint[:, :T], x_joint[:, T:]
    z_s = torch.where(mask_speaker, z_s, torch.full_like(z_s, -100))     # only the masked positions enter the CE (:307-308)
    z_l = torch.where(mask_listener, z_l, torch.full_like(z_l, -100))
    ctx_s = torch.cat([xj_s + patch_dec_s.reshape(1, 1, -1), v_audio], dim=-1)
    ctx_l = torch.cat([xj_l + patch_dec_l.reshape(1, 1, -1), v_audio], dim=-1)

    def tf(tokens, ctx):
        inp, target = tokens[:, :-1].clone(), tokens[:, 1:]
        inp[inp == -100] = 0
        logits = s2s_engine.teacher_forced(ctx, mask, inp, None)          # AutoregressiveWrapper default mask_prob = 0 in SLM (:165)
        ce = F.cross_entropy(logits.transpose(1, 2), target, ignore_index=-100)
        return ce, logits

    l_ce_s, px_s = tf(z_s, ctx_l)
    l_ce_l, px_l = tf(z_l, ctx_s)
    pred_s = speaker_vq_engine.decode(codes=torch.argmax(px_s, dim=-1))
    pred_l = listener_vq_engine.decode(codes=torch.argmax(px_l, dim=-1))
    l_cont_s = _masked_cont_loss(pred_s, v_speaker, mask_speaker)
    l_cont_l = _masked_cont_loss(pred_l, v_listener, mask_listener)
    total = l_ce_s + l_ce_l + l_cont_s + l_cont_l + nce
    d = {"l_ce_s": l_ce_s, "l_ce_l": l_ce_l, "l_cont_s": l_cont_s, "l_cont_l": l_cont_l, "nce": nce, "c_acc": c_acc}
    if return_parts:
        return total, d, dict(x_s=x_s, x_l=x_l, x_joint=x_joint, px_s=px_s, px_l=px_l, pred_s=pred_s, pred_l=pred_l, z_s=z_s, z_l=z_l)
    return total, d, None


# ---- EmocaConverter / SpeakerSLMFT mesh modules (seq2seq_pretrain.py:516-842) -------------------------------------------------
ACT_LEAKY = 1


@torch.no_grad()
def mesh_to_motion(v, template, map_w, map_b, conv_w, conv_b):
    """vertice_mapping + squasher (seq2seq_pretrain.py:710-713; modules :777-785): (B,T,size) vertices - template -> Linear(size,56)
    + LeakyReLU(0.2) -> Conv1d(k=5, replicate) + LeakyReLU(0.2) + InstanceNorm1d over time -> (B,T,56).
    Kernels: dim_linear_ragged_f32 (split over the 70110-wide K), dim_conv5_leaky_f32, dim_instance_norm_f32."""
    from . import ops
    B, T, size = v.shape
    x = (v - template.reshape(B, 1, size)).reshape(B * T, size).contiguous()
    y = ops.linear_ragged(x, map_w.detach().contiguous(), map_b.detach(), act=ACT_LEAKY, slope=0.2).view(B, T, -1)
    y = ops.conv5_leaky(y, ops.repack_conv_weight(conv_w.detach().contiguous()), conv_b.detach(), 0.2)
    return ops.instance_norm_(y)


@torch.no_grad()
def motion_to_mesh(dec, lstm_params, lin0_w, lin0_b, lin2_w, lin2_b, template=None):
    """vertice_map_reverse_lstm + vertice_map_reverse (seq2seq_pretrain.py:657-658, :823-825; modules :789-807): (B,L,56) ->
    2-layer bidirectional LSTM(384) -> Linear(768,768) + LeakyReLU(0.2) -> Linear(768,size) (+ template).
    Kernels: dim_lstm_layer_f32 per layer, dim_linear_f32, dim_linear_ragged_f32."""
    from . import ops
    B, L, _ = dec.shape
    h = ops.lstm(dec.contiguous(), lstm_params, hidden=lstm_params["weight_hh_l0"].shape[1])
    h = ops.linear(h.view(B * L, -1), lin0_w.detach().contiguous(), lin0_b.detach(), act=ACT_LEAKY, slope=0.2)
    out = ops.linear_ragged(h, lin2_w.detach().contiguous(), lin2_b.detach()).view(B, L, -1)
    return out if template is None else out + template.reshape(B, 1, -1)


@torch.no_grad()
def speaker_slmft_forward(s2s_engine, speaker_vq_engine, listener_vq_engine, mesh, v_speaker, v_speaker_emoca, v_audio, mask,
                          template, patch_dec_l, mouth_map, mode="train", speaker_rows=None, temperature=1.0, uniforms=None,
                          greedy=False):
    """SpeakerSLMFT.forward (seq2seq_pretrain.py:707-757), forward only.  `mesh` = (lstm_params, lin0_w, lin0_b, lin2_w, lin2_b) of
    the 'emoca' head.  The reference also runs vertice_mapping + squasher + a speaker-VQ encode of the result (:710-715) and never
    reads them (z_s is dead: :725 is commented out) -- skipped here like SURVEY F10; mesh_to_motion() is that path on its own.
    Returns (total_loss, dict, pred_cont_seq_s_emoca, parts)."""
    B, T, size = v_speaker.shape
    z = listener_codes(listener_vq_engine, v_speaker_emoca, mask)                       # forward_vq :701-703 (pad -100)
    x_l = torch.zeros(B, T, patch_dec_l.numel(), device=v_audio.device) if speaker_rows is None \
        else speaker_rows.unsqueeze(1).repeat(1, T, 1)                                   # :718-722
    ctx = torch.cat([x_l + patch_dec_l.reshape(1, 1, -1), v_audio], dim=-1).contiguous()  # forward_decoder :640-642
    if mode == "train":
        inp, target = z[:, :-1].clone(), z[:, 1:]
        inp[inp == -100] = 0
        logits = s2s_engine.teacher_forced(ctx, mask, inp, None)
        l_ce = F.cross_entropy(logits.transpose(1, 2), target, ignore_index=-100)
        codes = torch.argmax(logits, dim=-1)                                             # forward_vq_decoder :650-651
    else:
        if greedy or temperature == 0.0:
            codes = s2s_engine.generate(ctx, mask, z[:, 0:1].contiguous(), T - 1, temperature=0.0)
        else:
            if uniforms is None:
                uniforms = torch.rand(B, T - 1, device=ctx.device)
            codes = s2s_engine.generate(ctx, mask, z[:, 0:1].contiguous(), T - 1, temperature=temperature, uniforms=uniforms)
        l_ce, logits = 0.0, None
    pred_emoca = speaker_vq_engine.decode(codes=codes.contiguous())                      # :652-656: SPEAKER codebook rows, speaker decoder
    pred_mesh = motion_to_mesh(pred_emoca, *mesh, template=template)                     # :657-658, :731
    mse = F.mse_loss
    l_cont_mesh = mse(pred_mesh, v_speaker[:, 1:, :])                                    # :734 (overwritten at :749)
    nv = size // 3
    orig_mouth = v_speaker.reshape(-1, nv, 3)[:, mouth_map, :].reshape(-1, len(mouth_map) * 3)
    pred_mouth = pred_mesh.reshape(-1, nv, 3)[:, mouth_map, :].reshape(-1, len(mouth_map) * 3)
    l_mouth = mse(pred_mouth, orig_mouth[1:, :])                                         # :739: drops ONE row of the flattened batch
    l_emoca = mse(pred_emoca, v_speaker_emoca[:, 1:, :])
    total = l_ce + l_emoca
    d = {"l_ce_s": 0, "l_ce_l": l_ce, "l_cont_s": 1.0 * l_mouth, "l_cont_l": l_emoca, "nce": 0, "c_acc": 0}
    parts = dict(z=z, codes=codes, logits=logits, pred_mesh=pred_mesh, l_cont_mesh=l_cont_mesh)
    return total, d, pred_emoca, parts
