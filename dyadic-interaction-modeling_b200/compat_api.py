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
