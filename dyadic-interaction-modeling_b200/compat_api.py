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
