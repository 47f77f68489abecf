"""BASELINE.json configs[4]: end-to-end listener-generation eval on synthetic LM-Listener-shape inputs (chunks of T frames, audio =
zeros like dataset/data_loader.py:242, equal lengths), sharded data-parallel over the GPUs of one box, outputs diffed against the
CPU oracle (restated reference) within 1e-4.

    python scripts/lm_listener_eval.py [--clips 32] [--frames 1024] [--check 2]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/lm_listener_eval.py ...

Every rank decodes its contiguous shard (fp32-grade parity mode, greedy) with the GLOBAL batch positions, the generated codes are
all-gathered (the one collective), rank 0 (a) recomputes `check` clips of OTHER shards on its own GPU and asserts the gathered codes
are identical (sharded == unsharded), (b) runs the CPU oracle on `check` clips spread over the shards and asserts the decoded
FLAME coefficients agree within 1e-4 (codes token for token; a first difference must be a proven oracle near-tie), (c) writes the
predictions pickle test_l2l.py reads and evaluates dim_b200.metrics on it.  Prints one JSON line."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch

import dim_b200  # noqa: F401
from dim_b200 import dist as D
from dim_b200.compat_api import slmft_forward_val
from dim_b200.engine import PREC_FP32_TC, Handle, SLMFTEngine, VQEngine
from dim_b200.schema import S2SConfig, VQConfig


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--clips", type=int, default=32)           # dataset/l2l.py:103 batch size
    ap.add_argument("--frames", type=int, default=1024)        # chunk cap, data_loader.py:215-227
    ap.add_argument("--check", type=int, default=2)
    args = ap.parse_args()
    rank, world, local = D.init_from_env()
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    S2S, VQ = S2SConfig(), VQConfig()
    sd = dim_b200.synth.make_slmft_state_dict(131)
    h = Handle(local)
    h.register(sd)
    s2s, vq = SLMFTEngine(h, S2S, precision=PREC_FP32_TC), VQEngine(h, VQ, prefix="listener_vq.", precision=PREC_FP32_TC)
    B, T = args.clips, args.frames
    clips = dim_b200.synth.make_clips(B, T, seed=4242)
    clips["v_audio"] = torch.zeros_like(clips["v_audio"])                       # LmListenerDataset: audio_feats = zeros
    shard = D.shard_batch({k: clips[k] for k in ("v_speaker", "v_listener", "v_audio", "mask")}, rank, world)
    g = {k: v.to(dev) for k, v in shard.items()}

    def run(x):
        return slmft_forward_val(s2s, vq, x["v_speaker"], x["v_listener"], x["v_audio"], x["mask"], batch_index=x["batch_index"],
                                 return_codes=True)

    run(g)                                                                       # warm-up
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    t0 = time.perf_counter()
    _, _, pred, codes = run(g)
    all_codes = D.all_gather_codes(codes, B)
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    dt = time.perf_counter() - t0
    if rank != 0:                                   # nothing collective follows: the oracle check below is rank 0's alone (minutes of CPU)
        torch.distributed.destroy_process_group()
        return
    out = {"config": {"workload": "lm_listener", "clips": B, "frames_per_clip": T, "n_gpus": world, "precision": "fp32_tc (parity mode), greedy"},
           "frames_per_s_wall": B * (T - 1) / dt, "ms": dt * 1e3}
    # (a) sharded == unsharded: clips of the LAST shard recomputed here (16 rows: same kernel family as a full shard)
    s, e = D.shard_range(B, world - 1, world)
    n = min(16, e - s)
    sub = {k: clips[k][s:s + n].to(dev) for k in ("v_speaker", "v_listener", "v_audio", "mask")}
    sub["batch_index"] = torch.arange(s, s + n, dtype=torch.int32, device=dev)
    _, _, pred_f, codes_f = run(sub)
    out["sharded_equals_unsharded"] = bool(torch.equal(codes_f, all_codes[s:s + n]))
    # (b) CPU oracle on clips spread over the shards
    from oracle import slmft as OS
    from oracle import xt as OX
    from parity_util import explain_first_difference
    torch.set_num_threads(os.cpu_count() or 1)
    picks = sorted(set([0, B - 1] + [int(i * B / max(1, args.check)) for i in range(args.check)]))[: max(1, args.check)]
    worst, same = 0.0, 0
    t1 = time.perf_counter()
    for b in picks:
        one = {k: clips[k][b:b + 1] for k in ("v_speaker", "v_listener", "v_audio", "mask")}
        bi = torch.tensor([b])
        _, _, ref_pred, inter = OS.forward_val(sd, one["v_speaker"], one["v_listener"], one["v_audio"], one["mask"], S2S, VQ,
                                               batch_index=bi, return_intermediates=True)
        got = all_codes[b].cpu()
        if torch.equal(got, inter["codes"][0]):
            same += 1
            # this clip's frames: decoded by whichever rank owns it; recompute the decode here from the gathered codes (same kernel)
            pr = vq.decode(codes=all_codes[b:b + 1], batch_index=torch.tensor([b], dtype=torch.int32, device=dev))
            worst = max(worst, float((pr.cpu() - ref_pred).abs().max()))
        else:
            x = {k: v.to(dev) for k, v in one.items()}
            ctx = s2s.context(x["v_speaker"], x["v_audio"], x["mask"])
            c2, logits = s2s.generate(ctx, x["mask"], inter["z_l"][:, 0].to(dev), T - 1, return_logits=True)
            _, ref_logits = OX.generate(sd, "decoder_joint.net", inter["z_l"][:, 0:1], T - 1, S2S.depth, inter["ctx"], one["mask"], return_logits=True)
            explain_first_difference(c2.cpu()[0], logits[0].cpu(), inter["codes"][0], ref_logits[0])       # raises unless a near-tie
    out["oracle_check"] = {"clips": picks, "code_sequences_identical": same, "max_abs_coeff_diff": worst, "tolerance": 1e-4,
                           "oracle_seconds": time.perf_counter() - t1, "oracle": "oracle/slmft.py forward_val (restated reference), CPU fp32"}
    assert worst < 1e-4, worst
    # (c) the artefact test_l2l.py consumes + the metric suite on the device
    from dim_b200 import l2l_artifacts as A
    import tempfile
    with tempfile.TemporaryDirectory() as tmp:
        ids = [f"chunk_{i:03d}" for i in range(s, s + n)]
        A.write_l2l_predictions(os.path.join(tmp, "l2l_vico_predictions.pkl"), ids, [p.cpu().numpy() for p in pred_f])
        out["predictions_pickle_bytes"] = os.path.getsize(os.path.join(tmp, "l2l_vico_predictions.pkl"))
    print(json.dumps(out), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
