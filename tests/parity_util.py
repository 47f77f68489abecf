"""Shared proof obligations of the decode parity tests: a difference between a generated code sequence and the oracle's is accepted
only as a PROVEN tie, never because of where it occurs."""
import math

import torch

LOGIT_TOL = 5e-4
K = math.ceil(0.1 * 512)


def explain_first_difference(codes, logits, ref_codes, ref_logits, uniforms=None):
    """codes/logits vs the oracle's, one row.  Returns the number of steps that agree; asserts that the logits agree up to (and
    including) the first differing step and that the difference is a provable tie."""
    neq = (codes != ref_codes).nonzero()
    n = len(codes) if len(neq) == 0 else int(neq[0])
    upto = min(n + 1, len(codes))
    d = float((logits[:upto] - ref_logits[:upto]).abs().max())
    assert d < LOGIT_TOL, f"logits differ by {d} within the first {upto} steps"
    if n == len(codes):
        return n
    lr = ref_logits[n].double()
    if uniforms is None:
        top2 = torch.topk(lr, 2).values
        assert float(top2[0] - top2[1]) < 2 * LOGIT_TOL, f"greedy decode diverged at step {n} with oracle margin {float(top2[0] - top2[1])}"
    else:
        kth = torch.topk(lr, K).values[-1]
        p = torch.where(lr >= kth, (lr - lr.max()).exp(), torch.zeros_like(lr))
        cdf = (p / p.sum()).cumsum(0)
        gap = float((cdf - float(uniforms[n])).abs().min())
        # a logit error e moves a CDF boundary by at most ~e (d softmax <= p(1-p) e): 2 * LOGIT_TOL covers it
        assert gap < 2 * LOGIT_TOL, f"sampled decode diverged at step {n}: draw {float(uniforms[n])} is {gap} away from the nearest CDF boundary"
    return n
