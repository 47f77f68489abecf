"""CPU oracle for the DIM inference hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may import this
package, and only as the checker.  The product (dyadic-interaction-modeling_b200/) never imports it and has
no CPU fallback.

Contents
  vqvae.py  op-exact restatement of the reference VQ-VAE (models/stage1_BIWI.py, models/lib/*).
            PARITY PINNED: tests/golden/vq_*.pt were produced by importing the real reference from
            /root/reference/code (tests/golden/make_golden.py) and the restatement is bit-identical to
            it on CPU in the container that generated them.
  xt.py     restatement of the x-transformers==1.30.16 subset DIM uses (code/requirements.txt:99).
            The package is NOT in /root/reference and not installable offline.
            PARITY UNPINNED for this half: no upstream source, test or golden vector is reachable; the
            restatement follows SURVEY.md Appendix A and the reference's call sites
            (seq2seq_pretrain.py:369-452).  Golden files for it are self-generated regression pins.
  slmft.py  SLMFT.forward(mode='val') composed from the two (seq2seq_pretrain.py:431-514).

Everything is plain PyTorch on CPU in fp32, written as functions over a state_dict with the reference's
key names.
"""
