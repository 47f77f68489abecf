"""dim_b200 -- B200-native DIM inference hot path (VQ-VAE encode/quantize/decode + SLMFT generate).

Import as `dim_b200` (see ../dim_b200.py).  Sub-modules:
  schema   state_dict key/shape tables of the reference models
  synth    deterministic synthetic checkpoints / clips
  _lib     loader (and in-tree nvcc builder) of csrc/ -> libdimb200.so, the C-ABI in include/dimb200.h
  ops      one Python function per C-ABI kernel entry point (raw device pointers, current stream)
  engine   VQ-VAE / SLMFT engines: weight packing + model-level C-ABI calls
  dist     clip sharding over ranks + the single all-gather of generated codes
  compat/  a directory to put on sys.path in place of the reference's code/: same module and class names
"""
from . import schema, synth  # noqa: F401

__all__ = ["schema", "synth"]
