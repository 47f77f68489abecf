"""Names kept for import compatibility (reference: code/models/lib/base_models.py).  The layer arithmetic
(Norm/Residual/Attention/MLP/Transformer, PositionalEncoding with the batch-index quirk) is implemented in
csrc/ (engine.cu: vq_trunk); these module trees only hold parameters -- see dim_b200.paramtree."""
from dim_b200.paramtree import ParamTree  # noqa: F401
from dim_b200.synth import sinusoid_table  # noqa: F401
