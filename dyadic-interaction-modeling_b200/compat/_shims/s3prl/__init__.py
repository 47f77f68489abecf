"""Stand-in package for `s3prl` (HuBERT feature extraction in the reference's LM-Listener preprocessing)."""
