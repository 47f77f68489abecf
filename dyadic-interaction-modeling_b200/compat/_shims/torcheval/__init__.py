"""Stand-in package for `torcheval` (the reference's x_engine_pt.py imports Perplexity and never uses it on the eval path)."""
