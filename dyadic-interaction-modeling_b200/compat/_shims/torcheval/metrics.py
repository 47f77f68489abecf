class Perplexity:
    def __init__(self, *a, **k):
        raise NotImplementedError("torcheval is not available offline")
