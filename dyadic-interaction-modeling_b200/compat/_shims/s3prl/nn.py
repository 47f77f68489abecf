class S3PRLUpstream:
    def __init__(self, *a, **k):
        raise NotImplementedError("s3prl is not available offline; the eval path consumes pre-extracted 768-d HuBERT features")
