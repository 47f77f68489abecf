"""models.get_model (reference: code/models/__init__.py:1-16).  Only the DIM architecture is on the hot path."""


def get_model(cfg):
    if cfg.arch == "stage1_BIWI":
        from models.stage1_BIWI import VQAutoEncoder as Model
        return Model(args=cfg)
    raise Exception("architecture not supported by the B200 hot path: {} (only stage1_BIWI; SURVEY.md section 2)".format(cfg.arch))
