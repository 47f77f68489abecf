"""models.get_model (reference: code/models/__init__.py:1-16).  The two stage1_BIWI architectures run on the B200 kernels."""


def get_model(cfg):
    if cfg.arch == "stage1_BIWI":
        from models.stage1_BIWI import VQAutoEncoder as Model
        return Model(args=cfg)
    if cfg.arch == "stage1_BIWI_speaker":
        from models.stage1_BIWI import VQSpeakerAutoEncoder as Model
        return Model(args=cfg)
    raise Exception("architecture not supported by the B200 hot path: {} (stage1_BIWI, stage1_BIWI_speaker; SURVEY.md section 2)".format(cfg.arch))
