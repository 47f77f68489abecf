"""base.BaseModel (reference: code/base/base_model.py:5-30)."""
import torch.nn as nn


class BaseModel(nn.Module):
    def forward(self, *x):
        raise NotImplementedError

    def summary(self, logger, writer):
        n = sum(p.numel() for p in self.parameters() if p.requires_grad) / 1e6
        logger.info(self)
        logger.info("===>Trainable parameters: %.3f M" % n)
        if writer is not None:
            writer.add_text("Model Summary", "Trainable parameters: %.3f M" % n)
