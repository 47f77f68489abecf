"""Argument parsing / logging helpers used by train_vq.py (reference: code/base/utilities.py:11-66)."""
import argparse
import logging

from . import config


def get_parser():
    p = argparse.ArgumentParser(description=" ")
    p.add_argument("--config", type=str, default="config.yaml", help="path to config file")
    p.add_argument("opts", help=" ", default=None, nargs=argparse.REMAINDER)
    args = p.parse_args()
    cfg = config.load_cfg_from_cfg_file(args.config)
    if args.opts:
        cfg = config.merge_cfg_from_list(cfg, args.opts)
    return cfg


def get_logger():
    logger = logging.getLogger("main-logger")
    logger.setLevel(logging.INFO)
    if not logger.handlers:
        h = logging.StreamHandler()
        h.setFormatter(logging.Formatter("[%(asctime)s %(levelname)s %(filename)s line %(lineno)d %(process)d]=>%(message)s"))
        logger.addHandler(h)
    return logger


class AverageMeter:
    def __init__(self):
        self.reset()

    def reset(self):
        self.val = self.avg = self.sum = self.count = 0

    def update(self, val, n=1):
        self.val = val
        self.sum += val * n
        self.count += n
        self.avg = self.sum / self.count


def main_process(args):
    return not getattr(args, "multiprocessing_distributed", False) or \
        (args.multiprocessing_distributed and args.rank % args.ngpus_per_node == 0)
