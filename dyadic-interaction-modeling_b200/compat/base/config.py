"""YAML -> flat attribute dict, CLI overrides (reference: code/base/config.py:10-159).

`load_cfg_from_cfg_file` flattens the yaml's sections into one namespace exactly like the reference (:60-73);
`merge_cfg_from_list` applies `KEY VALUE` pairs with literal decoding and type coercion (:76-159)."""
import copy
import os
from ast import literal_eval

import yaml


class CfgNode(dict):
    def __init__(self, init_dict=None, key_list=None, new_allowed=False):
        init_dict = {} if init_dict is None else init_dict
        super().__init__({k: (CfgNode(v) if type(v) is dict else v) for k, v in init_dict.items()})

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name)

    def __setattr__(self, name, value):
        self[name] = value

    def __repr__(self):
        return "CfgNode(%s)" % dict.__repr__(self)


def load_cfg_from_cfg_file(file):
    assert os.path.isfile(file) and file.endswith(".yaml"), "{} is not a yaml file".format(file)
    with open(file, "r") as f:
        tree = yaml.safe_load(f)
    flat = {}
    for section in tree:
        for k, v in tree[section].items():
            flat[k] = v
    return CfgNode(flat)


def _coerce(new, old, key):
    if old is None or type(new) is type(old):
        return new
    for src, dst in ((list, tuple), (tuple, list), (int, float)):
        if type(new) is src and type(old) is dst:
            return dst(new)
    raise ValueError("type mismatch for config key {}: {} vs {}".format(key, type(old), type(new)))


def merge_cfg_from_list(cfg, cfg_list):
    out = copy.deepcopy(cfg)
    assert len(cfg_list) % 2 == 0
    for full_key, v in zip(cfg_list[0::2], cfg_list[1::2]):
        key = full_key.split(".")[-1]
        assert key in cfg, "Non-existent key: {}".format(full_key)
        if isinstance(v, str):
            try:
                v = literal_eval(v)
            except (ValueError, SyntaxError):
                pass
        out[key] = _coerce(v, cfg[key], key)
    return out
