"""mymetrics.print_metrics / print_metrics_full (reference: code/mymetrics.py:7-122), the calls test_s2s_pretrain.py:74-75 makes
after evaluate_test_epoch.  Same printed lines in the same order, same return value (fid_pose, fid_exp); the arithmetic runs as
device tensor ops (dim_b200.metrics) instead of numpy / scipy loops -- note that the reference's own
`calculate_frechet_distance` no longer runs on current scipy (`sqrtm(..., disp=False)`, eval_utils.py:28).
print_biwi_metrics (mymetrics.py:122-182: lip vertex error and upper-face dynamics deviation of the SpeakerSLMFT / BIWI path) likewise;
the two data paths it hard-codes are keyword arguments with the reference's values as defaults."""
import os
import pickle

import numpy as np
import torch

from dim_b200 import metrics as M
from dim_b200.compat_api import frechet_distance_torch


def _dev(seq):
    dev = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
    return [torch.as_tensor(np.asarray(a)).to(dev) for a in seq]


def print_metrics(y_true, y_pred, x):
    gt, pred, xs = _dev(y_true), _dev(y_pred), _dev(x)
    m = M.metrics_suite(gt, pred, xs)
    print('fid_pose: ', m["fid_pose"])
    print('fid_exp: ', m["fid_exp"])
    print('pfid_pose: ', m["pfid_pose"])
    print('pfid_exp: ', m["pfid_exp"])
    print('mse_pose: ', m["mse_pose"])
    print('mse_exp: ', m["mse_exp"])
    print('sid_pose: ', *m["sid_pose"])
    print('sid_exp: ', *m["sid_exp"])
    print('var_pose: ', *m["var_pose"])
    print('var_exp: ', *m["var_exp"])
    print('rpcc pose: ', m["rpcc_pose"])
    print('rpcc exp: ', m["rpcc_exp"])
    print('sts pose: ', m["sts_pose"])
    print('sts exp: ', m["sts_exp"])
    return m["fid_pose"], m["fid_exp"]


def print_metrics_full(y_true, y_pred, x):
    gt, pred, xs = _dev(y_true), _dev(y_pred), _dev(x)
    fid = torch.stack([frechet_distance_torch(g, p) for g, p in zip(gt, pred)]).mean()
    pfid = torch.stack([frechet_distance_torch(torch.cat([s, g], -1), torch.cat([s, p], -1)) for g, p, s in zip(gt, pred, xs)]).mean()
    mse = torch.stack([((g.double() - p.double()) ** 2).mean() for g, p in zip(gt, pred)]).mean()
    G, P = torch.cat(gt).double().reshape(-1, 56), torch.cat(pred).double().reshape(-1, 56)
    print('fid: ', float(fid))
    print('pfid: ', float(pfid))
    print('mse: ', float(mse))
    print('var: ', float(G.var(unbiased=False)), float(P.var(unbiased=False)))


def print_biwi_metrics(y_true, y_pred, file_names, templates_path="../data/BIWI_data/templates.pkl",
                       region_path="../data/CodeTalker/BIWI/regions/"):
    with open(templates_path, 'rb') as fin:
        templates = pickle.load(fin, encoding='latin1')
    read_map = lambda name: [int(i) for i in open(os.path.join(region_path, name)).read().split(", ")]
    mouth_map, upper_map = read_map("lve.txt"), read_map("fdd.txt")
    gt, pred = _dev(y_true), _dev(y_pred)
    tpl = _dev([templates[f.split("_")[0]] for f in file_names])
    lve, fdd = M.biwi_metrics(gt, pred, tpl, mouth_map, upper_map)
    print('Lip Vertex Error: {:.4e}'.format(lve))
    print('FDD: {:.4e}'.format(fdd))
    return lve, fdd
