"""Helpers to read tests/golden/*.npz (written by tests/golden/make_golden.py)."""
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
DEC_NAMES = ("joint", "joint_auto", "att", "ctc", "joint_pen")


def load(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    cfg = {k: int(v) for k, v in zip(z["cfg_keys"].tolist(), z["cfg_vals"].tolist())}
    if "encoder_type" in z.files:
        cfg["encoder"] = str(z["encoder_type"])
    weights = {k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("w:")}
    return z, cfg, weights


def decode_params(z, dn):
    beam, cw, mlr, minr, pen, nl = z[f"dec:{dn}:params"].tolist()
    return dict(beam_size=int(beam), ctc_weight=cw, maxlenratio=mlr, minlenratio=minr, penalty=pen,
                normalize_length=bool(nl))


def decode_results(z, dn):
    out = []
    for j in range(int(z[f"dec:{dn}:n"])):
        out.append((z[f"dec:{dn}:{j}:yseq"].tolist(), float(z[f"dec:{dn}:{j}:score"]), z[f"dec:{dn}:{j}:scores"]))
    return out
