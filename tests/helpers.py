import os
import numpy as np
import torch

from oracle import performer_oracle as PO

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def rel_err(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def rms_rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).pow(2).mean().sqrt() / (b.pow(2).mean().sqrt() + 1e-30))


def load_seeded(model, shapes, seed, extra=None):
    """Load the shared seeded synthetic weights into an emo model; returns the CPU state dict."""
    sd = PO.seeded_state(shapes, seed)
    if extra:
        sd.update(extra)
    msd = model.state_dict()
    msd.update({k: v for k, v in sd.items() if k in msd})
    model.load_state_dict(msd)
    return sd


def wsum(sd, shapes):
    return float(sum(sd[k].double().abs().sum() for k in shapes))
