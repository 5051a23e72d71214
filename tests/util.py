"""Shared helpers for the tests: golden fixture loading."""
import glob
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_npz(name):
    with np.load(os.path.join(GOLDEN, name)) as z:
        return {k: z[k] for k in z.files}


def golden_names(prefix):
    """e.g. prefix='lenia' -> ['abominable_mongolic', ...] (files <prefix>_<name>.npz)."""
    out = []
    for p in sorted(glob.glob(os.path.join(GOLDEN, f"{prefix}_*.npz"))):
        name = os.path.basename(p)[len(prefix) + 1:-4]
        if name.startswith("kmult"):
            continue  # synthetic k_mult fixture has its own tests
        out.append(name)
    return out


def t(a, dtype=None):
    x = torch.from_numpy(np.ascontiguousarray(a))
    return x.to(dtype) if dtype is not None else x


def family_case(tag, name):
    """Returns (fixture dict, params dict) for a Lenia-family golden case."""
    fx = load_npz(f"{tag}_{name}.npz")
    if name == "realstate":
        return fx, fx
    return fx, load_npz(f"params_{tag}_{name}.npz")
