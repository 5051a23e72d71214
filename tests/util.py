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


GARBI_CASES = ["underweight_skagway", "inaesthetic_maupassant", "unsuspected_lanthanotidae", "lenia_synth"]


def garbi_case(name):
    """g_arbi fixture -> (fixture, oracle `garbi` tuple, param dict for pyca_b200.LeniaParams)."""
    fx = load_npz(f"garbi_{name}.npz")
    rs = None if np.isnan(fx["g_rescale"]).any() else (float(fx["g_rescale"][0]), float(fx["g_rescale"][1]))
    cl = None if np.isnan(fx["g_clip"]) else float(fx["g_clip"])
    garbi = (t(fx["g_coeffs"]), t(fx["g_harmonics"]), rs, cl)
    return fx, garbi


INT_KEYS = ("k_size", "k_mult", "num_channels")
BOOL_KEYS = ("g_arbi", "k_arbi")
TUPLE_KEYS = ("g_rescale", "k_rescale")


def raw_param_dict(fx):
    """The parameter-file dict stored under p_* keys of a fixture, back in the reference's types."""
    d = {}
    for key, v in fx.items():
        if not key.startswith("p_"):
            continue
        k = key[2:]
        if k in INT_KEYS:
            d[k] = int(v)
        elif k in BOOL_KEYS:
            d[k] = bool(v)
        elif k in TUPLE_KEYS:
            d[k] = None if np.isnan(v).any() else (float(v[0]), float(v[1]))
        elif v.ndim == 0:
            d[k] = None if np.isnan(v) else float(v)
        else:
            d[k] = t(v.astype(np.float32))
    return d


def food_respawn_like_reference(H, W):
    """random_food_chan(food_amount=200, num_spots=5, food_size=7, add_to_exisitng=True) of macelenia.py:301-345 as a
    callback for oracle.mace_food_step: same `random.randint` calls in the same order."""
    import random

    def respawn(food, batches, food_amount=200, num_spots=5, food_size=7):
        dens = food_amount / (num_spots * food_size * food_size)
        places = [[random.randint(food_size, H - food_size), random.randint(food_size, W - food_size)] for _ in range(num_spots)]
        food = food.clone()
        for y, x in places:
            food[batches, :, y - food_size // 2: y + food_size // 2 + 1, x - food_size // 2: x + food_size // 2 + 1] = dens
        return food

    return respawn
