"""Host-side (cold path) parameter handling of the Lenia family.

Mirrors what the hot path consumes from the reference's parameter classes:
  * ``LeniaParams`` loading + ``_sanitize``   pyca/automata/models/lenia/leniaparams.py:28-87,
    pyca/automata/utils/batch_params.py:14-45 (``torch.load`` of a plain dict of batched tensors)
  * kernel construction ``compute_kernel`` / ``kernel_slice``   lenia/lenia.py:289-362
    (ring kernels on the full k x k square; ``k_arbi`` Fourier kernels with the smooth disc mask,
    lenia/funcgen.py:79-129, lenia/utils.py:3-17)
All of this runs once per ``update_params`` in plain torch; its output (``K``, ``mu``, ``sigma``,
``weights``) is what the CUDA step takes.
"""
from __future__ import annotations

import math

import torch

TENSOR_KEYS = ("mu", "sigma", "beta", "mu_k", "sigma_k", "weights")


class LeniaParams:
    """Dict of batched parameter tensors, same keys as the reference's files."""

    def __init__(self, param_dict=None, from_file=None, channels=3, device="cpu"):
        if from_file is not None:
            param_dict = torch.load(from_file, map_location="cpu", weights_only=True)
        if param_dict is None:
            raise ValueError("LeniaParams needs param_dict or from_file")
        if isinstance(param_dict, LeniaParams):
            param_dict = param_dict.param_dict
        self.param_dict = dict(param_dict)
        assert "weights" in self.param_dict, 'LeniaParams need "weights" tensor'
        assert "k_size" in self.param_dict, 'LeniaParams need "k_size" value'
        self.param_dict.setdefault("k_mult", 1)            # legacy files (leniaparams.py:61-64)
        self.param_dict.setdefault("num_channels", channels)
        self.batch_size = self.param_dict["weights"].shape[0]
        self._sanitize()
        self.to(device)

    def _sanitize(self):
        """leniaparams.py:68-87."""
        p = self.param_dict
        if "mu" in p:
            p["mu"] = torch.clamp(p["mu"], -2, 2)
        if "sigma" in p:
            p["sigma"] = torch.clamp(p["sigma"], 1e-4, 1.0)
        if "beta" in p:
            p["beta"] = torch.clamp(p["beta"], 0, None)
        if "mu_k" in p:
            p["mu_k"] = torch.clamp(p["mu_k"], 0.0, 2.0)
        if "sigma_k" in p:
            p["sigma_k"] = torch.clamp(p["sigma_k"], 1e-4, 1.0)
        w = torch.clamp(p["weights"], 0, None)
        n = w.sum(dim=1, keepdim=True)
        p["weights"] = torch.where(n > 1.0e-6, w / n, 0)

    def to(self, device):
        self.device = device
        for k, v in self.param_dict.items():
            if isinstance(v, torch.Tensor):
                self.param_dict[k] = v.to(device)
        return self

    def expand(self, batch_size):
        """Parameters with batch size 1 repeated to `batch_size` worlds (the reference's BatchParams.expand,
        automata/utils/batch_params.py:217-236): how PyCA steps several worlds with one parameter set, `size=(B,H,W)`."""
        assert self.batch_size == 1, f"Batch size must be 1 to expand, here is {self.batch_size}"
        new = {}
        for key, v in self.param_dict.items():
            new[key] = v.repeat(batch_size, *([1] * (v.dim() - 1))) if isinstance(v, torch.Tensor) else v
        channels = self.param_dict["weights"].shape[-1]
        return type(self)(param_dict=new, channels=channels, device=getattr(self, "device", "cpu"))

    def __getitem__(self, k):
        return self.param_dict[k]

    def __setitem__(self, k, v):
        self.param_dict[k] = v

    def __contains__(self, k):
        return k in self.param_dict

    def get(self, k, default=None):
        return self.param_dict.get(k, default)

    @property
    def k_size(self):
        return self.param_dict["k_size"]

    @property
    def k_mult(self):
        return self.param_dict["k_mult"]

    @staticmethod
    def default_gen(batch_size, num_channels=3, k_size=31, k_mult=1, device="cpu") -> "LeniaParams":
        """Random ring-kernel parameters with the reference's prior (leniaparams.py:309-353 + `_universal_params`
        :619-646), drawn from torch's global generator in the reference's order -- what `Lenia(params=None)` uses
        (lenia.py:68-70).  Same seed, same parameters."""
        CK = num_channels * k_mult
        if k_mult == 1:
            weights = torch.rand(batch_size, num_channels, num_channels, device=device) * (
                1 - 0.8 * torch.diag(torch.ones(num_channels, device=device)))
        else:
            weights = torch.rand(batch_size, CK, num_channels, device=device)
        mu = 0.7 * torch.rand((batch_size, CK, num_channels), device=device)
        sigma = mu / (math.sqrt(2 * math.log(2))) * 0.8 * torch.rand((batch_size, CK, num_channels), device=device) + 1e-4
        d = {
            "mu": mu,
            "sigma": sigma,
            "beta": torch.rand((batch_size, CK, num_channels, 3), device=device),
            "mu_k": 0.5 + 0.2 * torch.randn((batch_size, CK, num_channels, 3), device=device),
            "sigma_k": 0.05 * (1 + torch.clamp(0.3 * torch.randn((batch_size, CK, num_channels, 3), device=device), min=-0.9) + 1e-4),
            "k_size": k_size, "weights": weights, "k_mult": k_mult, "num_channels": num_channels,
        }
        # the reference builds LeniaParams twice on this path (default_gen returns one, the constructor wraps its dict in
        # another, leniaparams.py:47-52), so the weights are normalised twice: kept, it is a 1-ulp difference
        once = LeniaParams(param_dict=d, channels=num_channels, device=device)
        return LeniaParams(param_dict=once.param_dict, channels=num_channels, device=device)

    def slice_channels(self, c: int) -> "LeniaParams":
        """First `c` channels of every tensor (``v[:, :c*k_mult, :c]``); weights re-normalised by _sanitize."""
        km = self.k_mult
        d = {}
        for k, v in self.param_dict.items():
            d[k] = v[:, : c * km, :c].clone() if isinstance(v, torch.Tensor) and v.dim() >= 3 else v
        d["num_channels"] = c
        return LeniaParams(param_dict=d, channels=c, device=self.device)


def arbitrary_function(x, coeffs, harmonics, lo, hi, rescale=None, clip_min=None):
    """Batched Fourier-series function (lenia/funcgen.py:79-129).  x: (N,*), coeffs (N,2h) as (cos,sin)
    pairs, harmonics (N,h); defined on the period [lo,hi]."""
    N = coeffs.shape[0]
    cf = coeffs.reshape(N, -1, 2)
    shape = x.shape
    xf = x.reshape(N, 1, -1)
    arg = 2 * math.pi / (hi - lo) * harmonics[..., None] * (xf - lo)
    vals = torch.sum(cf[:, :, 0:1] * torch.cos(arg) + cf[:, :, 1:2] * torch.sin(arg), dim=1)
    if rescale is not None:
        mn = vals.min(dim=-1, keepdim=True)[0]
        mx = vals.max(dim=-1, keepdim=True)[0]
        vals = (vals - mn) / (mx - mn + 1e-8)
        vals = vals * (rescale[1] - rescale[0]) + rescale[0]
    if clip_min is not None:
        vals = torch.where(vals < clip_min, torch.full_like(vals, clip_min), vals)
    return vals.reshape(shape)


def smooth_circular_mask(K, radius):
    """lenia/utils.py:3-17."""
    H, W = K.shape[-2], K.shape[-1]
    y = torch.linspace(0, H - 1, H, device=K.device).view(-1, 1)
    x = torch.linspace(0, W - 1, W, device=K.device).view(1, -1)
    dist = ((y - (H - 1) / 2) ** 2 + (x - (W - 1) / 2) ** 2).sqrt()
    return K * torch.clamp(1 - (dist - radius) / 0.5, 0, 1)


def compute_kernel(params: LeniaParams, C: int) -> torch.Tensor:
    """(B, C*k_mult, C, k, k) normalised kernels (lenia.py:314-362)."""
    p = params.param_dict
    k = int(p["k_size"])
    if k % 2 == 0:
        k += 1
    km = int(p.get("k_mult", 1))
    dev = p["weights"].device
    ax = torch.linspace(-1, 1, k).to(dev)
    gx, gy = torch.meshgrid(ax, ax, indexing="xy")
    r = torch.sqrt(gx**2 + gy**2)
    B = p["weights"].shape[0]
    if p.get("k_arbi", False):
        n = B * C * km * C
        K = arbitrary_function(r[None].expand(n, -1, -1), p["k_coeffs"].reshape(n, -1), p["k_harmonics"].reshape(n, -1),
                               0.0, 1.0, rescale=p.get("k_rescale"), clip_min=0.0)
        K = smooth_circular_mask(K.reshape(B, C * km, C, k, k), k // 2)
    else:
        beta, mu_k, sigma_k = p["beta"], p["mu_k"], p["sigma_k"]
        rr = r[None, None, None, None].expand(*beta.shape, -1, -1)
        rings = torch.exp(-(((rr - mu_k[..., None, None]) / sigma_k[..., None, None]) ** 2) / 2)
        K = torch.sum(beta[..., None, None] * rings, dim=3)
    tot = torch.sum(K, dim=(-1, -2), keepdim=True)
    tot = torch.where(tot < 1e-6, 1, tot)
    return K / tot
