"""Developer tool behind profiles/r02_conv_paths.md: step time of the three Lenia convolution paths per kernel size.
    python tools/conv_paths.py > gpurun_out/conv_paths.txt"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pyca_b200  # noqa: E402
from tests.util import load_npz, t  # noqa: E402


def params(C, ks):
    p = load_npz("params_lenia_abominable_mongolic.npz")
    d = {k: t(p[k]) for k in ["mu", "sigma", "beta", "mu_k", "sigma_k", "weights"]}
    if C != 3:
        d = {k: v[:, :C, :C].clone() for k, v in d.items()}
    d["k_size"] = ks
    return pyca_b200.LeniaParams(param_dict=d, channels=C)


def timeit(C, shape, ks, path):
    per_world = shape[0] * shape[1] * C * 16
    nw = 1 if per_world > (126 << 20) else -(-(256 << 20) // per_world)
    g = torch.Generator().manual_seed(1)
    try:
        autos = [pyca_b200.Lenia(shape, num_channels=C, params=params(C, ks), state_init=torch.rand(1, C, *shape, generator=g),
                                 device="cuda", conv_path=path) for _ in range(nw)]
    except pyca_b200.B200CAError:
        return None
    for a in autos:
        a.step()
        a.step()
    torch.cuda.synchronize()
    n = max(48 if shape[0] >= 4096 else 240, 2 * nw) // nw * nw
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(3):
        e0.record()
        for i in range(n):
            autos[i % nw].step()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / n)
    return best * 1e3


if __name__ == "__main__":
    print("| world | C | k_size | direct us | marching FFT us | row-FFT hybrid us | fastest |")
    print("|---|---|---|---|---|---|---|")
    for C, shape in [(1, (1024, 1024)), (1, (4096, 4096)), (3, (1024, 1024)), (3, (4096, 4096))]:
        for ks in (5, 7, 11, 15, 19, 23, 31):
            r = {p: timeit(C, shape, ks, p) for p in ("direct", "fft", "fft_rows")}
            best = min((v, k) for k, v in r.items() if v is not None)[1]
            f = lambda v: "n/a" if v is None else f"{v:.1f}"
            print(f"| {shape[0]}x{shape[1]} | {C} | {ks} | {f(r['direct'])} | {f(r['fft'])} | {f(r['fft_rows'])} | "
                  f"{ {'direct': 'direct', 'fft': 'marching', 'fft_rows': 'hybrid'}[best]} |", flush=True)
