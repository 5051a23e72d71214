"""Developer tool: parity of the marching Lenia kernel against the oracle on a few shapes + step time against the
row-FFT hybrid (CUDA events, ring of worlds larger than L2 for the small shapes).
    python tools/march_check.py [--time-only]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pyca_b200  # noqa: E402
from oracle import pyca_oracle as orc  # noqa: E402
from tests.util import load_npz, t  # noqa: E402


def params(C):
    p = load_npz("params_lenia_abominable_mongolic.npz")
    d = {k: t(p[k]) for k in ["mu", "sigma", "beta", "mu_k", "sigma_k", "weights"]}
    d["k_size"] = int(p["k_size"])
    if C != 3:
        d = {k: (v[:, :C, :C].clone() if isinstance(v, torch.Tensor) else v) for k, v in d.items()}
    P = pyca_b200.LeniaParams(param_dict=d, channels=C)
    pd = P.param_dict
    K = orc.lenia_kernel(pd["beta"].cpu(), pd["mu_k"].cpu(), pd["sigma_k"].cpu(), 31)
    return P, K, pd["mu"].cpu(), pd["sigma"].cpu(), pd["weights"].cpu()


def check():
    ok = True
    for C, shape in [(1, (64, 256)), (1, (48, 224)), (1, (33, 449)), (3, (96, 300)), (1, (1024, 1024)), (3, (512, 1024)), (1, (600, 260))]:
        P, K, mu, sg, w = params(C)
        g = torch.Generator().manual_seed(7)
        s0 = torch.rand(1, C, *shape, generator=g)
        a = pyca_b200.Lenia(shape, num_channels=C, params=P, state_init=s0, device="cuda", conv_path="fft")
        a.step()
        ref = orc.lenia_step(s0, K, mu, sg, w, 0.1)
        e1 = (a.state.cpu() - ref).abs().max().item()
        a.step()
        ref2 = orc.lenia_step(ref, K, mu, sg, w, 0.1)
        e2 = (a.state.cpu() - ref2).abs().max().item()
        print(f"C={C} {shape}: step1 {e1:.2e} step2 {e2:.2e} {pyca_b200._lib.variants()}", flush=True)
        ok &= e1 <= 1e-5 and e2 <= 3e-5
    print("MARCH_CHECK", "PASS" if ok else "FAIL", flush=True)
    return ok


def timeit(C, shape, path, reps=40, cls=None, **kw):
    P, *_ = params(C)
    per_world = shape[0] * shape[1] * C * 16
    nw = 1 if per_world > (126 << 20) else -(-(256 << 20) // per_world)
    g = torch.Generator().manual_seed(1)
    cls = cls or pyca_b200.Lenia
    autos = [cls(shape, num_channels=C, params=P, state_init=torch.rand(1, C, *shape, generator=g), device="cuda", conv_path=path, **kw)
             for _ in range(nw)]
    for a in autos:
        a.step()
        a.step()
    torch.cuda.synchronize()
    n = max(reps, 2 * nw) // nw * nw
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(3):
        e0.record()
        for i in range(n):
            autos[i % nw].step()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / n)
    cells = shape[0] * shape[1]
    print(f"C={C} {shape} path={path}: {best * 1e3:.1f} us/step  {cells / best / 1e6:.1f} GCUPS  ({nw} worlds) {pyca_b200._lib.variants()[-2:]}", flush=True)
    del autos
    torch.cuda.empty_cache()


if __name__ == "__main__":
    if "--time-only" not in sys.argv:
        if not check():
            sys.exit(1)
    for C, shape in [(1, (1024, 1024)), (3, (1024, 1024)), (1, (4096, 4096)), (1, (16384, 16384)), (3, (4096, 4096))]:
        for path in ("fft", "fft_rows"):
            timeit(C, shape, path, reps=24 if shape[0] >= 4096 else 240)
