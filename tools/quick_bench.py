"""Developer micro-benchmark: device-side GCUPS of every model at the BASELINE shapes (CUDA events)."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pyca_b200  # noqa: E402
from tests.util import load_npz, t  # noqa: E402


def timeit(fn, steps, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps  # ms


def pd(name):
    p = load_npz(name)
    d = {k: t(p[k]) for k in ["mu", "sigma", "beta", "mu_k", "sigma_k", "weights"]}
    d["k_size"] = int(p["k_size"])
    return d


def main():
    which = sys.argv[1:] or ["gol", "lenia", "mace", "nca"]
    if "gol" in which:
        for (H, W) in [(250, 250), (1024, 1024), (4096, 4096), (16384, 16384), (65536, 65536)]:
            pw = pyca_b200.PackedWorld(H, W, "cuda")
            pw.random_fill(1)
            ms = timeit(lambda: pw.step(0b1100, 0b1000, 1), 50)
            ms10 = timeit(lambda: pw.step(0b1100, 0b1000, 10), 10) / 10
            ms_g = timeit(lambda: pw.step(0b1101, 0b1010, 10), 5) / 10
            print(f"GoL {H}x{W}: {ms*1e3:.1f} us/step (1/call) {ms10*1e3:.1f} us/step (10/call) -> {H*W/ms10/1e6:.1f} GCUPS;"
                  f" generic rule {H*W/ms_g/1e6:.1f} GCUPS; bytes {H*W/4/ms10/1e6:.1f} GB/s", flush=True)
            del pw
    if "lenia" in which:
        d3 = pd("params_lenia_abominable_mongolic.npz")
        d1 = {k: (v[:, :1, :1].clone() if isinstance(v, torch.Tensor) else v) for k, v in d3.items()}
        for (C, d, H, W) in [(1, d1, 1024, 1024), (3, d3, 1024, 1024), (1, d1, 4096, 4096), (3, d3, 4096, 4096), (1, d1, 16384, 16384)]:
            for path in ("direct", "fft"):
                a = pyca_b200.Lenia((H, W), num_channels=C, params=d, state_init=torch.rand(1, C, H, W), device="cuda", conv_path=path)
                ms = timeit(a.step, 10 if H >= 4096 else 50)
                fma = H * W * C * C * 496
                print(f"Lenia[{path}] C={C} {H}x{W}: {ms:.3f} ms/step -> {H*W/ms/1e6:.2f} GCUPS; {fma/ms/1e9:.2f} TFMA/s (folded-direct equiv);"
                      f" alg bytes {H*W*C*8/ms/1e6:.1f} GB/s", flush=True)
                del a
    if "mace" in which:
        d = pd("params_xchan_arborical_asbestosis.npz")
        for (H, W) in [(1024, 1024), (4096, 4096)]:
            a = pyca_b200.MaCELeniaXChan((H, W), params=d, state_init=torch.rand(1, 3, H, W), device="cuda")
            ms = timeit(a.step, 5 if H >= 4096 else 20)
            print(f"MaCE-XChan C=3 {H}x{W}: {ms:.3f} ms/step -> {H*W/ms/1e6:.2f} GCUPS", flush=True)
            fs = a._fs
            from pyca_b200 import _lib
            ms2 = timeit(lambda: _lib.call("b200ca_mace_redistribute", _lib.ptr(fs.buf), _lib.ptr(a._aff), _lib.ptr(fs.other), 1, 3, H, W,
                                           6.0, 1, 0.03, 1, _lib.current_stream("cuda")), 20)
            print(f"   redistribute kernel alone: {ms2*1e3:.1f} us -> {H*W*3*12/ms2/1e6:.1f} GB/s algorithmic", flush=True)
            del a
    if "nca" in which:
        fx = load_npz("nca_betta.npz")
        w = pyca_b200.NCAWeights(t(fx["W1"]), t(fx["b1"]), t(fx["W2"]), t(fx["b2"]), "cuda")
        for B in (16, 256):
            s = torch.rand(B, 16, 128, 128, device="cuda")
            o = torch.empty_like(s)
            scratch = torch.empty(pyca_b200.lib().b200ca_nca_scratch_bytes(B, 128, 128), dtype=torch.uint8, device="cuda")
            for impl in (0, 1):
                try:
                    ms = timeit(lambda: pyca_b200.nca_step(s, w, out=o, seed=1, step_index=0, scratch=scratch, impl=impl), 5)
                except pyca_b200.B200CAError as e:
                    print(f"NCA impl {impl}: {e}")
                    continue
                cells = B * 128 * 128
                print(f"NCA impl={impl} B={B}x128x128: {ms:.3f} ms/step -> {cells/ms/1e6:.2f} GCUPS; {cells*128/ms/1e6:.1f} GB/s;"
                      f" {cells*12288/ms/1e9:.2f} TFLOP/s (folded MLP)", flush=True)


if __name__ == "__main__":
    main()
