"""Developer tool: step time of a batch of B 1024^2 single-channel Lenia worlds (the headline workload) per batch size."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pyca_b200  # noqa: E402
from tools.march_check import params  # noqa: E402

P, *_ = params(1)
for B in (8, 16, 24, 48, 96):
    g = torch.Generator().manual_seed(B)
    a = pyca_b200.Lenia((B, 1024, 1024), num_channels=1, params=P.expand(B), state_init=torch.rand(B, 1, 1024, 1024, generator=g), device="cuda")
    for _ in range(3):
        a.step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(3):
        e0.record()
        for _ in range(40):
            a.step()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 40)
    print(f"B={B}: {best * 1e3:.1f} us/step = {best * 1e3 / B:.2f} us per world-step, {B * 1024 * 1024 / best / 1e6:.1f} GCUPS {pyca_b200._lib.variants()[-1:]}", flush=True)
    del a
    torch.cuda.empty_cache()
