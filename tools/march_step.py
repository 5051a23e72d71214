"""Developer tool: steps one Lenia world (or a batch) a few times on the marching kernel -- the thing to run under ncu.
    python tools/march_step.py C B H W [steps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pyca_b200  # noqa: E402
from tools.march_check import params  # noqa: E402

C, B, H, W = map(int, sys.argv[1:5])
steps = int(sys.argv[5]) if len(sys.argv) > 5 else 4
P, *_ = params(C)
if B > 1:
    P = P.expand(B)
g = torch.Generator().manual_seed(1)
a = pyca_b200.Lenia((B, H, W) if B > 1 else (H, W), num_channels=C, params=P, state_init=torch.rand(B, C, H, W, generator=g), device="cuda",
                    conv_path="fft")
for _ in range(steps):
    a.step()
torch.cuda.synchronize()
print(pyca_b200._lib.variants()[-1:])
