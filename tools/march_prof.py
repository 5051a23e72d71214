"""Developer tool: a few Lenia steps of one shape / path for ncu.   python tools/march_prof.py C H W path [steps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pyca_b200  # noqa: E402
from tools.march_check import params  # noqa: E402

C, H, W, path = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
steps = int(sys.argv[5]) if len(sys.argv) > 5 else 4
P, *_ = params(C)
g = torch.Generator().manual_seed(1)
a = pyca_b200.Lenia((H, W), num_channels=C, params=P, state_init=torch.rand(1, C, H, W, generator=g), device="cuda", conv_path=path)
for _ in range(steps):
    a.step()
torch.cuda.synchronize()
print("done", pyca_b200._lib.variants()[-1])
