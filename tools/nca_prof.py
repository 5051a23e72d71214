"""Developer tool: a few NCA steps at 256 x 128 x 128 for ncu.   python tools/nca_prof.py [steps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pyca_b200  # noqa: E402
from tests.util import load_npz, t  # noqa: E402

fx = load_npz("nca_betta.npz")
w = pyca_b200.NCAWeights(t(fx["W1"]), t(fx["b1"]), t(fx["W2"]), t(fx["b2"]), "cuda")
g = torch.Generator().manual_seed(0)
a = torch.rand(256, 16, 128, 128, generator=g).cuda()
b = torch.empty_like(a)
for i in range(int(sys.argv[1]) if len(sys.argv) > 1 else 4):
    pyca_b200.nca_step(a, w, out=b, seed=1, step_index=i, impl=1)
    a, b = b, a
torch.cuda.synchronize()
print("done")
