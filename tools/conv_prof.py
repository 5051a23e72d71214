"""Developer tool: one 4096^2 C=1 Lenia step per (path, k_size) for ncu (profiles/r02_conv_paths.md)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pyca_b200  # noqa: E402
from tools.conv_paths import params  # noqa: E402

g = torch.Generator().manual_seed(1)
for path, ks in (("direct", 7), ("direct", 15), ("direct", 31), ("fft", 31), ("fft_rows", 31)):
    a = pyca_b200.Lenia((4096, 4096), num_channels=1, params=params(1, ks), state_init=torch.rand(1, 1, 4096, 4096, generator=g), device="cuda",
                        conv_path=path)
    for _ in range(3):
        a.step()
    torch.cuda.synchronize()
    print(path, ks, pyca_b200._lib.variants()[-2:])
    del a
