import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pyca_b200
from tools.quick_bench import timeit, pd
d3 = pd("params_lenia_abominable_mongolic.npz")
d1 = {k: (v[:, :1, :1].clone() if isinstance(v, torch.Tensor) else v) for k, v in d3.items()}
H, W = 4096, 3976
a = pyca_b200.Lenia((H, W), num_channels=1, params=d1, state_init=torch.rand(1, 1, H, W), device="cuda", conv_path="fft")
ms = timeit(a.step, 20)
print(f"rows={os.environ.get('B200CA_FFT_RPC_ROWS')} inv={os.environ.get('B200CA_FFT_RPC_INV')}: {ms*1e3:.1f} us -> {H*W/ms/1e6:.1f} GCUPS")
