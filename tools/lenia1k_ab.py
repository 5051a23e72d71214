"""A/B of the Lenia 1024^2 step: device time per step over a ring of 12 worlds (inputs evicted from L2), C = 1 and C = 3."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pyca_b200
from bench import load_params

d3, _ = load_params("params_lenia_abominable_mongolic.npz")
for C in (1, 3):
    d = {k: (v[:, :C, :C].clone() if isinstance(v, torch.Tensor) else v) for k, v in d3.items()}
    params = pyca_b200.LeniaParams(param_dict=d, channels=C)
    autos = [pyca_b200.Lenia((1024, 1024), num_channels=C, params=params, state_init=torch.rand(1, C, 1024, 1024), device="cuda")
             for _ in range(12)]
    for a in autos:
        a.step()
    torch.cuda.synchronize()
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(240):
            autos[i % 12].step()
        e1.record()
        torch.cuda.synchronize()
        print(f"C={C} NT512={os.environ.get('B200CA_FFT_NT512', '1')}: {e0.elapsed_time(e1) / 240 * 1e3:.2f} us/step", flush=True)
