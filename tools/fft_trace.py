"""Developer tool: per-CTA phase timestamps of the fused FIR+inverse kernel (library built with -DB200CA_FFT_TRACE)."""
import ctypes, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pyca_b200
from pyca_b200 import _lib
from tools.quick_bench import pd
d3 = pd("params_lenia_abominable_mongolic.npz")
d1 = {k: (v[:, :1, :1].clone() if isinstance(v, torch.Tensor) else v) for k, v in d3.items()}
H, W = [int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "1024x1024").split("x")]
a = pyca_b200.Lenia((H, W), num_channels=1, params=d1, state_init=torch.rand(1, 1, H, W), device="cuda", conv_path="fft")
for _ in range(5):
    a.step()
torch.cuda.synchronize()
L = ctypes.CDLL(_lib.LIB_PATH)
n = 8 * 4096
out = np.zeros(n, dtype=np.int64)
rc = L.b200ca_debug_fft_trace(out.ctypes.data_as(ctypes.c_void_p), n)
tr = out.reshape(4096, 8)
tr = tr[tr[:, 0] != 0]
print("rc", rc, "ctas traced", len(tr))
names = ["prologue(tw+taps issue)", "pdl wait", "FIR", "sync", "IFFT..stage wait", "cp.async wait", "growth+store"]
d = np.diff(tr, axis=1) / 1.965e3  # us
for i, nm in enumerate(names):
    print(f"{nm:28s} median {np.median(d[:, i]):7.2f} us   p90 {np.percentile(d[:, i], 90):7.2f}   max {d[:, i].max():7.2f}")
print("total per CTA median", np.median((tr[:, 7] - tr[:, 0]) / 1.965e3), "us")
