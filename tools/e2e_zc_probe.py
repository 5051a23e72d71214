"""Developer tool: LeniaHostPipe with the copy engines against kernels that read / write the pinned host buffers directly
(B200CA_PIPE_ZEROCOPY bits: 1 = load, 2 = store; developer library).  One subprocess per setting."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE = r'''
import sys, time, os
sys.path.insert(0, %r)
import torch, pyca_b200
from pyca_b200.pipeline import LeniaHostPipe
from bench import load_params
H = W = 1024
g = torch.Generator().manual_seed(5)
hin = [torch.rand(1, 1, H, W, generator=g).pin_memory() for _ in range(8)]
hout = [torch.empty(1, 1, H, W).pin_memory() for _ in range(8)]
dct, _ = load_params("params_lenia_abominable_mongolic.npz")
dct = {k: (v[:, :1, :1].clone() if isinstance(v, torch.Tensor) else v) for k, v in dct.items()}
params = pyca_b200.LeniaParams(param_dict=dct, channels=1)
a = pyca_b200.Lenia((H, W), num_channels=1, params=params, state_init=hin[0], device="cuda")
ref = None
for depth in (3, 6, 8):
    cp = LeniaHostPipe(a, depth=depth)
    for r in range(3):
        for i in range(40):
            cp.submit(hin[i %% 8], hout[i %% 8])
        cp.drain()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 400
    for i in range(n):
        cp.submit(hin[i %% 8], hout[i %% 8])
    cp.drain()
    us = (time.perf_counter() - t0) / n * 1e6
    print(f"  depth {depth}: {us:.1f} us/world = {H * W / us / 1e3:.2f} GCUPS  checksum {float(hout[3].double().sum()):.6f}", flush=True)
    cp.close()
'''
for spec in sys.argv[1:] or ["0:0", "0:3", "1:0", "1:1"]:      # <store straight to the host buffer>:<copy mode: 1 upload stream, 2 download stream>
    zc, cm = spec.split(":")
    print(f"=== B200CA_PIPE_ZEROCOPY={zc} B200CA_PIPE_COPY_MODE={cm}", flush=True)
    env = dict(os.environ, B200CA_LIB=os.path.join(ROOT, "pyca_b200", "libb200ca_dev.so"), B200CA_PIPE_ZEROCOPY=zc, B200CA_PIPE_COPY_MODE=cm)
    r = subprocess.run([sys.executable, "-c", CODE % ROOT], env=env, capture_output=True, text=True)
    print(r.stdout.rstrip() or r.stderr[-1500:], flush=True)
