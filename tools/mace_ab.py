"""Developer tool: times b200ca_mace_redistribute alone (MaCE-XChan 4096^2 C=3, and a 512-row slab) for the library in
B200CA_LIB and every pyca_b200/variants/mace_*.so."""
import glob
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE = r'''
import sys, torch
sys.path.insert(0, %r)
import pyca_b200
from pyca_b200 import _lib
from tools.march_check import params
for H, W in [(4096, 4096), (512, 4096), (1024, 1024)]:
    P, *_ = params(3)
    g = torch.Generator().manual_seed(1)
    a = pyca_b200.MaCELeniaXChan((H, W), num_channels=3, params=P, state_init=torch.rand(1, 3, H, W, generator=g), device="cuda")
    a.step(); a.step()
    fs = a._fs
    st = _lib.current_stream(fs.device)
    def run():
        _lib.call("b200ca_mace_redistribute", _lib.ptr(fs.buf), _lib.ptr(a._aff), _lib.ptr(fs.other), fs.B, fs.C, fs.H, fs.W, float(a.b), 1, 0.03, 1, st)
    for _ in range(5): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(4):
        e0.record()
        for _ in range(20): run()
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 20)
    print(f"  {H}x{W}: mace_redistribute {best * 1e3:.1f} us  checksum {float(fs.other.double().sum()):.4f}  {_lib.variants()[-1:]}", flush=True)
'''
libs = [os.path.join(ROOT, "pyca_b200", "libb200ca.so")] + sorted(glob.glob(os.path.join(ROOT, "pyca_b200", "variants", "mace_*.so")))
for lib in libs:
    print("===", os.path.basename(lib), flush=True)
    r = subprocess.run([sys.executable, "-c", CODE % ROOT], env=dict(os.environ, B200CA_LIB=lib), capture_output=True, text=True)
    print(r.stdout.rstrip() or r.stderr[-1500:], flush=True)
