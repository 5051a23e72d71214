"""Developer tool (library built with `make DEV=1`, B200CA_LIB=pyca_b200/libb200ca_dev.so): time per generation of the open-rows
GoL kernel on slab-sized worlds for each rows-per-thread variant (B200CA_GOL_RPT)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import sys, torch
sys.path.insert(0, %r)
import pyca_b200
for H in (8320, 4096 + 128, 16384 + 128, 65536):
    pw = pyca_b200.PackedWorld(H, 65536, "cuda")
    pw.random_fill(seed=1)
    pw.step(0b1100, 0b1000, 64, vert_open=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(3):
        e0.record()
        pw.step(0b1100, 0b1000, 256, vert_open=True)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 256)
    print(f"  H={H}: {best * 1e3:.2f} us/generation  ({H * 65536 / best / 1e9:.1f} TCUPS)", pyca_b200._lib.variants()[-1:], flush=True)
    del pw
''' % ROOT
for evf, rpt in (("0", "8"), ("1", "8"), ("1", "16"), ("0", "16"), ("0", "4")):   # B200CA_GOL_EVF: evict-first loads (gol.cu)
    print("B200CA_GOL_EVF", evf, "RPT", rpt, flush=True)
    env = dict(os.environ, B200CA_LIB=os.path.join(ROOT, "pyca_b200", "libb200ca_dev.so"), B200CA_GOL_RPT=rpt, B200CA_GOL_EVF=evf)
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    print(r.stdout.rstrip() or r.stderr[-600:], flush=True)
