"""Developer tool: times every pyca_b200/variants/*.so (tools/march_ab.sh) on a few Lenia shapes, one subprocess per
variant (B200CA_LIB selects the library).   python tools/march_ab_run.py [shape spec ...]   spec = C:H:W"""
import glob
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
specs = sys.argv[1:] or ["1:4096:4096", "1:8192:8192", "1:1024:1024"]
code = r'''
import sys, torch
sys.path.insert(0, %r)
from tools.march_check import timeit
for sp in %r:
    C, H, W = map(int, sp.split(":"))
    timeit(C, (H, W), "fft", reps=24 if H >= 4096 else 240)
''' % (ROOT, specs)
for so in sorted(glob.glob(os.path.join(ROOT, "pyca_b200", "variants", "*.so"))):
    print("=== variant", os.path.basename(so), flush=True)
    env = dict(os.environ, B200CA_LIB=so)
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    print(r.stdout.strip() or r.stderr[-800:], flush=True)
