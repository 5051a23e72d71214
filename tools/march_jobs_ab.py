"""Developer tool: step time of the marching Lenia kernel with fixed-height bands (B200CA_MARCH_WAVES=0) against the balanced
launch with 1 / 2 / 4 waves of CTAs, on the shapes that matter (headline batch, 4096^2, an 8-GPU slab's shape, 16384^2, C = 3).
Needs the developer library (`make -C pyca_b200/csrc DEV=1`): one subprocess per setting.
    python tools/march_jobs_ab.py [waves ...]"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE = r'''
import sys, torch
sys.path.insert(0, %r)
import pyca_b200
from tools.march_check import params
def run(C, B, H, W, reps):
    P, *_ = params(C)
    if B > 1:
        P = P.expand(B)
    per = B * C * H * W * 16
    nw = 1 if per > (126 << 20) else -(-(256 << 20) // per)
    g = torch.Generator().manual_seed(1)
    size = (B, H, W) if B > 1 else (H, W)
    autos = [pyca_b200.Lenia(size, num_channels=C, params=P, state_init=torch.rand(B, C, H, W, generator=g), device="cuda", conv_path="fft")
             for _ in range(nw)]
    for a in autos:
        a.step(); a.step()
    torch.cuda.synchronize()
    n = max(reps, 2 * nw) // nw * nw
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(4):
        e0.record()
        for i in range(n):
            autos[i %% nw].step()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / n)
    print(f"  C={C} B={B} {H}x{W}: {best * 1e3:8.1f} us/step {B * H * W / best / 1e6:7.1f} GCUPS  {pyca_b200._lib.variants()[-1:]}", flush=True)
    del autos
    torch.cuda.empty_cache()
for spec in %r:
    run(*spec)
'''
SPECS = [(1, 24, 1024, 1024, 60), (1, 1, 4096, 4096, 60), (1, 1, 2048, 16384, 40), (1, 1, 16384, 16384, 12), (3, 1, 4096, 4096, 12),
         (3, 1, 1024, 1024, 100), (1, 1, 1024, 1024, 200)]
if os.environ.get("MARCH_AB_SHORT"):
    SPECS = SPECS[:4]
if os.environ.get("MARCH_AB_PRO"):
    SPECS = [SPECS[0], SPECS[1], SPECS[4], (3, 1, 2048, 2048, 40)]

if __name__ == "__main__":
    waves = sys.argv[1:] or ["0", "1", "2", "4"]      # "<waves>" or "<variant>:<waves>" (pyca_b200/variants/<variant>.so)
    for w0 in waves:
        w = w0
        lib = os.path.join(ROOT, "pyca_b200", "libb200ca_dev.so")
        skew, rest = None, []
        if ":" in w:                                   # "<variant>:<waves>[:<skew per mille>]"; variant "" = the developer library
            v, w, *rest = w.split(":")
            if v:
                lib = os.path.join(ROOT, "pyca_b200", "variants", v + ".so")
            skew = rest[0] if rest else None
        duo = rest[1] if len(rest) > 1 else None        # "<variant>:<waves>:<skew>:<duo 0/1>"
        print(f"=== {os.path.basename(lib)} B200CA_MARCH_WAVES={w} B200CA_MARCH_SKEW={skew} B200CA_MARCH_DUO={duo}", flush=True)
        env = dict(os.environ, B200CA_LIB=lib, B200CA_MARCH_WAVES=w)
        if skew:
            env["B200CA_MARCH_SKEW"] = skew
        if duo is not None:
            env["B200CA_MARCH_DUO"] = duo
        try:
            r = subprocess.run([sys.executable, "-c", CODE % (ROOT, SPECS)], env=env, capture_output=True, text=True, timeout=50)
        except subprocess.TimeoutExpired:
            print("  TIMEOUT", flush=True)
            continue
        print(r.stdout.rstrip() or r.stderr[-1500:], flush=True)
