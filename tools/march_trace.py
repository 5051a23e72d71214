"""Developer tool (developer library): per-CTA start / end times of one balanced launch of the marching Lenia kernel
(B200CA_MARCH_TRACE = device pointer of a 3 x grid uint64 buffer) next to the bands the plan gave each CTA.
    B200CA_LIB=pyca_b200/libb200ca_dev.so python tools/march_trace.py C B H W"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
C, B, H, W = map(int, sys.argv[1:5])
buf = torch.zeros(3 * 2048, dtype=torch.int64, device="cuda")
os.environ["B200CA_MARCH_TRACE"] = hex(buf.data_ptr())
import pyca_b200  # noqa: E402
from tools.march_check import params  # noqa: E402

P, *_ = params(C)
if B > 1:
    P = P.expand(B)
g = torch.Generator().manual_seed(1)
a = pyca_b200.Lenia((B, H, W) if B > 1 else (H, W), num_channels=C, params=P, state_init=torch.rand(B, C, H, W, generator=g), device="cuda",
                    conv_path="fft")
for _ in range(5):
    a.step()
torch.cuda.synchronize()
L = pyca_b200._lib.lib()
gr, k, units, cols = (ctypes.c_int() for _ in range(4))
L.b200ca_lenia_march_plan(H, W, B, ctypes.byref(gr), ctypes.byref(k), ctypes.byref(units), ctypes.byref(cols))
G = gr.value
t = buf.cpu().view(-1, 3)[:G]
t0 = int(t[:, 0].min())
start, end, sm = (t[:, 0] - t0).double() / 1e3, (t[:, 1] - t0).double() / 1e3, t[:, 2]
bands = (ctypes.c_int * (3 * 64))()
rows = []
for cta in range(G):
    n = L.b200ca_lenia_march_plan_bands(H, W, B, cta, bands, 64)
    nu = sum(bands[3 * i + 2] for i in range(n))
    top = sum(1 for i in range(n) if bands[3 * i + 1] == 0)
    bot = sum(1 for i in range(n) if bands[3 * i + 1] + bands[3 * i + 2] == units.value)
    rows.append((cta, n, nu, top, bot, float(start[cta]), float(end[cta]), int(sm[cta])))
print(f"grid {G} split_k {k.value} units {units.value} columns {cols.value}; launch span {float(end.max()):.1f} us; starts {float(start.min()):.1f}..{float(start.max()):.1f} us; "
      f"ends {float(end.min()):.1f}..{float(end.max()):.1f} us")
import collections
groups = collections.defaultdict(list)
for r in rows:
    groups[(r[1], r[2], r[3], r[4])].append(r[6] - r[5])
print("(bands, units, tops, bottoms): CTAs, mean / min / max duration us")
for key in sorted(groups):
    v = groups[key]
    print(f"  {key}: {len(v):4d}  {sum(v) / len(v):7.1f} {min(v):7.1f} {max(v):7.1f}")
late = sorted(rows, key=lambda r: -r[6])[:8]
print("latest finishers (cta, bands, units, tops, bottoms, start, end, sm):", [tuple(round(x, 1) if isinstance(x, float) else x for x in r) for r in late])
if len(sys.argv) > 5:
    import json
    json.dump(rows, open(sys.argv[5], "w"))
