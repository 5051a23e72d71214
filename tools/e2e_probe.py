"""Where does the end-to-end (host buffer) Lenia 1024^2 step spend its time?  PCIe copies alone, both directions at once,
the submit cost on the CPU, and HostPipeline at several depths."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pyca_b200
from pyca_b200.pipeline import HostPipeline
from bench import load_params

H = W = 1024
dev = "cuda"
hin = [torch.rand(1, 1, H, W).pin_memory() for _ in range(4)]
hout = [torch.empty(1, 1, H, W).pin_memory() for _ in range(4)]
d = torch.empty(1, 1, H, W, device=dev)
n = 200


def timeit(fn, n=n):
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e6


print("H2D 4MB us", timeit(lambda: d.copy_(hin[0], non_blocking=True)))
print("D2H 4MB us", timeit(lambda: hout[0].copy_(d, non_blocking=True)))
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
d2 = torch.empty_like(d)


def both():
    with torch.cuda.stream(s1):
        d.copy_(hin[0], non_blocking=True)
    with torch.cuda.stream(s2):
        hout[0].copy_(d2, non_blocking=True)


print("H2D+D2H concurrently us", timeit(both))
dct, _ = load_params("params_lenia_abominable_mongolic.npz")
dct = {k: (v[:, :1, :1].clone() if isinstance(v, torch.Tensor) else v) for k, v in dct.items()}
params = pyca_b200.LeniaParams(param_dict=dct, channels=1)
mk = lambda: pyca_b200.Lenia((H, W), num_channels=1, params=params, state_init=hin[0], device=dev)
a = mk()
a.step()
print("device step us", timeit(a.step))


def sub_nocopy():
    a.state = d
    a.step()


print("state= + step (device tensor) us", timeit(sub_nocopy))
for depth in (1, 2, 3, 4, 6):
    pipe = HostPipeline(mk, depth=depth)
    i = [0]

    def f():
        k = i[0] % 4
        i[0] += 1
        pipe.submit(hin[k], hout[k])

    us = timeit(f)
    pipe.drain()
    print(f"pipeline depth {depth}: {us:.1f} us/world")
# CPU-side submit cost: time only the submit calls of one burst shorter than the lanes' depth
pipe = HostPipeline(mk, depth=8)
pipe.drain()
torch.cuda.synchronize()
t0 = time.perf_counter()
for k in range(8):
    pipe.submit(hin[k % 4], hout[k % 4])
t1 = time.perf_counter()
pipe.drain()
print("CPU submit cost us/world", (t1 - t0) / 8 * 1e6)
from pyca_b200.pipeline import LeniaHostPipe
for depth in (2, 3, 4):
    cp = LeniaHostPipe(a, depth=depth)
    i = [0]

    def f():
        k = i[0] % 4
        i[0] += 1
        cp.submit(hin[k], hout[k])

    us = timeit(f)
    cp.drain()
    print(f"C pipe depth {depth}: {us:.1f} us/world")
    cp.close()
