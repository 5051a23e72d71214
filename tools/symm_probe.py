"""Developer tool: does memory from torch.distributed._symmetric_memory behave like cudaMalloc memory for the local kernels?
Times the marching Lenia step on an 8192 x 16384 slab whose buffers are (a) plain torch allocations, (b) symmetric memory."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pyca_b200  # noqa: E402
from pyca_b200 import _lib  # noqa: E402
from pyca_b200.slab import LeniaSlabCUDA, LeniaSlabP2P  # noqa: E402
from tools.march_check import params  # noqa: E402

os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
os.environ.setdefault("MASTER_PORT", "29533")
torch.cuda.set_device(0)
dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda:0"))
P, *_ = params(1)
H, W = 8192, 16384


def run(slab, label, n=20):
    g = torch.Generator(device="cuda").manual_seed(1)
    slab.set_interior(torch.rand(1, 1, slab.hl, W, generator=g, device="cuda"))
    for _ in range(3):
        slab.step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        slab.step()
    e1.record()
    torch.cuda.synchronize()
    print(f"{label}: {e0.elapsed_time(e1) / n * 1e3:.1f} us/step {_lib.variants()[-1:]}", flush=True)


a = pyca_b200.Lenia((H, W), num_channels=1, params=P, state_init=torch.rand(1, 1, H, W), device="cuda", conv_path="fft")
a.step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    a.step()
e1.record()
torch.cuda.synchronize()
print(f"torus, torch memory: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us/step", flush=True)
del a
sl = LeniaSlabP2P(P, 1, H, W, 0, 1, device="cuda")
run(sl, "world=1 slab, symmetric memory, p2p-halo kernel (self neighbour)")


def plain_step(bufs, label, n=20):
    """the ordinary marching step (row_wrap = 0, no peers) on the given pair of framed buffers"""
    st = torch.cuda.current_stream().cuda_stream
    def one(i):
        _lib.call("b200ca_lenia_step_march", bufs[i & 1].data_ptr(), bufs[(i & 1) ^ 1].data_ptr(), sl._kmarch.data_ptr(), sl.mu.data_ptr(),
                  sl.sigma.data_ptr(), sl.weights.data_ptr(), 1, 1, 1, H, W, 0.1, 0, 0, 0, 0, st)
    for i in range(4):
        one(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        one(i)
    e1.record()
    torch.cuda.synchronize()
    print(f"{label}: {e0.elapsed_time(e1) / n * 1e3:.1f} us/step", flush=True)


plain_step(sl.bufs, "ordinary marching step on the SYMMETRIC buffers")
mine = [torch.rand_like(sl.bufs[0]) for _ in range(2)]
plain_step(mine, "ordinary marching step on torch.rand buffers of the same shape")
dist.destroy_process_group()
