"""Developer tool (under torchrun, one rank per GPU): step time of the 16384^2 Lenia world in row slabs with the in-kernel halo
exchange (LeniaSlabP2P), for the library in B200CA_LIB.  Max over ranks of CUDA-event times, median of blocks.
    B200CA_LIB=... python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/slab_ab.py"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dev = f"cuda:{lr}"
    dist.init_process_group("nccl", device_id=torch.device(dev))
    import pyca_b200
    from pyca_b200.slab import LeniaSlabP2P
    from tools.march_check import params

    P, *_ = params(1)
    H = W = 16384
    sl = LeniaSlabP2P(P, 1, H, W, rank, world, device=dev)
    g = torch.Generator(device=dev).manual_seed(4242 + rank)
    sl.set_interior(torch.rand(1, 1, sl.hl, W, generator=g, device=dev))
    for _ in range(10):
        sl.step()
    torch.cuda.synchronize()
    times = []
    for _ in range(9):
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(40):
            sl.step()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / 40], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        times.append(float(t.item()))
    times.sort()
    if rank == 0:
        ms = times[len(times) // 2]
        print(f"{os.path.basename(os.environ.get('B200CA_LIB', 'libb200ca.so'))}: {world} slabs of {sl.hl} rows: {ms * 1e3:.1f} us/step = "
              f"{H * W / ms / 1e6:.1f} GCUPS  timed out: {sl.timed_out()}  {pyca_b200._lib.variants()[-1:]}", flush=True)
    sl.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
