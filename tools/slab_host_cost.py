"""Host-side cost of one Lenia slab step (launches, stream hand-offs, torch.distributed P2P batch), run under torchrun."""
import os, sys, time
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = f"cuda:{lr}"
dist.init_process_group("nccl", device_id=torch.device(dev))
import pyca_b200
from pyca_b200.slab import LeniaSlabCUDA
from bench import load_params
d, _ = load_params("params_lenia_abominable_mongolic.npz")
d = {k: (v[:, :1, :1].clone() if isinstance(v, torch.Tensor) else v) for k, v in d.items()}
H = int(os.environ.get("SLAB_H", 2048)) * world   # 2048 rows per rank = the 8-GPU slab of the 16384^2 world
sl = LeniaSlabCUDA(pyca_b200.LeniaParams(param_dict=d, channels=1), 1, H, 16384, rank, world, device=dev)
sl.set_interior(torch.rand(1, 1, sl.hl, 16384, device=dev))
for _ in range(5):
    sl.step()
torch.cuda.synchronize(); dist.barrier()
n = 40
t0 = time.perf_counter()
for _ in range(n):
    sl.step()
t_host = (time.perf_counter() - t0) / n * 1e6
torch.cuda.synchronize()
t_all = (time.perf_counter() - t0) / n * 1e6
if rank == 0:
    print(f"slab {sl.hl}x16384 per rank, world {world}: host issue {t_host:.0f} us/step, wall (device-bound if larger) {t_all:.0f} us/step", flush=True)
dist.barrier(); dist.destroy_process_group()
