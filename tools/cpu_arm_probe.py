"""Why does the CPU arm time differently in a fresh process and inside the CUDA bench process?  (run on the GPU box)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.dup2(2, 1)
import torch
if len(sys.argv) > 1 and sys.argv[1] == "cuda":
    torch.zeros(1, device="cuda")
    x = torch.empty(1 << 20).pin_memory()
import bench
wl = bench.make_workload("lenia_1k", 0, 1, "cpu")
wl.params = wl._params()
torch.set_num_threads(int(os.environ.get("NT", os.cpu_count())))
wl.cpu_setup()
for rep in range(4):
    t0 = time.perf_counter()
    n = 0
    while time.perf_counter() - t0 < 2.0:
        wl.cpu_step()
        n += 1
    print(f"mode={sys.argv[1:]} threads={torch.get_num_threads()} rep{rep}: {(time.perf_counter() - t0) / n * 1e3:.3f} ms/step", file=sys.stderr)
