"""Copy/compute overlap probe: which combination of H2D, kernels and D2H overlaps on this box?"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pyca_b200
from bench import load_params

H = W = 1024
hin = [torch.rand(1, 1, H, W).pin_memory() for _ in range(4)]
hout = [torch.empty(1, 1, H, W).pin_memory() for _ in range(4)]
L = 4
streams = [torch.cuda.Stream() for _ in range(L)]
din = [torch.empty(1, 1, H, W, device="cuda") for _ in range(L)]
dout = [torch.empty(1, 1, H, W, device="cuda") for _ in range(L)]
big = [torch.empty(8 << 20, device="cuda") for _ in range(L)]  # 32 MB fill ~ 10 us
dct, _ = load_params("params_lenia_abominable_mongolic.npz")
dct = {k: (v[:, :1, :1].clone() if isinstance(v, torch.Tensor) else v) for k, v in dct.items()}
params = pyca_b200.LeniaParams(param_dict=dct, channels=1)
autos = []
for s in streams:
    with torch.cuda.stream(s):
        a = pyca_b200.Lenia((H, W), num_channels=1, params=params, state_init=hin[0], device="cuda")
        a.step()
        autos.append(a)
torch.cuda.synchronize()


def run(h2d, kern, d2h, n=300):
    def one(i):
        k = i % L
        with torch.cuda.stream(streams[k]):
            if h2d:
                din[k].copy_(hin[k], non_blocking=True)
            if kern == "fill":
                big[k].fill_(1.0)
                big[k].fill_(2.0)
                big[k].fill_(3.0)
            elif kern == "lenia":
                autos[k].step()
            if d2h:
                hout[k].copy_(dout[k], non_blocking=True)
    for i in range(20):
        one(i)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(n):
        one(i)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e6


for cfg in [(1, None, 0), (0, None, 1), (1, None, 1), (0, "fill", 0), (0, "lenia", 0), (1, "fill", 0), (0, "fill", 1), (1, "fill", 1),
            (1, "lenia", 0), (0, "lenia", 1), (1, "lenia", 1)]:
    print(cfg, f"{run(*cfg):.1f} us")

# variant: one stream per ENGINE (up / compute / down) with events between them instead of one stream per lane
up, comp, down = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
ev_up = [torch.cuda.Event() for _ in range(L)]
ev_c = [torch.cuda.Event() for _ in range(L)]
ev_dn = [torch.cuda.Event() for _ in range(L)]


def run3(n=300, kern="lenia"):
    def one(i):
        k = i % L
        with torch.cuda.stream(up):
            up.wait_event(ev_c[k])      # lane k's previous input has been consumed
            din[k].copy_(hin[k], non_blocking=True)
            ev_up[k].record(up)
        with torch.cuda.stream(comp):
            comp.wait_event(ev_up[k])
            comp.wait_event(ev_dn[k])   # lane k's previous output has left
            if kern == "lenia":
                autos[k].step()
            else:
                big[k].fill_(1.0); big[k].fill_(2.0); big[k].fill_(3.0)
            ev_c[k].record(comp)
        with torch.cuda.stream(down):
            down.wait_event(ev_c[k])
            hout[k].copy_(dout[k], non_blocking=True)
            ev_dn[k].record(down)
    for i in range(20):
        one(i)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(n):
        one(i)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e6


print("engine streams, lenia", f"{run3():.1f} us")
print("engine streams, fill", f"{run3(kern='fill'):.1f} us")

# variant: lanes, but the D2H of job i is issued after the H2D of job i+1 (software-pipelined issue order)
def run_sp(n=300):
    def one(i):
        k = i % L
        with torch.cuda.stream(streams[k]):
            din[k].copy_(hin[k], non_blocking=True)
            autos[k].step()
        kp = (i - 1) % L
        with torch.cuda.stream(streams[kp]):
            hout[kp].copy_(dout[kp], non_blocking=True)
    for i in range(20):
        one(i)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(n):
        one(i)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e6


print("lanes, D2H issued one job late", f"{run_sp():.1f} us")
# chunked copies: 4 x 1 MB per direction
def run_chunk(n=300, parts=4):
    rows = H // parts
    def one(i):
        k = i % L
        with torch.cuda.stream(streams[k]):
            for c in range(parts):
                din[k][0, 0, c * rows:(c + 1) * rows].copy_(hin[k][0, 0, c * rows:(c + 1) * rows], non_blocking=True)
            autos[k].step()
            for c in range(parts):
                hout[k][0, 0, c * rows:(c + 1) * rows].copy_(dout[k][0, 0, c * rows:(c + 1) * rows], non_blocking=True)
    for i in range(20):
        one(i)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(n):
        one(i)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e6


print("lanes, 4 chunks per copy", f"{run_chunk():.1f} us")
