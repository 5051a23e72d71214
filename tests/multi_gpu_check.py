"""Multi-GPU parity check (run under torchrun, one rank per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py
N-GPU row-slab result == single-GPU torus result of the same kernels (bit-exact for GoL, <=1e-6 for Lenia)."""
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
    from pyca_b200.slab import GolSlabCUDA, LeniaSlabCUDA
    from tests.util import load_npz, t

    ok = True
    # ---- GoL: H x W torus, ghost 4, 23 generations (several exchanges, odd count)
    H, W, steps = 512 * world, 4096 + 64, 23
    slab = GolSlabCUDA(H, W, rank, world, ghost=4, device=dev)
    slab.random_fill(seed=77)
    slab.run(steps)
    ref = pyca_b200.PackedWorld(H, W, dev)
    ref.random_fill(seed=77)
    ref.step(0b1100, 0b1000, steps)
    same = torch.equal(slab.owned()[:, :ref.wpr], ref.bits[slab.r0:slab.r1, :ref.wpr])
    pop = torch.tensor([slab.population()], device=dev)
    dist.all_reduce(pop)
    ok &= same and int(pop.item()) == ref.population()
    print(f"[rank {rank}] GoL slab == torus: {same}; population {int(pop.item())} vs {ref.population()}", flush=True)

    # ---- the same with the ghost rows pushed by one P2P kernel per exchange (GolSlabP2P, csrc/halo.cu)
    from pyca_b200.slab import GolSlabP2P

    try:
        ps = GolSlabP2P(H, W, rank, world, ghost=4, device=dev)
        ps.random_fill(seed=77)
        ps.run(steps)
        torch.cuda.synchronize()
        dist.barrier()
        same = torch.equal(ps.owned()[:, :ref.wpr], ref.bits[ps.r0:ps.r1, :ref.wpr]) and not ps.timed_out()
        print(f"[rank {rank}] GoL P2P-halo slab == torus: {same} ({ps.epoch} exchanges)", flush=True)
        ok &= same
        ps.close()
    except Exception as e:
        print(f"[rank {rank}] GolSlabP2P unavailable: {e!r}", flush=True)
        ok = False

    # ---- Lenia: C=1 and C=3, 3 steps, overlap on and off
    p = load_npz("params_lenia_abominable_mongolic.npz")
    d = {k: t(p[k]) for k in ["mu", "sigma", "beta", "mu_k", "sigma_k", "weights"]}
    d["k_size"] = int(p["k_size"])
    for C in (1, 3):
        dd = {k: (v[:, :C, :C].clone() if isinstance(v, torch.Tensor) else v) for k, v in d.items()}
        Hh, Ww = 256 * world, 320
        g = torch.Generator().manual_seed(5)
        full = torch.rand(1, C, Hh, Ww, generator=g)
        single = pyca_b200.Lenia((Hh, Ww), num_channels=C, params=dict(dd), state_init=full, device=dev, conv_path="direct")
        for _ in range(3):
            single.step()
        for overlap, path in ((True, "direct"), (False, "direct"), (False, "fft"), (True, "fft")):
            params = pyca_b200.LeniaParams(param_dict=dict(dd), channels=C)
            sl = LeniaSlabCUDA(params, C, Hh, Ww, rank, world, device=dev, overlap=overlap, conv_path=path)
            sl.set_interior(full[:, :, sl.r0:sl.r1].to(dev))
            for _ in range(3):
                sl.step()
            torch.cuda.synchronize()
            err = (sl.interior() - single.state[:, :, sl.r0:sl.r1]).abs().max().item()
            print(f"[rank {rank}] Lenia C={C} path={path} overlap={sl.overlap}: max |slab - torus| = {err:.2e}", flush=True)
            ok &= err <= (1e-6 if path == "direct" else 5e-6)
    # ---- direct path, slab height NOT a multiple of the tile height (the clipped last tile row must still count as boundary:
    #      its rows are sent to the neighbour while the interior launch runs)
    dd = {k: (v[:, :1, :1].clone() if isinstance(v, torch.Tensor) else v) for k, v in d.items()}
    for hl in (104, 90):
        Hh, Ww = hl * world, 320
        g = torch.Generator().manual_seed(hl)
        full = torch.rand(1, 1, Hh, Ww, generator=g)
        single = pyca_b200.Lenia((Hh, Ww), num_channels=1, params=dict(dd), state_init=full, device=dev, conv_path="direct")
        params = pyca_b200.LeniaParams(param_dict=dict(dd), channels=1)
        sl = LeniaSlabCUDA(params, 1, Hh, Ww, rank, world, device=dev, overlap=True, conv_path="direct")
        sl.set_interior(full[:, :, sl.r0:sl.r1].to(dev))
        for _ in range(4):
            single.step()
            sl.step()
        torch.cuda.synchronize()
        err = (sl.interior() - single.state[:, :, sl.r0:sl.r1]).abs().max().item()
        print(f"[rank {rank}] Lenia direct hl={hl} (tile rows {sl.tile_rows}, boundary {sl.n_boundary}/{sl.n_boundary_bot}) "
              f"overlap={sl.overlap}: max |slab - torus| = {err:.2e}", flush=True)
        ok &= err <= 1e-6
    # ---- halo exchange inside the step kernel (LeniaSlabP2P: P2P stores + arrival counters over NVLink, one launch per step)
    from pyca_b200.slab import LeniaSlabP2P

    for C, hl, Ww, steps in ((1, 256, 480, 6), (3, 128, 320, 5), (1, 1024, 1200, 7)):
        dd = {k: (v[:, :C, :C].clone() if isinstance(v, torch.Tensor) else v) for k, v in d.items()}
        Hh = hl * world
        g = torch.Generator().manual_seed(100 + C + hl)
        full = torch.rand(1, C, Hh, Ww, generator=g)
        single = pyca_b200.Lenia((Hh, Ww), num_channels=C, params=dict(dd), state_init=full, device=dev, conv_path="fft")
        params = pyca_b200.LeniaParams(param_dict=dict(dd), channels=C)
        try:
            sl = LeniaSlabP2P(params, C, Hh, Ww, rank, world, device=dev)
        except Exception as e:  # no P2P / symmetric memory on this box: reported, not hidden
            print(f"[rank {rank}] LeniaSlabP2P unavailable: {e!r}", flush=True)
            ok = False
            break
        sl.set_interior(full[:, :, sl.r0:sl.r1].to(dev))
        for _ in range(steps):
            single.step()
            sl.step()
        torch.cuda.synchronize()
        dist.barrier()
        err = (sl.interior() - single.state[:, :, sl.r0:sl.r1]).abs().max().item()
        to = sl.timed_out()
        print(f"[rank {rank}] Lenia P2P-halo slab C={C} hl={hl} W={Ww} {steps} steps: max |slab - torus| = {err:.2e}, timed out: {to}, "
              f"kernels {pyca_b200._lib.variants()[-1:]}", flush=True)
        ok &= err <= 2e-6 and not to
        sl.close()
    # ---- MaCE-Lenia XChan slabs (two exchanges per step: affinity rows, then state rows), 3 steps, mass conserved
    from pyca_b200.slab import MaceSlabCUDA

    px = load_npz("params_xchan_arborical_asbestosis.npz")
    dx = {k: t(px[k]) for k in ["mu", "sigma", "beta", "mu_k", "sigma_k", "weights"]}
    dx["k_size"] = int(px["k_size"])
    Hh, Ww = 192 * world, 320
    g = torch.Generator().manual_seed(11)
    full = 1.5 * torch.rand(1, 3, Hh, Ww, generator=g)
    from pyca_b200.slab import MaceSlabP2P

    for cls in (MaceSlabCUDA, MaceSlabP2P):     # NCCL send/recv exchanges, then one P2P push kernel per exchange (csrc/halo.cu)
        single = pyca_b200.MaCELeniaXChan((Hh, Ww), params=dict(dx), state_init=full, device=dev)
        ms = cls(pyca_b200.LeniaParams(param_dict=dict(dx)), 3, Hh, Ww, rank, world, device=dev, beta=6.0, alpha=0.03)
        ms.set_interior(full[:, :, ms.r0:ms.r1].to(dev))
        for _ in range(5):
            single.step()
            ms.step()
        torch.cuda.synchronize()
        err = (ms.interior() - single.state[:, :, ms.r0:ms.r1]).abs().max().item()
        mass = ms.interior().double().sum().reshape(1)
        dist.all_reduce(mass)
        m0 = float(full.double().sum())
        to = ms.timed_out() if hasattr(ms, "timed_out") else False
        print(f"[rank {rank}] MaCE-XChan slab ({cls.__name__}): max |slab - torus| = {err:.2e}; mass {float(mass.item()):.4f} vs {m0:.4f}; "
              f"timed out: {to}", flush=True)
        ok &= err <= 1e-5 and abs(float(mass.item()) - m0) <= 1e-5 * m0 and not to
        if hasattr(ms, "close"):
            ms.close()
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("MULTI_GPU_CHECK", "PASS" if int(flag.item()) else "FAIL", flush=True)
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
