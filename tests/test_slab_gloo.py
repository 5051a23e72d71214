"""N>1 host logic on CPU: world_size-2 (and 3) gloo runs of the row-slab algorithms.

The slab bookkeeping and the halo exchange in pyca_b200/slab.py are device-agnostic; here the local
compute is the oracle (numpy / torch CPU) so the test checks that partition + ghost-row exchange +
multi-generation ghost zones reproduce the single-torus evolution exactly."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _pack(w):  # (H,W) 0/1 -> (H, stride) int32 words, bit x&31 of word x>>5
    H, W = w.shape
    wpr = (W + 31) // 32
    stride = (wpr + 3) // 4 * 4
    out = np.zeros((H, stride), dtype=np.uint32)
    for x in range(W):
        out[:, x >> 5] |= (w[:, x].astype(np.uint32) << np.uint32(x & 31))
    return out.view(np.int32)


def _unpack(bits, W):
    b = bits.view(np.uint32)
    out = np.zeros((b.shape[0], W), dtype=np.int32)
    for x in range(W):
        out[:, x] = (b[:, x >> 5] >> np.uint32(x & 31)) & 1
    return out


def _gol_worker(rank, world, port, H, W, ghost, nsteps, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import pyca_oracle as orc
    from pyca_b200.slab import GolSlab

    rng = np.random.default_rng(1234)
    full = (rng.random((H, W)) < 0.45).astype(np.int32)
    slab = GolSlab(H, W, rank, world, ghost, "cpu")
    slab.owned().copy_(torch.from_numpy(_pack(full[slab.r0:slab.r1])))

    def step_fn(src, dst, rows, n):
        # open-rows semantics of b200ca_gol_step(vert_open=1): rows outside the buffer are dead, columns wrap
        w = _unpack(src.numpy(), W)
        for _ in range(n):
            pad = np.zeros((rows + 2, W), dtype=np.int32)
            pad[1:-1] = w
            w = orc.gol_step(pad)[1:-1]
        dst.copy_(torch.from_numpy(_pack(w)))
        return dst

    slab.advance(step_fn, nsteps)
    got = _unpack(slab.owned().numpy(), W)
    ref = orc.gol_run(full, nsteps)[slab.r0:slab.r1]
    ret[rank] = bool(np.array_equal(got, ref))
    dist.destroy_process_group()


def _lenia_worker(rank, world, port, H, W, nsteps, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import pyca_oracle as orc
    from pyca_b200.slab import LeniaSlab
    from tests.util import load_npz, t

    p = load_npz("params_lenia_abominable_mongolic.npz")
    K, mu, sg, wt = t(p["K"]), t(p["mu_s"]), t(p["sigma_s"]), t(p["weights_s"])
    g = torch.Generator().manual_seed(7)
    full = torch.rand(1, 3, H, W, generator=g)
    F = 16
    pitch = (W + 2 * F + 3) // 4 * 4
    slab = LeniaSlab(1, 3, H, W, rank, world, F, pitch, "cpu")
    slab.interior().copy_(full[:, :, slab.r0:slab.r1])

    def fill_x_frame(tn):  # what the CUDA kernel does for its own rows: wrap columns into the frame
        rows = slice(F, F + slab.hl)
        tn[:, :, rows, :F] = tn[:, :, rows, W:W + F]
        tn[:, :, rows, W + F:W + 2 * F] = tn[:, :, rows, F:2 * F]

    fill_x_frame(slab.buf)
    slab.exchange()
    for _ in range(nsteps):
        # local operator on the slab + 15 ghost rows: zero-padded direct correlation == what the kernel computes
        src = slab.buf[:, :, :, :W + 2 * F]
        # explicit per-pair correlation (clear, small sizes)
        dx = torch.zeros(1, 3, slab.hl, W)
        for j in range(3):
            for i in range(3):
                inp = src[:, i:i + 1, 1:slab.rows - 1, 1:W + 2 * F - 1]  # rows/cols -15..+15 around the interior
                u = torch.nn.functional.conv2d(inp, K[0, i, j][None, None])
                gth = 2 * torch.exp(-((u - mu[0, i, j]) ** 2 / sg[0, i, j] ** 2) / 2) - 1
                dx[:, j:j + 1] += gth * wt[0, i, j]
        out = torch.clamp(slab.interior() + 0.1 * dx, 0, 1)
        slab.interior(slab.other).copy_(out)
        fill_x_frame(slab.other)
        slab.cur ^= 1
        slab.exchange()
    ref = full
    for _ in range(nsteps):
        ref = orc.lenia_step(ref, K, mu, sg, wt, 0.1)
    err = (slab.interior() - ref[:, :, slab.r0:slab.r1]).abs().max().item()
    ret[rank] = err
    dist.destroy_process_group()


def _spawn(fn, world, args):
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=fn, args=(r, world, port) + args + (ret,)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
        assert p.exitcode == 0, f"worker exit code {p.exitcode}"
    return dict(ret)


@pytest.mark.parametrize("world,ghost,nsteps", [(2, 1, 5), (2, 4, 11), (3, 3, 7)])
def test_gol_slab_matches_torus(world, ghost, nsteps):
    ret = _spawn(_gol_worker, world, (24 * world, 70, ghost, nsteps))
    assert all(ret[r] for r in range(world)), ret


@pytest.mark.parametrize("world", [2])
def test_lenia_slab_matches_torus(world):
    ret = _spawn(_lenia_worker, world, (64, 48, 2))
    assert all(ret[r] <= 2e-6 for r in range(world)), ret


def test_partition_helpers():
    from pyca_b200.slab import ring_neighbours, slab_rows

    assert slab_rows(64, 4, 1) == (16, 32)
    assert ring_neighbours(0, 4) == (3, 1) and ring_neighbours(3, 4) == (2, 0)
    with pytest.raises(ValueError):
        slab_rows(10, 4, 0)
