"""Spawns the multi-rank parity check when at least 2 GPUs are visible."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_slab_equals_torus_on_2_gpus():
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tests", "multi_gpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "MULTI_GPU_CHECK PASS" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


def test_single_rank_slab_paths():
    """world == 1 degenerates to local ghost copies: same result as the torus kernels."""
    import pyca_b200
    from pyca_b200.slab import GolSlabCUDA

    slab = GolSlabCUDA(300, 1000, 0, 1, ghost=5, device="cuda")
    slab.random_fill(seed=3)
    slab.run(17)
    ref = pyca_b200.PackedWorld(300, 1000, "cuda")
    ref.random_fill(seed=3)
    ref.step(0b1100, 0b1000, 17)
    assert torch.equal(slab.owned()[:, :ref.wpr], ref.bits[:, :ref.wpr])


def _local_exchange(slabs, pick):
    """The ring halo exchange of pyca_b200.slab done by direct copies between slab objects that live on ONE device: rank r's
    top frame rows <- last owned rows of rank r-1, bottom frame rows <- first owned rows of rank r+1 (full framed rows)."""
    n = len(slabs)
    for r, s in enumerate(slabs):
        up, down = slabs[(r - 1) % n], slabs[(r + 1) % n]
        t, tu, td = pick(s), pick(up), pick(down)
        F = s.F
        t[:, :, 0:F] = tu[:, :, up.hl:up.hl + F]
        t[:, :, s.hl + F:s.hl + 2 * F] = td[:, :, F:2 * F]


@pytest.mark.parametrize("alpha,path", [(None, "fft"), (0.03, "fft"), (0.03, "direct")])
def test_mace_slabs_equal_torus_single_device(alpha, path):
    """MaceSlabCUDA (two exchanges per step: affinity rows, then state rows) with three slabs emulated on one GPU -- the
    exchanges are plain copies between the slab objects -- against the single-torus MaCE / XChan step."""
    import pyca_b200
    from pyca_b200.slab import MaceSlabCUDA
    from tests.util import load_npz, t

    p = load_npz("params_xchan_arborical_asbestosis.npz")
    d = {k: t(p[k]) for k in ["mu", "sigma", "beta", "mu_k", "sigma_k", "weights"]}
    d["k_size"] = int(p["k_size"])
    params = pyca_b200.LeniaParams(param_dict=d)
    world, H, W = 3, 96 * 3, 160
    g = torch.Generator().manual_seed(17)
    full = 1.5 * torch.rand(1, 3, H, W, generator=g)
    cls = pyca_b200.MaCELenia if alpha is None else pyca_b200.MaCELeniaXChan
    ref = cls((H, W), params=params, state_init=full, device="cuda", conv_path=path)
    ref.b = 6.0
    if alpha is not None:
        ref.alpha = alpha
    slabs = [MaceSlabCUDA(params, 3, H, W, r, world, device="cuda", beta=6.0, alpha=alpha, conv_path=path) for r in range(world)]
    for s in slabs:
        s.world_emulated = True
        s.interior().copy_(full[:, :, s.r0:s.r1].cuda())
        pyca_b200._lib.call("b200ca_frame_refresh", pyca_b200._lib.ptr(s.buf), 1, 3, s.hl, W, 0,
                            pyca_b200._lib.current_stream(torch.device("cuda")))
    _local_exchange(slabs, lambda s: s.buf)
    for _ in range(3):
        ref.step()
        for s in slabs:
            s.affinity_phase()
        _local_exchange(slabs, lambda s: s.aff)
        for s in slabs:
            s.redistribute_phase()
        _local_exchange(slabs, lambda s: s.buf)
    for s in slabs:
        err = (s.interior() - ref.state[:, :, s.r0:s.r1]).abs().max().item()
        assert err <= 5e-6 * max(1.0, float(ref.state.max())), f"rank {s.rank}: {err}"
    mass = sum(float(s.interior().double().sum()) for s in slabs)
    assert abs(mass - float(full.double().sum())) <= 1e-5 * float(full.double().sum())
