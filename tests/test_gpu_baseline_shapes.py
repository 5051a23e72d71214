"""Oracle parity ON THE BASELINE.json SHAPES -- the kernel instantiations behind every bench number.

Small worlds (1024^2) are compared with the oracle directly (lenia.py:430-469, macelenia.py:89-120,
mace_lenia_xchan.py:68-76, nca.py:239-263).  Worlds the oracle cannot hold (4096^2 C=3, 16384^2, GoL 65536^2) use the
periodic-tiling oracle of SURVEY.md 8(e): a torus built as PxQ copies of a small torus stays tiled under a periodic CA, so
every tile of the big result must equal the oracle's step of the small torus.  Each test asserts, through
b200ca_debug_variants(), WHICH kernel instantiation the shape dispatched to, so that the coverage cannot drift away from
what bench.py times (tests/variants.py lists the expected names in one place).
"""
import numpy as np
import pytest
import torch

from oracle import pyca_oracle as orc
from tests.util import load_npz, t
from tests.variants import EXPECT

pytestmark = pytest.mark.gpu
TOL = 1e-5


def params_dict(p):
    d = {k: t(p[k]) for k in ["mu", "sigma", "beta", "mu_k", "sigma_k", "weights"]}
    d["k_size"] = int(p["k_size"])
    return d


def lenia_c(C):
    """demo_lenia/abominable_mongolic sliced to C channels, as bench.py does (BASELINE configs[1])."""
    import pyca_b200

    d = params_dict(load_npz("params_lenia_abominable_mongolic.npz"))
    if C != 3:
        d = {k: (v[:, :C, :C].clone() if isinstance(v, torch.Tensor) else v) for k, v in d.items()}
    P = pyca_b200.LeniaParams(param_dict=d, channels=C)
    pd = P.param_dict
    K = orc.lenia_kernel(pd["beta"].cpu(), pd["mu_k"].cpu(), pd["sigma_k"].cpu(), 31)
    return P, K, pd["mu"].cpu(), pd["sigma"].cpu(), pd["weights"].cpu()


def xchan_params():
    import pyca_b200

    d = params_dict(load_npz("params_xchan_arborical_asbestosis.npz"))
    P = pyca_b200.LeniaParams(param_dict=d)
    pd = P.param_dict
    K = orc.lenia_kernel(pd["beta"].cpu(), pd["mu_k"].cpu(), pd["sigma_k"].cpu(), 31)
    return P, K, pd["mu"].cpu(), pd["sigma"].cpu(), pd["weights"].cpu()


def launched():
    from pyca_b200 import _lib

    return _lib.variants(reset=True)


def assert_variants(got, key):
    want = EXPECT[key]
    for w in want:
        assert any(g.startswith(w) for g in got), f"{key}: expected a launch of {w}, the step launched {got}"


def tiles_err(big, ref_tile):
    """max |tile - ref| over all PxQ tiles of `big` (..., P*h, Q*w), on the device."""
    h, w = ref_tile.shape[-2:]
    P, Q = big.shape[-2] // h, big.shape[-1] // w
    lead = big.shape[:-2]
    v = big.reshape(*lead, P, h, Q, w)
    r = ref_tile.to(big.device).reshape(*lead, 1, h, 1, w)
    return (v - r).abs().max().item()


@pytest.mark.parametrize("path", ["fft", "fft_rows", "direct"])
@pytest.mark.parametrize("C", [1, 3])
def test_lenia_1k_vs_oracle(C, path):
    """BASELINE configs[1] (C=1) and the unsliced C=3 case, 1024x1024, one step, both convolution paths."""
    import pyca_b200

    P, K, mu, sg, w = lenia_c(C)
    g = torch.Generator().manual_seed(C)
    s0 = torch.rand(1, C, 1024, 1024, generator=g)
    a = pyca_b200.Lenia((1024, 1024), num_channels=C, params=P, state_init=s0, device="cuda", conv_path=path)
    launched()
    a.step()
    got = launched()
    assert_variants(got, f"lenia{C}_1k_{path}")
    ref = orc.lenia_step(s0, K, mu, sg, w, 0.1)
    err = (a.state.cpu() - ref).abs().max().item()
    assert err <= TOL, f"C={C} {path}: {err:.2e}"
    # second step from the kernel's own output (periodic frame of the output is what the next step reads)
    a.step()
    ref2 = orc.lenia_step(ref, K, mu, sg, w, 0.1)
    err2 = (a.state.cpu() - ref2).abs().max().item()
    assert err2 <= 3 * TOL, f"C={C} {path} step 2: {err2:.2e}"


def test_xchan_1k_vs_oracle():
    import pyca_b200

    P, K, mu, sg, w = xchan_params()
    g = torch.Generator().manual_seed(3)
    s0 = torch.rand(1, 3, 1024, 1024, generator=g)
    a = pyca_b200.MaCELeniaXChan((1024, 1024), params=P, state_init=s0, device="cuda")
    m0 = a.total_mass().cpu()
    launched()
    a.step()
    got = launched()
    assert_variants(got, "xchan_1k")
    ref = orc.xchan_step(s0, K, mu, sg, w, 6.0, 0.03)
    err = (a.state.cpu() - ref).abs().max().item()
    assert err <= TOL * max(1.0, float(ref.max())), f"{err:.2e}"
    assert torch.allclose(m0.sum(), a.total_mass().cpu().sum(), rtol=1e-6)


@pytest.mark.parametrize("path", ["auto", "fft_rows"])
def test_xchan_4k_periodic_tiling(path):
    """BASELINE configs[2]: 4096x4096 C=3 MaCE-XChan as 4x4 copies of a 1024^2 torus -- on the path `auto` picks (the marching
    kernel, three source launches) and on the two-kernel hybrid (the big-world instantiations of its fused kernel)."""
    import pyca_b200

    P, K, mu, sg, w = xchan_params()
    g = torch.Generator().manual_seed(4)
    s0 = torch.rand(1, 3, 1024, 1024, generator=g)
    big = s0.cuda().repeat(1, 1, 4, 4)
    a = pyca_b200.MaCELeniaXChan((4096, 4096), params=P, state_init=big, device="cuda", conv_path=path)
    m0 = float(a.total_mass().sum())
    launched()
    a.step()
    got = launched()
    assert_variants(got, "xchan_4k" if path == "auto" else "xchan_4k_hybrid")
    ref = orc.xchan_step(s0, K, mu, sg, w, 6.0, 0.03)
    err = tiles_err(a.state, ref)
    assert err <= TOL * max(1.0, float(ref.max())), f"{err:.2e}"
    assert abs(float(a.total_mass().sum()) - m0) <= 1e-6 * m0
    assert tiles_err(a.Aff, orc.lenia_affinity(s0, K, mu, sg, w)) <= TOL


def test_lenia_overlap_save_4k_periodic_tiling():
    """4096 x 3976 (five overlap-save segments, partial last one): 8 x 8 copies of a 512 x 497 torus."""
    import pyca_b200

    P, K, mu, sg, w = lenia_c(1)
    g = torch.Generator().manual_seed(5)
    s0 = torch.rand(1, 1, 512, 497, generator=g)
    big = s0.cuda().repeat(1, 1, 8, 8)
    a = pyca_b200.Lenia((4096, 3976), num_channels=1, params=P, state_init=big, device="cuda")
    launched()
    a.step()
    got = launched()
    assert_variants(got, "lenia_seg_4k")
    ref = orc.lenia_step(s0, K, mu, sg, w, 0.1)
    err = tiles_err(a.state, ref)
    assert err <= TOL, f"{err:.2e}"


def test_lenia_16k_periodic_tiling():
    """BASELINE configs[4]: 16384x16384 C=1 as 16x16 copies of a 1024^2 torus (two steps)."""
    import pyca_b200

    P, K, mu, sg, w = lenia_c(1)
    g = torch.Generator().manual_seed(6)
    s0 = torch.rand(1, 1, 1024, 1024, generator=g)
    big = s0.cuda().repeat(1, 1, 16, 16)
    a = pyca_b200.Lenia((16384, 16384), num_channels=1, params=P, state_init=big, device="cuda")
    del big
    launched()
    a.step()
    got = launched()
    assert_variants(got, "lenia_16k")
    ref = orc.lenia_step(s0, K, mu, sg, w, 0.1)
    err = tiles_err(a.state, ref)
    assert err <= TOL, f"{err:.2e}"
    a.step()
    ref2 = orc.lenia_step(ref, K, mu, sg, w, 0.1)
    err2 = tiles_err(a.state, ref2)
    assert err2 <= 3 * TOL, f"step 2: {err2:.2e}"


def test_gol_64k_periodic_tiling():
    """BASELINE configs[4]: 65536x65536 Game of Life as 32x32 copies of a 2048^2 torus, bit-exact over 6 generations."""
    import pyca_b200

    rng = np.random.default_rng(64)
    h = w = 2048
    P = 32
    steps = 6
    small = (rng.random((h, w)) < 0.5).astype(np.uint8)
    ref = orc.gol_run(small, steps)
    sp = pyca_b200.PackedWorld(h, w, "cuda")
    sp.pack_from(torch.from_numpy(small).cuda())
    words = sp.bits[:, :sp.wpr]
    pw = pyca_b200.PackedWorld(h * P, w * P, "cuda")
    assert pw.wpr == sp.wpr * P
    pw.bits[:, :pw.wpr] = words.repeat(P, P)
    assert pw.population() == int(small.sum()) * P * P
    launched()
    pw.step(0b1100, 0b1000, steps)
    got = launched()
    assert_variants(got, "gol_64k")
    rp = pyca_b200.PackedWorld(h, w, "cuda")
    rp.pack_from(torch.from_numpy(ref.astype(np.uint8)).cuda())
    rw = rp.bits[:, :rp.wpr]
    tiles = pw.bits[:, :pw.wpr].reshape(P, h, P, rp.wpr)
    assert bool((tiles == rw[None, :, None, :]).all())
    assert pw.population() == int(ref.sum()) * P * P


@pytest.mark.parametrize("impl", [1, 0])
def test_nca_256_batch_vs_oracle(impl):
    """BASELINE configs[3]: 256 worlds of 128x128 in one launch; worlds 0, 97 and 255 against the oracle."""
    import pyca_b200

    fx = load_npz("nca_betta.npz")
    wts = pyca_b200.NCAWeights(t(fx["W1"]), t(fx["b1"]), t(fx["W2"]), t(fx["b2"]), "cuda")
    g = torch.Generator().manual_seed(256)
    s0 = torch.rand(256, 16, 128, 128, generator=g)
    s0[:, 3] *= (torch.rand(256, 128, 128, generator=g) < 0.6)
    mask = torch.rand(256, 1, 128, 128, generator=g) <= 0.5
    launched()
    out = pyca_b200.nca_step(s0.cuda(), wts, update_mask=mask.cuda(), impl=impl)
    got = launched()
    assert_variants(got, f"nca_256_impl{impl}")
    pick = [0, 97, 255]
    ref = orc.nca_step(s0[pick], t(fx["W1"]), t(fx["b1"]), t(fx["W2"]), t(fx["b2"]), mask[pick])
    o = out[pick].cpu()
    err = (o - ref).abs().max().item()
    assert err <= TOL, f"impl {impl}: {err:.2e}"
    assert torch.equal(o == 0, ref == 0)
