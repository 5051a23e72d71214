"""GPU parity of the Lenia-family kernels (through the C ABI / Automaton drop-ins) against the golden
vectors produced by the unmodified reference and against the oracle.

Tolerance (BASELINE.json north_star): single step max-abs 1e-5 in fp32 on states in [0,1]; states with
values >> 1 (real saved MaCE states reach 247) use the same bound relative to the state maximum.
Multi-step runs compare mass / statistics only (chaotic dynamics).  MaCE must conserve mass to fp32 round-off.
"""
import numpy as np
import pytest
import torch

from oracle import pyca_oracle as orc
from tests.util import family_case, golden_names, load_npz, t

pytestmark = pytest.mark.gpu
TOL = 1e-5


def params_dict(p, C=3):
    d = {k: t(p[k]) for k in ["mu", "sigma", "beta", "mu_k", "sigma_k", "weights"]}
    d["k_size"] = int(p["k_size"])
    return d


def sliced(d, c):
    return {k: (v[:, :c, :c].clone() if isinstance(v, torch.Tensor) else v) for k, v in d.items()}


PATHS = ["direct", "fft", "fft_rows"]   # direct conv, marching FFT kernel (default), row-FFT + fused FIR/inverse hybrid


@pytest.mark.parametrize("path", PATHS)
@pytest.mark.parametrize("name", golden_names("lenia"))
def test_lenia_golden(name, path):
    import pyca_b200

    fx, p = family_case("lenia", name)
    s0 = t(fx["state_in"])
    a = pyca_b200.Lenia(tuple(s0.shape[2:]), dt=float(fx["dt"]), params=params_dict(p), state_init=s0, device="cuda",
                        conv_path=path)
    assert a.use_fft == (path != "direct") and a.conv_algo == {"direct": "direct", "fft": "march", "fft_rows": "fft_rows"}[path]
    assert torch.allclose(a.k.cpu(), t(p["K"]), atol=1e-8), "host kernel construction differs from the reference's"
    a.step()
    err = (a.state.cpu() - t(fx["state_out"])).abs().max().item()
    assert err <= TOL, f"{name}: max-abs {err:.2e}"
    assert a.frames == 1
    # single-channel slice of the same file (BASELINE configs[1])
    a1 = pyca_b200.Lenia(tuple(s0.shape[2:]), dt=float(fx["dt"]), num_channels=1, params=sliced(params_dict(p), 1),
                         state_init=s0[:, :1], device="cuda", conv_path=path)
    assert torch.allclose(a1.k.cpu(), t(fx["c1_K"]), atol=1e-8)
    a1.step()
    err1 = (a1.state.cpu() - t(fx["c1_out"])).abs().max().item()
    assert err1 <= TOL, f"{name} C=1: max-abs {err1:.2e}"


@pytest.mark.parametrize("path", PATHS)
@pytest.mark.parametrize("name", golden_names("mace"))
def test_mace_golden(name, path):
    import pyca_b200

    fx, p = family_case("mace", name)
    s0 = t(fx["state_in"])
    a = pyca_b200.MaCELenia(tuple(s0.shape[2:]), params=params_dict(p), state_init=s0, device="cuda", conv_path=path)
    assert a.b == float(fx["b"])
    m0 = a.total_mass().cpu()
    a.step()
    aff_err = (a.Aff.cpu() - t(fx["aff"])).abs().max().item()
    err = (a.state.cpu() - t(fx["state_out"])).abs().max().item()
    scale = max(1.0, float(t(fx["state_out"]).max()))
    assert err <= TOL * scale, f"{name}: state max-abs {err:.2e} (aff {aff_err:.2e})"
    m1 = a.total_mass().cpu()
    assert torch.allclose(m0, m1, rtol=1e-6), "mass not conserved"


@pytest.mark.parametrize("path", PATHS)
@pytest.mark.parametrize("name", golden_names("xchan"))
def test_xchan_golden(name, path):
    import pyca_b200

    fx, p = family_case("xchan", name)
    s0 = t(fx["state_in"])
    a = pyca_b200.MaCELeniaXChan(tuple(s0.shape[2:]), params=params_dict(p), state_init=s0, device="cuda", conv_path=path)
    a.alpha = float(fx["alpha_used"])
    assert a.b == float(fx["b"])
    m0 = a.total_mass().cpu()
    a.step()
    ref = t(fx["state_out"])
    scale = max(1.0, float(ref.max()))
    err = (a.state.cpu() - ref).abs().max().item()
    assert err <= TOL * scale, f"{name}: max-abs {err:.2e} (scale {scale:.1f})"
    assert torch.allclose(m0, a.total_mass().cpu(), rtol=1e-6)


def _oracle_params(name="abominable_mongolic", tag="lenia"):
    p = load_npz(f"params_{tag}_{name}.npz")
    return params_dict(p), t(p["K"]), t(p["mu_s"]), t(p["sigma_s"]), t(p["weights_s"])


@pytest.mark.parametrize("path", PATHS)
@pytest.mark.parametrize("shape", [(32, 32), (50, 70), (64, 64), (65, 129), (100, 47), (31 + 1, 200), (256, 256), (34, 500),
                                   (96, 1024), (40, 1100), (48, 224), (49, 225), (40, 448), (33, 449), (131, 300), (600, 260),
                                   (1030, 96)])
def test_lenia_odd_shapes_vs_oracle(shape, path):
    """Tile overhang, partial strips, worlds narrower than a tile; FFT path: periodic (W = 256, 1024) and overlap-save
    segments (one, several, partial last)."""
    import pyca_b200

    if path == "fft_rows" and pyca_b200.lib().b200ca_lenia_fft_length(*shape) == 0:
        pytest.skip("shape not supported by the FFT path")
    d, K, mu, sg, w = _oracle_params()
    g = torch.Generator().manual_seed(shape[0] * 1000 + shape[1])
    s0 = torch.rand(1, 3, *shape, generator=g)
    a = pyca_b200.Lenia(shape, params=d, state_init=s0, device="cuda", conv_path=path)
    a.step()
    ref = orc.lenia_step(s0, K, mu, sg, w, 0.1)
    err = (a.state.cpu() - ref).abs().max().item()
    assert err <= TOL, f"{shape}: {err:.2e}"


def test_lenia_batched_params_and_state_contract():
    import pyca_b200

    d0, K0, mu0, sg0, w0 = _oracle_params("abominable_mongolic")
    d1, K1, mu1, sg1, w1 = _oracle_params("bedfast_laplace")
    d = {k: (torch.cat([d0[k], d1[k]], 0) if isinstance(d0[k], torch.Tensor) else d0[k]) for k in d0}
    g = torch.Generator().manual_seed(5)
    s0 = torch.rand(2, 3, 64, 96, generator=g)
    a = pyca_b200.Lenia((2, 64, 96), params=d, state_init=s0, device="cuda")
    a.step()
    ref = torch.cat([orc.lenia_step(s0[:1], K0, mu0, sg0, w0), orc.lenia_step(s0[1:], K1, mu1, sg1, w1)], 0)
    assert (a.state.cpu() - ref).abs().max().item() <= TOL
    # in-place edit of the public state between steps (brush, lenia.py:681-690), including frame cells
    st = a.state
    st[:, :, :5, :] = 0.25
    st[:, 1, 20:30, 90:] += 0.1
    edited = ref.clone()
    edited[:, :, :5, :] = 0.25
    edited[:, 1, 20:30, 90:] += 0.1
    a.step()
    ref2 = torch.cat([orc.lenia_step(edited[:1], K0, mu0, sg0, w0), orc.lenia_step(edited[1:], K1, mu1, sg1, w1)], 0)
    assert (a.state.cpu() - ref2).abs().max().item() <= TOL
    # re-assignment
    a.state = s0.cuda()
    a.step()
    assert (a.state.cpu() - ref).abs().max().item() <= TOL
    a.draw() if a.batch == 1 else None


def test_small_kernel_size_embedding():
    import pyca_b200

    d, *_ = _oracle_params()
    d = dict(d)
    d["k_size"] = 13
    q = orc.sanitize_params({k: v for k, v in d.items() if isinstance(v, torch.Tensor)})
    K = orc.lenia_kernel(q["beta"], q["mu_k"], q["sigma_k"], 13)
    g = torch.Generator().manual_seed(9)
    s0 = torch.rand(1, 3, 40, 72, generator=g)
    a = pyca_b200.Lenia((40, 72), params=d, state_init=s0, device="cuda")
    a.step()
    ref = orc.lenia_step(s0, K, q["mu"], q["sigma"], q["weights"])
    assert (a.state.cpu() - ref).abs().max().item() <= TOL


def test_lenia_multistep_statistics():
    """Chaotic dynamics: compare mass and moments after 20 steps, not cells."""
    import pyca_b200

    d, K, mu, sg, w = _oracle_params()
    g = torch.Generator().manual_seed(21)
    s0 = torch.rand(1, 3, 128, 128, generator=g)
    a = pyca_b200.Lenia((128, 128), params=d, state_init=s0, device="cuda")
    ref = s0.clone()
    Kf = orc.kernel_fft(K, 128, 128)
    for _ in range(20):
        a.step()
        ref = orc.lenia_step(ref, K, mu, sg, w, 0.1, Kf)
    got = a.state.cpu()
    assert got.min() >= 0 and got.max() <= 1
    assert torch.allclose(a.mass().cpu(), ref.mean(dim=(-1, -2)), atol=2e-3)
    assert abs(got.std().item() - ref.std().item()) < 5e-3


@pytest.mark.parametrize("path", PATHS)
def test_vs_float64_noise_floor(path):
    """|cuda - ref64| must not exceed the reference's own fp32 rounding error |ref32 - ref64| by much
    (BASELINE.md section 6)."""
    import pyca_b200

    d, K, mu, sg, w = _oracle_params("arborical_asbestosis", "xchan")
    g = torch.Generator().manual_seed(33)
    s0 = torch.rand(1, 3, 128, 128, generator=g)
    a = pyca_b200.MaCELenia((128, 128), params=d, state_init=s0, device="cuda", conv_path=path)
    aff = a._compute_affinity().cpu()
    aff32 = orc.lenia_affinity(s0, K, mu, sg, w)
    aff64 = orc.lenia_affinity_f64(s0, K, mu, sg, w)
    e_cuda = (aff.double() - aff64).abs().max().item()
    e_ref = (aff32.double() - aff64).abs().max().item()
    print(f"[{path}] affinity error vs fp64: cuda {e_cuda:.2e}, reference fp32 {e_ref:.2e}")
    assert e_cuda <= max(4 * e_ref, 2e-6)


def test_xchan_realstate_relative():
    """Crop of a real saved MaCE-XChan state (values up to ~20): relative tolerance."""
    import pyca_b200

    fx = load_npz("xchan_realstate.npz")
    s0 = t(fx["state_in"])
    a = pyca_b200.MaCELeniaXChan((64, 64), params=params_dict(fx), state_init=s0, device="cuda")
    a.alpha = float(fx["alpha_used"])
    a.step()
    ref = t(fx["state_out"])
    err = (a.state.cpu() - ref).abs().max().item()
    assert err <= TOL * max(1.0, float(ref.max())), f"{err:.2e} vs max {float(ref.max()):.1f}"


def test_mace_mass_conservation_long_run():
    import pyca_b200

    d, *_ = _oracle_params("arborical_asbestosis", "xchan")
    g = torch.Generator().manual_seed(44)
    s0 = torch.rand(1, 3, 512, 512, generator=g)
    for cls in (pyca_b200.MaCELenia, pyca_b200.MaCELeniaXChan):
        a = cls((512, 512), params=d, state_init=s0, device="cuda")
        m0 = a.total_mass().item()
        for _ in range(25):
            a.step()
        m1 = a.total_mass().item()
        assert abs(m1 - m0) <= 2e-5 * m0, f"{cls.__name__}: mass {m0} -> {m1}"
        assert torch.isfinite(a.state).all() and a.state.min() >= 0


def test_periodic_tiling_identity():
    """A 2x2 tiling of a small torus must evolve as 4 copies of the small torus (exercises frame + tile seams)."""
    import pyca_b200

    d, *_ = _oracle_params("arborical_asbestosis", "xchan")
    g = torch.Generator().manual_seed(55)
    s0 = torch.rand(1, 3, 96, 160, generator=g)
    small = pyca_b200.MaCELeniaXChan((96, 160), params=d, state_init=s0, device="cuda")
    big = pyca_b200.MaCELeniaXChan((192, 320), params=d, state_init=s0.repeat(1, 1, 2, 2), device="cuda")
    for _ in range(3):
        small.step()
        big.step()
    sm = small.state
    bg = big.state
    for ty in range(2):
        for tx in range(2):
            tile = bg[:, :, ty * 96:(ty + 1) * 96, tx * 160:(tx + 1) * 160]
            assert (tile - sm).abs().max().item() <= 1e-5  # different tilings/segmentations round differently


def test_host_entry_point():
    import ctypes

    import pyca_b200
    from pyca_b200 import _lib

    d, K, mu, sg, w = _oracle_params()
    g = torch.Generator().manual_seed(66)
    s0 = torch.rand(1, 3, 64, 80, generator=g)
    a = pyca_b200.Lenia((64, 80), params=d, state_init=s0, device="cuda")  # prepares device params
    h_in = s0.numpy()
    h_out = np.empty_like(h_in)
    _lib.call("b200ca_lenia_run_host", h_in.ctypes.data_as(ctypes.c_void_p), h_out.ctypes.data_as(ctypes.c_void_p),
              _lib.ptr(a._kprep), _lib.ptr(a.mu), _lib.ptr(a.sigma), _lib.ptr(a.weights), 1, 3, 1, 64, 80, 0.1, 0, 0.0, 0.0, 2)
    ref = orc.lenia_step(orc.lenia_step(s0, K, mu, sg, w), K, mu, sg, w)
    assert (torch.from_numpy(h_out) - ref).abs().max().item() <= 2 * TOL


@pytest.mark.parametrize("path", PATHS)
def test_k_mult_2_vs_oracle(path):
    """k_mult > 1 (several kernels per source channel, lenia.py:77, index i*k_mult+m): synthetic parameters, since every
    demo_data file has k_mult = 1."""
    import pyca_b200

    torch.manual_seed(5)
    C, km, R = 2, 2, 3
    d = {"mu": 0.1 + 0.2 * torch.rand(1, C * km, C), "sigma": 0.03 + 0.1 * torch.rand(1, C * km, C),
         "beta": torch.rand(1, C * km, C, R), "mu_k": torch.rand(1, C * km, C, R), "sigma_k": 0.05 + 0.2 * torch.rand(1, C * km, C, R),
         "weights": torch.rand(1, C * km, C), "k_size": 31, "k_mult": km, "num_channels": C}
    q = orc.sanitize_params({k: v for k, v in d.items() if isinstance(v, torch.Tensor)})
    K = orc.lenia_kernel(q["beta"], q["mu_k"], q["sigma_k"], 31)
    g = torch.Generator().manual_seed(1)
    s0 = torch.rand(1, C, 64, 128, generator=g)
    a = pyca_b200.Lenia((64, 128), num_channels=C, params=dict(d), state_init=s0, device="cuda", conv_path=path)
    assert a.k_mult == km
    a.step()
    ref = orc.lenia_step(s0, K, q["mu"], q["sigma"], q["weights"], 0.1)
    assert (a.state.cpu() - ref).abs().max().item() <= TOL
    fx = load_npz("lenia_kmult2.npz")  # the same case run by the unmodified reference
    assert torch.equal(s0, t(fx["state_in"])) and (a.state.cpu() - t(fx["state_out"])).abs().max().item() <= TOL
    m = pyca_b200.MaCELeniaXChan((64, 128), num_channels=C, params=dict(d), state_init=s0, device="cuda", conv_path=path)
    m.step()
    refm = orc.xchan_step(s0, K, q["mu"], q["sigma"], q["weights"], 6.0, 0.03)
    assert (m.state.cpu() - refm).abs().max().item() <= TOL


@pytest.mark.parametrize("name", ["underweight_skagway", "inaesthetic_maupassant", "unsuspected_lanthanotidae", "lenia_synth"])
def test_garbi_growth_golden(name):
    """g_arbi Fourier growth functions: the three demo_data files that use them + a synthetic Lenia case,
    against the reference's own step (tests/golden/garbi_*.npz).  Built from the raw parameter file, so the k_arbi kernel
    construction, the coefficient tables and the two-pass global rescale are all on the tested path."""
    import pyca_b200
    from tests.util import garbi_case, raw_param_dict

    fx, garbi = garbi_case(name)
    s0 = t(fx["state_in"])
    C, H, W = s0.shape[1:]
    fam = int(fx["family"])
    params = pyca_b200.LeniaParams(param_dict=raw_param_dict(fx), channels=C)
    cls = [pyca_b200.Lenia, pyca_b200.MaCELenia, pyca_b200.MaCELeniaXChan][fam]
    kw = {"num_channels": C, "params": params, "state_init": s0, "device": "cuda"}
    a = cls((H, W), **kw)
    assert a.g_arbi and not a.use_fft
    assert (a.k.cpu() - t(fx["K"])).abs().max().item() <= 1e-6
    if fam == 2:
        a.alpha = float(fx["alpha_used"])
    a.step()
    # Steep Fourier growth functions amplify fp32 rounding of U (and MaCE's exp(b*Aff) amplifies it again): on
    # underweight_skagway the REFERENCE's own fp32 step is 2.3e-5 away from the same step in float64.  The bound is
    # therefore 1e-5 or 2.5x the reference's own distance to float64 on this input, whichever is larger.
    K64, w64 = t(fx["K"]).double(), t(fx["weights_s"]).double()
    g64 = (garbi[0].double(), garbi[1].double(), garbi[2], garbi[3])
    if fam == 0:
        o64 = orc.lenia_step(s0.double(), K64, None, None, w64, 0.1, garbi=g64)
    elif fam == 1:
        o64 = orc.mace_step(s0.double(), K64, None, None, w64, float(fx["b"]), garbi=g64)[0]
    else:
        o64 = orc.xchan_step(s0.double(), K64, None, None, w64, float(fx["b"]), float(fx["alpha_used"]), garbi=g64)
    ref = t(fx["state_out"])
    noise = (ref.double() - o64).abs().max().item()
    tol = max(1e-5 * max(1.0, float(ref.abs().max())), 2.5 * noise)
    assert (a.state.cpu() - ref).abs().max().item() <= tol
    assert (a.state.cpu().double() - o64).abs().max().item() <= tol
    if fam > 0:
        a64 = orc.lenia_affinity(s0.double(), K64, None, None, w64, garbi=g64)
        noise_a = (t(fx["aff"]).double() - a64).abs().max().item()
        assert (a.Aff.cpu() - t(fx["aff"])).abs().max().item() <= max(1e-5, 2.5 * noise_a)
        m0, m1 = float(s0.double().sum()), float(a.state.double().sum())
        assert abs(m1 - m0) <= 1e-6 * m0   # MaCE stays mass conserving whatever the growth function


def test_garbi_odd_shape_vs_oracle():
    """Global rescale statistics must ignore the overhang of edge tiles (world not a multiple of the 64-wide tiles)."""
    import pyca_b200
    from tests.util import garbi_case, raw_param_dict

    fx, garbi = garbi_case("lenia_synth")
    C = 2
    g = torch.Generator().manual_seed(12)
    s0 = torch.rand(1, C, 37, 83, generator=g)
    a = pyca_b200.Lenia((37, 83), num_channels=C, params=pyca_b200.LeniaParams(param_dict=raw_param_dict(fx), channels=C),
                        state_init=s0, device="cuda")
    a.step()
    ref = orc.lenia_step(s0, t(fx["K"]), None, None, t(fx["weights_s"]), 0.1, garbi=garbi)
    assert (a.state.cpu() - ref).abs().max().item() <= 1e-5


@pytest.mark.parametrize("tag", ["mace", "xchan"])
def test_mace_food_subsystem_golden(tag):
    """has_food=True steps: food sensing, decay, the respawn of food spots (host `random`, same seed as the
    reference run), consumption and -- for the cross-channel variant -- the mixing AFTER the food step; plus total_mass()
    and draw() with the food overlay.  Fixture: tests/golden/food.npz from the unmodified reference."""
    import random

    import pyca_b200

    fx = load_npz("food.npz")
    p = load_npz(f"params_{tag}_{str(fx[tag + '_file'])}.npz")
    d = {k: t(p[k]) for k in ["mu", "sigma", "beta", "mu_k", "sigma_k", "weights"]}
    d["k_size"] = int(p["k_size"])
    s0 = t(fx[f"{tag}_state0"])
    H, W = s0.shape[2:]
    cls = pyca_b200.MaCELenia if tag == "mace" else pyca_b200.MaCELeniaXChan
    a = cls((H, W), params=pyca_b200.LeniaParams(param_dict=d), state_init=s0, device="cuda")
    random.seed(1234)
    a.set_food(True)
    assert torch.equal(a.food_channel.cpu(), t(fx[f"{tag}_food0"]))   # same spot positions as the reference drew
    a.sense_food = tag == "xchan"
    a.cum_loss_mass = torch.full((1,), float(fx[f"{tag}_loss0"]), device="cuda")
    if tag == "xchan":
        a.alpha = float(fx["xchan_alpha"])
    random.seed(99)
    i = 1
    while f"{tag}_state{i}" in fx:
        a.step()
        assert (a.state.cpu() - t(fx[f"{tag}_state{i}"])).abs().max().item() <= 1e-5 * i
        assert (a.food_channel.cpu() - t(fx[f"{tag}_food{i}"])).abs().max().item() <= 1e-6
        assert (a.cum_loss_mass.cpu() - t(fx[f"{tag}_loss{i}"])).abs().max().item() <= 1e-4
        i += 1
    assert abs(float(a.total_mass()) - float(fx[f"{tag}_total_mass"])) <= 1e-5 * float(fx[f"{tag}_total_mass"])
    a.draw()
    assert (a._worldmap.cpu() - t(fx[f"{tag}_wm"])).abs().max().item() <= 1e-5 * i
    assert (a.worldmap.astype(np.int32) - fx[f"{tag}_frame"].astype(np.int32)).__abs__().max() <= 1   # uint8 truncation of ~1e-5 diffs


def test_constructors_without_params_or_state():
    """Lenia(size) as MainWindow builds it (no params, no state): random parameters with the reference's prior, a noise disc
    as initial state, and steps that stay finite and inside [0, 1]."""
    import pyca_b200

    torch.manual_seed(3)
    a = pyca_b200.Lenia((256, 320), device="cuda")      # disc radius 3 * 31 = 93 < distance of the corners
    assert tuple(a.state.shape) == (1, 3, 256, 320) and a.k_size == 31
    assert float(a.state[0, :, 0, 0].abs().max()) == 0.0 and float(a.state.max()) > 0.5   # disc in the centre, nothing outside
    for _ in range(3):
        a.step()
    s = a.state
    assert bool(torch.isfinite(s).all()) and float(s.min()) >= 0.0 and float(s.max()) <= 1.0
    m = pyca_b200.MaCELeniaXChan((256, 320), device="cuda")
    m0 = float(m.total_mass())
    m.step()
    assert abs(float(m.total_mass()) - m0) <= 1e-5 * m0


@pytest.mark.parametrize("path", PATHS)
@pytest.mark.parametrize("ks", [5, 7, 13, 21, 27])
def test_lenia_kernel_sizes_vs_oracle(ks, path):
    """k_size below 31 (Lenia.update_params(k_size_override), lenia.py:143-168): the direct kernel evaluates only the radius
    class of the kernel (3, 5, 7, 9, 15), the FFT paths embed it in their 31-tap footprint; all against the oracle."""
    import pyca_b200

    d, _, mu, sg, w = _oracle_params()
    d = dict(d)
    d["k_size"] = ks
    K = orc.lenia_kernel(d["beta"], d["mu_k"], d["sigma_k"], ks)
    g = torch.Generator().manual_seed(ks)
    shape = (96, 300)
    s0 = torch.rand(1, 3, *shape, generator=g)
    a = pyca_b200.Lenia(shape, params=d, state_init=s0, device="cuda", conv_path=path)
    assert a.k.shape[-1] == ks
    from pyca_b200 import _lib
    _lib.variants(reset=True)
    a.step()
    got = _lib.variants(reset=True)
    if path == "direct":
        rk = 3 if ks <= 7 else 5 if ks <= 11 else 7 if ks <= 15 else 9 if ks <= 19 else 15
        assert any(f"R={rk}>" in v for v in got), got
    ref = orc.lenia_step(s0, K, mu, sg, w, 0.1)
    err = (a.state.cpu() - ref).abs().max().item()
    assert err <= TOL, f"ks={ks} {path}: {err:.2e}"


@pytest.mark.parametrize("C,B,shape", [(1, 5, (1000, 500)), (1, 3, (1024, 1024)), (3, 4, (600, 460)), (1, 2, (2056, 230)), (1, 9, (40, 300))])
def test_march_balanced_runs_cross_columns(C, B, shape):
    """The marching kernel's balanced launch cuts the (world, strip, 16-row unit) space into equal runs, one per CTA
    (csrc/lenia_march.cu): a run starts and ends anywhere in a column of bands and may cross into the next strip or the next
    world.  Batches of worlds with their own parameters (lenia.py:19-29) and heights that are not multiples of 16, so that the
    runs hit every kind of seam; two steps, the second from the kernel's own framed output."""
    import pyca_b200

    sets = []
    for name in ("abominable_mongolic", "bedfast_laplace"):
        dd = params_dict(load_npz(f"params_lenia_{name}.npz"))
        if C != 3:
            dd = sliced(dd, C)
        pd = pyca_b200.LeniaParams(param_dict=dd, channels=C).param_dict   # sanitised as the constructor does (leniaparams.py)
        pd = {k: (v.cpu() if isinstance(v, torch.Tensor) else v) for k, v in pd.items()}
        sets.append((pd, orc.lenia_kernel(pd["beta"], pd["mu_k"], pd["sigma_k"], 31), pd["mu"], pd["sigma"], pd["weights"]))
    pick = [sets[i % 2] for i in range(B)]
    d = {k: torch.cat([p[0][k] for p in pick], 0) for k in ["mu", "sigma", "beta", "mu_k", "sigma_k", "weights"]}
    d["k_size"] = 31
    g = torch.Generator().manual_seed(B * 100 + C)
    s0 = torch.rand(B, C, *shape, generator=g)
    a = pyca_b200.Lenia((B, *shape), num_channels=C, params=d, state_init=s0, device="cuda", conv_path="fft")
    pyca_b200._lib.variants(reset=True)
    a.step()
    got = pyca_b200._lib.variants(reset=True)
    assert any(v.startswith(f"lenia_march<{'single' if C == 1 else 'multi'},balanced:") for v in got), got

    def ref_step(s):
        return torch.cat([orc.lenia_step(s[i:i + 1], p[1], p[2], p[3], p[4], 0.1) for i, p in enumerate(pick)], 0)

    ref = ref_step(s0)
    err = (a.state.cpu() - ref).abs().max().item()
    assert err <= TOL, f"step 1: {err:.2e}"
    a.step()
    err2 = (a.state.cpu() - ref_step(ref)).abs().max().item()
    assert err2 <= 3 * TOL, f"step 2: {err2:.2e}"
