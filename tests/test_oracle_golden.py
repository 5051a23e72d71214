"""Pins oracle/ against the fixtures produced by the UNMODIFIED reference (tests/golden/generate.py)."""
import numpy as np
import pytest
import torch

from oracle import pyca_oracle as orc
from tests.util import family_case, golden_names, load_npz, t

TOL = 1e-6  # oracle uses the reference's own torch primitives in the same order


def test_gol_bit_exact():
    z = load_npz("gol.npz")
    for ci in range(int(z["ncases"])):
        H, W, s, b, steps = [int(v) for v in z[f"c{ci}_meta"]]
        out = orc.gol_run(z[f"c{ci}_w0"], steps, s, b)
        assert out.shape == (H, W)
        assert np.array_equal(out, z[f"c{ci}_w1"].astype(np.int32)), f"case {ci}"


def test_rule_mask():
    assert orc.rule_mask("23") == 12 and orc.rule_mask("3") == 8 and orc.rule_mask("") == 0


@pytest.mark.parametrize("tag", ["lenia", "mace", "xchan"])
def test_kernel_and_sanitize(tag):
    for name in golden_names(f"params_{tag}"):
        p = load_npz(f"params_{tag}_{name}.npz")
        raw = {k: t(p[k]) for k in ["mu", "sigma", "beta", "mu_k", "sigma_k", "weights"]}
        q = orc.sanitize_params(raw)
        assert torch.equal(q["mu"], t(p["mu_s"]))
        assert torch.equal(q["sigma"], t(p["sigma_s"]))
        assert torch.allclose(q["weights"], t(p["weights_s"]), atol=0, rtol=0)
        K = orc.lenia_kernel(q["beta"], q["mu_k"], q["sigma_k"], int(p["k_size"]))
        assert torch.allclose(K, t(p["K"]), atol=1e-9, rtol=0)


@pytest.mark.parametrize("name", golden_names("lenia"))
def test_lenia_step(name):
    fx, p = family_case("lenia", name)
    out = orc.lenia_step(t(fx["state_in"]), t(p["K"]), t(p["mu_s"]), t(p["sigma_s"]), t(p["weights_s"]), float(fx["dt"]))
    assert (out - t(fx["state_out"])).abs().max().item() <= TOL
    out1 = orc.lenia_step(t(fx["state_in"][:, :1]), t(fx["c1_K"]), t(fx["c1_mu"]), t(fx["c1_sigma"]),
                          t(fx["c1_weights"]), float(fx["dt"]))
    assert (out1 - t(fx["c1_out"])).abs().max().item() <= TOL


@pytest.mark.parametrize("name", golden_names("mace"))
def test_mace_step(name):
    fx, p = family_case("mace", name)
    out, aff = orc.mace_step(t(fx["state_in"]), t(p["K"]), t(p["mu_s"]), t(p["sigma_s"]), t(p["weights_s"]), float(fx["b"]))
    assert (aff - t(fx["aff"])).abs().max().item() <= TOL
    assert (out - t(fx["state_out"])).abs().max().item() <= TOL


@pytest.mark.parametrize("name", golden_names("xchan"))
def test_xchan_step(name):
    fx, p = family_case("xchan", name)
    s0 = t(fx["state_in"])
    out = orc.xchan_step(s0, t(p["K"]), t(p["mu_s"]), t(p["sigma_s"]), t(p["weights_s"]), float(fx["b"]),
                         float(fx["alpha_used"]))
    scale = max(1.0, float(s0.max()))
    assert (out - t(fx["state_out"])).abs().max().item() <= TOL * scale
    # mass conservation of the reference semantics (macelenia.py:89-120)
    assert abs(out.double().sum().item() - s0.double().sum().item()) <= 1e-5 * s0.double().sum().item()


@pytest.mark.parametrize("name", golden_names("nca"))
def test_nca_step(name):
    fx = load_npz(f"nca_{name}.npz")
    assert torch.equal(orc.sobel_kernels(16), t(fx["sobelkern"]))
    s0 = t(fx["state_in"])
    B, _, H, W = s0.shape
    mask = orc.nca_draw_mask(B, H, W, int(fx["seed"]))
    assert torch.equal(mask, t(fx["mask"]).bool())
    out = orc.nca_step(s0, t(fx["W1"]), t(fx["b1"]), t(fx["W2"]), t(fx["b2"]), mask)
    assert (out - t(fx["state_out"])).abs().max().item() <= TOL
    # dead-cell zeroing: every non-zero output cell is alive before and after
    assert ((out != 0).any(dim=1, keepdim=True) <= orc.nca_live_mask(s0)).all()


def test_perception_channel_quirk():
    """Output o of the grouped conv reads input channel o//2 (nca.py:341)."""
    torch.manual_seed(0)
    s = torch.rand(1, 16, 8, 8)
    p = orc.nca_perception(s)
    sx = torch.tensor([[-1, 0, 1], [-2, 0, 2], [-1, 0, 1]], dtype=torch.float) / 8
    sy = sx.t()
    for o in range(32):
        k = sx if o < 16 else sy
        c = o // 2
        ref = torch.zeros(8, 8)
        for a in range(3):
            for b in range(3):
                ref += k[a, b] * torch.roll(s[0, c], (1 - a, 1 - b), dims=(0, 1))
        assert torch.allclose(p[0, o], ref, atol=1e-6)
    assert torch.equal(p[:, 32:], s)


def test_lenia_k_mult_2():
    fx = load_npz("lenia_kmult2.npz")
    raw = {k: t(fx[k]) for k in ["mu", "sigma", "beta", "mu_k", "sigma_k", "weights"]}
    q = orc.sanitize_params(raw)
    K = orc.lenia_kernel(q["beta"], q["mu_k"], q["sigma_k"], 31)
    assert torch.allclose(K, t(fx["K"]), atol=1e-9, rtol=0) and torch.equal(q["weights"], t(fx["weights_s"]))
    out = orc.lenia_step(t(fx["state_in"]), K, q["mu"], q["sigma"], q["weights"], float(fx["dt"]))
    assert (out - t(fx["state_out"])).abs().max().item() <= TOL


def test_draw_oracle_matches_reference_fixture():
    """draw() + Automaton.worldmap restatements against the reference's own frames (tests/golden/draw.npz)."""
    fx = load_npz("draw.npz")
    wm = torch.zeros(3, *fx["gol_w0"].shape)
    w = fx["gol_w0"].astype(np.int32)
    color, decay = t(fx["gol_color"]), float(fx["gol_decay"])
    for i in range(3):
        wm = orc.gol_draw(w, wm, decay, color)
        assert torch.equal(wm, t(fx[f"gol_wm{i}"]))
        assert np.array_equal(orc.frame_u8(wm), fx[f"gol_frame{i}"])
        w = orc.gol_step(w)
    s3 = t(fx["lenia3_state"])
    assert torch.equal(orc.lenia_draw(s3), t(fx["lenia3_wm"]))
    assert np.array_equal(orc.frame_u8(orc.lenia_draw(s3)), fx["lenia3_frame"])
    assert torch.equal(orc.lenia_draw(s3[:, :1]), t(fx["lenia1_wm"]))
    assert np.array_equal(orc.frame_u8(orc.lenia_draw(s3[:, :1])), fx["lenia1_frame"])
    assert torch.equal(orc.lenia_draw(t(fx["mace_state"])), t(fx["mace_wm"]))
    assert torch.equal(orc.nca_draw(t(fx["nca_state"])), t(fx["nca_wm"]))
    assert np.array_equal(orc.frame_u8(orc.nca_draw(t(fx["nca_state"]))), fx["nca_frame"])


@pytest.mark.parametrize("name", ["underweight_skagway", "inaesthetic_maupassant", "unsuspected_lanthanotidae", "lenia_synth"])
def test_garbi_oracle_matches_reference_fixture(name):
    """g_arbi Fourier growth (incl. the global min/max rescale) against the reference's own step on the demo_data files."""
    from tests.util import garbi_case

    fx, garbi = garbi_case(name)
    s0, K, w = t(fx["state_in"]), t(fx["K"]), t(fx["weights_s"])
    fam = int(fx["family"])
    if fam == 0:
        out = orc.lenia_step(s0, K, None, None, w, float(fx["dt"]), garbi=garbi)
    elif fam == 1:
        out, aff = orc.mace_step(s0, K, None, None, w, float(fx["b"]), garbi=garbi)
        assert (aff - t(fx["aff"])).abs().max().item() <= TOL
    else:
        out = orc.xchan_step(s0, K, None, None, w, float(fx["b"]), float(fx["alpha_used"]), garbi=garbi)
    assert (out - t(fx["state_out"])).abs().max().item() <= TOL


@pytest.mark.parametrize("tag", ["mace", "xchan"])
def test_food_oracle_matches_reference_fixture(tag):
    """MaCE food subsystem (sensing, decay, respawn, consumption, cross-channel order) against the reference's own steps."""
    import random

    from tests.util import food_respawn_like_reference

    fx = load_npz("food.npz")
    p = load_npz(f"params_{tag}_{str(fx[tag + '_file'])}.npz")
    K, mu, sg, w = t(p["K"]), t(p["mu_s"]), t(p["sigma_s"]), t(p["weights_s"])
    state, food = t(fx[f"{tag}_state0"]), t(fx[f"{tag}_food0"])
    H, W = state.shape[2:]
    loss = torch.full((1,), float(fx[f"{tag}_loss0"]))
    random.seed(99)
    i = 1
    while f"{tag}_state{i}" in fx:
        state, food, loss, _ = orc.mace_food_step(state, food, loss, K, mu, sg, w, float(fx[f"{tag}_b"]), sense_food=(tag == "xchan"),
                                                  alpha=float(fx[f"{tag}_alpha"]) if tag == "xchan" else None,
                                                  respawn=food_respawn_like_reference(H, W))
        assert (state - t(fx[f"{tag}_state{i}"])).abs().max().item() <= TOL
        assert (food - t(fx[f"{tag}_food{i}"])).abs().max().item() <= TOL
        assert (loss - t(fx[f"{tag}_loss{i}"])).abs().max().item() <= 1e-4
        i += 1
    assert i > 1


def test_flow_lenia_oracle_matches_reference_fixture():
    """FlowLenia (affinity -> Sobel flow -> reintegration tracking) restatement against the reference's own steps."""
    fx = load_npz("flow.npz")
    for ci in range(int(fx["ncases"])):
        p = load_npz(f"params_mace_{str(fx[f'c{ci}_file'])}.npz")
        K, mu, sg, w = t(p["K"]), t(p["mu_s"]), t(p["sigma_s"]), t(p["weights_s"])
        dt, dd, srt, thx, n = [float(v) for v in fx[f"c{ci}_meta"]]
        state = t(fx[f"c{ci}_state0"])
        aff = orc.lenia_affinity(state, K, mu, sg, w)
        assert (orc.flow_field(state, aff, thx, int(n)) - t(fx[f"c{ci}_flow0"])).abs().max().item() <= 1e-5
        i = 1
        while f"c{ci}_state{i}" in fx:
            state = orc.flow_step(state, K, mu, sg, w, dt, int(dd), srt, thx, int(n))
            assert (state - t(fx[f"c{ci}_state{i}"])).abs().max().item() <= TOL * i
            i += 1
