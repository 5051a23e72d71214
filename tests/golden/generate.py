"""Generate the golden fixtures in this directory FROM THE UNMODIFIED REFERENCE.

Run in the build container only (needs ``/root/reference``):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/generate.py

Every fixture is {seeded input, parameters, output of the reference's own
``step()``}.  The reference code is imported headless through ``ref_loader``;
no arithmetic is stubbed.  Outputs are ``.npz`` files (numpy arrays only) so
they load on the GPU box without the reference or ``torch.load`` pickles.

Files written
  gol.npz                       CA2D worlds before / after n reference steps (several shapes + rules)
  params_<family>_<name>.npz    raw parameter dict of a demo_data file + the reference's kernel K
  lenia_<name>.npz              Lenia C=3 single step (+ C=1 slice of the same file)
  mace_<name>.npz               MaCELenia single step (+ Aff)
  xchan_<name>.npz              MaCELeniaXChan single step (+ Aff)
  xchan_realstate.npz           XChan step on a crop of a real saved state (values >> 1)
  nca_<model>.npz               NeuralCA weights + one forward step with the seeded update mask
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(HERE)), "baseline"))
from ref_loader import load_reference  # noqa: E402

REF = load_reference()
DEMO = os.path.join(REF.root, "demo_data")
torch.set_grad_enabled(False)


def np_(t):
    return t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else np.asarray(t)


def save(name, **arrs):
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **{k: np_(v) for k, v in arrs.items()})
    print(f"wrote {name}: {os.path.getsize(path) / 1024:.0f} KiB")


# ------------------------------------------------------------------ GoL
def gen_gol():
    out = {}
    cases = [  # (H, W, s_rule, b_rule, steps)
        (16, 16, "23", "3", 1),
        (33, 70, "23", "3", 8),     # W not a multiple of 32
        (250, 250, "23", "3", 16),  # simulate.py default world (BASELINE configs[0])
        (64, 96, "1357", "1357", 5),    # replicator
        (40, 31, "45678", "3", 7),      # coral; W < 32
        (37, 129, "012345678", "0", 3),   # B0: dead cells with 0 neighbours are born
        (8, 257, "238", "357", 9),
    ]
    for ci, (H, W, s, b, steps) in enumerate(cases):
        torch.manual_seed(100 + ci)
        a = REF.CA2D((H, W), s_num=s, b_num=b)
        a.world = (torch.rand(H, W) < 0.45).to(torch.int)
        w0 = a.world.clone()
        for _ in range(steps):
            a.step()
        assert a.world.dtype == torch.int32
        out[f"c{ci}_w0"] = w0.to(torch.uint8)
        out[f"c{ci}_w1"] = a.world.to(torch.uint8)
        out[f"c{ci}_meta"] = np.array([H, W, a.s_num, a.b_num, steps], dtype=np.int64)
    # the reference's own init (centred random square, int16 world; ca2d.py:227-246)
    torch.manual_seed(7)
    a = REF.CA2D((250, 250))
    w0 = a.world.clone()
    assert w0.dtype == torch.int16
    for _ in range(32):
        a.step()
    ci = len(cases)
    out[f"c{ci}_w0"] = w0.to(torch.uint8)
    out[f"c{ci}_w1"] = a.world.to(torch.uint8)
    out[f"c{ci}_meta"] = np.array([250, 250, a.s_num, a.b_num, 32], dtype=np.int64)
    out["ncases"] = np.array(ci + 1)
    save("gol.npz", **out)


# ------------------------------------------------------------------ Lenia family
PARAM_KEYS = ["mu", "sigma", "beta", "mu_k", "sigma_k", "weights"]


def is_plain(d):
    return not d.get("k_arbi", False) and not d.get("g_arbi", False)


def pick(folder, n):
    files = sorted(f for f in os.listdir(os.path.join(DEMO, folder)) if f.endswith(".pt"))
    chosen = []
    for f in files:
        d = torch.load(os.path.join(DEMO, folder, f), weights_only=True, map_location="cpu")
        if is_plain(d):
            chosen.append(f)
        if len(chosen) == n:
            break
    return chosen


def raw_params(folder, f):
    d = torch.load(os.path.join(DEMO, folder, f), weights_only=True, map_location="cpu")
    out = {k: d[k] for k in PARAM_KEYS}
    out["k_size"] = np.array(d["k_size"])
    if "alpha" in d:
        out["alpha"] = np.array(float(d["alpha"]))
    return d, out


def gen_lenia_family():
    H, W = 48, 64
    for folder, cls, tag, n in [
        ("demo_lenia", REF.Lenia, "lenia", 5),
        ("demo_macelenia", REF.MaCELenia, "mace", 5),
        ("demo_macelenia_xchan", REF.MaCELeniaXChan, "xchan", 6),
    ]:
        for fi, f in enumerate(pick(folder, n)):
            name = f[:-3]
            d, raw = raw_params(folder, f)
            params = REF.LeniaParams(from_file=os.path.join(DEMO, folder, f))
            g = torch.Generator().manual_seed(1000 + fi)
            s0 = torch.rand(1, 3, H, W, generator=g)
            if tag != "lenia":
                s0 = s0 * (torch.rand(1, 3, H, W, generator=g) < 0.6)  # some empty cells
            auto = cls((H, W), params=params, state_init=s0.clone())
            save(f"params_{tag}_{name}.npz", K=auto.k, mu_s=auto.mu, sigma_s=auto.sigma, weights_s=auto.weights, **raw)
            extra = {}
            if tag == "lenia":
                auto.step()
                # C=1 slice of the same file (BASELINE configs[1]: single-channel Lenia)
                d1 = {k: (v[:, :1, :1] if isinstance(v, torch.Tensor) else v) for k, v in d.items()}
                p1 = REF.LeniaParams(param_dict=d1, channels=1)
                a1 = REF.Lenia((H, W), num_channels=1, params=p1, state_init=s0[:, :1].clone())
                a1.step()
                extra = dict(c1_out=a1.state, c1_K=a1.k, c1_mu=a1.mu, c1_sigma=a1.sigma, c1_weights=a1.weights)
                aff = auto.state.new_zeros(0)
            elif tag == "mace":
                aff = auto._compute_affinity()
                auto.step()
            else:
                aff = auto._compute_affinity()
                auto.step()
                extra = dict(b=np.array(float(auto.b)), alpha_used=np.array(float(auto.alpha)))
            if tag == "mace":
                extra = dict(b=np.array(float(auto.b)))
            save(f"{tag}_{name}.npz", state_in=s0, state_out=auto.state, aff=aff, dt=np.array(auto.dt), **extra)

    # realistic high-dynamic-range state: crop of a real saved MaCE-XChan state
    f = "arborical_asbestosis.pt"
    st = torch.load(os.path.join(DEMO, "demo_macelenia_xchan/state_saves/gilded_wok__state.pt"), weights_only=True, map_location="cpu")
    full = st["state"]
    crop = full[:, :, 150:150 + 64, 180:180 + 64].contiguous()
    pd = {k: v for k, v in st.items() if k != "state"}
    params = REF.LeniaParams(param_dict=pd)
    auto = REF.MaCELeniaXChan((64, 64), params=params, state_init=crop.clone())
    if "alpha" in pd:
        auto.alpha = float(pd["alpha"])
    aff = auto._compute_affinity()
    auto.step()
    raw = {k: pd[k] for k in PARAM_KEYS}
    save("xchan_realstate.npz", state_in=crop, state_out=auto.state, aff=aff, K=auto.k, mu_s=auto.mu,
         sigma_s=auto.sigma, weights_s=auto.weights, b=np.array(float(auto.b)),
         alpha_used=np.array(float(auto.alpha)), k_size=np.array(pd["k_size"]), **raw)


def gen_kmult():
    """k_mult = 2 (two kernels per source channel; no demo_data file has it): synthetic parameters, reference step."""
    torch.manual_seed(5)
    C, km, R = 2, 2, 3
    d = {"mu": 0.1 + 0.2 * torch.rand(1, C * km, C), "sigma": 0.03 + 0.1 * torch.rand(1, C * km, C),
         "beta": torch.rand(1, C * km, C, R), "mu_k": torch.rand(1, C * km, C, R),
         "sigma_k": 0.05 + 0.2 * torch.rand(1, C * km, C, R), "weights": torch.rand(1, C * km, C), "k_size": 31, "k_mult": km,
         "num_channels": C}
    g = torch.Generator().manual_seed(1)
    s0 = torch.rand(1, C, 64, 128, generator=g)
    raw = {k: v.clone() for k, v in d.items() if isinstance(v, torch.Tensor)}
    p = REF.LeniaParams(param_dict=dict(d), channels=C, k_mult=km)
    a = REF.Lenia((64, 128), num_channels=C, params=p, state_init=s0.clone())
    K, mu_s, sg_s, w_s = a.k.clone(), a.mu.clone(), a.sigma.clone(), a.weights.clone()
    a.step()
    save("lenia_kmult2.npz", state_in=s0, state_out=a.state, K=K, mu_s=mu_s, sigma_s=sg_s, weights_s=w_s, k_size=np.array(31),
         k_mult=np.array(km), dt=np.array(a.dt), **raw)


# ------------------------------------------------------------------ NCA
def gen_nca():
    folder = os.path.join(DEMO, "NCA_models")
    for mi, f in enumerate(sorted(os.listdir(folder))):
        name = f[:-3]
        sd = torch.load(os.path.join(folder, f), map_location="cpu")["state_dict"]
        model = REF.NCAModule(n_states=16, n_hidden=128)
        model.load_model(os.path.join(folder, f))
        model.eval()
        B, H, W = 2, 24, 40
        # world 0: dense random; world 1: grown from the all-ones seed pixel (sparse, exercises life masks)
        g = torch.Generator().manual_seed(50 + mi)
        dense = torch.rand(1, 16, H, W, generator=g)
        grown = torch.zeros(1, 16, H, W)
        grown[0, :, H // 2, W // 2] = 1.0
        torch.manual_seed(77 + mi)
        for _ in range(12):
            grown = model(grown)
        s0 = torch.cat([dense, grown], dim=0)
        seed = 900 + mi
        torch.manual_seed(seed)
        out = model(s0.clone())
        gm = torch.Generator().manual_seed(seed)
        mask = torch.rand(B, 1, H, W, generator=gm) <= 0.5
        save(f"nca_{name}.npz", W1=sd["computer.0.weight"], b1=sd["computer.0.bias"], W2=sd["computer.2.weight"],
             b2=sd["computer.2.bias"], sobelkern=sd["sobel.sobelkern"], state_in=s0, mask=mask.to(torch.uint8),
             state_out=out, seed=np.array(seed))




# ------------------------------------------------------------------ draw() / worldmap
def gen_draw():
    """The reference's own draw() + Automaton.worldmap on seeded states: float (3,H,W) `_worldmap` and the (W,H,3) uint8
    frame.  CA2D: three frames with two steps in between (the echo trail depends on the previous frame)."""
    out = {}
    H, W = 37, 70
    torch.manual_seed(21)
    a = REF.CA2D((H, W))
    a.world = (torch.rand(H, W) < 0.4).to(torch.int)
    out["gol_w0"] = a.world.to(torch.uint8)
    out["gol_color"] = a.highlight_color
    out["gol_decay"] = np.array(a.decay_speed, dtype=np.float64)
    for i in range(3):
        a.draw()
        out[f"gol_wm{i}"] = a._worldmap.clone()
        out[f"gol_frame{i}"] = a.worldmap
        a.step()
    # Lenia C=3 / C=1 and MaCE (values above 1 get clamped)
    f = "abominable_mongolic.pt"
    d = torch.load(os.path.join(DEMO, "demo_lenia", f), weights_only=True, map_location="cpu")
    g = torch.Generator().manual_seed(31)
    s3 = torch.rand(1, 3, 40, 52, generator=g) * 1.3 - 0.1
    au = REF.Lenia((40, 52), params=REF.LeniaParams(from_file=os.path.join(DEMO, "demo_lenia", f)), state_init=s3.clone())
    au.draw()
    out["lenia3_state"], out["lenia3_wm"], out["lenia3_frame"] = s3, au._worldmap.clone(), au.worldmap
    d1 = {k: (v[:, :1, :1] if isinstance(v, torch.Tensor) else v) for k, v in d.items()}
    a1 = REF.Lenia((40, 52), num_channels=1, params=REF.LeniaParams(param_dict=d1, channels=1), state_init=s3[:, :1].clone())
    a1.draw()
    out["lenia1_wm"], out["lenia1_frame"] = a1._worldmap.clone(), a1.worldmap
    fm = pick("demo_macelenia", 1)[0]
    sm = torch.rand(1, 3, 40, 52, generator=g) * 3.0
    am = REF.MaCELenia((40, 52), params=REF.LeniaParams(from_file=os.path.join(DEMO, "demo_macelenia", fm)), state_init=sm.clone())
    am.draw()
    out["mace_state"], out["mace_wm"], out["mace_frame"] = sm, am._worldmap.clone(), am.worldmap
    # NCA: draw() = clamp(argb) blend + a brush overlay at the GUI mouse position; the mouse is parked far outside the
    # world so the overlay is empty and `_worldmap` is the blend alone
    an = REF.NeuralCA((24, 40), os.path.join(DEMO, "NCA_models"))
    an.state = torch.rand(1, 16, 24, 40, generator=g) * 1.4 - 0.2
    an.m_pos = type("P", (), {"x": -1000, "y": -1000})()
    an.draw()
    out["nca_state"], out["nca_wm"], out["nca_frame"] = an.state, an._worldmap.clone(), an.worldmap
    save("draw.npz", **out)


# ------------------------------------------------------------------ g_arbi growth functions
def _raw_dict(d):
    """every entry of a parameter file as numpy under the key p_<name> (tuples -> arrays, None -> nan)."""
    out = {}
    for k, v in d.items():
        if k == "state":
            continue
        if v is None:
            v = np.nan
        out["p_" + k] = np_(v) if isinstance(v, torch.Tensor) else np.asarray(v, dtype=np.float64)
    return out


def gen_garbi():
    """Fourier-series growth functions (lenia.py:390-418, funcgen.py:79-129): the three demo_data files that use them
    (global min/max rescale on two of them, a k_arbi kernel on one) + a synthetic Lenia case with a clip."""
    H, W = 48, 64
    cases = [("demo_macelenia", "underweight_skagway", REF.MaCELenia, 1),
             ("demo_macelenia_xchan", "inaesthetic_maupassant", REF.MaCELeniaXChan, 2),
             ("demo_macelenia_xchan", "unsuspected_lanthanotidae", REF.MaCELeniaXChan, 2)]
    for ci, (folder, name, cls, family) in enumerate(cases):
        path = os.path.join(DEMO, folder, name + ".pt")
        d = torch.load(path, weights_only=True, map_location="cpu")
        g = torch.Generator().manual_seed(2000 + ci)
        s0 = torch.rand(1, 3, H, W, generator=g) * (torch.rand(1, 3, H, W, generator=g) < 0.6)
        auto = cls((H, W), params=REF.LeniaParams(from_file=path), state_init=s0.clone())
        if family == 2 and "alpha" in d:
            auto.alpha = float(d["alpha"])
        assert auto.g_arbi
        aff = auto._compute_affinity()
        K, w_s = auto.k.clone(), auto.weights.clone()
        auto.step()
        rs = d.get("g_rescale")
        save(f"garbi_{name}.npz", state_in=s0, state_out=auto.state, aff=aff, K=K, weights_s=w_s, g_coeffs=d["g_coeffs"],
             g_harmonics=d["g_harmonics"], g_rescale=np.array(rs if rs is not None else [np.nan, np.nan], dtype=np.float64),
             g_clip=np.array(np.nan if d.get("g_clip") is None else d["g_clip"], dtype=np.float64), family=np.array(family),
             b=np.array(float(auto.b)), alpha_used=np.array(float(getattr(auto, "alpha", 0.0))), k_size=np.array(d["k_size"]),
             **_raw_dict(d))
    # plain Lenia, C = 2, ring kernels + Fourier growth with rescale and clip
    torch.manual_seed(9)
    C = 2
    d = {"beta": torch.rand(1, C, C, 3), "mu_k": torch.rand(1, C, C, 3), "sigma_k": 0.05 + 0.2 * torch.rand(1, C, C, 3),
         "weights": torch.rand(1, C, C), "k_size": 31, "k_mult": 1, "num_channels": C, "g_arbi": True, "k_arbi": False,
         "g_coeffs": torch.randn(1, C, C, 8), "g_harmonics": torch.tensor([0.0, 1.0, 2.0, 3.0]).expand(1, C, C, 4).clone(),
         "g_rescale": (-1.0, 1.0), "g_clip": -0.5}
    g = torch.Generator().manual_seed(3)
    s0 = torch.rand(1, C, 40, 72, generator=g)
    raw = {k: (v.clone() if isinstance(v, torch.Tensor) else v) for k, v in d.items()}
    a = REF.Lenia((40, 72), num_channels=C, params=REF.LeniaParams(param_dict=dict(d), channels=C), state_init=s0.clone())
    assert a.g_arbi
    K, w_s = a.k.clone(), a.weights.clone()
    a.step()
    save("garbi_lenia_synth.npz", state_in=s0, state_out=a.state, aff=np.zeros(0, np.float32), K=K, weights_s=w_s,
         g_coeffs=raw["g_coeffs"], g_harmonics=raw["g_harmonics"], g_rescale=np.array([-1.0, 1.0]), g_clip=np.array(-0.5),
         family=np.array(0), b=np.array(0.0), alpha_used=np.array(0.0), k_size=np.array(31), dt=np.array(a.dt), **_raw_dict(raw))


# ------------------------------------------------------------------ MaCE food subsystem
def gen_food():
    """has_food=True steps of the reference (macelenia.py:72-87, :161-226): plain MaCE with a forced respawn (the world's
    cum_loss_mass starts just under the 200 threshold; spot positions come from Python's `random`, seeded), and the
    cross-channel variant with food sensing over two steps."""
    import random

    H, W = 48, 64
    out = {}
    for tag, folder, cls, sense, nsteps, loss0 in [("mace", "demo_macelenia", REF.MaCELenia, False, 1, 199.95),
                                                    ("xchan", "demo_macelenia_xchan", REF.MaCELeniaXChan, True, 2, 0.0)]:
        f = pick(folder, 1)[0]
        g = torch.Generator().manual_seed(4000 + len(out))
        s0 = torch.rand(1, 3, H, W, generator=g) * (torch.rand(1, 3, H, W, generator=g) < 0.6)
        auto = cls((H, W), params=REF.LeniaParams(from_file=os.path.join(DEMO, folder, f)), state_init=s0.clone())
        random.seed(1234)
        auto.set_food(True)
        auto.sense_food = sense
        auto.cum_loss_mass = torch.full((1,), loss0)
        out[f"{tag}_file"] = np.array(f[:-3])
        out[f"{tag}_state0"], out[f"{tag}_food0"] = s0, auto.food_channel.clone()
        out[f"{tag}_b"] = np.array(float(auto.b))
        out[f"{tag}_alpha"] = np.array(float(getattr(auto, "alpha", 0.0)))
        out[f"{tag}_loss0"] = np.array(loss0)
        random.seed(99)
        for i in range(nsteps):
            auto.step()
            out[f"{tag}_state{i + 1}"], out[f"{tag}_food{i + 1}"] = auto.state.clone(), auto.food_channel.clone()
            out[f"{tag}_loss{i + 1}"] = auto.cum_loss_mass.clone()
        out[f"{tag}_total_mass"] = np.array(float(auto.total_mass()))
        auto.draw()
        out[f"{tag}_wm"], out[f"{tag}_frame"] = auto._worldmap.clone(), auto.worldmap
    save("food.npz", **out)


# ------------------------------------------------------------------ FlowLenia
def gen_flow():
    """FlowLenia.step of the reference (flow_lenia.py:246-253) with the demo_macelenia parameters its DEFAULTS entry points
    at (interface/defaults.py:14-17): single steps on sparse random states whose values reach above 1 (alpha clips), two
    consecutive steps, and a non-square world."""
    out = {}
    files = pick("demo_macelenia", 2)
    cases = [(files[0], 48, 64, 1, 1.6), (files[1], 40, 72, 2, 2.5)]
    for ci, (f, H, W, nsteps, amp) in enumerate(cases):
        g = torch.Generator().manual_seed(5000 + ci)
        s0 = amp * torch.rand(1, 3, H, W, generator=g) * (torch.rand(1, 3, H, W, generator=g) < 0.5)
        a = REF.FlowLenia((H, W), params=REF.LeniaParams(from_file=os.path.join(DEMO, "demo_macelenia", f)), state_init=s0.clone())
        out[f"c{ci}_file"] = np.array(f[:-3])
        out[f"c{ci}_state0"] = s0
        out[f"c{ci}_flow0"] = a.compute_affinity()
        out[f"c{ci}_meta"] = np.array([a.dt, a.dd, a.sigma_rt, a.theta_x, a.n], dtype=np.float64)
        for i in range(nsteps):
            a.step()
            out[f"c{ci}_state{i + 1}"] = a.state.clone()
    out["ncases"] = np.array(len(cases))
    save("flow.npz", **out)


def gen_default_params():
    """LeniaParams(batch_size, k_size, channels) of the reference (random parameters, leniaparams.py:309-353) under a fixed
    torch seed, after its _sanitize: what Lenia(params=None) starts from."""
    out = {}
    for tag, (B, C, km) in {"b2c3": (2, 3, 1), "b1c2k2": (1, 2, 2)}.items():
        torch.manual_seed(123)
        p = REF.LeniaParams(batch_size=B, k_size=31, channels=C, k_mult=km)
        for k in PARAM_KEYS:
            out[f"{tag}_{k}"] = p[k]
        out[f"{tag}_meta"] = np.array([B, C, km, p["k_size"]], dtype=np.int64)
    save("default_params.npz", **out)


if __name__ == "__main__":
    only = [a[2:] for a in sys.argv[1:] if a.startswith("--")]   # e.g. --draw regenerates draw.npz alone
    gens = {"gol": gen_gol, "family": gen_lenia_family, "kmult": gen_kmult, "nca": gen_nca, "draw": gen_draw, "garbi": gen_garbi, "food": gen_food, "flow": gen_flow, "default_params": gen_default_params}
    for name, fn in gens.items():
        if not only or name in only:
            fn()
