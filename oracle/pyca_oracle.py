"""CPU oracle for the PyCA ``Automaton.step()`` hot path.  TEST INFRASTRUCTURE ONLY.

This file restates, in plain functional torch/numpy on the CPU, the arithmetic
of the reference's ``step()`` for the five models in scope.  Nothing under
``pyca_b200/`` imports it: only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` do, and only as the
checker / the timed CPU arm.

Pinning status: the reference ships no tests or golden vectors,
so the oracle is pinned against the reference ITSELF: ``tests/golden/generate.py``
imports the unmodified reference headless in the build container, runs its
``step()`` on seeded inputs and commits inputs+outputs as ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` checks every function below against those
fixtures (bit-exact for Game of Life, <=1e-6 for the float models; the float
functions use the same torch primitives as the reference -- ``fft2``, ``unfold``,
``conv2d``, ``Linear`` -- in the same order).

Reference lines followed (relative to the PyCA tree):
  gol_step              pyca/automata/models/ca2d.py:195-209, :221-225
  sanitize_params       pyca/automata/models/lenia/leniaparams.py:68-87
  lenia_kernel          pyca/automata/models/lenia/lenia.py:289-362 (ring kernels; k_arbi out of scope)
  kernel_fft            pyca/automata/models/lenia/lenia.py:364-380
  fft_conv              pyca/automata/models/lenia/lenia.py:453-469
  growth                pyca/automata/models/lenia/lenia.py:419-428
  growth_arbi           pyca/automata/models/lenia/lenia.py:390-418 + lenia/funcgen.py:79-129 (g_arbi Fourier growth)
  lenia_step            pyca/automata/models/lenia/lenia.py:430-451
  mace_affinity         pyca/automata/models/lenia/macelenia.py:122-159 (food off)
  mace_step             pyca/automata/models/lenia/macelenia.py:89-120
  xchan_step            pyca/automata/models/lenia/mace_lenia_xchan.py:59-76
  flow_sobel            pyca/automata/models/lenia/flow_lenia.py:6-47 (zero-padded, unnormalised)
  flow_field            pyca/automata/models/lenia/flow_lenia.py:207-228 (FlowLenia.compute_affinity)
  flow_reintegrate      pyca/automata/models/lenia/flow_lenia.py:97-153 (ReintegrationTracker.step / apply_flow)
  flow_step             pyca/automata/models/lenia/flow_lenia.py:246-253 (FlowLenia.step, food off)
  food_affinity         pyca/automata/models/lenia/macelenia.py:149-157 (sense_food)
  food_decay            pyca/automata/models/lenia/macelenia.py:171-185
  food_consume          pyca/automata/models/lenia/macelenia.py:194-226 (death_enabled=False)
  mace_food_step        pyca/automata/models/lenia/macelenia.py:72-87, :161-192; mace_lenia_xchan.py:59-66 (has_food)
  nca_perception        pyca/automata/models/nca.py:306-344
  nca_live_mask         pyca/automata/models/nca.py:226-237
  nca_step              pyca/automata/models/nca.py:239-263
  frame_u8              pyca/automata/automaton.py:209-216 (Automaton.worldmap)
  gol_draw              pyca/automata/models/ca2d.py:117-124
  lenia_draw            pyca/automata/models/lenia/lenia.py:481-500, macelenia.py:365-385 (single-world view)
  nca_draw              pyca/automata/models/nca.py:74-87 (blend; brush overlay is GUI)
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------- #
# Game of Life / outer-totalistic CA2D
# --------------------------------------------------------------------------- #
def rule_mask(rule) -> int:
    """'23' -> 0b1100 (ca2d.py:80-88).  Ints pass through."""
    if isinstance(rule, str):
        return sum(1 << int(ch) for ch in rule)
    return int(rule)


def gol_step(world: np.ndarray, s_mask=0b1100, b_mask=0b1000) -> np.ndarray:
    """One periodic Moore-neighbourhood outer-totalistic step on an (H,W) 0/1 array.

    ca2d.py:195-209: count = sum of the 8 rolled copies; next = bit `count` of
    s_mask where the cell is alive else of b_mask."""
    w = np.asarray(world).astype(np.int32)
    cnt = np.zeros_like(w)
    for dy in (-1, 0, 1):
        for dx in (-1, 0, 1):
            if dy or dx:
                cnt += np.roll(w, (dy, dx), axis=(0, 1))
    s = (int(s_mask) >> cnt) & 1
    b = (int(b_mask) >> cnt) & 1
    return np.where(w == 1, s, b).astype(np.int32)


def gol_run(world, steps, s_mask=0b1100, b_mask=0b1000):
    w = np.asarray(world).astype(np.int32)
    for _ in range(steps):
        w = gol_step(w, s_mask, b_mask)
    return w


# --------------------------------------------------------------------------- #
# Lenia family: parameters and kernels (cold path)
# --------------------------------------------------------------------------- #
def sanitize_params(p: dict) -> dict:
    """leniaparams.py:68-87: clamp ranges, weights >= 0 and normalised over dim 1."""
    q = dict(p)
    if "mu" in q:
        q["mu"] = torch.clamp(q["mu"], -2, 2)
    if "sigma" in q:
        q["sigma"] = torch.clamp(q["sigma"], 1e-4, 1.0)
    if "beta" in q:
        q["beta"] = torch.clamp(q["beta"], 0, None)
    if "mu_k" in q:
        q["mu_k"] = torch.clamp(q["mu_k"], 0.0, 2.0)
    if "sigma_k" in q:
        q["sigma_k"] = torch.clamp(q["sigma_k"], 1e-4, 1.0)
    w = torch.clamp(q["weights"], 0, None)
    n = w.sum(dim=1, keepdim=True)
    q["weights"] = torch.where(n > 1.0e-6, w / n, 0)
    return q


def lenia_kernel(beta, mu_k, sigma_k, k_size: int) -> torch.Tensor:
    """Ring kernels on the full k x k square, normalised to sum 1 (lenia.py:289-362).

    beta/mu_k/sigma_k: (B, C*k_mult, C, rings).  Returns (B, C*k_mult, C, k, k)."""
    ax = torch.linspace(-1, 1, k_size).to(beta.device)
    gx, gy = torch.meshgrid(ax, ax, indexing="xy")
    r = torch.sqrt(gx**2 + gy**2)
    rr = r[None, None, None, None].expand(*beta.shape, -1, -1)
    rings = torch.exp(-(((rr - mu_k[..., None, None]) / sigma_k[..., None, None]) ** 2) / 2)
    K = torch.sum(beta[..., None, None] * rings, dim=3)
    tot = torch.sum(K, dim=(-1, -2), keepdim=True)
    tot = torch.where(tot < 1e-6, 1, tot)
    return K / tot


def kernel_fft(K: torch.Tensor, h: int, w: int) -> torch.Tensor:
    """Zero-pad to (h,w), centre on the origin, fft2 (lenia.py:364-380)."""
    k = K.shape[-1]
    Kp = F.pad(K, [0, w - k, 0, h - k])
    Kp = Kp.roll((-(k // 2), -(k // 2)), dims=(-1, -2))
    return torch.fft.fft2(Kp)


def fft_conv(state: torch.Tensor, Kf: torch.Tensor, k_mult: int = 1) -> torch.Tensor:
    """Periodic convolution of every source channel with every kernel (lenia.py:453-469).

    state (B,C,H,W); Kf (B,C*k_mult,C,H,W) complex.  Returns U (B,C*k_mult,C,H,W) with
    U[b, i*k_mult+m, j] = K[b, i*k_mult+m, j] (*) state[b, i]."""
    B, C, H, W = state.shape
    S = torch.fft.fft2(state)[:, :, None, None]
    prod = S * Kf.reshape(B, C, k_mult, C, H, W)
    prod = prod.reshape(B, C * k_mult, C, H, W)
    return torch.real(torch.fft.ifft2(prod))


def growth(U, mu, sigma):
    """Gaussian growth 2*exp(-((u-mu)^2/sigma^2)/2)-1 (lenia.py:419-428)."""
    m = mu[..., None, None]
    s = sigma[..., None, None]
    return 2 * torch.exp(-((U - m) ** 2 / s**2) / 2) - 1


def growth_arbi(U, g_coeffs, g_harmonics, g_rescale=None, g_clip=None):
    """g_arbi growth (lenia.py:390-418): ArbitraryFunction.forward (funcgen.py:79-129) on the range (0, 2), one function
    per (b, i*k_mult+m, j) plane; `g_rescale` = global min/max rescale over each plane, `g_clip` = lower clip.
    U (B,CK,C,H,W); g_coeffs (B,CK,C,2*nh) as (cos,sin) pairs; g_harmonics (B,CK,C,nh)."""
    B, CK, C, H, W = U.shape
    n = B * CK * C
    co = g_coeffs.reshape(n, -1, 2)
    harm = g_harmonics.reshape(n, -1)
    ranges = torch.tensor([0.0, 2.0])[None, :].expand(n, -1)
    shifts, period = ranges[:, 0], ranges[:, 1] - ranges[:, 0]
    x = U.reshape(n, 1, -1)
    interior = 2 * torch.pi / period[:, None, None] * harm[..., None] * (x - shifts[:, None, None])
    values = torch.sum(co[:, :, 0][..., None] * torch.cos(interior) + co[:, :, 1][..., None] * torch.sin(interior), dim=1)
    if g_rescale is not None:
        mn, _ = values.min(dim=-1, keepdim=True)
        mx, _ = values.max(dim=-1, keepdim=True)
        values = (values - mn) / (mx - mn + 1e-8)
        values = values * (g_rescale[1] - g_rescale[0]) + g_rescale[0]
    if g_clip is not None:
        values[values < g_clip] = g_clip
    return values.reshape(B, CK, C, H, W)


def lenia_affinity(state, K, mu, sigma, weights, Kf=None, garbi=None):
    """Weighted source-sum of growths: (B,C,H,W) (lenia.py:436-446 / macelenia.py:144-150).
    garbi = (g_coeffs, g_harmonics, g_rescale, g_clip) replaces the Gaussian growth by the Fourier one."""
    B, C, H, W = state.shape
    k_mult = K.shape[1] // C
    if Kf is None:
        Kf = kernel_fft(K, H, W)
    U = fft_conv(state, Kf, k_mult)
    G = growth(U, mu, sigma) if garbi is None else growth_arbi(U, *garbi)
    return (G * weights[..., None, None]).sum(dim=1)


def lenia_step(state, K, mu, sigma, weights, dt=0.1, Kf=None, garbi=None):
    """lenia.py:430-451."""
    dx = lenia_affinity(state, K, mu, sigma, weights, Kf, garbi)
    return torch.clamp(state + dt * dx, 0, 1)


def _sum3x3_unfold(x):
    """Circular 3x3 box sum exactly as the reference does it: pad + unfold + sum
    (macelenia.py:107-109)."""
    B, C, H, W = x.shape
    p = F.pad(x, (1, 1, 1, 1), mode="circular")
    return F.unfold(p, kernel_size=(3, 3)).reshape(B, C, 9, H, W)


def mace_redistribute(state, aff, b):
    """macelenia.py:103-118: E=exp(b*Aff); Z=sum3x3(E); new = E * sum3x3(state/Z)."""
    E = torch.exp(b * aff)
    Z = _sum3x3_unfold(E).sum(dim=2)
    give = _sum3x3_unfold(state / Z)
    return (E[:, :, None] * give).sum(dim=2)


def mace_step(state, K, mu, sigma, weights, b=8.0, Kf=None, garbi=None):
    """macelenia.py:72-120 with food off.  Returns (new_state, Aff)."""
    aff = lenia_affinity(state, K, mu, sigma, weights, Kf, garbi)
    return mace_redistribute(state, aff, b), aff


def xchan_mix(state, aff, b, alpha):
    """mace_lenia_xchan.py:68-76."""
    mx = torch.max(aff, dim=1, keepdim=True)[0]
    num = torch.exp(b * (aff - mx))
    p = num / num.sum(dim=1, keepdim=True)
    target = state.sum(dim=1, keepdim=True) * p
    return state - (state - target) * alpha


def xchan_step(state, K, mu, sigma, weights, b=6.0, alpha=0.03, Kf=None, garbi=None):
    """mace_lenia_xchan.py:59-66 with food off."""
    new, aff = mace_step(state, K, mu, sigma, weights, b, Kf, garbi)
    return xchan_mix(new, aff, b, alpha)


# ---- MaCE food subsystem (off by default in the reference) ------------------------------------------------------
def food_affinity(aff, food, K, Kf=None):
    """macelenia.py:149-157: Aff + mask + sum over sources of K (*) mask, mask = (food > 0) on every channel."""
    B, C, H, W = aff.shape
    k_mult = K.shape[1] // C
    if Kf is None:
        Kf = kernel_fft(K, H, W)
    food_aff = (food > 0).float().expand(-1, C, -1, -1)
    food_aff = food_aff + fft_conv(food_aff, Kf, k_mult).sum(dim=1)
    return aff + food_aff


def food_decay(state):
    """macelenia.py:171-185 (decay variant 3): returns (state - decay, per-world mass lost)."""
    decay = torch.where(state > 0.02, 0.0003, torch.zeros_like(state))
    return state - decay, decay.view(state.shape[0], -1).sum(dim=1)


def food_consume(state, food, min_density=0.1, transfer_rate=0.1):
    """macelenia.py:194-226 with death_enabled=False: returns (state + transfer/3, food - transfer)."""
    where_food = food > 0
    where_contact = state.sum(dim=1, keepdim=True) >= min_density
    overlap = where_food & where_contact
    transfer = torch.minimum(food, torch.ones_like(where_food) * overlap * transfer_rate)
    return state + transfer / 3, food - transfer


def mace_food_step(state, food, cum_loss, K, mu, sigma, weights, b, sense_food=False, alpha=None, respawn=None, Kf=None,
                   garbi=None):
    """One MaCELenia.step (alpha None) or MaCELeniaXChan.step (alpha given) with has_food=True.
    `respawn(food, batches) -> food` stands in for random_food_chan(200, 5, 7, add_to_exisitng=True) (host RNG).
    Returns (state, food, cum_loss, Aff)."""
    aff = lenia_affinity(state, K, mu, sigma, weights, Kf, garbi)
    if sense_food:
        aff = food_affinity(aff, food, K, Kf)
    state = mace_redistribute(state, aff, b)
    state, lost = food_decay(state)
    cum_loss = cum_loss + lost
    idx = [p[0] for p in torch.argwhere(cum_loss >= 200).tolist()]
    if idx:
        cum_loss = cum_loss.clone()
        cum_loss[idx] = cum_loss[idx] - 200
        food = respawn(food, idx)
    state, food = food_consume(state, food, 0.3 if sense_food else 0.1, 0.1)
    if alpha is not None:
        state = xchan_mix(state, aff, b, alpha)
    return state, food, cum_loss, aff


# ---- FlowLenia -----------------------------------------------------------------------------------
def flow_sobel(x):
    """flow_lenia.py:6-47: per-channel 3x3 Sobel cross-correlations with ZERO padding ("same"), no 1/8 factor.
    x (B,C,H,W) -> (B,C,2,H,W): component 0 = derivative along the width, 1 = along the height."""
    C = x.shape[1]
    k_w = torch.tensor([[-1, 0, 1], [-2, 0, 2], [-1, 0, 1]], dtype=x.dtype, device=x.device).tile((C, 1, 1, 1))
    k_h = torch.tensor([[-1, -2, -1], [0, 0, 0], [1, 2, 1]], dtype=x.dtype, device=x.device).tile((C, 1, 1, 1))
    sw = F.conv2d(x, k_w, groups=C, stride=1, padding="same")
    sh = F.conv2d(x, k_h, groups=C, stride=1, padding="same")
    return torch.stack((sw, sh), dim=2)


def flow_field(state, aff, theta_x=2.0, n=2):
    """flow_lenia.py:212-226: F = grad(Aff) * (1 - alpha) - grad(total mass) * alpha, alpha = clip((mass/theta_x)^n, 0, 1)."""
    grad_u = flow_sobel(aff)
    mass = state.sum(dim=1, keepdims=True)
    grad_x = flow_sobel(mass)
    alpha = ((mass / theta_x) ** n).clip(0, 1)[:, :, None]
    return grad_u * (1 - alpha) - grad_x * alpha


def flow_reintegrate(state, flow, dt=0.1, dd=2, sigma=0.65):
    """flow_lenia.py:97-153: reintegration tracking.  Every cell's mass is a square of half-width sigma moved by
    clip(dt * flow); a destination collects the overlap with the squares of its (2dd+1)^2 rolled neighbours.  Positions
    are ABSOLUTE (x + 0.5), so a source rolled in across the world edge never overlaps: the flow is not periodic."""
    B, C, H, W = state.shape
    mx, my = torch.meshgrid(torch.arange(W, device=state.device), torch.arange(H, device=state.device), indexing="ij")
    pos = (torch.dstack((mx, my)) + 0.5).to(state.dtype)             # (W,H,2), pos[x,y] = (x+.5, y+.5)
    flow_perm = flow.permute(0, 1, 4, 3, 2)                          # (B,C,W,H,2)
    state_perm = state.permute(0, 1, 3, 2)                           # (B,C,W,H)
    max_flow = dd - sigma
    mu = pos[None, None] + torch.clip(dt * flow_perm, -max_flow, max_flow)
    parts = []
    for dx in range(-dd, dd + 1):
        for dy in range(-dd, dd + 1):
            st_r = torch.roll(state_perm, shifts=(dx, dy), dims=(2, 3))
            mu_r = torch.roll(mu, shifts=(dx, dy), dims=(2, 3))
            dist = torch.abs(mu_r - pos[None, None])
            amount = 0.5 + (sigma - dist)
            area = torch.clip(amount, 0, min(2 * sigma, 1)).prod(dim=-1)
            parts.append(st_r * (area / (4 * sigma**2)))
    return torch.stack(parts, dim=0).sum(dim=0).permute(0, 1, 3, 2)


def flow_step(state, K, mu, sigma, weights, dt=0.1, dd=2, sigma_rt=0.65, theta_x=2.0, n=2, Kf=None):
    """FlowLenia.step (flow_lenia.py:246-253, food off): Lenia affinity -> flow field -> reintegration tracking."""
    aff = lenia_affinity(state, K, mu, sigma, weights, Kf)
    return flow_reintegrate(state, flow_field(state, aff, theta_x, n), dt, dd, sigma_rt)


# float64 direct evaluation, for noise-floor reporting (not a restatement of reference lines)
def lenia_affinity_f64(state, K, mu, sigma, weights):
    s64 = state.double()
    K64, mu64, sg64, w64 = K.double(), mu.double(), sigma.double(), weights.double()
    return lenia_affinity(s64, K64, mu64, sg64, w64)


# --------------------------------------------------------------------------- #
# Neural CA inference
# --------------------------------------------------------------------------- #
def sobel_kernels(n_states: int) -> torch.Tensor:
    """(2n,1,3,3): n copies of sobel-x/8 then n copies of sobel-y/8 (nca.py:316-325)."""
    sx = torch.tensor([[-1, 0, 1], [-2, 0, 2], [-1, 0, 1]], dtype=torch.float)[None, None] / 8.0
    sy = torch.tensor([[-1, -2, -1], [0, 0, 0], [1, 2, 1]], dtype=torch.float)[None, None] / 8.0
    return torch.cat([sx.expand(n_states, -1, -1, -1), sy.expand(n_states, -1, -1, -1)], dim=0)


def nca_perception(state):
    """(B,n,H,W) -> (B,3n,H,W): grouped circular sobel conv then the state (nca.py:328-344).
    With groups=n and 2n output channels, output o reads input channel o//2."""
    n = state.shape[1]
    kern = sobel_kernels(n).to(device=state.device, dtype=state.dtype)
    p = F.pad(state, (1, 1, 1, 1), mode="circular")
    return torch.cat([F.conv2d(p, kern, groups=n), state], dim=1)


def nca_live_mask(state):
    """nca.py:226-237."""
    a = F.pad(state[:, 3:4], (1, 1, 1, 1), mode="circular")
    return F.max_pool2d(a, kernel_size=3, stride=1) > 0.1


def nca_step(state, W1, b1, W2, b2, update_mask):
    """nca.py:239-263 for one step; `update_mask` (B,1,H,W) bool replaces the in-step
    ``torch.rand(B,1,H,W) <= 0.5`` (nca.py:256) so both sides consume the same draw."""
    feat = nca_perception(state).permute(0, 2, 3, 1)
    hid = F.relu(F.linear(feat, W1, b1))
    dx = F.linear(hid, W2, b2).permute(0, 3, 1, 2)
    alive0 = nca_live_mask(state)
    new = torch.sigmoid(dx) * update_mask
    alive = alive0 & nca_live_mask(new)
    return torch.clamp(new * alive, 0, 1)


def nca_draw_mask(B, H, W, seed):
    """The mask the reference would draw with ``torch.manual_seed(seed)`` set right before
    ``forward`` (nca.py:256): one ``torch.rand(B,1,H,W)`` call on the CPU generator."""
    g = torch.Generator().manual_seed(seed)
    return torch.rand(B, 1, H, W, generator=g) <= 0.5


# --------------------------------------------------------------------------- #
# draw() / worldmap
# --------------------------------------------------------------------------- #
def frame_u8(worldmap: torch.Tensor) -> np.ndarray:
    """automaton.py:216: (3,H,W) float -> (W,H,3) uint8 (fp32 product, truncating cast)."""
    return (255 * worldmap.permute(2, 1, 0)).detach().cpu().numpy().astype(dtype=np.uint8)


def gol_draw(world, worldmap, decay_speed, highlight_color):
    """ca2d.py:121-124: echo of the previous frame fading by decay_speed * highlight_color, plus the live cells."""
    echo = torch.clamp(worldmap - decay_speed * highlight_color[:, None, None], min=0, max=1)
    w = torch.as_tensor(world)
    return torch.clamp(w[None, :, :].expand(3, -1, -1).to(dtype=torch.float) + echo, min=0, max=1)


def lenia_draw(state, batch_index=0, food=None):
    """lenia.py:487-500 / macelenia.py:371-385.  C == 2 shows (c0, c1, 0) (the reference concatenates a 2-plane zero
    tensor there, which yields 4 planes; the three-plane reading of its own comment is used)."""
    toshow = state[batch_index].clone()
    C = toshow.shape[0]
    if C == 1:
        toshow = toshow.expand(3, -1, -1).clone()
    elif C == 2:
        toshow = torch.cat([toshow, torch.zeros_like(toshow[:1])], dim=0)
    else:
        toshow = toshow[:3]
    if food is not None:
        toshow = toshow + food[batch_index]
    return torch.clamp(toshow, 0.0, 1.0)


def nca_draw(state):
    """nca.py:79-82: clamp(ARGB = state[:, :4]) then RGB * A (black background)."""
    pic = torch.clamp(state[:, :4].squeeze(0), 0, 1)
    return pic[:3] * pic[3:]
