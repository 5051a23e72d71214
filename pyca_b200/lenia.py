"""B200 drop-ins for PyCA's Lenia family: ``Lenia`` (pyca/automata/models/lenia/lenia.py),
``MaCELenia`` (lenia/macelenia.py) and ``MaCELeniaXChan`` (lenia/mace_lenia_xchan.py).

Same constructor signatures, attributes (``state``, ``mu``, ``sigma``, ``weights``, ``k``, ``dt``,
``b``, ``alpha``, ``frames``) and methods (``step``, ``draw``, ``update_params``, ``mass``,
``total_mass``).  Cold code -- parameter loading, sanitising, kernel construction -- stays torch on
the host side (``lenia_params.py``); ``step()`` is CUDA through the C ABI:

* Lenia.step (lenia.py:430-451)                  -> b200ca_lenia_step(mode=LENIA)
* MaCELenia.step (macelenia.py:72-120)           -> b200ca_lenia_step(mode=AFF) + b200ca_mace_redistribute
* MaCELeniaXChan.step (mace_lenia_xchan.py:59-76)-> same with alpha_on

State ownership: the reference exposes ``self.state`` (B,C,H,W) as a public tensor that event handlers
mutate in place between steps (lenia.py:681-690).  Here the state lives in two FRAMED device buffers
(interior + 16-cell periodic frame, see include/b200ca.h); ``self.state`` is a strided view of the
current buffer's interior.  In-place edits are detected through the view's version counter (the frame
is refreshed before the next step); assigning a new tensor makes the next step load it.
"""
from __future__ import annotations

from pathlib import Path

import torch

from . import _lib
from .automaton import Automaton
from .lenia_params import LeniaParams, compute_kernel

F = _lib.FRAME


def garbi_tables(p, B, C, k_mult):
    """Growth functions as Fourier series (lenia.py:390-418): coefficient tables for b200ca_lenia_step_garbi.  The argument
    of harmonic h is  (2 pi / period) * h * (u - 0)  on the period (0, 2), built in fp32 in the reference's order of
    operations (funcgen.py:99-104).  `p`: parameter dict with g_coeffs (B,C*k_mult,C,2*nh) [, g_harmonics, g_rescale, g_clip]."""
    if "g_coeffs" not in p:
        raise _lib.B200CAError("g_coeffs not in params, but g_arbi is True")
    n = B * C * k_mult * C
    co = p["g_coeffs"].float().reshape(n, -1, 2)
    nh = co.shape[1]
    if "g_harmonics" in p and p["g_harmonics"] is not None:
        harm = p["g_harmonics"].float().reshape(n, nh)
    else:
        harm = torch.arange(nh, device=co.device).float()[None].expand(n, -1)
    period = torch.full((n, 1), 2.0, device=co.device)
    freq = (2 * torch.pi / period * harm).contiguous()
    rescale, clip = p.get("g_rescale"), p.get("g_clip")
    return {"cos": co[:, :, 0].contiguous(), "sin": co[:, :, 1].contiguous(), "freq": freq, "nh": nh,
            "rescale": None if rescale is None else (float(rescale[0]), float(rescale[1])),
            "clip": None if clip is None else float(clip),
            "stats": torch.zeros((n, 2), dtype=torch.float32, device=co.device)}


def pick_conv_algo(B, NS, H, W, march_ok=True, rows_ok=True, k_size=31):
    """conv_path='auto': which convolution kernel steps a (B, C, H, W) world with NS = C * k_mult source kernels per
    destination.  Measured on B200 (profiles/r02_conv_paths.md; all paths agree with the oracle to <= 1e-5):
      * one source (C = 1): the marching kernel (lenia_march.cu; the spectra never leave shared memory, one balanced wave of
        CTAs) takes every launch of more than 1536^2 cells (95 us at 4096^2 against 141 us for the hybrid and 104 us for the
        direct kernel at k_size 5; 1.27 ms at 16384^2); at 1024^2 the whole world is a single wave of CTAs and the two-kernel
        hybrid (lenia_fft.cu, 512-thread small-world variant) has the shorter serial chain: 19.7 vs 22.0 us;
      * several sources: the hybrid's fused kernel keeps the per-destination sum over sources in registers, the marching kernel
        accumulates it through the output planes, one launch per source: 61 vs 140 us at 1024^2 (C = 3), but 848 vs 784 us at
        4096^2 -- the marching kernel from 2560^2 cells up;
      * the direct convolution (lenia.cu) is FMA-bound at R = 15 (385 us at 4096^2) but evaluates only the radius class of the
        kernel: small kernels on small worlds (k_size <= 11 at 1024^2: 16.5-18.5 us); it also serves what the FFT paths do not
        (Fourier growth functions, worlds narrower than 32)."""
    del march_ok, rows_ok   # (the library applies the same support checks)
    return ("direct", "fft_rows", "march")[_lib.lib().b200ca_lenia_pick_algo_ks(int(B), int(NS), int(H), int(W), int(k_size))]


def conv_garbi(fs, dst, kprep, weights, g, k_mult, dt, mode):
    """b200ca_lenia_step_garbi on a FramedState (direct path; two passes when the growth is globally rescaled)."""
    rs, cl = g["rescale"], g["clip"]
    _lib.call("b200ca_lenia_step_garbi", _lib.ptr(fs.buf), _lib.ptr(dst), _lib.ptr(kprep), _lib.ptr(weights), _lib.ptr(g["cos"]),
              _lib.ptr(g["sin"]), _lib.ptr(g["freq"]), g["nh"], 0 if rs is None else 1, 0.0 if rs is None else rs[0],
              0.0 if rs is None else rs[1], 0 if cl is None else 1, 0.0 if cl is None else cl, _lib.ptr(g["stats"]), fs.B, fs.C,
              k_mult, fs.H, fs.W, float(dt), mode, 1, _lib.current_stream(fs.device))


class FramedState:
    """Two framed (B,C,H,W) fp32 device buffers with ping-pong and an interior view."""

    def __init__(self, B, C, H, W, device, wrap_rows=True):
        self.B, self.C, self.H, self.W = int(B), int(C), int(H), int(W)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.B200CAError(f"Lenia-family automata need a CUDA device (got {device}); there is no CPU path")
        L = _lib.lib()
        self.pitch = L.b200ca_frame_pitch(self.W)
        self.rows = self.H + 2 * F
        self.wrap_rows = wrap_rows
        self.bufs = [torch.zeros((self.B, self.C, self.rows, self.pitch), dtype=torch.float32, device=self.device)
                     for _ in range(2)]
        self.cur = 0
        self._view = None
        self._view_version = -1

    def _stream(self):
        return _lib.current_stream(self.device)

    @property
    def buf(self):
        return self.bufs[self.cur]

    @property
    def other(self):
        return self.bufs[self.cur ^ 1]

    def swap(self):
        self.cur ^= 1
        self._view = None

    def view(self) -> torch.Tensor:
        if self._view is None:
            self._view = self.buf[:, :, F:F + self.H, F:F + self.W]
            self._view_version = self._view._version
        return self._view

    def load_dense(self, dense: torch.Tensor):
        d = dense.to(device=self.device, dtype=torch.float32).contiguous()
        if tuple(d.shape) != (self.B, self.C, self.H, self.W):
            raise _lib.B200CAError(f"state shape {tuple(d.shape)} != {(self.B, self.C, self.H, self.W)}")
        with _lib.device_guard(self.device):
            _lib.call("b200ca_frame_load", _lib.ptr(d), _lib.ptr(self.buf), self.B, self.C, self.H, self.W,
                      1 if self.wrap_rows else 0, self._stream())
        self._view = None

    def refresh_if_edited(self):
        """Re-derives the frame when the interior view was mutated in place since the last step."""
        if self._view is not None and self._view._version != self._view_version:
            with _lib.device_guard(self.device):
                _lib.call("b200ca_frame_refresh", _lib.ptr(self.buf), self.B, self.C, self.H, self.W,
                          1 if self.wrap_rows else 0, self._stream())
            self._view_version = self._view._version

    def dense(self) -> torch.Tensor:
        return self.view().contiguous()

    def mass(self) -> torch.Tensor:
        out = torch.zeros((self.B, self.C), dtype=torch.float64, device=self.device)
        with _lib.device_guard(self.device):
            _lib.call("b200ca_frame_mass", _lib.ptr(self.buf), _lib.ptr(out), self.B, self.C, self.H, self.W, self._stream())
        return out


class Lenia(Automaton):
    """
    Multi-channel lenia automaton. A multi-colored GoL-inspired continuous automaton. Introduced by Bert Chan.
    (B200 step: direct TMA-tiled convolution fused with growth, channel sum, Euler step and clamp.)
    """

    def __init__(self, size, dt=0.1, num_channels=3, params=None, state_init=None, device="cuda",
                 interest_files=None, save_dir=".", conv_path="auto"):
        """conv_path: 'direct' (TMA-tiled folded spatial convolution), 'fft' (marching row-FFT x column-FIR kernel),
        'fft_rows' (the two-kernel row-FFT + fused FIR/inverse hybrid) or 'auto' (fft when the world shape supports it, else
        direct)."""
        self.conv_path = conv_path
        if len(size) == 2:
            size = (1,) + tuple(size)
        Automaton.__init__(self, size[1:], device)
        self.batch = size[0]
        self.h, self.w = size[1:]
        self.C = num_channels
        self.dt = dt
        if params is None:
            # random parameters with the reference's prior (lenia.py:68-70 -> LeniaParams.default_gen)
            params = LeniaParams.default_gen(batch_size=self.batch, num_channels=self.C, k_size=31)
        self.g_arbi = False
        self.k_arbi = False
        self._fs = FramedState(self.batch, self.C, self.h, self.w, device)
        self._pending_state = None
        self._init_family_buffers()
        self.update_params(params)
        if state_init is None:
            state_init = self._init_circle()
        self.state = state_init
        self.display_kernel = False
        self.save_dir = save_dir
        if interest_files is not None:
            self.interest_files = sorted(p.as_posix() for p in Path(interest_files).rglob("*.pt"))
        else:
            self.interest_files = None
        self.chosen_interesting = 0
        self.frames = 0
        self._worldmap = self._worldmap.to(device)

    def _init_family_buffers(self):
        pass

    def _init_circle(self, radius=None) -> torch.Tensor:
        """Initial state when none is given: noise inside a centred disc of radius 3 * k_size, zero outside -- the shape of
        the reference's `set_init_circle` (lenia.py:254-287).  The reference fills the disc with Perlin noise from the
        third-party `pyperlin` package (absent here, and not part of the step path); uniform noise stands in for it."""
        radius = self.k_size * 3 if radius is None else radius
        X, Y = torch.meshgrid(torch.linspace(-self.h // 2, self.h // 2, self.h), torch.linspace(-self.w // 2, self.w // 2, self.w),
                              indexing="ij")
        disc = torch.sqrt(X**2 + Y**2) < radius
        return torch.rand((self.batch, self.C, self.h, self.w)) * disc

    # ---- public state tensor ------------------------------------------------------------------
    @property
    def state(self) -> torch.Tensor:
        if self._pending_state is not None:
            return self._pending_state
        return self._fs.view()

    @state.setter
    def state(self, value: torch.Tensor):
        self._pending_state = value

    def _sync_state(self):
        if self._pending_state is not None:
            self._fs.load_dense(self._pending_state)
            self._pending_state = None
        else:
            self._fs.refresh_if_edited()

    # ---- parameters (cold path, lenia.py:143-205) ----------------------------------------------
    def update_params(self, params, k_size_override=None):
        self.frames = 0
        if isinstance(params, (str, Path)):
            params = LeniaParams(from_file=params, channels=self.C)
        elif isinstance(params, dict):
            d = dict(params)
            st = d.pop("state", None)
            params = LeniaParams(param_dict=d, channels=self.C)
            if st is not None:
                self.state = st
        elif not isinstance(params, LeniaParams) and hasattr(params, "param_dict"):
            params = LeniaParams(param_dict=params.param_dict, channels=self.C)  # e.g. the reference's own LeniaParams
        self.params = params.to(self.device)
        p = self.params.param_dict
        if p["weights"].shape[0] != self.batch:
            raise _lib.B200CAError(f"params batch {p['weights'].shape[0]} != automaton batch {self.batch}")
        if k_size_override is not None:
            p["k_size"] = k_size_override
        if p["k_size"] % 2 == 0:
            p["k_size"] += 1
        self.k_size = p["k_size"]
        self.k_mult = p.get("k_mult", 1)
        self.g_arbi = bool(p.get("g_arbi", False))
        self.k_arbi = bool(p.get("k_arbi", False))
        self.weights = p["weights"].float().contiguous()
        exp_shape = (self.batch, self.C * self.k_mult, self.C)
        if tuple(self.weights.shape) != exp_shape:
            raise _lib.B200CAError(f"weights shape {tuple(self.weights.shape)} != {exp_shape}")
        # files with Fourier growth functions may carry no mu / sigma at all
        self.mu = p["mu"].float().contiguous() if "mu" in p else torch.zeros(exp_shape, device=self.device)
        self.sigma = p["sigma"].float().contiguous() if "sigma" in p else torch.ones(exp_shape, device=self.device)
        for key in ("beta", "mu_k", "sigma_k"):
            if key in p:
                setattr(self, key, p[key])
        if tuple(self.mu.shape) != exp_shape:
            raise _lib.B200CAError(f"mu shape {tuple(self.mu.shape)} != {exp_shape}")
        self._garbi = None
        if self.g_arbi:
            self._garbi = garbi_tables(p, self.batch, self.C, self.k_mult)
        self.k = compute_kernel(self.params, self.C).float().contiguous()  # (B, C*k_mult, C, ks, ks)
        L = _lib.lib()
        n = L.b200ca_lenia_kprep_elems(self.batch, self.C, self.k_mult)
        self._kprep = torch.empty(n, dtype=torch.float32, device=self.device)
        with _lib.device_guard(self.device):
            _lib.call("b200ca_lenia_prepare_kernel", _lib.ptr(self.k), _lib.ptr(self._kprep), self.batch, self.C, self.k_mult,
                      int(self.k_size), _lib.current_stream(self.device))
        # FFT paths: "fft" = the marching kernel (lenia_march.cu: one kernel, spectra stay in shared memory; every world with
        # H, W >= 32), "fft_rows" = the row-FFT + fused FIR/inverse hybrid (lenia_fft.cu: H even, W >= 64; kept for A/B runs
        # and for the phase-split row slabs); "auto" = b200ca_lenia_pick_algo (a rule fitted to measurements).
        self._kfft = self._fft_ws = self._kmarch = None
        march_ok = L.b200ca_lenia_march_supported(self.h, self.w) > 0 and not self.g_arbi  # Fourier growth: direct path only
        rows_ok = L.b200ca_lenia_fft_length(self.h, self.w) > 0 and not self.g_arbi
        if self.conv_path not in ("auto", "direct", "fft", "fft_rows"):
            raise _lib.B200CAError(f"conv_path must be 'auto', 'direct', 'fft' or 'fft_rows' (got {self.conv_path!r})")
        if self.conv_path == "fft" and not march_ok:
            raise _lib.B200CAError(f"conv_path='fft' is not supported for a {self.h}x{self.w} world (H, W >= 32, Gaussian growth)")
        if self.conv_path == "fft_rows" and not rows_ok:
            raise _lib.B200CAError(f"conv_path='fft_rows' is not supported for a {self.h}x{self.w} world (H even >= 32, W >= 64)")
        if self.g_arbi:
            self.conv_algo = "direct"
        elif self.conv_path == "auto":
            self.conv_algo = pick_conv_algo(self.batch, self.C * self.k_mult, self.h, self.w, march_ok, rows_ok, self.k_size)
        else:
            self.conv_algo = {"fft": "march", "fft_rows": "fft_rows", "direct": "direct"}[self.conv_path]
        self.use_fft = self.conv_algo != "direct"
        if self.conv_algo == "march":
            self._kmarch = torch.empty(L.b200ca_lenia_march_kernel_elems(self.batch, self.C, self.k_mult), dtype=torch.float32,
                                       device=self.device)
            with _lib.device_guard(self.device):
                _lib.call("b200ca_lenia_march_prepare_kernel", _lib.ptr(self.k), _lib.ptr(self._kmarch), self.batch, self.C, self.k_mult,
                          int(self.k_size), _lib.current_stream(self.device))
        elif self.conv_algo == "fft_rows":
            self._kfft = torch.empty(L.b200ca_lenia_fft_kernel_elems(self.batch, self.C, self.k_mult, self.h, self.w),
                                     dtype=torch.float32, device=self.device)
            with _lib.device_guard(self.device):
                _lib.call("b200ca_lenia_fft_prepare_kernel", _lib.ptr(self.k), _lib.ptr(self._kfft), self.batch, self.C, self.k_mult,
                          int(self.k_size), self.h, self.w, _lib.current_stream(self.device))
            self._fft_ws = torch.empty(L.b200ca_lenia_fft_workspace_elems(self.batch, self.C, self.k_mult, self.h, self.w),
                                       dtype=torch.float32, device=self.device)

    # ---- hot path ---------------------------------------------------------------------------------
    def _conv(self, dst: torch.Tensor, mode: int):
        fs = self._fs
        with _lib.device_guard(fs.device):
            if self._garbi is not None:
                conv_garbi(fs, dst, self._kprep, self.weights, self._garbi, self.k_mult, self.dt, mode)
            elif self.conv_algo == "march":
                _lib.call("b200ca_lenia_step_march", _lib.ptr(fs.buf), _lib.ptr(dst), _lib.ptr(self._kmarch), _lib.ptr(self.mu),
                          _lib.ptr(self.sigma), _lib.ptr(self.weights), fs.B, fs.C, self.k_mult, fs.H, fs.W, float(self.dt), mode, 1,
                          0, 0, _lib.current_stream(fs.device))
            elif self.conv_algo == "fft_rows":
                _lib.call("b200ca_lenia_step_fft", _lib.ptr(fs.buf), _lib.ptr(dst), _lib.ptr(self._kfft), _lib.ptr(self.mu),
                          _lib.ptr(self.sigma), _lib.ptr(self.weights), _lib.ptr(self._fft_ws), fs.B, fs.C, self.k_mult, fs.H, fs.W,
                          float(self.dt), mode, 1, _lib.current_stream(fs.device))
            else:
                _lib.call("b200ca_lenia_step", _lib.ptr(fs.buf), _lib.ptr(dst), _lib.ptr(self._kprep), _lib.ptr(self.mu),
                          _lib.ptr(self.sigma), _lib.ptr(self.weights), fs.B, fs.C, self.k_mult, fs.H, fs.W, float(self.dt), mode,
                          1, 0, 0, _lib.current_stream(fs.device))

    @torch.no_grad()
    def step(self):
        """lenia.py:430-451."""
        self._sync_state()
        self._conv(self._fs.other, _lib.MODE_LENIA)
        self._fs.swap()
        self.frames += 1

    def mass(self):
        """(B,C) mean mass (lenia.py:471-479)."""
        self._sync_state()
        return (self._fs.mass() / (self.h * self.w)).float()

    def _draw_fused(self, batch_index=0, food=None):
        """One kernel: framed state -> float `_worldmap` (3,H,W) + the (W,H,3) uint8 frame `worldmap` returns."""
        self._sync_state()
        fs = self._fs
        wm = self._worldmap
        if not (wm.is_cuda and wm.dtype == torch.float32 and wm.is_contiguous() and tuple(wm.shape) == (3, self.h, self.w)):
            wm = self._worldmap = torch.empty((3, self.h, self.w), dtype=torch.float32, device=fs.device)
        fb = self._frame()
        with _lib.device_guard(fs.device):
            _lib.call("b200ca_draw_lenia", _lib.ptr(fs.buf), _lib.ptr(food), _lib.ptr(wm), _lib.ptr(fb.dev), int(batch_index),
                      fs.C, fs.H, fs.W, _lib.current_stream(fs.device))
        fb.mark(wm)

    @torch.no_grad()
    def draw(self):
        """lenia.py:481-500 (C == 2 shows (c0, c1, 0); the kernel overlay of `display_kernel` is GUI and not drawn)."""
        assert self._fs.B == 1, "Batch size must be 1 to draw"
        self._draw_fused(0)


class MaCELenia(Lenia):
    """
    Mass conserving Lenia-like Alife model (B200 step).
    """

    def __init__(self, size, num_channels=3, params=None, state_init=None, device="cuda", has_food=False, sense_food=False,
                 interest_files=None, save_dir=".", conv_path="auto"):
        if len(size) == 2:
            size = (1,) + tuple(size)
        self.has_food = has_food
        self.initial_food = 0.1 * size[1] * size[2]   # macelenia.py:37
        self.sense_food = sense_food
        self._beta = 8
        super().__init__(size, dt=0.1, num_channels=num_channels, params=params, state_init=state_init, device=device,
                         interest_files=interest_files, save_dir=save_dir, conv_path=conv_path)
        self.show_batch = 0
        self.cum_loss_mass = torch.zeros(self.batch, device=device)
        self.food_channel = torch.zeros((self.batch, 1, self.h, self.w), device=device)
        if self.has_food:
            self.food_channel = self.random_food_chan(food_amount=self.initial_food)

    def _init_family_buffers(self):
        fs = self._fs
        self._aff = torch.zeros((fs.B, fs.C, fs.rows, fs.pitch), dtype=torch.float32, device=fs.device)
        self._food_bufs = None  # framed mask planes, framed food convolution, per-world loss (allocated on first use)

    @property
    def b(self):
        return self._beta

    @b.setter
    def b(self, value):
        self._beta = value

    @property
    def Aff(self) -> torch.Tensor:
        """Affinity of the last step, (B,C,H,W) view (macelenia.py:52)."""
        return self._aff[:, :, F:F + self.h, F:F + self.w]

    # ---- food subsystem (macelenia.py:161-226, :301-345; off by default) -----------------------------------------
    def _food_scratch(self):
        if self._food_bufs is None:
            fs = self._fs
            self._food_bufs = {"mask": torch.zeros_like(self._aff), "conv": torch.zeros_like(self._aff),
                               "loss": torch.zeros(fs.B, dtype=torch.float64, device=fs.device)}
        return self._food_bufs

    def _food_dense(self) -> torch.Tensor:
        f = self.food_channel
        if not (f.is_cuda and f.dtype == torch.float32 and f.is_contiguous()):
            f = self.food_channel = f.to(self._fs.device, torch.float32).contiguous()
        return f

    def random_food_chan(self, food_amount, num_spots=300, food_size=5, add_to_exisitng=False, batches=[]):
        """macelenia.py:301-345: `num_spots` squares of side `food_size` at positions drawn from Python's `random`."""
        import random

        food_density = food_amount / (num_spots * food_size * food_size)
        places = [[random.randint(food_size, self.h - food_size), random.randint(food_size, self.w - food_size)]
                  for _ in range(num_spots)]
        if add_to_exisitng:
            food_chan = self.food_channel.clone()
        else:
            food_chan = torch.zeros((self.batch, 1, self.h, self.w), device=self.device)
        for y, x in places:
            ys, xs = slice(y - food_size // 2, y + food_size // 2 + 1), slice(x - food_size // 2, x + food_size // 2 + 1)
            if add_to_exisitng:
                food_chan[batches, :, ys, xs] = food_density
            else:
                food_chan[:, :, ys, xs] = food_density
        return food_chan

    def set_food(self, has_food: bool):
        """macelenia.py:246-258."""
        self.has_food = has_food
        if self.has_food:
            self.food_channel = self.random_food_chan(food_amount=self.initial_food)
        else:
            self.food_channel = torch.zeros_like(self.food_channel)

    def _compute_affinity(self, sense_food=False) -> torch.Tensor:
        """macelenia.py:122-159."""
        self._sync_state()
        self._conv(self._aff, _lib.MODE_AFF)
        if self.has_food and sense_food:
            # Aff += mask + sum_src K_src,j (*) mask, mask = (food > 0) on every channel (:149-157)
            fs, sc = self._fs, self._food_scratch()
            food = self._food_dense()
            mask = (food > 0).float().expand(-1, fs.C, -1, -1).contiguous()
            with _lib.device_guard(fs.device):
                st = _lib.current_stream(fs.device)
                _lib.call("b200ca_frame_load", _lib.ptr(mask), _lib.ptr(sc["mask"]), fs.B, fs.C, fs.H, fs.W, 1, st)
                _lib.call("b200ca_lenia_conv_sum", _lib.ptr(sc["mask"]), _lib.ptr(sc["conv"]), _lib.ptr(self._kprep), fs.B, fs.C,
                          self.k_mult, fs.H, fs.W, st)
                _lib.call("b200ca_mace_food_sense", _lib.ptr(self._aff), _lib.ptr(sc["conv"]), _lib.ptr(food), fs.B, fs.C, fs.H, fs.W, st)
        return self.Aff

    def _mace_step(self, alpha_on=False, alpha=0.0, sense_food=None):
        """_mace_step (+ _food_step, + the cross-channel step of the XChan variant): macelenia.py:72-120, :161-226."""
        fs = self._fs
        sense = self.sense_food if sense_food is None else sense_food
        self._compute_affinity(sense_food=sense)
        with _lib.device_guard(fs.device):
            st = _lib.current_stream(fs.device)
            if not self.has_food:
                _lib.call("b200ca_mace_redistribute", _lib.ptr(fs.buf), _lib.ptr(self._aff), _lib.ptr(fs.other), fs.B, fs.C, fs.H,
                          fs.W, float(self.b), 1 if alpha_on else 0, float(alpha), 1, st)
                fs.swap()
                return self.Aff
            # with food the reference runs redistribution -> decay (+ respawn) -> consume -> cross-channel step, so the
            # mixing moves from the redistribution kernel's epilogue into the consume kernel's
            sc = self._food_scratch()
            food = self._food_dense()
            _lib.call("b200ca_mace_redistribute", _lib.ptr(fs.buf), _lib.ptr(self._aff), _lib.ptr(fs.other), fs.B, fs.C, fs.H, fs.W,
                      float(self.b), 0, 0.0, 1, st)
            sc["loss"].zero_()
            _lib.call("b200ca_mace_food_decay", _lib.ptr(fs.other), _lib.ptr(sc["loss"]), fs.B, fs.C, fs.H, fs.W, 0.02, 0.0003, st)
            self._respawn_food(sc["loss"])
            food = self._food_dense()
            _lib.call("b200ca_mace_food_consume", _lib.ptr(fs.other), _lib.ptr(fs.buf), _lib.ptr(food), _lib.ptr(self._aff), fs.B, fs.C,
                      fs.H, fs.W, 0.3 if self.sense_food else 0.1, 0.1, 1 if alpha_on else 0, float(self.b), float(alpha), st)
        fs._view = None  # same buffer, new contents: hand out a fresh view object
        return self.Aff

    def _respawn_food(self, loss64):
        """macelenia.py:183-192: a world that has lost 200 units of mass since the last respawn gets them back as food."""
        food_amount = 200
        self.cum_loss_mass = self.cum_loss_mass.to(loss64.device) + loss64.float()
        update_idxs = [p[0] for p in torch.argwhere(self.cum_loss_mass >= food_amount).tolist()]
        if update_idxs:
            self.cum_loss_mass[update_idxs] = self.cum_loss_mass[update_idxs] - food_amount
            self.food_channel = self.random_food_chan(food_amount=food_amount, num_spots=5, food_size=7, add_to_exisitng=True,
                                                      batches=update_idxs)

    @torch.no_grad()
    def step(self, sense_food=None):
        """macelenia.py:72-87."""
        self._mace_step(sense_food=sense_food)
        self.frames += 1

    def total_mass(self):
        """(B,) total mass (macelenia.py:421-428; the reference returns world 0's scalar when food is on)."""
        self._sync_state()
        m = self._fs.mass().sum(dim=1)
        if self.has_food:
            return (m + self.food_channel.double().sum(dim=(1, 2, 3)))[0]
        return m

    @torch.no_grad()
    def draw(self):
        """macelenia.py:365-385 (single-world view `show_batch`; the gridified `show_all` view is GUI)."""
        food = self._food_dense()[self.show_batch] if self.has_food else None
        self._draw_fused(self.show_batch, food=food)


class MaCELeniaXChan(MaCELenia):
    """
    MaCELenia with cross channel step (B200 step).
    """

    def __init__(self, size, num_channels=3, params=None, state_init=None, device="cuda", has_food=False, sense_food=False,
                 interest_files=None, save_dir=".", conv_path="auto"):
        super().__init__(size, num_channels, params, state_init, device=device, has_food=has_food, sense_food=sense_food,
                         interest_files=interest_files, save_dir=save_dir, conv_path=conv_path)
        # the reference constructor resets these after update_params ran (mace_lenia_xchan.py:47-49)
        self._beta = 6
        self.alpha = 0.03
        self.params["alpha"] = self.alpha

    def update_params(self, params, k_size_override=None):
        super().update_params(params, k_size_override=k_size_override)
        if "alpha" in self.params:
            self.alpha = float(self.params["alpha"])

    @torch.no_grad()
    def step(self):
        """mace_lenia_xchan.py:59-66: MaCE step, food step, cross-channel step (the reference does not advance `frames`)."""
        self._mace_step(alpha_on=True, alpha=self.alpha)


class FlowLenia(Lenia):
    """Pytorch port of mass conserving FlowLenia (B200 step: Lenia affinity kernels + fused flow / reintegration kernel)."""

    def __init__(self, size, dt=0.1, num_channels=3, params=None, state_init=None, device="cuda", dd=2, sigma_rt=0.65,
                 has_food=False, interest_files=None, save_dir=".", conv_path="auto"):
        if has_food:
            raise _lib.B200CAError("FlowLenia's experimental food mechanics (flow_lenia.py:230-244) are not built")
        self.dd = dd
        self.sigma_rt = sigma_rt
        self.theta_x = 2
        self.n = 2
        self.has_food = False
        super().__init__(size, dt, num_channels, params, state_init, device=device, interest_files=interest_files,
                         save_dir=save_dir, conv_path=conv_path)

    def _init_family_buffers(self):
        fs = self._fs
        self._aff = torch.zeros((fs.B, fs.C, fs.rows, fs.pitch), dtype=torch.float32, device=fs.device)

    @property
    def Aff(self) -> torch.Tensor:
        """Pre-flow affinity of the last step, (B,C,H,W) view."""
        return self._aff[:, :, F:F + self.h, F:F + self.w]

    @torch.no_grad()
    def step(self):
        """flow_lenia.py:246-253."""
        fs = self._fs
        self._sync_state()
        self._conv(self._aff, _lib.MODE_AFF)
        with _lib.device_guard(fs.device):
            _lib.call("b200ca_flow_reintegrate", _lib.ptr(fs.buf), _lib.ptr(self._aff), _lib.ptr(fs.other), fs.B, fs.C, fs.H, fs.W,
                      float(self.dt), int(self.dd), float(self.sigma_rt), float(self.theta_x), int(self.n),
                      _lib.current_stream(fs.device))
        fs.swap()
        self.frames += 1

    def total_mass(self):
        self._sync_state()
        return self._fs.mass().sum(dim=1)
