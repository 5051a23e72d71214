"""Registry override for a real PyCA checkout: `import pyca_b200.dropin; pyca_b200.dropin.install()` before
`MainWindow(...)` in `simulate.py` replaces the `step()` of CA2D / Lenia / MaCELenia / MaCELeniaXChan / FlowLenia / NeuralCA by the
B200 kernels and leaves everything else of the reference classes (widgets, event handlers, draw, save/load) untouched.

`AUTOMATAS` is keyed by class name and the last definition wins (pyca/automata/automaton.py:104-111), so defining
subclasses with the SAME names after `pyca` has been imported is all the registration there is.

The subclasses keep the reference's attribute contract without properties (Lenia registers `state` as an nn.Module
buffer, lenia.py:79, which rules a property out): `step()` reads the public tensor, decides whether it is still the
tensor the previous step produced (identity + version counter), reloads it if not, and assigns the new tensor back.
"""
from __future__ import annotations

import torch

from . import _lib
from .ca2d import PackedWorld
from .lenia import FramedState, conv_garbi, garbi_tables, pick_conv_algo
from .nca import NCAWeights, nca_step

F = _lib.FRAME


class _Tracked:
    """Remembers the tensor a step produced so the next step can tell whether the user replaced / edited it."""

    def __init__(self):
        self.tensor = None
        self.version = -1

    def same(self, t):
        return t is self.tensor and t._version == self.version

    def set(self, t):
        self.tensor, self.version = t, t._version


def _fast_worldmap(base):
    """`worldmap` (automaton.py:209-216) for a reference subclass: the float (3,H,W) `_worldmap` its own draw() left on the
    GPU is turned into the (W,H,3) uint8 frame there and copied as bytes (3 B/cell instead of 12 + a host-side cast)."""
    from .automaton import FrameBuffer

    def worldmap(self):
        wm = self._worldmap
        if not (torch.is_tensor(wm) and wm.is_cuda and wm.ndim == 3 and wm.shape[0] == 3):
            return base.worldmap.fget(self)
        fb = getattr(self, "_b200_fb", None)
        if fb is None or (fb.H, fb.W) != tuple(wm.shape[1:]) or fb.dev.device != wm.device:
            fb = FrameBuffer(wm.shape[1], wm.shape[2], wm.device)
            object.__setattr__(self, "_b200_fb", fb)
        fb.pack(wm)
        return fb.numpy()

    return property(worldmap)


def make_ca2d(base):
    class CA2D(base):
        __doc__ = base.__doc__

        def step(self):
            """ca2d.py:195-209 on the bit-packed kernel; `self.world` stays an int32 (H,W) tensor."""
            pw = getattr(self, "_b200_pw", None)
            if pw is None or (pw.H, pw.W) != tuple(self.world.shape):
                pw = self._b200_pw = PackedWorld(self.world.shape[0], self.world.shape[1], self.world.device)
                self._b200_trk = _Tracked()
            if not self._b200_trk.same(self.world):
                pw.pack_from(self.world)
            pw.step(int(self.s_num), int(self.b_num), 1)
            self.world = pw.unpack(torch.int32)
            self._b200_trk.set(self.world)

    CA2D.__name__ = CA2D.__qualname__ = "CA2D"
    if isinstance(getattr(base, "worldmap", None), property):
        CA2D.worldmap = _fast_worldmap(base)
    return CA2D


class _LeniaEngine:
    """Device-side state of one Lenia-family automaton built from the REFERENCE object's own tensors
    (`k`, `mu`, `sigma`, `weights`), so every kernel variant the reference can build (rings, k_arbi) is honoured."""

    def __init__(self, owner):
        st = owner.state
        B, C, H, W = st.shape
        self.fs = FramedState(B, C, H, W, st.device)
        self.trk = _Tracked()
        self.key = None
        self.aff = None
        self.L = _lib.lib()

    def prepare(self, owner):
        key = (id(owner.k), owner.k._version, id(getattr(owner, "mu", None)), id(getattr(owner, "sigma", None)), id(owner.weights),
               id(getattr(owner, "params", None)))
        if key == self.key:
            return
        fs = self.fs
        dev = fs.device
        self.k_mult = int(getattr(owner, "k_mult", 1))
        self.garbi = None
        if getattr(owner, "g_arbi", False):
            pd = {k: (v.detach().to(dev) if torch.is_tensor(v) else v) for k, v in owner.params.param_dict.items()}
            self.garbi = garbi_tables(pd, fs.B, fs.C, self.k_mult)
        self.K = owner.k.detach().to(dev, torch.float32).contiguous()
        self.mu = owner.mu.detach().to(dev, torch.float32).contiguous()
        self.sigma = owner.sigma.detach().to(dev, torch.float32).contiguous()
        self.weights = owner.weights.detach().to(dev, torch.float32).contiguous()
        ks = self.K.shape[-1]
        with torch.cuda.device(dev):
            self.kprep = torch.empty(self.L.b200ca_lenia_kprep_elems(fs.B, fs.C, self.k_mult), dtype=torch.float32, device=dev)
            _lib.call("b200ca_lenia_prepare_kernel", _lib.ptr(self.K), _lib.ptr(self.kprep), fs.B, fs.C, self.k_mult, ks,
                      _lib.current_stream(dev))
            rows_ok = self.L.b200ca_lenia_fft_length(fs.H, fs.W) > 0 and self.garbi is None
            march_ok = self.L.b200ca_lenia_march_supported(fs.H, fs.W) > 0 and self.garbi is None
            self.algo = "direct" if self.garbi is not None else pick_conv_algo(fs.B, fs.C * self.k_mult, fs.H, fs.W, march_ok, rows_ok, ks)
            self.use_fft = self.algo != "direct"
            if self.algo == "fft_rows":
                self.kfft = torch.empty(self.L.b200ca_lenia_fft_kernel_elems(fs.B, fs.C, self.k_mult, fs.H, fs.W), dtype=torch.float32,
                                        device=dev)
                _lib.call("b200ca_lenia_fft_prepare_kernel", _lib.ptr(self.K), _lib.ptr(self.kfft), fs.B, fs.C, self.k_mult, ks,
                          fs.H, fs.W, _lib.current_stream(dev))
                self.ws = torch.empty(self.L.b200ca_lenia_fft_workspace_elems(fs.B, fs.C, self.k_mult, fs.H, fs.W),
                                      dtype=torch.float32, device=dev)
            elif self.algo == "march":
                self.kmarch = torch.empty(self.L.b200ca_lenia_march_kernel_elems(fs.B, fs.C, self.k_mult), dtype=torch.float32, device=dev)
                _lib.call("b200ca_lenia_march_prepare_kernel", _lib.ptr(self.K), _lib.ptr(self.kmarch), fs.B, fs.C, self.k_mult, ks,
                          _lib.current_stream(dev))
        self.key = key

    def sync_state(self, owner):
        if not self.trk.same(owner.state):
            self.fs.load_dense(owner.state)

    def conv(self, dst, mode, dt):
        fs = self.fs
        with torch.cuda.device(fs.device):
            st = _lib.current_stream(fs.device)
            if self.garbi is not None:
                conv_garbi(fs, dst, self.kprep, self.weights, self.garbi, self.k_mult, dt, mode)
            elif self.algo == "march":
                _lib.call("b200ca_lenia_step_march", _lib.ptr(fs.buf), _lib.ptr(dst), _lib.ptr(self.kmarch), _lib.ptr(self.mu),
                          _lib.ptr(self.sigma), _lib.ptr(self.weights), fs.B, fs.C, self.k_mult, fs.H, fs.W, float(dt), mode, 1, 0, 0, st)
            elif self.algo == "fft_rows":
                _lib.call("b200ca_lenia_step_fft", _lib.ptr(fs.buf), _lib.ptr(dst), _lib.ptr(self.kfft), _lib.ptr(self.mu),
                          _lib.ptr(self.sigma), _lib.ptr(self.weights), _lib.ptr(self.ws), fs.B, fs.C, self.k_mult, fs.H, fs.W,
                          float(dt), mode, 1, st)
            else:
                _lib.call("b200ca_lenia_step", _lib.ptr(fs.buf), _lib.ptr(dst), _lib.ptr(self.kprep), _lib.ptr(self.mu),
                          _lib.ptr(self.sigma), _lib.ptr(self.weights), fs.B, fs.C, self.k_mult, fs.H, fs.W, float(dt), mode, 1, 0, 0, st)

    def publish(self, owner):
        self.fs.swap()
        owner.state = self.fs.view()
        self.trk.set(owner.state)


def _engine(owner) -> _LeniaEngine:
    eng = getattr(owner, "_b200_eng", None)
    if eng is None or (eng.fs.B, eng.fs.C, eng.fs.H, eng.fs.W) != tuple(owner.state.shape):
        eng = _LeniaEngine(owner)
        object.__setattr__(owner, "_b200_eng", eng)
    eng.prepare(owner)
    eng.sync_state(owner)
    return eng


def make_lenia(base):
    class Lenia(base):
        __doc__ = base.__doc__

        @torch.no_grad()
        def step(self):
            """lenia.py:430-451."""
            eng = _engine(self)
            eng.conv(eng.fs.other, _lib.MODE_LENIA, self.dt)
            eng.publish(self)
            self.frames += 1

    Lenia.__name__ = Lenia.__qualname__ = "Lenia"
    if isinstance(getattr(base, "worldmap", None), property):
        Lenia.worldmap = _fast_worldmap(base)
    return Lenia


def _mace_step(self, alpha_on, alpha):
    eng = _engine(self)
    fs = eng.fs
    if eng.aff is None:
        eng.aff = torch.zeros_like(fs.buf)
    eng.conv(eng.aff, _lib.MODE_AFF, 0.0)
    with torch.cuda.device(fs.device):
        _lib.call("b200ca_mace_redistribute", _lib.ptr(fs.buf), _lib.ptr(eng.aff), _lib.ptr(fs.other), fs.B, fs.C, fs.H, fs.W,
                  float(self.b), 1 if alpha_on else 0, float(alpha), 1, _lib.current_stream(fs.device))
    eng.publish(self)
    self.Aff = eng.aff[:, :, F:F + fs.H, F:F + fs.W]


def make_mace(base):
    class MaCELenia(base):
        __doc__ = base.__doc__

        @torch.no_grad()
        def step(self, sense_food=None):
            """macelenia.py:72-87 (food off: `has_food` automata keep the reference path)."""
            if getattr(self, "has_food", False):
                return base.step(self, sense_food)
            _mace_step(self, False, 0.0)
            self.frames += 1

    MaCELenia.__name__ = MaCELenia.__qualname__ = "MaCELenia"
    if isinstance(getattr(base, "worldmap", None), property):
        MaCELenia.worldmap = _fast_worldmap(base)
    return MaCELenia


def make_xchan(base):
    class MaCELeniaXChan(base):
        __doc__ = base.__doc__

        @torch.no_grad()
        def step(self):
            """mace_lenia_xchan.py:59-66."""
            if getattr(self, "has_food", False):
                return base.step(self)
            _mace_step(self, True, float(self.alpha))

    MaCELeniaXChan.__name__ = MaCELeniaXChan.__qualname__ = "MaCELeniaXChan"
    if isinstance(getattr(base, "worldmap", None), property):
        MaCELeniaXChan.worldmap = _fast_worldmap(base)
    return MaCELeniaXChan


def make_flow(base):
    class FlowLenia(base):
        __doc__ = base.__doc__

        @torch.no_grad()
        def step(self):
            """flow_lenia.py:246-253 (food off: `has_food` automata keep the reference path)."""
            if getattr(self, "has_food", False):
                return base.step(self)
            eng = _engine(self)
            fs = eng.fs
            if eng.aff is None:
                eng.aff = torch.zeros_like(fs.buf)
            eng.conv(eng.aff, _lib.MODE_AFF, 0.0)
            with torch.cuda.device(fs.device):
                _lib.call("b200ca_flow_reintegrate", _lib.ptr(fs.buf), _lib.ptr(eng.aff), _lib.ptr(fs.other), fs.B, fs.C, fs.H, fs.W,
                          float(self.dt), int(self.dd), float(self.sigma_rt), float(self.theta_x), int(self.n),
                          _lib.current_stream(fs.device))
            eng.publish(self)
            self.frames += 1

    FlowLenia.__name__ = FlowLenia.__qualname__ = "FlowLenia"
    if isinstance(getattr(base, "worldmap", None), property):
        FlowLenia.worldmap = _fast_worldmap(base)
    return FlowLenia


def make_nca(base):
    class NeuralCA(base):
        __doc__ = base.__doc__

        @torch.no_grad()
        def step(self):
            """nca.py:67-72: same torch.rand draw as NCAModule.forward (nca.py:256), fused CUDA update."""
            comp = self.model.computer
            key = (id(self.model), comp[0].weight._version)
            if getattr(self, "_b200_key", None) != key:
                self._b200_w = NCAWeights(comp[0].weight, comp[0].bias, comp[2].weight, comp[2].bias, self.state.device)
                self._b200_key = key
            st = self.state
            B, _, H, W = st.shape
            mask = torch.rand(B, 1, H, W, device=st.device) <= 0.5
            impl = _lib.NCA_TC
            self.state = nca_step(st, self._b200_w, update_mask=mask, impl=impl)

    NeuralCA.__name__ = NeuralCA.__qualname__ = "NeuralCA"
    if isinstance(getattr(base, "worldmap", None), property):
        NeuralCA.worldmap = _fast_worldmap(base)
    return NeuralCA


def install():
    """Re-registers the five models in PyCA's AUTOMATAS with B200 `step()`s.  Needs an importable `pyca`."""
    from pyca.automata import models as ref_models  # noqa: F401  (fills AUTOMATAS)
    from pyca.automata.automaton import AUTOMATAS

    made = {}
    for name, factory in [("CA2D", make_ca2d), ("Lenia", make_lenia), ("MaCELenia", make_mace), ("MaCELeniaXChan", make_xchan),
                          ("FlowLenia", make_flow), ("NeuralCA", make_nca)]:
        if name in AUTOMATAS:
            made[name] = factory(AUTOMATAS[name])  # __init_subclass__ re-registers under the same key
    return made
