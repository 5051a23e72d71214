"""B200 drop-in for PyCA's ``CA2D`` (pyca/automata/models/ca2d.py).

Same constructor, attributes and methods; ``step()`` (ca2d.py:195-209) runs the bit-packed CUDA
kernel through the C ABI.  The public ``world`` tensor of the reference -- an (H,W) int tensor that
``draw()`` reads and the brush handlers mutate in place (ca2d.py:121-124, :153-159) -- is kept as a
lazily materialised view of the packed state:

* reading ``self.world`` unpacks the current bits into an int32 tensor (cached until the next step);
* assigning ``self.world = t`` or mutating the cached tensor in place (detected through the tensor's
  version counter) makes the next ``step()`` re-pack it first.
"""
from __future__ import annotations

import colorsys

import torch

from . import _lib
from .automaton import Automaton

_ELEM = {torch.uint8: 1, torch.bool: 1, torch.int8: 1, torch.int16: 2, torch.int32: 4, torch.int64: 8}


def rule_to_mask(rule) -> int:
    """'23' -> 0b1100 (ca2d.py:80-88)."""
    if isinstance(rule, str):
        return sum(2 ** int(d) for d in rule) if rule != "" else 0
    return int(rule)


class PackedWorld:
    """Device-resident bit-packed (H,W) world with ping-pong buffers."""

    def __init__(self, H, W, device):
        self.H, self.W = int(H), int(W)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.B200CAError(f"CA2D needs a CUDA device (got {device}); there is no CPU path")
        self.wpr = _lib.lib().b200ca_gol_words_per_row(self.W)
        self.stride = (self.wpr + 3) // 4 * 4
        self.bufs = [torch.zeros((self.H, self.stride), dtype=torch.int32, device=self.device) for _ in range(2)]
        self.cur = 0

    @property
    def bits(self):
        return self.bufs[self.cur]

    def pack_from(self, world: torch.Tensor):
        if tuple(world.shape) != (self.H, self.W):
            raise _lib.B200CAError(f"world shape {tuple(world.shape)} != {(self.H, self.W)}")
        w = world.to(self.device)
        if w.dtype not in _ELEM:
            w = w.to(torch.int32)
        w = w.contiguous()
        with _lib.device_guard(self.device):
            _lib.call("b200ca_gol_pack", _lib.ptr(w), _ELEM[w.dtype], _lib.ptr(self.bits), self.H, self.W, self.stride,
                      _lib.current_stream(self.device))

    def unpack(self, dtype=torch.int32) -> torch.Tensor:
        out = torch.empty((self.H, self.W), dtype=dtype, device=self.device)
        with _lib.device_guard(self.device):
            _lib.call("b200ca_gol_unpack", _lib.ptr(self.bits), _lib.ptr(out), _ELEM[dtype], self.H, self.W, self.stride,
                      _lib.current_stream(self.device))
        return out

    def step(self, s_mask, b_mask, nsteps=1, vert_open=False):
        with _lib.device_guard(self.device):
            a, b = self.bufs[self.cur], self.bufs[self.cur ^ 1]
            _lib.call("b200ca_gol_run", _lib.ptr(a), _lib.ptr(b), self.H, self.W, self.stride, s_mask, b_mask,
                      1 if vert_open else 0, nsteps, _lib.current_stream(self.device))
        if nsteps & 1:
            self.cur ^= 1

    def population(self) -> int:
        cnt = torch.zeros(1, dtype=torch.int64, device=self.device)
        with _lib.device_guard(self.device):
            _lib.call("b200ca_gol_population", _lib.ptr(self.bits), self.H, self.W, self.stride, _lib.ptr(cnt),
                      _lib.current_stream(self.device))
        return int(cnt.item())

    def random_fill(self, seed: int, row0: int = 0):
        with _lib.device_guard(self.device):
            _lib.call("b200ca_gol_random_fill", _lib.ptr(self.bits), self.H, self.W, self.stride, seed, row0,
                      _lib.current_stream(self.device))


class CA2D(Automaton):
    """
    2D Cellular Automaton, with outer-holistic rules and binary values (B200 bit-packed step).
    """

    def __init__(self, size, s_num="23", b_num="3", dot=False, device="cuda"):
        super().__init__(size, device)
        self.s_num = self.get_num_from_rule(s_num)
        self.b_num = self.get_num_from_rule(b_num)
        self.dot = dot
        self.device = device
        self._packed = PackedWorld(self.h, self.w, device)
        self._world = None       # cached unpacked tensor (or a user-assigned one)
        self._world_version = -1
        self._dirty = False      # cached/assigned tensor is newer than the packed bits
        self.reset()
        self.change_highlight_color()
        self.decay_speed = 0.1
        self._worldmap = self._worldmap.to(device)

    # ---- the public world tensor (ca2d.py:29) ------------------------------------------------
    @property
    def world(self) -> torch.Tensor:
        if self._world is None:
            self._world = self._packed.unpack(torch.int32)  # the reference's world is int32 after a step (ca2d.py:209)
            self._world_version = self._world._version
            self._dirty = False
        return self._world

    @world.setter
    def world(self, value: torch.Tensor):
        self._world = value
        self._world_version = value._version
        self._dirty = True

    def _sync_to_device(self):
        if self._world is not None and (self._dirty or self._world._version != self._world_version):
            self._packed.pack_from(self._world)
        self._dirty = False

    # ---- reference API ------------------------------------------------------------------------
    def get_num_from_rule(self, num):
        return rule_to_mask(num)

    def get_rule_from_num(self, num: int):
        return "".join(str(i) for i in range(9) if (num >> i) & 1)

    def change_highlight_color(self):
        hue = torch.rand(1).item()
        self.highlight_color = torch.tensor(colorsys.hls_to_rgb(hue, 0.5, 0.5), dtype=torch.float, device=self.device)

    def reset(self):
        """ca2d.py:103-115."""
        self._worldmap = torch.zeros((3, self.h, self.w), device=self.device)
        if not self.dot:
            self.world = self.get_init_mat(0.5)
        else:
            w = torch.zeros((self.h, self.w), dtype=torch.int, device=self.device)
            w[self.h // 2 - 1: self.h // 2 + 1, self.w // 2 - 1: self.w // 2 + 1] = torch.randint(0, 2, (2, 2), device=self.device)
            self.world = w

    def get_init_mat(self, rand_portion):
        """Centred random square of side portion*size, int16 (ca2d.py:227-246)."""
        size = torch.tensor([self.h, self.w])
        randsize = (size * rand_portion).to(dtype=torch.int16)
        randstarts = (size * (1 - rand_portion) / 2).to(dtype=torch.int16)
        randsquare = torch.where(torch.randn(*randsize.tolist()) > 0, 1, 0)
        init_mat = torch.zeros((self.h, self.w), dtype=torch.int16)
        init_mat[randstarts[0]: randstarts[0] + randsize[0], randstarts[1]: randstarts[1] + randsize[1]] = randsquare
        return init_mat.to(torch.int16).to(self.device)

    def change_num(self, s_num, b_num):
        self.s_num = self.get_num_from_rule(s_num)
        self.b_num = self.get_num_from_rule(b_num)
        self.reset()

    def step(self, nsteps: int = 1):
        """One generation (ca2d.py:195-209) -- `nsteps` generations in one call as an extension."""
        self._sync_to_device()
        self._packed.step(self.s_num, self.b_num, nsteps)
        self._world = None

    def draw(self):
        """ca2d.py:117-124, fused: one kernel reads the packed bits and the previous `_worldmap` (the echo), rewrites
        `_worldmap` in place and emits the (W,H,3) uint8 frame that `worldmap` returns."""
        self._sync_to_device()
        wm = self._worldmap
        if not (wm.is_cuda and wm.dtype == torch.float32 and wm.is_contiguous() and tuple(wm.shape) == (3, self.h, self.w)):
            wm = self._worldmap = wm.to(self.device, torch.float32).contiguous()
        key = (id(self.highlight_color), self.highlight_color._version, self.decay_speed)
        if getattr(self, "_decay_key", None) != key:   # 3 floats to the host only when the colour or the speed changed
            self._decay = (self.decay_speed * self.highlight_color).tolist()  # fp32 product, as in the reference expression
            self._decay_key = key
        d = self._decay
        fb = self._frame()
        pw = self._packed
        with _lib.device_guard(pw.device):
            _lib.call("b200ca_draw_gol", _lib.ptr(pw.bits), _lib.ptr(wm), _lib.ptr(fb.dev), self.h, self.w, pw.stride,
                      d[0], d[1], d[2], _lib.current_stream(pw.device))
        fb.mark(wm)

    def name(self):
        return "2D Cellular Automaton"

    def get_string_state(self):
        return f"Rule: s:{self.get_rule_from_num(self.s_num)}, b:{self.get_rule_from_num(self.b_num)}"

    def get_brush_slice(self, x, y, radius=10):
        return (self.Y - y) ** 2 + (self.X - x) ** 2 < radius**2
