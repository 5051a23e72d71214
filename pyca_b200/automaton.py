"""Headless mirror of PyCA's ``Automaton`` plugin base (pyca/automata/automaton.py:13-304).

Same contract as the reference class -- ``step()``, ``draw()``, ``_worldmap`` (3,H,W) float,
``worldmap`` (W,H,3) uint8, the ``AUTOMATAS`` registry filled by ``__init_subclass__``
(automaton.py:104-111), ``name()``, ``get_string_state()`` -- without the pygame / pygame_gui
dependency, so the classes run on a GPU box that has no display stack.  The GUI hooks
(``process_event``, ``process_gui_change``, ``register_component`` ...) exist as no-ops; when the real
PyCA is importable, ``pyca_b200.dropin.install()`` builds subclasses of the reference's own classes
instead, so ``simulate.py`` keeps its widgets and only ``step()`` changes.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib

AUTOMATAS = {}


class FrameBuffer:
    """Device (W,H,3) uint8 frame + its pinned host mirror: what ``Automaton.worldmap`` returns (automaton.py:209-216),
    produced on the device so that 3 bytes per cell cross the PCIe link instead of a float (W,H,3) tensor."""

    def __init__(self, H, W, device):
        self.H, self.W = int(H), int(W)
        self.dev = torch.empty((self.W, self.H, 3), dtype=torch.uint8, device=device)
        self.host = torch.empty((self.W, self.H, 3), dtype=torch.uint8).pin_memory()
        self.fresh = False   # `dev` holds the frame of the current `_worldmap`
        self.source = None   # (tensor, version) of the `_worldmap` the frame was made from

    def pack(self, worldmap: torch.Tensor):
        """float (3,H,W) CUDA tensor -> `dev` (b200ca_frame_pack_rgb)."""
        wm = worldmap if (worldmap.dtype == torch.float32 and worldmap.is_contiguous()) else worldmap.float().contiguous()
        with _lib.device_guard(wm.device):
            _lib.call("b200ca_frame_pack_rgb", _lib.ptr(wm), _lib.ptr(self.dev), self.H, self.W, _lib.current_stream(wm.device))

    def mark(self, worldmap):
        self.fresh, self.source = True, (worldmap, worldmap._version)

    def valid_for(self, worldmap):
        return self.fresh and self.source is not None and self.source[0] is worldmap and self.source[1] == worldmap._version

    def numpy(self) -> np.ndarray:
        self.host.copy_(self.dev, non_blocking=True)
        torch.cuda.current_stream(self.dev.device).synchronize()
        return self.host.numpy()


class Automaton:
    def __init__(self, size, device="cuda"):
        self.h, self.w = size
        self.size = size
        self._worldmap = torch.zeros((3, self.h, self.w), dtype=torch.float)
        self._components = []
        self._changed_components = []
        self.X, self.Y = torch.meshgrid(torch.arange(self.w), torch.arange(self.h), indexing="xy")
        self.device = device

    def __init_subclass__(cls, **kwargs):
        super().__init_subclass__(**kwargs)
        AUTOMATAS[cls.__name__] = cls

    # -- the two methods a model must define (automaton.py:58-71)
    def step(self):
        raise NotImplementedError('Please subclass "Automaton" class, and define self.step')

    def draw(self):
        raise NotImplementedError('Please subclass "Automaton" class, and define self.draw')

    # -- GUI hooks: inert in the headless mirror
    def process_gui_change(self, component):
        pass

    def process_event(self, event, camera=None):
        pass

    def process_all_events(self, event, camera=None):
        self.process_event(event, camera)

    def register_component(self, component, keep_size=False):
        self._components.append(component)

    @property
    def changed_components(self):
        return self._changed_components

    def _frame(self) -> FrameBuffer:
        fb = getattr(self, "_framebuf", None)
        wm = self._worldmap
        if fb is None or (fb.H, fb.W) != tuple(wm.shape[1:]) or fb.dev.device != wm.device:
            fb = self._framebuf = FrameBuffer(wm.shape[1], wm.shape[2], wm.device)
        return fb

    @property
    def worldmap(self):
        """(W,H,3) uint8 numpy frame (automaton.py:209-216).  With `_worldmap` on the GPU the frame is built there (by the
        model's fused draw kernel, or from `_worldmap` when a subclass drew it with torch) and copied as bytes; the
        returned array aliases a pinned buffer that the next call overwrites."""
        wm = self._worldmap
        if not wm.is_cuda or wm.shape[0] != 3:
            return (255 * wm.permute(2, 1, 0)).detach().cpu().numpy().astype(dtype=np.uint8)
        fb = self._frame()
        if not fb.valid_for(wm):
            fb.pack(wm)
            fb.mark(wm)
        return fb.numpy()

    def name(self):
        return self.__class__.__name__

    def get_string_state(self):
        return "Live stuff!"
