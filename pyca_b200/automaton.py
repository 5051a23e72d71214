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

AUTOMATAS = {}


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

    @property
    def worldmap(self):
        """(W,H,3) uint8 numpy frame (automaton.py:209-216)."""
        return (255 * self._worldmap.permute(2, 1, 0)).detach().cpu().numpy().astype(dtype=np.uint8)

    def name(self):
        return self.__class__.__name__

    def get_string_state(self):
        return "Live stuff!"
