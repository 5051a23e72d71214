"""Headless loader for the UNMODIFIED PyCA reference.

Search order: ``$PYCA_REFERENCE``, ``/root/reference`` (build container only), ``baseline/_ref`` (the offline install made
by ``baseline/install_ref.py``: git-ignored, travels to the GPU box).  Used (a) in the build container to pin ``oracle/``
against the reference and to generate the fixtures in ``tests/golden/`` (``tests/golden/generate.py``), and (b) by
``bench.py --impl reference`` / its ``cpu_baseline`` leg, which time the reference's own ``step()`` on the host cores
from ``baseline/_ref``.  ``pyca_b200`` never imports this module.

The reference cannot be imported as-is (``pygame``, ``pygame_gui``,
``easydict``, ``torchenhanced``, ``pyperlin``, ``showtens``, ``matplotlib``
are absent).  None of these touches the arithmetic of ``step()``, so they are
replaced by inert stand-ins:

* GUI / plotting packages -> auto-attribute fake modules;
* ``easydict.EasyDict`` -> attribute dict;
* ``torchenhanced.DevModule/ConfigModule`` -> ``nn.Module`` with the
  ``_devtens`` buffer the NCA checkpoints carry (``nca.py:181,306``);
* ``pyca.interface`` -> permissive widget classes, because the real widgets
  need a live ``UIManager`` (``interface/ui_components/UIComponent.py:29``).
"""
from __future__ import annotations

import importlib
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))


def _find_root():
    for cand in (os.environ.get("PYCA_REFERENCE"), "/root/reference", os.path.join(_HERE, "_ref")):
        if cand and os.path.isdir(os.path.join(cand, "pyca")):
            return cand
    return os.environ.get("PYCA_REFERENCE", "/root/reference")


REF_ROOT = _find_root()


class _Fake(types.ModuleType):
    """Module whose every attribute is a permissive callable/class."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        obj = _Anything
        setattr(self, name, obj)
        return obj


class _AnythingMeta(type):
    def __getattr__(cls, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return 0


class _Anything(metaclass=_AnythingMeta):
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Anything()


def _fake_module(name):
    m = _Fake(name)
    m.__path__ = []  # behaves as a package for sub-imports
    sys.modules[name] = m
    return m


def _install_stubs():
    import torch
    import torch.nn as nn

    for name in [
        "pygame", "pygame.surfarray", "pygame_gui", "pygame_gui.core", "pygame_gui.core.utility",
        "pygame_gui.core.ui_element", "pygame_gui.core.ui_container", "pygame_gui.elements",
        "showtens", "pyperlin", "matplotlib", "matplotlib.pyplot", "matplotlib.colors", "cv2",
    ]:
        if name not in sys.modules:
            _fake_module(name)
    sys.modules["pygame_gui.core.utility"].get_default_manager = lambda: _Anything()

    # easydict
    ed = types.ModuleType("easydict")

    class EasyDict(dict):
        def __init__(self, d=None, **kw):
            super().__init__()
            for k, v in {**(d or {}), **kw}.items():
                self[k] = v

        def __getattr__(self, k):
            try:
                return self[k]
            except KeyError as e:
                raise AttributeError(k) from e

        def __setattr__(self, k, v):
            self[k] = v

    ed.EasyDict = EasyDict
    sys.modules["easydict"] = ed

    # torchenhanced
    te = types.ModuleType("torchenhanced")

    class DevModule(nn.Module):
        def __init__(self, device="cpu"):
            super().__init__()
            self.register_buffer("_devtens", torch.empty(0, device=device))

        @property
        def device(self):
            return self._devtens.device

        @property
        def paranum(self):
            return sum(p.numel() for p in self.parameters())

    class ConfigModule(DevModule):
        def __init__(self, device="cpu"):
            super().__init__(device)
            self._config = None

        def __init_subclass__(cls, **kw):
            super().__init_subclass__(**kw)
            orig = cls.__init__

            def wrapped(self, *a, **k):
                import inspect

                ba = inspect.signature(orig).bind(self, *a, **k)
                ba.apply_defaults()
                cfg = {n: v for n, v in ba.arguments.items() if n != "self"}
                orig(self, *a, **k)
                self._config = cfg

            cls.__init__ = wrapped

        @property
        def config(self):
            return self._config

    te.DevModule, te.ConfigModule = DevModule, ConfigModule
    sys.modules["torchenhanced"] = te

    # pyca namespace + fake interface package
    pyca = types.ModuleType("pyca")
    pyca.__path__ = [os.path.join(REF_ROOT, "pyca")]
    sys.modules["pyca"] = pyca

    class UIComponent:
        def __init__(self, *a, **k):
            self.value = k.get("initial_value", 0)
            self.active = False
            self.visible = True
            self.rel_size = (0.04, 1.0)
            self.state = 0

        def handle_event(self, event):
            return False

        def add_component(self, c):
            pass

        def toggle_visibility(self):
            pass

        def __getattr__(self, name):
            if name.startswith("__"):
                raise AttributeError(name)
            return _Anything()

    names = ["Button", "MultiToggle", "LabeledSlider", "Slider", "InputField", "DropDown", "TextBox",
             "TextLabel", "BoxHolder", "VertContainer", "HorizContainer", "SmartFont", "InputBox"]
    iface = types.ModuleType("pyca.interface")
    iface.__path__ = []
    iface.UIComponent = UIComponent
    iface.__all__ = ["UIComponent"] + names
    for n in names:
        setattr(iface, n, type(n, (UIComponent,), {}))
    sys.modules["pyca.interface"] = iface
    uic = types.ModuleType("pyca.interface.ui_components")
    uic.__path__ = []
    for n in ["UIComponent"] + names:
        setattr(uic, n, getattr(iface, n))
    sys.modules["pyca.interface.ui_components"] = uic
    uic2 = types.ModuleType("pyca.interface.ui_components.UIComponent")
    uic2.UIComponent = UIComponent
    sys.modules["pyca.interface.ui_components.UIComponent"] = uic2


_loaded = None


def load_reference():
    """Returns a namespace with the reference's own classes (byte-for-byte code)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not os.path.isdir(os.path.join(REF_ROOT, "pyca")):
        raise FileNotFoundError(f"reference tree not found at {REF_ROOT}")
    sys.dont_write_bytecode = True  # reference tree is read-only
    _install_stubs()
    ns = types.SimpleNamespace()
    ns.root = REF_ROOT
    ns.ca2d = importlib.import_module("pyca.automata.models.ca2d")
    ns.lenia = importlib.import_module("pyca.automata.models.lenia.lenia")
    ns.macelenia = importlib.import_module("pyca.automata.models.lenia.macelenia")
    ns.xchan = importlib.import_module("pyca.automata.models.lenia.mace_lenia_xchan")
    ns.leniaparams = importlib.import_module("pyca.automata.models.lenia.leniaparams")
    ns.nca = importlib.import_module("pyca.automata.models.nca")
    ns.flow = importlib.import_module("pyca.automata.models.lenia.flow_lenia")
    ns.FlowLenia = ns.flow.FlowLenia
    ns.CA2D = ns.ca2d.CA2D
    ns.Lenia = ns.lenia.Lenia
    ns.MaCELenia = ns.macelenia.MaCELenia
    ns.MaCELeniaXChan = ns.xchan.MaCELeniaXChan
    ns.LeniaParams = ns.leniaparams.LeniaParams
    ns.NeuralCA = ns.nca.NeuralCA
    ns.NCAModule = ns.nca.NCAModule
    _loaded = ns
    return ns


if __name__ == "__main__":
    import torch

    ref = load_reference()
    torch.manual_seed(0)
    a = ref.CA2D((50, 60))
    a.step()
    print("CA2D ok", a.world.dtype, int(a.world.sum()))
    p = ref.LeniaParams(from_file=os.path.join(REF_ROOT, "demo_data/demo_lenia/abominable_mongolic.pt"))
    l = ref.Lenia((64, 64), params=p, state_init=torch.rand(1, 3, 64, 64))
    l.step()
    print("Lenia ok", l.state.shape, float(l.state.mean()))
    p = ref.LeniaParams(from_file=os.path.join(REF_ROOT, "demo_data/demo_macelenia_xchan/arborical_asbestosis.pt"))
    m = ref.MaCELeniaXChan((64, 64), params=p, state_init=torch.rand(1, 3, 64, 64))
    m0 = float(m.state.sum())
    m.step()
    print("XChan ok", m0, float(m.state.sum()), m.b, m.alpha)
    n = ref.NeuralCA((32, 32), os.path.join(REF_ROOT, "demo_data/NCA_models"))
    n.step()
    print("NCA ok", n.state.shape, float(n.state.sum()))
