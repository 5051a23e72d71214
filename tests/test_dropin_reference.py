"""Registry override against the REAL reference classes.  Runs only where /root/reference exists (the build
container); the GPU box has no reference tree.  No step is executed here (the drop-in step needs a B200): the test pins
that the factories subclass the reference's own classes, that PyCA's AUTOMATAS registry (keyed by class name, last
definition wins, automaton.py:104-111) ends up pointing at the drop-ins, and that construction still runs the reference's
own constructors."""
import os
import sys

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("PYCA_REFERENCE", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF_ROOT, "pyca")), reason="reference tree not available")


def _ref():
    sys.path.insert(0, os.path.join(os.path.dirname(HERE), "baseline"))
    from ref_loader import load_reference

    return load_reference()


def test_factories_subclass_and_reregister_the_reference_models():
    REF = _ref()
    from pyca.automata.automaton import AUTOMATAS

    from pyca_b200 import dropin

    made = {}
    for name, factory, base in [("CA2D", dropin.make_ca2d, REF.CA2D), ("Lenia", dropin.make_lenia, REF.Lenia),
                                ("MaCELenia", dropin.make_mace, REF.MaCELenia), ("MaCELeniaXChan", dropin.make_xchan, REF.MaCELeniaXChan),
                                ("FlowLenia", dropin.make_flow, REF.FlowLenia), ("NeuralCA", dropin.make_nca, REF.NeuralCA)]:
        cls = factory(base)
        made[name] = cls
        assert issubclass(cls, base) and cls.__name__ == name
        assert AUTOMATAS[name] is cls, f"{name}: the registry must point at the drop-in (last definition wins)"
        assert cls.step is not base.step
        assert isinstance(cls.__dict__.get("worldmap"), property)      # device-side frame conversion is attached
    # construction goes through the reference's own __init__ (widgets, parameter loading, kernels): CPU is fine for that
    torch.manual_seed(0)
    a = made["CA2D"]((40, 50), device="cpu")
    assert tuple(a.world.shape) == (40, 50)
    p = REF.LeniaParams(from_file=os.path.join(REF_ROOT, "demo_data/demo_lenia/abominable_mongolic.pt"))
    le = made["Lenia"]((32, 64), params=p, state_init=torch.rand(1, 3, 32, 64), device="cpu")
    assert tuple(le.k.shape) == (1, 3, 3, 31, 31) and le.frames == 0
    # a CPU tensor falls back to the reference's own worldmap expression
    le.draw()
    assert le.worldmap.shape == (64, 32, 3)
