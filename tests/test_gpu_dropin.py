"""The registry-override subclasses of pyca_b200/dropin.py, exercised with stand-in base classes that carry the
reference classes' attribute contract (the real PyCA needs pygame and is not on the GPU box)."""
import numpy as np
import pytest
import torch

from oracle import pyca_oracle as orc
from tests.util import load_npz, t

pytestmark = pytest.mark.gpu


class FakeCA2D:
    def __init__(self, world, s_num=0b1100, b_num=0b1000):
        self.world, self.s_num, self.b_num = world, s_num, b_num


class FakeLenia(torch.nn.Module):
    """state registered as a buffer exactly like the reference (lenia.py:79)."""

    def __init__(self, state, K, mu, sigma, weights, dt=0.1):
        super().__init__()
        self.register_buffer("state", state)
        self.k, self.mu, self.sigma, self.weights = K, mu, sigma, weights
        self.dt, self.k_mult, self.frames, self.g_arbi, self.has_food = dt, 1, 0, False, False
        self.b, self.alpha = 6, 0.03


def test_ca2d_dropin():
    from pyca_b200.dropin import make_ca2d

    rng = np.random.default_rng(0)
    w0 = (rng.random((100, 256)) < 0.4).astype(np.int32)
    a = make_ca2d(FakeCA2D)(torch.from_numpy(w0).cuda())
    assert type(a).__name__ == "CA2D"
    a.step()
    a.world[5:9, 5:9] = 1  # brush edit between steps
    ref = orc.gol_step(w0)
    ref[5:9, 5:9] = 1
    a.step()
    a.step()
    assert np.array_equal(a.world.cpu().numpy(), orc.gol_run(ref, 2))


def test_lenia_family_dropin():
    from pyca_b200.dropin import make_lenia, make_mace, make_xchan

    p = load_npz("params_xchan_arborical_asbestosis.npz")
    K, mu, sg, w = t(p["K"]).cuda(), t(p["mu_s"]).cuda(), t(p["sigma_s"]).cuda(), t(p["weights_s"]).cuda()
    g = torch.Generator().manual_seed(3)
    s0 = torch.rand(1, 3, 64, 96, generator=g)
    le = make_lenia(FakeLenia)(s0.cuda(), K, mu, sg, w)
    le.step()
    ref = orc.lenia_step(s0, K.cpu(), mu.cpu(), sg.cpu(), w.cpu(), 0.1)
    assert (le.state.cpu() - ref).abs().max().item() <= 1e-5 and le.frames == 1
    le.state[:, :, :4] = 0.5  # in-place edit of the published tensor
    ref[:, :, :4] = 0.5
    le.step()
    ref2 = orc.lenia_step(ref, K.cpu(), mu.cpu(), sg.cpu(), w.cpu(), 0.1)
    assert (le.state.cpu() - ref2).abs().max().item() <= 1e-5
    for factory, fn in ((make_mace, lambda s: orc.mace_step(s, K.cpu(), mu.cpu(), sg.cpu(), w.cpu(), 6.0)[0]),
                        (make_xchan, lambda s: orc.xchan_step(s, K.cpu(), mu.cpu(), sg.cpu(), w.cpu(), 6.0, 0.03))):
        a = factory(FakeLenia)(s0.cuda(), K, mu, sg, w)
        a.step()
        assert (a.state.cpu() - fn(s0)).abs().max().item() <= 1e-5
        assert tuple(a.Aff.shape) == (1, 3, 64, 96)


def test_flow_lenia_dropin():
    from pyca_b200.dropin import make_flow

    p = load_npz("params_mace_anencephalic_brakeman.npz")
    K, mu, sg, w = t(p["K"]).cuda(), t(p["mu_s"]).cuda(), t(p["sigma_s"]).cuda(), t(p["weights_s"]).cuda()
    g = torch.Generator().manual_seed(8)
    s0 = 1.2 * torch.rand(1, 3, 64, 96, generator=g)

    class FakeFlow(FakeLenia):
        dd, sigma_rt, theta_x, n = 2, 0.65, 2, 2

    a = make_flow(FakeFlow)(s0.cuda(), K, mu, sg, w)
    assert type(a).__name__ == "FlowLenia"
    a.step()
    ref = orc.flow_step(s0, K.cpu(), mu.cpu(), sg.cpu(), w.cpu(), 0.1)
    assert (a.state.cpu() - ref).abs().max().item() <= 2e-5 and a.frames == 1
