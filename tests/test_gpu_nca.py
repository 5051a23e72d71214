"""GPU parity of the Neural-CA step (through the C ABI) against the golden vectors produced by the
unmodified reference and against the oracle.  Tolerance: max-abs 1e-5 (fp32), same update mask on both sides."""
import numpy as np
import pytest
import torch

from oracle import pyca_oracle as orc
from tests.util import golden_names, load_npz, t

pytestmark = pytest.mark.gpu
TOL = 1e-5
IMPLS = [0, 1]  # SIMT, tensor core


def _weights(fx):
    import pyca_b200

    return pyca_b200.NCAWeights(t(fx["W1"]), t(fx["b1"]), t(fx["W2"]), t(fx["b2"]), "cuda")


def _impl_or_skip(impl):
    if impl == 1:
        from pyca_b200 import _lib
        import pyca_b200

        # the tensor-core path reports B200CA_E_UNSUPPORTED until it is built
        try:
            w = pyca_b200.NCAWeights(torch.zeros(128, 48), torch.zeros(128), torch.zeros(16, 128), torch.zeros(16), "cuda")
            pyca_b200.nca_step(torch.zeros(1, 16, 8, 8, device="cuda"), w, impl=1, update_mask=torch.ones(1, 1, 8, 8))
        except _lib.B200CAError as e:
            if "not built" in str(e):
                pytest.skip("tensor-core NCA path not built")
            raise


@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("name", golden_names("nca"))
def test_nca_golden(name, impl):
    import pyca_b200

    _impl_or_skip(impl)
    fx = load_npz(f"nca_{name}.npz")
    w = _weights(fx)
    s0 = t(fx["state_in"]).cuda()
    out = pyca_b200.nca_step(s0, w, update_mask=t(fx["mask"]).cuda(), impl=impl).cpu()
    ref = t(fx["state_out"])
    err = (out - ref).abs().max().item()
    assert err <= TOL, f"{name} impl {impl}: max-abs {err:.2e}"
    assert torch.equal(out == 0, ref == 0), "dead-cell pattern differs"


@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("shape", [(1, 8, 8), (3, 37, 45), (2, 64, 128), (1, 33, 200), (5, 16, 31)])
def test_nca_random_vs_oracle(shape, impl):
    import pyca_b200

    _impl_or_skip(impl)
    fx = load_npz("nca_betta.npz")
    w = _weights(fx)
    B, H, W = shape
    g = torch.Generator().manual_seed(B * 10000 + H * 100 + W)
    s0 = torch.rand(B, 16, H, W, generator=g)
    s0[:, 3] *= (torch.rand(B, H, W, generator=g) < 0.3)  # sparse alpha so the life masks bite
    mask = torch.rand(B, 1, H, W, generator=g) <= 0.5
    out = pyca_b200.nca_step(s0.cuda(), w, update_mask=mask.cuda(), impl=impl).cpu()
    ref = orc.nca_step(s0, t(fx["W1"]), t(fx["b1"]), t(fx["W2"]), t(fx["b2"]), mask)
    err = (out - ref).abs().max().item()
    assert err <= TOL, f"{shape} impl {impl}: {err:.2e}"
    assert torch.equal(out == 0, ref == 0)


def test_neuralca_dropin_growth_and_rng_modes(tmp_path):
    """NeuralCA grows from the seed pixel; philox mode is deterministic; torch mode consumes torch's CUDA stream."""
    import pyca_b200

    fx = load_npz("nca_betta.npz")
    ck = {"config": {"n_states": 16, "n_hidden": 128, "device": "cpu"},
          "state_dict": {"computer.0.weight": t(fx["W1"]), "computer.0.bias": t(fx["b1"]), "computer.2.weight": t(fx["W2"]),
                         "computer.2.bias": t(fx["b2"])}}
    torch.save(ck, tmp_path / "betta.pt")
    runs = []
    for _ in range(2):
        a = pyca_b200.NeuralCA((64, 64), str(tmp_path), device="cuda", rng_seed=7)
        assert a.state.shape == (1, 16, 64, 64) and a.state[0, :, 31, 31].sum().item() == 16  # seed slice, nca.py:151-162
        for _ in range(24):
            a.step()
        runs.append(a.state.clone())
        assert a.state.min() >= 0 and a.state.max() <= 1
    assert torch.equal(runs[0], runs[1]), "philox mode must be deterministic"
    alive = (runs[0][0, 3] > 0.1).float().mean().item()
    assert alive > 0.02, f"pattern did not grow (alive fraction {alive})"
    # torch rng mode == oracle fed with the same torch.rand draws on the same device
    a = pyca_b200.NeuralCA((48, 48), str(tmp_path), device="cuda", rng="torch")
    ref = a.state.cpu().clone()
    torch.manual_seed(123)
    for _ in range(6):
        a.step()
    torch.manual_seed(123)
    for _ in range(6):
        m = (torch.rand(1, 1, 48, 48, device="cuda") <= 0.5).cpu()
        ref = orc.nca_step(ref, t(fx["W1"]), t(fx["b1"]), t(fx["W2"]), t(fx["b2"]), m)
    # 6 steps: allow accumulated rounding, but the live pattern must match
    assert (a.state.cpu() - ref).abs().max().item() <= 1e-4
    a.draw()
    assert tuple(a._worldmap.shape) == (3, 48, 48)


def test_philox_update_fraction():
    import pyca_b200

    fx = load_npz("nca_betta.npz")
    w = _weights(fx)
    s0 = torch.rand(4, 16, 64, 64).cuda()
    out = pyca_b200.nca_step(s0, w, seed=99, step_index=3)
    frac = (out[:, 0] != 0).float().mean().item()  # sigmoid > 0, dense alive input -> non-zero iff updated
    assert 0.47 < frac < 0.53
    out2 = pyca_b200.nca_step(s0, w, seed=99, step_index=4)
    assert not torch.equal(out == 0, out2 == 0)


def test_host_entry_point():
    import ctypes

    from pyca_b200 import _lib

    fx = load_npz("nca_whale.npz")
    w = _weights(fx)
    s0 = t(fx["state_in"])
    B, _, H, W = s0.shape
    g = torch.Generator().manual_seed(1)
    masks = (torch.rand(2, B, H, W, generator=g) <= 0.5).to(torch.uint8)
    h_in = s0.numpy()
    h_out = np.empty_like(h_in)
    hm = masks.numpy()
    _lib.call("b200ca_nca_run_host", h_in.ctypes.data_as(ctypes.c_void_p), h_out.ctypes.data_as(ctypes.c_void_p),
              _lib.ptr(w.prep), hm.ctypes.data_as(ctypes.c_void_p), 0, B, 16, 128, H, W, 0, 2)
    ref = s0
    for i in range(2):
        ref = orc.nca_step(ref, t(fx["W1"]), t(fx["b1"]), t(fx["W2"]), t(fx["b2"]), masks[i].bool()[:, None])
    assert (torch.from_numpy(h_out) - ref).abs().max().item() <= 2 * TOL


@pytest.mark.parametrize("impl", [0, 1])
def test_batch_shards_draw_the_global_mask(impl):
    """A batch split over GPUs: shards stepped with first_world = their offset reproduce the unsplit batch
    bit for bit, because the in-kernel Philox counter is the global cell index."""
    import pyca_b200

    _impl_or_skip(impl)
    fx = load_npz("nca_betta.npz")
    w = pyca_b200.NCAWeights(t(fx["W1"]), t(fx["b1"]), t(fx["W2"]), t(fx["b2"]), "cuda")
    g = torch.Generator().manual_seed(4)
    s = torch.rand(6, 16, 40, 72, generator=g).cuda()
    whole = pyca_b200.nca_step(s, w, seed=9, step_index=3, impl=impl)
    parts = [pyca_b200.nca_step(s[o:o + n].contiguous(), w, seed=9, step_index=3, impl=impl, first_world=o) for o, n in ((0, 2), (2, 3), (5, 1))]
    assert torch.equal(whole, torch.cat(parts, dim=0))
    assert not torch.equal(whole[2:5], pyca_b200.nca_step(s[2:5].contiguous(), w, seed=9, step_index=3, impl=impl))   # offset matters
