"""FlowLenia: Lenia affinity kernels + the fused flow-field / reintegration-tracking kernel against the
reference's own steps (tests/golden/flow.npz) and the oracle on further shapes."""
import pytest
import torch

from oracle import pyca_oracle as orc
from tests.util import load_npz, t

pytestmark = pytest.mark.gpu


def _params(name):
    import pyca_b200

    p = load_npz(f"params_mace_{name}.npz")
    d = {k: t(p[k]) for k in ["mu", "sigma", "beta", "mu_k", "sigma_k", "weights"]}
    d["k_size"] = int(p["k_size"])
    return pyca_b200.LeniaParams(param_dict=d), p


@pytest.mark.parametrize("path", ["fft", "fft_rows", "direct"])
def test_flow_lenia_golden(path):
    import pyca_b200

    fx = load_npz("flow.npz")
    for ci in range(int(fx["ncases"])):
        params, _ = _params(str(fx[f"c{ci}_file"]))
        s0 = t(fx[f"c{ci}_state0"])
        H, W = s0.shape[2:]
        dt, dd, srt, thx, n = [float(v) for v in fx[f"c{ci}_meta"]]
        a = pyca_b200.FlowLenia((H, W), dt=dt, params=params, state_init=s0, device="cuda", dd=int(dd), sigma_rt=srt,
                                conv_path=path)
        assert a.use_fft == (path != "direct")
        i = 1
        while f"c{ci}_state{i}" in fx:
            a.step()
            ref = t(fx[f"c{ci}_state{i}"])
            assert (a.state.cpu() - ref).abs().max().item() <= 1e-5 * i * max(1.0, float(ref.abs().max()))
            i += 1
        assert a.frames == i - 1


@pytest.mark.parametrize("shape,dd", [((33, 47), 2), ((64, 1030), 2), ((100, 96), 1), ((50, 70), 3)])
def test_flow_lenia_shapes_vs_oracle(shape, dd):
    """Odd shapes (partial tiles), wide worlds (positions ~1000: fp32 position rounding as the reference), dd = 1 and 3,
    in-place edit between steps, and mass conservation in the interior."""
    import pyca_b200

    H, W = shape
    params, p = _params("anencephalic_brakeman")
    K, mu, sg, w = t(p["K"]), t(p["mu_s"]), t(p["sigma_s"]), t(p["weights_s"])
    g = torch.Generator().manual_seed(H * 7 + W)
    s0 = 1.5 * torch.rand(1, 3, H, W, generator=g) * (torch.rand(1, 3, H, W, generator=g) < 0.5)
    a = pyca_b200.FlowLenia((H, W), params=params, state_init=s0, device="cuda", dd=dd)
    a.step()
    ref = orc.flow_step(s0, K, mu, sg, w, 0.1, dd, 0.65)
    # The reference tracks ABSOLUTE positions in fp32 (mu = x + 0.5 + clip(dt F)): at x ~ 1000 they sit on a grid of
    # 1.2e-4, so a 1e-7 difference in F can move a mass square by one grid step.  That quantum (times the state value) is
    # part of the bound; for the 48..128-wide worlds it is below 1e-5.
    import numpy as np

    quantum = float(np.spacing(np.float32(max(H, W)))) * float(s0.max())
    assert (a.state.cpu() - ref).abs().max().item() <= 2e-5 + quantum
    a.state[:, :, 5:9, 5:9] = 0.0
    ref[:, :, 5:9, 5:9] = 0.0
    a.step()
    ref2 = orc.flow_step(ref, K, mu, sg, w, 0.1, dd, 0.65)
    assert (a.state.cpu() - ref2).abs().max().item() <= 2 * (2e-5 + quantum)
    # matter that starts away from the edge stays in the world: total mass is conserved
    s1 = torch.zeros(1, 3, H, W)
    s1[:, :, 12:H - 12, 12:W - 12] = s0[:, :, 12:H - 12, 12:W - 12]
    b = pyca_b200.FlowLenia((H, W), params=params, state_init=s1, device="cuda", dd=dd)
    m0 = float(b.total_mass())
    for _ in range(3):
        b.step()
    assert abs(float(b.total_mass()) - m0) <= 1e-5 * m0


def test_flow_lenia_batch_of_two_parameter_sets():
    """B = 2 worlds with different parameter files (batched params, lenia.py:89-101) through the flow kernel's batch axis."""
    import pyca_b200

    pa, pb = load_npz("params_mace_anencephalic_brakeman.npz"), load_npz("params_mace_apparent_kenalog.npz")
    d = {k: torch.cat([t(pa[k]), t(pb[k])], dim=0) for k in ["mu", "sigma", "beta", "mu_k", "sigma_k", "weights"]}
    d["k_size"] = int(pa["k_size"])
    g = torch.Generator().manual_seed(2)
    s0 = 1.3 * torch.rand(2, 3, 48, 80, generator=g)
    a = pyca_b200.FlowLenia((2, 48, 80), params=pyca_b200.LeniaParams(param_dict=d), state_init=s0, device="cuda")
    a.step()
    for b, p in enumerate((pa, pb)):
        ref = orc.flow_step(s0[b:b + 1], t(p["K"]), t(p["mu_s"]), t(p["sigma_s"]), t(p["weights_s"]), 0.1)
        assert (a.state[b:b + 1].cpu() - ref).abs().max().item() <= 2e-5, f"world {b}"
