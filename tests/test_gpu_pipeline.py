"""pyca_b200.pipeline.HostPipeline: worlds streamed from pinned host buffers through several automata on their own
streams must come back exactly as from the one-world synchronous loop (same kernels, same inputs)."""
import numpy as np
import pytest
import torch

from oracle import pyca_oracle as orc
from tests.util import load_npz, t

pytestmark = pytest.mark.gpu


def _lenia_params():
    import pyca_b200

    p = load_npz("params_lenia_abominable_mongolic.npz")
    d = {k: t(p[k]) for k in ["mu", "sigma", "beta", "mu_k", "sigma_k", "weights"]}
    d["k_size"] = int(p["k_size"])
    return pyca_b200.LeniaParams(param_dict=d), p


def test_pipeline_lenia_matches_sync_and_oracle():
    import pyca_b200
    from pyca_b200.pipeline import HostPipeline

    params, p = _lenia_params()
    H, W = 128, 256
    g = torch.Generator().manual_seed(5)
    worlds = [torch.rand(1, 3, H, W, generator=g).pin_memory() for _ in range(7)]
    outs = [torch.empty(1, 3, H, W).pin_memory() for _ in range(7)]
    mk = lambda: pyca_b200.Lenia((H, W), params=params, state_init=worlds[0], device="cuda")
    pipe = HostPipeline(mk, depth=3)
    for wi, wo in zip(worlds, outs):
        pipe.submit(wi, wo, nsteps=2)
    pipe.drain()
    ref_auto = mk()
    K, mu, sg, w = t(p["K"]), t(p["mu_s"]), t(p["sigma_s"]), t(p["weights_s"])
    for wi, wo in zip(worlds, outs):
        ref_auto.state = wi.cuda()
        ref_auto.step()
        ref_auto.step()
        assert torch.equal(ref_auto.state.cpu(), wo)  # same kernels: bit-identical
        o = orc.lenia_step(orc.lenia_step(wi, K, mu, sg, w, 0.1), K, mu, sg, w, 0.1)
        assert (wo - o).abs().max().item() <= 2e-5


def test_pipeline_gol_world_attr():
    import pyca_b200
    from pyca_b200.pipeline import HostPipeline

    rng = np.random.default_rng(1)
    H, W = 96, 200
    worlds = [torch.from_numpy((rng.random((H, W)) < 0.4).astype(np.int32)).pin_memory() for _ in range(5)]
    outs = [torch.empty(H, W, dtype=torch.int32).pin_memory() for _ in range(5)]
    pipe = HostPipeline(lambda: pyca_b200.CA2D((H, W), device="cuda"), depth=2, attr="world")
    for wi, wo in zip(worlds, outs):
        pipe.submit(wi, wo, nsteps=3)
    pipe.drain()
    for wi, wo in zip(worlds, outs):
        assert np.array_equal(wo.numpy(), orc.gol_run(wi.numpy(), 3))


def test_pipeline_rejects_pageable_buffers():
    import pyca_b200
    from pyca_b200.pipeline import HostPipeline

    pipe = HostPipeline(lambda: pyca_b200.CA2D((32, 64), device="cuda"), depth=1, attr="world")
    with pytest.raises(pyca_b200.B200CAError):
        pipe.submit(torch.zeros(32, 64, dtype=torch.int32), torch.zeros(32, 64, dtype=torch.int32))


@pytest.mark.parametrize("family,shape", [(0, (128, 256)), (0, (50, 70)), (2, (64, 96)), (1, (64, 1024))])
def test_c_pipe_matches_sync(family, shape):
    """b200ca_lenia_pipe_*: worlds come back bit-identical to the one-world synchronous drop-in (same kernels)."""
    import pyca_b200
    from pyca_b200.pipeline import LeniaHostPipe

    H, W = shape
    tag = "lenia_abominable_mongolic" if family == 0 else "xchan_arborical_asbestosis"
    p = load_npz(f"params_{tag}.npz")
    d = {k: t(p[k]) for k in ["mu", "sigma", "beta", "mu_k", "sigma_k", "weights"]}
    d["k_size"] = int(p["k_size"])
    params = pyca_b200.LeniaParams(param_dict=d)
    cls = [pyca_b200.Lenia, pyca_b200.MaCELenia, pyca_b200.MaCELeniaXChan][family]
    g = torch.Generator().manual_seed(11)
    worlds = [torch.rand(1, 3, H, W, generator=g).pin_memory() for _ in range(5)]
    outs = [torch.empty(1, 3, H, W).pin_memory() for _ in range(5)]
    auto = cls((H, W), params=params, state_init=worlds[0], device="cuda")
    pipe = LeniaHostPipe(auto, depth=2)
    for wi, wo in zip(worlds, outs):
        pipe.submit(wi, wo, nsteps=2)
    pipe.drain()
    for wi, wo in zip(worlds, outs):
        auto.state = wi.cuda()
        auto.step()
        auto.step()
        assert torch.equal(auto.state.cpu(), wo)
    with pytest.raises(pyca_b200.B200CAError):
        pipe.submit(torch.zeros(1, 3, H, W), outs[0])
    pipe.close()


def test_step_graph_matches_eager_steps():
    """pyca_b200.graph.StepGraph: captured step() loops (PDL launches included) replay to the same bits as eager stepping."""
    import pyca_b200
    from pyca_b200.graph import StepGraph

    params, _ = _lenia_params()
    g = torch.Generator().manual_seed(21)
    states = [torch.rand(1, 3, 64, 1024, generator=g) for _ in range(3)]   # N = 1024 FFT path (PDL-chained kernels)
    mk = lambda s: pyca_b200.Lenia((64, 1024), params=params, state_init=s, device="cuda")
    eager = [mk(s) for s in states]
    graphed = [mk(s) for s in states]
    sg = StepGraph(graphed, reps=2)          # (the warm-up steps inside refresh() are undone: nothing has advanced yet)
    for a, b in zip(eager, graphed):
        assert torch.equal(a.state.cpu(), b.state.cpu()) and b.frames == 0
    for _ in range(3):
        sg.replay()
        for a in eager:
            a.step()
            a.step()
    for a, b in zip(eager, graphed):
        assert torch.equal(a.state, b.state) and a.frames == b.frames == 6
    graphed[1].state[:, :, :8] = 0.25        # edit between replays -> refresh() re-captures
    eager[1].state[:, :, :8] = 0.25
    sg.refresh()
    sg.replay()
    for a in eager:
        for _ in range(2):
            a.step()
    for a, b in zip(eager, graphed):
        assert torch.equal(a.state, b.state)
    ca = pyca_b200.CA2D((96, 200), device="cuda")
    w0 = ca.world.clone()
    sgc = StepGraph(ca, reps=4)
    sgc.replay()
    assert np.array_equal(ca.world.cpu().numpy(), orc.gol_run(w0.cpu().numpy(), 4))
    # NeuralCA is refused: its update mask is keyed by a step index passed by value at launch
    class _FakeNca:
        rng, device = "philox", "cuda"
    with pytest.raises(pyca_b200.B200CAError):
        StepGraph(_FakeNca(), reps=2)
