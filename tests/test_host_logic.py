"""Host-side (cold path) logic of the drop-ins that runs without a GPU: parameter handling, coefficient tables, kernel
construction against the reference's kernels stored in the golden fixtures."""
import math

import numpy as np
import pytest
import torch

from tests.util import garbi_case, load_npz, raw_param_dict, t


def test_garbi_tables_follow_the_reference_argument_order():
    from pyca_b200.lenia import garbi_tables

    fx, garbi = garbi_case("unsuspected_lanthanotidae")
    d = raw_param_dict(fx)
    g = garbi_tables(d, 1, 3, 1)
    assert g["nh"] == 8 and g["rescale"] is None and g["clip"] is None
    co = t(fx["g_coeffs"]).reshape(9, 8, 2)
    assert torch.equal(g["cos"], co[:, :, 0]) and torch.equal(g["sin"], co[:, :, 1])
    # funcgen.py:99-104: 2*pi / period (fp32 tensor) * harmonics, period = 2
    ref = 2 * torch.pi / torch.full((9, 1), 2.0) * t(fx["g_harmonics"]).reshape(9, 8)
    assert torch.equal(g["freq"], ref)
    assert abs(float(g["freq"][0, 1]) - math.pi) < 1e-6
    fx2, _ = garbi_case("lenia_synth")
    g2 = garbi_tables(raw_param_dict(fx2), 1, 2, 1)
    assert g2["rescale"] == (-1.0, 1.0) and g2["clip"] == -0.5 and tuple(g2["stats"].shape) == (4, 2)


@pytest.mark.parametrize("name", ["underweight_skagway", "inaesthetic_maupassant"])
def test_kernel_construction_matches_reference_kernels(name):
    """compute_kernel (ring kernels and the k_arbi Fourier kernels with the smooth disc mask) against the reference's `k`."""
    import pyca_b200

    fx, _ = garbi_case(name)
    p = pyca_b200.LeniaParams(param_dict=raw_param_dict(fx))
    K = pyca_b200.compute_kernel(p, 3)
    assert (K - t(fx["K"])).abs().max().item() <= 1e-6
    assert torch.allclose(K.sum(dim=(-1, -2)), torch.ones(1, 3, 3), atol=1e-5)
    assert torch.allclose(p["weights"].sum(dim=1), torch.ones(1, 3), atol=1e-6)   # _sanitize: source weights sum to 1


def test_cuda_only_classes_refuse_other_devices():
    import pyca_b200

    with pytest.raises(pyca_b200.B200CAError):
        pyca_b200.CA2D((32, 32), device="cpu")
    fx, _ = garbi_case("lenia_synth")
    with pytest.raises(pyca_b200.B200CAError):
        pyca_b200.Lenia((32, 64), num_channels=2, params=pyca_b200.LeniaParams(param_dict=raw_param_dict(fx), channels=2),
                        state_init=torch.rand(1, 2, 32, 64), device="cpu")


@pytest.mark.parametrize("tag", ["b2c3", "b1c2k2"])
def test_default_gen_draws_the_reference_parameters(tag):
    """Lenia(params=None) starts from LeniaParams.default_gen: same torch seed -> the reference's random parameters."""
    import pyca_b200

    fx = load_npz("default_params.npz")
    B, C, km, ks = [int(v) for v in fx[f"{tag}_meta"]]
    torch.manual_seed(123)
    p = pyca_b200.LeniaParams.default_gen(batch_size=B, num_channels=C, k_size=ks, k_mult=km)
    for k in ["mu", "sigma", "beta", "mu_k", "sigma_k", "weights"]:
        assert torch.equal(p[k], t(fx[f"{tag}_{k}"])), k
    assert p.k_size == 31 and p.k_mult == km


def test_reference_arm_of_bench_is_free_of_the_product():
    """bench.py --impl reference must not import pyca_b200 or load the CUDA library: its code path -- RefArm,
    run_reference, static_config -- names neither."""
    import importlib.util
    import inspect
    import os

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = open(os.path.join(root, "bench.py")).read()
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(root, "bench.py"))
    assert spec is not None
    import ast

    tree = ast.parse(src)
    wanted = {"RefArm", "run_reference", "static_config", "cpu_baseline", "lenia_ring_worlds"}
    for node in tree.body:
        if isinstance(node, (ast.FunctionDef, ast.ClassDef)) and node.name in wanted:
            seg = ast.get_source_segment(src, node)
            assert "pyca_b200" not in seg and "b200ca" not in seg.replace("B200CA", ""), node.name
            wanted.discard(node.name)
    assert not wanted, f"not found in bench.py: {wanted}"
