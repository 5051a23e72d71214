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


def test_conv_path_rule_and_band_heights_without_a_gpu():
    """b200ca_lenia_pick_algo_ks / b200ca_lenia_march_band_rows are host-side rules (profiles/r02_conv_paths.md): no device needed."""
    from pyca_b200 import _lib

    L = _lib.lib()
    DIRECT, ROWS, MARCH = 0, 1, 2
    assert L.b200ca_lenia_pick_algo_ks(1, 1, 1024, 1024, 31) == ROWS          # one small world: two-kernel hybrid
    assert L.b200ca_lenia_pick_algo_ks(24, 1, 1024, 1024, 31) == MARCH        # the headline batch
    assert L.b200ca_lenia_pick_algo_ks(1, 1, 16384, 16384, 31) == MARCH
    assert L.b200ca_lenia_pick_algo_ks(1, 3, 4096, 4096, 31) == MARCH         # several source channels: big worlds march,
    assert L.b200ca_lenia_pick_algo_ks(1, 3, 1024, 1024, 31) == ROWS          # small ones take the hybrid
    assert L.b200ca_lenia_pick_algo_ks(1, 3, 4097, 4096, 31) == MARCH         # odd height: the hybrid cannot, the marching kernel can
    assert L.b200ca_lenia_pick_algo_ks(1, 1, 4096, 4096, 7) == MARCH          # big single-channel worlds: marching at every kernel size
    assert L.b200ca_lenia_pick_algo_ks(1, 1, 1024, 1024, 7) == DIRECT         # small kernels on small worlds: direct
    assert L.b200ca_lenia_pick_algo_ks(1, 3, 4096, 4096, 5) == DIRECT
    assert L.b200ca_lenia_pick_algo_ks(1, 1, 4096, 4096, 11) == MARCH and L.b200ca_lenia_pick_algo_ks(1, 1, 1024, 1024, 11) == DIRECT
    assert L.b200ca_lenia_pick_algo_ks(1, 1, 24, 24, 31) == DIRECT            # too small for the FFT paths
    assert L.b200ca_lenia_pick_algo(1, 1, 1024, 1024) == ROWS
    assert L.b200ca_lenia_march_supported(33, 449) == 1 and L.b200ca_lenia_march_supported(16, 64) == 0
    for H, W, B in [(1024, 1024, 1), (1024, 1024, 24), (4096, 4096, 1), (16384, 16384, 1), (2048, 16384, 1)]:
        rows = L.b200ca_lenia_march_band_rows(H, W, B, 1)
        assert rows % 16 == 0 and 16 <= rows <= 256, (H, W, B, rows)
    assert L.b200ca_lenia_march_band_rows(16384, 16384, 1, 1) > L.b200ca_lenia_march_band_rows(1024, 1024, 1, 1)


def test_lenia_params_expand_matches_the_reference_semantics():
    """LeniaParams.expand (batch_params.py:217-236): batch-1 tensors repeated along dim 0, scalars kept."""
    import torch

    import pyca_b200
    from tests.util import load_npz, t

    p = load_npz("params_lenia_abominable_mongolic.npz")
    d = {k: t(p[k]) for k in ["mu", "sigma", "beta", "mu_k", "sigma_k", "weights"]}
    d["k_size"] = int(p["k_size"])
    P = pyca_b200.LeniaParams(param_dict=d)
    E = P.expand(5)
    assert E.batch_size == 5 and E["k_size"] == P["k_size"]
    for k in ["mu", "sigma", "beta", "mu_k", "sigma_k", "weights"]:
        assert E[k].shape[0] == 5 and torch.allclose(E[k][3], P[k][0], atol=1e-6)   # (the constructor re-sanitises, as the reference does)
    import pytest
    with pytest.raises(AssertionError):
        E.expand(2)


def test_both_bench_arms_print_the_same_config():
    import importlib.util
    import os

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = open(os.path.join(root, "bench.py")).read()
    ns = {}
    # the static description is a pure function of (workload name, GPU count): evaluate it in isolation
    import ast
    tree = ast.parse(src)
    keep = [n for n in tree.body if (isinstance(n, ast.FunctionDef) and n.name in ("static_config", "lenia_ring_worlds")) or
            (isinstance(n, ast.Assign) and any(getattr(tg, "id", "") in ("SHAPES", "REPLICA_WORKLOADS", "SLAB_WORKLOADS", "GOL_GHOST", "L2_BYTES",
                                                                           "LENIA1K_BATCH", "L2_FLUSH_BYTES") for tg in n.targets))]
    exec(compile(ast.Module(body=keep, type_ignores=[]), "bench_static", "exec"), ns)
    for name in ns["SHAPES"]:
        for n in (1, 2, 8):
            c = ns["static_config"](name, n)
            assert c["workload"] == name and "l2" in c and "partition" in c
    assert ns["static_config"]("lenia_1k", 1)["shape"] == [ns["LENIA1K_BATCH"], 1, 1024, 1024]


def test_marching_kernel_balanced_plan_tiles_every_column_once():
    """b200ca_lenia_march_plan / _plan_bands: the host copy of how a balanced launch of the marching Lenia kernel hands the
    (world, strip) columns of 16-row units to its CTAs (csrc/lenia_march.cu).  Every unit of every column must be in exactly
    one band, a CTA's bands must be consecutive, and the busiest CTA must not carry much more than its share."""
    import ctypes

    from pyca_b200 import _lib

    L = _lib.lib()
    for H, W, B in [(1024, 1024, 24), (1000, 500, 5), (4096, 4096, 1), (16384, 16384, 1), (2048, 16384, 1), (40, 300, 9), (2056, 230, 2),
                    (32, 32, 1), (1024, 1024, 1), (600, 460, 4)]:
        g, k, units, cols = (ctypes.c_int() for _ in range(4))
        assert L.b200ca_lenia_march_plan(H, W, B, ctypes.byref(g), ctypes.byref(k), ctypes.byref(units), ctypes.byref(cols)) == 0
        g, k, units, cols = g.value, k.value, units.value, cols.value
        assert units == -(-H // 16) and cols == -(-W // 224) * B and 1 <= g <= 296
        seen = np.zeros((cols, units), dtype=np.int32)
        buf = (ctypes.c_int * (3 * 64))()
        load = []
        for cta in range(g):
            n = L.b200ca_lenia_march_plan_bands(H, W, B, cta, buf, 64)
            assert 0 <= n < 64
            rounds = 0
            last = None
            for i in range(n):
                col, u0, nu = buf[3 * i], buf[3 * i + 1], buf[3 * i + 2]
                assert nu >= 1 and 0 <= u0 and u0 + nu <= units and 0 <= col < cols
                seen[col, u0:u0 + nu] += 1
                if last is not None:     # a run walks down a column and continues at the top of the next one
                    assert (col, u0) == (last[0] + 1, 0) and last[1] == units
                last = (col, u0 + nu)
                rounds += nu + 4         # PRO_UNITS rounds per band entered
            load.append(rounds)
            if k:
                assert n == 1
        assert (seen == 1).all(), (H, W, B)
        assert L.b200ca_lenia_march_plan_bands(H, W, B, g, buf, 64) == -1
        if g == 296 and units * cols >= 8 * g:
            # big launches: within a round or two of a perfect split inside each half of the grid; the first half (the CTA placed
            # first on every SM, which gets the larger share of its SM) carries 58 % of the rounds, or the split is even
            lo, hi = load[:g // 2], load[g // 2:]
            assert max(lo) - min(lo) <= 8 and max(hi) - min(hi) <= 8, (H, W, B, max(lo), min(lo), max(hi), min(hi))
            share = sum(lo) / (sum(lo) + sum(hi))
            assert abs(share - 0.5) < 0.01 or abs(share - 0.58) < 0.02, (H, W, B, share)
