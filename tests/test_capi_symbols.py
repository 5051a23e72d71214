"""CPU-side checks of the C ABI: the library loads without a GPU and exports every symbol the header declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "b200ca.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b200ca_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_expected_entry_points():
    syms = header_symbols()
    for must in ["b200ca_gol_step", "b200ca_lenia_step", "b200ca_mace_redistribute", "b200ca_nca_step",
                 "b200ca_gol_run_host", "b200ca_lenia_run_host", "b200ca_nca_run_host"]:
        assert must in syms


def test_library_exports_every_declared_symbol():
    from pyca_b200 import _lib

    assert os.path.exists(_lib.LIB_PATH), "libb200ca.so not built: run __graft_entry__.build()"
    L = ctypes.CDLL(_lib.LIB_PATH)
    for s in header_symbols():
        assert hasattr(L, s), f"{s} declared in include/b200ca.h but not exported"
    assert set(header_symbols()) == set(_lib.SIGNATURES), "ctypes signature table out of sync with the header"


def test_pure_host_helpers():
    from pyca_b200 import _lib

    L = _lib.lib()
    assert L.b200ca_version() == 100
    assert L.b200ca_gol_words_per_row(250) == 8 and L.b200ca_gol_words_per_row(256) == 8 and L.b200ca_gol_words_per_row(257) == 9
    assert L.b200ca_frame_pitch(1024) == 1056 and L.b200ca_frame_pitch(45) == 80
    assert L.b200ca_frame_elems(2, 3, 48, 64) == 2 * 3 * 80 * 96
    assert L.b200ca_nca_wprep_bytes(12, 128) == -1 and L.b200ca_nca_wprep_bytes(16, 128) > 0
    assert L.b200ca_nca_scratch_bytes(2, 10, 33) == 2 * 2 * 10 * 2 * 4


def test_no_cpu_fallback():
    """Without a GPU the product path must fail loudly, never compute on the CPU."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import pyca_b200

    with pytest.raises(Exception):
        pyca_b200.CA2D((32, 32), device="cuda")
    with pytest.raises(pyca_b200.B200CAError):
        pyca_b200.CA2D((32, 32), device="cpu")
    with pytest.raises(pyca_b200.B200CAError):
        pyca_b200.FramedState(1, 1, 32, 32, "cpu")


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "pyca_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "pyca_oracle" not in txt, f
