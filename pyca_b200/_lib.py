"""ctypes binding of ``libb200ca.so`` (C ABI declared in ``include/b200ca.h``).

This is the only place Python touches the CUDA code.  There is NO fallback: if the shared
library is missing, or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_int64, c_uint32, c_uint64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B200CA_LIB") or os.path.join(_HERE, "libb200ca.so")  # (B200CA_LIB: developer builds, e.g. -DB200CA_FFT_TRACE)

# name -> (restype, argtypes); mirrors include/b200ca.h one to one
SIGNATURES = {
    "b200ca_version": (c_int, []),
    "b200ca_last_error": (c_char_p, []),
    "b200ca_debug_variants": (c_char_p, [c_int]),
    "b200ca_device_info": (c_int, [ctypes.POINTER(c_int)] * 3),
    "b200ca_shutdown": (c_int, []),
    "b200ca_gol_words_per_row": (c_int, [c_int]),
    "b200ca_gol_pack": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_void_p]),
    "b200ca_gol_unpack": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "b200ca_gol_step": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_uint32, c_uint32, c_int, c_void_p]),
    "b200ca_gol_run": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_uint32, c_uint32, c_int, c_int, c_void_p]),
    "b200ca_gol_population": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "b200ca_gol_random_fill": (c_int, [c_void_p, c_int, c_int, c_int, c_uint64, c_int64, c_void_p]),
    "b200ca_gol_run_host": (c_int, [c_void_p, c_void_p, c_int, c_int, c_uint32, c_uint32, c_int]),
    "b200ca_frame_pitch": (c_int, [c_int]),
    "b200ca_frame_elems": (c_int64, [c_int, c_int, c_int, c_int]),
    "b200ca_frame_load": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "b200ca_frame_store": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "b200ca_frame_refresh": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "b200ca_lenia_kprep_elems": (c_int64, [c_int, c_int, c_int]),
    "b200ca_lenia_prepare_kernel": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "b200ca_lenia_tile_rows": (c_int, [c_int, c_int, c_int, c_int]),
    "b200ca_lenia_step": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                  c_int, c_float, c_int, c_int, c_int, c_int, c_void_p]),
    "b200ca_lenia_step_garbi": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_float,
                                        c_float, c_int, c_float, c_void_p, c_int, c_int, c_int, c_int, c_int, c_float, c_int,
                                        c_int, c_void_p]),
    "b200ca_lenia_fft_length": (c_int, [c_int, c_int]),
    "b200ca_lenia_fft_kernel_elems": (c_int64, [c_int, c_int, c_int, c_int, c_int]),
    "b200ca_lenia_fft_workspace_elems": (c_int64, [c_int, c_int, c_int, c_int, c_int]),
    "b200ca_lenia_fft_prepare_kernel": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "b200ca_lenia_step_fft": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                      c_int, c_int, c_float, c_int, c_int, c_void_p]),
    "b200ca_lenia_step_fft_phase": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
                                            c_int, c_int, c_int, c_float, c_int, c_int, c_int, c_void_p]),
    "b200ca_lenia_pick_algo": (c_int, [c_int, c_int, c_int, c_int]),
    "b200ca_lenia_pick_algo_ks": (c_int, [c_int, c_int, c_int, c_int, c_int]),
    "b200ca_lenia_march_supported": (c_int, [c_int, c_int]),
    "b200ca_lenia_march_kernel_elems": (c_int64, [c_int, c_int, c_int]),
    "b200ca_lenia_march_prepare_kernel": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "b200ca_lenia_march_band_rows": (c_int, [c_int, c_int, c_int, c_int]),
    "b200ca_lenia_march_plan": (c_int, [c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "b200ca_lenia_march_plan_bands": (c_int, [c_int, c_int, c_int, c_int, c_void_p, c_int]),
    "b200ca_lenia_step_march": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                        c_int, c_float, c_int, c_int, c_int, c_int, c_void_p]),
    "b200ca_lenia_march_slab_ctas": (c_int, [c_int, c_int, c_int, c_int]),
    "b200ca_lenia_step_march_slab": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                             c_int, c_float, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_uint32,
                                             c_void_p, c_void_p]),
    "b200ca_halo_push": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                 c_uint32, c_void_p, c_void_p, c_void_p]),
    "b200ca_halo_push_planes": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int64, c_void_p, c_void_p, c_void_p, c_void_p,
                                        c_void_p, c_void_p, c_uint32, c_void_p, c_void_p, c_void_p]),
    "b200ca_mace_redistribute": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float, c_int, c_float,
                                         c_int, c_void_p]),
    "b200ca_lenia_conv_sum": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "b200ca_mace_food_sense": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "b200ca_mace_food_decay": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float, c_float, c_void_p]),
    "b200ca_mace_food_consume": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float, c_float, c_int,
                                         c_float, c_float, c_void_p]),
    "b200ca_flow_reintegrate": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float, c_int, c_float, c_float,
                                        c_int, c_void_p]),
    "b200ca_frame_mass": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "b200ca_lenia_run_host": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                      c_int, c_float, c_int, c_float, c_float, c_int]),
    "b200ca_lenia_pipe_create": (c_int, [ctypes.POINTER(c_void_p), c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                         c_int, c_int, c_int, c_float, c_int, c_float, c_float, c_int]),
    "b200ca_lenia_pipe_submit": (c_int, [c_void_p, c_void_p, c_void_p, c_int]),
    "b200ca_lenia_pipe_drain": (c_int, [c_void_p]),
    "b200ca_lenia_pipe_destroy": (c_int, [c_void_p]),
    "b200ca_nca_wprep_bytes": (c_int64, [c_int, c_int]),
    "b200ca_nca_prepare_weights": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "b200ca_nca_scratch_bytes": (c_int64, [c_int, c_int, c_int]),
    "b200ca_nca_step": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_uint64, c_uint64, c_void_p, c_int, c_int, c_int,
                                c_int, c_int, c_int, c_void_p]),
    "b200ca_nca_step_shard": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_uint64, c_uint64, c_uint64, c_void_p, c_int, c_int,
                                      c_int, c_int, c_int, c_int, c_void_p]),
    "b200ca_nca_run_host": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_uint64, c_int, c_int, c_int, c_int, c_int,
                                    c_int, c_int]),
    "b200ca_draw_gol": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_float, c_float, c_void_p]),
    "b200ca_draw_lenia": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "b200ca_frame_pack_rgb": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "b200ca_draw_nca": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
}

FRAME = 16
MODE_LENIA, MODE_AFF = 0, 1
NCA_SIMT, NCA_TC = 0, 1


class B200CAError(RuntimeError):
    pass


_lib = None


def lib() -> ctypes.CDLL:
    """Loads the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise B200CAError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU / PyTorch fallback)")
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError if the .so does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def variants(reset: bool = True) -> list:
    """Kernel instantiations launched by this thread's compute calls since the last reset (test hook)."""
    v = lib().b200ca_debug_variants(1 if reset else 0)
    return [x for x in (v.decode() if v else "").split(";") if x]


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().b200ca_last_error()
        raise B200CAError(f"{what or 'b200ca call'} failed (rc={rc}): {msg.decode() if msg else ''}")


_fn_cache = {}


def call(name: str, *args) -> None:
    """Calls an int-returning entry point and raises on a non-zero return."""
    fn = _fn_cache.get(name)
    if fn is None:
        fn = _fn_cache[name] = getattr(lib(), name)
    rc = fn(*args)
    if rc != 0:
        check(rc, name)


class _NoGuard:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


_NOGUARD = _NoGuard()


def device_guard(device):
    """`torch.cuda.device(device)` only when `device` is not already current (the context manager costs microseconds,
    a Lenia 1024^2 step takes ~25)."""
    import torch

    if not isinstance(device, torch.device):
        device = torch.device(device)
    idx = device.index if device.index is not None else torch.cuda.current_device()
    if idx == torch.cuda.current_device():
        return _NOGUARD
    return torch.cuda.device(idx)


def ptr(t) -> int:
    """Device/host pointer of a torch tensor (0 for None)."""
    return 0 if t is None else t.data_ptr()


def current_stream(device=None) -> int:
    import torch

    return torch.cuda.current_stream(device).cuda_stream


def require_cuda(t, name="tensor"):
    if not t.is_cuda:
        raise B200CAError(f"{name} must live on a CUDA device (got {t.device}); this package has no CPU path")
