"""Streaming of HOST-resident worlds through the B200 step: the end-to-end path of ``bench.py``'s ``e2e`` figure.

The reference's headless loop on host data is ``auto.state = t; auto.step(); t = auto.state`` -- one world after the
other, every copy serialised with the step.  A world only depends on its own previous state, so independent worlds (a
batch of simulations, the frames of a parameter sweep) can overlap: while world ``i`` steps, world ``i+1`` is on its
way up the PCIe link and world ``i-1`` on its way down (the B200 has separate copy engines per direction).
``HostPipeline`` keeps ``depth`` automata, each bound to its own CUDA stream, and deals the submitted worlds to them
round-robin; a lane is reused only after its previous job -- including the device->host copy -- has completed.

Every job still performs, inside the pipeline, the full host->device copy of its input from pinned memory, ``nsteps``
calls of the automaton's own ``step()`` and the device->host copy of the resulting state.
"""
from __future__ import annotations

import torch

from . import _lib


class _Lane:
    def __init__(self, auto, device):
        self.auto = auto
        self.stream = torch.cuda.Stream(device=device)
        self.done = torch.cuda.Event()
        self.busy = False


class HostPipeline:
    """``HostPipeline(make_automaton, depth=3)``; ``submit(h_in, h_out)`` per world; ``drain()`` at the end.

    ``make_automaton`` builds one drop-in automaton (``Lenia``, ``MaCELeniaXChan``, ``CA2D``, ``NeuralCA`` ...);
    ``attr`` names its public state tensor (``state`` or ``world``).  ``h_in`` / ``h_out`` must be pinned host tensors.
    """

    def __init__(self, make_automaton, depth: int = 3, device="cuda", attr: str = "state"):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.B200CAError(f"HostPipeline needs a CUDA device (got {device}); there is no CPU path")
        if depth < 1:
            raise _lib.B200CAError("HostPipeline: depth must be >= 1")
        self.attr = attr
        self.lanes = []
        for _ in range(depth):
            lane = _Lane(None, self.device)
            with torch.cuda.stream(lane.stream):  # the automaton's own set-up kernels run on the lane's stream
                lane.auto = make_automaton()
            lane.stream.synchronize()
            self.lanes.append(lane)
        self.submitted = 0

    def submit(self, h_in: torch.Tensor, h_out: torch.Tensor, nsteps: int = 1):
        """Queues: H2D of ``h_in`` -> ``nsteps`` x ``step()`` -> D2H into ``h_out``.  Returns the lane's completion event."""
        if not (h_in.is_pinned() and h_out.is_pinned()):
            raise _lib.B200CAError("HostPipeline.submit: h_in and h_out must be pinned host tensors")
        lane = self.lanes[self.submitted % len(self.lanes)]
        self.submitted += 1
        if lane.busy:
            lane.done.synchronize()  # the lane's previous world has landed in its h_out
        with torch.cuda.stream(lane.stream):
            setattr(lane.auto, self.attr, h_in.to(self.device, non_blocking=True))
            for _ in range(nsteps):
                lane.auto.step()
            h_out.copy_(getattr(lane.auto, self.attr), non_blocking=True)
            lane.done.record(lane.stream)
        lane.busy = True
        return lane.done

    def drain(self):
        for lane in self.lanes:
            if lane.busy:
                lane.done.synchronize()
                lane.busy = False


class LeniaHostPipe:
    """The same pipeline for the Lenia family with the lanes inside the C library (``b200ca_lenia_pipe_*``): one C call
    per world queues H2D -> frame -> ``nsteps`` steps -> unframe -> D2H on the next lane, so the host spends ~20 us per
    world instead of a dozen Python-level tensor operations.

    ``auto`` is a ``Lenia`` / ``MaCELenia`` / ``MaCELeniaXChan`` drop-in (or any object with the reference's ``k``, ``mu``,
    ``sigma``, ``weights``, ``dt`` attributes); its parameters are copied into the pipe at construction.
    """

    def __init__(self, auto, depth: int = 3, family: int = None):
        import ctypes

        from .lenia import MaCELenia, MaCELeniaXChan

        # the pipe implements Lenia / MaCE / MaCE-XChan steps with Gaussian growth only; anything else would silently run
        # as plain Lenia (FlowLenia subclasses Lenia) or with placeholder mu / sigma (g_arbi)
        if type(auto).__name__ == "FlowLenia" or hasattr(auto, "_flow_args") or hasattr(auto, "dd"):
            raise _lib.B200CAError("LeniaHostPipe: FlowLenia is not supported by the host pipe")
        if getattr(auto, "g_arbi", False):
            raise _lib.B200CAError("LeniaHostPipe: Fourier growth functions (g_arbi) are not supported by the host pipe")
        if getattr(auto, "has_food", False):
            raise _lib.B200CAError("LeniaHostPipe: MaCE food (has_food=True) is not supported by the host pipe")
        if family is None:
            family = 2 if isinstance(auto, MaCELeniaXChan) else 1 if isinstance(auto, MaCELenia) else 0
        k = auto.k.float().contiguous()
        B, CK, C, ks, _ = k.shape
        self.shape = (B, C, auto.h, auto.w)
        self.device = torch.device(auto.device)
        self._h = ctypes.c_void_p()
        mu, sg, w = (t.float().contiguous() for t in (auto.mu, auto.sigma, auto.weights))
        with _lib.device_guard(self.device):
            torch.cuda.synchronize(self.device)
            _lib.call("b200ca_lenia_pipe_create", ctypes.byref(self._h), _lib.ptr(k), _lib.ptr(mu), _lib.ptr(sg), _lib.ptr(w),
                      B, C, CK // C, ks, auto.h, auto.w, float(auto.dt), family, float(getattr(auto, "b", 0.0)),
                      float(getattr(auto, "alpha", 0.0)), depth)

    def submit(self, h_in: torch.Tensor, h_out: torch.Tensor, nsteps: int = 1):
        for t in (h_in, h_out):
            if t.is_cuda or not t.is_pinned() or t.dtype != torch.float32 or not t.is_contiguous() or tuple(t.shape) != self.shape:
                raise _lib.B200CAError(f"LeniaHostPipe.submit: need pinned contiguous fp32 host tensors of shape {self.shape}")
        with _lib.device_guard(self.device):
            _lib.call("b200ca_lenia_pipe_submit", self._h, h_in.data_ptr(), h_out.data_ptr(), nsteps)

    def drain(self):
        with _lib.device_guard(self.device):
            _lib.call("b200ca_lenia_pipe_drain", self._h)

    def close(self):
        if self._h:
            _lib.call("b200ca_lenia_pipe_destroy", self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
