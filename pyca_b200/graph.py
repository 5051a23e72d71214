"""CUDA-graph replay of `step()` loops.

A 1024^2 Lenia step is two kernels and ~20 us of device time; issued one `step()` at a time, the Python / ctypes launch path
(and, with several ranks per box, its jitter) is of the same order.  `StepGraph` captures `reps` consecutive `step()` calls
of one or more automata -- the library's own launches, programmatic-dependent-launch attributes included -- into a CUDA
graph once and replays it with a single launch.  `reps` must bring every automaton's ping-pong buffers back to where they
started (an even number for the Lenia family and `CA2D`), so that a replay starts from the state the previous one left.

Only device-resident, steady-state stepping is captured: host-side decisions of a step (pending state assignments, in-place
edit detection, MaCE food respawn, torch-drawn NCA masks) are not part of a graph -- assign / edit between replays and call
`refresh()`; automata with `has_food` are refused, and so is `NeuralCA`: its stochastic update mask is keyed by the step
index, which `step()` passes to the kernel BY VALUE -- a capture would freeze it and every replay would redraw the same
`reps` masks instead of the reference's fresh mask per step (nca.py:256).

`refresh()` runs two warm-up steps outside the capture (lazy allocations, per-device function attributes); the state and
the frame counter are snapshotted before and restored after them, so constructing or refreshing a StepGraph does not advance
the automata.
"""
from __future__ import annotations

import torch

from . import _lib


class StepGraph:
    def __init__(self, autos, reps: int = 2):
        self.autos = list(autos) if isinstance(autos, (list, tuple)) else [autos]
        if reps < 2 or reps % 2:
            raise _lib.B200CAError("StepGraph: reps must be even (ping-pong buffers return to their start)")
        self.reps = reps
        for a in self.autos:
            if getattr(a, "has_food", False):
                raise _lib.B200CAError("StepGraph: this automaton takes host-side decisions inside step(); not capturable")
            if hasattr(a, "rng"):
                raise _lib.B200CAError("StepGraph: NeuralCA draws its update mask from the step index, passed by value at "
                                       "launch; a graph would replay the same masks forever -- step it from Python")
        self.device = torch.device(self.autos[0].device)
        self.graph = None
        self.refresh()

    def _prepare(self, a):
        if hasattr(a, "_sync_state"):
            a._sync_state()          # pending assignments / in-place edits are applied before the capture
        if hasattr(a, "_sync_to_device"):
            a._sync_to_device()

    def refresh(self):
        """(Re)captures the graph; call after assigning or editing a state between replays."""
        for a in self.autos:
            self._prepare(a)
            attr = "world" if hasattr(a, "world") else "state"
            saved, frames0 = getattr(a, attr).clone(), getattr(a, "frames", 0)
            a.step()                 # warm-up outside the capture: lazy allocations, function attributes
            a.step()
            setattr(a, attr, saved)  # ... and back to where the caller left the automaton
            if hasattr(a, "frames"):
                a.frames = frames0
            self._prepare(a)
        torch.cuda.synchronize(self.device)
        frames = [getattr(a, "frames", 0) for a in self.autos]
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(self.reps):
                for a in self.autos:
                    a.step()
        # nothing ran during the capture: take the frame counters back, remember what a replay advances them by
        self._frame_delta = [getattr(a, "frames", 0) - f for a, f in zip(self.autos, frames)]
        for a, f in zip(self.autos, frames):
            if hasattr(a, "frames"):
                a.frames = f
        self.graph = g

    def replay(self):
        """Advances every automaton by `reps` steps with one graph launch on the current stream."""
        self.graph.replay()
        for a, d in zip(self.autos, self._frame_delta):
            if hasattr(a, "frames"):
                a.frames += d

    @property
    def steps_per_replay(self) -> int:
        return self.reps * len(self.autos)
