#!/usr/bin/env python
"""Benchmark of the PyCA `Automaton.step()` hot path on B200 (contract in the task statement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME] [--no-extras]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One JSON line on stdout (rank 0).  metric = cell-updates/sec in GCUPS (BASELINE.json).
Workloads (synthetic data of the BASELINE.json shapes, parameters from tests/golden/params_*.npz which were
extracted from the reference's demo_data):
  lenia_1k        Lenia 1024x1024 single channel, demo_lenia/abominable_mongolic sliced to C=1   (default, N=1)
  lenia_16k       Lenia 16384x16384 single channel, row-slab sharded with NVLink halo exchange    (default, N>1)
  gol_64k         Game of Life 65536x65536, bit-packed, row-slab sharded for N>1
  gol_250         Game of Life 250x250 (simulate.py default world)
  mace_xchan_4k   MaCE-Lenia cross-channel 4096x4096 C=3, demo_macelenia_xchan/arborical_asbestosis
  nca_256         NCA inference 256 x 128x128 worlds (betta weights), batch-sharded for N>1
  flow_lenia_4k   FlowLenia 4096x4096 C=3, demo_macelenia/anencephalic_brakeman
`--impl reference` times the CPU implementation of the same path (the oracle port of the reference's torch
step, all host threads) on the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# Native libraries (NCCL's version banner, ...) write to file descriptor 1 behind Python's back; the contract is ONE JSON
# line on stdout, so fd 1 points at stderr while the benchmark runs and is restored for the final print.
_STDOUT_FD = os.dup(1)
os.dup2(2, 1)


def emit(line: dict):
    sys.stdout.flush()
    os.dup2(_STDOUT_FD, 1)
    print(json.dumps(line), flush=True)

import numpy as np  # noqa: E402
import torch  # noqa: E402

L2_FLUSH_BYTES = 256 << 20  # > 126 MB L2


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def load_params(name):
    from tests.util import load_npz, t

    p = load_npz(name)
    d = {k: t(p[k]) for k in ["mu", "sigma", "beta", "mu_k", "sigma_k", "weights"]}
    d["k_size"] = int(p["k_size"])
    return d, p


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            j = json.load(open(path))
            return {"hbm_gbs": float(j["hbm_gbs"]), "bf16_tflops": float(j.get("bf16_tflops_sustained", j["bf16_tflops"])),
                    "source": "measured"}
        except Exception:
            pass
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "source": "fallback"}


def ncu_traffic(key):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture (profiles/traffic.json), or None."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        return json.load(open(path)).get(key)
    except Exception:
        return None


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thr = threading.Thread(target=self._read, daemon=True)
            self.thr.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# workloads
# ------------------------------------------------------------------------------------------------
class Workload:
    name = ""
    dtype = "f32"
    bound = "hbm"
    scaling = "strong"

    def __init__(self, rank, world, device):
        self.rank, self.world, self.device = rank, world, device

    # ours
    def setup(self): ...
    def step(self): ...                 # one step of the hot path, state resident in HBM
    def cells_per_step(self): ...       # global cells updated per step (all ranks)
    def launches_per_step(self): ...    # kernels of ours per step on this rank
    def dominant(self): ...             # (kernel name, algorithmic bytes per launch, flops per launch, callable launching it once)
    def e2e_setup(self): ...
    def e2e_step(self): ...             # host buffers -> step -> host buffers; returns (h2d_bytes, d2h_bytes)
    def e2e_drain(self): pass           # waits for everything e2e_step queued (pipelined workloads)
    def config(self): ...
    # cpu (oracle port)
    def cpu_setup(self): ...
    def cpu_step(self): ...
    def cpu_cells_per_step(self): ...
    def cpu_sample(self): ...


E2E_DEPTH = 3


def _pipe_setup(wl, auto, shape):
    """End-to-end path: pyca_b200.LeniaHostPipe (b200ca_lenia_pipe_* in the C ABI) -- E2E_DEPTH lanes on their own streams;
    every world is copied up from its own pinned host buffer, stepped, and copied down into its own pinned host buffer."""
    from pyca_b200.pipeline import LeniaHostPipe

    wl.pipe = LeniaHostPipe(auto, depth=E2E_DEPTH)
    g = torch.Generator().manual_seed(77 + wl.rank)
    wl.pipe_in = [torch.rand(*shape, generator=g).pin_memory() for _ in range(E2E_DEPTH)]
    wl.pipe_out = [torch.empty(*shape).pin_memory() for _ in range(E2E_DEPTH)]
    wl.pipe_i = 0


def _pipe_step(wl):
    i = wl.pipe_i % E2E_DEPTH
    wl.pipe_i += 1
    wl.pipe.submit(wl.pipe_in[i], wl.pipe_out[i])
    n = wl.pipe_in[i].numel() * 4
    return n, n


class LeniaWL(Workload):
    """Lenia, single channel sliced from demo_lenia/abominable_mongolic (BASELINE configs[1]); slab-sharded for N>1."""

    def __init__(self, rank, world, device, H, W, C=1, name="lenia_1k", replicas=False):
        super().__init__(rank, world, device)
        self.H, self.W, self.C, self.name = H, W, C, name
        # replicas: one independent world per GPU (no communication, weak scaling); else one world in row slabs
        self.replicas = replicas and world > 1
        self.slab = None
        self.auto = None
        if replicas:
            self.scaling = "weak"  # one independent world per GPU at every N (N = 1 included)

    def _params(self):
        import pyca_b200

        d, _ = load_params("params_lenia_abominable_mongolic.npz")
        if self.C != 3:
            d = {k: (v[:, :self.C, :self.C].clone() if isinstance(v, torch.Tensor) else v) for k, v in d.items()}
        return pyca_b200.LeniaParams(param_dict=d, channels=self.C)

    def setup(self):
        import pyca_b200
        from pyca_b200.slab import LeniaSlabCUDA

        self.params = self._params()
        if self.world == 1 or self.replicas:
            g = torch.Generator().manual_seed(self.rank)
            self.h_state = torch.rand(1, self.C, self.H, self.W, generator=g).pin_memory()
            # A world smaller than L2 is benchmarked as a ring of independent worlds advanced round-robin, the ring sized
            # well above L2 (two state buffers + spectra workspace per world): every step then works on data that was
            # last touched `n_rot` steps ago, i.e. inputs larger than L2, with no flush kernel between the timed steps.
            per_world = self.H * self.W * self.C * 4 * (2 + 2 * max(1, self.C))
            self.n_rot = 1 if per_world > (126 << 20) else -(-(192 << 20) // per_world)
            self.autos = []
            for r in range(self.n_rot):
                st = self.h_state if r == 0 else torch.rand(1, self.C, self.H, self.W, generator=g)
                self.autos.append(pyca_b200.Lenia((self.H, self.W), num_channels=self.C, params=self.params, state_init=st,
                                                  device=self.device, conv_path=os.environ.get("B200CA_LENIA_PATH", "auto")))
            for a in self.autos:
                a.step()  # lazy initialisation (kernel spectra, workspace) happens outside the timed region for every world
            self.auto = self.autos[0]
            self.rot_i = 0
            self.rotating = self.n_rot > 1
            self.slab = None
            # small worlds: the ring's steps (one per world, twice round) are captured into a CUDA graph -- a 20 us step is
            # otherwise paced by the Python launch path and its jitter (B200CA_BENCH_GRAPH=0 issues every step from Python)
            self.graph = None
            if self.rotating and os.environ.get("B200CA_BENCH_GRAPH", "1") != "0":
                from pyca_b200.graph import StepGraph

                self.graph = StepGraph(self.autos, reps=2)
        else:
            self.slab = LeniaSlabCUDA(self.params, self.C, self.H, self.W, self.rank, self.world, device=self.device)
            g = torch.Generator(device=self.device).manual_seed(1000 + self.rank)
            dense = torch.rand(1, self.C, self.slab.hl, self.W, generator=g, device=self.device)
            self.slab.set_interior(dense)

    def step(self):
        if self.slab is None:
            self.autos[self.rot_i].step()
            self.rot_i = (self.rot_i + 1) % self.n_rot
        else:
            self.slab.step()

    def align(self):
        """(untimed, after the warm-up) finishes the current turn of the ring so that the timed steps start at world 0"""
        while getattr(self, "graph", None) is not None and self.rot_i != 0:
            self.step()

    def run(self, k):
        """exactly k steps of the ring (one world per step, round-robin): whole turns of the ring by graph replay, the rest
        step by step"""
        g = getattr(self, "graph", None)
        if g is not None:
            while k > 0 and self.rot_i != 0:   # finish the current turn of the ring
                self.step()
                k -= 1
            while k >= g.steps_per_replay:
                g.replay()
                k -= g.steps_per_replay
        for _ in range(k):
            self.step()

    def cells_per_step(self):
        return self.H * self.W * (self.world if self.replicas else 1)

    def _fft(self):
        if self.slab is not None:
            return self.slab.use_fft
        if self.auto is not None:
            return self.auto.use_fft
        import pyca_b200  # (reference arm on CPU: say which path the CUDA arm would take for this shape)

        return pyca_b200.lib().b200ca_lenia_fft_length(self.H, self.W) > 0

    def launches_per_step(self):
        if self.slab is not None:
            return self.slab.launches_per_step
        return self._fft_kernels() if self._fft() else 1

    def _fft_kernels(self):
        # N = 1024 transforms: row FFTs + fused (column FIR, inverse FFT, growth); N = 256: three kernels
        import pyca_b200

        return 2 if pyca_b200.lib().b200ca_lenia_fft_length(self.H, self.W) == 1024 else 3

    def dominant(self):
        cells = self.H * self.W if self.replicas else self.H * self.W // self.world
        if self._fft():
            # the kernels of the step are timed together (row FFTs, then column FIR + inverse FFT + growth)
            k = "lenia_fft_rows + lenia_fft_firinv (whole step)" if self._fft_kernels() == 2 else \
                "lenia_fft_rows + lenia_fft_fir + lenia_fft_inv (whole step)"
            return {"kernel": k, "bytes": cells * self.C * 8, "fma": 0}
        return {"kernel": "lenia_conv_kernel", "bytes": cells * self.C * 8, "fma": cells * self.C * self.C * 496}

    def config(self):
        path = f"row-FFT x column-FIR hybrid + growth + clamp, {self._fft_kernels()} kernels/step" if self._fft() else \
            "direct folded conv (TMA tiles) + growth + clamp, 1 kernel/step"
        return {"workload": self.name, "shape": [1, self.C, self.H, self.W], "k_size": 31,
                "params": "demo_lenia/abominable_mongolic" + (f" sliced to C={self.C}" if self.C != 3 else ""),
                "path": path,
                "partition": "single GPU" if self.world == 1 else (
                    f"{self.world} independent worlds, one per GPU, no communication" if self.replicas else
                    f"{self.world} row slabs, NCCL halo exchange" + (" overlapped with interior" if not (self.slab and self.slab.use_fft) else "")),
                "l2": self._l2_note()}

    def _l2_note(self):
        if getattr(self, "rotating", False):
            return (f"inputs larger than L2: {self.n_rot} independent worlds advanced round-robin (one world per step; a world is "
                    f"revisited after {self.n_rot} steps), no flush kernel between timed steps"
                    + ("; steps issued by CUDA-graph replay (2 turns of the ring per launch), remainder from Python"
                       if getattr(self, "graph", None) is not None else ""))
        if self.H * self.W * self.C * 8 // max(1, 1 if self.replicas else self.world) > (126 << 20):
            return "inputs larger than L2"
        return f"flushed between timed steps ({L2_FLUSH_BYTES >> 20} MiB write)"

    def e2e_setup(self):
        import pyca_b200

        self.h_out = torch.empty(1, self.C, self.H, self.W).pin_memory()
        _pipe_setup(self, self.auto, (1, self.C, self.H, self.W))

    def e2e_step(self):
        return _pipe_step(self)

    def e2e_drain(self):
        self.pipe.drain()

    def e2e_step_sync(self):
        # host tensor in -> public API (state assignment + step()) -> host tensor out, one world, nothing overlapped
        a = self.auto
        a.state = self.h_state.to(self.device, non_blocking=True)
        a.step()
        self.h_out.copy_(a.state, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        n = self.h_state.numel() * 4
        return n, n

    def cpu_setup(self):
        from oracle import pyca_oracle as orc
        from tests.util import load_npz, t

        self.orc = orc
        p = self.params.param_dict
        self.cK = orc.lenia_kernel(p["beta"].cpu(), p["mu_k"].cpu(), p["sigma_k"].cpu(), 31)
        self.cH, self.cW = min(self.H, 4096), min(self.W, 4096)
        g = torch.Generator().manual_seed(0)
        self.cstate = torch.rand(1, self.C, self.cH, self.cW, generator=g)
        self.cKf = orc.kernel_fft(self.cK, self.cH, self.cW)
        self.cmu, self.csg, self.cw = p["mu"].cpu(), p["sigma"].cpu(), p["weights"].cpu()

    def cpu_step(self):
        self.cstate = self.orc.lenia_step(self.cstate, self.cK, self.cmu, self.csg, self.cw, 0.1, self.cKf)

    def cpu_cells_per_step(self):
        return self.cH * self.cW

    def cpu_sample(self):
        return f"Lenia C={self.C} {self.cH}x{self.cW} oracle step (torch fft2 path of lenia.py:430-469)"


class GolWL(Workload):
    dtype = "u32 bit-planes"

    def __init__(self, rank, world, device, H, W, name):
        super().__init__(rank, world, device)
        self.H, self.W, self.name = H, W, name
        self.ghost = 64   # ghost rows per side = generations between halo exchanges (redundant rows: 2*64 / 8192 at 8 GPUs)
        # the 250x250 world fits in shared memory: CA2D.step(nsteps) runs 100 generations per launch (a "step" of this
        # workload is 100 generations; cells_per_step counts them all)
        self.steps_per_call = 100 if H * ((W + 31) // 32) <= 4096 and world == 1 else 1
        if world > 1:
            # a slab generation takes ~20 us at 8 GPUs: issue 16 generations per call (b200ca_gol_run loops the launches in C)
            # so the host does not become the bottleneck; a "step" of this workload is then 16 generations
            self.steps_per_call = 16

    def setup(self):
        import pyca_b200
        from pyca_b200.slab import GolSlabCUDA

        if self.world == 1:
            self.pw = pyca_b200.PackedWorld(self.H, self.W, self.device)
            self.pw.random_fill(seed=2024)
            self.slab = None
        else:
            self.slab = GolSlabCUDA(self.H, self.W, self.rank, self.world, self.ghost, self.device)
            self.slab.random_fill(seed=2024)

    def step(self):
        if self.slab is None:
            self.pw.step(0b1100, 0b1000, self.steps_per_call)
        else:
            self.slab.run(self.steps_per_call)

    def cells_per_step(self):
        return self.H * self.W * self.steps_per_call

    def launches_per_step(self):
        return 1

    def dominant(self):
        cells = self.H * self.W // self.world
        return {"kernel": ("gol_rows_kernel" if self.W <= 256 and self.H <= 1024 else "gol_small_kernel") if self.steps_per_call > 1
                else "gol_fast_kernel", "bytes": cells // 4 * self.steps_per_call,
                "fma": 0}

    def config(self):
        return {"workload": self.name, "shape": [self.H, self.W], "rule": "B3/S23", "layout": "1 bit/cell",
                "generations_per_launch": self.steps_per_call,
                "partition": "single GPU" if self.world == 1 else
                f"{self.world} row slabs, {self.ghost} ghost rows exchanged every {self.ghost} generations (NCCL)",
                "l2": "inputs larger than L2" if self.H * self.W // 8 // self.world > (126 << 20) else
                f"flushed between timed steps ({L2_FLUSH_BYTES >> 20} MiB write)"}

    def e2e_setup(self):
        import pyca_b200

        # the reference-facing call: CA2D with an int32 world on the host (ca2d.py:29)
        self.eH, self.eW = min(self.H, 8192), min(self.W, 8192)
        self.h_world = (torch.rand(self.eH, self.eW) < 0.5).to(torch.int32).pin_memory()
        self.h_wout = torch.empty_like(self.h_world).pin_memory()
        self.ca = pyca_b200.CA2D((self.eH, self.eW), device=self.device)

    def e2e_step(self):
        self.ca.world = self.h_world.to(self.device, non_blocking=True)
        self.ca.step()
        self.h_wout.copy_(self.ca.world, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        n = self.h_world.numel() * 4
        return n, n

    def e2e_cells(self):
        return self.eH * self.eW

    def cpu_setup(self):
        from oracle import pyca_oracle as orc

        self.orc = orc
        self.cH, self.cW = min(self.H, 4096), min(self.W, 4096)
        rng = np.random.default_rng(0)
        self.cworld = torch.from_numpy((rng.random((self.cH, self.cW)) < 0.5).astype(np.int32))

    def cpu_step(self):
        # the reference's own formulation in torch (rolls + where), which is what PyCA runs on CPU (ca2d.py:195-209)
        w = self.cworld
        a, b = w.roll(-1, 0), w.roll(1, 0)
        cnt = a + b + w.roll(-1, 1) + w.roll(1, 1) + a.roll(1, 1) + b.roll(1, 1) + a.roll(-1, 1) + b.roll(-1, 1)
        self.cworld = torch.where(w == 1, (0b1100 >> cnt) & 1, (0b1000 >> cnt) & 1).to(torch.int)

    def cpu_cells_per_step(self):
        return self.cH * self.cW

    def cpu_sample(self):
        return f"GoL {self.cH}x{self.cW} int32 torch roll/where step (ca2d.py:195-209)"


class MaceWL(Workload):
    def __init__(self, rank, world, device, H, W, name="mace_xchan_4k"):
        super().__init__(rank, world, device)
        self.H, self.W, self.name = H, W, name
        self.scaling = "weak"

    def setup(self):
        import pyca_b200

        d, _ = load_params("params_xchan_arborical_asbestosis.npz")
        self.params = pyca_b200.LeniaParams(param_dict=d)
        g = torch.Generator().manual_seed(self.rank)
        self.h_state = torch.rand(1, 3, self.H, self.W, generator=g).pin_memory()
        self.auto = pyca_b200.MaCELeniaXChan((self.H, self.W), params=self.params, state_init=self.h_state, device=self.device)

    def step(self):
        self.auto.step()

    def cells_per_step(self):
        return self.H * self.W * self.world  # independent worlds per rank

    def launches_per_step(self):
        return 4 if self.auto.use_fft else 2

    def dominant(self):
        cells = self.H * self.W
        if self.auto.use_fft:
            return {"kernel": "affinity: lenia_fft_rows + lenia_fft_fir + lenia_fft_inv", "bytes": cells * 3 * 8, "fma": 0}
        return {"kernel": "lenia_conv_kernel", "bytes": cells * 3 * 8, "fma": cells * 9 * 496}

    def config(self):
        return {"workload": self.name, "shape": [1, 3, self.H, self.W], "params": "demo_macelenia_xchan/arborical_asbestosis",
                "beta": 6, "alpha": 0.03, "partition": "one world per GPU",
                "l2": "inputs larger than L2" if self.H * self.W * 12 > (126 << 20) else "flushed between timed steps"}

    def e2e_setup(self):
        import pyca_b200

        self.h_out = torch.empty(1, 3, self.H, self.W).pin_memory()
        _pipe_setup(self, self.auto, (1, 3, self.H, self.W))

    def e2e_step(self):
        return _pipe_step(self)

    def e2e_drain(self):
        self.pipe.drain()

    def e2e_step_sync(self):
        a = self.auto
        a.state = self.h_state.to(self.device, non_blocking=True)
        a.step()
        self.h_out.copy_(a.state, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        n = self.h_state.numel() * 4
        return n, n

    def cpu_setup(self):
        from oracle import pyca_oracle as orc

        self.orc = orc
        p = self.params.param_dict
        self.cH, self.cW = min(self.H, 1024), min(self.W, 1024)
        self.cK = orc.lenia_kernel(p["beta"].cpu(), p["mu_k"].cpu(), p["sigma_k"].cpu(), 31)
        self.cKf = orc.kernel_fft(self.cK, self.cH, self.cW)
        g = torch.Generator().manual_seed(0)
        self.cstate = torch.rand(1, 3, self.cH, self.cW, generator=g)
        self.cmu, self.csg, self.cw = p["mu"].cpu(), p["sigma"].cpu(), p["weights"].cpu()

    def cpu_step(self):
        self.cstate = self.orc.xchan_step(self.cstate, self.cK, self.cmu, self.csg, self.cw, 6.0, 0.03, self.cKf)

    def cpu_cells_per_step(self):
        return self.cH * self.cW

    def cpu_sample(self):
        return f"MaCE-XChan C=3 {self.cH}x{self.cW} oracle step (fft2 + unfold path of macelenia.py:89-159)"


class FlowWL(MaceWL):
    """FlowLenia with the demo_macelenia parameters its DEFAULTS entry points at."""

    def setup(self):
        import pyca_b200

        d, _ = load_params("params_mace_anencephalic_brakeman.npz")
        self.params = pyca_b200.LeniaParams(param_dict=d)
        g = torch.Generator().manual_seed(self.rank)
        self.h_state = (1.5 * torch.rand(1, 3, self.H, self.W, generator=g)).pin_memory()
        self.auto = pyca_b200.FlowLenia((self.H, self.W), params=self.params, state_init=self.h_state, device=self.device)

    def launches_per_step(self):
        return 3 if self.auto.use_fft else 2

    def dominant(self):
        cells = self.H * self.W
        return {"kernel": "affinity (lenia_fft_rows + lenia_fft_firinv) + flow_rt_kernel", "bytes": cells * 3 * 8, "fma": 0}

    def config(self):
        return {"workload": self.name, "shape": [1, 3, self.H, self.W], "params": "demo_macelenia/anencephalic_brakeman",
                "dd": 2, "sigma_rt": 0.65, "partition": "one world per GPU",
                "l2": "inputs larger than L2" if self.H * self.W * 12 > (126 << 20) else "flushed between timed steps"}

    def e2e_setup(self):
        self.h_out = torch.empty(1, 3, self.H, self.W).pin_memory()

    def e2e_step(self):
        return self.e2e_step_sync()

    def e2e_drain(self):
        pass

    def cpu_setup(self):
        from oracle import pyca_oracle as orc

        self.orc = orc
        p = self.params.param_dict
        self.cH, self.cW = min(self.H, 1024), min(self.W, 1024)
        self.cK = orc.lenia_kernel(p["beta"].cpu(), p["mu_k"].cpu(), p["sigma_k"].cpu(), 31)
        self.cKf = orc.kernel_fft(self.cK, self.cH, self.cW)
        g = torch.Generator().manual_seed(0)
        self.cstate = 1.5 * torch.rand(1, 3, self.cH, self.cW, generator=g)
        self.cmu, self.csg, self.cw = p["mu"].cpu(), p["sigma"].cpu(), p["weights"].cpu()

    def cpu_step(self):
        self.cstate = self.orc.flow_step(self.cstate, self.cK, self.cmu, self.csg, self.cw, Kf=self.cKf)

    def cpu_sample(self):
        return f"FlowLenia C=3 {self.cH}x{self.cW} oracle step (fft2 + conv2d sobel + 25 rolls, flow_lenia.py:97-253)"


class NcaWL(Workload):
    def __init__(self, rank, world, device, B=256, H=128, W=128, name="nca_256", impl=None):
        super().__init__(rank, world, device)
        self.B, self.H, self.W, self.name = B, H, W, name
        self.impl = impl
        self.scaling = "strong"
        self.Bl = B // world

    def setup(self):
        import pyca_b200
        from tests.util import load_npz, t

        fx = load_npz("nca_betta.npz")
        self.fx = fx
        self.wts = pyca_b200.NCAWeights(t(fx["W1"]), t(fx["b1"]), t(fx["W2"]), t(fx["b2"]), self.device)
        self.Bl = self.B // self.world
        g = torch.Generator().manual_seed(self.rank)
        self.h_state = torch.rand(self.Bl, 16, self.H, self.W, generator=g).pin_memory()
        self.a = self.h_state.to(self.device)
        self.b = torch.empty_like(self.a)
        self.scratch = torch.empty(pyca_b200.lib().b200ca_nca_scratch_bytes(self.Bl, self.H, self.W), dtype=torch.uint8,
                                   device=self.device)
        self.t = 0
        if self.impl is None:
            self.impl = 1
            try:
                pyca_b200.nca_step(self.a, self.wts, out=self.b, seed=1, step_index=0, scratch=self.scratch, impl=1)
            except pyca_b200.B200CAError:
                self.impl = 0

    def step(self):
        import pyca_b200

        # (first_world: every rank draws its slice of the one global update mask of the 256-world batch)
        pyca_b200.nca_step(self.a, self.wts, out=self.b, seed=1, step_index=self.t, scratch=self.scratch, impl=self.impl,
                           first_world=self.rank * self.Bl)
        self.a, self.b = self.b, self.a
        self.t += 1

    def cells_per_step(self):
        return self.B * self.H * self.W

    def launches_per_step(self):
        return 2

    def dominant(self):
        cells = self.Bl * self.H * self.W
        return {"kernel": "nca_update_tc_kernel" if self.impl == 1 else "nca_update_simt_kernel", "bytes": cells * 128,
                "fma": 0 if self.impl == 1 else cells * 6144, "flops_algorithmic": cells * 16384,
                "flops_issued_tf32": cells * 2 * 3 * (128 * 40 + 16 * 128)}

    def config(self):
        return {"workload": self.name, "shape": [self.B, 16, self.H, self.W], "weights": "NCA_models/betta", "rng": "in-kernel Philox",
                "mlp": "tcgen05 split-tf32" if self.impl == 1 else "fp32 SIMT", "partition": f"batch sharded over {self.world} GPU(s)",
                "l2": "inputs larger than L2" if self.Bl * 16 * self.H * self.W * 4 > (126 << 20) else "flushed between timed steps"}

    def e2e_setup(self):
        self.h_out = torch.empty_like(self.h_state).pin_memory()

    def e2e_step(self):
        import pyca_b200

        d = self.h_state.to(self.device, non_blocking=True)
        out = pyca_b200.nca_step(d, self.wts, out=self.b, seed=1, step_index=self.t, scratch=self.scratch, impl=self.impl)
        self.h_out.copy_(out, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        n = self.h_state.numel() * 4
        return n, n

    def cpu_setup(self):
        from oracle import pyca_oracle as orc
        from tests.util import t

        self.orc = orc
        self.cB = min(self.B, 16)
        g = torch.Generator().manual_seed(0)
        self.cstate = torch.rand(self.cB, 16, self.H, self.W, generator=g)
        self.cw = [t(self.fx[k]) for k in ("W1", "b1", "W2", "b2")]

    def cpu_step(self):
        mask = torch.rand(self.cB, 1, self.H, self.W, device=self.cstate.device) <= 0.5
        self.cstate = self.orc.nca_step(self.cstate, *self.cw, mask)

    def cpu_cells_per_step(self):
        return self.cB * self.H * self.W

    def cpu_sample(self):
        return f"NCA {self.cB}x16x{self.H}x{self.W} oracle step (conv2d + Linear path of nca.py:239-263)"


def make_workload(name, rank, world, device):
    if name == "lenia_1k":
        return LeniaWL(rank, world, device, 1024, 1024, 1, name, replicas=True)
    if name == "lenia3_1k":
        return LeniaWL(rank, world, device, 1024, 1024, 3, name)
    if name == "lenia_4k":
        return LeniaWL(rank, world, device, 4096, 4096, 1, name)
    if name == "lenia_seg":
        return LeniaWL(rank, world, device, 4096, 3976, 1, name)
    if name == "lenia_16k":
        return LeniaWL(rank, world, device, 16384, 16384, 1, name)
    if name == "gol_64k":
        return GolWL(rank, world, device, 65536, 65536, name)
    if name == "gol_16k":
        return GolWL(rank, world, device, 16384, 16384, name)
    if name == "gol_250":
        return GolWL(rank, world, device, 250, 250, name)
    if name == "mace_xchan_4k":
        return MaceWL(rank, world, device, 4096, 4096, name)
    if name == "mace_xchan_1k":
        return MaceWL(rank, world, device, 1024, 1024, name)
    if name == "nca_256":
        return NcaWL(rank, world, device, 256, 128, 128, name)
    if name == "flow_lenia_1k":
        return FlowWL(rank, world, device, 1024, 1024, name)
    if name == "flow_lenia_4k":
        return FlowWL(rank, world, device, 4096, 4096, name)
    raise SystemExit(f"unknown workload {name}")


# ------------------------------------------------------------------------------------------------
# timing
# ------------------------------------------------------------------------------------------------
def barrier(world):
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()


def time_steps(wl, steps, warmup, world, flush):
    """W untimed + exactly K timed steps; per-step CUDA events (the L2 flush between steps is untimed); returns
    total ms over the K steps (max over ranks is taken by the caller)."""
    for _ in range(warmup):
        wl.step()
    barrier(world)
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for i in range(steps):
        if flush is not None:
            flush.zero_()
        evs[i][0].record()
        wl.step()
        evs[i][1].record()
    barrier(world)
    return sum(a.elapsed_time(b) for a, b in evs)


def time_block(wl, steps, warmup, world):
    """K steps back to back between one pair of events (used when the working set exceeds L2)."""
    for _ in range(warmup):
        wl.step()
    if hasattr(wl, "align"):
        wl.align()
    barrier(world)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    if hasattr(wl, "run"):
        wl.run(steps)
    else:
        for _ in range(steps):
            wl.step()
    e1.record()
    barrier(world)
    return e0.elapsed_time(e1)


def cpu_baseline(wl, budget_s=12.0):
    torch.set_num_threads(os.cpu_count() or 1)
    wl.cpu_setup()
    wl.cpu_step()  # warm-up
    t0 = time.perf_counter()
    n = 0
    while True:
        wl.cpu_step()
        n += 1
        el = time.perf_counter() - t0
        if el > budget_s or (n >= 200 and el > 0.8 * budget_s):
            break
    gcups = wl.cpu_cells_per_step() * n / el / 1e9
    return {"value": gcups, "unit": "GCUPS", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{n} x {wl.cpu_sample()}, {el:.1f} s", "ms_per_step": el / n * 1e3}


def torch_gpu_baseline(wl, device, budget_s=1.5):
    """Informative column: the same torch formulation of the step -- the oracle port of the reference's ops,
    i.e. what `python simulate.py -d cuda` executes -- on this GPU instead of the host cores."""
    wl.cpu_setup()
    for k, v in list(vars(wl).items()):
        if k.startswith("c") and isinstance(v, torch.Tensor):
            setattr(wl, k, v.to(device))
        elif k.startswith("c") and isinstance(v, list) and v and isinstance(v[0], torch.Tensor):
            setattr(wl, k, [x.to(device) for x in v])
    for _ in range(2):
        wl.cpu_step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 0
    while True:
        wl.cpu_step()
        n += 1
        if n % 4 == 0:
            torch.cuda.synchronize()
            if time.perf_counter() - t0 > budget_s or n >= 400:
                break
    torch.cuda.synchronize()
    el = time.perf_counter() - t0
    return {"value": wl.cpu_cells_per_step() * n / el / 1e9, "unit": "GCUPS", "ms_per_step": el / n * 1e3,
            "sample": f"{n} x {wl.cpu_sample()} with every tensor on the GPU"}


def run_reference(args, rank, world):
    """--impl reference: the CPU implementation of the path (oracle port), all host threads, rank 0 only."""
    if rank != 0:
        return
    name = args.workload or "lenia_1k"
    wl = make_workload(name, 0, 1, "cpu")
    if hasattr(wl, "_params"):
        wl.params = wl._params()
    elif isinstance(wl, MaceWL):
        import pyca_b200

        d, _ = load_params("params_mace_anencephalic_brakeman.npz" if isinstance(wl, FlowWL) else "params_xchan_arborical_asbestosis.npz")
        wl.params = pyca_b200.LeniaParams(param_dict=d)
    elif isinstance(wl, NcaWL):
        from tests.util import load_npz

        wl.fx = load_npz("nca_betta.npz")
    torch.set_num_threads(os.cpu_count() or 1)
    if torch.cuda.is_available():
        # Measured on the B200 box (tools/cpu_arm_probe.py): the same torch CPU step runs at 0.98 ms in a process that has
        # created a CUDA context and a pinned allocation and at 1.9 ms in one that has not (16 threads either way).  The CPU
        # arm gets the faster setting -- it is also the one `cpu_baseline` of the CUDA arm runs under.  No kernel is launched.
        torch.zeros(1, device="cuda")
        torch.empty(1 << 20).pin_memory()
    wl.cpu_setup()
    # one "step" of this arm = a bounded sample of `reps` consecutive CPU steps (>= 50 ms of work), so that K steps are not
    # dominated by start-up effects; a 4 s untimed spin-up precedes the W warm-up steps (on the box the step time keeps
    # falling for the first ~3 s of a fresh process: 1.5, 1.1, 0.98 ms in consecutive 2 s windows)
    t0 = time.perf_counter()
    n0 = 0
    while time.perf_counter() - t0 < 4.0 or n0 < 1:
        wl.cpu_step()
        n0 += 1
    t_one = (time.perf_counter() - t0) / n0
    reps = max(1, min(int(0.05 / t_one + 0.999), 1000))
    for _ in range(args.warmup):
        for _ in range(reps):
            wl.cpu_step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        for _ in range(reps):
            wl.cpu_step()
    el = time.perf_counter() - t0
    v = wl.cpu_cells_per_step() * args.steps * reps / el / 1e9
    cfg = dict(wl.config())
    cfg["cpu_sample"] = wl.cpu_sample()
    line = {"impl": "reference", "metric": "cell-updates/sec", "value": v, "unit": "GCUPS", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": el / (args.steps * reps) * 1e3, "higher_is_better": True, "scaling": wl.scaling,
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
            "cpu_baseline": {"value": v, "unit": "GCUPS", "cores": torch.get_num_threads(), "kind": "port",
                             "sample": f"{args.steps} steps x {reps} consecutive CPU updates of {wl.cpu_sample()}"},
            "e2e": {"value": v, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None)
    ap.add_argument("--no-extras", action="store_true", help="skip the per-model summary of the other BASELINE configs")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU path)"
    torch.cuda.set_device(local_rank)
    device = f"cuda:{local_rank}"
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.distributed.init_process_group("nccl", device_id=torch.device(device))
    name = args.workload or "lenia_1k"
    wl = make_workload(name, rank, world, device)
    wl.setup()
    dom = wl.dominant()
    big = dom["bytes"] > (126 << 20) or getattr(wl, "rotating", False)
    flush = None if big else torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=device)

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    total_ms = time_block(wl, args.steps, args.warmup, world) if big else time_steps(wl, args.steps, args.warmup, world, flush)
    clocks = sampler.stop() if sampler else None
    tms = torch.tensor([total_ms], dtype=torch.float64, device=device)
    if world > 1:
        torch.distributed.all_reduce(tms, op=torch.distributed.ReduceOp.MAX)
    total_ms = float(tms.item())
    ms_per_step = total_ms / args.steps
    value = wl.cells_per_step() / (ms_per_step * 1e-3) / 1e9

    # dominant kernel: average launch duration measured live with CUDA events on the launching stream
    peaks = measured_peaks()
    kern_ms = ms_per_step  # one kernel per step for the single-kernel workloads; refined below for MaCE/NCA
    if isinstance(wl, MaceWL) and not isinstance(wl, FlowWL) and rank == 0:
        from pyca_b200 import _lib

        a = wl.auto
        evs = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a._sync_state()
            e0.record()
            a._conv(a._aff, _lib.MODE_AFF)
            e1.record()
            evs.append((e0, e1))
        torch.cuda.synchronize()
        kern_ms = sum(x.elapsed_time(y) for x, y in evs) / len(evs)
    ach = dom["bytes"] / (kern_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": ach / peaks["hbm_gbs"],
                "traffic": ncu_traffic(name), "kernel": dom["kernel"], "kernel_ms": kern_ms, "peak_source": peaks["source"],
                "algorithmic_bytes_per_launch": dom["bytes"]}
    if dom.get("fma"):
        sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
        fma_peak = 148 * 128 * sm_mhz * 1e6 / 1e12
        roofline["fp32_fma"] = {"achieved_tfma_s": dom["fma"] / (kern_ms * 1e-3) / 1e12, "peak_tfma_s_at_sampled_clock": fma_peak,
                                "frac": dom["fma"] / (kern_ms * 1e-3) / 1e12 / fma_peak,
                                "note": "folded direct conv: 496 FMA per cell and channel pair; this pipe, not HBM, bounds the kernel"}
    if isinstance(wl, NcaWL) and wl.impl == 1:
        alg = dom["flops_algorithmic"] / (kern_ms * 1e-3) / 1e12
        iss = dom["flops_issued_tf32"] / (kern_ms * 1e-3) / 1e12
        roofline["tensor"] = {"algorithmic_tflops": alg, "issued_tf32_tflops": iss, "tf32_peak_tflops": peaks["bf16_tflops"] / 2,
                              "frac_issued_of_tf32_peak": iss / (peaks["bf16_tflops"] / 2),
                              "note": "split-tf32: 3 tcgen05.mma per product for fp32-grade accuracy; tf32 peak taken as half the "
                                      "measured bf16 peak"}

    e2e = None
    if world == 1 or getattr(wl, "replicas", False):
        # every rank drives its own world through the public API with pinned host buffers; max time over ranks
        wl.e2e_setup()

        def run_e2e(stepfn, n):
            for _ in range(3):
                stepfn()
            wl.e2e_drain()
            barrier(world)
            t0 = time.perf_counter()
            for _ in range(n):
                hb, db = stepfn()
            wl.e2e_drain()
            torch.cuda.synchronize()
            el = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=device)
            if world > 1:
                torch.distributed.all_reduce(el, op=torch.distributed.ReduceOp.MAX)
            return float(el.item()), hb, db

        n_e2e = max(6, min(4 * args.steps, 200))
        el, hb, db = run_e2e(wl.e2e_step, n_e2e)
        cells = wl.e2e_cells() if hasattr(wl, "e2e_cells") else wl.cells_per_step()
        piped = hasattr(wl, "pipe")
        e2e = {"value": cells * n_e2e / el / 1e9, "unit": "GCUPS", "h2d_bytes_per_step": hb, "d2h_bytes_per_step": db,
               "ms_per_step": el / n_e2e * 1e3, "steps": n_e2e,
               "api": (f"pyca_b200.LeniaHostPipe(auto, depth={E2E_DEPTH}).submit(pinned h_in, pinned h_out) = b200ca_lenia_pipe_submit: per "
                       "world H2D -> step -> D2H, independent worlds overlapped on separate streams") if piped else
                      "Automaton.state = <pinned host tensor>; step(); state -> pinned host tensor"}
        if piped:
            el_s, _, _ = run_e2e(wl.e2e_step_sync, max(3, min(args.steps, 20)))
            n_s = max(3, min(args.steps, 20))
            e2e["one_world_synchronous"] = {"value": cells * n_s / el_s / 1e9, "ms_per_step": el_s / n_s * 1e3,
                                            "api": "Automaton.state = <pinned host tensor>; step(); state -> pinned host tensor; "
                                                   "synchronize (no overlap)"}
    elif rank == 0:
        e2e = {"value": None, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
               "note": "end-to-end host-buffer path is measured for single-GPU / replica workloads (a sharded world is generated per slab on device)"}

    extras = None
    if rank == 0 and world == 1 and not args.no_extras:
        extras = {}
        for other in ["lenia3_1k", "gol_250", "gol_64k", "mace_xchan_4k", "nca_256", "flow_lenia_4k"]:
            if other == name:
                continue
            try:
                ow = make_workload(other, 0, 1, device)
                ow.setup()
                od = ow.dominant()
                k = 3 if other in ("mace_xchan_4k", "flow_lenia_4k") else 10
                ms = time_block(ow, k, 3, 1) / k
                ent = {"gcups": ow.cells_per_step() / (ms * 1e-3) / 1e9, "ms_per_step": ms,
                       "hbm_frac_algorithmic": od["bytes"] / (ms * 1e-3) / 1e9 / peaks["hbm_gbs"], "config": ow.config()}
                if not args.no_cpu:
                    cb = cpu_baseline(ow, budget_s=3.0)
                    ent["cpu_gcups"] = cb["value"]
                    ent["cpu_sample"] = cb["sample"]
                    try:
                        ent["torch_gpu_gcups"] = torch_gpu_baseline(ow, device)["value"]
                    except Exception as e:
                        ent["torch_gpu_gcups"] = None
                        ent["torch_gpu_error"] = repr(e)[:120]
                extras[other] = ent
                del ow
                torch.cuda.empty_cache()
            except Exception as e:  # keep the headline line even if an extra fails
                extras[other] = {"error": repr(e)[:200]}

        # draw() + Automaton.worldmap of a 1024^2 3-channel Lenia world, fused kernel + 3 B/cell D2H against
        # the reference's expression on the same GPU (torch clamp, float permute, 12 B/cell D2H, host-side uint8 cast)
        try:
            dw = make_workload("lenia3_1k", 0, 1, device)
            dw.setup()
            a = dw.auto

            def fused():
                a.draw()
                return a.worldmap

            def torch_way():
                wm = torch.clamp(a.state[0][:3].clone(), 0.0, 1.0)
                return (255 * wm.permute(2, 1, 0)).detach().cpu().numpy().astype(dtype=np.uint8)

            res = {}
            for nm, fn in (("fused_ms", fused), ("torch_reference_expression_ms", torch_way)):
                for _ in range(3):
                    fn()
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for _ in range(20):
                    fn()
                torch.cuda.synchronize()
                res[nm] = (time.perf_counter() - t0) / 20 * 1e3
            res["identical"] = bool(np.array_equal(fused(), torch_way()))
            res["workload"] = "draw() + worldmap, Lenia 1x3x1024x1024 -> (1024,1024,3) uint8 on the host"
            extras["draw_lenia3_1k"] = res
            del dw
            torch.cuda.empty_cache()
        except Exception as e:
            extras["draw_lenia3_1k"] = {"error": repr(e)[:200]}

    sharded = None
    if world > 1 and not args.no_extras:
        # BASELINE configs[4]: one big torus in row slabs with NVLink halo exchange; raw values at N GPUs and, for
        # reference, the same world on one GPU of this box (every rank runs it alone, rank 0's number is kept)
        sharded = {}
        for big in ["lenia_16k", "gol_64k"]:
            try:
                ent = {}
                sw = make_workload(big, rank, world, device)
                sw.setup()
                k = 30  # a slab step is 0.2-0.9 ms: a window of several ms, so that one host hiccup does not decide the figure
                t = torch.tensor([time_block(sw, k, 3, world) / k], dtype=torch.float64, device=device)
                torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
                ent["ms_per_step"] = float(t.item())
                ent["gcups"] = sw.cells_per_step() / (ent["ms_per_step"] * 1e-3) / 1e9
                ent["config"] = sw.config()
                del sw
                torch.cuda.empty_cache()
                s1 = make_workload(big, 0, 1, device)
                s1.setup()
                ms1 = time_block(s1, k, 3, 1) / k
                ent["single_gpu_ms_per_step"] = ms1
                ent["single_gpu_gcups"] = s1.cells_per_step() / (ms1 * 1e-3) / 1e9
                del s1
                torch.cuda.empty_cache()
                torch.distributed.barrier()
                sharded[big] = ent
            except Exception as e:
                sharded[big] = {"error": repr(e)[:300]}

    cpu = None
    torch_gpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_baseline(wl)
        try:
            torch_gpu = torch_gpu_baseline(wl, device)
        except Exception as e:
            torch_gpu = {"value": None, "error": repr(e)[:160]}

    if rank == 0:
        line = {"metric": "cell-updates/sec", "value": value, "unit": "GCUPS", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
                "scaling": wl.scaling, "vs_baseline": None, "dtype": wl.dtype, "data": "synthetic", "config": wl.config(),
                "clocks": clocks, "e2e": e2e, "gpu_launches": wl.launches_per_step() * args.steps, "roofline": roofline,
                "cpu_baseline": cpu, "torch_gpu_baseline": torch_gpu, "per_model": extras, "sharded": sharded}
        emit(line)
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
