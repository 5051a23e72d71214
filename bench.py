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
`--impl reference` times the reference's own CPU implementation of the same path on the same workload: the UNMODIFIED
reference from baseline/_ref (baseline/install_ref.py; cpu_baseline.kind "reference"), all host threads; the oracle port
only if that install is absent (kind "port").  That arm imports neither pyca_b200 nor the CUDA library.

Timing: W warm-up steps, then R blocks of EXACTLY K steps, each block between its own pair of CUDA events (barrier +
synchronize around the whole timed region); R is chosen so that the region lasts >= 250 ms whatever K is, and
ms_per_step = median block time (max over ranks per block) / K.  `timing` in the JSON line says what really ran.
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


def _pin_malloc():
    """glibc's dynamic mmap / trim thresholds make the CPU step time depend on the allocation HISTORY of the process (the torch
    CPU step allocates and frees 8-32 MB temporaries every step; whether they come from mmap, with a page fault per 4 KB, or
    from the heap depends on what was freed before): the same 16-thread Lenia step measured 0.97 ms inside the CUDA arm and
    1.81 ms in the reference arm in round 1.  Both arms therefore fix the thresholds up front (heap allocation up to 32 MB,
    no trimming), which is also the setting under which the CPU path is fastest."""
    try:
        import ctypes

        libc = ctypes.CDLL("libc.so.6")
        libc.mallopt(-3, 32 << 20)   # M_MMAP_THRESHOLD (its maximum)
        libc.mallopt(-1, 1 << 30)    # M_TRIM_THRESHOLD
        libc.mallopt(-2, 64 << 20)   # M_TOP_PAD
    except Exception:
        pass


_pin_malloc()

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


def cpu_params(npz_name, C=3):
    """Sanitised parameter tensors for the oracle port (no pyca_b200 import): leniaparams.py:68-87."""
    from oracle import pyca_oracle as orc

    d, _ = load_params(npz_name)
    if C != 3:
        d = {k: (v[:, :C, :C].clone() if isinstance(v, torch.Tensor) else v) for k, v in d.items()}
    return orc.sanitize_params(d)


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
# workload descriptions shared by BOTH arms (the reference arm prints the same `config` without importing pyca_b200)
# ------------------------------------------------------------------------------------------------
SHAPES = {
    "lenia_1k": ("lenia", 1, 1024, 1024), "lenia_1k_single": ("lenia", 1, 1024, 1024), "lenia3_1k": ("lenia", 3, 1024, 1024), "lenia_4k": ("lenia", 1, 4096, 4096),
    "lenia_seg": ("lenia", 1, 4096, 3976), "lenia_16k": ("lenia", 1, 16384, 16384),
    "gol_64k": ("gol", 0, 65536, 65536), "gol_16k": ("gol", 0, 16384, 16384), "gol_250": ("gol", 0, 250, 250),
    "mace_xchan_4k": ("xchan", 3, 4096, 4096), "mace_xchan_1k": ("xchan", 3, 1024, 1024),
    "nca_256": ("nca", 16, 128, 128), "flow_lenia_1k": ("flow", 3, 1024, 1024), "flow_lenia_4k": ("flow", 3, 4096, 4096),
}
REPLICA_WORKLOADS = ("lenia_1k",)          # independent worlds on every GPU at N > 1 (weak scaling, no communication)
# The headline workload steps a BATCH of independent 1024^2 worlds per GPU through the batch dimension of the Automaton API
# (`Lenia(size=(B,H,W))`, lenia.py:19-29, parameters by `LeniaParams.expand(B)`, batch_params.py:217-236): 24 worlds = 201 MB
# of state (in + out), i.e. inputs larger than L2 by construction, and enough CTAs to fill the machine -- a single 1024^2
# world is a latency measurement (one wave of CTAs, ~20 us per step whatever the kernel does; reported as `single_world`).
LENIA1K_BATCH = 24
SLAB_WORKLOADS = ("lenia_16k", "lenia_4k", "gol_64k", "gol_16k")
GOL_GHOST = 64
L2_BYTES = 126 << 20


def lenia_ring_worlds(C, H, W, steps=0):
    """Worlds in the round-robin ring of a Lenia workload smaller than L2 (1 = no ring)."""
    per_world = H * W * C * 4 * (2 + 2 * max(1, C))
    if per_world > L2_BYTES:
        return 1
    return max(-(-(192 << 20) // per_world), (steps + 1) // 2 if steps >= 18 else 0)


def static_config(name, n):
    fam, C, H, W = SHAPES[name]
    if fam == "lenia":
        per_rank = H * W * C * 8 // (1 if name in REPLICA_WORKLOADS else n)
        if name in REPLICA_WORKLOADS:
            return {"workload": name, "shape": [LENIA1K_BATCH, C, H, W], "k_size": 31,
                    "params": "demo_lenia/abominable_mongolic" + (f" sliced to C={C}" if C != 3 else "") + f", expanded to {LENIA1K_BATCH} worlds",
                    "partition": (f"{LENIA1K_BATCH} independent {H}x{W} worlds per GPU stepped together through the batch dimension of the "
                                  f"Automaton API (size=(B,H,W)), {n} GPU(s), no communication"),
                    "l2": f"inputs larger than L2: {LENIA1K_BATCH * H * W * C * 8 >> 20} MB of state read + written per step, no flush"}
        cfg = {"workload": name, "shape": [1, C, H, W], "k_size": 31,
               "params": "demo_lenia/abominable_mongolic" + (f" sliced to C={C}" if C != 3 else ""),
               "partition": "single GPU" if n == 1 else (f"{n} independent worlds, one per GPU, no communication"
                                                         if name in REPLICA_WORKLOADS else f"{n} row slabs, halo rows over NVLink"),
               "l2": ("inputs larger than L2: a ring of independent worlds (>= 192 MB of state + workspace) advanced round-robin, one "
                      "world per step, no flush kernel between timed steps") if lenia_ring_worlds(C, H, W) > 1 else
                     ("inputs larger than L2" if per_rank > L2_BYTES else f"flushed between timed steps ({L2_FLUSH_BYTES >> 20} MiB write)")}
    elif fam == "gol":
        small = H * ((W + 31) // 32) <= 4096
        cfg = {"workload": name, "shape": [H, W], "rule": "B3/S23", "layout": "1 bit/cell",
               "generations_per_step": 100 if (small and n == 1) else (16 if n > 1 else 1),
               "partition": "single GPU" if n == 1 else f"{n} row slabs, {GOL_GHOST} ghost rows exchanged every {GOL_GHOST} generations",
               "l2": "inputs larger than L2" if H * W // 8 // n > L2_BYTES else f"flushed between timed steps ({L2_FLUSH_BYTES >> 20} MiB write)"}
    elif fam == "xchan":
        cfg = {"workload": name, "shape": [1, 3, H, W], "params": "demo_macelenia_xchan/arborical_asbestosis", "beta": 6, "alpha": 0.03,
               "partition": "one world per GPU", "l2": "inputs larger than L2" if H * W * 12 > L2_BYTES else "flushed between timed steps"}
    elif fam == "flow":
        cfg = {"workload": name, "shape": [1, 3, H, W], "params": "demo_macelenia/anencephalic_brakeman", "dd": 2, "sigma_rt": 0.65,
               "partition": "one world per GPU", "l2": "inputs larger than L2" if H * W * 12 > L2_BYTES else "flushed between timed steps"}
    else:
        B = 256
        cfg = {"workload": name, "shape": [B, 16, H, W], "weights": "NCA_models/betta", "rng": "Bernoulli(1/2) update mask per cell and step",
               "partition": f"batch sharded over {n} GPU(s)",
               "l2": "inputs larger than L2" if B // n * 16 * H * W * 4 > L2_BYTES else "flushed between timed steps"}
    return cfg


# ------------------------------------------------------------------------------------------------
# CPU arm: the UNMODIFIED reference (baseline/_ref) stepped through its own public API on the host cores
# ------------------------------------------------------------------------------------------------
class RefArm:
    """One BASELINE workload on the reference's own classes (`Automaton.step()` on device='cpu'), at a size the host finishes in
    bounded time (the `sample`).  Falls back to the oracle port (tests infrastructure, oracle/pyca_oracle.py) only when
    baseline/_ref is absent -- kind says which."""

    def __init__(self, name):
        self.name = name
        self.fam, self.C, H, W = SHAPES[name]
        cap = {"lenia": 4096, "gol": 4096, "xchan": 1024, "flow": 1024, "nca": 128}[self.fam]
        self.H, self.W = min(H, cap), min(W, cap)
        self.B = 16 if self.fam == "nca" else (LENIA1K_BATCH if name in REPLICA_WORKLOADS else 1)
        self.kind = "reference"
        try:
            sys.path.insert(0, os.path.join(ROOT, "baseline"))
            from ref_loader import load_reference

            self.ref = load_reference()
        except Exception as e:
            log(f"reference install not usable ({e!r}): CPU arm falls back to the oracle port")
            self.ref = None
            self.kind = "port"

    def _param_file(self):
        sub = {"lenia": "demo_lenia/abominable_mongolic.pt", "xchan": "demo_macelenia_xchan/arborical_asbestosis.pt",
               "flow": "demo_macelenia/anencephalic_brakeman.pt"}[self.fam]
        return os.path.join(self.ref.root, "demo_data", sub)

    def setup(self):
        torch.set_num_threads(os.cpu_count() or 1)
        torch.set_grad_enabled(False)
        g = torch.Generator().manual_seed(0)
        if self.ref is None:
            self._port = make_workload(self.name, 0, 1, "cpu")
            self._port.cpu_setup()
            self.step = self._port.cpu_step
            self.H, self.W = getattr(self._port, "cH", self.H), getattr(self._port, "cW", self.W)
            self.B = getattr(self._port, "cB", self.B)
            return
        R = self.ref
        if self.fam == "gol":
            self.auto = R.CA2D((self.H, self.W), device="cpu")
            self.auto.world = (torch.rand(self.H, self.W, generator=g) < 0.5).to(torch.int)
        elif self.fam == "nca":
            self.auto = R.NeuralCA((self.H, self.W), os.path.join(R.root, "demo_data", "NCA_models"), device="cpu")
            self.auto.load_model(os.path.join(R.root, "demo_data", "NCA_models", "betta.pt"))
            self.auto.state = torch.rand(self.B, 16, self.H, self.W, generator=g)
        else:
            C = self.C
            P = R.LeniaParams(from_file=self._param_file())
            if C != 3:
                d = {k: (v[:, :C, :C].clone() if isinstance(v, torch.Tensor) else v) for k, v in P.param_dict.items()}
                P = R.LeniaParams(param_dict=d, channels=C)
            cls = {"lenia": R.Lenia, "xchan": R.MaCELeniaXChan, "flow": R.FlowLenia}[self.fam]
            kw = dict(num_channels=C, params=P, device="cpu")
            if self.fam == "lenia":
                kw["dt"] = 0.1
            if self.B > 1:
                # The batch of worlds, stepped ONE WORLD AT A TIME: on the host that is the reference's faster mode -- one
                # (B,H,W) automaton with expanded parameters works on 24 x 8 MB spectra at once, falls out of the caches and
                # ran 4x slower per cell here (0.056 vs 0.23 GCUPS on 8 cores) -- so the CPU arm gets its best case.
                self.autos = [cls((self.H, self.W), state_init=torch.rand(1, C, self.H, self.W, generator=g), **kw) for _ in range(self.B)]

                def step_all():
                    for a in self.autos:
                        a.step()

                self.step = step_all
                return
            s0 = torch.rand(1, C, self.H, self.W, generator=g) * (1.5 if self.fam == "flow" else 1.0)
            self.auto = cls((self.H, self.W), state_init=s0, **kw)
        self.step = self.auto.step

    def cells_per_step(self):
        return self.B * self.H * self.W

    def sample(self):
        what = {"lenia": (f"{self.B} worlds, one Lenia.step() each (world by world: the host's faster mode) of " if self.B > 1 else "") +
                         f"Lenia C={self.C} {self.H}x{self.W} Lenia.step() (lenia.py:430-469)",
                "xchan": f"MaCE-XChan C=3 {self.H}x{self.W} MaCELeniaXChan.step() (mace_lenia_xchan.py:59-76, macelenia.py:89-159)",
                "flow": f"FlowLenia C=3 {self.H}x{self.W} FlowLenia.step() (flow_lenia.py:246-253)",
                "gol": f"GoL {self.H}x{self.W} int32 CA2D.step() (ca2d.py:195-209)",
                "nca": f"NCA {self.B}x16x{self.H}x{self.W} NeuralCA.step() (nca.py:67-72, 239-263)"}[self.fam]
        return what + (" of the unmodified reference in baseline/_ref" if self.kind == "reference" else " -- oracle port")

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


E2E_DEPTH = 8


def _pipe_setup(wl, auto, shape):
    """End-to-end path: pyca_b200.LeniaHostPipe (b200ca_lenia_pipe_* in the C ABI) -- E2E_DEPTH lanes on their own streams;
    every world is copied up from its own pinned host buffer, stepped, and stored into its own pinned host buffer (by the
    unframing kernel, over PCIe: b200ca_lenia_pipe_submit)."""
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
        self.steps_hint = 0   # K of the timed blocks (set by main() before setup(): sizes the ring / the captured graph)
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
            self.n_rot = lenia_ring_worlds(self.C, self.H, self.W, self.steps_hint)
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
            # small worlds: a timed block of K steps is captured into ONE CUDA graph -- a 10-20 us step is otherwise paced by
            # the Python launch path and its jitter (B200CA_BENCH_GRAPH=0 issues every step from Python).  The graph steps the
            # first K/2 worlds of the ring twice round (2 x K/2 = K steps; an even count per world brings its ping-pong buffers
            # back, so the graph can be replayed); K/2 worlds of 16 MB must exceed L2, hence K >= 18.
            self.graph = None
            self.graph_steps = 0
            k = self.steps_hint
            if self.rotating and k >= 18 and os.environ.get("B200CA_BENCH_GRAPH", "1") != "0":
                from pyca_b200.graph import StepGraph

                self.graph = StepGraph(self.autos[:k // 2], reps=2)
                self.graph_steps = self.graph.steps_per_replay
        else:
            # one torus in row slabs: halo exchange inside the step kernel (P2P stores over NVLink, one launch per step);
            # NCCL send/recv slabs if symmetric memory / peer access is not available (B200CA_SLAB_HALO=nccl forces them)
            from pyca_b200.slab import LeniaSlabP2P

            self.halo = "p2p"
            try:
                if os.environ.get("B200CA_SLAB_HALO", "p2p") != "p2p":
                    raise RuntimeError("forced")
                self.slab = LeniaSlabP2P(self.params, self.C, self.H, self.W, self.rank, self.world, device=self.device)
            except Exception as e:
                log(f"P2P-halo slabs unavailable ({e!r}): NCCL send/recv slabs")
                self.halo = "nccl"
                self.slab = LeniaSlabCUDA(self.params, self.C, self.H, self.W, self.rank, self.world, device=self.device)
            self._fill_slab()

    def _full_world(self):
        """the whole torus, generated identically on every rank (same device generator seed)"""
        g = torch.Generator(device=self.device).manual_seed(4242)
        return torch.rand(1, self.C, self.H, self.W, generator=g, device=self.device)

    def _fill_slab(self):
        full = self._full_world()
        self.slab.set_interior(full[:, :, self.slab.r0:self.slab.r1])
        del full

    def check_against_single_gpu(self, steps=3):
        """N-GPU slabs == the same world stepped on ONE GPU (this rank's, the same kernels): max |diff| over this rank's rows."""
        import pyca_b200

        self._fill_slab()
        for _ in range(steps):
            self.slab.step()
        torch.cuda.synchronize()
        single = pyca_b200.Lenia((self.H, self.W), num_channels=self.C, params=self.params, state_init=self._full_world(),
                                 device=self.device)
        for _ in range(steps):
            single.step()
        err = (self.slab.interior() - single.state[:, :, self.slab.r0:self.slab.r1]).abs().max().item()
        res = {"steps": steps, "max_abs_err": err, "tolerance": 2e-6,
               "timed_out": bool(self.slab.timed_out()) if hasattr(self.slab, "timed_out") else False}
        self.single = single
        return res

    def step(self):
        if self.slab is None:
            self.autos[self.rot_i].step()
            self.rot_i = (self.rot_i + 1) % self.n_rot
        else:
            self.slab.step()

    def run(self, k):
        """exactly k steps (one world per step): one graph replay of 2 x (K/2) steps when the block was captured, the
        remainder (K odd) / everything else step by step from Python, round-robin over the ring"""
        g = getattr(self, "graph", None)
        if g is not None and k >= self.graph_steps:
            g.replay()
            k -= self.graph_steps
        for _ in range(k):
            self.step()

    def issue_note(self):
        if getattr(self, "graph", None) is not None:
            return (f"each timed block of K steps = one CUDA-graph replay of {self.graph_steps} steps ({self.graph_steps // 2} worlds of "
                    f"the {self.n_rot}-world ring, twice round)" + (" + 1 step from Python" if self.steps_hint % 2 else ""))
        return "steps issued one by one from Python (ctypes call per step)"

    def cells_per_step(self):
        return self.H * self.W * (self.world if self.replicas else 1)

    def _fft(self):
        if self.slab is not None:
            return self.slab.use_fft
        if self.auto is not None:
            return self.auto.use_fft
        return True

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
        return static_config(self.name, self.world)

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
        p = cpu_params("params_lenia_abominable_mongolic.npz", self.C)
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


class LeniaBatchWL(LeniaWL):
    """The headline workload (BASELINE configs[1]): LENIA1K_BATCH independent 1024x1024 single-channel worlds per GPU, stepped
    together through the batch dimension of the Automaton API -- `Lenia(size=(B,H,W), params=params.expand(B))`, the
    reference's own way of running several worlds (lenia.py:19-29, batch_params.py:217-236).  One step = one generation of
    every world = one kernel launch; 201 MB of state are read + written per step (inputs larger than L2)."""

    def __init__(self, rank, world, device, H, W, C=1, name="lenia_1k"):
        super().__init__(rank, world, device, H, W, C, name, replicas=True)
        self.replicas = world > 1
        self.Bw = LENIA1K_BATCH

    def setup(self):
        import pyca_b200

        self.params = self._params()
        g = torch.Generator().manual_seed(self.rank)
        st = torch.rand(self.Bw, self.C, self.H, self.W, generator=g)
        self.h_state = st[:1].clone().pin_memory()
        self.auto = pyca_b200.Lenia((self.Bw, self.H, self.W), num_channels=self.C, params=self.params.expand(self.Bw), state_init=st,
                                    device=self.device, conv_path=os.environ.get("B200CA_LENIA_PATH", "auto"))
        self.auto.step()
        self.slab = None
        self.rotating = True      # (inputs exceed L2 by construction: no flush)
        self.graph = None

    def step(self):
        self.auto.step()

    def run(self, k):
        for _ in range(k):
            self.auto.step()

    def issue_note(self):
        return "one launch per step (ctypes call from Python), all worlds of the batch in the same launch"

    def cells_per_step(self):
        return self.Bw * self.H * self.W * self.world

    def launches_per_step(self):
        return 1 if self.auto.conv_algo in ("march", "direct") else self._fft_kernels()

    def dominant(self):
        cells = self.Bw * self.H * self.W
        k = {"march": "lenia_march_kernel", "fft_rows": "lenia_fft_rows + lenia_fft_firinv (whole step)", "direct": "lenia_conv_kernel"}
        return {"kernel": k[self.auto.conv_algo], "bytes": cells * self.C * 8, "fma": 0}

    def e2e_setup(self):
        import pyca_b200

        # end to end is per WORLD: each 1024^2 world travels host -> device -> step -> host on its own (LeniaHostPipe lanes)
        self.auto1 = pyca_b200.Lenia((self.H, self.W), num_channels=self.C, params=self.params, state_init=self.h_state, device=self.device)
        self.h_out = torch.empty(1, self.C, self.H, self.W).pin_memory()
        _pipe_setup(self, self.auto1, (1, self.C, self.H, self.W))

    def e2e_cells(self):
        return self.H * self.W * self.world

    def e2e_step_sync(self):
        a = self.auto1
        a.state = self.h_state.to(self.device, non_blocking=True)
        a.step()
        self.h_out.copy_(a.state, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        n = self.h_state.numel() * 4
        return n, n


class GolWL(Workload):
    dtype = "u32 bit-planes"

    def __init__(self, rank, world, device, H, W, name):
        super().__init__(rank, world, device)
        self.H, self.W, self.name = H, W, name
        self.ghost = GOL_GHOST   # ghost rows per side = generations between halo exchanges (redundant rows: 2*64 / 8192 at 8 GPUs)
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
            # ghost rows pushed by one P2P kernel per exchange (symmetric memory); NCCL send/recv slabs as the fallback
            from pyca_b200.slab import GolSlabP2P

            self.halo = "p2p"
            try:
                if os.environ.get("B200CA_SLAB_HALO", "p2p") != "p2p":
                    raise RuntimeError("forced")
                self.slab = GolSlabP2P(self.H, self.W, self.rank, self.world, self.ghost, self.device)
            except Exception as e:
                log(f"P2P-halo GoL slabs unavailable ({e!r}): NCCL send/recv slabs")
                self.halo = "nccl"
                self.slab = GolSlabCUDA(self.H, self.W, self.rank, self.world, self.ghost, self.device)
            self.slab.random_fill(seed=2024)

    def step(self):
        if self.slab is None:
            self.pw.step(0b1100, 0b1000, self.steps_per_call)
        else:
            self.slab.run(self.steps_per_call)

    def cells_per_step(self):
        return self.H * self.W * self.steps_per_call

    def check_against_single_gpu(self, steps=48):
        """bit-exact: the slabs' owned rows and the global population against the whole torus on one GPU"""
        import pyca_b200

        self.slab.random_fill(seed=2024)
        self.slab.run(steps)
        ref = pyca_b200.PackedWorld(self.H, self.W, self.device)
        ref.random_fill(seed=2024)
        ref.step(0b1100, 0b1000, steps)
        same = bool(torch.equal(self.slab.owned()[:, :ref.wpr], ref.bits[self.slab.r0:self.slab.r1, :ref.wpr]))
        pop = torch.tensor([self.slab.population()], device=self.device)
        torch.distributed.all_reduce(pop)
        return {"steps": steps, "bit_exact": same, "population": int(pop.item()), "population_single_gpu": ref.population()}

    def launches_per_step(self):
        return 1

    def dominant(self):
        cells = self.H * self.W // self.world
        return {"kernel": ("gol_rows_kernel" if self.W <= 256 and self.H <= 1024 else "gol_small_kernel") if self.steps_per_call > 1
                else "gol_fast_kernel", "bytes": cells // 4 * self.steps_per_call,
                "fma": 0}

    def config(self):
        return static_config(self.name, self.world)

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


class MaceSlabWL(Workload):
    """MaCE-XChan 4096^2 C=3 as ONE torus in row slabs: two exchanges per step (affinity rows, state rows),
    each one P2P push kernel per rank (NCCL send/recv as the fallback); `sharded` block of the N > 1 runs."""
    scaling = "strong"

    def __init__(self, rank, world, device, H, W, name="mace_xchan_4k"):
        super().__init__(rank, world, device)
        self.H, self.W, self.name = H, W, name

    def _full_world(self):
        g = torch.Generator(device=self.device).manual_seed(777)
        return torch.rand(1, 3, self.H, self.W, generator=g, device=self.device)

    def setup(self):
        import pyca_b200
        from pyca_b200.slab import MaceSlabCUDA, MaceSlabP2P

        d, _ = load_params("params_xchan_arborical_asbestosis.npz")
        self.params = pyca_b200.LeniaParams(param_dict=d)
        self.halo = "p2p"
        try:
            if os.environ.get("B200CA_SLAB_HALO", "p2p") != "p2p":
                raise RuntimeError("forced")
            self.slab = MaceSlabP2P(self.params, 3, self.H, self.W, self.rank, self.world, device=self.device, beta=6.0, alpha=0.03)
        except Exception as e:
            log(f"P2P-halo MaCE slabs unavailable ({e!r}): NCCL send/recv slabs")
            self.halo = "nccl"
            self.slab = MaceSlabCUDA(self.params, 3, self.H, self.W, self.rank, self.world, device=self.device, beta=6.0, alpha=0.03)
        self.slab.set_interior(self._full_world()[:, :, self.slab.r0:self.slab.r1])

    def step(self):
        self.slab.step()

    def cells_per_step(self):
        return self.H * self.W

    def launches_per_step(self):
        return self.slab.launches_per_step

    def dominant(self):
        return {"kernel": "affinity + mace_fast", "bytes": self.H * self.W * 24 // self.world, "fma": 0}

    def config(self):
        c = static_config(self.name, self.world)
        c["partition"] = f"{self.world} row slabs, two halo exchanges per step (2 affinity rows, 16 state rows)"
        return c

    def check_against_single_gpu(self, steps=2):
        import pyca_b200

        full = self._full_world()
        self.slab.set_interior(full[:, :, self.slab.r0:self.slab.r1])
        for _ in range(steps):
            self.slab.step()
        single = pyca_b200.MaCELeniaXChan((self.H, self.W), params=self.params, state_init=full, device=self.device)
        for _ in range(steps):
            single.step()
        err = (self.slab.interior() - single.state[:, :, self.slab.r0:self.slab.r1]).abs().max().item()
        mass = self.slab.interior().double().sum().reshape(1)
        torch.distributed.all_reduce(mass)
        self.single = single
        return {"steps": steps, "max_abs_err": err, "tolerance": 1e-5, "mass": float(mass.item()), "mass_initial": float(full.double().sum().item()),
                "timed_out": bool(self.slab.timed_out()) if hasattr(self.slab, "timed_out") else False}


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
        return static_config(self.name, self.world)

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
        p = cpu_params("params_xchan_arborical_asbestosis.npz")
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


    def e2e_setup(self):
        self.h_out = torch.empty(1, 3, self.H, self.W).pin_memory()

    def e2e_step(self):
        return self.e2e_step_sync()

    def e2e_drain(self):
        pass

    def cpu_setup(self):
        from oracle import pyca_oracle as orc

        self.orc = orc
        p = cpu_params("params_mace_anencephalic_brakeman.npz")
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

    def check_against_single_gpu(self, steps=2):
        """this rank's shard of worlds against the same worlds inside the UNSPLIT 256-world batch stepped on one GPU (bit-exact:
        the in-kernel update mask is keyed by the global cell index)"""
        import pyca_b200

        g = torch.Generator(device=self.device).manual_seed(99)
        full = torch.rand(self.B, 16, self.H, self.W, generator=g, device=self.device)
        lo = self.rank * self.Bl
        a = full[lo:lo + self.Bl].contiguous()
        for t_ in range(steps):
            a = pyca_b200.nca_step(a, self.wts, seed=1, step_index=t_, impl=self.impl, first_world=lo)
        f = full
        for t_ in range(steps):
            f = pyca_b200.nca_step(f, self.wts, seed=1, step_index=t_, impl=self.impl)
        return {"steps": steps, "bit_exact": bool(torch.equal(a, f[lo:lo + self.Bl]))}

    def launches_per_step(self):
        return 2

    def dominant(self):
        cells = self.Bl * self.H * self.W
        return {"kernel": "nca_update_tc_kernel" if self.impl == 1 else "nca_update_simt_kernel", "bytes": cells * 128,
                "fma": 0 if self.impl == 1 else cells * 6144, "flops_algorithmic": cells * 16384,
                "flops_issued_tf32": cells * 2 * 3 * (128 * 40 + 16 * 128)}

    def config(self):
        return static_config(self.name, self.world)

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
        fx = getattr(self, "fx", None) or __import__("tests.util", fromlist=["load_npz"]).load_npz("nca_betta.npz")
        self.cw = [t(fx[k]) for k in ("W1", "b1", "W2", "b2")]

    def cpu_step(self):
        mask = torch.rand(self.cB, 1, self.H, self.W, device=self.cstate.device) <= 0.5
        self.cstate = self.orc.nca_step(self.cstate, *self.cw, mask)

    def cpu_cells_per_step(self):
        return self.cB * self.H * self.W

    def cpu_sample(self):
        return f"NCA {self.cB}x16x{self.H}x{self.W} oracle step (conv2d + Linear path of nca.py:239-263)"


def make_workload(name, rank, world, device):
    if name == "lenia_1k":
        return LeniaBatchWL(rank, world, device, 1024, 1024, 1, name)
    if name == "lenia_1k_single":
        return LeniaWL(rank, world, device, 1024, 1024, 1, name)
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


def _one_block(wl, steps, flush):
    """exactly `steps` steps; returns the CUDA events whose intervals add up to the block's device time"""
    if flush is not None:
        evs = []
        for _ in range(steps):
            flush.zero_()   # (untimed) evicts the previous step's lines
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            wl.step()
            b.record()
            evs.append((a, b))
        return evs
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    if hasattr(wl, "run"):
        wl.run(steps)
    else:
        for _ in range(steps):
            wl.step()
    b.record()
    return [(a, b)]


def timed_blocks(wl, steps, warmup, world, flush, device, min_ms=50.0, max_blocks=2000):
    """W untimed steps, then R blocks of EXACTLY K steps, every block between its own CUDA events on the launching stream, the
    whole region bracketed by barrier + synchronize.  R is sized from one untimed block so that the region lasts >= min_ms
    whatever K is (round 1's 20 x 21 us = 0.4 ms window let one host hiccup decide the figure).  Per block the MAX over ranks
    is taken; the reported step time is the MEDIAN block / K.  Returns (ms_per_step, info)."""
    for _ in range(warmup):
        wl.step()
    barrier(world)
    evs = _one_block(wl, steps, flush)
    torch.cuda.synchronize()
    est = max(sum(a.elapsed_time(b) for a, b in evs), 1e-3)
    R = torch.tensor([int(min(max(-(-min_ms // est), 3), max_blocks))], dtype=torch.int64, device=device)
    if world > 1:
        torch.distributed.all_reduce(R, op=torch.distributed.ReduceOp.MAX)
    R = int(R.item())
    barrier(world)
    t0 = time.perf_counter()
    blocks = [_one_block(wl, steps, flush) for _ in range(R)]
    barrier(world)
    wall = time.perf_counter() - t0
    bt = torch.tensor([sum(a.elapsed_time(b) for a, b in blk) for blk in blocks], dtype=torch.float64, device=device)
    if world > 1:
        torch.distributed.all_reduce(bt, op=torch.distributed.ReduceOp.MAX)
    bt = bt.cpu()
    med = float(bt.median())
    info = {"blocks": R, "steps_per_block": steps, "block_ms_median": med, "block_ms_min": float(bt.min()), "block_ms_max": float(bt.max()),
            "device_ms_total": float(bt.sum()), "wall_ms_total": wall * 1e3,
            "statistic": "median over blocks of (max over ranks of the block's CUDA-event time) / K",
            "issue": wl.issue_note() if hasattr(wl, "issue_note") else "steps issued one by one from Python (ctypes call per step)"}
    return med / steps, info


def flush_for(wl, device):
    """None when the workload's inputs exceed L2 by construction (big worlds, rings of worlds), else the flush buffer."""
    big = wl.dominant()["bytes"] > L2_BYTES or getattr(wl, "rotating", False)
    return None if big else torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=device)


def cpu_baseline(name, budget_s=12.0):
    """The reference's own step (baseline/_ref; oracle port if absent) on all host cores, a bounded sample of the workload."""
    arm = RefArm(name)
    arm.setup()
    t0 = time.perf_counter()
    n0 = 0
    while time.perf_counter() - t0 < min(2.0, 0.25 * budget_s) or n0 < 1:   # spin-up: thread pools, first-touch pages
        arm.step()
        n0 += 1
    t0 = time.perf_counter()
    n = 0
    while True:
        arm.step()
        n += 1
        el = time.perf_counter() - t0
        if el > budget_s or (n >= 200 and el > 0.8 * budget_s):
            break
    gcups = arm.cells_per_step() * n / el / 1e9
    return {"value": gcups, "unit": "GCUPS", "cores": torch.get_num_threads(), "kind": arm.kind,
            "sample": f"{n} x {arm.sample()}, {el:.1f} s", "ms_per_step": el / n * 1e3}


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
    """--impl reference: the reference's own CPU implementation of the path (baseline/_ref), all host threads, rank 0 only.
    Imports neither the product package nor the CUDA library, and creates no CUDA context."""
    if rank != 0:
        return
    name = args.workload or "lenia_1k"
    arm = RefArm(name)
    arm.setup()
    # one "step" of this arm = a bounded sample of `reps` consecutive CPU steps (>= 50 ms of work), so that K steps are not
    # dominated by start-up effects; an 8 s untimed spin-up precedes the W warm-up steps (thread pools, first-touch pages, host
    # clock ramp: the step time of a fresh process keeps falling for several seconds on the box)
    t0 = time.perf_counter()
    n0 = 0
    while time.perf_counter() - t0 < 8.0 or n0 < 1:
        arm.step()
        n0 += 1
    t_one = (time.perf_counter() - t0) / n0
    reps = max(1, min(int(0.05 / t_one + 0.999), 1000))
    for _ in range(args.warmup):
        for _ in range(reps):
            arm.step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        for _ in range(reps):
            arm.step()
    el = time.perf_counter() - t0
    v = arm.cells_per_step() * args.steps * reps / el / 1e9
    scaling = "weak" if name in REPLICA_WORKLOADS or SHAPES[name][0] in ("xchan", "flow") else "strong"
    line = {"impl": "reference", "metric": "cell-updates/sec", "value": v, "unit": "GCUPS", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": el / (args.steps * reps) * 1e3, "higher_is_better": True, "scaling": scaling,
            "vs_baseline": None, "dtype": "f32" if SHAPES[name][0] != "gol" else "int32", "data": "synthetic",
            "config": static_config(name, args.gpus),
            "cpu_baseline": {"value": v, "unit": "GCUPS", "cores": torch.get_num_threads(), "kind": arm.kind,
                             "sample": f"{args.steps} steps x {reps} consecutive CPU updates of {arm.sample()}"},
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
    from pyca_b200 import _lib as b200lib

    wl = make_workload(name, rank, world, device)
    wl.steps_hint = args.steps
    wl.setup()
    dom = wl.dominant()
    flush = flush_for(wl, device)
    b200lib.variants(reset=True)
    wl.step()   # (untimed) records which kernel instantiations one step of this workload launches
    kernels = b200lib.variants(reset=True)

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    ms_per_step, timing = timed_blocks(wl, args.steps, args.warmup, world, flush, device, min_ms=250.0)   # (long enough for >= 3 clock samples)
    clocks = sampler.stop() if sampler else None
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

        def run_e2e(stepfn, n, min_s=0.15, max_blocks=50):
            """blocks of n end-to-end steps (queue n worlds, drain), wall clock per block, max over ranks, MEDIAN block: the
            10 ms single window of round 1 moved by +-30 % from run to run"""
            for _ in range(3):
                stepfn()
            wl.e2e_drain()

            def one_block():
                barrier(world)
                t0 = time.perf_counter()
                for _ in range(n):
                    hb_, db_ = stepfn()
                wl.e2e_drain()
                torch.cuda.synchronize()
                el_ = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=device)
                if world > 1:
                    torch.distributed.all_reduce(el_, op=torch.distributed.ReduceOp.MAX)
                return float(el_.item()), hb_, db_

            first, hb, db = one_block()
            nb = torch.tensor([int(min(max(-(-min_s // max(first, 1e-6)), 5), max_blocks))], device=device)
            if world > 1:
                torch.distributed.all_reduce(nb, op=torch.distributed.ReduceOp.MAX)
            times = sorted(one_block()[0] for _ in range(int(nb.item())))
            return times[len(times) // 2], hb, db

        n_e2e = max(6, min(4 * args.steps, 200))
        el, hb, db = run_e2e(wl.e2e_step, n_e2e)
        cells = wl.e2e_cells() if hasattr(wl, "e2e_cells") else wl.cells_per_step()
        piped = hasattr(wl, "pipe")
        e2e = {"value": cells * n_e2e / el / 1e9, "unit": "GCUPS", "h2d_bytes_per_step": hb, "d2h_bytes_per_step": db,
               "ms_per_step": el / n_e2e * 1e3, "steps": n_e2e, "statistic": "median over >= 5 blocks of `steps` worlds (>= 0.15 s in all)",
               "api": (f"pyca_b200.LeniaHostPipe(auto, depth={E2E_DEPTH}).submit(pinned h_in, pinned h_out) = b200ca_lenia_pipe_submit: per "
                       "world H2D (copy engine) -> step -> the unframing kernel stores the result straight into the pinned host buffer over "
                       "PCIe; independent worlds overlapped on separate streams") if piped else
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
        for other in ["lenia_1k_single", "lenia3_1k", "gol_250", "gol_64k", "mace_xchan_4k", "nca_256", "flow_lenia_4k"]:
            if other == name:
                continue
            try:
                ow = make_workload(other, 0, 1, device)
                k = 4 if other in ("mace_xchan_4k", "flow_lenia_4k") else 20
                ow.steps_hint = k
                ow.setup()
                od = ow.dominant()
                b200lib.variants(reset=True)
                ow.step()
                okern = b200lib.variants(reset=True)
                ms, otiming = timed_blocks(ow, k, 3, 1, flush_for(ow, device), device, min_ms=30.0)
                ent = {"gcups": ow.cells_per_step() / (ms * 1e-3) / 1e9, "ms_per_step": ms,
                       "hbm_frac_algorithmic": od["bytes"] / (ms * 1e-3) / 1e9 / peaks["hbm_gbs"], "config": ow.config(),
                       "kernels": okern, "blocks": otiming["blocks"]}
                if not args.no_cpu:
                    cb = cpu_baseline(other, budget_s=3.0)
                    ent["cpu_gcups"] = cb["value"]
                    ent["cpu_sample"] = cb["sample"]
                    ent["cpu_kind"] = cb["kind"]
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
        # The sharded BASELINE configs at N GPUs: configs[4] (one big torus in row slabs, halo rows over NVLink), configs[3] (NCA
        # batch shards, no communication) and the MaCE-XChan world of configs[2] in row slabs.  Each entry carries the N-GPU time,
        # the same workload on ONE GPU of this box (every rank runs it alone; max over ranks) and a correctness check of the
        # N-GPU result against the single-GPU result made inside this run.
        sharded = {}

        def reduce_max(x):
            t_ = torch.tensor([float(x)], dtype=torch.float64, device=device)
            torch.distributed.all_reduce(t_, op=torch.distributed.ReduceOp.MAX)
            return float(t_.item())

        def reduce_all(flag):
            t_ = torch.tensor([1 if flag else 0], device=device)
            torch.distributed.all_reduce(t_, op=torch.distributed.ReduceOp.MIN)
            return bool(t_.item())

        for big in ["lenia_16k", "gol_64k", "nca_256", "mace_xchan_4k"]:
            try:
                ent = {}
                sw = MaceSlabWL(rank, world, device, 4096, 4096) if big == "mace_xchan_4k" else make_workload(big, rank, world, device)
                sw.setup()
                k = 30 if big != "mace_xchan_4k" else 8
                ms_n, tim_n = timed_blocks(sw, k, 3, world, flush_for(sw, device), device, min_ms=40.0)
                ent["ms_per_step"] = ms_n
                ent["blocks"] = tim_n["blocks"]
                ent["gcups"] = sw.cells_per_step() / (ms_n * 1e-3) / 1e9
                ent["config"] = sw.config()
                if hasattr(sw, "halo"):
                    ent["halo"] = ("P2P stores + arrival counters over NVLink (symmetric memory), no host-side exchange" if sw.halo == "p2p"
                                   else "NCCL send/recv on a side stream")
                b200lib.variants(reset=True)
                sw.step()
                ent["kernels"] = b200lib.variants(reset=True)
                chk = sw.check_against_single_gpu()
                if "max_abs_err" in chk:
                    chk["max_abs_err"] = reduce_max(chk["max_abs_err"])
                    chk["ok"] = reduce_all(chk["max_abs_err"] <= chk["tolerance"] and not chk.get("timed_out", False))
                else:
                    chk["ok"] = reduce_all(chk["bit_exact"] and chk.get("population", 0) == chk.get("population_single_gpu", 0))
                    chk["bit_exact"] = chk["ok"]
                ent["check_vs_single_gpu"] = chk
                if hasattr(getattr(sw, "slab", None), "close"):
                    sw.slab.close()
                del sw
                torch.cuda.empty_cache()
                s1 = make_workload(big, 0, 1, device)
                s1.setup()
                ms1, _ = timed_blocks(s1, k, 3, 1, flush_for(s1, device), device, min_ms=40.0)
                ms1 = reduce_max(ms1)
                ent["single_gpu_ms_per_step"] = ms1
                ent["single_gpu_gcups"] = s1.cells_per_step() / (ms1 * 1e-3) / 1e9
                del s1
                torch.cuda.empty_cache()
                torch.distributed.barrier()
                sharded[big] = ent
            except Exception as e:
                import traceback

                log(traceback.format_exc())
                sharded[big] = {"error": repr(e)[:300]}

    cpu = None
    torch_gpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_baseline(name)
        try:
            torch_gpu = torch_gpu_baseline(wl, device)
        except Exception as e:
            torch_gpu = {"value": None, "error": repr(e)[:160]}

    if rank == 0:
        line = {"metric": "cell-updates/sec", "value": value, "unit": "GCUPS", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
                "scaling": wl.scaling, "vs_baseline": None, "dtype": wl.dtype, "data": "synthetic", "config": wl.config(),
                "clocks": clocks, "e2e": e2e, "gpu_launches": wl.launches_per_step() * args.steps * timing["blocks"],
                "gpu_launches_per_step": wl.launches_per_step(), "kernels": kernels, "timing": timing, "roofline": roofline,
                "cpu_baseline": cpu, "torch_gpu_baseline": torch_gpu, "per_model": extras, "sharded": sharded}
        emit(line)
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
