"""Row-slab sharding of one large torus world across the GPUs of a box (one process per GPU).

The reference has no distributed code at all: it scales only by allocating a
bigger tensor on one device, and Lenia's global FFT (lenia.py:460-467) cannot be partitioned.  The
CUDA kernels here are local operators, so the world splits into row slabs:

    rank r owns global rows [r*H/N, (r+1)*H/N);  rank r-1 is "above", rank r+1 "below" (ring = torus).

* Game of Life: the slab carries `g` ghost rows on each side.  `g` generations are computed locally
  with rows outside the buffer treated as dead (`vert_open`); the error creeps one row per generation
  from the buffer edge, so after `g` generations the owned rows are still exact.  Then ONE exchange
  refreshes the ghost rows.  (Per-step latency of a send/recv pair would dominate a ~25 us step;
  trading 2g redundant rows for 1/g as many exchanges hides it.)
* Lenia: the framed layout already has 16 ghost rows; each step computes the slab-boundary tile rows
  first, then exchanges the fresh boundary rows on a side stream while the interior tiles run.

The exchange itself is `torch.distributed` point-to-point (NCCL over NVLink on GPUs, gloo in the CPU
tests) on contiguous row blocks; the functions below are device-agnostic so the N>1 logic is covered
on CPU with world_size 2.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


def slab_rows(H: int, world: int, rank: int) -> Tuple[int, int]:
    """Global row range [r0, r1) owned by `rank` (H must divide evenly: slabs are equal so ranks stay in step)."""
    if H % world != 0:
        raise ValueError(f"world height {H} is not divisible by the number of ranks {world}")
    hl = H // world
    return rank * hl, (rank + 1) * hl


def ring_neighbours(rank: int, world: int) -> Tuple[int, int]:
    """(rank above, rank below) on the periodic ring."""
    return (rank - 1) % world, (rank + 1) % world


def exchange_ghost_rows(blocks_up_send: List[torch.Tensor], blocks_up_recv: List[torch.Tensor],
                        blocks_down_send: List[torch.Tensor], blocks_down_recv: List[torch.Tensor],
                        rank: int, world: int, group=None) -> None:
    """One halo exchange.

    blocks_up_send:   my first owned rows   -> become the BOTTOM ghost rows of the rank above
    blocks_down_send: my last owned rows    -> become the TOP ghost rows of the rank below
    blocks_up_recv:   my TOP ghost rows     <- last owned rows of the rank above
    blocks_down_recv: my BOTTOM ghost rows  <- first owned rows of the rank below
    Every block must be a contiguous tensor (a row block of one plane).  world == 1: local copies.
    """
    if world == 1:
        for s, r in zip(blocks_down_send, blocks_up_recv):
            r.copy_(s)
        for s, r in zip(blocks_up_send, blocks_down_recv):
            r.copy_(s)
        return
    up, down = ring_neighbours(rank, world)
    # Post order matters when up == down (world == 2): messages between one pair of ranks match in posting
    # order, so send [last rows, first rows] against the peer's receives [top ghost, bottom ghost].
    ops = []
    for s in blocks_down_send:
        assert s.is_contiguous()
        ops.append(dist.P2POp(dist.isend, s, down, group))
    for s in blocks_up_send:
        assert s.is_contiguous()
        ops.append(dist.P2POp(dist.isend, s, up, group))
    for r in blocks_up_recv:
        assert r.is_contiguous()
        ops.append(dist.P2POp(dist.irecv, r, up, group))
    for r in blocks_down_recv:
        assert r.is_contiguous()
        ops.append(dist.P2POp(dist.irecv, r, down, group))
    for w in dist.batch_isend_irecv(ops):
        w.wait()


def build_exchange_ops(blocks_up_send, blocks_up_recv, blocks_down_send, blocks_down_recv, rank: int, world: int, group=None):
    """The P2P operation list of one halo exchange (same posting order as exchange_ghost_rows), built once per buffer
    and re-posted every step: the slab step is short enough at 8 GPUs for per-step Python object churn to show."""
    up, down = ring_neighbours(rank, world)
    ops = [dist.P2POp(dist.isend, s, down, group) for s in blocks_down_send]
    ops += [dist.P2POp(dist.isend, s, up, group) for s in blocks_up_send]
    ops += [dist.P2POp(dist.irecv, r, up, group) for r in blocks_up_recv]
    ops += [dist.P2POp(dist.irecv, r, down, group) for r in blocks_down_recv]
    return ops


class GolSlab:
    """Bit-packed Game-of-Life row slab with `ghost` ghost rows per side (device-agnostic bookkeeping).

    `step_fn(src, dst, rows, nsteps)` advances the whole local buffer (owned + ghost rows, rows outside
    dead) by nsteps generations ping-ponging src->dst and returns the tensor holding the result.
    """

    def __init__(self, H: int, W: int, rank: int, world: int, ghost: int, device, stride_words: Optional[int] = None):
        self.H, self.W, self.rank, self.world = H, W, rank, world
        self.r0, self.r1 = slab_rows(H, world, rank)
        self.hl = self.r1 - self.r0
        if ghost < 1 or ghost > self.hl:
            raise ValueError(f"ghost rows {ghost} must be in [1, {self.hl}]")
        self.g = ghost
        self.rows = self.hl + 2 * ghost
        wpr = (W + 31) // 32
        self.stride = stride_words or (wpr + 3) // 4 * 4
        self.bufs = [torch.zeros((self.rows, self.stride), dtype=torch.int32, device=device) for _ in range(2)]
        self.cur = 0
        self.fresh = 0  # generations the ghost rows are still good for

    @property
    def buf(self):
        return self.bufs[self.cur]

    def owned(self) -> torch.Tensor:
        return self.buf[self.g:self.g + self.hl]

    def exchange(self, group=None):
        b, g, hl = self.buf, self.g, self.hl
        exchange_ghost_rows([b[g:2 * g]], [b[0:g]], [b[hl:hl + g]], [b[hl + g:hl + 2 * g]], self.rank, self.world, group)
        self.fresh = g

    def advance(self, step_fn, nsteps: int, group=None):
        """nsteps generations, exchanging ghost rows whenever they run out."""
        done = 0
        while done < nsteps:
            if self.fresh == 0:
                self.exchange(group)
            n = nsteps - done
            if n > self.fresh:
                n = self.fresh
            out = step_fn(self.bufs[self.cur], self.bufs[self.cur ^ 1], self.rows, n)
            if out is not self.bufs[self.cur]:
                self.cur ^= 1
            self.fresh -= n
            done += n


class LeniaSlab:
    """Framed Lenia row slab: (B,C,hl+2F,pitch) fp32 ping-pong buffers whose top/bottom frame rows are
    ghost rows owned by the ring neighbours."""

    def __init__(self, B: int, C: int, H: int, W: int, rank: int, world: int, frame: int, pitch: int, device):
        self.B, self.C, self.H, self.W, self.rank, self.world, self.F = B, C, H, W, rank, world, frame
        self.r0, self.r1 = slab_rows(H, world, rank)
        self.hl = self.r1 - self.r0
        if self.hl < frame:
            raise ValueError(f"slab height {self.hl} < frame {frame}")
        self.pitch = pitch
        self.rows = self.hl + 2 * frame
        self.bufs = [torch.zeros((B, C, self.rows, pitch), dtype=torch.float32, device=device) for _ in range(2)]
        self.cur = 0

    @property
    def buf(self):
        return self.bufs[self.cur]

    @property
    def other(self):
        return self.bufs[self.cur ^ 1]

    def interior(self, t=None) -> torch.Tensor:
        t = self.buf if t is None else t
        if getattr(self, "_pending_exchange", False) and getattr(self, "comm_stream", None) is not None:
            torch.cuda.current_stream(self.device).wait_stream(self.comm_stream)  # readers see a settled buffer
        return t[:, :, self.F:self.F + self.hl, self.F:self.F + self.W]

    def ghost_blocks(self, t: torch.Tensor):
        """(up_send, up_recv, down_send, down_recv) lists of contiguous row blocks (full framed rows, so the
        x-frame travels with them)."""
        F, hl = self.F, self.hl
        us, ur, ds, dr = [], [], [], []
        for b in range(self.B):
            for c in range(self.C):
                p = t[b, c]
                us.append(p[F:2 * F])
                ur.append(p[0:F])
                ds.append(p[hl:hl + F])
                dr.append(p[hl + F:hl + 2 * F])
        return us, ur, ds, dr

    def exchange(self, t=None, group=None):
        t = self.buf if t is None else t
        if self.world == 1:
            us, ur, ds, dr = self.ghost_blocks(t)
            exchange_ghost_rows(us, ur, ds, dr, self.rank, self.world, group)
            return
        cache = self.__dict__.setdefault("_exchange_ops", {})
        key = (t.data_ptr(), id(group))
        ops = cache.get(key)
        if ops is None:
            ops = cache[key] = build_exchange_ops(*self.ghost_blocks(t), self.rank, self.world, group)
        for w in dist.batch_isend_irecv(ops):
            w.wait()


# ------------------------------------------------------------------------------------------------
# CUDA drivers of the slabs (one process per GPU; torch.distributed NCCL for the halo exchange)
# ------------------------------------------------------------------------------------------------
class GolSlabCUDA(GolSlab):
    """GolSlab whose local generations run through b200ca_gol_run(vert_open=1)."""

    def __init__(self, H, W, rank, world, ghost=16, device="cuda", s_mask=0b1100, b_mask=0b1000):
        from . import _lib

        super().__init__(H, W, rank, world, ghost, device)
        self.s_mask, self.b_mask = s_mask, b_mask
        self.device = torch.device(device)
        self._run = _lib.lib().b200ca_gol_run
        self._ptrs = [b.data_ptr() for b in self.bufs]

    def _step_fn(self, src, dst, rows, n):
        # hot: a 65536^2 slab generation takes ~20 us on 8 GPUs, so the host side of a launch has to stay well below that
        from . import _lib

        with _lib.device_guard(self.device):
            rc = self._run(self._ptrs[self.cur], self._ptrs[self.cur ^ 1], rows, self.W, self.stride, self.s_mask, self.b_mask, 1, n,
                           torch.cuda.current_stream(self.device).cuda_stream)
        if rc:
            _lib.check(rc, "b200ca_gol_run")
        return dst if (n & 1) else src

    def random_fill(self, seed: int):
        """Fills the OWNED rows with the rows [r0, r1) of the synthetic world `seed` (same bits for any N)."""
        from . import _lib

        own = self.owned()
        with torch.cuda.device(self.device):
            _lib.call("b200ca_gol_random_fill", _lib.ptr(own), self.hl, self.W, self.stride, seed, self.r0,
                      _lib.current_stream(self.device))
        self.fresh = 0

    def run(self, nsteps: int, group=None):
        self.advance(self._step_fn, nsteps, group)

    def population(self) -> int:
        from . import _lib

        cnt = torch.zeros(1, dtype=torch.int64, device=self.device)
        own = self.owned()
        with torch.cuda.device(self.device):
            _lib.call("b200ca_gol_population", _lib.ptr(own), self.hl, self.W, self.stride, _lib.ptr(cnt),
                      _lib.current_stream(self.device))
        return int(cnt.item())


class GolSlabP2P(GolSlabCUDA):
    """GolSlabCUDA whose ghost-row exchange is ONE kernel per rank (csrc/halo.cu): the boundary rows are stored straight into the
    neighbours' ghost rows over NVLink and the ranks synchronise through counters in symmetric memory -- no NCCL send/recv, no
    host synchronisation between the generations of consecutive batches."""

    def __init__(self, H, W, rank, world, ghost=16, device="cuda", s_mask=0b1100, b_mask=0b1000, group=None):
        import torch.distributed._symmetric_memory as symm

        super().__init__(H, W, rank, world, ghost, "meta", s_mask, b_mask)
        self.device = torch.device(device)
        self.group = group if group is not None else dist.group.WORLD
        n = self.rows * self.stride
        self._n = n
        self._sym = symm.empty(2 * n + 64, dtype=torch.int32, device=self.device)
        self._sym.zero_()
        self._hdl = symm.rendezvous(self._sym, self.group)
        self.bufs = [self._sym[i * n:(i + 1) * n].view(self.rows, self.stride) for i in range(2)]
        self._ptrs = [b.data_ptr() for b in self.bufs]
        self._counters = self._sym[2 * n:2 * n + 64]   # [0:2] ready, [2:4] arrived, [4] ticket, [5] error
        base = list(self._hdl.buffer_ptrs)
        self._base = base
        self._up, self._down = ring_neighbours(rank, world)
        self._cnt = 2 * n * 4
        self.epoch = 0

    def exchange(self, group=None):
        from . import _lib

        g, hl, row = self.g, self.hl, self.stride * 4
        me, up, dn = self._base[self.rank], self._base[self._up], self._base[self._down]
        buf = self.cur * self._n * 4
        self.epoch += 1
        with _lib.device_guard(self.device):
            _lib.call("b200ca_halo_push", me + buf + g * row, up + buf + (hl + g) * row, me + buf + hl * row, dn + buf, g * row,
                      up + self._cnt + 4, dn + self._cnt, me + self._cnt, up + self._cnt + 12, dn + self._cnt + 8, me + self._cnt + 8,
                      self.epoch & 0xFFFFFFFF, me + self._cnt + 16, me + self._cnt + 20, _lib.current_stream(self.device))
        self.fresh = g

    def timed_out(self) -> bool:
        return bool(self._counters[5].item())

    def close(self):
        if getattr(self, "_sym", None) is not None:
            torch.cuda.synchronize(self.device)
            dist.barrier(self.group)
            self.bufs = []
            self._counters = self._hdl = self._sym = None


class LeniaSlabCUDA(LeniaSlab):
    """Lenia row slab on one GPU: boundary tile rows first, halo exchange on a side stream overlapped with
    the interior tiles (Lenia.step semantics, lenia.py:430-451, on rows [r0, r1) of the global torus)."""

    def __init__(self, lenia_params, C, H, W, rank, world, device="cuda", dt=0.1, overlap=True, conv_path="auto"):
        from . import _lib
        from .lenia_params import compute_kernel

        L = _lib.lib()
        B = lenia_params["weights"].shape[0]
        super().__init__(B, C, H, W, rank, world, _lib.FRAME, L.b200ca_frame_pitch(W), device)
        self.device = torch.device(device)
        self.dt = dt
        p = lenia_params.to(device).param_dict
        self.k_mult = p.get("k_mult", 1)
        self.mu = p["mu"].float().contiguous()
        self.sigma = p["sigma"].float().contiguous()
        self.weights = p["weights"].float().contiguous()
        self.k = compute_kernel(lenia_params, C).float().contiguous()
        self._kprep = torch.empty(L.b200ca_lenia_kprep_elems(B, C, self.k_mult), dtype=torch.float32, device=device)
        with torch.cuda.device(self.device):
            _lib.call("b200ca_lenia_prepare_kernel", _lib.ptr(self.k), _lib.ptr(self._kprep), B, C, self.k_mult, int(p["k_size"]),
                      _lib.current_stream(self.device))
        # FFT path (row-FFT x column-FIR hybrid): the column direction is local, so a slab only needs its 15 ghost rows
        fft_ok = L.b200ca_lenia_fft_length(self.hl, W) > 0
        if conv_path == "fft" and not fft_ok:
            raise _lib.B200CAError(f"conv_path='fft' unsupported for a {self.hl}x{W} slab")
        self.use_fft = fft_ok and conv_path in ("fft", "auto")
        if self.use_fft:
            self._kfft = torch.empty(L.b200ca_lenia_fft_kernel_elems(B, C, self.k_mult, self.hl, W), dtype=torch.float32, device=device)
            with torch.cuda.device(self.device):
                _lib.call("b200ca_lenia_fft_prepare_kernel", _lib.ptr(self.k), _lib.ptr(self._kfft), B, C, self.k_mult, int(p["k_size"]),
                          self.hl, W, _lib.current_stream(self.device))
            self._fft_ws = torch.empty(L.b200ca_lenia_fft_workspace_elems(B, C, self.k_mult, self.hl, W), dtype=torch.float32,
                                       device=device)
        self.tile_rows = L.b200ca_lenia_tile_rows(self.hl, W, B, C)
        self.n_tile_rows = (self.hl + self.tile_rows - 1) // self.tile_rows
        # tile rows that cover the F owned rows next to each edge.  The LAST tile row is clipped to hl - (T-1)*TH rows
        # (lenia.cu clips the tile at row hl), so the bottom count comes from real row coverage: the smallest nb with
        # hl - (T - nb) * TH >= F  (a remainder tile of fewer than F rows needs one more tile row below it)
        self.n_boundary = (self.F + self.tile_rows - 1) // self.tile_rows
        T, TH = self.n_tile_rows, self.tile_rows
        self.n_boundary_bot = next((nb for nb in range(1, T + 1) if self.hl - (T - nb) * TH >= self.F), T)
        if self.use_fft:
            self.overlap = overlap and world > 1
        else:
            self.overlap = overlap and world > 1 and self.n_tile_rows > self.n_boundary + self.n_boundary_bot
        self.comm_stream = torch.cuda.Stream(device=self.device) if world > 1 else None
        self._pending_exchange = False  # FFT path: the ghost rows of `buf` are still in flight on comm_stream
        # FFT path: row FFTs (twice when the halo exchange is overlapped: interior rows, then the rows touching ghost rows)
        # + column FIR and inverse FFT, fused into one kernel for N = 1024 transforms
        nfft = 2 if self.use_fft and _lib.lib().b200ca_lenia_fft_length(self.hl, self.W) == 1024 else 3
        self.launches_per_step = (nfft + (1 if self.overlap else 0)) if self.use_fft else (3 if self.overlap else 1)

    def set_interior(self, dense: torch.Tensor, group=None):
        """dense: (B,C,hl,W) rows [r0,r1) of the global state."""
        from . import _lib

        self.interior().copy_(dense)
        with torch.cuda.device(self.device):
            _lib.call("b200ca_frame_refresh", _lib.ptr(self.buf), self.B, self.C, self.hl, self.W, 1 if self.world == 1 else 0,
                      _lib.current_stream(self.device))
        if self.world > 1:
            self.exchange(self.buf, group)

    def _launch(self, t0, t1):
        from . import _lib

        if self.use_fft:
            # (t0 is the phase here: 0 = whole step, 1 = row FFTs of owned rows, 2 = boundary row FFTs + FIR + inverse)
            _lib.call("b200ca_lenia_step_fft_phase", _lib.ptr(self.buf), _lib.ptr(self.other), _lib.ptr(self._kfft), _lib.ptr(self.mu),
                      _lib.ptr(self.sigma), _lib.ptr(self.weights), _lib.ptr(self._fft_ws), self.B, self.C, self.k_mult, self.hl,
                      self.W, float(self.dt), 0, 1 if self.world == 1 else 0, t0, _lib.current_stream(self.device))
            return
        _lib.call("b200ca_lenia_step", _lib.ptr(self.buf), _lib.ptr(self.other), _lib.ptr(self._kprep), _lib.ptr(self.mu),
                  _lib.ptr(self.sigma), _lib.ptr(self.weights), self.B, self.C, self.k_mult, self.hl, self.W, float(self.dt), 0,
                  1 if self.world == 1 else 0, t0, t1, _lib.current_stream(self.device))

    def step(self, group=None):
        from . import _lib

        with _lib.device_guard(self.device):
            if self.world == 1:
                self._launch(0, 0)
            elif self.use_fft and self.overlap:
                # the row FFTs of the owned rows run while the previous step's ghost rows are still travelling
                main = torch.cuda.current_stream(self.device)
                self._launch(1, 0)
                if self._pending_exchange:
                    main.wait_stream(self.comm_stream)
                self._launch(2, 0)
                self.comm_stream.wait_stream(main)
                with torch.cuda.stream(self.comm_stream):
                    self.exchange(self.other, group)
                self._pending_exchange = True
            elif not self.overlap:
                self._launch(0, 0)
                self.exchange(self.other, group)
            else:
                nb, nbb, T = self.n_boundary, self.n_boundary_bot, self.n_tile_rows
                main = torch.cuda.current_stream(self.device)
                self._launch(0, nb)
                self._launch(T - nbb, T)
                self.comm_stream.wait_stream(main)
                with torch.cuda.stream(self.comm_stream):
                    self.exchange(self.other, group)
                self._launch(nb, T - nbb)
                main.wait_stream(self.comm_stream)
        self.cur ^= 1


class LeniaSlabP2P(LeniaSlab):
    """Lenia row slab whose halo exchange happens INSIDE the step kernel (option B): the marching kernel
    (csrc/lenia_march.cu) stores the slab's first / last 16 rows straight into the neighbours' ghost rows over NVLink and bumps
    their arrival counters; the next step's boundary CTAs wait on this rank's counters.  The host issues ONE launch per step and
    source channel -- no send/recv, no side stream, no stream hand-off.  Buffers and counters live in symmetric memory
    (torch.distributed._symmetric_memory: every rank maps every peer's allocation); boundary bands are scheduled first, so the
    halo travels while the interior bands compute (Lenia.step semantics, lenia.py:430-451, on rows [r0, r1) of the torus).

    Needs P2P access between the GPUs (NVLink / NVSwitch) and a process group; `LeniaSlabCUDA` (NCCL send/recv) is the
    fallback.  Every rank must call `step()` the same number of times."""

    def __init__(self, lenia_params, C, H, W, rank, world, device="cuda", dt=0.1, group=None):
        import torch.distributed._symmetric_memory as symm

        from . import _lib
        from .lenia_params import compute_kernel

        L = _lib.lib()
        B = lenia_params["weights"].shape[0]
        self.device = torch.device(device)
        self.group = group if group is not None else dist.group.WORLD
        # the base class allocates two private framed buffers: replaced below by views of one symmetric allocation
        super().__init__(B, C, H, W, rank, world, _lib.FRAME, L.b200ca_frame_pitch(W), "meta")
        if not L.b200ca_lenia_march_supported(self.hl, W):
            raise _lib.B200CAError(f"a {self.hl}x{W} slab is not supported by the marching kernel")
        n = B * C * self.rows * self.pitch
        self._n = n
        self._sym = symm.empty(2 * n + 64, dtype=torch.float32, device=self.device)
        self._sym.zero_()
        self._hdl = symm.rendezvous(self._sym, self.group)
        self.bufs = [self._sym[i * n:(i + 1) * n].view(B, C, self.rows, self.pitch) for i in range(2)]
        self._counters = self._sym[2 * n:2 * n + 64].view(torch.int32)   # [0] top-ghost arrivals, [1] bottom-ghost arrivals, [2] error
        ptrs = list(self._hdl.buffer_ptrs)
        up, down = ring_neighbours(rank, world)
        cnt = 2 * n * 4
        self._peer_out = [(ptrs[up] + i * n * 4, ptrs[down] + i * n * 4) for i in range(2)]
        self._sig_up = ptrs[up] + cnt + 4       # the up neighbour's BOTTOM ghost rows are my first rows
        self._sig_dn = ptrs[down] + cnt         # the down neighbour's TOP ghost rows are my last rows
        self._wait_top, self._wait_bot, self._err = ptrs[rank] + cnt, ptrs[rank] + cnt + 4, ptrs[rank] + cnt + 8
        self.dt = dt
        p = lenia_params.to(device).param_dict
        self.k_mult = p.get("k_mult", 1)
        self.mu = p["mu"].float().contiguous()
        self.sigma = p["sigma"].float().contiguous()
        self.weights = p["weights"].float().contiguous()
        self.k = compute_kernel(lenia_params, C).float().contiguous()
        self._kmarch = torch.empty(L.b200ca_lenia_march_kernel_elems(B, C, self.k_mult), dtype=torch.float32, device=device)
        with torch.cuda.device(self.device):
            _lib.call("b200ca_lenia_march_prepare_kernel", _lib.ptr(self.k), _lib.ptr(self._kmarch), B, C, self.k_mult, int(p["k_size"]),
                      _lib.current_stream(self.device))
        self._ctas = L.b200ca_lenia_march_slab_ctas(self.hl, W, B, C)
        self.step_index = 0
        self.use_fft = True
        self.overlap = True
        self.launches_per_step = C * self.k_mult
        self._step_fn = L.b200ca_lenia_step_march_slab

    def set_interior(self, dense: torch.Tensor, group=None):
        """dense: (B,C,hl,W) rows [r0,r1) of the global state.  Collective: every rank calls it (the first ghost rows travel by
        send/recv, the arrival counters restart from zero)."""
        from . import _lib

        torch.cuda.synchronize(self.device)
        dist.barrier(self.group)                      # nobody is still stepping into this rank's buffers
        self.interior().copy_(dense)
        with torch.cuda.device(self.device):
            _lib.call("b200ca_frame_refresh", _lib.ptr(self.buf), self.B, self.C, self.hl, self.W, 0, _lib.current_stream(self.device))
        self._counters.zero_()
        self.step_index = 0
        self.exchange(self.buf, self.group)
        torch.cuda.synchronize(self.device)
        dist.barrier(self.group)

    def step(self, group=None):
        from . import _lib

        nxt = self.cur ^ 1
        pu, pd = self._peer_out[nxt]
        with _lib.device_guard(self.device):
            rc = self._step_fn(self.bufs[self.cur].data_ptr(), self.bufs[nxt].data_ptr(), self._kmarch.data_ptr(), self.mu.data_ptr(),
                               self.sigma.data_ptr(), self.weights.data_ptr(), self.B, self.C, self.k_mult, self.hl, self.W,
                               float(self.dt), 0, pu, pd, self._sig_up, self._sig_dn, self._wait_top, self._wait_bot,
                               (self.step_index * self._ctas) & 0xFFFFFFFF, self._err, torch.cuda.current_stream(self.device).cuda_stream)
        if rc:
            _lib.check(rc, "b200ca_lenia_step_march_slab")
        self.cur = nxt
        self.step_index += 1

    def timed_out(self) -> bool:
        """True if a boundary CTA gave up waiting for a neighbour (the results are then invalid)."""
        return bool(self._counters[2].item())

    def close(self):
        """Collective: no rank may free its buffers while a neighbour's kernel can still store into them."""
        if getattr(self, "_sym", None) is not None:
            torch.cuda.synchronize(self.device)
            dist.barrier(self.group)
            self.bufs = []
            self._counters = self._hdl = self._sym = None


class MaceSlabCUDA(LeniaSlabCUDA):
    """MaCE-Lenia (and the cross-channel variant) on a row slab.  The redistribution reads the affinity on a
    2-row halo and the state on a 1-row halo, so a step has TWO exchanges: affinity rows after the convolution, state rows
    after the redistribution (the 16-row frame carries both; the alternative -- a 17-row state halo and redundant boundary
    affinity -- does not fit the frame).  macelenia.py:72-120, mace_lenia_xchan.py:59-76 on rows [r0, r1) of the torus."""

    def __init__(self, lenia_params, C, H, W, rank, world, device="cuda", beta=8.0, alpha=None, conv_path="auto"):
        super().__init__(lenia_params, C, H, W, rank, world, device=device, dt=0.1, overlap=False, conv_path=conv_path)
        self.beta = float(beta)
        self.alpha = alpha  # None: MaCELenia; a float: MaCELeniaXChan mixing rate
        self.aff = torch.zeros_like(self.bufs[0])
        self.launches_per_step = (2 if self.use_fft else 1) + 1

    def affinity_phase(self):
        """Affinity of the owned rows into the framed `aff` tensor (its top / bottom frame rows belong to the neighbours)."""
        from . import _lib

        rw = 1 if self.world == 1 else 0
        with _lib.device_guard(self.device):
            st = _lib.current_stream(self.device)
            if self.use_fft:
                _lib.call("b200ca_lenia_step_fft", _lib.ptr(self.buf), _lib.ptr(self.aff), _lib.ptr(self._kfft), _lib.ptr(self.mu),
                          _lib.ptr(self.sigma), _lib.ptr(self.weights), _lib.ptr(self._fft_ws), self.B, self.C, self.k_mult, self.hl,
                          self.W, 0.1, _lib.MODE_AFF, rw, st)
            else:
                _lib.call("b200ca_lenia_step", _lib.ptr(self.buf), _lib.ptr(self.aff), _lib.ptr(self._kprep), _lib.ptr(self.mu),
                          _lib.ptr(self.sigma), _lib.ptr(self.weights), self.B, self.C, self.k_mult, self.hl, self.W, 0.1,
                          _lib.MODE_AFF, rw, 0, 0, st)

    def redistribute_phase(self):
        """New state of the owned rows into `other` (needs the neighbours' affinity rows in the frame of `aff`)."""
        from . import _lib

        with _lib.device_guard(self.device):
            _lib.call("b200ca_mace_redistribute", _lib.ptr(self.buf), _lib.ptr(self.aff), _lib.ptr(self.other), self.B, self.C, self.hl,
                      self.W, self.beta, 0 if self.alpha is None else 1, 0.0 if self.alpha is None else float(self.alpha),
                      1 if self.world == 1 else 0, _lib.current_stream(self.device))
        self.cur ^= 1

    def step(self, group=None):
        self.affinity_phase()
        if self.world > 1:
            self.exchange(self.aff, group)
        self.redistribute_phase()
        if self.world > 1:
            self.exchange(self.buf, group)


class MaceSlabP2P(MaceSlabCUDA):
    """MaceSlabCUDA whose two exchanges per step (affinity rows, state rows) are ONE kernel per rank each (csrc/halo.cu,
    `b200ca_halo_push_planes`): the rows are stored straight into the neighbours' frame rows over NVLink and the ranks synchronise
    through counters in symmetric memory.  With NCCL send/recv the two exchanges cost more than the 0.12 ms slab step itself on
    8 GPUs (efficiency 0.44).  The state buffers and the affinity buffer live in one symmetric allocation.  Collective: every
    rank calls `step()` / `set_interior()` / `close()` the same number of times."""

    AFF_ROWS = 2   # the redistribution reads the affinity on a 2-row halo (macelenia.py:103-118: 3x3 of a 3x3-normalised field)

    def __init__(self, lenia_params, C, H, W, rank, world, device="cuda", beta=8.0, alpha=None, conv_path="auto", group=None):
        import torch.distributed._symmetric_memory as symm

        super().__init__(lenia_params, C, H, W, rank, world, device=device, beta=beta, alpha=alpha, conv_path=conv_path)
        if world < 2:
            raise ValueError("MaceSlabP2P needs at least two ranks (MaceSlabCUDA runs a single slab)")
        self.group = group if group is not None else dist.group.WORLD
        n = self.B * self.C * self.rows * self.pitch
        self._n = n
        self._sym = symm.empty(3 * n + 64, dtype=torch.float32, device=self.device)
        self._sym.zero_()
        self._hdl = symm.rendezvous(self._sym, self.group)
        shape = (self.B, self.C, self.rows, self.pitch)
        self.bufs = [self._sym[i * n:(i + 1) * n].view(shape) for i in range(2)]
        self.aff = self._sym[2 * n:3 * n].view(shape)
        self._counters = self._sym[3 * n:3 * n + 64].view(torch.int32)   # [0:2] ready, [2:4] arrived, [4] ticket, [5] error
        self._base = list(self._hdl.buffer_ptrs)
        self._up, self._down = ring_neighbours(rank, world)
        self._cnt = 3 * n * 4
        self.epoch = 0
        self.__dict__.pop("_exchange_ops", None)

    def exchange(self, t=None, group=None):
        from . import _lib

        t = self.buf if t is None else t
        nrows = self.AFF_ROWS if t.data_ptr() == self.aff.data_ptr() else self.F
        F, hl, row = self.F, self.hl, self.pitch * 4
        off = t.data_ptr() - self._sym.data_ptr()
        me, up, dn = self._base[self.rank] + off, self._base[self._up] + off, self._base[self._down] + off
        cme, cup, cdn = self._base[self.rank] + self._cnt, self._base[self._up] + self._cnt, self._base[self._down] + self._cnt
        self.epoch += 1
        with _lib.device_guard(self.device):
            # my first rows -> the up neighbour's bottom frame rows; my last rows -> the down neighbour's top frame rows
            _lib.call("b200ca_halo_push_planes", me + F * row, up + (hl + F) * row, me + (hl + F - nrows) * row, dn + (F - nrows) * row,
                      nrows * row, self.B * self.C, self.rows * row, cup + 4, cdn, cme, cup + 12, cdn + 8, cme + 8,
                      self.epoch & 0xFFFFFFFF, cme + 16, cme + 20, _lib.current_stream(self.device))

    def timed_out(self) -> bool:
        return bool(self._counters[5].item())

    def close(self):
        if getattr(self, "_sym", None) is not None:
            torch.cuda.synchronize(self.device)
            dist.barrier(self.group)
            self.bufs = []
            self.aff = self._counters = self._hdl = self._sym = None
