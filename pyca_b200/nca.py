"""B200 drop-in for PyCA's ``NeuralCA`` (pyca/automata/models/nca.py).

Same constructor (``NeuralCA(size, models_folder, seed=None, device)``), ``state`` (B,n,H,W) NCHW
tensor, ``step``/``draw``/``insert_seed``/``reset``; checkpoints are the reference's own
``{'config', 'state_dict'}`` files (nca.py:286-303).  ``step()`` (nca.py:67-72 -> NCAModule.forward
:239-263) is one fused CUDA update pass + a tiny life-mask pass through the C ABI.

Update mask (``torch.rand(B,1,H,W) <= 0.5``, nca.py:256): ``rng='philox'`` (default) draws it inside the
kernel from Philox4x32-10(seed, step, cell); ``rng='torch'`` draws it with torch on the same device, i.e.
the exact stream the reference consumes on CUDA; ``step(update_mask=...)`` takes an explicit mask
(parity tests).
"""
from __future__ import annotations

from pathlib import Path

import torch

from . import _lib
from .automaton import Automaton


class NCAWeights:
    """Device-side prepared weights of one checkpoint (folded W1, transposed W2, tensor-core images)."""

    def __init__(self, W1, b1, W2, b2, device):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.B200CAError(f"NeuralCA needs a CUDA device (got {device}); there is no CPU path")
        self.n_states = W2.shape[0]
        self.n_hidden = W1.shape[0]
        if W1.shape[1] != 3 * self.n_states or W2.shape[1] != self.n_hidden:
            raise _lib.B200CAError(f"inconsistent NCA weight shapes {tuple(W1.shape)} / {tuple(W2.shape)}")
        nbytes = _lib.lib().b200ca_nca_wprep_bytes(self.n_states, self.n_hidden)
        if nbytes < 0:
            raise _lib.B200CAError(f"NCA with n_states={self.n_states}, n_hidden={self.n_hidden} is not built (only 16/128)")
        self.W1, self.b1, self.W2, self.b2 = [t.detach().to(self.device, torch.float32).contiguous() for t in (W1, b1, W2, b2)]
        self.prep = torch.zeros(nbytes, dtype=torch.uint8, device=self.device)
        with _lib.device_guard(self.device):
            _lib.call("b200ca_nca_prepare_weights", _lib.ptr(self.W1), _lib.ptr(self.b1), _lib.ptr(self.W2), _lib.ptr(self.b2),
                      _lib.ptr(self.prep), self.n_states, self.n_hidden, _lib.current_stream(self.device))

    @classmethod
    def from_checkpoint(cls, path, device):
        ck = torch.load(path, map_location="cpu", weights_only=False)
        sd = ck["state_dict"]
        return cls(sd["computer.0.weight"], sd["computer.0.bias"], sd["computer.2.weight"], sd["computer.2.bias"], device)


def nca_step(state: torch.Tensor, weights: NCAWeights, out: torch.Tensor = None, update_mask: torch.Tensor = None,
             seed: int = 0, step_index: int = 0, scratch: torch.Tensor = None, impl: int = _lib.NCA_SIMT,
             first_world: int = 0) -> torch.Tensor:
    """One NCA step on a dense (B,n,H,W) fp32 CUDA tensor; returns the new state (a different tensor).
    `first_world`: index of state[0] in a larger batch that is split over GPUs -- the in-kernel update mask is drawn from
    the global cell index, so the shards reproduce the unsplit batch bit for bit."""
    _lib.require_cuda(state, "state")
    B, n, H, W = state.shape
    if state.dtype != torch.float32 or not state.is_contiguous():
        state = state.float().contiguous()
    if out is None:
        out = torch.empty_like(state)
    if scratch is None:
        scratch = torch.empty(_lib.lib().b200ca_nca_scratch_bytes(B, H, W), dtype=torch.uint8, device=state.device)
    m = None
    if update_mask is not None:
        m = update_mask.to(device=state.device).reshape(B, H, W).to(torch.uint8).contiguous()
    with _lib.device_guard(state.device):
        _lib.call("b200ca_nca_step_shard", _lib.ptr(state), _lib.ptr(out), _lib.ptr(weights.prep), _lib.ptr(m), seed, step_index,
                  int(first_world), _lib.ptr(scratch), B, n, weights.n_hidden, H, W, impl, _lib.current_stream(state.device))
    return out


class NeuralCA(Automaton):
    """
    Neural Cellular Automaton. Can be trained to grow any image from a 'seed', using only local
    interactions, and a neural network.  (B200 inference step.)
    """

    def __init__(self, size, models_folder, seed=None, device="cuda", batch=1, rng="philox", rng_seed=0, impl=None):
        super().__init__(size, device)
        self.device = device
        self.models_paths = sorted(Path(models_folder).rglob("*.pt"))
        if not self.models_paths:
            raise FileNotFoundError(f"no NCA checkpoints under {models_folder}")
        self.cur_model = 0
        self.load_model(self.models_paths[self.cur_model])
        n = self.weights.n_states
        self.seed = torch.ones((n, 1, 1), device=device) if seed is None else seed
        self.seed_size = self.seed.shape[1:]
        self.batch = batch
        self.state = torch.zeros((batch, n, size[0], size[1]), device=device)
        self.insert_seed(size[0] // 2, size[1] // 2)
        self.brush_size = 4
        self.rng = rng
        self.rng_seed = rng_seed
        self.impl = _lib.NCA_SIMT if impl is None else impl
        self.frames = 0
        self._spare = None
        self._scratch = None
        self._worldmap = self._worldmap.to(device)

    def load_model(self, model_path):
        self.weights = NCAWeights.from_checkpoint(model_path, self.device)

    @torch.no_grad()
    def step(self, update_mask: torch.Tensor = None):
        """nca.py:67-72."""
        st = self.state
        B, n, H, W = st.shape
        if self._spare is None or self._spare.shape != st.shape or self._spare.data_ptr() == st.data_ptr():
            self._spare = torch.empty_like(st, memory_format=torch.contiguous_format)
        if self._scratch is None or self._scratch.numel() != _lib.lib().b200ca_nca_scratch_bytes(B, H, W):
            self._scratch = torch.empty(_lib.lib().b200ca_nca_scratch_bytes(B, H, W), dtype=torch.uint8, device=st.device)
        if update_mask is None and self.rng == "torch":
            update_mask = torch.rand(B, 1, H, W, device=st.device) <= 0.5
        new = nca_step(st, self.weights, out=self._spare, update_mask=update_mask, seed=self.rng_seed, step_index=self.frames,
                       scratch=self._scratch, impl=self.impl)
        self._spare = st if (st.is_contiguous() and st.dtype == torch.float32) else None
        self.state = new
        self.frames += 1

    @torch.no_grad()
    def draw(self):
        """nca.py:74-87 (without the brush overlay, which needs the GUI's mouse position): RGB = clamp(rgb) * clamp(alpha),
        one kernel writing the float `_worldmap` and the (W,H,3) uint8 frame `worldmap` returns."""
        st = self.state
        if not (st.is_contiguous() and st.dtype == torch.float32):
            st = self.state = st.float().contiguous()
        B, n, H, W = st.shape
        wm = self._worldmap
        if not (wm.is_cuda and wm.dtype == torch.float32 and wm.is_contiguous() and tuple(wm.shape) == (3, H, W)):
            wm = self._worldmap = torch.empty((3, H, W), dtype=torch.float32, device=st.device)
        fb = self._frame()
        with _lib.device_guard(st.device):
            _lib.call("b200ca_draw_nca", _lib.ptr(st), _lib.ptr(wm), _lib.ptr(fb.dev), 0, n, H, W, _lib.current_stream(st.device))
        fb.mark(wm)

    def insert_seed(self, h, w):
        """nca.py:151-162."""
        sh, sw = self.seed_size
        if h + sh // 2 < self.size[0] and w + sw // 2 < self.size[1]:
            hs = slice(max(0, h - (sh + 1) // 2), min(h + sh // 2, self.size[0]))
            ws = slice(max(0, w - (sw + 1) // 2), min(w + sw // 2, self.size[1]))
            self.state[:, :, hs, ws] = self.seed

    def reset(self):
        self.state = torch.zeros_like(self.state)
