"""pyca_b200 -- B200-native (sm_100a) `Automaton.step()` kernels behind PyCA's Automaton API.

Only the hot path of frotaur/PyCA is implemented here: Game of Life (CA2D), Lenia / multichannel
Lenia, MaCE-Lenia (+ cross-channel variant, food, Fourier growth functions), FlowLenia and Neural-CA inference.  The update runs in hand-written
CUDA reached through the C ABI in ``include/b200ca.h``; there is no CPU or PyTorch fallback.
"""
from . import _lib
from ._lib import B200CAError, lib
from .automaton import AUTOMATAS, Automaton
from .ca2d import CA2D, PackedWorld
from .lenia import FlowLenia, FramedState, Lenia, MaCELenia, MaCELeniaXChan
from .lenia_params import LeniaParams, compute_kernel
from .nca import NCAWeights, NeuralCA, nca_step

__all__ = ["AUTOMATAS", "Automaton", "B200CAError", "CA2D", "FlowLenia", "FramedState", "Lenia", "LeniaParams", "MaCELenia",
           "MaCELeniaXChan", "NCAWeights", "NeuralCA", "PackedWorld", "compute_kernel", "lib", "nca_step"]
