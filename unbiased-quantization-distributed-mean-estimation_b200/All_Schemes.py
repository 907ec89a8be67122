"""Drop-in for the reference's `All_Schemes` module (NMSE_Results/Codes/All_Schemes.py, AS).

Same function names, signatures `f(input_vector, bits_per_dimension=1)`, return types and `__name__`s
(the Flower scripts use `quantization_func.__name__` for directories and bit-depth tables, TU:191, TU:337), so
`from All_Schemes import *` (ND:2, TU:23) keeps working:

    import sys; sys.path.insert(0, "<repo>")
    import dme_b200.All_Schemes as All_Schemes; sys.modules["All_Schemes"] = All_Schemes

Every function runs the sm_100a kernels of libdme_b200.so on the current CUDA device; there is no CPU fallback.
Randomness: the reference draws from torch's global generator (AS:634, AS:735, AS:783, AS:800).  Here the draws are
Philox streams keyed by a seed derived from `torch.initial_seed()` and a per-process call counter, so
`torch.manual_seed(s)` still makes a run reproducible.  Results are statistically, not bitwise, equal to the
reference's (bitwise equality holds under injected draws: see tests/ and `dme_b200.api`).
"""
from __future__ import annotations

import itertools

import numpy as np
import torch

try:                                     # imported as dme_b200.All_Schemes
    from . import api as _api
except ImportError:                      # imported as top-level `All_Schemes` with the package dir on sys.path
    import dme_b200.api as _api

device = torch.device("cuda" if torch.cuda.is_available() else "cpu")     # AS:22 (no print: import stays silent)

# AS:539 (integer-rate copy of the table; the functions use the 20-entry one, AS:614-620)
Type_quantize_algo_rate_l_dict = {1: 0.21403, 2: 0.63752, 3: 1.41725, 4: 2.91504, 5: 5.87195, 6: 11.76507,
                                  7: 23.54075, 8: 47.0868, 9: 94.17625, 10: 188.35383}

#: "reference" reproduces the reference's DRIVE transform; "correct" is real DRIVE.
DRIVE_COMPAT = "reference"

_calls = itertools.count()
_keyed_to = None


def _seed() -> int:
    """Philox seed of the next call: torch's seed and a call counter that restarts whenever the seed changes, so that
    `torch.manual_seed(s)` reproduces a run from that point on (the reference's draws come from torch's global generator)."""
    global _calls, _keyed_to
    s0 = int(torch.initial_seed())
    if s0 != _keyed_to:
        _keyed_to, _calls = s0, itertools.count()
    return (s0 * 0x9E3779B97F4A7C15 + next(_calls) * 0xD1342543DE82EF95) & 0xFFFFFFFFFFFFFFFF


def _vec(input_vector):
    """AS:611: any array-like / tensor -> 1-D fp32 CUDA tensor (a copy; the input is never modified)."""
    if isinstance(input_vector, torch.Tensor):
        t = input_vector.detach()
    else:
        t = torch.as_tensor(np.asarray(input_vector))
    return t.to(device="cuda", dtype=torch.float32).reshape(-1)


def Type_unbiased_quantize(input_vector, bits_per_dimension=1):
    """AS:609-641.  Unknown rates raise KeyError like the reference's table lookup (AS:623)."""
    return _api.type_quantize(_vec(input_vector), bits_per_dimension, mode="unbiased", seed=_seed())["deq"]


def Type_biased_quantize(input_vector, bits_per_dimension=1):
    """AS:669-687 (Reznik rounding AS:644-666; ties in the mass repair go to the lowest index)."""
    return _api.type_quantize(_vec(input_vector), bits_per_dimension, mode="biased")["deq"]


def DRIVE_quantize_Hadamard(input_vector, bits_per_dimension=1):
    """AS:707-752 (`bits_per_dimension` is ignored there too)."""
    return _api.drive(_vec(input_vector), seed=_seed(), compat=DRIVE_COMPAT)


def Scalar_quantize(input_vector, bits_per_dimension=1):
    """AS:755-790."""
    return _api.scalar_quantize(_vec(input_vector), bits_per_dimension, seed=_seed())


def EDEN_quantize_Hadamard(input_vector, bits_per_dimension=1):
    """AS:793-812.  Returns a numpy array on the host like the reference (AS:812).  Rates: 1, 2, below 1 (1-bit quantization,
    the receiver drops coordinates, AS:413-421) and fractional rates between 1 and 2 (AS:352-368); any other rate has no
    centroids in the reference either (AS:301-320) and raises KeyError like its table lookup."""
    x = _vec(input_vector)
    seed = int(torch.randint(0, 100, (1,)).item())                          # AS:800: per-call rotation seed
    return _api.eden(x, bits_per_dimension, seed=seed).cpu().numpy()


def QUICFL_quantize(input_vector, bits_per_dimension=1):
    """AS:814-832: QuicFLSender.compress + QuicFLReceiver.decompress, rotation seed 123, per-call seed in [0, 100) (AS:820); returns
    numpy (AS:832).  The reference itself raises FileNotFoundError here, because its sender tables are not in its tree;
    `dme_b200.quicfl` derives them from the shipped receiver tables (quicfl_tables.py).  Rates 1-4 bits (AS:430); other rates raise KeyError
    like the reference's table lookup."""
    if bits_per_dimension not in (1, 2, 3, 4):
        raise KeyError(bits_per_dimension)
    seed = int(torch.randint(0, 100, (1,)).item())                          # AS:820
    return _api.quicfl(_vec(input_vector), int(bits_per_dimension), seed=seed, rotation_seed=123).cpu().numpy()


def Kashin_quantize(input_vector, bits_per_dimension=1):
    """AS:834-854: Kashin frame coefficients (AS:191-239, eta=0.9, delta=1, pad_threshold=0.85, 3 iterations) then
    min/max stochastic quantization (AS:62-91) -- `dme_b200.kashin`.  Returns numpy (AS:854)."""
    seed = int(torch.randint(0, 100, (1,)).item())                          # AS:841: one of 100 Bernoulli streams (AS:67)
    return _api.kashin(_vec(input_vector), bits_per_dimension, seed=seed, rotation_seed=123).cpu().numpy()


def No_quantize(input_vector, bits_per_dimension=1):
    """AS:856-859."""
    return _vec(input_vector).clone()


def _write_back(arg, result):
    """The reference's transforms work in place; keep that for tensors we can write to (same shape, fp32, CUDA)."""
    if isinstance(arg, torch.Tensor) and arg.is_cuda and arg.dtype == torch.float32 and arg.shape == result.shape:
        arg.copy_(result)
        return arg
    return result


class Hadamard:
    """AS:94-120 (rotation helper classes of the upstream QUIC-FL code)."""

    def __init__(self, device=device):
        self.device = device

    def hadamard(self, vec):
        """AS:100-115 transforms its argument in place and returns it: a CUDA fp32 tensor is written back."""
        return _write_back(vec, _api.hadamard(vec))

    def random_diagonal(self, size, seed):
        return _api.rademacher(size, seed)


class HadamardSender(Hadamard):
    def randomized_hadamard_transform(self, vec, seed):
        return _api.rht(vec, seed)


class HadamardReceiver(Hadamard):
    def randomized_inverse_hadamard_transform(self, vec, seed):
        """AS:151-156 mutates its argument (AS:153) and returns it: a CUDA fp32 tensor is written back."""
        return _write_back(vec, _api.irht(vec, seed))


__all__ = ["Type_unbiased_quantize", "Type_biased_quantize", "DRIVE_quantize_Hadamard", "Scalar_quantize",
           "EDEN_quantize_Hadamard", "QUICFL_quantize", "Kashin_quantize", "No_quantize", "Hadamard", "HadamardSender",
           "HadamardReceiver", "Type_quantize_algo_rate_l_dict", "device", "torch", "np"]
