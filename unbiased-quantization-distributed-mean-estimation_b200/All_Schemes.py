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


def _seed() -> int:
    return (int(torch.initial_seed()) * 0x9E3779B97F4A7C15 + next(_calls) * 0xD1342543DE82EF95) & 0xFFFFFFFFFFFFFFFF


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
    """AS:793-812.  Returns a numpy array on the host like the reference (AS:812)."""
    x = _vec(input_vector)
    seed = int(torch.randint(0, 100, (1,)).item())                          # AS:800: per-call rotation seed
    if bits_per_dimension not in (1, 2):
        if bits_per_dimension == round(bits_per_dimension):
            raise KeyError(int(bits_per_dimension))                        # AS:301-320: centroids for 1 and 2 bits only
        raise NotImplementedError("fractional EDEN rates (AS:352-368) are not built yet")
    return _api.eden(x, int(bits_per_dimension), seed=seed).cpu().numpy()


def QUICFL_quantize(input_vector, bits_per_dimension=1):
    """AS:814-832.  The reference cannot run this either: its sender tables are not shipped.  The
    receiver is available as `dme_b200.quicfl_decode`."""
    raise FileNotFoundError("QUIC-FL sender tables (*_sender_table_X.pt / *_sender_table_p.pt) are not part of the "
                            "reference tree; only the receiver (dme_b200.quicfl_decode) can be built from it")


def Kashin_quantize(input_vector, bits_per_dimension=1):
    """AS:834-854: Kashin frame coefficients (AS:191-239, eta=0.9, delta=1, pad_threshold=0.85, 3 iterations) then
    min/max stochastic quantization (AS:62-91).  Rotations and the quantizer are the sm_100a kernels; the clamp /
    residual updates between them are elementwise torch ops on the device.  Returns numpy (AS:854)."""
    x = _vec(input_vector)
    dim = x.numel()
    seed, rot_seed = int(torch.randint(0, 100, (1,)).item()), 123
    eta, delta, pad_threshold, niters = 0.9, 1.0, 0.85, 3
    pdim = 1 << int(np.ceil(np.log2(dim))) if dim & (dim - 1) else 2 * dim                      # AS:203-211
    if dim & (dim - 1) and dim / pdim > pad_threshold:
        pdim *= 2
    coeff = torch.zeros(pdim, device=x.device)
    resid = x.clone()
    M = torch.norm(resid) / np.sqrt(delta * pdim)                                                # AS:221
    for i in range(niters):
        padded = torch.zeros(pdim, device=x.device)
        padded[:dim] = resid
        b = _api.rht(padded, rot_seed)                                                           # AS:225
        b_hat = torch.clamp(b, min=-M, max=M)
        coeff += b_hat                                                                           # AS:229
        if i < niters - 1:
            resid = resid - _api.irht(b_hat, rot_seed)[:dim]                                     # AS:232-233
            M = M * eta
        err = (x - _api.irht(coeff, rot_seed)[:dim]).norm(2) / resid.norm(2)                     # AS:236
        if err < 1e-6:
            break
    q = _api.scalar_quantize(coeff, bits_per_dimension, seed=seed * 1000003 + 17)                # AS:62-91 ("standard" step)
    return _api.irht(q, rot_seed)[:dim].cpu().numpy()                                            # AS:262-267, AS:854


def No_quantize(input_vector, bits_per_dimension=1):
    """AS:856-859."""
    return _vec(input_vector).clone()


class Hadamard:
    """AS:94-120 (rotation helper classes of the upstream QUIC-FL code)."""

    def __init__(self, device=device):
        self.device = device

    def hadamard(self, vec):
        return _api.hadamard(vec)

    def random_diagonal(self, size, seed):
        return _api.rademacher(size, seed)


class HadamardSender(Hadamard):
    def randomized_hadamard_transform(self, vec, seed):
        return _api.rht(vec, seed)


class HadamardReceiver(Hadamard):
    def randomized_inverse_hadamard_transform(self, vec, seed):
        return _api.irht(vec, seed)


__all__ = ["Type_unbiased_quantize", "Type_biased_quantize", "DRIVE_quantize_Hadamard", "Scalar_quantize",
           "EDEN_quantize_Hadamard", "QUICFL_quantize", "Kashin_quantize", "No_quantize", "Hadamard", "HadamardSender",
           "HadamardReceiver", "Type_quantize_algo_rate_l_dict", "device", "torch", "np"]
