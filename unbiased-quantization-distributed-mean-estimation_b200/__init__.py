"""B200-native (sm_100a) distributed-mean-estimation hot path.

Import as `dme_b200` (the repo-root alias package; this directory's name is not a Python identifier).
The compute path is libdme_b200.so (hand-written CUDA behind a C ABI, include/dme_b200.h); PyTorch is used
for device memory, streams and torch.distributed only.  There is no CPU fallback.
"""
from . import _cabi  # noqa: F401
from .api import (RATE_TABLE, DmeError, m_for_rate, l1_norms, type_quantize, type_encode, decode_mean, quantize_mean,
                  quantize_mean_host, decode_mean_host, codes_to_host, quantize_mean_sliced, rotated_type_quantize, rotated_quantize_mean, hadamard, rht, irht, rademacher, pair_transform, drive, eden, eden_encode, eden_decode,
                  quicfl_decode, quicfl_encode, quicfl_decode_dense, quicfl, scalar_quantize, kashin, kashin_padded_dim, mean_accumulate, client_uniforms, set_unbiased_path, set_biased_path, profile_kernels, PackedCodes, Workspace, MeanGraph)

__all__ = ["RATE_TABLE", "DmeError", "m_for_rate", "l1_norms", "type_quantize", "type_encode", "decode_mean", "quantize_mean",
           "quantize_mean_host", "decode_mean_host", "codes_to_host", "quantize_mean_sliced", "rotated_type_quantize", "rotated_quantize_mean", "hadamard", "rht", "irht", "rademacher", "pair_transform", "drive", "eden", "eden_encode",
           "eden_decode", "quicfl_decode", "quicfl_encode", "quicfl_decode_dense", "quicfl", "scalar_quantize", "kashin", "kashin_padded_dim", "mean_accumulate", "client_uniforms", "set_unbiased_path", "set_biased_path", "profile_kernels", "PackedCodes", "Workspace", "MeanGraph"]
