"""The Flower hook of the reference (SImulation_Results_datasets/CIFAR10/Codes/Type_unbiased.py, TU) on the packed code.

The reference's client quantizes its model delta, immediately DEquantizes it and ships the dense fp32 result (TU:164-229); the
server is flwr's stock FedAvg (TU:260-269, flwr 1.11.1: sum_c num_examples_c * w_c / sum_c num_examples_c).  Here the same round
is an encode -> wire -> decode path: `ClientCodec.encode` turns the delta into one DMEP1 message (`PackedCodes.to_messages`),
`aggregate_fit` decodes all messages with one launch of the fused decode-mean kernel, weighted by num_examples, and adds the
global parameters back.  flwr itself is not needed (it is not installed here): `TypeCodecStrategy.aggregate_fit` takes the same
`results` shape flwr hands to a Strategy -- a list of (client, FitRes-like) whose `.parameters` carry the message bytes and
`.num_examples` the weight -- so it can be mixed into `fl.server.strategy.FedAvg` unchanged where flwr is available.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np
import torch

from . import api


def flatten(arrays: Sequence[np.ndarray]):
    """TU:166-168: parameter arrays -> (flat fp32 vector, shapes, sizes)."""
    shapes = [a.shape for a in arrays]
    sizes = [int(a.size) for a in arrays]
    return np.concatenate([np.asarray(a, np.float32).reshape(-1) for a in arrays]), shapes, sizes


def unflatten(vec: np.ndarray, shapes, sizes) -> List[np.ndarray]:
    """TU:221-227."""
    out, off = [], 0
    for shp, sz in zip(shapes, sizes):
        out.append(vec[off: off + sz].reshape(shp))
        off += sz
    return out


class ClientCodec:
    """Client side of a round: delta = params - global (TU:169), quantize with the type quantizer, ship the CODE."""

    def __init__(self, bits_per_dimension=1, mode="unbiased"):
        self.bits, self.mode = bits_per_dimension, mode

    def encode(self, params: Sequence[np.ndarray], global_flat: np.ndarray, seed: int, client_id: int = 0) -> bytes:
        flat, _, _ = flatten(params)
        delta = torch.from_numpy(flat - np.asarray(global_flat, np.float32)).cuda()                   # TU:169, TU:173
        pc = api.type_encode(delta, self.bits, mode=self.mode, seed=seed, client0=client_id)
        return pc.to_messages(seed=seed, client0=client_id)[0]

    def encode_many(self, deltas: torch.Tensor, seed: int, client0: int = 0) -> List[bytes]:
        """All clients of a simulated round at once (rows of `deltas`): one quantize launch."""
        return api.type_encode(deltas, self.bits, mode=self.mode, seed=seed, client0=client0).to_messages(seed=seed, client0=client0)


def aggregate_fit(messages: Sequence[bytes], num_examples: Sequence[int], global_flat: np.ndarray, shapes=None, sizes=None):
    """Server side (replaces FedAvg.aggregate_fit's weighted average over dense arrays, TU:260-269): decode the clients' codes,
    average them weighted by num_examples on the GPU, add the global parameters back (TU:220).  -> flat vector or list of arrays."""
    pc = api.PackedCodes.from_messages(list(messages))
    mean_delta = api.decode_mean(pc, weights=list(num_examples)).cpu().numpy()
    new_flat = mean_delta + np.asarray(global_flat, np.float32)
    return unflatten(new_flat, shapes, sizes) if shapes is not None else new_flat


class TypeCodecStrategy:
    """Duck-typed Strategy mix-in: `aggregate_fit(server_round, results, failures)` with flwr's argument shapes.  `results` is a
    list of (client_proxy, fit_res) where fit_res.parameters is the DMEP1 message (bytes, or an object with `.tensors[0]`) and
    fit_res.num_examples the client's weight (TU:236: len(trainloader))."""

    def __init__(self, initial_parameters: Sequence[np.ndarray]):
        self.global_flat, self.shapes, self.sizes = flatten(initial_parameters)

    def aggregate_fit(self, server_round: int, results: Sequence[Tuple[object, object]], failures=()):
        if not results:
            return None, {}
        msgs, weights = [], []
        for _, res in results:
            p = res.parameters
            msgs.append(p if isinstance(p, (bytes, bytearray)) else bytes(p.tensors[0]))
            weights.append(int(res.num_examples))
        arrays = aggregate_fit(msgs, weights, self.global_flat, self.shapes, self.sizes)
        self.global_flat, _, _ = flatten(arrays)
        return arrays, {"clients": len(msgs), "bytes_up": int(sum(len(m) for m in msgs))}
