"""CPU: host-side logic -- the C ABI exports every symbol the header declares, the ctypes table matches the header,
the rate table / drop-in names mirror the reference, client sharding, and the product refuses to run without CUDA."""
import ctypes
import inspect
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    hdr = open(os.path.join(ROOT, "include", "dme_b200.h")).read()
    return sorted(set(re.findall(r"DME_API\s+[\w\s\*]+?\b(dme_\w+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as ge
    ge.build()
    from dme_b200 import _cabi
    names = _declared()
    assert len(names) >= 25
    lib = ctypes.CDLL(_cabi.LIB_PATH)
    for nm in names:
        assert hasattr(lib, nm), f"{nm} declared in include/dme_b200.h but not exported"
        assert nm in _cabi.SIGNATURES, f"{nm} missing from the ctypes table"
    assert sorted(_cabi.SIGNATURES) == names
    L = _cabi.lib()
    assert L.dme_version() >= 100 and L.dme_launch_count() == 0


def test_host_only_entry_points():
    from dme_b200 import _cabi
    L = _cabi.lib()
    a = [L.dme_uniform_x(7, c) for c in range(1000)]
    assert a == [L.dme_uniform_x(7, c) for c in range(1000)]                     # counter-based: pure function
    assert all(0.0 <= v < 1.0 for v in a) and 0.45 < float(np.mean(a)) < 0.55
    assert len(set(a)) > 990 and L.dme_uniform_x(8, 0) != a[0]
    assert all(abs(v * 2 ** 24 - round(v * 2 ** 24)) == 0 for v in a)            # on torch.rand's 2^-24 grid
    assert L.dme_workspace_bytes(128, 1 << 24) > 128 * 4096 * 24     # 8-byte L1 partial + 16-byte look-back record per tile
    assert L.dme_dir_entries(3, 4097) == 3 * 5                      # one entry per 1024-coordinate code tile
    assert L.dme_codes_bytes(128, 1 << 24, 3590827, 1) < L.dme_codes_bytes(128, 1 << 24, 3590827, 0)
    assert L.dme_codes_bytes(128, 1 << 24, 3590827, 0) >= 128 * (1 << 24) * 4


def test_rate_table_matches_reference_api():
    import dme_b200 as dme
    assert len(dme.RATE_TABLE) == 20 and dme.RATE_TABLE[1] == 0.21403 and dme.RATE_TABLE[10] == 188.35383
    assert dme.m_for_rate(1, 1 << 24) == 3590827 and dme.m_for_rate(2, 1024) == 652 and dme.m_for_rate(4, 8) == 23
    with pytest.raises(KeyError):
        dme.m_for_rate(0.7, 100)                                                 # AS:623: dict lookup of an unknown rate


def test_dropin_module_mirrors_reference_names():
    import dme_b200.All_Schemes as AS
    for nm in ("Type_unbiased_quantize", "Type_biased_quantize", "DRIVE_quantize_Hadamard", "Scalar_quantize",
               "EDEN_quantize_Hadamard", "QUICFL_quantize", "Kashin_quantize", "No_quantize"):
        fn = getattr(AS, nm)
        assert fn.__name__ == nm                                                 # TU:191 / TU:337 use __name__
        sig = inspect.signature(fn)
        assert list(sig.parameters) == ["input_vector", "bits_per_dimension"] and sig.parameters["bits_per_dimension"].default == 1
    assert AS.Type_quantize_algo_rate_l_dict[2] == 0.63752
    with pytest.raises(KeyError):
        AS.QUICFL_quantize(np.zeros(4, np.float32), 5)                           # tables exist for 1-4 bits (AS:430)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import dme_b200 as dme
    with pytest.raises(dme.DmeError, match="no CUDA device"):
        dme.type_quantize(np.ones(16, np.float32), 1)
    with pytest.raises(dme.DmeError):
        dme.quantize_mean(np.ones((2, 16), np.float32), 1)


def test_shard_clients():
    from dme_b200.distributed import shard_clients
    for n_total, world in ((1024, 8), (10, 4), (3, 8), (128, 1)):
        blocks = [shard_clients(n_total, r, world) for r in range(world)]
        assert sum(b[1] for b in blocks) == n_total
        pos = 0
        for c0, nl in blocks:
            assert c0 == pos
            pos += nl
        assert max(b[1] for b in blocks) - min(b[1] for b in blocks) <= 1
    with pytest.raises(ValueError):
        shard_clients(10, 4, 4)
