"""Imports the UNMODIFIED reference (`/root/reference/NMSE_Results/Codes/All_Schemes.py`) so that
golden vectors can be generated from it -- TEST INFRASTRUCTURE ONLY, build container only.

`/root/reference` does not exist on the GPU box; nothing under tests/ -m gpu, smoke() or bench.py
imports this module.  It is used by tests/golden/make_golden.py (fixtures are committed) and by the
optional CPU tests that are skipped when the reference tree is absent.

The reference imports matplotlib (AS:8) without using it; matplotlib is not installed here, so an
empty stub module is injected.  No reference source is edited or copied.
"""
from __future__ import annotations

import contextlib
import os
import sys
import types

REF_DIR = "/root/reference/NMSE_Results/Codes"


def available() -> bool:
    return os.path.exists(os.path.join(REF_DIR, "All_Schemes.py"))


_mod = None


def load():
    """Return the reference's All_Schemes module (imported once)."""
    global _mod
    if _mod is None:
        if not available():
            raise FileNotFoundError(REF_DIR)
        for name in ("matplotlib", "matplotlib.pyplot"):
            sys.modules.setdefault(name, types.ModuleType(name))
        sys.path.insert(0, REF_DIR)
        try:
            import io
            with contextlib.redirect_stdout(io.StringIO()):   # AS:23 prints the device
                import All_Schemes as _m
        finally:
            sys.path.remove(REF_DIR)
        _mod = _m
    return _mod


@contextlib.contextmanager
def inject(rand=None, rand_like=None, randint=None):
    """Replace torch.rand / torch.rand_like / torch.randint (the reference's global-RNG draws:
    AS:634, AS:735, AS:783, AS:800) by fixed values for the duration of the block."""
    import torch
    saved = (torch.rand, torch.rand_like, torch.randint)
    try:
        if rand is not None:
            torch.rand = lambda *a, **k: torch.tensor([rand], dtype=torch.float32)
        if rand_like is not None:
            it = iter(rand_like) if isinstance(rand_like, (list, tuple)) else None

            def _rl(t, *a, **k):
                src = next(it) if it is not None else rand_like
                return torch.as_tensor(src, dtype=torch.float32)[: t.numel()].reshape(t.shape).clone()
            torch.rand_like = _rl
        if randint is not None:
            torch.randint = lambda *a, **k: torch.tensor([randint])
        yield
    finally:
        torch.rand, torch.rand_like, torch.randint = saved
