"""Importable alias of the product package.

The product lives in `unbiased-quantization-distributed-mean-estimation_b200/` (the name the build contract
fixes), which is not a valid Python identifier; this stub makes it importable as `dme_b200`.
"""
import os as _os

_real = _os.path.normpath(_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), _os.pardir,
                                        "unbiased-quantization-distributed-mean-estimation_b200"))
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _f
