"""CPU: libb200rbf.so loads, exports every symbol include/b200rbf.h declares, and fails loudly without a GPU."""
import ctypes
import os
import re

import pytest

from conftest import ROOT
from pysot_b200 import _lib


def header_symbols():
    text = open(os.path.join(ROOT, "include", "b200rbf.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200rbf_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    syms = header_symbols()
    assert len(syms) >= 20
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for s in syms:
        assert hasattr(lib, s), "libb200rbf.so does not export %s" % s


def test_python_binding_covers_header():
    assert sorted(_lib.PROTOTYPES) == header_symbols()


def test_version_and_error_string_need_no_gpu():
    lib = _lib.load()
    assert lib.b200rbf_version() >= 100
    assert isinstance(lib.b200rbf_last_error(), bytes)


def test_no_cpu_fallback():
    """Without a visible B200 creating a handle must raise; nothing silently computes on the host."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is visible")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.Handle(0)
    from pysot_b200 import GpuRBFInterpolant
    import numpy as np

    s = GpuRBFInterpolant(dim=2, lb=np.zeros(2), ub=np.ones(2))
    s.add_points(np.random.rand(5, 2), np.random.rand(5))
    with pytest.raises(RuntimeError):
        s.predict(np.zeros(2))


def test_product_code_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "pysot_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("no oracle", ""), "%s mentions the oracle" % f
