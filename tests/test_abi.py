"""The C-ABI library must load and export every entry point include/pclean_b200.h declares,
and must fail loudly (no CPU fallback) when no CUDA device is present."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from pclean_b200 import engine
    engine.build()
    return C.CDLL(engine.LIB_PATH)


def test_exports_every_declared_symbol(lib):
    header = open(os.path.join(ROOT, "include", "pclean_b200.h")).read()
    names = sorted(set(re.findall(r"\b(pclean_[a-z0-9_]+)\s*\(", header)))
    assert len(names) >= 20
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_python_binding_lists_match(lib):
    from pclean_b200 import engine
    assert all(hasattr(lib, n) for n in engine.EXPORTS)


def test_create_fails_loudly_without_gpu(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from pclean_b200.lowering import Config
    cfg = Config(1, 2, 1, 1, 0, 50, 100)
    h = C.c_void_p()
    rc = lib.pclean_create(C.byref(cfg), 0, C.byref(h))
    assert rc == -2 and not h.value          # PCLEAN_ERR_CUDA, no handle


def test_product_path_does_not_import_oracle():
    import subprocess, sys
    code = "import sys; import pclean_b200.engine, pclean_b200.lowering, pclean_b200.host_fixture.synth, pclean_b200.host_fixture.analysis; print('oracle' in sys.modules)"
    out = subprocess.check_output([sys.executable, "-c", code], cwd=ROOT).decode().strip()
    assert out == "False"
    for f in ("engine.cu", "device.cuh", "lower.hpp", "osa_bitpar.cuh"):
        src = open(os.path.join(ROOT, "pclean_b200", "csrc", f)).read()
        assert "oracle/" not in src.replace("oracle/pclean_oracle.cpp addtypos_score", "")
