"""The C-ABI library loads on a CPU-only box and exports every symbol include/yue_b200.h
declares (no compute calls here)."""
import ctypes
import os
import re

import pytest

from yue_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "yue_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(yue_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    syms = declared_symbols()
    assert len(syms) >= 25
    assert sorted(_lib.SIGNATURES) == syms


def test_library_exports_every_declared_symbol():
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for s in declared_symbols():
        assert hasattr(lib, s), s
    lib.yue_version.restype = ctypes.c_char_p
    assert b"sm_100a" in lib.yue_version()


def test_no_cpu_fallback_without_device():
    """On a box without a GPU the product path must fail loudly, not compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from yue_b200.engine import Engine, YueError
    with pytest.raises(YueError):
        Engine(0)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "yue_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(d, f), encoding="utf8").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), os.path.join(d, f)
