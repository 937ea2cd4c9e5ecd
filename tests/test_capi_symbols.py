"""CPU test: libringo_b200.so loads and exports every symbol include/ringo_cuda.h declares
(no compute calls: there is no GPU here).  Also checks the loud failure without a device."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    names = set()
    for fn in os.listdir(os.path.join(ROOT, "include")):
        if fn.endswith(".h"):
            src = open(os.path.join(ROOT, "include", fn)).read()
            src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
            names |= set(re.findall(r"\b(rs_[a-z0-9_]+)\s*\(", src))
    return sorted(names)


def test_library_exports_every_declared_symbol(product):
    L = product.load()
    syms = header_symbols()
    assert len(syms) >= 40
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/*.h but not exported"
    # the ctypes table binds exactly the declared set
    assert sorted(product.capi.declared_symbols()) == syms


def test_no_cpu_fallback(product):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    L = product.load()
    assert L.rs_init(0) == -5
    assert b"no CPU fallback" in L.rs_last_error()
