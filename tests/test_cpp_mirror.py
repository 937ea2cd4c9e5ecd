"""The C++ host mirror (include/ringo.hpp): compiles, links against the C-ABI library, and (GPU) gives the
oracle's results through the reference's operator names.  The test program links the oracle C library as
the checker (tests/cpp/mirror_parity.cpp)."""
import os
import subprocess

import pytest

from oracle import pyref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "ringo-snark_b200")
BIN = os.path.join(ROOT, "tests", "cpp", "_build", "mirror_parity")


def _build():
    import oracle  # noqa: F401  (builds oracle/_build/liboracle.so if missing)
    lib = os.path.join(PKG, "libringo_b200.so")
    if not os.path.exists(lib):
        pytest.fail("libringo_b200.so not built: run python ringo-snark_b200/build.py")
    os.makedirs(os.path.dirname(BIN), exist_ok=True)
    src = os.path.join(ROOT, "tests", "cpp", "mirror_parity.cpp")
    deps = [src, os.path.join(ROOT, "include", "ringo.hpp"), lib]
    if os.path.exists(BIN) and all(os.path.getmtime(BIN) >= os.path.getmtime(d) for d in deps):
        return
    obuild = os.path.join(ROOT, "oracle", "_build")
    cmd = ["/usr/bin/g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "oracle"),
           src, "-o", BIN, "-L", PKG, "-lringo_b200", "-L", obuild, "-loracle",
           f"-Wl,-rpath,{PKG}", f"-Wl,-rpath,{obuild}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_mirror_compiles_and_links():
    _build()
    r = subprocess.run([BIN, "--link"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stdout + r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["zp128", "zp255", "zp440"])
def test_mirror_parity_on_gpu(name):
    _build()
    q = pyref.modulus(name)
    L = (q.bit_length() + 63) // 64
    limbs = [hex((q >> (64 * i)) & ((1 << 64) - 1)) for i in range(L)]
    r = subprocess.run([BIN, str(L)] + limbs, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stdout + r.stderr
