"""CPU tests: the device field arithmetic (ringo-snark_b200/csrc/field.cuh) compiled for the
host with an emulated carry flag, against Python integers.  This checks the exact mad/carry
instruction sequence the sm_100a kernels execute (all three modulus-shape variants)."""
import ctypes
import os
import random
import subprocess

import numpy as np
import pytest

from oracle import pyref

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB = os.path.join(HERE, "_build", "libhostfield.so")


@pytest.fixture(scope="module")
def hf():
    src = os.path.join(HERE, "host_field_test.cpp")
    hdr = os.path.join(ROOT, "ringo-snark_b200", "csrc", "field.cuh")
    if (not os.path.exists(LIB)) or os.path.getmtime(LIB) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        os.makedirs(os.path.dirname(LIB), exist_ok=True)
        subprocess.check_call(["/usr/bin/g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-o", LIB, src])
    return ctypes.CDLL(LIB)


def to32(xs, NL):
    out = np.zeros((len(xs), NL), dtype=np.uint32)
    for i, x in enumerate(xs):
        for k in range(NL):
            out[i, k] = (x >> (32 * k)) & 0xFFFFFFFF
    return out


def from32(a):
    return [sum(int(v) << (32 * k) for k, v in enumerate(row)) for row in a]


def qs_variants(q):
    v = [0]
    if q % (1 << 32) == 1:
        v.append(1)
    if q % (1 << 64) == 1:
        v.append(2)
    return v


def call(hf, NL, qs, op, q, xs, ys):
    qa = to32([q], NL)[0]
    qinv = (-pow(q, -1, 1 << 32)) % (1 << 32)
    x = to32(xs, NL); y = to32(ys, NL); z = np.zeros_like(x)
    P = ctypes.POINTER(ctypes.c_uint32)
    rc = hf.hf_op(NL, qs, op, qa.ctypes.data_as(P), ctypes.c_uint32(qinv), x.ctypes.data_as(P),
                  y.ctypes.data_as(P), z.ctypes.data_as(P), ctypes.c_size_t(len(xs)))
    assert rc == 0
    return from32(z)


EXTRA = {
    # generic odd moduli (not of jindo shape) exercise QS = 0 with qinv != -1 and no spare bit
    "odd64": (1 << 64) - 59,
    "odd128_full": (1 << 128) - 159,
    "odd256_full": (1 << 256) - 189,
    "p2_31": 12640 ** 4 * 2 ** 3 * 0 + (3 * 2 ** 31 + 1) * 2 ** 64 + 2 ** 31 + 1,  # q == 1 mod 2^31 only
}


@pytest.mark.parametrize("name", sorted(pyref.MODULI) + sorted(EXTRA))
def test_device_field_sequence(hf, name):
    q = pyref.modulus(name) if name in pyref.MODULI else EXTRA[name] | 1
    NL = 2 * ((q.bit_length() + 63) // 64)
    R = 1 << (32 * NL)
    rinv = pow(R, -1, q)
    rng = random.Random(hash(name) & 0xFFFF)
    edge = [0, 1, 2, q - 1, q - 2, R % q, (R * R) % q, (q - 1) // 2, (q + 1) // 2,
            (1 << 32) - 1, 1 << 32, (1 << 64) - 1, (1 << (32 * NL - 1)) % q, q >> 1, (q >> 32) << 32]
    edge = [e % q for e in edge]
    xs = [a for a in edge for _ in edge] + [rng.randrange(q) for _ in range(3000)]
    ys = [b for _ in edge for b in edge] + [rng.randrange(q) for _ in range(3000)]
    for qs in qs_variants(q):
        assert call(hf, NL, qs, 0, q, xs, ys) == [(a * b * rinv) % q for a, b in zip(xs, ys)], (name, qs)
    assert call(hf, NL, 0, 1, q, xs, ys) == [(a + b) % q for a, b in zip(xs, ys)]
    assert call(hf, NL, 0, 2, q, xs, ys) == [(a - b) % q for a, b in zip(xs, ys)]
    assert call(hf, NL, 0, 3, q, xs, ys) == [(-a) % q for a in xs]


@pytest.mark.parametrize("name", ["zp110", "zp220", "zp240", "zp255"])
def test_lazy_2q_arithmetic(hf, name):
    """[0, 2q) variants: no overflow word, no final subtraction (moduli with 3q < 2^(32 NL))."""
    q = pyref.modulus(name)
    NL = 2 * ((q.bit_length() + 63) // 64)
    R = 1 << (32 * NL)
    assert 3 * q * (1 << 32) + 2 * q < R << 32
    rinv = pow(R, -1, q)
    rng = random.Random(77)
    P = ctypes.POINTER(ctypes.c_uint32)
    qa = to32([q], NL)[0]; q2a = to32([2 * q], NL)[0]
    qinv = (-pow(q, -1, 1 << 32)) % (1 << 32)

    def run(qs, op, xs, ys):
        x = to32(xs, NL); y = to32(ys, NL); z = np.zeros_like(x)
        assert hf.hf_lazy(NL, qs, op, qa.ctypes.data_as(P), ctypes.c_uint32(qinv), q2a.ctypes.data_as(P),
                          x.ctypes.data_as(P), y.ctypes.data_as(P), z.ctypes.data_as(P), ctypes.c_size_t(len(xs))) == 0
        return from32(z)
    edge2 = [0, 1, q - 1, q, q + 1, 2 * q - 1, 2 * q - 2, R % q, (1 << 32) - 1]
    edge1 = [0, 1, q - 1, q - 2, R % q, (q - 1) // 2]
    xs = [a for a in edge2 for _ in edge1] + [rng.randrange(2 * q) for _ in range(3000)]
    ys = [b for _ in edge2 for b in edge1] + [rng.randrange(q) for _ in range(3000)]
    for qs in qs_variants(q):
        got = run(qs, 0, xs, ys)
        assert all(g < 2 * q and g % q == a * b * rinv % q for g, a, b in zip(got, xs, ys)), (name, qs)
    ys2 = [rng.randrange(2 * q) for _ in xs]
    assert run(0, 1, xs, ys2) == [(a + b) % (2 * q) for a, b in zip(xs, ys2)]
    assert run(0, 2, xs, ys2) == [(a - b) % (2 * q) for a, b in zip(xs, ys2)]
    assert run(0, 3, xs, ys2) == [a % q for a in xs]
