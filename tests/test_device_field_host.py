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


# ---- reduced-radix lazy arithmetic (rr_field.cuh) ------------------------------------------------
B = 29


def rr_limbs(x, W):
    return [(x >> (B * j)) & ((1 << B) - 1) for j in range(W - 1)] + [x >> (B * (W - 1))]


def rr_value(l):
    return sum(int(v) << (B * j) for j, v in enumerate(l))


def rr_params(q, W):
    M = 3 * q
    m = rr_limbs(M, W)
    c3 = [m[0] + (1 << B)] + [m[j] + (1 << B) - 1 for j in range(1, W - 1)] + [m[W - 1] - 1]
    assert rr_value(c3) == M
    return np.array(rr_limbs(q, W) + c3 + rr_limbs(2 * q, W) + rr_limbs((1 << (B * W)) % q, W), dtype=np.uint32)


@pytest.mark.parametrize("name,W", [("zp255", 9), ("zp240", 9), ("zp220", 8)])
def test_reduced_radix_arithmetic(hf, name, W):
    q = pyref.modulus(name)
    NL = 8
    Rp = 1 << (B * W)
    assert q % (1 << B) == 1
    P = ctypes.POINTER(ctypes.c_uint32)
    params = rr_params(q, W)
    rng = random.Random(29 + W)
    for QZ in ([0, 1] if q % (1 << (2 * B)) == 1 else [0]):
        def run(op, x, y, out_cols, n):
            x = np.ascontiguousarray(x, dtype=np.uint32); y = np.ascontiguousarray(y, dtype=np.uint32)
            z = np.zeros((n, out_cols), dtype=np.uint32)
            assert hf.hf_rr(NL, W, QZ, op, params.ctypes.data_as(P), x.ctypes.data_as(P), y.ctypes.data_as(P),
                            z.ctypes.data_as(P), ctypes.c_size_t(n)) == 0
            return z
        n = 2000
        vals = [0, 1, q - 1, q - 2, Rp % q, (q - 1) // 2] + [rng.randrange(q) for _ in range(n - 6)]
        words = to32(vals, NL)
        rr = run(1, words, words, W, n)
        assert [rr_value(r) for r in rr] == vals and all(max(r[:-1]) < (1 << B) for r in rr)
        assert from32(run(2, rr, rr, NL, n)) == vals
        # lazy multiplicand: limb-wise sum of several normalised values (what two lazy stages produce)
        bound = (Rp // q) - 1 if W * B - q.bit_length() < 8 else 60
        lazy, lazy_vals = [], []
        for _ in range(n):
            k = rng.randrange(1, 4)
            parts = [rng.randrange(q) for _ in range(k)] + [rng.randrange(min(bound - k, 40) * q)]
            limbs = [sum(rr_limbs(p, W)[j] for p in parts) for j in range(W)]
            extra = [rng.randrange(1 << 30) if rng.random() < 0.3 else 0 for _ in range(W - 1)] + [0]   # stray limb excess
            limbs = [a + e for a, e in zip(limbs, extra)]
            limbs = [min(v, int(2.5 * (1 << 30)) - 1) for v in limbs]
            lazy.append(limbs); lazy_vals.append(rr_value(limbs))
        bvals = [rng.randrange(q) for _ in range(n)]
        b = np.array([rr_limbs(v, W) for v in bvals], dtype=np.uint32)
        z = run(0, np.array(lazy, dtype=np.uint32), b, W, n)
        rinv = pow(Rp, -1, q)
        for zi, av, bv in zip(z, lazy_vals, bvals):
            if av * bv >= q * Rp:
                continue                       # outside the documented input range
            v = rr_value(zi)
            assert v < 2 * q and v % q == av * bv * rinv % q and max(zi[:-1]) < (1 << B)
        # canonicalisation of lazy values
        can = run(3, np.array(lazy, dtype=np.uint32), b, NL, n)
        assert from32(can) == [av % q for av in lazy_vals]
        # three chained lazy butterfly stages + canonicalisation vs Python mod arithmetic
        m = 500
        us = [rng.randrange(q) for _ in range(m)]; vs = [rng.randrange(q) for _ in range(m)]
        ws = [[rng.randrange(q) for _ in range(3)] for _ in range(m)]
        xin = np.zeros((m, 2 * NL), dtype=np.uint32)
        xin[:, :NL] = to32(us, NL); xin[:, NL:] = to32(vs, NL)
        tw = np.array([rr_limbs(w, W) for row in ws for w in row], dtype=np.uint32)
        out = run(4, xin, tw, 2 * NL, m)
        for i in range(m):
            u, v = us[i], vs[i]
            for st in range(3):
                t = v * ws[i][st] * rinv % q
                u, v = (u + t) % q, (u - t) % q
            assert from32(out[i:i + 1, :NL])[0] == u and from32(out[i:i + 1, NL:])[0] == v


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
