"""CPU tests: the Shoup butterfly product and the predicated lazy add / sub of the big-field NTT
(ringo-snark_b200/csrc/shoup.cuh) compiled for the host with an emulated carry flag, against Python integers.
This checks the exact multiply-add sequence the sm_100a kernels execute, the truncated-quotient error bound
(including products engineered onto its slow path) and the lazy ranges the NTT relies on."""
import ctypes
import os
import random
import subprocess

import numpy as np
import pytest

from oracle import pyref

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB = os.path.join(HERE, "_build", "libhostshoup.so")
P32 = ctypes.POINTER(ctypes.c_uint32)


@pytest.fixture(scope="module")
def hs():
    src = os.path.join(HERE, "host_shoup_test.cpp")
    hdrs = [os.path.join(ROOT, "ringo-snark_b200", "csrc", h) for h in ("field.cuh", "shoup.cuh")]
    if (not os.path.exists(LIB)) or os.path.getmtime(LIB) < max(os.path.getmtime(p) for p in [src] + hdrs):
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


def qs_of(q):
    return 2 if q % (1 << 64) == 1 else (1 if q % (1 << 32) == 1 else 0)


def call(hs, NL, qs, op, q, xs, ys, ywords):
    qa = to32([q], NL)[0]
    q2 = to32([2 * q], NL)[0]
    x = to32(xs, NL)
    y = to32(ys, ywords)
    z = np.zeros_like(x)
    fr = np.zeros(len(xs), dtype=np.uint32)
    rc = hs.hs_op(NL, qs, op, qa.ctypes.data_as(P32), q2.ctypes.data_as(P32), x.ctypes.data_as(P32), y.ctypes.data_as(P32),
                  z.ctypes.data_as(P32), fr.ctypes.data_as(P32), ctypes.c_size_t(len(xs)))
    assert rc == 0
    return from32(z), [int(v) for v in fr]


# fields the Shoup path serves: 3.6 q < beta (all of them are q == 1 mod 2^32 at least)
FIELDS = [("zp255", 8), ("zp240", 8), ("zp220", 8), ("zp110", 4)]


@pytest.mark.parametrize("name,NL", FIELDS)
def test_shoup_product_matches_integers(hs, name, NL):
    q = pyref.modulus(name)
    beta = 1 << (32 * NL)
    assert 36 * q < 10 * beta
    rng = random.Random(11)
    delta = 1 << (32 * (NL - 2))
    xs, ws = [], []
    edge = [0, 1, q - 1, q, 2 * q - 1, 2 * q, 2 * q + delta, beta - 1]
    for x in edge:
        for w in (0, 1, 2, q - 1, q // 2, rng.randrange(q)):
            xs.append(x); ws.append(w)
    for _ in range(3000):
        xs.append(rng.randrange(2 * q + delta)); ws.append(rng.randrange(q))
    packed = [w | (((w * beta) // q) << (32 * NL)) for w in ws]
    for qs in range(qs_of(q) + 1):
        z, _ = call(hs, NL, qs, 0, q, xs, packed, 2 * NL)
        _, fr = call(hs, NL, qs, 4, q, xs, packed, 2 * NL)
        for x, w, got, f in zip(xs, ws, z, fr):
            h = (x * ((w * beta) // q)) // beta
            v = x * w - h * q
            if f <= (1 << 32) - 1 - NL:
                assert got == v                   # the exact quotient: the textbook Shoup value
            else:
                assert got == (v - q if v >= q else v)   # slow path (x = beta - 1 with tiny w lands there)
            assert got % q == (x * w) % q and got * beta < (beta + x) * q + beta


@pytest.mark.parametrize("name,NL", FIELDS)
def test_shoup_slow_path(hs, name, NL):
    """Products whose truncated high half has its fractional limb within NL of 2^32 - 1 (engineered: x odd and
    wp = target / x mod beta, target = x wp mod beta), where the truncated quotient may be one short: the contract is
    z == x w - h~ q (- q if that is >= q) with h~ in {h - 1, h}."""
    q = pyref.modulus(name)
    beta = 1 << (32 * NL)
    rng = random.Random(5)
    xs, packed, meta = [], [], []
    for t in range(400):
        x = rng.randrange(2 * q) | 1
        if t % 2:      # x wp just above a multiple of beta: the dropped low columns pull the kept sum below it
            target = rng.randrange((NL - 2) << (32 * (NL - 1)))
        else:          # fractional limb straddling the threshold from below
            target = (((1 << 32) - 1 - rng.randrange(NL + 4)) << (32 * (NL - 1))) | rng.randrange(1 << (32 * (NL - 1)))
        wp = (target * pow(x, -1, beta)) % beta
        w = rng.randrange(q)
        xs.append(x); packed.append(w | (wp << (32 * NL))); meta.append((w, wp))
    qs = qs_of(q)
    z, _ = call(hs, NL, qs, 0, q, xs, packed, 2 * NL)
    h_tr, fr = call(hs, NL, qs, 4, q, xs, packed, 2 * NL)
    short = 0
    for x, (w, wp), got, ht, f in zip(xs, meta, z, h_tr, fr):
        h = (x * wp) // beta
        assert ht in (h, h - 1)
        if f <= (1 << 32) - 1 - NL:
            assert ht == h                                            # fast path: the quotient is exact
            want = (x * w - h * q) % beta
        else:
            want = (x * w - ht * q) % beta
            if want >= q:
                want -= q
        short += ht == h - 1
        assert got == want
    assert short > 0                                                  # the one-short case really occurred


@pytest.mark.parametrize("name,NL", FIELDS)
def test_tri_hi_error_bound(hs, name, NL):
    q = pyref.modulus(name)
    beta = 1 << (32 * NL)
    rng = random.Random(3)
    xs = [rng.randrange(beta) for _ in range(2000)] + [beta - 1] * 4
    wps = [rng.randrange(beta) for _ in range(2000)] + [beta - 1, beta - 2, 1, 0]
    h_tr, fr = call(hs, NL, 0, 4, q, xs, [wp << (32 * NL) for wp in wps], 2 * NL)
    for x, wp, ht, f in zip(xs, wps, h_tr, fr):
        S = x * wp
        SK = sum(((x >> (32 * j)) & 0xFFFFFFFF) * ((wp >> (32 * i)) & 0xFFFFFFFF) << (32 * (i + j))
                 for i in range(NL) for j in range(NL) if i + j >= NL - 2)
        assert ht == SK >> (32 * NL) and f == (SK >> (32 * (NL - 1))) & 0xFFFFFFFF
        assert 0 <= S - SK < (NL - 2) << (32 * (NL - 1)) | 1 << 40


@pytest.mark.parametrize("name,NL", FIELDS)
def test_lazy_add_sub_ranges(hs, name, NL):
    q = pyref.modulus(name)
    beta = 1 << (32 * NL)
    delta = 1 << (32 * (NL - 2))
    rng = random.Random(9)
    B = 2 * q + delta
    edges = [0, 1, q, 2 * q - 1, 2 * q, 2 * q + 1, 2 * q + delta - 2, B - 1]
    # forward add: x in [0, B), y a product below 1.56 q; the sum never wraps
    xs = [rng.randrange(B) for _ in range(2000)] + edges * len(edges)
    ys = [rng.randrange(156 * q // 100) for _ in range(2000)] + [min(e, 156 * q // 100) for e in edges for _ in edges]
    z, _ = call(hs, NL, 0, 1, q, xs, ys, NL)
    for x, y, got in zip(xs, ys, z):
        assert got % q == (x + y) % q and got in (x + y, x + y - 2 * q)
        assert got < 2 * q + delta
        if x + y >= 2 * q + delta:
            assert got == x + y - 2 * q
    # sub: the subtrahend is at most 2q (forward: a product; inverse: an exactly reduced value)
    ys = [rng.randrange(2 * q + 1) for _ in range(2000)] + [min(e, 2 * q) for e in edges for _ in edges]
    z, _ = call(hs, NL, 0, 3, q, xs, ys, NL)
    for x, y, got in zip(xs, ys, z):
        assert got == (x - y if x >= y else x - y + 2 * q) and 0 <= got < B
