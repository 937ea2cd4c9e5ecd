"""CPU tests: the FP64-pipe field arithmetic (ringo-snark_b200/csrc/field52.cuh) and the lazy butterfly rounds
of ntt52.cuh, compiled for the host and run under round-towards-zero so that fma() is __fma_rz, against Python
integers.  Checks exactness of the limb-product split, the Montgomery product with radix 2^260, the lazy
reduction bounds and whole radix-8 / radix-4 / radix-2 rounds (forward, inverse, scaled, unit-twiddle variants)
including the value bounds the kernel's bookkeeping relies on."""
import ctypes
import os
import random
import subprocess

import numpy as np
import pytest

from oracle import pyref

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB = os.path.join(HERE, "_build", "libhostfield52.so")
ND, NL = 5, 8
RP = 1 << (52 * ND)
M52 = (1 << 52) - 1
PLL = ctypes.POINTER(ctypes.c_longlong)
PU32 = ctypes.POINTER(ctypes.c_uint32)
PD = ctypes.POINTER(ctypes.c_double)


@pytest.fixture(scope="module")
def lib():
    src = os.path.join(HERE, "host_field52_test.cpp")
    deps = [src] + [os.path.join(ROOT, "ringo-snark_b200", "csrc", h) for h in ("field52.cuh", "ntt52.cuh", "ntt.cuh", "field.cuh")]
    if (not os.path.exists(LIB)) or os.path.getmtime(LIB) < max(os.path.getmtime(d) for d in deps):
        os.makedirs(os.path.dirname(LIB), exist_ok=True)
        subprocess.check_call(["/usr/bin/g++", "-O1", "-std=c++17", "-frounding-math", "-ffp-contract=off", "-fPIC", "-shared",
                               "-o", LIB, src])
    return ctypes.CDLL(LIB)


def limbs(x):
    """normalised 52-bit limbs; the top limb takes everything above 2^208"""
    return [(x >> (52 * k)) & M52 for k in range(ND - 1)] + [x >> (52 * (ND - 1))]


def value(l):
    return sum(int(v) << (52 * k) for k, v in enumerate(l))


def words(x):
    return [(x >> (32 * k)) & 0xFFFFFFFF for k in range(NL)]


def setup(lib, q, N=1 << 16):
    ninv = pow(N, -1, q) * RP % q
    w1 = 7 * RP % q
    info = (ctypes.c_int * 2)()
    qa = np.array(words(q), dtype=np.uint32)
    a = np.array(limbs(ninv), dtype=np.int64); b = np.array(limbs(ninv * 7 % q), dtype=np.int64)
    rc = lib.hf52_setup(qa.ctypes.data_as(PU32), a.ctypes.data_as(PLL), b.ctypes.data_as(PLL), info)
    assert rc == 0
    return info[0], info[1], ninv, ninv * 7 % q


FIELDS = ["zp255", "zp240", "zp220"]


@pytest.mark.parametrize("name", FIELDS)
def test_setup_and_words(lib, name):
    q = pyref.modulus(name)
    q0one, max_lazy, _, _ = setup(lib, q)
    assert q0one == (1 if q % (1 << 52) == 1 else 0)
    assert max_lazy == min(RP // q, 1 << 20) and max_lazy >= 36
    rng = random.Random(52)
    vals = [0, 1, q - 1, q, 2 * q - 1, (1 << 256) - 1, M52, M52 + 1] + [rng.randrange(1 << 256) for _ in range(2000)]
    w = np.array([words(v) for v in vals], dtype=np.uint32)
    l = np.zeros((len(vals), ND), dtype=np.int64)
    lib.hf52_from_words(w.ctypes.data_as(PU32), l.ctypes.data_as(PLL), ctypes.c_size_t(len(vals)))
    assert [list(map(int, r)) for r in l] == [limbs(v) for v in vals]
    w2 = np.zeros_like(w)
    lib.hf52_to_words(l.ctypes.data_as(PLL), w2.ctypes.data_as(PU32), ctypes.c_size_t(len(vals)))
    assert (w2 == w).all()


@pytest.mark.parametrize("name", FIELDS)
def test_mont52_exact(lib, name):
    q = pyref.modulus(name)
    q0one, _, _, _ = setup(lib, q)
    rinv = pow(RP, -1, q)
    rng = random.Random(104)
    edge_a = [0, 1, q - 1, q, 2 * q, RP - 1, RP - q, M52, (M52 << 208) | (M52 << 104), sum(M52 << (52 * k) for k in range(ND))]
    edge_b = [0, 1, q - 1, q - 2, RP % q, (q - 1) // 2, M52, M52 << 52]
    a_vals = [a for a in edge_a for _ in edge_b] + [rng.randrange(RP) for _ in range(4000)] + [rng.randrange(q) for _ in range(1000)]
    b_vals = [b % q for _ in edge_a for b in edge_b] + [rng.randrange(q) for _ in range(5000)]
    a = np.array([limbs(v) for v in a_vals], dtype=np.int64); b = np.array([limbs(v) for v in b_vals], dtype=np.int64)
    for generic in ([0, 1] if q0one else [1]):
        r = np.zeros_like(a)
        lib.hf52_mont(generic, a.ctypes.data_as(PLL), b.ctypes.data_as(PLL), r.ctypes.data_as(PLL), ctypes.c_size_t(len(a_vals)))
        for row, av, bv in zip(r, a_vals, b_vals):
            v = value(row)
            assert all(0 <= int(x) < (1 << 58) for x in row)
            assert v % q == av * bv * rinv % q and v < q + (av * bv >> (52 * ND)) + 1 and v < 2 * q, (name, generic)


@pytest.mark.parametrize("name", FIELDS)
def test_reduce_bounds(lib, name):
    q = pyref.modulus(name)
    _, max_lazy, _, _ = setup(lib, q)
    rng = random.Random(7)
    top = min(max_lazy, 64)
    vals = [0, 1, q - 1, q, q + 1, 2 * q - 1, 2 * q, 3 * q - 1, top * q - 1] + [k * q + d for k in range(top) for d in (-1, 0, 1) if k * q + d >= 0]
    vals += [rng.randrange(top * q) for _ in range(6000)]
    rows = []
    for v in vals:
        l = limbs(v)
        # spread the value over unnormalised signed limbs the way lazy sums and differences do
        for k in range(ND - 1):
            d = rng.randrange(-(1 << 9), 1 << 9) if rng.random() < 0.7 else 0
            l[k] += d << 52
            l[k + 1] -= d
        assert value(l) == v
        rows.append(l)
    arr = np.array(rows, dtype=np.int64)
    lib.hf52_reduce(arr.ctypes.data_as(PLL), ctypes.c_size_t(len(vals)))
    worst = 0
    for row, v in zip(arr, vals):
        r = value(row)
        assert all(0 <= int(x) <= M52 for x in row[:-1]) and int(row[-1]) >= 0
        assert r % q == v % q and 0 <= r < 2 * q + (q >> 13), (name, v // q, r // q)
        worst = max(worst, r)
    # finish: canonical words
    w = np.zeros((len(vals), NL), dtype=np.uint32)
    arr = np.array(rows, dtype=np.int64)
    lib.hf52_finish(arr.ctypes.data_as(PLL), w.ctypes.data_as(PU32), 1, ctypes.c_size_t(len(vals)))
    assert [sum(int(x) << (32 * k) for k, x in enumerate(r)) for r in w] == [v % q for v in vals]


def run_round(lib, RR, inv, scale, triv, xs, tw, s_top, hi):
    arr = np.array(xs, dtype=np.int64)
    t = np.ascontiguousarray(tw, dtype=np.float64)
    rc = lib.hf52_round(RR, inv, scale, triv, arr.ctypes.data_as(PLL), t.ctypes.data_as(PD), s_top, ctypes.c_uint(hi),
                        ctypes.c_size_t(arr.shape[0]))
    assert rc == 0
    return arr


def ref_round(q, RR, inv, scale, x, tw_vals, s_top, hi, ninv, ninv_w1):
    """radix_round of ntt.cuh on Python integers mod q; twiddles are plain field elements here (x * w mod q)"""
    x = list(x)
    for bb in range(RR):
        beta = bb if inv else RR - 1 - bb
        s = s_top + (RR - 1 - beta)
        tbase = (1 << s) + (hi << (RR - 1 - beta))
        for g in range(1 << (RR - 1 - beta)):
            w = tw_vals[tbase + g]
            for k in range(1 << beta):
                u0 = (g << (beta + 1)) + k; u1 = u0 + (1 << beta)
                if not inv:
                    v = x[u1] * w % q
                    x[u0], x[u1] = (x[u0] + v) % q, (x[u0] - v) % q
                else:
                    d = (x[u0] - x[u1]) % q
                    x[u0] = (x[u0] + x[u1]) % q
                    if scale and beta == RR - 1:
                        x[u0] = x[u0] * ninv % q; x[u1] = d * ninv_w1 % q
                    else:
                        x[u1] = d * w % q
    return x


@pytest.mark.parametrize("name", FIELDS)
@pytest.mark.parametrize("RR,inv,scale,triv", [(3, 0, 0, 0), (3, 0, 0, 1), (3, 1, 0, 0), (3, 1, 1, 0), (3, 1, 1, 1),
                                               (2, 0, 0, 0), (2, 1, 0, 0), (2, 1, 1, 0), (1, 0, 0, 0), (1, 1, 0, 0), (1, 1, 1, 0),
                                               (2, 0, 0, 1), (2, 1, 1, 1)])
def test_rounds_against_integers(lib, name, RR, inv, scale, triv):
    q = pyref.modulus(name)
    _, max_lazy, ninv_m, ninv_w1_m = setup(lib, q)
    rinv = pow(RP, -1, q)
    rng = random.Random(RR * 100 + inv * 10 + scale + 3 * triv)
    E = 1 << RR
    # a twiddle table in R' Montgomery form (TW52 stride 6 doubles); unit twiddles where the TRIV variant assumes them
    s_top, hi = (0, 0) if triv else (4, 5)
    T = 1 << (s_top + RR + 1)
    tw_vals = [rng.randrange(q) for _ in range(T)]
    if triv:
        for s in range(RR):
            tw_vals[1 << s] = 1
    tw = np.zeros((T, 6), dtype=np.float64)
    for i, w in enumerate(tw_vals):
        tw[i, :ND] = limbs(w * RP % q)
    n = 300
    groups, vals = [], []
    for i in range(n):
        # inputs below 2.0001 q like a freshly loaded tile, some at the extremes
        ev = [rng.choice([0, q - 1, q, 2 * q - 1, 2 * q + (q >> 14)]) if rng.random() < 0.2 else rng.randrange(2 * q) for _ in range(E)]
        vals.append(ev)
        groups.append([limbs(v) for v in ev])
    out = run_round(lib, RR, inv, scale, triv, groups, tw, s_top, hi)
    ninv = ninv_m * rinv % q; ninv_w1 = ninv_w1_m * rinv % q
    worst = 0
    for i in range(n):
        exp = ref_round(q, RR, inv, scale, vals[i], tw_vals, s_top, hi, ninv, ninv_w1)
        for u in range(E):
            got = value(out[i, u])
            assert got % q == exp[u], (name, RR, inv, scale, triv, i, u)
            assert 0 <= got < min(max_lazy, 36) * q
            worst = max(worst, got // q)
            assert all(abs(int(l)) < (1 << 62) for l in out[i, u])
    # the bounds ntt52.cuh documents
    if inv:
        assert worst <= 2
    elif triv:
        assert worst <= 18
    else:
        assert worst <= 2 + 2 * RR


@pytest.mark.parametrize("name", ["zp255"])
def test_forward_pass_growth(lib, name):
    """11 forward stages in a row on lazy values (the longest pass), then the store path: canonical result."""
    q = pyref.modulus(name)
    _, max_lazy, _, _ = setup(lib, q)
    rng = random.Random(11)
    n = 200
    tw_vals = [rng.randrange(q) for _ in range(1 << 12)]
    tw = np.zeros((len(tw_vals), 6), dtype=np.float64)
    for i, w in enumerate(tw_vals):
        tw[i, :ND] = limbs(w * RP % q)
    vals = [[rng.randrange(2 * q) for _ in range(8)] for _ in range(n)]
    cur = np.array([[limbs(v) for v in ev] for ev in vals], dtype=np.int64)
    exp = [list(ev) for ev in vals]
    stages = 0
    for rnd, RR in enumerate([3, 3, 3, 2]):
        cur = run_round(lib, RR, 0, 0, 0, cur if RR == 3 else cur.reshape(-1, 4, ND), tw, 2, 1).reshape(n, 8, ND)
        if RR == 3:
            exp = [ref_round(q, 3, 0, 0, e, tw_vals, 2, 1, 0, 0) for e in exp]
        else:
            exp = [ref_round(q, 2, 0, 0, e[:4], tw_vals, 2, 1, 0, 0) + ref_round(q, 2, 0, 0, e[4:], tw_vals, 2, 1, 0, 0) for e in exp]
        stages += RR
    assert stages == 11
    worst = max(value(cur[i, u]) // q for i in range(n) for u in range(8))
    assert worst < 24 and worst < max_lazy
    w = np.zeros((n * 8, NL), dtype=np.uint32)
    flat = np.ascontiguousarray(cur.reshape(n * 8, ND))
    lib.hf52_finish(flat.ctypes.data_as(PLL), w.ctypes.data_as(PU32), 1, ctypes.c_size_t(n * 8))
    got = [sum(int(x) << (32 * k) for k, x in enumerate(r)) for r in w]
    assert got == [v for e in exp for v in e]
