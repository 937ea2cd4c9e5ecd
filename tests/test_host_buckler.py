"""CPU tests of the host code the compiled Buckler prover / verifier is made of (csrc/buckler_host.cpp, the scalar helpers of
csrc/capi.cu, the mask helpers of csrc/host_sampler.cpp): no GPU involved.

  rs_sha256 / rs_shake128        FIPS 180-4 / FIPS 202 against hashlib, all block-boundary lengths (the SHA-NI and the portable path)
  rs_field_set_bytes / marshal   element.go:873-915 against Python integers, every shipped modulus
  rs_field_mul / add / sub       element.go:400-479 against Python integers
  rs_tape_builder_sum_check_mask prover.go:381-397: maskSum is the first random coefficient, the folded coefficients subtract"""
import ctypes
import hashlib
import random

import numpy as np
import pytest

from oracle import pyref

MODS = ["zp110", "zp128", "zp220", "zp240", "zp255", "zp440", "zp880"]


def test_sha256_and_shake128_match_hashlib(product):
    L = product.load()
    rng = random.Random(7)
    for n in [0, 1, 3, 55, 56, 57, 63, 64, 65, 119, 120, 127, 128, 129, 167, 168, 169, 335, 336, 337, 1000, 4095, 4096, 100003]:
        d = bytes(rng.getrandbits(8) for _ in range(n))
        o = ctypes.create_string_buffer(32)
        L.rs_sha256(d, n, o)
        assert o.raw == hashlib.sha256(d).digest(), n
        for m in [1, 16, 32, 167, 168, 169, 400]:
            o = ctypes.create_string_buffer(m)
            L.rs_shake128(d, n, o, m)
            assert o.raw == hashlib.shake_128(d).digest(m), (n, m)


def _limbs(x, Lw):
    return np.array([(x >> (64 * i)) & ((1 << 64) - 1) for i in range(Lw)], dtype=np.uint64)


def _int(a):
    return sum(int(v) << (64 * i) for i, v in enumerate(a))


@pytest.mark.parametrize("name", MODS)
def test_set_bytes_marshal_and_scalar_ops(product, name):
    rs = product
    L = rs.load()
    q = pyref.modulus(name)
    f = rs.Field(q)
    Lw = f.L
    R = 1 << (64 * Lw)
    u64p = rs.capi.c_u64p
    rng = random.Random(hash(name) & 0xffff)
    out = np.zeros(Lw, dtype=np.uint64)
    for n in [0, 1, 8, 31, 32, 33, 8 * Lw - 1, 8 * Lw, 8 * Lw + 1, 64, 128]:
        if n > 128:
            continue
        b = bytes(rng.getrandbits(8) for _ in range(n))
        assert L.rs_field_set_bytes(f.h, b, n, out.ctypes.data_as(u64p)) == 0
        want = int.from_bytes(b, "big") % q
        assert _int(out) == want * R % q, (name, n)                    # Montgomery form of the reduced value
        be = ctypes.create_string_buffer(8 * Lw)
        assert L.rs_field_marshal(f.h, out.ctypes.data_as(u64p), be) == 0
        assert be.raw == want.to_bytes(8 * Lw, "big")
    assert L.rs_field_set_bytes(f.h, b"\x00" * 129, 129, out.ctypes.data_as(u64p)) != 0      # longer than the reference ever passes
    for _ in range(20):
        a, b = rng.randrange(q), rng.randrange(q)
        am, bm = _limbs(a * R % q, Lw), _limbs(b * R % q, Lw)
        for fn, want in ((L.rs_field_mul, a * b % q), (L.rs_field_add, (a + b) % q), (L.rs_field_sub, (a - b) % q)):
            assert fn(f.h, out.ctypes.data_as(u64p), am.ctypes.data_as(u64p), bm.ctypes.data_as(u64p)) == 0
            assert _int(out) == want * R % q
    # edge values: 0, 1, q - 1, and aliasing out with an input
    for a, b in ((0, 0), (q - 1, q - 1), (q - 1, 1), (0, q - 1), (1, q - 1)):
        am, bm = _limbs(a * R % q, Lw), _limbs(b * R % q, Lw)
        assert L.rs_field_sub(f.h, am.ctypes.data_as(u64p), am.ctypes.data_as(u64p), bm.ctypes.data_as(u64p)) == 0
        assert _int(am) == (a - b) % q * R % q


def test_sum_check_mask_folds_like_the_reference(product):
    rs = product
    L = rs.load()
    from oracle import jindo_ref as J
    q = pyref.modulus("zp128")
    Lw = 2
    p = J.new_parameters(q, 1024, 1)
    import sys, os
    sys.path.insert(0, os.path.dirname(__file__))
    from jindo_util import params_dict
    ecd = J.Encoder(p, J.Tape(b"mask"))
    pd = params_dict(p, ecd)
    noise = {"ecd_blind_std": p.ecd_blind_std, "mask_blind_std": p.mask_blind_std, "mask_std": p.mask_std, "mask_mlwe_std": p.mask_mlwe_std}
    tb = rs.jindo.TapeBuilder(pd, len(ecd.twin.tables[0]), noise, q, threads=1)
    N, MR = 64, 200
    mask = np.zeros((MR, Lw), dtype=np.uint64); s = np.zeros(Lw, dtype=np.uint64)
    assert L.rs_tape_builder_sum_check_mask(tb.h, N, MR, mask.ctypes.data_as(rs.capi.c_vp), s.ctypes.data_as(rs.capi.c_vp)) == 0
    vals = [_int(r) for r in mask]
    assert all(v < q for v in vals) and len(set(vals)) > MR - 3
    # undo the folding: the first coefficient before the subtractions is maskSum; sum over i = 0 mod N of the final mask telescopes
    # back to it (what the verifier's sum-check equation relies on: prover.go:381-397)
    tot = 0
    for i in range(0, MR, N):
        tot = (tot + vals[i]) % q
    assert tot == _int(s)
    assert L.rs_tape_builder_sum_check_mask(tb.h, 0, MR, mask.ctypes.data_as(rs.capi.c_vp), s.ctypes.data_as(rs.capi.c_vp)) != 0
