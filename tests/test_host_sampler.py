"""CPU tests: the product's host-side csprng restatement (ringo-snark_b200/csrc/host_sampler.cpp, include/ringo_sampler.h)
against the oracle's (oracle/jindo_ref.py, itself a line-by-line restatement of math/csprng/*.go): same AES-CTR stream
with the XOR-accumulating refill, same ziggurat, same COSAC accept/reject decisions, and the tape builder's output equal
to the randomness trace of the oracle's Prover.Commit on the same seeds and crypto/rand bytes."""
import ctypes
import hashlib

import numpy as np
import pytest

import oracle
from oracle import jindo_ref as J
from oracle import pyref
import jindo_util as U


@pytest.fixture(scope="module")
def rs(product):
    product.load()
    return product


def k48(seed):
    return hashlib.sha384(seed).digest()


def test_uniform_sampler_stream(rs):
    L = rs.lib()
    seed = b"uniform-seed-0123456789abcdef...."
    ref = J.UniformSampler(seed)
    u = L.rs_uniform_create(ctypes.c_char_p(k48(seed)))
    try:
        # across several 8 KiB refills (the buffer XOR-accumulates, uniform.go:64-70)
        n = 5000
        out = np.zeros(n, dtype=np.uint64)
        L.rs_uniform_sample_many(u, out.ctypes.data_as(ctypes.c_void_p), n)
        assert np.array_equal(out, ref.sample_many(n))
        assert [L.rs_uniform_sample_float(u) for _ in range(300)] == [ref.sample_float() for _ in range(300)]
        for bound in (3, 1 << 40, (1 << 63) + 12345, 60272):
            assert [L.rs_uniform_sample_n(u, bound) for _ in range(50)] == [ref.sample_n(bound) for _ in range(50)]
        assert L.rs_uniform_sample(u) == ref.sample()
    finally:
        L.rs_uniform_destroy(u)
    # Read(): plain keystream into zeroed caller bytes, odd lengths (cipher.Stream keeps the unused block bytes)
    ref = J.UniformSampler(seed)
    u = L.rs_uniform_create(ctypes.c_char_p(k48(seed)))
    try:
        for n in (1, 15, 16, 17, 31, 100, 4097):
            buf = ctypes.create_string_buffer(n)
            L.rs_uniform_read(u, buf, n)
            assert buf.raw == ref.read(n), n
    finally:
        L.rs_uniform_destroy(u)


def test_rounded_gaussian_sampler(rs):
    L = rs.lib()
    seed = b"rounded-seed"
    ref = J.RoundedGaussianSampler(seed)
    r = L.rs_rounded_create(ctypes.c_char_p(k48(seed)))
    try:
        # 40 000 draws reach the ziggurat's wedge and tail branches many times
        got = [L.rs_rounded_norm_float(r) for _ in range(40000)]
        exp = [ref.norm_float() for _ in range(40000)]
        assert got == exp
        assert max(abs(x) for x in got) > 3.442619855899          # the i == 0 tail branch was taken
        for sigma in (0.8, 3.19, 1234.5, 7.0e9):
            assert [L.rs_rounded_sample(r, 0.0, sigma) for _ in range(500)] == [ref.sample(0, sigma) for _ in range(500)]
    finally:
        L.rs_rounded_destroy(r)


def test_cosac_sampler(rs):
    L = rs.lib()
    sb, sr = b"cosac-base", b"cosac-rounded"
    ref = J.COSACSampler(sb, sr)
    c = L.rs_cosac_create(ctypes.c_char_p(k48(sb)), ctypes.c_char_p(k48(sr)))
    rng = np.random.default_rng(5)
    try:
        for sigma in (0.6, 2.5, 41.0, 1.0e4, 3.3e7):
            centres = np.concatenate([rng.uniform(-50, 50, 300), rng.uniform(-1e9, 1e9, 100), [0.0, 0.5, -0.5, 1e-12, 123456.499999]])
            got = [L.rs_cosac_sample(c, float(x), sigma) for x in centres]
            exp = [ref.sample(float(x), sigma) for x in centres]
            assert got == exp, sigma
    finally:
        L.rs_cosac_destroy(c)


@pytest.mark.parametrize("sigma", [0.7978845608, 3.19, 12.77, 33.0])
def test_compute_cdt(rs, sigma):
    for centre in (0.0, 1 / 128, 0.5, 127 / 128):
        tab, lo = rs.jindo.compute_cdt(centre, sigma)
        assert [int(x) for x in tab] == J.compute_cdt(centre, sigma)
        assert lo == -int(np.ceil(9 * sigma))


@pytest.mark.parametrize("name,N,batch,length", [("zp255", 1024, 1, 1024), ("zp255", 1024, 1, 700), ("zp128", 512, 2, 512),
                                                 ("zp240", 1024, 8, 1000), ("zp255", 1024, 1, 1)])
def test_tape_builder_equals_oracle_commit_trace(rs, name, N, batch, length):
    """One worker, the oracle prover's sampler seeds and crypto/rand bytes: every host entry of the tape equals what the
    restated Prover.Commit consumed (first / last / mask rows, COSAC integers in call order, the mask column's MLWE noise)."""
    q = pyref.modulus(name)
    fo = oracle.Field(q)
    p, tape, prv = U.setup(q, N, batch, seed=b"tape-builder-" + name.encode())
    pos0 = tape._pos                                             # crypto/rand position after NewProver
    rng = np.random.default_rng(N + length)
    v = [int.from_bytes(rng.bytes(40), "little") % q for _ in range(length)]
    prv.commit(v)
    rec = prv.trace
    exp = U.tape_arrays(fo, p, rec)
    # the same crypto/rand bytes for the product's builder
    clone = J.Tape(b"tape-builder-" + name.encode())
    clone.read(pos0)
    rand_bytes = clone.read(tape._pos - pos0)
    pd = U.params_dict(p, prv.ecd)
    noise = {"ecd_blind_std": p.ecd_blind_std, "mask_blind_std": p.mask_blind_std, "mask_std": p.mask_std, "mask_mlwe_std": p.mask_mlwe_std}
    seeds = [prv.ecd.cosac.base.seed, prv.ecd.cosac.rounded.base.seed, prv.rounded.base.seed]
    tb = rs.jindo.TapeBuilder(pd, len(prv.ecd.twin.tables[0]), noise, q, threads=1, seeds=seeds)
    tb.set_random_bytes(rand_bytes)
    tb.start(fo.to_mont(v), length)
    T, ms = tb.wait()
    got = tb.arrays(T)
    assert T.twin_u is None and T.mlwe_mask_only == 1 and T.on_device == 0
    assert np.array_equal(got["last_row"], exp["last_row"])
    assert np.array_equal(got["first_row"], exp["first_row"])
    assert np.array_equal(got["mask_rows"], exp["mask_rows"])
    assert T.n_cosac_polys * p.d == exp["cosac_g"].size
    assert np.array_equal(got["cosac_g"].reshape(-1), exp["cosac_g"])
    nm = p.in_msis + p.mlwe
    assert np.array_equal(got["mlwe_mask"], exp["mlwe"][p.cols * nm:])
    # a second Commit continues every stream where the first one stopped
    v2 = [int.from_bytes(rng.bytes(40), "little") % q for _ in range(length)]
    pos1 = tape._pos
    prv.commit(v2)
    exp2 = U.tape_arrays(fo, p, prv.trace)
    clone2 = J.Tape(b"tape-builder-" + name.encode()); clone2.read(pos1)
    tb.set_random_bytes(clone2.read(tape._pos - pos1))
    tb.start(fo.to_mont(v2), length, slot=1)
    T2, _ = tb.wait(1)
    got2 = tb.arrays(T2)
    assert np.array_equal(got2["cosac_g"].reshape(-1), exp2["cosac_g"]) and np.array_equal(got2["mask_rows"], exp2["mask_rows"])
    assert np.array_equal(got2["mlwe_mask"], exp2["mlwe"][p.cols * nm:])


def test_tape_builder_threads_produce_valid_tapes(rs):
    """Several workers with OS-seeded samplers (the SafeCopy model): shapes, ranges and the first-row relation hold, and the
    COSAC integers are centred where the encoder's float64 centres put them."""
    q = pyref.modulus("zp255")
    fo = oracle.Field(q)
    p, tape, prv = U.setup(q, 1024, 1)
    pd = U.params_dict(p, prv.ecd)
    noise = {"ecd_blind_std": p.ecd_blind_std, "mask_blind_std": p.mask_blind_std, "mask_std": p.mask_std, "mask_mlwe_std": p.mask_mlwe_std}
    tb = rs.jindo.TapeBuilder(pd, len(prv.ecd.twin.tables[0]), noise, q, threads=4)
    rng = np.random.default_rng(9)
    v = [int.from_bytes(rng.bytes(40), "little") % q for _ in range(1024)]
    vm = fo.to_mont(v)
    tb.start(vm, 1024, 0)
    tb.start(vm, 1024, 1)                      # both slots in flight
    (T0, _), (T1, _) = tb.wait(0), tb.wait(1)
    a, b = tb.arrays(T0), tb.arrays(T1)
    assert not np.array_equal(a["last_row"], b["last_row"]) and not np.array_equal(a["cosac_g"], b["cosac_g"])
    for t in (a, b):
        last = oracle.array_to_ints(t["last_row"]); first = oracle.array_to_ints(t["first_row"])
        assert all(x < q for x in last) and last[-1] == 0
        R = fo.R
        vals = [x * R % q for x in v]                # Montgomery representatives as integers
        assert first[0] == vals[0]
        cs = p.cols * p.slots
        assert all(first[i] == (vals[i] - last[i - 1]) % q for i in range(1, cs))
        assert all(x < q for x in oracle.array_to_ints(t["mask_rows"]))
        # data-column first rows: |g + centre| within 12 sigma of zero
        enc = prv.ecd
        for col in range(p.cols):
            canon = [x * pow(R, -1, q) % q for x in first[col * p.slots:(col + 1) * p.slots]]
            fp = enc.centre(enc.base_encode(canon)[0])
            g = t["cosac_g"][col]
            assert max(abs(int(gi) + f) for gi, f in zip(g, fp)) < 12 * p.ecd_blind_std + 1
        assert abs(t["mlwe_mask"]).max() < 12 * p.mask_mlwe_std + 1
