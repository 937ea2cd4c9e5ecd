"""GPU parity tests (run with -m gpu on the B200): the CUDA path, called through the C ABI,
against the CPU oracle on the same seeded inputs -- bit-exact."""
import random

import numpy as np
import pytest

import oracle
from oracle import pyref

pytestmark = pytest.mark.gpu

NAMES = sorted(pyref.MODULI)


@pytest.fixture(scope="module")
def rs(product):
    L = product.load()
    assert L.rs_init(0) == 0, L.rs_last_error()
    return product


def rand_mont(f, n, seed):
    rng = np.random.default_rng(seed)
    q = f.q
    nb = (q.bit_length() + 7) // 8 + 8
    return f.to_mont([int.from_bytes(rng.bytes(nb), "little") % q for _ in range(n)])


def edge_polys(f, N):
    q = f.q
    z = f.to_mont([0] * N)
    top = f.to_mont([q - 1] * N)
    imp0 = f.to_mont([1] + [0] * (N - 1))
    impN = f.to_mont([0] * (N - 1) + [1])
    return [z, top, imp0, impN]


@pytest.mark.parametrize("name", NAMES)
def test_field_ops(rs, name):
    q = pyref.modulus(name)
    fo = oracle.Field(q)
    fg = rs.Field(q)
    n = 4096
    R = fo.R
    edge = [0, 1, 2, q - 1, q - 2, R % q, (R * R) % q, (q - 1) // 2, (q + 1) // 2, (1 << 64) - 1]
    xs = [a for a in edge for _ in edge]
    ys = [b for _ in edge for b in edge]
    a = np.concatenate([fo.to_mont(xs), rand_mont(fo, n - len(xs), 1)])
    b = np.concatenate([fo.to_mont(ys), rand_mont(fo, n - len(ys), 2)])
    c = rand_mont(fo, 1, 3)[0]
    # host scalar helpers of the ABI
    assert np.array_equal(fg.to_mont(oracle.ints_to_array(xs, fo.L)), fo.to_mont(xs))
    assert np.array_equal(fg.from_mont(fo.to_mont(xs)), oracle.ints_to_array([x % q for x in xs], fo.L))
    B = rs.Batch(fg, 4, n)
    B.upload(np.stack([a, b, a, b]))
    ev = rs.CyclicEvaluator.__new__(rs.CyclicEvaluator)
    rs.bigpoly._BaseOperator.__init__(ev, fg, n, None)
    P = [B.poly(i) for i in range(4)]
    ev.AddTo(P[2], P[0], P[1]); assert np.array_equal(P[2].Coeffs, fo.vec_add(a, b))
    ev.SubTo(P[2], P[0], P[1]); assert np.array_equal(P[2].Coeffs, fo.vec_sub(a, b))
    ev.NegTo(P[2], P[0]); assert np.array_equal(P[2].Coeffs, fo.vec_neg(a))
    ev.ScalarMulTo(P[2], P[0], c); assert np.array_equal(P[2].Coeffs, fo.vec_scalar_mul(a, c))
    for p in P:
        p.IsNTT = True
    ev.MulTo(P[2], P[0], P[1]); assert np.array_equal(P[2].Coeffs, fo.vec_mul(a, b))
    ev.MulAddTo(P[3], P[0], P[1]); assert np.array_equal(P[3].Coeffs, fo.vec_add(b, fo.vec_mul(a, b)))
    P[3].set(b)
    ev.MulSubTo(P[3], P[0], P[1]); assert np.array_equal(P[3].Coeffs, fo.vec_sub(b, fo.vec_mul(a, b)))
    P[3].set(b)
    ev.ScalarMulAddTo(P[3], P[0], c); assert np.array_equal(P[3].Coeffs, fo.vec_add(b, fo.vec_scalar_mul(a, c)))
    P[3].set(b)
    ev.ScalarMulSubTo(P[3], P[0], c); assert np.array_equal(P[3].Coeffs, fo.vec_sub(b, fo.vec_scalar_mul(a, c)))
    # domain / shape misuse raises like the reference panics
    P[0].IsNTT = False
    with pytest.raises(ValueError):
        ev.MulTo(P[2], P[0], P[1])


def _ntt_case(rs, name, logN, neg, n_polys, seed):
    q = pyref.modulus(name)
    fo = oracle.Field(q); fg = rs.Field(q)
    N = 1 << logN
    to = oracle.NTT(fo, logN, neg)
    tg = (rs.CyclotomicTransformer if neg else rs.CyclicTransformer)(fg, N)
    # twiddle tables identical to the reference construction
    tw_o, twinv_o, rinv_o = to.tables()
    tw_g, twinv_g, rinv_g = tg.tables()
    lo = 0 if neg else 1      # cyclic tw[0] is never written by the reference
    assert np.array_equal(tw_g[lo:], tw_o[lo:]) and np.array_equal(twinv_g[lo:], twinv_o[lo:])
    assert np.array_equal(rinv_g, rinv_o)
    polys = edge_polys(fo, N)[: max(0, min(4, n_polys - 1))]
    data = np.stack(polys + [rand_mont(fo, N, seed + i) for i in range(n_polys - len(polys))])
    src = rs.Batch(fg, n_polys, N).upload(data)
    dst = rs.Batch(fg, n_polys, N)
    tg.FwdNTTTo(dst, src)
    exp = to.fwd(data)
    got = dst.download()
    assert np.array_equal(got, exp), (name, logN, neg, "fwd")
    # inverse, in place
    tg.InvNTTTo(dst, dst)
    assert np.array_equal(dst.download(), data), (name, logN, neg, "inv roundtrip")
    # inverse of arbitrary data, out of place
    tg.InvNTTTo(dst, src)
    assert np.array_equal(dst.download(), to.inv(data)), (name, logN, neg, "inv")
    # forward in place on a sub-range of the batch
    if n_polys >= 3:
        tg.FwdNTTTo(src, src, dst_poly=1, src_poly=1, n_polys=2)
        d = src.download()
        assert np.array_equal(d[1:3], exp[1:3]) and np.array_equal(d[0], data[0])
        if n_polys > 3:
            assert np.array_equal(d[3:], data[3:])


@pytest.mark.parametrize("name", NAMES)
@pytest.mark.parametrize("neg", [False, True])
def test_ntt_parity_small_ranks(rs, name, neg):
    for logN, n_polys in [(0, 3), (1, 5), (2, 1), (3, 7), (4, 2), (5, 9), (7, 3), (8, 6), (9, 5), (10, 3), (11, 2), (12, 3)]:
        _ntt_case(rs, name, logN, neg, n_polys, 100 * logN)


@pytest.mark.parametrize("name,logN,n_polys", [("zp128", 13, 3), ("zp255", 13, 2), ("zp255", 14, 1), ("zp128", 15, 1),
                                               ("zp255", 16, 2), ("zp128", 16, 1), ("zp220", 16, 1), ("zp240", 14, 2),
                                               ("zp110", 17, 1), ("zp255", 17, 1), ("zp440", 13, 1), ("zp880", 12, 1),
                                               ("zp880", 13, 1), ("zp440", 14, 1)])
@pytest.mark.parametrize("neg", [False, True])
def test_ntt_parity_multi_pass(rs, name, logN, n_polys, neg):
    _ntt_case(rs, name, logN, neg, n_polys, 7 * logN)


@pytest.mark.parametrize("name", ["zp128", "zp255"])
def test_ntt_full_size_properties(rs, name):
    """BASELINE config 2 size (rank 2^16, batch 64): round trip + linearity + one poly vs oracle."""
    q = pyref.modulus(name)
    fo = oracle.Field(q); fg = rs.Field(q)
    logN, n_polys = 16, 64
    N = 1 << logN
    rng = np.random.default_rng(0x4E5454)
    # uniform limbs reduced into range by clearing top bits (canonical < q)
    raw = rng.integers(0, 1 << 63, size=(2, n_polys, N, fo.L), dtype=np.uint64)
    raw[..., fo.L - 1] &= np.uint64((1 << ((q.bit_length() - 1) % 64)) - 1)
    a, b = raw[0], raw[1]
    for neg in (False, True):
        tg = (rs.CyclotomicTransformer if neg else rs.CyclicTransformer)(fg, N)
        A = rs.Batch(fg, n_polys, N).upload(a)
        Bt = rs.Batch(fg, n_polys, N).upload(b)
        S = rs.Batch(fg, n_polys, N)
        lib = rs.lib()
        assert lib.rs_vec_add(S.h, 0, A.h, 0, Bt.h, 0, n_polys) == 0
        tg.FwdNTTTo(A, A); tg.FwdNTTTo(Bt, Bt); tg.FwdNTTTo(S, S)
        nA = A.download()
        # one polynomial against the oracle at full size
        to = oracle.NTT(fo, logN, neg)
        assert np.array_equal(nA[5], to.fwd(a[5]))
        assert lib.rs_vec_add(A.h, 0, A.h, 0, Bt.h, 0, n_polys) == 0
        assert np.array_equal(A.download(), S.download())           # linearity: NTT(a)+NTT(b) == NTT(a+b)
        tg.InvNTTTo(Bt, Bt)
        assert np.array_equal(Bt.download(), b)                       # round trip


def test_quorem_aut_reverse_evaluate(rs):
    q = pyref.modulus("zp220")
    fo = oracle.Field(q); fg = rs.Field(q)
    for rank, N in [(64, 16), (4096, 1024), (1 << 15, 1 << 13), (96, 32), (64, 64), (50, 16)]:
        p = rand_mont(fo, rank, rank)
        B = rs.Batch(fg, 3, rank).upload(np.stack([p, p, p]))
        ev = rs.CyclicEvaluator.__new__(rs.CyclicEvaluator)
        rs.bigpoly._BaseOperator.__init__(ev, fg, rank, None)
        quo, rem = ev.QuoRemByVanishing(B.poly(0), N, B.poly(1), B.poly(2))
        eq, er = fo.quorem_vanishing(p, N)
        assert np.array_equal(quo.Coeffs, eq) and np.array_equal(rem.Coeffs, er), (rank, N)
        with pytest.raises(ValueError):
            ev.QuoRemByVanishing(B.poly(0, is_ntt=True), N, B.poly(1), B.poly(2))
    for logN in (0, 1, 5, 10, 13):
        n = 1 << logN
        a = rand_mont(fo, n, 77 + logN)
        ce = rs.CyclotomicEvaluator(fg, n)
        B = rs.Batch(fg, 2, n).upload(np.stack([a, a]))
        for idx in (3, 5, -1, 2 * n - 1, 7 + 4 * n, -(2 * n + 3)):
            for is_ntt in (False, True):
                src = B.poly(0, is_ntt); dst = B.poly(1)
                ce.AutTo(dst, src, idx)
                assert np.array_equal(dst.Coeffs, fo.aut(a, idx, is_ntt)), (logN, idx, is_ntt)
        with pytest.raises(ValueError):
            ce.AutTo(B.poly(1), B.poly(0), 4)
    # nttChecker.TransposeTo pre-step + Horner evaluation
    n = 1 << 12
    a = rand_mont(fo, n, 5)
    B = rs.Batch(fg, 2, n).upload(np.stack([a, a]))
    scale = fo.to_mont([n])[0]
    assert rs.lib().rs_reverse_scale(B.h, 1, B.h, 0, scale.ctypes.data_as(rs.capi.c_u64p)) == 0
    assert np.array_equal(B.download(1, 1)[0], fo.vec_scalar_mul(a[::-1].copy(), scale))
    x = rand_mont(fo, 1, 9)[0]
    for cnt in (n, n - 1, 1, 65, 64, 63):
        assert np.array_equal(B.poly(0).Evaluate(x, cnt), fo.poly_evaluate(a[:cnt], x)), cnt
    # long vectors (> 2^16 coefficients: chunked Horner, then the partial values evaluated at x^64 by the one-CTA kernel)
    nl = (1 << 17) + 77
    al = rand_mont(fo, nl, 15)
    BL = rs.Batch(fg, 1, nl).upload(al[None])
    for cnt in (nl, (1 << 16) + 1, (1 << 16) + 64, 1 << 16):
        assert np.array_equal(BL.poly(0).Evaluate(x, cnt), fo.poly_evaluate(al[:cnt], x)), cnt
    # the batched form (jindo.Prover.Evaluate's closing loop): ragged lengths incl. 1, < 256 and non-multiples of 256
    import ctypes
    lens = [n, n - 1, 1, 255, 256, 257, 1000, 3]
    hs = (ctypes.c_void_p * len(lens))(*[B.h.value] * len(lens))
    ps = (ctypes.c_size_t * len(lens))(*[k % 2 for k in range(len(lens))])
    ls = (ctypes.c_size_t * len(lens))(*lens)
    out = np.zeros((len(lens), fo.L), dtype=np.uint64)
    a1 = B.download(1, 1)[0]
    assert rs.lib().rs_poly_evaluate_many(hs, ps, ls, len(lens), x.ctypes.data_as(rs.capi.c_u64p), out.ctypes.data_as(rs.capi.c_u64p)) == 0, rs.lib().rs_last_error()
    for k, cnt in enumerate(lens):
        assert np.array_equal(out[k], fo.poly_evaluate((a if k % 2 == 0 else a1)[:cnt], x)), ("many", cnt)


@pytest.mark.parametrize("name", ["zp128", "zp240"])
def test_eval_circuit(rs, name):
    """MultCircuit-shaped constraint set (examples/mult/main.go:35-65) through the fused kernel."""
    q = pyref.modulus(name)
    fo = oracle.Field(q); fg = rs.Field(q)
    rank = 1 << 12
    pw = np.stack([rand_mont(fo, rank, 1)])
    w = np.stack([rand_mont(fo, rank, 10 + i) for i in range(7)])
    coeffs = fo.to_mont([1, q - 1, q - 3, 5])
    terms = [(0, 0, -1, [1, 0]), (0, 1, -1, [2]),                      # Y*X - Z
             (1, 0, -1, [4, 4, 4]), (1, 1, -1, [4]),                   # w^3 - w (ternary)
             (2, 0, -1, [5, 5, 5]), (2, 1, -1, [5]),
             (3, 0, -1, [6, 6, 6]), (3, 1, -1, [6]),
             (4, 0, -1, [0]), (4, 2, -1, [4]), (4, 1, -1, [5]), (4, 1, -1, [6]),   # X - 3 w0 - w1 - w2
             (5, 3, 0, [3, 3])]                                         # 5 * pw * w3^2
    bc = rand_mont(fo, 1, 99)[0]
    exp = fo.eval_circuit(list(pw), list(w), terms, coeffs, bc, rank)
    PW = rs.Batch(fg, 1, rank).upload(pw); W = rs.Batch(fg, 7, rank).upload(w); O = rs.Batch(fg, 1, rank)
    T = (rs.capi.Term * len(terms))()
    for i, (c, co, p, ws) in enumerate(terms):
        T[i].constraint, T[i].coeff, T[i].pw, T[i].n_w = c, co, p, len(ws)
        for k, x in enumerate(ws):
            T[i].w[k] = x
    rc = rs.lib().rs_eval_circuit(O.h, 0, PW.h, W.h, T, len(terms), coeffs.ctypes.data_as(rs.capi.c_u64p), len(coeffs),
                                  bc.ctypes.data_as(rs.capi.c_u64p))
    assert rc == 0, rs.lib().rs_last_error()
    assert np.array_equal(O.download()[0], exp)


def test_shape_errors(rs):
    q = pyref.modulus("zp255")
    fg = rs.Field(q)
    t = rs.CyclicTransformer(fg, 256)
    small = rs.Batch(fg, 1, 128); ok = rs.Batch(fg, 1, 256)
    with pytest.raises(rs.RingoError) as e:
        t.FwdNTTTo(ok, small)
    assert e.value.code == -3 and "inconsistent input(s)" in str(e.value)
    with pytest.raises(rs.RingoError):
        t.FwdNTTTo(small, ok)
    with pytest.raises(ValueError):
        rs.CyclicTransformer(fg, 100)          # "rank must be a power of two" (ntt.go:27-29)


@pytest.mark.parametrize("env", [{"RS_NTT52": "1"}, {"RS_NTT_MIX": "1/4"}, {"RS_NTT_MIX": "2/5"}])
@pytest.mark.parametrize("name", ["zp255", "zp240", "zp220"])
def test_fp64_pipe_paths_equal_integer_path(rs, name, env, monkeypatch):
    """The opt-in FP64-pipe kernels (ntt52.cuh: 52-bit limbs in doubles, lazy reduction) and the two-pipe kernels that
    share a batch between integer and FP64 tiles give the integer kernels' bytes: every kind of transform, ranks with
    one and two passes, partial tiles, unit-twiddle rounds, q == 1 mod 2^52 (zp255, zp220) and the generic modulus (zp240)."""
    q = pyref.modulus(name)
    fg = rs.Field(q); fo = oracle.Field(q)
    cases = [(1, 3), (4, 5), (9, 3), (10, 2), (11, 3), (13, 2), (16, 8)]
    for logN, n_polys in cases:
        N = 1 << logN
        rng = np.random.default_rng(logN * 31 + n_polys)
        raw = rng.integers(0, 1 << 63, size=(n_polys, N, fo.L), dtype=np.uint64)
        raw[..., fo.L - 1] &= np.uint64((1 << ((q.bit_length() - 1) % 64)) - 1)
        raw[0, :min(4, N)] = fo.to_mont([0, 1, q - 1, q - 2])[:min(4, N)]          # edge values
        for neg in (False, True):
            cls = rs.CyclotomicTransformer if neg else rs.CyclicTransformer
            ref_t = cls(fg, N)
            for k, v in env.items():
                monkeypatch.setenv(k, v)
            alt_t = cls(fg, N)                                     # the path is chosen when the tables are built
            for k in env:
                monkeypatch.delenv(k)
            for inv in (False, True):
                A = rs.Batch(fg, n_polys, N).upload(raw)
                Bt = rs.Batch(fg, n_polys, N).upload(raw)
                (ref_t.InvNTTTo if inv else ref_t.FwdNTTTo)(A, A)
                (alt_t.InvNTTTo if inv else alt_t.FwdNTTTo)(Bt, Bt)
                assert np.array_equal(A.download(), Bt.download()), (name, env, logN, neg, inv)
        if logN == 9:
            to = oracle.NTT(fo, logN, True)
            assert np.array_equal(Bt.download()[1], to.inv(raw[1]))   # and the oracle itself once


@pytest.mark.parametrize("name", ["zp255", "zp240", "zp220", "zp110"])
@pytest.mark.parametrize("neg", [False, True])
def test_ntt_raw_extremes(rs, name, neg):
    """The kernels work on the raw Montgomery words, and the Shoup kernels of the 8-limb fields (csrc/shoup.cuh) keep
    values in a lazy range just above 2q: polynomials whose RAW words are extreme (all q - 1, alternating 0 / q - 1,
    q - 1 in one half, top limbs saturated) drive the lazy sums to the top of that range.  Forward and inverse against
    the oracle, single-pass and multi-pass ranks."""
    q = pyref.modulus(name)
    fo = oracle.Field(q); fg = rs.Field(q)
    for logN in (6, 10, 13):
        N = 1 << logN
        top = oracle.ints_to_array([q - 1] * N, fo.L)
        zero = oracle.ints_to_array([0] * N, fo.L)
        alt = oracle.ints_to_array([(q - 1) if i & 1 else 0 for i in range(N)], fo.L)
        half = oracle.ints_to_array([(q - 1) if i < N // 2 else 0 for i in range(N)], fo.L)
        blocks = oracle.ints_to_array([(q - 1) if (i >> 3) & 1 else 1 for i in range(N)], fo.L)
        rng = random.Random(logN)
        hi = oracle.ints_to_array([q - 1 - rng.randrange(1 << 40) for _ in range(N)], fo.L)
        data = np.stack([top, zero, alt, half, blocks, hi])
        to = oracle.NTT(fo, logN, neg)
        tg = (rs.CyclotomicTransformer if neg else rs.CyclicTransformer)(fg, N)
        src = rs.Batch(fg, len(data), N).upload(data)
        dst = rs.Batch(fg, len(data), N)
        tg.FwdNTTTo(dst, src)
        assert np.array_equal(dst.download(), to.fwd(data)), (name, logN, neg, "fwd")
        tg.InvNTTTo(dst, src)
        assert np.array_equal(dst.download(), to.inv(data)), (name, logN, neg, "inv")
        tg.FwdNTTTo(dst, dst)
        assert np.array_equal(dst.download(), data), (name, logN, neg, "round trip")


def test_shoup_and_montgomery_kernels_agree_over_many_products(rs, monkeypatch):
    """The Shoup kernels take a slow path when the truncated quotient of a product may be one short (about 8 * 2^-32 of the
    products, csrc/shoup.cuh).  tests/test_device_shoup_host.py checks that path on engineered operands; here both kernel
    families run the same 1.3e9 butterfly products of the 255-bit field (a handful of them land on the slow path) and must
    agree bit for bit.  The Montgomery kernels are selected per field handle by RS_NTT_NO_SHOUP (read by rs_field_create)."""
    q = pyref.modulus("zp255")
    logN, n_polys = 16, 256
    N = 1 << logN
    monkeypatch.setenv("RS_NTT_NO_SHOUP", "1")
    f_mont = rs.Field(q)
    t_mont = rs.CyclotomicTransformer(f_mont, N)
    monkeypatch.delenv("RS_NTT_NO_SHOUP")
    f_sh = rs.Field(q)
    t_sh = rs.CyclotomicTransformer(f_sh, N)
    rng = np.random.default_rng(0x5348)
    raw = rng.integers(0, 1 << 63, size=(n_polys, N, f_sh.L), dtype=np.uint64)
    raw[..., f_sh.L - 1] &= np.uint64((1 << ((q.bit_length() - 1) % 64)) - 1)       # canonical: below q
    A = rs.Batch(f_sh, n_polys, N).upload(raw)
    B = rs.Batch(f_mont, n_polys, N).upload(raw)
    for it in range(10):                       # each output is the next input: fresh canonical values every round
        t_sh.FwdNTTTo(A, A)
        t_mont.FwdNTTTo(B, B)
        a = A.download()
        assert np.array_equal(a, B.download()), ("forward", it)
    for it in range(3):
        t_sh.InvNTTTo(A, A)
        t_mont.InvNTTTo(B, B)
        assert np.array_equal(A.download(), B.download()), ("inverse", it)
