"""GPU parity tests for the Jindo path (rs_jindo_* through the C ABI) against the oracle
(oracle/jindo_ref.py) on the same parameters, commit key and randomness tape -- bit-exact on
every polynomial of the Opening, the Commitment and the Evaluate responses."""
import os

import numpy as np
import pytest

import oracle
from oracle import jindo_ref as J, pyref
from jindo_util import cdt_array, params_dict, polys_array, tape_arrays

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rs(product):
    L = product.load()
    assert L.rs_init(0) == 0, L.rs_last_error()
    return product


CASES = [
    ("zp255", 1024, 1, 1024),    # TestJindo single (jindo_test.go:16-52): 2 limbs -> 1 limb
    ("zp255", 1024, 1, 700),     # short vector: rows past len(v) stay zero (prover.go:118-125)
    ("zp255", 1024, 1, 1),       # single coefficient
    ("zp128", 700, 2, 700),      # 3-limb rings, 128-bit field, batch parameters
    ("zp220", 2049, 9, 2049),    # TestPublicKey commitment shape
    ("zp240", 600, 3, 513),
]


@pytest.mark.parametrize("name,N,batch,length", CASES)
def test_commit_and_evaluate_parity(rs, name, N, batch, length):
    q = pyref.modulus(name)
    fo = oracle.Field(q); fg = rs.Field(q)
    p = J.new_parameters(q, N, batch)
    tape = J.Tape(b"gpu-parity-" + name.encode())
    prv = J.Prover(p, b"Jindo!", tape)
    gp = rs.jindo.Prover(fg, params_dict(p, prv.ecd), polys_array(prv.ck.In), polys_array(prv.ck.MLWE),
                         polys_array(prv.ck.Out), cdt_array(prv.ecd))
    v = [J.field_set_random(tape, q) for _ in range(length)]
    com_o, open_o = prv.commit(v)
    rec = prv.trace
    # committed vector lives in a device batch, at a non-zero poly index
    rank_pad = max(length, 8)
    B = rs.Batch(fg, 2, rank_pad)
    host = np.zeros((2, rank_pad, fo.L), dtype=np.uint64)
    host[1, :length] = fo.to_mont(v)
    B.upload(host)
    com_g, open_g = gp.Commit(B, 1, length, tape_arrays(fo, p, rec))
    # Opening: every Encode / MLWE / InCommit polynomial
    for i in range(p.cols + 1):
        for j in range(p.rows):
            assert np.array_equal(open_g.Encode(i, j), polys_array(open_o.Encode[i][j])), ("Encode", i, j)
        for j in range(p.mlwe + p.in_msis):
            assert np.array_equal(open_g.MLWE(i, j), polys_array(open_o.MLWE[i][j])), ("MLWE", i, j)
    for k in range(p.in_com_dcmp_len):
        assert np.array_equal(open_g.InCommit(k), polys_array(open_o.InCommit[k])), ("InCommit", k)
    assert np.array_equal(com_g, polys_array(com_o))
    # Evaluate (batch-1 flow of prover.go:267-316 on this opening)
    x = J.field_set_random(tape, q)
    left = polys_array([prv.ecd.encode([le]) for le in J.left_vec(p, x)])
    part = gp.Partial(open_g, left)
    ring = p.ringQ
    exp_part = []
    for i in range(p.cols + 1):
        acc = ring.new_poly()
        for j in range(p.rows):
            acc = ring.mul_mont_then_add(left[j].tolist(), open_o.Encode[i][j], acc)
        exp_part.append(acc)
    assert np.array_equal(part, polys_array(exp_part))
    chal_bytes = [bytes([17 * (i + 1) % 256] * 16) for i in range(p.cols)]
    chals = [J.encode_challenge(p, ring, cb) for cb in chal_bytes]
    enc_g, mlwe_g = gp.Respond(open_g, polys_array(chals))
    for i in range(p.rows):
        acc = [list(c) for c in open_o.Encode[p.cols][i]]
        for j in range(p.cols):
            acc = ring.mul_mont_then_add(chals[j], open_o.Encode[j][i], acc)
        assert np.array_equal(enc_g[i], polys_array(acc)), ("resp Encode", i)
    for i in range(p.mlwe + p.in_msis):
        acc = [list(c) for c in open_o.MLWE[p.cols][i]]
        for j in range(p.cols):
            acc = ring.mul_mont_then_add(chals[j], open_o.MLWE[j][i], acc)
        assert np.array_equal(mlwe_g[i], polys_array(acc)), ("resp MLWE", i)
    # batch combine (prover.go:251-265): acc = 0 + open (.) challenge
    cq = J.encode_challenge(p, ring, b"\x05" * 16); co = J.encode_challenge(p, p.ringQOut, b"\x05" * 16)
    acc_g = gp.NewOpening()
    gp.OpeningMulAdd(acc_g, open_g, polys_array(cq), polys_array(co))
    gp.OpeningMulAdd(acc_g, open_g, polys_array(cq), polys_array(co))
    z = ring.new_poly()
    e = ring.mul_mont_then_add(open_o.Encode[0][1], cq, ring.mul_mont_then_add(open_o.Encode[0][1], cq, z))
    assert np.array_equal(acc_g.Encode(0, 1), polys_array(e))
    zo = p.ringQOut.new_poly()
    e = p.ringQOut.mul_mont_then_add(open_o.InCommit[3], co, p.ringQOut.mul_mont_then_add(open_o.InCommit[3], co, zo))
    assert np.array_equal(acc_g.InCommit(3), polys_array(e))
    # the same loop in one pass (rs_opening_combine_bytes): two openings, two different challenges
    b2 = b"\x05" * 16 + bytes(range(16))
    seq = gp.NewOpening()
    gp.OpeningMulAddBytes(seq, open_g, b2[:16]); gp.OpeningMulAddBytes(seq, acc_g, b2[16:])
    one = gp.OpeningCombineBytes([open_g, acc_g], b2)
    for k in range((p.cols + 1) * p.in_msis):
        assert np.array_equal(one.InCommit(k), seq.InCommit(k)), ("combine InCommit", k)
    for c in range(p.cols + 1):
        for r in (0, 1, p.rows - 1):
            assert np.array_equal(one.Encode(c, r), seq.Encode(c, r)), ("combine Encode", c, r)
        for r in (0, p.mlwe + p.in_msis - 1):
            assert np.array_equal(one.MLWE(c, r), seq.MLWE(c, r)), ("combine MLWE", c, r)


def test_commit_shape_errors(rs):
    q = pyref.modulus("zp255")
    fo = oracle.Field(q); fg = rs.Field(q)
    p = J.new_parameters(q, 1024, 1)
    tape = J.Tape(b"err")
    prv = J.Prover(p, b"Jindo!", tape)
    gp = rs.jindo.Prover(fg, params_dict(p, prv.ecd), polys_array(prv.ck.In), polys_array(prv.ck.MLWE),
                         polys_array(prv.ck.Out), cdt_array(prv.ecd))
    v = [J.field_set_random(tape, q) for _ in range(64)]
    prv.commit(v)
    t = tape_arrays(fo, p, prv.trace)
    B = rs.Batch(fg, 1, 4096)
    with pytest.raises(rs.RingoError) as e:
        gp.Commit(B, 0, p.rank + 1, t)                       # "len(v) > params.rank" (prover.go:47-49)
    assert "len(v) > params.rank" in str(e.value)
    t["twin_u"] = t["twin_u"][:-1]
    with pytest.raises(rs.RingoError):
        gp.Commit(B, 0, 64, t)


@pytest.mark.parametrize("batch", [1, 3])
def test_gpu_proof_verifies(rs, batch):
    """TestJindo (jindo/jindo_test.go:16-52) with the prover's arithmetic on the GPU: Commit and
    Evaluate go through the C ABI (host keeps the SHAKE transcript, prover.go:220-302), and the
    restated jindo.Verifier must accept the resulting proof."""
    import hashlib
    q = pyref.modulus("zp255")
    N = 512
    fo = oracle.Field(q); fg = rs.Field(q)
    p = J.new_parameters(q, N, batch)
    tape = J.Tape(b"gpu-verify-%d" % batch)
    prv = J.Prover(p, b"Jindo!", tape)                      # used only as the randomness / key source
    vrf = J.Verifier(p, b"Jindo!", tape, commit_key=prv.ck)
    gp = rs.jindo.Prover(fg, params_dict(p, prv.ecd), polys_array(prv.ck.In), polys_array(prv.ck.MLWE),
                         polys_array(prv.ck.Out), cdt_array(prv.ecd))
    ring, ro = p.ringQ, p.ringQOut
    vs, coms, opens = [], [], []
    B = rs.Batch(fg, batch, N)
    for b in range(batch):
        v = [J.field_set_random(tape, q) for _ in range(N)]
        prv.commit(v)                                       # draws the randomness in reference order
        B.upload(fo.to_mont(v).reshape(1, N, fo.L), b)
        com, op = gp.Commit(B, b, N, tape_arrays(fo, p, prv.trace))
        vs.append(v); coms.append(com.tolist()); opens.append(op)
    x = J.field_set_random(tape, q)
    L = (q.bit_length() + 63) // 64
    prefix = b"Jindo!" + b"".join(J.commitment_write_raw(p, c) for c in coms) + x.to_bytes(8 * L, "big")
    if batch > 1:
        bb = hashlib.shake_128(prefix).digest(16 * batch)
        prefix += bb
        ob = gp.NewOpening()
        for i in range(batch):
            cq = J.encode_challenge(p, ring, bb[16 * i:16 * i + 16]); co = J.encode_challenge(p, ro, bb[16 * i:16 * i + 16])
            gp.OpeningMulAdd(ob, opens[i], polys_array(cq), polys_array(co))
    else:
        ob = opens[0]
    left = polys_array([prv.ecd.encode([le]) for le in J.left_vec(p, x)])
    part = gp.Partial(ob, left)
    pf = {"Partial": part[:p.cols].tolist(), "PartialMask": part[p.cols].tolist(),
          "InCommit": [ob.InCommit(k).tolist() for k in range(p.in_com_dcmp_len)]}
    tr = prefix + b"".join(J.poly_write_to(c) for c in pf["Partial"]) + J.poly_write_to(pf["PartialMask"])
    cs = hashlib.shake_128(tr).digest(16 * p.cols)
    chals = [J.encode_challenge(p, ring, cs[16 * i:16 * i + 16]) for i in range(p.cols)]
    enc, ml = gp.Respond(ob, polys_array(chals))
    pf["Encode"], pf["MLWE"] = enc.tolist(), ml.tolist()
    # evaluations by the chunked-Horner kernel (poly.go:64-76)
    xm = fo.to_mont([x])[0]
    ys = [fo.from_mont(B.poly(b).Evaluate(xm, N).reshape(1, -1))[0] for b in range(batch)]
    assert ys == [sum(c * pow(x, i, q) for i, c in enumerate(v)) % q for v in vs]
    assert vrf.verify(x, coms, ys, pf)
    bad = list(ys); bad[0] = (bad[0] + 1) % q
    assert not vrf.verify(x, coms, bad, pf)
    # ---- the same decisions from the device verifier (rs_jindo_verify), check by check
    ymont = lambda vals: fo.to_mont(vals)
    bbytes = bb if batch > 1 else b""
    dv = lambda ys_, pf_, coms_=coms: gp.Verify(xm, coms_, bbytes, cs, ymont(ys_), pf_, p.in_com_dcmp_two_nm, p.res_two_nm)
    assert dv(ys, pf) == 15
    assert dv(bad, pf) == 7                                   # only verifyEval fails
    import copy
    t1 = copy.deepcopy(pf); t1["Encode"][1][0][5] = (t1["Encode"][1][0][5] + 1) % p.q_primes[0]
    assert not vrf.verify(x, coms, ys, t1)
    c1 = dv(ys, t1); assert c1 != 15 and not (c1 & 4)         # breaks the consistency equation (and the inner norm)
    t2 = copy.deepcopy(pf); t2["InCommit"][3][0][7] = (t2["InCommit"][3][0][7] + 12345) % p.qout_primes[0]
    assert not vrf.verify(x, coms, ys, t2)
    c2 = dv(ys, t2); assert not (c2 & 1) or not (c2 & 2)      # outer / inner commitment equations leave the norm ball
    t3 = copy.deepcopy(pf); t3["Partial"][0][0][0] = (t3["Partial"][0][0][0] + 1) % p.q_primes[0]
    assert not vrf.verify(x, coms, ys, t3)
    c3 = dv(ys, t3); assert c3 != 15
    coms_bad = copy.deepcopy(coms); coms_bad[0][0][0][1] = (coms_bad[0][0][0][1] + 1) % p.qout_primes[0]
    assert not vrf.verify(x, coms_bad, ys, pf) and not (dv(ys, pf, coms_bad) & 1)


def _prover_pair(rs, name, N, batch, seed):
    q = pyref.modulus(name)
    fo = oracle.Field(q); fg = rs.Field(q)
    p = J.new_parameters(q, N, batch)
    tape = J.Tape(seed)
    prv = J.Prover(p, b"Jindo!", tape)
    gp = rs.jindo.Prover(fg, params_dict(p, prv.ecd), polys_array(prv.ck.In), polys_array(prv.ck.MLWE),
                         polys_array(prv.ck.Out), cdt_array(prv.ecd))
    return q, fo, fg, p, tape, prv, gp


def test_device_uniform_stream_matches_csprng(rs):
    """rs_jindo_set_twin_stream / _draw against the restated csprng.UniformSampler (uniform.go:26-70):
    AES-256-CTR + the XOR-accumulating 8 KiB refill, across refill boundaries and unaligned call sizes."""
    q, fo, fg, p, tape, prv, gp = _prover_pair(rs, "zp255", 1024, 1, b"stream")
    for seed in (b"\x00" * 32, bytes(range(32)), b"short"):
        ref = J.UniformSampler(seed)
        gp.SetTwinSeed(seed)
        for n in (1, 100, 923, 1024, 5000, 0, 7, 300000, 1024 * 33, 1, 4097 * 1024 + 5):
            got = gp.TwinStreamDraw(n)
            assert np.array_equal(got, ref.sample_many(n)), (seed, n)


def test_commit_with_device_twin_stream(rs, monkeypatch):
    """Consecutive Commits drawing the TwinCDT uniforms on the device (tape.twin_u = NULL) give the same Opening /
    Commitment as the oracle, whose encoder owns one stateful sampler (encoder.go:56-62).  With RS_JINDO_PREFETCH every Commit pre-generates
    the next Commit's draws (uniform_stream_prefetch): the lengths below make the prefetch miss (different number of
    rows), hit twice, and miss again, and a host-side draw in between moves the stream under a pending prefetch."""
    monkeypatch.setenv("RS_JINDO_PREFETCH", "1")       # (read by the library when the prover handle is created)
    q, fo, fg, p, tape, prv, gp = _prover_pair(rs, "zp128", 700, 2, b"dev-twin")
    gp.SetTwinSeed(prv.ecd.twin.base.seed)
    for step, length in enumerate((700, 333, 333, 333, 40, 700)):
        if step == 4:                                  # the oracle's sampler hands 777 draws to somebody else
            assert np.array_equal(gp.TwinStreamDraw(777), prv.ecd.twin.base.sample_many(777))
        v = [J.field_set_random(tape, q) for _ in range(length)]
        com_o, open_o = prv.commit(v)
        B = rs.Batch(fg, 1, 1024)
        host = np.zeros((1, 1024, fo.L), dtype=np.uint64)
        host[0, :length] = fo.to_mont(v)
        B.upload(host)
        t = tape_arrays(fo, p, prv.trace)
        assert len(t["twin_u"]) > 0
        t["twin_u"] = None
        com_g, open_g = gp.Commit(B, 0, length, t)
        assert np.array_equal(com_g, polys_array(com_o))
        for i in range(p.cols + 1):
            for j in range(p.rows):
                assert np.array_equal(open_g.Encode(i, j), polys_array(open_o.Encode[i][j])), ("Encode", i, j)


def test_commit_full_size_properties(rs):
    """BASELINE configs[2]: one Commit of 2^22 coefficients over the 255-bit field (NewParameters(2^22, 1)),
    synthetic key and tape.  The Python restatement cannot commit at this size, so the check is through
    size-independent properties of the result:
      (1) decode: every encoded polynomial evaluates, at X^slots = b, to the field elements it carries
          (encoder.go:120-146 digits + (X^slots - b) g, and b^exp = -1 mod q) -- sampled over data rows,
          first / last rows, the mask column and the ragged tail;
      (2) one inner-commitment column recomputed from the downloaded Opening with the oracle's ring
          arithmetic and cut-off pipeline (prover.go:144-176) equals Opening.InCommit;
      (3) the same tape gives the same commitment (determinism)."""
    name, N = "zp255", 1 << 22
    length = N - 12345                                    # ragged: the last data row is partial
    q = pyref.modulus(name)
    fo = oracle.Field(q); fg = rs.Field(q)
    p = J.new_parameters(q, N, 1)
    ecd = J.Encoder(p, J.Tape(b"full-size"))
    rng = np.random.default_rng(2022)

    def rpolys(shape, primes):
        a = np.empty(shape + (len(primes), p.d), dtype=np.uint64)
        for l, pr in enumerate(primes):
            a[..., l, :] = rng.integers(0, pr, size=shape + (p.d,), dtype=np.uint64)
        return a

    def felems(n):                                        # Montgomery limbs of canonical elements (top limb < 2^62 < q's)
        return rng.integers(0, 1 << 62, size=(n, fo.L), dtype=np.uint64)
    cs = p.cols * p.slots
    data_rows = [r for r in range(1, p.rows - 1) if r * cs <= length]          # mask column rule (prover.go:103)
    # data columns: row r of column i exists while r*cs + i*slots <= len(v)   (note '>' in prover.go:121)
    col_rows = [sum(1 for r in range(1, p.rows - 1) if r * cs + i * p.slots <= length) for i in range(p.cols)]
    n_twin = sum(col_rows) + p.cols; n_mask = len(data_rows) + 2; n_cosac = p.cols + n_mask
    n_noise = (p.cols + 1) * (p.in_msis + p.mlwe)
    ck_in, ck_mlwe = rpolys((p.in_msis, p.rows), p.q_primes), rpolys((p.in_msis, p.mlwe), p.q_primes)
    ck_out = rpolys((p.out_msis, p.in_com_dcmp_len), p.qout_primes)
    tape = {"first_row": felems(cs), "last_row": felems(cs), "mask_rows": felems(n_mask * p.slots),
            "twin_u": rng.integers(0, 1 << 63, size=n_twin * p.d, dtype=np.int64).view(np.uint64),
            "cosac_g": rng.integers(-5000, 5000, size=n_cosac * p.d, dtype=np.int64),
            "mlwe": rng.integers(-20, 20, size=n_noise * p.d, dtype=np.int64)}
    gp = rs.jindo.Prover(fg, params_dict(p, ecd), ck_in, ck_mlwe, ck_out, cdt_array(ecd))
    v = felems(length)
    B = rs.Batch(fg, 1, N)
    host = np.zeros((1, N, fo.L), dtype=np.uint64); host[0, :length] = v
    B.upload(host)
    com, op = gp.Commit(B, 0, length, tape)

    ring = p.ringQ
    rec = J.RNSReconstructor(ring)
    R = 1 << (64 * fo.L)

    def to_int(limbs):                                    # Montgomery limbs -> canonical value
        x = sum(int(w) << (64 * k) for k, w in enumerate(limbs))
        return x * pow(R, -1, q) % q

    def decode(poly):                                     # NTT+Montgomery ring polynomial -> slots field values
        c = rec.reconstruct(ring.intt(ring.imform([[int(x) for x in limb] for limb in poly])))
        return [sum(c[k * p.slots + s] * pow(p.base, k, q) for k in range(p.exp)) % q for s in range(p.slots)]

    # (1) decode property on sampled polynomials
    last_data = data_rows[-1]
    samples = [(0, 1), (p.cols - 1, 1), (7, 100), (p.cols // 2, last_data // 2), (0, last_data), (p.cols - 1, last_data),
               (3, 0), (p.cols - 1, 0), (5, p.rows - 1), (p.cols, 0), (p.cols, 3), (p.cols, p.rows - 1)]
    samples += [(int(rng.integers(0, p.cols)), int(rng.integers(1, last_data))) for _ in range(12)]
    mask_index = {0: 0, p.rows - 1: n_mask - 1}
    for (i, j) in samples:
        got = decode(op.Encode(i, j))
        if i == p.cols:
            m = mask_index.get(j, j)
            exp = [to_int(e) for e in tape["mask_rows"][m * p.slots:(m + 1) * p.slots]]
        elif j == 0:
            exp = [to_int(e) for e in tape["first_row"][i * p.slots:(i + 1) * p.slots]]
        elif j == p.rows - 1:
            exp = [to_int(e) for e in tape["last_row"][i * p.slots:(i + 1) * p.slots]]
        else:
            s0 = j * cs + i * p.slots
            if s0 > length:                                   # no such row: zero polynomial
                assert not op.Encode(i, j).any()
                continue
            exp = [to_int(v[s0 + s]) if s0 + s < length else 0 for s in range(p.slots)]
        assert got == exp, ("decode", i, j)
    # rows past len(v) stay zero polynomials (prover.go:118-125)
    if last_data + 1 < p.rows - 1:
        assert not op.Encode(0, last_data + 1).any()
    short = [i for i in range(p.cols) if col_rows[i] < len(data_rows)]
    if short:
        assert not op.Encode(short[-1], last_data).any()

    # (2) inner commitment of one column, one output index, from the downloaded Opening
    i_col, o_idx = 5, 3
    prv = J.Prover.__new__(J.Prover)                      # only cut_pipeline's state is needed
    prv.p = p; prv.rns_out = J.RNSReconstructor(p.ringQOut)
    acc = ring.new_poly()
    for k in range(p.rows):
        e = op.Encode(i_col, k)
        if e.any():
            acc = ring.mul_mont_then_add(ck_in[o_idx, k].tolist(), e.tolist(), acc)
    for k in range(p.mlwe):
        acc = ring.mul_mont_then_add(ck_mlwe[o_idx, k].tolist(), op.MLWE(i_col, k).tolist(), acc)
    acc = ring.add(op.MLWE(i_col, p.mlwe + o_idx).tolist(), acc)
    expect = prv.cut_pipeline(ring, rec, acc, p.log_in_cut)
    assert np.array_equal(op.InCommit(i_col * p.in_msis + o_idx), polys_array(expect))

    # (3) determinism
    com2, _ = gp.Commit(B, 0, length, tape)
    assert np.array_equal(com, com2)


@pytest.mark.parametrize("name,N,batch,length", [("zp255", 1024, 1, 700), ("zp220", 2049, 9, 2049), ("zp128", 700, 2, 1)])
def test_commit_host_matches_commit(rs, name, N, batch, length):
    """rs_jindo_commit_host (v uploaded piecewise behind the encode kernels) == rs_jindo_commit == oracle."""
    import ctypes
    q, fo, fg, p, tape, prv, gp = _prover_pair(rs, name, N, batch, b"host-v-" + name.encode())
    v = [J.field_set_random(tape, q) for _ in range(length)]
    com_o, open_o = prv.commit(v)
    t = tape_arrays(fo, p, prv.trace)
    vm = np.ascontiguousarray(fo.to_mont(v), dtype=np.uint64)
    keep = {k: np.ascontiguousarray(t[k], dtype=(np.int64 if k in ("cosac_g", "mlwe") else np.uint64)) for k in t}
    T = rs.capi.JindoTape()
    vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    T.first_row, T.last_row, T.mask_rows = vp(keep["first_row"]), vp(keep["last_row"]), vp(keep["mask_rows"])
    T.n_mask_rows = keep["mask_rows"].size // (fo.L * p.slots)
    T.twin_u, T.n_twin_u = vp(keep["twin_u"]), keep["twin_u"].size
    T.cosac_g, T.n_cosac_polys = vp(keep["cosac_g"]), keep["cosac_g"].size // p.d
    T.mlwe = vp(keep["mlwe"])
    op = gp.NewOpening()
    com = np.zeros((p.out_msis, len(p.qout_primes), p.d), dtype=np.uint64)
    gp.CommitHostInto(vm.ctypes.data, length, T, op, com)
    assert np.array_equal(com, polys_array(com_o))
    for i in range(p.cols + 1):
        for j in range(p.rows):
            assert np.array_equal(op.Encode(i, j), polys_array(open_o.Encode[i][j])), ("Encode", i, j)
    for k in range(p.in_com_dcmp_len):
        assert np.array_equal(op.InCommit(k), polys_array(open_o.InCommit[k])), ("InCommit", k)


@pytest.mark.parametrize("name,logn,length_off,huge", [("zp255", 16, 0, False), ("zp255", 18, 4321, False), ("zp128", 14, 77, False),
                                                       ("zp220", 15, 1, False), ("zp255", 14, 5, True), ("zp220", 13, 0, True)])
def test_commit_mid_size_bit_exact_vs_c_oracle(rs, name, logn, length_off, huge):
    """Sizes the Python restatement cannot reach: the whole Opening and the Commitment against the C
    restatement (oracle/jindo_oracle.c, pinned to the Python one by tests/test_oracle_jindo_c.py) on a
    synthetic key / tape, including ragged lengths and 3-limb rings.  `huge`: half of the supplied noise integers span
    +-2^62, far beyond what a sampler produces -- setCoeffSigned (jindo/utils.go:49-60) is defined for any int64, and this
    is what drives encode_kernel's term-by-term fallback instead of its combined int64 path."""
    from oracle import jindo_c
    q = pyref.modulus(name)
    fo = oracle.Field(q); fg = rs.Field(q)
    N = 1 << logn
    length = N - length_off
    p = J.new_parameters(q, N, 1)
    ecd = J.Encoder(p, J.Tape(b"mid-" + name.encode()))
    rng = np.random.default_rng(logn)

    def rpolys(shape, primes):
        a = np.empty(shape + (len(primes), p.d), dtype=np.uint64)
        for l, pr in enumerate(primes):
            a[..., l, :] = rng.integers(0, pr, size=shape + (p.d,), dtype=np.uint64)
        return a

    top_bits = (q.bit_length() - 1) % 64          # canonical Montgomery limbs: top limb below q's top limb
    def felems(n):
        a = rng.integers(0, 1 << 63, size=(n, fo.L), dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, size=(n, fo.L), dtype=np.uint64)
        a[:, fo.L - 1] &= np.uint64((1 << top_bits) - 1)
        return a
    cs = p.cols * p.slots
    col_rows = [sum(1 for r in range(1, p.rows - 1) if r * cs + i * p.slots <= length) for i in range(p.cols)]
    n_mask = 2 + sum(1 for r in range(1, p.rows - 1) if r * cs <= length)
    n_twin = sum(col_rows) + p.cols; n_cosac = p.cols + n_mask; n_noise = (p.cols + 1) * (p.in_msis + p.mlwe)
    ck_in, ck_mlwe = rpolys((p.in_msis, p.rows), p.q_primes), rpolys((p.in_msis, p.mlwe), p.q_primes)
    ck_out = rpolys((p.out_msis, p.in_com_dcmp_len), p.qout_primes)
    tape = {"first_row": felems(cs), "last_row": felems(cs), "mask_rows": felems(n_mask * p.slots),
            "twin_u": rng.integers(0, 1 << 63, size=n_twin * p.d, dtype=np.int64).view(np.uint64) * np.uint64(2),
            "cosac_g": rng.integers(-5000, 5000, size=n_cosac * p.d, dtype=np.int64),
            "mlwe": rng.integers(-20, 20, size=n_noise * p.d, dtype=np.int64)}
    if huge:
        for key in ("cosac_g", "mlwe"):
            big = rng.integers(-(1 << 62), 1 << 62, size=tape[key].size, dtype=np.int64)
            tape[key] = np.where(rng.integers(0, 2, size=tape[key].size) == 1, big, tape[key])
        tape["cosac_g"][:4] = [np.iinfo(np.int64).max, np.iinfo(np.int64).min + 1, -int(p.q_primes[0]), int(p.q_primes[0])]
    v = felems(length)
    pd = params_dict(p, ecd)
    exp = jindo_c.commit(fo, pd, ck_in, ck_mlwe, ck_out, cdt_array(ecd), v, tape, threads=8)
    gp = rs.jindo.Prover(fg, pd, ck_in, ck_mlwe, ck_out, cdt_array(ecd))
    B = rs.Batch(fg, 1, N)
    host = np.zeros((1, N, fo.L), dtype=np.uint64); host[0, :length] = v
    B.upload(host)
    if logn >= 16:
        # the same Commit with the mat-vec's TMA loader (cp.async.bulk + mbarrier; opt-in because it measured slower)
        os.environ["RS_MV_TMA"] = "1"
        try:
            gt = rs.jindo.Prover(fg, pd, ck_in, ck_mlwe, ck_out, cdt_array(ecd))
        finally:
            del os.environ["RS_MV_TMA"]
        com_t, _ = gt.Commit(B, 0, length, tape)
        assert np.array_equal(com_t, exp["commitment"])
    com, op = gp.Commit(B, 0, length, tape)
    assert np.array_equal(com, exp["commitment"])
    for k in range(p.in_com_dcmp_len):
        assert np.array_equal(op.InCommit(k), exp["incommit"][k]), ("InCommit", k)
    for i in range(p.cols + 1):
        for j in range(p.rows):
            assert np.array_equal(op.Encode(i, j), exp["encode"][i, j]), ("Encode", i, j)
        for j in range(p.mlwe + p.in_msis):
            assert np.array_equal(op.MLWE(i, j), exp["mlwe"][i, j]), ("MLWE", i, j)


@pytest.mark.parametrize("name,N,batch,length,cuts", [("zp255", 1024, 1, 700, (0, 1, 2)), ("zp128", 700, 2, 700, (0, 2)), ("zp220", 2049, 9, 2049, (0, 1, 4))])
def test_column_sharded_commit_equals_commit(rs, name, N, batch, length, cuts):
    """rs_jindo_commit_columns over a partition of the columns (what the ranks of a sharded Commit run, here one
    after the other into one Opening) + rs_jindo_outer_commit == rs_jindo_commit == oracle, incl. the tape offsets."""
    q, fo, fg, p, tape, prv, gp = _prover_pair(rs, name, N, batch, b"shard-" + name.encode())
    v = [J.field_set_random(tape, q) for _ in range(length)]
    com_o, open_o = prv.commit(v)
    t = tape_arrays(fo, p, prv.trace)
    B = rs.Batch(fg, 1, max(length, 8))
    host = np.zeros((1, max(length, 8), fo.L), dtype=np.uint64); host[0, :length] = fo.to_mont(v)
    B.upload(host)
    bounds = [c for c in cuts if c <= p.cols] + [p.cols + 1]
    op = None
    for a, b in zip(bounds[:-1], bounds[1:]):
        if b > a:
            op = gp.CommitColumns(B, 0, length, t, a, b, op)
    com = gp.OuterCommit(op)
    assert np.array_equal(com, polys_array(com_o))
    for i in range(p.cols + 1):
        for j in range(p.rows):
            assert np.array_equal(op.Encode(i, j), polys_array(open_o.Encode[i][j])), ("Encode", i, j)
    for k in range(p.in_com_dcmp_len):
        assert np.array_equal(op.InCommit(k), polys_array(open_o.InCommit[k])), ("InCommit", k)
    with pytest.raises(rs.RingoError):
        gp.CommitColumns(B, 0, length, t, 2, 2)


@pytest.mark.parametrize("name,N,batch", [("zp255", 1024, 1), ("zp220", 2049, 9), ("zp128", 700, 2), ("zp240", 600, 3)])
def test_commit_key_generated_on_device_matches_new_commit_key(rs, name, N, batch):
    """NewCommitKey (jindo/entities.go:21-61) evaluated in HBM by rs_jindo_create_crs against the oracle's restatement:
    every coefficient of ck.In / ck.MLWE / ck.Out, including the stream shifts caused by SampleN's rejected draws
    (the 52-55-bit primes of the zp220 / zp240 sets reject a few hundred of the ~10^6 draws; the oracle counts them)."""
    q = pyref.modulus(name)
    fg = rs.Field(q)
    p = J.new_parameters(q, N, batch)
    crs = b"Jindo!-" + name.encode()
    ck = J.CommitKey(p, crs)
    ecd = J.Encoder(p, J.Tape(b"k"))
    gp = rs.jindo.Prover(fg, params_dict(p, ecd), None, None, None, cdt_array(ecd), crs=crs, key_from_crs=True)
    g_in, g_mlwe, g_out = gp.CommitKey()
    assert np.array_equal(g_in, polys_array(ck.In))
    assert np.array_equal(g_mlwe, polys_array(ck.MLWE))
    assert np.array_equal(g_out, polys_array(ck.Out))
    # how many draws the reference's SampleN loop threw away on this key (makes the rejection path's coverage visible)
    u = J.UniformSampler(crs)
    n_samples = g_in.size + g_mlwe.size + g_out.size
    mods = [(g_in.shape[0] * g_in.shape[1], p.q_primes), (g_mlwe.shape[0] * g_mlwe.shape[1], p.q_primes),
            (g_out.shape[0] * g_out.shape[1], p.qout_primes)]
    drawn = 0
    for n_polys, primes in mods:
        for _ in range(n_polys * p.d):
            for pr in primes:
                bound = J.MASK64 - J.MASK64 % pr
                while True:
                    drawn += 1
                    if u.sample() < bound:
                        break
    print(f"{name}: {n_samples} key coefficients, {drawn - n_samples} rejected draws")
    if name in ("zp220", "zp240"):
        assert drawn > n_samples


@pytest.mark.parametrize("name,N,batch", [("zp255", 1024, 1), ("zp128", 700, 2), ("zp220", 2049, 9), ("zp240", 600, 3)])
def test_evaluate_vectors_encoded_on_device(rs, name, N, batch):
    """Evaluate's host-side vectors produced by the library: leftVec(x) through Encoder.EncodeTo (jindo/utils.go:63-73,
    encoder.go:113-117) and encodeChallengeTo (utils.go:21-46) over both rings, against the oracle's restatements --
    including challenge bytes whose digits land on both sides of B/2 and the all-0x00 / all-0xff edge cases."""
    q = pyref.modulus(name)
    fo = oracle.Field(q); fg = rs.Field(q)
    p = J.new_parameters(q, N, batch)
    tape = J.Tape(b"eval-vec-" + name.encode())
    prv = J.Prover(p, b"Jindo!", tape)
    gp = rs.jindo.Prover(fg, params_dict(p, prv.ecd), polys_array(prv.ck.In), polys_array(prv.ck.MLWE),
                         polys_array(prv.ck.Out), cdt_array(prv.ecd))
    for x in (J.field_set_random(tape, q), 0, 1, q - 1):
        want = polys_array([prv.ecd.encode([le]) for le in J.left_vec(p, x)])
        got = gp.LeftVec(fo.to_mont([x])[0])
        assert np.array_equal(got, want), ("leftVec", x)
    rng = np.random.default_rng(5)
    chal = [bytes(16), b"\xff" * 16] + [rng.bytes(16) for _ in range(6)]
    blob = b"".join(chal)
    for ring_out, ring in ((False, p.ringQ), (True, p.ringQOut)):
        want = polys_array([J.encode_challenge(p, ring, cb) for cb in chal])
        got = gp.EncodeChallenges(blob, ring_out)
        assert np.array_equal(got, want), ("challenge", ring_out)


@pytest.mark.parametrize("env", ["RS_JINDO_NOSLOWLB", "RS_JINDO_NOF64", "RS_JINDO_NOPAIR", "RS_JINDO_NOGUIDE"])
def test_commit_parity_with_alternative_kernel_paths(rs, monkeypatch, env):
    """encode_kernel keeps its earlier code paths behind environment switches (read when the prover is created / at every
    launch): no slow-path lower bound (every slow sample takes the CDF sum), integer instead of FP64 warp NTT, one digit per
    long division, plain binary search instead of the guide table.  Each must give the oracle's Opening as well."""
    monkeypatch.setenv(env, "1")
    test_commit_and_evaluate_parity(rs, "zp255", 1024, 1, 1024)
    test_commit_and_evaluate_parity(rs, "zp220", 2049, 9, 2049)


def test_device_mlwe_sampler_matches_oracle(rs):
    """Prover.mlweSampler (TwinCDT, centre 0) on the device: the AES-CTR stream keyed by the sampler's seed and the
    binary search in table 0 give the oracle's Sample(0) sequence, across calls (the stream is stateful)."""
    q = pyref.modulus("zp255")
    fg = rs.Field(q)
    p = J.new_parameters(q, 1024, 1)
    tape = J.Tape(b"mlwe-stream")
    prv = J.Prover(p, b"Jindo!", tape)
    gp = rs.jindo.Prover(fg, params_dict(p, prv.ecd), polys_array(prv.ck.In), polys_array(prv.ck.MLWE),
                         polys_array(prv.ck.Out), cdt_array(prv.ecd))
    seed = prv.mlwe.base.seed
    gp.SetMlweSeed(seed, p.mlwe_std)
    ref = J.TwinCDTGaussianSampler(p.mlwe_std, seed)
    for n in (1, 1023, 5000, 3 * 8192 + 5):
        got = gp.MlweStreamSample(n)
        assert got.tolist() == [ref.sample(0) for _ in range(n)], n
    # exact-hit rule of slices.BinarySearch (gaussian_twin_cdt.go:86-89): a draw equal to a table entry steps one back;
    # checked on the table itself through a host-built sampler with the same table
    tab, lo = rs.jindo.compute_cdt(0.0, p.mlwe_std)
    assert [int(x) for x in tab] == ref.tables[0] and lo == ref.tail_lo


@pytest.mark.parametrize("name,N,batch,length", [("zp255", 1024, 1, 1024), ("zp255", 1024, 1, 600), ("zp128", 700, 2, 700), ("zp240", 600, 3, 513)])
def test_commit_from_seeds_only(rs, name, N, batch, length):
    """The end-to-end form of Commit: nothing but the vector, the sampler seeds and the crypto/rand bytes go in.  The
    host half of the tape comes from the product's tape builder (COSAC rows, mask MLWE, MustSetRandom elements), the two
    TwinCDT samplers (encoder rows, mlweSampler) run on the device from their seeds -- and the Opening and Commitment
    are the oracle's, bit for bit, for two Commits in a row (every stream continues where it stopped)."""
    q = pyref.modulus(name)
    fo = oracle.Field(q); fg = rs.Field(q)
    sd = b"seeds-only-" + name.encode()
    p = J.new_parameters(q, N, batch)
    tape = J.Tape(sd)
    prv = J.Prover(p, b"Jindo!", tape)
    pd = params_dict(p, prv.ecd)
    gp = rs.jindo.Prover(fg, pd, polys_array(prv.ck.In), polys_array(prv.ck.MLWE), polys_array(prv.ck.Out), cdt_array(prv.ecd))
    gp.SetTwinSeed(prv.ecd.twin.base.seed)
    gp.SetMlweSeed(prv.mlwe.base.seed, p.mlwe_std)
    noise = {"ecd_blind_std": p.ecd_blind_std, "mask_blind_std": p.mask_blind_std, "mask_std": p.mask_std, "mask_mlwe_std": p.mask_mlwe_std}
    tb = rs.jindo.TapeBuilder(pd, len(prv.ecd.twin.tables[0]), noise, q, threads=1,
                              seeds=[prv.ecd.cosac.base.seed, prv.ecd.cosac.rounded.base.seed, prv.rounded.base.seed])
    rng = np.random.default_rng(N + length)
    for it in range(2):
        v = [int.from_bytes(rng.bytes(40), "little") % q for _ in range(length)]
        pos0 = tape._pos
        com_o, open_o = prv.commit(v)
        clone = J.Tape(sd); clone.read(pos0)
        tb.set_random_bytes(clone.read(tape._pos - pos0))
        vm = fo.to_mont(v)
        tb.start(vm, length, it)
        T, _ = tb.wait(it)
        B = rs.Batch(fg, 1, max(length, 8))
        host = np.zeros((1, max(length, 8), fo.L), dtype=np.uint64); host[0, :length] = vm
        B.upload(host)
        com_g, open_g = gp.Commit(B, 0, length, T)
        assert np.array_equal(com_g, polys_array(com_o)), it
        for i in range(p.cols + 1):
            for j in range(p.rows):
                assert np.array_equal(open_g.Encode(i, j), polys_array(open_o.Encode[i][j])), ("Encode", it, i, j)
            for j in range(p.mlwe + p.in_msis):
                assert np.array_equal(open_g.MLWE(i, j), polys_array(open_o.MLWE[i][j])), ("MLWE", it, i, j)


def test_wire_format_through_the_abi(rs):
    """rs_jindo_commitment_write_raw / rs_jindo_poly_write_to give the bytes of the restated Commitment.WriteRawTo /
    ring.Poly.WriteTo (entities.go:96-100; the ringQ limb-count quirk of NewCommitment included: zp255 at N = 1024 has
    2 ringQ limbs and 1 ringQOut limb)."""
    import ctypes
    q, fo, fg, p, tape, prv, gp = _prover_pair(rs, "zp255", 1024, 1, b"wire")
    assert len(p.q_primes) == 2 and len(p.qout_primes) == 1
    v = [J.field_set_random(tape, q) for _ in range(100)]
    com_o, _ = prv.commit(v)
    B = rs.Batch(fg, 1, 128)
    host = np.zeros((1, 128, fo.L), dtype=np.uint64); host[0, :100] = fo.to_mont(v)
    B.upload(host)
    com_g, _ = gp.Commit(B, 0, 100, tape_arrays(fo, p, prv.trace))
    assert gp.CommitmentWriteRaw(com_g) == J.commitment_write_raw(p, com_o)
    poly = np.ascontiguousarray(polys_array(prv.ck.In[0][0]))
    n = rs.lib().rs_jindo_poly_write_to(poly.ctypes.data_as(ctypes.c_void_p), poly.shape[0], poly.shape[1], None, 0)
    buf = ctypes.create_string_buffer(int(n))
    assert rs.lib().rs_jindo_poly_write_to(poly.ctypes.data_as(ctypes.c_void_p), poly.shape[0], poly.shape[1], ctypes.cast(buf, ctypes.c_void_p), int(n)) == n
    assert buf.raw == J.poly_write_to(prv.ck.In[0][0])


def _host_cosac_expected(rs, p, pd, q, L, keys, noise, first_row, mask_rows, worker0):
    """What the device COSAC sampler must produce: for worker w = worker0 + 256 t + c a fresh host COSACSampler
    (rs_cosac_create: csprng.COSACSampler restated, pinned to the oracle in test_host_sampler.py) keyed like the device's --
    the stream's two keys, counters IV + (w << 16) -- asked for Sample(-fp[c], stdDev) once, with the base-b digits and float64
    centres of encoder.go:120-164 computed here in plain Python floats (IEEE doubles, unfused, same order)."""
    import ctypes
    lib = rs.lib()
    kb, kr = keys
    R_inv = pow(1 << (64 * L), -1, q)
    slots, exp, base, cols, N = p.slots, p.exp, p.base, p.cols, 256
    n_mask = len(mask_rows) // slots
    out = np.zeros((cols + n_mask, N), dtype=np.int64)
    ivb, ivr = int.from_bytes(kb[32:], "big"), int.from_bytes(kr[32:], "big")
    lib.rs_cosac_sample.restype = ctypes.c_int64
    for t in range(cols + n_mask):
        rows = first_row[t * slots:(t + 1) * slots] if t < cols else mask_rows[(t - cols) * slots:(t - cols + 1) * slots]
        digits = [0] * N
        for i, limbs in enumerate(rows):
            c = sum(int(v) << (64 * k) for k, v in enumerate(limbs)) * R_inv % q
            for j in range(exp - 1):
                digits[j * slots + i] = c % base
                c //= base
            digits[(exp - 1) * slots + i] = c & ((1 << 64) - 1)
        fp = [0.0] * N
        for i in range(exp):
            di = pd["delta_inv"][i]
            if di == 0:
                continue
            dd = N - (i + 1) * slots
            for j, jj in zip(range(0, N - dd), range(dd, N)):
                fp[jj] += di * float(digits[j])
            for j, jj in zip(range(N - dd, N), range(0, dd)):
                fp[jj] -= di * float(digits[j])
        std = noise["ecd_blind_std"] if t < cols else (noise["mask_blind_std"] if t == cols else noise["mask_std"])
        for c in range(N):
            w = worker0 + 256 * t + c
            kbw = kb[:32] + ((ivb + (w << 16)) % (1 << 128)).to_bytes(16, "big")
            krw = kr[:32] + ((ivr + (w << 16)) % (1 << 128)).to_bytes(16, "big")
            h = lib.rs_cosac_create(ctypes.c_char_p(kbw), ctypes.c_char_p(krw))
            out[t, c] = lib.rs_cosac_sample(h, ctypes.c_double(-fp[c]), ctypes.c_double(std))
            lib.rs_cosac_destroy(h)
    return out


@pytest.mark.parametrize("name,N", [("zp255", 1024), ("zp128", 700)])
def test_device_cosac_equals_the_host_samplers(rs, name, N):
    """rs_jindo_set_cosac_stream: the encoder's COSAC rows sampled on the device, one fresh sampler per coefficient.  Every sample
    equals what the host restatement of csprng.COSACSampler returns for the same keys / counters / centre / stdDev; a Commit that
    draws them on the device equals the Commit that is handed the same samples on its tape."""
    q = pyref.modulus(name)
    fo = oracle.Field(q); fg = rs.Field(q)
    p = J.new_parameters(q, N, 1)
    tape = J.Tape(b"device-cosac-" + name.encode())
    prv = J.Prover(p, b"Jindo!", tape)
    pd = params_dict(p, prv.ecd)
    gp = rs.jindo.Prover(fg, pd, polys_array(prv.ck.In), polys_array(prv.ck.MLWE), polys_array(prv.ck.Out), cdt_array(prv.ecd))
    noise = {"ecd_blind_std": p.ecd_blind_std, "mask_blind_std": p.mask_blind_std, "mask_std": p.mask_std, "mask_mlwe_std": p.mask_mlwe_std}
    gp.SetCosacSeeds(b"cosac-base-seed", b"cosac-rounded-seed", noise)
    rng = np.random.default_rng(11)
    rnd_elems = lambda n: fo.to_mont([int.from_bytes(rng.bytes(40), "little") % q for _ in range(n)])
    first = rnd_elems(p.cols * p.slots)
    mask = rnd_elems(p.rows * p.slots)
    # small values too: the digits above the first are zero, the centre is tiny
    first[0] = fo.to_mont([0])[0]; first[1] = fo.to_mont([1])[0]; mask[0] = fo.to_mont([q - 1])[0]
    for rep in range(2):                                  # the worker counter runs on between calls
        g, w0, ov = gp.CosacSample(first, mask)
        assert ov == 0 and w0 == rep * (p.cols + p.rows) * 256
        want = _host_cosac_expected(rs, p, pd, q, fo.L, gp._cosac_keys, noise, first, mask, w0)
        assert np.array_equal(g, want), (rep, np.argwhere(g != want)[:5])
    assert abs(float(g[p.cols + 1:].std()) - p.mask_std) < 0.1 * p.mask_std      # sanity: the spread is the requested one
    # ---- a whole Commit: cosac_g = NULL on the tape against the same samples supplied
    v = [int.from_bytes(rng.bytes(40), "little") % q for _ in range(N)]
    com_o, open_o = prv.commit(v)
    t_ex = tape_arrays(fo, p, prv.trace)
    B = rs.Batch(fg, 1, max(N, 8))
    host = np.zeros((1, max(N, 8), fo.L), dtype=np.uint64); host[0, :N] = fo.to_mont(v)
    B.upload(host)
    w0 = 2 * (p.cols + p.rows) * 256
    n_mask_rows = t_ex["mask_rows"].shape[0] // p.slots
    want = _host_cosac_expected(rs, p, pd, q, fo.L, gp._cosac_keys, noise, t_ex["first_row"], t_ex["mask_rows"], w0)
    t_dev = dict(t_ex); t_dev["cosac_g"] = None
    com_d, op_d = gp.Commit(B, 0, N, t_dev)
    t_sup = dict(t_ex); t_sup["cosac_g"] = want.reshape(-1)
    assert want.shape[0] == p.cols + n_mask_rows
    com_s, op_s = gp.Commit(B, 0, N, t_sup)
    assert np.array_equal(com_d, com_s)
    for i in range(p.cols + 1):
        for jrow in (0, 1, p.rows - 1):
            assert np.array_equal(op_d.Encode(i, jrow), op_s.Encode(i, jrow)), (i, jrow)
    assert not np.array_equal(com_d, polys_array(com_o))          # (different COSAC draws than the oracle's own tape)
