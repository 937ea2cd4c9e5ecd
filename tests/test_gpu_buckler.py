"""GPU parity for the Buckler prover's polynomial pipeline (buckler/prover.go), composed exactly as
Prove composes it, against the oracle's composition of the same steps:

  Encode            cyclic INTT at rank N into an embRank poly, zero padded   (buckler/encoder.go:32-38)
  RandEncode        + blinding coefficient N                                  (encoder.go:50-54)
  polyEval.NTT      cyclic NTT at embRank                                     (prover.go:128,150)
  arithCheck        evalCircuit -> InvNTT -> QuoRemByVanishing                (prover.go:355-379, :399-404)
  linCheck pieces   nttChecker.TransposeTo (reverse*N + negacyclic INTT), MulTo / MulSubTo / ScalarMulTo / AddTo
                                                                              (linear.go:38-43, prover.go:406-441)
MultCircuit constraint set of examples/mult/main.go:35-65 (N = 2^11, embRank = 4N, 128-bit field)."""
import numpy as np
import pytest

import oracle
from oracle import pyref

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rs(product):
    L = product.load()
    assert L.rs_init(0) == 0, L.rs_last_error()
    return product


def rand_vals(q, n, seed, small=None):
    rng = np.random.default_rng(seed)
    if small is not None:
        return [int(x) for x in rng.integers(0, small, size=n)]
    return [int.from_bytes(rng.bytes(24), "little") % q for _ in range(n)]


@pytest.mark.parametrize("name,logN", [("zp128", 11), ("zp240", 10)])
def test_prover_pipeline(rs, name, logN):
    q = pyref.modulus(name)
    fo = oracle.Field(q); fg = rs.Field(q)
    N = 1 << logN
    E = 4 * N                                            # embRank = nextpow2(3N + 1) (compile.go:180)
    lib = rs.lib()
    u64p = rs.capi.c_u64p
    # witnesses: X in [0,5], Y 63-bit, Z = X*Y, ternary digits of X (decomposeBase(5) = [3,1,1])
    X = rand_vals(q, N, 1, small=6); Y = rand_vals(q, N, 2, small=1 << 62)
    Z = [(a * b) % q for a, b in zip(X, Y)]
    dig = {0: (0, 0, 0), 1: (0, 1, 0), 2: (0, 1, 1), 3: (1, 0, 0), 4: (1, 1, 0), 5: (1, 1, 1)}
    W = [X, Y, Z, rand_vals(q, N, 3)] + [[dig[x][k] for x in X] for k in range(3)]
    PW = [rand_vals(q, N, 4)]
    blind = rand_vals(q, len(W), 5)                      # MustSetRandom of RandEncodeTo (host randomness)
    n_w = len(W)
    # ---- oracle composition
    enc_o = oracle.NTT(fo, logN, False); emb_o = oracle.NTT(fo, logN + 2, False)
    def encode_o(v, r=None):
        c = np.zeros((E, fo.L), dtype=np.uint64)
        c[:N] = enc_o.inv(fo.to_mont(v))
        if r is not None:
            c[N] = fo.to_mont([r])[0]
            c[0] = fo.vec_sub(c[0:1], c[N:N + 1])[0]
        return c
    w_ecd_o = [encode_o(W[i], blind[i]) for i in range(n_w)]
    w_ntt_o = [emb_o.fwd(c) for c in w_ecd_o]
    pw_ntt_o = [emb_o.fwd(encode_o(PW[0]))]
    coeffs = fo.to_mont([1, q - 1, q - 3])
    terms = [(0, 0, -1, [1, 0]), (0, 1, -1, [2]),                                   # Y*X - Z
             (1, 0, -1, [4, 4, 4]), (1, 1, -1, [4]), (2, 0, -1, [5, 5, 5]), (2, 1, -1, [5]),
             (3, 0, -1, [6, 6, 6]), (3, 1, -1, [6]),                                 # w^3 - w
             (4, 0, -1, [0]), (4, 2, -1, [4]), (4, 1, -1, [5]), (4, 1, -1, [6]),     # X - 3w0 - w1 - w2
             (5, 0, 0, [3])]                                                         # pw * w3
    bc = fo.to_mont(rand_vals(q, 1, 6))[0]
    ev_o = fo.eval_circuit(pw_ntt_o, w_ntt_o, terms, coeffs, bc, E)
    quo_o, rem_o = fo.quorem_vanishing(emb_o.inv(ev_o), N)
    # ---- GPU composition through the C ABI
    enc_g = rs.CyclicTransformer(fg, N); emb_g = rs.CyclicEvaluator(fg, E)
    vals = rs.Batch(fg, n_w + 1, N).upload(np.stack([fo.to_mont(w) for w in W] + [fo.to_mont(PW[0])]))
    ecd = rs.Batch(fg, n_w + 1, E)
    assert lib.rs_ntt_inverse_embed(enc_g.h, ecd.h, 0, vals.h, 0, n_w + 1, 1) == 0, lib.rs_last_error()
    # RandEncodeTo: c[N] <- r ; c[0] -= c[N]   (one-element uploads / vec ops on 1-coefficient views)
    one = rs.Batch(fg, 2, 1)
    for i in range(n_w):
        one.upload(np.stack([fo.to_mont([blind[i]]), fo.to_mont([0])]).reshape(2, 1, fo.L))
        assert lib.rs_batch_copy(ecd.h, i, N, one.h, 0, 0, 1) == 0
        assert lib.rs_batch_copy(one.h, 1, 0, ecd.h, i, 0, 1) == 0
        assert lib.rs_vec_sub(one.h, 1, one.h, 1, one.h, 0, 1) == 0
        assert lib.rs_batch_copy(ecd.h, i, 0, one.h, 1, 0, 1) == 0
    got_ecd = ecd.download()
    for i in range(n_w):
        assert np.array_equal(got_ecd[i], w_ecd_o[i]), ("RandEncode", i)
    nt = rs.Batch(fg, n_w + 1, E)
    emb_g.ntt.FwdNTTTo(nt, ecd)
    got_ntt = nt.download()
    for i in range(n_w):
        assert np.array_equal(got_ntt[i], w_ntt_o[i]), ("NTT", i)
    assert np.array_equal(got_ntt[n_w], pw_ntt_o[0])
    wb = rs.Batch(fg, n_w, E, device_ptr=lib.rs_batch_device_ptr(nt.h), keepalive=nt)
    pwb = rs.Batch(fg, 1, E); assert lib.rs_batch_copy(pwb.h, 0, 0, nt.h, n_w, 0, E) == 0
    T = (rs.capi.Term * len(terms))()
    for i, (c, co, p, ws) in enumerate(terms):
        T[i].constraint, T[i].coeff, T[i].pw, T[i].n_w = c, co, p, len(ws)
        for k, x in enumerate(ws):
            T[i].w[k] = x
    work = rs.Batch(fg, 3, E)
    assert lib.rs_eval_circuit(work.h, 0, pwb.h, wb.h, T, len(terms), coeffs.ctypes.data_as(u64p), len(coeffs),
                               bc.ctypes.data_as(u64p)) == 0, lib.rs_last_error()
    assert np.array_equal(work.download(0, 1)[0], ev_o)
    ev = work.poly(0, is_ntt=True)
    emb_g.InvNTTTo(ev, ev)
    quo, rem = emb_g.QuoRemByVanishing(ev, N, work.poly(1), work.poly(2))
    assert np.array_equal(quo.Coeffs, quo_o) and np.array_equal(rem.Coeffs, rem_o)
    # the quotient of an honest witness has degree < arithCheckMaxRank - N = 2N + 1 and rem == 0 on the mult constraint part
    # ---- linCheck pieces: nttChecker.TransposeTo + the per-constraint MulTo / MulSubTo / ScalarMulTo / AddTo
    neg_o = oracle.NTT(fo, logN, True); neg_g = rs.CyclotomicTransformer(fg, N)
    lin = fo.to_mont(rand_vals(q, N, 9))
    scale = fo.to_mont([N])[0]
    tr_o = neg_o.inv(fo.vec_scalar_mul(lin[::-1].copy(), scale))
    LB = rs.Batch(fg, 2, N).upload(np.stack([lin, lin]))
    assert lib.rs_reverse_scale(LB.h, 1, LB.h, 0, scale.ctypes.data_as(u64p)) == 0
    neg_g.InvNTTTo(LB, LB, dst_poly=1, src_poly=1, n_polys=1)
    assert np.array_equal(LB.download(1, 1)[0], tr_o)
    a, b, c3 = w_ntt_o[0], w_ntt_o[1], w_ntt_o[2]
    term_o = fo.vec_sub(fo.vec_mul(a, b), fo.vec_mul(c3, a))
    eval_o = fo.vec_add(fo.vec_scalar_mul(term_o, bc), term_o)
    P = [work.poly(i, is_ntt=True) for i in range(3)]
    A_, B_, C_ = nt.poly(0, True), nt.poly(1, True), nt.poly(2, True)
    emb_g.MulTo(P[0], A_, B_); emb_g.MulSubTo(P[0], C_, A_)
    emb_g.ScalarMulTo(P[1], P[0], bc); emb_g.AddTo(P[1], P[1], P[0])
    assert np.array_equal(P[1].Coeffs, eval_o)


@pytest.mark.parametrize("native", [False, True], ids=["python-mirror", "compiled"])
@pytest.mark.parametrize("circuit,name,N", [("mult", "zp128", 256), ("pk", "zp220", 256), ("ct", "zp240", 256), ("norm", "zp220", 256)])
def test_full_prove_bit_exact_and_verifies(rs, circuit, name, N, native):
    """buckler.Prover.Prove end to end on the GPU (ringo-snark_b200/buckler.py: host transcript, device
    polynomials, Jindo commits/evaluation through the C ABI) against the oracle's Prove on the same
    circuit, witness, CRS and randomness tape: every commitment, evaluation and proof polynomial is
    bit-identical, and the restated buckler.Verifier accepts the GPU proof (buckler_test.go:82-99).
    native: the same Prove as one call into the compiled prover (rs_buckler_prove, include/ringo_buckler.h)."""
    import sys, os
    sys.path.insert(0, os.path.dirname(__file__))
    from test_oracle_buckler import circuit_and_witness
    from oracle import buckler_ref as B, jindo_ref as J
    from jindo_util import cdt_array, params_dict, polys_array, tape_arrays
    q = pyref.modulus(name)
    fo = oracle.Field(q); fg = rs.Field(q)
    (define, n_pw, n_w), (pw, w) = circuit_and_witness(circuit, q, N, 11)    # "ct": examples/bfv CiphertextCircuit (BASELINE configs[3])
    tape = J.Tape(b"buckler-gpu-" + circuit.encode())
    C = B.compile_circuit(q, N, define, n_pw, n_w, b"Buckler!", tape)
    op = B.Prover(C, tape)
    pf_o = op.prove(pw, w)
    # ---- the same statement for the GPU prover: plain data a Go caller would pass
    ctx = C.ctx
    terms, coeffs = ctx.terms(ctx.arith_constraints)
    desc = {"rank": N, "emb_rank": C.emb_rank, "w_cnt": ctx.w_cnt, "pw_cnt": ctx.pw_cnt, "arith_max_rank": ctx.arith_max_rank,
            "arith_terms": terms, "arith_coeffs": coeffs,
            "lin": [("ntt" if isinstance(chk, B.NTTChecker) else ("aut", chk.idx_inv, chk.is_ntt), ctx.lin_constraints[id(chk)])
                    for chk in ctx.lin_checkers]}
    if ctx.has_sum():                                # sum-check rounds ("norm": automorphism + two-norm + explicit sum-check)
        desc["sum_terms"], desc["sum_coeffs"] = ctx.terms(ctx.sumcheck_constraints)
        desc["sum_max_rank"] = ctx.sumcheck_max_rank
    prv = C.poly_prover
    gp = rs.jindo.Prover(fg, params_dict(C.jp, prv.ecd), polys_array(prv.ck.In), polys_array(prv.ck.MLWE),
                         polys_array(prv.ck.Out), cdt_array(prv.ecd), crs=b"Buckler!")
    rand = {"blind": op.trace["blind"], "mask": op.trace["mask"], "mask_sum": op.trace["mask_sum"],
            "tapes": [tape_arrays(fo, C.jp, t) for t in op.trace["commit_traces"]]}
    gprv = (rs.buckler.NativeProver if native else rs.buckler.Prover)(fg, q, desc, C.jp, gp)
    pf_g = gprv.prove(np.stack([fo.to_mont(v) for v in op.trace["pw_full"]]),
                      np.stack([fo.to_mont(v) for v in op.trace["w_full"]]), rand)
    assert pf_g["_eval_point"] == pf_o["_eval_point"]
    assert pf_g["Witness"] == [[[list(map(int, l)) for l in poly] for poly in cm] for cm in pf_o["Witness"]]
    assert pf_g["Evals"] == pf_o["Evals"] and pf_g["LinCheckMaskSum"] == pf_o["LinCheckMaskSum"]
    assert pf_g["SumCheckMaskSum"] == pf_o["SumCheckMaskSum"]
    for k in ("Partial", "Encode", "MLWE", "InCommit"):
        assert pf_g["EvalProof"][k] == [[list(map(int, l)) for l in poly] for poly in pf_o["EvalProof"][k]], k
    assert pf_g["EvalProof"]["PartialMask"] == [list(map(int, l)) for l in pf_o["EvalProof"]["PartialMask"]]
    assert B.Verifier(C).verify(pw, pf_g)
    if native:
        # the compiled Verifier (rs_buckler_verify) takes the same decisions as the restated one
        pwm = np.stack([fo.to_mont(v) for v in op.trace["pw_full"]]) if ctx.pw_cnt else np.zeros((0, N, fo.L), dtype=np.uint64)
        raw = gprv.last_raw
        assert gprv.verify_raw(pwm, raw) == 0
        import copy
        t1 = copy.deepcopy(raw); t1["Evals"][0, 0] ^= np.uint64(1)                 # a wrong evaluation: polyVerifier.Verify rejects
        assert gprv.verify_raw(pwm, t1) == 1
        t2 = copy.deepcopy(raw); t2["Encode"][1, 0, 5] = (int(t2["Encode"][1, 0, 5]) + 1) % C.jp.q_primes[0]
        assert gprv.verify_raw(pwm, t2) == 1
        if ctx.has_lin():
            t3 = copy.deepcopy(raw); t3["lin_mask_sum"][0] ^= np.uint64(1)         # re-derives every later challenge
            assert gprv.verify_raw(pwm, t3) != 0
        if ctx.pw_cnt:
            bad_pw = pwm.copy(); bad_pw[0, 2] = fo.to_mont([777])[0]               # a different statement
            assert gprv.verify_raw(bad_pw, raw) in (2, 4, 6)


@pytest.mark.parametrize("circuit,name,N", [("mult", "zp128", 256), ("norm", "zp220", 256)])
def test_self_sampled_prove_verifies(rs, circuit, name, N):
    """rs_buckler_prove with rand == NULL: every random value is drawn inside the library (host samplers on worker threads +
    the device TwinCDT / MLWE streams), as the reference's Prove does with crypto/rand.  No trace to compare with, so the
    check is the reference's own (buckler_test.go:82-99): the restated Verifier accepts, two proofs of the same statement
    differ, and a proof for a wrong witness is rejected."""
    import os, sys
    sys.path.insert(0, os.path.dirname(__file__))
    from test_oracle_buckler import circuit_and_witness
    from oracle import buckler_ref as B, jindo_ref as J
    from jindo_util import cdt_array, params_dict, polys_array
    q = pyref.modulus(name)
    fo = oracle.Field(q); fg = rs.Field(q)
    (define, n_pw, n_w), (pw, w) = circuit_and_witness(circuit, q, N, 23)
    tape = J.Tape(b"buckler-self-" + circuit.encode())
    C = B.compile_circuit(q, N, define, n_pw, n_w, b"Buckler!", tape)
    op = B.Prover(C, tape)
    op.prove(pw, w)                                   # only for the completed witness vectors (decomposition witnesses)
    ctx = C.ctx
    terms, coeffs = ctx.terms(ctx.arith_constraints)
    desc = {"rank": N, "emb_rank": C.emb_rank, "w_cnt": ctx.w_cnt, "pw_cnt": ctx.pw_cnt, "arith_max_rank": ctx.arith_max_rank,
            "arith_terms": terms, "arith_coeffs": coeffs,
            "lin": [("ntt" if isinstance(chk, B.NTTChecker) else ("aut", chk.idx_inv, chk.is_ntt), ctx.lin_constraints[id(chk)])
                    for chk in ctx.lin_checkers]}
    if ctx.has_sum():
        desc["sum_terms"], desc["sum_coeffs"] = ctx.terms(ctx.sumcheck_constraints)
        desc["sum_max_rank"] = ctx.sumcheck_max_rank
    prv = C.poly_prover
    p = C.jp
    gp = rs.jindo.Prover(fg, params_dict(p, prv.ecd), polys_array(prv.ck.In), polys_array(prv.ck.MLWE),
                         polys_array(prv.ck.Out), cdt_array(prv.ecd), crs=b"Buckler!")
    gp.SetTwinSeed(os.urandom(32))
    gp.SetMlweSeed(os.urandom(32), p.mlwe_std)
    nprv = rs.buckler.NativeProver(fg, q, desc, p, gp)
    nprv.enable_sampling({"ecd_blind_std": p.ecd_blind_std, "mask_blind_std": p.mask_blind_std, "mask_std": p.mask_std,
                          "mask_mlwe_std": p.mask_mlwe_std}, threads=3)
    pwm = np.stack([fo.to_mont(v) for v in op.trace["pw_full"]])
    wm = np.stack([fo.to_mont(v) for v in op.trace["w_full"]])
    pf1 = nprv.prove(pwm, wm, None)
    raw1 = nprv.last_raw
    pf2 = nprv.prove(pwm, wm, None)
    V = B.Verifier(C)
    assert V.verify(pw, pf1) and V.verify(pw, pf2)
    assert nprv.verify_raw(pwm, raw1) == 0 and nprv.verify_raw(pwm, nprv.last_raw) == 0     # the compiled Verifier agrees
    assert pf1["Witness"] != pf2["Witness"] and pf1["_eval_point"] != pf2["_eval_point"]
    bad = wm.copy()
    bad[0, 3] = fo.to_mont([12345])[0]                # breaks a constraint on one coefficient
    pf_bad = nprv.prove(pwm, bad, None)
    assert not V.verify(pw, pf_bad)
    assert nprv.verify_raw(pwm, nprv.last_raw) != 0
    # the same with the COSAC rows drawn by the device sampler (rs_jindo_set_cosac_stream): both Verifiers still accept
    gp.SetCosacSeeds(os.urandom(32), os.urandom(32), {"ecd_blind_std": p.ecd_blind_std, "mask_blind_std": p.mask_blind_std, "mask_std": p.mask_std})
    nprv.set_device_cosac(True)
    pf3 = nprv.prove(pwm, wm, None)
    assert V.verify(pw, pf3) and nprv.verify_raw(pwm, nprv.last_raw) == 0
    assert gp.CosacStatus()[1] == 0 and gp.CosacStatus()[0] > 0


def test_compiled_prove_at_the_example_size_from_the_cached_trace(rs):
    """examples/mult at its default rank N = 2^13 (BASELINE configs[0]): the compiled prover's proof against the oracle's proof
    cached in oracle_trace/mult13.pkl (written once by tests/perf/bench_buckler.py --trace: 93 s of pure-Python oracle, so the
    test only runs where the cache exists), every polynomial bit for bit, and the compiled Verifier accepts it."""
    import os, pickle, sys
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle_trace", "mult13.pkl")
    if not os.path.exists(path):
        pytest.skip("oracle_trace/mult13.pkl not present (python tests/perf/bench_buckler.py 13 mult --trace=oracle_trace/mult13.pkl --oracle-only)")
    sys.path.insert(0, os.path.dirname(__file__))
    from oracle import buckler_ref as B, jindo_ref as J
    from jindo_util import cdt_array, params_dict
    with open(path, "rb") as fh:
        saved = pickle.load(fh)
    assert saved["key"] == (13, "mult", "zp128")
    rand, pwm, wm = saved["packed"]
    pf_o = saved["proof"]
    N = 1 << 13
    q = pyref.modulus("zp128")
    fg = rs.Field(q)
    desc, commit_rank, batch = rs.circuits.mult_circuit(q, N)                      # product-side compiled form of the circuit
    p = J.new_parameters(q, commit_rank, batch)
    ecd = J.Encoder(p, J.Tape(b"unused"))
    gp = rs.jindo.Prover(fg, params_dict(p, ecd), None, None, None, cdt_array(ecd), crs=b"Buckler!", key_from_crs=True)
    nprv = rs.buckler.NativeProver(fg, q, desc, p, gp)
    pf = nprv.prove(pwm, wm, rand)
    assert pf["_eval_point"] == pf_o["_eval_point"] and pf["Evals"] == pf_o["Evals"] and pf["LinCheckMaskSum"] == pf_o["LinCheckMaskSum"]
    assert pf["Witness"] == [[[list(map(int, l)) for l in poly] for poly in cm] for cm in pf_o["Witness"]]
    for k in ("Partial", "Encode", "MLWE", "InCommit"):
        assert pf["EvalProof"][k] == [[list(map(int, l)) for l in poly] for poly in pf_o["EvalProof"][k]], k
    assert pf["EvalProof"]["PartialMask"] == [list(map(int, l)) for l in pf_o["EvalProof"]["PartialMask"]]
    assert nprv.verify_raw(pwm, nprv.last_raw) == 0
    assert nprv.timing["total"] < 500.0                                             # milliseconds, with the phase synchronisations (6-8 ms on a B200)


def test_compiled_prover_reproduces_the_committed_digest(rs):
    """tests/golden/buckler_proof_digests.json holds SHA-256 of the restated prover's proof of the seeded MultCircuit statement
    (written by tests/golden/make_buckler_digests.py, checked on the CPU by test_oracle_buckler.py).  The compiled GPU prover, fed
    the same witness and randomness, must produce a proof with that digest."""
    import json, os, sys
    sys.path.insert(0, os.path.dirname(__file__)); sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    from make_buckler_digests import proof_digest
    from test_oracle_buckler import circuit_and_witness
    from oracle import buckler_ref as B, jindo_ref as J
    from jindo_util import cdt_array, params_dict, polys_array, tape_arrays
    want = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "buckler_proof_digests.json")))["mult"]
    q = pyref.modulus(want["field"]); N = want["rank"]
    fo = oracle.Field(q); fg = rs.Field(q)
    (define, n_pw, n_w), (pw, w) = circuit_and_witness("mult", q, N, want["witness_seed"])
    tape = J.Tape(want["tape_seed"].encode())
    C = B.compile_circuit(q, N, define, n_pw, n_w, want["crs"].encode(), tape)
    op = B.Prover(C, tape)
    op.prove(pw, w)                                   # for the randomness trace
    desc, commit_rank, batch = rs.circuits.mult_circuit(q, N)
    prv = C.poly_prover
    gp = rs.jindo.Prover(fg, params_dict(C.jp, prv.ecd), polys_array(prv.ck.In), polys_array(prv.ck.MLWE), polys_array(prv.ck.Out),
                         cdt_array(prv.ecd), crs=want["crs"].encode())
    rand = {"blind": op.trace["blind"], "mask": op.trace["mask"], "mask_sum": op.trace["mask_sum"],
            "tapes": [tape_arrays(fo, C.jp, t) for t in op.trace["commit_traces"]]}
    nprv = rs.buckler.NativeProver(fg, q, desc, C.jp, gp)
    pf = nprv.prove(np.stack([fo.to_mont(v) for v in op.trace["pw_full"]]), np.stack([fo.to_mont(v) for v in op.trace["w_full"]]), rand)
    assert proof_digest(pf, fo.L) == want["sha256"]
