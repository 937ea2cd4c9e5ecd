"""CPU test of the Buckler oracle: the reference's own check (buckler/buckler_test.go:82-99),
Verify(Prove(circuit)) == true, on the restatement -- for PublicKeyCircuit (the reference's test
circuit, zp220) and MultCircuit (examples/mult, the 128-bit field), plus a rejected bad witness."""
import random

import numpy as np
import pytest

import oracle
from oracle import buckler_ref as B, jindo_ref as J, pyref


def mult_witness(q, N, seed):
    f = oracle.Field(q)
    rng = random.Random(seed)
    ntt = oracle.NTT(f, N.bit_length() - 1, True)
    X = [rng.randrange(6) for _ in range(N)]                       # |X| <= 5 (examples/mult/main.go:75-78)
    Y = [rng.getrandbits(63) for _ in range(N)]
    Xn = ntt.fwd(f.to_mont(X)); Yn = ntt.fwd(f.to_mont(Y))
    Zn = f.vec_mul(Xn, Yn)
    Z = ntt.inv(Zn)
    g = lambda a: f.from_mont(a)
    return [g(Yn)], [X, g(Z), g(Xn), g(Zn)]


def pk_witness(q, N, seed):
    f = oracle.Field(q)
    rng = random.Random(seed)
    ntt = oracle.NTT(f, N.bit_length() - 1, True)
    sk = [rng.randrange(3) - 1 for _ in range(N)]; noise = [rng.randrange(3) - 1 for _ in range(N)]
    skn = ntt.fwd(f.to_mont(sk)); nn = ntt.fwd(f.to_mont(noise))
    pk0 = f.to_mont([rng.randrange(q) for _ in range(N)])
    pk1 = f.vec_add(f.vec_mul(pk0, skn), nn)
    g = lambda a: f.from_mont(a)
    return [g(pk0), g(pk1)], [[x % q for x in sk], g(skn), [x % q for x in noise], g(nn)]


BFV_T = (1 << 16) + 1                    # PlaintextModulus of examples/bfv/main.go:91-95


def ct_params(q, N):
    """Delta and ErrBound of examples/bfv/main.go:160-174 for a synthetic ciphertext that lives in zp already
    (the second option of main.go:70-86): Delta = floor(q / t), ErrBound = N + t."""
    return q // BFV_T, N + BFV_T


def ct_witness(q, N, seed):
    """Synthetic BFV ciphertext over zp: c1 uniform, sk ternary, m centred mod t (ringT.PolyToBigintCentered,
    main.go:143: |m| <= t/2), |e| <= N + t, c0 = -c1*s + Delta*m + e (examples/bfv/main.go:22-28); all relations
    hold in the NTT domain.  (Quirk kept by the restatement: for a power-of-two bound such as t - 1 = 2^16,
    decomposeBase (buckler/utils.go:7-32) drops one digit and the digits sum to bound - 1, so a coefficient equal
    to +-(t - 1) itself has no decomposition and the proof is rejected; centred plaintexts never reach it.)"""
    f = oracle.Field(q)
    rng = random.Random(seed)
    ntt = oracle.NTT(f, N.bit_length() - 1, True)
    delta, err_bound = ct_params(q, N)
    sk = [rng.randrange(3) - 1 for _ in range(N)]
    m = [rng.randrange(-(BFV_T // 2), BFV_T // 2 + 1) for _ in range(N)]
    e = [rng.randrange(-err_bound, err_bound + 1) for _ in range(N)]
    skn = ntt.fwd(f.to_mont(sk)); mn = ntt.fwd(f.to_mont(m)); en = ntt.fwd(f.to_mont(e))
    c1 = f.to_mont([rng.randrange(q) for _ in range(N)])
    dm = f.vec_scalar_mul(mn, f.to_mont([delta])[0])
    c0 = f.vec_add(f.vec_sub(dm, f.vec_mul(c1, skn)), en)
    g = lambda a: f.from_mont(a)
    return [g(c0), g(c1)], [g(skn), g(mn), [x % q for x in m], g(en), [x % q for x in e]]


AUT_IDX = 5                              # the Galois element of the automorphism constraints (odd, generates rotations)


def norm_aut_witness(q, N, seed):
    """S ternary; SAut = S(X^idx) on coefficients (cyclotomic.go:49-64), SNTT = NTT(S), SNTTAut = the same automorphism on
    NTT slots (:66-82).  Returns the declared witnesses; the two-norm decomposition witness is filled by Prove."""
    f = oracle.Field(q)
    rng = random.Random(seed)
    ntt = oracle.NTT(f, N.bit_length() - 1, True)
    S = [rng.randrange(3) - 1 for _ in range(N)]
    Sm = f.to_mont(S)
    g = lambda a: f.from_mont(a)
    Sn = ntt.fwd(Sm)
    return [], [[x % q for x in S], g(f.aut(Sm, AUT_IDX, False)), g(Sn), g(f.aut(Sn, AUT_IDX, True))]


def circuit_and_witness(circuit, q, N, seed):
    if circuit == "norm":
        return B.norm_aut_circuit(N, AUT_IDX, N), norm_aut_witness(q, N, seed)
    if circuit == "ct":
        delta, err_bound = ct_params(q, N)
        return B.ciphertext_circuit(N, delta, BFV_T, err_bound), ct_witness(q, N, seed)
    if circuit == "pk":
        return B.public_key_circuit(N), pk_witness(q, N, seed)
    return B.mult_circuit(N), mult_witness(q, N, seed)


def test_aut_checker_is_the_adjoint():
    """autChecker.TransposeTo is the transpose of TransformTo (what linCheck relies on, linear.go:63-73):
    <u, Aut(v)> == <Aut^T(u), v> in both domains, and the NTT-domain automorphism commutes with the NTT."""
    q = pyref.modulus("zp220"); N = 64
    f = oracle.Field(q)
    rng = random.Random(1)
    u = [rng.randrange(q) for _ in range(N)]; v = [rng.randrange(q) for _ in range(N)]
    ntt = oracle.NTT(f, 6, True)
    for is_ntt in (False, True):
        chk = B.AutChecker(f, N, AUT_IDX, is_ntt)
        assert chk.idx * chk.idx_inv % (2 * N) == 1
        a = f.from_mont(chk.transform(f.to_mont(v))); b = f.from_mont(chk.transpose(f.to_mont(u)))
        assert sum(x * y for x, y in zip(u, a)) % q == sum(x * y for x, y in zip(b, v)) % q
    vm = f.to_mont(v)
    assert np.array_equal(ntt.fwd(f.aut(vm, AUT_IDX, False)), f.aut(ntt.fwd(vm), AUT_IDX, True))


def test_norm_aut_circuit_metadata():
    N = 256
    q = pyref.modulus("zp220")
    define, n_pw, n_w = B.norm_aut_circuit(N, AUT_IDX, N)
    ctx = B.Context(q, N, n_pw, n_w)
    define(ctx, oracle.Field(q))
    # 4 declared + 1 two-norm decomposition witness; 2 public (base, mask); sum-check adds 4 commitments
    assert (ctx.w_cnt, ctx.pw_cnt) == (5, 2) and len(ctx.lin_checkers) == 3 and len(ctx.sumcheck_constraints) == 2
    assert ctx.batch() == 5 + 1 + 4 + 4
    assert ctx.sumcheck_max_rank == 3 * N + 1 and ctx.commit_rank() == 3 * N + 1 + N + 1


@pytest.mark.parametrize("circuit,name,N", [("pk", "zp220", 256), ("mult", "zp128", 256), ("ct", "zp240", 256), ("norm", "zp220", 256)])
def test_prove_verify(circuit, name, N):
    q = pyref.modulus(name)
    (define, n_pw, n_w), (pw, w) = circuit_and_witness(circuit, q, N, 3)
    tape = J.Tape(b"buckler-" + circuit.encode())
    C = B.compile_circuit(q, N, define, n_pw, n_w, b"Buckler!", tape)
    pf = B.Prover(C, tape).prove(pw, w)
    assert len(pf["Witness"]) == C.ctx.batch() == len(pf["Evals"])
    assert B.Verifier(C).verify(pw, pf)
    # the committed digest of exactly this proof (tests/golden/make_buckler_digests.py): the restatement, its samplers and its
    # transcripts have not drifted since the fixture was written
    import json, os, sys
    gdir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    sys.path.insert(0, gdir)
    from make_buckler_digests import proof_digest
    want = json.load(open(os.path.join(gdir, "buckler_proof_digests.json")))[circuit]
    assert (want["field"], want["rank"], want["commitments"]) == (name, N, len(pf["Witness"]))
    assert proof_digest(pf, (q.bit_length() + 63) // 64) == want["sha256"]
    if circuit == "ct":                              # (43 commits per proof: the negative cases run on the two smaller circuits)
        return
    if circuit == "norm":
        assert pf["SumCheckMaskSum"] is not None
        # a non-ternary coefficient breaks the explicit sum-check (and the ternary constraint); a wrong automorphism image the linear check
        for wi, pos in ((0, 7), (1, 3), (3, 9)):
            w_bad = [list(v) for v in w]
            w_bad[wi][pos] = (w_bad[wi][pos] + 2) % q
            assert not B.Verifier(C).verify(pw, B.Prover(C, tape).prove(pw, w_bad)), wi
        # a tampered sum-check mask sum is rejected
        pf2 = dict(pf); pf2["SumCheckMaskSum"] = (pf["SumCheckMaskSum"] + 1) % q
        assert not B.Verifier(C).verify(pw, pf2)
        return
    # a witness violating the arithmetic constraint is rejected
    w_bad = [list(v) for v in w]
    w_bad[3][5] = (w_bad[3][5] + 1) % q              # ZNTT / NoiseNTT / ErrorNTT: breaks the arithmetic constraint
    pf_bad = B.Prover(C, tape).prove(pw, w_bad)
    assert not B.Verifier(C).verify(pw, pf_bad)
    # a proof does not verify against other public inputs
    pw2 = [list(v) for v in pw]; pw2[0][0] = (pw2[0][0] + 1) % q
    assert not B.Verifier(C).verify(pw2, pf)


def test_context_metadata_matches_survey():
    """MultCircuit -> 7 witnesses, 1 public, arithCheckMaxRank 3N+1, embRank 4N, batch 12,
    commitRank 2N+1; decomposeBase(5) = [3, 1, 1]."""
    N = 1 << 13
    q = pyref.modulus("zp128")
    define, n_pw, n_w = B.mult_circuit(N)
    ctx = B.Context(q, N, n_pw, n_w)
    define(ctx, oracle.Field(q))
    assert B.decompose_base(5) == [3, 1, 1]
    assert (ctx.w_cnt, ctx.pw_cnt, ctx.arith_max_rank, ctx.batch(), ctx.commit_rank()) == (7, 1, 3 * N + 1, 12, 2 * N + 1)
    assert 1 << (ctx.arith_max_rank - 1).bit_length() == 4 * N
    assert len(ctx.arith_constraints) == 5 and sum(len(c.witness) for c in ctx.arith_constraints) == 12
    for x in range(-5, 6):
        d = B.decompose_big(x % q, [3, 1, 1], q)
        assert sum(a * b for a, b in zip(d, [3, 1, 1])) == x and all(v in (-1, 0, 1) for v in d)


def test_ciphertext_circuit_metadata_matches_survey():
    """examples/bfv CiphertextCircuit -> 2 public, 38 witnesses (5 declared + 16 ternary
    digits for the bound t - 1 = 2^16 + 17 for N + t), 36 arithmetic constraints, embRank 4N, batch 43."""
    N = 1 << 13
    q = pyref.modulus("zp240")
    delta, err_bound = ct_params(q, N)
    define, n_pw, n_w = B.ciphertext_circuit(N, delta, BFV_T, err_bound)
    ctx = B.Context(q, N, n_pw, n_w)
    define(ctx, oracle.Field(q))
    assert (ctx.w_cnt, ctx.pw_cnt, ctx.batch()) == (38, 2, 43)
    assert len(ctx.arith_constraints) == 36 and sum(len(c.witness) for c in ctx.arith_constraints) == 105
    assert 1 << (ctx.arith_max_rank - 1).bit_length() == 4 * N
