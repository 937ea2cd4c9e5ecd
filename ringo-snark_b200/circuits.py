"""What buckler.Compile leaves in Prover.ctx for the reference's example circuits, as the plain data rs_buckler_create takes
(include/ringo_buckler.h), plus a witness generator for benchmarks.  Product side: no oracle involved.

examples/mult MultCircuit (examples/mult/main.go:35-65):
    public   YNTT(0)
    witness  XCoeffs(0) ZCoeffs(1) XNTT(2) ZNTT(3) + the three ternary digits of XCoeffs (4, 5, 6) that AddInfNormConstraint(X, 5)
             adds (context.go:120-163: decomposeBase(5) = [3, 1, 1])
    linear   XNTT = NTT(XCoeffs), ZNTT = NTT(ZCoeffs)                               (nttChecker, linear.go:20-43)
    arith    YNTT * XNTT - ZNTT = 0 ; d^3 - d = 0 for each digit ; X - 3 d0 - d1 - d2 = 0
"""
import numpy as np

from . import capi
from .capi import check, lib


def mult_circuit(q, N):
    """-> (description dict for buckler.Prover / NativeProver, commit rank, batch) of MultCircuit at rank N"""
    terms = [(0, 0, 0, [2]), (0, 1, -1, [3]),
             (1, 2, -1, [4, 4, 4]), (1, 3, -1, [4]), (2, 4, -1, [5, 5, 5]), (2, 5, -1, [5]), (3, 6, -1, [6, 6, 6]), (3, 7, -1, [6]),
             (4, 8, -1, [0]), (4, 9, -1, [4]), (4, 10, -1, [5]), (4, 11, -1, [6])]
    coeffs = [c % q for c in (1, -1, 1, -1, 1, -1, 1, -1, 1, -3, -1, -1)]
    arith_max_rank = 3 * N + 1                         # degree-3 constraints over rank-(N+1) encodings (context.go:96-118)
    desc = {"rank": N, "emb_rank": 1 << (arith_max_rank - 1).bit_length(), "w_cnt": 7, "pw_cnt": 1, "arith_max_rank": arith_max_rank,
            "arith_terms": terms, "arith_coeffs": coeffs, "lin": [("ntt", [(2, 0), (3, 1)])]}
    commit_rank = 2 * N + 1                            # the largest committed vector: the arithmetic quotient (compile.go:176-184)
    batch = 7 + 1 + 4                                  # witnesses + arithmetic quotient + linCheck mask, quo, remLo, remHi
    return desc, commit_rank, batch


_DIG = np.array([(0, 0, 0), (0, 1, 0), (0, 1, 1), (1, 0, 0), (1, 1, 0), (1, 1, 1)], dtype=np.uint64)   # x = 3 d0 + d1 + d2


def mult_witness(field, q, N, seed, transformer=None):
    """A random MultCircuit instance through the library's own NTT: X uniform in [0, 5], Y uniform 63-bit, Z = X * Y in
    Z_q[X]/(X^N + 1).  Returns (pw [1, N, L], w [7, N, L]) as Montgomery limbs -- what the reference's user puts into the circuit
    struct plus the digit witnesses Prove derives (prover.go:77-86)."""
    from .bigpoly import Batch, CyclotomicTransformer
    L = field.L
    neg = transformer or CyclotomicTransformer(field, N)
    rng = np.random.default_rng(seed)
    X = rng.integers(0, 6, size=N, dtype=np.uint64)
    Y = rng.integers(0, 1 << 63, size=N, dtype=np.uint64)
    raw = np.zeros((8, N, L), dtype=np.uint64)
    raw[0, :, 0] = X; raw[1, :, 0] = Y
    for k in range(3):
        raw[2 + k, :, 0] = _DIG[X, k]
    bt = Batch(field, 8, N).upload(raw)
    R2 = pow(1 << (64 * L), 2, q)                       # canonical small integers -> Montgomery form: one product by R^2
    r2 = np.array([(R2 >> (64 * i)) & ((1 << 64) - 1) for i in range(L)], dtype=np.uint64)
    check(lib().rs_vec_scalar_mul(bt.h, 0, bt.h, 0, r2.ctypes.data_as(capi.c_u64p), 5))
    neg.FwdNTTTo(bt, bt, dst_poly=5, src_poly=0, n_polys=2)       # 5: XNTT, 6: YNTT
    check(lib().rs_vec_mul(bt.h, 7, bt.h, 5, bt.h, 6, 1))         # 7: ZNTT
    out = bt.download()
    zb = Batch(field, 1, N).upload(out[7:8])
    neg.InvNTTTo(zb, zb)
    w = np.stack([out[0], zb.download()[0], out[5], out[7], out[2], out[3], out[4]])
    return out[6:7].copy(), w
