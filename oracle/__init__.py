"""oracle -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement of the reference (sp301415/ringo-snark) hot path.  Only
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
leg import this package; the CUDA product (ringo-snark_b200/) never does.

Layout:
  field_oracle.[ch]  plain-C big-field + math/bigpoly restatement (this file binds it)
  pyref.py           Python-integer cross-checks (the math/big equivalent)
  ring_oracle.[ch]   plain-C lattigo-compatible small-ring arithmetic + Jindo commit pipeline
  jindo_ref.py       Python restatement of jindo params / csprng / encoder / prover / verifier

Parity pinning status (see DESIGN.md "Oracle"):
  * field arithmetic: pinned -- constants checked against the reference's
    vendored zp packages (tests/golden/field_constants.json), ops against
    Python ints (the reference's own gopter tests pin against math/big).
  * bigpoly NTT: the reference has no test; pinned by the algebraic contract
    out[j] = a(g^brv(j)) / a(g^(2 brv(j)+1)) with the reference's generator
    search, checked here against a naive O(N^2) evaluation.
  * lattigo-dependent bytes (RNS primes, psi, serialization): PARITY UNPINNED
    (third-party module not in /root/reference, Go toolchain absent).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")
ORC_MAX_L = 16


def build(force=False):
    """Compile the C oracle with the committed Makefile (gcc only)."""
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".c", ".h"))]
    if (not force and os.path.exists(_LIB_PATH)
            and all(os.path.getmtime(_LIB_PATH) >= os.path.getmtime(s) for s in srcs)):
        return _LIB_PATH
    subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


class _Field(ctypes.Structure):
    _fields_ = [("L", ctypes.c_int),
                ("q", ctypes.c_uint64 * ORC_MAX_L),
                ("qinvneg", ctypes.c_uint64),
                ("rsquare", ctypes.c_uint64 * ORC_MAX_L),
                ("one", ctypes.c_uint64 * ORC_MAX_L)]


class _Ntt(ctypes.Structure):
    _fields_ = [("f", ctypes.POINTER(_Field)),
                ("logN", ctypes.c_int),
                ("negacyclic", ctypes.c_int),
                ("N", ctypes.c_size_t),
                ("tw", ctypes.POINTER(ctypes.c_uint64)),
                ("twinv", ctypes.POINTER(ctypes.c_uint64)),
                ("rankinv", ctypes.c_uint64 * ORC_MAX_L),
                ("gen_x", ctypes.c_uint64)]


class Term(ctypes.Structure):
    _fields_ = [("constraint", ctypes.c_int32), ("coeff", ctypes.c_int32),
                ("pw", ctypes.c_int32), ("n_w", ctypes.c_int32),
                ("w", ctypes.c_int32 * 8)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64))


def int_to_limbs(x, L):
    return np.array([(x >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(L)], dtype=np.uint64)


def limbs_to_int(a):
    return sum(int(v) << (64 * i) for i, v in enumerate(np.asarray(a).reshape(-1)))


def ints_to_array(xs, L):
    out = np.zeros((len(xs), L), dtype=np.uint64)
    for i, x in enumerate(xs):
        for k in range(L):
            out[i, k] = (x >> (64 * k)) & 0xFFFFFFFFFFFFFFFF
    return out


def array_to_ints(a):
    a = np.asarray(a)
    L = a.shape[-1]
    flat = a.reshape(-1, L)
    return [sum(int(flat[i, k]) << (64 * k) for k in range(L)) for i in range(flat.shape[0])]


class Field:
    """Montgomery prime field with L 64-bit limbs (gnark-crypto generated zp semantics)."""

    def __init__(self, q):
        self.q = int(q)
        self.L = (self.q.bit_length() + 63) // 64
        self._f = _Field()
        qa = int_to_limbs(self.q, self.L)
        rc = lib().orc_field_init(ctypes.byref(self._f), _p(qa), self.L)
        if rc != 0:
            raise ValueError("orc_field_init failed")
        self.R = 1 << (64 * self.L)

    # -- constants --------------------------------------------------------
    @property
    def qinvneg(self):
        return int(self._f.qinvneg)

    @property
    def rsquare(self):
        return sum(int(self._f.rsquare[i]) << (64 * i) for i in range(self.L))

    # -- conversion helpers (Python ints <-> Montgomery limb arrays) -------
    def to_mont(self, xs):
        return ints_to_array([(x % self.q) * self.R % self.q for x in xs], self.L)

    def from_mont(self, a):
        rinv = pow(self.R, -1, self.q)
        return [v * rinv % self.q for v in array_to_ints(a)]

    def _bin(self, name, a, b):
        a = np.ascontiguousarray(a, dtype=np.uint64); b = np.ascontiguousarray(b, dtype=np.uint64)
        o = np.empty_like(a)
        getattr(lib(), name)(ctypes.byref(self._f), _p(o), _p(a), _p(b), ctypes.c_size_t(a.size // self.L))
        return o

    def vec_add(self, a, b): return self._bin("orc_vec_add", a, b)
    def vec_sub(self, a, b): return self._bin("orc_vec_sub", a, b)
    def vec_mul(self, a, b): return self._bin("orc_vec_mul", a, b)

    def vec_neg(self, a):
        a = np.ascontiguousarray(a, dtype=np.uint64)
        o = np.empty_like(a)
        lib().orc_vec_neg(ctypes.byref(self._f), _p(o), _p(a), ctypes.c_size_t(a.size // self.L))
        return o

    def vec_scalar_mul(self, a, c):
        a = np.ascontiguousarray(a, dtype=np.uint64); c = np.ascontiguousarray(c, dtype=np.uint64)
        o = np.empty_like(a)
        lib().orc_vec_scalar_mul(ctypes.byref(self._f), _p(o), _p(a), _p(c), ctypes.c_size_t(a.size // self.L))
        return o

    def quorem_vanishing(self, p, N):
        p = np.ascontiguousarray(p, dtype=np.uint64)
        quo = np.empty_like(p); rem = np.empty_like(p)
        lib().orc_quorem_vanishing(ctypes.byref(self._f), _p(quo), _p(rem), _p(p),
                                   ctypes.c_size_t(p.shape[0]), ctypes.c_size_t(N))
        return quo, rem

    def aut(self, p, idx, is_ntt):
        p = np.ascontiguousarray(p, dtype=np.uint64)
        o = np.empty_like(p)
        lib().orc_aut(ctypes.byref(self._f), _p(o), _p(p), ctypes.c_size_t(p.shape[0]),
                      ctypes.c_long(idx), ctypes.c_int(1 if is_ntt else 0))
        return o

    def poly_evaluate(self, coeffs, x):
        coeffs = np.ascontiguousarray(coeffs, dtype=np.uint64); x = np.ascontiguousarray(x, dtype=np.uint64)
        z = np.zeros(self.L, dtype=np.uint64)
        lib().orc_poly_evaluate(ctypes.byref(self._f), _p(z), _p(coeffs), ctypes.c_size_t(coeffs.shape[0]), _p(x))
        return z

    def eval_circuit(self, pw_ntt, w_ntt, terms, coeffs, batch_const, rank):
        """pw_ntt / w_ntt: lists of (rank, L) arrays; terms: list of (constraint, coeff, pw, [w...])."""
        tarr = (Term * len(terms))()
        for i, (c, co, pw, ws) in enumerate(terms):
            tarr[i].constraint = c; tarr[i].coeff = co; tarr[i].pw = pw; tarr[i].n_w = len(ws)
            for k, w in enumerate(ws):
                tarr[i].w[k] = w
        P = ctypes.POINTER(ctypes.c_uint64)
        pw_keep = [np.ascontiguousarray(a, dtype=np.uint64) for a in pw_ntt]
        w_keep = [np.ascontiguousarray(a, dtype=np.uint64) for a in w_ntt]
        pwp = (P * max(1, len(pw_keep)))(*[_p(a) for a in pw_keep])
        wp = (P * max(1, len(w_keep)))(*[_p(a) for a in w_keep])
        coeffs = np.ascontiguousarray(coeffs, dtype=np.uint64)
        bc = np.ascontiguousarray(batch_const, dtype=np.uint64)
        out = np.zeros((rank, self.L), dtype=np.uint64)
        lib().orc_eval_circuit(ctypes.byref(self._f), _p(out), ctypes.c_size_t(rank), pwp, wp,
                               tarr, ctypes.c_size_t(len(terms)), _p(coeffs), _p(bc))
        return out


class NTT:
    """math/bigpoly CyclicTransformer (negacyclic=False) / CyclotomicTransformer (True)."""

    def __init__(self, field, logN, negacyclic):
        self.field = field
        self.logN = logN
        self.N = 1 << logN
        self.negacyclic = bool(negacyclic)
        self._t = _Ntt()
        rc = lib().orc_ntt_init(ctypes.byref(self._t), ctypes.byref(field._f), logN, int(self.negacyclic))
        if rc != 0:
            raise ValueError("NTT not supported")

    def __del__(self):
        try:
            lib().orc_ntt_free(ctypes.byref(self._t))
        except Exception:
            pass

    @property
    def gen_x(self):
        return int(self._t.gen_x)

    def tables(self):
        n = self.N * self.field.L
        tw = np.ctypeslib.as_array(self._t.tw, shape=(n,)).reshape(self.N, self.field.L).copy()
        twinv = np.ctypeslib.as_array(self._t.twinv, shape=(n,)).reshape(self.N, self.field.L).copy()
        rinv = np.array([self._t.rankinv[i] for i in range(self.field.L)], dtype=np.uint64)
        return tw, twinv, rinv

    def fwd(self, v):
        v = np.ascontiguousarray(v, dtype=np.uint64)
        o = np.empty_like(v)
        flat_in = v.reshape(-1, self.N, self.field.L); flat_out = o.reshape(-1, self.N, self.field.L)
        for i in range(flat_in.shape[0]):
            lib().orc_ntt_fwd(ctypes.byref(self._t), _p(flat_out[i]), _p(flat_in[i]))
        return o

    def inv(self, v):
        v = np.ascontiguousarray(v, dtype=np.uint64)
        o = np.empty_like(v)
        flat_in = v.reshape(-1, self.N, self.field.L); flat_out = o.reshape(-1, self.N, self.field.L)
        for i in range(flat_in.shape[0]):
            lib().orc_ntt_inv(ctypes.byref(self._t), _p(flat_out[i]), _p(flat_in[i]))
        return o

    def fwd_batch_inplace(self, v, threads=1):
        assert v.flags.c_contiguous and v.dtype == np.uint64
        n = v.size // (self.N * self.field.L)
        lib().orc_ntt_fwd_batch(ctypes.byref(self._t), _p(v), ctypes.c_size_t(n), ctypes.c_int(threads))

    def inv_batch_inplace(self, v, threads=1):
        assert v.flags.c_contiguous and v.dtype == np.uint64
        n = v.size // (self.N * self.field.L)
        lib().orc_ntt_inv_batch(ctypes.byref(self._t), _p(v), ctypes.c_size_t(n), ctypes.c_int(threads))
