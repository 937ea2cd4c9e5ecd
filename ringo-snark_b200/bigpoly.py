"""Host-side mirror of the reference's math/bigpoly operator interface over the C ABI.

Names and semantics follow the reference so parity tests read like its own code:
  bigpoly.Poly{Coeffs, IsNTT}                 math/bigpoly/poly.go:11-14
  CyclicTransformer / CyclotomicTransformer   math/bigpoly/ntt.go:17-24, :144-151  (FwdNTTTo / InvNTTTo)
  CyclicEvaluator / CyclotomicEvaluator       math/bigpoly/cyclic.go, cyclotomic.go + base_op.go method set
Misuse that panics in the reference raises RingoError / ValueError here.

Polynomials live on the device: a Poly is (Batch, index).  Host data crosses the boundary
only through Batch.upload / Batch.download as uint64 arrays of shape (n_polys, rank, L) in
Montgomery form -- byte-identical to the reference's []*Uint.
"""
import ctypes

import numpy as np

from . import capi
from .capi import check, lib, c_vp, c_u64p


# The moduli the reference ships (jindo-modulus output, q = b^k + 1): package -> (b, k).  SURVEY.md Appendix A2;
# q itself is what the Go shim reads from the element type (SetInt64(-1).BigInt() + 1, ntt.go:31-38).
MODULI = {
    "zp110": (12640, 8),      # buckler/internal/zp110
    "zp128": (60256, 8),      # examples/mult/zp
    "zp220": (13216, 16),     # buckler/internal/zp220
    "zp240": (31432, 16),     # examples/bfv/zp
    "zp255": (60272, 16),     # jindo/internal/zp
    "zp440": (13512, 32),     # buckler/internal/zp440
    "zp880": (13694, 64),     # buckler/internal/zp880
}


def modulus(name):
    b, k = MODULI[name]
    return b ** k + 1


def _u64p(a):
    return a.ctypes.data_as(c_u64p)


class Field:
    """bignum.Uint[E] constants for one modulus (math/bignum/bignum.go:8-31)."""

    def __init__(self, q):
        self.q = int(q)
        self.L = (self.q.bit_length() + 63) // 64
        arr = np.array([(self.q >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(self.L)], dtype=np.uint64)
        h = c_vp()
        check(lib().rs_field_create(_u64p(arr), self.L, ctypes.byref(h)))
        self.h = h

    def __del__(self):
        try:
            if getattr(self, "h", None):
                lib().rs_field_destroy(self.h)
        except Exception:
            pass

    def to_mont(self, canonical):
        canonical = np.ascontiguousarray(canonical, dtype=np.uint64)
        out = np.empty_like(canonical)
        flat_in = canonical.reshape(-1, self.L); flat_out = out.reshape(-1, self.L)
        for i in range(flat_in.shape[0]):
            check(lib().rs_field_to_mont(self.h, _u64p(flat_out[i]), _u64p(flat_in[i])))
        return out

    def from_mont(self, mont):
        mont = np.ascontiguousarray(mont, dtype=np.uint64)
        out = np.empty_like(mont)
        flat_in = mont.reshape(-1, self.L); flat_out = out.reshape(-1, self.L)
        for i in range(flat_in.shape[0]):
            check(lib().rs_field_from_mont(self.h, _u64p(flat_out[i]), _u64p(flat_in[i])))
        return out


class Batch:
    """n_polys x rank device-resident field elements."""

    def __init__(self, field, n_polys, rank, device_ptr=None, keepalive=None):
        self.field = field
        self.n_polys = int(n_polys)
        self.rank = int(rank)
        self._keep = keepalive
        h = c_vp()
        if device_ptr is None:
            check(lib().rs_batch_alloc(field.h, self.n_polys, self.rank, ctypes.byref(h)))
        else:
            check(lib().rs_batch_wrap(field.h, c_vp(device_ptr), self.n_polys, self.rank, ctypes.byref(h)))
        self.h = h

    @classmethod
    def from_torch(cls, field, tensor, n_polys, rank):
        """Wrap a CUDA torch tensor (int64/uint64 storage of n_polys*rank*L words) without copying."""
        assert tensor.is_cuda and tensor.is_contiguous()
        assert tensor.numel() * tensor.element_size() == n_polys * rank * field.L * 8
        return cls(field, n_polys, rank, device_ptr=tensor.data_ptr(), keepalive=tensor)

    def __del__(self):
        try:
            if getattr(self, "h", None):
                lib().rs_batch_free(self.h)
        except Exception:
            pass

    def upload(self, host, poly0=0):
        host = np.ascontiguousarray(host, dtype=np.uint64)
        n = host.size // (self.rank * self.field.L)
        assert n * self.rank * self.field.L == host.size
        check(lib().rs_batch_upload(self.h, poly0, n, host.ctypes.data_as(c_vp)))
        check(lib().rs_sync())   # the numpy buffer may be pageable / temporary
        return self

    def download(self, poly0=0, n=None):
        n = self.n_polys - poly0 if n is None else n
        out = np.empty((n, self.rank, self.field.L), dtype=np.uint64)
        check(lib().rs_batch_download(self.h, poly0, n, out.ctypes.data_as(c_vp)))
        return out

    def poly(self, i, is_ntt=False):
        return Poly(self, i, is_ntt)


class Poly:
    """bigpoly.Poly: Coeffs live in batch[index]; IsNTT is host-side state (poly.go:11-14)."""

    def __init__(self, batch, index, is_ntt=False):
        self.batch = batch
        self.index = index
        self.IsNTT = is_ntt

    def Rank(self):
        return self.batch.rank

    @property
    def Coeffs(self):
        return self.batch.download(self.index, 1)[0]

    def set(self, coeffs):
        self.batch.upload(np.ascontiguousarray(coeffs, dtype=np.uint64).reshape(1, self.batch.rank, -1), self.index)
        return self

    def CopyFrom(self, p0):
        check(lib().rs_batch_copy(self.batch.h, self.index, 0, p0.batch.h, p0.index, 0, self.batch.rank))
        self.IsNTT = p0.IsNTT

    def Clear(self):
        check(lib().rs_batch_zero(self.batch.h, self.index, 0, self.batch.rank))

    def Evaluate(self, x, n=None):
        """poly.go:64-76 (x: one Montgomery element)."""
        if self.IsNTT:
            raise ValueError("Evaluate: p is in NTT form")
        x = np.ascontiguousarray(x, dtype=np.uint64)
        out = np.zeros(self.batch.field.L, dtype=np.uint64)
        check(lib().rs_poly_evaluate(self.batch.h, self.index, self.batch.rank if n is None else n, _u64p(x), _u64p(out)))
        return out


class _Transformer:
    negacyclic = False

    def __init__(self, field, rank):
        if rank <= 0 or rank & (rank - 1):
            raise ValueError("rank must be a power of two")
        self.field = field
        self.rank = rank
        self.logN = rank.bit_length() - 1
        h = c_vp()
        check(lib().rs_ntt_create(field.h, self.logN, int(self.negacyclic), ctypes.byref(h)))
        self.h = h

    def __del__(self):
        try:
            if getattr(self, "h", None):
                lib().rs_ntt_destroy(self.h)
        except Exception:
            pass

    def Rank(self):
        return self.rank

    def tables(self):
        L = self.field.L
        tw = np.zeros((self.rank, L), dtype=np.uint64); twinv = np.zeros_like(tw); rinv = np.zeros(L, dtype=np.uint64)
        check(lib().rs_ntt_tables(self.h, _u64p(tw), _u64p(twinv), _u64p(rinv)))
        return tw, twinv, rinv

    # batch-level entry points (n_polys polynomials per call)
    def FwdNTTTo(self, dst, src, dst_poly=0, src_poly=0, n_polys=None):
        n = src.n_polys - src_poly if n_polys is None else n_polys
        check(lib().rs_ntt_forward(self.h, dst.h, dst_poly, src.h, src_poly, n))

    def InvNTTTo(self, dst, src, dst_poly=0, src_poly=0, n_polys=None):
        n = src.n_polys - src_poly if n_polys is None else n_polys
        check(lib().rs_ntt_inverse(self.h, dst.h, dst_poly, src.h, src_poly, n))


class CyclicTransformer(_Transformer):
    """math/bigpoly/ntt.go:17-95"""
    negacyclic = False


class CyclotomicTransformer(_Transformer):
    """math/bigpoly/ntt.go:144-203"""
    negacyclic = True


class _BaseOperator:
    """math/bigpoly/base_op.go:42-207 over device polys."""

    def __init__(self, field, rank, ntt):
        self.field = field
        self.rank = rank
        self.ntt = ntt

    def Rank(self):
        return self.rank

    def NewBatch(self, n_polys):
        return Batch(self.field, n_polys, self.rank)

    def _unary(self, pOut, p):
        if pOut.Rank() != self.rank or p.Rank() != self.rank:
            raise ValueError("inconsistent input(s)")

    def _binary(self, pOut, p0, p1):
        if pOut.Rank() != self.rank or p0.Rank() != self.rank or p1.Rank() != self.rank:
            raise ValueError("inconsistent input(s)")
        if p0.IsNTT != p1.IsNTT:
            raise ValueError("inconsistent input(s)")

    def AddTo(self, pOut, p0, p1):
        self._binary(pOut, p0, p1)
        check(lib().rs_vec_add(pOut.batch.h, pOut.index, p0.batch.h, p0.index, p1.batch.h, p1.index, 1))
        pOut.IsNTT = p0.IsNTT

    def SubTo(self, pOut, p0, p1):
        self._binary(pOut, p0, p1)
        check(lib().rs_vec_sub(pOut.batch.h, pOut.index, p0.batch.h, p0.index, p1.batch.h, p1.index, 1))
        pOut.IsNTT = p0.IsNTT

    def NegTo(self, pOut, p):
        self._unary(pOut, p)
        check(lib().rs_vec_neg(pOut.batch.h, pOut.index, p.batch.h, p.index, 1))
        pOut.IsNTT = p.IsNTT

    def ScalarMulTo(self, pOut, p, c):
        self._unary(pOut, p)
        check(lib().rs_vec_scalar_mul(pOut.batch.h, pOut.index, p.batch.h, p.index, _u64p(np.ascontiguousarray(c, dtype=np.uint64)), 1))
        pOut.IsNTT = p.IsNTT

    def ScalarMulAddTo(self, pOut, p, c):
        self._unary(pOut, p)
        check(lib().rs_vec_scalar_mul_add(pOut.batch.h, pOut.index, p.batch.h, p.index, _u64p(np.ascontiguousarray(c, dtype=np.uint64)), 1))
        pOut.IsNTT = p.IsNTT

    def ScalarMulSubTo(self, pOut, p, c):
        self._unary(pOut, p)
        check(lib().rs_vec_scalar_mul_sub(pOut.batch.h, pOut.index, p.batch.h, p.index, _u64p(np.ascontiguousarray(c, dtype=np.uint64)), 1))
        pOut.IsNTT = p.IsNTT

    def _need_ntt(self, p0, p1):
        if not p0.IsNTT or not p1.IsNTT:
            raise ValueError("input(s) not in NTT domain")

    def MulTo(self, pOut, p0, p1):
        self._binary(pOut, p0, p1); self._need_ntt(p0, p1)
        check(lib().rs_vec_mul(pOut.batch.h, pOut.index, p0.batch.h, p0.index, p1.batch.h, p1.index, 1))
        pOut.IsNTT = True

    def MulAddTo(self, pOut, p0, p1):
        self._binary(pOut, p0, p1); self._need_ntt(p0, p1)
        check(lib().rs_vec_mul_add(pOut.batch.h, pOut.index, p0.batch.h, p0.index, p1.batch.h, p1.index, 1))
        pOut.IsNTT = True

    def MulSubTo(self, pOut, p0, p1):
        self._binary(pOut, p0, p1); self._need_ntt(p0, p1)
        check(lib().rs_vec_mul_sub(pOut.batch.h, pOut.index, p0.batch.h, p0.index, p1.batch.h, p1.index, 1))
        pOut.IsNTT = True

    def NTTTo(self, pOut, p):
        self._unary(pOut, p)
        if p.IsNTT:
            raise ValueError("input already in NTT domain")
        check(lib().rs_ntt_forward(self.ntt.h, pOut.batch.h, pOut.index, p.batch.h, p.index, 1))
        pOut.IsNTT = True

    def InvNTTTo(self, pOut, p):
        self._unary(pOut, p)
        if not p.IsNTT:
            raise ValueError("input not in NTT domain")
        check(lib().rs_ntt_inverse(self.ntt.h, pOut.batch.h, pOut.index, p.batch.h, p.index, 1))
        pOut.IsNTT = False


class CyclicEvaluator(_BaseOperator):
    """math/bigpoly/cyclic.go:8-16"""

    def __init__(self, field, rank):
        super().__init__(field, rank, CyclicTransformer(field, rank))

    def QuoRemByVanishing(self, p, N, quo, rem):
        """cyclic.go:18-37 (quo, rem are caller-provided device polys)."""
        if p.Rank() != self.rank:
            raise ValueError("inputs not consistent")
        if p.IsNTT:
            raise ValueError("input in NTT domain")
        check(lib().rs_quorem_vanishing(quo.batch.h, quo.index, rem.batch.h, rem.index, p.batch.h, p.index, N))
        quo.IsNTT = False; rem.IsNTT = False
        return quo, rem


class CyclotomicEvaluator(_BaseOperator):
    """math/bigpoly/cyclotomic.go:11-20"""

    def __init__(self, field, rank):
        super().__init__(field, rank, CyclotomicTransformer(field, rank))

    def AutTo(self, pOut, p, idx):
        """cyclotomic.go:29-47"""
        self._unary(pOut, p)
        if idx % 2 == 0:
            raise ValueError("AutTo: idx must be odd")
        check(lib().rs_aut(pOut.batch.h, pOut.index, p.batch.h, p.index, 1, idx, int(p.IsNTT)))
        pOut.IsNTT = p.IsNTT
