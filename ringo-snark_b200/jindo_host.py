"""Host-side inputs of the Jindo C ABI (include/ringo_jindo.h) for callers that do not have the Go library at hand -- benchmarks,
examples, the Python mirror of buckler.Prover.  The derivation itself is compiled code inside the library
(rs_jindo_new_parameters, csrc/jindo_params.cpp: jindo.NewParameters, jindo/params.go:18-321, with lattigo's prime / psi
conventions restated); this module only turns its result into the objects the bindings take, plus the two wire-format helpers
the Python mirror's transcript uses.  No device work happens here."""
import ctypes

import numpy as np

from . import capi
from .capi import check, lib


class Parameters:
    """jindo.Parameters (jindo/params.go:53-124) as plain attributes"""


def new_parameters(q_field, target_n, batch):
    """jindo.NewParameters(q, targetN, batch) through rs_jindo_new_parameters"""
    if target_n < 1 or batch < 1:
        raise ValueError("NewParameters: bad arguments")
    L = (q_field.bit_length() + 63) // 64
    qa = np.array([(q_field >> (64 * i)) & ((1 << 64) - 1) for i in range(L)], dtype=np.uint64)
    D = capi.JindoDerived()
    rc = lib().rs_jindo_new_parameters(qa.ctypes.data_as(capi.c_u64p), L, int(target_n), int(batch), ctypes.byref(D))
    if rc != 0:
        raise ValueError(lib().rs_last_error().decode())
    p = Parameters()
    P = D.p
    p.q_field, p.batch, p.rank = q_field, batch, int(D.rank)
    for k in ("d", "slots", "exp", "base", "rows", "cols", "in_msis", "mlwe", "out_msis", "log_in_cut", "log_out_cut"):
        setattr(p, k, int(getattr(P, k)))
    p.in_com_dcmp_len = int(D.in_com_dcmp_len)
    p.q_primes, p.qout_primes = [int(v) for v in P.q[:P.n_q]], [int(v) for v in P.qout[:P.n_qout]]
    p.psi_q, p.psi_qout = [int(v) for v in P.psi_q[:P.n_q]], [int(v) for v in P.psi_qout[:P.n_qout]]
    p.q_bits, p.qout_bits = int(D.q_bits), int(D.qout_bits)
    p.ecd_std, p.mlwe_std = float(D.ecd_std), float(D.mlwe_std)
    p.ecd_blind_std, p.mask_blind_std = float(D.noise.ecd_blind_std), float(D.noise.mask_blind_std)
    p.mask_std, p.mask_mlwe_std = float(D.noise.mask_std), float(D.noise.mask_mlwe_std)
    p.res_two_nm, p.in_com_dcmp_two_nm = float(D.res_two_nm), float(D.in_com_dcmp_two_nm)
    p.com_size, p.pf_size = float(D.com_size), float(D.pf_size)
    p.delta_inv = [float(v) for v in P.delta_inv[:P.exp]]
    p.cdt_len, p.cdt_tail_lo = int(P.cdt_len), int(P.cdt_tail_lo)
    p._derived = D
    return p


def prover_inputs(p):
    """dict for ringo_snark_b200.jindo.Prover + the 128 x cdt_len TwinCDT tables (sigma = ecdStdDev)"""
    d = {"d": p.d, "slots": p.slots, "exp": p.exp, "base": p.base, "rows": p.rows, "cols": p.cols,
         "in_msis": p.in_msis, "mlwe": p.mlwe, "out_msis": p.out_msis, "log_in_cut": p.log_in_cut, "log_out_cut": p.log_out_cut,
         "q": list(p.q_primes), "qout": list(p.qout_primes), "psi_q": list(p.psi_q), "psi_qout": list(p.psi_qout),
         "delta_inv": list(p.delta_inv), "cdt_tail_lo": p.cdt_tail_lo, "cdt_sigma": p.ecd_std}
    tables = np.zeros((128, p.cdt_len), dtype=np.uint64)
    check(lib().rs_jindo_cdt_tables(ctypes.byref(p._derived), tables.ctypes.data_as(capi.c_vp)))
    return d, tables


def poly_write_to(poly):
    """[lattigo] ring.Poly.WriteTo framing (restated; see DESIGN.md "unpinned")."""
    import struct
    out = [struct.pack("<Q", len(poly))]
    for c in poly:
        out.append(struct.pack("<Q", len(c)))
        out.append(struct.pack("<%dQ" % len(c), *[int(v) for v in c]))
    return b"".join(out)


def commitment_write_raw(p, value):
    """Commitment.WriteRawTo (jindo/entities.go:85-100): polys allocated with ringQ's limb count."""
    nl = len(p.q_primes)
    out = []
    if isinstance(value, np.ndarray):                      # [out_msis, n_qout, d] uint64: same bytes without int lists
        import struct
        head, lim = struct.pack("<Q", max(nl, value.shape[1])), struct.pack("<Q", p.d)
        for v in value:
            out.append(head)
            for l in range(max(nl, value.shape[1])):
                out.append(lim)
                out.append(np.ascontiguousarray(v[l], dtype="<u8").tobytes() if l < value.shape[1] else bytes(8 * p.d))
        return b"".join(out)
    for v in value:
        limbs = [list(c) for c in v] + [[0] * p.d for _ in range(nl - len(v))]
        out.append(poly_write_to(limbs))
    return b"".join(out)
