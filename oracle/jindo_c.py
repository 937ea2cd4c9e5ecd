"""ctypes front end of oracle/jindo_oracle.c (C restatement of jindo.Prover.Commit).

TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this.
Takes the same inputs as the CUDA path's C ABI (parameters, commit key, TwinCDT tables, randomness tape),
so one set of arrays drives both sides.  Pinned against oracle/jindo_ref.py by tests/test_oracle_jindo_c.py."""
import ctypes

import numpy as np

import oracle

MAX_LIMBS, MAX_EXP = 4, 64


class Params(ctypes.Structure):
    """orj_params: same layout as rs_jindo_params (include/ringo_jindo.h)."""
    _fields_ = [("d", ctypes.c_int32), ("slots", ctypes.c_int32), ("exp", ctypes.c_int32), ("base", ctypes.c_uint64),
                ("rows", ctypes.c_int32), ("cols", ctypes.c_int32),
                ("in_msis", ctypes.c_int32), ("mlwe", ctypes.c_int32), ("out_msis", ctypes.c_int32),
                ("log_in_cut", ctypes.c_int32), ("log_out_cut", ctypes.c_int32),
                ("n_q", ctypes.c_int32), ("n_qout", ctypes.c_int32),
                ("q", ctypes.c_uint64 * MAX_LIMBS), ("qout", ctypes.c_uint64 * MAX_LIMBS),
                ("psi_q", ctypes.c_uint64 * MAX_LIMBS), ("psi_qout", ctypes.c_uint64 * MAX_LIMBS),
                ("delta_inv", ctypes.c_double * MAX_EXP),
                ("cdt_len", ctypes.c_int32), ("cdt_tail_lo", ctypes.c_int64), ("cdt_sigma", ctypes.c_double)]


def make_params(pd, cdt_len):
    """pd: dict with the fields of tests/jindo_util.params_dict / jindo_host.prover_inputs."""
    P = Params()
    for k in ("d", "slots", "exp", "base", "rows", "cols", "in_msis", "mlwe", "out_msis", "log_in_cut", "log_out_cut"):
        setattr(P, k, int(pd[k]))
    P.n_q, P.n_qout = len(pd["q"]), len(pd["qout"])
    for i, v in enumerate(pd["q"]): P.q[i] = v
    for i, v in enumerate(pd["qout"]): P.qout[i] = v
    for i, v in enumerate(pd["psi_q"]): P.psi_q[i] = v
    for i, v in enumerate(pd["psi_qout"]): P.psi_qout[i] = v
    for i, v in enumerate(pd["delta_inv"]): P.delta_inv[i] = v
    P.cdt_len = cdt_len
    P.cdt_tail_lo = int(pd["cdt_tail_lo"]); P.cdt_sigma = float(pd["cdt_sigma"])
    return P


def commit(field, pd, ck_in, ck_mlwe, ck_out, cdt, v_mont, tape, threads=1):
    """Prover.Commit (prover.go:45-62).  field: oracle.Field; v_mont: [len, L] Montgomery limbs; tape: dict of
    arrays first_row, last_row, mask_rows, twin_u, cosac_g, mlwe.  Returns dict of uint64 arrays
    encode [(cols+1), rows, n_q, 256], mlwe [(cols+1), mlwe+in_msis, n_q, 256], incommit [(cols+1)*in_msis, n_qout, 256],
    commitment [out_msis, n_qout, 256]."""
    L = oracle.lib()
    u64 = lambda a: np.ascontiguousarray(a, dtype=np.uint64)
    i64 = lambda a: np.ascontiguousarray(a, dtype=np.int64)
    cdt = u64(cdt)
    P = make_params(pd, cdt.shape[1])
    ck_in, ck_mlwe, ck_out, v_mont = u64(ck_in), u64(ck_mlwe), u64(ck_out), u64(v_mont)
    t = {k: (i64(tape[k]) if k in ("cosac_g", "mlwe") else u64(tape[k])) for k in ("first_row", "last_row", "mask_rows", "twin_u", "cosac_g", "mlwe")}
    nq, nqo, d = P.n_q, P.n_qout, P.d
    out = {
        "encode": np.zeros((P.cols + 1, P.rows, nq, d), dtype=np.uint64),
        "mlwe": np.zeros((P.cols + 1, P.mlwe + P.in_msis, nq, d), dtype=np.uint64),
        "incommit": np.zeros(((P.cols + 1) * P.in_msis, nqo, d), dtype=np.uint64),
        "commitment": np.zeros((P.out_msis, nqo, d), dtype=np.uint64),
    }
    vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    fn = L.orj_commit
    fn.restype = ctypes.c_int
    fn.argtypes = [ctypes.c_void_p, ctypes.POINTER(Params)] + [ctypes.c_void_p] * 5 + [ctypes.c_size_t] + [ctypes.c_void_p] * 10 + [ctypes.c_int]
    rc = fn(ctypes.addressof(field._f), ctypes.byref(P), vp(ck_in), vp(ck_mlwe), vp(ck_out), vp(cdt), vp(v_mont), v_mont.shape[0],
            vp(t["first_row"]), vp(t["last_row"]), vp(t["mask_rows"]), vp(t["twin_u"]), vp(t["cosac_g"]), vp(t["mlwe"]),
            vp(out["encode"]), vp(out["mlwe"]), vp(out["incommit"]), vp(out["commitment"]), int(threads))
    if rc != 0:
        raise ValueError(f"orj_commit failed ({rc}): parameters outside the C restatement's limits")
    return out
