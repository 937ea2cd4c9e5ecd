"""Test helpers: turn the oracle's Jindo objects (oracle/jindo_ref.py) into the plain arrays the
C ABI takes (what the Go shim would pass from jindo.Parameters / CommitKey / csprng)."""
import numpy as np

import oracle
from oracle import jindo_ref as J


def params_dict(p, encoder):
    return {
        "d": p.d, "slots": p.slots, "exp": p.exp, "base": p.base, "rows": p.rows, "cols": p.cols,
        "in_msis": p.in_msis, "mlwe": p.mlwe, "out_msis": p.out_msis,
        "log_in_cut": p.log_in_cut, "log_out_cut": p.log_out_cut,
        "q": list(p.q_primes), "qout": list(p.qout_primes),
        "psi_q": [s.psi for s in p.ringQ.sub], "psi_qout": [s.psi for s in p.ringQOut.sub],
        "delta_inv": list(encoder.delta_inv),
        "cdt_tail_lo": encoder.twin.tail_lo, "cdt_sigma": encoder.twin.std_dev,
    }


def cdt_array(encoder):
    return np.array(encoder.twin.tables, dtype=np.uint64)


def polys_array(polys):
    """nested lists [...][limb][coeff] of ints -> uint64 array"""
    return np.array(polys, dtype=np.uint64)


def tape_arrays(fo, p, rec):
    """fo: oracle.Field of the committed field; rec: Prover.trace of one commit."""
    n_noise = (p.cols + 1) * (p.in_msis + p.mlwe)
    mlwe = np.array([rec["mlwe"][(i, j)] for i in range(p.cols + 1) for j in range(p.in_msis + p.mlwe)], dtype=np.int64)
    assert mlwe.shape == (n_noise, p.d)
    mask = [x for row in rec["mask"] for x in row]
    return {
        "first_row": fo.to_mont(rec["first"]), "last_row": fo.to_mont(rec["last"]),
        "mask_rows": fo.to_mont(mask),
        "twin_u": np.array(rec["twin_u"], dtype=np.uint64),
        "cosac_g": np.array(rec["cosac_g"], dtype=np.int64),
        "mlwe": mlwe,
    }


def setup(q_field, N, batch, crs=b"Jindo!", seed=b"tape-seed"):
    p = J.new_parameters(q_field, N, batch)
    tape = J.Tape(seed)
    prv = J.Prover(p, crs, tape)
    return p, tape, prv
