"""Pins the C restatement of Commit (oracle/jindo_oracle.c) against the Python restatement
(oracle/jindo_ref.py) on the same parameters, key and randomness tape: every polynomial of the Opening
and the Commitment, for 1-3 RNS limbs, ragged vectors and multi-threaded column splitting."""
import numpy as np
import pytest

import oracle
from oracle import jindo_c, jindo_ref as J, pyref
from jindo_util import cdt_array, params_dict, polys_array, tape_arrays

CASES = [("zp255", 1024, 1, 1024, 1), ("zp255", 1024, 1, 700, 3), ("zp255", 1024, 1, 1, 2),
         ("zp128", 700, 2, 700, 4), ("zp220", 2049, 9, 2049, 2)]


@pytest.mark.parametrize("name,N,batch,length,threads", CASES)
def test_c_commit_matches_python_restatement(name, N, batch, length, threads):
    q = pyref.modulus(name)
    fo = oracle.Field(q)
    p = J.new_parameters(q, N, batch)
    tape = J.Tape(b"c-commit-" + name.encode())
    prv = J.Prover(p, b"Jindo!", tape)
    v = [J.field_set_random(tape, q) for _ in range(length)]
    com_o, open_o = prv.commit(v)
    out = jindo_c.commit(fo, params_dict(p, prv.ecd), polys_array(prv.ck.In), polys_array(prv.ck.MLWE), polys_array(prv.ck.Out),
                         cdt_array(prv.ecd), fo.to_mont(v), tape_arrays(fo, p, prv.trace), threads=threads)
    assert np.array_equal(out["commitment"], polys_array(com_o))
    for i in range(p.cols + 1):
        for j in range(p.rows):
            assert np.array_equal(out["encode"][i, j], polys_array(open_o.Encode[i][j])), ("Encode", i, j)
        for j in range(p.mlwe + p.in_msis):
            assert np.array_equal(out["mlwe"][i, j], polys_array(open_o.MLWE[i][j])), ("MLWE", i, j)
    for k in range(p.in_com_dcmp_len):
        assert np.array_equal(out["incommit"][k], polys_array(open_o.InCommit[k])), ("InCommit", k)
