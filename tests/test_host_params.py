"""CPU test: rs_jindo_new_parameters (csrc/jindo_params.cpp, jindo.NewParameters compiled behind the C ABI) against the oracle's
Python restatement of jindo/params.go:18-321 -- every integer equal, every float64 bit-equal (both run glibc's libm in the
reference's operation order), RNS primes / psi / deltaInv / TwinCDT tables equal -- for every parameter set the test suite and
the benchmarks use, all seven shipped moduli, batch 1..43."""
import ctypes

import numpy as np
import pytest

from oracle import jindo_ref as J, pyref

CASES = [("zp255", 1024, 1), ("zp128", 700, 2), ("zp220", 2049, 9), ("zp240", 600, 3), ("zp255", 1 << 22, 1), ("zp128", 2 * 8192 + 1, 12),
         ("zp240", 3 * (1 << 14) + 1 + (1 << 14) + 1, 43), ("zp110", 513, 12), ("zp440", 4096, 1), ("zp880", 300, 2), ("zp255", 1, 1),
         ("zp128", 1 << 19, 8)]


@pytest.mark.parametrize("name,target_n,batch", CASES)
def test_new_parameters_equal_the_oracle(product, name, target_n, batch):
    rs = product
    L = rs.load()
    q = pyref.modulus(name)
    Lw = (q.bit_length() + 63) // 64
    qa = np.array([(q >> (64 * i)) & ((1 << 64) - 1) for i in range(Lw)], dtype=np.uint64)
    D = rs.capi.JindoDerived()
    assert L.rs_jindo_new_parameters(qa.ctypes.data_as(rs.capi.c_u64p), Lw, target_n, batch, ctypes.byref(D)) == 0, L.rs_last_error()
    p = J.new_parameters(q, target_n, batch)
    P = D.p
    for k in ("d", "slots", "exp", "base", "rows", "cols", "in_msis", "mlwe", "out_msis", "log_in_cut", "log_out_cut"):
        assert getattr(P, k) == getattr(p, k), k
    assert (D.rank, D.batch, D.in_com_dcmp_len) == (p.rank, p.batch, p.in_com_dcmp_len)
    assert list(P.q[:P.n_q]) == list(p.q_primes) and list(P.qout[:P.n_qout]) == list(p.qout_primes)
    assert list(P.psi_q[:P.n_q]) == [s.psi for s in p.ringQ.sub] and list(P.psi_qout[:P.n_qout]) == [s.psi for s in p.ringQOut.sub]
    for got, want in ((D.ecd_std, p.ecd_std), (D.mlwe_std, p.mlwe_std), (D.noise.ecd_blind_std, p.ecd_blind_std), (D.noise.mask_blind_std, p.mask_blind_std),
                      (D.noise.mask_std, p.mask_std), (D.noise.mask_mlwe_std, p.mask_mlwe_std), (D.res_two_nm, p.res_two_nm),
                      (D.in_com_dcmp_two_nm, p.in_com_dcmp_two_nm)):
        assert got == want                                     # bit-equal float64
    ecd = J.Encoder(p, J.Tape(b"params"))
    assert list(P.delta_inv[:P.exp]) == list(ecd.delta_inv)
    assert (P.cdt_len, P.cdt_tail_lo, P.cdt_sigma) == (len(ecd.twin.tables[0]), ecd.twin.tail_lo, ecd.twin.std_dev)
    tabs = np.zeros((128, P.cdt_len), dtype=np.uint64)
    assert L.rs_jindo_cdt_tables(ctypes.byref(D), tabs.ctypes.data_as(rs.capi.c_vp)) == 0
    assert np.array_equal(tabs, np.array(ecd.twin.tables, dtype=np.uint64))


def test_new_parameters_rejects_what_the_reference_rejects(product):
    rs = product
    L = rs.load()
    D = rs.capi.JindoDerived()
    q = pyref.modulus("zp255")
    qa = np.array([(q >> (64 * i)) & ((1 << 64) - 1) for i in range(4)], dtype=np.uint64)
    assert L.rs_jindo_new_parameters(qa.ctypes.data_as(rs.capi.c_u64p), 4, 0, 1, ctypes.byref(D)) != 0         # targetN < 1
    assert L.rs_jindo_new_parameters(qa.ctypes.data_as(rs.capi.c_u64p), 4, 16, 0, ctypes.byref(D)) != 0        # batch < 1
    bad = qa.copy(); bad[0] += np.uint64(2)                                                                     # not b^k + 1
    assert L.rs_jindo_new_parameters(bad.ctypes.data_as(rs.capi.c_u64p), 4, 16, 1, ctypes.byref(D)) != 0
    assert b"jindo-friendly" in L.rs_last_error()
