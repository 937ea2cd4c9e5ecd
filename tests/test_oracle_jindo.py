"""CPU tests for the Jindo oracle (oracle/jindo_ref.py).

The reference's own tests only assert Verify(Prove) == true (jindo/jindo_test.go:16-52); the
same self-check is run on the restatement, together with parameter-table and sampler checks."""
import hashlib

import numpy as np
import pytest
from cryptography.hazmat.primitives.ciphers import Cipher, algorithms, modes

from oracle import jindo_ref as J, pyref

# derived Jindo parameter sets of SURVEY.md section 8 (rows x cols, slots, limbs x bits, ranks, cut-offs)
PARAM_TABLE = [
    ("zp128", 16385, 12, (130, 4, 32), (3, 51), (3, 52), (21, 32, 22), (41, 112)),
    ("zp255", 1 << 22, 1, (2049, 128, 16), (2, 38), (1, 44), (10, 32, 6), (44, 35)),
    ("zp240", 32769, 43, (258, 8, 16), (2, 55), (2, 51), (15, 32, 14), (40, 70)),
    ("zp240", 16385, 43, (258, 4, 16), (2, 55), (2, 50), (15, 32, 14), (40, 70)),
    ("zp255", 1024, 1, (33, 2, 16), (2, 35), (1, 38), (10, 32, 6), (41, 29)),
    ("zp255", 1024, 8, (65, 1, 16), (2, 54), (2, 48), (15, 32, 13), (41, 66)),
    ("zp220", 2049, 9, (66, 2, 16), (2, 52), (2, 48), (14, 32, 13), (37, 67)),
]


@pytest.mark.parametrize("name,N,batch,shape,rq,rqo,ranks,cuts", PARAM_TABLE)
def test_parameter_search(name, N, batch, shape, rq, rqo, ranks, cuts):
    p = J.new_parameters(pyref.modulus(name), N, batch)
    assert (p.rows, p.cols, p.slots) == shape
    assert (len(p.q_primes), p.q_bits) == rq and (len(p.qout_primes), p.qout_bits) == rqo
    assert (p.in_msis, p.mlwe, p.out_msis) == ranks
    assert (p.log_in_cut, p.log_out_cut) == cuts
    for q in p.q_primes + p.qout_primes:
        assert J.is_prime64(q) and q % (2 * p.d) == 1
    assert p.base ** p.exp + 1 == pyref.modulus(name)


def test_uniform_sampler_stream_quirk():
    """uniform.go:64-70: refill n of the 8 KiB buffer is the XOR of keystream chunks 0..n."""
    seed = b"Jindo!"
    h = hashlib.sha384(seed).digest()
    ks = Cipher(algorithms.AES(h[:32]), modes.CTR(h[32:48])).encryptor().update(bytes(3 * 8192))
    u = J.UniformSampler(seed)
    first = [u.sample() for _ in range(1024)]
    assert first[0] == int.from_bytes(ks[:8], "little") and first[-1] == int.from_bytes(ks[8184:8192], "little")
    second = u.sample()
    assert second == int.from_bytes(ks[:8], "little") ^ int.from_bytes(ks[8192:8200], "little")
    u2 = J.UniformSampler(seed)
    assert list(u2.sample_many(1500)) == first + [second] + [u.sample() for _ in range(475)]


def test_twin_cdt_moments_and_ring_ntt():
    tape = J.Tape(b"x")
    s = J.TwinCDTGaussianSampler(4.787, tape.read(32))
    xs = [s.sample(0.3) for _ in range(20000)]
    mean = sum(xs) / len(xs)
    var = sum((x - mean) ** 2 for x in xs) / len(xs)
    assert abs(mean - 0.3) < 0.15 and abs(var - 4.787 ** 2) < 1.5
    # [lattigo]-style NTT: evaluation at psi^(2 brv(i)+1), and round trip
    q = J.next_upstream_primes(38, 512, 1)[0]
    sr = J.SubRing(256, q)
    a = [(i * 7919 + 13) % q for i in range(256)]
    A = sr.ntt(a)
    for i in (0, 1, 77, 255):
        x = pow(sr.psi, 2 * J.brv(i, 8) + 1, q)
        assert A[i] == sum(c * pow(x, k, q) for k, c in enumerate(a)) % q
    assert sr.intt(A) == a


@pytest.mark.parametrize("name,N,batch", [("zp255", 1024, 1), ("zp255", 256, 8), ("zp128", 700, 2)])
def test_jindo_commit_evaluate_verify(name, N, batch):
    """TestJindo (jindo/jindo_test.go:16-52) on the restatement: Verify(Evaluate(Commit)) == true."""
    q = pyref.modulus(name)
    p = J.new_parameters(q, N, batch)
    tape = J.Tape(b"jindo-test-%d" % batch)
    prv = J.Prover(p, b"Jindo!", tape)
    vrf = J.Verifier(p, b"Jindo!", tape, commit_key=prv.ck)
    vs = [[J.field_set_random(tape, q) for _ in range(N)] for _ in range(batch)]
    coms, opens = [], []
    for v in vs:
        c, o = prv.commit(v)
        coms.append(c); opens.append(o)
    x = J.field_set_random(tape, q)
    ys, pf = prv.evaluate(x, vs, coms, opens)
    assert ys == [sum(c * pow(x, i, q) for i, c in enumerate(v)) % q for v in vs]
    assert vrf.verify(x, coms, ys, pf)
    bad = list(ys); bad[0] = (bad[0] + 1) % q
    assert not vrf.verify(x, coms, bad, pf)
    # decoding an encoding returns the encoded values (encoder.go:204-219)
    enc = prv.ecd.encode(vs[0][:p.slots])
    assert prv.ecd.decode(p.ringQ.intt(p.ringQ.imform(enc)))[:p.slots] == vs[0][:p.slots]


def test_product_host_parameter_mirror_matches_oracle(product):
    """ringo-snark_b200/jindo_host.py (inputs of the C ABI) derives the same parameters, primes, psi,
    deltaInv and CDT tables as the oracle's restatement."""
    JH = product.jindo_host
    for name, N, batch in [("zp255", 1 << 22, 1), ("zp128", 16385, 12), ("zp220", 2049, 9)]:
        q = pyref.modulus(name)
        po = J.new_parameters(q, N, batch); ph = JH.new_parameters(q, N, batch)
        for k in ("rows", "cols", "slots", "exp", "base", "in_msis", "mlwe", "out_msis", "log_in_cut", "log_out_cut",
                  "q_primes", "qout_primes", "rank", "in_com_dcmp_len", "ecd_std", "mlwe_std"):
            assert getattr(po, k) == getattr(ph, k), (name, k)
        enc = J.Encoder(po, J.Tape(b"x"))
        d, tables = JH.prover_inputs(ph)
        assert d["delta_inv"] == enc.delta_inv and d["cdt_tail_lo"] == enc.twin.tail_lo
        assert d["psi_q"] == [s.psi for s in po.ringQ.sub] and d["psi_qout"] == [s.psi for s in po.ringQOut.sub]
        assert np.array_equal(tables, np.array(enc.twin.tables, dtype=np.uint64))
