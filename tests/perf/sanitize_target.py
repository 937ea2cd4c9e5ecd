"""[tests/perf: developer harness that takes inputs from / checks against the oracle, hence under tests/ -- the product never imports oracle/]
Small end-to-end exercise of every kernel family for compute-sanitizer (tools/sanitize.sh): big-field NTTs (integer, FP64 and
two-pipe kernels), element-wise ops / evalCircuit / QuoRem / Aut, one Jindo Commit (encode_kernel with its named barriers and
shared queues, matvec_pipe_kernel with its cp.async ring incl. split-K, cut-off kernels, device samplers) and Evaluate.
Results are checked against the oracle where cheap, so a sanitizer run is also a parity run."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_product
import oracle
from oracle import pyref, jindo_ref as J
from jindo_util import cdt_array, params_dict, polys_array, tape_arrays

rs = load_product(); L = rs.load()
assert L.rs_init(0) == 0
big = "--big" in sys.argv
q = pyref.modulus("zp255"); fg = rs.Field(q); fo = oracle.Field(q)
rng = np.random.default_rng(1)


def polys(n, N):
    a = rng.integers(0, 1 << 63, size=(n, N, fo.L), dtype=np.uint64)
    a[..., fo.L - 1] &= np.uint64((1 << ((q.bit_length() - 1) % 64)) - 1)
    return a


# ---- NTT: two-pass rank on every path
for env, n_polys, logN in (({}, 3, 12), ({"RS_NTT52": "1"}, 3, 12), ({"RS_NTT_MIX": "1/4"}, 8 if big else 5, 16 if big else 12)):
    os.environ.update(env)
    N = 1 << logN
    data = polys(n_polys, N)
    for neg in (False, True):
        t = (rs.CyclotomicTransformer if neg else rs.CyclicTransformer)(fg, N)
        A = rs.Batch(fg, n_polys, N).upload(data)
        t.FwdNTTTo(A, A)
        if logN <= 12:
            assert np.array_equal(A.download()[0], oracle.NTT(fo, logN, neg).fwd(data[0])), (env, neg)
        t.InvNTTTo(A, A)
        assert np.array_equal(A.download(), data), (env, neg)
    for k in env:
        del os.environ[k]
# ---- element-wise, QuoRem, Aut
ev = rs.CyclotomicEvaluator(fg, 1024)
B = rs.Batch(fg, 4, 1024).upload(polys(4, 1024))
P = [B.poly(i, is_ntt=True) for i in range(4)]
ev.MulTo(P[2], P[0], P[1]); ev.MulAddTo(P[2], P[0], P[1]); ev.AddTo(P[3], P[0], P[1]); ev.AutTo(P[3], P[0], 5)
assert L.rs_quorem_vanishing(B.h, 2, B.h, 3, B.h, 0, 256) == 0
# ---- Jindo: Commit (host tape, device TwinCDT / mlwe samplers), Evaluate passes
N = 4096 if big else 1024
p = J.new_parameters(q, N, 1)
tape = J.Tape(b"sanitize")
prv = J.Prover(p, b"Jindo!", tape)
gp = rs.jindo.Prover(fg, params_dict(p, prv.ecd), polys_array(prv.ck.In), polys_array(prv.ck.MLWE), polys_array(prv.ck.Out), cdt_array(prv.ecd))
v = [J.field_set_random(tape, q) for _ in range(N)]
com_o, open_o = prv.commit(v)
Bv = rs.Batch(fg, 1, N).upload(fo.to_mont(v).reshape(1, N, fo.L))
t = tape_arrays(fo, p, prv.trace)
com_g, open_g = gp.Commit(Bv, 0, N, t)
assert np.array_equal(com_g, polys_array(com_o))
gp.SetTwinSeed(prv.ecd.twin.base.seed)
t2 = dict(t); t2["twin_u"] = None
gp.Commit(Bv, 0, N, t2)
x = fo.to_mont([12345])[0]
gp.PartialAt(open_g, x)
gp.RespondBytes(open_g, bytes(range(16)) * p.cols)
print("sanitize target: ok", flush=True)
