"""[tests/perf: developer harness that takes inputs from / checks against the oracle, hence under tests/ -- the product never imports oracle/]
Host-side cost of the small C-ABI calls a Buckler Prove is made of (mult circuit, N = 2^13 shapes): wall time per call with
the stream drained before and after.  python tests/perf/hostcost.py [trace.pkl]"""
import json, os, pickle, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_product
import oracle
from oracle import buckler_ref as B, jindo_ref as J, pyref
from jindo_util import cdt_array, params_dict
from test_oracle_buckler import circuit_and_witness

trace = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "oracle_trace", "mult13.pkl")
saved = pickle.load(open(trace, "rb"))
logn, circuit, name = saved["key"]
rand, pwm, wm = saved["packed"]
N = 1 << logn
q = pyref.modulus(name)
(define, n_pw, n_w), _ = circuit_and_witness(circuit, q, N, 11)
C = B.compile_circuit(q, N, define, n_pw, n_w, b"Buckler!", J.Tape(b"buckler-bench-" + circuit.encode()))
rs = load_product(); L = rs.load()
assert L.rs_init(0) == 0
fg = rs.Field(q)
gp = rs.jindo.Prover(fg, params_dict(C.jp, C.poly_prover.ecd), None, None, None, cdt_array(C.poly_prover.ecd), crs=b"Buckler!", key_from_crs=True)
v = rs.Batch(fg, 1, 4 * N).upload(np.random.default_rng(1).integers(0, 1 << 60, size=(1, 4 * N, fg.L), dtype=np.uint64))
T, keep = gp.tape_struct(rand["tapes"][0])


def timed(fn, reps=20):
    fn(); L.rs_sync()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    L.rs_sync()
    return (time.perf_counter() - t0) * 1e3 / reps

out = {}
com = np.zeros((gp.p["out_msis"], gp.n_qout, gp.d), dtype=np.uint64)
op0 = gp.Commit(v, 0, N + 1, T)[1]
out["commit_reuse_opening_ms"] = timed(lambda: gp.CommitInto(v, 0, N + 1, T, op0, com))
ops = []
out["commit_new_opening_ms"] = timed(lambda: ops.append(gp.Commit(v, 0, N + 1, T)[1]))
acc = gp.NewOpening()
out["opening_mul_add_bytes_ms"] = timed(lambda: gp.OpeningMulAddBytes(acc, op0, b"0123456789abcdef"))
out["encode_challenges_ms"] = timed(lambda: gp.EncodeChallenges(b"0123456789abcdef"))
cq = gp.EncodeChallenges(b"0123456789abcdef"); co = gp.EncodeChallenges(b"0123456789abcdef", ring_out=True)
out["opening_mul_add_hostvec_ms"] = timed(lambda: gp.OpeningMulAdd(acc, op0, cq, co))
xm = np.ones(fg.L, dtype=np.uint64)
out["partial_at_ms"] = timed(lambda: gp.PartialAt(acc, xm))
res = np.zeros(fg.L, dtype=np.uint64)
from ctypes import c_void_p
u64p = rs.capi.c_u64p
out["poly_evaluate_ms"] = timed(lambda: L.rs_poly_evaluate(v.h, 0, N + 1, xm.ctypes.data_as(u64p), res.ctypes.data_as(u64p)))
out["sync_only_ms"] = timed(lambda: L.rs_sync())
out["opening_words"] = {"encode": (gp.p["cols"] + 1) * gp.p["rows"] * gp.n_q * gp.d, "n_q": gp.n_q, "n_qout": gp.n_qout}
print(json.dumps(out))
