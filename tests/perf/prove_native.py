"""[tests/perf: developer harness that takes inputs from / checks against the oracle, hence under tests/ -- the product never imports oracle/]
The compiled Buckler prover alone (no oracle run, no Python mirror): loads a cached oracle trace, proves `reps` times through
rs_buckler_prove and prints the wall times -- the command ncu / the launch list is taken of.
  python tests/perf/prove_native.py [trace.pkl] [reps] [--pinned]"""
import ctypes, json, os, pickle, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_product
from oracle import buckler_ref as B, jindo_ref as J, pyref
from jindo_util import cdt_array, params_dict
from test_oracle_buckler import circuit_and_witness

pos = [a for a in sys.argv[1:] if not a.startswith("--")]
trace = pos[0] if pos else os.path.join(ROOT, "oracle_trace", "mult13.pkl")
reps = int(pos[1]) if len(pos) > 1 else 5
saved = pickle.load(open(trace, "rb"))
logn, circuit, name = saved["key"]
rand, pwm, wm = saved["packed"]
N = 1 << logn
q = pyref.modulus(name)
(define, n_pw, n_w), _ = circuit_and_witness(circuit, q, N, 11)
C = B.compile_circuit(q, N, define, n_pw, n_w, b"Buckler!", J.Tape(b"buckler-bench-" + circuit.encode()))
ctx = C.ctx
terms, coeffs = ctx.terms(ctx.arith_constraints)
desc = {"rank": N, "emb_rank": C.emb_rank, "w_cnt": ctx.w_cnt, "pw_cnt": ctx.pw_cnt, "arith_max_rank": ctx.arith_max_rank,
        "arith_terms": terms, "arith_coeffs": coeffs, "lin": [("ntt", ctx.lin_constraints[id(chk)]) for chk in ctx.lin_checkers]}
rs = load_product(); L = rs.load()
assert L.rs_init(0) == 0
fg = rs.Field(q)
gp = rs.jindo.Prover(fg, params_dict(C.jp, C.poly_prover.ecd), None, None, None, cdt_array(C.poly_prover.ecd), crs=b"Buckler!", key_from_crs=True)
nprv = rs.buckler.NativeProver(fg, q, desc, C.jp, gp)
R, keep = nprv.rand_struct(rand, pinned="--pinned" in sys.argv)
nprv.prove_raw(pwm, wm, R)
times = []
for k in range(reps):
    t0 = time.perf_counter()
    out = nprv.prove_raw(pwm, wm, R, phases=False, reuse=True)
    times.append((time.perf_counter() - t0) * 1e3)
nprv.prove_raw(pwm, wm, R, phases=True)
ok = out["eval_point"].tolist() == [int(v) for v in nprv.mont_many([saved["proof"]["_eval_point"]])[0]]
print(json.dumps({"workload": f"{circuit} 2^{logn} {name}", "pinned_tapes": "--pinned" in sys.argv, "prove_ms": times, "best_ms": min(times),
                  "phases_ms_with_syncs": nprv.timing, "eval_point_matches_oracle": bool(ok), "launches_total": int(L.rs_launch_count())}))
