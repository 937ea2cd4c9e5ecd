"""Buckler Prove on the GPU at the reference's own example size (BASELINE configs[0] / SURVEY 8d cfg1:
examples/mult MultCircuit, N = 2^13, 128-bit field): wall time with its split, checked bit for bit against the
Python restatement of buckler.Prover.Prove on the same witness / CRS / randomness tape and accepted by the
restated Verifier.  The oracle's own time is printed for context only (pure Python, not a performance baseline).

  python tools/bench_buckler.py [logN] [circuit: mult|pk] [field]  ->  one JSON line"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_product
import oracle
from oracle import buckler_ref as B, jindo_ref as J, pyref
from jindo_util import cdt_array, params_dict, polys_array, tape_arrays
from test_oracle_buckler import mult_witness, pk_witness

logn = int(sys.argv[1]) if len(sys.argv) > 1 else 13
circuit = sys.argv[2] if len(sys.argv) > 2 else "mult"
name = sys.argv[3] if len(sys.argv) > 3 else ("zp128" if circuit == "mult" else "zp220")
N = 1 << logn
rs = load_product(); L = rs.load()
assert L.rs_init(0) == 0, L.rs_last_error()
q = pyref.modulus(name)
fo = oracle.Field(q); fg = rs.Field(q)
define, n_pw, n_w = (B.public_key_circuit if circuit == "pk" else B.mult_circuit)(N)
tape = J.Tape(b"buckler-bench-" + circuit.encode())
t0 = time.perf_counter()
C = B.compile_circuit(q, N, define, n_pw, n_w, b"Buckler!", tape)
t_compile = time.perf_counter() - t0
pw, w = (pk_witness if circuit == "pk" else mult_witness)(q, N, 11)
op = B.Prover(C, tape)
t0 = time.perf_counter()
pf_o = op.prove(pw, w)
t_oracle = time.perf_counter() - t0
ctx = C.ctx
terms, coeffs = ctx.terms(ctx.arith_constraints)
desc = {"rank": N, "emb_rank": C.emb_rank, "w_cnt": ctx.w_cnt, "pw_cnt": ctx.pw_cnt, "arith_max_rank": ctx.arith_max_rank,
        "arith_terms": terms, "arith_coeffs": coeffs,
        "lin": [("ntt", ctx.lin_constraints[id(chk)]) for chk in ctx.lin_checkers]}
prv = C.poly_prover
gp = rs.jindo.Prover(fg, params_dict(C.jp, prv.ecd), polys_array(prv.ck.In), polys_array(prv.ck.MLWE),
                     polys_array(prv.ck.Out), cdt_array(prv.ecd), crs=b"Buckler!")
rand = {"blind": op.trace["blind"], "mask": op.trace["mask"], "mask_sum": op.trace["mask_sum"],
        "tapes": [tape_arrays(fo, C.jp, t) for t in op.trace["commit_traces"]]}
gprv = rs.buckler.Prover(fg, q, desc, C.jp, gp)
pwm = np.stack([fo.to_mont(v) for v in op.trace["pw_full"]]); wm = np.stack([fo.to_mont(v) for v in op.trace["w_full"]])
gprv.prove(pwm, wm, rand)                                  # warm-up (twiddle tables, descriptor caches)
times, split = [], None
for _ in range(3):
    t0 = time.perf_counter()
    pf_g = gprv.prove(pwm, wm, rand)
    times.append((time.perf_counter() - t0) * 1e3)
    split = dict(gprv.timing)
same = (pf_g["_eval_point"] == pf_o["_eval_point"]
        and pf_g["Witness"] == [[[list(map(int, l)) for l in poly] for poly in cm] for cm in pf_o["Witness"]]
        and pf_g["Evals"] == pf_o["Evals"] and pf_g["LinCheckMaskSum"] == pf_o["LinCheckMaskSum"]
        and all(pf_g["EvalProof"][k] == [[list(map(int, l)) for l in poly] for poly in pf_o["EvalProof"][k]]
                for k in ("Partial", "Encode", "MLWE", "InCommit")))
t0 = time.perf_counter()
ok = bool(B.Verifier(C).verify(pw, pf_g))
t_verify = time.perf_counter() - t0
print(json.dumps({
    "workload": f"buckler.Prove, {circuit} circuit, N = 2^{logn}, field {name}, embRank {C.emb_rank}, witnesses {ctx.w_cnt}+{ctx.pw_cnt}, "
                f"{len(rand['tapes'])} Jindo commits (rows {C.jp.rows}, cols {C.jp.cols}, slots {C.jp.slots})",
    "gpu_prove_ms": min(times), "gpu_prove_ms_all": times, "split_ms": split,
    "note": "split: device work + the host-side Python glue of ringo-snark_b200/buckler.py (transcript, tape marshalling) between marks",
    "proof_bit_identical_to_oracle": bool(same), "oracle_verifier_accepts": ok,
    "oracle_python_prove_s": t_oracle, "oracle_python_verify_s": t_verify, "oracle_compile_s": t_compile,
}))
