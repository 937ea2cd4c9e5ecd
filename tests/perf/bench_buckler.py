"""[tests/perf: measurement harness that checks against / times the oracle, hence under tests/ -- the product never imports oracle/]
Buckler Prove on the GPU at the reference's own example size (BASELINE configs[0] / SURVEY 8d cfg1:
examples/mult MultCircuit, N = 2^13, 128-bit field): wall time with its split, checked bit for bit against the
Python restatement of buckler.Prover.Prove on the same witness / CRS / randomness tape and accepted by the
restated Verifier.  The oracle's own time is printed for context only (pure Python, not a performance baseline).

  python tests/perf/bench_buckler.py [logN] [circuit: mult|pk|ct] [field]  ->  one JSON line"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_product
import oracle
from oracle import buckler_ref as B, jindo_ref as J, pyref
from jindo_util import cdt_array, params_dict, polys_array, tape_arrays
from test_oracle_buckler import circuit_and_witness

import pickle
pos = [a for a in sys.argv[1:] if not a.startswith("--")]
# --trace=FILE: the oracle's proof and randomness trace are cached there (the pure-Python oracle takes minutes at the bfv
# sizes; it can run once on a CPU-only machine and the GPU box then only loads it).  --oracle-only: stop after writing it.
trace_path = next((a[8:] for a in sys.argv[1:] if a.startswith("--trace=")), None)
oracle_only = "--oracle-only" in sys.argv
logn = int(pos[0]) if len(pos) > 0 else 13
circuit = pos[1] if len(pos) > 1 else "mult"
name = pos[2] if len(pos) > 2 else {"mult": "zp128", "pk": "zp220", "ct": "zp240"}[circuit]
N = 1 << logn
q = pyref.modulus(name)
fo = oracle.Field(q)
(define, n_pw, n_w), (pw, w) = circuit_and_witness(circuit, q, N, 11)
tape = J.Tape(b"buckler-bench-" + circuit.encode())
t0 = time.perf_counter()
C = B.compile_circuit(q, N, define, n_pw, n_w, b"Buckler!", tape)
t_compile = time.perf_counter() - t0
def slim(pf):
    """the proof without the oracle's debugging attachments (whole Openings)"""
    return {k: v for k, v in pf.items() if not k.startswith("_") or k == "_eval_point"}


def pack(trace):
    """the oracle's randomness trace as the plain arrays the GPU prover takes (numpy: ~5x smaller than the int lists)"""
    rand = {"blind": trace["blind"], "mask": trace["mask"], "mask_sum": trace["mask_sum"],
            "tapes": [tape_arrays(fo, C.jp, t) for t in trace["commit_traces"]]}
    return rand, np.stack([fo.to_mont(v) for v in trace["pw_full"]]), np.stack([fo.to_mont(v) for v in trace["w_full"]])


if trace_path and os.path.exists(trace_path):
    with open(trace_path, "rb") as fh:
        saved = pickle.load(fh)
    assert saved["key"] == (logn, circuit, name), "trace file belongs to another workload"
    pf_o, t_oracle = slim(saved["proof"]), saved["t_oracle"]
    rand, pwm, wm = saved["packed"] if "packed" in saved else pack(saved["trace"])
    if "packed" not in saved or len(pf_o) != len(saved["proof"]):      # rewrite an old-format file compactly
        with open(trace_path, "wb") as fh:
            pickle.dump({"key": saved["key"], "proof": pf_o, "packed": (rand, pwm, wm), "t_oracle": t_oracle}, fh, protocol=4)
else:
    op = B.Prover(C, tape)
    t0 = time.perf_counter()
    pf_o = slim(op.prove(pw, w))
    t_oracle = time.perf_counter() - t0
    rand, pwm, wm = pack(op.trace)
    if trace_path:
        os.makedirs(os.path.dirname(os.path.abspath(trace_path)), exist_ok=True)
        with open(trace_path, "wb") as fh:
            pickle.dump({"key": (logn, circuit, name), "proof": pf_o, "packed": (rand, pwm, wm), "t_oracle": t_oracle}, fh, protocol=4)
if oracle_only:
    print(json.dumps({"oracle_python_prove_s": t_oracle, "trace": trace_path})); sys.exit(0)
rs = load_product(); L = rs.load()
assert L.rs_init(0) == 0, L.rs_last_error()
fg = rs.Field(q)
ctx = C.ctx
terms, coeffs = ctx.terms(ctx.arith_constraints)
desc = {"rank": N, "emb_rank": C.emb_rank, "w_cnt": ctx.w_cnt, "pw_cnt": ctx.pw_cnt, "arith_max_rank": ctx.arith_max_rank,
        "arith_terms": terms, "arith_coeffs": coeffs,
        "lin": [("ntt", ctx.lin_constraints[id(chk)]) for chk in ctx.lin_checkers]}
prv = C.poly_prover
# commit key: NewCommitKey(params, crs) evaluated on the device (rs_jindo_create_crs); the oracle's key only sizes nothing here
gp = rs.jindo.Prover(fg, params_dict(C.jp, prv.ecd), None, None, None, cdt_array(prv.ecd), crs=b"Buckler!", key_from_crs=True)
gprv = rs.buckler.Prover(fg, q, desc, C.jp, gp)
gprv.prove(pwm, wm, rand)                                  # warm-up (twiddle tables, descriptor caches)
times, split = [], None
for _ in range(3):
    t0 = time.perf_counter()
    pf_g = gprv.prove(pwm, wm, rand)
    times.append((time.perf_counter() - t0) * 1e3)
    split = dict(gprv.timing)
# the compiled prover (rs_buckler_prove): inputs marshalled once, as a Go caller holding its tapes in C memory would
nprv = rs.buckler.NativeProver(fg, q, desc, C.jp, gp)
R, keep = nprv.rand_struct(rand)
nprv.prove_raw(pwm, wm, R)
n_times, n_split = [], None
for k in range(5):
    t0 = time.perf_counter()
    nprv.prove_raw(pwm, wm, R, phases=(k == 0))
    n_times.append((time.perf_counter() - t0) * 1e3)
    if k == 0:
        n_split = dict(nprv.timing)
pf_n = nprv.prove(pwm, wm, rand)
same_native = all(pf_n[k] == pf_g[k] for k in ("Witness", "Evals", "LinCheckMaskSum", "_eval_point")) and pf_n["EvalProof"] == pf_g["EvalProof"]
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
    "compiled_prove_ms": min(n_times[1:]), "compiled_prove_ms_all": n_times, "compiled_split_ms": n_split,
    "compiled_proof_equals_mirror": bool(same_native),
    "gpu_prove_ms": min(times), "gpu_prove_ms_all": times, "split_ms": split,
    "note": "split: device work + the host-side Python glue of ringo-snark_b200/buckler.py (transcript, tape marshalling) between marks",
    "proof_bit_identical_to_oracle": bool(same), "oracle_verifier_accepts": ok,
    "oracle_python_prove_s": t_oracle, "oracle_python_verify_s": t_verify, "oracle_compile_s": t_compile,
}))
