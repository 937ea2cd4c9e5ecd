"""[tests/perf: measurement harness that takes its inputs from the oracle, hence under tests/ -- the product never imports oracle/]
BASELINE configs[4]: Buckler multi-circuit batched proving, whole circuits sharded across the GPUs of one box
(SURVEY 8e ii: circuits are independent; every Prove is sequential inside through its Fiat-Shamir rounds), with the
NCCL all-gather of the per-circuit first-round commitments that the host transcripts bind.

    python tests/perf/bench_multi.py [--instances=256] [--logn=13] [--circuit=mult] --trace=oracle_trace/mult13.pkl
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tests/perf/bench_multi.py ...

256 MultCircuit instances at N = 2^13 are 256 * 8 * 2^13 = 2^24 witness coefficients.  The oracle
produces ONE witness / randomness trace (minutes of pure Python per instance), so every instance replays that trace:
the device work and the host glue per instance are those of distinct instances, the data are not.  Rank r proves the
instances shard_range(instances, r, world); weak/strong: total work is fixed (strong scaling).  One JSON line on rank 0.
"""
import json, os, pickle, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_product
import oracle
from oracle import buckler_ref as B, jindo_ref as J, pyref
from jindo_util import cdt_array, params_dict
from test_oracle_buckler import ct_params, BFV_T
import torch
import torch.distributed as dist


def opt(name, default):
    return next((a[len(name) + 3:] for a in sys.argv[1:] if a.startswith("--" + name + "=")), default)


instances = int(opt("instances", "256")); logn = int(opt("logn", "13")); circuit = opt("circuit", "mult")
name = {"mult": "zp128", "pk": "zp220", "ct": "zp240"}[circuit]
trace_path = opt("trace", os.path.join(ROOT, "oracle_trace", f"{circuit}{logn}.pkl"))
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
N = 1 << logn
q = pyref.modulus(name)
fo = oracle.Field(q)
if circuit == "ct":                                            # (the witness itself comes from the trace)
    delta, eb = ct_params(q, N)
    define, n_pw, n_w = B.ciphertext_circuit(N, delta, BFV_T, eb)
else:
    define, n_pw, n_w = {"mult": B.mult_circuit, "pk": B.public_key_circuit}[circuit](N)
C = B.compile_circuit(q, N, define, n_pw, n_w, b"Buckler!", J.Tape(b"buckler-bench-" + circuit.encode()))
with open(trace_path, "rb") as fh:
    saved = pickle.load(fh)
assert saved["key"] == (logn, circuit, name)
rand, pwm, wm = saved["packed"]
pf_o = saved["proof"]

rs = load_product(); L = rs.load()
rs.check(L.rs_init(local))
fg = rs.Field(q)
ctx = C.ctx
terms, coeffs = ctx.terms(ctx.arith_constraints)
desc = {"rank": N, "emb_rank": C.emb_rank, "w_cnt": ctx.w_cnt, "pw_cnt": ctx.pw_cnt, "arith_max_rank": ctx.arith_max_rank,
        "arith_terms": terms, "arith_coeffs": coeffs,
        "lin": [("ntt", ctx.lin_constraints[id(chk)]) for chk in ctx.lin_checkers]}
gp = rs.jindo.Prover(fg, params_dict(C.jp, C.poly_prover.ecd), None, None, None, cdt_array(C.poly_prover.ecd), crs=b"Buckler!", key_from_crs=True)
gprv = rs.buckler.Prover(fg, q, desc, C.jp, gp)
par = rs.parallel
a, b = par.shard_range(instances, rank, world)


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()


pf = gprv.prove(pwm, wm, rand)                                  # warm-up
ok = pf["_eval_point"] == pf_o["_eval_point"] and pf["Evals"] == pf_o["Evals"]
barrier()
t0 = time.perf_counter()
firsts = []
split = {}
for _ in range(a, b):
    pf = gprv.prove(pwm, wm, rand)
    firsts.append(np.array(pf["Witness"][0], dtype=np.uint64).reshape(-1))
    for k, v in gprv.timing.items():
        split[k] = split.get(k, 0.0) + v
# the per-circuit commitments every rank's transcript would bind (prover.go:155-157): one all-gather
allc = par.all_gather_commitments(np.stack(firsts) if firsts else np.zeros((0, 1), dtype=np.uint64), instances, rank, world,
                                  device=torch.device("cuda", local) if world > 1 else None)
barrier()
dt = time.perf_counter() - t0
t = torch.tensor([dt], device="cuda")
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
dt = float(t[0])
if rank == 0:
    coeffs_total = instances * (ctx.w_cnt + ctx.pw_cnt) * N
    print(json.dumps({
        "workload": f"buckler multi-circuit proving: {instances} x {circuit} circuit, N = 2^{logn}, field {name} "
                    f"({ctx.w_cnt}+{ctx.pw_cnt} witnesses, {len(rand['tapes'])} Jindo commits each) = 2^{np.log2(coeffs_total):.2f} witness coefficients",
        "n_gpus": world, "scaling": "strong", "seconds": dt, "proofs_per_s": instances / dt,
        "witness_mcoeff_per_s": coeffs_total / dt / 1e6, "ms_per_proof_per_gpu": dt * 1e3 / max(1, b - a),
        "rank0_split_ms_per_proof": {k: v / max(1, b - a) for k, v in split.items()},
        "gathered_commitments": list(allc.shape), "transcript_digest": par.transcript_digest(allc),
        "first_proof_matches_oracle": bool(ok),
        "data": "one oracle witness / randomness trace replayed for every instance (distinct instances would take minutes of pure-Python oracle each)",
        "note": "wall time incl. the Python rendering of the Go shim (ringo-snark_b200/buckler.py); commit key generated on the device from the CRS",
    }))
if world > 1:
    dist.destroy_process_group()
