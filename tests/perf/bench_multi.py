"""[tests/perf: measurement harness that checks a sample of its proofs with the oracle's Verifier, hence under tests/ -- the product never imports oracle/]
BASELINE configs[4]: Buckler multi-circuit batched proving, whole circuits sharded across the GPUs of one box
(SURVEY 8e ii: circuits are independent; every Prove is sequential inside through its Fiat-Shamir rounds), with the
NCCL all-gather of the per-circuit first-round commitments that the host transcripts bind (rs_allgather_u64, C ABI).

    python tests/perf/bench_multi.py [--instances=256] [--logn=13] [--threads=4] [--verify=2]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tests/perf/bench_multi.py ...

256 examples/mult MultCircuit instances at N = 2^13 are 256 * 8 * 2^13 = 2^24 witness coefficients.
Every instance is DISTINCT: its own X (|X| <= 5), Y (63 bit), Z = X * Y in R_q (computed with the library's negacyclic NTT
before the timed region, like the user of the reference fills the circuit before Prove), and its own randomness -- the compiled
prover (rs_buckler_prove with rand = NULL) draws it inside the timed region: COSAC rows, mask noise and MustSetRandom
elements on `threads` host worker threads per rank, TwinCDT / MLWE rows from the device streams.  Rank r proves the
instances shard_range(instances, r, world); total work is fixed (strong scaling).  `verify` proofs of rank 0 are checked
by the restated buckler.Verifier after the timed region.  One JSON line on rank 0.
"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_product
import oracle
from oracle import buckler_ref as B, jindo_ref as J, pyref
from jindo_util import cdt_array, params_dict
import torch
import torch.distributed as dist


def opt(name, default):
    return next((a[len(name) + 3:] for a in sys.argv[1:] if a.startswith("--" + name + "=")), default)


instances = int(opt("instances", "256")); logn = int(opt("logn", "13")); threads = int(opt("threads", "4")); n_verify = int(opt("verify", "2"))
device_cosac = opt("device-cosac", "1") != "0"   # COSAC rows sampled on the device (rs_jindo_set_cosac_stream) instead of the host threads
n_streams = int(opt("streams", "1"))     # concurrent provers per GPU: one host thread + CUDA stream + prover handle each
circuit, name = "mult", "zp128"
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
N = 1 << logn
q = pyref.modulus(name)
fo = oracle.Field(q)
define, n_pw, n_w = B.mult_circuit(N)
C = B.compile_circuit(q, N, define, n_pw, n_w, b"Buckler!", J.Tape(b"buckler-bench-" + circuit.encode()))

rs = load_product(); L = rs.load()
rs.check(L.rs_init(local))
fg = rs.Field(q)
ctx = C.ctx
terms, coeffs = ctx.terms(ctx.arith_constraints)
desc = {"rank": N, "emb_rank": C.emb_rank, "w_cnt": ctx.w_cnt, "pw_cnt": ctx.pw_cnt, "arith_max_rank": ctx.arith_max_rank,
        "arith_terms": terms, "arith_coeffs": coeffs,
        "lin": [("ntt", ctx.lin_constraints[id(chk)]) for chk in ctx.lin_checkers]}
p = C.jp


def make_prover():
    """one jindo.Prover + buckler.Prover pair, like Prover.SafeCopy hands every goroutine (jindo/prover.go:327-338)"""
    gp = rs.jindo.Prover(fg, params_dict(p, C.poly_prover.ecd), None, None, None, cdt_array(C.poly_prover.ecd), crs=b"Buckler!", key_from_crs=True)
    gp.SetTwinSeed(os.urandom(32)); gp.SetMlweSeed(os.urandom(32), p.mlwe_std)
    pr = rs.buckler.NativeProver(fg, q, desc, p, gp)
    noise = {"ecd_blind_std": p.ecd_blind_std, "mask_blind_std": p.mask_blind_std, "mask_std": p.mask_std, "mask_mlwe_std": p.mask_mlwe_std}
    pr.enable_sampling(noise, threads=threads)
    if device_cosac:
        gp.SetCosacSeeds(os.urandom(32), os.urandom(32), noise)
        pr.set_device_cosac(True)
    return pr


provers = [make_prover() for _ in range(n_streams)]
nprv = provers[0]
par = rs.parallel
a, b = par.shard_range(instances, rank, world)

# ---- distinct instances (outside the timed region: the reference's user fills the circuit before calling Prove):
# ringo-snark_b200/circuits.py: X uniform in [0, 5], Y uniform 63-bit, Z = X * Y through the library's negacyclic NTT, ternary digits
neg = rs.CyclotomicTransformer(fg, N)


def make_instance(seed):
    return rs.circuits.mult_witness(fg, q, N, seed, transformer=neg)


inst = [make_instance(1000 + i) for i in range(a, b)]


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()


comm = par.NcclComm(rank, world) if world > 1 else None
for pr in provers:
    pr.prove_raw(inst[0][0], inst[0][1], None)                 # warm-up (twiddle tables, descriptor caches, pinned arenas)
streams = [torch.cuda.Stream() for _ in range(n_streams)]
firsts, proofs, raws = [None] * len(inst), [None] * len(inst), [None] * len(inst)


def run(s):
    """prover s on its own CUDA stream (the library's current stream is per host thread; ctypes drops the GIL during calls)"""
    rs.check(L.rs_init(local)); rs.check(L.rs_set_stream(streams[s].cuda_stream))
    for k in range(s, len(inst), n_streams):
        o = provers[s].prove_raw(inst[k][0], inst[k][1], None, phases=False)
        firsts[k] = o["Witness"][0].reshape(-1).copy()
        raws[k] = o
        if k < n_verify:
            proofs[k] = o


barrier()
t0 = time.perf_counter()
if n_streams == 1:
    run(0)
else:
    import threading
    ths = [threading.Thread(target=run, args=(s,)) for s in range(n_streams)]
    for th in ths: th.start()
    for th in ths: th.join()
    rs.check(L.rs_set_stream(None))
proofs = [o for o in proofs if o is not None]
# the per-circuit commitments every rank's transcript would bind (prover.go:155-157): one all-gather through the C ABI
local_c = np.stack(firsts)
if world > 1:
    per = max(bb - aa for aa, bb in (par.shard_range(instances, r, world) for r in range(world)))
    pad = np.zeros((per, local_c.shape[1]), dtype=np.uint64); pad[:len(firsts)] = local_c
    g = comm.all_gather_u64(pad).reshape(world, per, -1)
    allc = np.concatenate([g[r, :bb - aa] for r, (aa, bb) in enumerate(par.shard_range(instances, r, world) for r in range(world))])
else:
    allc = local_c
barrier()
dt = time.perf_counter() - t0
t = torch.tensor([dt], device="cuda")
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
dt = float(t[0])
# every proof of this rank through the compiled Verifier (rs_buckler_verify), outside the timed region
tv = time.perf_counter()
n_ok = sum(1 for k, o in enumerate(raws) if nprv.verify_raw(inst[k][0], o) == 0)
verify_ms = (time.perf_counter() - tv) * 1e3 / max(1, len(raws))
tt = torch.tensor([n_ok, len(raws)], device="cuda")
if world > 1:
    dist.all_reduce(tt)
if rank == 0:
    # a sample of the proofs through the restated Verifier (pure Python, ~1.5 s each; outside the timed region)
    V = B.Verifier(C)
    accepted = []
    for k, o in enumerate(proofs):
        pf = {"Witness": o["Witness"].tolist(), "LinCheckMaskSum": nprv.from_mont_arr(o["lin_mask_sum"]), "SumCheckMaskSum": None,
              "Evals": [nprv.from_mont_arr(e) for e in o["Evals"]],
              "EvalProof": {kk: o[kk].tolist() for kk in ("Partial", "PartialMask", "InCommit", "Encode", "MLWE")}}
        pw_ints = [fo.from_mont(inst[k][0][0])]
        accepted.append(bool(V.verify(pw_ints, pf)))
    distinct = len({c.tobytes() for c in allc}) == len(allc)
    coeffs_total = instances * (ctx.w_cnt + ctx.pw_cnt) * N
    print(json.dumps({
        "workload": f"buckler multi-circuit proving: {instances} distinct {circuit} circuits, N = 2^{logn}, field {name} "
                    f"({ctx.w_cnt}+{ctx.pw_cnt} witnesses, {ctx.batch()} Jindo commits each) = 2^{np.log2(coeffs_total):.2f} witness coefficients",
        "n_gpus": world, "scaling": "strong", "seconds": dt, "proofs_per_s": instances / dt,
        "witness_mcoeff_per_s": coeffs_total / dt / 1e6, "ms_per_proof_per_gpu": dt * 1e3 / max(1, b - a),
        "cosac": "device" if device_cosac else "host threads", "cosac_overflow": int(sum(pr.gp.CosacStatus()[1] for pr in provers)) if device_cosac else None,
        "concurrent_provers_per_gpu": n_streams, "host_sampler_threads_per_prover": threads, "host_cpus": os.cpu_count(),
        "gathered_commitments": list(allc.shape), "all_commitments_distinct": bool(distinct), "transcript_digest": par.transcript_digest(allc),
        "compiled_verifier_accepted": int(tt[0]), "compiled_verifier_checked": int(tt[1]), "verify_ms_per_proof": verify_ms,
        "restated_verifier_sample": accepted,
        "prover": "rs_buckler_prove (compiled, include/ringo_buckler.h), rand = NULL: randomness drawn inside the timed region",
        "exchange": "rs_allgather_u64 (NCCL behind the C ABI)" if world > 1 else "none (1 GPU)",
        "data": "distinct synthetic instances: X uniform in [0,5], Y uniform 63-bit, Z = X*Y in R_q (library NTT), ternary digits of X",
    }))
if comm is not None:
    comm.close()
if world > 1:
    dist.destroy_process_group()
