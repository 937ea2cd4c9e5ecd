#!/usr/bin/env python3
"""bench.py -- contract benchmark (see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A step is one pass of the hot path over one batch of synthetic input: the negacyclic forward
NTT (math/bigpoly FwdNTTTo) of `batch` polynomials of rank 2^16 over the 255-bit
jindo-modulus field (BASELINE.json configs[1] headline: logN=16, 4 x uint64 limbs, batch 512),
followed -- when the Jindo path is built -- by the Commit leg reported under "commit".
Prints ONE JSON line on rank 0.

  value     device-resident throughput (inputs in HBM before the timed region), CUDA events
  e2e       same metric through the C ABI with HOST buffers: pinned host -> H2D -> NTT -> D2H
  roofline  dominant kernel vs measured HBM peak (schema-mandated) and, in roofline_imad, vs the
            measured integer-pipe peak, which is what actually binds this kernel
  cpu_baseline  the C oracle (restated reference) on a bounded sample, on the box's host cores

`--impl reference` times the restated reference CPU implementation (oracle/; the Go
toolchain is absent, so the reference itself cannot run) with all host threads.
"""
import argparse
import ctypes
import importlib.util
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FIELD = "zp255"
LOGN = 16
BATCH = 512
METRIC = "bigpoly NTT throughput (rank 2^16, 255-bit jindo field, negacyclic forward)"
UNIT = "Mcoeff/s"


def product():
    name = "ringo_snark_b200"
    if name in sys.modules:
        return sys.modules[name]
    pkg = os.path.join(ROOT, "ringo-snark_b200")
    spec = importlib.util.spec_from_file_location(name, os.path.join(pkg, "__init__.py"),
                                                  submodule_search_locations=[pkg])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def synthetic_polys(q, L, n_polys, N, seed):
    """uniform canonical-range limbs (top bits cleared so every value is < q), Montgomery form is
    just another uniform residue, so the raw limbs are used directly as Montgomery residues."""
    import numpy as np
    rng = np.random.default_rng(seed)
    a = rng.integers(0, 1 << 63, size=(n_polys, N, L), dtype=np.uint64)
    a[..., L - 1] &= np.uint64((1 << ((q.bit_length() - 1) % 64)) - 1)
    return a


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self.stop_flag = False

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0]))
                self.max_mhz = float(out[1])
                for n, v in zip(names, out[2:]):
                    if v.strip().lower() == "active":
                        self.reasons.add(n)
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def bench_config():
    """The workload both arms are quoted on (identical dict in the product line and the --impl reference line; the reference arm
    says in cpu_baseline.sample how much of it one of its steps covers)."""
    nbytes = BATCH * (1 << LOGN) * 32
    return {"workload": f"bigpoly negacyclic forward NTT, {FIELD} (jindo/internal/zp), rank 2^{LOGN}, "
                        f"batch {BATCH} polys per GPU (BASELINE configs[1] headline)",
            "l2": f"input+output {2 * nbytes >> 20} MiB per step, larger than the 126 MB L2 (no flush needed)",
            "sharding": "independent polynomials per rank, no data-path collective"}


def cpu_reference_rate(threads, seconds_target=12.0):
    """The restated reference (oracle C) on a bounded sample of the same workload."""
    import numpy as np
    import oracle
    from oracle import pyref
    q = pyref.modulus(FIELD)
    f = oracle.Field(q)
    t = oracle.NTT(f, LOGN, True)
    N = 1 << LOGN
    n_polys = max(threads, 8)
    data = synthetic_polys(q, f.L, n_polys, N, 7)
    t0 = time.perf_counter(); t.fwd_batch_inplace(data, threads); dt = time.perf_counter() - t0
    reps = max(1, int(seconds_target / max(dt, 1e-3)))
    reps = min(reps, 50)
    t0 = time.perf_counter()
    for _ in range(reps):
        t.fwd_batch_inplace(data, threads)
    dt = (time.perf_counter() - t0) / reps
    return n_polys * N / dt / 1e6, n_polys, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    import numpy as np
    import oracle
    from oracle import pyref
    q = pyref.modulus(FIELD)
    f = oracle.Field(q)
    t = oracle.NTT(f, LOGN, True)
    N = 1 << LOGN
    n_polys = max(threads, 8)          # bounded sample of the 512-poly batch
    data = synthetic_polys(q, f.L, n_polys, N, 7)
    for _ in range(args.warmup):
        t.fwd_batch_inplace(data, threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        t.fwd_batch_inplace(data, threads)
    dt = (time.perf_counter() - t0) / args.steps
    val = n_polys * N / dt / 1e6
    # second half of the metric: one Jindo Commit of 2^22 coefficients by the C restatement, all host threads
    from oracle import jindo_ref as J
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from jindo_util import cdt_array, params_dict
    p = J.new_parameters(q, 1 << COMMIT_LOGN, 1)
    ecd = J.Encoder(p, J.Tape(b"bench"))
    rng = np.random.default_rng(0x4A494E444F)
    key, tape, _ = synthetic_commit_inputs(p, f.L, rng)
    v = rng.integers(0, 1 << 62, size=(1 << COMMIT_LOGN, f.L), dtype=np.uint64)
    tie_first_row(tape, v, q, f.L, p.cols * p.slots)
    _, commit_cpu = commit_cpu_baseline(q, key, tape, v, params_dict(p, ecd), cdt_array(ecd))
    emit(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64x4 (255-bit Montgomery)", "data": "synthetic",
        "config": bench_config(),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{n_polys} of the {BATCH} polys x 2^{LOGN} per step (throughput is per coefficient, so the sample "
                                   f"size does not change it); C restatement of math/bigpoly/ntt.go (Go toolchain unavailable), "
                                   f"batch-parallel over {threads} pthreads -- the reference itself is single-threaded "
                                   f"(1 thread: ~1/{threads} of this figure), so the ratio against it is conservative"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "commit": {"metric": "Jindo Commit throughput (one 2^22-coefficient polynomial, 255-bit field, NewParameters(2^22, 1))",
                   "value": commit_cpu["value"], "unit": "polys/s", "cpu_baseline": commit_cpu},
    }))


COMMIT_LOGN = 22


def synthetic_commit_inputs(p, L, rng):
    """Synthetic commit key and randomness tape of the right shapes (uniform residues / small integers)."""
    import numpy as np

    def rpolys(shape, primes):
        a = np.empty(shape + (len(primes), p.d), dtype=np.uint64)
        for l, pr in enumerate(primes):
            a[..., l, :] = rng.integers(0, pr, size=shape + (p.d,), dtype=np.uint64)
        return a

    def felems(n):
        return rng.integers(0, 1 << 62, size=(n, L), dtype=np.uint64)
    cs = p.cols * p.slots
    n_twin = p.cols * (p.rows - 1); n_cosac = p.cols + p.rows; n_noise = (p.cols + 1) * (p.in_msis + p.mlwe)
    key = (rpolys((p.in_msis, p.rows), p.q_primes), rpolys((p.in_msis, p.mlwe), p.q_primes),
           rpolys((p.out_msis, p.in_com_dcmp_len), p.qout_primes))
    tape = {"first_row": felems(cs), "last_row": felems(cs), "mask_rows": felems(p.rows * p.slots),
            "twin_u": rng.integers(0, 1 << 63, size=n_twin * p.d, dtype=np.int64).view(np.uint64),
            "cosac_g": rng.integers(-5000, 5000, size=n_cosac * p.d, dtype=np.int64),
            "mlwe": rng.integers(-20, 20, size=n_noise * p.d, dtype=np.int64)}
    return key, tape, (n_twin, n_cosac, n_noise)


def tie_first_row(tape, v, q, L, cs):
    """genFirstLastRow (jindo/prover.go:65-86) on the synthetic tape: the last row keeps its random elements (its final one
    is zero) and the first row becomes v[i] - last[i-1], so that the committed matrix really evaluates to v and the proof of the
    timed Commit passes verifyEval.  Montgomery representatives subtract like the values they stand for."""
    import numpy as np
    def ints(a):
        return [sum(int(x) << (64 * k) for k, x in enumerate(row)) for row in a]
    last = ints(tape["last_row"][:cs]); vv = ints(v[:cs])
    last[cs - 1] = 0
    first = [vv[0]] + [(vv[i] - last[i - 1]) % q for i in range(1, cs)]
    def limbs(vals):
        return np.array([[(x >> (64 * k)) & ((1 << 64) - 1) for k in range(L)] for x in vals], dtype=np.uint64)
    tape["first_row"] = limbs(first); tape["last_row"] = limbs(last)


def commit_cpu_baseline(q, key, tape, v, pdict, cdt):
    """The same 2^22-coefficient Commit (same key, tape and vector) by the C restatement of jindo.Prover.Commit
    (oracle/jindo_oracle.c), columns split over the host threads.  One commit is 5-60 s of CPU work."""
    import numpy as np
    import oracle
    from oracle import jindo_c
    threads = max(1, min(os.cpu_count() or 1, 64))
    fo = oracle.Field(q)
    t0 = time.perf_counter()
    res = jindo_c.commit(fo, pdict, key[0], key[1], key[2], np.array(cdt, dtype=np.uint64), v, tape, threads=threads)
    dt = time.perf_counter() - t0
    return res["commitment"], {"value": 1.0 / dt, "unit": "polys/s", "cores": threads, "kind": "port",
            "sample": f"one Commit of 2^{COMMIT_LOGN} coefficients (synthetic key, tape and vector generated like the GPU leg's) in {dt:.1f} s; C restatement of "
                      "jindo.Prover.Commit, pinned to the Python restatement; columns in parallel over pthreads "
                      "(the reference itself is single-threaded)"}


def commit_leg(rs, L, f, q, args, world, rank, barrier, dist):
    """Second half of BASELINE.json's metric: Jindo Commit polys/s on one 2^22-coefficient polynomial
    (BASELINE configs[2], NewParameters(2^22, 1) over the 255-bit field), key/tape synthetic."""
    import numpy as np
    import torch
    JH = rs.jindo_host                          # host-side parameter derivation (product side, no oracle)
    N = 1 << COMMIT_LOGN
    p = JH.new_parameters(q, N, 1)
    pdict, cdt = JH.prover_inputs(p)
    rng = np.random.default_rng(0x4A494E444F)       # same inputs on every rank (the column-sharded leg needs that)
    key, tape, (n_twin, n_cosac, n_noise) = synthetic_commit_inputs(p, f.L, rng)
    # commit key: NewCommitKey(params, "Jindo!") evaluated in HBM (rs_jindo_create_crs); read back once for the CPU leg
    t0 = time.perf_counter()
    gp = rs.jindo.Prover(f, pdict, None, None, None, np.array(cdt, dtype=np.uint64), crs=b"Jindo!", key_from_crs=True)
    key_ms = (time.perf_counter() - t0) * 1e3
    key = gp.CommitKey()
    v_np = rng.integers(0, 1 << 62, size=(N, f.L), dtype=np.uint64)
    tie_first_row(tape, v_np, q, f.L, p.cols * p.slots)
    v_host = torch.from_numpy(v_np.view(np.int64)).pin_memory()
    v = v_host.cuda()
    B = rs.Batch.from_torch(f, v, 1, N)
    dev = {k: torch.from_numpy(np.ascontiguousarray(a).view(np.int64).reshape(-1)).cuda() for k, a in tape.items()}
    pin = {k: torch.from_numpy(np.ascontiguousarray(a).view(np.int64).reshape(-1)).pin_memory() for k, a in tape.items()}

    def tape_struct(src, on_device):
        T = rs.capi.JindoTape()
        T.first_row, T.last_row, T.mask_rows = src["first_row"].data_ptr(), src["last_row"].data_ptr(), src["mask_rows"].data_ptr()
        T.n_mask_rows = p.rows; T.twin_u = src["twin_u"].data_ptr(); T.n_twin_u = n_twin * p.d
        T.cosac_g = src["cosac_g"].data_ptr(); T.n_cosac_polys = n_cosac; T.mlwe = src["mlwe"].data_ptr(); T.on_device = on_device
        return T
    Td, Th = tape_struct(dev, 1), tape_struct(pin, 0)
    op = gp.NewOpening()
    com = np.zeros((p.out_msis, len(p.qout_primes), p.d), dtype=np.uint64)
    steps = max(1, min(args.steps, 5))
    for _ in range(3):
        gp.CommitInto(B, 0, N, Td, op, com)
    barrier()
    l0 = L.rs_launch_count()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        gp.CommitInto(B, 0, N, Td, op, com)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1) / steps
    launches = (L.rs_launch_count() - l0) // steps
    com_ref = com.copy()
    # "open" (Prover.Evaluate's two passes over the Opening, prover.go:281-316) through the C ABI: left vector /
    # challenges from the host, partial sums / responses back to the host
    def rp(n):
        a = np.empty((n, len(p.q_primes), p.d), dtype=np.uint64)
        for l, pr in enumerate(p.q_primes):
            a[:, l, :] = rng.integers(0, pr, size=(n, p.d), dtype=np.uint64)
        return a
    left_h, chal_h = rp(p.rows), rp(p.cols)

    def wall(fn):
        fn(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) * 1e3 / steps
    # ---- N > 1: ONE commit split over the GPUs by columns of the committed matrix (SURVEY 8e iii): encode +
    #      inner commit per shard, InCommit slices exchanged over NCCL in the Opening's device buffer, outer commit
    sharded = None
    if world > 1:
        par = rs.parallel
        comm = par.NcclComm(rank, world)                 # the library's own communicator (rs_nccl_comm_create)
        op2 = gp.NewOpening()
        for _ in range(3):                               # warm-up (NCCL channels)
            com_s, _ = par.sharded_commit(gp, B, 0, N, Td, rank, world, op2, comm=comm)
        barrier()
        t0 = time.perf_counter()
        for _ in range(max(steps, 5)):
            com_s, _ = par.sharded_commit(gp, B, 0, N, Td, rank, world, op2, comm=comm)
        barrier()
        sh_ms = (time.perf_counter() - t0) * 1e3 / max(steps, 5)
        ts = torch.tensor([sh_ms], device="cuda")
        dist.all_reduce(ts, op=dist.ReduceOp.MAX)
        # the gathered InCommit equals the unsharded Opening's
        inc_ok = all(np.array_equal(op2.InCommit(k), op.InCommit(k)) for k in (0, p.in_msis * (p.cols // 2) + 1, p.in_com_dcmp_len - 1))
        sharded = {"ms_per_commit": float(ts[0]), "polys_per_s": 1e3 / float(ts[0]), "scaling": "strong",
                   "commitment_matches_unsharded": bool(np.array_equal(com_s, com_ref)), "incommit_gathered_matches": bool(inc_ok),
                   "exchange": "rs_jindo_outer_commit_sharded: all-reduce of the out_msis partial outer sums (%d KiB), cut-off replicated; "
                               "InCommit slices all-gathered in place afterwards (rs_jindo_allgather_incommit, %d KiB), NCCL on the compute stream"
                               % (p.out_msis * len(p.qout_primes) * p.d * 8 // 1024, p.in_com_dcmp_len * len(p.qout_primes) * p.d * 8 // 1024)}
        del op2
        comm.close()
    partial_ms = wall(lambda: gp.Partial(op, left_h))
    respond_ms = wall(lambda: gp.Respond(op, chal_h))
    # the same two passes with leftVec(x) / encodeChallengeTo evaluated on the device (only x and the SHAKE bytes go in)
    x_mont = rng.integers(0, 1 << 62, size=f.L, dtype=np.uint64)
    chal_bytes = rng.bytes(16 * p.cols)
    partial_at_ms = wall(lambda: gp.PartialAt(op, x_mont))
    # responses land in pinned host memory (what the Go shim stages through, rs_pinned_alloc)
    nq = len(p.q_primes)
    enc_pin = torch.empty((p.rows, nq, p.d), dtype=torch.int64).pin_memory()
    ml_pin = torch.empty((p.mlwe + p.in_msis, nq, p.d), dtype=torch.int64).pin_memory()
    resp_out = (enc_pin.numpy().view(np.uint64), ml_pin.numpy().view(np.uint64))
    respond_bytes_ms = wall(lambda: gp.RespondBytes(op, chal_bytes, out=resp_out))
    # prove -> verify: Verifier.Verify's arithmetic on the device (rs_jindo_verify) for the proof just produced
    # (batch 1: the transcript bytes are the ones the responses were computed with, y = v(x) by the Horner kernel)
    part_v = gp.PartialAt(op, x_mont)
    enc_v, ml_v = gp.RespondBytes(op, chal_bytes)
    pf_v = {"Partial": part_v[:p.cols], "PartialMask": part_v[p.cols], "Encode": enc_v, "MLWE": ml_v,
            "InCommit": np.stack([op.InCommit(k) for k in range(p.in_com_dcmp_len)])}
    y_mont = np.zeros(f.L, dtype=np.uint64)
    rs.check(L.rs_poly_evaluate(B.h, 0, N, x_mont.ctypes.data_as(rs.capi.c_u64p), y_mont.ctypes.data_as(rs.capi.c_u64p)))
    verify_checks = gp.Verify(x_mont, com_ref[None], b"", chal_bytes, y_mont[None], pf_v, p.in_com_dcmp_two_nm, p.res_two_nm)
    verify_ms = wall(lambda: gp.Verify(x_mont, com_ref[None], b"", chal_bytes, y_mont[None], pf_v, p.in_com_dcmp_two_nm, p.res_two_nm))
    # e2e: committed vector + the whole tape start in pinned host memory
    nb = v_host.numel() * 8
    tape_bytes = sum(t.numel() * 8 for t in pin.values())

    def timed_e2e(T):
        def one():                                  # v and the tape start in pinned host memory
            gp.CommitHostInto(v_host.data_ptr(), N, T, op, com)
        one()
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            one()
        barrier()
        return (time.perf_counter() - t0) * 1e3 / steps
    host_tape_ms = timed_e2e(Th)
    deterministic = bool(np.array_equal(com, com_ref))      # same tape through host and device pointers
    # the drop-in default: the encoder's TwinCDT uniforms are drawn in HBM from the prover's own AES-CTR
    # stream (csprng.UniformSampler restated on the device), everything else comes from the host
    gp.SetTwinSeed(bytes(range(32)))
    Ts = tape_struct(pin, 0); Ts.twin_u = None; Ts.n_twin_u = 0
    e2e_ms = timed_e2e(Ts)
    twin_bytes = pin["twin_u"].numel() * 8
    # ---- e2e from "vector + seeds": EVERY tape entry is produced inside the timed region.  The two TwinCDT samplers
    # (encoder rows, mlweSampler) run on the device from their seeds; the COSAC rows, the mask column's rounded-Gaussian
    # noise and the MustSetRandom elements come from the product's host sampler (csprng restated in C++, include/
    # ringo_sampler.h) on worker threads, double buffered: the tape of Commit k+1 is sampled while Commit k runs.
    gp.SetMlweSeed(bytes(range(32, 64)), p.mlwe_std)
    noise = {"ecd_blind_std": p.ecd_blind_std, "mask_blind_std": p.mask_blind_std, "mask_std": p.mask_std, "mask_mlwe_std": p.mask_mlwe_std}
    n_thr = max(1, min((os.cpu_count() or 1) // max(1, world), 32))
    tb = rs.jindo.TapeBuilder(pdict, len(cdt[0]), noise, q, threads=n_thr)
    vp = v_host.data_ptr()

    def seeds_run(n_steps, overlap):
        samp = []
        slot = 0
        if overlap:
            tb.start(vp, N, slot)
        t0 = time.perf_counter()
        for _ in range(n_steps):
            if not overlap:
                tb.start(vp, N, slot)
            T, hms = tb.wait(slot)
            samp.append(hms)
            if overlap:
                tb.start(vp, N, slot ^ 1)                 # next Commit's tape, sampled behind this Commit's kernels
            gp.CommitHostInto(vp, N, T, op, com)
            slot ^= 1
        dt = (time.perf_counter() - t0) * 1e3 / n_steps
        if overlap:
            tb.wait(slot)                                  # drain the tape started last
        return dt, sum(samp) / len(samp)
    seeds_run(2, True)
    barrier()
    seeds_ms, sampling_ms = seeds_run(max(steps, 4), True)
    barrier()
    serial_ms, _ = seeds_run(2, False)
    one_thread = rs.jindo.TapeBuilder(pdict, len(cdt[0]), noise, q, threads=1)
    one_thread.start(vp, N, 0); _, sampling_1t_ms = one_thread.wait(0)
    del one_thread
    # ---- the same with the COSAC rows sampled on the device too (rs_jindo_set_cosac_stream): the host only draws the
    # MustSetRandom elements and the mask column's MLWE noise
    gp.SetCosacSeeds(os.urandom(32), os.urandom(32), noise)
    tb.set_device_cosac(True)
    seeds_run(2, True)
    barrier()
    devc_ms, devc_sampling_ms = seeds_run(max(steps, 4), True)
    barrier()
    cosac_workers, cosac_overflow = gp.CosacStatus()
    tb.set_device_cosac(False)
    t = torch.tensor([ms, e2e_ms, host_tape_ms, seeds_ms, serial_ms, devc_ms], device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_ms, host_tape_ms, seeds_ms, serial_ms, devc_ms = (float(x) for x in t)
    # algorithmic work per commit: mod-mul-adds of the inner/outer products and NTT butterflies
    macs = (p.cols + 1) * p.in_msis * (p.rows + p.mlwe) * p.d * len(p.q_primes) + p.out_msis * p.in_com_dcmp_len * p.d * len(p.qout_primes)
    out = {
        "metric": "Jindo Commit throughput (one 2^22-coefficient polynomial, 255-bit field, NewParameters(2^22, 1))",
        "value": world * 1e3 / ms, "unit": "polys/s", "ms_per_commit": ms, "mcoeff_per_s": world * N / ms / 1e3,
        "params": {"rows": p.rows, "cols": p.cols, "slots": p.slots, "q_bits": p.q_bits, "q_limbs": len(p.q_primes),
                   "qout_limbs": len(p.qout_primes), "in_msis": p.in_msis, "mlwe": p.mlwe, "out_msis": p.out_msis},
        "ring_polys_encoded": (p.cols + 1) * p.rows + (p.cols + 1) * (p.in_msis + p.mlwe),
        "mod_mul_adds": macs, "gmac_per_s": world * macs / ms / 1e6,
        "alg_imad": 16 * (macs + ((p.cols + 1) * (p.rows + p.in_msis + p.mlwe) * len(p.q_primes)
                                  + (p.cols + 1) * p.in_msis * (len(p.q_primes) + len(p.qout_primes))
                                  + 2 * p.out_msis * len(p.qout_primes)) * (p.d // 2) * 8),
        "commit_key": {"ms": key_ms, "note": "NewCommitKey(params, crs) on the device: AES-CTR SampleN stream incl. rejected draws (rs_jindo_create_crs), "
                                             "cold (includes table setup of the prover handle); no key bytes cross PCIe"},
        "open": {"partial_ms": partial_ms, "respond_ms": respond_ms, "partial_at_ms": partial_at_ms, "respond_bytes_ms": respond_bytes_ms,
                 "commit_plus_open_polys_per_s": world * 1e3 / (ms + partial_at_ms + respond_bytes_ms),
                 "verify_ms": verify_ms, "verify_checks": int(verify_checks),
                 "verify_note": "rs_jindo_verify on the proof of this Commit + open (host proof arrays in, 4 check bits out; 15 = accepted; "
                                "the synthetic tape's first row is tied to v by genFirstLastRow's rule, so all four checks are meaningful)",
                 "note": "Evaluate's two passes over the 1.08 GB Opening: partial/respond = host-encoded vectors in, pageable results out; "
                         "partial_at/respond_bytes = x and SHAKE bytes in (vectors encoded on the device), results into pinned memory"},
        "sharded": sharded,
        "gpu_launches_per_commit": int(launches), "deterministic": deterministic,
        "e2e": {"value": world * 1e3 / devc_ms, "unit": "polys/s", "ms_per_commit": devc_ms,
                "h2d_bytes_per_step": nb + tape_bytes - twin_bytes - (n_noise - (p.in_msis + p.mlwe)) * p.d * 8 - n_cosac * p.d * 8,
                "d2h_bytes_per_step": int(com.nbytes),
                "inputs": "the committed vector in pinned host memory + sampler seeds; nothing pre-sampled",
                "sampling": {"host_threads": n_thr, "host_ms_per_commit": devc_sampling_ms,
                             "device": "every Gaussian sampler: encoder TwinCDT rows (%d polys), mlweSampler (%d polys), COSAC rows (%d polys, one fresh "
                                       "COSACSampler per coefficient, cosac_dev.cu)" % (n_twin, p.cols * (p.in_msis + p.mlwe), n_cosac),
                             "host": "MustSetRandom elements of the first / last / mask rows and the mask column's rounded-Gaussian MLWE noise",
                             "cosac_workers": int(cosac_workers), "cosac_overflow": int(cosac_overflow)},
                "call": "rs_tape_builder_start/wait (device COSAC) + rs_jindo_commit_host: v uploaded in 8 pieces behind the encode kernels"},
        "e2e_host_cosac": {"value": world * 1e3 / seeds_ms, "unit": "polys/s", "ms_per_commit": seeds_ms,
                "h2d_bytes_per_step": nb + tape_bytes - twin_bytes - (n_noise - (p.in_msis + p.mlwe)) * p.d * 8, "d2h_bytes_per_step": int(com.nbytes),
                "inputs": "the committed vector in pinned host memory + sampler seeds; nothing pre-sampled; COSAC rows on the host worker threads",
                "sampling": {"host_threads": n_thr, "host_ms_per_commit": sampling_ms, "host_ms_per_commit_1_thread": sampling_1t_ms,
                             "overlapped": "tape k+1 is sampled on the worker threads while Commit k runs on the GPU",
                             "no_overlap_ms_per_commit": serial_ms,
                             "host": "COSAC rows (%d polys), mask-column rounded-Gaussian MLWE noise, MustSetRandom elements: csprng restated in C++ "
                                     "(rs_tape_builder, every worker owns independently seeded samplers like Prover.SafeCopy)" % n_cosac,
                             "device": "encoder TwinCDT rows (%d polys) and mlweSampler (%d polys) from AES-256-CTR streams in HBM" % (n_twin, p.cols * (p.in_msis + p.mlwe))},
                "call": "rs_tape_builder_start/wait + rs_jindo_commit_host: v uploaded in 8 pieces behind the encode kernels"},
        "e2e_presampled_tape": {"value": world * 1e3 / e2e_ms, "unit": "polys/s", "ms_per_commit": e2e_ms,
                "note": "host entries of the tape already sampled and resident in pinned memory (what round 1 reported as e2e); TwinCDT uniforms on the device"},
        "e2e_host_tape": {"value": world * 1e3 / host_tape_ms, "unit": "polys/s", "ms_per_commit": host_tape_ms,
                          "h2d_bytes_per_step": nb + tape_bytes, "d2h_bytes_per_step": int(com.nbytes),
                          "sampler": "one u64 per TwinCDT sample uploaded from pinned host memory"},
        "data": "commit key = NewCommitKey(params, crs \"Jindo!\") generated on the device; synthetic randomness tape (uniform residues / small integers), resident in HBM for `value`",
    }
    if rank == 0 and world == 1:
        cpu_com, out["cpu_baseline"] = commit_cpu_baseline(q, key, tape, v_host.numpy().view(np.uint64).reshape(N, f.L), pdict, cdt)
        # full-size parity: the GPU commitment of the timed leg against the C restatement on the same inputs
        out["commitment_matches_cpu_oracle"] = bool(np.array_equal(com_ref, cpu_com))
    return out


def buckler_leg(rs, L, args):
    """BASELINE configs[0]: buckler Prove (and Verify) of examples/mult's MultCircuit at its default rank N = 2^13 over the
    128-bit field, through the compiled prover (rs_buckler_prove, include/ringo_buckler.h) with every random value drawn inside
    the timed region (host samplers on worker threads + the device TwinCDT / MLWE streams), as the reference's Prove does.
    Product side only: circuit description, Jindo parameters and the witness come from the package."""
    import numpy as np
    N = 1 << 13
    q = rs.modulus("zp128")
    f = rs.Field(q)
    JH = rs.jindo_host
    desc, commit_rank, batch = rs.circuits.mult_circuit(q, N)
    p = JH.new_parameters(q, commit_rank, batch)
    pdict, cdt = JH.prover_inputs(p)
    gp = rs.jindo.Prover(f, pdict, None, None, None, np.array(cdt, dtype=np.uint64), crs=b"Buckler!", key_from_crs=True)
    gp.SetTwinSeed(os.urandom(32)); gp.SetMlweSeed(os.urandom(32), p.mlwe_std)
    threads = max(1, min((os.cpu_count() or 2) - 1, 15))
    pr = rs.buckler.NativeProver(f, q, desc, p, gp)
    pr.enable_sampling({"ecd_blind_std": p.ecd_blind_std, "mask_blind_std": p.mask_blind_std, "mask_std": p.mask_std,
                        "mask_mlwe_std": p.mask_mlwe_std}, threads=threads)
    insts = [rs.circuits.mult_witness(f, q, N, 77 + k) for k in range(4)]
    steps = max(4, min(args.steps, 10))

    def timed_prove():
        pr.prove_raw(insts[0][0], insts[0][1], None)
        l0 = L.rs_launch_count()
        t0 = time.perf_counter()
        for k in range(steps):
            raw = pr.prove_raw(insts[k % 4][0], insts[k % 4][1], None, phases=False, reuse=True)   # proof buffers owned by the caller, as in Go
        return (time.perf_counter() - t0) * 1e3 / steps, (L.rs_launch_count() - l0) // steps, raw
    host_cosac_ms, _, _ = timed_prove()                        # COSAC rows on the host worker threads
    gp.SetCosacSeeds(os.urandom(32), os.urandom(32), {"ecd_blind_std": p.ecd_blind_std, "mask_blind_std": p.mask_blind_std, "mask_std": p.mask_std})
    pr.set_device_cosac(True)                                  # ... and on the device: the default of this leg
    prove_ms, launches, raw = timed_prove()
    res = pr.verify_raw(insts[(steps - 1) % 4][0], raw)
    t0 = time.perf_counter()
    for _ in range(3):
        pr.verify_raw(insts[(steps - 1) % 4][0], raw)
    verify_ms = (time.perf_counter() - t0) * 1e3 / 3
    wrong = pr.verify_raw(insts[steps % 4][0], raw)            # the proof against another instance's public witness
    pr.prove_raw(insts[0][0], insts[0][1], None, phases=True)
    return {"workload": f"buckler.Prove + Verify, examples/mult MultCircuit, N = 2^13, zp128, {desc['w_cnt']}+{desc['pw_cnt']} witnesses, "
                        f"{batch} Jindo commits (rows {p.rows}, cols {p.cols}, slots {p.slots}) (BASELINE configs[0])",
            "prove_ms": prove_ms, "proofs_per_s": 1e3 / prove_ms, "prove_ms_host_cosac": host_cosac_ms,
            "cosac_overflow": int(gp.CosacStatus()[1]), "verify_ms": verify_ms, "verifier_accepts": res == 0,
            "verifier_rejects_other_statement": wrong != 0, "gpu_launches_per_proof": int(launches),
            "host_sampler_threads": threads,
            "phases_ms_with_syncs": {k: round(v, 3) for k, v in pr.timing.items() if k != "_"},
            "randomness": "drawn inside the timed region (rs_buckler_prove with rand = NULL): every Gaussian sampler on the device, "
                          "MustSetRandom elements and the mask column's noise on the host worker threads",
            "parity": "the same prover with explicit randomness is bit-identical to the restated reference prover "
                      "(tests/test_gpu_buckler.py; tests/perf/bench_buckler.py at this size: 6.0 ms)"}


def bind_to_gpu_numa_node(local):
    """Run this rank's host threads -- and therefore place its pinned staging buffers (first touch) -- on the CPUs of
    the NUMA node the GPU's PCIe root port hangs off.  On an 8-GPU box the e2e leg moves 2 GiB per step per rank
    through host memory; with unpinned ranks half of that crosses the socket interconnect.  Returns what was done."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(local)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        base = "/sys/bus/pci/devices/" + bdf
        node = int(open(base + "/numa_node").read().strip())
        cpus = set()
        for part in open(base + "/local_cpulist").read().strip().split(","):
            if part:
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0)
        use = cpus & allowed
        if node < 0 or not use or use == allowed:
            return {"node": node, "bound": False, "cpus": len(allowed)}
        os.sched_setaffinity(0, use)
        return {"node": node, "bound": True, "cpus": len(use)}
    except Exception as e:                                 # no sysfs / not permitted: run unbound
        return {"bound": False, "why": type(e).__name__}


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product has no CPU fallback; use --impl reference)")
    torch.cuda.set_device(local)
    numa = bind_to_gpu_numa_node(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    rs = product()
    L = rs.load()
    rs.check(L.rs_init(local))
    stream = torch.cuda.current_stream()
    rs.check(L.rs_set_stream(ctypes.c_void_p(stream.cuda_stream)))

    q = product().modulus(FIELD)
    f = rs.Field(q)
    N = 1 << LOGN
    ntt = rs.CyclotomicTransformer(f, N)
    words = BATCH * N * f.L
    nbytes = words * 8

    # ---- device-resident leg: inputs live in HBM before the timed region (1 GiB >> 126 MB L2) ----
    host = synthetic_polys(q, f.L, BATCH, N, 0x4E5454 + rank)
    src_t = torch.from_numpy(host.view(np.int64).reshape(-1)).cuda()
    dst_t = torch.empty_like(src_t)
    src = rs.Batch.from_torch(f, src_t, BATCH, N)
    dst = rs.Batch.from_torch(f, dst_t, BATCH, N)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # spin-up (setup, not part of the W warm-up steps): ~0.2 s of the same work, so that table uploads, lazy module loading
    # and the clock ramp of an idle GPU are over before the contract's warm-up / timed steps begin
    for _ in range(40):
        ntt.FwdNTTTo(dst, src)
    barrier()
    for _ in range(args.warmup):
        ntt.FwdNTTTo(dst, src)
    barrier()
    sampler = ClockSampler(local); sampler.start()
    launches0 = L.rs_launch_count()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        ntt.FwdNTTTo(dst, src)
    e1.record()
    barrier()
    launches = L.rs_launch_count() - launches0
    ms = e0.elapsed_time(e1) / args.steps
    sampler.stop_flag = True
    tmax = torch.tensor([ms], device="cuda")
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms_max = float(tmax.item())
    value = world * BATCH * N / (ms_max * 1e-3) / 1e6

    # ---- per-kernel times (CUDA events on the launching stream, inside the library) ----
    pass_ms = (ctypes.c_float * 16)(); n_pass = ctypes.c_int(16)
    rs.check(L.rs_ntt_profile(ntt.h, dst.h, BATCH, 0, max(3, args.steps), pass_ms, ctypes.byref(n_pass)))
    passes = [float(pass_ms[i]) for i in range(n_pass.value)]
    dom = max(range(len(passes)), key=lambda i: passes[i])
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    alg_bytes = BATCH * N * 16 * f.L            # every pass reads and writes each coefficient once
    hbm_ach = alg_bytes / (passes[dom] * 1e-3) / 1e9
    g = ctypes.c_double()
    # a 32x32 -> 64-bit multiply-add is one IMAD.WIDE.U32 with a 64-bit addend: measured with operands that change
    # every iteration (mode 5).  Mode 1 -- used as the denominator until the end of round 1 -- has loop-invariant
    # multiplicands, ptxas hoists the product and the loop measures IADD3 pairs (16.7 T/s); the real rate is 9.3 T/s,
    # the same with or without carry-in (mode 2): 32 lanes per clock per SM.
    rs.check(L.rs_bench_imad_peak(5, ctypes.byref(g)))
    imad_peak = g.value
    nl = 2 * f.L
    imad_alg_total = BATCH * N * (LOGN / 2) * (2 * nl * nl + nl)       # (log2N/2)(2l^2+l) per coeff
    imad_ach = imad_alg_total / (sum(passes) * 1e-3) / 1e9

    # ---- the other transforms of the sweep (BASELINE configs[1]), same batch, device resident ----
    def timed(fn):
        fn(); torch.cuda.synchronize()
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3):
            fn()
        b.record(); torch.cuda.synchronize()
        return BATCH * N / (a.elapsed_time(b) / 3 * 1e-3) / 1e6
    cyc = rs.CyclicTransformer(f, N)             # buckler.Encoder / polyEval transform (encoder.go:17, compile.go:186)
    variants = {
        "negacyclic_inverse": timed(lambda: ntt.InvNTTTo(dst, src)),
        "cyclic_forward": timed(lambda: cyc.FwdNTTTo(dst, src)),
        "cyclic_inverse": timed(lambda: cyc.InvNTTTo(dst, src)),
        "unit": "Mcoeff/s per GPU",
        "note": "cyclic tables have tw[2^s] = 1: the opening radix-8 round skips 7 of its 12 products",
    }

    # ---- e2e leg: host buffers through the C ABI, copies inside the timed region ----
    # The batch is cut into chunks that ride three streams (upload / transform / download of different
    # chunks overlap; PCIe is full duplex), exactly what a host caller streaming witnesses would do.
    hin = torch.from_numpy(host.view(np.int64).reshape(-1)).pin_memory()
    hout = torch.empty_like(hin).pin_memory()
    e2e_steps = max(1, min(args.steps, 5))
    n_streams = 4                      # tools/e2e_probe.py: 8x3 24.5 ms, 16x4 23.5 ms, 64x4 25.1 ms; duplex copies alone 21.6 ms
    # chunk sizes in polynomials: 32 in the middle (16 chunks of 32 was the best uniform cut), small chunks at both ends so that the
    # download engine idles only while the first 4 polynomials upload and the upload engine only while the last 4 download
    # (a uniform cut leaves 1/16 of each direction un-overlapped)
    if BATCH >= 128 and BATCH % 32 == 0:
        ramp = [4, 8, 20]
        sizes = ramp + [32] * ((BATCH - 2 * sum(ramp)) // 32) + ramp[::-1]
    else:
        sizes = [max(1, BATCH // 16)] * 16 if BATCH >= 16 else [BATCH]
        sizes[-1] += BATCH - sum(sizes)
    assert sum(sizes) == BATCH and min(sizes) >= 1
    starts = [sum(sizes[:i]) for i in range(len(sizes))]
    n_chunks = len(sizes)
    pw_poly = N * f.L                      # words per polynomial
    cw = sizes[0] * pw_poly                # words of the first chunk (result check below)
    streams = [torch.cuda.Stream() for _ in range(n_streams)]

    def e2e_step():
        for c in range(n_chunks):
            st = streams[c % n_streams]
            p0, cp = starts[c], sizes[c]
            rs.check(L.rs_set_stream(ctypes.c_void_p(st.cuda_stream)))
            rs.check(L.rs_batch_upload(src.h, p0, cp, ctypes.c_void_p(hin.data_ptr() + p0 * pw_poly * 8)))
            rs.check(L.rs_ntt_forward(ntt.h, dst.h, p0, src.h, p0, cp))
            rs.check(L.rs_batch_download_async(dst.h, p0, cp, ctypes.c_void_p(hout.data_ptr() + p0 * pw_poly * 8)))

    def e2e_drain():
        for st in streams:
            st.synchronize()
        rs.check(L.rs_set_stream(ctypes.c_void_p(stream.cuda_stream)))

    # The steps of the timed region are enqueued back to back, as a caller streaming batch after batch would: chunk c always rides
    # stream c % n_streams, so step k + 1's upload of a chunk is ordered behind step k's transform of it and its transform behind
    # step k's download (same buffers every step); every step moves its own 1 GiB in and 1 GiB out; the host waits once, at the end.
    e2e_step(); e2e_drain()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    e2e_drain()
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
    t2 = torch.tensor([e2e_ms], device="cuda")
    if world > 1:
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
    e2e_val = world * BATCH * N / (float(t2.item()) * 1e-3) / 1e6
    # (every chunk came back: compare the whole batch with the device-resident result of the same transform)
    e2e_ok = bool(np.array_equal(hout.numpy(), dst_t.cpu().numpy()))
    # the leg's own roof: the same bytes as pinned copies in both directions at once, no kernel (PCIe full duplex)
    def duplex():
        with torch.cuda.stream(streams[0]):
            src_t.copy_(hin, non_blocking=True)
        with torch.cuda.stream(streams[1]):
            hout.copy_(dst_t, non_blocking=True)
        streams[0].synchronize(); streams[1].synchronize()
    duplex()
    t0 = time.perf_counter()
    for _ in range(2):
        duplex()
    duplex_ms = (time.perf_counter() - t0) * 1e3 / 2

    commit = commit_leg(rs, L, f, q, args, world, rank, barrier, dist if world > 1 else None)
    buckler = buckler_leg(rs, L, args) if rank == 0 else None
    barrier()

    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_max, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u64x4 (255-bit Montgomery-form elements, IMAD.WIDE on 32-bit limbs, Shoup twiddle products)", "data": "synthetic",
        "config": bench_config(), "host_numa": numa,
        "roofline": {"bound": "hbm", "achieved": hbm_ach, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_ach / hbm_peak,
                     # dram__bytes_read.sum + dram__bytes_write.sum of the dominant (contiguous) pass in the ncu --set full capture
                     # profiles/ntt_sh_r2_ncu_batch512_summary.csv (tools/ntt_shoup_probe.py zp255 16 512, Shoup kernels): 1.080 + 1.039 GB
                     "traffic": 2.119e9 if (BATCH, LOGN, FIELD) == (512, 16, "zp255") else None,
                     "algorithmic_bytes": alg_bytes, "kernel": f"ntt_pass_kernel pass {dom} of {len(passes)}",
                     "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s",
                     "note": "kernel is integer-pipe bound, see roofline_imad"},
        "roofline_imad": {"bound": "imad", "achieved": imad_ach, "peak": imad_peak, "unit": "G IMAD.WIDE/s",
                          "frac": imad_ach / imad_peak, "pass_ms": passes,
                          "algorithmic_imad_per_coeff": (LOGN / 2) * (2 * nl * nl + nl),
                          "peak_source": "rs_bench_imad_peak mode 5 (IMAD.WIDE.U32 with 64-bit addend, changing operands, this run)",
                          "nominal_32bit_imad_peak": 148 * 64 * 1.965, "frac_of_nominal_pairs": imad_ach / (148 * 64 * 1.965 / 2),
                          "note": "one 32x32->64 multiply-add = one IMAD.WIDE (measured 32 lanes/clk/SM) or an IMAD.LO + IMAD.HI pair "
                                  "(nominal 64 lanes/clk/SM each): the same roof either way",
                          # `achieved` counts SURVEY 8(d)'s ALGORITHMIC multiply-adds (2 l^2 + l = 136 per butterfly).  The forward
                          # kernels of the 255-bit field execute fewer: Shoup products by the (w, w') twiddle pairs need 86 wide + 16
                          # half (low-word) multiply-adds = 94 wide equivalents per butterfly (profiles/ntt_sh_r2_opcode_mix.txt),
                          # so the share of the pipe's peak actually spent multiplying is frac * 94 / 136
                          "executed_wide_equiv_per_butterfly": 94 if FIELD == "zp255" else None,
                          "frac_executed": (imad_ach / imad_peak) * 94 / 136 if FIELD == "zp255" else None},
        "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": nbytes,
                "ms_per_step": float(t2.item()), "pipeline": f"{n_chunks} chunks ({'/'.join(str(x) for x in sizes[:4])}/../{sizes[-1]} polys) on {n_streams} streams (pinned host memory), {e2e_steps} steps enqueued back to back, one host wait at the end",
                "result_checked": e2e_ok,
                "pcie_roof": {"duplex_copy_ms": duplex_ms, "gbs_each_way": nbytes / duplex_ms / 1e6,
                              "frac": duplex_ms / e2e_ms,
                              "note": "the step's 2 x 1 GiB as plain pinned copies in both directions at once on this rank, no kernel: "
                                      "the host-buffer leg is PCIe bound, a field element is 255 bits in 32 bytes (nothing to compress)"}},
        "ntt_variants": variants,
        "commit": commit,
        "buckler": buckler,
        "gpu_launches": int(launches),
        "clocks": sampler.summary(),
    }
    if commit and "alg_imad" in commit:
        # Commit against the same integer roof: algorithmic 32-bit multiply-adds (16 per 64-bit mod-mul-add or
        # NTT butterfly) over the measured IMAD.WIDE peak, per GPU
        ach = commit["alg_imad"] / (commit["ms_per_commit"] * 1e-3) / 1e9
        commit["roofline_imad"] = {"bound": "imad", "achieved": ach, "peak": imad_peak, "unit": "G IMAD.WIDE/s", "frac": ach / imad_peak,
                                   "algorithmic_imad_per_commit": commit.pop("alg_imad"),
                                   # what the kernels execute is far below SURVEY 8(d)'s count: the RNS residues have 38 bits, a lazy
                                   # multiply-accumulate is 3 IMAD.WIDE (+ one reduction per output), so `frac` says nothing about pipe
                                   # occupancy.  Executed: the inner mat-vec's 3 IMAD.WIDE per mod-mul-add over its own launch time
                                   # (0.88 ms) is 0.5 of the peak; over the whole Commit:
                                   "executed_imad_wide_per_mac": 3,        # Karatsuba split of the 38-bit residues: 3 IMAD.WIDE
                                   "frac_executed_matvec_over_commit": 3 * commit["mod_mul_adds"] / (commit["ms_per_commit"] * 1e-3) / 1e9 / imad_peak,
                                   "note": "frac uses SURVEY 8(d)'s algorithmic 16 multiply-adds per 64-bit mod-mul-add; the kernels execute 3 "
                                           "(38-bit residues, lazy accumulation): see frac_executed_matvec_over_commit and DESIGN.md section 6"}
    if rank == 0:
        if world == 1:
            threads = os.cpu_count() or 1
            v, n_polys, dt = cpu_reference_rate(threads)
            out["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                                   "sample": f"{n_polys} polys x 2^{LOGN} ({dt * 1e3:.0f} ms per sample); C restatement of "
                                             f"math/bigpoly/ntt.go, batch-parallel over {threads} pthreads"}
        emit(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    # stdout carries exactly one JSON line: anything libraries print there meanwhile (e.g. NCCL's version
    # banner) is sent to stderr, and the real stdout is put back just before the line is printed
    sys.stdout.flush()
    global _REAL_STDOUT
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


_REAL_STDOUT = None


def emit(line):
    sys.stdout.flush()
    if _REAL_STDOUT is not None:
        os.dup2(_REAL_STDOUT, 1)
    print(line, flush=True)


if __name__ == "__main__":
    main()
