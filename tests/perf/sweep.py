"""[tests/perf: measurement harness that checks against / times the oracle, hence under tests/ -- the product never imports oracle/]
BASELINE configs[1] / [2] as tables: the bigpoly NTT sweep (rank 2^12..2^17,
128-bit and 255-bit jindo-modulus fields, batch 1..4096, cyclic and negacyclic, forward and inverse) and the Jindo
Commit / open size sweep (N = 2^13..2^22 batch 1, the shapes of the reference's BenchmarkSingle, jindo_test.go:54-94),
each GPU figure next to the restated-reference CPU figure (C oracle, ONE thread: the reference is single-threaded).

    python tests/perf/sweep.py [ntt] [commit] [--out FILE]     ->  one JSON document

GPU times are CUDA events inside the library (rs_ntt_profile) or around the C-ABI call; inputs are resident in HBM.
CPU samples are bounded (one transform / one commit per point, skipped above a size cap).
"""
import ctypes, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
from oracle import pyref
import oracle

args = [a for a in sys.argv[1:] if not a.startswith("--")]
out_path = next((a[6:] for a in sys.argv[1:] if a.startswith("--out=")), None)
do_ntt = not args or "ntt" in args
do_commit = not args or "commit" in args
do_k3 = not args or "k3" in args

import torch
rs = bench.product(); L = rs.load()
rs.check(L.rs_init(0))
rs.check(L.rs_set_stream(ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
doc = {"gpu": torch.cuda.get_device_name(0), "host_cores": os.cpu_count(),
       "cpu_label": "restated-reference CPU baseline (Go toolchain unavailable): C oracle, 1 thread"}


def ntt_sweep():
    rows = []
    for name in ("zp128", "zp255"):
        q = pyref.modulus(name); f = rs.Field(q); fo = oracle.Field(q)
        for logN in (12, 13, 14, 15, 16, 17):
            N = 1 << logN
            cpu = {}
            for neg in (False, True):                     # one single-thread transform per (rank, kind, direction)
                t = oracle.NTT(fo, logN, neg)
                d = bench.synthetic_polys(q, f.L, 1, N, 5)
                t.fwd_batch_inplace(d, 1)
                t0 = time.perf_counter(); t.fwd_batch_inplace(d, 1); dt_f = time.perf_counter() - t0
                t0 = time.perf_counter(); t.inv_batch_inplace(d, 1); dt_i = time.perf_counter() - t0
                cpu[neg] = (N / dt_f / 1e6, N / dt_i / 1e6)
            for batch in (1, 8, 64, 512, 4096):
                if batch * N * f.L * 8 > (4 << 30):        # 4 GiB per buffer is plenty to be far beyond L2
                    continue
                host = bench.synthetic_polys(q, f.L, min(batch, 64), N, 0x4E5454)
                buf = torch.from_numpy(np.ascontiguousarray(np.resize(host, (batch, N, f.L))).view(np.int64).reshape(-1)).cuda()
                B = rs.Batch.from_torch(f, buf, batch, N)
                for neg in (False, True):
                    tr = (rs.CyclotomicTransformer if neg else rs.CyclicTransformer)(f, N)
                    for inv in (0, 1):
                        ms = (ctypes.c_float * 16)(); n = ctypes.c_int(16)
                        rs.check(L.rs_ntt_profile(tr.h, B.h, batch, inv, 2, ms, ctypes.byref(n)))
                        n = ctypes.c_int(16)
                        rs.check(L.rs_ntt_profile(tr.h, B.h, batch, inv, 5 if batch * N >= (1 << 22) else 20, ms, ctypes.byref(n)))
                        tot = sum(ms[i] for i in range(n.value))
                        nl = 2 * f.L
                        rows.append({"field": name, "logN": logN, "batch": batch, "kind": "negacyclic" if neg else "cyclic",
                                     "dir": "inverse" if inv else "forward", "passes": n.value, "ms": round(tot, 4),
                                     "gpu_mcoeff_s": round(batch * N / tot / 1e3, 1),
                                     "alg_gbs": round(batch * N * f.L * 16 / tot / 1e6, 1),
                                     "alg_gimad_s": round(batch * N * (logN / 2) * (2 * nl * nl + nl) / tot / 1e6, 1),
                                     "cpu_mcoeff_s_1thread": round(cpu[neg][inv], 3)})
                del B, buf
                torch.cuda.empty_cache()
    return rows


def commit_sweep():
    from oracle import jindo_c
    JH = rs.jindo_host
    q = pyref.modulus("zp255"); f = rs.Field(q); fo = oracle.Field(q)
    rows = []
    for logN, batch in ((13, 1), (15, 1), (17, 1), (19, 1), (19, 8), (22, 1)):
        N = 1 << logN
        p = JH.new_parameters(q, N, batch)
        pdict, cdt = JH.prover_inputs(p)
        rng = np.random.default_rng(0x4A494E444F)
        key, tape, (n_twin, n_cosac, n_noise) = bench.synthetic_commit_inputs(p, f.L, rng)
        gp = rs.jindo.Prover(f, pdict, key[0], key[1], key[2], np.array(cdt, dtype=np.uint64))
        v_np = rng.integers(0, 1 << 62, size=(N, f.L), dtype=np.uint64)
        v = torch.from_numpy(v_np.view(np.int64)).cuda()
        B = rs.Batch.from_torch(f, v, 1, N)
        dev = {k: torch.from_numpy(np.ascontiguousarray(a).view(np.int64).reshape(-1)).cuda() for k, a in tape.items()}
        T = rs.capi.JindoTape()
        T.first_row, T.last_row, T.mask_rows = dev["first_row"].data_ptr(), dev["last_row"].data_ptr(), dev["mask_rows"].data_ptr()
        T.n_mask_rows = p.rows; T.twin_u = dev["twin_u"].data_ptr(); T.n_twin_u = n_twin * p.d
        T.cosac_g = dev["cosac_g"].data_ptr(); T.n_cosac_polys = n_cosac; T.mlwe = dev["mlwe"].data_ptr(); T.on_device = 1
        op = gp.NewOpening()
        com = np.zeros((p.out_msis, len(p.qout_primes), p.d), dtype=np.uint64)
        for _ in range(3):
            gp.CommitInto(B, 0, N, T, op, com)
        torch.cuda.synchronize()
        reps = 5 if logN >= 19 else 20
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            gp.CommitInto(B, 0, N, T, op, com)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps

        def rp(n):
            a = np.empty((n, len(p.q_primes), p.d), dtype=np.uint64)
            for l, pr in enumerate(p.q_primes):
                a[:, l, :] = rng.integers(0, pr, size=(n, p.d), dtype=np.uint64)
            return a
        left_h, chal_h = rp(p.rows), rp(p.cols)

        def wall(fn):
            fn(); torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(reps):
                fn()
            torch.cuda.synchronize()
            return (time.perf_counter() - t0) * 1e3 / reps
        partial_ms = wall(lambda: gp.Partial(op, left_h)); respond_ms = wall(lambda: gp.Respond(op, chal_h))
        row = {"logN": logN, "batch_param": batch, "rows": p.rows, "cols": p.cols, "q_limbs": len(p.q_primes), "qout_limbs": len(p.qout_primes),
               "gpu_commit_ms": round(ms, 4), "gpu_polys_s": round(1e3 / ms, 1), "gpu_mcoeff_s": round(N / ms / 1e3, 1),
               "gpu_partial_ms": round(partial_ms, 4), "gpu_respond_ms": round(respond_ms, 4)}
        if logN <= 19:                                     # one single-thread C-oracle commit on the same inputs
            t0 = time.perf_counter()
            res = jindo_c.commit(fo, pdict, key[0], key[1], key[2], np.array(cdt, dtype=np.uint64), v_np, tape, threads=1)
            dt = time.perf_counter() - t0
            row["cpu_commit_s_1thread"] = round(dt, 3)
            row["commitment_matches_cpu_oracle"] = bool(np.array_equal(res["commitment"], com))
            row["speedup_vs_1thread"] = round(dt * 1e3 / ms, 1)
        rows.append(row)
        del gp, op, B, v, dev
        torch.cuda.empty_cache()
    return rows


def k3_sweep():
    """SURVEY 8d "pointwise ops": single element-wise ops (HBM bound), the fused evalCircuit of the two example circuits
    (integer bound), QuoRemByVanishing (HBM) -- both fractions for each, at the circuits' own embedding ranks."""
    from oracle import buckler_ref as B
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_oracle_buckler import ct_params, BFV_T
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    g = ctypes.c_double(); rs.check(L.rs_bench_imad_peak(5, ctypes.byref(g))); imad = g.value
    rows = []

    def timed(fn, reps=20):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps
    for name, circuit, N in (("zp128", "mult", 1 << 13), ("zp240", "ct", 1 << 14)):
        q = pyref.modulus(name); f = rs.Field(q); fo = oracle.Field(q)
        if circuit == "mult":
            define, n_pw, n_w = B.mult_circuit(N)
        else:
            delta, eb = ct_params(q, N)
            define, n_pw, n_w = B.ciphertext_circuit(N, delta, BFV_T, eb)
        ctx = B.Context(q, N, n_pw, n_w); define(ctx, fo)
        E = 1 << (ctx.arith_max_rank - 1).bit_length()
        terms, coeffs = ctx.terms(ctx.arith_constraints)
        nl = 2 * f.L; mul_imad = 2 * nl * nl + nl
        n_mul = sum((1 if pw >= 0 else 0) + len(ws) for (_, _, pw, ws) in terms) + len(ctx.arith_constraints)   # MulTo + ScalarMulTo
        inputs = len({("w", x) for (_, _, _, ws) in terms for x in ws} | {("pw", pw) for (_, _, pw, _) in terms if pw >= 0})
        rng = np.random.default_rng(3)
        wb = rs.Batch(f, ctx.w_cnt, E).upload(bench.synthetic_polys(q, f.L, ctx.w_cnt, E, 1))
        pwb = rs.Batch(f, ctx.pw_cnt, E).upload(bench.synthetic_polys(q, f.L, ctx.pw_cnt, E, 2))
        out = rs.Batch(f, 4, E)
        T = (rs.capi.Term * len(terms))()
        for k, (ci, co, pw, ws) in enumerate(terms):
            T[k].constraint, T[k].coeff, T[k].pw, T[k].n_w = ci, co, pw, len(ws)
            for j, x in enumerate(ws):
                T[k].w[j] = x
        cm = fo.to_mont(coeffs); bc = fo.to_mont([12345])[0]
        u64p = rs.capi.c_u64p
        ms = timed(lambda: rs.check(L.rs_eval_circuit(out.h, 0, pwb.h, wb.h, T, len(terms), cm.ctypes.data_as(u64p), len(coeffs), bc.ctypes.data_as(u64p))))
        byts = (inputs + 1) * 8 * f.L * E
        rows.append({"op": f"evalCircuit ({circuit}, {len(ctx.arith_constraints)} constraints, {len(terms)} terms)", "field": name, "rank": E, "ms": round(ms, 4),
                     "alg_gbs": round(byts / ms / 1e6, 1), "hbm_frac": round(byts / ms / 1e6 / hbm, 3),
                     "alg_gimad_s": round(n_mul * mul_imad * E / ms / 1e6, 1), "imad_frac": round(n_mul * mul_imad * E / ms / 1e6 / imad, 3),
                     "bound": "integer", "unfused_passes": 2 * n_mul + len(terms)})
        for opname, fn, reads, muls in (("MulTo", lambda: rs.check(L.rs_vec_mul(out.h, 1, wb.h, 0, wb.h, 1, 1)), 2, 1),
                                        ("MulAddTo", lambda: rs.check(L.rs_vec_mul_add(out.h, 1, wb.h, 0, wb.h, 1, 1)), 3, 1),
                                        ("AddTo", lambda: rs.check(L.rs_vec_add(out.h, 1, wb.h, 0, wb.h, 1, 1)), 2, 0),
                                        ("QuoRemByVanishing", lambda: rs.check(L.rs_quorem_vanishing(out.h, 2, out.h, 3, wb.h, 0, N)), 1, 0)):
            ms = timed(fn, 50)
            writes = 2 if opname.startswith("QuoRem") else 1
            byts = (reads + writes) * 8 * f.L * E
            rows.append({"op": opname, "field": name, "rank": E, "ms": round(ms, 4), "alg_gbs": round(byts / ms / 1e6, 1),
                         "hbm_frac": round(byts / ms / 1e6 / hbm, 3), "alg_gimad_s": round(muls * mul_imad * E / ms / 1e6, 1),
                         "imad_frac": round(muls * mul_imad * E / ms / 1e6 / imad, 3), "bound": "hbm (one polynomial = %.1f MB: launch latency at this size)" % (8 * f.L * E / 1e6)})
        # the same single ops on a batch large enough to leave L2 and hide the launch (64 polynomials)
        big = rs.Batch(f, 192, E).upload(np.resize(bench.synthetic_polys(q, f.L, 8, E, 5), (192, E, f.L)))
        for opname, fn, reads, muls in (("MulTo x64", lambda: rs.check(L.rs_vec_mul(big.h, 128, big.h, 0, big.h, 64, 64)), 2, 1),
                                        ("AddTo x64", lambda: rs.check(L.rs_vec_add(big.h, 128, big.h, 0, big.h, 64, 64)), 2, 0)):
            ms = timed(fn, 20)
            byts = (reads + 1) * 8 * f.L * E * 64
            rows.append({"op": opname, "field": name, "rank": E, "ms": round(ms, 4), "alg_gbs": round(byts / ms / 1e6, 1),
                         "hbm_frac": round(byts / ms / 1e6 / hbm, 3), "alg_gimad_s": round(muls * mul_imad * E * 64 / ms / 1e6, 1),
                         "imad_frac": round(muls * mul_imad * E * 64 / ms / 1e6 / imad, 3), "bound": "hbm"})
        del wb, pwb, out, big
        torch.cuda.empty_cache()
    return {"hbm_peak_gbs": hbm, "imad_wide_peak_g_s": round(imad, 1), "rows": rows}


if do_ntt:
    doc["ntt_sweep"] = ntt_sweep()
if do_k3:
    doc["k3_sweep"] = k3_sweep()
if do_commit:
    doc["commit_sweep"] = commit_sweep()
s = json.dumps(doc)
if out_path:
    open(out_path, "w").write(s + "\n")
print(s)
