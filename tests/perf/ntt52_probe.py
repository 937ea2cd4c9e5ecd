"""[tests/perf: developer harness that takes inputs from / checks against the oracle, hence under tests/ -- the product never imports oracle/]
FP64-pipe NTT probe (one GPU): pipe microbenchmarks, then for every ntt52 tuning variant the headline transform
(zp255, rank 2^16, batch 512, all four kinds) timed with CUDA events inside the library and compared bit for bit
with the integer-pipe kernels (RS_NTT_NO52) on the same input.

    python tests/perf/ntt52_probe.py [--out=FILE] [--batch=512] [--logn=16] [--variants=0,1,2,3]
"""
import ctypes, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
import torch

opt = {a.split("=")[0][2:]: a.split("=")[1] for a in sys.argv[1:] if a.startswith("--") and "=" in a}
batch = int(opt.get("batch", 512)); logN = int(opt.get("logn", 16)); N = 1 << logN
variants = [int(v) for v in opt.get("variants", "2").split(",") if v != ""]
mixes = [m for m in opt.get("mix", "1/4,1/3").split(",") if m != ""]
micro = opt.get("micro", "1") == "1"
fields = opt.get("fields", "zp255").split(",")

rs = bench.product(); L = rs.load()
rs.check(L.rs_init(0))
rs.check(L.rs_set_stream(ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
doc = {"gpu": torch.cuda.get_device_name(0), "batch": batch, "logN": logN, "micro": {}, "ntt": []}

names = {5: "IMAD.WIDE.U32 64-bit accumulate", 3: "IADD3.X chains", 6: "DFMA.RZ", 7: "DFMA.RZ + IMAD.WIDE 1:1 (both counted)",
         8: "52x52 limb products (DFMA.RZ, DADD, DFMA.RZ + 2 pattern adds)"}
g = ctypes.c_double()
for m in ((5, 3, 6, 7, 8, 6) if micro else ()):
    rs.check(L.rs_bench_imad_peak(m, ctypes.byref(g)))
    doc["micro"][f"mode{m}"] = {"what": names[m], "g_per_s": round(g.value, 1)}
    print(f"mode {m}: {g.value:9.1f} G/s   {names[m]}", flush=True)


def time_ntt(tr, B, inv, reps):
    ms = (ctypes.c_float * 16)(); n = ctypes.c_int(16)
    rs.check(L.rs_ntt_profile(tr.h, B.h, batch, inv, 2, ms, ctypes.byref(n)))
    n = ctypes.c_int(16)
    rs.check(L.rs_ntt_profile(tr.h, B.h, batch, inv, reps, ms, ctypes.byref(n)))
    return [round(ms[i], 4) for i in range(n.value)]


from oracle import pyref   # moduli table only (test-side tool)
for name in fields:
    q = pyref.modulus(name); f = rs.Field(q)
    host = bench.synthetic_polys(q, f.L, min(batch, 16), N, 0x4E5454)
    full = np.ascontiguousarray(np.resize(host, (batch, N, f.L)))
    src = torch.from_numpy(full.view(np.int64).reshape(-1)).cuda()
    ref = {}
    for neg in (True, False):
        tr = (rs.CyclotomicTransformer if neg else rs.CyclicTransformer)(f, N)
        for inv in (0, 1):
            buf = src.clone(); B = rs.Batch.from_torch(f, buf, batch, N)
            rs.check((L.rs_ntt_inverse if inv else L.rs_ntt_forward)(tr.h, B.h, 0, B.h, 0, batch)); torch.cuda.synchronize()
            ref[(neg, inv)] = buf
            t = time_ntt(tr, rs.Batch.from_torch(f, src.clone(), batch, N), inv, 5)
            row = {"field": name, "path": "int", "kind": "negacyclic" if neg else "cyclic", "dir": "inverse" if inv else "forward",
                   "pass_ms": t, "ms": round(sum(t), 4), "mcoeff_s": round(batch * N / sum(t) / 1e3, 1)}
            doc["ntt"].append(row); print(row, flush=True)
    for v in variants:
        os.environ["RS_NTT52"] = "1"
        for neg in (True, False):
            tr = (rs.CyclotomicTransformer if neg else rs.CyclicTransformer)(f, N)
            for inv in (0, 1):
                buf = src.clone(); B = rs.Batch.from_torch(f, buf, batch, N)
                rs.check((L.rs_ntt_inverse if inv else L.rs_ntt_forward)(tr.h, B.h, 0, B.h, 0, batch)); torch.cuda.synchronize()
                ok = bool(torch.equal(buf, ref[(neg, inv)]))
                t = time_ntt(tr, rs.Batch.from_torch(f, src.clone(), batch, N), inv, 5)
                row = {"field": name, "path": f"fp64 v{v}", "kind": "negacyclic" if neg else "cyclic", "dir": "inverse" if inv else "forward",
                       "bit_exact_vs_int": ok, "pass_ms": t, "ms": round(sum(t), 4), "mcoeff_s": round(batch * N / sum(t) / 1e3, 1)}
                doc["ntt"].append(row); print(row, flush=True)
    if "RS_NTT52" in os.environ:
        del os.environ["RS_NTT52"]
    for mx in mixes:
        os.environ["RS_NTT_MIX"] = mx
        for neg in (True, False):
            tr = (rs.CyclotomicTransformer if neg else rs.CyclicTransformer)(f, N)
            for inv in (0, 1):
                buf = src.clone(); B = rs.Batch.from_torch(f, buf, batch, N)
                rs.check((L.rs_ntt_inverse if inv else L.rs_ntt_forward)(tr.h, B.h, 0, B.h, 0, batch)); torch.cuda.synchronize()
                ok = bool(torch.equal(buf, ref[(neg, inv)]))
                t = time_ntt(tr, rs.Batch.from_torch(f, src.clone(), batch, N), inv, 5)
                row = {"field": name, "path": f"mix {mx}", "kind": "negacyclic" if neg else "cyclic", "dir": "inverse" if inv else "forward",
                       "bit_exact_vs_int": ok, "pass_ms": t, "ms": round(sum(t), 4), "mcoeff_s": round(batch * N / sum(t) / 1e3, 1)}
                doc["ntt"].append(row); print(row, flush=True)
    if "RS_NTT_MIX" in os.environ:
        del os.environ["RS_NTT_MIX"]
    del src, ref
    torch.cuda.empty_cache()
if "out" in opt:
    with open(opt["out"], "w") as fh:
        json.dump(doc, fh, indent=1)
