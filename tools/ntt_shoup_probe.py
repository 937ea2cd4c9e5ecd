"""zp255 (or argv[1]) rank 2^16 batch 512 negacyclic NTT, forward and inverse, out of place (like bench.py), timed with CUDA
events: ms per transform and an xor digest of the first polynomials, for the kernels the environment selects
(RS_NTT_NO_SHOUP=1: Montgomery lazy kernels; default: Shoup kernels where valid).  python tools/ntt_shoup_probe.py [field] [logN] [batch]"""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_product
rs = load_product(); L = rs.load()
assert L.rs_init(0) == 0
name = sys.argv[1] if len(sys.argv) > 1 else "zp255"
logN = int(sys.argv[2]) if len(sys.argv) > 2 else 16
batch = int(sys.argv[3]) if len(sys.argv) > 3 else 512
q = rs.modulus(name); f = rs.Field(q); N = 1 << logN
rng = np.random.default_rng(1)
host = rng.integers(0, 1 << 62, size=(batch, N, f.L), dtype=np.uint64)
host[..., f.L - 1] %= (q >> (64 * (f.L - 1)))      # canonical: below q
src = rs.Batch(f, batch, N).upload(host); dst = rs.Batch(f, batch, N)
out = {"field": name, "logN": logN, "batch": batch, "no_shoup": os.environ.get("RS_NTT_NO_SHOUP", "0")}
for neg in (True, False):
    t = (rs.CyclotomicTransformer if neg else rs.CyclicTransformer)(f, N)
    for inv in (False, True):
        fn = t.InvNTTTo if inv else t.FwdNTTTo
        for _ in range(3):
            fn(dst, src)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            fn(dst, src)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        chk = int(np.bitwise_xor.reduce(dst.download(0, 4).reshape(-1)))
        out[("neg" if neg else "cyc") + ("_inv" if inv else "_fwd")] = {"ms": round(ms, 4), "mcoeff_s": round(batch * N / ms / 1e3, 1), "xor": chk}
print(json.dumps(out))
