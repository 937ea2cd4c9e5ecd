"""zp255 rank 2^16 batch 512 negacyclic forward NTT, out of place (like bench.py), timed with CUDA events through torch:
ms per transform for the RS_NTT_CHUNK_MB value in the environment.  python tools/ntt_chunk_probe.py"""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_product
rs = load_product(); L = rs.load()
assert L.rs_init(0) == 0
q = rs.modulus("zp255"); f = rs.Field(q); N = 1 << 16; batch = 512
rng = np.random.default_rng(1)
host = rng.integers(0, 1 << 62, size=(batch, N, f.L), dtype=np.uint64)
src = rs.Batch(f, batch, N).upload(host); dst = rs.Batch(f, batch, N)
t = rs.CyclotomicTransformer(f, N)
for _ in range(3):
    t.FwdNTTTo(dst, src)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    t.FwdNTTTo(dst, src)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
chk = int(np.bitwise_xor.reduce(dst.download(0, 2).reshape(-1)))
print(json.dumps({"chunk_mb": os.environ.get("RS_NTT_CHUNK_MB", "0"), "ms": ms, "mcoeff_s": batch * N / ms / 1e3, "xor_check": chk}))
