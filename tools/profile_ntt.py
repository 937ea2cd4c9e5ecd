"""Small fixed workload for ncu captures: zp255 negacyclic forward+inverse NTT, rank 2^16, batch 64."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_product
pyref = load_product()  # product-side modulus table (bigpoly.MODULI)
rs = load_product(); L = rs.load()
assert L.rs_init(0) == 0
name = sys.argv[1] if len(sys.argv) > 1 else "zp255"
logN = int(sys.argv[2]) if len(sys.argv) > 2 else 16
batch = int(sys.argv[3]) if len(sys.argv) > 3 else 64
q = pyref.modulus(name); f = rs.Field(q); N = 1 << logN
rng = np.random.default_rng(1)
host = rng.integers(0, 1 << 62, size=(batch, N, f.L), dtype=np.uint64)
host[..., f.L - 1] &= np.uint64((1 << ((q.bit_length() - 1) % 64)) - 1)
B = rs.Batch(f, batch, N).upload(host)
t = rs.CyclotomicTransformer(f, N)
c = rs.CyclicTransformer(f, N)
for _ in range(3):
    t.FwdNTTTo(B, B)
    t.InvNTTTo(B, B)
for _ in range(2):
    c.FwdNTTTo(B, B)
    c.InvNTTTo(B, B)
assert L.rs_sync() == 0
assert np.array_equal(B.download(), host)
print("ok")
