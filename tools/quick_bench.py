"""Developer timing script (not the contract bench): NTT throughput per field / rank."""
import ctypes, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_product
pyref = load_product()  # product-side modulus table (bigpoly.MODULI)

rs = load_product(); L = rs.load()
assert L.rs_init(0) == 0, L.rs_last_error()
g = ctypes.c_double()
for wide in (1, 0, 2, 3, 4):
    assert L.rs_bench_imad_peak(wide, ctypes.byref(g)) == 0
    print(f"imad peak wide={wide}: {g.value:.1f} G/s", flush=True)
cases = [("zp255", 16, 512), ("zp128", 16, 512), ("zp255", 12, 4096), ("zp255", 14, 1024), ("zp255", 17, 256),
         ("zp220", 16, 256), ("zp240", 16, 256), ("zp440", 14, 512), ("zp880", 13, 256)]
if len(sys.argv) > 1:
    cases = [c for c in cases if c[0] in sys.argv[1:]]
for name, logN, batch in cases:
    q = pyref.modulus(name); f = rs.Field(q); N = 1 << logN
    rng = np.random.default_rng(1)
    host = rng.integers(0, 1 << 62, size=(batch, N, f.L), dtype=np.uint64)
    host[..., f.L - 1] &= np.uint64((1 << ((q.bit_length() - 1) % 64)) - 1)
    B = rs.Batch(f, batch, N).upload(host)
    for kind in ("cyclic", "cyclotomic"):
        t = rs.CyclicTransformer(f, N) if kind == "cyclic" else rs.CyclotomicTransformer(f, N)
        for inv in (0, 1):
            ms = (ctypes.c_float * 16)(); n = ctypes.c_int(16)
            assert L.rs_ntt_profile(t.h, B.h, batch, inv, 2, ms, ctypes.byref(n)) == 0
            assert L.rs_ntt_profile(t.h, B.h, batch, inv, 5, ms, ctypes.byref(n)) == 0, L.rs_last_error()
            tot = sum(ms[i] for i in range(n.value))
            print(f"{name} {kind} logN={logN} batch={batch} inv={inv}: passes={[round(ms[i],3) for i in range(n.value)]} ms "
                  f"total={tot:.3f} ms  {batch*N/tot/1e3:.0f} Mcoeff/s  {batch*N*f.L*16/tot/1e6:.0f} GB/s(alg)", flush=True)
