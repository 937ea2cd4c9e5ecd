"""Device AES-CTR keystream rate: draws 2^26 uniform u64 (537 MB, what a 2^22 Commit consumes) from the TwinCDT stream and times it.
RS_AES_CTAS = resident CTAs per SM of aes_keystream_kernel.  python tools/aes_probe.py"""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_product
rs = load_product(); L = rs.load()
assert L.rs_init(0) == 0
q = rs.modulus("zp255"); f = rs.Field(q)
JH = rs.jindo_host
p = JH.new_parameters(q, 1 << 22, 1)
pdict, cdt = JH.prover_inputs(p)
gp = rs.jindo.Prover(f, pdict, None, None, None, cdt, crs=b"Jindo!", key_from_crs=True)
gp.SetTwinSeed(bytes(32))
out = gp.TwinStreamDraw(1 << 10)
n = 1 << 26
import ctypes
buf = torch.empty(n, dtype=torch.int64).pin_memory()
for k in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    rs.check(L.rs_jindo_twin_stream_draw(gp.h, n, ctypes.c_void_p(buf.data_ptr())))
    dt = (time.perf_counter() - t0) * 1e3
    print("ctas", os.environ.get("RS_AES_CTAS", "6"), "draw+D2H ms", round(dt, 2), "xor", hex(int(buf[:1000].numpy().view(np.uint64)[-1])))
