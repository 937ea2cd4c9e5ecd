"""[tests/perf: measurement harness that checks against / times the oracle, hence under tests/ -- the product never imports oracle/]
Developer timing script: Jindo Commit at BASELINE configs[2] scale (2^22 coefficients, 255-bit
field, NewParameters(2^22, 1)) with synthetic key / tapes resident in HBM."""
import ctypes, os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_product
from oracle import pyref, jindo_ref as J

logN = int(sys.argv[1]) if len(sys.argv) > 1 else 22
rs = load_product(); L = rs.load()
assert L.rs_init(0) == 0
q = pyref.modulus("zp255"); f = rs.Field(q)
N = 1 << logN
p = J.new_parameters(q, N, 1)
enc = J.Encoder(p, J.Tape(b"s"))
print("params", p.rows, p.cols, p.slots, p.q_primes, p.qout_primes, p.in_msis, p.mlwe, p.out_msis, flush=True)
rng = np.random.default_rng(1)
def rpolys(shape, primes):
    a = np.empty(shape + (len(primes), p.d), dtype=np.uint64)
    for l, pr in enumerate(primes):
        a[..., l, :] = rng.integers(0, pr, size=shape + (p.d,), dtype=np.uint64)
    return a
sys.path.insert(0, os.path.join(ROOT, "tests"))
from jindo_util import params_dict, cdt_array
gp = rs.jindo.Prover(f, params_dict(p, enc), rpolys((p.in_msis, p.rows), p.q_primes), rpolys((p.in_msis, p.mlwe), p.q_primes),
                     rpolys((p.out_msis, p.in_com_dcmp_len), p.qout_primes), cdt_array(enc))
def felems(n):
    a = rng.integers(0, 1 << 62, size=(n, f.L), dtype=np.uint64); return a
v = torch.from_numpy(felems(N).view(np.int64)).cuda()
B = rs.Batch.from_torch(f, v, 1, N)
cs = p.cols * p.slots
n_twin = p.cols * (p.rows - 1); n_cosac = p.cols + p.rows; n_noise = (p.cols + 1) * (p.in_msis + p.mlwe)
dev = {
 "first_row": torch.from_numpy(felems(cs).view(np.int64)).cuda(), "last_row": torch.from_numpy(felems(cs).view(np.int64)).cuda(),
 "mask_rows": torch.from_numpy(felems(p.rows * p.slots).view(np.int64)).cuda(),
 "twin_u": torch.from_numpy(rng.integers(0, 1 << 63, size=n_twin * p.d, dtype=np.int64)).cuda(),
 "cosac_g": torch.from_numpy(rng.integers(-5000, 5000, size=n_cosac * p.d, dtype=np.int64)).cuda(),
 "mlwe": torch.from_numpy(rng.integers(-20, 20, size=n_noise * p.d, dtype=np.int64)).cuda(),
}
T = rs.capi.JindoTape()
T.first_row, T.last_row, T.mask_rows = dev["first_row"].data_ptr(), dev["last_row"].data_ptr(), dev["mask_rows"].data_ptr()
T.n_mask_rows = p.rows; T.twin_u = dev["twin_u"].data_ptr(); T.n_twin_u = n_twin * p.d
T.cosac_g = dev["cosac_g"].data_ptr(); T.n_cosac_polys = n_cosac; T.mlwe = dev["mlwe"].data_ptr(); T.on_device = 1
op = gp.NewOpening()
com = np.zeros((p.out_msis, len(p.qout_primes), p.d), dtype=np.uint64)
for _ in range(2):
    gp.CommitInto(B, 0, N, T, op, com)
torch.cuda.synchronize()
l0 = L.rs_launch_count(); t0 = time.perf_counter()
reps = 5
for _ in range(reps):
    gp.CommitInto(B, 0, N, T, op, com)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / reps
print(f"commit 2^{logN}: {dt*1e3:.3f} ms/commit  {1/dt:.1f} polys/s  {N/dt/1e6:.1f} Mcoeff/s  launches/commit={(L.rs_launch_count()-l0)//reps}")
com2 = com.copy(); gp.CommitInto(B, 0, N, T, op, com); assert np.array_equal(com, com2)
# ---- Evaluate inner loops (prover.go:281-316) on the same opening
left = rpolys((p.rows,), p.q_primes); chals = rpolys((p.cols,), p.q_primes)
for fn, name in ((lambda: gp.Partial(op, left), "Partial"), (lambda: gp.Respond(op, chals), "Respond")):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5): fn()
    torch.cuda.synchronize()
    print(f"{name}: {(time.perf_counter() - t0) / 5 * 1e3:.3f} ms", flush=True)
# ---- column shards on one GPU: how the per-rank share of a sharded Commit scales (no exchange here)
if "--shards" in sys.argv:
    for world in (1, 2, 4, 8):
        per = -(-(p.cols + 1) // world)
        for r in sorted({0, world - 1}):
            c0, c1 = min(r * per, p.cols + 1), min((r + 1) * per, p.cols + 1)
            for _ in range(2):
                gp.CommitColumns(B, 0, N, T, c0, c1, op)
            torch.cuda.synchronize()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter(); e0.record()
            for _ in range(reps):
                gp.CommitColumns(B, 0, N, T, c0, c1, op)
            e1.record(); torch.cuda.synchronize()
            print(f"world {world} rank {r}: columns [{c0},{c1})  wall {(time.perf_counter() - t0) / reps * 1e3:.3f} ms  device {e0.elapsed_time(e1) / reps:.3f} ms", flush=True)
