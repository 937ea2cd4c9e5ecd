"""One Jindo Commit of 2^22 coefficients from "vector in pinned memory + seeds" with every sampler on the device (the bench's
commit.e2e path), a few times: the command the launch list of that path is taken of.  python tools/e2e_commit_probe.py [reps]"""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_product
rs = load_product(); L = rs.load()
assert L.rs_init(0) == 0
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
q = rs.modulus("zp255"); f = rs.Field(q); N = 1 << 22
JH = rs.jindo_host
p = JH.new_parameters(q, N, 1)
pdict, cdt = JH.prover_inputs(p)
gp = rs.jindo.Prover(f, pdict, None, None, None, np.array(cdt, dtype=np.uint64), crs=b"Jindo!", key_from_crs=True)
noise = {"ecd_blind_std": p.ecd_blind_std, "mask_blind_std": p.mask_blind_std, "mask_std": p.mask_std, "mask_mlwe_std": p.mask_mlwe_std}
gp.SetTwinSeed(bytes(32)); gp.SetMlweSeed(bytes(range(32)), p.mlwe_std); gp.SetCosacSeeds(b"a", b"b", noise)
tb = rs.jindo.TapeBuilder(pdict, len(cdt[0]), noise, q, threads=8)
tb.set_device_cosac(True)
v = torch.from_numpy(np.random.default_rng(1).integers(0, 1 << 62, size=(N, f.L), dtype=np.uint64).view(np.int64)).pin_memory()
op = gp.NewOpening(); com = np.zeros((p.out_msis, len(p.qout_primes), p.d), dtype=np.uint64)
for k in range(reps + 1):
    tb.start(v.data_ptr(), N, 0); T, _ = tb.wait(0)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    gp.CommitHostInto(v.data_ptr(), N, T, op, com)
    torch.cuda.synchronize()
    print("commit_host ms", (time.perf_counter() - t0) * 1e3)
print("cosac status", gp.CosacStatus())
