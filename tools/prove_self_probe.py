"""Self-sampled compiled Buckler Prove of examples/mult at N = 2^13 (product side only): wall time per proof with / without the
phase synchronisations, COSAC on the host threads / on the device.  python tools/prove_self_probe.py [threads]"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_product
rs = load_product(); L = rs.load()
assert L.rs_init(0) == 0
threads = int(sys.argv[1]) if len(sys.argv) > 1 else 15
N = 1 << 13
q = rs.modulus("zp128"); f = rs.Field(q); JH = rs.jindo_host
desc, commit_rank, batch = rs.circuits.mult_circuit(q, N)
p = JH.new_parameters(q, commit_rank, batch)
pdict, cdt = JH.prover_inputs(p)
gp = rs.jindo.Prover(f, pdict, None, None, None, cdt, crs=b"Buckler!", key_from_crs=True)
noise = {"ecd_blind_std": p.ecd_blind_std, "mask_blind_std": p.mask_blind_std, "mask_std": p.mask_std, "mask_mlwe_std": p.mask_mlwe_std}
gp.SetTwinSeed(os.urandom(32)); gp.SetMlweSeed(os.urandom(32), p.mlwe_std); gp.SetCosacSeeds(os.urandom(32), os.urandom(32), noise)
pr = rs.buckler.NativeProver(f, q, desc, p, gp)
pr.enable_sampling(noise, threads=threads)
insts = [rs.circuits.mult_witness(f, q, N, 77 + k) for k in range(4)]
out = {"streams": os.environ.get("RS_BUCKLER_STREAMS", "4"), "threads": threads}
for dev in (False, True):
    pr.set_device_cosac(dev)
    for phases in (False, True):
        pr.prove_raw(insts[0][0], insts[0][1], None, phases=phases)
        ts = []
        for k in range(8):
            t0 = time.perf_counter()
            pr.prove_raw(insts[k % 4][0], insts[k % 4][1], None, phases=phases, reuse=True)
            ts.append((time.perf_counter() - t0) * 1e3)
        out[f"cosac_{'device' if dev else 'host'}_phases_{phases}"] = [round(min(ts), 2), round(sorted(ts)[len(ts) // 2], 2)]
    out[f"split_{'device' if dev else 'host'}"] = {k: round(v, 2) for k, v in pr.timing.items() if v}
print(json.dumps(out))
