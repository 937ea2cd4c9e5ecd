"""Integer-pipe microbenchmarks (rs_bench_imad_peak): G lane-ops/s for each mode."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_product
rs = load_product(); L = rs.load()
assert L.rs_init(0) == 0
names = {0: "32-bit IMAD (loop-invariant product: may be hoisted)", 1: "mad.wide.u32, loop-invariant operands (hoisted -> IADD3 pairs)",
         2: "carry-chained IMAD.WIDE.U32.X", 3: "IADD3.X chains", 4: "1:1 IMAD.WIDE + LOP3 (invariant operands)",
         5: "IMAD.WIDE.U32 64-bit accumulate, changing operands", 6: "DFMA.RZ", 7: "DFMA.RZ + IMAD.WIDE.U32 1:1 (both counted)",
         8: "52x52 limb products (DFMA.RZ / DADD / DFMA.RZ + two 64-bit pattern adds)", 9: "DFMA.RZ + 64-bit add (IADD3 + IADD3.X = one op) 1:1",
         10: "DFMA.RZ + LOP3 1:1", 11: "DFMA.RZ + 32-bit IMAD 1:1", 12: "DADD"}
g = ctypes.c_double()
for m in (5, 2, 1, 0, 3, 4, 5, 6, 12, 7, 9, 10, 11, 8, 6):
    assert L.rs_bench_imad_peak(m, ctypes.byref(g)) == 0
    print(f"mode {m}: {g.value:9.1f} G/s   {names[m]}", flush=True)
