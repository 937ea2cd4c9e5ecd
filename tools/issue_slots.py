"""How many issue cycles do IMAD.WIDE.U32 and DFMA cost?  One primary instruction followed by K LOP3 (independent chains),
cycles per group from the measured group rate (rs_bench_imad_peak modes 100 + 10 P + K).  Needs the SM clock: read from nvidia-smi
while the kernels run is overkill here -- the pure-LOP3-bound and pure-primary rows calibrate it."""
import ctypes, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_product
rs = load_product(); L = rs.load()
assert L.rs_init(0) == 0
g = ctypes.c_double()
clk = 1.965e9
for P, name, ks in ((0, "IMAD.WIDE.U32 (64-bit addend)", range(0, 7)), (1, "DFMA.RZ", range(0, 5))):
    for K in ks:
        assert L.rs_bench_imad_peak(100 + 10 * P + K, ctypes.byref(g)) == 0
        groups = g.value * 1e9          # counted as 8 per inner iteration = one primary each
        cyc = 32 * 592 * clk / groups
        print(f"{name:32s} + {K} LOP3: {g.value:9.1f} G groups/s = {cyc:5.2f} cycles per group at 1965 MHz ({(1 + K) / cyc:4.2f} instr/cycle)", flush=True)
