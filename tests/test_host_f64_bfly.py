"""CPU check of the exactness claim behind the FP64 warp NTT (csrc/jindo.cu f64_bfly): the fma-based modular product is
compared with 128-bit integer arithmetic for primes up to the 2^44 limit, |x| < 9q (tests/host_f64_bfly_test.c)."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_f64_modular_product_is_exact(tmp_path):
    src = os.path.join(ROOT, "tests", "host_f64_bfly_test.c")
    exe = str(tmp_path / "f64_bfly")
    # -ffp-contract=off: the device code uses explicit __dmul_rn / __fma_rn; the host must not fuse x*w + c on its own
    r = subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-o", exe, src, "-lm"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe, "1000000"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.strip() == "ok", r.stdout + r.stderr
