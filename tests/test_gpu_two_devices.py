"""GPU test for one process driving two devices (needs 2 GPUs; skipped otherwise): per-device one-time setup of the library
-- the AES S-box in __constant__ memory, the opt-in to > 48 KB of dynamic shared memory per kernel -- must happen on every
device a context is created on.  Two OS threads bind to device 0 and device 1 with rs_init (the binding is per thread, as
INTEGRATION.md section 5 describes for goroutines), and each runs an NTT and a Jindo Commit (device-generated commit key, encode,
mat-vec, cut-off) against the oracle."""
import threading

import numpy as np
import pytest

import oracle
from oracle import pyref

pytestmark = pytest.mark.gpu


draws = {}


def _work(rs, device, out):
    try:
        import __graft_entry__ as ge
        L = rs.load()
        assert L.rs_init(device) == 0, L.rs_last_error()
        q = pyref.modulus("zp255")
        fo = oracle.Field(q); fg = rs.Field(q)
        N = 1 << 12
        rng = np.random.default_rng(100 + device)
        data = np.stack([fo.to_mont([int.from_bytes(rng.bytes(40), "little") % q for _ in range(N)]) for _ in range(2)])
        for neg in (False, True):
            to = oracle.NTT(fo, 12, neg)
            tg = (rs.CyclotomicTransformer if neg else rs.CyclicTransformer)(fg, N)
            src = rs.Batch(fg, 2, N).upload(data); dst = rs.Batch(fg, 2, N)
            tg.FwdNTTTo(dst, src)
            assert np.array_equal(dst.download(), to.fwd(data))
            tg.InvNTTTo(dst, dst)
            assert np.array_equal(dst.download(), data)
        ge._smoke_commit(rs, fo, fg, q)          # Commit + InCommit + leftVec against the C restatement
        # the AES-CTR source (S-box in __constant__ memory, uploaded per device): commit key from the CRS + TwinCDT draws
        JH = rs.jindo_host
        p = JH.new_parameters(q, 1 << 12, 1)
        pdict, cdt = JH.prover_inputs(p)
        gp = rs.jindo.Prover(fg, pdict, None, None, None, np.array(cdt, dtype=np.uint64), crs=b"Jindo!", key_from_crs=True)
        gp.SetTwinSeed(bytes(range(32)))
        draws[device] = np.array(gp.TwinStreamDraw(1 << 12), dtype=np.uint64).copy()
        out[device] = "ok"
    except BaseException as e:   # noqa: BLE001 -- reported by the main thread
        out[device] = repr(e)


def test_two_devices_from_one_process(product):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    out = {}
    # device 1 FIRST: a process-wide "already set up" flag left by device 0 is exactly the bug this guards against
    for dev in (1, 0):
        t = threading.Thread(target=_work, args=(product, dev, out))
        t.start(); t.join()
    assert out == {0: "ok", 1: "ok"}, out
    assert np.array_equal(draws[0], draws[1]) and len(np.unique(draws[1])) > 4000     # same keystream on both devices
    # and both at once
    out2 = {}
    ts = [threading.Thread(target=_work, args=(product, dev, out2)) for dev in (0, 1)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert out2 == {0: "ok", 1: "ok"}, out2
