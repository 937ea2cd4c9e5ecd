"""Where the host-buffer NTT leg's time goes: pure pinned copies (one way, duplex) against the pipelined upload -> NTT -> download
for several chunk / stream counts (zp255, rank 2^16, batch 512 = 1 GiB each way).  python tools/e2e_probe.py"""
import ctypes, json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_product
rs = load_product(); L = rs.load()
assert L.rs_init(0) == 0
q = rs.modulus("zp255"); f = rs.Field(q); N = 1 << 16; BATCH = 512
words = BATCH * N * f.L
hin = torch.randint(0, 1 << 60, (words,), dtype=torch.int64).pin_memory()
hout = torch.empty_like(hin).pin_memory()
src_t = torch.empty(words, dtype=torch.int64, device="cuda"); dst_t = torch.empty_like(src_t)
src = rs.Batch.from_torch(f, src_t, BATCH, N); dst = rs.Batch.from_torch(f, dst_t, BATCH, N)
ntt = rs.CyclotomicTransformer(f, N)
main = torch.cuda.current_stream()
res = {}


def wall(fn, reps=4):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) * 1e3 / reps


s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def h2d():
    with torch.cuda.stream(s1): src_t.copy_(hin, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): hout.copy_(dst_t, non_blocking=True)
res["h2d_only_ms"] = wall(h2d); res["d2h_only_ms"] = wall(d2h)
res["duplex_ms"] = wall(lambda: (h2d(), d2h()))
gib = words * 8 / 2**30
res["h2d_gbs"] = words * 8 / res["h2d_only_ms"] / 1e6; res["d2h_gbs"] = words * 8 / res["d2h_only_ms"] / 1e6
res["duplex_gbs_each_way"] = words * 8 / res["duplex_ms"] / 1e6

for n_chunks, n_streams in ((8, 3), (16, 3), (16, 4), (32, 4), (64, 4), (32, 8)):
    cp = BATCH // n_chunks; cw = cp * N * f.L
    streams = [torch.cuda.Stream() for _ in range(n_streams)]
    def step():
        for c in range(n_chunks):
            st = streams[c % n_streams]
            L.rs_set_stream(ctypes.c_void_p(st.cuda_stream))
            L.rs_batch_upload(src.h, c * cp, cp, ctypes.c_void_p(hin.data_ptr() + c * cw * 8))
            L.rs_ntt_forward(ntt.h, dst.h, c * cp, src.h, c * cp, cp)
            L.rs_batch_download_async(dst.h, c * cp, cp, ctypes.c_void_p(hout.data_ptr() + c * cw * 8))
        for st in streams:
            st.synchronize()
        L.rs_set_stream(ctypes.c_void_p(main.cuda_stream))
    res[f"pipeline_{n_chunks}x{n_streams}_ms"] = wall(step)
# dedicated copy streams: uploads on one stream, transforms on a second, downloads on a third, chained by events
for n_chunks in (16, 32):
    cp = BATCH // n_chunks; cw = cp * N * f.L
    su, sc, sd = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
    def step3():
        for c in range(n_chunks):
            L.rs_set_stream(ctypes.c_void_p(su.cuda_stream))
            L.rs_batch_upload(src.h, c * cp, cp, ctypes.c_void_p(hin.data_ptr() + c * cw * 8))
            eu = torch.cuda.Event(); eu.record(su)
            sc.wait_event(eu)
            L.rs_set_stream(ctypes.c_void_p(sc.cuda_stream))
            L.rs_ntt_forward(ntt.h, dst.h, c * cp, src.h, c * cp, cp)
            ec = torch.cuda.Event(); ec.record(sc)
            sd.wait_event(ec)
            L.rs_set_stream(ctypes.c_void_p(sd.cuda_stream))
            L.rs_batch_download_async(dst.h, c * cp, cp, ctypes.c_void_p(hout.data_ptr() + c * cw * 8))
        for st in (su, sc, sd):
            st.synchronize()
        L.rs_set_stream(ctypes.c_void_p(main.cuda_stream))
    res[f"three_role_streams_{n_chunks}_ms"] = wall(step3)
print(json.dumps(res))
