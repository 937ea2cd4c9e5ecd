// ntt_rr_l8.cu -- launcher of the reduced-radix lazy forward NTT (ntt_rr.cuh) for 8-limb fields.
#include "ntt_launch.cuh"
#include "ntt_rr.cuh"

#ifndef RS_RR_R
#define RS_RR_R 3          // radix-2^R rounds (tuning knob)
#endif
#ifndef RS_RR_MINB
#define RS_RR_MINB 2       // resident CTAs per SM the register allocation is capped for
#endif

namespace rs {

template <int W, int QZ, bool IN_RR, bool OUT_RR>
static int launch_rr_pass(const rs_ntt *t, uint32_t *out, const uint32_t *in, const RRConsts<W> &C, const PassGeom &G,
                          unsigned long long tiles) {
    constexpr int NL = 8, R = RS_RR_R, THREADS = 256, TS = THREADS << R;
    auto kern = ntt_rr_pass_kernel<NL, W, QZ, R, THREADS, RS_RR_MINB, IN_RR, OUT_RR>;
    constexpr int smem = RRSmem<W, TS>::BYTES;
    static bool configured = false;
    if (!configured) {
        RS_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = true;
    }
    RS_REQUIRE(tiles <= 0x7fffffffull, RS_ERR_UNSUPPORTED, "ntt: batch too large for one launch");
    kern<<<(unsigned)tiles, THREADS, smem, current_stream()>>>(out, in, t->d_tw_rr, C, G);
    count_launch();
    RS_CUDA_CHECK(cudaGetLastError());
    if (g_pass_profile && g_pass_profile->n < 16) RS_CUDA_CHECK(cudaEventRecord(g_pass_profile->ev[g_pass_profile->n++], current_stream()));
    return RS_OK;
}

template <int W, int QZ>
static int rr_run(rs_ntt *t, uint32_t *out, const uint32_t *in, size_t n_polys) {
    constexpr int LOG_TS = 8 + RS_RR_R, TS = 1 << LOG_TS;
    const int logN = t->logN;
    if (n_polys == 0) return RS_OK;
    RRConsts<W> C;
    for (int j = 0; j < W; j++) { C.P.q[j] = t->rr_q[j]; C.P.c3q[j] = t->rr_c3q[j]; C.P.q2[j] = t->rr_q2[j]; C.P.one[j] = t->rr_one[j]; }
    const unsigned long long total = (unsigned long long)n_polys << logN;
    const int KB = logN < LOG_TS ? logN : LOG_TS;
    const int top = logN - KB;
    const int KA_MAX = LOG_TS - 3;
    const int n_row = top ? (top + KA_MAX - 1) / KA_MAX : 0;
    PassGeom passes[8];
    int np = 0, s = 0;
    for (int r = 0; r < n_row; r++) {
        const int K = (top - s + (n_row - r) - 1) / (n_row - r);
        PassGeom G{};
        G.logN = logN; G.K = K; G.s0 = s; G.pl = logN - s - K; G.logW = LOG_TS - K; G.col_major = 0; G.total_elems = total;
        passes[np++] = G;
        s += K;
    }
    {
        PassGeom G{};
        G.logN = logN; G.K = KB; G.s0 = top; G.pl = 0; G.logW = LOG_TS - KB; G.col_major = 1; G.total_elems = total;
        passes[np++] = G;
    }
    const unsigned long long tiles_col = (total + TS - 1) / TS, tiles_row = total / TS;
    if (np > 1) {
        const size_t need = (size_t)total * W;
        if (t->scratch_words < need) {
            if (t->d_scratch) { RS_CUDA_CHECK(cudaStreamSynchronize(current_stream())); cudaFree(t->d_scratch); t->d_scratch = nullptr; t->scratch_words = 0; }
            RS_CUDA_CHECK(cudaMalloc((void **)&t->d_scratch, need * 4));
            t->scratch_words = need;
        }
    }
    for (int p = 0; p < np; p++) {
        const PassGeom &G = passes[p];
        const unsigned long long tiles = G.col_major ? tiles_col : tiles_row;
        const bool first = (p == 0), last = (p == np - 1);
        int rc;
        if (first && last) rc = launch_rr_pass<W, QZ, false, false>(t, out, in, C, G, tiles);
        else if (first) rc = launch_rr_pass<W, QZ, false, true>(t, t->d_scratch, in, C, G, tiles);
        else if (last) rc = launch_rr_pass<W, QZ, true, false>(t, out, t->d_scratch, C, G, tiles);
        else rc = launch_rr_pass<W, QZ, true, true>(t, t->d_scratch, t->d_scratch, C, G, tiles);
        if (rc) return rc;
    }
    return RS_OK;
}

int ntt_rr_run_l8(rs_ntt *t, uint32_t *out, const uint32_t *in, size_t n_polys) {
    if (t->f->NL != 8) return RS_ERR_UNSUPPORTED;
    if (t->rr_W == 9) return t->rr_QZ ? rr_run<9, 1>(t, out, in, n_polys) : rr_run<9, 0>(t, out, in, n_polys);
    if (t->rr_W == 8) return t->rr_QZ ? rr_run<8, 1>(t, out, in, n_polys) : rr_run<8, 0>(t, out, in, n_polys);
    return RS_ERR_UNSUPPORTED;
}

}  // namespace rs
