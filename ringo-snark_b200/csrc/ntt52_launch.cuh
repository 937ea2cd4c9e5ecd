// ntt52_launch.cuh -- pass planning + launches for ntt52_pass_kernel (FP64-pipe butterflies, ntt52.cuh).
// Same plan as ntt_run (ntt_launch.cuh): column passes over the top stages, then the contiguous pass; the inverse
// runs them in reverse.  One translation unit per tuning variant (ntt52_v*.cu).
#pragma once
#include <mutex>

#include "common.cuh"
#include "ntt52.cuh"

namespace rs {

template <int R, int LOG_THREADS, int MINB_FWD, int MINB_INV>
struct Ntt52Variant {
    static constexpr int kR = R, kLogThreads = LOG_THREADS, kThreads = 1 << LOG_THREADS, kMinbFwd = MINB_FWD, kMinbInv = MINB_INV;
    static constexpr int kLogTS = LOG_THREADS + R, kTS = 1 << (LOG_THREADS + R);
};

template <class V, bool Q0ONE, bool INV, bool LAST>
int launch_pass52(const rs_ntt *t, uint32_t *out, const uint32_t *in, const PassGeom &G, unsigned long long tiles) {
    constexpr int MINB = INV ? V::kMinbInv : V::kMinbFwd;
    auto kern = ntt52_pass_kernel<8, 5, Q0ONE, V::kR, V::kThreads, MINB, INV, LAST>;
    constexpr int smem = Smem52<5, V::kTS>::BYTES;
    RS_CUDA_CHECK(ensure_smem_attr((const void *)kern, smem));
    RS_REQUIRE(tiles <= 0x7fffffffull, RS_ERR_UNSUPPORTED, "ntt: batch too large for one launch");
    kern<<<(unsigned)tiles, V::kThreads, smem, current_stream()>>>(out, in, INV ? t->d_twinv52 : t->d_tw52, t->c52, G);
    count_launch();
    RS_CUDA_CHECK(cudaGetLastError());
    if (g_pass_profile && g_pass_profile->n < 16) RS_CUDA_CHECK(cudaEventRecord(g_pass_profile->ev[g_pass_profile->n++], current_stream()));
    return RS_OK;
}

template <class V, bool Q0ONE>
int ntt52_run(const rs_ntt *t, uint32_t *out, const uint32_t *in, size_t n_polys, bool inverse) {
    const int logN = t->logN;
    if (n_polys == 0) return RS_OK;
    const unsigned long long total = (unsigned long long)n_polys << logN;
    const int KB = logN < V::kLogTS ? logN : V::kLogTS;
    const int top = logN - KB;
    const int KA_MAX = V::kLogTS - 3;
    const int n_row = top ? (top + KA_MAX - 1) / KA_MAX : 0;
    PassGeom passes[8];
    int np = 0, s = 0;
    for (int r = 0; r < n_row; r++) {
        const int K = (top - s + (n_row - r) - 1) / (n_row - r);
        PassGeom G{};
        G.logN = logN; G.K = K; G.s0 = s; G.pl = logN - s - K; G.logW = V::kLogTS - K; G.col_major = 0;
        G.total_elems = total; G.triv = !t->negacyclic;
        passes[np++] = G;
        s += K;
    }
    {
        PassGeom G{};
        G.logN = logN; G.K = KB; G.s0 = top; G.pl = 0; G.logW = V::kLogTS - KB; G.col_major = 1;
        G.total_elems = total; G.triv = !t->negacyclic;
        passes[np++] = G;
    }
    const unsigned long long tiles_col = (total + V::kTS - 1) / V::kTS;
    const unsigned long long tiles_row = total / V::kTS;
    const uint32_t *src = in;
    for (int k = 0; k < np; k++) {
        const int p = inverse ? np - 1 - k : k;
        const PassGeom &G = passes[p];
        const unsigned long long tiles = G.col_major ? tiles_col : tiles_row;
        const bool last = (k == np - 1);
        int rc;
        if (!inverse) rc = last ? launch_pass52<V, Q0ONE, false, true>(t, out, src, G, tiles) : launch_pass52<V, Q0ONE, false, false>(t, out, src, G, tiles);
        else rc = last ? launch_pass52<V, Q0ONE, true, true>(t, out, src, G, tiles) : launch_pass52<V, Q0ONE, true, false>(t, out, src, G, tiles);
        if (rc) return rc;
        src = out;
    }
    return RS_OK;
}

template <class V>
int ntt52_run_variant(const rs_ntt *t, uint32_t *out, const uint32_t *in, size_t n_polys, bool inverse) {
    return t->c52.F.q0one ? ntt52_run<V, true>(t, out, in, n_polys, inverse) : ntt52_run<V, false>(t, out, in, n_polys, inverse);
}

}  // namespace rs
