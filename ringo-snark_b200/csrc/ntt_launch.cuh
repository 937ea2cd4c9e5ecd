// ntt_launch.cuh -- host-side pass planning + launches for ntt_pass_kernel (see ntt.cuh).
#pragma once
#include <stdlib.h>

#include "common.cuh"
#include "ntt.cuh"

namespace rs {

#ifndef RS_NTT_R_SMALL
#define RS_NTT_R_SMALL 3      // radix-2^R rounds for <= 8 limbs (tuning knob)
#endif
#ifndef RS_NTT_R_L4
#define RS_NTT_R_L4 3         // same knob for the 4-limb (128-bit) fields
#endif
#ifndef RS_NTT_MINB_SMALL
#define RS_NTT_MINB_SMALL 4   // resident CTAs per SM the register allocation is capped for (<= 8 limbs)
#endif
#ifndef RS_NTT_MINB_FWD
#define RS_NTT_MINB_FWD 5     // forward passes of the 8-limb fields: 102 registers (8 bytes of spill), 20 warps per SM: +3-5 %;
#endif                        // the inverse passes (multiply after the add/sub pair, N^-1 folded in) lose 3 % there and keep 4
#ifndef RS_NTT_MINB_L4
#define RS_NTT_MINB_L4 8        // 4-limb (128-bit) fields: 64 registers suffice, 8 CTAs of 128 threads per SM (+4 %)
#endif
#ifndef RS_NTT_LOG_THREADS
// log2 threads per CTA for <= 8 limbs.  128 threads x 4 CTAs per SM instead of 256 x 2 (same registers, same shared
// memory, one stage moves from the contiguous pass to the column pass): the CTAs' load / exchange / store phases
// interleave better and the FMA pipe idles less -- measured +7 % (255-bit) to +14 % (128-bit) at rank 2^16.
#define RS_NTT_LOG_THREADS 7
#endif

#ifndef RS_NTT_LOG_THREADS_L14
#define RS_NTT_LOG_THREADS_L14 7   // 14-limb (440-bit) fields: 128 threads x 3 CTAs per SM instead of 256 x 1: +35 %
#endif
#ifndef RS_NTT_MINB_L14
#define RS_NTT_MINB_L14 3
#endif
#ifndef RS_NTT_LOG_THREADS_L28
#define RS_NTT_LOG_THREADS_L28 7   // 28-limb (880-bit) fields: 128 threads x 3 CTAs per SM: +8 % forward, +21 % inverse
#endif
#ifndef RS_NTT_MINB_L28
#define RS_NTT_MINB_L28 3
#endif

template <int NL> struct NttTuning {
    // 2^R elements per thread between shared-memory exchanges
    static constexpr int R = (NL <= 4) ? RS_NTT_R_L4 : ((NL <= 8) ? RS_NTT_R_SMALL : (NL <= 14 ? 2 : 1));
    static constexpr int LOG_THREADS = (NL <= 8) ? RS_NTT_LOG_THREADS : (NL <= 14 ? RS_NTT_LOG_THREADS_L14 : RS_NTT_LOG_THREADS_L28);
    static constexpr int THREADS = 1 << LOG_THREADS;
    static constexpr int TS = THREADS << R;
    static constexpr int LOG_TS = LOG_THREADS + R;
    static constexpr int MINB = (NL <= 4) ? RS_NTT_MINB_L4 : ((NL <= 8) ? RS_NTT_MINB_SMALL : (NL <= 14 ? RS_NTT_MINB_L14 : RS_NTT_MINB_L28));
};

#ifndef RS_NTT_MINB_SHOUP
#define RS_NTT_MINB_SHOUP 4   // Shoup passes of the 8-limb fields (two twiddle words per product live next to the tile registers)
#endif

template <int NL, int QS, bool INV, bool LAST, int MODE, int FID = FID_RUNTIME>
int launch_pass(const rs_ntt *t, uint32_t *out, const uint32_t *in, const NttConsts<NL> &C, const PassGeom &G,
                unsigned long long tiles) {
    using T = NttTuning<NL>;
    constexpr int MINB = (MODE == NTT_SHOUP && NL > 4) ? RS_NTT_MINB_SHOUP : ((!INV && NL > 4 && NL <= 8) ? RS_NTT_MINB_FWD : T::MINB);
    auto kern = ntt_pass_kernel<NL, QS, T::R, T::THREADS, MINB, INV, (INV && LAST), MODE, LAST, FID>;
    constexpr int smem = SmemLayout<NL, T::TS>::BYTES;
    RS_CUDA_CHECK(ensure_smem_attr((const void *)kern, smem));
    RS_REQUIRE(tiles <= 0x7fffffffull, RS_ERR_UNSUPPORTED, "ntt: batch too large for one launch");
    const uint32_t *tw = (MODE == NTT_SHOUP) ? (INV ? t->d_twinv_sh : t->d_tw_sh) : (INV ? t->d_twinv : t->d_tw);
    kern<<<(unsigned)tiles, T::THREADS, smem, current_stream()>>>(out, in, tw, C, G);
    count_launch();
    RS_CUDA_CHECK(cudaGetLastError());
    if (g_pass_profile && g_pass_profile->n < 16) RS_CUDA_CHECK(cudaEventRecord(g_pass_profile->ev[g_pass_profile->n++], current_stream()));
    return RS_OK;
}

// Plans and launches the passes of one batched transform.
//   forward : row-major passes over the top stages (s0 = 0 upward), then the contiguous pass
//   inverse : the same passes in reverse order; the pass that owns stage 0 folds N^-1 in
//   SHOUP : this translation unit carries the NTT_SHOUP kernels only (ntt_sh_l*.cu; the caller has checked t->d_tw_sh)
//   FID   : modulus compiled into the kernels (ntt.cuh FixedField; the caller has compared t->f->q with it)
template <int NL, int QS, bool SHOUP = false, int FID = FID_RUNTIME>
int ntt_run(const rs_ntt *t, uint32_t *out, const uint32_t *in, size_t n_polys, bool inverse) {
    using T = NttTuning<NL>;
    const int logN = t->logN;
    if (n_polys == 0) return RS_OK;
    if (logN == 0) {   // N = 1: both transforms are the identity (1^-1 = 1)
        if (out != in)
            RS_CUDA_CHECK(cudaMemcpyAsync(out, in, n_polys * NL * 4, cudaMemcpyDeviceToDevice, current_stream()));
        return RS_OK;
    }
    NttConsts<NL> C;
    C.F = make_params<NL>(t->f);
    for (int i = 0; i < NL; i++) { C.ninv[i] = t->ninv[i]; C.ninv_w1[i] = t->ninv_w1[i]; }
    for (int i = 0; i < 2 * NL; i++) { C.ninv_sh[i] = t->ninv_sh[i]; C.ninv_w1_sh[i] = t->ninv_w1_sh[i]; }

    // L2-resident sub-batches (RS_NTT_CHUNK_MB, off by default): a multi-pass transform re-reads what the previous pass wrote;
    // run over sub-batches whose output fits the 126 MB L2, the later passes hit L2 instead of HBM (one HBM round trip per
    // coefficient instead of one per pass).  The kernels are integer bound, so this changes DRAM traffic, not time.
    {
        static const long chunk_mb = [] { const char *e = getenv("RS_NTT_CHUNK_MB"); return e ? atol(e) : 0L; }();
        const size_t poly_bytes = ((size_t)NL * 4) << logN;
        if (chunk_mb > 0 && logN > T::LOG_TS && !g_pass_profile && n_polys * poly_bytes > (size_t)chunk_mb << 20) {
            size_t per = ((size_t)chunk_mb << 20) / poly_bytes;
            if (per < 1) per = 1;
            for (size_t p0 = 0; p0 < n_polys; p0 += per) {
                const size_t n = n_polys - p0 < per ? n_polys - p0 : per;
                const size_t off = (p0 << logN) * NL;
                // (the recursion sees n * poly_bytes <= chunk and runs the passes)
                const int rc = ntt_run<NL, QS, SHOUP, FID>(t, out + off, in + off, n, inverse);
                if (rc) return rc;
            }
            return RS_OK;
        }
    }
    const unsigned long long total = (unsigned long long)n_polys << logN;
    const int KB = logN < T::LOG_TS ? logN : T::LOG_TS;   // stages of the contiguous pass
    const int top = logN - KB;                              // stages left for column passes
    const int KA_MAX = T::LOG_TS - 3;                       // keep >= 8 contiguous columns per row
    const int n_row = top ? (top + KA_MAX - 1) / KA_MAX : 0;

    PassGeom passes[8];
    int np = 0;
    int s = 0;
    for (int r = 0; r < n_row; r++) {
        const int K = (top - s + (n_row - r) - 1) / (n_row - r);   // even split of the remaining stages
        PassGeom G{};
        G.logN = logN; G.K = K; G.s0 = s; G.pl = logN - s - K; G.logW = T::LOG_TS - K; G.col_major = 0;
        G.total_elems = total; G.triv = !t->negacyclic;
        passes[np++] = G;
        s += K;
    }
    {
        PassGeom G{};
        G.logN = logN; G.K = KB; G.s0 = top; G.pl = 0; G.logW = T::LOG_TS - KB; G.col_major = 1;
        G.total_elems = total; G.triv = !t->negacyclic;
        passes[np++] = G;
    }
    const unsigned long long tiles_col = (total + T::TS - 1) / T::TS;
    const unsigned long long tiles_row = total / T::TS;   // logN > LOG_TS here, so it divides

    // lazy [0, 2q) arithmetic when the modulus leaves room (3q < 2^(32 NL)); compiled for <= 8 limbs
    const bool lazy = (NL <= 8) && t->f->lazy_ok;
    const uint32_t *src = in;
    for (int k = 0; k < np; k++) {
        const int p = inverse ? np - 1 - k : k;
        const PassGeom &G = passes[p];
        const unsigned long long tiles = G.col_major ? tiles_col : tiles_row;
        const bool last = (k == np - 1);
        int rc;
#define RS_PASS(INV, LAST, LAZY) launch_pass<NL, QS, INV, LAST, (((NL <= 8) && LAZY) ? NTT_LAZY : NTT_CANON)>(t, out, src, C, G, tiles)
        if constexpr (SHOUP) {
            // (inverse Shoup kernels are instantiated for the compiled-in moduli only: FID > 0)
            if (!inverse) rc = last ? launch_pass<NL, QS, false, true, NTT_SHOUP, FID>(t, out, src, C, G, tiles)
                                    : launch_pass<NL, QS, false, false, NTT_SHOUP, FID>(t, out, src, C, G, tiles);
            else if constexpr (FID != FID_RUNTIME) rc = last ? launch_pass<NL, QS, true, true, NTT_SHOUP, FID>(t, out, src, C, G, tiles)
                                                             : launch_pass<NL, QS, true, false, NTT_SHOUP, FID>(t, out, src, C, G, tiles);
            else return set_error(RS_ERR_UNSUPPORTED, "ntt: no inverse Shoup kernels for this modulus");
        } else if (!inverse) rc = last ? (lazy ? RS_PASS(false, true, true) : RS_PASS(false, true, false))
                                : (lazy ? RS_PASS(false, false, true) : RS_PASS(false, false, false));
        else rc = last ? (lazy ? RS_PASS(true, true, true) : RS_PASS(true, true, false))
                       : (lazy ? RS_PASS(true, false, true) : RS_PASS(true, false, false));
#undef RS_PASS
        if (rc) return rc;
        src = out;
    }
    return RS_OK;
}

}  // namespace rs
