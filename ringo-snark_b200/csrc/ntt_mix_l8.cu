// ntt_mix_l8.cu -- launcher of ntt_mix_pass_kernel (ntt52.cuh): one batched transform of an 8-word field with the
// polynomials of the batch shared between integer-pipe tiles and FP64-pipe tiles inside the same launches.
#include <stdlib.h>

#include "ntt_launch.cuh"

namespace rs {

namespace {
// Tuning variant of the two-pipe kernels.  The unrolled round bodies of BOTH roles have to stay resident in the 32 KB
// L1.5 instruction cache together (radix-8 bodies, 49 KB + 73 KB of SASS, thrashed it: 4.5 no-instruction stall cycles per
// issue), and small CTAs give every SM a fine-grained mix of the two kinds of warps.
template <int RI_, int RF_, int LOG_THREADS_, int LOG_TS_, int MINB_>
struct MixVariant {
    static constexpr int RI = RI_, RF = RF_, LOG_THREADS = LOG_THREADS_, THREADS = 1 << LOG_THREADS_, LOG_TS = LOG_TS_, TS = 1 << LOG_TS_, MINB = MINB_;
};
constexpr int MIN_LOG_W = 2;   // a row of a column tile is >= 4 contiguous elements = one 128-byte line

template <class V, int QS, bool Q0ONE, bool INV, bool LAST, bool LAZY>
int launch_mix(const rs_ntt *t, uint32_t *out, const uint32_t *in, const NttConsts<8> &C, const PassGeom &G, const MixSplit &S,
               unsigned long long blocks) {
    auto kern = ntt_mix_pass_kernel<8, 5, QS, Q0ONE, V::RI, V::RF, V::THREADS, V::TS, V::MINB, INV, LAST, LAZY>;
    constexpr int smem_i = SmemLayout<8, V::TS>::BYTES, smem_f = Smem52<5, V::TS>::BYTES;
    constexpr int smem = smem_i > smem_f ? smem_i : smem_f;
    RS_CUDA_CHECK(ensure_smem_attr((const void *)kern, smem));
    RS_REQUIRE(blocks <= 0x7fffffffull, RS_ERR_UNSUPPORTED, "ntt: batch too large for one launch");
    kern<<<(unsigned)blocks, V::THREADS, smem, current_stream()>>>(out, in, INV ? t->d_twinv : t->d_tw, INV ? t->d_twinv52 : t->d_tw52,
                                                                     C, t->c52, G, S);
    count_launch();
    RS_CUDA_CHECK(cudaGetLastError());
    if (g_pass_profile && g_pass_profile->n < 16) RS_CUDA_CHECK(cudaEventRecord(g_pass_profile->ev[g_pass_profile->n++], current_stream()));
    return RS_OK;
}

template <class V, int QS, bool Q0ONE>
int mix_run(const rs_ntt *t, uint32_t *out, const uint32_t *in, size_t n_polys, bool inverse, size_t n_fp) {
    constexpr int LOG_TS = V::LOG_TS, TS = V::TS;
    const int logN = t->logN;
    NttConsts<8> C;
    C.F = make_params<8>(t->f);
    for (int i = 0; i < 8; i++) { C.ninv[i] = t->ninv[i]; C.ninv_w1[i] = t->ninv_w1[i]; }
    const size_t n_int = n_polys - n_fp;
    MixSplit S;
    S.pf = t->mix_pf; S.p = t->mix_p;
    S.int_elems = (unsigned long long)n_int << logN;
    S.fp_elems = (unsigned long long)n_fp << logN;

    const int KB = logN < LOG_TS ? logN : LOG_TS;
    const int top = logN - KB;
    const int KA_MAX = LOG_TS - MIN_LOG_W;
    const int n_row = top ? (top + KA_MAX - 1) / KA_MAX : 0;
    PassGeom passes[8];
    int np = 0, s = 0;
    for (int r = 0; r < n_row; r++) {
        const int K = (top - s + (n_row - r) - 1) / (n_row - r);
        PassGeom G{};
        G.logN = logN; G.K = K; G.s0 = s; G.pl = logN - s - K; G.logW = LOG_TS - K; G.col_major = 0; G.triv = !t->negacyclic;
        passes[np++] = G;
        s += K;
    }
    {
        PassGeom G{};
        G.logN = logN; G.K = KB; G.s0 = top; G.pl = 0; G.logW = LOG_TS - KB; G.col_major = 1; G.triv = !t->negacyclic;
        passes[np++] = G;
    }
    // tiles per share (the same for every pass: logN > LOG_TS whenever there is a row pass)
    const unsigned long long t_int = (S.int_elems + TS - 1) / TS, t_fp = (S.fp_elems + TS - 1) / TS;
    unsigned long long blocks = 0;
    {
        const unsigned long long a = S.pf ? (t_fp * S.p + S.pf - 1) / S.pf : 0;
        const unsigned long long b = (t_int * S.p + (S.p - S.pf) - 1) / (S.p - S.pf);
        blocks = a > b ? a : b;
        while (blocks * S.pf / S.p < t_fp || blocks - blocks * S.pf / S.p < t_int) blocks++;
    }
    const bool lazy = t->f->lazy_ok;
    const uint32_t *src = in;
    for (int k = 0; k < np; k++) {
        const int p = inverse ? np - 1 - k : k;
        const PassGeom &G = passes[p];
        const bool last = (k == np - 1);
        int rc;
#define RS_MIX(INV, LAST) (lazy ? launch_mix<V, QS, Q0ONE, INV, LAST, true>(t, out, src, C, G, S, blocks) \
                                : launch_mix<V, QS, Q0ONE, INV, LAST, false>(t, out, src, C, G, S, blocks))
        if (!inverse) rc = last ? RS_MIX(false, true) : RS_MIX(false, false);
        else rc = last ? RS_MIX(true, true) : RS_MIX(true, false);
#undef RS_MIX
        if (rc) return rc;
        src = out;
    }
    return RS_OK;
}

template <class V>
int mix_dispatch(const rs_ntt *t, uint32_t *out, const uint32_t *in, size_t n_polys, bool inverse, size_t n_fp) {
    // the shipped 8-word moduli: q == 1 mod 2^64 (zp220, zp255) or q == 1 mod 2^32 only (zp240)
    const int qs = t->f->qs;
    if (t->c52.F.q0one && qs == 2) return mix_run<V, 2, true>(t, out, in, n_polys, inverse, n_fp);
    if (!t->c52.F.q0one && qs == 1) return mix_run<V, 1, false>(t, out, in, n_polys, inverse, n_fp);
    return RS_ERR_UNSUPPORTED;
}
}  // namespace

// RS_ERR_UNSUPPORTED: the caller falls back to the single-pipe kernels (small batches, unsupported modulus shape)
int ntt_mix_run_l8(const rs_ntt *t, uint32_t *out, const uint32_t *in, size_t n_polys, bool inverse) {
    if (!t->has52 || !t->mix_pf || t->logN < 1) return RS_ERR_UNSUPPORTED;
    // the FP64 share: n_polys * pf / p polynomials, at least one of each kind
    size_t n_fp = (n_polys * t->mix_pf + t->mix_p / 2) / t->mix_p;
    if (((unsigned long long)n_polys << t->logN) < (1ull << 10) * 148 * 2 || n_fp == 0 || n_fp == n_polys) return RS_ERR_UNSUPPORTED;
    // radix-4 integer rounds next to radix-2 FP64 rounds, 128 threads, 512-element tiles, 5 CTAs per SM: the best of the four
    // shapes measured (profiles/ntt52_r2c_probe.json); radix-2 / radix-2, radix-4 / radix-4 and 256-thread CTAs were 4-15 % slower
    return mix_dispatch<MixVariant<2, 1, 7, 9, 5>>(t, out, in, n_polys, inverse, n_fp);
}

}  // namespace rs
