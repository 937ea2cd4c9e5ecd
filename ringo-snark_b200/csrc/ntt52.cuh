// ntt52.cuh -- the pass kernel of ntt.cuh with the butterflies on the FP64 pipe (field52.cuh).
//
// Same decomposition, tile geometry and twiddle indexing as ntt_pass_kernel (math/bigpoly/ntt.go:246-466 restated
// as passes of radix-2^R rounds); what changes is the element held in registers and shared memory: ND signed
// 64-bit limbs of 52 bits, lazily reduced, instead of NL canonical 32-bit words.
//
// Bound bookkeeping (values are representatives mod q; "B q" = the value is below B q; max_lazy = 2^(52 ND) / q):
//   * a tile is loaded with values below 2.0001 q (canonical input, or the previous pass's f52_reduce output);
//   * mont52 accepts any normalised a < 2^(52 ND) and returns less than 2q;
//   * forward (Cooley-Tukey, ntt.go:254-259): u' = u + v, v' = u - v + 2q with v fresh from mont52, so the bound
//     grows by 2q per stage: at most 2.0001 + 2 * 11 = 24.0001 q in a pass.  In the opening round of a cyclic
//     transform the first twiddle of every stage is 1 (ntt.go:56-84) and v is the unmultiplied element: offsets
//     4q / 6q / 10q keep the difference non-negative and the round ends below 18.0004 q (then + 2q per stage);
//   * inverse (Gentleman-Sande, ntt.go:365-370): u' = u + v doubles the bound per stage, d = u - v + off is
//     multiplied at once.  A round of <= 3 stages starts below 2.0001 q, so its inputs stay below 8.0004 q,
//     off = 4q / 6q / 10q by stage covers v, d < 18.0004 q, and the round ends by reducing the elements that were
//     not multiplied in its last stage.  The closing round of a cyclic transform (unit first twiddles, N^-1 folded
//     into its last stage) leaves some differences unmultiplied: off = 14q in its last stage, everything < 26.0002 q;
//   * the host enables this path only when max_lazy >= 36 (f52_setup).
// Limb magnitudes stay below 2^62: mont52 columns are below 2^58 and at most 11 of them accumulate in a limb.
#pragma once
#include "field52.cuh"
#include "ntt.cuh"

namespace rs {

constexpr int TW52_STRIDE = 6;   // doubles per twiddle-table entry (ND = 5 limbs + 1 pad: 48 bytes, 16-byte aligned)

template <int NL, int ND>
struct Ntt52Consts {
    Field52<ND> F;
    double ninv[ND];       // N^-1            * 2^(52 ND) mod q
    double ninv_w1[ND];    // N^-1 * twinv[1] * 2^(52 ND) mod q
    uint32_t qw[NL];       // q as words (final canonicalisation)
};

template <int ND>
RS_HD void load_tw52(double (&w)[ND], const double *tw, unsigned long long idx) {
    static_assert(ND <= TW52_STRIDE, "twiddle stride");
    const double *src = tw + idx * TW52_STRIDE;
#if defined(__CUDA_ARCH__)
#pragma unroll
    for (int h = 0; h < (ND + 1) / 2; h++) {
        const double2 v = __ldg(reinterpret_cast<const double2 *>(src) + h);
        w[2 * h] = v.x;
        if (2 * h + 1 < ND) w[2 * h + 1] = v.y;
    }
#else
    for (int k = 0; k < ND; k++) w[k] = src[k];
#endif
}

template <int ND>
RS_HD void f52_mul_inplace(long long (&x)[ND], const double (&w)[ND], const Field52<ND> &F, bool q0one) {
    f52_normalize<ND>(x);
    double a[ND];
    f52_to_doubles<ND>(a, x);
    if (q0one) mont52<ND, true>(x, a, w, F); else mont52<ND, false>(x, a, w, F);
}

// One radix-2^RR round on 2^RR register-resident lazy elements (see radix_round in ntt.cuh for the index algebra).
template <int ND, bool Q0ONE, int RR, bool INV, bool SCALE, bool TRIV, class C52>
RS_HD void radix_round52(long long (&x)[1 << RR][ND], const double *tw, int s_top, unsigned hi, const C52 &C) {
    const Field52<ND> &F = C.F;
#pragma unroll
    for (int bb = 0; bb < RR; bb++) {
        const int beta = INV ? bb : (RR - 1 - bb);
        const int s = s_top + (RR - 1 - beta);
        const unsigned long long tbase = (1ull << s) + ((unsigned long long)hi << (RR - 1 - beta));
#pragma unroll
        for (int g = 0; g < (1 << (RR - 1 - beta)); g++) {
            double w[ND];
            const bool last_scaled = INV && SCALE && (beta == RR - 1);
            const bool triv = TRIV && (g == 0);
            if (!last_scaled && !triv) load_tw52<ND>(w, tw, tbase + g);
#pragma unroll
            for (int k = 0; k < (1 << beta); k++) {
                const int u0 = (g << (beta + 1)) + k, u1 = u0 + (1 << beta);
                if (!INV) {
                    const int oi = triv ? (bb + 1) : 0;                  // 4q / 6q / 10q for unit twiddles, else 2q
                    if (!triv) f52_mul_inplace<ND>(x[u1], w, F, Q0ONE);
#pragma unroll
                    for (int l = 0; l < ND; l++) {
                        const long long v = x[u1][l];
                        x[u1][l] = x[u0][l] - v + F.off[oi][l];
                        x[u0][l] += v;
                    }
                } else {
                    const int oi = (TRIV && bb == RR - 1) ? 4 : (bb + 1);   // 4q / 6q / 10q by stage; 14q closing a TRIV round
#pragma unroll
                    for (int l = 0; l < ND; l++) {
                        const long long v = x[u1][l];
                        x[u1][l] = x[u0][l] - v + F.off[oi][l];
                        x[u0][l] += v;
                    }
                    if (last_scaled) {
                        f52_mul_inplace<ND>(x[u0], C.ninv, F, Q0ONE);
                        f52_mul_inplace<ND>(x[u1], C.ninv_w1, F, Q0ONE);
                    } else if (!triv) {
                        f52_mul_inplace<ND>(x[u1], w, F, Q0ONE);
                    }
                }
            }
        }
    }
    if (INV && !SCALE) {
        // elements that were not multiplied in the last stage carry sums: back below 2.0001 q
#pragma unroll
        for (int u = 0; u < (1 << (RR - 1)); u++) f52_reduce<ND>(x[u], F);
    }
}

#ifdef __CUDACC__

// shared memory: ND planes of 8 bytes, one pad slot per 8 elements (conflict-free LDS.64 / STS.64 in every round)
template <int ND, int TS>
struct Smem52 {
    static constexpr int SLOTS = TS + TS / 8;
    static constexpr int PLANE_STRIDE = SLOTS * 2 + 16;     // words
    static constexpr int BYTES = PLANE_STRIDE * ND * 4;
};

template <int ND, int TS>
__device__ __forceinline__ void smem52_load(long long (&x)[ND], const uint32_t *sm, unsigned idx) {
    const unsigned p = pad_idx(idx);
#pragma unroll
    for (int h = 0; h < ND; h++)
        x[h] = *reinterpret_cast<const long long *>(sm + h * Smem52<ND, TS>::PLANE_STRIDE + p * 2);
}
template <int ND, int TS>
__device__ __forceinline__ void smem52_store(uint32_t *sm, unsigned idx, const long long (&x)[ND]) {
    const unsigned p = pad_idx(idx);
#pragma unroll
    for (int h = 0; h < ND; h++)
        *reinterpret_cast<long long *>(sm + h * Smem52<ND, TS>::PLANE_STRIDE + p * 2) = x[h];
}

template <int ND, bool Q0ONE, int RR, bool INV, bool SCALE, int THREADS, int TS, bool TRIV, class C52>
__device__ __forceinline__ void run_round52(uint32_t *sm, const double *__restrict__ tw, const PassGeom &G, int lb,
                                            unsigned long long tile, const C52 &C) {
    const int K = G.K;
    const int lgu = lb + (G.col_major ? 0 : G.logW);
    const unsigned ustride = 1u << lgu;
    const int hi_bits = K - lb - RR;
    const int s_top = G.s0 + (K - lb - RR);
    constexpr unsigned NGROUPS = TS >> RR;
    for (unsigned gid = threadIdx.x; gid < NGROUPS; gid += THREADS) {
        const unsigned qd = gid >> lgu, rem = gid & (ustride - 1);
        const unsigned base = (qd << (lgu + RR)) + rem;
        unsigned hi;
        bool valid = true;
        if (G.col_major) {
            const unsigned cl = qd >> hi_bits;
            const unsigned i_hi = qd & ((1u << hi_bits) - 1);
            const unsigned long long chunk = tile * (unsigned long long)(TS >> K) + cl;
            valid = (chunk << K) < G.total_elems;
            const unsigned a = (unsigned)(chunk & ((1ull << (G.logN - K)) - 1));
            hi = (a << hi_bits) + i_hi;
        } else {
            const unsigned tpp = 1u << (G.logN - K - G.logW);
            const unsigned c0 = (unsigned)(tile % tpp) << G.logW;
            const unsigned a = c0 >> G.pl;
            hi = (a << hi_bits) + qd;
        }
        if (!valid) continue;
        long long x[1 << RR][ND];
#pragma unroll
        for (int u = 0; u < (1 << RR); u++) smem52_load<ND, TS>(x[u], sm, base + u * ustride);
        radix_round52<ND, Q0ONE, RR, INV, SCALE, TRIV>(x, tw, s_top, hi, C);
#pragma unroll
        for (int u = 0; u < (1 << RR); u++) smem52_store<ND, TS>(sm, base + u * ustride, x[u]);
    }
}

// global <-> shared: one thread moves one whole element (256-bit access), converting words <-> limbs on the way.
// Store: lazy value -> below 2.0001 q (f52_reduce) -> words -> (LAST) two conditional subtractions -> [0, q).
template <int NL, int ND, int THREADS, int TS, bool STORE, bool LAST, class C52>
__device__ __forceinline__ void move_tile52(uint32_t *sm, const uint32_t *gin, uint32_t *gout, const PassGeom &G,
                                            unsigned long long tile, const C52 &C) {
    static_assert(NL == 8, "256-bit element access");
    const unsigned tpp = G.col_major ? 1u : (1u << (G.logN - G.K - G.logW));
    const unsigned long long poly = G.col_major ? 0ull : tile / tpp;
    const unsigned c0 = G.col_major ? 0u : ((unsigned)(tile % tpp) << G.logW);
    const unsigned long long a_sh = G.col_major ? 0ull : ((unsigned long long)(c0 >> G.pl) << (G.pl + G.K));
    const unsigned b0 = G.col_major ? 0u : (c0 & ((1u << G.pl) - 1));
    for (unsigned e = threadIdx.x; e < (unsigned)TS; e += THREADS) {
        unsigned long long flat;
        if (G.col_major) {
            flat = tile * (unsigned long long)TS + e;
        } else {
            const unsigned cl = e & ((1u << G.logW) - 1), i = e >> G.logW;
            flat = (poly << G.logN) + a_sh + ((unsigned long long)i << G.pl) + b0 + cl;
        }
        if (flat >= G.total_elems) continue;
        uint32_t w[NL];
        long long l[ND];
        if (STORE) {
            smem52_load<ND, TS>(l, sm, e);
            f52_reduce<ND>(l, C.F);
            f52_to_words<NL, ND>(w, l);
            if (LAST) { cond_sub<NL>(w, C.qw); cond_sub<NL>(w, C.qw); }
            const unsigned long long a0 = w[0] | ((unsigned long long)w[1] << 32), a1 = w[2] | ((unsigned long long)w[3] << 32),
                                     a2 = w[4] | ((unsigned long long)w[5] << 32), a3 = w[6] | ((unsigned long long)w[7] << 32);
            asm volatile("st.global.v4.u64 [%0], {%1, %2, %3, %4};" ::"l"(gout + flat * NL), "l"(a0), "l"(a1), "l"(a2), "l"(a3) : "memory");
        } else {
            unsigned long long a0, a1, a2, a3;
            asm volatile("ld.global.v4.u64 {%0, %1, %2, %3}, [%4];" : "=l"(a0), "=l"(a1), "=l"(a2), "=l"(a3) : "l"(gin + flat * NL));
            w[0] = (uint32_t)a0; w[1] = (uint32_t)(a0 >> 32); w[2] = (uint32_t)a1; w[3] = (uint32_t)(a1 >> 32);
            w[4] = (uint32_t)a2; w[5] = (uint32_t)(a2 >> 32); w[6] = (uint32_t)a3; w[7] = (uint32_t)(a3 >> 32);
            f52_from_words<NL, ND>(l, w);
            smem52_store<ND, TS>(sm, e, l);
        }
    }
}

// LAST: this pass ends the transform (forward: owns stage logN-1; inverse: owns stage 0 and folds N^-1 in)
template <int NL, int ND, bool Q0ONE, int R, int THREADS, int TS, bool INV, bool LAST, class SYNC>
__device__ __forceinline__ void ntt52_pass_body(uint32_t *sm, uint32_t *out, const uint32_t *in, const double *__restrict__ tw,
                                                const Ntt52Consts<NL, ND> &C, const PassGeom &G, unsigned long long tile, SYNC sync) {
    constexpr bool SCALE = INV && LAST;
    static_assert(R <= 3, "offset table covers rounds of at most 3 stages");
    move_tile52<NL, ND, THREADS, TS, false, false>(sm, in, out, G, tile, C);
    sync();

    const int K = G.K;
    const int full = K / R, rem = K % R;
    for (int r = 0; r < full; r++) {
        const int lb = INV ? r * R : K - (r + 1) * R;
        if (SCALE && rem == 0 && r == full - 1) {
            if (G.triv) run_round52<ND, Q0ONE, R, INV, SCALE, THREADS, TS, SCALE>(sm, tw, G, lb, tile, C);
            else run_round52<ND, Q0ONE, R, INV, SCALE, THREADS, TS, false>(sm, tw, G, lb, tile, C);
        } else if (!INV && r == 0 && G.s0 == 0 && G.triv) run_round52<ND, Q0ONE, R, false, false, THREADS, TS, !INV>(sm, tw, G, lb, tile, C);
        else run_round52<ND, Q0ONE, R, INV, false, THREADS, TS, false>(sm, tw, G, lb, tile, C);
        sync();
    }
    if (rem) {
        const int lb = INV ? full * R : 0;
        if (R >= 2 && rem == 1) run_round52<ND, Q0ONE, 1, INV, SCALE, THREADS, TS, false>(sm, tw, G, lb, tile, C);
        if (R >= 3 && rem == 2) run_round52<ND, Q0ONE, (R >= 3 ? 2 : 1), INV, SCALE, THREADS, TS, false>(sm, tw, G, lb, tile, C);
        sync();
    }
    move_tile52<NL, ND, THREADS, TS, true, LAST>(sm, in, out, G, tile, C);
}

template <int NL, int ND, bool Q0ONE, int R, int THREADS, int MINB, bool INV, bool LAST>
__global__ void __launch_bounds__(THREADS, MINB) ntt52_pass_kernel(uint32_t *out, const uint32_t *in, const double *__restrict__ tw,
                                                                   const __grid_constant__ Ntt52Consts<NL, ND> C,
                                                                   const __grid_constant__ PassGeom G) {
    extern __shared__ __align__(16) uint32_t sm[];
    ntt52_pass_body<NL, ND, Q0ONE, R, THREADS, (THREADS << R), INV, LAST>(sm, out, in, tw, C, G, blockIdx.x, CtaSync());
}

// ---- both pipes at once ---------------------------------------------------------------------------------------
// A DFMA / DADD holds the issue port of its sub-partition for two cycles (measured: DFMA next to LOP3 or IMAD 1:1
// runs at 2 instructions per 3 cycles, tools/imad_peaks.py modes 10 / 11), so a kernel made of FP64 products alone is
// issue bound and no faster than the IMAD.WIDE kernels, whose limit is the FMA pipe (one IMAD.WIDE per 4 cycles) with
// more than half of the issue slots idle.  The two are complementary: this kernel gives some CTAs integer tiles and
// the others FP64 tiles (disjoint polynomials of the batch, same pass geometry), so every SM holds warps of both
// kinds and the warp scheduler fills the integer warps' idle issue slots with DFMA work.
struct MixSplit {
    unsigned pf, p;                   // of every p consecutive CTAs, pf take FP64 tiles (Bresenham pattern)
    unsigned long long int_elems;     // elements (polys * N) of the integer share: polys [0, n_int)
    unsigned long long fp_elems;      // elements of the FP64 share: polys [n_int, n)
};

// RI / RF: radix-2^R rounds of the integer / FP64 role; TS: tile size of both (elements)
template <int NL, int ND, int QS, bool Q0ONE, int RI, int RF, int THREADS, int TS, int MINB, bool INV, bool LAST, bool LAZY>
__global__ void __launch_bounds__(THREADS, MINB) ntt_mix_pass_kernel(uint32_t *out, const uint32_t *in, const uint32_t *__restrict__ tw,
                                                                     const double *__restrict__ tw52,
                                                                     const __grid_constant__ NttConsts<NL> C,
                                                                     const __grid_constant__ Ntt52Consts<NL, ND> C52,
                                                                     const __grid_constant__ PassGeom G0,
                                                                     const __grid_constant__ MixSplit S) {
    extern __shared__ __align__(16) uint32_t sm[];
    const unsigned long long b = blockIdx.x;
    const unsigned long long fp_before = b * S.pf / S.p, fp_after = (b + 1) * S.pf / S.p;
    const bool fp_role = fp_after != fp_before;
    PassGeom G = G0;
    if (fp_role) {
        G.total_elems = S.fp_elems;
        const unsigned long long tile = fp_before;
        if (tile * TS >= (G.col_major ? ((S.fp_elems + TS - 1) / TS) * TS : S.fp_elems)) return;
        ntt52_pass_body<NL, ND, Q0ONE, RF, THREADS, TS, INV, LAST>(sm, out + S.int_elems * NL, in + S.int_elems * NL, tw52, C52, G, tile, CtaSync());
    } else {
        G.total_elems = S.int_elems;
        const unsigned long long tile = b - fp_before;
        if (tile * TS >= (G.col_major ? ((S.int_elems + TS - 1) / TS) * TS : S.int_elems)) return;
        ntt_pass_body<NL, QS, RI, THREADS, TS, INV, (INV && LAST), LAZY, LAST>(sm, out, in, tw, C, G, tile, CtaSync());
    }
}

#endif  // __CUDACC__
}  // namespace rs
