// ntt.cuh -- batched big-field NTT for sm_100a.
//
// Replaces math/bigpoly/ntt.go: nttInPlace(Unroll) :246-355, inttInPlace(Unroll) :357-466,
// FwdNTTTo/InvNTTTo :98-136 (cyclic) and :206-244 (negacyclic).  Same twiddle tables
// (tw[m+i], bit-reversed powers, Montgomery form; ntt.go:56-84, :183-192), same radix-2
// butterflies (:254-259, :365-370), so outputs are bit-identical: forward output in
// bit-reversed order, inverse ends with * N^-1 (folded into the last stage here).
//
// Decomposition: the log2 N stages are split into "passes"; a pass owns K consecutive
// stages and works on tiles of TS = THREADS * 2^R elements staged in shared memory:
//   * column tiles (row-major, early stages): 2^K rows x W contiguous columns, rows
//     strided in HBM by 2^pl elements, twiddles shared by all columns;
//   * contiguous tiles (col-major, last K stages): W consecutive chunks of 2^K elements.
// Inside a pass, every thread keeps 2^R elements in registers and runs R stages
// (radix-2^R round) between shared-memory exchanges.  Shared memory holds 16-byte (or
// 8-byte) "planes" of the elements with one pad slot per 8 elements, so that every LDS/STS
// of a round is conflict-free.
#pragma once
#include "field.cuh"
#include "shoup.cuh"

namespace rs {

template <int NL>
struct NttConsts {
    FieldParams<NL> F;
    uint32_t ninv[NL];      // N^-1            (Montgomery)   ntt.go:86-87
    uint32_t ninv_w1[NL];   // N^-1 * twinv[1] (Montgomery)
    // Shoup pairs (plain value, floor(value 2^(32 NL) / q)) of the two constants above (NTT_SHOUP inverse kernels, shoup.cuh)
    uint32_t ninv_sh[2 * NL];
    uint32_t ninv_w1_sh[2 * NL];
};

// arithmetic of a pass: canonical [0, q) Montgomery; lazy [0, 2q) Montgomery (field.cuh); lazy with Shoup products by the
// (w, w') twiddle pairs and predicated corrections (shoup.cuh; the table has 2 NL words per entry)
enum { NTT_CANON = 0, NTT_LAZY = 1, NTT_SHOUP = 2 };

// Moduli compiled into their kernels (FID > 0), as in the reference, whose zp packages are generated per modulus with q as
// constants (jindo/internal/zp/element.go:32, :78): the limbs of q and 2q become immediates of the multiply-adds and of the
// predicated corrections (with q a kernel parameter ptxas reloads them under the predicate: 13 LDC per butterfly).
// FID 0: q is the kernel parameter.  The launcher compares the field's q with the table before it picks FID > 0.
enum { FID_RUNTIME = 0, FID_ZP255 = 1 };
template <int NL, int FID> struct FixedField;
template <> struct FixedField<8, FID_ZP255> {
    // q = 0x430d45996b62afc2d65643d9e6fb65558e9630dc8c3732810000000000000001 (jindo/internal/zp/element.go:32)
    RS_HD static constexpr uint32_t q(int i) {
        constexpr uint32_t v[8] = {0x00000001u, 0x00000000u, 0x8c373281u, 0x8e9630dcu, 0xe6fb6555u, 0xd65643d9u, 0x6b62afc2u, 0x430d4599u};
        return v[i];
    }
    RS_HD static constexpr uint32_t q2(int i) {
        constexpr uint32_t v[8] = {0x00000002u, 0x00000000u, 0x186e6502u, 0x1d2c61b9u, 0xcdf6caabu, 0xacac87b3u, 0xd6c55f85u, 0x861a8b32u};
        return v[i];
    }
};
template <int NL, int FID>
RS_HD bool fixed_field_matches(const uint32_t *q) {
    if constexpr (FID == FID_RUNTIME) return true;
    else {
        for (int i = 0; i < NL; i++) if (q[i] != FixedField<NL, FID>::q(i)) return false;
        return true;
    }
}
// the field a kernel computes in: the parameter block, or a local copy whose q / 2q are the compiled-in constants
// (every access is at a compile-time index after unrolling, so the copy never materialises)
#ifdef __CUDACC__
template <int NL, int FID>
__device__ __forceinline__ FieldParams<NL> kernel_field(const FieldParams<NL> &P) {
    if constexpr (FID == FID_RUNTIME) return P;
    else {
        FieldParams<NL> F = P;
#pragma unroll
        for (int i = 0; i < NL; i++) { F.q[i] = FixedField<NL, FID>::q(i); F.q2[i] = FixedField<NL, FID>::q2(i); }
        return F;
    }
}
#endif

struct PassGeom {
    int logN;        // log2 of the transform size
    int K;           // stages in this pass
    int s0;          // first stage (m = 2^s0) of this pass
    int pl;          // bit position of the pass's lowest butterfly stride (logN - s0 - K)
    int logW;        // log2 columns per tile
    int col_major;   // 1: contiguous tiles (pl == 0)
    int triv;        // 1: tw[2^s] == 1 for every stage (cyclic tables, ntt.go:56-84) -> stage-0 rounds skip those products
    unsigned long long total_elems;   // n_polys * N
};

template <int NL> struct Planes {
    static constexpr int BYTES = (NL % 4 == 0) ? 16 : 8;
    static constexpr int WORDS = BYTES / 4;
    static constexpr int COUNT = NL / WORDS;
};

RS_HD unsigned pad_idx(unsigned idx) { return idx + (idx >> 3); }

template <int NL, int TS>
struct SmemLayout {
    static constexpr int PW = Planes<NL>::WORDS;
    static constexpr int PC = Planes<NL>::COUNT;
    static constexpr int SLOTS = TS + TS / 8;
    // plane stride in words; +16 words (64 B) so that consecutive planes start 16 banks apart
    static constexpr int PLANE_STRIDE = SLOTS * PW + 16;
    static constexpr int BYTES = PLANE_STRIDE * PC * 4;
};

#ifdef __CUDACC__

template <int NL, int TS>
__device__ __forceinline__ void smem_load_elem(uint32_t *x, const uint32_t *sm, unsigned idx) {
    using L = SmemLayout<NL, TS>;
    const unsigned p = pad_idx(idx);
#pragma unroll
    for (int h = 0; h < L::PC; h++) {
        const uint32_t *src = sm + h * L::PLANE_STRIDE + p * L::PW;
        if (L::PW == 4) {
            uint4 v = *reinterpret_cast<const uint4 *>(src);
            x[4 * h] = v.x; x[4 * h + 1] = v.y; x[4 * h + 2] = v.z; x[4 * h + 3] = v.w;
        } else {
            uint2 v = *reinterpret_cast<const uint2 *>(src);
            x[2 * h] = v.x; x[2 * h + 1] = v.y;
        }
    }
}

template <int NL, int TS>
__device__ __forceinline__ void smem_store_elem(uint32_t *sm, unsigned idx, const uint32_t *x) {
    using L = SmemLayout<NL, TS>;
    const unsigned p = pad_idx(idx);
#pragma unroll
    for (int h = 0; h < L::PC; h++) {
        uint32_t *dst = sm + h * L::PLANE_STRIDE + p * L::PW;
        if (L::PW == 4) *reinterpret_cast<uint4 *>(dst) = make_uint4(x[4 * h], x[4 * h + 1], x[4 * h + 2], x[4 * h + 3]);
        else *reinterpret_cast<uint2 *>(dst) = make_uint2(x[2 * h], x[2 * h + 1]);
    }
}

template <int NL>
__device__ __forceinline__ void load_twiddle(uint32_t *w, const uint32_t *__restrict__ tw, unsigned long long idx) {
    constexpr int PW = Planes<NL>::WORDS;
    const uint32_t *src = tw + idx * NL;
#pragma unroll
    for (int h = 0; h < NL / PW; h++) {
        if (PW == 4) {
            uint4 v = __ldg(reinterpret_cast<const uint4 *>(src) + h);
            w[4 * h] = v.x; w[4 * h + 1] = v.y; w[4 * h + 2] = v.z; w[4 * h + 3] = v.w;
        } else {
            uint2 v = __ldg(reinterpret_cast<const uint2 *>(src) + h);
            w[2 * h] = v.x; w[2 * h + 1] = v.y;
        }
    }
}

// (w, w') pair of the Shoup table
template <int NL>
__device__ __forceinline__ void load_twiddle_pair(uint32_t *w, uint32_t *wp, const uint32_t *__restrict__ tw, unsigned long long idx) {
    load_twiddle<NL>(w, tw, 2 * idx);
    load_twiddle<NL>(wp, tw, 2 * idx + 1);
}

// One radix-2^RR round on the 2^RR register-resident elements x[u][*].
//   s_top : global stage index of the round's highest local bit (u-bit RR-1)
//   hi    : index of this group among the groups of the round inside its polynomial
//   SCALE : (inverse only) the round's last stage is stage 0 of the transform: fold N^-1 in
//   MODE  : NTT_LAZY / NTT_SHOUP: values are kept in [0, 2q) resp. [0, 2q + delta) (field.cuh, shoup.cuh); the caller canonicalises at the end
//   TRIV  : the round starts at stage 0 of the transform (s_top == 0, hi == 0): the first twiddle of every
//           stage is tw[2^s] = 1 (Montgomery one), so those products are skipped -- 7 of the 12 in a
//           radix-8 round.  x * R * R^-1 = x exactly, so the values are the ones the reference computes.
template <int NL, int QS, int RR, bool INV, bool SCALE, int MODE, bool TRIV = false, int FID = FID_RUNTIME>
__device__ __forceinline__ void radix_round(uint32_t (&x)[1 << RR][NL], const uint32_t *__restrict__ tw,
                                            int s_top, unsigned hi, const NttConsts<NL> &C) {
    constexpr bool LAZY = MODE >= NTT_LAZY, SH = MODE == NTT_SHOUP;
    const FieldParams<NL> F = kernel_field<NL, FID>(C.F);
#pragma unroll
    for (int bb = 0; bb < RR; bb++) {
        const int beta = INV ? bb : (RR - 1 - bb);         // u-bit handled by this stage
        const int s = s_top + (RR - 1 - beta);             // global stage index (m = 2^s)
        const unsigned long long tbase = (1ull << s) + ((unsigned long long)hi << (RR - 1 - beta));
#pragma unroll
        for (int g = 0; g < (1 << (RR - 1 - beta)); g++) {
            uint32_t w[NL], wp[SH ? NL : 1];
            constexpr bool SC = INV && SCALE;
            const bool last_scaled = SC && (beta == RR - 1);   // compile-time after unrolling
            const bool triv = TRIV && (g == 0);                 // compile-time after unrolling
            if (!last_scaled && !triv) {
                if constexpr (SH) load_twiddle_pair<NL>(w, wp, tw, tbase + g); else load_twiddle<NL>(w, tw, tbase + g);
            }
#pragma unroll
            for (int k = 0; k < (1 << beta); k++) {
                const int u0 = (g << (beta + 1)) + k, u1 = u0 + (1 << beta);
                if (!INV) {
                    // ntt.go:254-259   v = v*w ; u' = u+v ; v' = u-v
                    uint32_t v[NL];
                    if (triv) {
#pragma unroll
                        for (int l = 0; l < NL; l++) v[l] = x[u1][l];
                    }
                    bool done = false;
                    if constexpr (SH) {
                        if (!triv) {
                            // v < 1.56 q: u + v cannot wrap and u - v + 2q cannot be negative (shoup.cuh)
                            shoup_mul_lazy<NL, QS>(v, x[u1], w, wp, F);
                            lazy_sub_p<NL>(x[u1], x[u0], v, F.q2);
                            lazy_add_p<NL>(x[u0], x[u0], v, F.q2);
                            done = true;
                        }
                    }
                    if (done) {
                    } else if (LAZY) {
                        // (MODE 2: the trivial-twiddle butterflies only ever see the exact [0, 2q) outputs of other trivial ones)
                        if (!triv) mont_mul_lazy<NL, QS>(v, x[u1], w, F);
                        mod_sub_m<NL>(x[u1], x[u0], v, F.q2);
                        mod_add_m<NL>(x[u0], x[u0], v, F.q2);
                    } else {
                        if (!triv) mont_mul<NL, QS>(v, x[u1], w, F);
                        mod_sub<NL>(x[u1], x[u0], v, F);
                        mod_add<NL>(x[u0], x[u0], v, F);
                    }
                } else {
                    // ntt.go:365-370   u' = u+v ; v' = (u-v)*w
                    uint32_t d[NL];
                    if constexpr (SH) {
                        // the subtrahend is an arbitrary value: sums keep the exact reduction to [0, 2q)
                        lazy_sub_p<NL>(d, x[u0], x[u1], F.q2);
                        mod_add_m<NL>(x[u0], x[u0], x[u1], F.q2);
                    } else if (LAZY) {
                        mod_sub_m<NL>(d, x[u0], x[u1], F.q2);
                        mod_add_m<NL>(x[u0], x[u0], x[u1], F.q2);
                    } else {
                        mod_sub<NL>(d, x[u0], x[u1], F);
                        mod_add<NL>(x[u0], x[u0], x[u1], F);
                    }
                    if (last_scaled) {
                        // stage m = 1 with the trailing scalarMulVecTo(N^-1) (ntt.go:134-135, :242-243) folded in
                        uint32_t t[NL];
                        if constexpr (SH) shoup_mul_lazy<NL, QS>(t, x[u0], C.ninv_sh, C.ninv_sh + NL, F);
                        else if (LAZY) mont_mul_lazy<NL, QS>(t, x[u0], C.ninv, F);
                        else mont_mul<NL, QS>(t, x[u0], C.ninv, F);
#pragma unroll
                        for (int l = 0; l < NL; l++) x[u0][l] = t[l];
                        if constexpr (SH) shoup_mul_lazy<NL, QS>(x[u1], d, C.ninv_w1_sh, C.ninv_w1_sh + NL, F);
                        else if (LAZY) mont_mul_lazy<NL, QS>(x[u1], d, C.ninv_w1, F);
                        else mont_mul<NL, QS>(x[u1], d, C.ninv_w1, F);
                    } else if (triv) {
#pragma unroll
                        for (int l = 0; l < NL; l++) x[u1][l] = d[l];
                    } else {
                        if constexpr (SH) shoup_mul_lazy<NL, QS>(x[u1], d, w, wp, F);
                        else if (LAZY) mont_mul_lazy<NL, QS>(x[u1], d, w, F);
                        else mont_mul<NL, QS>(x[u1], d, w, F);
                    }
                }
            }
        }
    }
}

// rounds of one pass over the tile in shared memory
//   FINAL : (LAZY only) this round contains the last stage of the transform: bring values back to [0, q)
template <int NL, int QS, int RR, bool INV, bool SCALE, int MODE, bool FINAL, int THREADS, int TS, bool TRIV = false, int FID = FID_RUNTIME>
__device__ __forceinline__ void run_round(uint32_t *sm, const uint32_t *__restrict__ tw, const PassGeom &G,
                                          int lb, unsigned long long tile, const NttConsts<NL> &C) {
    constexpr bool LAZY = MODE >= NTT_LAZY;
    // group geometry (see header comment): element u of group gid sits at base + u * ustride
    const int K = G.K;
    const int lgu = lb + (G.col_major ? 0 : G.logW);
    const unsigned ustride = 1u << lgu;
    const int hi_bits = K - lb - RR;                       // bits of i above the round's field
    const int s_top = G.s0 + (K - lb - RR);                // stage of the round's highest bit
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
            const unsigned tpp = 1u << (G.logN - K - G.logW);              // tiles per polynomial
            const unsigned c0 = (unsigned)(tile % tpp) << G.logW;
            const unsigned a = c0 >> G.pl;
            hi = (a << hi_bits) + qd;
        }
        if (!valid) continue;
        uint32_t x[1 << RR][NL];
#pragma unroll
        for (int u = 0; u < (1 << RR); u++) smem_load_elem<NL, TS>(x[u], sm, base + u * ustride);
        radix_round<NL, QS, RR, INV, SCALE, MODE, TRIV, FID>(x, tw, s_top, hi, C);
        const FieldParams<NL> F = kernel_field<NL, FID>(C.F);
#pragma unroll
        for (int u = 0; u < (1 << RR); u++) {
            // (Shoup kernels, forward: sums may sit up to delta above 2q; the inverse ends with products, below 1.56 q)
            if (MODE == NTT_SHOUP && FINAL && !INV) cond_sub<NL>(x[u], F.q2);
            if (LAZY && FINAL) cond_sub<NL>(x[u], F.q);
            smem_store_elem<NL, TS>(sm, base + u * ustride, x[u]);
        }
    }
}

// global <-> shared tile movement in plane-sized pieces, fully coalesced
template <int NL, int THREADS, int TS, bool STORE>
__device__ __forceinline__ void move_tile(uint32_t *sm, const uint32_t *gin, uint32_t *gout, const PassGeom &G,
                                          unsigned long long tile) {
    using L = SmemLayout<NL, TS>;
    constexpr int PC = L::PC, PW = L::PW;
    const unsigned tpp = G.col_major ? 1u : (1u << (G.logN - G.K - G.logW));
    const unsigned long long poly = G.col_major ? 0ull : tile / tpp;
    const unsigned c0 = G.col_major ? 0u : ((unsigned)(tile % tpp) << G.logW);
    const unsigned long long a_sh = G.col_major ? 0ull : ((unsigned long long)(c0 >> G.pl) << (G.pl + G.K));
    const unsigned b0 = G.col_major ? 0u : (c0 & ((1u << G.pl) - 1));
    for (unsigned p = threadIdx.x; p < (unsigned)(TS * PC); p += THREADS) {
        const unsigned e = p / PC, h = p % PC;
        unsigned long long flat;
        if (G.col_major) {
            flat = tile * (unsigned long long)TS + e;
        } else {
            const unsigned cl = e & ((1u << G.logW) - 1), i = e >> G.logW;
            flat = (poly << G.logN) + a_sh + ((unsigned long long)i << G.pl) + b0 + cl;
        }
        if (flat >= G.total_elems) continue;
        uint32_t *s = sm + h * L::PLANE_STRIDE + pad_idx(e) * PW;
        if (PW == 4) {
            if (STORE) reinterpret_cast<uint4 *>(gout + flat * NL)[h] = *reinterpret_cast<const uint4 *>(s);
            else *reinterpret_cast<uint4 *>(s) = reinterpret_cast<const uint4 *>(gin + flat * NL)[h];
        } else {
            if (STORE) reinterpret_cast<uint2 *>(gout + flat * NL)[h] = *reinterpret_cast<const uint2 *>(s);
            else *reinterpret_cast<uint2 *>(s) = reinterpret_cast<const uint2 *>(gin + flat * NL)[h];
        }
    }
}

// LAST: this pass ends the transform (forward: owns stage logN-1; inverse: owns stage 0, == SCALE)
// (in == out for every pass after the first: plain loads on the data, no __restrict__)
// TS: tile size in elements (THREADS << R in the single-pipe kernels: one group of 2^R elements per thread and round;
// a larger TS makes every thread walk several groups per round)
template <int NL, int QS, int R, int THREADS, int TS, bool INV, bool SCALE, int LAZY, bool LAST, class SYNC, int FID = FID_RUNTIME>
__device__ __forceinline__ void ntt_pass_body(uint32_t *sm, uint32_t *out, const uint32_t *in, const uint32_t *__restrict__ tw,
                                              const NttConsts<NL> &C, const PassGeom &G, unsigned long long tile, SYNC sync) {
    move_tile<NL, THREADS, TS, false>(sm, in, out, G, tile);
    sync();

    // rounds: forward walks the local bit field from the top, inverse from the bottom;
    // the one partial round (K mod R stages) comes last in both directions.
    const int K = G.K;
    const int full = K / R, rem = K % R;
    // The round executed last in the LAST pass owns the final stage of the transform
    // (inverse: stage 0 with the N^-1 scaling; lazy mode: the canonicalising subtraction).
    for (int r = 0; r < full; r++) {
        const int lb = INV ? r * R : K - (r + 1) * R;
        // (the inverse's closing round and the forward's opening round sit at stage 0: trivial first twiddles)
        if (LAST && rem == 0 && r == full - 1) {
            if (INV && SCALE && G.triv) run_round<NL, QS, R, INV, SCALE, LAZY, true, THREADS, TS, INV && SCALE, FID>(sm, tw, G, lb, tile, C);
            else run_round<NL, QS, R, INV, SCALE, LAZY, true, THREADS, TS, false, FID>(sm, tw, G, lb, tile, C);
        } else if (!INV && r == 0 && G.s0 == 0 && G.triv) run_round<NL, QS, R, false, false, LAZY, false, THREADS, TS, true, FID>(sm, tw, G, lb, tile, C);
        else run_round<NL, QS, R, INV, false, LAZY, false, THREADS, TS, false, FID>(sm, tw, G, lb, tile, C);
        sync();
    }
    if (rem) {
        const int lb = INV ? full * R : 0;
        if (R >= 2 && rem == 1) run_round<NL, QS, 1, INV, SCALE, LAZY, LAST, THREADS, TS, false, FID>(sm, tw, G, lb, tile, C);
        if (R >= 3 && rem == 2) run_round<NL, QS, (R >= 3 ? 2 : 1), INV, SCALE, LAZY, LAST, THREADS, TS, false, FID>(sm, tw, G, lb, tile, C);
        if (R >= 4 && rem == 3) run_round<NL, QS, (R >= 4 ? 3 : 1), INV, SCALE, LAZY, LAST, THREADS, TS, false, FID>(sm, tw, G, lb, tile, C);
        sync();
    }
    move_tile<NL, THREADS, TS, true>(sm, in, out, G, tile);
}

struct CtaSync { __device__ __forceinline__ void operator()() const { __syncthreads(); } };

template <int NL, int QS, int R, int THREADS, int MINB, bool INV, bool SCALE, int LAZY, bool LAST, int FID = FID_RUNTIME>
__global__ void __launch_bounds__(THREADS, MINB) ntt_pass_kernel(uint32_t *out, const uint32_t *in,
                                                           const uint32_t *__restrict__ tw,
                                                           const __grid_constant__ NttConsts<NL> C,
                                                           const __grid_constant__ PassGeom G) {
    extern __shared__ __align__(16) uint32_t sm[];
    ntt_pass_body<NL, QS, R, THREADS, (THREADS << R), INV, SCALE, LAZY, LAST, CtaSync, FID>(sm, out, in, tw, C, G, blockIdx.x, CtaSync());
}

#endif  // __CUDACC__
}  // namespace rs
