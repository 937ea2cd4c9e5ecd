// ntt_rr.cuh -- forward big-field NTT on reduced-radix lazy arithmetic (rr_field.cuh).
//
// Same transform, tables, pass/tile geometry and output as ntt.cuh (math/bigpoly/ntt.go:246-355),
// but inside the kernels an element is W 29-bit limbs kept lazy: the butterfly
//     v' = v * w ; u' = u + v' ; v'' = u + 3q - v'
// uses one carry-free multiplication (144 plain IMAD.WIDE for the 255-bit field) and limb-wise adds
// with no carry propagation or conditional subtraction; limbs are re-normalised once per radix-8
// round and values are brought back to the canonical [0, q) representation only in the final stage.
// A value grows by at most 3q per stage, so rank 2^s needs (1 + 3 s) q < 2^(29 W): up to 2^21 for
// the 255-bit field.  Between passes the lazy elements live in a scratch buffer laid out in
// 16-byte planes (coalesced LDG.128/STG.128).
#pragma once
#include "ntt.cuh"
#include "rr_field.cuh"

namespace rs {

template <int W>
struct RRConsts {
    RRParams<W> P;
};

// shared-memory / scratch plane geometry of a W-limb element: W/4 planes of 16 bytes + W%4 planes of 4 bytes
template <int W> struct RRPlanes {
    static constexpr int P16 = W / 4;
    static constexpr int P4 = W % 4;
};

template <int W, int TS>
struct RRSmem {
    static constexpr int SLOTS = TS + TS / 8;
    static constexpr int P16 = RRPlanes<W>::P16, P4 = RRPlanes<W>::P4;
    static constexpr int STRIDE16 = SLOTS * 4 + 16;       // words; +16 so planes start 16 banks apart
    static constexpr int STRIDE4 = SLOTS + 8;
    static constexpr int WORDS = P16 * STRIDE16 + P4 * STRIDE4;
    static constexpr int BYTES = WORDS * 4;
};

#ifdef __CUDACC__

template <int W, int TS>
__device__ __forceinline__ void rr_smem_load(uint32_t *x, const uint32_t *sm, unsigned idx) {
    using L = RRSmem<W, TS>;
    const unsigned p = pad_idx(idx);
#pragma unroll
    for (int h = 0; h < L::P16; h++) {
        const uint4 v = *reinterpret_cast<const uint4 *>(sm + h * L::STRIDE16 + p * 4);
        x[4 * h] = v.x; x[4 * h + 1] = v.y; x[4 * h + 2] = v.z; x[4 * h + 3] = v.w;
    }
#pragma unroll
    for (int h = 0; h < L::P4; h++) x[4 * L::P16 + h] = sm[L::P16 * L::STRIDE16 + h * L::STRIDE4 + p];
}

template <int W, int TS>
__device__ __forceinline__ void rr_smem_store(uint32_t *sm, unsigned idx, const uint32_t *x) {
    using L = RRSmem<W, TS>;
    const unsigned p = pad_idx(idx);
#pragma unroll
    for (int h = 0; h < L::P16; h++)
        *reinterpret_cast<uint4 *>(sm + h * L::STRIDE16 + p * 4) = make_uint4(x[4 * h], x[4 * h + 1], x[4 * h + 2], x[4 * h + 3]);
#pragma unroll
    for (int h = 0; h < L::P4; h++) sm[L::P16 * L::STRIDE16 + h * L::STRIDE4 + p] = x[4 * L::P16 + h];
}

// twiddle table entry: 12 words (48 B) per twiddle, W used
constexpr int RR_TW_STRIDE = 12;
template <int W>
__device__ __forceinline__ void rr_load_twiddle(uint32_t *w, const uint32_t *__restrict__ tw, unsigned long long idx) {
    const uint32_t *src = tw + idx * RR_TW_STRIDE;
#pragma unroll
    for (int h = 0; h < W / 4; h++) {
        const uint4 v = __ldg(reinterpret_cast<const uint4 *>(src) + h);
        w[4 * h] = v.x; w[4 * h + 1] = v.y; w[4 * h + 2] = v.z; w[4 * h + 3] = v.w;
    }
#pragma unroll
    for (int h = 0; h < W % 4; h++) w[4 * (W / 4) + h] = __ldg(src + 4 * (W / 4) + h);
}

// forward radix-2^RR round on lazy elements (see ntt.cuh radix_round for the index conventions)
template <int W, int QZ, int RR>
__device__ __forceinline__ void rr_radix_round(uint32_t (&x)[1 << RR][W], const uint32_t *__restrict__ tw,
                                               int s_top, unsigned hi, const RRParams<W> &P) {
#pragma unroll
    for (int bb = 0; bb < RR; bb++) {
        const int beta = RR - 1 - bb;
        const int s = s_top + (RR - 1 - beta);
        const unsigned long long tbase = (1ull << s) + ((unsigned long long)hi << (RR - 1 - beta));
#pragma unroll
        for (int g = 0; g < (1 << (RR - 1 - beta)); g++) {
            uint32_t w[W];
            rr_load_twiddle<W>(w, tw, tbase + g);
#pragma unroll
            for (int k = 0; k < (1 << beta); k++) {
                const int u0 = (g << (beta + 1)) + k, u1 = u0 + (1 << beta);
                uint32_t v[W];
                rr_mul<W, QZ>(v, x[u1], w, P);          // ntt.go:254-259  v = v*w
                rr_sub<W>(x[u1], x[u0], v, P);          //                 v' = u - v  (+3q)
                rr_add<W>(x[u0], x[u0], v);             //                 u' = u + v
            }
        }
    }
}

template <int W, int QZ, int RR, bool FINAL, int THREADS, int TS>
__device__ __forceinline__ void rr_run_round(uint32_t *sm, const uint32_t *__restrict__ tw, const PassGeom &G, int lb,
                                             unsigned long long tile, const RRParams<W> &P) {
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
        uint32_t x[1 << RR][W];
#pragma unroll
        for (int u = 0; u < (1 << RR); u++) rr_smem_load<W, TS>(x[u], sm, base + u * ustride);
        rr_radix_round<W, QZ, RR>(x, tw, s_top, hi, P);
#pragma unroll
        for (int u = 0; u < (1 << RR); u++) {
            if (FINAL) rr_canonical<W, QZ>(x[u], P);     // last stage of the transform: back to [0, q)
            else rr_normalise<W>(x[u]);
            rr_smem_store<W, TS>(sm, base + u * ustride, x[u]);
        }
    }
}

// scratch layout of one polynomial batch: per 16-byte plane a contiguous array over all elements,
// then the 4-byte planes
template <int W>
__device__ __forceinline__ size_t rr_scratch_words(size_t total) { return total * W; }

// tile movement.  IN_RR / OUT_RR select the scratch (lazy, planes) format, otherwise canonical AoS words.
template <int NL, int W, int THREADS, int TS, bool STORE, bool RRFMT>
__device__ __forceinline__ void rr_move_tile(uint32_t *sm, const uint32_t *gin, uint32_t *gout, const PassGeom &G,
                                             unsigned long long tile) {
    const unsigned tpp = G.col_major ? 1u : (1u << (G.logN - G.K - G.logW));
    const unsigned long long poly = G.col_major ? 0ull : tile / tpp;
    const unsigned c0 = G.col_major ? 0u : ((unsigned)(tile % tpp) << G.logW);
    const unsigned long long a_sh = G.col_major ? 0ull : ((unsigned long long)(c0 >> G.pl) << (G.pl + G.K));
    const unsigned b0 = G.col_major ? 0u : (c0 & ((1u << G.pl) - 1));
    constexpr int P16 = RRPlanes<W>::P16, P4 = RRPlanes<W>::P4;
    for (unsigned e = threadIdx.x; e < (unsigned)TS; e += THREADS) {
        unsigned long long flat;
        if (G.col_major) {
            flat = tile * (unsigned long long)TS + e;
        } else {
            const unsigned cl = e & ((1u << G.logW) - 1), i = e >> G.logW;
            flat = (poly << G.logN) + a_sh + ((unsigned long long)i << G.pl) + b0 + cl;
        }
        if (flat >= G.total_elems) continue;
        uint32_t x[W];
        if (!STORE) {
            if (RRFMT) {
#pragma unroll
                for (int h = 0; h < P16; h++) {
                    const uint4 v = __ldg(reinterpret_cast<const uint4 *>(gin + (size_t)h * G.total_elems * 4) + flat);
                    x[4 * h] = v.x; x[4 * h + 1] = v.y; x[4 * h + 2] = v.z; x[4 * h + 3] = v.w;
                }
#pragma unroll
                for (int h = 0; h < P4; h++) x[4 * P16 + h] = __ldg(gin + (size_t)P16 * G.total_elems * 4 + (size_t)h * G.total_elems + flat);
            } else {
                uint32_t wds[NL];
#pragma unroll
                for (int h = 0; h < NL / 4; h++) {
                    const uint4 v = __ldg(reinterpret_cast<const uint4 *>(gin + flat * NL) + h);
                    wds[4 * h] = v.x; wds[4 * h + 1] = v.y; wds[4 * h + 2] = v.z; wds[4 * h + 3] = v.w;
                }
                rr_from_words<NL, W>(x, wds);
            }
            rr_smem_store<W, TS>(sm, e, x);
        } else {
            rr_smem_load<W, TS>(x, sm, e);
            if (RRFMT) {
#pragma unroll
                for (int h = 0; h < P16; h++)
                    reinterpret_cast<uint4 *>(gout + (size_t)h * G.total_elems * 4)[flat] = make_uint4(x[4 * h], x[4 * h + 1], x[4 * h + 2], x[4 * h + 3]);
#pragma unroll
                for (int h = 0; h < P4; h++) gout[(size_t)P16 * G.total_elems * 4 + (size_t)h * G.total_elems + flat] = x[4 * P16 + h];
            } else {
                uint32_t wds[NL];
                rr_to_words<NL, W>(wds, x);
#pragma unroll
                for (int h = 0; h < NL / 4; h++)
                    reinterpret_cast<uint4 *>(gout + flat * NL)[h] = make_uint4(wds[4 * h], wds[4 * h + 1], wds[4 * h + 2], wds[4 * h + 3]);
            }
        }
    }
}

template <int NL, int W, int QZ, int R, int THREADS, int MINB, bool IN_RR, bool OUT_RR>
__global__ void __launch_bounds__(THREADS, MINB) ntt_rr_pass_kernel(uint32_t *__restrict__ out, const uint32_t *__restrict__ in,
                                                                    const uint32_t *__restrict__ tw,
                                                                    const __grid_constant__ RRConsts<W> C,
                                                                    const __grid_constant__ PassGeom G) {
    constexpr int TS = THREADS << R;
    extern __shared__ __align__(16) uint32_t sm[];
    const unsigned long long tile = blockIdx.x;
    rr_move_tile<NL, W, THREADS, TS, false, IN_RR>(sm, in, out, G, tile);
    __syncthreads();
    const int K = G.K;
    const int full = K / R, rem = K % R;
    // the pass that writes canonical words owns the last stage of the transform: its last round canonicalises
    for (int r = 0; r < full; r++) {
        const int lb = K - (r + 1) * R;
        if (!OUT_RR && rem == 0 && r == full - 1) rr_run_round<W, QZ, R, true, THREADS, TS>(sm, tw, G, lb, tile, C.P);
        else rr_run_round<W, QZ, R, false, THREADS, TS>(sm, tw, G, lb, tile, C.P);
        __syncthreads();
    }
    if (rem) {
        if (rem == 1) rr_run_round<W, QZ, 1, !OUT_RR, THREADS, TS>(sm, tw, G, 0, tile, C.P);
        if (R >= 3 && rem == 2) rr_run_round<W, QZ, (R >= 3 ? 2 : 1), !OUT_RR, THREADS, TS>(sm, tw, G, 0, tile, C.P);
        __syncthreads();
    }
    rr_move_tile<NL, W, THREADS, TS, true, OUT_RR>(sm, in, out, G, tile);
}

#endif  // __CUDACC__
}  // namespace rs
