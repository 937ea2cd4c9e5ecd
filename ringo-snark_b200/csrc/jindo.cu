// jindo.cu -- Jindo commitment hot path (include/ringo_jindo.h).
//
// Reference: jindo/prover.go:45-324 (Commit, commitColTo, outerCommitTo, Evaluate),
// jindo/encoder.go:113-201 (baseEncodeTo, randEncodeTo), jindo/rns.go:76-114, jindo/utils.go:12-60,
// and the lattigo ring.Ring calls they make.  Ring polynomials are Coeffs[limb][256] uint64.
//
// Kernels
//   encode_kernel   one CTA per ring polynomial: fromMont + base-b digits -> float64 centre ->
//                   TwinCDT sample (or supplied noise) -> MForm(digits) + (X^slots - b) g -> 256-point
//                   NTT per RNS limb -> Opening.  One HBM round trip per polynomial.
//   matvec_kernel   NTT-domain ring matrix-vector products (inner/outer commit, Partial, responses):
//                   per (limb, coefficient) slot, 128-bit lazy accumulation of 64x64 products, one
//                   Montgomery reduction per output.
//   cut_kernel      IMForm/INTT -> CRT to a centred multi-word integer -> arithmetic >> cut ->
//                   residues mod the outer primes -> MForm/NTT  (prover.go:164-176, :190-200).
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "../../include/ringo_jindo.h"
#include "common.cuh"

typedef unsigned long long u64;
typedef long long i64;

#define JD 256            // ring degree handled by these kernels
#define JMAXL RS_JINDO_MAX_LIMBS
#define JW 5              // 64-bit words of the centred integers (Q <= 240 bits, + products)

struct RingDev {
    int n;
    u64 q[JMAXL];
    u64 qinvneg[JMAXL];     // -q^-1 mod 2^64
    u64 r2[JMAXL];          // 2^128 mod q  (MForm multiplier)
    u64 ninv[JMAXL];        // d^-1 mod q, plain (INTT scale that also leaves Montgomery form)
    u64 r1[JMAXL];          // 2^64 mod q, and its Shoup constant floor(r1 2^64 / q): MForm of an unreduced 64-bit value
    u64 r1s[JMAXL];         //   x * r1 - hi(x * r1s) * q  in [0, 2q)  for any x < 2^64
    const u64 *roots;       // [limb][d]  psi^brv(j) * 2^64 mod q   (forward twiddles)
    const u64 *roots_inv;   // [limb][d]  psi^-brv(j) * 2^64 mod q
    const u64 *roots_plain; // [limb][d]  psi^brv(j) mod q             (Shoup butterflies of the warp NTT)
    const u64 *roots_shoup; // [limb][d]  floor(psi^brv(j) * 2^64 / q)
    // float64 butterflies (primes below 2^44: the 35-38 bit rings of the 255-bit field): psi^brv(j), q and 1/q as doubles
    const double *roots_f64;
    double qf[JMAXL], qinvf[JMAXL];
    int f64_ok;
    // kernel-path switches (tuning / the parity tests of the alternative paths): read from the environment ONCE, when the
    // prover handle is created, never on the Commit path
    int no_rowsum, no_pipe, no_split, no_small, use_tma;
};

struct CrtDev {             // RNSReconstructor constants of the *source* ring (rns.go:33-55)
    u64 Q[JW];              // product of the primes
    u64 Qhalf[JW];          // Q >> 1
    u64 Qdiv[JMAXL][JW];    // Q / q_l
    u64 inv_mont[JMAXL];    // ((Q/q_l)^-1 mod q_l) * 2^64 mod q_l
};

namespace rs { struct EncDesc; }
constexpr int RS_VCHUNKS = 8;
struct DescCache {            // encode descriptors of one committed length, resident on the device
    size_t len;
    int c0, c1;               // column range [c0, c1) the descriptors cover (0 .. cols+1 for a whole Commit)
    rs::EncDesc *d_descs;
    size_t n_descs, twin, cosac, mask;
    // descriptors are ordered by row of the committed matrix, i.e. by ascending offset into v: chunk k
    // (descriptors [chunk_desc[k], chunk_desc[k+1])) needs v[0, chunk_vhi[k]) -- lets a host-resident v
    // upload in pieces while earlier pieces are already being encoded (rs_jindo_commit_host)
    size_t chunk_desc[RS_VCHUNKS + 1], chunk_vhi[RS_VCHUNKS];
};

namespace rs {
struct UniformStream;
int uniform_stream_init(UniformStream **out, const uint8_t key[32], const uint8_t iv[16]);
void uniform_stream_free(UniformStream *s);
int uniform_stream_draw(UniformStream *s, size_t n, const u64 **d_samples);
void uniform_stream_release(UniformStream *s);
int uniform_stream_prefetch(UniformStream *s, size_t n);
}  // namespace rs

namespace rs {
struct CosacStream;
int cosac_stream_init(CosacStream **out, const uint8_t kb[48], const uint8_t kr[48], double std_first, double std_mask0, double std_mask);
void cosac_stream_free(CosacStream *s);
u64 cosac_stream_workers(const CosacStream *s);
const int *cosac_stream_overflow_flag(const CosacStream *s);
int cosac_generate(CosacStream *s, const rs_field *f, const rs_jindo_params &p, const uint32_t *d_first_row, const uint32_t *d_mask_rows,
                   size_t n_polys, long long *d_out);
}  // namespace rs

struct rs_jindo {
    std::vector<DescCache> *cache;
    rs::EncDesc *d_ndescs;    // descriptors of the MLWE noise polynomials
    rs::EncDesc *d_ldescs;    // descriptors of the encoded leftVec (rs_jindo_partial_at)
    long long *d_zero_g;      // 256 zero "noise" samples (deterministic Encoder.EncodeTo through encode_kernel)
    rs_field *f;
    rs_jindo_params p;
    RingDev rq, rqo;
    CrtDev crt_q, crt_qo;
    u64 *d_roots;           // all twiddle tables
    u64 *d_ck_in, *d_ck_mlwe, *d_ck_out;
    u64 *d_cdt;
    uint8_t *d_cdt_guide;   // [128][257]: number of table entries below b << 56 (search window per top byte of u)
    double *d_slow_lb;      // [16][cdt_len]: rigorous lower bounds of the slow path's CDF sum (see encode_kernel)
    u64 base_mont_q[JMAXL]; // base * 2^64 mod q_l
    rs::UniformStream *twin_stream;   // GPU-resident uniform source of the encoder's TwinCDT sampler (aes_ctr.cu)
    rs::UniformStream *mlwe_stream;   // same for Prover.mlweSampler (centre 0: table 0 only)
    rs::CosacStream *cosac_stream;    // GPU-resident COSAC sampler of the encoder (cosac_dev.cu)
    u64 *d_mlwe_cdt; int mlwe_cdt_len; long long mlwe_tail_lo;
    int no_pair, prefetch;            // RS_JINDO_NOPAIR / RS_JINDO_PREFETCH, read when the handle is created
    cudaStream_t copy_stream;         // rs_jindo_commit_host: uploads of v pieces
    size_t ck_rejected;               // rs_jindo_create_crs: draws SampleN threw away while generating the key
    cudaEvent_t ev_alloc, ev_piece[RS_VCHUNKS];
};

#define RS_MAX_SHARDS 64      // ranks a Commit can be split over (rs_jindo_shard_columns)

struct rs_opening {
    rs_jindo *j;
    u64 *d_encode;          // (cols+1)*rows polys over ringQ
    u64 *d_mlwe;            // (cols+1)*(mlwe+in_msis) polys over ringQ
    u64 *d_incommit;        // (cols+1)*in_msis polys over ringQOut
};

namespace rs {

// ---------------------------------------------------------------------------------------------
// word-size Montgomery arithmetic (lattigo MRed semantics: x*y*2^-64 mod q, canonical)
// ---------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ u64 hmulhi(u64 a, u64 b) {
#ifdef __CUDA_ARCH__
    return __umul64hi(a, b);
#else
    return (u64)(((unsigned __int128)a * b) >> 64);
#endif
}
__host__ __device__ __forceinline__ u64 redc128(u64 hi, u64 lo, u64 q, u64 qinvneg) {
    // (hi:lo) < q * 2^64
    const u64 m = lo * qinvneg;
    u64 r = hi + hmulhi(m, q) + (lo != 0);
    return r >= q ? r - q : r;
}
__host__ __device__ __forceinline__ u64 mont64(u64 a, u64 b, u64 q, u64 qinvneg) {
    return redc128(hmulhi(a, b), a * b, q, qinvneg);
}
__host__ __device__ __forceinline__ u64 addmod(u64 a, u64 b, u64 q) { u64 s = a + b; return s >= q ? s - q : s; }
__host__ __device__ __forceinline__ u64 submod(u64 a, u64 b, u64 q) { return a >= b ? a - b : a + q - b; }

// exact uint32 -> double without a conversion instruction: (2^52 + x) - 2^52
__device__ __forceinline__ double u32_to_f64(uint32_t x) {
    return __dsub_rn(__hiloint2double(0x43300000, (int)x), 4503599627370496.0);
}

// MForm of a signed 64-bit integer: (c mod q) * 2^64 mod q, canonical -- setCoeffSigned + MForm in one Shoup product
// (valid for every |c| < 2^64: x r1 - hi(x r1s) q < 2q).  The same ring element as mont64(signed_res(c, q), r2).
__device__ __forceinline__ u64 mform_signed(i64 c, u64 q, u64 r1, u64 r1s) {
    const u64 m = (u64)(c < 0 ? -c : c);
    u64 t = m * r1 - __umul64hi(m, r1s) * q;
    t = t >= q ? t - q : t;
    return (c < 0 && t) ? q - t : t;
}

// signed 64-bit value -> canonical residue (setCoeffSigned utils.go:49-60 followed by MForm's reduction)
__device__ __forceinline__ u64 signed_res(i64 c, u64 q) {
    u64 m = (u64)(c < 0 ? -c : c);
    if (__builtin_expect(m >= q, 0)) m %= q;              // never for Gaussian-sized values; kept for the contract
    return (c < 0 && m) ? q - m : m;
}

// 256-point negacyclic NTT of one limb in shared memory, Montgomery domain, 128 threads (t in [0,128)).
// [lattigo] Cooley-Tukey, out[i] = a(psi^(2 brv(i)+1)).
__device__ __forceinline__ void ntt256(u64 *a, const u64 *__restrict__ roots, u64 q, u64 qi, int t, int bar, int nthreads) {
    int tt = JD;
#pragma unroll 1
    for (int m = 1; m < JD; m <<= 1) {
        tt >>= 1;
        const int i = t / tt, j = 2 * i * tt + (t % tt);
        const u64 w = __ldg(roots + m + i);
        const u64 u = a[j], v = mont64(a[j + tt], w, q, qi);
        a[j] = addmod(u, v, q);
        a[j + tt] = submod(u, v, q);
        asm volatile("bar.sync %0, %1;" ::"r"(bar), "r"(nthreads));
    }
}
// [lattigo] Gentleman-Sande inverse; `scale` = d^-1 (plain) also strips the Montgomery factor,
// or d^-1 * 2^64 (Montgomery) to stay in Montgomery form.
__device__ __forceinline__ void intt256(u64 *a, const u64 *__restrict__ roots_inv, u64 q, u64 qi, u64 scale, int t, int bar, int nthreads) {
    int tt = 1;
#pragma unroll 1
    for (int m = JD; m > 1; m >>= 1) {
        const int h = m >> 1;
        const int i = t / tt, j = 2 * i * tt + (t % tt);
        const u64 w = __ldg(roots_inv + h + i);
        const u64 u = a[j], v = a[j + tt];
        a[j] = addmod(u, v, q);
        a[j + tt] = mont64(submod(u, v, q), w, q, qi);
        tt <<= 1;
        asm volatile("bar.sync %0, %1;" ::"r"(bar), "r"(nthreads));
    }
    a[t] = mont64(a[t], scale, q, qi);
    a[t + 128] = mont64(a[t + 128], scale, q, qi);
    asm volatile("bar.sync %0, %1;" ::"r"(bar), "r"(nthreads));
}

// ---------------------------------------------------------------------------------------------
// Warp-level 256-point forward NTT (one warp, 8 coefficients per lane, radix-8 / radix-8 / radix-4
// rounds in registers with two shared-memory exchanges, no block barrier).  Same transform as ntt256
// ([lattigo] Cooley-Tukey, psi powers in bit-reversed order); butterflies are Harvey-lazy with Shoup
// constants: w' = floor(w 2^64 / q), x*w mod q = x*w - hi(x*w')*q in [0, 2q); values stay in [0, 4q)
// (q < 2^62) and are reduced to [0, q) at the end, so the output equals ntt256's bit for bit.
// The polynomial sits in shared memory with 4 pad elements per 32 (index j + (j >> 5) * 4).
// ---------------------------------------------------------------------------------------------
constexpr int NTTW_STRIDE = JD + (JD / 32) * 4;           // padded length of one limb polynomial
__device__ __forceinline__ int nttw_idx(int j) { return j + ((j >> 5) << 2); }

__device__ __forceinline__ void harvey_bfly(u64 &X, u64 &Y, u64 w, u64 ws, u64 q, u64 q2) {
    const u64 x = X >= q2 ? X - q2 : X;                    // [0, 2q)
    const u64 qe = __umul64hi(Y, ws);
    const u64 t = Y * w - qe * q;                          // [0, 2q)
    X = x + t;
    Y = x - t + q2;
}

__device__ __forceinline__ void ntt256_warp(u64 *a, const u64 *__restrict__ rp, const u64 *__restrict__ rs, u64 q, int lane) {
    const u64 q2 = 2 * q;
    u64 v[8];
    // round 1: stages t = 128, 64, 32 on elements j = u*32 + lane
#pragma unroll
    for (int u = 0; u < 8; u++) v[u] = a[nttw_idx(u * 32 + lane)];
#pragma unroll
    for (int u = 0; u < 4; u++) harvey_bfly(v[u], v[u + 4], __ldg(rp + 1), __ldg(rs + 1), q, q2);
#pragma unroll
    for (int u = 0; u < 8; u++) if ((u & 2) == 0) harvey_bfly(v[u], v[u + 2], __ldg(rp + 2 + (u >> 2)), __ldg(rs + 2 + (u >> 2)), q, q2);
#pragma unroll
    for (int u = 0; u < 8; u += 2) harvey_bfly(v[u], v[u + 1], __ldg(rp + 4 + (u >> 1)), __ldg(rs + 4 + (u >> 1)), q, q2);
#pragma unroll
    for (int u = 0; u < 8; u++) a[nttw_idx(u * 32 + lane)] = v[u];
    __syncwarp();
    // round 2: stages t = 16, 8, 4 on elements j = hi*32 + u*4 + lo, hi = lane >> 2, lo = lane & 3
    const int hi = lane >> 2, lo = lane & 3;
#pragma unroll
    for (int u = 0; u < 8; u++) v[u] = a[nttw_idx(hi * 32 + u * 4 + lo)];
    {
        const u64 w = __ldg(rp + 8 + hi), ws = __ldg(rs + 8 + hi);
#pragma unroll
        for (int u = 0; u < 4; u++) harvey_bfly(v[u], v[u + 4], w, ws, q, q2);
    }
#pragma unroll
    for (int u = 0; u < 8; u++) if ((u & 2) == 0) harvey_bfly(v[u], v[u + 2], __ldg(rp + 16 + hi * 2 + (u >> 2)), __ldg(rs + 16 + hi * 2 + (u >> 2)), q, q2);
#pragma unroll
    for (int u = 0; u < 8; u += 2) harvey_bfly(v[u], v[u + 1], __ldg(rp + 32 + hi * 4 + (u >> 1)), __ldg(rs + 32 + hi * 4 + (u >> 1)), q, q2);
#pragma unroll
    for (int u = 0; u < 8; u++) a[nttw_idx(hi * 32 + u * 4 + lo)] = v[u];
    __syncwarp();
    // round 3: stages t = 2, 1 on the 8 consecutive elements j = lane*8 + u
#pragma unroll
    for (int u = 0; u < 8; u++) v[u] = a[nttw_idx(lane * 8 + u)];
#pragma unroll
    for (int u = 0; u < 8; u++) if ((u & 2) == 0) harvey_bfly(v[u], v[u + 2], __ldg(rp + 64 + lane * 2 + (u >> 2)), __ldg(rs + 64 + lane * 2 + (u >> 2)), q, q2);
#pragma unroll
    for (int u = 0; u < 8; u += 2) harvey_bfly(v[u], v[u + 1], __ldg(rp + 128 + lane * 4 + (u >> 1)), __ldg(rs + 128 + lane * 4 + (u >> 1)), q, q2);
#pragma unroll
    for (int u = 0; u < 8; u++) {
        u64 x = v[u];
        x = x >= q2 ? x - q2 : x;
        x = x >= q ? x - q : x;
        a[nttw_idx(lane * 8 + u)] = x;
    }
    __syncwarp();
}

// ---------------------------------------------------------------------------------------------
// The same warp-level transform on the FP64 pipe, for primes below 2^44.  Values are integer-valued doubles; a
// product x*w mod q is h = rn(x*w), l = fma(x, w, -h) (the exact rounding error), b = rint(h / q), r = fma(-b, q, h) + l:
// every step is exact (|x| < 9q < 2^48, h - b q and l are integers far below 2^53), r == x*w (mod q) with |r| < q.
// No reduction is needed between the 8 stages (|value| < (stage + 1) q); the last loop maps to the canonical [0, q).
// Exact modular arithmetic throughout, so the output equals ntt256_warp's / ntt256's bit for bit -- 9 DFMA-class
// instructions per butterfly instead of ~27 integer ones, on a pipe the rest of the kernel leaves idle.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void f64_bfly(double &X, double &Y, double w, double q, double qinv) {
    const double h = __dmul_rn(Y, w);
    const double l = __fma_rn(Y, w, -h);
    const double b = __dadd_rn(__dadd_rn(__dmul_rn(h, qinv), 6755399441055744.0), -6755399441055744.0);   // rint, |h/q| < 2^51
    const double r = __dadd_rn(__fma_rn(-b, q, h), l);
    Y = __dsub_rn(X, r);
    X = __dadd_rn(X, r);
}

__device__ __forceinline__ void ntt256_warp_f64(double *a, const double *__restrict__ rf, double q, double qinv, int lane) {
    double v[8];
#pragma unroll
    for (int u = 0; u < 8; u++) v[u] = a[nttw_idx(u * 32 + lane)];
#pragma unroll
    for (int u = 0; u < 4; u++) f64_bfly(v[u], v[u + 4], __ldg(rf + 1), q, qinv);
#pragma unroll
    for (int u = 0; u < 8; u++) if ((u & 2) == 0) f64_bfly(v[u], v[u + 2], __ldg(rf + 2 + (u >> 2)), q, qinv);
#pragma unroll
    for (int u = 0; u < 8; u += 2) f64_bfly(v[u], v[u + 1], __ldg(rf + 4 + (u >> 1)), q, qinv);
#pragma unroll
    for (int u = 0; u < 8; u++) a[nttw_idx(u * 32 + lane)] = v[u];
    __syncwarp();
    const int hi = lane >> 2, lo = lane & 3;
#pragma unroll
    for (int u = 0; u < 8; u++) v[u] = a[nttw_idx(hi * 32 + u * 4 + lo)];
    {
        const double w = __ldg(rf + 8 + hi);
#pragma unroll
        for (int u = 0; u < 4; u++) f64_bfly(v[u], v[u + 4], w, q, qinv);
    }
#pragma unroll
    for (int u = 0; u < 8; u++) if ((u & 2) == 0) f64_bfly(v[u], v[u + 2], __ldg(rf + 16 + hi * 2 + (u >> 2)), q, qinv);
#pragma unroll
    for (int u = 0; u < 8; u += 2) f64_bfly(v[u], v[u + 1], __ldg(rf + 32 + hi * 4 + (u >> 1)), q, qinv);
#pragma unroll
    for (int u = 0; u < 8; u++) a[nttw_idx(hi * 32 + u * 4 + lo)] = v[u];
    __syncwarp();
#pragma unroll
    for (int u = 0; u < 8; u++) v[u] = a[nttw_idx(lane * 8 + u)];
#pragma unroll
    for (int u = 0; u < 8; u++) if ((u & 2) == 0) f64_bfly(v[u], v[u + 2], __ldg(rf + 64 + lane * 2 + (u >> 2)), q, qinv);
#pragma unroll
    for (int u = 0; u < 8; u += 2) f64_bfly(v[u], v[u + 1], __ldg(rf + 128 + lane * 4 + (u >> 1)), q, qinv);
    u64 *au = reinterpret_cast<u64 *>(a);
#pragma unroll
    for (int u = 0; u < 8; u++) {
        const double x = v[u];                             // integer, |x| < 9q
        const double b = __dadd_rn(__dadd_rn(__dmul_rn(x, qinv), 6755399441055744.0), -6755399441055744.0);
        double r = __fma_rn(-b, q, x);                     // [-q/2, q/2]
        r = r < 0.0 ? __dadd_rn(r, q) : r;
        au[nttw_idx(lane * 8 + u)] = (u64)__double2ll_rn(r);
    }
    __syncwarp();
}

// ---------------------------------------------------------------------------------------------
// encode kernel
// ---------------------------------------------------------------------------------------------
struct EncDesc {
    int src_kind;        // 0: elements from the committed vector, 1: from the aux element buffer, 2: no digits (noise poly)
    int cnt;             // number of field elements (<= slots)
    i64 src_off;         // element offset
    int noise_kind;      // 0: TwinCDT from uniform tape, 1: supplied integers
    i64 noise_off;       // offset (in samples) into twin_u / given noise
    i64 out_poly;        // destination polynomial index
};

struct EncConsts {
    int slots, exp, cdt_len, group;   // group: polynomials per CTA
    int subgroup, cdt_pow2;           // polynomials per NTT staging round; largest power of two <= cdt_len
    u64 base, base_magic;       // base_magic = floor(2^64 / base) + 1  (base is not a power of two)
    i64 tail_lo;
    double sigma;
    double delta_inv[RS_JINDO_MAX_EXP];   // non-zero entries of Encoder.deltaInv, ascending i (encoder.go:152-164 skips zeros)
    int delta_w[RS_JINDO_MAX_EXP];        // (i + 1) * slots of the same entries
    int n_delta;
    u64 base_mont[JMAXL];
    uint32_t base2, base_magic32;   // base^2 when it fits 32 bits (else 0); floor(2^32 / base) + 1
    u64 base2_magic;                // floor(2^64 / base^2) + 1
    i64 g_lim;                  // |g| below this: digits + X^slots g - base g fits an int64 with room (2^62 / base)
};

// One CTA handles G = 256 / slots consecutive descriptors.  Phase 1 (all 256 threads, one field
// element each): fromMont + long division into base-b digits.  Phase 2 (per polynomial, one thread
// per coefficient): centre, sample, combine, NTT.
template <int NL>
__global__ void __launch_bounds__(256) encode_kernel(u64 *__restrict__ out, const EncDesc *__restrict__ descs, int n_descs,
                                                     const uint32_t *__restrict__ vec, const uint32_t *__restrict__ aux,
                                                     const u64 *__restrict__ twin_u, const i64 *__restrict__ given,
                                                     const u64 *__restrict__ cdt, const uint8_t *__restrict__ guide,
                                                     const double *__restrict__ slow_lb,
                                                     const __grid_constant__ RingDev R,
                                                     const __grid_constant__ EncConsts E,
                                                     const __grid_constant__ FieldParams<NL> F) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // layout: digits [G][256] u32 | g [SG][256] i64 | poly [SG * limbs][NTTW_STRIDE] u64
    const int G = E.group;
    uint32_t *s_dig = reinterpret_cast<uint32_t *>(smem_raw);
    i64 *s_gbuf = reinterpret_cast<i64 *>(smem_raw + (size_t)G * JD * 4);
    u64 *s_poly = reinterpret_cast<u64 *>(smem_raw + (size_t)G * JD * 4 + (size_t)E.subgroup * JD * 8);
    const int c = threadIdx.x;
    const int d0 = blockIdx.x * G;
    // TwinCDT slow-path samples (~1 %) are parked here and served by all eight warps after the fast pass, one sample
    // per warp at a time, instead of stalling the warp that happened to draw them (the barrier after phase 2a waits for
    // the unluckiest warp otherwise).  Overflowing samples fall back to the in-warp loop; results do not depend on which.
    constexpr int QCAP = 48;
    __shared__ double q_cfrac[QCAP];
    __shared__ u64 q_u[QCAP];
    __shared__ int q_v0[QCAP], q_v1[QCAP], q_idx[QCAP];
    __shared__ int q_n;
    if (c == 0) q_n = 0;

    // ---- phase 1: baseEncodeTo (encoder.go:120-146) for G polynomials at once
    for (int k = c; k < G * JD; k += 256) s_dig[k] = 0;
    __syncthreads();
    {
        const int g = c / E.slots, i = c % E.slots;
        if (g < G && d0 + g < n_descs) {
            const EncDesc D = descs[d0 + g];
            if (D.src_kind != 2 && i < D.cnt) {
                const uint32_t *src = (D.src_kind == 0 ? vec : aux) + (size_t)(D.src_off + i) * NL;
                uint32_t x[NL], one[NL], y[NL];
#pragma unroll
                for (int l = 0; l < NL; l++) { x[l] = src[l]; one[l] = 0; }
                one[0] = 1;
                mont_mul<NL, 0>(y, x, one, F);                 // Slice(): fromMont (element.go:1769-1773)
                const uint32_t b = (uint32_t)E.base;
                uint32_t *dst = s_dig + g * JD + i;
                int j = 0;
                if (E.base2) {
                    // two digits per long division: divide by base^2 (< 2^32), then split the remainder; limbs that have
                    // become zero are skipped (the quotient loses ~one limb per pass)
                    const uint32_t b2 = E.base2;
                    int top = NL - 1;
                    for (; j + 1 < E.exp - 1; j += 2) {
                        while (top > 0 && y[top] == 0) top--;
                        uint32_t r = 0;
#pragma unroll
                        for (int l = NL - 1; l >= 0; l--) {
                            if (l > top) continue;
                            const u64 cur = ((u64)r << 32) | y[l];
                            u64 qd = hmulhi(cur, E.base2_magic);   // floor(cur / b2) or one more (cur < b2 * 2^32)
                            i64 rem = (i64)(cur - qd * b2);
                            if (rem < 0) { rem += b2; qd--; }
                            y[l] = (uint32_t)qd;
                            r = (uint32_t)rem;
                        }
                        uint32_t hi = __umulhi(r, E.base_magic32);  // floor(r / b) or one more (r < 2^32)
                        int lo = (int)(r - hi * b);
                        if (lo < 0) { lo += (int)b; hi--; }
                        dst[j * E.slots] = (uint32_t)lo;
                        dst[(j + 1) * E.slots] = hi;
                    }
                }
                for (; j < E.exp - 1; j++) {
                    uint32_t r = 0;                            // divMod64 (utils.go:12-18) on 32-bit limbs
#pragma unroll
                    for (int l = NL - 1; l >= 0; l--) {
                        const u64 cur = ((u64)r << 32) | y[l];
                        u64 qd = hmulhi(cur, E.base_magic);    // floor(cur / b) or one more (cur < b * 2^32)
                        i64 rem = (i64)(cur - qd * b);
                        if (rem < 0) { rem += b; qd--; }
                        y[l] = (uint32_t)qd;
                        r = (uint32_t)rem;
                    }
                    dst[j * E.slots] = r;
                }
                dst[(E.exp - 1) * E.slots] = y[0];             // remaining quotient (<= base)
            }
        }
    }
    __syncthreads();

    // Phase 2 runs in sub-groups of SG polynomials so that the NTT staging buffer stays small
    // (shared memory per CTA ~40 KB: 4-5 CTAs per SM and room in L1 for the TwinCDT tables).
    const int SG = E.subgroup;
    const int n_poly = (n_descs - d0 < G) ? n_descs - d0 : G;
    for (int g0 = 0; g0 < n_poly; g0 += SG) {
    const int g1 = (g0 + SG < n_poly) ? g0 + SG : n_poly;
    // ---- phase 2a: one thread per coefficient.  First the Gaussian vectors of all SG polynomials
    //      (no barrier in between), then one barrier, then the combination into s_poly.
    for (int g = g0; g < g1; g++) {
        const EncDesc D = descs[d0 + g];
        const uint32_t *dig = s_dig + g * JD;
        i64 *s_g = s_gbuf + (g - g0) * JD;
        i64 gv;
        if (D.noise_kind == 1) {
            gv = given[D.noise_off + c];
        } else {
            const u64 u = __ldg(twin_u + D.noise_off + c);     // issued first: its HBM latency hides behind the centre sum
            // float64 centre (encoder.go:152-164), unfused multiply-add in the reference's order
            double fp = 0.0;
            for (int i = 0; i < E.n_delta; i++) {          // the non-zero deltaInv terms, ascending i
                const double di = E.delta_inv[i];
                const int w = E.delta_w[i];                // N - d
                // fp[c] += di * dig[c + w - N] for c >= N - w, fp[c] -= di * dig[c + w] below: one load, sign by select
                const double t = __dmul_rn(di, u32_to_f64(dig[(c + w) & (JD - 1)]));
                fp = __dadd_rn(fp, (c >= JD - w) ? t : -t);
            }
            // TwinCDT (gaussian_twin_cdt.go:77-111)
            const double center = -fp;
            const double cfloor = floor(center);
            const double cfrac = __dsub_rn(center, cfloor);
            const double sc = __dmul_rn(128.0, cfrac);
            // 0 <= sc < 128 (cfrac in [0, 1)): floor / ceil fit an int, "% BLOCK" only matters for ceil == 128
            const int c0 = __double2int_rd(sc) & 127, c1 = __double2int_ru(sc) & 127;
            // slices.BinarySearch on both tables at once (branch-free lower bound: number of entries < u),
            // minus one on an exact hit (gaussian_twin_cdt.go:86-93)
            int v[2];
            {
                const u64 *t0 = cdt + (size_t)c0 * E.cdt_len, *t1 = cdt + (size_t)c1 * E.cdt_len;
                int l0 = 0, l1 = 0;
                int e0 = E.cdt_len, e1 = E.cdt_len;            // an exact hit can only sit below these indices
                if (guide) {
                    // the top byte of u narrows the lower bound to [guide[b], guide[b+1]]: 0-1 probes for the
                    // bulk of the distribution, log2(tail length) for the two tail buckets
                    const unsigned b = (unsigned)(u >> 56);
                    const uint8_t *g0p = guide + c0 * 257 + b, *g1p = guide + c1 * 257 + b;
                    int h0 = g0p[1], h1 = g1p[1];
                    l0 = g0p[0]; l1 = g1p[0];
                    e0 = h0; e1 = h1;                          // entries sharing u's top byte end here: beyond, table > u
                    while (l0 < h0 || l1 < h1) {
                        const int m0 = (l0 + h0) >> 1, m1 = (l1 + h1) >> 1;
                        const u64 a0 = (l0 < h0) ? __ldg(t0 + m0) : ~0ull;
                        const u64 a1 = (l1 < h1) ? __ldg(t1 + m1) : ~0ull;
                        if (l0 < h0) { if (a0 < u) l0 = m0 + 1; else h0 = m0; }
                        if (l1 < h1) { if (a1 < u) l1 = m1 + 1; else h1 = m1; }
                    }
                } else {
                    for (int half = E.cdt_pow2; half >= 1; half >>= 1) {
                        const int p0 = l0 + half, p1 = l1 + half;
                        const u64 a0 = (p0 <= E.cdt_len) ? __ldg(t0 + p0 - 1) : ~0ull;
                        const u64 a1 = (p1 <= E.cdt_len) ? __ldg(t1 + p1 - 1) : ~0ull;
                        if (p0 <= E.cdt_len && a0 < u) l0 = p0;
                        if (p1 <= E.cdt_len && a1 < u) l1 = p1;
                    }
                }
                if (l0 < e0 && __ldg(t0 + l0) == u) l0 -= 1;
                if (l1 < e1 && __ldg(t1 + l1) == u) l1 -= 1;
                v[0] = l0; v[1] = l1;
            }
            // Slow path (gaussian_twin_cdt.go:95-110): CDF at v0 for the exact centre.  ~1% of the samples
            // take it; the warp serves its slow lanes one at a time, all 32 lanes evaluating the terms
            // rho(x), x = tail_lo .. v0, in parallel (lane-strided partial sums, then a shuffle tree).
            int pick = v[0];
            bool is_slow = v[0] != v[1];
            if (is_slow && slow_lb && v[0] >= 0 && v[0] < E.cdt_len) {
                // The slow path keeps v0 iff u / 2^64 < sum_{x = tail_lo .. v0} rho(x - cfrac) / norm.  slow_lb[k][v0] bounds
                // that sum from below, term by term, for every cfrac in [k/16, (k+1)/16) (minus a 1e-9 margin): below it
                // the outcome is v0 without evaluating the sum.  (x runs up to the table INDEX v0, as in the reference,
                // so the sum is close to 1 and this settles ~99 % of the slow samples.)
                int kc = (int)(cfrac * 16.0); kc = kc > 15 ? 15 : kc;
                const double p = __dmul_rn(__ull2double_rn(u), 5.421010862427522e-20);
                if (p < __ldg(slow_lb + kc * E.cdt_len + v[0])) is_slow = false;
            }
            if (is_slow) {
                const int slot = atomicAdd(&q_n, 1);
                if (slot < QCAP) {
                    q_cfrac[slot] = cfrac; q_u[slot] = u; q_v0[slot] = v[0]; q_v1[slot] = v[1]; q_idx[slot] = (g - g0) * JD + c;
                    is_slow = false;                       // decided after the barrier; s_g holds the v0 guess until then
                }
            }
            unsigned slow = __ballot_sync(0xffffffffu, is_slow);
            const int lane = c & 31;
            const double norm = __dmul_rn(sqrt(6.283185307179586), E.sigma);
            const double den = __dmul_rn(__dmul_rn(2.0, E.sigma), E.sigma);
            while (slow) {
                const int src = __ffs(slow) - 1;
                slow &= slow - 1;
                const double cf = __shfl_sync(0xffffffffu, cfrac, src);
                const int top = __shfl_sync(0xffffffffu, v[0], src);
                double part = 0.0;
                for (i64 x = E.tail_lo + lane; x <= (i64)top; x += 32) {
                    const double t = __dsub_rn((double)x, cf);
                    part = __dadd_rn(part, exp(__ddiv_rn(__dmul_rn(-t, t), den)));
                }
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) part = __dadd_rn(part, __shfl_xor_sync(0xffffffffu, part, off));
                if (lane == src) {
                    // u / 2^64 < sum rho(x) / norm, with the division by norm moved to the other side
                    const double p = __dmul_rn(__dmul_rn(__ull2double_rn(u), 5.421010862427522e-20), norm);
                    pick = (p < part) ? v[0] : v[1];
                }
            }
            gv = (i64)pick + (i64)cfloor + E.tail_lo;
        }
        s_g[c] = gv;
    }
    __syncthreads();
    {
        // parked slow samples: the CDF at v0 for the exact centre, 32 lanes per sample (same sums as the in-warp loop)
        const int nq = q_n < QCAP ? q_n : QCAP;
        const int warp = c >> 5, lane = c & 31;
        const double norm = __dmul_rn(sqrt(6.283185307179586), E.sigma);
        const double den = __dmul_rn(__dmul_rn(2.0, E.sigma), E.sigma);
        for (int e = warp; e < nq; e += 8) {
            const double cf = q_cfrac[e];
            const int top = q_v0[e];
            double part = 0.0;
            for (i64 x = E.tail_lo + lane; x <= (i64)top; x += 32) {
                const double t = __dsub_rn((double)x, cf);
                part = __dadd_rn(part, exp(__ddiv_rn(__dmul_rn(-t, t), den)));
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) part = __dadd_rn(part, __shfl_xor_sync(0xffffffffu, part, off));
            if (lane == 0) {
                const double p = __dmul_rn(__dmul_rn(__ull2double_rn(q_u[e]), 5.421010862427522e-20), norm);   // u / 2^64 * norm
                if (!(p < part)) s_gbuf[q_idx[e]] += (i64)(q_v1[e] - top);                     // the guess was v0
            }
        }
    }
    __syncthreads();
    if (c == 0) q_n = 0;                                   // (the barriers of phase 2b order this before the next pushes)
    for (int g = g0; g < g1; g++) {
        const int kind = descs[d0 + g].src_kind;
        const uint32_t *dig = s_dig + g * JD;
        const i64 *s_g = s_gbuf + (g - g0) * JD;
        const i64 gv = s_g[c];
        // MForm(digits) + MForm(X^slots g - base g)  (encoder.go:166-199), per RNS limb.  The three terms are small
        // integers, so they are combined in int64 first and reduced once per limb (same ring element: every step of
        // the reference is exact arithmetic mod q); values too large for that keep the term-by-term path.
        const bool wrap = c < E.slots;
        const i64 gnb = (kind == 2) ? 0 : (wrap ? s_g[c + JD - E.slots] : s_g[c - E.slots]);
        const bool small = gv > -E.g_lim && gv < E.g_lim && gnb > -(1ll << 61) && gnb < (1ll << 61);
        // (base * g: one signed 32x32 -> 64 multiply when g fits 32 bits, as every Gaussian sample does)
        const i64 bg = (gv == (i64)(int)gv) ? (i64)(int)gv * (i64)(int)E.base : (i64)E.base * gv;
        const i64 comb = (kind == 2) ? gv : (i64)dig[c] + (wrap ? -gnb : gnb) - bg;
        for (int l = 0; l < R.n; l++) {
            const u64 q = R.q[l], qi = R.qinvneg[l];
            u64 mf;
            if (kind == 2 || small) {
                mf = mform_signed(comb, q, R.r1[l], R.r1s[l]);      // noise polynomial: setCoeffSigned (prover.go:130-139)
            } else {
                const u64 gm = signed_res(gv, q);
                const u64 gs = wrap ? submod(0, signed_res(gnb, q), q) : signed_res(gnb, q);
                const u64 bg = mont64(E.base_mont[l], gm, q, qi);        // base * g mod q
                const u64 val = submod(addmod((u64)dig[c], gs, q), bg, q);   // digits <= base < 2^31 < q
                mf = mont64(val, R.r2[l], q, qi);                        // MForm
            }
            s_poly[(size_t)((g - g0) * R.n + l) * NTTW_STRIDE + nttw_idx(c)] = R.f64_ok ? (u64)__double_as_longlong(__ull2double_rn(mf)) : mf;
        }
    }
    __syncthreads();
    // ---- phase 2b: one warp per (polynomial, limb): 256-point NTT in registers, then the coalesced store
    {
        const int warp = c >> 5, lane = c & 31;
        for (int task = warp; task < (g1 - g0) * R.n; task += 8) {
            const int g = g0 + task / R.n, l = task % R.n;
            u64 *a = s_poly + (size_t)task * NTTW_STRIDE;
            if (R.f64_ok) ntt256_warp_f64(reinterpret_cast<double *>(a), R.roots_f64 + (size_t)l * JD, R.qf[l], R.qinvf[l], lane);
            else ntt256_warp(a, R.roots_plain + (size_t)l * JD, R.roots_shoup + (size_t)l * JD, R.q[l], lane);
            u64 *dst = out + ((size_t)descs[d0 + g].out_poly * R.n + l) * JD;
#pragma unroll
            for (int u = 0; u < 8; u++) dst[u * 32 + lane] = a[nttw_idx(u * 32 + lane)];
        }
    }
    __syncthreads();
    }   // sub-groups
}

// ---------------------------------------------------------------------------------------------
// matvec kernel:  out[cc][o] (+)= REDC( sum_k key[o][k] * x[cc][k] )  per slot
//   key poly index  = o * key_os + k * key_ks      x poly index = cc * x_cs + k * x_ks
//   out poly index  = cc * out_cs + o * out_os     addend poly index = cc * add_cs + o * add_os
// ---------------------------------------------------------------------------------------------
struct MatvecGeom {
    int n_limbs, O, C, K;
    i64 key_os, key_ks, key_cs, x_cs, x_ks, out_cs, out_os, add_cs, add_os;
    int accumulate;      // add the reduced sum to the existing out value
    int has_addend;
    int fold_every;      // reduce the high word mod q every this many terms (0 = never)
    int kara_sh;         // matvec_pipe_kernel<.., KARA>: residues are split at this many bits (0: unused)
    int k_split;         // matvec_pipe_kernel: the K range is cut into k_split pieces (grid.y carries the piece); piece s writes its
    i64 split_stride;    // reduced partial sums at out + s * split_stride words, summed afterwards by reduce_split_kernel (0/1: off)
    u64 q[JMAXL], qinvneg[JMAXL], r2[JMAXL];
};

// block = (32 slots, WARPS column tiles): the warps of a CTA work on the same slots and the same
// output rows, so the key values they load are shared through L1.  Measured alternatives (B200,
// 2^22 commit): CT = 1 per thread or 8+ column tiles per CTA run 5-10x slower (many polynomial streams
// ~8 MB apart per SM); one column tile per CTA with 4 slot chunks is 1.3x slower (4x the key traffic).
// The kernel is memory-latency bound (12 warps/SM, ~170 registers); the planned next step is a
// cp.async.bulk + mbarrier ring of key/x rows in shared memory.
//
// Each 64x64 product is accumulated lazily, split by 32-bit halves so that no multiply needs a
// carry-in (the carry-in form IMAD.WIDE.X runs at half rate on sm_100):
//   S0 (96 bit)  += k0*x0      one IMAD.WIDE with carry-out + one IADD3.X
//   S1 (64 bit)  += k0*x1 + k1*x0,   S2 (64 bit) += k1*x1     plain IMAD.WIDE (k1, x1 < 2^(bits-32))
// FOLD: spill the high words of S1/S2 every fold_every terms (only needed for primes > ~50 bits and
// long sums).  One Montgomery reduction per output at the end: the result is the same field element
// as the reference's chain of MulCoeffsMontgomeryThenAdd.
// 64-bit accumulator; the multiply-add is written as a mad.lo.cc / madc.hi pair on its halves so that ptxas
// emits ONE IMAD.WIDE.U32 with the 64-bit addend (it splits a plain mad.wide.u32 into a multiply and a
// two-instruction add)
struct Acc64 {
    u64 v;
    __device__ __forceinline__ u64 value() const { return v; }
};
__device__ __forceinline__ void acc_mad(Acc64 &s, uint32_t a, uint32_t b) {
    asm("{\n\t.reg .u32 lo, hi;\n\tmov.b64 {lo, hi}, %0;\n\tmad.lo.cc.u32 lo, %1, %2, lo;\n\tmadc.hi.u32 hi, %1, %2, hi;\n\tmov.b64 %0, {lo, hi};\n\t}"
        : "+l"(s.v) : "r"(a), "r"(b));
}
// one lazy multiply-accumulate of two 64-bit residues k = k1 2^32 + k0, x = x1 2^32 + x0:
//   S0 (96 bit) += k0 x0 ;  S1 += k0 x1 + k1 x0 ;  S2 += k1 x1       -- 5 instructions (IMAD.WIDE + IADD3.X, 3 x IMAD.WIDE)
__device__ __forceinline__ void mac_lazy(uint32_t &s0a, uint32_t &s0b, uint32_t &s0c, Acc64 &s1, Acc64 &s2, u64 kv, u64 xv) {
    const uint32_t k0 = (uint32_t)kv, k1 = (uint32_t)(kv >> 32);
    const uint32_t x0 = (uint32_t)xv, x1 = (uint32_t)(xv >> 32);
    asm("mad.lo.cc.u32 %0, %3, %4, %0;\n\tmadc.hi.cc.u32 %1, %3, %4, %1;\n\taddc.u32 %2, %2, 0;"
        : "+r"(s0a), "+r"(s0b), "+r"(s0c) : "r"(k0), "r"(x0));
    acc_mad(s1, k0, x1);
    acc_mad(s1, k1, x0);
    acc_mad(s2, k1, x1);
}
// the same with the top product k1 x1 accumulated in 32 bits by a plain IMAD (half the pipe cycles of an IMAD.WIDE): valid while
// K 2^(2 (bits - 32)) < 2^32, i.e. for the 38-bit RNS primes of the 255-bit parameter sets up to K = 2^20 (launch_matvec checks)
__device__ __forceinline__ void mac_lazy_n(uint32_t &s0a, uint32_t &s0b, uint32_t &s0c, Acc64 &s1, uint32_t &s2, u64 kv, u64 xv) {
    const uint32_t k0 = (uint32_t)kv, k1 = (uint32_t)(kv >> 32);
    const uint32_t x0 = (uint32_t)xv, x1 = (uint32_t)(xv >> 32);
    asm("mad.lo.cc.u32 %0, %3, %4, %0;\n\tmadc.hi.cc.u32 %1, %3, %4, %1;\n\taddc.u32 %2, %2, 0;"
        : "+r"(s0a), "+r"(s0b), "+r"(s0c) : "r"(k0), "r"(x0));
    acc_mad(s1, k0, x1);
    acc_mad(s1, k1, x0);
    s2 += k1 * x1;
}

template <int OT, int CT, int WARPS, bool FOLD, int UNR = 2>
__global__ void __launch_bounds__(32 * WARPS) matvec_kernel(u64 *__restrict__ out, const u64 *__restrict__ key,
                                                           const u64 *__restrict__ x, const u64 *__restrict__ addend,
                                                           const __grid_constant__ MatvecGeom G) {
    const int slot = blockIdx.x * 32 + threadIdx.x;          // limb * 256 + coeff
    const int n_slots = G.n_limbs * JD;
    const int o0 = blockIdx.y * OT;
    const int c0 = (blockIdx.z * WARPS + threadIdx.y) * CT;
    if (slot >= n_slots || c0 >= G.C) return;
    const int limb = slot / JD;
    const u64 q = G.q[limb], qi = G.qinvneg[limb];
    const size_t pw = (size_t)n_slots;                                // words per polynomial
    uint32_t s0a[OT][CT], s0b[OT][CT], s0c[OT][CT];
    Acc64 s1[OT][CT], s2[OT][CT]; u64 s1h[OT][CT], s2h[OT][CT];
#pragma unroll
    for (int a = 0; a < OT; a++)
#pragma unroll
        for (int b = 0; b < CT; b++) { s0a[a][b] = s0b[a][b] = s0c[a][b] = 0; s1[a][b] = s2[a][b] = Acc64{0}; s1h[a][b] = s2h[a][b] = 0; }
    int since = 0;
    // running pointers (ragged tiles clamp to the last valid row / column; their results are not stored)
    const u64 *xp[CT], *kp[OT][CT];
#pragma unroll
    for (int b = 0; b < CT; b++) {
        const int cc = (c0 + b < G.C) ? c0 + b : G.C - 1;
        xp[b] = x + (size_t)(cc * G.x_cs) * pw + slot;
#pragma unroll
        for (int a = 0; a < OT; a++) {
            const int oo = (o0 + a < G.O) ? o0 + a : G.O - 1;
            kp[a][b] = key + (size_t)(oo * G.key_os + cc * G.key_cs) * pw + slot;
        }
    }
    const size_t x_step = (size_t)G.x_ks * pw, k_step = (size_t)G.key_ks * pw;
    const bool key_per_col = (G.key_cs != 0);
#pragma unroll(UNR)
    for (int k = 0; k < G.K; k++) {
        u64 kv[OT][CT], xv[CT];
#pragma unroll
        for (int b = 0; b < CT; b++) { xv[b] = __ldg(xp[b]); xp[b] += x_step; }
#pragma unroll
        for (int a = 0; a < OT; a++)
#pragma unroll
            for (int b = 0; b < CT; b++) {
                if (!key_per_col && b > 0) { kv[a][b] = kv[a][0]; continue; }
                kv[a][b] = __ldg(kp[a][b]);
                kp[a][b] += k_step;
            }
#pragma unroll
        for (int a = 0; a < OT; a++)
#pragma unroll
            for (int b = 0; b < CT; b++) {
                mac_lazy(s0a[a][b], s0b[a][b], s0c[a][b], s1[a][b], s2[a][b], kv[a][b], xv[b]);
            }
        if (FOLD && ++since == G.fold_every) {
            since = 0;
#pragma unroll
            for (int a = 0; a < OT; a++)
#pragma unroll
                for (int b = 0; b < CT; b++) {
                    s1h[a][b] += s1[a][b].v >> 32; s1[a][b].v &= 0xffffffffull;
                    s2h[a][b] += s2[a][b].v >> 32; s2[a][b].v &= 0xffffffffull;
                }
        }
    }
#pragma unroll
    for (int a = 0; a < OT; a++)
#pragma unroll
        for (int b = 0; b < CT; b++) {
            if (o0 + a >= G.O || c0 + b >= G.C) continue;
            // T = S0 + (S1 + S1h 2^32) 2^32 + (S2 + S2h 2^32) 2^64   as three 64-bit words (w2 small)
            u64 w0 = ((u64)s0b[a][b] << 32) | s0a[a][b], w1 = s0c[a][b], w2 = 0;
            auto add_at32 = [&](u64 v) {       // T += v << 32
                const u64 lo = v << 32, hi = v >> 32;
                const u64 n0 = w0 + lo; const u64 cy = n0 < lo; w0 = n0;
                const u64 n1 = w1 + hi; u64 cy2 = n1 < hi; const u64 n1b = n1 + cy; cy2 += n1b < n1; w1 = n1b; w2 += cy2;
            };
            auto add_at64 = [&](u64 v) { const u64 n1 = w1 + v; w2 += n1 < v; w1 = n1; };
            add_at32(s1[a][b].value());
            add_at64(s2[a][b].value());
            if (FOLD) { add_at64(s1h[a][b]); const u64 v = s2h[a][b]; add_at64(v << 32); w2 += v >> 32; }
            // h = (w2 2^64 + w1) mod q ;  result = (h 2^64 + w0) 2^-64 mod q
            u64 h = w1 >= q ? w1 % q : w1;
            if (w2) h = addmod(h, mont64(w2 >= q ? w2 % q : w2, G.r2[limb], q, qi), q);
            u64 r = redc128(h, w0, q, qi);                 // == sum_k key*x * 2^-64 mod q (MulCoeffsMontgomeryThenAdd chain)
            u64 *dst = out + ((c0 + b) * G.out_cs + (o0 + a) * G.out_os) * pw + slot;
            if (G.accumulate) r = addmod(r, *dst, q);
            if (G.has_addend) r = addmod(r, __ldg(addend + ((c0 + b) * G.add_cs + (o0 + a) * G.add_os) * pw + slot), q);
            *dst = r;
        }
}

// ---------------------------------------------------------------------------------------------
// matvec_pipe_kernel: same product as matvec_kernel, with the key / x rows streamed through a
// multi-stage shared-memory ring by cp.async (LDGSTS) so that the memory-level parallelism needed to
// cover HBM latency lives in shared memory instead of registers.
//   CTA tile : MS = 64 slots x OT key rows x CTC = 16 columns; 512 threads = 64 slots x 8 column pairs,
//              10 outputs per thread.  A stage holds KS = 4 consecutive k-steps (42 KB); 3 stages.
//   Per k-step a thread reads 5 key values (broadcast across the 8 column-pair groups) and 2 x values
//   from shared memory and issues 10 lazy multiply-accumulates.
// ---------------------------------------------------------------------------------------------
#ifndef RS_MP_CTC
#define RS_MP_CTC 16          // columns per CTA tile: 16 -> 512 threads, one CTA per SM.  8 -> 256 threads, two CTAs per SM was
#endif                        // measured slower (Commit 2^22: 3.29 -> 3.53 ms: twice the key traffic, half the reuse per key load)
constexpr int MP_MS = 64, MP_OT = 5, MP_CTC = RS_MP_CTC, MP_KS = 4, MP_STAGES = 3, MP_THREADS = MP_MS * MP_CTC / 2;
constexpr int MP_MINB = (MP_CTC <= 8) ? 2 : 1;
constexpr int MP_ROWS = MP_OT + MP_CTC;                         // polynomial rows per k-step
constexpr int MP_STAGE_WORDS = MP_KS * MP_ROWS * MP_MS;         // u64 words per stage
constexpr int MP_SMEM_BYTES = MP_STAGES * MP_STAGE_WORDS * 8;

// TMA == true: the stage's (k, row) segments -- 512 contiguous bytes each -- are fetched by the copy engine (cp.async.bulk,
// SASS UBLKCP) issued by one warp and counted on an mbarrier per stage (complete_tx), instead of 16-byte LDGSTS pieces issued
// by every thread with their address arithmetic and a cp.async.wait_group.
// S2N: narrow accumulator for the k1 x1 products (mac_lazy_n; never together with FOLD)
// KARA: (never together with FOLD / S2N) residues below 2^(2 sh) are split at sh bits, k = k1 2^sh + k0, and a multiply-accumulate
//       is THREE plain 64-bit IMAD.WIDE: A += k0 x0, B += k1 x1, C += (k0 + k1)(x0 + x1); the middle term is C - A - B.  No carry
//       word (every product is below 2^(2 sh + 2), launch_matvec checks K 2^(2 sh + 2) < 2^64); the splits cost three ALU
//       instructions per loaded value, shared by the 2 resp. 5 products the value takes part in.
template <bool FOLD, bool TMA, bool S2N = false, bool KARA = false>
__global__ void __launch_bounds__(MP_THREADS, MP_MINB) matvec_pipe_kernel(u64 *__restrict__ out, const u64 *__restrict__ key,
                                                                   const u64 *__restrict__ x, const u64 *__restrict__ addend,
                                                                   const __grid_constant__ MatvecGeom G) {
    extern __shared__ __align__(16) u64 ring[];
    const int tx = threadIdx.x & (MP_MS - 1), ty = threadIdx.x / MP_MS;     // slot in tile, column pair
    const int slot0 = blockIdx.x * MP_MS;
    const int n_slots = G.n_limbs * JD;
    // split-K (a column shard of a Commit has too few column tiles to fill the chip): grid.y = o-tiles x pieces
    const int n_ot = (G.O + MP_OT - 1) / MP_OT;
    const int ksp = G.k_split > 1 ? (int)blockIdx.y / n_ot : 0;
    const int k_per = G.k_split > 1 ? (G.K + G.k_split - 1) / G.k_split : G.K;
    const int k_begin = ksp * k_per, k_cnt = (k_begin + k_per <= G.K ? k_per : (G.K > k_begin ? G.K - k_begin : 0));
    const int o0 = ((int)blockIdx.y - ksp * n_ot) * MP_OT, cb = blockIdx.z * MP_CTC;
    const size_t pw = (size_t)n_slots;
    const int limb = (slot0 + tx) / JD;
    const u64 q = G.q[limb], qi = G.qinvneg[limb];
    out += (size_t)ksp * (size_t)G.split_stride;

    // ---- loader.  A stage is MP_KS k-steps x MP_ROWS polynomial rows (OT key rows, then CTC x columns) x
    //      PIECES 16-byte pieces.  Per CTA the rows' base pointers and per-k strides are fixed: they sit in a
    //      small shared table, and a thread's chunk list (row, k_local) is static, so one chunk costs a table
    //      read, a 32x32+64 multiply-add for the address and the cp.async.
    constexpr int PIECES = MP_MS / 2;                               // 16-byte pieces per row
    constexpr int PAIRS = MP_KS * MP_ROWS;                          // (k_local, row) pairs per stage
    constexpr int PAIRS_PER_ROUND = MP_THREADS / PIECES;
    __shared__ unsigned long long row_base[MP_ROWS];                // global address of (row, k = 0, this tile's first slot)
    __shared__ unsigned row_step[MP_ROWS];                          // bytes per k-step
    __shared__ __align__(8) unsigned long long full_bar[MP_STAGES]; // TMA: one transaction barrier per stage
    if (TMA && threadIdx.x == 0) {
        for (int st = 0; st < MP_STAGES; st++)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((unsigned)__cvta_generic_to_shared(&full_bar[st])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < MP_ROWS) {
        const int row = threadIdx.x;
        const u64 *src; size_t step;
        if (row < MP_OT) {
            const int oo = (o0 + row < G.O) ? o0 + row : G.O - 1;
            src = key + (size_t)(oo * G.key_os) * pw; step = (size_t)G.key_ks * pw * 8;
        } else {
            const int cc = (cb + row - MP_OT < G.C) ? cb + row - MP_OT : G.C - 1;
            src = x + (size_t)(cc * G.x_cs) * pw; step = (size_t)G.x_ks * pw * 8;
        }
        row_base[row] = (unsigned long long)(src + slot0) + (unsigned long long)k_begin * step;
        row_step[row] = (unsigned)step;                             // < 2^32: launch_matvec checks
    }
    __syncthreads();
    const int piece = threadIdx.x % PIECES, pair0 = threadIdx.x / PIECES;
    const unsigned ring_s = (unsigned)__cvta_generic_to_shared(ring);
    auto load_stage = [&](int it, int buf) {
        if (TMA) {
            if (threadIdx.x < 32) {
                const unsigned bar = (unsigned)__cvta_generic_to_shared(&full_bar[buf]);
                if (threadIdx.x == 0)
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((unsigned)(PAIRS * MP_MS * 8)) : "memory");
                __syncwarp();
                for (int pair = threadIdx.x; pair < PAIRS; pair += 32) {
                    const int kl = pair / MP_ROWS, row = pair - kl * MP_ROWS;
                    int k = it * MP_KS + kl;
                    if (k >= k_cnt) k = k_cnt > 0 ? k_cnt - 1 : 0;          // tail: harmless duplicate, not accumulated
                    const unsigned long long src = row_base[row] + (unsigned long long)(unsigned)k * row_step[row];
                    const unsigned dstp = ring_s + (unsigned)buf * (MP_STAGE_WORDS * 8) + (unsigned)pair * (MP_MS * 8);
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 ::"r"(dstp), "l"(src), "r"((unsigned)(MP_MS * 8)), "r"(bar) : "memory");
                }
            }
            return;
        }
        const unsigned dst = ring_s + (unsigned)buf * (MP_STAGE_WORDS * 8) + piece * 16;
#pragma unroll
        for (int r = 0; r < (PAIRS + PAIRS_PER_ROUND - 1) / PAIRS_PER_ROUND; r++) {
            const int pair = pair0 + r * PAIRS_PER_ROUND;
            if (pair >= PAIRS) break;
            const int kl = pair / MP_ROWS, row = pair - kl * MP_ROWS;
            int k = it * MP_KS + kl;
            if (k >= k_cnt) k = k_cnt > 0 ? k_cnt - 1 : 0;          // tail: harmless duplicate, not accumulated
            const unsigned long long src = row_base[row] + (unsigned long long)(unsigned)k * row_step[row] + piece * 16;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + (unsigned)pair * (MP_MS * 8)), "l"(src));
        }
        asm volatile("cp.async.commit_group;");
    };

    uint32_t s0a[MP_OT][2], s0b[MP_OT][2], s0c[MP_OT][2];
    Acc64 s1[MP_OT][2], s2[MP_OT][2]; u64 s1h[MP_OT][2], s2h[MP_OT][2];
    uint32_t s2n[MP_OT][2];
    Acc64 sk[MP_OT][2];
    const int kara_sh = G.kara_sh;
    const uint32_t kara_mask = (1u << kara_sh) - 1u;
#pragma unroll
    for (int a = 0; a < MP_OT; a++)
#pragma unroll
        for (int b = 0; b < 2; b++) { s0a[a][b] = s0b[a][b] = s0c[a][b] = 0; s1[a][b] = s2[a][b] = sk[a][b] = Acc64{0}; s1h[a][b] = s2h[a][b] = 0; s2n[a][b] = 0; }

    const int n_it = (k_cnt + MP_KS - 1) / MP_KS;
    for (int st = 0; st < MP_STAGES - 1; st++) {
        if (st < n_it) load_stage(st, st); else if (!TMA) asm volatile("cp.async.commit_group;");
    }
    int since = 0;
    for (int it = 0; it < n_it; it++) {
        // groups it and it+1 are in flight: wait for group it, then ONE barrier per iteration -- it publishes stage it and, since
        // every thread has finished iteration it-1 by now, frees that buffer for the load of iteration it + STAGES - 1
        if (TMA) {
            // stage it has landed when its barrier has completed phase (it / STAGES): every thread waits on it itself
            const unsigned bar = (unsigned)__cvta_generic_to_shared(&full_bar[it % MP_STAGES]);
            const unsigned parity = (unsigned)((it / MP_STAGES) & 1);
            asm volatile("{\n\t.reg .pred p;\n\tRS_MP_WAIT:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra RS_MP_DONE;\n\tbra RS_MP_WAIT;\n\tRS_MP_DONE:\n\t}"
                         ::"r"(bar), "r"(parity) : "memory");
        } else {
            asm volatile("cp.async.wait_group %0;" ::"n"(MP_STAGES - 2));
        }
        __syncthreads();
        if (it + MP_STAGES - 1 < n_it) load_stage(it + MP_STAGES - 1, (it + MP_STAGES - 1) % MP_STAGES);
        else if (!TMA) asm volatile("cp.async.commit_group;");
        const u64 *buf = ring + (size_t)(it % MP_STAGES) * MP_STAGE_WORDS;
#pragma unroll
        for (int kl = 0; kl < MP_KS; kl++) {
            if (it * MP_KS + kl >= k_cnt) break;
            const u64 *rowp = buf + (size_t)kl * MP_ROWS * MP_MS + tx;
            u64 kv[MP_OT], xv[2];
#pragma unroll
            for (int a = 0; a < MP_OT; a++) kv[a] = rowp[a * MP_MS];
#pragma unroll
            for (int b = 0; b < 2; b++) xv[b] = rowp[(MP_OT + ty * 2 + b) * MP_MS];
            if (KARA) {
                uint32_t k0[MP_OT], k1[MP_OT], ks[MP_OT], x0[2], x1[2], xs[2];
#pragma unroll
                for (int a = 0; a < MP_OT; a++) { k0[a] = (uint32_t)kv[a] & kara_mask; k1[a] = (uint32_t)(kv[a] >> kara_sh); ks[a] = k0[a] + k1[a]; }
#pragma unroll
                for (int b = 0; b < 2; b++) { x0[b] = (uint32_t)xv[b] & kara_mask; x1[b] = (uint32_t)(xv[b] >> kara_sh); xs[b] = x0[b] + x1[b]; }
#pragma unroll
                for (int a = 0; a < MP_OT; a++)
#pragma unroll
                    for (int b = 0; b < 2; b++) {
                        acc_mad(s1[a][b], k0[a], x0[b]);      // A
                        acc_mad(s2[a][b], k1[a], x1[b]);      // B
                        acc_mad(sk[a][b], ks[a], xs[b]);      // C
                    }
                continue;
            }
#pragma unroll
            for (int a = 0; a < MP_OT; a++)
#pragma unroll
                for (int b = 0; b < 2; b++) {
                    if (S2N) mac_lazy_n(s0a[a][b], s0b[a][b], s0c[a][b], s1[a][b], s2n[a][b], kv[a], xv[b]);
                    else mac_lazy(s0a[a][b], s0b[a][b], s0c[a][b], s1[a][b], s2[a][b], kv[a], xv[b]);
                }
            if (FOLD && ++since == G.fold_every) {
                since = 0;
#pragma unroll
                for (int a = 0; a < MP_OT; a++)
#pragma unroll
                    for (int b = 0; b < 2; b++) {
                        s1h[a][b] += s1[a][b].v >> 32; s1[a][b].v &= 0xffffffffull;
                        s2h[a][b] += s2[a][b].v >> 32; s2[a][b].v &= 0xffffffffull;
                    }
            }
        }
    }
    if (S2N) {
#pragma unroll
        for (int a = 0; a < MP_OT; a++)
#pragma unroll
            for (int b = 0; b < 2; b++) s2[a][b].v = s2n[a][b];
    }
    const int slot = slot0 + tx;
#pragma unroll
    for (int a = 0; a < MP_OT; a++)
#pragma unroll
        for (int b = 0; b < 2; b++) {
            const int oo = o0 + a, cc = cb + ty * 2 + b;
            if (oo >= G.O || cc >= G.C) continue;
            u64 w0 = ((u64)s0b[a][b] << 32) | s0a[a][b], w1 = s0c[a][b], w2 = 0;
            auto add_at32 = [&](u64 v) {
                const u64 lo = v << 32, hi = v >> 32;
                const u64 n0 = w0 + lo; const u64 cy = n0 < lo; w0 = n0;
                const u64 n1 = w1 + hi; u64 cy2 = n1 < hi; const u64 n1b = n1 + cy; cy2 += n1b < n1; w1 = n1b; w2 += cy2;
            };
            auto add_at64 = [&](u64 v) { const u64 n1 = w1 + v; w2 += n1 < v; w1 = n1; };
            if (KARA) {
                // sum = A + (C - A - B) 2^sh + B 2^(2 sh), below 2^128
                const u64 A = s1[a][b].value(), B = s2[a][b].value(), M = sk[a][b].value() - A - B;
                const unsigned __int128 t = (unsigned __int128)A + ((unsigned __int128)M << kara_sh) + ((unsigned __int128)B << (2 * kara_sh));
                w0 = (u64)t; w1 = (u64)(t >> 64);
            } else {
                add_at32(s1[a][b].value());
                add_at64(s2[a][b].value());
            }
            if (FOLD) { add_at64(s1h[a][b]); const u64 v = s2h[a][b]; add_at64(v << 32); w2 += v >> 32; }
            u64 h = w1 >= q ? w1 % q : w1;
            if (w2) h = addmod(h, mont64(w2 >= q ? w2 % q : w2, G.r2[limb], q, qi), q);
            u64 r = redc128(h, w0, q, qi);
            u64 *dst = out + ((size_t)cc * G.out_cs + (size_t)oo * G.out_os) * pw + slot;
            if (G.accumulate) r = addmod(r, *dst, q);
            if (G.has_addend) r = addmod(r, __ldg(addend + ((size_t)cc * G.add_cs + (size_t)oo * G.add_os) * pw + slot), q);
            *dst = r;
        }
}

// out[o] = sum_c part[c][o] mod q   (split-K reduction of the outer commit)
__global__ void __launch_bounds__(256) reduce_cols_kernel(u64 *__restrict__ out, const u64 *__restrict__ part, int C, int O,
                                                          const __grid_constant__ MatvecGeom G) {
    const size_t pw = (size_t)G.n_limbs * JD;
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= (size_t)O * pw) return;
    const size_t o = i / pw, slot = i % pw;
    const u64 q = G.q[slot / JD];
    u64 acc = 0;
    for (int c = 0; c < C; c++) acc = addmod(acc, part[((size_t)c * O + o) * pw + slot], q);
    out[i] = acc;
}

// out[i] = sum_s part[s][i] mod q   (split-K reduction of matvec_pipe_kernel; out and part share the [C][O] polynomial layout)
__global__ void __launch_bounds__(256) reduce_split_kernel(u64 *__restrict__ out, const u64 *__restrict__ part, int S, size_t n,
                                                           const __grid_constant__ MatvecGeom G) {
    const size_t pw = (size_t)G.n_limbs * JD;
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const u64 q = G.q[(i % pw) / JD];
    u64 acc = 0;
    for (int s = 0; s < S; s++) acc = addmod(acc, part[(size_t)s * n + i], q);
    out[i] = acc;
}

// ---------------------------------------------------------------------------------------------
// cut-off pipeline (prover.go:164-176, :190-200; rns.go:76-114)
// ---------------------------------------------------------------------------------------------
struct CutConsts {
    int log_cut;
    RingDev in, outr;
    CrtDev crt;
};

__device__ __forceinline__ bool mw_geq(const u64 *a, const u64 *b) {
    for (int i = JW - 1; i >= 0; i--) { if (a[i] > b[i]) return true; if (a[i] < b[i]) return false; }
    return true;
}
__device__ __forceinline__ void mw_sub(u64 *a, const u64 *b) {
    u64 br = 0;
    for (int i = 0; i < JW; i++) {
        const u64 d = a[i] - b[i], d2 = d - br;
        br = (a[i] < b[i]) | (d < br);
        a[i] = d2;
    }
}

// Verifier side (jindo/verifier.go:100-113, :262-282): with ints_out / sq_out set the kernel also hands back the reconstructed
// centred integers (JW two's-complement words per coefficient) and, per polynomial, the sum of their squares (2 JW words,
// what verifyNorm accumulates); out == NULL stops after the reconstruction.
__global__ void __launch_bounds__(256) cut_kernel(u64 *__restrict__ out, const u64 *__restrict__ in,
                                                  const __grid_constant__ CutConsts K, u64 *__restrict__ ints_out, u64 *__restrict__ sq_out) {
    __shared__ u64 s[JMAXL][JD];
    const int c = threadIdx.x;
    const RingDev &A = K.in, &B = K.outr;
    const u64 *src = in + (size_t)blockIdx.x * A.n * JD;
    for (int l = 0; l < A.n; l++) s[l][c] = src[(size_t)l * JD + c];
    __syncthreads();
    // IMForm + INTT (plain canonical coefficients)
    for (int l0 = 0; l0 < A.n; l0 += 2) {
        const int l = l0 + (c >> 7);
        if (l < A.n) intt256(s[l], A.roots_inv + (size_t)l * JD, A.q[l], A.qinvneg[l], A.ninv[l], c & 127, 1 + (c >> 7), 128);
    }
    __syncthreads();
    // reconstructTo (rns.go:76-105) -> centred integer V (two's complement, JW words)
    u64 V[JW];
    bool neg = false;
    {
        bool small = true;
        const u64 x0 = s[0][c];
        const i64 bal0 = (x0 > (A.q[0] >> 1)) ? (i64)x0 - (i64)A.q[0] : (i64)x0;       // toBalanced rns.go:68-73
        for (int l = 1; l < A.n; l++) {
            const u64 xl = s[l][c];
            const i64 bl = (xl > (A.q[l] >> 1)) ? (i64)xl - (i64)A.q[l] : (i64)xl;
            if (bl != bal0) { small = false; break; }
        }
        if (small) {
            neg = bal0 < 0;
            V[0] = (u64)bal0;
            for (int i = 1; i < JW; i++) V[i] = neg ? ~0ull : 0ull;
        } else {
            // acc = sum_l c_l * gad_l mod Q, computed as sum_l y_l * (Q/q_l) with y_l = c_l (Q/q_l)^-1 mod q_l
            for (int i = 0; i < JW; i++) V[i] = 0;
            for (int l = 0; l < A.n; l++) {
                const u64 y = mont64(s[l][c], K.crt.inv_mont[l], A.q[l], A.qinvneg[l]);
                u64 carry = 0;
                for (int i = 0; i < JW; i++) {
                    const u64 pl = y * K.crt.Qdiv[l][i], ph = __umul64hi(y, K.crt.Qdiv[l][i]);
                    u64 t = V[i] + pl;
                    u64 cy = t < pl;
                    const u64 t2 = t + carry;
                    cy += t2 < t;
                    V[i] = t2;
                    carry = ph + cy;
                }
            }
            for (int it = 0; it < JMAXL && mw_geq(V, K.crt.Q); it++) mw_sub(V, K.crt.Q);
            if (mw_geq(V, K.crt.Qhalf)) { mw_sub(V, K.crt.Q); neg = true; }             // acc >= Q>>1 -> acc - Q
        }
    }
    if (ints_out) {
        u64 *dst = ints_out + ((size_t)blockIdx.x * JD + c) * JW;
        for (int i = 0; i < JW; i++) dst[i] = V[i];
    }
    if (sq_out) {
        __shared__ u64 s_sq[JD][2 * JW + 1];
        u64 M[JW];
        if (neg) { u64 cy = 1; for (int i = 0; i < JW; i++) { const u64 t = ~V[i] + cy; cy = (cy && t == 0); M[i] = t; } }
        else for (int i = 0; i < JW; i++) M[i] = V[i];
        u64 P[2 * JW];
        for (int i = 0; i < 2 * JW; i++) P[i] = 0;
        for (int a = 0; a < JW; a++) {
            u64 carry = 0;
            for (int b = 0; b < JW; b++) {
                const u64 lo = M[a] * M[b], hi = __umul64hi(M[a], M[b]);
                u64 t = P[a + b] + lo; u64 cy = t < lo;
                const u64 t2 = t + carry; cy += t2 < t;
                P[a + b] = t2;
                carry = hi + cy;
            }
            P[a + JW] += carry;
        }
        for (int i = 0; i < 2 * JW; i++) s_sq[c][i] = P[i];
        __syncthreads();
        for (int st = JD / 2; st >= 1; st >>= 1) {
            if (c < st) {
                u64 cy = 0;
                for (int i = 0; i < 2 * JW; i++) {
                    const u64 a = s_sq[c][i], b = s_sq[c + st][i];
                    const u64 t = a + b; u64 c1 = t < b;
                    const u64 t2 = t + cy; c1 += t2 < t;
                    s_sq[c][i] = t2; cy = c1;
                }
            }
            __syncthreads();
        }
        if (c < 2 * JW) sq_out[(size_t)blockIdx.x * 2 * JW + c] = s_sq[0][c];
    }
    if (!out) return;
    // Rsh(log_cut): arithmetic shift (floor for negatives, like big.Int.Rsh)
    {
        const int ws = K.log_cut / 64, bs = K.log_cut % 64;
        const u64 fill = neg ? ~0ull : 0ull;
        u64 Wv[JW];
        for (int i = 0; i < JW; i++) {
            const u64 a = (i + ws < JW) ? V[i + ws] : fill;
            const u64 b = (i + ws + 1 < JW) ? V[i + ws + 1] : fill;
            Wv[i] = bs ? ((a >> bs) | (b << (64 - bs))) : a;
        }
        for (int i = 0; i < JW; i++) V[i] = Wv[i];
    }
    // setBigCoeffTo (rns.go:108-114): Euclidean residue mod each outer prime
    {
        u64 M[JW];                                                                      // |V|
        if (neg) {
            u64 cy = 1;
            for (int i = 0; i < JW; i++) { const u64 t = ~V[i] + cy; cy = (cy && t == 0); M[i] = t; }
        } else {
            for (int i = 0; i < JW; i++) M[i] = V[i];
        }
        __syncthreads();
        for (int l = 0; l < B.n; l++) {
            const u64 q = B.q[l];
            u64 r = 0;
            // Horner over 64-bit words: r <- r * 2^64 + M[i] (mod q); r * 2^64 = mont64(r, 2^128 mod q)
            for (int i = JW - 1; i >= 0; i--) r = addmod(mont64(r, B.r2[l], q, B.qinvneg[l]), M[i] % q, q);
            if (neg && r) r = q - r;
            s[l][c] = mont64(r, B.r2[l], q, B.qinvneg[l]);                              // MForm
        }
    }
    __syncthreads();
    for (int l0 = 0; l0 < B.n; l0 += 2) {
        const int l = l0 + (c >> 7);
        if (l < B.n) ntt256(s[l], B.roots + (size_t)l * JD, B.q[l], B.qinvneg[l], c & 127, 1 + (c >> 7), 128);
    }
    __syncthreads();
    u64 *dst = out + (size_t)blockIdx.x * B.n * JD;
    for (int l = 0; l < B.n; l++) dst[(size_t)l * JD + c] = s[l][c];
}

// acc[p] += o[p] (.) chal   for n_polys polynomials (Evaluate batch combine, prover.go:251-265)
__global__ void __launch_bounds__(256) muladd_bcast_kernel(u64 *__restrict__ acc, const u64 *__restrict__ o,
                                                           const u64 *__restrict__ chal, size_t n_polys, int n_limbs,
                                                           const __grid_constant__ MatvecGeom G) {
    const size_t pw = (size_t)n_limbs * JD;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n_polys * pw; i += (size_t)gridDim.x * blockDim.x) {
        const size_t slot = i % pw;
        const int limb = (int)(slot / JD);
        const u64 q = G.q[limb];
        acc[i] = addmod(acc[i], mont64(o[i], __ldg(chal + slot), q, G.qinvneg[limb]), q);
    }
}

// acc[p] = sum_k srcs[k][p] (.) chal[k]   -- the whole batch combination of Evaluate (prover.go:251-265) in one pass: every
// opening is read once and the accumulator written once (the reference adds the products one opening at a time; the sum of
// canonical residues mod q does not depend on the order)
__global__ void __launch_bounds__(256) combine_kernel(u64 *__restrict__ acc, const u64 *const *__restrict__ srcs, const u64 *__restrict__ chals,
                                                      int n, size_t n_polys, int n_limbs, const __grid_constant__ MatvecGeom G) {
    const size_t pw = (size_t)n_limbs * JD;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n_polys * pw; i += (size_t)gridDim.x * blockDim.x) {
        const size_t slot = i % pw;
        const int limb = (int)(slot / JD);
        const u64 q = G.q[limb], qi = G.qinvneg[limb];
        u64 a = 0;
        for (int k = 0; k < n; k++) a = addmod(a, mont64(__ldg(srcs[k] + i), __ldg(chals + (size_t)k * pw + slot), q, qi), q);
        acc[i] = a;
    }
}

// out = a * s - b [- c]  (mod q), s a per-limb Montgomery scalar: the verifier's  commitment * 2^cut - key (.) proof  lines
// (MulRNSScalarMontgomery + MulCoeffsMontgomeryThenSub + Sub, verifier.go:146-157, :183-196)
__global__ void __launch_bounds__(256) scale_sub_kernel(u64 *__restrict__ out, const u64 *__restrict__ a, const u64 *__restrict__ b,
                                                        const u64 *__restrict__ c, size_t n_polys, const __grid_constant__ MatvecGeom G,
                                                        const __grid_constant__ MatvecGeom S /* S.q[l] = scalar in Montgomery form */) {
    const size_t pw = (size_t)G.n_limbs * JD;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n_polys * pw; i += (size_t)gridDim.x * blockDim.x) {
        const int limb = (int)((i % pw) / JD);
        const u64 q = G.q[limb];
        u64 r = mont64(a[i], S.q[limb], q, G.qinvneg[limb]);
        r = submod(r, b[i], q);
        if (c) r = submod(r, c[i], q);
        out[i] = r;
    }
}
// flag |= (a - b - c != 0) anywhere
__global__ void __launch_bounds__(256) nonzero_kernel(unsigned *__restrict__ flag, const u64 *__restrict__ a, const u64 *__restrict__ b,
                                                      const u64 *__restrict__ c, size_t n, const __grid_constant__ MatvecGeom G) {
    const size_t pw = (size_t)G.n_limbs * JD;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const u64 q = G.q[(i % pw) / JD];
        if (submod(submod(a[i], b[i], q), c[i], q) != 0) atomicOr(flag, 1u);
    }
}

}  // namespace rs

// =============================================================================================
// host side
// =============================================================================================
using namespace rs;
typedef unsigned __int128 u128;

static u64 h_mulmod(u64 a, u64 b, u64 q) { return (u64)(((u128)a * b) % q); }
static u64 h_powmod(u64 a, u64 e, u64 q) {
    u64 r = 1 % q;
    a %= q;
    while (e) { if (e & 1) r = h_mulmod(r, a, q); a = h_mulmod(a, a, q); e >>= 1; }
    return r;
}
static u64 h_brv(u64 i, int bits) { u64 r = 0; for (int k = 0; k < bits; k++) { r = (r << 1) | (i & 1); i >>= 1; } return r; }

// multi-word helpers (JW words)
static void mw_mul_word(u64 *a, u64 w) {
    u64 carry = 0;
    for (int i = 0; i < JW; i++) { u128 t = (u128)a[i] * w + carry; a[i] = (u64)t; carry = (u64)(t >> 64); }
}
static u64 mw_mod_word(const u64 *a, u64 q) {
    u128 r = 0;
    for (int i = JW - 1; i >= 0; i--) r = ((r << 64) | a[i]) % q;
    return (u64)r;
}

static int ring_setup(RingDev &R, CrtDev &C, int n, const u64 *q, const u64 *psi, u64 *h_roots /* 5*n*JD */) {
    memset(&R, 0, sizeof(R));
    memset(&C, 0, sizeof(C));
    R.n = n;
    R.f64_ok = getenv("RS_JINDO_NOF64") ? 0 : 1;
    R.no_rowsum = getenv("RS_MV_NOROWSUM") ? 1 : 0; R.no_pipe = getenv("RS_MV_NOPIPE") ? 1 : 0; R.no_split = getenv("RS_MV_NOSPLIT") ? 1 : 0;
    R.no_small = getenv("RS_MV_NOSMALL") ? 1 : 0;
    { const char *e = getenv("RS_MV_TMA"); R.use_tma = e ? atoi(e) : 0; }
    for (int l = 0; l < n; l++) {
        const u64 ql = q[l];
        if (ql >= (1ull << 44)) R.f64_ok = 0;
        R.qf[l] = (double)ql; R.qinvf[l] = 1.0 / (double)ql;
        if (!(ql & 1) || ql >= (1ull << 62) || ql < (1ull << 33)) return set_error(RS_ERR_UNSUPPORTED, "jindo: RNS prime out of the supported range (2^33, 2^62)");
        R.q[l] = ql;
        u64 inv = 1;
        for (int i = 0; i < 6; i++) inv *= 2 - ql * inv;
        R.qinvneg[l] = 0 - inv;
        const u64 r1 = (u64)((((u128)1) << 64) % ql);
        R.r2[l] = h_mulmod(r1, r1, ql);
        R.r1[l] = r1;
        R.r1s[l] = (u64)((((u128)r1) << 64) / ql);
        R.ninv[l] = h_powmod(JD, ql - 2, ql);
        if (h_powmod(psi[l], JD, ql) != ql - 1) return set_error(RS_ERR_ARG, "jindo: psi is not a primitive 2d-th root of unity");
        const u64 psi_inv = h_powmod(psi[l], ql - 2, ql);
        u64 pw = 1, pwi = 1;
        for (int j = 0; j < JD; j++) {
            const u64 b = h_brv(j, 8);
            h_roots[(size_t)l * JD + b] = h_mulmod(pw, r1, ql);
            h_roots[(size_t)(n + l) * JD + b] = h_mulmod(pwi, r1, ql);
            h_roots[(size_t)(2 * n + l) * JD + b] = pw;
            h_roots[(size_t)(3 * n + l) * JD + b] = (u64)((((u128)pw) << 64) / ql);
            { const double pf = (double)pw; memcpy(&h_roots[(size_t)(4 * n + l) * JD + b], &pf, 8); }
            pw = h_mulmod(pw, psi[l], ql);
            pwi = h_mulmod(pwi, psi_inv, ql);
        }
    }
    // CRT constants (rns.go:33-55)
    C.Q[0] = 1;
    for (int l = 0; l < n; l++) mw_mul_word(C.Q, q[l]);
    for (int i = 0; i < JW; i++) C.Qhalf[i] = (C.Q[i] >> 1) | (i + 1 < JW ? (C.Q[i + 1] << 63) : 0);
    for (int l = 0; l < n; l++) {
        C.Qdiv[l][0] = 1;
        for (int k = 0; k < n; k++) if (k != l) mw_mul_word(C.Qdiv[l], q[k]);
        const u64 dm = mw_mod_word(C.Qdiv[l], q[l]);
        const u64 r1 = (u64)((((u128)1) << 64) % q[l]);
        C.inv_mont[l] = h_mulmod(h_powmod(dm, q[l] - 2, q[l]), r1, q[l]);
    }
    return RS_OK;
}

static void fill_geom_ring(MatvecGeom &G, const RingDev &R) {
    G.n_limbs = R.n;
    for (int l = 0; l < R.n; l++) { G.q[l] = R.q[l]; G.qinvneg[l] = R.qinvneg[l]; G.r2[l] = R.r2[l]; }
    // S1 receives two terms below 2^(32+hb) per product (hb = bits-32): it can take 2^(31-hb) products
    int bits = 0;
    for (int l = 0; l < R.n; l++) { int b = 64 - __builtin_clzll(R.q[l]); if (b > bits) bits = b; }
    const int hb = bits > 32 ? bits - 32 : 0;
    const long long cap = 1ll << (31 - hb);
    G.fold_every = (G.K < cap) ? 0 : (int)(cap > 1 ? cap - 1 : 1);
}

// ---------------------------------------------------------------------------------------------
// rowsum_kernel: the O == 1 products of Evaluate (Partial: prover.go:281-289, responses: :304-316):
//   out[cc] (+ addend[cc]) = sum_k key[k] (.) x[cc * x_cs + k * x_ks]
// A pure stream over the Opening: one thread per slot of one output polynomial, 8 rows in flight per thread,
// the K range optionally split over blockIdx.z into partial sums (reduced mod q, added up by
// rowsum_reduce_kernel).  Same lazy accumulation and final REDC as matvec_kernel.
// ---------------------------------------------------------------------------------------------
// V: slots per thread (2: 16-byte loads, half the load instructions and 512-byte warp requests -- the kernel is a latency-bound
// stream, and wider requests carry more bytes per outstanding load)
template <bool FOLD, int V = 1>
__global__ void __launch_bounds__(256) rowsum_kernel(u64 *__restrict__ out, const u64 *__restrict__ key, const u64 *__restrict__ x,
                                                     const u64 *__restrict__ addend, const __grid_constant__ MatvecGeom G, int ksplit) {
    const int slot = (blockIdx.x * 256 + threadIdx.x) * V;
    const int cc = blockIdx.y, part = blockIdx.z;
    const int n_slots = G.n_limbs * JD;
    if (slot >= n_slots) return;
    const int limb = slot / JD;                           // (V == 2: JD is even, both slots sit in the same limb)
    const u64 q = G.q[limb], qi = G.qinvneg[limb];
    const size_t pw = (size_t)n_slots;
    const int per = (G.K + ksplit - 1) / ksplit;
    const int k0 = part * per, k1 = (k0 + per < G.K) ? k0 + per : G.K;
    uint32_t s0a[V], s0b[V], s0c[V];
    Acc64 s1[V], s2[V]; u64 s1h[V], s2h[V];
#pragma unroll
    for (int v = 0; v < V; v++) { s0a[v] = s0b[v] = s0c[v] = 0; s1[v] = s2[v] = Acc64{0}; s1h[v] = s2h[v] = 0; }
    int since = 0;
    const u64 *xp = x + ((size_t)cc * G.x_cs + (size_t)k0 * G.x_ks) * pw + slot;
    const u64 *kp = key + (size_t)k0 * G.key_ks * pw + slot;
    const size_t x_step = (size_t)G.x_ks * pw, k_step = (size_t)G.key_ks * pw;
    constexpr int U = 8;
    for (int k = k0; k < k1; k += U) {
        u64 kv[U][V], xv[U][V];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const bool ok = k + u < k1;
            if (V == 2) {
                const ulonglong2 a = ok ? __ldg(reinterpret_cast<const ulonglong2 *>(xp + (size_t)u * x_step)) : make_ulonglong2(0, 0);
                const ulonglong2 b = ok ? __ldg(reinterpret_cast<const ulonglong2 *>(kp + (size_t)u * k_step)) : make_ulonglong2(0, 0);
                xv[u][0] = a.x; xv[u][V - 1] = a.y; kv[u][0] = b.x; kv[u][V - 1] = b.y;
            } else {
                xv[u][0] = ok ? __ldg(xp + (size_t)u * x_step) : 0;
                kv[u][0] = ok ? __ldg(kp + (size_t)u * k_step) : 0;
            }
        }
        xp += (size_t)U * x_step; kp += (size_t)U * k_step;
#pragma unroll
        for (int u = 0; u < U; u++) {
#pragma unroll
            for (int v = 0; v < V; v++) mac_lazy(s0a[v], s0b[v], s0c[v], s1[v], s2[v], kv[u][v], xv[u][v]);
            if (FOLD && ++since == G.fold_every) {
                since = 0;
#pragma unroll
                for (int v = 0; v < V; v++) {
                    s1h[v] += s1[v].v >> 32; s1[v].v &= 0xffffffffull;
                    s2h[v] += s2[v].v >> 32; s2[v].v &= 0xffffffffull;
                }
            }
        }
    }
#pragma unroll
    for (int v = 0; v < V; v++) {
        u64 w0 = ((u64)s0b[v] << 32) | s0a[v], w1 = s0c[v], w2 = 0;
        auto add_at32 = [&](u64 t) {
            const u64 lo = t << 32, hi = t >> 32;
            const u64 n0 = w0 + lo; const u64 cy = n0 < lo; w0 = n0;
            const u64 n1 = w1 + hi; u64 cy2 = n1 < hi; const u64 n1b = n1 + cy; cy2 += n1b < n1; w1 = n1b; w2 += cy2;
        };
        auto add_at64 = [&](u64 t) { const u64 n1 = w1 + t; w2 += n1 < t; w1 = n1; };
        add_at32(s1[v].value());
        add_at64(s2[v].value());
        if (FOLD) { add_at64(s1h[v]); const u64 t = s2h[v]; add_at64(t << 32); w2 += t >> 32; }
        u64 h = w1 >= q ? w1 % q : w1;
        if (w2) h = addmod(h, mont64(w2 >= q ? w2 % q : w2, G.r2[limb], q, qi), q);
        u64 r = redc128(h, w0, q, qi);
        if (ksplit == 1 && G.has_addend) r = addmod(r, __ldg(addend + (size_t)cc * G.add_cs * pw + slot + v), q);
        out[((size_t)part * G.C + cc) * pw + slot + v] = r;       // ksplit == 1: out[cc]
    }
}

__global__ void rowsum_reduce_kernel(u64 *__restrict__ out, const u64 *__restrict__ parts, const u64 *__restrict__ addend,
                                     const __grid_constant__ MatvecGeom G, int ksplit) {
    const size_t pw = (size_t)G.n_limbs * JD;
    const size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (idx >= (size_t)G.C * pw) return;
    const int slot = (int)(idx % pw), cc = (int)(idx / pw);
    const u64 q = G.q[slot / JD];
    u64 r = 0;
    for (int s = 0; s < ksplit; s++) r = addmod(r, parts[(size_t)s * G.C * pw + idx], q);
    if (G.has_addend) r = addmod(r, __ldg(addend + (size_t)cc * G.add_cs * pw + slot), q);
    out[idx] = r;
}

static int launch_matvec(u64 *out, const u64 *key, const u64 *x, const u64 *addend, MatvecGeom G, const RingDev &R) {
    fill_geom_ring(G, R);
    const int n_slots = R.n * JD;
    // single-output products (Evaluate): streaming row sums
    if (G.O == 1 && G.key_cs == 0 && G.out_cs == 1 && !G.accumulate && !R.no_rowsum) {
        // enough CTAs for >= 4 per SM; K is split when there are few output polynomials
        // two slots per thread (16-byte loads) when the operands allow it
        static const bool v1_only = getenv("RS_JINDO_ROWSUM_V1") != nullptr;
        const bool v2 = !v1_only && n_slots % 2 == 0 && ((((uintptr_t)key | (uintptr_t)x) & 15) == 0);
        const int bx = v2 ? (n_slots / 2 + 255) / 256 : (n_slots + 255) / 256;
        int ksplit = 1;
        while ((long long)bx * G.C * ksplit < 148 * 4 && ksplit * 64 < G.K) ksplit *= 2;
        u64 *d_parts = nullptr;
        if (ksplit > 1) RS_CUDA_CHECK(cudaMallocAsync((void **)&d_parts, (size_t)ksplit * G.C * n_slots * 8, current_stream()));
        dim3 grid(bx, G.C, ksplit);
        u64 *dst = ksplit > 1 ? d_parts : out;
        if (v2) {
            if (G.fold_every) rowsum_kernel<true, 2><<<grid, 256, 0, current_stream()>>>(dst, key, x, addend, G, ksplit);
            else rowsum_kernel<false, 2><<<grid, 256, 0, current_stream()>>>(dst, key, x, addend, G, ksplit);
        } else if (G.fold_every) rowsum_kernel<true><<<grid, 256, 0, current_stream()>>>(dst, key, x, addend, G, ksplit);
        else rowsum_kernel<false><<<grid, 256, 0, current_stream()>>>(dst, key, x, addend, G, ksplit);
        count_launch();
        if (ksplit > 1) {
            const size_t tot = (size_t)G.C * n_slots;
            rowsum_reduce_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, current_stream()>>>(out, d_parts, addend, G, ksplit);
            count_launch();
            cudaFreeAsync(d_parts, current_stream());
        }
        RS_CUDA_CHECK(cudaGetLastError());
        return RS_OK;
    }
    // large products (inner commit): shared-memory ring version
    // (also a single column with a long K -- the verifier's ck.In x Encode product: split-K fills the chip, the register-tile kernel
    //  below would walk K = rows sequentially in 160 CTAs)
    if (G.key_cs == 0 && ((long long)G.K * G.C >= 4096 || (G.C == 1 && G.K >= 1024)) && !R.no_pipe &&
        (unsigned long long)G.key_ks * n_slots * 8 < (1ull << 32) && (unsigned long long)G.x_ks * n_slots * 8 < (1ull << 32)) {
        RS_CUDA_CHECK(ensure_smem_attr((const void *)matvec_pipe_kernel<true, false>, MP_SMEM_BYTES));
        RS_CUDA_CHECK(ensure_smem_attr((const void *)matvec_pipe_kernel<false, false>, MP_SMEM_BYTES));
        RS_CUDA_CHECK(ensure_smem_attr((const void *)matvec_pipe_kernel<true, true>, MP_SMEM_BYTES));
        RS_CUDA_CHECK(ensure_smem_attr((const void *)matvec_pipe_kernel<false, true>, MP_SMEM_BYTES));
        RS_CUDA_CHECK(ensure_smem_attr((const void *)matvec_pipe_kernel<false, false, true>, MP_SMEM_BYTES));
        RS_CUDA_CHECK(ensure_smem_attr((const void *)matvec_pipe_kernel<false, false, false, true>, MP_SMEM_BYTES));
        // k1 x1 < 2^(2 hb) per term (hb = prime bits - 32): a 32-bit accumulator takes K terms while K 2^(2 hb) < 2^32
        bool s2_narrow = false, kara = false;
        {
            int bits = 0;
            for (int l = 0; l < G.n_limbs; l++) { int b = 64 - __builtin_clzll(G.q[l]); if (b > bits) bits = b; }
            const int hb = bits > 32 ? bits - 32 : 0;
            static const bool off = getenv("RS_JINDO_NO_S2N") != nullptr;
            s2_narrow = !off && !G.fold_every && 2 * hb < 32 && (unsigned long long)G.K < (1ull << (32 - 2 * hb));
            // Karatsuba split at sh = ceil(bits / 2): three IMAD.WIDE per multiply-accumulate, valid while K 2^(2 sh + 2) < 2^64
            static const bool kara_on = getenv("RS_JINDO_NO_KARA") == nullptr;
            const int sh = (bits + 1) / 2;
            kara = kara_on && !G.fold_every && sh >= 1 && sh <= 30 && (unsigned long long)G.K < (1ull << (62 - 2 * sh));
            G.kara_sh = kara ? sh : 0;
        }
        dim3 grid(n_slots / MP_MS, (G.O + MP_OT - 1) / MP_OT, (G.C + MP_CTC - 1) / MP_CTC);
        // few column tiles (a column shard of a Commit, a short Commit): cut K so that the grid still covers the chip
        const unsigned ctas = grid.x * grid.y * grid.z;
        int S = 1;
        if (ctas < 148 && !G.accumulate && !G.has_addend && G.out_os == 1 && G.out_cs == G.O && !R.no_split) {
            // one 512-thread CTA per SM: the launch lasts ceil(ctas S / 148) waves of K / S steps.  Pick the S with the shortest
            // launch (80 CTAs: S = 1 and S = 2 both take K steps, S = 3 two waves of K / 3), smaller S on near ties.
            double best = 1.0;
            for (int s = 2; s <= 16 && G.K / s >= 64; s++) {
                const double cost = (double)((ctas * s + 147) / 148) / s;
                if (cost < best * 0.97) { best = cost; S = s; }
            }
        }
        u64 *d_parts = nullptr;
        const size_t n_out = (size_t)G.C * G.O * n_slots;
        if (S > 1) {
            if (cudaMallocAsync((void **)&d_parts, (size_t)S * n_out * 8, current_stream()) != cudaSuccess)
                return set_error(RS_ERR_CUDA, "matvec: cudaMallocAsync failed");
            G.k_split = S; G.split_stride = (i64)n_out;
            grid.y *= S;
        }
        u64 *dst = S > 1 ? d_parts : out;
        // rows must be 16-byte aligned 512-byte segments for the bulk copies (they are: polynomial strides are multiples of 2 KiB)
        const bool tma = R.use_tma && (((uintptr_t)key | (uintptr_t)x) & 15) == 0;
        if (tma) {
            if (G.fold_every) matvec_pipe_kernel<true, true><<<grid, MP_THREADS, MP_SMEM_BYTES, current_stream()>>>(dst, key, x, addend, G);
            else matvec_pipe_kernel<false, true><<<grid, MP_THREADS, MP_SMEM_BYTES, current_stream()>>>(dst, key, x, addend, G);
        } else if (G.fold_every) matvec_pipe_kernel<true, false><<<grid, MP_THREADS, MP_SMEM_BYTES, current_stream()>>>(dst, key, x, addend, G);
        else if (kara) matvec_pipe_kernel<false, false, false, true><<<grid, MP_THREADS, MP_SMEM_BYTES, current_stream()>>>(dst, key, x, addend, G);
        else if (s2_narrow) matvec_pipe_kernel<false, false, true><<<grid, MP_THREADS, MP_SMEM_BYTES, current_stream()>>>(dst, key, x, addend, G);
        else matvec_pipe_kernel<false, false><<<grid, MP_THREADS, MP_SMEM_BYTES, current_stream()>>>(dst, key, x, addend, G);
        count_launch();
        if (S > 1) {
            reduce_split_kernel<<<(unsigned)((n_out + 255) / 256), 256, 0, current_stream()>>>(out, d_parts, S, n_out, G);
            count_launch();
            cudaFreeAsync(d_parts, current_stream());
        }
        RS_CUDA_CHECK(cudaGetLastError());
        return RS_OK;
    }
#define RS_MV(OT, CT, WARPS)                                                                                         \
    {                                                                                                                \
        dim3 grid((n_slots + 31) / 32, (G.O + OT - 1) / OT, (G.C + CT * WARPS - 1) / (CT * WARPS));                   \
        if (G.fold_every) matvec_kernel<OT, CT, WARPS, true><<<grid, dim3(32, WARPS), 0, current_stream()>>>(out, key, x, addend, G); \
        else matvec_kernel<OT, CT, WARPS, false><<<grid, dim3(32, WARPS), 0, current_stream()>>>(out, key, x, addend, G);             \
    }
    // a small product (the Commits of a Buckler proof: K*C in the hundreds) would leave two thirds of the SMs without a CTA
    // with the 5 x 2 register tile and spends its time waiting for loads: one output per thread, eight loads in flight
    if ((long long)((n_slots + 31) / 32) * ((G.O + 4) / 5) * ((G.C + 7) / 8) < 148 * 2 && !R.no_small) {
        dim3 grid((n_slots + 31) / 32, G.O, (G.C + 3) / 4);
        if (G.fold_every) matvec_kernel<1, 1, 4, true, 8><<<grid, dim3(32, 4), 0, current_stream()>>>(out, key, x, addend, G);
        else matvec_kernel<1, 1, 4, false, 8><<<grid, dim3(32, 4), 0, current_stream()>>>(out, key, x, addend, G);
    } else
        RS_MV(5, 2, 4)
#undef RS_MV
    count_launch();
    RS_CUDA_CHECK(cudaGetLastError());
    return RS_OK;
}

template <int NL>
static int launch_encode(rs_jindo *j, u64 *out, const EncDesc *d_descs, size_t n, const uint32_t *vec, const uint32_t *aux,
                         const u64 *twin_u, const i64 *given) {
    if (n == 0) return RS_OK;
    EncConsts E;
    memset(&E, 0, sizeof(E));
    E.slots = j->p.slots; E.exp = j->p.exp; E.cdt_len = j->p.cdt_len;
    E.base = j->p.base; E.base_magic = (u64)((((u128)1) << 64) / j->p.base) + 1;
    E.tail_lo = j->p.cdt_tail_lo; E.sigma = j->p.cdt_sigma;
    E.n_delta = 0;
    for (int i = 0; i < j->p.exp; i++)
        if (j->p.delta_inv[i] != 0.0) { E.delta_inv[E.n_delta] = j->p.delta_inv[i]; E.delta_w[E.n_delta] = (i + 1) * j->p.slots; E.n_delta++; }
    for (int l = 0; l < j->rq.n; l++) E.base_mont[l] = j->base_mont_q[l];
    E.g_lim = (i64)((1ull << 62) / j->p.base);
    if (j->p.base < 65536 && !j->no_pair) {
        E.base2 = (uint32_t)(j->p.base * j->p.base);
        E.base2_magic = (u64)((((u128)1) << 64) / E.base2) + 1;
        E.base_magic32 = (uint32_t)((1ull << 32) / j->p.base) + 1;
    }
    FieldParams<NL> F = make_params<NL>(j->f);
    int G = JD / j->p.slots;
    if (G > 16) G = 16;
    // small commits (the 12-43 per Buckler proof: 650-2300 encodes): a CTA walks its G polynomials in sequence, so with
    // fewer than ~4 CTAs per SM the launch lasts as long as one CTA.  Shrink G until the grid covers the chip.
    while (G > 1 && (n + G - 1) / G < (size_t)148 * 4) G >>= 1;
    E.group = G;
    E.subgroup = 8 / j->rq.n > 0 ? 8 / j->rq.n : 1;
    if (E.subgroup > G) E.subgroup = G;
    E.cdt_pow2 = 1;
    while (E.cdt_pow2 * 2 <= E.cdt_len) E.cdt_pow2 *= 2;
    const size_t smem = (size_t)G * JD * 4 + (size_t)E.subgroup * JD * 8 + (size_t)E.subgroup * j->rq.n * NTTW_STRIDE * 8;
    RS_CUDA_CHECK(ensure_smem_attr((const void *)encode_kernel<NL>, 200 * 1024));
    encode_kernel<NL><<<(unsigned)((n + G - 1) / G), 256, smem, current_stream()>>>(out, d_descs, (int)n, vec, aux, twin_u, given,
                                                                                   j->d_cdt, j->d_cdt_guide, j->d_slow_lb, j->rq, E, F);
    count_launch();
    RS_CUDA_CHECK(cudaGetLastError());
    return RS_OK;
}

static int encode_dispatch(rs_jindo *j, u64 *out, const EncDesc *d, size_t n, const uint32_t *vec, const uint32_t *aux,
                           const u64 *tw, const i64 *gv) {
    switch (j->f->NL) {
        case 4: return launch_encode<4>(j, out, d, n, vec, aux, tw, gv);
        case 8: return launch_encode<8>(j, out, d, n, vec, aux, tw, gv);
        case 14: return launch_encode<14>(j, out, d, n, vec, aux, tw, gv);
        case 28: return launch_encode<28>(j, out, d, n, vec, aux, tw, gv);
    }
    return set_error(RS_ERR_UNSUPPORTED, "unsupported limb count");
}

static int launch_cut(rs_jindo *j, u64 *out, const u64 *in, size_t n_polys, bool outer) {
    CutConsts K;
    K.log_cut = outer ? j->p.log_out_cut : j->p.log_in_cut;
    K.in = outer ? j->rqo : j->rq;
    K.outr = j->rqo;
    K.crt = outer ? j->crt_qo : j->crt_q;
    cut_kernel<<<(unsigned)n_polys, 256, 0, current_stream()>>>(out, in, K, nullptr, nullptr);
    count_launch();
    RS_CUDA_CHECK(cudaGetLastError());
    return RS_OK;
}

// verifier forms: src ring -> (optional) re-RNS into dst ring with no shift (BasisExtender.ModUpQtoP of a centred lift),
// (optional) centred integers, (optional) per-polynomial sums of squares
static int launch_recon(rs_jindo *j, bool src_out, bool dst_q, u64 *out, const u64 *in, size_t n_polys, u64 *ints_out, u64 *sq_out) {
    if (n_polys == 0) return RS_OK;
    CutConsts K;
    K.log_cut = 0;
    K.in = src_out ? j->rqo : j->rq;
    K.outr = dst_q ? j->rq : j->rqo;
    K.crt = src_out ? j->crt_qo : j->crt_q;
    cut_kernel<<<(unsigned)n_polys, 256, 0, current_stream()>>>(out, in, K, ints_out, sq_out);
    count_launch();
    RS_CUDA_CHECK(cudaGetLastError());
    return RS_OK;
}

static size_t pq(const rs_jindo *j) { return (size_t)j->rq.n * JD; }       // words per ringQ poly
static size_t pqo(const rs_jindo *j) { return (size_t)j->rqo.n * JD; }

template <typename T>
static int upload(T **d, const T *h, size_t count) {
    *d = nullptr;
    RS_CUDA_CHECK(cudaMallocAsync((void **)d, (count ? count : 1) * sizeof(T), current_stream()));
    if (count) RS_CUDA_CHECK(cudaMemcpyAsync(*d, h, count * sizeof(T), cudaMemcpyHostToDevice, current_stream()));
    return RS_OK;
}

static int opening_alloc(rs_jindo *j, rs_opening **out, bool zero = true) {
    rs_opening *o = new rs_opening{j, nullptr, nullptr, nullptr};
    const rs_jindo_params &p = j->p;
    const size_t ne = (size_t)(p.cols + 1) * p.rows * pq(j), nm = (size_t)(p.cols + 1) * (p.mlwe + p.in_msis) * pq(j),
                 ni = (size_t)(p.cols + 1 + RS_MAX_SHARDS) * p.in_msis * pqo(j);   // + slack: padded in-place all-gather of column shards
    // stream-ordered pool (rs_init keeps it cached): a Buckler proof allocates and frees 12-43 Openings
    cudaError_t e = cudaMallocAsync((void **)&o->d_encode, ne * 8, current_stream());
    if (e == cudaSuccess) e = cudaMallocAsync((void **)&o->d_mlwe, nm * 8, current_stream());
    if (e == cudaSuccess) e = cudaMallocAsync((void **)&o->d_incommit, ni * 8, current_stream());
    if (zero && e == cudaSuccess) e = cudaMemsetAsync(o->d_encode, 0, ne * 8, current_stream());
    if (zero && e == cudaSuccess) e = cudaMemsetAsync(o->d_mlwe, 0, nm * 8, current_stream());
    if (zero && e == cudaSuccess) e = cudaMemsetAsync(o->d_incommit, 0, ni * 8, current_stream());
    if (e != cudaSuccess) { rs_opening_free(o); return set_error(RS_ERR_CUDA, "opening alloc: %s", cudaGetErrorString(e)); }
    *out = o;
    return RS_OK;
}

extern "C" {
#pragma GCC visibility push(default)

// ---------------------------------------------------------------------------------------------
// Evaluate's host-side vectors on the device (jindo/utils.go).
//   leftVec (utils.go:63-73): left[i] = (x^(cols*slots))^i, except left[rows-1] = x.  One thread per entry,
//   square-and-multiply on Montgomery elements (exact field arithmetic: the same canonical values as the
//   reference's running product).  The entries are then encoded by encode_kernel as Encoder.EncodeTo does
//   (encoder.go:113-117: one element in slot 0, no noise).
//   encodeChallengeTo (utils.go:21-46): 16 bytes -> two big-endian words, read as a little-endian 128-bit number
//   (bytes 8..15 are the HIGH limb), exp balanced digits in base B = ChallengeBound() at coefficients i*slots,
//   MForm, NTT.  One CTA per (challenge, RNS limb).
// ---------------------------------------------------------------------------------------------
extern "C++" {
namespace rs {
template <int NL>
__global__ void left_vec_kernel(uint32_t *__restrict__ out, const uint32_t *__restrict__ xin, int rows, unsigned skip_exp,
                                const __grid_constant__ FieldParams<NL> F) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows) return;
    uint32_t b[NL], skip[NL], r[NL];
#pragma unroll
    for (int l = 0; l < NL; l++) { b[l] = xin[l]; skip[l] = F.one[l]; r[l] = F.one[l]; }
    if (i == rows - 1) {                                    // utils.go:71
#pragma unroll
        for (int l = 0; l < NL; l++) out[(size_t)i * NL + l] = b[l];
        return;
    }
    for (unsigned e = skip_exp; e; e >>= 1) {               // skip = x^(cols*slots)
        if (e & 1) mont_mul<NL, 0>(skip, skip, b, F);
        mont_mul<NL, 0>(b, b, b, F);
    }
    for (unsigned e = (unsigned)i; e; e >>= 1) {            // left[i] = skip^i
        if (e & 1) mont_mul<NL, 0>(r, r, skip, F);
        mont_mul<NL, 0>(skip, skip, skip, F);
    }
#pragma unroll
    for (int l = 0; l < NL; l++) out[(size_t)i * NL + l] = r[l];
}

__global__ void __launch_bounds__(128) challenge_kernel(u64 *__restrict__ out, const uint8_t *__restrict__ bytes, int exp, int slots,
                                                        uint32_t bound, const __grid_constant__ RingDev R) {
    __shared__ u64 a[JD];
    __shared__ uint32_t dig[RS_JINDO_MAX_EXP];
    const int ch = blockIdx.x, l = blockIdx.y, t = threadIdx.x;
    const u64 q = R.q[l], qi = R.qinvneg[l];
    if (t == 0) {
        const uint8_t *b = bytes + (size_t)ch * 16;
        uint32_t w[4];                                      // little-endian 32-bit limbs of c = [BE64(b[0:8]), BE64(b[8:16])]
        w[1] = (uint32_t)b[0] << 24 | (uint32_t)b[1] << 16 | (uint32_t)b[2] << 8 | b[3];
        w[0] = (uint32_t)b[4] << 24 | (uint32_t)b[5] << 16 | (uint32_t)b[6] << 8 | b[7];
        w[3] = (uint32_t)b[8] << 24 | (uint32_t)b[9] << 16 | (uint32_t)b[10] << 8 | b[11];
        w[2] = (uint32_t)b[12] << 24 | (uint32_t)b[13] << 16 | (uint32_t)b[14] << 8 | b[15];
        for (int i = 0; i < exp; i++) {                     // divMod64 (utils.go:12-18)
            uint32_t r = 0;
            for (int k = 3; k >= 0; k--) {
                const u64 cur = ((u64)r << 32) | w[k];
                w[k] = (uint32_t)(cur / bound);
                r = (uint32_t)(cur % bound);
            }
            dig[i] = r;
        }
    }
    a[t] = 0; a[t + 128] = 0;
    __syncthreads();
    if (t < exp) {
        const uint32_t r = dig[t];
        const u64 val = (r > bound / 2) ? q - (u64)(bound - r) : (u64)r;
        a[t * slots] = mont64(val, R.r2[l], q, qi);         // MForm
    }
    __syncthreads();
    ntt256(a, R.roots + (size_t)l * JD, q, qi, t, 1, 128);
    u64 *dst = out + ((size_t)ch * R.n + l) * JD;
    dst[t] = a[t]; dst[t + 128] = a[t + 128];
}
}  // namespace rs
}  // extern "C++"

// ---------------------------------------------------------------------------------------------
// NewCommitKey on the device (jindo/entities.go:21-61).  The key is the SampleN stream of one
// csprng.UniformSampler seeded with the CRS: polynomial -> coefficient k -> limb l, each value the first draw below
// bound_l = MaxUint64 - MaxUint64 % q_l, reduced mod q_l (uniform.go:87-95).  Draws come from aes_ctr.cu (same XOR-
// accumulated keystream).  A rejected draw shifts every later sample by one stream position, and whether a draw is
// rejected depends on the limb it lands on, so: (1) flag the draws that ANY limb would reject (rare: < q_max / 2^64),
// (2) the host walks that short sorted list once, deciding each with the shift accumulated so far, (3) one kernel
// writes every polynomial, finding its shift by binary search in the list of rejected positions.
// ---------------------------------------------------------------------------------------------
struct KeyGeom {
    u64 n_in, n_ml, n_out;          // samples of ck.In, ck.MLWE (ringQ) and ck.Out (ringQOut)
    int nq, nqo;
    u64 q[JMAXL], qo[JMAXL];
};

namespace rs {
__global__ void key_candidates_kernel(const u64 *__restrict__ smp, u64 n, u64 min_bound, u64 *__restrict__ list, unsigned *count, unsigned cap) {
    for (u64 s = blockIdx.x * (u64)blockDim.x + threadIdx.x; s < n; s += (u64)gridDim.x * blockDim.x)
        if (smp[s] >= min_bound) { const unsigned at = atomicAdd(count, 1u); if (at < cap) { list[2 * at] = s; list[2 * at + 1] = smp[s]; } }
}

__global__ void key_sample_kernel(u64 *__restrict__ ck_in, u64 *__restrict__ ck_mlwe, u64 *__restrict__ ck_out, const u64 *__restrict__ smp,
                                  const u64 *__restrict__ rejected, unsigned n_rej, const __grid_constant__ KeyGeom G) {
    const u64 total = G.n_in + G.n_ml + G.n_out;
    for (u64 s = blockIdx.x * (u64)blockDim.x + threadIdx.x; s < total + n_rej; s += (u64)gridDim.x * blockDim.x) {
        unsigned lo = 0, hi = n_rej;                       // lo = #{ rejected positions <= s }
        while (lo < hi) { const unsigned mid = (lo + hi) >> 1; if (rejected[mid] <= s) lo = mid + 1; else hi = mid; }
        if (lo && rejected[lo - 1] == s) continue;         // this draw was thrown away
        u64 o = s - lo;
        u64 *dst; int nl; const u64 *mod;
        if (o < G.n_in) { dst = ck_in; nl = G.nq; mod = G.q; }
        else if (o < G.n_in + G.n_ml) { o -= G.n_in; dst = ck_mlwe; nl = G.nq; mod = G.q; }
        else { o -= G.n_in + G.n_ml; dst = ck_out; nl = G.nqo; mod = G.qo; }
        const u64 poly = o / ((u64)JD * nl), rem = o % ((u64)JD * nl);
        const int k = (int)(rem / nl), l = (int)(rem % nl);
        dst[(poly * nl + l) * JD + k] = smp[s] % mod[l];
    }
}
}  // namespace rs

static u64 sample_bound(u64 n) { return ~0ull - (~0ull % n); }

// fills j->d_ck_* from the CRS stream; returns the number of rejected draws through *n_rejected
static int generate_commit_key(rs_jindo *j, const uint8_t *key32, const uint8_t *iv16, size_t *n_rejected) {
    using namespace rs;
    const rs_jindo_params &p = j->p;
    KeyGeom G;
    memset(&G, 0, sizeof(G));
    G.nq = p.n_q; G.nqo = p.n_qout;
    G.n_in = (u64)p.in_msis * p.rows * pq(j); G.n_ml = (u64)p.in_msis * p.mlwe * pq(j);
    G.n_out = (u64)p.out_msis * (p.cols + 1) * p.in_msis * pqo(j);
    u64 min_bound = ~0ull;
    for (int l = 0; l < p.n_q; l++) { G.q[l] = p.q[l]; if (sample_bound(p.q[l]) < min_bound) min_bound = sample_bound(p.q[l]); }
    for (int l = 0; l < p.n_qout; l++) { G.qo[l] = p.qout[l]; if (sample_bound(p.qout[l]) < min_bound) min_bound = sample_bound(p.qout[l]); }
    const u64 total = G.n_in + G.n_ml + G.n_out;
    const u64 slack = total / 64 + 4096;                    // rejection rate < q_max / 2^64 <= 2^-4 at the 60-bit limit; typically 2^-9 .. 2^-26
    UniformStream *us = nullptr;
    int rc = uniform_stream_init(&us, key32, iv16);
    if (rc) return rc;
    const u64 *smp = nullptr;
    u64 *d_list = nullptr; unsigned *d_count = nullptr;
    std::vector<u64> cand, vals, rej;
    unsigned n_cand = 0;
    const unsigned cap = (unsigned)slack;
    rc = uniform_stream_draw(us, total + slack, &smp);
    cudaError_t e = cudaSuccess;
    if (rc == RS_OK) {
        e = cudaMalloc((void **)&d_list, (size_t)cap * 16);
        if (e == cudaSuccess) e = cudaMalloc((void **)&d_count, 4);
        if (e == cudaSuccess) e = cudaMemsetAsync(d_count, 0, 4, current_stream());
    }
    if (rc == RS_OK && e == cudaSuccess) {
        key_candidates_kernel<<<148 * 8, 256, 0, current_stream()>>>(smp, total + slack, min_bound, d_list, d_count, cap);
        count_launch();
        e = cudaMemcpyAsync(&n_cand, d_count, 4, cudaMemcpyDeviceToHost, current_stream());
        if (e == cudaSuccess) e = cudaStreamSynchronize(current_stream());
    }
    if (rc == RS_OK && e == cudaSuccess && n_cand > cap) rc = set_error(RS_ERR_UNSUPPORTED, "rs_jindo_create_crs: %u rejection candidates exceed the slack of %llu draws", n_cand, (unsigned long long)slack);
    if (rc == RS_OK && e == cudaSuccess && n_cand) {
        std::vector<u64> pairs((size_t)2 * n_cand);
        e = cudaMemcpy(pairs.data(), d_list, pairs.size() * 8, cudaMemcpyDeviceToHost);
        std::vector<std::pair<u64, u64>> pv(n_cand);
        for (unsigned i = 0; i < n_cand; i++) pv[i] = {pairs[2 * i], pairs[2 * i + 1]};
        std::sort(pv.begin(), pv.end());
        cand.resize(n_cand); vals.resize(n_cand);
        for (unsigned i = 0; i < n_cand; i++) { cand[i] = pv[i].first; vals[i] = pv[i].second; }
        if (e == cudaSuccess) {
            for (unsigned i = 0; i < n_cand; i++) {
                const u64 s = cand[i];
                if (s >= total + rej.size()) break;          // past the last draw the key consumes
                u64 o = s - rej.size();
                int nl; const u64 *mod;
                if (o < G.n_in) { nl = G.nq; mod = G.q; }
                else if (o < G.n_in + G.n_ml) { o -= G.n_in; nl = G.nq; mod = G.q; }
                else { o -= G.n_in + G.n_ml; nl = G.nqo; mod = G.qo; }
                const int l = (int)(o % (u64)nl);
                if (vals[i] >= sample_bound(mod[l])) rej.push_back(s);
            }
        }
    }
    if (rc == RS_OK && e == cudaSuccess && rej.size() > slack) rc = set_error(RS_ERR_UNSUPPORTED, "rs_jindo_create_crs: more rejected draws than slack");
    if (rc == RS_OK && e == cudaSuccess) {
        if (!rej.empty()) e = cudaMemcpyAsync(d_list, rej.data(), rej.size() * 8, cudaMemcpyHostToDevice, current_stream());
        if (e == cudaSuccess) {
            key_sample_kernel<<<148 * 8, 256, 0, current_stream()>>>(j->d_ck_in, j->d_ck_mlwe, j->d_ck_out, smp, d_list, (unsigned)rej.size(), G);
            count_launch();
            uniform_stream_release(us);
            e = cudaStreamSynchronize(current_stream());
            if (e == cudaSuccess) e = cudaGetLastError();
        }
    }
    cudaFree(d_list); cudaFree(d_count);
    uniform_stream_free(us);
    if (rc) return rc;
    if (e != cudaSuccess) return set_error(RS_ERR_CUDA, "rs_jindo_create_crs: %s", cudaGetErrorString(e));
    if (n_rejected) *n_rejected = rej.size();
    return RS_OK;
}

static int jindo_create_impl(rs_field *f, const rs_jindo_params *p, const uint64_t *ck_in, const uint64_t *ck_mlwe,
                             const uint64_t *ck_out, const uint64_t *cdt_tables, const uint8_t *key32, const uint8_t *iv16, rs_jindo **out);

int rs_jindo_create(rs_field *f, const rs_jindo_params *p, const uint64_t *ck_in, const uint64_t *ck_mlwe,
                    const uint64_t *ck_out, const uint64_t *cdt_tables, rs_jindo **out) {
    RS_REQUIRE(f && p && ck_in && ck_mlwe && ck_out && cdt_tables && out, RS_ERR_ARG, "rs_jindo_create: null argument");
    return jindo_create_impl(f, p, ck_in, ck_mlwe, ck_out, cdt_tables, nullptr, nullptr, out);
}

int rs_jindo_create_crs(rs_field *f, const rs_jindo_params *p, const uint8_t *key32, const uint8_t *iv16,
                        const uint64_t *cdt_tables, rs_jindo **out) {
    RS_REQUIRE(f && p && key32 && iv16 && cdt_tables && out, RS_ERR_ARG, "rs_jindo_create_crs: null argument");
    return jindo_create_impl(f, p, nullptr, nullptr, nullptr, cdt_tables, key32, iv16, out);
}

int rs_jindo_get_params(const rs_jindo *j, rs_jindo_params *out) {
    RS_REQUIRE(j && out, RS_ERR_ARG, "rs_jindo_get_params: null argument");
    *out = j->p;
    return RS_OK;
}

int rs_jindo_commit_key_download(const rs_jindo *j, int which, uint64_t *out) {
    RS_REQUIRE(j && out && which >= 0 && which <= 2, RS_ERR_ARG, "rs_jindo_commit_key_download: bad argument");
    const rs_jindo_params &p = j->p;
    const size_t n = which == 0 ? (size_t)p.in_msis * p.rows * pq(j) : which == 1 ? (size_t)p.in_msis * p.mlwe * pq(j)
                                : (size_t)p.out_msis * (p.cols + 1) * p.in_msis * pqo(j);
    const u64 *src = which == 0 ? j->d_ck_in : which == 1 ? j->d_ck_mlwe : j->d_ck_out;
    RS_CUDA_CHECK(cudaMemcpy(out, src, n * 8, cudaMemcpyDeviceToHost));
    return RS_OK;
}

static int jindo_create_impl(rs_field *f, const rs_jindo_params *p, const uint64_t *ck_in, const uint64_t *ck_mlwe,
                             const uint64_t *ck_out, const uint64_t *cdt_tables, const uint8_t *key32, const uint8_t *iv16, rs_jindo **out) {
    RS_REQUIRE(p->d == JD, RS_ERR_UNSUPPORTED, "rs_jindo_create: ring degree must be 256");
    RS_REQUIRE(p->n_q >= 1 && p->n_q <= JMAXL && p->n_qout >= 1 && p->n_qout <= JMAXL, RS_ERR_UNSUPPORTED, "rs_jindo_create: 1..4 RNS limbs");
    RS_REQUIRE(p->exp >= 1 && p->exp <= RS_JINDO_MAX_EXP && p->slots * p->exp == JD, RS_ERR_ARG, "rs_jindo_create: slots*exp != d");
    RS_REQUIRE(p->base >= 2 && p->base < (1ull << 31) && (p->base & (p->base - 1)) != 0, RS_ERR_UNSUPPORTED, "rs_jindo_create: base must be < 2^31 and not a power of two");
    RS_REQUIRE(p->rows >= 2 && p->cols >= 1 && p->in_msis >= 1 && p->out_msis >= 1 && p->mlwe >= 1, RS_ERR_ARG, "rs_jindo_create: bad shape");
    RS_REQUIRE(p->cdt_len >= 1, RS_ERR_ARG, "rs_jindo_create: bad CDT");
    rs_jindo *j = new rs_jindo();
    memset(j, 0, sizeof(*j));
    j->no_pair = getenv("RS_JINDO_NOPAIR") ? 1 : 0; j->prefetch = getenv("RS_JINDO_PREFETCH") ? 1 : 0;
    j->cache = new std::vector<DescCache>();
    j->f = f;
    j->p = *p;
    std::vector<u64> roots((size_t)5 * (p->n_q + p->n_qout) * JD);
    u64 *rq_h = roots.data(), *rqo_h = roots.data() + (size_t)5 * p->n_q * JD;
    int rc = ring_setup(j->rq, j->crt_q, p->n_q, (const u64 *)p->q, (const u64 *)p->psi_q, rq_h);
    if (rc == RS_OK) rc = ring_setup(j->rqo, j->crt_qo, p->n_qout, (const u64 *)p->qout, (const u64 *)p->psi_qout, rqo_h);
    if (rc != RS_OK) { delete j; return rc; }
    for (int l = 0; l < p->n_q; l++) j->base_mont_q[l] = h_mulmod(p->base % p->q[l], (u64)((((u128)1) << 64) % p->q[l]), p->q[l]);
    const size_t n_in = (size_t)p->in_msis * p->rows * pq(j), n_ml = (size_t)p->in_msis * p->mlwe * pq(j),
                 n_out = (size_t)p->out_msis * (p->cols + 1) * p->in_msis * pqo(j), n_cdt = (size_t)128 * p->cdt_len;
    cudaError_t e = cudaMalloc((void **)&j->d_roots, roots.size() * 8);
    if (e == cudaSuccess) e = cudaMalloc((void **)&j->d_ck_in, n_in * 8);
    if (e == cudaSuccess) e = cudaMalloc((void **)&j->d_ck_mlwe, n_ml * 8);
    if (e == cudaSuccess) e = cudaMalloc((void **)&j->d_ck_out, n_out * 8);
    if (e == cudaSuccess) e = cudaMalloc((void **)&j->d_cdt, n_cdt * 8);
    if (e == cudaSuccess) e = cudaMemcpy(j->d_roots, roots.data(), roots.size() * 8, cudaMemcpyHostToDevice);
    if (ck_in) {
        if (e == cudaSuccess) e = cudaMemcpy(j->d_ck_in, ck_in, n_in * 8, cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(j->d_ck_mlwe, ck_mlwe, n_ml * 8, cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(j->d_ck_out, ck_out, n_out * 8, cudaMemcpyHostToDevice);
    } else if (e == cudaSuccess) {
        rc = generate_commit_key(j, key32, iv16, &j->ck_rejected);
        if (rc != RS_OK) { rs_jindo_destroy(j); return rc; }
    }
    if (e == cudaSuccess) e = cudaMemcpy(j->d_cdt, cdt_tables, n_cdt * 8, cudaMemcpyHostToDevice);
    if (e == cudaSuccess && p->cdt_len <= 255 && !getenv("RS_JINDO_NOGUIDE")) {
        // guide[t][b] = #{ i : table_t[i] < b << 56 }, b = 0..256 (tables are sorted: gaussian_twin_cdt.go:13-35)
        std::vector<uint8_t> guide((size_t)128 * 257);
        for (int t = 0; t < 128; t++) {
            const uint64_t *tab = cdt_tables + (size_t)t * p->cdt_len;
            int i = 0;
            for (int b = 0; b <= 256; b++) {
                while (i < p->cdt_len && (b == 256 || tab[i] < ((uint64_t)b << 56))) i++;
                guide[(size_t)t * 257 + b] = (uint8_t)i;
            }
        }
        e = cudaMalloc((void **)&j->d_cdt_guide, guide.size());
        if (e == cudaSuccess) e = cudaMemcpy(j->d_cdt_guide, guide.data(), guide.size(), cudaMemcpyHostToDevice);
    }
    if (e == cudaSuccess && !getenv("RS_JINDO_NOSLOWLB")) {
        // lower bounds for the TwinCDT slow path (gaussian_twin_cdt.go:95-110): for cfrac in [a, b) = [k/16, (k+1)/16),
        // rho(x - cfrac) >= rho(x - b) for x <= 0 and >= rho(x - a) for x >= 1, so
        //   sum_{x = tail_lo .. v0} rho(x - cfrac) >= sum_{x = tail_lo .. min(v0, 0)} rho(x - b) + sum_{x = 1 .. v0} rho(x - a)
        const double sg = p->cdt_sigma, den = 2.0 * sg * sg, norm = sqrt(6.283185307179586) * sg;
        std::vector<double> lb((size_t)16 * p->cdt_len);
        for (int k = 0; k < 16; k++) {
            const double a = k / 16.0, b = (k + 1) / 16.0;
            double neg = 0.0;                                   // x <= 0 part, accumulated up to min(v0, 0); v0 >= 0 so it is complete
            for (long long x = p->cdt_tail_lo; x <= 0; x++) neg += exp(-((double)x - b) * ((double)x - b) / den);
            double pos = 0.0;
            for (int v0 = 0; v0 < p->cdt_len; v0++) {
                if (v0 >= 1) pos += exp(-((double)v0 - a) * ((double)v0 - a) / den);
                lb[(size_t)k * p->cdt_len + v0] = (neg + pos) / norm - 1e-9;
            }
        }
        e = cudaMalloc((void **)&j->d_slow_lb, lb.size() * 8);
        if (e == cudaSuccess) e = cudaMemcpy(j->d_slow_lb, lb.data(), lb.size() * 8, cudaMemcpyHostToDevice);
    }
    if (e != cudaSuccess) { rs_jindo_destroy(j); return set_error(RS_ERR_CUDA, "rs_jindo_create: %s", cudaGetErrorString(e)); }
    j->rq.roots = j->d_roots;
    j->rq.roots_inv = j->d_roots + (size_t)p->n_q * JD;
    j->rq.roots_plain = j->d_roots + (size_t)2 * p->n_q * JD;
    j->rq.roots_shoup = j->d_roots + (size_t)3 * p->n_q * JD;
    j->rq.roots_f64 = reinterpret_cast<const double *>(j->d_roots + (size_t)4 * p->n_q * JD);
    j->rqo.roots = j->d_roots + (size_t)5 * p->n_q * JD;
    j->rqo.roots_inv = j->rqo.roots + (size_t)p->n_qout * JD;
    j->rqo.roots_plain = j->rqo.roots + (size_t)2 * p->n_qout * JD;
    j->rqo.roots_shoup = j->rqo.roots + (size_t)3 * p->n_qout * JD;
    j->rqo.roots_f64 = reinterpret_cast<const double *>(j->rqo.roots + (size_t)4 * p->n_qout * JD);
    cudaStreamCreateWithFlags(&j->copy_stream, cudaStreamNonBlocking);
    cudaEventCreateWithFlags(&j->ev_alloc, cudaEventDisableTiming);
    for (int k = 0; k < RS_VCHUNKS; k++) cudaEventCreateWithFlags(&j->ev_piece[k], cudaEventDisableTiming);
    *out = j;
    return RS_OK;
}

void rs_jindo_destroy(rs_jindo *j) {
    if (!j) return;
    if (j->copy_stream) {
        cudaStreamSynchronize(j->copy_stream);
        cudaStreamDestroy(j->copy_stream); cudaEventDestroy(j->ev_alloc);
        for (int k = 0; k < RS_VCHUNKS; k++) cudaEventDestroy(j->ev_piece[k]);
    }
    cudaFree(j->d_roots); cudaFree(j->d_ck_in); cudaFree(j->d_ck_mlwe); cudaFree(j->d_ck_out); cudaFree(j->d_cdt); cudaFree(j->d_cdt_guide); cudaFree(j->d_slow_lb);
    if (j->cache) { for (auto &c : *j->cache) cudaFree(c.d_descs); delete j->cache; }
    cudaFree(j->d_ndescs); cudaFree(j->d_ldescs); cudaFree(j->d_zero_g);
    rs::uniform_stream_free(j->twin_stream);
    rs::cosac_stream_free(j->cosac_stream);
    rs::uniform_stream_free(j->mlwe_stream);
    cudaFree(j->d_mlwe_cdt);
    delete j;
}

int rs_jindo_set_twin_stream(rs_jindo *j, const uint8_t *key32, const uint8_t *iv16) {
    RS_REQUIRE(j && key32 && iv16, RS_ERR_ARG, "rs_jindo_set_twin_stream: null argument");
    rs::uniform_stream_free(j->twin_stream);
    j->twin_stream = nullptr;
    return rs::uniform_stream_init(&j->twin_stream, key32, iv16);
}

int rs_jindo_twin_stream_draw(rs_jindo *j, size_t n, uint64_t *out) {
    RS_REQUIRE(j && j->twin_stream && (out || n == 0), RS_ERR_ARG, "rs_jindo_twin_stream_draw: no stream / null argument");
    const u64 *smp = nullptr;
    int rc = rs::uniform_stream_draw(j->twin_stream, n, &smp);
    if (rc || n == 0) return rc;
    cudaError_t e = cudaMemcpyAsync(out, smp, n * 8, cudaMemcpyDeviceToHost, rs::current_stream());
    rs::uniform_stream_release(j->twin_stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(rs::current_stream());
    if (e != cudaSuccess) return rs::set_error(RS_ERR_CUDA, "rs_jindo_twin_stream_draw: %s", cudaGetErrorString(e));
    return RS_OK;
}

namespace rs {
// TwinCDTGaussianSampler.Sample(0) (gaussian_twin_cdt.go:77-97): v0 = slices.BinarySearch(table0, u), minus one on an exact
// hit; c0 == c1 == 0, so v0 == v1 and the result is v0 + tailLo.  One thread per sample, table in shared memory.
__global__ void __launch_bounds__(256) cdt0_sample_kernel(i64 *__restrict__ out, const u64 *__restrict__ draws, const u64 *__restrict__ table,
                                                         int len, i64 tail_lo, size_t n) {
    extern __shared__ u64 s_tab[];
    for (int i = threadIdx.x; i < len; i += blockDim.x) s_tab[i] = table[i];
    __syncthreads();
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const u64 u = draws[i];
        int lo = 0, hi = len;                       // first index with table[idx] >= u
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (s_tab[mid] < u) lo = mid + 1; else hi = mid; }
        const int v0 = (lo < len && s_tab[lo] == u) ? lo - 1 : lo;
        out[i] = (i64)v0 + tail_lo;
    }
}
static int mlwe_stream_sample(rs_jindo *j, size_t n, i64 *d_out) {
    if (n == 0) return RS_OK;
    const u64 *smp = nullptr;
    int rc = uniform_stream_draw(j->mlwe_stream, n, &smp);
    if (rc) return rc;
    size_t blocks = (n + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    cdt0_sample_kernel<<<(unsigned)blocks, 256, (size_t)j->mlwe_cdt_len * 8, current_stream()>>>(d_out, smp, j->d_mlwe_cdt, j->mlwe_cdt_len,
                                                                                             (i64)j->mlwe_tail_lo, n);
    count_launch();
    uniform_stream_release(j->mlwe_stream);
    if (cudaGetLastError() != cudaSuccess) return set_error(RS_ERR_CUDA, "cdt0_sample_kernel launch failed");
    return RS_OK;
}
}  // namespace rs

int rs_jindo_set_mlwe_stream(rs_jindo *j, const uint8_t *key32, const uint8_t *iv16, const uint64_t *table0, int len, int64_t tail_lo) {
    RS_REQUIRE(j && key32 && iv16 && table0 && len > 0 && len <= 4096, RS_ERR_ARG, "rs_jindo_set_mlwe_stream: bad argument");
    rs::uniform_stream_free(j->mlwe_stream);
    j->mlwe_stream = nullptr;
    cudaFree(j->d_mlwe_cdt); j->d_mlwe_cdt = nullptr;
    RS_CUDA_CHECK(cudaMalloc((void **)&j->d_mlwe_cdt, (size_t)len * 8));
    RS_CUDA_CHECK(cudaMemcpy(j->d_mlwe_cdt, table0, (size_t)len * 8, cudaMemcpyHostToDevice));
    j->mlwe_cdt_len = len; j->mlwe_tail_lo = tail_lo;
    return rs::uniform_stream_init(&j->mlwe_stream, key32, iv16);
}

int rs_jindo_set_cosac_stream(rs_jindo *j, const uint8_t *key48_base, const uint8_t *key48_rounded, double ecd_blind_std, double mask_blind_std,
                              double mask_std) {
    RS_REQUIRE(j && key48_base && key48_rounded, RS_ERR_ARG, "rs_jindo_set_cosac_stream: null argument");
    RS_REQUIRE(ecd_blind_std > 0 && mask_blind_std > 0 && mask_std > 0, RS_ERR_ARG, "rs_jindo_set_cosac_stream: standard deviation not positive");
    if (j->cosac_stream) { rs::cosac_stream_free(j->cosac_stream); j->cosac_stream = nullptr; }
    return rs::cosac_stream_init(&j->cosac_stream, key48_base, key48_rounded, ecd_blind_std, mask_blind_std, mask_std);
}

int rs_jindo_cosac_status(rs_jindo *j, uint64_t *workers, int *overflowed) {
    RS_REQUIRE(j && j->cosac_stream, RS_ERR_STATE, "rs_jindo_cosac_status: no stream was set (rs_jindo_set_cosac_stream)");
    if (workers) *workers = rs::cosac_stream_workers(j->cosac_stream);
    if (overflowed) {
        int flag = 0;
        if (cudaMemcpyAsync(&flag, rs::cosac_stream_overflow_flag(j->cosac_stream), 4, cudaMemcpyDeviceToHost, rs::current_stream()) != cudaSuccess ||
            cudaStreamSynchronize(rs::current_stream()) != cudaSuccess)
            return rs::set_error(RS_ERR_CUDA, "rs_jindo_cosac_status: %s", cudaGetErrorString(cudaGetLastError()));
        *overflowed = flag;
    }
    return RS_OK;
}

int rs_jindo_cosac_sample(rs_jindo *j, const uint64_t *first_row, const uint64_t *mask_rows, size_t n_mask_rows, int64_t *out, uint64_t *worker0,
                          int *overflowed) {
    RS_REQUIRE(j && first_row && mask_rows && out, RS_ERR_ARG, "rs_jindo_cosac_sample: null argument");
    RS_REQUIRE(j->cosac_stream, RS_ERR_STATE, "rs_jindo_cosac_sample: no stream was set (rs_jindo_set_cosac_stream)");
    const rs_jindo_params &p = j->p;
    const size_t NLw = (size_t)j->f->NL, cs = (size_t)p.cols * p.slots, n_polys = (size_t)p.cols + n_mask_rows;
    if (worker0) *worker0 = rs::cosac_stream_workers(j->cosac_stream);
    uint32_t *d_el = nullptr;
    i64 *d_g = nullptr;
    RS_CUDA_CHECK(cudaMallocAsync((void **)&d_el, (cs + n_mask_rows * p.slots) * NLw * 4 + 16, rs::current_stream()));
    int rc = RS_OK;
    if (cudaMallocAsync((void **)&d_g, n_polys * JD * 8, rs::current_stream()) != cudaSuccess) rc = rs::set_error(RS_ERR_CUDA, "rs_jindo_cosac_sample: cudaMallocAsync failed");
    if (rc == RS_OK && (cudaMemcpyAsync(d_el, first_row, cs * NLw * 4, cudaMemcpyHostToDevice, rs::current_stream()) != cudaSuccess ||
                        cudaMemcpyAsync(d_el + cs * NLw, mask_rows, n_mask_rows * p.slots * NLw * 4, cudaMemcpyHostToDevice, rs::current_stream()) != cudaSuccess))
        rc = rs::set_error(RS_ERR_CUDA, "rs_jindo_cosac_sample: upload failed");
    if (rc == RS_OK) rc = rs::cosac_generate(j->cosac_stream, j->f, p, d_el, d_el + cs * NLw, n_polys, d_g);
    int flag = 0;
    if (rc == RS_OK && (cudaMemcpyAsync(out, d_g, n_polys * JD * 8, cudaMemcpyDeviceToHost, rs::current_stream()) != cudaSuccess ||
                        cudaMemcpyAsync(&flag, rs::cosac_stream_overflow_flag(j->cosac_stream), 4, cudaMemcpyDeviceToHost, rs::current_stream()) != cudaSuccess ||
                        cudaStreamSynchronize(rs::current_stream()) != cudaSuccess))
        rc = rs::set_error(RS_ERR_CUDA, "rs_jindo_cosac_sample: %s", cudaGetErrorString(cudaGetLastError()));
    if (overflowed) *overflowed = flag;
    cudaFreeAsync(d_el, rs::current_stream());
    if (d_g) cudaFreeAsync(d_g, rs::current_stream());
    return rc;
}

int rs_jindo_mlwe_stream_sample(rs_jindo *j, size_t n, int64_t *out) {
    RS_REQUIRE(j && j->mlwe_stream && (out || n == 0), RS_ERR_ARG, "rs_jindo_mlwe_stream_sample: no stream / null argument");
    if (n == 0) return RS_OK;
    i64 *d = nullptr;
    RS_CUDA_CHECK(cudaMallocAsync((void **)&d, n * 8, rs::current_stream()));
    int rc = rs::mlwe_stream_sample(j, n, d);
    cudaError_t e = cudaSuccess;
    if (rc == RS_OK) e = cudaMemcpyAsync(out, d, n * 8, cudaMemcpyDeviceToHost, rs::current_stream());
    cudaFreeAsync(d, rs::current_stream());
    if (rc) return rc;
    if (e == cudaSuccess) e = cudaStreamSynchronize(rs::current_stream());
    if (e != cudaSuccess) return rs::set_error(RS_ERR_CUDA, "rs_jindo_mlwe_stream_sample: %s", cudaGetErrorString(e));
    return RS_OK;
}

void rs_opening_free(rs_opening *o) {
    if (!o) return;
    if (o->d_encode) cudaFreeAsync(o->d_encode, rs::current_stream());
    if (o->d_mlwe) cudaFreeAsync(o->d_mlwe, rs::current_stream());
    if (o->d_incommit) cudaFreeAsync(o->d_incommit, rs::current_stream());
    delete o;
}

// =============================================================================================
// Verifier.Verify on the device (jindo/verifier.go:50-282)
// =============================================================================================
static int encode_left_vec(rs_jindo *j, const uint64_t *x_mont, u64 **d_left);
static int encode_challenges(rs_jindo *j, const uint8_t *bytes, size_t n, const RingDev &ring, u64 **d_out);
namespace {
// sum of n_parts numbers of 2 JW words -> total (2 JW + 2 words)
void sq_total(const u64 *parts, size_t n_parts, u64 *tot /* 2*JW+2 */) {
    for (int i = 0; i < 2 * JW + 2; i++) tot[i] = 0;
    for (size_t k = 0; k < n_parts; k++) {
        u64 cy = 0;
        for (int i = 0; i < 2 * JW + 2; i++) {
            const u64 b = i < 2 * JW ? parts[k * 2 * JW + i] : 0;
            const u64 t = tot[i] + b; u64 c1 = t < b;
            const u64 t2 = t + cy; c1 += t2 < t;
            tot[i] = t2; cy = c1;
        }
    }
}
// float64(isqrt(S)) < nm, as verifyNorm does (big.Int.Sqrt then Float64, verifier.go:278-281).  S: W little-endian words.
bool isqrt_lt(const u64 *S, int W, double nm) {
    const int HW = (W + 1) / 2;
    std::vector<u64> r(HW, 0), cand(HW), sq(2 * HW + 1);
    for (int bit = 64 * HW - 1; bit >= 0; bit--) {
        cand = r;
        cand[bit / 64] |= 1ull << (bit % 64);
        std::fill(sq.begin(), sq.end(), 0);
        for (int a = 0; a < HW; a++) {
            u64 carry = 0;
            for (int b = 0; b < HW; b++) {
                const u128 t = (u128)cand[a] * cand[b] + sq[a + b] + carry;
                sq[a + b] = (u64)t; carry = (u64)(t >> 64);
            }
            sq[a + HW] += carry;
        }
        bool le = true;                                              // cand^2 <= S ?
        for (int i = 2 * HW; i >= 0; i--) {
            const u64 sv = i < W ? S[i] : 0;
            if (sq[i] != sv) { le = sq[i] < sv; break; }
        }
        if (le) r = cand;
    }
    // big.Int.Float64: nearest, ties to even
    int top = HW - 1;
    while (top > 0 && r[top] == 0) top--;
    if (top == 0 && r[0] < (1ull << 53)) return (double)r[0] < nm;
    const int hb = 63 - __builtin_clzll(r[top]);                      // highest set bit of the top word
    const int total_bits = 64 * top + hb + 1;
    const int shift = total_bits - 53;                                // bits dropped
    auto bit_at = [&](int b) { return b < 0 ? 0ull : (r[b / 64] >> (b % 64)) & 1ull; };
    u64 mant = 0;
    for (int b = 0; b < 53; b++) mant |= bit_at(shift + b) << b;
    bool sticky = false;
    for (int b = 0; b < shift - 1 && !sticky; b++) sticky = bit_at(b);
    const u64 half = bit_at(shift - 1);
    if (half && (sticky || (mant & 1))) mant++;
    return ldexp((double)mant, shift) < nm;
}
}  // namespace

int rs_jindo_verify(rs_jindo *j, const uint64_t *x_mont, const uint64_t *commitments, int batch, const uint8_t *batch_bytes,
                    const uint8_t *chal_bytes, const uint64_t *ys_mont, const rs_jindo_proof *pf, double in_com_dcmp_two_nm,
                    double res_two_nm, int *checks) {
    RS_REQUIRE(j && x_mont && commitments && chal_bytes && ys_mont && pf && checks && batch >= 1, RS_ERR_ARG, "rs_jindo_verify: null argument");
    RS_REQUIRE(pf->partial && pf->partial_mask && pf->encode && pf->mlwe && pf->incommit, RS_ERR_ARG, "rs_jindo_verify: incomplete proof");
    RS_REQUIRE(batch == 1 || batch_bytes, RS_ERR_ARG, "rs_jindo_verify: batch challenges missing");
    const rs_jindo_params &p = j->p;
    const int cols = p.cols, rows = p.rows, nm = p.mlwe + p.in_msis, NL = j->f->NL;
    const size_t n_com = (size_t)(cols + 1) * p.in_msis, wq = pq(j), wo = pqo(j);
    *checks = 0;
    int rc = RS_OK;
    cudaStream_t st = current_stream();
    // ---- device copies of the proof and of the commitments
    u64 *d_part = nullptr, *d_enc = nullptr, *d_ml = nullptr, *d_inc = nullptr, *d_coms = nullptr, *d_chal = nullptr, *d_bq = nullptr, *d_bo = nullptr;
    u64 *d_left = nullptr, *d_tmp = nullptr, *d_sq = nullptr, *d_ints = nullptr;
    unsigned *d_flag = nullptr;
    std::vector<u64> h_sq, h_ints;
    MatvecGeom Gq, Go, Sq, So;
    memset(&Gq, 0, sizeof(Gq)); memset(&Go, 0, sizeof(Go)); memset(&Sq, 0, sizeof(Sq)); memset(&So, 0, sizeof(So));
    fill_geom_ring(Gq, j->rq); fill_geom_ring(Go, j->rqo);
    for (int l = 0; l < j->rq.n; l++) {       // MForm(2^log_in_cut mod q_l): the RNS scalar of verifier.go:28-37
        const u64 q = j->rq.q[l];
        Sq.q[l] = h_mulmod(h_powmod(2, (u64)p.log_in_cut, q), h_powmod(2, 64, q), q);
    }
    for (int l = 0; l < j->rqo.n; l++) {
        const u64 q = j->rqo.q[l];
        So.q[l] = h_mulmod(h_powmod(2, (u64)p.log_out_cut, q), h_powmod(2, 64, q), q);
    }
#define RS_V(x) do { if (rc == RS_OK) { rc = (x); } } while (0)
#define RS_VA(ptr, words) do { if (rc == RS_OK && cudaMallocAsync((void **)&(ptr), (size_t)(words) * 8, st) != cudaSuccess) rc = set_error(RS_ERR_CUDA, "rs_jindo_verify: cudaMallocAsync failed"); } while (0)
#define RS_VU(dst, src, words) do { if (rc == RS_OK && cudaMemcpyAsync(dst, src, (size_t)(words) * 8, cudaMemcpyHostToDevice, st) != cudaSuccess) rc = set_error(RS_ERR_CUDA, "rs_jindo_verify: upload failed"); } while (0)
    RS_VA(d_part, (size_t)(cols + 1) * wq); RS_VA(d_enc, (size_t)rows * wq); RS_VA(d_ml, (size_t)nm * wq); RS_VA(d_inc, n_com * wo);
    RS_VA(d_coms, (size_t)batch * p.out_msis * wo);
    RS_VU(d_part, pf->partial, (size_t)cols * wq); RS_VU(d_part + (size_t)cols * wq, pf->partial_mask, wq);
    RS_VU(d_enc, pf->encode, (size_t)rows * wq); RS_VU(d_ml, pf->mlwe, (size_t)nm * wq); RS_VU(d_inc, pf->incommit, n_com * wo);
    RS_VU(d_coms, commitments, (size_t)batch * p.out_msis * wo);
    // challenge polynomials (encodeChallengeTo, utils.go:21-46) and the encoded leftVec, on the device
    RS_V(encode_challenges(j, chal_bytes, (size_t)cols, j->rq, &d_chal));
    if (batch > 1) { RS_V(encode_challenges(j, batch_bytes, (size_t)batch, j->rq, &d_bq)); RS_V(encode_challenges(j, batch_bytes, (size_t)batch, j->rqo, &d_bo)); }
    RS_V(encode_left_vec(j, x_mont, &d_left));
    // scratch: cut-off polynomials, mod-up images, partial sums
    const size_t tmp_q = (size_t)(3 * p.in_msis + 2) * wq + n_com * wq, tmp_o = (size_t)(3 * p.out_msis) * wo + (size_t)(cols + 1) * p.out_msis * wo;
    RS_VA(d_tmp, tmp_q + tmp_o);
    const size_t n_sq_polys = n_com + p.out_msis + rows + nm + p.in_msis;
    RS_VA(d_sq, n_sq_polys * 2 * JW);
    RS_VA(d_ints, (size_t)(cols + batch) * JD * JW);
    if (rc == RS_OK && cudaMallocAsync((void **)&d_flag, 4, st) != cudaSuccess) rc = set_error(RS_ERR_CUDA, "rs_jindo_verify: cudaMallocAsync failed");
    if (rc == RS_OK) cudaMemsetAsync(d_flag, 0, 4, st);
    if (rc == RS_OK) {
        u64 *q_acc = d_tmp, *q_T = q_acc + (size_t)p.in_msis * wq, *q_cut = q_T + (size_t)p.in_msis * wq, *q_t1 = q_cut + (size_t)p.in_msis * wq,
            *q_t2 = q_t1 + wq, *q_up = q_t2 + wq;
        u64 *o_comb = d_tmp + tmp_q, *o_S = o_comb + (size_t)p.out_msis * wo, *o_cut = o_S + (size_t)p.out_msis * wo, *o_part = o_cut + (size_t)p.out_msis * wo;
        u64 *sq_o = d_sq, *sq_q = d_sq + (n_com + p.out_msis) * 2 * JW;
        MatvecGeom G;
        // ---- outer commitment (verifier.go:135-163)
        if (batch > 1) {         // cutoff_i = sum_b com[b][i] (.) batchOut[b]
            memset(&G, 0, sizeof(G));
            G.O = 1; G.C = p.out_msis; G.K = batch; G.key_os = 0; G.key_ks = 1; G.x_cs = 1; G.x_ks = p.out_msis; G.out_cs = 1; G.out_os = 0;
            RS_V(launch_matvec(o_comb, d_bo, d_coms, nullptr, G, j->rqo));
        } else if (cudaMemcpyAsync(o_comb, d_coms, (size_t)p.out_msis * wo * 8, cudaMemcpyDeviceToDevice, st) != cudaSuccess) {
            rc = set_error(RS_ERR_CUDA, "rs_jindo_verify: copy failed");
        }
        memset(&G, 0, sizeof(G));
        G.O = p.out_msis; G.C = cols + 1; G.K = p.in_msis;
        G.key_os = (i64)n_com; G.key_cs = p.in_msis; G.key_ks = 1; G.x_cs = p.in_msis; G.x_ks = 1; G.out_cs = p.out_msis; G.out_os = 1;
        RS_V(launch_matvec(o_part, j->d_ck_out, d_inc, nullptr, G, j->rqo));
        if (rc == RS_OK) {
            const size_t tot = (size_t)p.out_msis * wo;
            reduce_cols_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(o_S, o_part, cols + 1, p.out_msis, Go);
            scale_sub_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(o_cut, o_comb, o_S, nullptr, (size_t)p.out_msis, Go, So);
            count_launch(2);
        }
        RS_V(launch_recon(j, true, false, nullptr, d_inc, n_com, nullptr, sq_o));
        RS_V(launch_recon(j, true, false, nullptr, o_cut, (size_t)p.out_msis, nullptr, sq_o + n_com * 2 * JW));
        // ---- inner commitment (verifier.go:165-205)
        RS_V(launch_recon(j, true, true, q_up, d_inc, n_com, nullptr, nullptr));                 // ModUpQtoP + MForm + NTT
        memset(&G, 0, sizeof(G));
        G.O = 1; G.C = p.in_msis; G.K = cols; G.key_os = 0; G.key_ks = 1; G.x_cs = 1; G.x_ks = p.in_msis; G.out_cs = 1; G.out_os = 0;
        G.has_addend = 1; G.add_cs = 1; G.add_os = 0;
        RS_V(launch_matvec(q_acc, d_chal, q_up, q_up + (size_t)cols * p.in_msis * wq, G, j->rq));
        memset(&G, 0, sizeof(G));
        G.O = p.in_msis; G.C = 1; G.K = rows; G.key_os = rows; G.key_ks = 1; G.x_cs = rows; G.x_ks = 1; G.out_cs = p.in_msis; G.out_os = 1;
        RS_V(launch_matvec(q_T, j->d_ck_in, d_enc, nullptr, G, j->rq));
        G.K = p.mlwe; G.key_os = p.mlwe; G.x_cs = nm; G.accumulate = 1; G.has_addend = 1; G.add_cs = nm; G.add_os = 1;
        RS_V(launch_matvec(q_T, j->d_ck_mlwe, d_ml, d_ml + (size_t)p.mlwe * wq, G, j->rq));
        if (rc == RS_OK) {
            const size_t tot = (size_t)p.in_msis * wq;
            scale_sub_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(q_cut, q_acc, q_T, nullptr, (size_t)p.in_msis, Gq, Sq);
            count_launch();
        }
        RS_V(launch_recon(j, false, false, nullptr, d_enc, (size_t)rows, nullptr, sq_q));
        RS_V(launch_recon(j, false, false, nullptr, d_ml, (size_t)nm, nullptr, sq_q + (size_t)rows * 2 * JW));
        RS_V(launch_recon(j, false, false, nullptr, q_cut, (size_t)p.in_msis, nullptr, sq_q + (size_t)(rows + nm) * 2 * JW));
        // ---- consistency (verifier.go:207-227): sum_i left_i (.) Encode_i - sum_i chal_i (.) Partial_i - PartialMask == 0
        memset(&G, 0, sizeof(G));
        G.O = 1; G.C = 1; G.K = rows; G.key_os = 0; G.key_ks = 1; G.x_cs = rows; G.x_ks = 1; G.out_cs = 1; G.out_os = 0;
        RS_V(launch_matvec(q_t1, d_left, d_enc, nullptr, G, j->rq));
        G.K = cols; G.x_cs = cols;
        RS_V(launch_matvec(q_t2, d_chal, d_part, nullptr, G, j->rq));
        if (rc == RS_OK) {
            nonzero_kernel<<<(unsigned)((wq + 255) / 256), 256, 0, st>>>(d_flag, q_t1, q_t2, d_part + (size_t)cols * wq, wq, Gq);
            count_launch();
        }
        // ---- evaluation (verifier.go:229-260): centred coefficients of the Partial (and batch challenge) polynomials
        RS_V(launch_recon(j, false, false, nullptr, d_part, (size_t)cols, d_ints, nullptr));
        if (batch > 1) RS_V(launch_recon(j, false, false, nullptr, d_bq, (size_t)batch, d_ints + (size_t)cols * JD * JW, nullptr));
    }
    unsigned h_flag = 1;
    if (rc == RS_OK) {
        h_sq.resize(n_sq_polys * 2 * JW); h_ints.resize((size_t)(cols + batch) * JD * JW);
        if (cudaMemcpyAsync(h_sq.data(), d_sq, h_sq.size() * 8, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
            cudaMemcpyAsync(h_ints.data(), d_ints, (size_t)(cols + (batch > 1 ? batch : 0)) * JD * JW * 8, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
            cudaMemcpyAsync(&h_flag, d_flag, 4, cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess)
            rc = set_error(RS_ERR_CUDA, "rs_jindo_verify: %s", cudaGetErrorString(cudaGetLastError()));
    }
    for (u64 *ptr : {d_part, d_enc, d_ml, d_inc, d_coms, d_chal, d_bq, d_bo, d_left, d_tmp, d_sq, d_ints}) if (ptr) cudaFreeAsync(ptr, st);
    if (d_flag) cudaFreeAsync(d_flag, st);
#undef RS_V
#undef RS_VA
#undef RS_VU
    if (rc) return rc;
    // ---- host tail: norms, and the decode + inner product of verifyEval in the committed field
    u64 tot[2 * JW + 2];
    sq_total(h_sq.data(), n_com + p.out_msis, tot);
    if (isqrt_lt(tot, 2 * JW + 2, in_com_dcmp_two_nm)) *checks |= 1;                  // verifyOuterCommitment
    sq_total(h_sq.data() + (n_com + p.out_msis) * 2 * JW, (size_t)rows + nm + p.in_msis, tot);
    if (isqrt_lt(tot, 2 * JW + 2, res_two_nm)) *checks |= 2;                          // verifyInnerCommitment
    if (h_flag == 0) *checks |= 4;                                                     // verifyConsistency
    {
        const rs_field *f = j->f;
        const int L64 = f->L64;
        auto to_field = [&](const u64 *V, uint32_t *out /* Montgomery */) {          // SetBigInt of a centred JW-word integer
            const bool neg = (V[JW - 1] >> 63) != 0;
            u64 M[JW];
            if (neg) { u64 cy = 1; for (int i = 0; i < JW; i++) { const u64 t = ~V[i] + cy; cy = (cy && t == 0); M[i] = t; } }
            else for (int i = 0; i < JW; i++) M[i] = V[i];
            uint32_t acc[RS_MAX_NL] = {0}, chunk[RS_MAX_NL], t[RS_MAX_NL];
            bool one_word = true;
            for (int i = 1; i < JW; i++) one_word &= (M[i] == 0);
            if (one_word) {                                                            // |V| < 2^64 (every honest digit): one product
                memset(chunk, 0, sizeof(chunk));
                chunk[0] = (uint32_t)M[0]; chunk[1] = (uint32_t)(M[0] >> 32);
                host_mul(f, acc, chunk, f->r2);
            } else {
            const int n_chunks = (JW + L64 - 1) / L64;
            for (int ch = n_chunks - 1; ch >= 0; ch--) {                               // value = sum chunk_k R^k, Horner in Montgomery form
                host_mul(f, acc, acc, f->r2);                                          // acc * R
                memset(chunk, 0, sizeof(chunk));
                for (int w = 0; w < L64 && ch * L64 + w < JW; w++) { chunk[2 * w] = (uint32_t)M[ch * L64 + w]; chunk[2 * w + 1] = (uint32_t)(M[ch * L64 + w] >> 32); }
                host_mul(f, t, chunk, f->r2);                                          // to Montgomery (chunk < R; result canonical)
                host_add(f, acc, acc, t);
            }
            }
            if (neg) { uint32_t z[RS_MAX_NL] = {0}; bool any = false; for (int i = 0; i < NL; i++) any |= acc[i] != 0;
                       if (any) { u64 br = 0; for (int i = 0; i < NL; i++) { const u64 dd = (u64)f->q[i] - acc[i] - br; z[i] = (uint32_t)dd; br = (dd >> 63) & 1; } memcpy(acc, z, 4 * NL); } }
            memcpy(out, acc, 4 * NL);
        };
        uint32_t base_m[RS_MAX_NL];
        host_set_u64(f, base_m, p.base);
        auto decode_slot = [&](const u64 *poly_ints, int slot, uint32_t *out) {       // Encoder.DecodeTo (encoder.go:204-219)
            uint32_t acc[RS_MAX_NL] = {0}, c[RS_MAX_NL];
            for (int e = p.exp - 1; e >= 0; e--) {
                host_mul(f, acc, acc, base_m);
                to_field(poly_ints + (size_t)(e * p.slots + slot) * JW, c);
                host_add(f, acc, acc, c);
            }
            memcpy(out, acc, 4 * NL);
        };
        uint32_t yb[RS_MAX_NL] = {0}, t[RS_MAX_NL], dc[RS_MAX_NL];
        if (batch > 1) {
            for (int b = 0; b < batch; b++) {
                decode_slot(h_ints.data() + (size_t)(cols + b) * JD * JW, 0, dc);
                host_mul(f, t, dc, (const uint32_t *)(ys_mont + (size_t)b * L64));
                host_add(f, yb, yb, t);
            }
        } else memcpy(yb, ys_mont, 4 * NL);
        uint32_t test[RS_MAX_NL] = {0}, xp[RS_MAX_NL];
        memcpy(xp, f->one, 4 * NL);                                                    // rightVec[k] = x^k (utils.go:75-82)
        for (int i = 0; i < cols; i++)
            for (int s2 = 0; s2 < p.slots; s2++) {
                decode_slot(h_ints.data() + (size_t)i * JD * JW, s2, dc);
                host_mul(f, t, xp, dc);
                host_add(f, test, test, t);
                host_mul(f, xp, xp, (const uint32_t *)x_mont);
            }
        if (memcmp(test, yb, 4 * NL) == 0) *checks |= 8;                                // verifyEval
    }
    return RS_OK;
}

static void put_u64le(uint8_t *&o, uint64_t v) { for (int i = 0; i < 8; i++) *o++ = (uint8_t)(v >> (8 * i)); }

long long rs_jindo_poly_write_to(const uint64_t *poly, int n_limbs, int d, uint8_t *out, size_t cap) {
    if ((!poly && out) || n_limbs < 0 || d < 0) return rs::set_error(RS_ERR_ARG, "rs_jindo_poly_write_to: bad argument");
    const size_t need = 8 + (size_t)n_limbs * (8 + (size_t)d * 8);
    if (!out) return (long long)need;
    if (cap < need) return rs::set_error(RS_ERR_ARG, "rs_jindo_poly_write_to: buffer too small");
    uint8_t *o = out;
    put_u64le(o, (uint64_t)n_limbs);
    for (int l = 0; l < n_limbs; l++) {
        put_u64le(o, (uint64_t)d);
        for (int c = 0; c < d; c++) put_u64le(o, poly[(size_t)l * d + c]);
    }
    return (long long)need;
}

long long rs_jindo_commitment_write_raw(const rs_jindo *j, const uint64_t *commitment, uint8_t *out, size_t cap) {
    if (!j || (!commitment && out)) return rs::set_error(RS_ERR_ARG, "rs_jindo_commitment_write_raw: null argument");
    const rs_jindo_params &p = j->p;
    const int nl = p.n_q > p.n_qout ? p.n_q : p.n_qout, d = JD;
    const size_t per = 8 + (size_t)nl * (8 + (size_t)d * 8), need = per * p.out_msis;
    if (!out) return (long long)need;
    if (cap < need) return rs::set_error(RS_ERR_ARG, "rs_jindo_commitment_write_raw: buffer too small");
    uint8_t *o = out;
    for (int i = 0; i < p.out_msis; i++) {
        put_u64le(o, (uint64_t)nl);
        for (int l = 0; l < nl; l++) {
            put_u64le(o, (uint64_t)d);
            for (int c = 0; c < d; c++) put_u64le(o, l < p.n_qout ? commitment[((size_t)i * p.n_qout + l) * d + c] : 0);
        }
    }
    return (long long)need;
}

int rs_opening_alloc_zero(rs_jindo *j, rs_opening **out) {
    RS_REQUIRE(j && out, RS_ERR_ARG, "rs_opening_alloc_zero: null argument");
    return opening_alloc(j, out);
}

// every polynomial of one kind in one copy: kind 2 = InCommit[(cols+1)*in_msis] (what Proof.InCommit carries)
int rs_opening_download_all(const rs_opening *o, int kind, uint64_t *out) {
    RS_REQUIRE(o && out && kind == 2, RS_ERR_ARG, "rs_opening_download_all: bad argument");
    const rs_jindo *j = o->j;
    const size_t words = (size_t)(j->p.cols + 1) * j->p.in_msis * pqo(j);
    cudaError_t e = cudaMemcpyAsync(out, o->d_incommit, words * 8, cudaMemcpyDeviceToHost, rs::current_stream());
    if (e == cudaSuccess) e = cudaStreamSynchronize(rs::current_stream());
    if (e != cudaSuccess) return rs::set_error(RS_ERR_CUDA, "rs_opening_download_all: %s", cudaGetErrorString(e));
    return RS_OK;
}

int rs_opening_download(const rs_opening *o, int kind, size_t col, size_t idx, uint64_t *poly_out) {
    RS_REQUIRE(o && poly_out, RS_ERR_ARG, "rs_opening_download: null argument");
    const rs_jindo *j = o->j;
    const rs_jindo_params &p = j->p;
    const u64 *src; size_t words;
    if (kind == 0) {
        RS_REQUIRE(col <= (size_t)p.cols && idx < (size_t)p.rows, RS_ERR_SHAPE, "rs_opening_download: out of range");
        words = pq(j); src = o->d_encode + (col * p.rows + idx) * words;
    } else if (kind == 1) {
        RS_REQUIRE(col <= (size_t)p.cols && idx < (size_t)(p.mlwe + p.in_msis), RS_ERR_SHAPE, "rs_opening_download: out of range");
        words = pq(j); src = o->d_mlwe + (col * (p.mlwe + p.in_msis) + idx) * words;
    } else if (kind == 2) {
        RS_REQUIRE(idx < (size_t)(p.cols + 1) * p.in_msis, RS_ERR_SHAPE, "rs_opening_download: out of range");
        words = pqo(j); src = o->d_incommit + idx * words;
    } else {
        return set_error(RS_ERR_ARG, "rs_opening_download: bad kind");
    }
    RS_CUDA_CHECK(cudaMemcpyAsync(poly_out, src, words * 8, cudaMemcpyDeviceToHost, current_stream()));
    RS_CUDA_CHECK(cudaStreamSynchronize(current_stream()));
    return RS_OK;
}

// vec: the committed vector in HBM, or nullptr with v_host set (then it is uploaded piecewise, see DescCache)
static int outer_commit(rs_jindo *j, const rs_opening *o, uint64_t *commitment_out, bool sync = true);

// col0/col1: the columns of the committed matrix to process (the mask column is index cols); a whole Commit
// is [0, cols+1) and ends with the outer commit, a partial one leaves InCommit[col0*in_msis, col1*in_msis) filled
// sync == false (rs_jindo_commit_async): nothing waits for the stream; commitment_out must be pinned and, like a host-resident
// tape, stay untouched until the caller has synchronised
static int commit_impl(rs_jindo *j, const uint32_t *vec, const uint64_t *v_host, size_t len, const rs_jindo_tape *tape,
                       rs_opening **open, uint64_t *commitment_out, int col0, int col1, bool sync = true) {
    const rs_jindo_params &p = j->p;
    RS_REQUIRE(col0 >= 0 && col0 < col1 && col1 <= p.cols + 1, RS_ERR_ARG, "rs_jindo_commit: bad column range");
    const bool whole = (commitment_out != nullptr);       // rs_jindo_commit_columns never runs the outer commit
    const size_t rank = (size_t)p.cols * (p.rows - 1) * p.slots;
    RS_REQUIRE(len <= rank, RS_ERR_SHAPE, "len(v) > params.rank");                    // prover.go:47-49
    RS_REQUIRE(len >= 1, RS_ERR_SHAPE, "rs_jindo_commit: vector range out of bounds");
    RS_REQUIRE(tape->first_row && tape->last_row && tape->mask_rows && tape->mlwe, RS_ERR_ARG, "rs_jindo_commit: incomplete tape");
    const bool gen_cosac = (tape->cosac_g == nullptr);
    if (gen_cosac) RS_REQUIRE(j->cosac_stream, RS_ERR_ARG, "rs_jindo_commit: tape carries no COSAC samples and no stream was set (rs_jindo_set_cosac_stream)");
    const int NL = j->f->NL, d = JD, slots = p.slots, rows = p.rows, cols = p.cols;
    const size_t cs = (size_t)cols * slots;

    // ---- descriptors, in the order of commitColTo (prover.go:89-128); cached per committed length
    const DescCache *dc = nullptr;
    for (auto &c : *j->cache) if (c.len == len && c.c0 == col0 && c.c1 == col1) dc = &c;
    if (!dc) {
        std::vector<EncDesc> descs;
        size_t twin = 0, cosac = 0, mask = 0;
        auto add = [&](int kind, size_t off, int cnt, int nk, size_t noff, size_t col, size_t row) {
            // tape offsets are global (the counters advance for every column): a column shard consumes exactly
            // the draws the whole Commit would give it
            if ((int)col < col0 || (int)col >= col1) return;
            descs.push_back(EncDesc{kind, cnt, (i64)off, nk, (i64)(noff * d), (i64)(col * rows + row)});
        };
        for (int i = 0; i < cols; i++) add(1, (size_t)i * slots, slots, 1, (size_t)i, i, 0);   // firstRow, COSAC (ecdBlindStdDev)
        cosac = cols;
        for (int i = 0; i < cols; i++) {
            for (int r = 1; r < rows - 1; r++) {
                const size_t s0 = (size_t)r * cs + (size_t)i * slots, e0 = s0 + slots;
                if (s0 > len) break;                                                     // note '>' (prover.go:121)
                const size_t hi = e0 < len ? e0 : len;
                add(0, s0, (int)(hi - s0), 0, twin++, i, r);
            }
            add(1, cs + (size_t)i * slots, slots, 0, twin++, i, rows - 1);               // lastRow, TwinCDT (ecdStdDev)
        }
        // mask column (prover.go:93-115): every row COSAC, fresh random elements per row
        add(1, 2 * cs + (mask++) * slots, slots, 1, cosac++, cols, 0);
        for (int r = 1; r < rows - 1; r++) {
            if ((size_t)r * cs > len) break;
            add(1, 2 * cs + (mask++) * slots, slots, 1, cosac++, cols, r);
        }
        add(1, 2 * cs + (mask++) * slots, slots, 1, cosac++, cols, rows - 1);
        // row-major order (every descriptor is independent: sources, noise offsets and the output slot are
        // explicit), so that consecutive descriptors read consecutive pieces of v
        std::stable_sort(descs.begin(), descs.end(), [rows](const EncDesc &a, const EncDesc &b) {
            const i64 ra = a.out_poly % rows, rb = b.out_poly % rows;
            return ra != rb ? ra < rb : a.out_poly < b.out_poly;
        });
        DescCache nc{};
        nc.len = len; nc.c0 = col0; nc.c1 = col1; nc.d_descs = nullptr; nc.n_descs = descs.size(); nc.twin = twin; nc.cosac = cosac; nc.mask = mask;
        {
            size_t hi = 0;
            for (int k = 0; k < RS_VCHUNKS; k++) {
                nc.chunk_desc[k] = descs.size() * k / RS_VCHUNKS;
                const size_t e = descs.size() * (k + 1) / RS_VCHUNKS;
                for (size_t t = nc.chunk_desc[k]; t < e; t++)
                    if (descs[t].src_kind == 0 && (size_t)(descs[t].src_off + descs[t].cnt) > hi) hi = (size_t)(descs[t].src_off + descs[t].cnt);
                nc.chunk_vhi[k] = hi;
            }
            nc.chunk_desc[RS_VCHUNKS] = descs.size();
        }
        RS_CUDA_CHECK(cudaMalloc((void **)&nc.d_descs, descs.size() * sizeof(EncDesc)));
        RS_CUDA_CHECK(cudaMemcpy(nc.d_descs, descs.data(), descs.size() * sizeof(EncDesc), cudaMemcpyHostToDevice));
        if (j->cache->size() >= 64) { cudaFree(j->cache->front().d_descs); j->cache->erase(j->cache->begin()); }
        j->cache->push_back(nc);
        dc = &j->cache->back();
    }
    const size_t twin = dc->twin, cosac = dc->cosac, mask = dc->mask, n_descs = dc->n_descs;
    const size_t n_noise = (size_t)(cols + 1) * (p.in_msis + p.mlwe);
    if (!j->d_ndescs) {
        std::vector<EncDesc> ndescs(n_noise);
        for (size_t k = 0; k < n_noise; k++) ndescs[k] = EncDesc{2, 0, 0, 1, (i64)(k * d), (i64)k};
        RS_CUDA_CHECK(cudaMalloc((void **)&j->d_ndescs, n_noise * sizeof(EncDesc)));
        RS_CUDA_CHECK(cudaMemcpy(j->d_ndescs, ndescs.data(), n_noise * sizeof(EncDesc), cudaMemcpyHostToDevice));
    }
    RS_REQUIRE(tape->n_mask_rows == mask, RS_ERR_SHAPE, "rs_jindo_commit: tape has the wrong number of mask rows");
    RS_REQUIRE(gen_cosac || tape->n_cosac_polys == cosac, RS_ERR_SHAPE, "rs_jindo_commit: tape has the wrong number of COSAC polynomials");
    const bool gen_mlwe = tape->mlwe_mask_only != 0;
    if (gen_mlwe) RS_REQUIRE(j->mlwe_stream, RS_ERR_ARG, "rs_jindo_commit: tape has mlwe_mask_only set and no stream was set (rs_jindo_set_mlwe_stream)");
    const bool gen_twin = (tape->twin_u == nullptr && twin != 0);
    if (gen_twin)
        RS_REQUIRE(j->twin_stream, RS_ERR_ARG, "rs_jindo_commit: tape carries no TwinCDT draws and no stream was set (rs_jindo_set_twin_stream)");
    else
        RS_REQUIRE(tape->n_twin_u == twin * d, RS_ERR_SHAPE, "rs_jindo_commit: tape has the wrong number of TwinCDT draws");

    const bool reuse = (*open != nullptr);
    rs_opening *o = *open;
    int rc = RS_OK;
    if (!reuse) {
        rc = opening_alloc(j, &o);
        if (rc) return rc;
    } else {
        RS_REQUIRE(o->j == j, RS_ERR_ARG, "rs_jindo_commit: opening belongs to another prover");
        if (n_descs < (size_t)(col1 - col0) * rows)          // rows past len(v) must read as zero polynomials
            RS_CUDA_CHECK(cudaMemsetAsync(o->d_encode + (size_t)col0 * rows * pq(j), 0, (size_t)(col1 - col0) * rows * pq(j) * 8, current_stream()));
    }

    // ---- uploads (skipped when the tape is already device resident)
    const bool dev = tape->on_device != 0;
    uint32_t *d_aux = nullptr; u64 *d_twin = nullptr; bool twin_drawn = false, mlwe_owned = false, cosac_owned = false; i64 *d_cosac = nullptr, *d_mlwe_s = nullptr;
    u64 *d_com = nullptr;
    uint32_t *d_v = nullptr;          // device copy of a host-resident v
#define RS_TRY(x) do { rc = (x); if (rc) goto fail; } while (0)
    {
        const size_t aux_words = (2 * cs + mask * slots) * NL;
        if (cudaMallocAsync((void **)&d_aux, aux_words * 4, current_stream()) != cudaSuccess) { rc = set_error(RS_ERR_CUDA, "rs_jindo_commit: cudaMallocAsync failed"); goto fail; }
        const cudaMemcpyKind kind = dev ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
        cudaMemcpyAsync(d_aux, tape->first_row, cs * NL * 4, kind, current_stream());
        cudaMemcpyAsync(d_aux + cs * NL, tape->last_row, cs * NL * 4, kind, current_stream());
        cudaMemcpyAsync(d_aux + 2 * cs * NL, tape->mask_rows, mask * slots * NL * 4, kind, current_stream());
    }
    if (dev && !gen_mlwe) {
        d_twin = (u64 *)tape->twin_u; d_cosac = (i64 *)tape->cosac_g; d_mlwe_s = (i64 *)tape->mlwe;
    } else {
        if (dev) { d_twin = (u64 *)tape->twin_u; d_cosac = (i64 *)tape->cosac_g; }
        else {
            if (!gen_twin) RS_TRY(upload(&d_twin, (const u64 *)tape->twin_u, twin * d));
            if (!gen_cosac) { RS_TRY(upload(&d_cosac, (const i64 *)tape->cosac_g, cosac * d)); cosac_owned = true; }
        }
        if (gen_mlwe) {
            // data columns: Prover.mlweSampler on the device (the whole Commit's draws, so that a column shard sees the
            // stream positions the unsharded Commit gives its columns); mask column: the host's roundedSampler samples
            const size_t nmd = (size_t)(p.in_msis + p.mlwe) * d;
            if (cudaMallocAsync((void **)&d_mlwe_s, n_noise * d * 8, current_stream()) != cudaSuccess) { rc = set_error(RS_ERR_CUDA, "rs_jindo_commit: cudaMallocAsync failed"); goto fail; }
            mlwe_owned = true;
            cudaMemcpyAsync(d_mlwe_s + (size_t)cols * nmd, tape->mlwe, nmd * 8, dev ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, current_stream());
            RS_TRY(mlwe_stream_sample(j, (size_t)cols * nmd, d_mlwe_s));
        } else {
            RS_TRY(upload(&d_mlwe_s, (const i64 *)tape->mlwe, n_noise * d));
            mlwe_owned = true;
        }
    }
    if (gen_cosac) {
        // the COSAC rows (row 0 of the data columns, every row of the mask column) sampled in HBM from the elements just uploaded:
        // all of them, also for a column shard (the worker numbering is global, like the tape offsets)
        if (cudaMallocAsync((void **)&d_cosac, cosac * d * 8, current_stream()) != cudaSuccess) { rc = set_error(RS_ERR_CUDA, "rs_jindo_commit: cudaMallocAsync failed"); goto fail; }
        cosac_owned = true;
        RS_TRY(rs::cosac_generate(j->cosac_stream, j->f, p, d_aux, d_aux + 2 * cs * NL, cosac, d_cosac));
    }
    if (v_host) {
        // upload v in RS_VCHUNKS pieces on the copy stream; piece k is encoded as soon as it has landed.  Enqueued right behind
        // the (small) tape uploads and BEFORE the main stream starts waiting for the sampler's AES kernels: the 134 MB of v
        // must not queue up behind that wait.
        if (cudaMallocAsync((void **)&d_v, len * NL * 4, current_stream()) != cudaSuccess) { rc = set_error(RS_ERR_CUDA, "rs_jindo_commit_host: cudaMallocAsync failed"); goto fail; }
        cudaEventRecord(j->ev_alloc, current_stream());
        cudaStreamWaitEvent(j->copy_stream, j->ev_alloc, 0);
        size_t lo = 0;
        for (int k = 0; k < RS_VCHUNKS; k++) {
            const size_t hi = dc->chunk_vhi[k];
            if (hi > lo) cudaMemcpyAsync(d_v + lo * NL, (const uint32_t *)v_host + lo * NL, (hi - lo) * NL * 4, cudaMemcpyHostToDevice, j->copy_stream);
            if (hi > lo) lo = hi;
            cudaEventRecord(j->ev_piece[k], j->copy_stream);
        }
    }
    if (gen_twin) {       // draws come from the prover's own AES-CTR stream, generated in HBM
        const u64 *smp = nullptr;
        RS_TRY(uniform_stream_draw(j->twin_stream, twin * d, &smp));
        twin_drawn = true;
        d_twin = (u64 *)smp;
    }
    {
        const size_t n_com = (size_t)(cols + 1) * p.in_msis;
        const int nc = col1 - col0, nm = p.mlwe + p.in_msis;
        if (cudaMallocAsync((void **)&d_com, n_com * pq(j) * 8, current_stream()) != cudaSuccess) {
            rc = set_error(RS_ERR_CUDA, "rs_jindo_commit: cudaMallocAsync failed");
            goto fail;
        }
        // ---- encodes + noise polynomials
        if (v_host) {
            for (int k = 0; k < RS_VCHUNKS; k++) {
                cudaStreamWaitEvent(current_stream(), j->ev_piece[k], 0);
                const size_t c0 = dc->chunk_desc[k], c1 = dc->chunk_desc[k + 1];
                if (c1 > c0) RS_TRY(encode_dispatch(j, o->d_encode, dc->d_descs + c0, c1 - c0, d_v, d_aux, d_twin, d_cosac));
            }
            vec = d_v;
        } else {
            RS_TRY(encode_dispatch(j, o->d_encode, dc->d_descs, n_descs, vec, d_aux, d_twin, d_cosac));
        }
        RS_TRY(encode_dispatch(j, o->d_mlwe, j->d_ndescs + (size_t)col0 * nm, (size_t)nc * nm, vec, d_aux, d_twin, d_mlwe_s));
        // ---- inner commit (prover.go:144-157) of the columns [col0, col1)
        MatvecGeom G;
        memset(&G, 0, sizeof(G));
        G.O = p.in_msis; G.C = nc; G.K = rows;
        G.key_os = rows; G.key_ks = 1; G.x_cs = rows; G.x_ks = 1; G.out_cs = p.in_msis; G.out_os = 1;
        u64 *com0 = d_com + (size_t)col0 * p.in_msis * pq(j);
        RS_TRY(launch_matvec(com0, j->d_ck_in, o->d_encode + (size_t)col0 * rows * pq(j), nullptr, G, j->rq));
        G.K = p.mlwe; G.key_os = p.mlwe; G.x_cs = nm; G.accumulate = 1;
        G.has_addend = 1; G.add_cs = nm; G.add_os = 1;
        {
            const u64 *ml0 = o->d_mlwe + (size_t)col0 * nm * pq(j);
            RS_TRY(launch_matvec(com0, j->d_ck_mlwe, ml0, ml0 + (size_t)p.mlwe * pq(j), G, j->rq));
        }
        // ---- cut-off pipeline (prover.go:164-176)
        RS_TRY(launch_cut(j, o->d_incommit + (size_t)col0 * p.in_msis * pqo(j), com0, (size_t)nc * p.in_msis, false));
        if (whole) RS_TRY(outer_commit(j, o, commitment_out, sync));
        // a column shard ends without the outer commit: host-resident tape buffers must have been consumed before the caller
        // may reuse them; with a device-resident tape nothing needs the host to wait (the exchange runs on the same stream)
        else if (!dev && cudaStreamSynchronize(current_stream()) != cudaSuccess) { rc = set_error(RS_ERR_CUDA, "rs_jindo_commit_columns: %s", cudaGetErrorString(cudaGetLastError())); goto fail; }
    }
fail:
#undef RS_TRY
    cudaFreeAsync(d_aux, current_stream());
    if (d_v) cudaFreeAsync(d_v, current_stream());
    if (!dev && !gen_twin) cudaFreeAsync(d_twin, current_stream());
    if (cosac_owned) cudaFreeAsync(d_cosac, current_stream());
    if (mlwe_owned) cudaFreeAsync(d_mlwe_s, current_stream());
    if (twin_drawn) {
        uniform_stream_release(j->twin_stream);                  // every reader of the draws has been launched
        // optional (RS_JINDO_PREFETCH=1): generate the next Commit's draws now, behind this Commit's remaining kernels.
        // Measured neutral at 2^22 (5.9 ms e2e either way: the AES kernels take the same SM time wherever they run), so
        // it is off by default; it pays when the host leaves gaps between Commits.
        if (rc == RS_OK && j->prefetch) uniform_stream_prefetch(j->twin_stream, twin * d);
    }
    cudaFreeAsync(d_com, current_stream());
    if (rc) { if (!reuse) rs_opening_free(o); return rc; }
    *open = o;
    return RS_OK;
}

// outer commit (prover.go:180-202) from Opening.InCommit: split over the cols+1 column blocks, partial sums
// reduced mod q afterwards (same field element as the sequential sum), then the cut-off pipeline
static int outer_commit(rs_jindo *j, const rs_opening *o, uint64_t *commitment_out, bool sync) {
    const rs_jindo_params &p = j->p;
    const int cols = p.cols;
    const size_t n_com = (size_t)(cols + 1) * p.in_msis;
    u64 *d_com2 = nullptr, *d_value = nullptr, *d_part = nullptr;
    int rc = RS_OK;
    if (cudaMallocAsync((void **)&d_com2, (size_t)p.out_msis * pqo(j) * 8, current_stream()) != cudaSuccess ||
        cudaMallocAsync((void **)&d_part, (size_t)(cols + 1) * p.out_msis * pqo(j) * 8, current_stream()) != cudaSuccess ||
        cudaMallocAsync((void **)&d_value, (size_t)p.out_msis * pqo(j) * 8, current_stream()) != cudaSuccess)
        rc = set_error(RS_ERR_CUDA, "rs_jindo_commit: cudaMallocAsync failed");
    if (rc == RS_OK) {
        MatvecGeom G;
        memset(&G, 0, sizeof(G));
        G.O = p.out_msis; G.C = cols + 1; G.K = p.in_msis;
        G.key_os = (i64)n_com; G.key_cs = p.in_msis; G.key_ks = 1; G.x_cs = p.in_msis; G.x_ks = 1;
        G.out_cs = p.out_msis; G.out_os = 1;
        rc = launch_matvec(d_part, j->d_ck_out, o->d_incommit, nullptr, G, j->rqo);
    }
    if (rc == RS_OK) {
        MatvecGeom Gr;
        memset(&Gr, 0, sizeof(Gr));
        fill_geom_ring(Gr, j->rqo);
        const size_t tot = (size_t)p.out_msis * pqo(j);
        reduce_cols_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, current_stream()>>>(d_com2, d_part, cols + 1, p.out_msis, Gr);
        count_launch();
        rc = launch_cut(j, d_value, d_com2, p.out_msis, true);
    }
    if (rc == RS_OK &&
        (cudaMemcpyAsync(commitment_out, d_value, (size_t)p.out_msis * pqo(j) * 8, cudaMemcpyDeviceToHost, current_stream()) != cudaSuccess ||
         (sync && cudaStreamSynchronize(current_stream()) != cudaSuccess)))
        rc = set_error(RS_ERR_CUDA, "rs_jindo_commit: %s", cudaGetErrorString(cudaGetLastError()));
    cudaFreeAsync(d_part, current_stream()); cudaFreeAsync(d_com2, current_stream()); cudaFreeAsync(d_value, current_stream());
    return rc;
}

int rs_jindo_commit(rs_jindo *j, const rs_batch *v, size_t v_poly, size_t len, const rs_jindo_tape *tape,
                    rs_opening **open, uint64_t *commitment_out) {
    RS_REQUIRE(j && v && tape && open && commitment_out, RS_ERR_ARG, "rs_jindo_commit: null argument");
    RS_REQUIRE(len >= 1 && v_poly < v->n_polys && len <= v->rank, RS_ERR_SHAPE, "rs_jindo_commit: vector range out of bounds");
    RS_REQUIRE(v->f->NL == j->f->NL, RS_ERR_SHAPE, "rs_jindo_commit: field mismatch");
    return commit_impl(j, v->d + v_poly * v->rank * j->f->NL, nullptr, len, tape, open, commitment_out, 0, j->p.cols + 1);
}

int rs_jindo_commit_async(rs_jindo *j, const rs_batch *v, size_t v_poly, size_t len, const rs_jindo_tape *tape,
                          rs_opening **open, uint64_t *commitment_out_pinned) {
    RS_REQUIRE(j && v && tape && open && commitment_out_pinned, RS_ERR_ARG, "rs_jindo_commit_async: null argument");
    RS_REQUIRE(len >= 1 && v_poly < v->n_polys && len <= v->rank, RS_ERR_SHAPE, "rs_jindo_commit_async: vector range out of bounds");
    RS_REQUIRE(v->f->NL == j->f->NL, RS_ERR_SHAPE, "rs_jindo_commit_async: field mismatch");
    return commit_impl(j, v->d + v_poly * v->rank * j->f->NL, nullptr, len, tape, open, commitment_out_pinned, 0, j->p.cols + 1, false);
}

int rs_jindo_commit_columns(rs_jindo *j, const rs_batch *v, size_t v_poly, size_t len, const rs_jindo_tape *tape,
                            int col_begin, int col_end, rs_opening **open) {
    RS_REQUIRE(j && v && tape && open, RS_ERR_ARG, "rs_jindo_commit_columns: null argument");
    RS_REQUIRE(len >= 1 && v_poly < v->n_polys && len <= v->rank, RS_ERR_SHAPE, "rs_jindo_commit_columns: vector range out of bounds");
    RS_REQUIRE(v->f->NL == j->f->NL, RS_ERR_SHAPE, "rs_jindo_commit_columns: field mismatch");
    return commit_impl(j, v->d + v_poly * v->rank * j->f->NL, nullptr, len, tape, open, nullptr, col_begin, col_end);
}

int rs_jindo_outer_commit(rs_jindo *j, const rs_opening *open, uint64_t *commitment_out) {
    RS_REQUIRE(j && open && commitment_out, RS_ERR_ARG, "rs_jindo_outer_commit: null argument");
    RS_REQUIRE(open->j == j, RS_ERR_ARG, "rs_jindo_outer_commit: opening belongs to another prover");
    return outer_commit(j, open, commitment_out);
}

void *rs_opening_incommit_device(rs_opening *open) { return open ? (void *)open->d_incommit : nullptr; }

// ---- one Commit over several GPUs (SURVEY 8e iii) ------------------------------------------------------------------
// Columns are dealt in equal contiguous blocks of ceil((cols+1) / world): every rank except possibly the last owns a full
// block, so the ranks' InCommit slices, laid end to end, ARE the InCommit array and the exchange is one in-place all-gather.
int rs_jindo_shard_columns(const rs_jindo *j, int rank, int world, int *col_begin, int *col_end) {
    RS_REQUIRE(j && col_begin && col_end && world >= 1 && world <= RS_MAX_SHARDS && rank >= 0 && rank < world, RS_ERR_ARG,
               "rs_jindo_shard_columns: bad argument");
    const int n = j->p.cols + 1, per = (n + world - 1) / world;
    const int b = rank * per < n ? rank * per : n, e = (rank + 1) * per < n ? (rank + 1) * per : n;
    *col_begin = b; *col_end = e;
    return RS_OK;
}

namespace rs {
__global__ void mod_reduce_kernel(u64 *__restrict__ x, size_t n, const __grid_constant__ MatvecGeom G) {
    const size_t pw = (size_t)G.n_limbs * JD;
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    x[i] %= G.q[(i % pw) / JD];
}
}  // namespace rs

// Outer commit of a column-sharded Commit without gathering InCommit first: every rank multiplies ck.Out by ITS columns'
// InCommit polynomials (prover.go:186-189 is a sum over j, split here by column block), the out_msis partial sums (a few
// KiB) are all-reduced as plain uint64 sums (world * q < 2^64) and reduced mod q, and the cut-off pipeline (:190-200) runs
// replicated.  The same field elements as the sequential sum, so Commitment.Value is bit-identical to the unsharded one.
int rs_jindo_outer_commit_sharded(rs_jindo *j, const rs_opening *o, int rank, int world, void *nccl_comm, uint64_t *commitment_out) {
    RS_REQUIRE(j && o && o->j == j && nccl_comm && commitment_out, RS_ERR_ARG, "rs_jindo_outer_commit_sharded: bad argument");
    int c0 = 0, c1 = 0;
    int rc = rs_jindo_shard_columns(j, rank, world, &c0, &c1);
    if (rc) return rc;
    const rs_jindo_params &p = j->p;
    for (int l = 0; l < j->rqo.n; l++)
        RS_REQUIRE(j->rqo.q[l] < (~0ull) / (u64)world, RS_ERR_UNSUPPORTED, "rs_jindo_outer_commit_sharded: world * q overflows 64 bits");
    const size_t n_com = (size_t)(p.cols + 1) * p.in_msis, tot = (size_t)p.out_msis * pqo(j);
    const int nc = c1 - c0;
    u64 *d_part = nullptr, *d_sum = nullptr, *d_value = nullptr;
    if (cudaMallocAsync((void **)&d_sum, tot * 8, current_stream()) != cudaSuccess ||
        cudaMallocAsync((void **)&d_value, tot * 8, current_stream()) != cudaSuccess ||
        (nc > 0 && cudaMallocAsync((void **)&d_part, (size_t)nc * tot * 8, current_stream()) != cudaSuccess))
        rc = set_error(RS_ERR_CUDA, "rs_jindo_outer_commit_sharded: cudaMallocAsync failed");
    MatvecGeom Gr;
    memset(&Gr, 0, sizeof(Gr));
    fill_geom_ring(Gr, j->rqo);
    if (rc == RS_OK) {
        if (nc > 0) {
            MatvecGeom G;
            memset(&G, 0, sizeof(G));
            G.O = p.out_msis; G.C = nc; G.K = p.in_msis;
            G.key_os = (i64)n_com; G.key_cs = p.in_msis; G.key_ks = 1; G.x_cs = p.in_msis; G.x_ks = 1;
            G.out_cs = p.out_msis; G.out_os = 1;
            rc = launch_matvec(d_part, j->d_ck_out + (size_t)c0 * p.in_msis * pqo(j), o->d_incommit + (size_t)c0 * p.in_msis * pqo(j), nullptr, G, j->rqo);
            if (rc == RS_OK) {
                reduce_cols_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, current_stream()>>>(d_sum, d_part, nc, p.out_msis, Gr);
                count_launch();
            }
        } else if (cudaMemsetAsync(d_sum, 0, tot * 8, current_stream()) != cudaSuccess) {
            rc = set_error(RS_ERR_CUDA, "rs_jindo_outer_commit_sharded: memset failed");
        }
    }
    if (rc == RS_OK) rc = nccl_allreduce_sum_u64(nccl_comm, (uint64_t *)d_sum, tot);
    if (rc == RS_OK) {
        mod_reduce_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, current_stream()>>>(d_sum, tot, Gr);
        count_launch();
        rc = launch_cut(j, d_value, d_sum, p.out_msis, true);
    }
    if (rc == RS_OK && (cudaMemcpyAsync(commitment_out, d_value, tot * 8, cudaMemcpyDeviceToHost, current_stream()) != cudaSuccess ||
                        cudaStreamSynchronize(current_stream()) != cudaSuccess))
        rc = set_error(RS_ERR_CUDA, "rs_jindo_outer_commit_sharded: %s", cudaGetErrorString(cudaGetLastError()));
    cudaFreeAsync(d_part, current_stream()); cudaFreeAsync(d_sum, current_stream()); cudaFreeAsync(d_value, current_stream());
    return rc;
}

// The InCommit polynomials of every shard, gathered in place in the Opening (Proof.InCommit / Evaluate need all of them):
// one ncclAllGather on the compute stream, no host synchronisation.
int rs_jindo_allgather_incommit(rs_jindo *j, rs_opening *o, int rank, int world, void *nccl_comm) {
    RS_REQUIRE(j && o && o->j == j && nccl_comm && world >= 1 && world <= RS_MAX_SHARDS && rank >= 0 && rank < world, RS_ERR_ARG,
               "rs_jindo_allgather_incommit: bad argument");
    const int per = (j->p.cols + 1 + world - 1) / world;
    return nccl_allgather_inplace_u64(nccl_comm, (uint64_t *)o->d_incommit, (size_t)per * j->p.in_msis * pqo(j), rank);
}

int rs_jindo_commit_host(rs_jindo *j, const uint64_t *v_host, size_t len, const rs_jindo_tape *tape,
                         rs_opening **open, uint64_t *commitment_out) {
    RS_REQUIRE(j && v_host && tape && open && commitment_out, RS_ERR_ARG, "rs_jindo_commit_host: null argument");
    RS_REQUIRE(!tape->on_device, RS_ERR_ARG, "rs_jindo_commit_host: the tape must be host resident like v");
    return commit_impl(j, nullptr, v_host, len, tape, open, commitment_out, 0, j->p.cols + 1);
}

static int partial_impl(rs_jindo *j, const rs_opening *o, const u64 *d_left, uint64_t *partial_out) {
    const rs_jindo_params &p = j->p;
    u64 *d_out = nullptr;
    RS_CUDA_CHECK(cudaMallocAsync((void **)&d_out, (size_t)(p.cols + 1) * pq(j) * 8, current_stream()));
    MatvecGeom G;
    memset(&G, 0, sizeof(G));
    G.O = 1; G.C = p.cols + 1; G.K = p.rows;
    G.key_os = 0; G.key_ks = 1; G.x_cs = p.rows; G.x_ks = 1; G.out_cs = 1; G.out_os = 0;
    int rc = launch_matvec(d_out, d_left, o->d_encode, nullptr, G, j->rq);
    if (rc == RS_OK) {
        cudaError_t e = cudaMemcpyAsync(partial_out, d_out, (size_t)(p.cols + 1) * pq(j) * 8, cudaMemcpyDeviceToHost, current_stream());
        if (e == cudaSuccess) e = cudaStreamSynchronize(current_stream());
        if (e != cudaSuccess) rc = set_error(RS_ERR_CUDA, "rs_jindo_partial: %s", cudaGetErrorString(e));
    }
    cudaFreeAsync(d_out, current_stream());
    return rc;
}

int rs_jindo_partial(rs_jindo *j, const rs_opening *o, const uint64_t *left, uint64_t *partial_out) {
    RS_REQUIRE(j && o && left && partial_out, RS_ERR_ARG, "rs_jindo_partial: null argument");
    RS_REQUIRE(o->j == j, RS_ERR_ARG, "rs_jindo_partial: opening belongs to another prover");
    u64 *d_left = nullptr;
    int rc = upload(&d_left, (const u64 *)left, (size_t)j->p.rows * pq(j));
    if (rc == RS_OK) rc = partial_impl(j, o, d_left, partial_out);
    cudaFreeAsync(d_left, current_stream());
    return rc;
}

extern "C++" {
template <int NL>
static int launch_left_vec(rs_jindo *j, uint32_t *d_elems, const uint32_t *d_x) {
    const rs_jindo_params &p = j->p;
    FieldParams<NL> F = make_params<NL>(j->f);
    left_vec_kernel<NL><<<(p.rows + 127) / 128, 128, 0, current_stream()>>>(d_elems, d_x, p.rows, (unsigned)(p.cols * p.slots), F);
    count_launch();
    RS_CUDA_CHECK(cudaGetLastError());
    return RS_OK;
}
}  // extern "C++"

// leftVec(x) encoded into rows polynomials over ringQ at *d_left (cudaMallocAsync'ed here)
static int encode_left_vec(rs_jindo *j, const uint64_t *x_mont, u64 **d_left) {
    const rs_jindo_params &p = j->p;
    const int NL = j->f->NL;
    *d_left = nullptr;
    if (!j->d_ldescs) {                                     // one descriptor per row: aux element i -> polynomial i, no noise
        std::vector<EncDesc> ld((size_t)p.rows);
        for (int i = 0; i < p.rows; i++) ld[i] = EncDesc{1, 1, (i64)i, 1, 0, (i64)i};
        RS_CUDA_CHECK(cudaMalloc((void **)&j->d_ldescs, ld.size() * sizeof(EncDesc)));
        RS_CUDA_CHECK(cudaMemcpy(j->d_ldescs, ld.data(), ld.size() * sizeof(EncDesc), cudaMemcpyHostToDevice));
        RS_CUDA_CHECK(cudaMalloc((void **)&j->d_zero_g, JD * sizeof(i64)));
        RS_CUDA_CHECK(cudaMemset(j->d_zero_g, 0, JD * sizeof(i64)));
    }
    uint32_t *d_elems = nullptr;
    RS_CUDA_CHECK(cudaMallocAsync((void **)&d_elems, ((size_t)p.rows + 1) * NL * 4, current_stream()));
    uint32_t *d_x = d_elems + (size_t)p.rows * NL;
    int rc = RS_OK;
    if (cudaMemcpyAsync(d_x, x_mont, (size_t)NL * 4, cudaMemcpyHostToDevice, current_stream()) != cudaSuccess)
        rc = set_error(RS_ERR_CUDA, "rs_jindo_partial_at: upload failed");
    if (rc == RS_OK) {
        switch (NL) {
            case 4: rc = launch_left_vec<4>(j, d_elems, d_x); break;
            case 8: rc = launch_left_vec<8>(j, d_elems, d_x); break;
            case 14: rc = launch_left_vec<14>(j, d_elems, d_x); break;
            case 28: rc = launch_left_vec<28>(j, d_elems, d_x); break;
            default: rc = set_error(RS_ERR_UNSUPPORTED, "unsupported limb count");
        }
    }
    if (rc == RS_OK && cudaMallocAsync((void **)d_left, (size_t)p.rows * pq(j) * 8, current_stream()) != cudaSuccess)
        rc = set_error(RS_ERR_CUDA, "rs_jindo_partial_at: cudaMallocAsync failed");
    if (rc == RS_OK) rc = encode_dispatch(j, *d_left, j->d_ldescs, (size_t)p.rows, nullptr, d_elems, nullptr, j->d_zero_g);
    cudaFreeAsync(d_elems, current_stream());
    return rc;
}

int rs_jindo_partial_at(rs_jindo *j, const rs_opening *o, const uint64_t *x_mont, uint64_t *partial_out) {
    RS_REQUIRE(j && o && x_mont && partial_out, RS_ERR_ARG, "rs_jindo_partial_at: null argument");
    RS_REQUIRE(o->j == j, RS_ERR_ARG, "rs_jindo_partial_at: opening belongs to another prover");
    u64 *d_left = nullptr;
    int rc = encode_left_vec(j, x_mont, &d_left);
    if (rc == RS_OK) rc = partial_impl(j, o, d_left, partial_out);
    if (d_left) cudaFreeAsync(d_left, current_stream());
    return rc;
}

int rs_jindo_left_vec(rs_jindo *j, const uint64_t *x_mont, uint64_t *left_out) {
    RS_REQUIRE(j && x_mont && left_out, RS_ERR_ARG, "rs_jindo_left_vec: null argument");
    u64 *d_left = nullptr;
    int rc = encode_left_vec(j, x_mont, &d_left);
    if (rc == RS_OK) {
        cudaError_t e = cudaMemcpyAsync(left_out, d_left, (size_t)j->p.rows * pq(j) * 8, cudaMemcpyDeviceToHost, current_stream());
        if (e == cudaSuccess) e = cudaStreamSynchronize(current_stream());
        if (e != cudaSuccess) rc = set_error(RS_ERR_CUDA, "rs_jindo_left_vec: %s", cudaGetErrorString(e));
    }
    if (d_left) cudaFreeAsync(d_left, current_stream());
    return rc;
}

static uint32_t challenge_bound(const rs_jindo_params &p) {              // params.go:358-361
    const int sh = 120 / p.exp;
    const u64 cap = sh >= 63 ? ~0ull : (1ull << sh);
    return (uint32_t)((p.base < cap ? p.base : cap) / 2);
}

// n challenges of 16 bytes (host) -> n polynomials over `ring` at *d_out (cudaMallocAsync'ed here)
static int encode_challenges(rs_jindo *j, const uint8_t *bytes, size_t n, const RingDev &ring, u64 **d_out) {
    *d_out = nullptr;
    uint8_t *d_b = nullptr;
    RS_CUDA_CHECK(cudaMallocAsync((void **)&d_b, n * 16, current_stream()));
    int rc = RS_OK;
    if (cudaMemcpyAsync(d_b, bytes, n * 16, cudaMemcpyHostToDevice, current_stream()) != cudaSuccess ||
        cudaMallocAsync((void **)d_out, n * (size_t)ring.n * JD * 8, current_stream()) != cudaSuccess)
        rc = set_error(RS_ERR_CUDA, "encode_challenges: allocation / upload failed");
    if (rc == RS_OK) {
        challenge_kernel<<<dim3((unsigned)n, (unsigned)ring.n), 128, 0, current_stream()>>>(*d_out, d_b, j->p.exp, j->p.slots, challenge_bound(j->p), ring);
        count_launch();
        if (cudaGetLastError() != cudaSuccess) rc = set_error(RS_ERR_CUDA, "challenge_kernel launch failed");
    }
    cudaFreeAsync(d_b, current_stream());
    return rc;
}

static int respond_impl(rs_jindo *j, const rs_opening *o, const u64 *d_ch, uint64_t *enc_out, uint64_t *mlwe_out) {
    const rs_jindo_params &p = j->p;
    const int nm = p.mlwe + p.in_msis;
    u64 *d_e = nullptr, *d_m = nullptr;
    RS_CUDA_CHECK(cudaMallocAsync((void **)&d_e, (size_t)p.rows * pq(j) * 8, current_stream()));
    RS_CUDA_CHECK(cudaMallocAsync((void **)&d_m, (size_t)nm * pq(j) * 8, current_stream()));
    MatvecGeom G;
    memset(&G, 0, sizeof(G));
    // Encode[i] = Enc[cols][i] + sum_j chal[j] (.) Enc[j][i]: "columns" of the product are the rows i
    G.O = 1; G.C = p.rows; G.K = p.cols;
    G.key_os = 0; G.key_ks = 1; G.x_cs = 1; G.x_ks = p.rows; G.out_cs = 1; G.out_os = 0;
    G.has_addend = 1; G.add_cs = 1; G.add_os = 0;
    int rc = launch_matvec(d_e, d_ch, o->d_encode, o->d_encode + (size_t)p.cols * p.rows * pq(j), G, j->rq);
    if (rc == RS_OK) {
        G.C = nm; G.x_ks = nm;
        rc = launch_matvec(d_m, d_ch, o->d_mlwe, o->d_mlwe + (size_t)p.cols * nm * pq(j), G, j->rq);
    }
    if (rc == RS_OK) {
        cudaError_t e = cudaMemcpyAsync(enc_out, d_e, (size_t)p.rows * pq(j) * 8, cudaMemcpyDeviceToHost, current_stream());
        if (e == cudaSuccess) e = cudaMemcpyAsync(mlwe_out, d_m, (size_t)nm * pq(j) * 8, cudaMemcpyDeviceToHost, current_stream());
        if (e == cudaSuccess) e = cudaStreamSynchronize(current_stream());
        if (e != cudaSuccess) rc = set_error(RS_ERR_CUDA, "rs_jindo_respond: %s", cudaGetErrorString(e));
    }
    cudaFreeAsync(d_e, current_stream()); cudaFreeAsync(d_m, current_stream());
    return rc;
}

int rs_jindo_respond(rs_jindo *j, const rs_opening *o, const uint64_t *chals, uint64_t *enc_out, uint64_t *mlwe_out) {
    RS_REQUIRE(j && o && chals && enc_out && mlwe_out, RS_ERR_ARG, "rs_jindo_respond: null argument");
    RS_REQUIRE(o->j == j, RS_ERR_ARG, "rs_jindo_respond: opening belongs to another prover");
    u64 *d_ch = nullptr;
    int rc = upload(&d_ch, (const u64 *)chals, (size_t)j->p.cols * pq(j));
    if (rc == RS_OK) rc = respond_impl(j, o, d_ch, enc_out, mlwe_out);
    cudaFreeAsync(d_ch, current_stream());
    return rc;
}

int rs_jindo_respond_bytes(rs_jindo *j, const rs_opening *o, const uint8_t *chal_bytes, uint64_t *enc_out, uint64_t *mlwe_out) {
    RS_REQUIRE(j && o && chal_bytes && enc_out && mlwe_out, RS_ERR_ARG, "rs_jindo_respond_bytes: null argument");
    RS_REQUIRE(o->j == j, RS_ERR_ARG, "rs_jindo_respond_bytes: opening belongs to another prover");
    u64 *d_ch = nullptr;
    int rc = encode_challenges(j, chal_bytes, (size_t)j->p.cols, j->rq, &d_ch);
    if (rc == RS_OK) rc = respond_impl(j, o, d_ch, enc_out, mlwe_out);
    if (d_ch) cudaFreeAsync(d_ch, current_stream());
    return rc;
}

int rs_jindo_encode_challenges(rs_jindo *j, const uint8_t *chal_bytes, size_t n, int ring_out, uint64_t *polys_out) {
    RS_REQUIRE(j && chal_bytes && polys_out && n >= 1, RS_ERR_ARG, "rs_jindo_encode_challenges: bad argument");
    const RingDev &ring = ring_out ? j->rqo : j->rq;
    u64 *d_ch = nullptr;
    int rc = encode_challenges(j, chal_bytes, n, ring, &d_ch);
    if (rc == RS_OK) {
        cudaError_t e = cudaMemcpyAsync(polys_out, d_ch, n * (size_t)ring.n * JD * 8, cudaMemcpyDeviceToHost, current_stream());
        if (e == cudaSuccess) e = cudaStreamSynchronize(current_stream());
        if (e != cudaSuccess) rc = set_error(RS_ERR_CUDA, "rs_jindo_encode_challenges: %s", cudaGetErrorString(e));
    }
    if (d_ch) cudaFreeAsync(d_ch, current_stream());
    return rc;
}

static int opening_mul_add_impl(rs_jindo *j, rs_opening *acc, const rs_opening *o, const u64 *d_cq, const u64 *d_co) {
    const rs_jindo_params &p = j->p;
    MatvecGeom Gq, Go;
    memset(&Gq, 0, sizeof(Gq)); memset(&Go, 0, sizeof(Go));
    fill_geom_ring(Gq, j->rq); fill_geom_ring(Go, j->rqo);
    const size_t ne = (size_t)(p.cols + 1) * p.rows, nm = (size_t)(p.cols + 1) * (p.mlwe + p.in_msis), ni = (size_t)(p.cols + 1) * p.in_msis;
    muladd_bcast_kernel<<<148 * 8, 256, 0, current_stream()>>>(acc->d_encode, o->d_encode, d_cq, ne, j->rq.n, Gq);
    muladd_bcast_kernel<<<148 * 2, 256, 0, current_stream()>>>(acc->d_mlwe, o->d_mlwe, d_cq, nm, j->rq.n, Gq);
    muladd_bcast_kernel<<<148 * 2, 256, 0, current_stream()>>>(acc->d_incommit, o->d_incommit, d_co, ni, j->rqo.n, Go);
    count_launch(3);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_error(RS_ERR_CUDA, "rs_opening_mul_add: %s", cudaGetErrorString(e));
    return RS_OK;
}

int rs_opening_mul_add_bytes(rs_jindo *j, rs_opening *acc, const rs_opening *o, const uint8_t *chal16) {
    RS_REQUIRE(j && acc && o && chal16, RS_ERR_ARG, "rs_opening_mul_add_bytes: null argument");
    RS_REQUIRE(o->j == j && acc->j == j, RS_ERR_ARG, "rs_opening_mul_add_bytes: opening belongs to another prover");
    u64 *d_cq = nullptr, *d_co = nullptr;
    int rc = encode_challenges(j, chal16, 1, j->rq, &d_cq);
    if (rc == RS_OK) rc = encode_challenges(j, chal16, 1, j->rqo, &d_co);
    if (rc == RS_OK) rc = opening_mul_add_impl(j, acc, o, d_cq, d_co);
    if (d_cq) cudaFreeAsync(d_cq, current_stream());
    if (d_co) cudaFreeAsync(d_co, current_stream());
    return rc;
}

int rs_opening_combine_bytes(rs_jindo *j, const rs_opening *const *opens, size_t n, const uint8_t *chal_bytes, rs_opening **out) {
    RS_REQUIRE(j && opens && chal_bytes && out && n >= 1, RS_ERR_ARG, "rs_opening_combine_bytes: bad argument");
    for (size_t k = 0; k < n; k++) RS_REQUIRE(opens[k] && opens[k]->j == j, RS_ERR_ARG, "rs_opening_combine_bytes: opening belongs to another prover");
    const rs_jindo_params &p = j->p;
    rs_opening *acc = nullptr;
    int rc = opening_alloc(j, &acc, false);
    if (rc) return rc;
    u64 *d_cq = nullptr, *d_co = nullptr;
    const u64 **d_ptrs = nullptr;
    std::vector<const u64 *> h(3 * n);
    for (size_t k = 0; k < n; k++) { h[k] = opens[k]->d_encode; h[n + k] = opens[k]->d_mlwe; h[2 * n + k] = opens[k]->d_incommit; }
    rc = encode_challenges(j, chal_bytes, n, j->rq, &d_cq);
    if (rc == RS_OK) rc = encode_challenges(j, chal_bytes, n, j->rqo, &d_co);
    if (rc == RS_OK && (cudaMallocAsync((void **)&d_ptrs, 3 * n * sizeof(u64 *), current_stream()) != cudaSuccess ||
                        cudaMemcpyAsync(d_ptrs, h.data(), 3 * n * sizeof(u64 *), cudaMemcpyHostToDevice, current_stream()) != cudaSuccess))
        rc = set_error(RS_ERR_CUDA, "rs_opening_combine_bytes: upload failed");      // pageable source: staged before the call returns
    if (rc == RS_OK) {
        MatvecGeom Gq, Go;
        memset(&Gq, 0, sizeof(Gq)); memset(&Go, 0, sizeof(Go));
        fill_geom_ring(Gq, j->rq); fill_geom_ring(Go, j->rqo);
        const size_t ne = (size_t)(p.cols + 1) * p.rows, nm = (size_t)(p.cols + 1) * (p.mlwe + p.in_msis), ni = (size_t)(p.cols + 1) * p.in_msis;
        combine_kernel<<<148 * 8, 256, 0, current_stream()>>>(acc->d_encode, d_ptrs, d_cq, (int)n, ne, j->rq.n, Gq);
        combine_kernel<<<148 * 2, 256, 0, current_stream()>>>(acc->d_mlwe, d_ptrs + n, d_cq, (int)n, nm, j->rq.n, Gq);
        combine_kernel<<<148 * 2, 256, 0, current_stream()>>>(acc->d_incommit, d_ptrs + 2 * n, d_co, (int)n, ni, j->rqo.n, Go);
        count_launch(3);
        if (cudaGetLastError() != cudaSuccess) rc = set_error(RS_ERR_CUDA, "rs_opening_combine_bytes: launch failed");
    }
    if (d_cq) cudaFreeAsync(d_cq, current_stream());
    if (d_co) cudaFreeAsync(d_co, current_stream());
    if (d_ptrs) cudaFreeAsync(d_ptrs, current_stream());
    if (rc) { rs_opening_free(acc); return rc; }
    *out = acc;
    return RS_OK;
}

int rs_opening_mul_add(rs_jindo *j, rs_opening *acc, const rs_opening *o, const uint64_t *chal_q, const uint64_t *chal_qout) {
    RS_REQUIRE(j && acc && o && chal_q && chal_qout, RS_ERR_ARG, "rs_opening_mul_add: null argument");
    RS_REQUIRE(o->j == j && acc->j == j, RS_ERR_ARG, "rs_opening_mul_add: opening belongs to another prover");
    u64 *d_cq = nullptr, *d_co = nullptr;
    int rc = upload(&d_cq, (const u64 *)chal_q, pq(j));
    if (rc == RS_OK) rc = upload(&d_co, (const u64 *)chal_qout, pqo(j));
    if (rc == RS_OK) rc = opening_mul_add_impl(j, acc, o, d_cq, d_co);
    cudaFreeAsync(d_cq, current_stream()); cudaFreeAsync(d_co, current_stream());
    return rc;
}

#pragma GCC visibility pop
}  // extern "C"
