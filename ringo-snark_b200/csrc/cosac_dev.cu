// cosac_dev.cu -- the encoder's COSAC rows sampled in HBM (the last sampler that lived on the host).
//
// A Commit's COSAC polynomials (row 0 of every data column, every row of the mask column: jindo/prover.go:93-128,
// encoder.go:149-171) each need 256 samples of COSACSampler.Sample(center_i, stdDev) (math/csprng/gaussian_cosac.go:23-57), and
// every sample runs data-dependent rejection loops over two uniform streams.  One sampler object is therefore sequential; the
// reference's own answer to parallelism is one sampler per worker (Prover.SafeCopy gives every goroutine fresh samplers,
// prover.go:327-338).  Here a worker is one coefficient: worker w owns a fresh COSACSampler whose two UniformSamplers are
// AES-256-CTR streams (uniform.go:26-70) with the stream's two keys and the counters IV + (w << 16) -- i.e. exactly the object
// csprng.COSACSampler{baseSampler: {key, IV_b + (w << 16)}, roundedSampler: {baseSampler: {key', IV_r + (w << 16)}}} before its
// first refill.  The test builds that object on the host (rs_cosac_create) for every worker and compares sample for sample.
// A worker that would need a second 8 KiB refill (1024 draws; a sample takes ~5) raises an overflow flag instead.
//
//   phase 1  fromMont + base-b digits of the row's field elements            encoder.go:120-146
//   phase 2  float64 centres, unfused, in the reference's order              encoder.go:152-164
//   phase 3  Sample(-fp[i], stdDev): ziggurat normFloat (gaussian_rounded.go:77-117) + the COSAC accept loop
//
// Every floating-point step is written with the _rn intrinsics: the host code is built with -ffp-contract=off because Go on
// amd64 never fuses a*b+c, and nvcc would otherwise contract.  exp / log are CUDA's (<= 1 ulp, like glibc's): a comparison
// can differ from the host's only when a 53-bit uniform lands within an ulp of the threshold.
#include <math.h>
#include <string.h>

#include "common.cuh"
#include "../../include/ringo_jindo.h"

typedef unsigned long long u64;
typedef long long i64;

extern "C" void rs_zig_tables(uint64_t *kn, double *wn, double *fn);   // host_sampler.cpp: the package-init tables of gaussian_rounded.go

namespace rs {

void aes256_expand_key(const uint8_t key[32], uint32_t rk[60]);       // aes_ctr.cu
const uint8_t *aes_sbox_host();

struct CosacKeys { uint32_t rkb[60], rkr[60]; };
struct CosacParams {
    int slots, exp, cols, n_polys;
    u64 base;
    double delta_inv[RS_JINDO_MAX_EXP];
    double std_first, std_mask0, std_mask;
    u64 ivb_hi, ivb_lo, ivr_hi, ivr_lo;       // IVs of the two streams as 128-bit big-endian integers
    u64 worker0;                              // workers handed out before this launch
};
struct ZigTab { u64 kn[128]; double wn[128], fn[128]; };

struct CosacStream {
    CosacKeys K;
    u64 ivb_hi, ivb_lo, ivr_hi, ivr_lo;
    u64 workers;
    double std_first, std_mask0, std_mask;
    ZigTab *d_zig;
    uint8_t *d_sbox;
    int *d_overflow;
};

// ---- AES-256 block on two round tables in 256-byte rows (the lookup scheme of aes_keystream4_scan_kernel, aes_ctr.cu) -----------
// Shared layout: byte offset of (table t in {0, 1}, value b, lane l) = b * 256 + t * 128 + l * 4  (64 KiB; every lane reads its
// own bank).  A lookup address is ONE byte-permute: byte 0 = 4 * lane, byte 1 = the state byte; Te2 = ROT16(Te0) and
// Te3 = ROT16(Te1), and the rotation is linear over XOR, so a column costs 4 PRMT + 4 LDS + 1 PRMT + 3 LOP3.
// T points at the lane's slot of row 0 (Te + 4 * lane bytes); `Sb` is unused (kept in the signatures): the last round
// takes S[b] out of bytes 1 / 2 of Te0[b].
#define RS_TOFF(s, SEL) __byte_perm((s), lane4, SEL)
#define RS_T0(off) (*reinterpret_cast<const uint32_t *>(reinterpret_cast<const char *>(T0) + (off)))
#define RS_T1(off) (*reinterpret_cast<const uint32_t *>(reinterpret_cast<const char *>(T0) + (off) + 128))
#define RS_COL(a, b, c, d, rkw) (RS_T0(RS_TOFF(a, 0x5534)) ^ RS_T1(RS_TOFF(b, 0x5524)) ^ \
                                 __byte_perm(RS_T0(RS_TOFF(c, 0x5514)) ^ RS_T1(RS_TOFF(d, 0x5504)), 0, 0x1032) ^ (rkw))
#define RS_LAST(a, b, c, d, rkw) (__byte_perm(__byte_perm(RS_T0(RS_TOFF(a, 0x5534)), RS_T0(RS_TOFF(b, 0x5524)), 0x2600), \
                                              __byte_perm(RS_T0(RS_TOFF(c, 0x5514)), RS_T0(RS_TOFF(d, 0x5504)), 0x0015), 0x3254) ^ (rkw))
__device__ __forceinline__ void aes_block(const uint32_t *rk, u64 hi, u64 lo, const uint32_t *T, const uint32_t *Sb, u64 &w0, u64 &w1) {
    (void)Sb;
    const uint32_t lane4 = (threadIdx.x & 31) * 4;
    const uint32_t *T0 = T - (threadIdx.x & 31);                          // row 0, lane 0 (the permute puts 4 * lane back in)
    uint32_t s0 = (uint32_t)(hi >> 32) ^ rk[0], s1 = (uint32_t)hi ^ rk[1], s2 = (uint32_t)(lo >> 32) ^ rk[2], s3 = (uint32_t)lo ^ rk[3];
#pragma unroll 1
    for (int r = 1; r < 14; r++) {
        const uint32_t t0 = RS_COL(s0, s1, s2, s3, rk[4 * r]);
        const uint32_t t1 = RS_COL(s1, s2, s3, s0, rk[4 * r + 1]);
        const uint32_t t2 = RS_COL(s2, s3, s0, s1, rk[4 * r + 2]);
        const uint32_t t3 = RS_COL(s3, s0, s1, s2, rk[4 * r + 3]);
        s0 = t0; s1 = t1; s2 = t2; s3 = t3;
    }
    const uint32_t o0 = RS_LAST(s0, s1, s2, s3, rk[56]);
    const uint32_t o1 = RS_LAST(s1, s2, s3, s0, rk[57]);
    const uint32_t o2 = RS_LAST(s2, s3, s0, s1, rk[58]);
    const uint32_t o3 = RS_LAST(s3, s0, s1, s2, rk[59]);
    // the block's 16 bytes are the big-endian words o0..o3; Sample() reads them 8 at a time, little endian (uniform.go:72-77)
    w0 = (u64)__byte_perm(o0, 0, 0x0123) | ((u64)__byte_perm(o1, 0, 0x0123) << 32);
    w1 = (u64)__byte_perm(o2, 0, 0x0123) | ((u64)__byte_perm(o3, 0, 0x0123) << 32);
}
#undef RS_TOFF
#undef RS_T0
#undef RS_T1
#undef RS_COL
#undef RS_LAST

// a fresh csprng.UniformSampler: the first 8 KiB of its buffer are the plain keystream
struct DevUniform {
    const uint32_t *rk;
    u64 hi, lo;          // counter of the next block
    u64 spare;
    uint32_t blocks;
    bool has_spare;
};
__device__ __forceinline__ void uni_init(DevUniform &u, const uint32_t *rk, u64 iv_hi, u64 iv_lo, u64 worker) {
    // counter = IV + (worker << 16), 128-bit
    const u64 add_lo = worker << 16, add_hi = worker >> 48;
    u.lo = iv_lo + add_lo;
    u.hi = iv_hi + add_hi + (u.lo < add_lo);
    u.rk = rk; u.blocks = 0; u.has_spare = false; u.spare = 0;
}
__device__ __forceinline__ u64 uni_sample(DevUniform &u, const uint32_t *T, const uint32_t *Sb, int *overflow) {
    if (u.has_spare) { u.has_spare = false; return u.spare; }
    if (u.blocks >= 512) *overflow = 1;                       // a second refill would XOR into the first buffer (uniform.go:64-70)
    u64 w0, w1;
    aes_block(u.rk, u.hi, u.lo, T, Sb, w0, w1);
    u.lo += 1; u.hi += (u.lo == 0);
    u.blocks++;
    u.spare = w1; u.has_spare = true;
    return w0;
}
__device__ __forceinline__ double uni_sample_float(DevUniform &u, const uint32_t *T, const uint32_t *Sb, int *overflow) {   // uniform.go:96-100
    const u64 r = uni_sample(u, T, Sb, overflow) & ((1ull << 52) - 1);
    const double rf = __longlong_as_double((long long)(r | ((u64)(1023 + 52) << 52)));
    return __dsub_rn(__ddiv_rn(rf, 4503599627370496.0), 1.0);
}

// RoundedGaussianSampler.normFloat (gaussian_rounded.go:77-117)
__device__ double norm_float(DevUniform &u, const ZigTab *__restrict__ Z, const uint32_t *T, const uint32_t *Sb, int *overflow) {
    const double RN = 3.442619855899, INV_RN = 1.0 / 3.442619855899;
    for (;;) {
        const u64 r = uni_sample(u, T, Sb, overflow);
        const u64 b = r >> 63;
        const u64 i = r & 127;
        const u64 j = (r >> 7) & ((1ull << 52) - 1);
        const double x = __dmul_rn(__ll2double_rn((long long)((j ^ (0 - b)) + b)), __ldg(&Z->wn[i]));
        if (j < __ldg(&Z->kn[i])) return x;
        if (i == 0) {
            double a, v;
            for (;;) {
                a = __dmul_rn(-log(uni_sample_float(u, T, Sb, overflow)), INV_RN);
                v = -log(uni_sample_float(u, T, Sb, overflow));
                if (__dadd_rn(v, v) >= __dmul_rn(a, a)) break;
                if (*overflow) return 0.0;
            }
            a = __dadd_rn(a, RN);
            return b == 1 ? -a : a;
        }
        const double f0 = __ldg(&Z->fn[i - 1]), f1 = __ldg(&Z->fn[i]);
        if (__dmul_rn(uni_sample_float(u, T, Sb, overflow), __dsub_rn(f0, f1)) < __dsub_rn(exp(__dmul_rn(__dmul_rn(-0.5, x), x)), f1)) return x;
        if (*overflow) return 0.0;
    }
}

// COSACSampler.Sample (gaussian_cosac.go:48-57) and sampleRound (:23-46), with the host restatement's exp-free shortcuts
__device__ i64 cosac_sample(DevUniform &base, DevUniform &rnd, double center, double std_dev, const ZigTab *__restrict__ Z,
                            const uint32_t *T, const uint32_t *Sb, int *overflow) {
    const double c_int = round(center);
    const double c_frac = __dsub_rn(c_int, center);
    const double r0 = uni_sample_float(base, T, Sb, overflow);
    const double norm = __dmul_rn(sqrt(6.283185307179586), std_dev);
    const double two_s2 = __dmul_rn(__dmul_rn(2.0, std_dev), std_dev);
    if (r0 < __ddiv_rn(1.0, norm) && r0 < __ddiv_rn(exp(__ddiv_rn(-__dmul_rn(c_frac, c_frac), two_s2)), norm)) return (i64)c_int;
    for (;;) {
        const double y = __dmul_rn(std_dev, norm_float(rnd, Z, T, Sb, overflow));
        const u64 b = uni_sample(base, T, Sb, overflow) & 1;
        double y_round;
        bool cmp;
        if (b == 0) { y_round = __dsub_rn(round(y), 1.0); cmp = y_round <= 0.5; }
        else { y_round = __dadd_rn(round(y), 1.0); cmp = y_round >= -0.5; }
        if (cmp) {
            const double r = uni_sample_float(base, T, Sb, overflow);
            const double s = __dadd_rn(y_round, c_frac);
            const double z = __ddiv_rn(__dsub_rn(__dmul_rn(s, s), __dmul_rn(y, y)), two_s2);
            if (z <= 0 || r < __dsub_rn(__dsub_rn(1.0, z), 1e-12)) return (i64)y_round + (i64)c_int;
            if (r < exp(-z)) return (i64)y_round + (i64)c_int;
        }
        if (*overflow) return 0;
    }
}

template <int NL, int QS>
__global__ void __launch_bounds__(256) cosac_kernel(i64 *__restrict__ out, const uint32_t *__restrict__ first_row, const uint32_t *__restrict__ mask_rows,
                                                    const __grid_constant__ CosacKeys K, const __grid_constant__ CosacParams P,
                                                    const __grid_constant__ FieldParams<NL> F, const ZigTab *__restrict__ Z,
                                                    const uint8_t *__restrict__ sbox, int *overflow) {
    extern __shared__ uint32_t Te[];                 // [256 values][Te0 x 32 lanes | Te1 x 32 lanes] (see aes_block)
    __shared__ uint32_t Sb[1];                       // (unused: the last round reads S[b] out of Te0)
    __shared__ double dig[256];
    __shared__ int s_over;
    const int c = threadIdx.x;
    {
        const uint32_t s = sbox[c], s2 = ((s << 1) ^ (((s >> 7) & 1) * 0x1b)) & 0xff, s3 = s2 ^ s;
        const uint32_t t = (s2 << 24) | (s << 16) | (s << 8) | s3, t1 = __byte_perm(t, 0, 0x0321);
#pragma unroll
        for (int k = 0; k < 32; k++) { Te[c * 64 + k] = t; Te[c * 64 + 32 + k] = t1; }
        if (c == 0) s_over = 0;
    }
    for (int t = blockIdx.x; t < P.n_polys; t += gridDim.x) {
        dig[c] = 0.0;
        __syncthreads();
        const uint32_t *elems = t < P.cols ? first_row + (size_t)t * P.slots * NL : mask_rows + (size_t)(t - P.cols) * P.slots * NL;
        if (c < P.slots) {                           // fromMont + base-b digits (encoder.go:120-146): exp - 1 divisions, the quotient's low word last
            uint32_t x[NL], one[NL], v[NL];
#pragma unroll
            for (int l = 0; l < NL; l++) x[l] = elems[(size_t)c * NL + l];
#pragma unroll
            for (int l = 0; l < NL; l++) one[l] = (l == 0);
            mont_mul<NL, QS>(v, x, one, F);
            const uint32_t base = (uint32_t)P.base;
            for (int j = 0; j < P.exp - 1; j++) {
                u64 rem = 0;
#pragma unroll
                for (int l = NL - 1; l >= 0; l--) {
                    const u64 cur = (rem << 32) | v[l];
                    v[l] = (uint32_t)(cur / base);
                    rem = cur - (u64)v[l] * base;
                }
                dig[j * P.slots + c] = __ull2double_rn(rem);
            }
            dig[(P.exp - 1) * P.slots + c] = __ull2double_rn(((u64)v[1] << 32) | v[0]);
        }
        __syncthreads();
        double fp = 0.0;                             // encoder.go:152-164, one coefficient per thread, terms in the order of i
        for (int i = 0; i < P.exp; i++) {
            const double di = P.delta_inv[i];
            if (di == 0) continue;
            const int dd = 256 - (i + 1) * P.slots;
            if (c >= dd) fp = __dadd_rn(fp, __dmul_rn(di, dig[c - dd]));
            else fp = __dsub_rn(fp, __dmul_rn(di, dig[256 - dd + c]));
        }
        const double std_dev = t < P.cols ? P.std_first : (t == P.cols ? P.std_mask0 : P.std_mask);
        const u64 w = P.worker0 + (u64)t * 256 + c;
        DevUniform ub, ur;
        uni_init(ub, K.rkb, P.ivb_hi, P.ivb_lo, w);
        uni_init(ur, K.rkr, P.ivr_hi, P.ivr_lo, w);
        const uint32_t *T = Te + (c & 31);
        out[(size_t)t * 256 + c] = cosac_sample(ub, ur, -fp, std_dev, Z, T, Sb, &s_over);
        __syncthreads();
    }
    if (c == 0 && s_over) *overflow = 1;
}

static void iv_words(const uint8_t iv[16], u64 &hi, u64 &lo) {
    hi = lo = 0;
    for (int i = 0; i < 8; i++) { hi = (hi << 8) | iv[i]; lo = (lo << 8) | iv[8 + i]; }
}

int cosac_stream_init(CosacStream **out, const uint8_t kb[48], const uint8_t kr[48], double std_first, double std_mask0, double std_mask) {
    CosacStream *s = new CosacStream();
    memset(s, 0, sizeof(*s));
    aes256_expand_key(kb, s->K.rkb);
    aes256_expand_key(kr, s->K.rkr);
    iv_words(kb + 32, s->ivb_hi, s->ivb_lo);
    iv_words(kr + 32, s->ivr_hi, s->ivr_lo);
    s->std_first = std_first; s->std_mask0 = std_mask0; s->std_mask = std_mask;
    ZigTab z;
    rs_zig_tables((uint64_t *)z.kn, z.wn, z.fn);
    cudaError_t e = cudaMalloc((void **)&s->d_zig, sizeof(ZigTab));
    if (e == cudaSuccess) e = cudaMemcpy(s->d_zig, &z, sizeof(ZigTab), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMalloc((void **)&s->d_sbox, 256);
    if (e == cudaSuccess) e = cudaMemcpy(s->d_sbox, aes_sbox_host(), 256, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMalloc((void **)&s->d_overflow, 4);
    if (e == cudaSuccess) e = cudaMemset(s->d_overflow, 0, 4);
    if (e != cudaSuccess) {
        cudaFree(s->d_zig); cudaFree(s->d_sbox); cudaFree(s->d_overflow);
        delete s;
        return set_error(RS_ERR_CUDA, "cosac stream: %s", cudaGetErrorString(e));
    }
    *out = s;
    return RS_OK;
}

void cosac_stream_free(CosacStream *s) {
    if (!s) return;
    cudaFree(s->d_zig); cudaFree(s->d_sbox); cudaFree(s->d_overflow);
    delete s;
}

u64 cosac_stream_workers(const CosacStream *s) { return s->workers; }
const int *cosac_stream_overflow_flag(const CosacStream *s) { return s->d_overflow; }

// samples the n_polys COSAC polynomials of one Commit into d_out (n_polys x 256): polys [0, cols) take their elements from
// d_first_row (slots each), the others from d_mask_rows; advances the stream by n_polys * 256 workers
int cosac_generate(CosacStream *s, const rs_field *f, const rs_jindo_params &p, const uint32_t *d_first_row, const uint32_t *d_mask_rows,
                   size_t n_polys, i64 *d_out) {
    if (n_polys == 0) return RS_OK;
    RS_REQUIRE(p.d == 256 && p.exp * p.slots == 256, RS_ERR_UNSUPPORTED, "device COSAC: ring degree must be 256 = exp * slots");
    RS_REQUIRE(p.base < (1ull << 32), RS_ERR_UNSUPPORTED, "device COSAC: encoder base must fit 32 bits");
    CosacParams P;
    memset(&P, 0, sizeof(P));
    P.slots = p.slots; P.exp = p.exp; P.cols = p.cols; P.n_polys = (int)n_polys; P.base = p.base;
    for (int i = 0; i < p.exp; i++) P.delta_inv[i] = p.delta_inv[i];
    P.std_first = s->std_first; P.std_mask0 = s->std_mask0; P.std_mask = s->std_mask;
    P.ivb_hi = s->ivb_hi; P.ivb_lo = s->ivb_lo; P.ivr_hi = s->ivr_hi; P.ivr_lo = s->ivr_lo;
    P.worker0 = s->workers;
    const int smem = 256 * 64 * 4;                   // 64 KiB of round tables per CTA: three 256-thread CTAs per SM
    const unsigned grid = (unsigned)(n_polys < 148 * 3 ? n_polys : 148 * 3);
#define CALL(NL, QS)                                                                                                      \
    {                                                                                                                     \
        FieldParams<NL> F = make_params<NL>(f);                                                                           \
        RS_CUDA_CHECK(ensure_smem_attr((const void *)cosac_kernel<NL, QS>, smem));                                        \
        cosac_kernel<NL, QS><<<grid, 256, smem, current_stream()>>>(d_out, d_first_row, d_mask_rows, s->K, P, F, s->d_zig, s->d_sbox, s->d_overflow); \
    }
    switch (f->NL) {
        case 4: if (f->qs >= 1) { CALL(4, 1); } else { CALL(4, 0); } break;
        case 8: if (f->qs == 2) { CALL(8, 2); } else if (f->qs == 1) { CALL(8, 1); } else { CALL(8, 0); } break;
        case 14: if (f->qs == 2) { CALL(14, 2); } else { CALL(14, 0); } break;
        case 28: if (f->qs == 2) { CALL(28, 2); } else { CALL(28, 0); } break;
        default: return set_error(RS_ERR_UNSUPPORTED, "unsupported limb count %d", f->L64);
    }
#undef CALL
    count_launch();
    RS_CUDA_CHECK(cudaGetLastError());
    s->workers += (u64)n_polys * 256;
    return RS_OK;
}

}  // namespace rs
