// host_sampler.cpp -- the host half of a Commit's randomness (include/ringo_sampler.h): bit-exact restatements of
// math/csprng's UniformSampler / RoundedGaussianSampler / COSACSampler, the MustSetRandom draws, and a threaded
// builder that turns "vector + samplers" into the tape entries rs_jindo_commit_host consumes.
// Plain host C++ (g++, -ffp-contract=off: Go on amd64 never fuses a*b+c); no CUDA here except the pinned staging
// buffers, which come from rs_pinned_alloc.
#include <math.h>
#include <stdint.h>
#include <string.h>
#include <sys/random.h>

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#if defined(__x86_64__)
#include <immintrin.h>
#endif

#include <stdlib.h>

#include "../../include/ringo_sampler.h"
#include "field.cuh"

#define RS_MAX_NL_HOST 28

namespace {

// ---------------------------------------------------------------------------------------------------------------
// AES-256 (FIPS-197) in CTR mode, Go's cipher.NewCTR convention: 128-bit big-endian counter starting at the IV
// ---------------------------------------------------------------------------------------------------------------
const uint8_t SBOX[256] = {
    0x63,0x7c,0x77,0x7b,0xf2,0x6b,0x6f,0xc5,0x30,0x01,0x67,0x2b,0xfe,0xd7,0xab,0x76,0xca,0x82,0xc9,0x7d,0xfa,0x59,0x47,0xf0,0xad,0xd4,0xa2,0xaf,0x9c,0xa4,0x72,0xc0,
    0xb7,0xfd,0x93,0x26,0x36,0x3f,0xf7,0xcc,0x34,0xa5,0xe5,0xf1,0x71,0xd8,0x31,0x15,0x04,0xc7,0x23,0xc3,0x18,0x96,0x05,0x9a,0x07,0x12,0x80,0xe2,0xeb,0x27,0xb2,0x75,
    0x09,0x83,0x2c,0x1a,0x1b,0x6e,0x5a,0xa0,0x52,0x3b,0xd6,0xb3,0x29,0xe3,0x2f,0x84,0x53,0xd1,0x00,0xed,0x20,0xfc,0xb1,0x5b,0x6a,0xcb,0xbe,0x39,0x4a,0x4c,0x58,0xcf,
    0xd0,0xef,0xaa,0xfb,0x43,0x4d,0x33,0x85,0x45,0xf9,0x02,0x7f,0x50,0x3c,0x9f,0xa8,0x51,0xa3,0x40,0x8f,0x92,0x9d,0x38,0xf5,0xbc,0xb6,0xda,0x21,0x10,0xff,0xf3,0xd2,
    0xcd,0x0c,0x13,0xec,0x5f,0x97,0x44,0x17,0xc4,0xa7,0x7e,0x3d,0x64,0x5d,0x19,0x73,0x60,0x81,0x4f,0xdc,0x22,0x2a,0x90,0x88,0x46,0xee,0xb8,0x14,0xde,0x5e,0x0b,0xdb,
    0xe0,0x32,0x3a,0x0a,0x49,0x06,0x24,0x5c,0xc2,0xd3,0xac,0x62,0x91,0x95,0xe4,0x79,0xe7,0xc8,0x37,0x6d,0x8d,0xd5,0x4e,0xa9,0x6c,0x56,0xf4,0xea,0x65,0x7a,0xae,0x08,
    0xba,0x78,0x25,0x2e,0x1c,0xa6,0xb4,0xc6,0xe8,0xdd,0x74,0x1f,0x4b,0xbd,0x8b,0x8a,0x70,0x3e,0xb5,0x66,0x48,0x03,0xf6,0x0e,0x61,0x35,0x57,0xb9,0x86,0xc1,0x1d,0x9e,
    0xe1,0xf8,0x98,0x11,0x69,0xd9,0x8e,0x94,0x9b,0x1e,0x87,0xe9,0xce,0x55,0x28,0xdf,0x8c,0xa1,0x89,0x0d,0xbf,0xe6,0x42,0x68,0x41,0x99,0x2d,0x0f,0xb0,0x54,0xbb,0x16};

inline uint8_t xt(uint8_t x) { return (uint8_t)((x << 1) ^ ((x >> 7) * 0x1b)); }

struct Aes256 {
    alignas(16) uint8_t rk[15][16];     // round keys as byte strings
    bool ni;
    void init(const uint8_t key[32]) {
        uint32_t w[60];
        for (int i = 0; i < 8; i++) w[i] = (uint32_t)key[4 * i] << 24 | (uint32_t)key[4 * i + 1] << 16 | (uint32_t)key[4 * i + 2] << 8 | key[4 * i + 3];
        uint32_t rcon = 1;
        for (int i = 8; i < 60; i++) {
            uint32_t t = w[i - 1];
            if (i % 8 == 0) {
                t = (t << 8) | (t >> 24);
                t = (uint32_t)SBOX[t >> 24] << 24 | (uint32_t)SBOX[(t >> 16) & 0xff] << 16 | (uint32_t)SBOX[(t >> 8) & 0xff] << 8 | SBOX[t & 0xff];
                t ^= rcon << 24;
                rcon = ((rcon << 1) ^ ((rcon >> 7) * 0x11b)) & 0xff;
            } else if (i % 8 == 4) {
                t = (uint32_t)SBOX[t >> 24] << 24 | (uint32_t)SBOX[(t >> 16) & 0xff] << 16 | (uint32_t)SBOX[(t >> 8) & 0xff] << 8 | SBOX[t & 0xff];
            }
            w[i] = w[i - 8] ^ t;
        }
        for (int r = 0; r < 15; r++)
            for (int c = 0; c < 4; c++) {
                rk[r][4 * c] = (uint8_t)(w[4 * r + c] >> 24); rk[r][4 * c + 1] = (uint8_t)(w[4 * r + c] >> 16);
                rk[r][4 * c + 2] = (uint8_t)(w[4 * r + c] >> 8); rk[r][4 * c + 3] = (uint8_t)w[4 * r + c];
            }
#if defined(__x86_64__)
        ni = __builtin_cpu_supports("aes");
#else
        ni = false;
#endif
    }
    void encrypt_portable(const uint8_t in[16], uint8_t out[16]) const {
        uint8_t s[16], t[16];
        for (int i = 0; i < 16; i++) s[i] = in[i] ^ rk[0][i];
        for (int r = 1; r <= 14; r++) {
            // SubBytes + ShiftRows (state is column-major: byte index = 4*col + row)
            for (int c = 0; c < 4; c++)
                for (int row = 0; row < 4; row++) t[4 * c + row] = SBOX[s[4 * ((c + row) & 3) + row]];
            if (r < 14) {
                for (int c = 0; c < 4; c++) {
                    const uint8_t a0 = t[4 * c], a1 = t[4 * c + 1], a2 = t[4 * c + 2], a3 = t[4 * c + 3];
                    s[4 * c] = (uint8_t)(xt(a0) ^ (xt(a1) ^ a1) ^ a2 ^ a3);
                    s[4 * c + 1] = (uint8_t)(a0 ^ xt(a1) ^ (xt(a2) ^ a2) ^ a3);
                    s[4 * c + 2] = (uint8_t)(a0 ^ a1 ^ xt(a2) ^ (xt(a3) ^ a3));
                    s[4 * c + 3] = (uint8_t)((xt(a0) ^ a0) ^ a1 ^ a2 ^ xt(a3));
                }
            } else {
                memcpy(s, t, 16);
            }
            for (int i = 0; i < 16; i++) s[i] ^= rk[r][i];
        }
        memcpy(out, s, 16);
    }
#if defined(__x86_64__)
    // 8 blocks at a time keeps the AES units' pipelines full (AESENC: latency 4, throughput 1-2 per cycle)
    __attribute__((target("aes,sse4.1"))) void encrypt8_ni(const uint8_t *in /* 8 blocks */, uint8_t *out) const {
        const __m128i *k = (const __m128i *)rk;
        __m128i b0 = _mm_xor_si128(_mm_loadu_si128((const __m128i *)in + 0), k[0]), b1 = _mm_xor_si128(_mm_loadu_si128((const __m128i *)in + 1), k[0]);
        __m128i b2 = _mm_xor_si128(_mm_loadu_si128((const __m128i *)in + 2), k[0]), b3 = _mm_xor_si128(_mm_loadu_si128((const __m128i *)in + 3), k[0]);
        __m128i b4 = _mm_xor_si128(_mm_loadu_si128((const __m128i *)in + 4), k[0]), b5 = _mm_xor_si128(_mm_loadu_si128((const __m128i *)in + 5), k[0]);
        __m128i b6 = _mm_xor_si128(_mm_loadu_si128((const __m128i *)in + 6), k[0]), b7 = _mm_xor_si128(_mm_loadu_si128((const __m128i *)in + 7), k[0]);
#define RS_AES_ROUND(r)                                                                                               \
        b0 = _mm_aesenc_si128(b0, k[r]); b1 = _mm_aesenc_si128(b1, k[r]); b2 = _mm_aesenc_si128(b2, k[r]); b3 = _mm_aesenc_si128(b3, k[r]); \
        b4 = _mm_aesenc_si128(b4, k[r]); b5 = _mm_aesenc_si128(b5, k[r]); b6 = _mm_aesenc_si128(b6, k[r]); b7 = _mm_aesenc_si128(b7, k[r]);
        RS_AES_ROUND(1) RS_AES_ROUND(2) RS_AES_ROUND(3) RS_AES_ROUND(4) RS_AES_ROUND(5) RS_AES_ROUND(6) RS_AES_ROUND(7)
        RS_AES_ROUND(8) RS_AES_ROUND(9) RS_AES_ROUND(10) RS_AES_ROUND(11) RS_AES_ROUND(12) RS_AES_ROUND(13)
#undef RS_AES_ROUND
        __m128i *o = (__m128i *)out;
        _mm_storeu_si128(o + 0, _mm_aesenclast_si128(b0, k[14])); _mm_storeu_si128(o + 1, _mm_aesenclast_si128(b1, k[14]));
        _mm_storeu_si128(o + 2, _mm_aesenclast_si128(b2, k[14])); _mm_storeu_si128(o + 3, _mm_aesenclast_si128(b3, k[14]));
        _mm_storeu_si128(o + 4, _mm_aesenclast_si128(b4, k[14])); _mm_storeu_si128(o + 5, _mm_aesenclast_si128(b5, k[14]));
        _mm_storeu_si128(o + 6, _mm_aesenclast_si128(b6, k[14])); _mm_storeu_si128(o + 7, _mm_aesenclast_si128(b7, k[14]));
    }
#endif
#if defined(__x86_64__)
    // p[0 .. 16 n) ^= keystream for the counters ctr, ctr+1, ... (128-bit big endian), n a multiple of 8; returns through
    // hi/lo the counter after the last block
    __attribute__((target("aes,sse4.1,ssse3"))) void ctr_xor8_ni(uint8_t *p, size_t n, uint64_t &hi, uint64_t &lo) const {
        const __m128i *k = (const __m128i *)rk;
        const __m128i bswap = _mm_set_epi8(8, 9, 10, 11, 12, 13, 14, 15, 0, 1, 2, 3, 4, 5, 6, 7);   // two u64 -> big-endian bytes (hi first)
        for (size_t i = 0; i < n; i += 8, p += 128) {
            __m128i b[8];
            for (int j = 0; j < 8; j++) {
                const uint64_t l = lo + (uint64_t)j, h = hi + (l < lo);
                b[j] = _mm_xor_si128(_mm_shuffle_epi8(_mm_set_epi64x((long long)l, (long long)h), bswap), k[0]);
            }
            { const uint64_t l = lo + 8; hi += (l < lo); lo = l; }
            for (int r = 1; r < 14; r++)
                for (int j = 0; j < 8; j++) b[j] = _mm_aesenc_si128(b[j], k[r]);
            for (int j = 0; j < 8; j++) {
                const __m128i ks = _mm_aesenclast_si128(b[j], k[14]);
                _mm_storeu_si128((__m128i *)p + j, _mm_xor_si128(_mm_loadu_si128((const __m128i *)p + j), ks));
            }
        }
    }
#endif
    // n blocks
    void encrypt_blocks(const uint8_t *in, uint8_t *out, size_t n) const {
        size_t i = 0;
#if defined(__x86_64__)
        if (ni) {
            for (; i + 8 <= n; i += 8) encrypt8_ni(in + 16 * i, out + 16 * i);
            if (i < n) {
                uint8_t tin[128] = {0}, tout[128];
                memcpy(tin, in + 16 * i, 16 * (n - i));
                encrypt8_ni(tin, tout);
                memcpy(out + 16 * i, tout, 16 * (n - i));
            }
            return;
        }
#endif
        for (; i < n; i++) encrypt_portable(in + 16 * i, out + 16 * i);
    }
};

void os_random(uint8_t *p, size_t n) {
    size_t got = 0;
    while (got < n) {
        const ssize_t r = getrandom(p + got, n - got, 0);
        if (r > 0) got += (size_t)r;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// csprng.UniformSampler (uniform.go)
// ---------------------------------------------------------------------------------------------------------------
constexpr size_t BUF = 8192;         // uniform.go:12
constexpr int FLOAT_PREC = 52;       // gaussian_rounded.go:9

struct Uniform {
    Aes256 aes;
    uint8_t ctr[16];
    uint8_t ks[16];
    int ks_pos = 16;                 // unused bytes of the last keystream block (cipher.Stream semantics)
    alignas(16) uint8_t buf[BUF];
    size_t ptr = BUF;                // uniform.go:51-55

    void init(const uint8_t *key48) {
        uint8_t seed[48];
        if (key48) memcpy(seed, key48, 48); else os_random(seed, 48);
        aes.init(seed);
        memcpy(ctr, seed + 32, 16);
        memset(buf, 0, BUF);
        ptr = BUF;
        ks_pos = 16;
    }
    inline void bump() {                 // 128-bit big-endian increment
        if (__builtin_expect(++ctr[15] != 0, 1)) return;
        for (int i = 14; i >= 0; i--) if (++ctr[i]) break;
    }
    // XORKeyStream(p, p)
    void xor_stream(uint8_t *p, size_t n) {
        size_t i = 0;
        while (i < n && ks_pos < 16) p[i++] ^= ks[ks_pos++];
#if defined(__x86_64__)
        if (aes.ni && n - i >= 128) {
            uint64_t hi = 0, lo = 0;
            for (int b = 0; b < 8; b++) { hi = (hi << 8) | ctr[b]; lo = (lo << 8) | ctr[8 + b]; }
            const size_t blocks = ((n - i) / 128) * 8;
            aes.ctr_xor8_ni(p + i, blocks, hi, lo);
            for (int b = 7; b >= 0; b--) { ctr[b] = (uint8_t)hi; hi >>= 8; ctr[8 + b] = (uint8_t)lo; lo >>= 8; }
            i += 16 * blocks;
        }
#endif
        alignas(16) uint8_t cb[128 * 16], kb[128 * 16];
        while (n - i >= 16) {
            size_t blocks = (n - i) / 16;
            if (blocks > 128) blocks = 128;
            for (size_t b = 0; b < blocks; b++) { memcpy(cb + 16 * b, ctr, 16); bump(); }
            aes.encrypt_blocks(cb, kb, blocks);
            for (size_t b = 0; b < 2 * blocks; b++) {           // 8 bytes at a time (p may be unaligned)
                uint64_t x, y;
                memcpy(&x, p + i + 8 * b, 8); memcpy(&y, kb + 8 * b, 8);
                x ^= y;
                memcpy(p + i + 8 * b, &x, 8);
            }
            i += 16 * blocks;
        }
        if (i < n) {
            aes.encrypt_blocks(ctr, ks, 1);
            bump();
            ks_pos = 0;
            while (i < n) p[i++] ^= ks[ks_pos++];
        }
    }
    inline uint64_t sample() {       // uniform.go:64-84
        if (ptr == BUF) { xor_stream(buf, BUF); ptr = 0; }
        uint64_t r;
        memcpy(&r, buf + ptr, 8);    // little endian
        ptr += 8;
        return r;
    }
    uint64_t sample_n(uint64_t n) {  // :87-95
        const uint64_t bound = ~0ull - (~0ull % n);
        for (;;) { const uint64_t r = sample(); if (r < bound) return r % n; }
    }
    inline double sample_float() {   // :96-100
        const uint64_t r = sample() % (1ull << FLOAT_PREC);
        const uint64_t bits = r | ((uint64_t)(1023 + FLOAT_PREC) << FLOAT_PREC);
        double rf;
        memcpy(&rf, &bits, 8);
        return (rf / (double)(1ull << FLOAT_PREC)) - 1;
    }
};

// ---------------------------------------------------------------------------------------------------------------
// RoundedGaussianSampler (gaussian_rounded.go): ziggurat of Marsaglia & Tsang with 128 blocks
// ---------------------------------------------------------------------------------------------------------------
constexpr int BLOCK = 128;
constexpr double RN = 3.442619855899;

struct Zig {
    uint64_t kn[BLOCK];
    double wn[BLOCK], fn[BLOCK];
    Zig() {                                               // package init, gaussian_rounded.go:22-39
        auto normal = [](double x) { return exp(-0.5 * x * x); };
        auto normal_integral = [](double x) { return sqrt(M_PI / 2) * erfc(x / sqrt(2.0)); };
        auto normal_inv = [](double x) { return sqrt(-2 * log(x)); };
        const double v = RN * normal(RN) + normal_integral(RN);
        double xn[BLOCK] = {0};
        xn[BLOCK - 1] = RN;
        for (int i = BLOCK - 2; i >= 1; i--) xn[i] = normal_inv(v / xn[i + 1] + normal(xn[i + 1]));
        const double scale = (double)(1ull << FLOAT_PREC);
        kn[0] = 0; wn[0] = 0; fn[0] = 0;
        for (int i = 1; i < BLOCK; i++) {
            kn[i] = (uint64_t)((xn[i - 1] / xn[i]) * scale);
            wn[i] = xn[i] / scale;
            fn[i] = normal(xn[i]);
        }
        kn[0] = (uint64_t)((RN * normal(RN) / v) * scale);
        wn[0] = (v / normal(RN)) / scale;
    }
};
const Zig &zig() { static const Zig z; return z; }

struct Rounded {
    Uniform base;
    void init(const uint8_t *key48) { base.init(key48); zig(); }
    double norm_float() {                                 // :77-117
        const Zig &Z = zig();
        for (;;) {
            const uint64_t r = base.sample();
            const uint64_t b = r >> 63;
            const uint64_t i = r % (1ull << 7);
            const uint64_t j = (r >> 7) % (1ull << FLOAT_PREC);
            const double x = (double)(int64_t)((j ^ (0 - b)) + b) * Z.wn[i];
            if (j < Z.kn[i]) return x;
            if (i == 0) {
                double u, v;
                for (;;) {
                    u = -log(base.sample_float()) * (1.0 / RN);
                    v = -log(base.sample_float());
                    if (v + v >= u * u) break;
                }
                u += RN;
                return b == 1 ? -u : u;
            }
            const double f0 = Z.fn[i - 1], f1 = Z.fn[i];
            if (base.sample_float() * (f0 - f1) < exp(-0.5 * x * x) - f1) return x;
        }
    }
    int64_t sample(double center, double std_dev) { return (int64_t)round(center + norm_float() * std_dev); }   // :119-125
};

// ---------------------------------------------------------------------------------------------------------------
// COSACSampler (gaussian_cosac.go)
// ---------------------------------------------------------------------------------------------------------------
struct Cosac {
    Uniform base;
    Rounded rounded;
    void init(const uint8_t *kb, const uint8_t *kr) { base.init(kb); rounded.init(kr); }
    int64_t sample_round(double c_frac, double std_dev) {           // :23-46
        for (;;) {
            const double y = std_dev * rounded.norm_float();
            const uint64_t b = base.sample() & 1;
            double y_round;
            bool cmp;
            if (b == 0) { y_round = round(y) - 1; cmp = y_round <= 0.5; }
            else { y_round = round(y) + 1; cmp = y_round >= -0.5; }
            if (cmp) {
                const double r = base.sample_float();
                const double z = ((y_round + c_frac) * (y_round + c_frac) - y * y) / (2 * std_dev * std_dev);   // test: r < exp(-z)
                // same decision without the exponential when it cannot matter: r < 1 <= exp(-z) for z <= 0, and
                // exp(-z) >= 1 - z, with a margin far above the rounding errors of both sides
                if (z <= 0 || r < (1 - z) - 1e-12) return (int64_t)y_round;
                if (r < exp(-z)) return (int64_t)y_round;
            }
        }
    }
    int64_t sample(double center, double std_dev) {                 // :48-57
        const double c_int = round(center);
        const double c_frac = c_int - center;
        const double r = base.sample_float();
        const double norm = sqrt(2 * M_PI) * std_dev;
        // exp(.) <= 1 and the division is monotone: r >= 1 / norm decides the test below without the exponential
        if (r < 1.0 / norm && r < exp(-(c_frac * c_frac) / (2 * std_dev * std_dev)) / norm) return (int64_t)c_int;
        return sample_round(c_frac, std_dev) + (int64_t)c_int;
    }
};

// ---------------------------------------------------------------------------------------------------------------
// field helpers on Montgomery limbs (the generated zp packages: element.go:303-346 SetRandom, :413-446 Sub, fromMont)
// ---------------------------------------------------------------------------------------------------------------
struct FieldHost {
    int L64, NL, bits;
    uint32_t q[RS_MAX_NL_HOST];
    uint32_t qinv;
};

template <int NL>
void from_mont_t(const FieldHost &f, uint32_t *z, const uint32_t *x) {
    rs::FieldParams<NL> F;
    for (int i = 0; i < NL; i++) F.q[i] = f.q[i];
    F.qinv = f.qinv;
    uint32_t one[NL] = {1};
    uint32_t t[NL];
    rs::mont_mul<NL, 0>(t, x, one, F);
    memcpy(z, t, 4 * NL);
}
template <int NL>
void sub_t(const FieldHost &f, uint32_t *z, const uint32_t *x, const uint32_t *y) {
    rs::FieldParams<NL> F;
    for (int i = 0; i < NL; i++) F.q[i] = f.q[i];
    F.qinv = f.qinv;
    uint32_t t[NL];
    rs::mod_sub<NL>(t, x, y, F);
    memcpy(z, t, 4 * NL);
}
void from_mont(const FieldHost &f, uint32_t *z, const uint32_t *x) {
    switch (f.NL) {
        case 4: from_mont_t<4>(f, z, x); break;
        case 8: from_mont_t<8>(f, z, x); break;
        case 14: from_mont_t<14>(f, z, x); break;
        case 28: from_mont_t<28>(f, z, x); break;
    }
}
void field_sub(const FieldHost &f, uint32_t *z, const uint32_t *x, const uint32_t *y) {
    switch (f.NL) {
        case 4: sub_t<4>(f, z, x, y); break;
        case 8: sub_t<8>(f, z, x, y); break;
        case 14: sub_t<14>(f, z, x, y); break;
        case 28: sub_t<28>(f, z, x, y); break;
    }
}

// source of the bytes crypto/rand.Reader would hand out
struct ByteSource {
    const uint8_t *fixed = nullptr;      // test mode: preset bytes
    size_t fixed_n = 0, fixed_pos = 0;
    Uniform drbg;                        // default: AES-CTR generator keyed from the OS
    bool drbg_ready = false;
    bool read(uint8_t *p, size_t n) {
        if (fixed) {
            if (fixed_pos + n > fixed_n) return false;
            memcpy(p, fixed + fixed_pos, n);
            fixed_pos += n;
            return true;
        }
        if (!drbg_ready) { drbg.init(nullptr); drbg_ready = true; }
        // the same keystream bytes in the same order, fetched 16 KiB at a time (a MustSetRandom element asks for 16-110 bytes: the
        // per-call block bookkeeping cost more than the AES)
        while (n) {
            if (pool_pos == pool_len) {
                memset(pool, 0, sizeof(pool));
                drbg.xor_stream(pool, sizeof(pool));
                pool_pos = 0; pool_len = sizeof(pool);
            }
            const size_t take = n < pool_len - pool_pos ? n : pool_len - pool_pos;
            memcpy(p, pool + pool_pos, take);
            pool_pos += take; p += take; n -= take;
        }
        return true;
    }
    alignas(16) uint8_t pool[16384];
    size_t pool_pos = 0, pool_len = 0;
};

// SetRandom (element.go:303-346): ceil(bits/8) bytes, unused top bits cleared, retry while >= q; the raw limbs ARE the
// Montgomery representation
bool set_random(const FieldHost &f, ByteSource &src, uint64_t *out) {
    const int k = (f.bits + 7) / 8;
    const int b = f.bits % 8 ? f.bits % 8 : 8;
    uint8_t bytes[RS_MAX_NL_HOST * 4];
    for (;;) {
        memset(bytes, 0, 8 * (size_t)f.L64);
        if (!src.read(bytes, k)) return false;
        bytes[k - 1] &= (uint8_t)((1u << b) - 1);
        uint64_t limbs[RS_MAX_NL_HOST / 2];
        memcpy(limbs, bytes, 8 * f.L64);
        bool lt = false;
        for (int i = f.L64 - 1; i >= 0; i--) {
            const uint64_t qi = (uint64_t)f.q[2 * i] | ((uint64_t)f.q[2 * i + 1] << 32);
            if (limbs[i] < qi) { lt = true; break; }
            if (limbs[i] > qi) break;
        }
        if (lt) { memcpy(out, limbs, 8 * f.L64); return true; }
    }
}

inline uint64_t div_mod64(uint64_t *x, int n, uint64_t y) {      // jindo/utils.go:12-18
    if (y < (1ull << 32)) {
        // the encoder's base is below 2^25: two 64-bit hardware divisions per limb instead of a 128-bit one
        uint64_t r = 0;
        for (int i = n - 1; i >= 0; i--) {
            const uint64_t t1 = (r << 32) | (x[i] >> 32);
            const uint64_t q1 = t1 / y;
            r = t1 - q1 * y;
            const uint64_t t0 = (r << 32) | (x[i] & 0xffffffffull);
            const uint64_t q0 = t0 / y;
            r = t0 - q0 * y;
            x[i] = (q1 << 32) | q0;
        }
        return r;
    }
    unsigned __int128 r = 0;
    for (int i = n - 1; i >= 0; i--) {
        const unsigned __int128 cur = (r << 64) | x[i];
        x[i] = (uint64_t)(cur / y);
        r = cur % y;
    }
    return (uint64_t)r;
}

// staging buffers: pinned when a CUDA device is there (async uploads), plain aligned memory otherwise (host-only use of the
// samplers, e.g. the CPU tests); `pinned` remembers which
struct Staging { void *p; bool pinned; };
void *stage_alloc(size_t bytes, std::vector<Staging> &all) {
    void *p = rs_pinned_alloc(bytes);
    const bool pinned = p != nullptr;
    if (!p && posix_memalign(&p, 64, bytes ? bytes : 64) != 0) p = nullptr;
    if (p) all.push_back(Staging{p, pinned});
    return p;
}

struct Samplers {
    Cosac cosac;          // Encoder.cosac
    Rounded rounded;      // Prover.roundedSampler
    ByteSource rnd;       // crypto/rand
};

}  // namespace

// the ziggurat tables of gaussian_rounded.go's package init, for the device sampler (cosac_dev.cu)
extern "C" void rs_zig_tables(uint64_t *kn, double *wn, double *fn) {
    const Zig &Z = zig();
    memcpy(kn, Z.kn, sizeof(Z.kn)); memcpy(wn, Z.wn, sizeof(Z.wn)); memcpy(fn, Z.fn, sizeof(Z.fn));
}

struct rs_uniform { Uniform u; };
struct rs_rounded { Rounded r; };
struct rs_cosac { Cosac c; };

struct rs_tape_builder {
    rs_jindo_params p;
    rs_jindo_noise nz;
    FieldHost f;
    int T;
    std::vector<Samplers *> smp;                  // one set per worker
    bool deterministic;
    // pinned buffers, n_slots slots (two for the double-buffered Commit loop; one per Commit of a proof for the Buckler prover)
    struct Slot {
        uint64_t *first = nullptr, *last = nullptr, *mask = nullptr;
        int64_t *cosac = nullptr, *mlwe = nullptr;
        size_t n_mask_rows = 0, n_cosac = 0;
        std::atomic<int> pending{0};
        std::chrono::steady_clock::time_point t0;
        double ms = 0;
        std::atomic<int> error{0};
    };
    std::deque<Slot> slot;
    int n_slots = 2;
    bool device_cosac = false;                    // the COSAC samples are drawn on the device (rs_jindo_set_cosac_stream): skip them here
    ByteSource main_rnd;                          // genFirstLastRow's draws when the workers own their own generators
    std::vector<Staging> staging;
    size_t cap_mask_rows;
    // worker pool
    std::vector<std::thread> threads;
    std::mutex mu;
    std::condition_variable cv, cv_done;
    std::deque<std::function<void(int)>> jobs;
    bool stop = false;
};

namespace {

void worker_loop(rs_tape_builder *b, int w) {
    for (;;) {
        std::function<void(int)> job;
        {
            std::unique_lock<std::mutex> lk(b->mu);
            b->cv.wait(lk, [&] { return b->stop || !b->jobs.empty(); });
            if (b->stop && b->jobs.empty()) return;
            job = std::move(b->jobs.front());
            b->jobs.pop_front();
        }
        job(w);
    }
}

// COSAC polynomial of `cnt` elements (Montgomery limbs): base-b digits (encoder.go:120-146), float64 centres
// (:152-164), one COSACSampler.Sample per coefficient (:166-171)
void cosac_poly(const rs_tape_builder *b, Cosac &cs, const uint64_t *elems, int cnt, double std_dev, int64_t *g) {
    const rs_jindo_params &p = b->p;
    const int N = p.d, slots = p.slots, L = b->f.L64;
    uint64_t digits[256];
    double fp[256];
    for (int i = 0; i < N; i++) { digits[i] = 0; fp[i] = 0.0; }
    for (int i = 0; i < cnt; i++) {
        uint64_t c[RS_MAX_NL_HOST / 2];
        from_mont(b->f, (uint32_t *)c, (const uint32_t *)(elems + (size_t)i * L));     // Slice(): canonical limbs
        int top = L;                                             // limbs above `top` are zero: the quotient shrinks
        for (int j = 0; j < p.exp - 1; j++) {
            while (top > 1 && c[top - 1] == 0) top--;
            digits[j * slots + i] = div_mod64(c, top, p.base);
        }
        digits[(p.exp - 1) * slots + i] = c[0];
    }
    for (int i = 0; i < p.exp; i++) {
        const double di = p.delta_inv[i];
        if (di == 0) continue;
        const int dd = N - (i + 1) * slots;
        for (int j = 0, jj = dd; jj < N; j++, jj++) fp[jj] += di * (double)digits[j];
        for (int j = N - dd, jj = 0; j < N; j++, jj++) fp[jj] -= di * (double)digits[j];
    }
    for (int i = 0; i < N; i++) g[i] = cs.sample(-fp[i], std_dev);
}

void finish_one(rs_tape_builder *b, rs_tape_builder::Slot &s) {
    if (s.pending.fetch_sub(1) == 1) {
        s.ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - s.t0).count();
        std::lock_guard<std::mutex> lk(b->mu);
        b->cv_done.notify_all();
    }
}

}  // namespace

extern "C" {
#pragma GCC visibility push(default)

rs_uniform *rs_uniform_create(const uint8_t *key48) { rs_uniform *u = new rs_uniform(); u->u.init(key48); return u; }
void rs_uniform_destroy(rs_uniform *u) { delete u; }
uint64_t rs_uniform_sample(rs_uniform *u) { return u->u.sample(); }
void rs_uniform_sample_many(rs_uniform *u, uint64_t *out, size_t n) { for (size_t i = 0; i < n; i++) out[i] = u->u.sample(); }
uint64_t rs_uniform_sample_n(rs_uniform *u, uint64_t n) { return u->u.sample_n(n); }
double rs_uniform_sample_float(rs_uniform *u) { return u->u.sample_float(); }
void rs_uniform_read(rs_uniform *u, uint8_t *p, size_t n) { u->u.xor_stream(p, n); }

rs_rounded *rs_rounded_create(const uint8_t *key48) { rs_rounded *r = new rs_rounded(); r->r.init(key48); return r; }
void rs_rounded_destroy(rs_rounded *r) { delete r; }
double rs_rounded_norm_float(rs_rounded *r) { return r->r.norm_float(); }
int64_t rs_rounded_sample(rs_rounded *r, double center, double std_dev) { return r->r.sample(center, std_dev); }

rs_cosac *rs_cosac_create(const uint8_t *kb, const uint8_t *kr) { rs_cosac *c = new rs_cosac(); c->c.init(kb, kr); return c; }
void rs_cosac_destroy(rs_cosac *c) { delete c; }
int64_t rs_cosac_sample(rs_cosac *c, double center, double std_dev) { return c->c.sample(center, std_dev); }

int rs_compute_cdt(double center, double sigma, uint64_t *table, int cap, int64_t *tail_lo) {
    const int64_t tail_hi = (int64_t)ceil(9 * sigma);
    const int64_t lo = -tail_hi;
    const int64_t size = tail_hi - lo + 1;
    if (!table || size > cap) return RS_ERR_ARG;
    double cdf = 0.0;
    const double norm = sqrt(2 * M_PI) * sigma;
    int i = 0;
    for (int64_t x = lo; x <= tail_hi; x++, i++) {
        const double xf = (double)x;
        const double rho = exp(-(xf - center) * (xf - center) / (2 * sigma * sigma)) / norm;
        cdf += rho;
        if (cdf > 1) table[i] = ~0ull;
        else {
            // uint64(math.Round(cdf * 2^64)): a product of exactly 2^64 saturates like the Go conversion on amd64 does not
            // (it is undefined there); the tables of the shipped parameters never reach it before cdf > 1
            const double v = round(cdf * 18446744073709551616.0);
            table[i] = v >= 18446744073709551616.0 ? ~0ull : (uint64_t)v;
        }
    }
    if (tail_lo) *tail_lo = lo;
    return (int)size;
}

int rs_tape_builder_create(const rs_jindo_params *p, const rs_jindo_noise *noise, const uint64_t *q, int L64, int n_threads,
                           const uint8_t *seeds, rs_tape_builder **out) {
    return rs_tape_builder_create_slots(p, noise, q, L64, n_threads, 2, seeds, out);
}

int rs_tape_builder_create_slots(const rs_jindo_params *p, const rs_jindo_noise *noise, const uint64_t *q, int L64, int n_threads,
                                 int n_slots, const uint8_t *seeds, rs_tape_builder **out) {
    if (!p || !noise || !q || !out || n_threads < 1 || n_slots < 1 || n_slots > 4096 || p->d != 256) return RS_ERR_ARG;
    if (L64 != 2 && L64 != 4 && L64 != 7 && L64 != 14) return RS_ERR_UNSUPPORTED;
    if (seeds && n_threads != 1) return RS_ERR_ARG;
    rs_tape_builder *b = new rs_tape_builder();
    b->p = *p; b->nz = *noise; b->T = n_threads; b->deterministic = seeds != nullptr;
    b->n_slots = n_slots;
    b->slot.resize((size_t)n_slots);
    b->f.L64 = L64; b->f.NL = 2 * L64;
    for (int i = 0; i < L64; i++) { b->f.q[2 * i] = (uint32_t)q[i]; b->f.q[2 * i + 1] = (uint32_t)(q[i] >> 32); }
    uint32_t inv = 1;
    for (int i = 0; i < 5; i++) inv *= 2 - b->f.q[0] * inv;
    b->f.qinv = 0u - inv;
    b->f.bits = 0;
    for (int i = 2 * L64 - 1; i >= 0; i--) if (b->f.q[i]) { b->f.bits = 32 * i + (32 - __builtin_clz(b->f.q[i])); break; }
    for (int t = 0; t < n_threads; t++) {
        Samplers *s = new Samplers();
        s->cosac.init(seeds ? seeds : nullptr, seeds ? seeds + 48 : nullptr);
        s->rounded.init(seeds ? seeds + 96 : nullptr);
        b->smp.push_back(s);
    }
    const size_t cs = (size_t)p->cols * p->slots, L = (size_t)L64;
    b->cap_mask_rows = (size_t)p->rows;
    const size_t nm = (size_t)(p->in_msis + p->mlwe);
    for (int k = 0; k < n_slots; k++) {
        auto &s = b->slot[k];
        s.first = (uint64_t *)stage_alloc(cs * L * 8, b->staging);
        s.last = (uint64_t *)stage_alloc(cs * L * 8, b->staging);
        s.mask = (uint64_t *)stage_alloc(b->cap_mask_rows * p->slots * L * 8, b->staging);
        s.cosac = (int64_t *)stage_alloc(((size_t)p->cols + b->cap_mask_rows) * 256 * 8, b->staging);
        s.mlwe = (int64_t *)stage_alloc(nm * 256 * 8, b->staging);
        if (!s.first || !s.last || !s.mask || !s.cosac || !s.mlwe) { rs_tape_builder_destroy(b); return RS_ERR_CUDA; }
    }
    for (int t = 0; t < n_threads; t++) b->threads.emplace_back(worker_loop, b, t);
    *out = b;
    return RS_OK;
}

void rs_tape_builder_destroy(rs_tape_builder *b) {
    if (!b) return;
    {
        std::lock_guard<std::mutex> lk(b->mu);
        b->stop = true;
    }
    b->cv.notify_all();
    for (auto &t : b->threads) t.join();
    for (auto &st : b->staging) { if (st.pinned) rs_pinned_free(st.p); else free(st.p); }
    for (auto *s : b->smp) delete s;
    delete b;
}

int rs_tape_builder_set_device_cosac(rs_tape_builder *b, int on) {
    if (!b) return RS_ERR_ARG;
    b->device_cosac = on != 0;
    return RS_OK;
}

int rs_tape_builder_set_random_bytes(rs_tape_builder *b, const uint8_t *bytes, size_t n) {
    if (!b || b->T != 1) return RS_ERR_ARG;
    b->smp[0]->rnd.fixed = bytes; b->smp[0]->rnd.fixed_n = n; b->smp[0]->rnd.fixed_pos = 0;
    return RS_OK;
}

int rs_tape_builder_start(rs_tape_builder *b, const uint64_t *v_host, size_t len, int k) {
    if (!b || !v_host || k < 0 || k >= b->n_slots || len < 1) return RS_ERR_ARG;
    const rs_jindo_params &p = b->p;
    const size_t cs = (size_t)p.cols * p.slots, L = (size_t)b->f.L64;
    if (len > cs * (size_t)(p.rows - 1)) return RS_ERR_SHAPE;                         // prover.go:47-49
    auto &s = b->slot[k];
    if (s.pending.load() != 0) return RS_ERR_STATE;
    if (b->deterministic) {
        // one sequential crypto/rand stream shared with worker 0: wait until the other slots' jobs have drained
        std::unique_lock<std::mutex> lk(b->mu);
        b->cv_done.wait(lk, [&] { for (auto &o : b->slot) if (o.pending.load() != 0) return false; return true; });
    }
    s.t0 = std::chrono::steady_clock::now();
    s.error = 0;
    // genFirstLastRow (prover.go:65-86) on the calling thread: cols*slots - 1 draws, a few microseconds
    ByteSource &r0 = b->deterministic ? b->smp[0]->rnd : b->main_rnd;
    for (size_t i = 0; i + 1 < cs; i++)
        if (!set_random(b->f, r0, s.last + i * L)) return RS_ERR_ARG;
    memset(s.last + (cs - 1) * L, 0, L * 8);
    memcpy(s.first, v_host, L * 8);
    uint64_t zero[RS_MAX_NL_HOST / 2] = {0};
    for (size_t i = 1; i < cs; i++)
        field_sub(b->f, (uint32_t *)(s.first + i * L), (const uint32_t *)(i < len ? v_host + i * L : zero), (const uint32_t *)(s.last + (i - 1) * L));
    // rows of the mask column that are encoded (prover.go:93-115): row 0, rows 1..rows-2 while r*cols*slots <= len, row rows-1
    size_t M = 1;
    for (int r = 1; r < p.rows - 1; r++) { if ((size_t)r * cs > len) break; M++; }
    M++;
    s.n_mask_rows = M;
    s.n_cosac = (size_t)p.cols + M;
    const size_t nm = (size_t)(p.in_msis + p.mlwe);
    // jobs: COSAC polynomials in chunks, then the mask column's MLWE noise
    const size_t total = s.n_cosac;
    const size_t chunk = b->T == 1 ? total : (total + (size_t)b->T * 4 - 1) / ((size_t)b->T * 4);
    std::vector<std::function<void(int)>> jobs;
    for (size_t a = 0; a < total; a += chunk) {
        const size_t e = a + chunk < total ? a + chunk : total;
        jobs.push_back([b, &s, a, e, L](int w) {
            const rs_jindo_params &p = b->p;
            Samplers &S = *b->smp[w];
            for (size_t t = a; t < e && !s.error.load(); t++) {
                if (t < (size_t)p.cols) {
                    if (!b->device_cosac) cosac_poly(b, S.cosac, s.first + t * p.slots * L, p.slots, b->nz.ecd_blind_std, s.cosac + t * 256);
                } else {
                    const size_t r = t - p.cols;
                    uint64_t *row = s.mask + r * p.slots * L;
                    for (int i = 0; i < p.slots; i++)
                        if (!set_random(b->f, S.rnd, row + (size_t)i * L)) { s.error = RS_ERR_ARG; break; }
                    if (!b->device_cosac) cosac_poly(b, S.cosac, row, p.slots, r == 0 ? b->nz.mask_blind_std : b->nz.mask_std, s.cosac + t * 256);
                }
            }
            finish_one(b, s);
        });
    }
    jobs.push_back([b, &s, nm](int w) {
        Samplers &S = *b->smp[w];
        for (size_t i = 0; i < nm * 256; i++) s.mlwe[i] = S.rounded.sample(0, b->nz.mask_mlwe_std);     // prover.go:131-134
        finish_one(b, s);
    });
    s.pending.store((int)jobs.size());
    {
        std::lock_guard<std::mutex> lk(b->mu);
        for (auto &j : jobs) b->jobs.push_back(std::move(j));
    }
    b->cv.notify_all();
    return RS_OK;
}

// MustSetRandom (jindo-modulus element.go, rejection sampling over crypto/rand bytes) for n elements, Montgomery limbs out
int rs_tape_builder_random_elements(rs_tape_builder *b, uint64_t *out_mont, size_t n) {
    if (!b || (!out_mont && n)) return RS_ERR_ARG;
    ByteSource &r0 = b->deterministic ? b->smp[0]->rnd : b->main_rnd;
    for (size_t i = 0; i < n; i++)
        if (!set_random(b->f, r0, out_mont + i * (size_t)b->f.L64)) return RS_ERR_ARG;
    return RS_OK;
}

// Prover.sumCheckMask (buckler/prover.go:381-397): rank random coefficients, maskSum = the first one, then every further
// coefficient is random and folded back: mask[i - rank] -= mask[i]
int rs_tape_builder_sum_check_mask(rs_tape_builder *b, size_t rank, size_t mask_rank, uint64_t *mask_out, uint64_t *mask_sum_out) {
    if (!b || !mask_out || !mask_sum_out || rank < 1 || mask_rank < rank) return RS_ERR_ARG;
    const size_t L = (size_t)b->f.L64;
    // the draws in the reference's order (coefficient 0 .. mask_rank - 1), then the fold: coefficient i - rank only ever
    // subtracts coefficient i, which is final by then
    const int rc = rs_tape_builder_random_elements(b, mask_out, mask_rank);
    if (rc) return rc;
    memcpy(mask_sum_out, mask_out, L * 8);
    for (size_t i = rank; i < mask_rank; i++)
        field_sub(b->f, (uint32_t *)(mask_out + (i - rank) * L), (const uint32_t *)(mask_out + (i - rank) * L), (const uint32_t *)(mask_out + i * L));
    return RS_OK;
}

int rs_tape_builder_wait(rs_tape_builder *b, int k, rs_jindo_tape *tape, double *host_ms) {
    if (!b || !tape || k < 0 || k >= b->n_slots) return RS_ERR_ARG;
    auto &s = b->slot[k];
    {
        std::unique_lock<std::mutex> lk(b->mu);
        b->cv_done.wait(lk, [&] { return s.pending.load() == 0; });
    }
    if (s.error.load()) return s.error.load();
    memset(tape, 0, sizeof(*tape));
    tape->first_row = s.first; tape->last_row = s.last; tape->mask_rows = s.mask; tape->n_mask_rows = s.n_mask_rows;
    tape->twin_u = nullptr; tape->n_twin_u = 0;
    tape->cosac_g = b->device_cosac ? nullptr : s.cosac; tape->n_cosac_polys = s.n_cosac;
    tape->mlwe = s.mlwe; tape->mlwe_mask_only = 1;
    tape->on_device = 0;
    if (host_ms) *host_ms = s.ms;
    return RS_OK;
}

#pragma GCC visibility pop
}
