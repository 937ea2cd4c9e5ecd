// buckler_host.cpp -- buckler.Prover.Prove (buckler/prover.go:42-353) as compiled host code over the library's own C ABI
// (include/ringo_buckler.h).  Host side only: SHA-256 / SHAKE128 transcripts, round structure, launch order.  Every
// polynomial lives in HBM between the first upload of the witnesses and the download of the proof.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <chrono>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#if defined(__x86_64__)
#include <cpuid.h>
#include <immintrin.h>
#endif

#include "../../include/ringo_buckler.h"
#include "../../include/ringo_sampler.h"

namespace rs { int set_error(int code, const char *fmt, ...); }

namespace {

// ---------------------------------------------------------------------------------------------------------------
// SHA-256 (FIPS 180-4)
// ---------------------------------------------------------------------------------------------------------------
const uint32_t K256[64] = {
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01, 0x243185be,
    0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa,
    0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967, 0x27b70a85,
    0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85, 0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3,
    0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f,
    0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};

inline uint32_t rotr(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }

#if defined(__x86_64__)
bool have_sha_ni() {
    static const bool ok = [] {
        unsigned a, b, c, d;
        if (!__get_cpuid_count(7, 0, &a, &b, &c, &d)) return false;
        const bool sha = (b >> 29) & 1;
        if (!__get_cpuid(1, &a, &b, &c, &d)) return false;
        return sha && ((c >> 19) & 1) && ((c >> 9) & 1);      // SHA, SSE4.1, SSSE3
    }();
    return ok;
}

// the compression function on the SHA extensions (two rounds per SHA256RNDS2, message schedule by SHA256MSG1/2)
__attribute__((target("sha,sse4.1,ssse3"))) void sha256_blocks_ni(uint32_t state[8], const uint8_t *p, size_t n_blocks) {
    const __m128i shuf = _mm_set_epi64x(0x0c0d0e0f08090a0bULL, 0x0405060700010203ULL);
    __m128i tmp = _mm_loadu_si128((const __m128i *)&state[0]);        // DCBA
    __m128i st1 = _mm_loadu_si128((const __m128i *)&state[4]);        // HGFE
    tmp = _mm_shuffle_epi32(tmp, 0xB1);                               // CDAB
    st1 = _mm_shuffle_epi32(st1, 0x1B);                               // EFGH
    __m128i st0 = _mm_alignr_epi8(tmp, st1, 8);                       // ABEF
    st1 = _mm_blend_epi16(st1, tmp, 0xF0);                            // CDGH
    for (; n_blocks; n_blocks--, p += 64) {
        const __m128i save0 = st0, save1 = st1;
        __m128i m[4];
        for (int i = 0; i < 4; i++) m[i] = _mm_shuffle_epi8(_mm_loadu_si128((const __m128i *)(p + 16 * i)), shuf);
        for (int r = 0; r < 16; r++) {
            __m128i w;
            if (r < 4) {
                w = m[r];
            } else {                                                  // W[4r .. 4r+3] from the previous sixteen words
                __m128i t = _mm_sha256msg1_epu32(m[0], m[1]);
                t = _mm_add_epi32(t, _mm_alignr_epi8(m[3], m[2], 4));
                w = _mm_sha256msg2_epu32(t, m[3]);
                m[0] = m[1]; m[1] = m[2]; m[2] = m[3]; m[3] = w;
            }
            __m128i wk = _mm_add_epi32(w, _mm_loadu_si128((const __m128i *)&K256[4 * r]));
            st1 = _mm_sha256rnds2_epu32(st1, st0, wk);
            wk = _mm_shuffle_epi32(wk, 0x0E);
            st0 = _mm_sha256rnds2_epu32(st0, st1, wk);
        }
        st0 = _mm_add_epi32(st0, save0);
        st1 = _mm_add_epi32(st1, save1);
    }
    tmp = _mm_shuffle_epi32(st0, 0x1B);                               // FEBA
    st1 = _mm_shuffle_epi32(st1, 0xB1);                               // DCHG
    st0 = _mm_blend_epi16(tmp, st1, 0xF0);                            // DCBA
    st1 = _mm_alignr_epi8(st1, tmp, 8);                               // HGFE
    _mm_storeu_si128((__m128i *)&state[0], st0);
    _mm_storeu_si128((__m128i *)&state[4], st1);
}
#endif

struct Sha256 {
    uint32_t h[8];
    uint8_t buf[64];
    size_t fill = 0;
    uint64_t total = 0;
    Sha256() {
        static const uint32_t iv[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
        memcpy(h, iv, sizeof(iv));
    }
    void block(const uint8_t *p) {
        uint32_t w[64];
        for (int i = 0; i < 16; i++) w[i] = (uint32_t)p[4 * i] << 24 | (uint32_t)p[4 * i + 1] << 16 | (uint32_t)p[4 * i + 2] << 8 | p[4 * i + 3];
        for (int i = 16; i < 64; i++) {
            const uint32_t s0 = rotr(w[i - 15], 7) ^ rotr(w[i - 15], 18) ^ (w[i - 15] >> 3);
            const uint32_t s1 = rotr(w[i - 2], 17) ^ rotr(w[i - 2], 19) ^ (w[i - 2] >> 10);
            w[i] = w[i - 16] + s0 + w[i - 7] + s1;
        }
        uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
        for (int i = 0; i < 64; i++) {
            const uint32_t S1 = rotr(e, 6) ^ rotr(e, 11) ^ rotr(e, 25), ch = (e & f) ^ (~e & g);
            const uint32_t t1 = hh + S1 + ch + K256[i] + w[i];
            const uint32_t S0 = rotr(a, 2) ^ rotr(a, 13) ^ rotr(a, 22), mj = (a & b) ^ (a & c) ^ (b & c);
            const uint32_t t2 = S0 + mj;
            hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
        }
        h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
    }
    void update(const uint8_t *p, size_t n) {
        total += n;
        if (fill) {
            const size_t take = n < 64 - fill ? n : 64 - fill;
            memcpy(buf + fill, p, take);
            fill += take; p += take; n -= take;
            if (fill == 64) { block(buf); fill = 0; }
        }
#if defined(__x86_64__)
        if (n >= 64 && have_sha_ni()) {
            sha256_blocks_ni(h, p, n / 64);
            p += n & ~(size_t)63; n &= 63;
        }
#endif
        for (; n >= 64; p += 64, n -= 64) block(p);
        if (n) { memcpy(buf, p, n); fill = n; }
    }
    void final(uint8_t out[32]) {
        const uint64_t bits = total * 8;
        uint8_t pad[72] = {0x80};
        const size_t padlen = (fill < 56 ? 56 - fill : 120 - fill);
        uint8_t len[8];
        for (int i = 0; i < 8; i++) len[i] = (uint8_t)(bits >> (56 - 8 * i));
        update(pad, padlen);
        update(len, 8);
        for (int i = 0; i < 8; i++) { out[4 * i] = h[i] >> 24; out[4 * i + 1] = h[i] >> 16; out[4 * i + 2] = h[i] >> 8; out[4 * i + 3] = h[i]; }
    }
};

// ---------------------------------------------------------------------------------------------------------------
// SHAKE128 (FIPS 202): Keccak-f[1600], rate 168, domain byte 0x1f; absorb then squeeze, both incremental
// ---------------------------------------------------------------------------------------------------------------
#define RS_ROL(x, n) (((x) << (n)) | ((x) >> (64 - (n))))
// one round from lanes a** into lanes e** (theta, rho, pi, chi, iota), lanes named a[y][x] = a<y><x>
#define RS_KROUND(A, E, rc)                                                                                           \
    do {                                                                                                              \
        const uint64_t c0 = A##00 ^ A##10 ^ A##20 ^ A##30 ^ A##40, c1 = A##01 ^ A##11 ^ A##21 ^ A##31 ^ A##41,         \
                       c2 = A##02 ^ A##12 ^ A##22 ^ A##32 ^ A##42, c3 = A##03 ^ A##13 ^ A##23 ^ A##33 ^ A##43,         \
                       c4 = A##04 ^ A##14 ^ A##24 ^ A##34 ^ A##44;                                                     \
        const uint64_t d0 = c4 ^ RS_ROL(c1, 1), d1 = c0 ^ RS_ROL(c2, 1), d2 = c1 ^ RS_ROL(c3, 1), d3 = c2 ^ RS_ROL(c4, 1), \
                       d4 = c3 ^ RS_ROL(c0, 1);                                                                        \
        uint64_t b0, b1, b2, b3, b4;                                                                                  \
        b0 = A##00 ^ d0; b1 = RS_ROL(A##11 ^ d1, 44); b2 = RS_ROL(A##22 ^ d2, 43); b3 = RS_ROL(A##33 ^ d3, 21); b4 = RS_ROL(A##44 ^ d4, 14); \
        E##00 = b0 ^ (~b1 & b2) ^ (rc); E##01 = b1 ^ (~b2 & b3); E##02 = b2 ^ (~b3 & b4); E##03 = b3 ^ (~b4 & b0); E##04 = b4 ^ (~b0 & b1); \
        b0 = RS_ROL(A##03 ^ d3, 28); b1 = RS_ROL(A##14 ^ d4, 20); b2 = RS_ROL(A##20 ^ d0, 3); b3 = RS_ROL(A##31 ^ d1, 45); b4 = RS_ROL(A##42 ^ d2, 61); \
        E##10 = b0 ^ (~b1 & b2); E##11 = b1 ^ (~b2 & b3); E##12 = b2 ^ (~b3 & b4); E##13 = b3 ^ (~b4 & b0); E##14 = b4 ^ (~b0 & b1); \
        b0 = RS_ROL(A##01 ^ d1, 1); b1 = RS_ROL(A##12 ^ d2, 6); b2 = RS_ROL(A##23 ^ d3, 25); b3 = RS_ROL(A##34 ^ d4, 8); b4 = RS_ROL(A##40 ^ d0, 18); \
        E##20 = b0 ^ (~b1 & b2); E##21 = b1 ^ (~b2 & b3); E##22 = b2 ^ (~b3 & b4); E##23 = b3 ^ (~b4 & b0); E##24 = b4 ^ (~b0 & b1); \
        b0 = RS_ROL(A##04 ^ d4, 27); b1 = RS_ROL(A##10 ^ d0, 36); b2 = RS_ROL(A##21 ^ d1, 10); b3 = RS_ROL(A##32 ^ d2, 15); b4 = RS_ROL(A##43 ^ d3, 56); \
        E##30 = b0 ^ (~b1 & b2); E##31 = b1 ^ (~b2 & b3); E##32 = b2 ^ (~b3 & b4); E##33 = b3 ^ (~b4 & b0); E##34 = b4 ^ (~b0 & b1); \
        b0 = RS_ROL(A##02 ^ d2, 62); b1 = RS_ROL(A##13 ^ d3, 55); b2 = RS_ROL(A##24 ^ d4, 39); b3 = RS_ROL(A##30 ^ d0, 41); b4 = RS_ROL(A##41 ^ d1, 2); \
        E##40 = b0 ^ (~b1 & b2); E##41 = b1 ^ (~b2 & b3); E##42 = b2 ^ (~b3 & b4); E##43 = b3 ^ (~b4 & b0); E##44 = b4 ^ (~b0 & b1); \
    } while (0)

void keccak_f(uint64_t s[25]) {
    static const uint64_t RC[24] = {0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808aULL, 0x8000000080008000ULL, 0x000000000000808bULL,
                                    0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL, 0x000000000000008aULL, 0x0000000000000088ULL,
                                    0x0000000080008009ULL, 0x000000008000000aULL, 0x000000008000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL,
                                    0x8000000000008003ULL, 0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800aULL, 0x800000008000000aULL,
                                    0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
    // lane [x + 5y] of the state is a<y><x>
    uint64_t a00 = s[0], a01 = s[1], a02 = s[2], a03 = s[3], a04 = s[4], a10 = s[5], a11 = s[6], a12 = s[7], a13 = s[8], a14 = s[9],
             a20 = s[10], a21 = s[11], a22 = s[12], a23 = s[13], a24 = s[14], a30 = s[15], a31 = s[16], a32 = s[17], a33 = s[18], a34 = s[19],
             a40 = s[20], a41 = s[21], a42 = s[22], a43 = s[23], a44 = s[24];
    uint64_t e00, e01, e02, e03, e04, e10, e11, e12, e13, e14, e20, e21, e22, e23, e24, e30, e31, e32, e33, e34, e40, e41, e42, e43, e44;
    for (int r = 0; r < 24; r += 2) {
        RS_KROUND(a, e, RC[r]);
        RS_KROUND(e, a, RC[r + 1]);
    }
    s[0] = a00; s[1] = a01; s[2] = a02; s[3] = a03; s[4] = a04; s[5] = a10; s[6] = a11; s[7] = a12; s[8] = a13; s[9] = a14;
    s[10] = a20; s[11] = a21; s[12] = a22; s[13] = a23; s[14] = a24; s[15] = a30; s[16] = a31; s[17] = a32; s[18] = a33; s[19] = a34;
    s[20] = a40; s[21] = a41; s[22] = a42; s[23] = a43; s[24] = a44;
}

struct Shake128 {
    static constexpr size_t RATE = 168;
    uint64_t s[25];
    size_t pos = 0;
    bool squeezing = false;
    Shake128() { memset(s, 0, sizeof(s)); }
    void xor_byte(size_t i, uint8_t v) { s[i / 8] ^= (uint64_t)v << (8 * (i % 8)); }
    void absorb(const uint8_t *p, size_t n) {
        while (n) {
            if (pos == 0 && n >= RATE) {                       // whole blocks, word-wise
                for (size_t k = 0; k < RATE / 8; k++) { uint64_t w; memcpy(&w, p + 8 * k, 8); s[k] ^= w; }
                keccak_f(s);
                p += RATE; n -= RATE;
                continue;
            }
            xor_byte(pos++, *p++); n--;
            if (pos == RATE) { keccak_f(s); pos = 0; }
        }
    }
    void squeeze(uint8_t *out, size_t n) {
        if (!squeezing) {
            xor_byte(pos, 0x1f);
            xor_byte(RATE - 1, 0x80);
            keccak_f(s);
            pos = 0;
            squeezing = true;
        }
        while (n) {
            if (pos == RATE) { keccak_f(s); pos = 0; }
            *out++ = (uint8_t)(s[pos / 8] >> (8 * (pos % 8)));
            pos++; n--;
        }
    }
};

// ---------------------------------------------------------------------------------------------------------------
// gnark-crypto fiat-shamir Transcript over SHA-256 (consensys/gnark-crypto v0.19.2 fiat-shamir/transcript.go, as the reference
// uses it at buckler/prover.go:113-121): challenge_i = H(name_i || challenge_{i-1} || bindings_i in Bind order)
// ---------------------------------------------------------------------------------------------------------------
const char *const CHAL_NAMES[6] = {"projConst", "arithBatchConst", "linCheckBatchConst", "linCheckConst", "sumCheckBatchConst", "evalPoint"};
enum { PROJ = 0, ARITH_BC, LIN_BC, LIN_C, SUM_BC, EVAL_PT };

struct Transcript {
    std::vector<uint8_t> bindings[6];
    uint8_t value[6][32];
    void bind(int which, const uint8_t *p, size_t n) { bindings[which].insert(bindings[which].end(), p, p + n); }
    const uint8_t *compute(int which) {
        Sha256 h;
        h.update((const uint8_t *)CHAL_NAMES[which], strlen(CHAL_NAMES[which]));
        if (which) h.update(value[which - 1], 32);
        h.update(bindings[which].data(), bindings[which].size());
        h.final(value[which]);
        return value[which];
    }
};

struct Lin { int kind, is_ntt; long idx_inv; std::vector<int32_t> pairs; };

// Evaluate's SHAKE128 transcript starts with the commit key and every commitment (jindo/prover.go:221-226) -- hundreds of KB
// at ~0.3 GB/s.  None of it depends on the evaluation point, so a worker thread absorbs each commitment as soon as it is
// out while the main thread keeps the GPU busy with the next rounds.
struct ShakeWorker {
    Shake128 xof;
    std::thread th;
    std::mutex m;
    std::condition_variable cv;
    std::deque<std::vector<uint8_t>> q;
    bool closing = false, running = false;
    void start() {
        running = true;
        th = std::thread([this] {
            for (;;) {
                std::unique_lock<std::mutex> lk(m);
                cv.wait(lk, [&] { return !q.empty() || closing; });
                if (q.empty()) return;
                std::vector<uint8_t> v = std::move(q.front());
                q.pop_front();
                lk.unlock();
                xof.absorb(v.data(), v.size());
            }
        });
    }
    void post(const uint8_t *p, size_t n) {
        { std::lock_guard<std::mutex> lk(m); q.emplace_back(p, p + n); }
        cv.notify_one();
    }
    void finish() {
        if (!running) return;
        { std::lock_guard<std::mutex> lk(m); closing = true; }
        cv.notify_one();
        th.join();
        running = false;
    }
    ~ShakeWorker() { finish(); }
};

double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

}  // namespace

struct rs_buckler {
    rs_field *f = nullptr;
    rs_jindo *j = nullptr;
    rs_jindo_params jp;
    int L = 0;
    size_t N = 0, E = 0, wc = 0, pc = 0, arith_max_rank = 0, sum_max_rank = 0, jrank = 0;
    std::vector<rs_term> arith_terms, sum_terms;
    std::vector<uint64_t> arith_coeffs, sum_coeffs;
    std::vector<Lin> lin;
    std::string crs;
    rs_ntt *enc = nullptr, *emb = nullptr, *neg = nullptr;
    // persistent device batches
    rs_batch *vals = nullptr, *ecd = nullptr, *ntt = nullptr, *w_ntt = nullptr, *pw_ntt = nullptr, *work = nullptr, *quo = nullptr, *lv = nullptr,
             *le = nullptr, *qr = nullptr, *rem = nullptr, *qs = nullptr, *rems = nullptr;
    uint64_t *pin = nullptr;        // pinned staging: witnesses and masks in, commitments out
    size_t pin_words = 0, off_lin_mask = 0, off_sum_mask = 0, off_coms = 0;
    uint64_t *tape_pin = nullptr;   // pinned staging of the tapes' host arrays (grown on demand)
    size_t tape_pin_words = 0;
    // self-sampled proofs (rs_buckler_enable_sampling): host samplers on worker threads, one tape slot per Commit
    rs_tape_builder *tb = nullptr;
    uint64_t *heads = nullptr;      // pinned: the first cols*slots coefficients of every committed vector (genFirstLastRow's input)
    void *side[4] = {nullptr, nullptr, nullptr, nullptr};   // the Commits of a round run side by side on these streams
    int n_side = 0;
    rs_batch *vs = nullptr;         // verifier scratch: Encode(linCheckVec) and one Encode(transposed vec) per checker, rank N
    std::vector<uint64_t> r_blind, r_lin_mask, r_sum_mask;
    uint64_t r_lin_sum[32], r_sum_sum[32];
    uint64_t n_mont[32];            // N as a Montgomery element (nttChecker.TransposeTo's scale)
    size_t batch() const { return wc + (arith_terms.empty() ? 0 : 1) + (lin.empty() ? 0 : 4) + (sum_terms.empty() ? 0 : 4); }
};

#define RS_B(call)                      \
    do {                                \
        int _rc = (call);               \
        if (_rc != RS_OK) return _rc;   \
    } while (0)

#pragma GCC visibility push(default)
extern "C" {

void rs_sha256(const uint8_t *data, size_t len, uint8_t out32[32]) {
    Sha256 h;
    h.update(data, len);
    h.final(out32);
}

void rs_shake128(const uint8_t *data, size_t len, uint8_t *out, size_t out_len) {
    Shake128 x;
    x.absorb(data, len);
    x.squeeze(out, out_len);
}

void rs_buckler_destroy(rs_buckler *b) {
    if (!b) return;
    for (rs_batch *x : {b->vals, b->ecd, b->ntt, b->w_ntt, b->pw_ntt, b->work, b->quo, b->lv, b->le, b->qr, b->rem, b->qs, b->rems, b->vs})
        if (x) rs_batch_free(x);
    for (rs_ntt *t : {b->enc, b->emb, b->neg})
        if (t) rs_ntt_destroy(t);
    for (void *st : b->side) rs_stream_destroy(st);
    if (b->pin) rs_pinned_free(b->pin);
    if (b->tape_pin) rs_pinned_free(b->tape_pin);
    if (b->heads) rs_pinned_free(b->heads);
    if (b->tb) rs_tape_builder_destroy(b->tb);
    delete b;
}

size_t rs_buckler_batch(const rs_buckler *b) { return b ? b->batch() : 0; }

int rs_buckler_create(rs_field *f, rs_jindo *j, const rs_buckler_circuit *c, const uint8_t *crs, size_t crs_len, rs_buckler **out) {
    if (!f || !j || !c || !out || (!crs && crs_len)) return rs::set_error(RS_ERR_ARG, "rs_buckler_create: null argument");
    auto log2_exact = [](size_t v) { int l = 0; while (((size_t)1 << l) < v) l++; return ((size_t)1 << l) == v ? l : -1; };
    const int lN = log2_exact(c->rank), lE = log2_exact(c->emb_rank);
    if (lN < 0 || lE < 0 || c->emb_rank < 2 * c->rank) return rs::set_error(RS_ERR_SHAPE, "rs_buckler_create: rank / emb_rank must be powers of two, emb_rank >= 2 rank");
    if (c->w_cnt == 0) return rs::set_error(RS_ERR_SHAPE, "rs_buckler_create: circuit has no witness");
    rs_buckler *b = new rs_buckler();
    b->f = f; b->j = j; b->L = rs_field_limbs(f);
    b->N = c->rank; b->E = c->emb_rank; b->wc = c->w_cnt; b->pc = c->pw_cnt;
    b->arith_max_rank = c->arith_max_rank; b->sum_max_rank = c->sum_max_rank;
    b->arith_terms.assign(c->arith_terms, c->arith_terms + c->n_arith_terms);
    b->arith_coeffs.assign(c->arith_coeffs, c->arith_coeffs + c->n_arith_coeffs * b->L);
    b->sum_terms.assign(c->sum_terms, c->sum_terms + c->n_sum_terms);
    b->sum_coeffs.assign(c->sum_coeffs, c->sum_coeffs + c->n_sum_coeffs * b->L);
    for (size_t i = 0; i < c->n_lin; i++) {
        Lin l;
        l.kind = c->lin[i].kind; l.is_ntt = c->lin[i].is_ntt; l.idx_inv = (long)c->lin[i].idx_inv;
        l.pairs.assign(c->lin[i].pairs, c->lin[i].pairs + 2 * c->lin[i].n_pairs);
        b->lin.push_back(l);
    }
    b->crs.assign((const char *)crs, crs_len);
    int rc = rs_jindo_get_params(j, &b->jp);
    const size_t N = b->N, E = b->E;
    b->jrank = (size_t)b->jp.cols * (b->jp.rows - 1) * b->jp.slots;
    auto term_ok = [&](const std::vector<rs_term> &ts, size_t n_coeffs) {
        for (const rs_term &t : ts) {
            if (t.coeff < 0 || (size_t)t.coeff >= n_coeffs || t.pw >= (int)b->pc || t.n_w < 0 || t.n_w > RS_TERM_MAX_W) return false;
            for (int k = 0; k < t.n_w; k++) if (t.w[k] < 0 || (size_t)t.w[k] >= b->wc) return false;
        }
        return true;
    };
    if (rc == RS_OK && (!term_ok(b->arith_terms, c->n_arith_coeffs) || !term_ok(b->sum_terms, c->n_sum_coeffs)))
        rc = rs::set_error(RS_ERR_SHAPE, "rs_buckler_create: constraint term refers to a witness / coefficient out of range");
    for (const Lin &l : b->lin)
        for (int32_t w : l.pairs) if (rc == RS_OK && (w < 0 || (size_t)w >= b->wc)) rc = rs::set_error(RS_ERR_SHAPE, "rs_buckler_create: linear constraint refers to a witness out of range");
    if (rc == RS_OK && (N + 1 > b->jrank || 2 * N > b->jrank || (!b->arith_terms.empty() && (b->arith_max_rank < N + 1 || b->arith_max_rank - N > E)) ||
                        (!b->sum_terms.empty() && (b->sum_max_rank < N + 1 || b->sum_max_rank > E))))
        rc = rs::set_error(RS_ERR_SHAPE, "rs_buckler_create: committed lengths exceed the Jindo rank / emb_rank");
    if (rc == RS_OK) rc = rs_ntt_create(f, lN, 0, &b->enc);        // Encoder.ntt (encoder.go:15-20)
    if (rc == RS_OK) rc = rs_ntt_create(f, lE, 0, &b->emb);        // polyEval (compile.go:186)
    if (rc == RS_OK && !b->lin.empty()) rc = rs_ntt_create(f, lN, 1, &b->neg);   // nttChecker (linear.go:26-33)
    const size_t np = b->wc + b->pc, rr = b->jrank > N ? b->jrank : N;
    if (rc == RS_OK) rc = rs_batch_alloc(f, np, N, &b->vals);
    if (rc == RS_OK) rc = rs_batch_alloc(f, np, E, &b->ecd);
    if (rc == RS_OK) rc = rs_batch_alloc(f, np, E, &b->ntt);
    if (rc == RS_OK) rc = rs_batch_wrap(f, rs_batch_device_ptr(b->ntt), b->wc, E, &b->w_ntt);
    if (rc == RS_OK) rc = rs_batch_alloc(f, b->pc ? b->pc : 1, E, &b->pw_ntt);
    if (rc == RS_OK) rc = rs_batch_alloc(f, 8, E, &b->work);       // 0 linCheckMask, 1 eval, 4 term, 5 sumCheckMask
    if (rc == RS_OK) rc = rs_batch_alloc(f, 2, E, &b->quo);
    if (rc == RS_OK) rc = rs_batch_alloc(f, 2, N, &b->lv);
    if (rc == RS_OK) rc = rs_batch_alloc(f, 2, E, &b->le);
    if (rc == RS_OK) rc = rs_batch_alloc(f, 2, E, &b->qr);
    if (rc == RS_OK) rc = rs_batch_alloc(f, 2, rr, &b->rem);
    if (rc == RS_OK) rc = rs_batch_alloc(f, 2, E, &b->qs);
    if (rc == RS_OK) rc = rs_batch_alloc(f, 2, rr, &b->rems);
    if (rc == RS_OK) {
        uint64_t n_can[32] = {0};
        n_can[0] = N;
        rc = rs_field_to_mont(f, b->n_mont, n_can);
    }
    {
        const char *e = getenv("RS_BUCKLER_STREAMS");              // 0: every Commit on the caller's stream
        const int want = e ? atoi(e) : 4;
        for (int k = 0; rc == RS_OK && k < want && k < 4; k++) { rc = rs_stream_create(&b->side[k]); if (rc == RS_OK) b->n_side = k + 1; }
    }
    if (rc == RS_OK) {
        b->off_lin_mask = np * N * b->L;
        b->off_sum_mask = b->off_lin_mask + 2 * N * b->L;
        b->off_coms = b->off_sum_mask + b->sum_max_rank * b->L;
        b->pin_words = b->off_coms + b->batch() * (size_t)b->jp.out_msis * b->jp.n_qout * b->jp.d;
        b->pin = (uint64_t *)rs_pinned_alloc(b->pin_words * 8);
        if (!b->pin) rc = rs::set_error(RS_ERR_CUDA, "rs_buckler_create: pinned allocation failed");
    }
    if (rc != RS_OK) { rs_buckler_destroy(b); return rc; }
    *out = b;
    return RS_OK;
}

int rs_buckler_enable_sampling(rs_buckler *b, const rs_jindo_noise *noise, const uint64_t *q_le, int n_threads) {
    if (!b || !noise || !q_le || n_threads < 1) return rs::set_error(RS_ERR_ARG, "rs_buckler_enable_sampling: bad argument");
    if (b->tb) { rs_tape_builder_destroy(b->tb); b->tb = nullptr; }
    int rc = rs_tape_builder_create_slots(&b->jp, noise, q_le, b->L, n_threads, (int)b->batch(), nullptr, &b->tb);
    if (rc != RS_OK) return rs::set_error(rc, "rs_buckler_enable_sampling: tape builder creation failed");
    if (!b->heads) {
        b->heads = (uint64_t *)rs_pinned_alloc(b->batch() * (size_t)b->jp.cols * b->jp.slots * b->L * 8);
        if (!b->heads) return rs::set_error(RS_ERR_CUDA, "rs_buckler_enable_sampling: pinned allocation failed");
    }
    return RS_OK;
}

int rs_buckler_set_device_cosac(rs_buckler *b, int on) {
    if (!b || !b->tb) return rs::set_error(RS_ERR_STATE, "rs_buckler_set_device_cosac: call rs_buckler_enable_sampling first");
    return rs_tape_builder_set_device_cosac(b->tb, on);
}

int rs_buckler_prove(rs_buckler *b, const uint64_t *pw, const uint64_t *w, const rs_buckler_rand *rd, rs_buckler_proof *pf, double *phase_ms) {
    if (!b || !w || !pf || (b->pc && !pw)) return rs::set_error(RS_ERR_ARG, "rs_buckler_prove: null argument");
    const bool sampling = (rd == nullptr);       // the prover draws its own randomness, like the reference's Prove
    rs_buckler_rand own;
    if (sampling) {
        if (!b->tb) return rs::set_error(RS_ERR_STATE, "rs_buckler_prove: no randomness given and rs_buckler_enable_sampling was not called");
        memset(&own, 0, sizeof(own));
        rd = &own;
    }
    const size_t N = b->N, E = b->E, wc = b->wc, pc = b->pc, n_com = b->batch();
    const int L = b->L;
    const rs_jindo_params &jp = b->jp;
    const bool has_arith = !b->arith_terms.empty(), has_lin = !b->lin.empty(), has_sum = !b->sum_terms.empty();
    if (!pf->witness || !pf->evals || !pf->partial || !pf->partial_mask || !pf->encode || !pf->mlwe || !pf->incommit)
        return rs::set_error(RS_ERR_ARG, "rs_buckler_prove: proof buffers missing");
    if (sampling) {      // RandEncode's blinding elements and the two sumCheckMask polynomials (encoder.go:50-54, prover.go:381-397)
        b->r_blind.resize(b->wc * L);
        RS_B(rs_tape_builder_random_elements(b->tb, b->r_blind.data(), b->wc));
        own.blind = b->r_blind.data();
        if (has_lin) {
            b->r_lin_mask.resize(2 * b->N * L);
            RS_B(rs_tape_builder_sum_check_mask(b->tb, b->N, 2 * b->N, b->r_lin_mask.data(), b->r_lin_sum));
            own.lin_mask = b->r_lin_mask.data(); own.lin_mask_sum = b->r_lin_sum;
        }
        if (has_sum) {
            b->r_sum_mask.resize(b->sum_max_rank * L);
            RS_B(rs_tape_builder_sum_check_mask(b->tb, b->N, b->sum_max_rank, b->r_sum_mask.data(), b->r_sum_sum));
            own.sum_mask = b->r_sum_mask.data(); own.sum_mask_sum = b->r_sum_sum;
        }
    } else {
        if (rd->n_tapes != n_com || !rd->tapes || !rd->blind) return rs::set_error(RS_ERR_ARG, "rs_buckler_prove: one tape per commitment and the blinding elements are required");
        if ((has_lin && (!rd->lin_mask || !rd->lin_mask_sum)) || (has_sum && (!rd->sum_mask || !rd->sum_mask_sum)))
            return rs::set_error(RS_ERR_ARG, "rs_buckler_prove: mask polynomial missing");
    }
    double t_phase[16] = {0};
    double t_last = now_ms();
    const double t_start = t_last;
    auto mark = [&](int k) {
        if (!phase_ms) return;
        void *m = rs_get_stream();
        for (int i = 0; i < b->n_side; i++) rs_stream_wait(m, b->side[i]);
        rs_sync();
        const double t = now_ms();
        t_phase[k] += t - t_last;
        t_last = t;
    };
    const size_t com_words = (size_t)jp.out_msis * jp.n_qout * jp.d;
    uint64_t *coms_pin = b->pin + b->off_coms;

    // ---- tapes: the Commits are enqueued without waiting for each other, so their host arrays must be pinned (a copy from
    // pageable memory would make every cudaMemcpyAsync wait for the stream).  Pageable tapes are staged here first.
    std::vector<rs_jindo_tape> tapes(n_com);
    if (!sampling) tapes.assign(rd->tapes, rd->tapes + n_com);
    if (!sampling && !rd->tapes_pinned) {
        const size_t cs = (size_t)jp.cols * jp.slots, nmd = (size_t)(jp.in_msis + jp.mlwe) * jp.d;
        auto tape_words = [&](const rs_jindo_tape &t) {
            return 2 * cs * L + t.n_mask_rows * jp.slots * L + (t.twin_u ? t.n_twin_u : 0) + t.n_cosac_polys * jp.d +
                   (t.mlwe_mask_only ? nmd : (size_t)(jp.cols + 1) * nmd);
        };
        size_t total = 0;
        for (const rs_jindo_tape &t : tapes) if (!t.on_device) total += tape_words(t);
        if (total > b->tape_pin_words) {
            rs_sync();
            if (b->tape_pin) rs_pinned_free(b->tape_pin);
            b->tape_pin_words = total + total / 8;
            b->tape_pin = (uint64_t *)rs_pinned_alloc(b->tape_pin_words * 8);
            if (!b->tape_pin) { b->tape_pin_words = 0; return rs::set_error(RS_ERR_CUDA, "rs_buckler_prove: pinned allocation failed"); }
        }
        uint64_t *o = b->tape_pin;
        auto stage = [&](const void *src, size_t words) -> uint64_t * {
            uint64_t *dst = o;
            memcpy(dst, src, words * 8);
            o += words;
            return dst;
        };
        for (rs_jindo_tape &t : tapes) {
            if (t.on_device) continue;
            if (!t.first_row || !t.last_row || !t.mask_rows || !t.cosac_g || !t.mlwe) return rs::set_error(RS_ERR_ARG, "rs_buckler_prove: incomplete tape");
            t.first_row = stage(t.first_row, cs * L);
            t.last_row = stage(t.last_row, cs * L);
            t.mask_rows = stage(t.mask_rows, t.n_mask_rows * jp.slots * L);
            if (t.twin_u) t.twin_u = stage(t.twin_u, t.n_twin_u);
            t.cosac_g = (const int64_t *)stage(t.cosac_g, t.n_cosac_polys * jp.d);
            t.mlwe = (const int64_t *)stage(t.mlwe, t.mlwe_mask_only ? nmd : (size_t)(jp.cols + 1) * nmd);
        }
    }
    mark(6);

    std::vector<uint8_t> raw;
    auto raw_of = [&](size_t idx) -> int {
        const long long n = rs_jindo_commitment_write_raw(b->j, coms_pin + idx * com_words, nullptr, 0);
        if (n < 0) return (int)n;
        raw.resize((size_t)n);
        const long long got = rs_jindo_commitment_write_raw(b->j, coms_pin + idx * com_words, raw.data(), raw.size());
        return got == n ? RS_OK : (int)(got < 0 ? got : RS_ERR_ARG);
    };
    struct View { rs_batch *bt; size_t poly, len; };
    std::vector<View> views(n_com);
    std::vector<rs_opening *> opens(n_com, nullptr);
    struct Cleanup {
        std::vector<rs_opening *> &o; rs_opening *extra = nullptr;
        rs_buckler *bb; void *m;
        ~Cleanup() {
            rs_set_stream(m);
            for (int k = 0; k < bb->n_side; k++) rs_stream_wait(m, bb->side[k]);
            rs_sync();
            for (rs_opening *x : o) if (x) rs_opening_free(x);
            if (extra) rs_opening_free(extra);
        }
    } cleanup{opens, nullptr, b, rs_get_stream()};
    Transcript tr;
    ShakeWorker shake;
    shake.xof.absorb((const uint8_t *)b->crs.data(), b->crs.size());             // ck.WriteRawTo (jindo/prover.go:222)
    shake.start();
    size_t n_issued = 0, n_bound = 0;
    const size_t head_cap = (size_t)jp.cols * jp.slots;
    std::vector<size_t> pending;
    std::vector<int> bind_of(n_com);
    // one Commit, enqueued only, on the next side stream: it starts once everything enqueued on the caller's stream so far (the
    // committed vector) has run, and runs beside the other Commits of its round and the next round's arithmetic
    void *const main_stream = rs_get_stream();
    struct StreamGuard { void *m; ~StreamGuard() { rs_set_stream(m); } } stream_guard{main_stream};
    auto issue = [&](size_t idx) -> int {
        void *st = b->n_side ? b->side[idx % (size_t)b->n_side] : main_stream;
        if (st != main_stream) { RS_B(rs_stream_wait(st, main_stream)); rs_set_stream(st); }
        const int rc = rs_jindo_commit_async(b->j, views[idx].bt, views[idx].poly, views[idx].len, &tapes[idx], &opens[idx], coms_pin + idx * com_words);
        rs_set_stream(main_stream);
        return rc;
    };
    auto join_sides = [&]() -> int {
        for (int k = 0; k < b->n_side; k++) RS_B(rs_stream_wait(main_stream, b->side[k]));
        return RS_OK;
    };
    // polyProver.Commit, enqueued only; the commitment lands in pinned memory
    auto commit = [&](size_t idx, rs_batch *bt, size_t poly, size_t len, int bind) -> int {
        if (idx != n_issued) return rs::set_error(RS_ERR_ARG, "rs_buckler_prove: commit order");
        views[idx] = View{bt, poly, len};
        bind_of[idx] = bind;
        n_issued++;
        if (!sampling) return issue(idx);
        // genFirstLastRow (jindo/prover.go:65-86) and the first rows' COSAC centres need the head of the vector on the host
        const size_t head = len < head_cap ? len : head_cap;
        rs_batch *view = nullptr;
        RS_B(rs_batch_wrap(b->f, (uint8_t *)rs_batch_device_ptr(bt) + poly * rs_batch_rank(bt) * L * 8, 1, head, &view));
        const int rc = rs_batch_download_async(view, 0, 1, b->heads + idx * head_cap * L);
        rs_batch_free(view);
        pending.push_back(idx);
        return rc;
    };
    // self-sampled mode: sample the pending Commits' tapes on the worker threads and enqueue each Commit as its tape completes
    auto flush = [&]() -> int {
        if (pending.empty()) return RS_OK;
        RS_B(rs_sync());
        for (size_t i : pending)
            if (rs_tape_builder_start(b->tb, b->heads + i * head_cap * L, views[i].len, (int)i) != RS_OK)
                return rs::set_error(RS_ERR_ARG, "rs_buckler_prove: tape builder start failed");
        for (size_t i : pending) {
            if (rs_tape_builder_wait(b->tb, (int)i, &tapes[i], nullptr) != RS_OK) return rs::set_error(RS_ERR_ARG, "rs_buckler_prove: tape builder failed");
            RS_B(issue(i));
        }
        pending.clear();
        return RS_OK;
    };
    // wait for the commitments issued so far, bind them into the fiat-shamir transcript in order (coms[i].WriteRawTo,
    // oracle.Bind) and hand their bytes to the SHAKE worker; `extra` runs right after commitment `extra_after` is bound
    uint8_t el[8 * 32];
    auto bind_ready = [&](long extra_after_a, const uint64_t *sum_a, long extra_after_b, const uint64_t *sum_b) -> int {
        RS_B(flush());
        RS_B(join_sides());
        RS_B(rs_sync());
        for (; n_bound < n_issued; n_bound++) {
            RS_B(raw_of(n_bound));
            tr.bind(bind_of[n_bound], raw.data(), raw.size());
            shake.post(raw.data(), raw.size());
            memcpy(pf->witness + n_bound * com_words, coms_pin + n_bound * com_words, com_words * 8);
            for (int k = 0; k < 2; k++) {
                const long after = k ? extra_after_b : extra_after_a;
                const uint64_t *sum = k ? sum_b : sum_a;
                if (after == (long)n_bound && sum) {                                  // oracle.Bind("arithBatchConst", maskSum.Marshal())
                    RS_B(rs_field_marshal(b->f, sum, el));
                    tr.bind(ARITH_BC, el, 8 * L);
                }
            }
        }
        return RS_OK;
    };

    // ---- Encode + NTT of the public witnesses and the (first-round) witnesses (prover.go:123-160)
    memcpy(b->pin, w, wc * N * L * 8);
    if (pc) memcpy(b->pin + wc * N * L, pw, pc * N * L * 8);
    RS_B(rs_batch_upload(b->vals, 0, wc + pc, b->pin));
    RS_B(rs_ntt_inverse_embed(b->enc, b->ecd, 0, b->vals, 0, wc + pc, 1));
    RS_B(rs_rand_encode_blind(b->ecd, 0, wc, N, rd->blind));
    RS_B(rs_ntt_forward(b->emb, b->ntt, 0, b->ecd, 0, wc + pc));
    for (size_t k = 0; k < pc; k++) RS_B(rs_batch_copy(b->pw_ntt, k, 0, b->ntt, wc + k, 0, E));
    mark(0);
    for (size_t i = 0; i < wc; i++) RS_B(commit(i, b->ecd, i, N + 1, PROJ));
    size_t idx = wc;
    long lin_mask_idx = -1, sum_mask_idx = -1;
    auto upload_mask = [&](size_t work_poly, const uint64_t *mask, size_t words_off, size_t len) -> int {
        RS_B(rs_batch_zero(b->work, work_poly, 0, E));
        memcpy(b->pin + words_off, mask, len * L * 8);
        rs_batch *view = nullptr;      // the mask occupies the first `len` coefficients of an emb_rank polynomial
        RS_B(rs_batch_wrap(b->f, (uint8_t *)rs_batch_device_ptr(b->work) + work_poly * E * L * 8, 1, len, &view));
        const int rc = rs_batch_upload(view, 0, 1, b->pin + words_off);
        rs_batch_free(view);
        return rc;
    };
    if (has_lin) {      // linCheckMask (prover.go:209-224)
        RS_B(upload_mask(0, rd->lin_mask, b->off_lin_mask, 2 * N));
        lin_mask_idx = (long)idx;
        RS_B(commit(idx++, b->work, 0, 2 * N, ARITH_BC));
        if (pf->lin_mask_sum) memcpy(pf->lin_mask_sum, rd->lin_mask_sum, 8 * L);
    }
    if (has_sum) {      // sumCheckMask (prover.go:226-241)
        RS_B(upload_mask(5, rd->sum_mask, b->off_sum_mask, b->sum_max_rank));
        sum_mask_idx = (long)idx;
        RS_B(commit(idx++, b->work, 5, b->sum_max_rank, ARITH_BC));
        if (pf->sum_mask_sum) memcpy(pf->sum_mask_sum, rd->sum_mask_sum, 8 * L);
    }
    // projConst only feeds the projection checker; its digest is needed as the prefix of arithBatchConst.  Bindings are per
    // challenge name, so binding the masks' commitments before computing projConst changes nothing.
    RS_B(bind_ready(lin_mask_idx, rd->lin_mask_sum, sum_mask_idx, rd->sum_mask_sum));
    tr.compute(PROJ);
    mark(1);
    uint64_t bc[32], lc[32];
    RS_B(rs_field_set_bytes(b->f, tr.compute(ARITH_BC), 32, bc));
    // From here to evalPoint no challenge depends on a new commitment (linCheckBatchConst, linCheckConst and
    // sumCheckBatchConst have no bindings, prover.go:262-304): every remaining round is enqueued back to back.
    if (has_arith) {    // arithCheck (prover.go:399-404)
        RS_B(rs_eval_circuit(b->work, 1, b->pw_ntt, b->w_ntt, b->arith_terms.data(), b->arith_terms.size(), b->arith_coeffs.data(),
                             b->arith_coeffs.size() / L, bc));
        RS_B(rs_ntt_inverse(b->emb, b->work, 1, b->work, 1, 1));
        RS_B(rs_quorem_vanishing(b->quo, 0, b->quo, 1, b->work, 1, N));
        mark(2);
        RS_B(commit(idx++, b->quo, 0, b->arith_max_rank - N, EVAL_PT));
        mark(1);
    }
    RS_B(rs_field_set_bytes(b->f, tr.compute(LIN_BC), 32, bc));
    RS_B(rs_field_set_bytes(b->f, tr.compute(LIN_C), 32, lc));
    // quo / remLo / remHi of a sum-check style round (prover.go:443-456, :473-484)
    auto split_rem = [&](rs_batch *q, rs_batch *r) -> int {
        const size_t rr = rs_batch_rank(r);
        RS_B(rs_batch_zero(r, 0, 0, rr));
        RS_B(rs_batch_zero(r, 1, 0, rr));
        RS_B(rs_batch_copy(r, 0, 0, q, 1, 1, N - 1));
        RS_B(rs_batch_copy(r, 1, b->jrank - (N - 1), q, 1, 1, N - 1));
        return RS_OK;
    };
    if (has_lin) {      // linCheck (prover.go:406-459)
        RS_B(rs_vec_powers(b->lv, 0, N, lc));
        RS_B(rs_ntt_inverse_embed(b->enc, b->le, 0, b->lv, 0, 1, 1));
        RS_B(rs_ntt_forward(b->emb, b->le, 0, b->le, 0, 1));                      // linCheckEcd
        RS_B(rs_batch_zero(b->work, 1, 0, E));                                    // eval
        for (const Lin &l : b->lin) {
            if (l.kind == 0) {                                                    // nttChecker.TransposeTo (linear.go:38-43)
                RS_B(rs_reverse_scale(b->lv, 1, b->lv, 0, b->n_mont));
                RS_B(rs_ntt_inverse(b->neg, b->lv, 1, b->lv, 1, 1));
            } else {                                                              // autChecker.TransposeTo (linear.go:69-73)
                RS_B(rs_aut(b->lv, 1, b->lv, 0, 1, l.idx_inv, l.is_ntt));
            }
            RS_B(rs_ntt_inverse_embed(b->enc, b->le, 1, b->lv, 1, 1, 1));
            RS_B(rs_ntt_forward(b->emb, b->le, 1, b->le, 1, 1));                  // linCheckEcdTr
            for (size_t k = 0; k < l.pairs.size(); k += 2) {
                const size_t w_out = l.pairs[k], w_in = l.pairs[k + 1];
                RS_B(rs_vec_mul(b->work, 4, b->le, 1, b->w_ntt, w_in, 1));
                RS_B(rs_vec_mul_sub(b->work, 4, b->le, 0, b->w_ntt, w_out, 1));
                RS_B(rs_vec_scalar_mul(b->work, 1, b->work, 1, bc, 1));
                RS_B(rs_vec_add(b->work, 1, b->work, 1, b->work, 4, 1));
            }
        }
        RS_B(rs_vec_scalar_mul(b->work, 1, b->work, 1, bc, 1));
        RS_B(rs_ntt_inverse(b->emb, b->work, 1, b->work, 1, 1));
        RS_B(rs_vec_add(b->work, 1, b->work, 1, b->work, 0, 1));                  // + linCheckMask
        RS_B(rs_quorem_vanishing(b->qr, 0, b->qr, 1, b->work, 1, N));
        RS_B(split_rem(b->qr, b->rem));
        mark(3);
        RS_B(commit(idx, b->qr, 0, N, EVAL_PT));
        RS_B(commit(idx + 1, b->rem, 0, N - 1, EVAL_PT));
        RS_B(commit(idx + 2, b->rem, 1, b->jrank, EVAL_PT));
        idx += 3;
        mark(1);
    }
    RS_B(rs_field_set_bytes(b->f, tr.compute(SUM_BC), 32, bc));
    if (has_sum) {      // sumCheck (prover.go:461-485)
        RS_B(rs_eval_circuit(b->work, 1, b->pw_ntt, b->w_ntt, b->sum_terms.data(), b->sum_terms.size(), b->sum_coeffs.data(),
                             b->sum_coeffs.size() / L, bc));
        RS_B(rs_vec_scalar_mul(b->work, 1, b->work, 1, bc, 1));
        RS_B(rs_ntt_inverse(b->emb, b->work, 1, b->work, 1, 1));
        RS_B(rs_vec_add(b->work, 1, b->work, 1, b->work, 5, 1));                  // + sumCheckMask
        RS_B(rs_quorem_vanishing(b->qs, 0, b->qs, 1, b->work, 1, N));
        RS_B(split_rem(b->qs, b->rems));
        mark(4);
        RS_B(commit(idx, b->qs, 0, b->sum_max_rank - N, EVAL_PT));
        RS_B(commit(idx + 1, b->rems, 0, N - 1, EVAL_PT));
        RS_B(commit(idx + 2, b->rems, 1, b->jrank, EVAL_PT));
        idx += 3;
        mark(1);
    }
    RS_B(bind_ready(-1, nullptr, -1, nullptr));
    mark(1);
    uint64_t x[32];
    RS_B(rs_field_set_bytes(b->f, tr.compute(EVAL_PT), 32, x));
    if (pf->eval_point) memcpy(pf->eval_point, x, 8 * L);

    // ---- polyProver.Evaluate (jindo/prover.go:205-324)
    shake.finish();                                                               // ck and every commitment absorbed
    Shake128 &xof = shake.xof;
    RS_B(rs_field_marshal(b->f, x, el));
    xof.absorb(el, 8 * L);
    rs_opening *ob = opens[0];
    if (n_com > 1) {
        std::vector<uint8_t> bb(16 * n_com);
        Shake128 fork = xof;                                                      // Read, Reset, re-absorb the same prefix (prover.go:232-248)
        fork.squeeze(bb.data(), bb.size());
        xof.absorb(bb.data(), bb.size());
        RS_B(rs_opening_combine_bytes(b->j, opens.data(), n_com, bb.data(), &cleanup.extra));
        ob = cleanup.extra;
    }
    mark(8);
    const size_t pq = (size_t)jp.n_q * jp.d;
    std::vector<uint64_t> part((size_t)(jp.cols + 1) * pq);
    RS_B(rs_jindo_partial_at(b->j, ob, x, part.data()));                          // leftVec(x) encoded on the device
    mark(9);
    memcpy(pf->partial, part.data(), (size_t)jp.cols * pq * 8);
    memcpy(pf->partial_mask, part.data() + (size_t)jp.cols * pq, pq * 8);
    {
        const long long n = rs_jindo_poly_write_to(part.data(), jp.n_q, jp.d, nullptr, 0);
        if (n < 0) return (int)n;
        raw.resize((size_t)n);
        for (int i = 0; i <= jp.cols; i++) {                                      // Partial[i].WriteTo, PartialMask.WriteTo
            if (rs_jindo_poly_write_to(part.data() + (size_t)i * pq, jp.n_q, jp.d, raw.data(), raw.size()) != n)
                return rs::set_error(RS_ERR_ARG, "rs_buckler_prove: poly_write_to failed");
            xof.absorb(raw.data(), raw.size());
        }
    }
    std::vector<uint8_t> cs(16 * (size_t)jp.cols);
    xof.squeeze(cs.data(), cs.size());
    mark(10);
    RS_B(rs_jindo_respond_bytes(b->j, ob, cs.data(), pf->encode, pf->mlwe));
    RS_B(rs_opening_download_all(ob, 2, pf->incommit));
    mark(11);
    {                                                                             // Poly.Evaluate of every committed vector
        std::vector<const rs_batch *> bs(n_com);
        std::vector<size_t> ps(n_com), ls(n_com);
        for (size_t i = 0; i < n_com; i++) { bs[i] = views[i].bt; ps[i] = views[i].poly; ls[i] = views[i].len; }
        RS_B(rs_poly_evaluate_many(bs.data(), ps.data(), ls.data(), n_com, x, pf->evals));
    }
    mark(12);
    if (phase_ms) {
        t_phase[5] = t_phase[8] + t_phase[9] + t_phase[10] + t_phase[11] + t_phase[12];
        t_phase[7] = now_ms() - t_start;
        memcpy(phase_ms, t_phase, sizeof(t_phase));
    }
    return RS_OK;
}

// Verifier.Verify (buckler/verifier.go:29-315).  The transcripts are recomputed from the proof, the Jindo evaluation proof is
// checked on the device (rs_jindo_verify), Encode(.).Evaluate(x) of the public witnesses and of the linear-check vectors runs
// on the device (INTT + one batched evaluation), and the closing equations are a few dozen field operations on the host.
// *result: 0 accepted; 1 polyVerifier.Verify failed, 2 arithmetic check, 3 / 5 remHi != x^k remLo (linear / sum check),
// 4 linear check, 6 sum check.
int rs_buckler_verify(rs_buckler *b, const uint64_t *pw, const rs_buckler_proof *pf, double in_com_dcmp_two_nm, double res_two_nm, int *result) {
    if (!b || !pf || !result || (b->pc && !pw)) return rs::set_error(RS_ERR_ARG, "rs_buckler_verify: null argument");
    if (!pf->witness || !pf->evals || !pf->partial || !pf->partial_mask || !pf->encode || !pf->mlwe || !pf->incommit)
        return rs::set_error(RS_ERR_ARG, "rs_buckler_verify: proof buffers missing");
    const size_t N = b->N, wc = b->wc, pc = b->pc, n_com = b->batch();
    const int L = b->L;
    const rs_jindo_params &jp = b->jp;
    const bool has_arith = !b->arith_terms.empty(), has_lin = !b->lin.empty(), has_sum = !b->sum_terms.empty();
    if ((has_lin && !pf->lin_mask_sum) || (has_sum && !pf->sum_mask_sum)) return rs::set_error(RS_ERR_ARG, "rs_buckler_verify: mask sum missing");
    *result = -1;
    const size_t com_words = (size_t)jp.out_msis * jp.n_qout * jp.d;
    std::vector<uint8_t> raw;
    auto raw_of = [&](size_t idx) -> int {
        const long long n = rs_jindo_commitment_write_raw(b->j, pf->witness + idx * com_words, nullptr, 0);
        if (n < 0) return (int)n;
        raw.resize((size_t)n);
        const long long got = rs_jindo_commitment_write_raw(b->j, pf->witness + idx * com_words, raw.data(), raw.size());
        return got == n ? RS_OK : (int)(got < 0 ? got : RS_ERR_ARG);
    };
    // ---- the fiat-shamir transcript, bound as Prove binds it (verifier.go:70-160), and Evaluate's SHAKE prefix
    Transcript tr;
    Shake128 xof;
    xof.absorb((const uint8_t *)b->crs.data(), b->crs.size());
    uint8_t el[8 * 32];
    size_t idx = wc;
    const size_t lin_mask_idx = idx;
    if (has_lin) idx++;
    const size_t sum_mask_idx = idx;
    if (has_sum) idx++;
    const size_t first_late = idx;
    for (size_t i = 0; i < n_com; i++) {
        RS_B(raw_of(i));
        xof.absorb(raw.data(), raw.size());
        const int which = i < wc ? PROJ : (i < first_late ? ARITH_BC : EVAL_PT);
        tr.bind(which, raw.data(), raw.size());
        if ((has_lin && i == lin_mask_idx) || (has_sum && i == sum_mask_idx)) {
            RS_B(rs_field_marshal(b->f, (has_lin && i == lin_mask_idx) ? pf->lin_mask_sum : pf->sum_mask_sum, el));
            tr.bind(ARITH_BC, el, 8 * L);
        }
    }
    uint64_t arith_bc[32], lin_bc[32], lin_c[32], sum_bc[32], x[32];
    tr.compute(PROJ);
    RS_B(rs_field_set_bytes(b->f, tr.compute(ARITH_BC), 32, arith_bc));
    RS_B(rs_field_set_bytes(b->f, tr.compute(LIN_BC), 32, lin_bc));
    RS_B(rs_field_set_bytes(b->f, tr.compute(LIN_C), 32, lin_c));
    RS_B(rs_field_set_bytes(b->f, tr.compute(SUM_BC), 32, sum_bc));
    RS_B(rs_field_set_bytes(b->f, tr.compute(EVAL_PT), 32, x));
    // ---- polyVerifier.Verify (jindo/verifier.go:50-282): SHAKE bytes here, arithmetic on the device
    RS_B(rs_field_marshal(b->f, x, el));
    xof.absorb(el, 8 * L);
    std::vector<uint8_t> bb(16 * n_com);
    if (n_com > 1) {
        Shake128 fork = xof;
        fork.squeeze(bb.data(), bb.size());
        xof.absorb(bb.data(), bb.size());
    }
    {
        const size_t pq = (size_t)jp.n_q * jp.d;
        const long long n = rs_jindo_poly_write_to(pf->partial, jp.n_q, jp.d, nullptr, 0);
        if (n < 0) return (int)n;
        raw.resize((size_t)n);
        for (int i = 0; i <= jp.cols; i++) {
            const uint64_t *poly = i < jp.cols ? pf->partial + (size_t)i * pq : pf->partial_mask;
            if (rs_jindo_poly_write_to(poly, jp.n_q, jp.d, raw.data(), raw.size()) != n) return rs::set_error(RS_ERR_ARG, "rs_buckler_verify: poly_write_to failed");
            xof.absorb(raw.data(), raw.size());
        }
    }
    std::vector<uint8_t> cs(16 * (size_t)jp.cols);
    xof.squeeze(cs.data(), cs.size());
    rs_jindo_proof jpf{pf->partial, pf->partial_mask, pf->encode, pf->mlwe, pf->incommit};
    int checks = 0;
    RS_B(rs_jindo_verify(b->j, x, pf->witness, (int)n_com, n_com > 1 ? bb.data() : nullptr, cs.data(), pf->evals, &jpf, in_com_dcmp_two_nm, res_two_nm, &checks));
    if (checks != 15) { *result = 1; return RS_OK; }
    // ---- Encode(v).Evaluate(x) for the public witnesses, linCheckVec and its transposes (verifier.go:191-193, :244-262)
    const size_t n_lin = b->lin.size();
    if (has_lin && !b->vs) RS_B(rs_batch_alloc(b->f, 1 + n_lin, N, &b->vs));
    std::vector<const rs_batch *> bs;
    std::vector<size_t> ps, ls;
    if (pc) {
        memcpy(b->pin + wc * N * L, pw, pc * N * L * 8);
        RS_B(rs_batch_upload(b->vals, wc, pc, b->pin + wc * N * L));
        RS_B(rs_ntt_inverse(b->enc, b->vals, wc, b->vals, wc, pc));
        for (size_t k = 0; k < pc; k++) { bs.push_back(b->vals); ps.push_back(wc + k); ls.push_back(N); }
    }
    if (has_lin) {
        RS_B(rs_vec_powers(b->lv, 0, N, lin_c));
        RS_B(rs_ntt_inverse(b->enc, b->vs, 0, b->lv, 0, 1));
        bs.push_back(b->vs); ps.push_back(0); ls.push_back(N);
        for (size_t k = 0; k < n_lin; k++) {
            const Lin &l = b->lin[k];
            if (l.kind == 0) {
                RS_B(rs_reverse_scale(b->lv, 1, b->lv, 0, b->n_mont));
                RS_B(rs_ntt_inverse(b->neg, b->lv, 1, b->lv, 1, 1));
            } else {
                RS_B(rs_aut(b->lv, 1, b->lv, 0, 1, l.idx_inv, l.is_ntt));
            }
            RS_B(rs_ntt_inverse(b->enc, b->vs, 1 + k, b->lv, 1, 1));
            bs.push_back(b->vs); ps.push_back(1 + k); ls.push_back(N);
        }
    }
    std::vector<uint64_t> dev_evals(bs.size() * L + 1);
    if (!bs.empty()) RS_B(rs_poly_evaluate_many(bs.data(), ps.data(), ls.data(), bs.size(), x, dev_evals.data()));
    const uint64_t *pw_evals = dev_evals.data();
    const uint64_t *lin_eval = dev_evals.data() + pc * L, *tr_evals = lin_eval + L;
    // ---- host scalars
    rs_field *f = b->f;
    auto E = [&](size_t i) { return pf->evals + i * L; };
    uint64_t one[32] = {0}, one_c[32] = {0};
    one_c[0] = 1;
    RS_B(rs_field_to_mont(f, one, one_c));
    auto fpow = [&](uint64_t *out, const uint64_t *base, size_t e) {
        uint64_t r[32], bsq[32];
        memcpy(r, one, 8 * L); memcpy(bsq, base, 8 * L);
        for (; e; e >>= 1) {
            if (e & 1) rs_field_mul(f, r, r, bsq);
            rs_field_mul(f, bsq, bsq, bsq);
        }
        memcpy(out, r, 8 * L);
    };
    auto eq = [&](const uint64_t *a, const uint64_t *c) { return memcmp(a, c, 8 * L) == 0; };
    uint64_t vanish[32], t0[32], t1[32], ev[32], test[32];
    fpow(vanish, x, N);
    rs_field_sub(f, vanish, vanish, one);                              // x^N - 1
    // evalCircuit at the point (verifier.go:191-215): sum_c bc * sum_terms coeff * pw(x) * prod w(x)
    auto eval_circuit = [&](uint64_t *out, const uint64_t *bc, const std::vector<rs_term> &terms, const std::vector<uint64_t> &coeffs) {
        uint64_t acc[32] = {0}, cur[32] = {0}, t[32];
        int cur_c = terms.empty() ? 0 : terms[0].constraint;
        auto close = [&]() { rs_field_mul(f, cur, cur, bc); rs_field_add(f, acc, acc, cur); memset(cur, 0, sizeof(cur)); };
        for (const rs_term &tm : terms) {
            if (tm.constraint != cur_c) { close(); cur_c = tm.constraint; }
            memcpy(t, coeffs.data() + (size_t)tm.coeff * L, 8 * L);
            if (tm.pw >= 0) rs_field_mul(f, t, t, pw_evals + (size_t)tm.pw * L);
            for (int k = 0; k < tm.n_w; k++) rs_field_mul(f, t, t, E((size_t)tm.w[k]));
            rs_field_add(f, cur, cur, t);
        }
        if (!terms.empty()) close();
        memcpy(out, acc, 8 * L);
    };
    idx = first_late;
    if (has_arith) {                                                   // verifier.go:217-225
        eval_circuit(ev, arith_bc, b->arith_terms, b->arith_coeffs);
        rs_field_mul(f, test, E(idx), vanish);
        if (!eq(ev, test)) { *result = 2; return RS_OK; }
        idx++;
    }
    uint64_t xk[32];
    if (has_lin || has_sum) fpow(xk, x, b->jrank - (N - 1));
    // quo * (x^N - 1) + remLo * x + maskSum, and remHi == x^(rank - (N - 1)) remLo (verifier.go:227-242, :264-293)
    auto closing = [&](size_t i0, const uint64_t *mask_sum, int fail_hi) -> int {
        rs_field_mul(f, t0, xk, E(i0 + 1));
        if (!eq(t0, E(i0 + 2))) return fail_hi;
        rs_field_mul(f, test, E(i0), vanish);
        rs_field_mul(f, t0, E(i0 + 1), x);
        rs_field_add(f, test, test, t0);
        rs_field_add(f, test, test, mask_sum);
        return 0;
    };
    if (has_lin) {
        if (int r = closing(idx, pf->lin_mask_sum, 3)) { *result = r; return RS_OK; }
        memset(ev, 0, sizeof(ev));
        for (size_t k = 0; k < n_lin; k++) {
            const Lin &l = b->lin[k];
            for (size_t j = 0; j < l.pairs.size(); j += 2) {
                rs_field_mul(f, t0, tr_evals + k * L, E((size_t)l.pairs[j + 1]));      // trEval * wIn(x)
                rs_field_mul(f, t1, lin_eval, E((size_t)l.pairs[j]));                  // linEval * wOut(x)
                rs_field_sub(f, t0, t0, t1);
                rs_field_mul(f, ev, ev, lin_bc);
                rs_field_add(f, ev, ev, t0);
            }
        }
        rs_field_mul(f, ev, ev, lin_bc);
        rs_field_add(f, ev, ev, E(lin_mask_idx));
        if (!eq(ev, test)) { *result = 4; return RS_OK; }
        idx += 3;
    }
    if (has_sum) {
        if (int r = closing(idx, pf->sum_mask_sum, 5)) { *result = r; return RS_OK; }
        eval_circuit(ev, sum_bc, b->sum_terms, b->sum_coeffs);
        rs_field_mul(f, ev, ev, sum_bc);
        rs_field_add(f, ev, ev, E(sum_mask_idx));
        if (!eq(ev, test)) { *result = 6; return RS_OK; }
        idx += 3;
    }
    *result = 0;
    return RS_OK;
}

}  // extern "C"
#pragma GCC visibility pop
