// buckler_host.cpp -- buckler.Prover.Prove (buckler/prover.go:42-353) as compiled host code over the library's own C ABI
// (include/ringo_buckler.h).  Host side only: SHA-256 / SHAKE128 transcripts, round structure, launch order.  Every
// polynomial lives in HBM between the first upload of the witnesses and the download of the proof.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <chrono>
#include <string>
#include <vector>

#include "../../include/ringo_buckler.h"

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
inline uint64_t rotl64(uint64_t x, int n) { return n ? (x << n) | (x >> (64 - n)) : x; }

void keccak_f(uint64_t s[25]) {
    static const uint64_t RC[24] = {0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808aULL, 0x8000000080008000ULL, 0x000000000000808bULL,
                                    0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL, 0x000000000000008aULL, 0x0000000000000088ULL,
                                    0x0000000080008009ULL, 0x000000008000000aULL, 0x000000008000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL,
                                    0x8000000000008003ULL, 0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800aULL, 0x800000008000000aULL,
                                    0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
    static const int ROT[25] = {0, 1, 62, 28, 27, 36, 44, 6, 55, 20, 3, 10, 43, 25, 39, 41, 45, 15, 21, 8, 18, 2, 61, 56, 14};   // [x + 5y]
    for (int r = 0; r < 24; r++) {
        uint64_t c[5], d[5], b[25];
        for (int x = 0; x < 5; x++) c[x] = s[x] ^ s[x + 5] ^ s[x + 10] ^ s[x + 15] ^ s[x + 20];
        for (int x = 0; x < 5; x++) d[x] = c[(x + 4) % 5] ^ rotl64(c[(x + 1) % 5], 1);
        for (int i = 0; i < 25; i++) s[i] ^= d[i % 5];
        for (int x = 0; x < 5; x++)
            for (int y = 0; y < 5; y++) b[y + 5 * ((2 * x + 3 * y) % 5)] = rotl64(s[x + 5 * y], ROT[x + 5 * y]);   // rho + pi
        for (int y = 0; y < 5; y++)
            for (int x = 0; x < 5; x++) s[x + 5 * y] = b[x + 5 * y] ^ (~b[(x + 1) % 5 + 5 * y] & b[(x + 2) % 5 + 5 * y]);
        s[0] ^= RC[r];
    }
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
    uint64_t *pin = nullptr;        // pinned staging: witnesses in, commitments out
    size_t pin_words = 0;
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
    for (rs_batch *x : {b->vals, b->ecd, b->ntt, b->w_ntt, b->pw_ntt, b->work, b->quo, b->lv, b->le, b->qr, b->rem, b->qs, b->rems})
        if (x) rs_batch_free(x);
    for (rs_ntt *t : {b->enc, b->emb, b->neg})
        if (t) rs_ntt_destroy(t);
    if (b->pin) rs_pinned_free(b->pin);
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
    if (rc == RS_OK) {
        b->pin_words = np * N * b->L;
        const size_t mask_words = (2 * N > b->sum_max_rank ? 2 * N : b->sum_max_rank) * b->L;
        if (b->pin_words < mask_words) b->pin_words = mask_words;
        b->pin = (uint64_t *)rs_pinned_alloc(b->pin_words * 8);
        if (!b->pin) rc = rs::set_error(RS_ERR_CUDA, "rs_buckler_create: pinned allocation failed");
    }
    if (rc != RS_OK) { rs_buckler_destroy(b); return rc; }
    *out = b;
    return RS_OK;
}

int rs_buckler_prove(rs_buckler *b, const uint64_t *pw, const uint64_t *w, const rs_buckler_rand *rd, rs_buckler_proof *pf, double *phase_ms) {
    if (!b || !w || !rd || !pf || (b->pc && !pw)) return rs::set_error(RS_ERR_ARG, "rs_buckler_prove: null argument");
    const size_t N = b->N, E = b->E, wc = b->wc, pc = b->pc, n_com = b->batch();
    const int L = b->L;
    const rs_jindo_params &jp = b->jp;
    const bool has_arith = !b->arith_terms.empty(), has_lin = !b->lin.empty(), has_sum = !b->sum_terms.empty();
    if (!pf->witness || !pf->evals || !pf->partial || !pf->partial_mask || !pf->encode || !pf->mlwe || !pf->incommit)
        return rs::set_error(RS_ERR_ARG, "rs_buckler_prove: proof buffers missing");
    if (rd->n_tapes != n_com || !rd->tapes || !rd->blind) return rs::set_error(RS_ERR_ARG, "rs_buckler_prove: one tape per commitment and the blinding elements are required");
    if ((has_lin && (!rd->lin_mask || !rd->lin_mask_sum)) || (has_sum && (!rd->sum_mask || !rd->sum_mask_sum)))
        return rs::set_error(RS_ERR_ARG, "rs_buckler_prove: mask polynomial missing");
    double t_phase[8] = {0};
    double t_last = now_ms();
    const double t_start = t_last;
    auto mark = [&](int k) {
        if (!phase_ms) return;
        rs_sync();
        const double t = now_ms();
        t_phase[k] += t - t_last;
        t_last = t;
    };
    const size_t com_words = (size_t)jp.out_msis * jp.n_qout * jp.d;
    std::vector<uint8_t> raw;
    auto raw_of = [&](size_t idx) -> int {
        const long long n = rs_jindo_commitment_write_raw(b->j, pf->witness + idx * com_words, nullptr, 0);
        if (n < 0) return (int)n;
        raw.resize((size_t)n);
        const long long got = rs_jindo_commitment_write_raw(b->j, pf->witness + idx * com_words, raw.data(), raw.size());
        return got == n ? RS_OK : (int)(got < 0 ? got : RS_ERR_ARG);
    };
    struct View { rs_batch *bt; size_t poly, len; };
    std::vector<View> views(n_com);
    std::vector<rs_opening *> opens(n_com, nullptr);
    struct Cleanup {
        std::vector<rs_opening *> &o; rs_opening *extra = nullptr;
        ~Cleanup() { for (rs_opening *x : o) if (x) rs_opening_free(x); if (extra) rs_opening_free(extra); }
    } cleanup{opens};
    Transcript tr;
    size_t tape_i = 0;
    auto commit = [&](size_t idx, rs_batch *bt, size_t poly, size_t len, int bind_to) -> int {
        views[idx] = View{bt, poly, len};
        RS_B(rs_jindo_commit(b->j, bt, poly, len, &rd->tapes[tape_i++], &opens[idx], pf->witness + idx * com_words));
        RS_B(raw_of(idx));
        tr.bind(bind_to, raw.data(), raw.size());
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
    tr.compute(PROJ);
    size_t idx = wc;
    uint8_t el[8 * 32];
    if (has_lin) {      // linCheckMask (prover.go:209-224)
        RS_B(rs_batch_zero(b->work, 0, 0, E));
        memcpy(b->pin, rd->lin_mask, 2 * N * L * 8);
        {   // the mask occupies the first 2N coefficients of an emb_rank polynomial
            rs_batch *view = nullptr;
            RS_B(rs_batch_wrap(b->f, rs_batch_device_ptr(b->work), 1, 2 * N, &view));
            const int rc = rs_batch_upload(view, 0, 1, b->pin);
            rs_batch_free(view);
            RS_B(rc);
        }
        RS_B(commit(idx, b->work, 0, 2 * N, ARITH_BC));
        RS_B(rs_field_marshal(b->f, rd->lin_mask_sum, el));
        tr.bind(ARITH_BC, el, 8 * L);
        if (pf->lin_mask_sum) memcpy(pf->lin_mask_sum, rd->lin_mask_sum, 8 * L);
        idx++;
    }
    if (has_sum) {      // sumCheckMask (prover.go:226-241)
        const size_t SR = b->sum_max_rank;
        RS_B(rs_batch_zero(b->work, 5, 0, E));
        memcpy(b->pin, rd->sum_mask, SR * L * 8);
        {
            rs_batch *view = nullptr;
            RS_B(rs_batch_wrap(b->f, (uint8_t *)rs_batch_device_ptr(b->work) + 5 * E * L * 8, 1, SR, &view));
            const int rc = rs_batch_upload(view, 0, 1, b->pin);
            rs_batch_free(view);
            RS_B(rc);
        }
        RS_B(commit(idx, b->work, 5, SR, ARITH_BC));
        RS_B(rs_field_marshal(b->f, rd->sum_mask_sum, el));
        tr.bind(ARITH_BC, el, 8 * L);
        if (pf->sum_mask_sum) memcpy(pf->sum_mask_sum, rd->sum_mask_sum, 8 * L);
        idx++;
    }
    mark(1);
    uint64_t bc[32], lc[32];
    RS_B(rs_field_set_bytes(b->f, tr.compute(ARITH_BC), 32, bc));
    if (has_arith) {    // arithCheck (prover.go:399-404)
        RS_B(rs_eval_circuit(b->work, 1, b->pw_ntt, b->w_ntt, b->arith_terms.data(), b->arith_terms.size(), b->arith_coeffs.data(),
                             b->arith_coeffs.size() / L, bc));
        RS_B(rs_ntt_inverse(b->emb, b->work, 1, b->work, 1, 1));
        RS_B(rs_quorem_vanishing(b->quo, 0, b->quo, 1, b->work, 1, N));
        mark(2);
        RS_B(commit(idx, b->quo, 0, b->arith_max_rank - N, EVAL_PT));
        idx++;
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
    uint64_t x[32];
    RS_B(rs_field_set_bytes(b->f, tr.compute(EVAL_PT), 32, x));
    if (pf->eval_point) memcpy(pf->eval_point, x, 8 * L);

    // ---- polyProver.Evaluate (jindo/prover.go:205-324)
    Shake128 xof;
    xof.absorb((const uint8_t *)b->crs.data(), b->crs.size());                   // ck.WriteRawTo
    for (size_t i = 0; i < n_com; i++) {
        RS_B(raw_of(i));
        xof.absorb(raw.data(), raw.size());
    }
    RS_B(rs_field_marshal(b->f, x, el));
    xof.absorb(el, 8 * L);
    rs_opening *ob = opens[0];
    if (n_com > 1) {
        std::vector<uint8_t> bb(16 * n_com);
        Shake128 fork = xof;                                                      // Read, Reset, re-absorb the same prefix (prover.go:232-248)
        fork.squeeze(bb.data(), bb.size());
        xof.absorb(bb.data(), bb.size());
        RS_B(rs_opening_alloc_zero(b->j, &cleanup.extra));
        ob = cleanup.extra;
        for (size_t i = 0; i < n_com; i++) RS_B(rs_opening_mul_add_bytes(b->j, ob, opens[i], bb.data() + 16 * i));
    }
    const size_t pq = (size_t)jp.n_q * jp.d;
    std::vector<uint64_t> part((size_t)(jp.cols + 1) * pq);
    RS_B(rs_jindo_partial_at(b->j, ob, x, part.data()));                          // leftVec(x) encoded on the device
    memcpy(pf->partial, part.data(), (size_t)jp.cols * pq * 8);
    memcpy(pf->partial_mask, part.data() + (size_t)jp.cols * pq, pq * 8);
    RS_B(rs_opening_download_all(ob, 2, pf->incommit));
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
    RS_B(rs_jindo_respond_bytes(b->j, ob, cs.data(), pf->encode, pf->mlwe));
    for (size_t i = 0; i < n_com; i++)                                            // Poly.Evaluate of every committed vector
        RS_B(rs_poly_evaluate(views[i].bt, views[i].poly, views[i].len, x, pf->evals + i * L));
    mark(5);
    if (phase_ms) {
        t_phase[7] = now_ms() - t_start;
        memcpy(phase_ms, t_phase, sizeof(t_phase));
    }
    return RS_OK;
}

}  // extern "C"
#pragma GCC visibility pop
