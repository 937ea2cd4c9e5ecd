// capi.cu -- the extern "C" surface declared in include/ringo_cuda.h.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <mutex>
#include <set>
#include <utility>

#include "common.cuh"

namespace rs {

thread_local PassProfile *g_pass_profile = nullptr;
static thread_local char g_err[512] = "";
static thread_local cudaStream_t g_stream = nullptr;
static std::atomic<unsigned long long> g_launches{0};

int set_error(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
cudaStream_t current_stream() { return g_stream; }
void count_launch(unsigned n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

static std::mutex g_once_mu;
static std::map<std::pair<int, const void *>, int> g_smem_attr;      // (device, kernel) -> bytes opted in
static std::set<std::pair<int, const void *>> g_first_use;
cudaError_t ensure_smem_attr(const void *func, int bytes) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    std::lock_guard<std::mutex> lk(g_once_mu);
    auto it = g_smem_attr.find({dev, func});
    if (it != g_smem_attr.end() && it->second >= bytes) return cudaSuccess;
    e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) g_smem_attr[{dev, func}] = bytes;
    return e;
}
bool first_use_on_device(const void *tag) {
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lk(g_once_mu);
    return g_first_use.insert({dev, tag}).second;
}

// implemented in pointwise.cu
int pointwise(int op, const rs_field *f, uint32_t *out, const uint32_t *a, const uint32_t *b, const uint32_t *c, size_t n);
int eval_circuit(const rs_field *f, uint32_t *out, const uint32_t *pw, const uint32_t *w, const rs_term *d_terms,
                 int n_terms, const uint32_t *d_coeffs, int n_coeffs, const uint32_t *bc, size_t rank);
int quorem(const rs_field *f, uint32_t *quo, uint32_t *rem, const uint32_t *p, size_t rank, size_t N);
int aut(const rs_field *f, uint32_t *out, const uint32_t *p, int logN, unsigned long long idx,
        unsigned long long idx_inv, int is_ntt, size_t n_polys);
int reverse_scale(const rs_field *f, uint32_t *out, const uint32_t *v, const uint32_t *scale, size_t N);
int horner(const rs_field *f, uint32_t *d_result, uint32_t *d_partial, const uint32_t *coeffs, size_t n,
           const uint32_t *x, const uint32_t *x_pow_ch, int CH);
int blind(const rs_field *f, uint32_t *polys, const uint32_t *d_blind, size_t n, size_t rank, size_t N);
int powers(const rs_field *f, uint32_t *out, const uint32_t *c, size_t n);
int evaluate_many(const rs_field *f, uint32_t *d_results, const void *d_descs, size_t n_polys, const uint32_t *x);
int twiddles(const rs_field *f, uint32_t *d_tw, uint32_t *d_twinv, const uint32_t *g, const uint32_t *ginv, int logN, int negacyclic);
int shoup_tables(const rs_field *f, uint32_t *d_sh, const uint32_t *d_tw, size_t n);
bool shoup_inverse_available(const rs_field *f);   // ntt_sh_l8.cu
enum PwOp { OP_ADD, OP_SUB, OP_NEG, OP_MUL, OP_MULADD, OP_MULSUB, OP_SMUL, OP_SMULADD, OP_SMULSUB };

// ---- host-side field helpers (setup only) -------------------------------------------------
template <int NL> static void host_mul_t(const rs_field *f, uint32_t *z, const uint32_t *x, const uint32_t *y) {
    FieldParams<NL> F = make_params<NL>(f);
    uint32_t t[NL];
    mont_mul<NL, 0>(t, x, y, F);
    memcpy(z, t, 4 * NL);
}
template <int NL> static void host_add_t(const rs_field *f, uint32_t *z, const uint32_t *x, const uint32_t *y) {
    FieldParams<NL> F = make_params<NL>(f);
    uint32_t t[NL];
    mod_add<NL>(t, x, y, F);
    memcpy(z, t, 4 * NL);
}
template <int NL> static void host_sub_t(const rs_field *f, uint32_t *z, const uint32_t *x, const uint32_t *y) {
    FieldParams<NL> F = make_params<NL>(f);
    uint32_t t[NL];
    mod_sub<NL>(t, x, y, F);
    memcpy(z, t, 4 * NL);
}
void host_sub(const rs_field *f, uint32_t *z, const uint32_t *x, const uint32_t *y) {
    switch (f->NL) {
        case 4: host_sub_t<4>(f, z, x, y); break;
        case 8: host_sub_t<8>(f, z, x, y); break;
        case 14: host_sub_t<14>(f, z, x, y); break;
        case 28: host_sub_t<28>(f, z, x, y); break;
    }
}
void host_mul(const rs_field *f, uint32_t *z, const uint32_t *x, const uint32_t *y) {
    switch (f->NL) {
        case 4: host_mul_t<4>(f, z, x, y); break;
        case 8: host_mul_t<8>(f, z, x, y); break;
        case 14: host_mul_t<14>(f, z, x, y); break;
        case 28: host_mul_t<28>(f, z, x, y); break;
    }
}
void host_add(const rs_field *f, uint32_t *z, const uint32_t *x, const uint32_t *y) {
    switch (f->NL) {
        case 4: host_add_t<4>(f, z, x, y); break;
        case 8: host_add_t<8>(f, z, x, y); break;
        case 14: host_add_t<14>(f, z, x, y); break;
        case 28: host_add_t<28>(f, z, x, y); break;
    }
}
void host_pow(const rs_field *f, uint32_t *z, const uint32_t *x, const uint32_t *e, int e_limbs) {
    uint32_t r[RS_MAX_NL], b[RS_MAX_NL];
    memcpy(r, f->one, 4 * f->NL);
    memcpy(b, x, 4 * f->NL);
    for (int i = 0; i < e_limbs; i++) {
        uint32_t w = e[i];
        for (int k = 0; k < 32; k++) {
            if (w & 1) host_mul(f, r, r, b);
            host_mul(f, b, b, b);
            w >>= 1;
        }
    }
    memcpy(z, r, 4 * f->NL);
}
void host_inverse(const rs_field *f, uint32_t *z, const uint32_t *x) {
    // x^(q-2): value-level identical to Inverse (element.go:1396-1548), 0 -> 0
    uint32_t e[RS_MAX_NL];
    uint64_t borrow = 2;
    for (int i = 0; i < f->NL; i++) {
        uint64_t d = (uint64_t)f->q[i] - borrow;
        e[i] = (uint32_t)d;
        borrow = (d >> 63) & 1;
    }
    host_pow(f, z, x, e, f->NL);
}
void host_set_u64(const rs_field *f, uint32_t *z, uint64_t v) {
    uint32_t t[RS_MAX_NL] = {0};
    t[0] = (uint32_t)v;
    t[1] = (uint32_t)(v >> 32);
    host_mul(f, z, t, f->r2);
}

static bool geq(const uint32_t *a, const uint32_t *b, int n) {
    for (int i = n - 1; i >= 0; i--) {
        if (a[i] > b[i]) return true;
        if (a[i] < b[i]) return false;
    }
    return true;
}

// z = x * y mod 2^(32 n)
static void host_mul_lo(uint32_t *z, const uint32_t *x, const uint32_t *y, int n) {
    uint32_t r[RS_MAX_NL] = {0};
    for (int i = 0; i < n; i++) {
        uint64_t c = 0;
        for (int j = 0; i + j < n; j++) {
            c += (uint64_t)x[j] * y[i] + r[i + j];
            r[i + j] = (uint32_t)c;
            c >>= 32;
        }
    }
    memcpy(z, r, 4 * n);
}
// Shoup pair of a Montgomery-form element e = w R mod q (shoup.cuh): (w, floor(w R / q)) = (e R^-1 mod q, e * (-q^-1) mod R)
static void host_shoup_pair(const rs_field *f, uint32_t *pair, const uint32_t *e) {
    uint32_t one[RS_MAX_NL] = {1};
    host_mul(f, pair, e, one);
    host_mul_lo(pair + f->NL, e, f->qinv_full, f->NL);
}

static bool is_pow2(size_t x) { return x && !(x & (x - 1)); }
static int ilog2(size_t x) { int l = 0; while (((size_t)1 << l) < x) l++; return l; }

static int ntt_dispatch(const rs_ntt *t, uint32_t *out, const uint32_t *in, size_t n_polys, bool inverse) {
    if (t->has52 && t->mix_pf) {
        const int rc = ntt_mix_run_l8(t, out, in, n_polys, inverse);
        if (rc != RS_ERR_UNSUPPORTED) return rc;
    }
    if (t->has52 && t->variant52 >= 0 && t->logN >= 1) {
        return ntt52_run_v2(t, out, in, n_polys, inverse);
    }
    switch (t->f->NL) {
        case 4: return ntt_run_l4(t, out, in, n_polys, inverse);
        case 8: return (t->d_tw_sh && t->logN >= 1 && (!inverse || t->d_twinv_sh)) ? ntt_run_sh_l8(t, out, in, n_polys, inverse)
                                                                                  : ntt_run_l8(t, out, in, n_polys, inverse);
        case 14: return ntt_run_l14(t, out, in, n_polys, inverse);
        case 28: return ntt_run_l28(t, out, in, n_polys, inverse);
    }
    return set_error(RS_ERR_UNSUPPORTED, "unsupported limb count %d", t->f->L64);
}

}  // namespace rs

using namespace rs;

static inline uint32_t *elem_ptr(const rs_batch *b, size_t poly, size_t off = 0) {
    return b->d + (poly * b->rank + off) * b->f->NL;
}

extern "C" {
#pragma GCC visibility push(default)

const char *rs_last_error(void) { return g_err; }
const char *rs_version(void) { return "ringo-b200 0.1 (sm_100a)"; }
unsigned long long rs_launch_count(void) { return g_launches.load(); }

int rs_init(int device) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return set_error(RS_ERR_CUDA, "no CUDA device: %s (this library has no CPU fallback)", cudaGetErrorString(e));
    RS_REQUIRE(device >= 0 && device < n, RS_ERR_ARG, "rs_init: bad device index");
    RS_CUDA_CHECK(cudaSetDevice(device));
    cudaDeviceProp p;
    RS_CUDA_CHECK(cudaGetDeviceProperties(&p, device));
    if (p.major != 10) return set_error(RS_ERR_CUDA, "device %s is sm_%d%d; kernels are built for sm_100a only", p.name, p.major, p.minor);
    // per-call scratch (tapes, partial sums, sampler output: hundreds of MB per Commit) comes from the
    // stream-ordered pool; keep it cached across synchronisations instead of returning it to the driver
    {
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            unsigned long long keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
    }
    return RS_OK;
}

int rs_set_stream(void *s) { g_stream = (cudaStream_t)s; return RS_OK; }
void *rs_get_stream(void) { return (void *)g_stream; }
int rs_stream_create(void **out) {
    RS_REQUIRE(out, RS_ERR_ARG, "rs_stream_create: null argument");
    cudaStream_t st;
    RS_CUDA_CHECK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    *out = (void *)st;
    return RS_OK;
}
void rs_stream_destroy(void *s) { if (s) cudaStreamDestroy((cudaStream_t)s); }
// `waiter` continues only after everything enqueued on `signaller` so far has run (an event recorded now; no host wait)
int rs_stream_wait(void *waiter, void *signaller) {
    cudaEvent_t e;
    RS_CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    cudaError_t r = cudaEventRecord(e, (cudaStream_t)signaller);
    if (r == cudaSuccess) r = cudaStreamWaitEvent((cudaStream_t)waiter, e, 0);
    cudaEventDestroy(e);
    if (r != cudaSuccess) return set_error(RS_ERR_CUDA, "rs_stream_wait: %s", cudaGetErrorString(r));
    return RS_OK;
}
int rs_sync(void) { RS_CUDA_CHECK(cudaStreamSynchronize(g_stream)); return RS_OK; }

void *rs_pinned_alloc(size_t bytes) {
    void *p = nullptr;
    if (cudaMallocHost(&p, bytes) != cudaSuccess) { set_error(RS_ERR_CUDA, "cudaMallocHost(%zu) failed", bytes); return nullptr; }
    return p;
}
void rs_pinned_free(void *p) { if (p) cudaFreeHost(p); }

// ---- field ----------------------------------------------------------------------------------
int rs_field_create(const uint64_t *q, int L64, rs_field **out) {
    RS_REQUIRE(q && out, RS_ERR_ARG, "rs_field_create: null argument");
    RS_REQUIRE(L64 == 2 || L64 == 4 || L64 == 7 || L64 == 14, RS_ERR_UNSUPPORTED,
               "rs_field_create: supported limb counts are 2, 4, 7, 14");
    RS_REQUIRE(q[0] & 1, RS_ERR_ARG, "rs_field_create: modulus must be odd");
    RS_REQUIRE(q[L64 - 1] != 0, RS_ERR_ARG, "rs_field_create: top limb of the modulus is zero");
    rs_field *f = new rs_field();
    memset(f, 0, sizeof(*f));
    f->L64 = L64;
    f->NL = 2 * L64;
    for (int i = 0; i < L64; i++) { f->q[2 * i] = (uint32_t)q[i]; f->q[2 * i + 1] = (uint32_t)(q[i] >> 32); }
    uint32_t inv = 1;
    for (int i = 0; i < 5; i++) inv *= 2 - f->q[0] * inv;   // Newton: q^-1 mod 2^32
    f->qinv = 0u - inv;
    f->qs = 0;
    if (f->q[0] == 1) f->qs = (f->q[1] == 0) ? 2 : 1;
    // R mod q and R^2 mod q by doubling (same values as the generated rSquare, element.go:784)
    uint32_t t[RS_MAX_NL] = {1};
    const int n = f->NL;
    for (int i = 0; i < 64 * n; i++) {
        uint32_t c = 0;
        for (int k = 0; k < n; k++) { uint32_t nc = t[k] >> 31; t[k] = (t[k] << 1) | c; c = nc; }
        if (c || geq(t, f->q, n)) {
            uint64_t br = 0;
            for (int k = 0; k < n; k++) { uint64_t d = (uint64_t)t[k] - f->q[k] - br; t[k] = (uint32_t)d; br = (d >> 63) & 1; }
        }
        if (i == 32 * n - 1) memcpy(f->one, t, 4 * n);
    }
    memcpy(f->r2, t, 4 * n);
    // lazy [0, 2q) arithmetic needs 3q 2^32 + 2q < 2^(32 (n+1)); checked as 3q + 1 < 2^(32 n) (and 2q fits n limbs)
    {
        uint64_t c = 0; uint32_t q3[RS_MAX_NL + 1];
        for (int k = 0; k < n; k++) { c += (uint64_t)f->q[k] * 3 + (k == 0 ? 1 : 0); q3[k] = (uint32_t)c; c >>= 32; }
        f->lazy_ok = (c == 0) && !getenv("RS_NTT_NO_LAZY");
        c = 0;
        for (int k = 0; k < n; k++) { c += (uint64_t)f->q[k] * 2; f->q2[k] = (uint32_t)c; c >>= 32; }
    }
    // -q^-1 mod 2^(32 n) by Newton (y <- y (2 - q y) doubles the correct low bits), and the validity of the Shoup NTT
    // kernels (shoup.cuh): 18 q < 5 2^(32 n), the top 64 bits of q at least 2^8, q == 1 mod 2^32; compiled for 8 limbs
    {
        uint32_t y[RS_MAX_NL] = {1}, t2[RS_MAX_NL];
        for (int it = 0; it < 10; it++) {
            host_mul_lo(t2, f->q, y, n);
            uint64_t br = 0;
            for (int k = 0; k < n; k++) { uint64_t d = (uint64_t)(k == 0 ? 2 : 0) - t2[k] - br; t2[k] = (uint32_t)d; br = (d >> 63) & 1; }
            host_mul_lo(y, y, t2, n);
        }
        uint64_t br = 0;
        for (int k = 0; k < n; k++) { uint64_t d = (uint64_t)0 - y[k] - br; f->qinv_full[k] = (uint32_t)d; br = (d >> 63) & 1; }
        uint64_t c = 0;
        for (int k = 0; k < n; k++) { c += (uint64_t)f->q[k] * 18; c >>= 32; }
        const uint64_t top64 = ((uint64_t)f->q[n - 1] << 32) | f->q[n - 2];
        f->shoup_ok = f->lazy_ok && n == 8 && f->qs >= 1 && c < 5 && top64 >= 256 && !getenv("RS_NTT_NO_SHOUP");
    }
    *out = f;
    return RS_OK;
}
void rs_field_destroy(rs_field *f) { delete f; }
int rs_field_limbs(const rs_field *f) { return f ? f->L64 : RS_ERR_ARG; }

int rs_field_to_mont(const rs_field *f, uint64_t *out, const uint64_t *canonical) {
    RS_REQUIRE(f && out && canonical, RS_ERR_ARG, "rs_field_to_mont: null argument");
    uint32_t z[RS_MAX_NL];
    host_mul(f, z, (const uint32_t *)canonical, f->r2);
    memcpy(out, z, 4 * f->NL);
    return RS_OK;
}
int rs_field_from_mont(const rs_field *f, uint64_t *out, const uint64_t *mont) {
    RS_REQUIRE(f && out && mont, RS_ERR_ARG, "rs_field_from_mont: null argument");
    uint32_t z[RS_MAX_NL], one[RS_MAX_NL] = {1};
    host_mul(f, z, (const uint32_t *)mont, one);
    memcpy(out, z, 4 * f->NL);
    return RS_OK;
}

// scalar arithmetic on Montgomery elements (host): what a verifier-side caller needs for a handful of field elements
int rs_field_mul(const rs_field *f, uint64_t *out, const uint64_t *a, const uint64_t *b) {
    RS_REQUIRE(f && out && a && b, RS_ERR_ARG, "rs_field_mul: null argument");
    host_mul(f, (uint32_t *)out, (const uint32_t *)a, (const uint32_t *)b);
    return RS_OK;
}
int rs_field_add(const rs_field *f, uint64_t *out, const uint64_t *a, const uint64_t *b) {
    RS_REQUIRE(f && out && a && b, RS_ERR_ARG, "rs_field_add: null argument");
    host_add(f, (uint32_t *)out, (const uint32_t *)a, (const uint32_t *)b);
    return RS_OK;
}
int rs_field_sub(const rs_field *f, uint64_t *out, const uint64_t *a, const uint64_t *b) {
    RS_REQUIRE(f && out && a && b, RS_ERR_ARG, "rs_field_sub: null argument");
    host_sub(f, (uint32_t *)out, (const uint32_t *)a, (const uint32_t *)b);
    return RS_OK;
}

// SetBytes (zp element.go:891-915): the bytes as a big-endian integer, reduced mod q (any length up to 128 bytes), in
// Montgomery form.  Horner over chunks of L64 words, each step a multiplication by R = 2^(64 L64) done as a Montgomery
// product with R^2.
int rs_field_set_bytes(const rs_field *f, const uint8_t *be, size_t len, uint64_t *out_mont) {
    RS_REQUIRE(f && (be || len == 0) && out_mont && len <= 128, RS_ERR_ARG, "rs_field_set_bytes: bad argument");
    const int L = f->L64, NL = f->NL;
    uint64_t words[16] = {0};
    for (size_t i = 0; i < len; i++) words[i / 8] |= (uint64_t)be[len - 1 - i] << (8 * (i % 8));
    const int n_words = (int)((len + 7) / 8), n_chunks = (n_words + L - 1) / L;
    uint32_t acc[RS_MAX_NL] = {0}, chunk[RS_MAX_NL], t[RS_MAX_NL];
    for (int ch = n_chunks - 1; ch >= 0; ch--) {
        host_mul(f, acc, acc, f->r2);
        memset(chunk, 0, sizeof(chunk));
        for (int w = 0; w < L && ch * L + w < n_words; w++) { chunk[2 * w] = (uint32_t)words[ch * L + w]; chunk[2 * w + 1] = (uint32_t)(words[ch * L + w] >> 32); }
        host_mul(f, t, chunk, f->r2);
        host_add(f, acc, acc, t);
    }
    memcpy(out_mont, acc, 4 * NL);
    return RS_OK;
}
// Marshal (element.go:873-883): the canonical value as 8 L64 big-endian bytes
int rs_field_marshal(const rs_field *f, const uint64_t *mont, uint8_t *out_be) {
    RS_REQUIRE(f && mont && out_be, RS_ERR_ARG, "rs_field_marshal: null argument");
    uint64_t c[RS_MAX_NL / 2];
    int rc = rs_field_from_mont(f, c, mont);
    if (rc) return rc;
    const int L = f->L64;
    for (int i = 0; i < L; i++)
        for (int b = 0; b < 8; b++) out_be[8 * (L - 1 - i) + (7 - b)] = (uint8_t)(c[i] >> (8 * b));
    return RS_OK;
}

// ---- batches ----------------------------------------------------------------------------------
int rs_batch_alloc(rs_field *f, size_t n_polys, size_t rank, rs_batch **out) {
    RS_REQUIRE(f && out, RS_ERR_ARG, "rs_batch_alloc: null argument");
    rs_batch *b = new rs_batch{f, n_polys, rank, nullptr, true};
    const size_t bytes = n_polys * rank * f->NL * 4;
    cudaError_t e = cudaMalloc((void **)&b->d, bytes ? bytes : 16);
    if (e != cudaSuccess) { delete b; return set_error(RS_ERR_CUDA, "cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e)); }
    *out = b;
    return RS_OK;
}
int rs_batch_wrap(rs_field *f, void *dptr, size_t n_polys, size_t rank, rs_batch **out) {
    RS_REQUIRE(f && out && dptr, RS_ERR_ARG, "rs_batch_wrap: null argument");
    RS_REQUIRE(((uintptr_t)dptr & 15) == 0, RS_ERR_ARG, "rs_batch_wrap: device pointer must be 16-byte aligned");
    *out = new rs_batch{f, n_polys, rank, (uint32_t *)dptr, false};
    return RS_OK;
}
void rs_batch_free(rs_batch *b) {
    if (!b) return;
    if (b->owned && b->d) cudaFree(b->d);
    delete b;
}
void *rs_batch_device_ptr(rs_batch *b) { return b ? b->d : nullptr; }
size_t rs_batch_polys(const rs_batch *b) { return b ? b->n_polys : 0; }
size_t rs_batch_rank(const rs_batch *b) { return b ? b->rank : 0; }

int rs_batch_upload(rs_batch *b, size_t poly0, size_t n, const uint64_t *host) {
    RS_REQUIRE(b && host, RS_ERR_ARG, "rs_batch_upload: null argument");
    RS_REQUIRE(poly0 + n <= b->n_polys, RS_ERR_SHAPE, "rs_batch_upload: poly range out of bounds");
    RS_CUDA_CHECK(cudaMemcpyAsync(elem_ptr(b, poly0), host, n * b->rank * b->f->NL * 4, cudaMemcpyHostToDevice, g_stream));
    return RS_OK;
}
int rs_batch_download(const rs_batch *b, size_t poly0, size_t n, uint64_t *host) {
    RS_REQUIRE(b && host, RS_ERR_ARG, "rs_batch_download: null argument");
    RS_REQUIRE(poly0 + n <= b->n_polys, RS_ERR_SHAPE, "rs_batch_download: poly range out of bounds");
    RS_CUDA_CHECK(cudaMemcpyAsync(host, elem_ptr(b, poly0), n * b->rank * b->f->NL * 4, cudaMemcpyDeviceToHost, g_stream));
    RS_CUDA_CHECK(cudaStreamSynchronize(g_stream));
    return RS_OK;
}
int rs_batch_download_async(const rs_batch *b, size_t poly0, size_t n, uint64_t *host) {
    RS_REQUIRE(b && host, RS_ERR_ARG, "rs_batch_download_async: null argument");
    RS_REQUIRE(poly0 + n <= b->n_polys, RS_ERR_SHAPE, "rs_batch_download_async: poly range out of bounds");
    RS_CUDA_CHECK(cudaMemcpyAsync(host, elem_ptr(b, poly0), n * b->rank * b->f->NL * 4, cudaMemcpyDeviceToHost, g_stream));
    return RS_OK;
}
int rs_batch_copy(rs_batch *dst, size_t dp, size_t doff, const rs_batch *src, size_t sp, size_t soff, size_t n) {
    RS_REQUIRE(dst && src, RS_ERR_ARG, "rs_batch_copy: null argument");
    RS_REQUIRE(dst->f->NL == src->f->NL, RS_ERR_SHAPE, "rs_batch_copy: field mismatch");
    RS_REQUIRE(dp < dst->n_polys && sp < src->n_polys && doff + n <= dst->rank && soff + n <= src->rank, RS_ERR_SHAPE,
               "rs_batch_copy: range out of bounds");
    if (n) RS_CUDA_CHECK(cudaMemcpyAsync(elem_ptr(dst, dp, doff), elem_ptr(src, sp, soff), n * src->f->NL * 4,
                                         cudaMemcpyDeviceToDevice, g_stream));
    return RS_OK;
}
int rs_batch_zero(rs_batch *dst, size_t poly, size_t off, size_t n) {
    RS_REQUIRE(dst, RS_ERR_ARG, "rs_batch_zero: null argument");
    RS_REQUIRE(poly < dst->n_polys && off + n <= dst->rank, RS_ERR_SHAPE, "rs_batch_zero: range out of bounds");
    if (n) RS_CUDA_CHECK(cudaMemsetAsync(elem_ptr(dst, poly, off), 0, n * dst->f->NL * 4, g_stream));
    return RS_OK;
}

// ---- NTT ---------------------------------------------------------------------------------------
int rs_ntt_create(rs_field *f, int logN, int negacyclic, rs_ntt **out) {
    RS_REQUIRE(f && out, RS_ERR_ARG, "rs_ntt_create: null argument");
    RS_REQUIRE(logN >= 0 && logN <= 26, RS_ERR_UNSUPPORTED, "rs_ntt_create: logN out of range");
    const size_t N = (size_t)1 << logN;
    const int n = f->NL;
    // (q-1) mod 2N == 0  (ntt.go:35-37, :162-164)
    {
        uint64_t q0 = ((uint64_t)f->q[1] << 32) | f->q[0];
        if (((q0 - 1) & (2 * N - 1)) != 0) return set_error(RS_ERR_NTT, "NTT not supported");
    }
    rs_ntt *t = new rs_ntt();
    t->f = f; t->logN = logN; t->negacyclic = negacyclic; t->d_tw = t->d_twinv = nullptr;
    t->d_tw_sh = t->d_twinv_sh = nullptr;
    t->h_tw.assign(N * n, 0);
    t->h_twinv.assign(N * n, 0);
    // generator search: smallest x >= 2 with (x^((q-1)/order))^(order/2) != 1  (ntt.go:40-53, :167-180)
    const uint64_t order = negacyclic ? 2 * N : N;
    uint32_t g[RS_MAX_NL], ginv[RS_MAX_NL];
    if (order >= 2) {
        uint32_t e[RS_MAX_NL], qm1[RS_MAX_NL];
        memcpy(qm1, f->q, 4 * n);
        qm1[0] -= 1;   // q odd
        const int sh = ilog2(order);
        for (int i = 0; i < n; i++) {
            const int w = sh / 32, b = sh % 32;
            uint32_t lo = (i + w < n) ? qm1[i + w] : 0, hi = (i + w + 1 < n) ? qm1[i + w + 1] : 0;
            e[i] = b ? ((lo >> b) | (hi << (32 - b))) : lo;
        }
        uint32_t half[2] = {(uint32_t)(order >> 1), (uint32_t)((order >> 1) >> 32)};
        for (uint64_t x = 2;; x++) {
            uint32_t xm[RS_MAX_NL], gp[RS_MAX_NL];
            host_set_u64(f, xm, x);
            host_pow(f, g, xm, e, n);
            host_pow(f, gp, g, half, 2);
            if (memcmp(gp, f->one, 4 * n) != 0) break;
        }
        host_inverse(f, ginv, g);
    } else {
        memcpy(g, f->one, 4 * n);
        memcpy(ginv, f->one, 4 * n);
    }
    // the tables themselves are generated in HBM (twiddle_kernel: one square-and-multiply per entry instead of the reference's N
    // sequential products) and read back once: rs_ntt_tables, the N^-1 * tw fold below and the FP64 variant want them on the host
    const size_t bytes = N * n * 4;
    {
        cudaError_t e1 = cudaMalloc((void **)&t->d_tw, bytes), e2 = cudaMalloc((void **)&t->d_twinv, bytes);
        int rc = (e1 == cudaSuccess && e2 == cudaSuccess) ? RS_OK : set_error(RS_ERR_CUDA, "rs_ntt_create: cudaMalloc failed: %s", cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
        if (rc == RS_OK) rc = twiddles(f, t->d_tw, t->d_twinv, g, ginv, logN, negacyclic);
        if (rc == RS_OK && (cudaMemcpyAsync(t->h_tw.data(), t->d_tw, bytes, cudaMemcpyDeviceToHost, g_stream) != cudaSuccess ||
                            cudaMemcpyAsync(t->h_twinv.data(), t->d_twinv, bytes, cudaMemcpyDeviceToHost, g_stream) != cudaSuccess ||
                            cudaStreamSynchronize(g_stream) != cudaSuccess))
            rc = set_error(RS_ERR_CUDA, "rs_ntt_create: twiddle generation failed: %s", cudaGetErrorString(cudaGetLastError()));
        if (rc != RS_OK) {
            if (t->d_tw) cudaFree(t->d_tw);
            if (t->d_twinv) cudaFree(t->d_twinv);
            delete t;
            return rc;
        }
    }
    uint32_t nm[RS_MAX_NL];
    host_set_u64(f, nm, (uint64_t)N);
    host_inverse(f, t->ninv, nm);
    if (N >= 2) host_mul(f, t->ninv_w1, t->ninv, &t->h_twinv[1 * n]);
    else memcpy(t->ninv_w1, t->ninv, 4 * n);
    // Shoup tables (shoup.cuh): every butterfly product has a table entry as one operand; (w, floor(w R / q)) pairs made in HBM
    if (f->shoup_ok && N >= 2) {
        // (the inverse kernels exist for the moduli compiled into their kernels only, see ntt_sh_l8.cu; RS_NTT_NO_SHOUP_INV: off)
        const bool inv_too = shoup_inverse_available(f) && !getenv("RS_NTT_NO_SHOUP_INV");
        host_shoup_pair(f, t->ninv_sh, t->ninv);
        host_shoup_pair(f, t->ninv_w1_sh, t->ninv_w1);
        cudaError_t e1 = cudaMalloc((void **)&t->d_tw_sh, 2 * bytes), e2 = inv_too ? cudaMalloc((void **)&t->d_twinv_sh, 2 * bytes) : cudaSuccess;
        int rc = (e1 == cudaSuccess && e2 == cudaSuccess) ? RS_OK : set_error(RS_ERR_CUDA, "rs_ntt_create: cudaMalloc failed: %s", cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
        if (rc == RS_OK) rc = shoup_tables(f, t->d_tw_sh, t->d_tw, N);
        if (rc == RS_OK && inv_too) rc = shoup_tables(f, t->d_twinv_sh, t->d_twinv, N);
        if (rc == RS_OK && cudaStreamSynchronize(g_stream) != cudaSuccess)
            rc = set_error(RS_ERR_CUDA, "rs_ntt_create: Shoup table generation failed: %s", cudaGetErrorString(cudaGetLastError()));
        if (rc != RS_OK) {
            cudaFree(t->d_tw); cudaFree(t->d_twinv);
            if (t->d_tw_sh) cudaFree(t->d_tw_sh);
            if (t->d_twinv_sh) cudaFree(t->d_twinv_sh);
            delete t;
            return rc;
        }
    }
    // ---- FP64-pipe paths (ntt52.cuh), 8-word fields with 2^260 >= 36 q.  Both are bit-exact and both measured no faster than
    // the integer kernels on B200 (a DFMA costs two dispatch cycles, and half-rate instructions of different
    // pipes barely overlap), so they are opt-in: RS_NTT52=1 runs the FP64 kernels alone, RS_NTT_MIX=pf/p the two-pipe kernels.
    // RS_NTT52_VARIANT=k: FP64 kernels alone (tuning variant k); RS_NTT_MIX=pf/p: share of FP64 CTAs in the two-pipe kernels
    t->has52 = 0; t->d_tw52 = t->d_twinv52 = nullptr; t->variant52 = -1; t->mix_pf = 0; t->mix_p = 4;
    std::vector<double> tw52, twinv52;
    if (n == 8 && N >= 2 && (getenv("RS_NTT52") || getenv("RS_NTT_MIX")) && f52_setup<8, 5>(t->c52.F, f->q)) {
        if (const char *v = getenv("RS_NTT52")) { t->variant52 = atoi(v) ? 2 : -1; }
        if (const char *v = getenv("RS_NTT_MIX")) {
            unsigned a = 0, b = 1;
            if (sscanf(v, "%u/%u", &a, &b) >= 1 && b > a) { t->mix_pf = a; t->mix_p = b; }
        }
        // x R -> x R' = x R 2^(260 - 256): four modular doublings
        auto to52 = [&](const uint32_t *w, double *out) {
            uint32_t x[RS_MAX_NL];
            memcpy(x, w, 4 * n);
            for (int d = 0; d < 52 * 5 - 32 * 8; d++) host_add(f, x, x, x);
            long long l[5];
            f52_from_words<8, 5>(l, x);
            for (int k = 0; k < 5; k++) out[k] = (double)l[k];
        };
        tw52.assign(N * TW52_STRIDE, 0.0); twinv52.assign(N * TW52_STRIDE, 0.0);
        for (size_t i = 0; i < N; i++) { to52(&t->h_tw[i * n], &tw52[i * TW52_STRIDE]); to52(&t->h_twinv[i * n], &twinv52[i * TW52_STRIDE]); }
        to52(t->ninv, t->c52.ninv);
        to52(t->ninv_w1, t->c52.ninv_w1);
        memcpy(t->c52.qw, f->q, 32);
        t->has52 = 1;
    }
    cudaError_t e1 = cudaSuccess, e2 = cudaSuccess;
    if (t->has52) {
        e1 = cudaMalloc((void **)&t->d_tw52, tw52.size() * 8);
        e2 = cudaMalloc((void **)&t->d_twinv52, twinv52.size() * 8);
        if (e1 == cudaSuccess) e1 = cudaMemcpy(t->d_tw52, tw52.data(), tw52.size() * 8, cudaMemcpyHostToDevice);
        if (e2 == cudaSuccess) e2 = cudaMemcpy(t->d_twinv52, twinv52.data(), twinv52.size() * 8, cudaMemcpyHostToDevice);
    }
    if (e1 != cudaSuccess || e2 != cudaSuccess) {
        if (t->d_tw) cudaFree(t->d_tw);
        if (t->d_twinv) cudaFree(t->d_twinv);
        if (t->d_tw52) cudaFree(t->d_tw52);
        if (t->d_twinv52) cudaFree(t->d_twinv52);
        if (t->d_tw_sh) cudaFree(t->d_tw_sh);
        if (t->d_twinv_sh) cudaFree(t->d_twinv_sh);
        delete t;
        return set_error(RS_ERR_CUDA, "rs_ntt_create: cudaMalloc failed: %s", cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
    }
    *out = t;
    return RS_OK;
}
void rs_ntt_destroy(rs_ntt *t) {
    if (!t) return;
    if (t->d_tw) cudaFree(t->d_tw);
    if (t->d_twinv) cudaFree(t->d_twinv);
    if (t->d_tw52) cudaFree(t->d_tw52);
    if (t->d_twinv52) cudaFree(t->d_twinv52);
    if (t->d_tw_sh) cudaFree(t->d_tw_sh);
    if (t->d_twinv_sh) cudaFree(t->d_twinv_sh);
    delete t;
}
int rs_ntt_rank_log2(const rs_ntt *t) { return t ? t->logN : RS_ERR_ARG; }
int rs_ntt_tables(const rs_ntt *t, uint64_t *tw, uint64_t *twinv, uint64_t *rankinv) {
    RS_REQUIRE(t, RS_ERR_ARG, "rs_ntt_tables: null argument");
    if (tw) memcpy(tw, t->h_tw.data(), t->h_tw.size() * 4);
    if (twinv) memcpy(twinv, t->h_twinv.data(), t->h_twinv.size() * 4);
    if (rankinv) memcpy(rankinv, t->ninv, 4 * t->f->NL);
    return RS_OK;
}

static int ntt_call(rs_ntt *t, rs_batch *dst, size_t dp, const rs_batch *src, size_t sp, size_t n, bool inverse) {
    RS_REQUIRE(t && dst && src, RS_ERR_ARG, "ntt: null argument");
    const size_t N = (size_t)1 << t->logN;
    RS_REQUIRE(dst->rank == N && src->rank == N, RS_ERR_SHAPE, "inconsistent input(s)");
    RS_REQUIRE(dst->f->NL == t->f->NL && src->f->NL == t->f->NL, RS_ERR_SHAPE, "ntt: field mismatch");
    RS_REQUIRE(dp + n <= dst->n_polys && sp + n <= src->n_polys, RS_ERR_SHAPE, "ntt: poly range out of bounds");
    return ntt_dispatch(t, elem_ptr(dst, dp), elem_ptr(src, sp), n, inverse);
}
int rs_ntt_forward(rs_ntt *t, rs_batch *dst, size_t dp, const rs_batch *src, size_t sp, size_t n) {
    return ntt_call(t, dst, dp, src, sp, n, false);
}
int rs_ntt_inverse(rs_ntt *t, rs_batch *dst, size_t dp, const rs_batch *src, size_t sp, size_t n) {
    return ntt_call(t, dst, dp, src, sp, n, true);
}
int rs_ntt_profile(rs_ntt *t, rs_batch *b, size_t n_polys, int inverse, int reps, float *pass_ms, int *n_passes) {
    RS_REQUIRE(t && b && pass_ms && n_passes && reps > 0, RS_ERR_ARG, "rs_ntt_profile: bad argument");
    RS_REQUIRE(b->rank == ((size_t)1 << t->logN) && n_polys <= b->n_polys, RS_ERR_SHAPE, "inconsistent input(s)");
    PassProfile prof;
    cudaEvent_t start;
    RS_CUDA_CHECK(cudaEventCreate(&start));
    for (int i = 0; i < 16; i++) RS_CUDA_CHECK(cudaEventCreate(&prof.ev[i]));
    const int cap = *n_passes;
    double acc[16] = {0};
    int np = 0, rc = RS_OK;
    for (int r = 0; r < reps && rc == RS_OK; r++) {
        prof.n = 0;
        RS_CUDA_CHECK(cudaEventRecord(start, g_stream));
        g_pass_profile = &prof;
        rc = ntt_dispatch(t, b->d, b->d, n_polys, inverse != 0);
        g_pass_profile = nullptr;
        if (rc) break;
        RS_CUDA_CHECK(cudaStreamSynchronize(g_stream));
        np = prof.n;
        for (int i = 0; i < np; i++) {
            float ms = 0;
            RS_CUDA_CHECK(cudaEventElapsedTime(&ms, i ? prof.ev[i - 1] : start, prof.ev[i]));
            acc[i] += ms;
        }
    }
    cudaEventDestroy(start);
    for (int i = 0; i < 16; i++) cudaEventDestroy(prof.ev[i]);
    if (rc) return rc;
    RS_REQUIRE(np <= cap, RS_ERR_ARG, "rs_ntt_profile: pass_ms too small");
    for (int i = 0; i < np; i++) pass_ms[i] = (float)(acc[i] / reps);
    *n_passes = np;
    return RS_OK;
}
int rs_ntt_inverse_embed(rs_ntt *t, rs_batch *dst, size_t dp, const rs_batch *src, size_t sp, size_t n, int zero_pad) {
    RS_REQUIRE(t && dst && src, RS_ERR_ARG, "ntt: null argument");
    const size_t N = (size_t)1 << t->logN;
    RS_REQUIRE(dst->rank >= N && src->rank >= N, RS_ERR_SHAPE, "inconsistent input(s)");
    RS_REQUIRE(dst->f->NL == t->f->NL && src->f->NL == t->f->NL, RS_ERR_SHAPE, "ntt: field mismatch");
    RS_REQUIRE(dp + n <= dst->n_polys && sp + n <= src->n_polys, RS_ERR_SHAPE, "ntt: poly range out of bounds");
    if (n == 0) return RS_OK;
    if (src->rank == N && dst->rank == N) return ntt_dispatch(t, elem_ptr(dst, dp), elem_ptr(src, sp), n, true);
    if (src->rank == N && dst != src) {
        // Encoder.EncodeTo for a whole batch of witnesses (buckler/encoder.go:32-38): ONE batched transform of the packed
        // source into a stream-ordered temporary, then one strided copy into the low N coefficients of the wider
        // polynomials and one strided zero-fill of their tails -- `passes` + 2 launches instead of (passes + 1) per witness
        const size_t eb = (size_t)t->f->NL * 4;
        uint32_t *tmp = nullptr;
        RS_CUDA_CHECK(cudaMallocAsync((void **)&tmp, n * N * eb, g_stream));
        int rc = ntt_dispatch(t, tmp, elem_ptr(src, sp), n, true);
        if (rc == RS_OK && cudaMemcpy2DAsync(elem_ptr(dst, dp), dst->rank * eb, tmp, N * eb, N * eb, n, cudaMemcpyDeviceToDevice, g_stream) != cudaSuccess)
            rc = set_error(RS_ERR_CUDA, "rs_ntt_inverse_embed: %s", cudaGetErrorString(cudaGetLastError()));
        if (rc == RS_OK && zero_pad && dst->rank > N &&
            cudaMemset2DAsync(elem_ptr(dst, dp, N), dst->rank * eb, 0, (dst->rank - N) * eb, n, g_stream) != cudaSuccess)
            rc = set_error(RS_ERR_CUDA, "rs_ntt_inverse_embed: %s", cudaGetErrorString(cudaGetLastError()));
        cudaFreeAsync(tmp, g_stream);
        return rc;
    }
    for (size_t i = 0; i < n; i++) {
        int rc = ntt_dispatch(t, elem_ptr(dst, dp + i), elem_ptr(src, sp + i), 1, true);
        if (rc) return rc;
        if (zero_pad && dst->rank > N) {
            rc = rs_batch_zero(dst, dp + i, N, dst->rank - N);
            if (rc) return rc;
        }
    }
    return RS_OK;
}

// ---- element-wise --------------------------------------------------------------------------------
static int pw_check(const rs_batch *o, size_t op, const rs_batch *a, size_t ap, const rs_batch *b, size_t bp, size_t n) {
    RS_REQUIRE(o && a, RS_ERR_ARG, "vec op: null argument");
    RS_REQUIRE(o->rank == a->rank && (!b || b->rank == a->rank), RS_ERR_SHAPE, "inconsistent input(s)");
    RS_REQUIRE(o->f->NL == a->f->NL && (!b || b->f->NL == a->f->NL), RS_ERR_SHAPE, "vec op: field mismatch");
    RS_REQUIRE(op + n <= o->n_polys && ap + n <= a->n_polys && (!b || bp + n <= b->n_polys), RS_ERR_SHAPE,
               "vec op: poly range out of bounds");
    return RS_OK;
}
#define RS_BINARY(NAME, OPC)                                                                                       \
    int NAME(rs_batch *o, size_t op, const rs_batch *a, size_t ap, const rs_batch *b, size_t bp, size_t n) {        \
        RS_REQUIRE(b, RS_ERR_ARG, "vec op: null argument");                                                         \
        int rc = pw_check(o, op, a, ap, b, bp, n);                                                                  \
        if (rc) return rc;                                                                                          \
        return pointwise(OPC, o->f, elem_ptr(o, op), elem_ptr(a, ap), elem_ptr(b, bp), nullptr, n * o->rank);       \
    }
RS_BINARY(rs_vec_add, OP_ADD)
RS_BINARY(rs_vec_sub, OP_SUB)
RS_BINARY(rs_vec_mul, OP_MUL)
RS_BINARY(rs_vec_mul_add, OP_MULADD)
RS_BINARY(rs_vec_mul_sub, OP_MULSUB)
int rs_vec_neg(rs_batch *o, size_t op, const rs_batch *a, size_t ap, size_t n) {
    int rc = pw_check(o, op, a, ap, nullptr, 0, n);
    if (rc) return rc;
    return pointwise(OP_NEG, o->f, elem_ptr(o, op), elem_ptr(a, ap), nullptr, nullptr, n * o->rank);
}
#define RS_SCALAR(NAME, OPC)                                                                                        \
    int NAME(rs_batch *o, size_t op, const rs_batch *a, size_t ap, const uint64_t *c, size_t n) {                    \
        RS_REQUIRE(c, RS_ERR_ARG, "vec op: null scalar");                                                            \
        int rc = pw_check(o, op, a, ap, nullptr, 0, n);                                                              \
        if (rc) return rc;                                                                                           \
        return pointwise(OPC, o->f, elem_ptr(o, op), elem_ptr(a, ap), nullptr, (const uint32_t *)c, n * o->rank);    \
    }
RS_SCALAR(rs_vec_scalar_mul, OP_SMUL)
RS_SCALAR(rs_vec_scalar_mul_add, OP_SMULADD)
RS_SCALAR(rs_vec_scalar_mul_sub, OP_SMULSUB)

// ---- Buckler ---------------------------------------------------------------------------------------
int rs_eval_circuit(rs_batch *out, size_t out_poly, const rs_batch *pw, const rs_batch *w, const rs_term *terms,
                    size_t n_terms, const uint64_t *coeffs, size_t n_coeffs, const uint64_t *batch_const) {
    RS_REQUIRE(out && w && terms && coeffs && batch_const, RS_ERR_ARG, "rs_eval_circuit: null argument");
    RS_REQUIRE(out_poly < out->n_polys, RS_ERR_SHAPE, "rs_eval_circuit: out poly out of bounds");
    RS_REQUIRE(w->rank == out->rank && (!pw || pw->rank == out->rank), RS_ERR_SHAPE, "inconsistent input(s)");
    const int NL = out->f->NL;
    for (size_t k = 0; k < n_terms; k++) {
        const rs_term &t = terms[k];
        RS_REQUIRE(t.coeff >= 0 && (size_t)t.coeff < n_coeffs, RS_ERR_ARG, "rs_eval_circuit: coeff index out of range");
        RS_REQUIRE(t.n_w >= 0 && t.n_w <= RS_TERM_MAX_W, RS_ERR_ARG, "rs_eval_circuit: too many witness factors");
        RS_REQUIRE(t.pw < 0 || (pw && (size_t)t.pw < pw->n_polys), RS_ERR_ARG, "rs_eval_circuit: public witness index out of range");
        for (int x = 0; x < t.n_w; x++)
            RS_REQUIRE(t.w[x] >= 0 && (size_t)t.w[x] < w->n_polys, RS_ERR_ARG, "rs_eval_circuit: witness index out of range");
        RS_REQUIRE(k == 0 || t.constraint >= terms[k - 1].constraint, RS_ERR_ARG, "rs_eval_circuit: terms must be grouped by constraint");
    }
    // one stream-ordered temporary for both tables; freed on every path
    const size_t tbytes = ((n_terms ? n_terms : 1) * sizeof(rs_term) + 15) & ~(size_t)15;
    char *d_tmp = nullptr;
    RS_CUDA_CHECK(cudaMallocAsync((void **)&d_tmp, tbytes + (n_coeffs ? n_coeffs : 1) * NL * 4, g_stream));
    rs_term *d_terms = (rs_term *)d_tmp;
    uint32_t *d_coeffs = (uint32_t *)(d_tmp + tbytes);
    int rc = RS_OK;
    if (cudaMemcpyAsync(d_terms, terms, n_terms * sizeof(rs_term), cudaMemcpyHostToDevice, g_stream) != cudaSuccess ||
        cudaMemcpyAsync(d_coeffs, coeffs, n_coeffs * NL * 4, cudaMemcpyHostToDevice, g_stream) != cudaSuccess)
        rc = set_error(RS_ERR_CUDA, "rs_eval_circuit: upload failed: %s", cudaGetErrorString(cudaGetLastError()));
    if (rc == RS_OK)
        rc = eval_circuit(out->f, elem_ptr(out, out_poly), pw ? pw->d : nullptr, w->d, d_terms, (int)n_terms, d_coeffs, (int)n_coeffs,
                          (const uint32_t *)batch_const, out->rank);
    cudaFreeAsync(d_tmp, g_stream);
    return rc;
}

int rs_vec_powers(rs_batch *out, size_t poly, size_t n, const uint64_t *c_mont) {
    RS_REQUIRE(out && c_mont, RS_ERR_ARG, "rs_vec_powers: null argument");
    RS_REQUIRE(poly < out->n_polys && n <= out->rank, RS_ERR_SHAPE, "rs_vec_powers: range out of bounds");
    if (n == 0) return RS_OK;
    return powers(out->f, elem_ptr(out, poly), (const uint32_t *)c_mont, n);
}

int rs_rand_encode_blind(rs_batch *b, size_t poly0, size_t n, size_t N, const uint64_t *blind_mont) {
    RS_REQUIRE(b && blind_mont, RS_ERR_ARG, "rs_rand_encode_blind: null argument");
    RS_REQUIRE(poly0 + n <= b->n_polys && N < b->rank, RS_ERR_SHAPE, "rs_rand_encode_blind: range out of bounds");
    if (n == 0) return RS_OK;
    const int NL = b->f->NL;
    uint32_t *d = nullptr;
    RS_CUDA_CHECK(cudaMallocAsync((void **)&d, n * NL * 4, g_stream));
    int rc = RS_OK;
    if (cudaMemcpyAsync(d, blind_mont, n * NL * 4, cudaMemcpyHostToDevice, g_stream) != cudaSuccess)
        rc = set_error(RS_ERR_CUDA, "rs_rand_encode_blind: upload failed");
    if (rc == RS_OK) rc = blind(b->f, elem_ptr(b, poly0), d, n, b->rank, N);
    cudaFreeAsync(d, g_stream);
    return rc;
}

int rs_quorem_vanishing(rs_batch *quo, size_t qp, rs_batch *rem, size_t rp, const rs_batch *p, size_t pp, size_t N) {
    RS_REQUIRE(quo && rem && p, RS_ERR_ARG, "rs_quorem_vanishing: null argument");
    RS_REQUIRE(quo->rank == p->rank && rem->rank == p->rank, RS_ERR_SHAPE, "inputs not consistent");
    RS_REQUIRE(qp < quo->n_polys && rp < rem->n_polys && pp < p->n_polys, RS_ERR_SHAPE, "rs_quorem_vanishing: poly out of bounds");
    RS_REQUIRE(N >= 1, RS_ERR_ARG, "rs_quorem_vanishing: N must be >= 1");
    RS_REQUIRE(elem_ptr(quo, qp) != elem_ptr(p, pp), RS_ERR_ARG, "rs_quorem_vanishing: quo must not alias p");
    return quorem(p->f, elem_ptr(quo, qp), elem_ptr(rem, rp), elem_ptr(p, pp), p->rank, N);
}

int rs_aut(rs_batch *out, size_t op, const rs_batch *p, size_t pp, size_t n, long idx, int is_ntt) {
    RS_REQUIRE(out && p, RS_ERR_ARG, "rs_aut: null argument");
    RS_REQUIRE(out->rank == p->rank, RS_ERR_SHAPE, "inconsistent input(s)");
    RS_REQUIRE(is_pow2(p->rank), RS_ERR_SHAPE, "rs_aut: rank must be a power of two");
    RS_REQUIRE(idx % 2 != 0, RS_ERR_ARG, "AutTo: idx must be odd");
    RS_REQUIRE(op + n <= out->n_polys && pp + n <= p->n_polys, RS_ERR_SHAPE, "rs_aut: poly range out of bounds");
    RS_REQUIRE(elem_ptr(out, op) != elem_ptr(p, pp), RS_ERR_ARG, "rs_aut: out must not alias p");
    const long twoN = 2 * (long)p->rank;
    long i = idx % twoN;
    if (i < 0) i += twoN;
    // idx^-1 mod 2N (idx odd): Newton iteration in Z/2^k
    unsigned long long inv = 1, ui = (unsigned long long)i;
    for (int k = 0; k < 6; k++) inv *= 2 - ui * inv;
    inv &= (unsigned long long)(twoN - 1);
    return aut(p->f, elem_ptr(out, op), elem_ptr(p, pp), ilog2(p->rank), ui, inv, is_ntt, n);
}

int rs_reverse_scale(rs_batch *out, size_t op, const rs_batch *v, size_t vp, const uint64_t *scale) {
    RS_REQUIRE(out && v && scale, RS_ERR_ARG, "rs_reverse_scale: null argument");
    RS_REQUIRE(out->rank == v->rank, RS_ERR_SHAPE, "inconsistent input(s)");
    RS_REQUIRE(op < out->n_polys && vp < v->n_polys, RS_ERR_SHAPE, "rs_reverse_scale: poly out of bounds");
    RS_REQUIRE(elem_ptr(out, op) != elem_ptr(v, vp), RS_ERR_ARG, "rs_reverse_scale: out must not alias v");
    return reverse_scale(v->f, elem_ptr(out, op), elem_ptr(v, vp), (const uint32_t *)scale, v->rank);
}

int rs_poly_evaluate_many(const rs_batch *const *batches, const size_t *polys, const size_t *lens, size_t n, const uint64_t *x, uint64_t *results);

int rs_poly_evaluate(const rs_batch *p, size_t poly, size_t n_elems, const uint64_t *x, uint64_t *result) {
    RS_REQUIRE(p && x && result, RS_ERR_ARG, "rs_poly_evaluate: null argument");
    RS_REQUIRE(poly < p->n_polys && n_elems <= p->rank, RS_ERR_SHAPE, "rs_poly_evaluate: range out of bounds");
    const rs_field *f = p->f;
    const int NL = f->NL;
    if (n_elems == 0) { memset(result, 0, 4 * NL); return RS_OK; }
    // up to 2^16 coefficients: one CTA, 256 chunks (evaluate_many_kernel)
    if (n_elems <= ((size_t)1 << 16)) return rs_poly_evaluate_many(&p, &poly, &n_elems, 1, x, result);
    // longer: every thread Horner-evaluates 64 coefficients, then the n / 64 partial values are themselves a polynomial in
    // x^64, evaluated by the same one-CTA kernel (the one-thread combine this replaces took 31 ms at 2^22 coefficients)
    const int CH = 64;
    const size_t n_chunks = (n_elems + CH - 1) / CH;
    uint32_t xch[RS_MAX_NL];
    uint32_t e[2] = {(uint32_t)CH, 0};
    host_pow(f, xch, (const uint32_t *)x, e, 1);
    struct Desc { const uint32_t *coeffs; size_t n; };
    uint8_t *d_buf = nullptr;
    const size_t part_bytes = (n_chunks * NL * 4 + 15) & ~(size_t)15;
    RS_CUDA_CHECK(cudaMallocAsync((void **)&d_buf, part_bytes + 16 + NL * 4, g_stream));
    uint32_t *d_partial = (uint32_t *)d_buf, *d_result = (uint32_t *)(d_buf + part_bytes + 16);
    const Desc h{d_partial, n_chunks};
    int rc = horner(f, nullptr, d_partial, elem_ptr(p, poly), n_elems, (const uint32_t *)x, xch, CH);
    cudaError_t e2 = cudaMemcpyAsync(d_buf + part_bytes, &h, sizeof(h), cudaMemcpyHostToDevice, g_stream);
    if (rc == RS_OK && e2 != cudaSuccess) rc = set_error(RS_ERR_CUDA, "rs_poly_evaluate: %s", cudaGetErrorString(e2));
    if (rc == RS_OK) rc = evaluate_many(f, d_result, d_buf + part_bytes, 1, xch);
    if (rc == RS_OK) {
        e2 = cudaMemcpyAsync(result, d_result, NL * 4, cudaMemcpyDeviceToHost, g_stream);
        if (e2 == cudaSuccess) e2 = cudaStreamSynchronize(g_stream);
        if (e2 != cudaSuccess) rc = set_error(RS_ERR_CUDA, "rs_poly_evaluate: %s", cudaGetErrorString(e2));
    }
    cudaFreeAsync(d_buf, g_stream);
    return rc;
}

int rs_poly_evaluate_many(const rs_batch *const *batches, const size_t *polys, const size_t *lens, size_t n, const uint64_t *x, uint64_t *results) {
    RS_REQUIRE(batches && polys && lens && x && results, RS_ERR_ARG, "rs_poly_evaluate_many: null argument");
    if (n == 0) return RS_OK;
    const rs_field *f = batches[0]->f;
    const int NL = f->NL;
    struct Desc { const uint32_t *coeffs; size_t n; };
    std::vector<Desc> h(n);
    for (size_t i = 0; i < n; i++) {
        RS_REQUIRE(batches[i] && batches[i]->f->NL == NL, RS_ERR_SHAPE, "rs_poly_evaluate_many: field mismatch");
        RS_REQUIRE(polys[i] < batches[i]->n_polys && lens[i] <= batches[i]->rank, RS_ERR_SHAPE, "rs_poly_evaluate_many: range out of bounds");
        h[i] = Desc{elem_ptr(batches[i], polys[i]), lens[i]};
    }
    uint8_t *d_buf = nullptr;
    const size_t desc_bytes = (n * sizeof(Desc) + 15) & ~(size_t)15;
    RS_CUDA_CHECK(cudaMallocAsync((void **)&d_buf, desc_bytes + n * NL * 4, g_stream));
    cudaError_t e = cudaMemcpyAsync(d_buf, h.data(), n * sizeof(Desc), cudaMemcpyHostToDevice, g_stream);   // pageable: staged before the call returns
    int rc = e == cudaSuccess ? evaluate_many(f, (uint32_t *)(d_buf + desc_bytes), d_buf, n, (const uint32_t *)x) : set_error(RS_ERR_CUDA, "rs_poly_evaluate_many: %s", cudaGetErrorString(e));
    if (rc == RS_OK) {
        e = cudaMemcpyAsync(results, d_buf + desc_bytes, n * NL * 4, cudaMemcpyDeviceToHost, g_stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(g_stream);
        if (e != cudaSuccess) rc = set_error(RS_ERR_CUDA, "rs_poly_evaluate_many: %s", cudaGetErrorString(e));
    }
    cudaFreeAsync(d_buf, g_stream);
    return rc;
}

#pragma GCC visibility pop
}  // extern "C"
