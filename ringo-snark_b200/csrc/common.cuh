// common.cuh -- handle definitions and error plumbing shared by the C-ABI translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>
#include <string>
#include <vector>

#include "../../include/ringo_cuda.h"
#include "field.cuh"

#define RS_MAX_NL 28

struct rs_field {
    int L64;
    int NL;          // 32-bit limbs
    int qs;          // modulus shape: 0 generic, 1: q == 1 mod 2^32, 2: q == 1 mod 2^64
    int lazy_ok;     // 3q 2^32 + 2q < 2^(32 (NL+1)): the lazy [0, 2q) NTT arithmetic is valid
    uint32_t q2[RS_MAX_NL];   // 2q
    uint32_t q[RS_MAX_NL];
    uint32_t qinv;   // -q^-1 mod 2^32
    uint32_t one[RS_MAX_NL];
    uint32_t r2[RS_MAX_NL];
};

struct rs_ntt {
    rs_field *f;
    int logN;
    int negacyclic;
    uint32_t *d_tw;      // N elements
    uint32_t *d_twinv;   // N elements
    std::vector<uint32_t> h_tw, h_twinv;
    uint32_t ninv[RS_MAX_NL];
    uint32_t ninv_w1[RS_MAX_NL];
    // reduced-radix lazy forward path (rr_field.cuh / ntt_rr.cuh); rr_W == 0: not available
    int rr_W, rr_QZ, rr_max_bound;
    uint32_t rr_q[12], rr_c3q[12], rr_q2[12], rr_one[12];
    uint32_t *d_tw_rr;       // N x 12 words: omega * 2^(29 W) mod q in 29-bit limbs
    uint32_t *d_scratch;     // lazy elements between passes (planes)
    size_t scratch_words;
};

struct rs_batch {
    rs_field *f;
    size_t n_polys;
    size_t rank;
    uint32_t *d;
    bool owned;
};

namespace rs {

// when non-null, the NTT launcher records an event after every pass (rs_ntt_profile)
struct PassProfile { cudaEvent_t ev[16]; int n; };
extern thread_local PassProfile *g_pass_profile;

int set_error(int code, const char *fmt, ...);
cudaStream_t current_stream();
void count_launch(unsigned n = 1);

template <int NL>
inline FieldParams<NL> make_params(const rs_field *f) {
    FieldParams<NL> F;
    for (int i = 0; i < NL; i++) { F.q[i] = f->q[i]; F.one[i] = f->one[i]; F.r2[i] = f->r2[i]; F.q2[i] = f->q2[i]; }
    F.qinv = f->qinv;
    return F;
}

// host-side generic Montgomery helpers on rs_field (setup only: twiddles, scalars)
void host_mul(const rs_field *f, uint32_t *z, const uint32_t *x, const uint32_t *y);
void host_add(const rs_field *f, uint32_t *z, const uint32_t *x, const uint32_t *y);
void host_pow(const rs_field *f, uint32_t *z, const uint32_t *x, const uint32_t *e, int e_limbs);
void host_inverse(const rs_field *f, uint32_t *z, const uint32_t *x);
void host_set_u64(const rs_field *f, uint32_t *z, uint64_t v);

// NTT launchers, one translation unit per limb count (ntt_l*.cu)
int ntt_run_l4(const rs_ntt *t, uint32_t *out, const uint32_t *in, size_t n_polys, bool inverse);
int ntt_run_l8(const rs_ntt *t, uint32_t *out, const uint32_t *in, size_t n_polys, bool inverse);
int ntt_run_l14(const rs_ntt *t, uint32_t *out, const uint32_t *in, size_t n_polys, bool inverse);
int ntt_run_l28(const rs_ntt *t, uint32_t *out, const uint32_t *in, size_t n_polys, bool inverse);
// forward transform on the reduced-radix lazy kernels (8-limb fields); RS_ERR_UNSUPPORTED if not applicable
int ntt_rr_run_l8(rs_ntt *t, uint32_t *out, const uint32_t *in, size_t n_polys);

}  // namespace rs

#define RS_CUDA_CHECK(expr)                                                                          \
    do {                                                                                             \
        cudaError_t _e = (expr);                                                                     \
        if (_e != cudaSuccess)                                                                       \
            return rs::set_error(RS_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                                 __FILE__, __LINE__);                                                \
    } while (0)

#define RS_REQUIRE(cond, code, msg)                              \
    do {                                                         \
        if (!(cond)) return rs::set_error((code), "%s", (msg));  \
    } while (0)
