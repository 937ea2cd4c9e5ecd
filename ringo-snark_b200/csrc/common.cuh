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
#include "ntt52.cuh"

#define RS_MAX_NL 28

struct rs_field {
    int L64;
    int NL;          // 32-bit limbs
    int qs;          // modulus shape: 0 generic, 1: q == 1 mod 2^32, 2: q == 1 mod 2^64
    int lazy_ok;     // 3q 2^32 + 2q < 2^(32 (NL+1)): the lazy [0, 2q) NTT arithmetic is valid
    int shoup_ok;    // 3.6 q < 2^(32 NL), q >= 2^(32 (NL-2) + 8), q == 1 mod 2^32: the Shoup NTT kernels are valid (shoup.cuh)
    uint32_t qinv_full[RS_MAX_NL];   // -q^-1 mod 2^(32 NL)
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
    // Shoup tables (shoup.cuh): N entries of 2 NL words (w, floor(w 2^(32 NL) / q)); null: Montgomery kernels
    uint32_t *d_tw_sh, *d_twinv_sh;
    uint32_t ninv_sh[2 * RS_MAX_NL], ninv_w1_sh[2 * RS_MAX_NL];   // the same pairs for N^-1 and N^-1 twinv[1]
    // FP64-pipe path (ntt52.cuh), 8-word fields only: twiddles re-scaled to the 2^260 Montgomery radix, 52-bit limbs
    // as doubles, TW52_STRIDE doubles per entry; has52 == 0: not available for this field
    int has52, variant52;
    unsigned mix_pf, mix_p;   // share of the CTAs that take FP64 tiles in the two-pipe kernels (0: off)
    double *d_tw52, *d_twinv52;
    rs::Ntt52Consts<8, 5> c52;
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
// One-time per (device, kernel) opt-in to more than 48 KB of dynamic shared memory.  The attribute is per device, so
// the cache is keyed by the current device (rs_init may bind different threads to different GPUs); mutex-guarded.
cudaError_t ensure_smem_attr(const void *func, int bytes);
// true exactly once per (current device, tag): per-device uploads such as __constant__ tables
bool first_use_on_device(const void *tag);

// NCCL collectives on the current stream (nccl_glue.cu; libnccl is loaded at run time)
int nccl_allgather_inplace_u64(void *comm, uint64_t *d_buf, size_t count, int rank);
int nccl_allreduce_sum_u64(void *comm, uint64_t *d_buf, size_t count);

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
int ntt_run_sh_l8(const rs_ntt *t, uint32_t *out, const uint32_t *in, size_t n_polys, bool inverse);   // Shoup kernels (ntt_sh_l8.cu)
int ntt_run_l14(const rs_ntt *t, uint32_t *out, const uint32_t *in, size_t n_polys, bool inverse);
int ntt_run_l28(const rs_ntt *t, uint32_t *out, const uint32_t *in, size_t n_polys, bool inverse);
// FP64-pipe kernels for the 8-word fields, one translation unit per tuning variant (ntt52_v*.cu)
int ntt52_run_v2(const rs_ntt *t, uint32_t *out, const uint32_t *in, size_t n_polys, bool inverse);
// both pipes in one launch (ntt_mix_l8.cu); RS_ERR_UNSUPPORTED: use the single-pipe kernels
int ntt_mix_run_l8(const rs_ntt *t, uint32_t *out, const uint32_t *in, size_t n_polys, bool inverse);

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
