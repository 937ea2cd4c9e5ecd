// Host build of ringo-snark_b200/csrc/field52.cuh and of the butterfly rounds of ntt52.cuh: the exact
// fused-multiply-add sequence the sm_100a kernels execute on the FP64 pipe, run under FE_TOWARDZERO so that
// fma() rounds like __fma_rz.  Checked against Python integers by tests/test_device_field52_host.py.
// Test infrastructure only.
#include <cfenv>
#include <cstddef>
#include <cstring>

#include "../ringo-snark_b200/csrc/ntt52.cuh"

using namespace rs;

static Ntt52Consts<8, 5> g_C;

struct RZ {
    int old;
    RZ() : old(fegetround()) { fesetround(FE_TOWARDZERO); }
    ~RZ() { fesetround(old); }
};

extern "C" int hf52_setup(const uint32_t *q, const long long *ninv, const long long *ninv_w1, int *info /* q0one, max_lazy */) {
    if (!f52_setup<8, 5>(g_C.F, q)) return -1;
    for (int k = 0; k < 5; k++) { g_C.ninv[k] = (double)ninv[k]; g_C.ninv_w1[k] = (double)ninv_w1[k]; }
    memcpy(g_C.qw, q, 32);
    info[0] = g_C.F.q0one; info[1] = g_C.F.max_lazy;
    return 0;
}

// op 0: r = mont52(a, b) (a: normalised limbs, b: normalised limbs of a value < q); generic = 1 forces the q[0] != 1 path
extern "C" int hf52_mont(int generic, const long long *a, const long long *b, long long *r, size_t n) {
    RZ rz;
    for (size_t i = 0; i < n; i++) {
        double ad[5], bd[5];
        long long out[5];
        for (int k = 0; k < 5; k++) { ad[k] = (double)a[5 * i + k]; bd[k] = (double)b[5 * i + k]; }
        if (!generic && g_C.F.q0one) mont52<5, true>(out, ad, bd, g_C.F); else mont52<5, false>(out, ad, bd, g_C.F);
        memcpy(r + 5 * i, out, 40);
    }
    return 0;
}

extern "C" int hf52_from_words(const uint32_t *w, long long *l, size_t n) {
    for (size_t i = 0; i < n; i++) { long long t[5]; f52_from_words<8, 5>(t, w + 8 * i); memcpy(l + 5 * i, t, 40); }
    return 0;
}
extern "C" int hf52_to_words(const long long *l, uint32_t *w, size_t n) {
    for (size_t i = 0; i < n; i++) { long long t[5]; memcpy(t, l + 5 * i, 40); f52_to_words<8, 5>(w + 8 * i, t); }
    return 0;
}
extern "C" int hf52_reduce(long long *l, size_t n) {
    for (size_t i = 0; i < n; i++) { long long t[5]; memcpy(t, l + 5 * i, 40); f52_reduce<5>(t, g_C.F); memcpy(l + 5 * i, t, 40); }
    return 0;
}
// the store path of move_tile52: reduce -> words -> (last) two conditional subtractions
extern "C" int hf52_finish(const long long *l, uint32_t *w, int last, size_t n) {
    for (size_t i = 0; i < n; i++) {
        long long t[5]; memcpy(t, l + 5 * i, 40);
        f52_reduce<5>(t, g_C.F);
        f52_to_words<8, 5>(w + 8 * i, t);
        if (last) { cond_sub<8>(w + 8 * i, g_C.qw); cond_sub<8>(w + 8 * i, g_C.qw); }
    }
    return 0;
}

template <int RR, bool INV, bool SCALE, bool TRIV>
static void round_t(long long *x, const double *tw, int s_top, unsigned hi, size_t n) {
    for (size_t i = 0; i < n; i++) {
        long long e[1 << RR][5];
        memcpy(e, x + i * (5 << RR), sizeof(e));
        if (g_C.F.q0one) radix_round52<5, true, RR, INV, SCALE, TRIV>(e, tw, s_top, hi, g_C);
        else radix_round52<5, false, RR, INV, SCALE, TRIV>(e, tw, s_top, hi, g_C);
        memcpy(x + i * (5 << RR), e, sizeof(e));
    }
}

// x: n groups of 2^RR elements (5 limbs each), transformed in place by one radix-2^RR round
extern "C" int hf52_round(int RR, int inv, int scale, int triv, long long *x, const double *tw, int s_top, unsigned hi, size_t n) {
    RZ rz;
#define RD(R, I, S, T) if (RR == R && inv == I && scale == S && triv == T) { round_t<R, I, S, T>(x, tw, s_top, hi, n); return 0; }
    RD(3, 0, 0, 0) RD(3, 0, 0, 1) RD(3, 1, 0, 0) RD(3, 1, 1, 0) RD(3, 1, 1, 1)
    RD(2, 0, 0, 0) RD(2, 1, 0, 0) RD(2, 1, 1, 0) RD(1, 0, 0, 0) RD(1, 1, 0, 0) RD(1, 1, 1, 0)
    RD(2, 0, 0, 1) RD(2, 1, 1, 1)
#undef RD
    return -1;
}
