/* Host check of the FP64 modular product used by the warp NTT of encode_kernel (csrc/jindo.cu: f64_bfly):
 *   h = rn(x*w); l = fma(x, w, -h); b = rint(h / q); r = fma(-b, q, h) + l
 * must give r == x*w (mod q) with |r| < q for every integer |x| < 9q, 0 <= w < q, q < 2^44 -- compared with exact
 * 128-bit integer arithmetic.  Same IEEE-754 operations as the device intrinsics (__dmul_rn, __fma_rn, __dadd_rn).
 * Test infrastructure (compiled by tests/test_host_f64_bfly.py). */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

static uint64_t rng_state = 0x9e3779b97f4a7c15ull;
static uint64_t rnd(void) { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return rng_state; }

static int check(uint64_t q, int64_t x, uint64_t w) {
    const double qf = (double)q, qinv = 1.0 / qf, xf = (double)x, wf = (double)w;
    const double h = xf * wf;
    const double l = fma(xf, wf, -h);
    const double b = (h * qinv + 6755399441055744.0) - 6755399441055744.0;
    const double r = fma(-b, qf, h) + l;
    if (r != floor(r) || fabs(r) >= qf) return 1;
    __int128 exact = (__int128)x * (__int128)w;
    __int128 m = exact % (__int128)q; if (m < 0) m += q;
    __int128 rr = (__int128)(int64_t)r % (__int128)q; if (rr < 0) rr += q;
    return m != rr;
}

int main(int argc, char **argv) {
    /* primes == 1 (mod 512): the largest below 2^44, the 2^22 commit's 38-bit primes, a 35-bit one */
    const uint64_t primes[] = {17592186037249ull, 274877908993ull, 274877916673ull, 34359753217ull};
    const long n = argc > 1 ? atol(argv[1]) : 2000000;
    for (unsigned p = 0; p < sizeof(primes) / sizeof(primes[0]); p++) {
        const uint64_t q = primes[p];
        const int64_t lim = 9 * (int64_t)q - 1;
        const int64_t xs[] = {lim, -lim, (int64_t)q, -(int64_t)q, 1, -1, 0, (int64_t)(q / 2), (int64_t)(q / 2) + 1, 8 * (int64_t)q + (int64_t)(q / 2)};
        const uint64_t ws[] = {q - 1, 1, 0, q / 2, q / 2 + 1, 2, q - 2};
        for (unsigned i = 0; i < sizeof(xs) / sizeof(xs[0]); i++)
            for (unsigned j = 0; j < sizeof(ws) / sizeof(ws[0]); j++)
                if (check(q, xs[i], ws[j])) { printf("FAIL q=%llu x=%lld w=%llu\n", (unsigned long long)q, (long long)xs[i], (unsigned long long)ws[j]); return 1; }
        for (long k = 0; k < n; k++) {
            const int64_t x = (int64_t)(rnd() % (uint64_t)(2 * lim + 1)) - lim;
            const uint64_t w = rnd() % q;
            if (check(q, x, w)) { printf("FAIL q=%llu x=%lld w=%llu\n", (unsigned long long)q, (long long)x, (unsigned long long)w); return 1; }
        }
    }
    printf("ok\n");
    return 0;
}
