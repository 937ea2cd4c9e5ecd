// jindo_params.cpp -- jindo.NewParameters (jindo/params.go:18-321) and the host-library conventions a caller without
// lattigo needs, as compiled host code behind the C ABI (rs_jindo_new_parameters, include/ringo_jindo.h):
//   encodeParameters            q = b^k + 1  ->  (b, k)                                         params.go:18-40
//   the parameter search        float64 formulas in the reference's order                       params.go:126-321
//   RNS primes                  [lattigo v6.1.0] ring.NewNTTFriendlyPrimesGenerator(bits, 2d).NextUpstreamPrimes, restated
//   psi                         [lattigo] SubRing: g^((q-1)/2d), g the smallest primitive root of Z_q^*, restated
//   Encoder.deltaInv            -(b^i) / q rounded to float64, thresholded                      encoder.go:50-67
// Built with -ffp-contract=off like the samplers (Go on amd64 never fuses a*b+c).  The lattigo conventions are unpinned
// (DESIGN.md section 4): they are what oracle/jindo_ref.py restates too, and what a Go caller would take from lattigo itself.
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "../../include/ringo_jindo.h"
#include "../../include/ringo_sampler.h"

namespace rs { int set_error(int code, const char *fmt, ...); }

namespace {

typedef unsigned __int128 u128;

// ---- little multi-precision helpers on 32-bit limbs (values up to ~2000 bits, setup only) -----------------------------
typedef std::vector<uint32_t> Big;

void trim(Big &a) { while (a.size() > 1 && a.back() == 0) a.pop_back(); }
int cmp(const Big &a, const Big &b) {
    if (a.size() != b.size()) return a.size() < b.size() ? -1 : 1;
    for (size_t i = a.size(); i-- > 0;) if (a[i] != b[i]) return a[i] < b[i] ? -1 : 1;
    return 0;
}
Big mul(const Big &a, const Big &b) {
    Big r(a.size() + b.size(), 0);
    for (size_t i = 0; i < a.size(); i++) {
        uint64_t carry = 0;
        for (size_t j = 0; j < b.size(); j++) {
            const uint64_t t = (uint64_t)a[i] * b[j] + r[i + j] + carry;
            r[i + j] = (uint32_t)t; carry = t >> 32;
        }
        r[i + b.size()] += (uint32_t)carry;
    }
    trim(r);
    return r;
}
Big from_u64(uint64_t v) { Big r{(uint32_t)v, (uint32_t)(v >> 32)}; trim(r); return r; }
Big pow_big(uint64_t base, int e) {
    Big r = from_u64(1), b = from_u64(base);
    for (; e; e >>= 1) { if (e & 1) r = mul(r, b); if (e > 1) b = mul(b, b); }
    return r;
}
void shl1(Big &a) {
    uint32_t c = 0;
    for (size_t i = 0; i < a.size(); i++) { const uint32_t n = a[i] >> 31; a[i] = (a[i] << 1) | c; c = n; }
    if (c) a.push_back(c);
}
void sub_in_place(Big &a, const Big &b) {      // a >= b
    int64_t borrow = 0;
    for (size_t i = 0; i < a.size(); i++) {
        int64_t t = (int64_t)a[i] - (i < b.size() ? b[i] : 0) - borrow;
        borrow = t < 0;
        a[i] = (uint32_t)(t + (borrow ? (1ll << 32) : 0));
    }
    trim(a);
}
bool is_zero(const Big &a) { return a.size() == 1 && a[0] == 0; }

// ---- 64-bit number theory ---------------------------------------------------------------------------------------------
uint64_t mulmod(uint64_t a, uint64_t b, uint64_t m) { return (uint64_t)((u128)a * b % m); }
uint64_t powmod(uint64_t a, uint64_t e, uint64_t m) {
    uint64_t r = 1 % m;
    for (a %= m; e; e >>= 1) { if (e & 1) r = mulmod(r, a, m); a = mulmod(a, a, m); }
    return r;
}
bool is_prime64(uint64_t n) {
    if (n < 2) return false;
    static const uint64_t small[] = {2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37};
    for (uint64_t p : small) if (n % p == 0) return n == p;
    uint64_t d = n - 1;
    int s = 0;
    while ((d & 1) == 0) { d >>= 1; s++; }
    for (uint64_t a : small) {                                   // deterministic for n < 3.3e24
        uint64_t x = powmod(a, d, n);
        if (x == 1 || x == n - 1) continue;
        bool comp = true;
        for (int r = 1; r < s; r++) { x = mulmod(x, x, n); if (x == n - 1) { comp = false; break; } }
        if (comp) return false;
    }
    return true;
}
uint64_t gcd64(uint64_t a, uint64_t b) { while (b) { const uint64_t t = a % b; a = b; b = t; } return a; }
void factor(uint64_t n, std::vector<uint64_t> &fs) {
    for (uint64_t d = 2; d * d <= n && d < (1u << 20); d += (d == 2 ? 1 : 2))
        if (n % d == 0) { fs.push_back(d); while (n % d == 0) n /= d; }
    std::vector<uint64_t> stack;
    if (n > 1) stack.push_back(n);
    while (!stack.empty()) {
        const uint64_t m = stack.back();
        stack.pop_back();
        if (m == 1) continue;
        if (is_prime64(m)) { fs.push_back(m); continue; }
        for (uint64_t c = 1;; c++) {                             // Pollard rho
            uint64_t x = 2, y = 2, g = 1;
            while (g == 1) {
                x = (mulmod(x, x, m) + c) % m;
                y = (mulmod(y, y, m) + c) % m; y = (mulmod(y, y, m) + c) % m;
                g = gcd64(x > y ? x - y : y - x, m);
            }
            if (g != m) { stack.push_back(g); stack.push_back(m / g); break; }
        }
    }
    std::sort(fs.begin(), fs.end());
    fs.erase(std::unique(fs.begin(), fs.end()), fs.end());
}
// [lattigo] ring.PrimitiveRoot: the smallest generator of Z_q^*
uint64_t primitive_root(uint64_t q) {
    std::vector<uint64_t> fs;
    factor(q - 1, fs);
    for (uint64_t g = 2;; g++) {
        bool ok = true;
        for (uint64_t f : fs) if (powmod(g, (q - 1) / f, q) == 1) { ok = false; break; }
        if (ok) return g;
    }
}
// [lattigo] NTTFriendlyPrimesGenerator.NextUpstreamPrimes: ascending primes 2^bits + 1 + k * nth_root
bool next_upstream_primes(int bits, uint64_t nth_root, int count, uint64_t *out) {
    if (bits < 2 || bits > 61) return false;
    uint64_t c = (1ull << bits) + 1;
    int n = 0;
    while (n < count) {
        if (c >> (bits + 1)) return false;                       // ran out of primes below 2^(bits+1)
        if (is_prime64(c)) out[n++] = c;
        c += nth_root;
    }
    return true;
}

// encodeParameters (params.go:18-40): repeated integer square roots of q - 1 while they are exact
bool encode_parameters(const Big &qm1, uint64_t &base, int &exp) {
    for (int e = 1024; e >= 1; e >>= 1) {
        // the e-th root, if it is an integer below 2^64: binary search on r^e <=> q - 1
        uint64_t lo = 1, hi = ~0ull;
        const size_t bits = (qm1.size() - 1) * 32 + (32 - __builtin_clz(qm1.back()));
        if ((bits + e - 1) / e < 64) hi = (1ull << ((bits + e - 1) / e)) + 1;
        if (bits < (size_t)e) continue;
        while (lo < hi) {
            const uint64_t mid = lo + (hi - lo + 1) / 2;
            if (cmp(pow_big(mid, e), qm1) <= 0) lo = mid; else hi = mid - 1;
        }
        if (cmp(pow_big(lo, e), qm1) == 0) { base = lo; exp = e; return true; }
    }
    return false;
}

int find_msis_rank(double d, double q, double beta, bool &ok) {
    if (beta > q) { ok = false; return 0; }
    const double log_beta = log2(beta), log_q = log2(q), log_delta = log2(1.005);
    return (int)ceil((log_beta * log_beta) / (4 * d * log_q * log_delta));
}

// -(b^i) / q as the nearest float64 (the reference converts the exact rational, encoder.go:56-63)
double neg_pow_over_q(uint64_t base, int i, const Big &q) {
    Big r = pow_big(base, i);                                    // 0 < b^i < q
    uint64_t mant = 0;
    int got = 0, shift = 0;                                      // value = 0.b1 b2 b3 ... ; shift = position of the bit just produced
    bool started = false;
    while (got < 55) {
        shl1(r);
        shift++;
        int bit = 0;
        if (cmp(r, q) >= 0) { sub_in_place(r, q); bit = 1; }
        if (bit) started = true;
        if (started) { mant = (mant << 1) | (uint64_t)bit; got++; }
        if (shift > 4000) return 0.0;
    }
    const bool sticky = !is_zero(r);
    // mant holds 55 significant bits, the last at weight 2^-shift: round to 53 (nearest, ties to even; sticky breaks ties upward)
    const uint64_t low = mant & 3, keep = mant >> 2;
    uint64_t rounded = keep;
    if (low > 2 || (low == 2 && (sticky || (keep & 1)))) rounded++;
    return -ldexp((double)rounded, -(shift - 2));
}

const int RLWE_RANK = 1 << 13, MAX_LOG_Q = 240, ETA = 6, TAILCUT = 5, TAILCUT_CSPRNG = 9;

}  // namespace

#pragma GCC visibility push(default)
extern "C" {

int rs_jindo_new_parameters(const uint64_t *q_le, int L64, uint64_t target_n, int batch, rs_jindo_derived *out) {
    if (!q_le || !out || L64 < 1 || L64 > 16) return rs::set_error(RS_ERR_ARG, "rs_jindo_new_parameters: bad argument");
    if (target_n < 1 || batch < 1) return rs::set_error(RS_ERR_ARG, "NewParameters: bad arguments");
    memset(out, 0, sizeof(*out));
    Big q(2 * (size_t)L64), qm1;
    for (int i = 0; i < L64; i++) { q[2 * i] = (uint32_t)q_le[i]; q[2 * i + 1] = (uint32_t)(q_le[i] >> 32); }
    trim(q);
    qm1 = q;
    qm1[0] -= 1;                                                  // q is odd
    uint64_t base = 0;
    int exp = 0;
    if (!encode_parameters(qm1, base, exp) || exp > RS_JINDO_MAX_EXP) return rs::set_error(RS_ERR_UNSUPPORTED, "modulus not jindo-friendly");
    const double t = (double)batch, b = (double)base, k = (double)exp;
    const double d = k > 256 ? k : 256, l = d / k;
    const double nu = RLWE_RANK / d;
    const uint64_t max_cols = (uint64_t)ceil((double)target_n / l);
    double min_size = INFINITY;
    bool found = false;
    for (uint64_t nn = 1; nn <= max_cols; nn <<= 1) {
        const double n = (double)nn;
        const double m = ceil((double)target_n / (n * l));
        const double x_one = sqrt(k) * b;
        const double c_one = sqrt(k) * fmin(b, pow(2.0, 120 / k)) / 2;
        const double ecd_std = 2 / (b - 1) * (b + 1) * ETA;
        const double ecd_blind_std = 2 * x_one / (b - 1) * (b + 1) * ETA;
        const double mask_std = 2 * c_one / (b - 1) * (b + 1) * ETA;
        const double mask_blind_std = 2 * c_one * x_one / (b - 1) * (b + 1) * ETA;
        const double mlwe_std = 2 * sqrt(2.0) * ETA;
        const double mask_mlwe_std = 2 * c_one * sqrt(2.0) * ETA;
        const double fij_inf = TAILCUT * (b + 1) * ecd_std;
        const double f0j_inf = TAILCUT * (b + 1) * sqrt(m + 1) * ecd_blind_std;
        const double fin_inf = TAILCUT * (b + 1) * sqrt(n + 1) * mask_std;
        const double f0n_inf = TAILCUT * (b + 1) * sqrt((m + 1) * n + 1) * mask_blind_std;
        double res_ecdi_inf = sqrt(n) * c_one * fij_inf + fin_inf;
        double res_ecd0_inf = sqrt(n) * c_one * f0j_inf + f0n_inf;
        double pr_inf = sqrt(m) * x_one * fij_inf + f0j_inf;
        if (t > 1) {
            res_ecdi_inf *= sqrt(t) * c_one;
            res_ecd0_inf *= sqrt(t) * c_one;
            pr_inf *= sqrt(t) * c_one;
        }
        const double res_ecd_two = sqrt(d * (m * res_ecdi_inf * res_ecdi_inf + res_ecd0_inf * res_ecd0_inf));
        const double mlwe_inf = TAILCUT * mlwe_std;
        const double mask_mlwe_inf = TAILCUT * sqrt(n + 1) * mask_mlwe_std;
        double res_mlwe_inf = sqrt(n) * c_one * mlwe_inf + mask_mlwe_inf;
        if (t > 1) res_mlwe_inf *= sqrt(t) * c_one;
        double res_two = 0, in_cut_two = 0, in_msis_beta = 0, qf = 0, c_ext_one = 0, d_ext_one = 0, in_msis_rank = 0;
        bool ok = true;
        for (int mu = 1;; mu++) {
            if (mu > 1 << 20) { ok = false; break; }
            const double res_mlwe_two = sqrt(d * ((double)mu + nu)) * res_mlwe_inf;
            res_two = sqrt(res_ecd_two * res_ecd_two + res_mlwe_two * res_mlwe_two);
            in_cut_two = res_two;
            double ext_beta;
            if (t == 1) { ext_beta = 2 * (res_two + in_cut_two); c_ext_one = 2 * c_one; d_ext_one = 1.0; }
            else { ext_beta = 2 * (2 * c_one) * (res_two + in_cut_two); c_ext_one = (2 * c_one) * (2 * c_one); d_ext_one = 2 * c_one; }
            in_msis_beta = 2 * d_ext_one * c_ext_one * ext_beta;
            const double log_q = ceil(log2(in_msis_beta));
            const int q_limbs = (int)ceil(log_q / 60.0);
            const int q_bits = (int)ceil(log_q / (double)q_limbs);
            qf = pow(2.0, (double)(q_bits * q_limbs));
            if (log2(qf) > MAX_LOG_Q) continue;
            bool fok = true;
            if (find_msis_rank(d, qf, in_msis_beta, fok) == mu && fok) { in_msis_rank = (double)mu; break; }
            if (!fok) { ok = false; break; }
        }
        if (!ok) return rs::set_error(RS_ERR_UNSUPPORTED, "NewParameters: findMSISRank: beta > q");
        double in_cut_inf = in_cut_two / ((1 + sqrt(n) * c_one) * sqrt(in_msis_rank * d));
        if (t > 1) in_cut_inf /= sqrt(t) * c_one;
        double in_dcmp_inf = qf / in_cut_inf;
        if (t > 1) in_dcmp_inf *= sqrt(t) * c_one;
        const double in_dcmp_two = sqrt((n + 1) * in_msis_rank * d) * in_dcmp_inf;
        const double out_cut_two = in_dcmp_two;
        const double out_msis_beta = 2 * d_ext_one * (2 * (in_dcmp_two + out_cut_two));
        const double log_qq = ceil(log2(out_msis_beta));
        const int qq_limbs = (int)ceil(log_qq / 60.0);
        const int qq_bits = (int)ceil(log_qq / (double)qq_limbs);
        const double qq = pow(2.0, (double)(qq_bits * qq_limbs));
        if (log2(qq) > MAX_LOG_Q) continue;
        bool fok = true;
        const double out_msis_rank = (double)find_msis_rank(d, qq, out_msis_beta, fok);
        if (!fok) return rs::set_error(RS_ERR_UNSUPPORTED, "NewParameters: findMSISRank: beta > q");
        double out_cut_inf = out_cut_two / sqrt(out_msis_rank * d);
        if (t > 1) out_cut_inf /= sqrt(t) * c_one;
        const double com_size = t * out_msis_rank * d * log2(qq / out_cut_inf);
        double pf_size = 0.0;
        pf_size += n * d * log2(pr_inf);
        pf_size += d * log2(qf);
        pf_size += m * d * log2(res_ecdi_inf);
        pf_size += d * log2(res_ecd0_inf);
        pf_size += (in_msis_rank + nu) * d * log2(res_mlwe_inf);
        pf_size += ((n + 1) * in_msis_rank * d) * log2(in_dcmp_inf);
        if (com_size + pf_size < min_size) {
            min_size = com_size + pf_size;
            found = true;
            rs_jindo_derived &o = *out;
            memset(&o, 0, sizeof(o));
            o.batch = batch;
            o.rank = (uint64_t)n * (uint64_t)m * (uint64_t)l;
            o.p.rows = (int)m + 1; o.p.cols = (int)n;
            o.p.base = base; o.p.exp = exp; o.p.d = (int)d; o.p.slots = (int)d / exp;
            o.p.in_msis = (int)in_msis_rank; o.p.out_msis = (int)out_msis_rank; o.p.mlwe = (int)nu;
            o.p.log_in_cut = (int)floor(log2(in_cut_inf));
            o.p.log_out_cut = (int)floor(log2(out_cut_inf));
            o.in_com_dcmp_len = (int)((n + 1) * in_msis_rank);
            const int ql = (int)ceil(log2(qf) / 60), qb = (int)ceil(log2(qf) / (double)ql);
            const int qql = (int)ceil(log2(qq) / 60), qqb = (int)ceil(log2(qq) / (double)qql);
            if (ql > RS_JINDO_MAX_LIMBS || qql > RS_JINDO_MAX_LIMBS) return rs::set_error(RS_ERR_UNSUPPORTED, "NewParameters: more than 4 RNS limbs");
            o.p.n_q = ql; o.p.n_qout = qql; o.q_bits = qb; o.qout_bits = qqb;
            if (!next_upstream_primes(qb, 2 * (uint64_t)d, ql, o.p.q) || !next_upstream_primes(qqb, 2 * (uint64_t)d, qql, o.p.qout))
                return rs::set_error(RS_ERR_UNSUPPORTED, "NewParameters: ran out of NTT-friendly primes");
            const double s2pi = sqrt(2 * M_PI);
            o.ecd_std = ecd_std / s2pi; o.noise.ecd_blind_std = ecd_blind_std / s2pi;
            o.noise.mask_std = mask_std / s2pi; o.noise.mask_blind_std = mask_blind_std / s2pi;
            o.mlwe_std = mlwe_std / s2pi; o.noise.mask_mlwe_std = mask_mlwe_std / s2pi;
            o.res_two_nm = res_two + in_cut_two;
            o.in_com_dcmp_two_nm = in_dcmp_two + out_cut_two;
            o.com_size = com_size; o.pf_size = pf_size;
        }
    }
    if (!found) return rs::set_error(RS_ERR_UNSUPPORTED, "NewParameters: no parameter set below the modulus cap");
    rs_jindo_derived &o = *out;
    for (int i = 0; i < o.p.n_q; i++) o.p.psi_q[i] = powmod(primitive_root(o.p.q[i]), (o.p.q[i] - 1) / (2 * (uint64_t)o.p.d), o.p.q[i]);
    for (int i = 0; i < o.p.n_qout; i++) o.p.psi_qout[i] = powmod(primitive_root(o.p.qout[i]), (o.p.qout[i] - 1) / (2 * (uint64_t)o.p.d), o.p.qout[i]);
    // Encoder.deltaInv (encoder.go:50-67)
    const double thr = pow(2.0, -50) / ((double)base * (double)exp);
    for (int i = 0; i < exp; i++) {
        const double v = neg_pow_over_q(base, i, q);
        o.p.delta_inv[i] = fabs(v) < thr ? 0.0 : v;
    }
    // TwinCDT tables of the encoder (sigma = ecdStdDev): geometry here, contents by rs_jindo_cdt_tables
    const int64_t tail_hi = (int64_t)ceil(TAILCUT_CSPRNG * o.ecd_std);
    o.p.cdt_tail_lo = -tail_hi;
    o.p.cdt_len = (int32_t)(2 * tail_hi + 1);
    o.p.cdt_sigma = o.ecd_std;
    return RS_OK;
}

int rs_jindo_cdt_tables(const rs_jindo_derived *d, uint64_t *tables) {
    if (!d || !tables) return rs::set_error(RS_ERR_ARG, "rs_jindo_cdt_tables: null argument");
    for (int i = 0; i < 128; i++) {                               // NewTwinCDTGaussianSampler: centres i / 128 (gaussian_twin_cdt.go:47-60)
        int64_t lo = 0;
        const int n = rs_compute_cdt((double)i / 128, d->p.cdt_sigma, tables + (size_t)i * d->p.cdt_len, d->p.cdt_len, &lo);
        if (n != d->p.cdt_len) return rs::set_error(RS_ERR_ARG, "rs_jindo_cdt_tables: table geometry mismatch");
    }
    return RS_OK;
}

}  // extern "C"
#pragma GCC visibility pop
