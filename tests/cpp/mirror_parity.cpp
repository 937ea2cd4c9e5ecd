// Parity test of the C++ host mirror (include/ringo.hpp) against the C oracle (oracle/field_oracle.h).
// Test infrastructure: built and run by tests/test_cpp_mirror.py.
//   mirror_parity --link                      no GPU: checks that the mirror compiles and the library loads
//   mirror_parity L limb0 .. limb(L-1)        GPU: runs the bigpoly operator set on modulus q and compares
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <vector>

#include "ringo.hpp"
#include "field_oracle.h"

using namespace ringo;
using u64 = uint64_t;

static int failures = 0;
#define EXPECT(cond, what) do { if (!(cond)) { std::printf("FAIL %s (%s:%d)\n", what, __FILE__, __LINE__); failures++; } } while (0)

template <typename F>
static bool panics(F f) {
    try { f(); } catch (const Panic &) { return true; }
    return false;
}

int main(int argc, char **argv) {
    if (argc >= 2 && std::string(argv[1]) == "--link") {
        std::printf("launches=%llu\n", (unsigned long long)rs_launch_count());
        std::printf("ok\n");
        return 0;
    }
    if (argc < 3) { std::printf("usage\n"); return 2; }
    const int L = std::atoi(argv[1]);
    std::vector<u64> q(L);
    for (int i = 0; i < L; i++) q[i] = std::strtoull(argv[2 + i], nullptr, 0);
    Init(0);
    bigpoly::Field F(q);
    orc_field of;
    EXPECT(orc_field_init(&of, q.data(), L) == 0, "oracle field");
    EXPECT(F.Limb() == L, "Limb()");

    const size_t N = 1024, NP = 6;
    std::mt19937_64 rng(12345);
    auto random_elems = [&](size_t n) {            // canonical values, then Montgomery form via the oracle
        std::vector<u64> v(n * L);
        for (size_t i = 0; i < n; i++) {
            u64 *e = &v[i * L];
            for (;;) {
                for (int k = 0; k < L; k++) e[k] = rng();
                int top = L - 1;
                while (top > 0 && q[top] == 0) top--;
                e[top] &= (q[top] == ~0ull) ? ~0ull : ((1ull << (64 - __builtin_clzll(q[top]))) - 1);
                for (int k = top + 1; k < L; k++) e[k] = 0;
                bool lt = false;
                for (int k = L - 1; k >= 0; k--) { if (e[k] != q[k]) { lt = e[k] < q[k]; break; } }
                if (lt) break;
            }
            orc_to_mont(&of, e, e);
        }
        return v;
    };
    std::vector<u64> host = random_elems(NP * N);
    bigpoly::Batch B(F, NP, N), O(F, NP, N);
    B.Upload(host.data(), 0, NP);
    auto poly = [&](const std::vector<u64> &h, size_t i) { return h.data() + i * N * L; };
    std::vector<u64> got(N * L), exp(N * L);
    auto same = [&](bigpoly::Batch &b, size_t i, const std::vector<u64> &e) {
        b.Download(got.data(), i, 1);
        return std::memcmp(got.data(), e.data(), N * L * 8) == 0;
    };

    // ---- transformers (ntt.go:98-136, :206-244)
    for (int neg = 0; neg < 2; neg++) {
        orc_ntt ot;
        EXPECT(orc_ntt_init(&ot, &of, 10, neg) == 0, "oracle ntt");
        bigpoly::CyclicTransformer ct(F, N);
        bigpoly::CyclotomicTransformer zt(F, N);
        const bigpoly::Transformer &t = neg ? (const bigpoly::Transformer &)zt : (const bigpoly::Transformer &)ct;
        t.FwdNTTTo(O, 0, B, 0, NP);
        for (size_t i = 0; i < NP; i++) { orc_ntt_fwd(&ot, exp.data(), poly(host, i)); EXPECT(same(O, i, exp), "FwdNTTTo"); }
        t.InvNTTTo(O, 0, O, 0, NP);
        for (size_t i = 0; i < NP; i++) { std::memcpy(exp.data(), poly(host, i), N * L * 8); EXPECT(same(O, i, exp), "InvNTTTo round trip"); }
        orc_ntt_free(&ot);
    }

    // ---- evaluator operator set (base_op.go:49-207)
    bigpoly::CyclicEvaluator ev(F, N);
    bigpoly::Poly p0(B, 0), p1(B, 1), out(O, 2), acc(O, 3);
    ev.AddTo(out, p0, p1);  orc_vec_add(&of, exp.data(), poly(host, 0), poly(host, 1), N); EXPECT(same(O, 2, exp), "AddTo");
    ev.SubTo(out, p0, p1);  orc_vec_sub(&of, exp.data(), poly(host, 0), poly(host, 1), N); EXPECT(same(O, 2, exp), "SubTo");
    ev.NegTo(out, p0);      orc_vec_neg(&of, exp.data(), poly(host, 0), N);                 EXPECT(same(O, 2, exp), "NegTo");
    std::vector<u64> c = random_elems(1);
    ev.ScalarMulTo(out, p0, c.data()); orc_vec_scalar_mul(&of, exp.data(), poly(host, 0), c.data(), N); EXPECT(same(O, 2, exp), "ScalarMulTo");
    EXPECT(panics([&] { ev.MulTo(out, p0, p1); }), "MulTo must panic outside the NTT domain");
    {
        // product through the NTT domain == oracle: NTT both, multiply, (MulAdd), inverse
        orc_ntt ot; orc_ntt_init(&ot, &of, 10, 0);
        bigpoly::Poly a(O, 4), b(O, 5);
        ev.NTTTo(a, p0); ev.NTTTo(b, p1);
        EXPECT(a.IsNTT && b.IsNTT, "NTTTo sets IsNTT");
        EXPECT(panics([&] { ev.NTTTo(out, a); }), "NTTTo must panic on an NTT-domain operand");
        EXPECT(panics([&] { ev.AddTo(out, a, p1); }), "AddTo must panic on mixed domains");
        ev.MulTo(out, a, b);
        std::vector<u64> ea(N * L), eb(N * L);
        orc_ntt_fwd(&ot, ea.data(), poly(host, 0)); orc_ntt_fwd(&ot, eb.data(), poly(host, 1));
        orc_vec_mul(&of, exp.data(), ea.data(), eb.data(), N);
        EXPECT(same(O, 2, exp), "MulTo");
        ev.MulAddTo(out, a, b);                      // out += a*b
        orc_vec_add(&of, exp.data(), exp.data(), exp.data(), N);
        EXPECT(same(O, 2, exp), "MulAddTo");
        ev.MulSubTo(out, a, b);                      // back to a*b
        orc_vec_mul(&of, exp.data(), ea.data(), eb.data(), N);
        EXPECT(same(O, 2, exp), "MulSubTo");
        ev.InvNTTTo(acc, out);
        orc_ntt_inv(&ot, exp.data(), exp.data());
        EXPECT(same(O, 3, exp) && !acc.IsNTT, "InvNTTTo");
        orc_ntt_free(&ot);
    }
    // ---- QuoRemByVanishing (cyclic.go:18-37): rank-N polynomial divided by X^(N/4) - 1
    {
        const size_t M = N / 4;
        bigpoly::Poly quo(O, 2), rem(O, 3);
        ev.QuoRemByVanishing(p0, M, quo, rem);
        std::vector<u64> eq(N * L), er(N * L);
        orc_quorem_vanishing(&of, eq.data(), er.data(), poly(host, 0), N, M);
        EXPECT(same(O, 2, eq), "QuoRemByVanishing quo");
        EXPECT(same(O, 3, er), "QuoRemByVanishing rem");
    }
    // ---- AutTo (cyclotomic.go:29-82) and Evaluate (poly.go:64-76)
    {
        bigpoly::CyclotomicEvaluator zev(F, N);
        bigpoly::Poly o2(O, 2);
        for (long idx : {3L, 5L, (long)(2 * N - 1)}) {
            zev.AutTo(o2, p0, idx);
            orc_aut(&of, exp.data(), poly(host, 0), N, idx, 0);
            EXPECT(same(O, 2, exp), "AutTo");
        }
        std::vector<u64> x = random_elems(1), ez(L);
        std::vector<u64> z = p0.Evaluate(x.data());
        orc_poly_evaluate(&of, ez.data(), poly(host, 0), N, x.data());
        EXPECT(z == ez, "Evaluate");
    }
    // ---- constructor panics (ntt.go:27-37)
    EXPECT(panics([&] { bigpoly::CyclicTransformer bad(F, 1000); }), "non power-of-two rank must panic");
    EXPECT(panics([&] { bigpoly::CyclicTransformer bad(F, (size_t)1 << 62); }), "unsupported rank must panic");

    Sync();
    std::printf(failures ? "failed %d\n" : "ok\n", failures);
    return failures ? 1 : 0;
}
