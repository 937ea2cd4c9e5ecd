// ringo.hpp -- C++17 host-side mirror of the reference's operator interface for the hot path, written
// over the C ABI (ringo_cuda.h / ringo_jindo.h).  Header only; link with libringo_b200.so.
//
// The reference is Go and has no Go toolchain on the build / GPU boxes, so this is the compiled-language
// host side a C++ caller uses; the cgo shim of INTEGRATION.md binds the same C entry points.  Names,
// argument order and failure behaviour follow the Go API:
//
//   ringo::bigpoly::Poly / CyclicTransformer / CyclotomicTransformer   math/bigpoly/poly.go, ntt.go:26-244
//   ringo::bigpoly::CyclicEvaluator / CyclotomicEvaluator              math/bigpoly/base_op.go:42-207,
//                                                                      cyclic.go:9-37, cyclotomic.go:14-82
//   ringo::jindo::Prover                                               jindo/prover.go:28-324, verifier.go:50-282
//   ringo::buckler::Prover                                             buckler/prover.go:42-353
//
// Differences that follow from the boundary: polynomials live in device-resident
// batches (`Batch`), a `Poly` is a (batch, index, IsNTT) view, scalars are L64-limb Montgomery values on
// the host.  Where the Go code panics (rank mismatch, NTT-domain mismatch, unsupported rank) these
// classes throw ringo::Panic carrying the library's message.
#ifndef RINGO_HPP
#define RINGO_HPP

#include <cstddef>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "ringo_buckler.h"
#include "ringo_cuda.h"
#include "ringo_jindo.h"

namespace ringo {

struct Panic : std::runtime_error {
    int code;
    Panic(int c, const std::string &what) : std::runtime_error(what), code(c) {}
};

inline void check(int rc) {
    if (rc != 0) throw Panic(rc, rs_last_error());
}

// rs_init / stream plumbing
inline void Init(int device = 0) { check(rs_init(device)); }
inline void Sync() { check(rs_sync()); }

namespace bigpoly {

// bignum.Uint[E] parameters of one modulus (math/bignum/bignum.go:8-31)
class Field {
  public:
    Field(const uint64_t *modulus_le, int limbs) : L_(limbs) { check(rs_field_create(modulus_le, limbs, &h_)); }
    explicit Field(const std::vector<uint64_t> &modulus_le) : Field(modulus_le.data(), (int)modulus_le.size()) {}
    ~Field() { rs_field_destroy(h_); }
    Field(const Field &) = delete;
    Field &operator=(const Field &) = delete;
    int Limb() const { return L_; }                                       // element.go:1776-1778
    std::vector<uint64_t> ToMont(const uint64_t *canonical) const {
        std::vector<uint64_t> out(L_);
        check(rs_field_to_mont(h_, out.data(), canonical));
        return out;
    }
    std::vector<uint64_t> FromMont(const uint64_t *mont) const {
        std::vector<uint64_t> out(L_);
        check(rs_field_from_mont(h_, out.data(), mont));
        return out;
    }
    rs_field *handle() const { return h_; }

  private:
    rs_field *h_ = nullptr;
    int L_;
};

// []*bigpoly.Poly resident in HBM: n_polys polynomials of `rank` coefficients, [poly][coeff][limb]
class Batch {
  public:
    Batch(const Field &f, size_t n_polys, size_t rank) : f_(&f) { check(rs_batch_alloc(f.handle(), n_polys, rank, &h_)); }
    ~Batch() { rs_batch_free(h_); }
    Batch(const Batch &) = delete;
    Batch &operator=(const Batch &) = delete;
    size_t Polys() const { return rs_batch_polys(h_); }
    size_t Rank() const { return rs_batch_rank(h_); }
    const Field &field() const { return *f_; }
    void Upload(const uint64_t *host, size_t poly0, size_t n) { check(rs_batch_upload(h_, poly0, n, host)); }
    void Download(uint64_t *host, size_t poly0, size_t n) const { check(rs_batch_download(h_, poly0, n, host)); }
    rs_batch *handle() const { return h_; }

  private:
    const Field *f_;
    rs_batch *h_ = nullptr;
};

// bigpoly.Poly (poly.go:11-28): a view of one polynomial of a batch + the IsNTT flag
struct Poly {
    Batch *batch;
    size_t index;
    bool IsNTT;
    Poly(Batch &b, size_t i, bool is_ntt = false) : batch(&b), index(i), IsNTT(is_ntt) {}
    size_t Rank() const { return batch->Rank(); }
    // CopyFrom (poly.go:48-62): copies min(rank) coefficients, clears the rest, takes the flag
    void CopyFrom(const Poly &p) {
        const size_t n = p.Rank() < Rank() ? p.Rank() : Rank();
        check(rs_batch_copy(batch->handle(), index, 0, p.batch->handle(), p.index, 0, n));
        if (n < Rank()) check(rs_batch_zero(batch->handle(), index, n, Rank() - n));
        IsNTT = p.IsNTT;
    }
    void Clear() { check(rs_batch_zero(batch->handle(), index, 0, Rank())); }          // poly.go:79-84
    // Evaluate (poly.go:64-76): Horner over the coefficients, x and result Montgomery values on the host
    std::vector<uint64_t> Evaluate(const uint64_t *x) const {
        std::vector<uint64_t> out(batch->field().Limb());
        check(rs_poly_evaluate(batch->handle(), index, Rank(), x, out.data()));
        return out;
    }
};

// bigpoly.transformer (ntt.go:11-14)
class Transformer {
  public:
    virtual ~Transformer() { rs_ntt_destroy(h_); }
    Transformer(const Transformer &) = delete;
    Transformer &operator=(const Transformer &) = delete;
    size_t Rank() const { return rank_; }
    // FwdNTTTo / InvNTTTo on vOut <- v (ntt.go:98-136, :206-244); batches of polynomials at once
    void FwdNTTTo(Batch &out, size_t out_poly, const Batch &in, size_t in_poly, size_t n_polys) const {
        check(rs_ntt_forward(h_, out.handle(), out_poly, in.handle(), in_poly, n_polys));
    }
    void InvNTTTo(Batch &out, size_t out_poly, const Batch &in, size_t in_poly, size_t n_polys) const {
        check(rs_ntt_inverse(h_, out.handle(), out_poly, in.handle(), in_poly, n_polys));
    }
    rs_ntt *handle() const { return h_; }

  protected:
    Transformer(const Field &f, size_t rank, bool negacyclic) : rank_(rank) {
        if (rank == 0 || (rank & (rank - 1)) != 0) throw Panic(RS_ERR_ARG, "rank must be a power of two");   // ntt.go:27-29
        int logn = 0;
        while (((size_t)1 << logn) < rank) logn++;
        check(rs_ntt_create(f.handle(), logn, negacyclic ? 1 : 0, &h_));   // "NTT not supported" when (q-1) % 2N != 0
    }
    rs_ntt *h_ = nullptr;
    size_t rank_;
};
struct CyclicTransformer : Transformer {           // NewCyclicTransformer ntt.go:26-95
    CyclicTransformer(const Field &f, size_t rank) : Transformer(f, rank, false) {}
};
struct CyclotomicTransformer : Transformer {       // NewCyclotomicTransformer ntt.go:153-203
    CyclotomicTransformer(const Field &f, size_t rank) : Transformer(f, rank, true) {}
};

// baseOperator (base_op.go:42-207).  The *To methods write pOut; p0/p1 may alias pOut.
class BaseOperator {
  public:
    size_t Rank() const { return rank_; }
    void AddTo(Poly &pOut, const Poly &p0, const Poly &p1) const {               // base_op.go:49-60
        sameDomain(p0, p1);
        check(rs_vec_add(h(pOut), pOut.index, h(p0), p0.index, h(p1), p1.index, 1));
        pOut.IsNTT = p0.IsNTT;
    }
    void SubTo(Poly &pOut, const Poly &p0, const Poly &p1) const {               // base_op.go:62-73
        sameDomain(p0, p1);
        check(rs_vec_sub(h(pOut), pOut.index, h(p0), p0.index, h(p1), p1.index, 1));
        pOut.IsNTT = p0.IsNTT;
    }
    void NegTo(Poly &pOut, const Poly &p) const {                                // base_op.go:75-80
        check(rs_vec_neg(h(pOut), pOut.index, h(p), p.index, 1));
        pOut.IsNTT = p.IsNTT;
    }
    void ScalarMulTo(Poly &pOut, const Poly &p, const uint64_t *c) const {       // base_op.go:82-87
        check(rs_vec_scalar_mul(h(pOut), pOut.index, h(p), p.index, c, 1));
        pOut.IsNTT = p.IsNTT;
    }
    void ScalarMulAddTo(Poly &pOut, const Poly &p, const uint64_t *c) const {    // base_op.go:89-96
        sameDomain(pOut, p);
        check(rs_vec_scalar_mul_add(h(pOut), pOut.index, h(p), p.index, c, 1));
    }
    void ScalarMulSubTo(Poly &pOut, const Poly &p, const uint64_t *c) const {    // base_op.go:98-105
        sameDomain(pOut, p);
        check(rs_vec_scalar_mul_sub(h(pOut), pOut.index, h(p), p.index, c, 1));
    }
    // MulTo / MulAddTo / MulSubTo (base_op.go:108-171): operands must be in the NTT domain
    void MulTo(Poly &pOut, const Poly &p0, const Poly &p1) const {
        needNTT(p0, p1);
        check(rs_vec_mul(h(pOut), pOut.index, h(p0), p0.index, h(p1), p1.index, 1));
        pOut.IsNTT = true;
    }
    void MulAddTo(Poly &pOut, const Poly &p0, const Poly &p1) const {
        needNTT(p0, p1);
        check(rs_vec_mul_add(h(pOut), pOut.index, h(p0), p0.index, h(p1), p1.index, 1));
    }
    void MulSubTo(Poly &pOut, const Poly &p0, const Poly &p1) const {
        needNTT(p0, p1);
        check(rs_vec_mul_sub(h(pOut), pOut.index, h(p0), p0.index, h(p1), p1.index, 1));
    }
    // NTTTo / InvNTTTo (base_op.go:173-207): panic when the operand is already in the target domain
    void NTTTo(Poly &pOut, const Poly &p) const {
        if (p.IsNTT) throw Panic(RS_ERR_ARG, "polynomial already in NTT domain");
        ntt_->FwdNTTTo(*pOut.batch, pOut.index, *p.batch, p.index, 1);
        pOut.IsNTT = true;
    }
    void InvNTTTo(Poly &pOut, const Poly &p) const {
        if (!p.IsNTT) throw Panic(RS_ERR_ARG, "polynomial not in NTT domain");
        ntt_->InvNTTTo(*pOut.batch, pOut.index, *p.batch, p.index, 1);
        pOut.IsNTT = false;
    }
    const Transformer &transformer() const { return *ntt_; }

  protected:
    BaseOperator(const Transformer *t, size_t rank) : ntt_(t), rank_(rank) {}
    static rs_batch *h(const Poly &p) { return p.batch->handle(); }
    static void sameDomain(const Poly &a, const Poly &b) {
        if (a.IsNTT != b.IsNTT) throw Panic(RS_ERR_ARG, "polynomials are in different domains");
    }
    static void needNTT(const Poly &a, const Poly &b) {
        if (!a.IsNTT || !b.IsNTT) throw Panic(RS_ERR_ARG, "polynomials must be in NTT domain");
    }
    const Transformer *ntt_;
    size_t rank_;
};

class CyclicEvaluator : public BaseOperator {          // cyclic.go:9-16
  public:
    CyclicEvaluator(const Field &f, size_t rank) : BaseOperator(&t_, rank), t_(f, rank) {}
    // QuoRemByVanishing (cyclic.go:18-37): p = quo * (X^N - 1) + rem, coefficient domain
    void QuoRemByVanishing(const Poly &p, size_t N, Poly &quo, Poly &rem) const {
        if (p.IsNTT) throw Panic(RS_ERR_ARG, "polynomial must be in coefficient domain");
        check(rs_quorem_vanishing(h(quo), quo.index, h(rem), rem.index, h(p), p.index, N));
        quo.IsNTT = rem.IsNTT = false;
    }

  private:
    CyclicTransformer t_;
};

class CyclotomicEvaluator : public BaseOperator {      // cyclotomic.go:14-27
  public:
    CyclotomicEvaluator(const Field &f, size_t rank) : BaseOperator(&t_, rank), t_(f, rank) {}
    // AutTo (cyclotomic.go:29-82): X -> X^idx in either domain; pOut must not alias p
    void AutTo(Poly &pOut, const Poly &p, long idx) const {
        check(rs_aut(h(pOut), pOut.index, h(p), p.index, 1, idx, p.IsNTT ? 1 : 0));
        pOut.IsNTT = p.IsNTT;
    }

  private:
    CyclotomicTransformer t_;
};

}  // namespace bigpoly

namespace jindo {

// jindo.Prover (prover.go:14-40) over a commit key already expanded by the host (entities.go:30-77)
class Prover {
  public:
    Prover(const bigpoly::Field &f, const rs_jindo_params &p, const uint64_t *ck_in, const uint64_t *ck_mlwe,
           const uint64_t *ck_out, const uint64_t *cdt_tables)
        : p_(p) {
        check(rs_jindo_create(f.handle(), &p, ck_in, ck_mlwe, ck_out, cdt_tables, &h_));
    }
    // NewProver with NewCommitKey(params, crs) evaluated on the device (entities.go:21-61); key || iv = SHA-384(crs)
    Prover(const bigpoly::Field &f, const rs_jindo_params &p, const uint8_t key[32], const uint8_t iv[16], const uint64_t *cdt_tables)
        : p_(p) {
        check(rs_jindo_create_crs(f.handle(), &p, key, iv, cdt_tables, &h_));
    }
    ~Prover() { rs_jindo_destroy(h_); }
    Prover(const Prover &) = delete;
    Prover &operator=(const Prover &) = delete;

    struct Opening {
        rs_opening *h = nullptr;
        Opening() = default;
        Opening(Opening &&o) noexcept : h(o.h) { o.h = nullptr; }
        Opening &operator=(Opening &&o) noexcept { std::swap(h, o.h); return *this; }
        ~Opening() { rs_opening_free(h); }
    };

    // the encoder's TwinCDT uniform source, key || iv = SHA-384(seed) (math/csprng/uniform.go:38-46)
    void SetTwinStream(const uint8_t key[32], const uint8_t iv[16]) { check(rs_jindo_set_twin_stream(h_, key, iv)); }

    // Commit (prover.go:45-62): v = the first `len` coefficients of poly `v_poly` of a device batch.
    // Returns the Commitment.Value polynomials (out_msis x n_qout x d words); fills `open`.
    std::vector<uint64_t> Commit(const bigpoly::Batch &v, size_t v_poly, size_t len, const rs_jindo_tape &tape, Opening &open) {
        std::vector<uint64_t> com((size_t)p_.out_msis * p_.n_qout * p_.d);
        check(rs_jindo_commit(h_, v.handle(), v_poly, len, &tape, &open.h, com.data()));
        return com;
    }
    // Evaluate (prover.go:205-324) split around the host transcript
    std::vector<uint64_t> Partial(const Opening &open, const uint64_t *left) {             // prover.go:281-289
        std::vector<uint64_t> out((size_t)(p_.cols + 1) * p_.n_q * p_.d);
        check(rs_jindo_partial(h_, open.h, left, out.data()));
        return out;
    }
    void Respond(const Opening &open, const uint64_t *chals, std::vector<uint64_t> &enc, std::vector<uint64_t> &mlwe) {   // :304-316
        enc.assign((size_t)p_.rows * p_.n_q * p_.d, 0);
        mlwe.assign((size_t)(p_.mlwe + p_.in_msis) * p_.n_q * p_.d, 0);
        check(rs_jindo_respond(h_, open.h, chals, enc.data(), mlwe.data()));
    }
    // the same two phases from the transcript's raw inputs: leftVec(x) (utils.go:63-73) and encodeChallengeTo
    // (utils.go:21-46) are evaluated on the device
    std::vector<uint64_t> PartialAt(const Opening &open, const uint64_t *x_mont) {         // prover.go:267-289
        std::vector<uint64_t> out((size_t)(p_.cols + 1) * p_.n_q * p_.d);
        check(rs_jindo_partial_at(h_, open.h, x_mont, out.data()));
        return out;
    }
    void RespondBytes(const Opening &open, const uint8_t *chal_bytes, std::vector<uint64_t> &enc, std::vector<uint64_t> &mlwe) {   // :296-316
        enc.assign((size_t)p_.rows * p_.n_q * p_.d, 0);
        mlwe.assign((size_t)(p_.mlwe + p_.in_msis) * p_.n_q * p_.d, 0);
        check(rs_jindo_respond_bytes(h_, open.h, chal_bytes, enc.data(), mlwe.data()));
    }
    void OpeningMulAddBytes(Opening &acc, const Opening &open, const uint8_t chal16[16]) {   // :230-265
        check(rs_opening_mul_add_bytes(h_, acc.h, open.h, chal16));
    }
    // the whole batch combination of prover.go:230-265 in one pass
    Opening CombineBytes(const std::vector<const Opening *> &opens, const uint8_t *chal_bytes) {
        std::vector<const rs_opening *> hs;
        for (const Opening *o : opens) hs.push_back(o->h);
        Opening out;
        check(rs_opening_combine_bytes(h_, hs.data(), hs.size(), chal_bytes, &out.h));
        return out;
    }
    // Verifier.Verify's arithmetic on the device (verifier.go:50-282): the four check bits, 15 = accepted
    int Verify(const uint64_t *x_mont, const uint64_t *commitments, int batch, const uint8_t *batch_bytes, const uint8_t *chal_bytes,
               const uint64_t *ys_mont, const rs_jindo_proof &pf, double in_com_dcmp_two_nm, double res_two_nm) {
        int checks = 0;
        check(rs_jindo_verify(h_, x_mont, commitments, batch, batch_bytes, chal_bytes, ys_mont, &pf, in_com_dcmp_two_nm, res_two_nm, &checks));
        return checks;
    }
    // Prover.mlweSampler on the device (prover.go:38, :136-138)
    void SetMlweStream(const uint8_t key[32], const uint8_t iv[16], const uint64_t *table0, int len, int64_t tail_lo) {
        check(rs_jindo_set_mlwe_stream(h_, key, iv, table0, len, tail_lo));
    }
    std::vector<uint64_t> LeftVec(const uint64_t *x_mont) {
        std::vector<uint64_t> out((size_t)p_.rows * p_.n_q * p_.d);
        check(rs_jindo_left_vec(h_, x_mont, out.data()));
        return out;
    }
    // ck.In (0) / ck.MLWE (1) / ck.Out (2) as the device holds them
    std::vector<uint64_t> CommitKey(int which) const {
        const size_t n = which == 0 ? (size_t)p_.in_msis * p_.rows * p_.n_q : which == 1 ? (size_t)p_.in_msis * p_.mlwe * p_.n_q
                                    : (size_t)p_.out_msis * (p_.cols + 1) * p_.in_msis * p_.n_qout;
        std::vector<uint64_t> out(n * p_.d);
        check(rs_jindo_commit_key_download(h_, which, out.data()));
        return out;
    }
    const rs_jindo_params &Parameters() const { return p_; }
    rs_jindo *handle() const { return h_; }

  private:
    rs_jindo_params p_;
    rs_jindo *h_ = nullptr;
};

}  // namespace jindo

namespace buckler {

// buckler.Proof (buckler/entities.go) in host vectors
struct Proof {
    std::vector<uint64_t> Witness;                    // batch x out_msis polys over ringQOut
    std::vector<uint64_t> LinCheckMaskSum, SumCheckMaskSum, EvalPoint;   // L64 Montgomery limbs each
    std::vector<uint64_t> Evals;                      // batch x L64 Montgomery limbs
    std::vector<uint64_t> Partial, PartialMask, Encode, MLWE, InCommit;  // jindo.Proof
    double phase_ms[16] = {0};
};

// buckler.Prover (prover.go:16-40): Prove is one call into the compiled prover (ringo_buckler.h)
class Prover {
  public:
    Prover(const bigpoly::Field &f, jindo::Prover &poly_prover, const rs_buckler_circuit &c, const std::string &crs)
        : L_(f.Limb()), jp_(poly_prover.Parameters()) {
        check(rs_buckler_create(f.handle(), poly_prover.handle(), &c, (const uint8_t *)crs.data(), crs.size(), &h_));
    }
    ~Prover() { rs_buckler_destroy(h_); }
    Prover(const Prover &) = delete;
    Prover &operator=(const Prover &) = delete;
    size_t Batch() const { return rs_buckler_batch(h_); }
    // draw all randomness inside the library, as the reference's Prove does (crypto/rand + the device streams)
    void EnableSampling(const rs_jindo_noise &noise, const uint64_t *q_le, int n_threads) {
        check(rs_buckler_enable_sampling(h_, &noise, q_le, n_threads));
    }
    // Prove (prover.go:42-353).  pw / w: pw_cnt x rank and w_cnt x rank Montgomery elements; rand == nullptr: self-sampled
    Proof Prove(const uint64_t *pw, const uint64_t *w, const rs_buckler_rand *rand = nullptr, bool phases = false) {
        Proof pf;
        const size_t B = Batch(), pq = (size_t)jp_.n_q * jp_.d, pqo = (size_t)jp_.n_qout * jp_.d;
        pf.Witness.resize(B * jp_.out_msis * pqo); pf.Evals.resize(B * L_);
        pf.LinCheckMaskSum.resize(L_); pf.SumCheckMaskSum.resize(L_); pf.EvalPoint.resize(L_);
        pf.Partial.resize((size_t)jp_.cols * pq); pf.PartialMask.resize(pq); pf.Encode.resize((size_t)jp_.rows * pq);
        pf.MLWE.resize((size_t)(jp_.mlwe + jp_.in_msis) * pq); pf.InCommit.resize((size_t)(jp_.cols + 1) * jp_.in_msis * pqo);
        rs_buckler_proof out{pf.Witness.data(), pf.LinCheckMaskSum.data(), pf.SumCheckMaskSum.data(), pf.Evals.data(), pf.Partial.data(),
                             pf.PartialMask.data(), pf.Encode.data(), pf.MLWE.data(), pf.InCommit.data(), pf.EvalPoint.data()};
        check(rs_buckler_prove(h_, pw, w, rand, &out, phases ? pf.phase_ms : nullptr));
        return pf;
    }

  private:
    int L_;
    rs_jindo_params jp_;
    rs_buckler *h_ = nullptr;
};

}  // namespace buckler
}  // namespace ringo

#endif  // RINGO_HPP
