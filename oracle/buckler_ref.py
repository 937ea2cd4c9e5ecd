"""oracle/buckler_ref.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Restatement of the Buckler PIOP layer of the reference (sp301415/ringo-snark):
  buckler/context.go, constraint.go, compile.go (ID bookkeeping, ranks), encoder.go, linear.go
  (nttChecker), utils.go, prover.go, verifier.go
and of gnark-crypto v0.19.2 `fiat-shamir.Transcript` over SHA-256 (third-party, not under
/root/reference: the framing H(name || previous challenge || bindings) is restated from the published
algorithm -- PARITY UNPINNED for challenge bytes).

Polynomials are numpy uint64 arrays in Montgomery form handled by the C oracle (oracle.Field /
oracle.NTT); scalars are Python ints (canonical values).  Randomness comes from the same explicit
`Tape` as oracle/jindo_ref.py, consumed in the reference's call order.
The reference's own test only asserts Verify(Prove(.)) == true (buckler/buckler_test.go:82-99);
tests/test_oracle_buckler.py runs that self-check on MultCircuit and PublicKeyCircuit.
"""
import hashlib

import numpy as np

import oracle
from oracle import jindo_ref as J


# ---------------------------------------------------------------------------------------------
# gnark-crypto fiat-shamir (restated)
# ---------------------------------------------------------------------------------------------
class Transcript:
    def __init__(self, names):
        self.names = list(names)
        self.bindings = {n: [] for n in names}
        self.values = {}

    def bind(self, name, data):
        if name in self.values:
            raise ValueError("challenge already computed, cannot be binded to other values")
        self.bindings[name].append(bytes(data))

    def compute_challenge(self, name):
        if name in self.values:
            return self.values[name]
        pos = self.names.index(name)
        h = hashlib.sha256()
        h.update(name.encode())
        if pos != 0:
            prev = self.names[pos - 1]
            if prev not in self.values:
                raise ValueError("the previous challenge is needed and has not been computed")
            h.update(self.values[prev])
        for b in self.bindings[name]:
            h.update(b)
        self.values[name] = h.digest()
        return self.values[name]


CHAL_NAMES = ["projConst", "arithBatchConst", "linCheckBatchConst", "linCheckConst", "sumCheckBatchConst", "evalPoint"]


# ---------------------------------------------------------------------------------------------
# buckler/utils.go
# ---------------------------------------------------------------------------------------------
def decompose_base(x):
    """utils.go:7-32"""
    dcmp_len = x.bit_length()
    if x == (x & -x):                           # power of two
        dcmp_len -= 1
    base = []
    for _ in range(dcmp_len - 1):
        b = x - sum(base)
        base.append((b >> 1) + (b & 1))
    base.append(1)
    return base


def decompose_big(x, base, q):
    """utils.go:34-56"""
    xs = x - q if x > (q >> 1) else x
    out = []
    for b in base:
        if xs >= b:
            out.append(1); xs -= b
        elif xs <= -b:
            out.append(-1); xs += b
        else:
            out.append(0)
    return out


# ---------------------------------------------------------------------------------------------
# buckler/constraint.go, context.go
# ---------------------------------------------------------------------------------------------
class ArithmeticConstraint:
    def __init__(self):
        self.coeffs, self.has_pw, self.pw, self.witness = [], [], [], []

    def add_term_with_const(self, coeff, pw, *ws):
        self.coeffs.append(coeff)
        self.has_pw.append(pw is not None)
        self.pw.append(pw if pw is not None else 0)
        self.witness.append(list(ws))

    def max_rank(self, rank):
        md = 0
        for i in range(len(self.witness)):
            deg = (rank - 1 if self.has_pw[i] else 0) + len(self.witness[i]) * rank
            md = max(md, deg)
        return md + 1


class NTTChecker:
    """linear.go:20-43 (negacyclic NTT; transpose = reverse * N then inverse NTT)."""

    def __init__(self, field, rank):
        self.f = field
        self.rank = rank
        self.ntt = oracle.NTT(field, rank.bit_length() - 1, True)
        self.scale = field.to_mont([rank])[0]

    def transform(self, v):
        return self.ntt.fwd(v)

    def transpose(self, v):
        return self.ntt.inv(self.f.vec_scalar_mul(np.ascontiguousarray(v[::-1]), self.scale))


def mod_inverse(x, m):
    """linear.go:75-92 (extended Euclid on uint64; the result is reduced mod m)"""
    x %= m
    a, b, u, v = x, m, 1, 0
    while b != 0:
        qq = a // b
        a, b = b, a - qq * b
        u, v = v, (u - qq * v) % (1 << 64)          # uint64 wrap-around as in the reference
    if a != 1:
        raise ValueError("modular inverse does not exist")
    return u % m


class AutChecker:
    """linear.go:46-73: X -> X^idx on coefficient (is_ntt False) or NTT-slot (True) vectors; the transpose is the
    automorphism with idx^-1 mod 2N (cyclotomic.go:29-82)."""

    def __init__(self, field, rank, idx, is_ntt):
        self.f, self.rank, self.idx, self.is_ntt = field, rank, idx, is_ntt
        self.idx_inv = mod_inverse(idx, 2 * rank)

    def transform(self, v):
        return self.f.aut(np.ascontiguousarray(v), self.idx, self.is_ntt)

    def transpose(self, v):
        return self.f.aut(np.ascontiguousarray(v), self.idx_inv, self.is_ntt)


class Context:
    def __init__(self, q, rank, pw_cnt, w_cnt):
        self.q, self.rank, self.pw_cnt, self.w_cnt = q, rank, pw_cnt, w_cnt
        self.w_second = []
        self.arith_max_rank = 0
        self.sumcheck_max_rank = 0
        self.arith_constraints = []
        self.sumcheck_constraints, self.sumcheck_sums = [], []
        self.lin_checkers, self.lin_constraints = [], {}
        self.inf_dcmp_bound, self.inf_dcmp_witness = {}, {}
        self.two_dcmp_bound, self.two_dcmp_base, self.two_dcmp_mask, self.two_dcmp_witness = {}, {}, {}, {}

    def add_sumcheck_constraint(self, c, total):
        """context.go:84-94"""
        self.sumcheck_constraints.append(c)
        self.sumcheck_sums.append(total)
        self.sumcheck_max_rank = max(self.sumcheck_max_rank, c.max_rank(self.rank))

    def add_sq_two_norm_constraint(self, w, bound):
        """context.go:165-196: one decomposition witness, two public witnesses (base, mask), a binary constraint
        wDcmp^2 - mask * wDcmp = 0 and the sum-check constraint sum(w^2 - base * wDcmp) = 0."""
        q = self.q
        wd = self.w_cnt; self.w_cnt += 1
        pw_base, pw_mask = self.pw_cnt, self.pw_cnt + 1; self.pw_cnt += 2
        self.two_dcmp_bound[w], self.two_dcmp_base[w], self.two_dcmp_mask[w], self.two_dcmp_witness[w] = bound, pw_base, pw_mask, wd
        c = ArithmeticConstraint()
        c.add_term_with_const(1, None, wd, wd); c.add_term_with_const(q - 1, pw_mask, wd)
        self.add_arithmetic_constraint(c)
        c = ArithmeticConstraint()
        c.add_term_with_const(1, None, w, w); c.add_term_with_const(q - 1, pw_base, wd)
        self.add_sumcheck_constraint(c, 0)

    def add_arithmetic_constraint(self, c):
        self.arith_constraints.append(c)
        self.arith_max_rank = max(self.arith_max_rank, c.max_rank(self.rank))

    def add_linear_constraint(self, w_out, w_in, checker):
        if self.arith_max_rank < 2 * self.rank - 1:
            self.arith_max_rank = 2 * self.rank - 1
        if id(checker) not in self.lin_constraints:
            self.lin_checkers.append(checker)
            self.lin_constraints[id(checker)] = []
        self.lin_constraints[id(checker)].append((w_out, w_in))

    def add_inf_norm_constraint(self, w, bound):
        """context.go:116-163"""
        q = self.q
        if bound < 0:
            return
        if bound == 0:
            c = ArithmeticConstraint(); c.add_term_with_const(1, None, w)
            self.add_arithmetic_constraint(c); return
        if bound == 1:
            c = ArithmeticConstraint()
            c.add_term_with_const(1, None, w, w, w); c.add_term_with_const(q - 1, None, w)
            self.add_arithmetic_constraint(c); return
        base = decompose_base(bound)
        wd = list(range(self.w_cnt, self.w_cnt + len(base)))
        self.inf_dcmp_witness[w] = wd
        self.inf_dcmp_bound[w] = bound
        self.w_cnt += len(base)
        for x in wd:
            c = ArithmeticConstraint()
            c.add_term_with_const(1, None, x, x, x); c.add_term_with_const(q - 1, None, x)
            self.add_arithmetic_constraint(c)
        c = ArithmeticConstraint()
        c.add_term_with_const(1, None, w)
        for b, x in zip(base, wd):
            c.add_term_with_const((-b) % q, None, x)
        self.add_arithmetic_constraint(c)

    def has_arith(self): return len(self.arith_constraints) > 0
    def has_lin(self): return len(self.lin_constraints) > 0
    def has_sum(self): return len(self.sumcheck_constraints) > 0

    def batch(self):
        return self.w_cnt + (1 if self.has_arith() else 0) + (4 if self.has_lin() else 0) + (4 if self.has_sum() else 0)

    def commit_rank(self):
        rank = self.rank + 1 if self.w_cnt > 0 else 0
        for c in self.arith_constraints:
            rank = max(rank, c.max_rank(self.rank) - self.rank)
        if self.has_lin():
            rank = max(rank, 2 * self.rank)
        for c in self.sumcheck_constraints:
            rank = max(rank, c.max_rank(self.rank) + self.rank + 1)
        return rank

    def terms(self, constraints):
        """flatten to (constraint, coeff index, pw or -1, [w...]) + the coefficient list"""
        coeffs, terms = [], []
        for ci, c in enumerate(constraints):
            for i in range(len(c.witness)):
                coeffs.append(c.coeffs[i])
                terms.append((ci, len(coeffs) - 1, c.pw[i] if c.has_pw[i] else -1, list(c.witness[i])))
        return terms, coeffs


# ---------------------------------------------------------------------------------------------
# compile (compile.go:164-207) + encoder.go
# ---------------------------------------------------------------------------------------------
class Compiled:
    pass


def compile_circuit(q, rank, define, n_pw, n_w, crs, tape):
    """define(ctx): records constraints using witness IDs 0..n_w-1 / public IDs 0..n_pw-1 (the IDs the
    reference's reflection walker assigns in struct-field order, compile.go:49-86)."""
    f = oracle.Field(q)
    ctx = Context(q, rank, n_pw, n_w)
    define(ctx, f)
    C = Compiled()
    C.q, C.f, C.ctx, C.crs = q, f, ctx, bytes(crs)
    C.jp = J.new_parameters(q, ctx.commit_rank(), ctx.batch())
    C.emb_rank = 1 << (max(ctx.arith_max_rank, ctx.sumcheck_max_rank) - 1).bit_length()
    C.logN = rank.bit_length() - 1
    C.enc_ntt = oracle.NTT(f, C.logN, False)
    C.emb_ntt = oracle.NTT(f, C.emb_rank.bit_length() - 1, False)
    C.poly_prover = J.Prover(C.jp, crs, tape)                              # compile.go:188
    C.poly_verifier = J.Verifier(C.jp, crs, tape, commit_key=C.poly_prover.ck)   # :200
    return C


def encode(C, v):
    """Encoder.EncodeTo (encoder.go:32-38)"""
    out = np.zeros((C.emb_rank, C.f.L), dtype=np.uint64)
    out[:C.ctx.rank] = C.enc_ntt.inv(np.ascontiguousarray(v[:C.ctx.rank]))
    return out


def _marshal(x, q):
    return x.to_bytes(8 * ((q.bit_length() + 63) // 64), "big")


def _set_bytes(b, q):
    return int.from_bytes(b, "big") % q


# ---------------------------------------------------------------------------------------------
# prover.go
# ---------------------------------------------------------------------------------------------
class Prover:
    def __init__(self, C, tape):
        self.C, self.tape = C, tape
        self.trace = {"blind": [], "mask": [], "commit_traces": []}

    def _mont(self, x):
        return self.C.f.to_mont([x])[0]

    def _commit(self, coeffs_mont):
        vals = self.C.f.from_mont(coeffs_mont)
        com, op = self.C.poly_prover.commit(vals)
        self.trace["commit_traces"].append(self.C.poly_prover.trace)
        return vals, com, op

    def _rand_encode(self, v):
        """encoder.go:50-54"""
        C = self.C; N = C.ctx.rank
        p = encode(C, v)
        r = J.field_set_random(self.tape, C.q)
        self.trace["blind"].append(r)
        p[N] = self._mont(r)
        p[0] = C.f.vec_sub(p[0:1], p[N:N + 1])[0]
        return p

    def _sum_check_mask(self, mask_rank):
        """prover.go:381-397"""
        C = self.C; N = C.ctx.rank; q = C.q
        vals = [0] * C.emb_rank
        for i in range(N):
            vals[i] = J.field_set_random(self.tape, q)
        mask_sum = vals[0]
        for i in range(N, mask_rank):
            vals[i] = J.field_set_random(self.tape, q)
            vals[i - N] = (vals[i - N] - vals[i]) % q
        self.trace["mask"].append(list(vals[:mask_rank]))
        self.trace.setdefault("mask_sum", []).append(mask_sum)
        return C.f.to_mont(vals), mask_sum

    def prove(self, pw_vals, w_vals):
        """Prover.Prove (prover.go:42-353).  pw_vals / w_vals: lists of length-N lists of canonical ints
        for the declared witnesses (decomposition witnesses are appended here, prover.go:77-86)."""
        C = self.C; ctx = C.ctx; f = C.f; q = C.q; N = ctx.rank
        if any(len(v) != N for v in list(pw_vals) + list(w_vals)):
            raise ValueError("witness rank mismatch")
        w = [list(v) for v in w_vals] + [[0] * N for _ in range(ctx.w_cnt - len(w_vals))]
        pw = [list(v) for v in pw_vals] + [[0] * N for _ in range(ctx.pw_cnt - len(pw_vals))]
        for wid, wds in ctx.inf_dcmp_witness.items():
            base = decompose_base(ctx.inf_dcmp_bound[wid])
            for i in range(N):
                d = decompose_big(w[wid][i], base, q)
                for j, x in enumerate(wds):
                    w[x][i] = d[j] % q
        sq_nm = 0                                     # (not reset between constraints, like the reference: prover.go:88)
        for wid, bound in ctx.two_dcmp_bound.items():
            base = decompose_base(bound)                # prover.go:88-110
            pb, pm = ctx.two_dcmp_base[wid], ctx.two_dcmp_mask[wid]
            for i, b in enumerate(base):
                pw[pb][i] = b % q
                pw[pm][i] = 1
            for i in range(N):
                sq_nm += w[wid][i] * w[wid][i]
            sq_nm %= q
            d = decompose_big(sq_nm, base, q)
            for i, x in enumerate(d):
                w[ctx.two_dcmp_witness[wid]][i] = x % q
        self.trace["w_full"], self.trace["pw_full"] = w, pw
        tr = Transcript(CHAL_NAMES)
        pw_ecd_ntt = [C.emb_ntt.fwd(encode(C, f.to_mont(v))) for v in pw]
        w_ecd, w_ntt = [None] * ctx.w_cnt, [None] * ctx.w_cnt
        B = ctx.batch()
        coms, opens, com_polys = [None] * B, [None] * B, [None] * B
        for i in range(ctx.w_cnt):
            if i in ctx.w_second:
                continue
            w_ecd[i] = self._rand_encode(f.to_mont(w[i]))
            w_ntt[i] = C.emb_ntt.fwd(w_ecd[i])
            com_polys[i], coms[i], opens[i] = self._commit(w_ecd[i][:N + 1])
            tr.bind("projConst", J.commitment_write_raw(C.jp, coms[i]))
        tr.compute_challenge("projConst")
        idx = ctx.w_cnt
        lin_mask = lin_mask_sum = None
        if ctx.has_lin():
            lin_mask, lin_mask_sum = self._sum_check_mask(2 * N)
            com_polys[idx], coms[idx], opens[idx] = self._commit(lin_mask[:2 * N])
            tr.bind("arithBatchConst", J.commitment_write_raw(C.jp, coms[idx]))
            tr.bind("arithBatchConst", _marshal(lin_mask_sum, q))
            idx += 1
        sum_mask = sum_mask_sum = None
        if ctx.has_sum():                              # prover.go:223-238
            sum_mask, sum_mask_sum = self._sum_check_mask(ctx.sumcheck_max_rank)
            com_polys[idx], coms[idx], opens[idx] = self._commit(sum_mask[:ctx.sumcheck_max_rank])
            tr.bind("arithBatchConst", J.commitment_write_raw(C.jp, coms[idx]))
            tr.bind("arithBatchConst", _marshal(sum_mask_sum, q))
            idx += 1
        arith_bc = _set_bytes(tr.compute_challenge("arithBatchConst"), q)
        if ctx.has_arith():
            terms, coeffs = ctx.terms(ctx.arith_constraints)
            ev = f.eval_circuit(pw_ecd_ntt, w_ntt, terms, f.to_mont(coeffs), self._mont(arith_bc), C.emb_rank)
            quo, _ = f.quorem_vanishing(C.emb_ntt.inv(ev), N)                  # prover.go:399-404
            com_polys[idx], coms[idx], opens[idx] = self._commit(quo[:ctx.arith_max_rank - N])
            tr.bind("evalPoint", J.commitment_write_raw(C.jp, coms[idx]))
            idx += 1
        lin_bc = _set_bytes(tr.compute_challenge("linCheckBatchConst"), q)
        lin_c = _set_bytes(tr.compute_challenge("linCheckConst"), q)
        if ctx.has_lin():
            quo, rem_lo, rem_hi = self._lin_check(lin_bc, lin_c, lin_mask, w_ntt)
            for poly in (quo, rem_lo, rem_hi):
                com_polys[idx], coms[idx], opens[idx] = self._commit(poly)
                tr.bind("evalPoint", J.commitment_write_raw(C.jp, coms[idx]))
                idx += 1
        sum_bc = _set_bytes(tr.compute_challenge("sumCheckBatchConst"), q)
        if ctx.has_sum():                              # prover.go:307-331, :461-485
            terms, coeffs = ctx.terms(ctx.sumcheck_constraints)
            bcm = self._mont(sum_bc)
            ev = f.eval_circuit(pw_ecd_ntt, w_ntt, terms, f.to_mont(coeffs), bcm, C.emb_rank)
            ev = f.vec_scalar_mul(ev, bcm)
            ev = f.vec_add(C.emb_ntt.inv(ev), sum_mask)
            quo, rem = f.quorem_vanishing(ev, N)
            rem_lo = rem[1:N].copy()
            R = C.jp.rank
            rem_hi = np.zeros((R, f.L), dtype=np.uint64)
            rem_hi[R - (N - 1):] = rem_lo
            for poly in (quo[:ctx.sumcheck_max_rank - N].copy(), rem_lo, rem_hi):
                com_polys[idx], coms[idx], opens[idx] = self._commit(poly)
                tr.bind("evalPoint", J.commitment_write_raw(C.jp, coms[idx]))
                idx += 1
        x = _set_bytes(tr.compute_challenge("evalPoint"), q)
        evals, pf = C.poly_prover.evaluate(x, com_polys, coms, opens)
        return {"Witness": coms, "LinCheckMaskSum": lin_mask_sum, "SumCheckMaskSum": sum_mask_sum, "Evals": evals,
                "EvalProof": pf, "_eval_point": x, "_opens": opens, "_com_polys": com_polys,
                "_chal": {"arith": arith_bc, "lin_bc": lin_bc, "lin_c": lin_c, "sum_bc": sum_bc}}

    def _lin_check(self, bc, lin_c, mask, w_ntt):
        """prover.go:406-459"""
        C = self.C; ctx = C.ctx; f = C.f; q = C.q; N = ctx.rank
        vec = [1]
        for _ in range(1, N):
            vec.append(vec[-1] * lin_c % q)
        vec_m = f.to_mont(vec)
        lin_ecd = C.emb_ntt.fwd(encode(C, vec_m))
        ev = np.zeros((C.emb_rank, f.L), dtype=np.uint64)
        bcm = self._mont(bc)
        for chk in ctx.lin_checkers:
            tr_ecd = C.emb_ntt.fwd(encode(C, chk.transpose(vec_m)))
            for (w_out, w_in) in ctx.lin_constraints[id(chk)]:
                term = f.vec_sub(f.vec_mul(tr_ecd, w_ntt[w_in]), f.vec_mul(lin_ecd, w_ntt[w_out]))
                ev = f.vec_add(f.vec_scalar_mul(ev, bcm), term)
        ev = f.vec_scalar_mul(ev, bcm)
        ev = f.vec_add(C.emb_ntt.inv(ev), mask)
        quo, rem = f.quorem_vanishing(ev, N)
        rem_lo = rem[1:N].copy()
        R = C.jp.rank
        rem_hi = np.zeros((R, f.L), dtype=np.uint64)
        rem_hi[R - (N - 1):] = rem_lo
        return quo[:N].copy(), rem_lo, rem_hi


# ---------------------------------------------------------------------------------------------
# verifier.go
# ---------------------------------------------------------------------------------------------
class Verifier:
    def __init__(self, C):
        self.C = C

    def _encode_eval(self, vals_mont, x_m):
        """Encode(v).Evaluate(x) (encoder.go:23-38, poly.go:64-76)"""
        C = self.C
        return C.f.from_mont(C.f.poly_evaluate(encode(C, vals_mont)[:C.ctx.rank], x_m).reshape(1, -1))[0]

    def _eval_circuit(self, bc, constraints, evals, pw_evals):
        q = self.C.q
        out = 0
        for c in constraints:
            ev = 0
            for i in range(len(c.witness)):
                t = c.coeffs[i]
                if c.has_pw[i]:
                    t = t * pw_evals[c.pw[i]] % q
                for wi in c.witness[i]:
                    t = t * evals[wi] % q
                ev = (ev + t) % q
            out = (out + ev * bc) % q
        return out

    def verify(self, pw_vals, pf):
        """Verifier.Verify (verifier.go:29-189)"""
        C = self.C; ctx = C.ctx; f = C.f; q = C.q; N = ctx.rank
        pw = [list(v) for v in pw_vals] + [[0] * N for _ in range(ctx.pw_cnt - len(pw_vals))]
        for wid, bound in ctx.two_dcmp_bound.items():          # verifier.go:48-57
            for i, b in enumerate(decompose_base(bound)):
                pw[ctx.two_dcmp_base[wid]][i] = b % q
                pw[ctx.two_dcmp_mask[wid]][i] = 1
        tr = Transcript(CHAL_NAMES)
        for i in range(ctx.w_cnt):
            if i not in ctx.w_second:
                tr.bind("projConst", J.commitment_write_raw(C.jp, pf["Witness"][i]))
        tr.compute_challenge("projConst")
        idx = ctx.w_cnt
        lin_mask_eval = None
        if ctx.has_lin():
            lin_mask_eval = pf["Evals"][idx]
            tr.bind("arithBatchConst", J.commitment_write_raw(C.jp, pf["Witness"][idx]))
            tr.bind("arithBatchConst", _marshal(pf["LinCheckMaskSum"], q))
            idx += 1
        sum_mask_eval = None
        if ctx.has_sum():                              # verifier.go:128-136
            sum_mask_eval = pf["Evals"][idx]
            tr.bind("arithBatchConst", J.commitment_write_raw(C.jp, pf["Witness"][idx]))
            tr.bind("arithBatchConst", _marshal(pf["SumCheckMaskSum"], q))
            idx += 1
        arith_bc = _set_bytes(tr.compute_challenge("arithBatchConst"), q)
        lin_bc = _set_bytes(tr.compute_challenge("linCheckBatchConst"), q)
        lin_c = _set_bytes(tr.compute_challenge("linCheckConst"), q)
        sum_bc = _set_bytes(tr.compute_challenge("sumCheckBatchConst"), q)
        for i in range(idx, len(pf["Witness"])):
            tr.bind("evalPoint", J.commitment_write_raw(C.jp, pf["Witness"][i]))
        x = _set_bytes(tr.compute_challenge("evalPoint"), q)
        if not C.poly_verifier.verify(x, pf["Witness"], pf["Evals"], pf["EvalProof"]):
            return False
        x_m = f.to_mont([x])[0]
        vanish = (pow(x, N, q) - 1) % q
        pw_evals = [self._encode_eval(f.to_mont(v), x_m) for v in pw]
        evals = pf["Evals"]
        if ctx.has_arith():
            ev = self._eval_circuit(arith_bc, ctx.arith_constraints, evals, pw_evals)
            if ev != evals[idx] * vanish % q:
                return False
            idx += 1
        if ctx.has_lin():
            quo_e, lo_e, hi_e = evals[idx], evals[idx + 1], evals[idx + 2]
            if hi_e != pow(x, C.jp.rank - (N - 1), q) * lo_e % q:
                return False
            vec = [1]
            for _ in range(1, N):
                vec.append(vec[-1] * lin_c % q)
            vec_m = f.to_mont(vec)
            lin_eval = self._encode_eval(vec_m, x_m)
            ev = 0
            for chk in ctx.lin_checkers:
                tr_eval = self._encode_eval(chk.transpose(vec_m), x_m)
                for (w_out, w_in) in ctx.lin_constraints[id(chk)]:
                    term = (tr_eval * evals[w_in] - lin_eval * evals[w_out]) % q
                    ev = (ev * lin_bc + term) % q
            ev = (ev * lin_bc + lin_mask_eval) % q
            test = (quo_e * vanish + lo_e * x + pf["LinCheckMaskSum"]) % q
            if ev != test:
                return False
            idx += 3
        if ctx.has_sum():                              # verifier.go:295-315
            quo_e, lo_e, hi_e = evals[idx], evals[idx + 1], evals[idx + 2]
            if hi_e != pow(x, C.jp.rank - (N - 1), q) * lo_e % q:
                return False
            ev = self._eval_circuit(sum_bc, ctx.sumcheck_constraints, evals, pw_evals)
            ev = (ev * sum_bc + sum_mask_eval) % q
            test = (quo_e * vanish + lo_e * x + pf["SumCheckMaskSum"]) % q
            if ev != test:
                return False
            idx += 3
        return True


# ---------------------------------------------------------------------------------------------
# the circuits of the reference's examples / tests
# ---------------------------------------------------------------------------------------------
def mult_circuit(rank):
    """examples/mult/main.go:35-65.  public: YNTT(0); witnesses: XCoeffs(0) ZCoeffs(1) XNTT(2) ZNTT(3)."""
    def define(ctx, f):
        chk = NTTChecker(f, rank)
        ctx.add_linear_constraint(2, 0, chk)
        ctx.add_linear_constraint(3, 1, chk)
        c = ArithmeticConstraint()
        c.add_term_with_const(1, 0, 2)
        c.add_term_with_const(ctx.q - 1, None, 3)
        ctx.add_arithmetic_constraint(c)
        ctx.add_inf_norm_constraint(0, 5)
    return define, 1, 4


def public_key_circuit(rank):
    """buckler/buckler_test.go:17-44.  witnesses: Sk(0) SkNTT(1) Noise(2) NoiseNTT(3); public: PkNTT[0], PkNTT[1]."""
    def define(ctx, f):
        chk = NTTChecker(f, rank)
        ctx.add_linear_constraint(1, 0, chk)
        ctx.add_linear_constraint(3, 2, chk)
        c = ArithmeticConstraint()
        c.add_term_with_const(1, 1)
        c.add_term_with_const(ctx.q - 1, 0, 1)
        c.add_term_with_const(ctx.q - 1, None, 3)
        ctx.add_arithmetic_constraint(c)
        ctx.add_inf_norm_constraint(0, 1)
        ctx.add_inf_norm_constraint(2, 1)
    return define, 2, 4


def norm_aut_circuit(rank, aut_idx, sq_bound):
    """A circuit over the remaining constraint kinds of buckler (none of the shipped examples uses them; BASELINE
    configs[3] names them): an automorphism linear constraint in the coefficient domain and one in the NTT domain
    (NewAutChecker, linear.go:54-73), a squared two-norm constraint (AddSqTwoNormConstraint, context.go:165-196: binary
    decomposition + sum-check) and an explicit sum-check constraint (AddSumCheckConstraint, context.go:84-94).
    witnesses: S(0) SAut(1) SNTT(2) SNTTAut(3); no declared public witness."""
    def define(ctx, f):
        ctx.add_linear_constraint(1, 0, AutChecker(f, rank, aut_idx, False))      # SAut = Aut(S)
        ctx.add_linear_constraint(2, 0, NTTChecker(f, rank))                      # SNTT = NTT(S)
        ctx.add_linear_constraint(3, 2, AutChecker(f, rank, aut_idx, True))       # SNTTAut = AutNTT(SNTT)
        ctx.add_inf_norm_constraint(0, 1)                                         # ternary S
        ctx.add_sq_two_norm_constraint(0, sq_bound)                               # |S|_2^2 <= bound
        c = ArithmeticConstraint()                                                # sum_i (S_i^3 - S_i) = 0  (true for ternary S)
        c.add_term_with_const(1, None, 0, 0, 0); c.add_term_with_const(ctx.q - 1, None, 0)
        ctx.add_sumcheck_constraint(c, 0)
    return define, 0, 4


def ciphertext_circuit(rank, delta, plaintext_modulus, err_bound):
    """examples/bfv/main.go:30-68 (CiphertextCircuit).  public: CiphertextNTT[0], CiphertextNTT[1];
    witnesses in declaration order: SecretKeyNTT(0) MessageNTT(1) MessageCoeffs(2) ErrorNTT(3) ErrorCoeffs(4).
    Body + Mask*sk - Delta*Message - Error = 0 (:56-62), |m| <= t - 1 (:65), |e| <= ErrBound (:67)."""
    def define(ctx, f):
        chk = NTTChecker(f, rank)
        ctx.add_linear_constraint(1, 2, chk)
        ctx.add_linear_constraint(3, 4, chk)
        c = ArithmeticConstraint()
        c.add_term_with_const(1, 0)
        c.add_term_with_const(1, 1, 0)
        c.add_term_with_const((-delta) % ctx.q, None, 1)
        c.add_term_with_const(ctx.q - 1, None, 3)
        ctx.add_arithmetic_constraint(c)
        ctx.add_inf_norm_constraint(2, plaintext_modulus - 1)
        ctx.add_inf_norm_constraint(4, err_bound)
    return define, 2, 5
