"""Host-side mirror of buckler.Prover.Prove (buckler/prover.go:42-353) over the C ABI.

The orchestration (round structure, Fiat-Shamir transcript, which polynomial is committed when)
follows the reference line by line; every polynomial operation is a device call:

  Encode / RandEncode      rs_ntt_inverse_embed (+ one blinding coefficient)      buckler/encoder.go:32-54
  polyEval.NTT             rs_ntt_forward at embRank                               prover.go:128,150
  polyProver.Commit        rs_jindo_commit on the device-resident coefficients    prover.go:152
  arithCheck               rs_eval_circuit -> rs_ntt_inverse -> rs_quorem_vanishing   prover.go:355-404
  linCheck                 rs_reverse_scale + negacyclic INTT (TransposeTo), rs_vec_* , QuoRem   prover.go:406-459
  polyProver.Evaluate      rs_opening_mul_add / rs_jindo_partial / rs_jindo_respond / rs_poly_evaluate  jindo/prover.go:205-324

What stays on the host is what the reference keeps on the host in any deployment: the transcript
(SHA-256 / SHAKE128), big-integer witness decomposition (done by the caller: `w` must already contain
the decomposition witnesses, prover.go:77-86), powers of the linCheck constant, and all randomness,
which is an explicit input (`rand`: blinding elements, mask polynomials and one Jindo tape per Commit,
in call order) exactly like the tapes of include/ringo_jindo.h.

`circuit` is the compiled description a Go caller gets from buckler.Compile (plain data):
  rank, emb_rank, w_cnt, pw_cnt, arith_max_rank, arith_terms [(constraint, coeff_idx, pw|-1, [w..])],
  arith_coeffs [ints], lin [(checker, [(w_out, w_in), ...])] with checker == "ntt" (nttChecker, linear.go:20-43) or
  ("aut", idx_inv, is_ntt) (autChecker, linear.go:46-73: TransposeTo is AutTo with idx^-1 mod 2N),
  and, when the circuit has sum-check constraints (context.go:84-94; AddSqTwoNormConstraint adds one, :165-196),
  sum_terms / sum_coeffs in the same flattened form plus sum_max_rank.
"""
import ctypes
import hashlib

import numpy as np

from . import capi, jindo_host as JH
from .bigpoly import Batch, CyclicTransformer, CyclotomicTransformer
from .capi import check, lib, c_u64p

CHAL_NAMES = ["projConst", "arithBatchConst", "linCheckBatchConst", "linCheckConst", "sumCheckBatchConst", "evalPoint"]


class Transcript:
    """gnark-crypto fiat-shamir over SHA-256: H(name || previous challenge || bindings) (restated)."""

    def __init__(self, names):
        self.names, self.bindings, self.values = list(names), {n: [] for n in names}, {}

    def bind(self, name, data):
        self.bindings[name].append(bytes(data))

    def compute_challenge(self, name):
        pos = self.names.index(name)
        h = hashlib.sha256(name.encode())
        if pos:
            h.update(self.values[self.names[pos - 1]])
        for b in self.bindings[name]:
            h.update(b)
        self.values[name] = h.digest()
        return self.values[name]


def _p(a):
    return a.ctypes.data_as(c_u64p)


class Prover:
    def __init__(self, field, q, circuit, jindo_params, jindo_prover):
        self.f, self.q, self.c, self.jp, self.gp = field, q, circuit, jindo_params, jindo_prover
        self.L = field.L
        self.R = 1 << (64 * self.L)
        N, E = circuit["rank"], circuit["emb_rank"]
        self.enc = CyclicTransformer(field, N)           # Encoder.ntt (encoder.go:15-20)
        self.emb = CyclicTransformer(field, E)           # polyEval (compile.go:186)
        self.neg = CyclotomicTransformer(field, N)       # nttChecker (linear.go:26-33)

    # ---- scalars (host) -------------------------------------------------------------------------
    def mont(self, x):
        return np.array([((x % self.q) * self.R % self.q >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(self.L)], dtype=np.uint64)

    def from_mont_arr(self, a):
        rinv = pow(self.R, -1, self.q)
        return sum(int(v) << (64 * i) for i, v in enumerate(a)) * rinv % self.q

    def _set_elem(self, batch, poly, off, value):
        if getattr(self, "_one", None) is None:
            self._one = Batch(self.f, 1, 1)
        self._one.upload(self.mont(value).reshape(1, 1, self.L))
        check(lib().rs_batch_copy(batch.h, poly, off, self._one.h, 0, 0, 1))

    def mont_many(self, values):
        """canonical ints -> [n, L] Montgomery limbs (one big-int product each; limbs split by numpy)"""
        R, q, nb = self.R, self.q, 8 * self.L
        buf = b"".join(((v % q) * R % q).to_bytes(nb, "little") for v in values)
        return np.frombuffer(buf, dtype="<u8").reshape(len(values), self.L).copy()

    # ---- Jindo commit of coefficients [0, length) of batch[poly] ---------------------------------
    def _commit(self, batch, poly, length, tape):
        com, op = self.gp.Commit(batch, poly, length, tape)
        return com, op

    def prove(self, pw_mont, w_mont, rand):
        """pw_mont / w_mont: uint64 arrays [pw_cnt, N, L] / [w_cnt, N, L] (Montgomery).  rand: dict with
        "blind" (ints, one per witness), "mask" (list of int lists) and "tapes" (one Jindo tape dict per
        Commit, in call order).  Returns the proof as plain data."""
        c, f, q = self.c, self.f, self.q
        N, E, wc, pc = c["rank"], c["emb_rank"], c["w_cnt"], c["pw_cnt"]
        L = lib()
        tapes = iter(rand["tapes"])
        tr = Transcript(CHAL_NAMES)
        # wall-time split of the last Prove (ms; the stream is synchronised at every mark)
        import time as _time
        self.timing = {}
        _t = [_time.perf_counter()]

        def mark(name):
            check(L.rs_sync())
            now = _time.perf_counter()
            self.timing[name] = self.timing.get(name, 0.0) + (now - _t[0]) * 1e3
            _t[0] = now
        # ---- Encode + NTT of all (public) witnesses: 1 upload, 2 batched transforms
        vals = Batch(f, wc + pc, N).upload(np.concatenate([w_mont, pw_mont]))
        ecd = Batch(f, wc + pc, E)
        check(L.rs_ntt_inverse_embed(self.enc.h, ecd.h, 0, vals.h, 0, wc + pc, 1))
        bl = self.mont_many(rand["blind"][:wc])            # RandEncodeTo: c[N] <- r ; c[0] -= c[N], all witnesses in one launch
        check(L.rs_rand_encode_blind(ecd.h, 0, wc, N, _p(bl)))
        ntt = Batch(f, wc + pc, E)
        self.emb.FwdNTTTo(ntt, ecd)
        w_ntt = Batch(f, wc, E, device_ptr=L.rs_batch_device_ptr(ntt.h), keepalive=ntt)
        pw_ntt = Batch(f, max(pc, 1), E)
        for k in range(pc):
            check(L.rs_batch_copy(pw_ntt.h, k, 0, ntt.h, wc + k, 0, E))
        mark("encode_ntt")
        has_sum = bool(c.get("sum_terms"))
        n_com = wc + (1 if c["arith_terms"] else 0) + (4 if c["lin"] else 0) + (4 if has_sum else 0)
        coms, opens, views = [None] * n_com, [None] * n_com, [None] * n_com
        for i in range(wc):                              # first-round witnesses (no second-round witnesses here)
            coms[i], opens[i] = self._commit(ecd, i, N + 1, next(tapes))
            views[i] = (ecd, i, N + 1)
            tr.bind("projConst", JH.commitment_write_raw(self.jp, coms[i]))
        tr.compute_challenge("projConst")
        mark("commit")
        idx = wc
        work = Batch(f, 8, E)                            # 0 linCheckMask, 1 eval, 2 quo, 3 rem, 4 term, 5 sumCheckMask, 6-7 scratch
        lin_mask_sum = sum_mask_sum = None
        masks = iter(zip(rand["mask"], rand.get("mask_sum", [])))
        if c["lin"]:
            mask, lin_mask_sum = next(masks)
            host = np.zeros((1, E, self.L), dtype=np.uint64)
            host[0, :2 * N] = self.mont_many(mask)
            work.upload(host, 0)
            coms[idx], opens[idx] = self._commit(work, 0, 2 * N, next(tapes)); views[idx] = (work, 0, 2 * N)
            tr.bind("arithBatchConst", JH.commitment_write_raw(self.jp, coms[idx]))
            tr.bind("arithBatchConst", lin_mask_sum.to_bytes(8 * self.L, "big"))
            idx += 1
        if has_sum:                                      # sumCheckMask (prover.go:223-238, :381-397)
            mask, sum_mask_sum = next(masks)
            SR = c["sum_max_rank"]
            host = np.zeros((1, E, self.L), dtype=np.uint64)
            host[0, :SR] = self.mont_many(mask)
            work.upload(host, 5)
            coms[idx], opens[idx] = self._commit(work, 5, SR, next(tapes)); views[idx] = (work, 5, SR)
            tr.bind("arithBatchConst", JH.commitment_write_raw(self.jp, coms[idx]))
            tr.bind("arithBatchConst", sum_mask_sum.to_bytes(8 * self.L, "big"))
            idx += 1
        mark("commit")
        arith_bc = int.from_bytes(tr.compute_challenge("arithBatchConst"), "big") % q
        keep = []                                        # committed polynomials must stay alive until Evaluate
        if c["arith_terms"]:
            T = (capi.Term * len(c["arith_terms"]))()
            for k, (ci, co, pw, ws) in enumerate(c["arith_terms"]):
                T[k].constraint, T[k].coeff, T[k].pw, T[k].n_w = ci, co, pw, len(ws)
                for j, x in enumerate(ws):
                    T[k].w[j] = x
            coeffs = self.mont_many(c["arith_coeffs"])
            check(L.rs_eval_circuit(work.h, 1, pw_ntt.h, w_ntt.h, T, len(c["arith_terms"]), _p(coeffs), len(coeffs), _p(self.mont(arith_bc))))
            check(L.rs_ntt_inverse(self.emb.h, work.h, 1, work.h, 1, 1))
            quo_b = Batch(f, 2, E)
            check(L.rs_quorem_vanishing(quo_b.h, 0, quo_b.h, 1, work.h, 1, N))
            keep.append(quo_b)
            mark("arith_check")
            coms[idx], opens[idx] = self._commit(quo_b, 0, c["arith_max_rank"] - N, next(tapes)); views[idx] = (quo_b, 0, c["arith_max_rank"] - N)
            tr.bind("evalPoint", JH.commitment_write_raw(self.jp, coms[idx]))
            idx += 1
        mark("commit")
        lin_bc = int.from_bytes(tr.compute_challenge("linCheckBatchConst"), "big") % q
        lin_c = int.from_bytes(tr.compute_challenge("linCheckConst"), "big") % q
        if c["lin"]:
            vec = [1]
            for _ in range(1, N):
                vec.append(vec[-1] * lin_c % q)
            lvm = self.mont_many(vec)
            lv = Batch(f, 2, N).upload(np.stack([lvm, lvm]))
            le = Batch(f, 2, E)
            check(L.rs_ntt_inverse_embed(self.enc.h, le.h, 0, lv.h, 0, 1, 1))
            check(L.rs_ntt_forward(self.emb.h, le.h, 0, le.h, 0, 1))                     # linCheckEcd (NTT form)
            check(L.rs_batch_zero(work.h, 1, 0, E))                                      # eval = 0
            bcm = self.mont(lin_bc)
            for kind, pairs in c["lin"]:
                if kind == "ntt":
                    check(L.rs_reverse_scale(lv.h, 1, lv.h, 0, _p(self.mont(N))))        # nttChecker.TransposeTo (linear.go:38-43)
                    check(L.rs_ntt_inverse(self.neg.h, lv.h, 1, lv.h, 1, 1))
                else:
                    _, idx_inv, is_ntt = kind                                            # autChecker.TransposeTo (linear.go:69-73)
                    check(L.rs_aut(lv.h, 1, lv.h, 0, 1, int(idx_inv), 1 if is_ntt else 0))
                check(L.rs_ntt_inverse_embed(self.enc.h, le.h, 1, lv.h, 1, 1, 1))
                check(L.rs_ntt_forward(self.emb.h, le.h, 1, le.h, 1, 1))                 # linCheckEcdTr
                for (w_out, w_in) in pairs:
                    check(L.rs_vec_mul(work.h, 4, le.h, 1, w_ntt.h, w_in, 1))            # term = EcdTr * wIn
                    check(L.rs_vec_mul_sub(work.h, 4, le.h, 0, w_ntt.h, w_out, 1))       # term -= Ecd * wOut
                    check(L.rs_vec_scalar_mul(work.h, 1, work.h, 1, _p(bcm), 1))         # eval *= batchConst
                    check(L.rs_vec_add(work.h, 1, work.h, 1, work.h, 4, 1))              # eval += term
            check(L.rs_vec_scalar_mul(work.h, 1, work.h, 1, _p(bcm), 1))
            check(L.rs_ntt_inverse(self.emb.h, work.h, 1, work.h, 1, 1))
            check(L.rs_vec_add(work.h, 1, work.h, 1, work.h, 0, 1))                      # + linCheckMask
            qr = Batch(f, 2, E)
            check(L.rs_quorem_vanishing(qr.h, 0, qr.h, 1, work.h, 1, N))
            Rk = self.jp.rank
            rem = Batch(f, 2, max(Rk, N))                                                # 0: remLo, 1: remHi
            check(L.rs_batch_zero(rem.h, 0, 0, rem.rank)); check(L.rs_batch_zero(rem.h, 1, 0, rem.rank))
            check(L.rs_batch_copy(rem.h, 0, 0, qr.h, 1, 1, N - 1))
            check(L.rs_batch_copy(rem.h, 1, Rk - (N - 1), qr.h, 1, 1, N - 1))
            keep += [qr, rem]
            mark("lin_check")
            for (b, p, ln) in ((qr, 0, N), (rem, 0, N - 1), (rem, 1, Rk)):
                coms[idx], opens[idx] = self._commit(b, p, ln, next(tapes)); views[idx] = (b, p, ln)
                tr.bind("evalPoint", JH.commitment_write_raw(self.jp, coms[idx]))
                idx += 1
        sum_bc = int.from_bytes(tr.compute_challenge("sumCheckBatchConst"), "big") % q
        if has_sum:                                      # sumCheck (prover.go:307-331, :461-485)
            T = (capi.Term * len(c["sum_terms"]))()
            for k, (ci, co, pw, ws) in enumerate(c["sum_terms"]):
                T[k].constraint, T[k].coeff, T[k].pw, T[k].n_w = ci, co, pw, len(ws)
                for j, xw in enumerate(ws):
                    T[k].w[j] = xw
            coeffs = self.mont_many(c["sum_coeffs"])
            bcm = self.mont(sum_bc)
            check(L.rs_eval_circuit(work.h, 1, pw_ntt.h, w_ntt.h, T, len(c["sum_terms"]), _p(coeffs), len(coeffs), _p(bcm)))
            check(L.rs_vec_scalar_mul(work.h, 1, work.h, 1, _p(bcm), 1))
            check(L.rs_ntt_inverse(self.emb.h, work.h, 1, work.h, 1, 1))
            check(L.rs_vec_add(work.h, 1, work.h, 1, work.h, 5, 1))                      # + sumCheckMask
            qs = Batch(f, 2, E)
            check(L.rs_quorem_vanishing(qs.h, 0, qs.h, 1, work.h, 1, N))
            Rk = self.jp.rank
            rems = Batch(f, 2, max(Rk, N))
            check(L.rs_batch_zero(rems.h, 0, 0, rems.rank)); check(L.rs_batch_zero(rems.h, 1, 0, rems.rank))
            check(L.rs_batch_copy(rems.h, 0, 0, qs.h, 1, 1, N - 1))
            check(L.rs_batch_copy(rems.h, 1, Rk - (N - 1), qs.h, 1, 1, N - 1))
            keep += [qs, rems]
            mark("sum_check")
            for (b, p, ln) in ((qs, 0, c["sum_max_rank"] - N), (rems, 0, N - 1), (rems, 1, Rk)):
                coms[idx], opens[idx] = self._commit(b, p, ln, next(tapes)); views[idx] = (b, p, ln)
                tr.bind("evalPoint", JH.commitment_write_raw(self.jp, coms[idx]))
                idx += 1
        x = int.from_bytes(tr.compute_challenge("evalPoint"), "big") % q
        mark("commit")
        evals, pf = self.evaluate(x, views, coms, opens)
        mark("evaluate")
        return {"Witness": [cm.tolist() for cm in coms], "LinCheckMaskSum": lin_mask_sum, "SumCheckMaskSum": sum_mask_sum,
                "Evals": evals, "EvalProof": pf, "_eval_point": x}

    def evaluate(self, x, views, coms, opens):
        """jindo.Prover.Evaluate (jindo/prover.go:205-324); views: (batch, poly, length) of each committed vector."""
        jp, gp, q = self.jp, self.gp, self.q
        batch = len(coms)
        prefix = gp.crs + b"".join(JH.commitment_write_raw(jp, cm) for cm in coms) + x.to_bytes(8 * self.L, "big")
        if batch > 1:
            bb = hashlib.shake_128(prefix).digest(16 * batch)
            prefix += bb
            ob = gp.NewOpening()
            for i in range(batch):                       # challenge polynomials are encoded on the device from their bytes
                gp.OpeningMulAddBytes(ob, opens[i], bb[16 * i:16 * i + 16])
        else:
            ob = opens[0]
        part = gp.PartialAt(ob, self.mont(x))            # leftVec(x) + Encoder.EncodeTo on the device
        pf = {"Partial": part[:jp.cols].tolist(), "PartialMask": part[jp.cols].tolist(),
              "InCommit": [ob.InCommit(k).tolist() for k in range(jp.in_com_dcmp_len)]}
        tr = prefix + b"".join(JH.poly_write_to(c) for c in pf["Partial"]) + JH.poly_write_to(pf["PartialMask"])
        cs = hashlib.shake_128(tr).digest(16 * jp.cols)
        enc, ml = gp.RespondBytes(ob, cs)
        pf["Encode"], pf["MLWE"] = enc.tolist(), ml.tolist()
        xm = self.mont(x)
        evals = []
        for (b, p, ln) in views:
            out = np.zeros(self.L, dtype=np.uint64)
            check(lib().rs_poly_evaluate(b.h, p, ln, _p(xm), _p(out)))
            evals.append(self.from_mont_arr(out))
        return evals, pf


class _PinnedArray:
    """a numpy copy of `src` living in rs_pinned_alloc memory"""
    _live = []

    def __init__(self, src):
        n = max(src.nbytes, 8)
        self.ptr = lib().rs_pinned_alloc(n)
        if not self.ptr:
            raise capi.RingoError(-5, "rs_pinned_alloc failed")
        buf = (ctypes.c_uint8 * n).from_address(self.ptr)
        self.a = np.frombuffer(buf, dtype=src.dtype, count=src.size).reshape(src.shape)
        self.a[...] = src
        _PinnedArray._live.append(self)      # pinned for the life of the process (benchmark helper)


class NativeProver:
    """The same Prove as ONE C-ABI call: rs_buckler_prove (include/ringo_buckler.h) runs the round structure, both transcripts
    and the launch sequence in compiled code inside the library.  Same constructor and `prove` arguments / result as `Prover`
    (which stays as the readable mirror and as the second implementation the tests compare against)."""

    def __init__(self, field, q, circuit, jindo_params, jindo_prover):
        self.f, self.q, self.c, self.jp, self.gp = field, q, circuit, jindo_params, jindo_prover
        self.L = field.L
        self.R = 1 << (64 * self.L)
        c = circuit
        self._keep = []

        def terms(ts):
            T = (capi.Term * max(len(ts), 1))()
            for k, (ci, co, pw, ws) in enumerate(ts):
                T[k].constraint, T[k].coeff, T[k].pw, T[k].n_w = ci, co, pw, len(ws)
                for j, x in enumerate(ws):
                    T[k].w[j] = x
            self._keep.append(T)
            return ctypes.cast(T, capi.c_vp), len(ts)

        def elems(vals):
            a = self.mont_many(vals) if len(vals) else np.zeros((1, self.L), dtype=np.uint64)
            self._keep.append(a)
            return a.ctypes.data_as(capi.c_vp), len(vals)
        C = capi.BucklerCircuit()
        C.rank, C.emb_rank, C.w_cnt, C.pw_cnt = c["rank"], c["emb_rank"], c["w_cnt"], c["pw_cnt"]
        C.arith_max_rank, C.sum_max_rank = c.get("arith_max_rank", 0), c.get("sum_max_rank", 0) if c.get("sum_terms") else 0
        C.arith_terms, C.n_arith_terms = terms(c["arith_terms"])
        C.arith_coeffs, C.n_arith_coeffs = elems(c["arith_coeffs"])
        C.sum_terms, C.n_sum_terms = terms(c.get("sum_terms") or [])
        C.sum_coeffs, C.n_sum_coeffs = elems(c.get("sum_coeffs") or [])
        lin = (capi.LinChecker * max(len(c["lin"]), 1))()
        for k, (kind, pairs) in enumerate(c["lin"]):
            pa = np.ascontiguousarray(np.array(pairs, dtype=np.int32).reshape(-1, 2))
            self._keep.append(pa)
            lin[k].pairs, lin[k].n_pairs = pa.ctypes.data_as(capi.c_vp), len(pairs)
            if kind == "ntt":
                lin[k].kind = 0
            else:
                lin[k].kind, lin[k].idx_inv, lin[k].is_ntt = 1, int(kind[1]), 1 if kind[2] else 0
        self._keep.append(lin)
        C.lin, C.n_lin = ctypes.cast(lin, capi.c_vp), len(c["lin"])
        h = capi.c_vp()
        crs = bytes(jindo_prover.crs)
        check(lib().rs_buckler_create(field.h, jindo_prover.h, ctypes.byref(C), ctypes.c_char_p(crs), len(crs), ctypes.byref(h)))
        self.h = h
        self.batch = lib().rs_buckler_batch(h)
        self.timing = {}

    def __del__(self):
        try:
            if getattr(self, "h", None):
                lib().rs_buckler_destroy(self.h); self.h = None
        except Exception:
            pass

    mont_many = Prover.mont_many

    def set_device_cosac(self, on=True):
        """self-sampled proofs draw their COSAC rows on the device (the Jindo prover needs SetCosacSeeds first)"""
        check(lib().rs_buckler_set_device_cosac(self.h, 1 if on else 0))

    def enable_sampling(self, noise, threads=4):
        """rs_buckler_enable_sampling: prove_raw(pw, w, None) then draws every random value inside the library (host samplers on
        `threads` worker threads + the Jindo prover's device streams, which must have been seeded: SetTwinSeed / SetMlweSeed)."""
        Nz = capi.JindoNoise()
        for k in ("ecd_blind_std", "mask_blind_std", "mask_std", "mask_mlwe_std"):
            setattr(Nz, k, float(noise[k]))
        qa = np.array([(self.q >> (64 * i)) & ((1 << 64) - 1) for i in range(self.L)], dtype=np.uint64)
        check(lib().rs_buckler_enable_sampling(self.h, ctypes.byref(Nz), qa.ctypes.data_as(capi.c_u64p), int(threads)))

    def from_mont_arr(self, a):
        return sum(int(v) << (64 * i) for i, v in enumerate(a)) * pow(self.R, -1, self.q) % self.q

    def rand_struct(self, rand, pinned=False):
        """rs_buckler_rand over host arrays (kept alive in the returned list).  pinned: the tapes' arrays are copied into
        pinned memory once here (what a Go caller holding its tapes in rs_pinned_alloc / rs_tape_builder buffers passes) and
        rs_buckler_prove reads them in place instead of staging them on every call."""
        keep = []
        R = capi.BucklerRand()

        def el(vals):
            a = self.mont_many(vals)
            keep.append(a)
            return a.ctypes.data_as(capi.c_vp)
        R.blind = el(rand["blind"][:self.c["w_cnt"]])
        masks = list(zip(rand["mask"], rand.get("mask_sum", [])))
        if self.c["lin"]:
            m, s = masks.pop(0)
            R.lin_mask, R.lin_mask_sum = el(m), el([s])
        if self.c.get("sum_terms"):
            m, s = masks.pop(0)
            R.sum_mask, R.sum_mask_sum = el(m), el([s])
        T = (capi.JindoTape * len(rand["tapes"]))()
        for k, t in enumerate(rand["tapes"]):
            if isinstance(t, capi.JindoTape):
                T[k] = t
                continue
            if pinned:
                t = {kk: (_PinnedArray(v).a if isinstance(v, np.ndarray) else v) for kk, v in
                     ((kk, (np.ascontiguousarray(v, dtype=(np.int64 if kk in ("cosac_g", "mlwe") else np.uint64)) if v is not None and kk in
                            ("first_row", "last_row", "mask_rows", "twin_u", "cosac_g", "mlwe") else v)) for kk, v in t.items())}
            T[k], kp = self.gp.tape_struct(t)
            keep.append(kp)
        keep.append(T)
        R.tapes_pinned = 1 if pinned else 0
        R.tapes, R.n_tapes = ctypes.cast(T, capi.c_vp), len(rand["tapes"])
        return R, keep

    def prove_raw(self, pw_mont, w_mont, R, phases=True, reuse=False):
        """rs_buckler_prove on prepared inputs; returns the proof as numpy arrays (no Python-object conversion).
        phases: also fill self.timing (the library synchronises the stream at every phase boundary to do so).
        reuse: write into the buffers of the previous call (what a caller with its own proof buffers does; the returned arrays
        are overwritten by the next call)"""
        p, d, L = self.gp.p, self.gp.d, self.L
        nq, nqo, B = self.gp.n_q, self.gp.n_qout, self.batch
        out = getattr(self, "_out_bufs", None) if reuse else None
        if out is None:
            out = self._new_out_bufs(p, d, L, nq, nqo, B)
            if reuse:
                self._out_bufs = out
        return self._prove_into(out, pw_mont, w_mont, R, phases)

    @staticmethod
    def _new_out_bufs(p, d, L, nq, nqo, B):
        out = {"Witness": np.zeros((B, p["out_msis"], nqo, d), dtype=np.uint64), "Evals": np.zeros((B, L), dtype=np.uint64),
               "Partial": np.zeros((p["cols"], nq, d), dtype=np.uint64), "PartialMask": np.zeros((nq, d), dtype=np.uint64),
               "Encode": np.zeros((p["rows"], nq, d), dtype=np.uint64), "MLWE": np.zeros((p["mlwe"] + p["in_msis"], nq, d), dtype=np.uint64),
               "InCommit": np.zeros(((p["cols"] + 1) * p["in_msis"], nqo, d), dtype=np.uint64),
               "lin_mask_sum": np.zeros(L, dtype=np.uint64), "sum_mask_sum": np.zeros(L, dtype=np.uint64), "eval_point": np.zeros(L, dtype=np.uint64)}
        return out

    def _prove_into(self, out, pw_mont, w_mont, R, phases):
        P = capi.BucklerProof()
        vp = lambda a: a.ctypes.data_as(capi.c_vp)
        P.witness, P.evals, P.partial, P.partial_mask = vp(out["Witness"]), vp(out["Evals"]), vp(out["Partial"]), vp(out["PartialMask"])
        P.encode, P.mlwe, P.incommit = vp(out["Encode"]), vp(out["MLWE"]), vp(out["InCommit"])
        P.lin_mask_sum, P.sum_mask_sum, P.eval_point = vp(out["lin_mask_sum"]), vp(out["sum_mask_sum"]), vp(out["eval_point"])
        ms = (ctypes.c_double * 16)()
        pw = np.ascontiguousarray(pw_mont, dtype=np.uint64); w = np.ascontiguousarray(w_mont, dtype=np.uint64)
        check(lib().rs_buckler_prove(self.h, vp(pw) if pw.size else None, vp(w), ctypes.byref(R) if R is not None else None, ctypes.byref(P),
                                     ms if phases else None))
        self.last_raw = out
        self.timing = dict(zip(("encode_ntt", "commit", "arith_check", "lin_check", "sum_check", "evaluate", "tape_staging", "total", "ev_combine", "ev_partial",
                                "ev_incommit", "ev_respond", "ev_poly_eval"), ms))
        return out

    def verify_raw(self, pw_mont, o, in_com_dcmp_two_nm=None, res_two_nm=None):
        """rs_buckler_verify on a proof in prove_raw's layout: 0 = accepted, else the first failing check
        (1 polyVerifier.Verify, 2 arithmetic, 3 / 4 linear, 5 / 6 sum check)"""
        P = capi.BucklerProof()
        vp = lambda a: np.ascontiguousarray(a, dtype=np.uint64).ctypes.data_as(capi.c_vp)
        keep = {k: np.ascontiguousarray(o[k], dtype=np.uint64) for k in ("Witness", "Evals", "Partial", "PartialMask", "Encode", "MLWE", "InCommit",
                                                                         "lin_mask_sum", "sum_mask_sum")}
        P.witness, P.evals, P.partial, P.partial_mask = vp(keep["Witness"]), vp(keep["Evals"]), vp(keep["Partial"]), vp(keep["PartialMask"])
        P.encode, P.mlwe, P.incommit = vp(keep["Encode"]), vp(keep["MLWE"]), vp(keep["InCommit"])
        P.lin_mask_sum, P.sum_mask_sum = vp(keep["lin_mask_sum"]), vp(keep["sum_mask_sum"])
        pw = np.ascontiguousarray(pw_mont, dtype=np.uint64)
        res = ctypes.c_int(-1)
        a = self.jp.in_com_dcmp_two_nm if in_com_dcmp_two_nm is None else in_com_dcmp_two_nm
        b = self.jp.res_two_nm if res_two_nm is None else res_two_nm
        check(lib().rs_buckler_verify(self.h, vp(pw) if pw.size else None, ctypes.byref(P), float(a), float(b), ctypes.byref(res)))
        return res.value

    def prove(self, pw_mont, w_mont, rand):
        """rand = None: self-sampled (enable_sampling first)"""
        R, keep = self.rand_struct(rand) if rand is not None else (None, None)
        o = self.prove_raw(pw_mont, w_mont, R)
        has_lin, has_sum = bool(self.c["lin"]), bool(self.c.get("sum_terms"))
        return {"Witness": o["Witness"].tolist(), "LinCheckMaskSum": self.from_mont_arr(o["lin_mask_sum"]) if has_lin else None,
                "SumCheckMaskSum": self.from_mont_arr(o["sum_mask_sum"]) if has_sum else None,
                "Evals": [self.from_mont_arr(e) for e in o["Evals"]],
                "EvalProof": {k: o[k].tolist() for k in ("Partial", "PartialMask", "InCommit", "Encode", "MLWE")},
                "_eval_point": self.from_mont_arr(o["eval_point"])}
