"""Host-side mirror of jindo.Prover (jindo/prover.go) over the C ABI of include/ringo_jindo.h.

`Prover.Commit(v)` / `Prover.Evaluate(...)` keep the reference's names and argument meaning.
Everything that is a host-library convention or a random draw in the reference is an explicit
input here (`params`: primes, psi, deltaInv, CDT tables; `tape`: the randomness of one Commit),
exactly as the Go shim would pass them.
"""
import ctypes

import numpy as np

from . import capi
from .capi import check, lib, c_vp


def _vp(a):
    return a.ctypes.data_as(c_vp)


def params_struct(p, cdt_len):
    """rs_jindo_params from the plain dict the host glue / the tests carry around"""
    P = capi.JindoParams()
    for k in ("d", "slots", "exp", "base", "rows", "cols", "in_msis", "mlwe", "out_msis", "log_in_cut", "log_out_cut"):
        setattr(P, k, int(p[k]))
    P.n_q, P.n_qout = len(p["q"]), len(p["qout"])
    for i, v in enumerate(p["q"]): P.q[i] = v
    for i, v in enumerate(p["qout"]): P.qout[i] = v
    for i, v in enumerate(p["psi_q"]): P.psi_q[i] = v
    for i, v in enumerate(p["psi_qout"]): P.psi_qout[i] = v
    for i, v in enumerate(p["delta_inv"]): P.delta_inv[i] = v
    P.cdt_len = int(cdt_len)
    P.cdt_tail_lo = int(p["cdt_tail_lo"])
    P.cdt_sigma = float(p["cdt_sigma"])
    return P


def key48(seed):
    """SHA-384(seed) = AES key || IV of a csprng.UniformSampler (math/csprng/uniform.go:38-46)"""
    import hashlib
    return hashlib.sha384(bytes(seed)).digest()


def compute_cdt(center, sigma):
    """computeCDT (gaussian_twin_cdt.go:13-35) through the library: (table, tail_lo)"""
    import math
    cap = 2 * int(math.ceil(9 * sigma)) + 1
    tab = np.zeros(cap, dtype=np.uint64)
    lo = ctypes.c_int64()
    n = lib().rs_compute_cdt(float(center), float(sigma), _vp(tab), cap, ctypes.byref(lo))
    if n < 0:
        raise capi.RingoError(n, "rs_compute_cdt")
    return tab[:n], lo.value


class TapeBuilder:
    """rs_tape_builder (include/ringo_sampler.h): the host half of a Commit's randomness tape -- COSAC rows, the mask
    column's rounded-Gaussian MLWE noise and the MustSetRandom elements -- produced on worker threads.
    noise: dict with ecd_blind_std, mask_blind_std, mask_std, mask_mlwe_std (jindo/params.go:303-309).
    seeds: None (OS randomness, any number of threads) or three 32-byte sampler seeds in NewProver's order
    (encoder COSAC base, COSAC rounded, prover rounded; one thread: the reference's sequential streams)."""

    def __init__(self, p, cdt_len, noise, q_field, threads=1, seeds=None):
        self.p = p
        self.L = (q_field.bit_length() + 63) // 64
        P = params_struct(p, cdt_len)
        Nz = capi.JindoNoise()
        for k in ("ecd_blind_std", "mask_blind_std", "mask_std", "mask_mlwe_std"):
            setattr(Nz, k, float(noise[k]))
        qa = np.array([(q_field >> (64 * i)) & ((1 << 64) - 1) for i in range(self.L)], dtype=np.uint64)
        kb = b"".join(key48(s) for s in seeds) if seeds is not None else None
        h = c_vp()
        check(lib().rs_tape_builder_create(ctypes.byref(P), ctypes.byref(Nz), qa.ctypes.data_as(capi.c_u64p), self.L, int(threads),
                                           ctypes.c_char_p(kb) if kb else None, ctypes.byref(h)))
        self.h = h
        self._keep = None

    def __del__(self):
        try:
            if getattr(self, "h", None):
                lib().rs_tape_builder_destroy(self.h); self.h = None
        except Exception:
            pass

    def set_device_cosac(self, on=True):
        check(lib().rs_tape_builder_set_device_cosac(self.h, 1 if on else 0))

    def set_random_bytes(self, data):
        self._rand = ctypes.create_string_buffer(bytes(data), len(data))
        check(lib().rs_tape_builder_set_random_bytes(self.h, ctypes.cast(self._rand, c_vp), len(data)))

    def start(self, v_host, length, slot=0):
        """v_host: numpy uint64 array [>= length, L] (kept alive until wait) or a raw host pointer"""
        if isinstance(v_host, np.ndarray):
            self._keep = np.ascontiguousarray(v_host, dtype=np.uint64)
            ptr = _vp(self._keep)
        else:
            ptr = c_vp(int(v_host))
        check(lib().rs_tape_builder_start(self.h, ptr, int(length), int(slot)))

    def wait(self, slot=0):
        """-> (rs_jindo_tape struct pointing into the builder's staging buffers, host milliseconds)"""
        T = capi.JindoTape()
        ms = ctypes.c_double()
        check(lib().rs_tape_builder_wait(self.h, int(slot), ctypes.byref(T), ctypes.byref(ms)))
        return T, ms.value

    def arrays(self, T):
        """numpy copies of a finished tape (tests)"""
        d, L, slots = self.p["d"], self.L, self.p["slots"]
        cs = self.p["cols"] * slots
        nm = self.p["in_msis"] + self.p["mlwe"]

        def arr(ptr, n, dt):
            return np.frombuffer((ctypes.c_uint64 * n).from_address(ptr), dtype=dt).copy()
        return {"first_row": arr(T.first_row, cs * L, np.uint64).reshape(cs, L), "last_row": arr(T.last_row, cs * L, np.uint64).reshape(cs, L),
                "mask_rows": arr(T.mask_rows, T.n_mask_rows * slots * L, np.uint64).reshape(-1, L),
                "cosac_g": arr(T.cosac_g, T.n_cosac_polys * d, np.int64).reshape(-1, d),
                "mlwe_mask": arr(T.mlwe, nm * d, np.int64).reshape(nm, d)}


class Opening:
    def __init__(self, prover, handle):
        self.prover = prover
        self.h = handle

    def __del__(self):
        try:
            if getattr(self, "h", None):
                lib().rs_opening_free(self.h)
        except Exception:
            pass

    def _get(self, kind, col, idx, n_limbs):
        out = np.zeros((n_limbs, self.prover.d), dtype=np.uint64)
        check(lib().rs_opening_download(self.h, kind, col, idx, _vp(out)))
        return out

    def Encode(self, col, row): return self._get(0, col, row, self.prover.n_q)
    def MLWE(self, col, idx): return self._get(1, col, idx, self.prover.n_q)
    def InCommit(self, idx): return self._get(2, 0, idx, self.prover.n_qout)

    def incommit_device_ptr(self):
        """Device address of Opening.InCommit: (cols+1)*in_msis polynomials of n_qout*256 uint64 (for the
        all-gather of a column-sharded Commit)."""
        return int(lib().rs_opening_incommit_device(self.h) or 0)


class Prover:
    """jindo.NewProver (prover.go:28-40).  `p` is a dict with the jindo.Parameters fields, the
    RNS primes / psi, deltaInv and the TwinCDT tables; `ck_*` are the commit-key polynomials as
    uint64 arrays [.., limb, 256]."""

    def __init__(self, field, p, ck_in, ck_mlwe, ck_out, cdt_tables, crs=b"", key_from_crs=False):
        """key_from_crs: ignore ck_* and evaluate NewCommitKey(params, crs) on the device (rs_jindo_create_crs)."""
        self.field = field
        self.crs = bytes(crs)           # CommitKey.WriteRawTo (entities.go:75-77): bound into the Evaluate transcript
        self.p = dict(p)
        self.d = p["d"]
        self.n_q, self.n_qout = len(p["q"]), len(p["qout"])
        cdt = np.ascontiguousarray(cdt_tables, dtype=np.uint64)
        P = params_struct(p, cdt.shape[1])
        h = c_vp()
        if key_from_crs:
            import hashlib
            dg = hashlib.sha384(self.crs).digest()                     # uniform.go:38-46: key = h[:32], iv = h[32:48]
            check(lib().rs_jindo_create_crs(field.h, ctypes.byref(P), ctypes.c_char_p(dg[:32]), ctypes.c_char_p(dg[32:48]),
                                            _vp(cdt), ctypes.byref(h)))
        else:
            ck_in = np.ascontiguousarray(ck_in, dtype=np.uint64); ck_mlwe = np.ascontiguousarray(ck_mlwe, dtype=np.uint64)
            ck_out = np.ascontiguousarray(ck_out, dtype=np.uint64)
            check(lib().rs_jindo_create(field.h, ctypes.byref(P), _vp(ck_in), _vp(ck_mlwe), _vp(ck_out), _vp(cdt), ctypes.byref(h)))
        self.h = h

    def CommitKey(self):
        """(In, MLWE, Out) as uint64 arrays [i][j][limb][256] (jindo.CommitKey, entities.go:15-19), read back from HBM."""
        p = self.p
        shapes = [(p["in_msis"], p["rows"], self.n_q, self.d), (p["in_msis"], p["mlwe"], self.n_q, self.d),
                  (p["out_msis"], (p["cols"] + 1) * p["in_msis"], self.n_qout, self.d)]
        out = []
        for which, shp in enumerate(shapes):
            a = np.zeros(shp, dtype=np.uint64)
            check(lib().rs_jindo_commit_key_download(self.h, which, _vp(a)))
            out.append(a)
        return tuple(out)

    def __del__(self):
        try:
            if getattr(self, "h", None):
                lib().rs_jindo_destroy(self.h)
        except Exception:
            pass

    def SetTwinSeed(self, seed):
        """Seeds the encoder's TwinCDT uniform source on the device, as csprng.NewUniformSamplerWithSeed
        does on the host (uniform.go:38-46): key || iv = SHA-384(seed).  Commits whose tape has no
        `twin_u` then draw from this stream in HBM instead of uploading one u64 per sample."""
        import hashlib
        h = hashlib.sha384(bytes(seed)).digest()
        key = np.frombuffer(h[:32], dtype=np.uint8).copy(); iv = np.frombuffer(h[32:48], dtype=np.uint8).copy()
        check(lib().rs_jindo_set_twin_stream(self.h, _vp(key), _vp(iv)))

    def Verify(self, x_mont, coms, batch_bytes, chal_bytes, ys_mont, pf, in_com_dcmp_two_nm, res_two_nm):
        """Verifier.Verify's arithmetic on the device (rs_jindo_verify; jindo/verifier.go:50-282).  coms: [batch, out_msis,
        n_qout, 256]; pf: dict Partial / PartialMask / Encode / MLWE / InCommit (as Evaluate returns them); the SHAKE128
        challenge bytes come from the caller's transcript.  Returns the check bits (15 = accepted)."""
        arr = lambda a: np.ascontiguousarray(a, dtype=np.uint64)
        keep = [arr(pf["Partial"]), arr(pf["PartialMask"]), arr(pf["Encode"]), arr(pf["MLWE"]), arr(pf["InCommit"]), arr(coms),
                arr(x_mont), arr(ys_mont)]
        P = capi.JindoProof()
        P.partial, P.partial_mask, P.encode, P.mlwe, P.incommit = (_vp(a) for a in keep[:5])
        out = ctypes.c_int(0)
        batch = keep[5].shape[0]
        check(lib().rs_jindo_verify(self.h, _vp(keep[6]), _vp(keep[5]), batch, ctypes.c_char_p(bytes(batch_bytes)) if batch > 1 else None,
                                    ctypes.c_char_p(bytes(chal_bytes)), _vp(keep[7]), ctypes.byref(P), float(in_com_dcmp_two_nm),
                                    float(res_two_nm), ctypes.byref(out)))
        return out.value

    def CommitmentWriteRaw(self, com):
        """Commitment.WriteRawTo bytes (jindo/entities.go:96-100) through the C ABI (rs_jindo_commitment_write_raw)"""
        com = np.ascontiguousarray(com, dtype=np.uint64)
        n = lib().rs_jindo_commitment_write_raw(self.h, _vp(com), None, 0)
        buf = ctypes.create_string_buffer(int(n))
        got = lib().rs_jindo_commitment_write_raw(self.h, _vp(com), ctypes.cast(buf, c_vp), int(n))
        if got != n:
            raise capi.RingoError(int(got), "rs_jindo_commitment_write_raw")
        return buf.raw

    def SetMlweSeed(self, seed, mlwe_std):
        """Prover.mlweSampler on the device (rs_jindo_set_mlwe_stream): seed of its UniformSampler, sigma = params.mlweStdDev"""
        tab, lo = compute_cdt(0.0, mlwe_std)
        dg = key48(seed)
        check(lib().rs_jindo_set_mlwe_stream(self.h, ctypes.c_char_p(dg[:32]), ctypes.c_char_p(dg[32:48]), _vp(tab), len(tab), lo))

    def SetCosacSeeds(self, seed_base, seed_rounded, noise):
        """Encoder.cosac on the device (rs_jindo_set_cosac_stream): seeds of its two UniformSamplers (hashed with SHA-384 like
        NewUniformSamplerWithSeed); noise: dict with ecd_blind_std, mask_blind_std, mask_std"""
        kb, kr = key48(seed_base), key48(seed_rounded)
        self._cosac_keys = (kb, kr)
        check(lib().rs_jindo_set_cosac_stream(self.h, ctypes.c_char_p(kb), ctypes.c_char_p(kr), float(noise["ecd_blind_std"]),
                                              float(noise["mask_blind_std"]), float(noise["mask_std"])))

    def CosacStatus(self):
        """(workers handed out so far, sticky overflow flag)"""
        w, ov = ctypes.c_uint64(), ctypes.c_int()
        check(lib().rs_jindo_cosac_status(self.h, ctypes.byref(w), ctypes.byref(ov)))
        return w.value, ov.value

    def CosacSample(self, first_row, mask_rows):
        """the COSAC polynomials a Commit with these rows draws: (g [cols + n_mask_rows, 256], first worker index, overflow flag)"""
        fr = np.ascontiguousarray(first_row, dtype=np.uint64); mr = np.ascontiguousarray(mask_rows, dtype=np.uint64)
        n_mask = mr.size // (self.field.L * self.p["slots"])
        out = np.zeros((self.p["cols"] + n_mask, self.d), dtype=np.int64)
        w0, ov = ctypes.c_uint64(), ctypes.c_int()
        check(lib().rs_jindo_cosac_sample(self.h, _vp(fr), _vp(mr), n_mask, _vp(out), ctypes.byref(w0), ctypes.byref(ov)))
        return out, w0.value, ov.value

    def MlweStreamSample(self, n):
        out = np.zeros(n, dtype=np.int64)
        check(lib().rs_jindo_mlwe_stream_sample(self.h, n, _vp(out)))
        return out

    def TwinStreamDraw(self, n):
        """The next n Sample() values of the device stream (advances it)."""
        out = np.zeros(n, dtype=np.uint64)
        check(lib().rs_jindo_twin_stream_draw(self.h, n, _vp(out)))
        return out

    def Commit(self, batch, poly, length, tape):
        """prover.go:45-62.  batch/poly/length: device-resident committed vector; tape: dict of
        numpy arrays first_row, last_row, mask_rows (uint64 Montgomery limbs), twin_u (uint64),
        cosac_g, mlwe (int64).  Returns (commitment [out_msis, n_qout, 256], Opening)."""
        if isinstance(tape, capi.JindoTape):       # e.g. straight from TapeBuilder.wait()
            T = tape
        else:
            T, keep = self.tape_struct(tape)
        com = np.zeros((self.p["out_msis"], self.n_qout, self.d), dtype=np.uint64)
        oh = c_vp()
        check(lib().rs_jindo_commit(self.h, batch.h, poly, length, ctypes.byref(T), ctypes.byref(oh), _vp(com)))
        return com, Opening(self, oh)

    def CommitInto(self, batch, poly, length, tape_struct, opening, com_out):
        """Same call with a prebuilt rs_jindo_tape (host or device pointers) that overwrites an existing
        Opening -- the allocation-free form used by bench.py."""
        oh = c_vp(opening.h.value)
        check(lib().rs_jindo_commit(self.h, batch.h, poly, length, ctypes.byref(tape_struct), ctypes.byref(oh), _vp(com_out)))

    def tape_struct(self, tape):
        """rs_jindo_tape over host numpy arrays (returns the struct and the arrays that must stay alive)."""
        keep = {k: np.ascontiguousarray(tape[k], dtype=(np.int64 if k in ("cosac_g", "mlwe") else np.uint64))
                for k in ("first_row", "last_row", "mask_rows", "twin_u", "cosac_g", "mlwe") if tape.get(k) is not None}
        L = self.field.L
        T = capi.JindoTape()
        T.first_row, T.last_row, T.mask_rows = _vp(keep["first_row"]), _vp(keep["last_row"]), _vp(keep["mask_rows"])
        T.n_mask_rows = keep["mask_rows"].size // (L * self.p["slots"])
        if "twin_u" in keep:
            T.twin_u, T.n_twin_u = _vp(keep["twin_u"]), keep["twin_u"].size
        if "cosac_g" in keep:
            T.cosac_g, T.n_cosac_polys = _vp(keep["cosac_g"]), keep["cosac_g"].size // self.d
        T.mlwe = _vp(keep["mlwe"])
        T.mlwe_mask_only = 1 if tape.get("mlwe_mask_only") else 0
        return T, keep

    def CommitColumns(self, batch, poly, length, tape, col_begin, col_end, opening=None):
        """rs_jindo_commit_columns: commitColTo (prover.go:89-177) for the columns [col_begin, col_end) only,
        with the whole tape.  Returns the Opening (its other columns are untouched / zero)."""
        T, keep = self.tape_struct(tape) if isinstance(tape, dict) else (tape, None)
        oh = c_vp(opening.h.value) if opening is not None else c_vp()
        check(lib().rs_jindo_commit_columns(self.h, batch.h, poly, length, ctypes.byref(T), col_begin, col_end, ctypes.byref(oh)))
        return opening if opening is not None else Opening(self, oh)

    def OuterCommit(self, opening):
        """rs_jindo_outer_commit: Commitment.Value from the (gathered) Opening.InCommit (prover.go:180-202)."""
        com = np.zeros((self.p["out_msis"], self.n_qout, self.d), dtype=np.uint64)
        check(lib().rs_jindo_outer_commit(self.h, opening.h, _vp(com)))
        return com

    def CommitHostInto(self, v_host_ptr, length, tape_struct, opening, com_out):
        """rs_jindo_commit_host: v in (pinned) host memory, uploaded piecewise behind the encode kernels."""
        oh = c_vp(opening.h.value)
        check(lib().rs_jindo_commit_host(self.h, c_vp(v_host_ptr), length, ctypes.byref(tape_struct), ctypes.byref(oh), _vp(com_out)))

    def Partial(self, opening, left):
        """prover.go:281-289: returns [cols+1, n_q, 256]; index cols is PartialMask."""
        left = np.ascontiguousarray(left, dtype=np.uint64)
        out = np.zeros((self.p["cols"] + 1, self.n_q, self.d), dtype=np.uint64)
        check(lib().rs_jindo_partial(self.h, opening.h, _vp(left), _vp(out)))
        return out

    def Respond(self, opening, chals):
        """prover.go:304-316: returns (Encode [rows,..], MLWE [mlwe+in_msis,..])."""
        chals = np.ascontiguousarray(chals, dtype=np.uint64)
        enc = np.zeros((self.p["rows"], self.n_q, self.d), dtype=np.uint64)
        ml = np.zeros((self.p["mlwe"] + self.p["in_msis"], self.n_q, self.d), dtype=np.uint64)
        check(lib().rs_jindo_respond(self.h, opening.h, _vp(chals), _vp(enc), _vp(ml)))
        return enc, ml

    def PartialAt(self, opening, x_mont):
        """prover.go:267-289 with leftVec(x) encoded on the device; x_mont: L64 Montgomery limbs of the evaluation point."""
        xm = np.ascontiguousarray(x_mont, dtype=np.uint64)
        out = np.zeros((self.p["cols"] + 1, self.n_q, self.d), dtype=np.uint64)
        check(lib().rs_jindo_partial_at(self.h, opening.h, _vp(xm), _vp(out)))
        return out

    def RespondBytes(self, opening, chal_bytes, out=None):
        """prover.go:296-316 with encodeChallengeTo on the device; chal_bytes: 16 * cols bytes of the SHAKE output.
        out: optional (enc, mlwe) uint64 arrays to fill (e.g. views of pinned memory: the 8 MB of responses at 2^22
        then come back at PCIe speed instead of through the driver's pageable staging)."""
        assert len(chal_bytes) == 16 * self.p["cols"]
        if out is not None:
            enc, ml = out
            assert enc.shape == (self.p["rows"], self.n_q, self.d) and ml.shape == (self.p["mlwe"] + self.p["in_msis"], self.n_q, self.d)
            assert enc.dtype == np.uint64 and ml.dtype == np.uint64 and enc.flags.c_contiguous and ml.flags.c_contiguous
        else:
            enc = np.zeros((self.p["rows"], self.n_q, self.d), dtype=np.uint64)
            ml = np.zeros((self.p["mlwe"] + self.p["in_msis"], self.n_q, self.d), dtype=np.uint64)
        check(lib().rs_jindo_respond_bytes(self.h, opening.h, ctypes.c_char_p(bytes(chal_bytes)), _vp(enc), _vp(ml)))
        return enc, ml

    def LeftVec(self, x_mont):
        xm = np.ascontiguousarray(x_mont, dtype=np.uint64)
        out = np.zeros((self.p["rows"], self.n_q, self.d), dtype=np.uint64)
        check(lib().rs_jindo_left_vec(self.h, _vp(xm), _vp(out)))
        return out

    def EncodeChallenges(self, chal_bytes, ring_out=False):
        n = len(chal_bytes) // 16
        out = np.zeros((n, self.n_qout if ring_out else self.n_q, self.d), dtype=np.uint64)
        check(lib().rs_jindo_encode_challenges(self.h, ctypes.c_char_p(bytes(chal_bytes)), n, 1 if ring_out else 0, _vp(out)))
        return out

    def OpeningMulAddBytes(self, acc, opening, chal16):
        """prover.go:230-265: acc += opening (.) encodeChallenge(16 bytes), both rings, encoded on the device."""
        check(lib().rs_opening_mul_add_bytes(self.h, acc.h, opening.h, ctypes.c_char_p(bytes(chal16))))

    def OpeningCombineBytes(self, openings, chal_bytes):
        """prover.go:230-265 in one pass (rs_opening_combine_bytes): sum_i openings[i] (.) encodeChallenge(16 bytes each)."""
        assert len(chal_bytes) == 16 * len(openings)
        hs = (c_vp * len(openings))(*[o.h.value for o in openings])
        oh = c_vp()
        check(lib().rs_opening_combine_bytes(self.h, hs, len(openings), ctypes.c_char_p(bytes(chal_bytes)), ctypes.byref(oh)))
        return Opening(self, oh)

    def NewOpening(self):
        oh = c_vp()
        check(lib().rs_opening_alloc_zero(self.h, ctypes.byref(oh)))
        return Opening(self, oh)

    def OpeningMulAdd(self, acc, opening, chal_q, chal_qout):
        """prover.go:251-265: acc += opening (.) challenge (both rings)."""
        cq = np.ascontiguousarray(chal_q, dtype=np.uint64); co = np.ascontiguousarray(chal_qout, dtype=np.uint64)
        check(lib().rs_opening_mul_add(self.h, acc.h, opening.h, _vp(cq), _vp(co)))
