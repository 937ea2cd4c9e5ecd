"""Host-side mirror of the parts of jindo.NewParameters / NewProver that only *derive inputs* for the
C ABI (include/ringo_jindo.h): the parameter search (jindo/params.go:126-321), the RNS primes and their
2d-th roots psi ([lattigo] ring.NewNTTFriendlyPrimesGenerator / ring.NewRing conventions, restated --
see DESIGN.md "unpinned"), the encoder constants deltaInv (jindo/encoder.go:50-67) and the TwinCDT
tables (math/csprng/gaussian_twin_cdt.go:13-35).  In a Go deployment these values come from the
reference's own host code; bench.py and examples use this module.  No device work happens here."""
import math

import numpy as np
from fractions import Fraction

MASK64 = (1 << 64) - 1
TAILCUT_CSPRNG = 9
BLOCK = 128

def _go_round(x):
    """math.Round: half away from zero."""
    # (not floor(x + 0.5): that sum rounds to even above 2^52 and turns 0.49999999999999994 into 1)
    t = float(math.trunc(x))
    if abs(x - t) >= 0.5:
        t += math.copysign(1.0, x)
    return t


def compute_cdt(center, sigma):
    """gaussian_twin_cdt.go:13-35"""
    tail_hi = int(math.ceil(TAILCUT_CSPRNG * sigma))
    tail_lo = -tail_hi
    table = []
    cdf = 0.0
    norm = math.sqrt(2 * math.pi) * sigma
    for x in range(tail_lo, tail_hi + 1):
        xf = float(x)
        rho = math.exp(-(xf - center) * (xf - center) / (2 * sigma * sigma)) / norm
        cdf += rho
        if cdf > 1:
            table.append(MASK64)
        else:
            v = _go_round(cdf * 18446744073709551616.0)
            table.append(min(int(v), MASK64) if v < 18446744073709551616.0 else MASK64)
    return table



def is_prime64(n):
    if n < 2:
        return False
    for p in (2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37):
        if n % p == 0:
            return n == p
    d, s = n - 1, 0
    while d % 2 == 0:
        d //= 2; s += 1
    for a in (2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37):      # deterministic for n < 3.3e24
        x = pow(a, d, n)
        if x in (1, n - 1):
            continue
        for _ in range(s - 1):
            x = x * x % n
            if x == n - 1:
                break
        else:
            return False
    return True


def next_upstream_primes(bits, nth_root, count):
    """[lattigo] NTTFriendlyPrimesGenerator.NextUpstreamPrimes: ascending primes 2^bits + 1 + k*nth_root."""
    out = []
    c = (1 << bits) + 1
    while len(out) < count:
        if c.bit_length() > bits + 1:
            raise ValueError("ran out of primes")
        if is_prime64(c):
            out.append(c)
        c += nth_root
    return out


def _factor(n):
    fs = set()
    d = 2
    while d * d <= n and d < 1 << 20:
        while n % d == 0:
            fs.add(d); n //= d
        d += 1 if d == 2 else 2
    if n > 1:
        # Pollard rho for the remaining cofactor
        stack = [n]
        while stack:
            m = stack.pop()
            if m == 1:
                continue
            if is_prime64(m):
                fs.add(m); continue
            c = 1
            while True:
                x = y = 2; g = 1
                while g == 1:
                    x = (x * x + c) % m
                    y = (y * y + c) % m; y = (y * y + c) % m
                    g = math.gcd(abs(x - y), m)
                if g != m:
                    break
                c += 1
            stack += [g, m // g]
    return sorted(fs)


def primitive_root(q):
    """[lattigo] ring.PrimitiveRoot: smallest generator of Z_q^*."""
    fs = _factor(q - 1)
    g = 2
    while True:
        if all(pow(g, (q - 1) // f, q) != 1 for f in fs):
            return g
        g += 1


def brv(i, bits):
    r = 0
    for _ in range(bits):
        r = (r << 1) | (i & 1)
        i >>= 1
    return r



RLWE_RANK = 1 << 13
MAX_LOG_Q = 240
ETA = 6
TAILCUT = 5


def encode_parameters(q):
    """params.go:18-40"""
    base = q - 1
    log_exp = 0
    while True:
        s = math.isqrt(base)
        if s * s != base:
            break
        base = s
        log_exp += 1
    if base >= 1 << 64:
        raise ValueError("modulus not jindo-friendly")
    return base, 1 << log_exp


def find_msis_rank(d, q, beta):
    if beta > q:
        raise ValueError("findMSISRank: beta > q")
    log_beta = math.log2(beta); log_q = math.log2(q); log_delta = math.log2(1.005)
    return int(math.ceil((log_beta * log_beta) / (4 * d * log_q * log_delta)))


class Parameters:
    pass


def new_parameters(q_field, target_n, batch):
    """params.go:126-321, float64 formulas in the reference's order."""
    if target_n < 1 or batch < 1:
        raise ValueError("NewParameters: bad arguments")
    params = Parameters()
    base, exp = encode_parameters(q_field)
    t = float(batch); b = float(base); k = float(exp)
    d = float(max(k, 256)); l = d / k
    nu = RLWE_RANK / d
    max_cols = int(math.ceil(float(target_n) / l))
    min_size = math.inf
    sqrt = math.sqrt
    nn = 1
    while nn <= max_cols:
        n = float(nn)
        m = math.ceil(float(target_n) / (n * l))
        x_one = sqrt(k) * b
        c_one = sqrt(k) * min(b, math.pow(2.0, 120 / k)) / 2
        ecd_std = 2 / (b - 1) * (b + 1) * ETA
        ecd_blind_std = 2 * x_one / (b - 1) * (b + 1) * ETA
        mask_std = 2 * c_one / (b - 1) * (b + 1) * ETA
        mask_blind_std = 2 * c_one * x_one / (b - 1) * (b + 1) * ETA
        mlwe_std = 2 * math.sqrt(2) * ETA
        mask_mlwe_std = 2 * c_one * math.sqrt(2) * ETA
        fij_inf = TAILCUT * (b + 1) * ecd_std
        f0j_inf = TAILCUT * (b + 1) * sqrt(m + 1) * ecd_blind_std
        fin_inf = TAILCUT * (b + 1) * sqrt(n + 1) * mask_std
        f0n_inf = TAILCUT * (b + 1) * sqrt((m + 1) * n + 1) * mask_blind_std
        res_ecdi_inf = sqrt(n) * c_one * fij_inf + fin_inf
        res_ecd0_inf = sqrt(n) * c_one * f0j_inf + f0n_inf
        pr_inf = sqrt(m) * x_one * fij_inf + f0j_inf
        if t > 1:
            res_ecdi_inf *= sqrt(t) * c_one
            res_ecd0_inf *= sqrt(t) * c_one
            pr_inf *= sqrt(t) * c_one
        res_ecd_two = sqrt(d * (m * res_ecdi_inf * res_ecdi_inf + res_ecd0_inf * res_ecd0_inf))
        mlwe_inf = TAILCUT * mlwe_std
        mask_mlwe_inf = TAILCUT * sqrt(n + 1) * mask_mlwe_std
        res_mlwe_inf = sqrt(n) * c_one * mlwe_inf + mask_mlwe_inf
        if t > 1:
            res_mlwe_inf *= sqrt(t) * c_one
        mu = 1
        while True:
            res_mlwe_two = sqrt(d * (float(mu) + nu)) * res_mlwe_inf
            res_two = sqrt(res_ecd_two * res_ecd_two + res_mlwe_two * res_mlwe_two)
            in_cut_two = res_two
            if t == 1:
                ext_beta = 2 * (res_two + in_cut_two); c_ext_one = 2 * c_one; d_ext_one = 1.0
            else:
                ext_beta = 2 * (2 * c_one) * (res_two + in_cut_two)
                c_ext_one = (2 * c_one) * (2 * c_one); d_ext_one = 2 * c_one
            in_msis_beta = 2 * d_ext_one * c_ext_one * ext_beta
            log_q = math.ceil(math.log2(in_msis_beta))
            q_limbs = int(math.ceil(log_q / 60.0))
            q_bits = int(math.ceil(log_q / float(q_limbs)))
            qf = math.pow(2.0, float(q_bits * q_limbs))
            if math.log2(qf) > MAX_LOG_Q:
                mu += 1
                continue
            if find_msis_rank(d, qf, in_msis_beta) == mu:
                in_msis_rank = float(mu)
                break
            mu += 1
        in_cut_inf = in_cut_two / ((1 + sqrt(n) * c_one) * sqrt(in_msis_rank * d))
        if t > 1:
            in_cut_inf /= sqrt(t) * c_one
        in_dcmp_inf = qf / in_cut_inf
        if t > 1:
            in_dcmp_inf *= sqrt(t) * c_one
        in_dcmp_two = sqrt((n + 1) * in_msis_rank * d) * in_dcmp_inf
        out_cut_two = in_dcmp_two
        out_msis_beta = 2 * d_ext_one * (2 * (in_dcmp_two + out_cut_two))
        log_qq = math.ceil(math.log2(out_msis_beta))
        qq_limbs = int(math.ceil(log_qq / 60.0))
        qq_bits = int(math.ceil(log_qq / float(qq_limbs)))
        qq = math.pow(2.0, float(qq_bits * qq_limbs))
        if math.log2(qq) > MAX_LOG_Q:
            nn <<= 1
            continue
        out_msis_rank = float(find_msis_rank(d, qq, out_msis_beta))
        out_cut_inf = out_cut_two / sqrt(out_msis_rank * d)
        if t > 1:
            out_cut_inf /= sqrt(t) * c_one
        com_size = t * out_msis_rank * d * math.log2(qq / out_cut_inf)
        pf_size = 0.0
        pf_size += n * d * math.log2(pr_inf)
        pf_size += d * math.log2(qf)
        pf_size += m * d * math.log2(res_ecdi_inf)
        pf_size += d * math.log2(res_ecd0_inf)
        pf_size += (in_msis_rank + nu) * d * math.log2(res_mlwe_inf)
        pf_size += ((n + 1) * in_msis_rank * d) * math.log2(in_dcmp_inf)
        if com_size + pf_size < min_size:
            min_size = com_size + pf_size
            p = params
            p.q_field = q_field
            p.batch = batch
            p.rank = int(n) * int(m) * int(l)
            p.rows = int(m) + 1
            p.cols = int(n)
            p.base, p.exp = base, exp
            p.slots = int(d) // exp
            p.d = int(d)
            p.in_msis = int(in_msis_rank); p.out_msis = int(out_msis_rank); p.mlwe = int(nu)
            p.log_in_cut = int(math.floor(math.log2(in_cut_inf)))
            p.log_out_cut = int(math.floor(math.log2(out_cut_inf)))
            p.in_com_dcmp_len = int((n + 1) * in_msis_rank)
            ql = int(math.ceil(math.log2(qf) / 60)); qb = int(math.ceil(math.log2(qf) / float(ql)))
            p.q_primes = next_upstream_primes(qb, 2 * int(d), ql)
            qql = int(math.ceil(math.log2(qq) / 60)); qqb = int(math.ceil(math.log2(qq) / float(qql)))
            p.qout_primes = next_upstream_primes(qqb, 2 * int(d), qql)
            p.q_bits, p.qout_bits = qb, qqb
            s2pi = sqrt(2 * math.pi)
            p.ecd_std = ecd_std / s2pi; p.ecd_blind_std = ecd_blind_std / s2pi
            p.mask_std = mask_std / s2pi; p.mask_blind_std = mask_blind_std / s2pi
            p.mlwe_std = mlwe_std / s2pi; p.mask_mlwe_std = mask_mlwe_std / s2pi
            p.res_two_nm = res_two + in_cut_two
            p.in_com_dcmp_two_nm = in_dcmp_two + out_cut_two
            p.com_size, p.pf_size = com_size, pf_size
        nn <<= 1
    return params




def psi_for(q, d):
    """[lattigo] SubRing: psi = g^((q-1)/2d) with g the smallest primitive root of Z_q^*."""
    return pow(primitive_root(q), (q - 1) // (2 * d), q)


def delta_inv(p):
    """Encoder.deltaInv (jindo/encoder.go:50-67): -(b^i)/p as float64, thresholded."""
    p_big = p.base ** p.exp + 1
    thr = math.pow(2.0, -50) / (float(p.base) * float(p.exp))
    out = []
    for i in range(p.exp):
        v = float(Fraction(-(p.base ** i), p_big))
        out.append(0.0 if abs(v) < thr else v)
    return out


def prover_inputs(p):
    """dict for ringo_snark_b200.jindo.Prover + the 128 x len TwinCDT tables (sigma = ecdStdDev)."""
    tail_hi = int(math.ceil(TAILCUT_CSPRNG * p.ecd_std))
    tables = [compute_cdt(i / BLOCK, p.ecd_std) for i in range(BLOCK)]
    d = {
        "d": p.d, "slots": p.slots, "exp": p.exp, "base": p.base, "rows": p.rows, "cols": p.cols,
        "in_msis": p.in_msis, "mlwe": p.mlwe, "out_msis": p.out_msis,
        "log_in_cut": p.log_in_cut, "log_out_cut": p.log_out_cut,
        "q": list(p.q_primes), "qout": list(p.qout_primes),
        "psi_q": [psi_for(q, p.d) for q in p.q_primes], "psi_qout": [psi_for(q, p.d) for q in p.qout_primes],
        "delta_inv": delta_inv(p), "cdt_tail_lo": -tail_hi, "cdt_sigma": p.ecd_std,
    }
    return d, tables


# ---------------------------------------------------------------------------------------------
# Host-side ring helpers of jindo.Prover.Evaluate (jindo/utils.go:21-82, encoder.go:113-146): the
# challenge and left-vector polynomials are tiny and sequentially derived from the transcript, so the
# reference computes them on the host; so does this mirror ([lattigo] NTT conventions restated).
# ---------------------------------------------------------------------------------------------
def brv(i, bits):
    r = 0
    for _ in range(bits):
        r = (r << 1) | (i & 1)
        i >>= 1
    return r


class HostRing:
    def __init__(self, d, primes, psis):
        self.d, self.primes = d, list(primes)
        logd = d.bit_length() - 1
        self.roots = []
        for q, psi in zip(primes, psis):
            r = [0] * d
            pw = 1
            for j in range(d):
                r[brv(j, logd)] = pw
                pw = pw * psi % q
            self.roots.append(r)

    def ntt_mform(self, poly):
        """NTT(MForm(poly)): poly = [limb][coeff] canonical ints."""
        out = []
        for q, roots, c in zip(self.primes, self.roots, poly):
            a = [x * (1 << 64) % q for x in c]
            t = self.d; m = 1
            while m < self.d:
                t >>= 1
                for i in range(m):
                    w = roots[m + i]
                    for j in range(2 * i * t, 2 * i * t + t):
                        v = a[j + t] * w % q
                        u = a[j]
                        a[j] = (u + v) % q
                        a[j + t] = (u - v) % q
                m <<= 1
            out.append(a)
        return out


def div_mod64(x, y):
    r = 0
    for i in range(len(x) - 1, -1, -1):
        cur = (r << 64) | x[i]
        x[i] = cur // y
        r = cur % y
    return r


def encode_challenge(p, ring, chal_bytes):
    """jindo/utils.go:21-46"""
    poly = [[0] * p.d for _ in ring.primes]
    c = [int.from_bytes(chal_bytes[:8], "big"), int.from_bytes(chal_bytes[8:16], "big")]
    bound = min(p.base, 1 << (120 // p.exp)) // 2
    for i in range(p.exp):
        r = div_mod64(c, bound)
        for k, q in enumerate(ring.primes):
            poly[k][i * p.slots] = (q - (bound - r)) if r > bound // 2 else r
    return ring.ntt_mform(poly)


def encode_elements(p, ring, q_field, values):
    """Encoder.encode (jindo/encoder.go:106-146) of up to `slots` canonical field values."""
    L = (q_field.bit_length() + 63) // 64
    coeffs = [0] * p.d
    for i, x in enumerate(values):
        limbs = [(x >> (64 * k)) & MASK64 for k in range(L)]
        for j in range(p.exp - 1):
            coeffs[j * p.slots + i] = div_mod64(limbs, p.base)
        coeffs[(p.exp - 1) * p.slots + i] = limbs[0]
    return ring.ntt_mform([list(coeffs) for _ in ring.primes])


def left_vec(p, q_field, x):
    """jindo/utils.go:63-73"""
    skip = pow(x, p.cols * p.slots, q_field)
    left = [1]
    for _ in range(1, p.rows):
        left.append(left[-1] * skip % q_field)
    left[p.rows - 1] = x
    return left


def poly_write_to(poly):
    """[lattigo] ring.Poly.WriteTo framing (restated; see DESIGN.md "unpinned")."""
    import struct
    out = [struct.pack("<Q", len(poly))]
    for c in poly:
        out.append(struct.pack("<Q", len(c)))
        out.append(struct.pack("<%dQ" % len(c), *[int(v) for v in c]))
    return b"".join(out)


def commitment_write_raw(p, value):
    """Commitment.WriteRawTo (jindo/entities.go:85-100): polys allocated with ringQ's limb count."""
    nl = len(p.q_primes)
    out = []
    if isinstance(value, np.ndarray):                      # [out_msis, n_qout, d] uint64: same bytes without int lists
        import struct
        head, lim = struct.pack("<Q", max(nl, value.shape[1])), struct.pack("<Q", p.d)
        for v in value:
            out.append(head)
            for l in range(max(nl, value.shape[1])):
                out.append(lim)
                out.append(np.ascontiguousarray(v[l], dtype="<u8").tobytes() if l < value.shape[1] else bytes(8 * p.d))
        return b"".join(out)
    for v in value:
        limbs = [list(c) for c in v] + [[0] * p.d for _ in range(nl - len(v))]
        out.append(poly_write_to(limbs))
    return b"".join(out)
