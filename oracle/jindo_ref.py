"""oracle/jindo_ref.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Python restatement of the Jindo polynomial-commitment hot path of the reference
(sp301415/ringo-snark, paths relative to its root):

  math/csprng/uniform.go, gaussian_twin_cdt.go, gaussian_cosac.go, gaussian_rounded.go
  jindo/params.go, entities.go, encoder.go, rns.go, utils.go, prover.go, verifier.go

plus the parts of github.com/tuneinsight/lattigo/v6 v6.1.0 `ring` that Jindo calls
(prime generation, NTT, Montgomery forms, Poly.WriteTo).  lattigo is NOT under
/root/reference and there is no Go toolchain: its conventions are restated from the
published algorithm and flagged `# [lattigo]`.  PARITY UNPINNED for every byte that depends
on them (RNS prime values, psi, serialization, hence SHAKE challenges); value-level
semantics (which ring element is committed; Verify(Prove) == true, the only thing the
reference's own tests assert, jindo/jindo_test.go:50-51) are pinned by the algebra and
checked in tests/test_oracle_jindo.py.  Float transcendental last-ulp equality with Go's
math.Exp/Log is likewise unpinned.

Randomness: the reference draws from crypto/rand.  Here every draw comes from an explicit
`Tape` object consumed in the reference's exact call order, so the GPU path can
be fed the same samples.
"""
import hashlib
import math
import struct
from fractions import Fraction

from cryptography.hazmat.primitives.ciphers import Cipher, algorithms, modes

MASK64 = (1 << 64) - 1

# ---------------------------------------------------------------------------------------------
# math/csprng/uniform.go
# ---------------------------------------------------------------------------------------------


class UniformSampler:
    BUF = 8192  # uniform.go:12

    def __init__(self, seed):
        self.seed = bytes(seed)
        h = hashlib.sha384(bytes(seed)).digest()            # uniform.go:38-39
        self._enc = Cipher(algorithms.AES(h[:32]), modes.CTR(h[32:48])).encryptor()   # :41-46
        self.buf = bytearray(self.BUF)
        self.ptr = self.BUF

    def read(self, n):
        """Read(p): XORKeyStream(p, p) over zero-initialised caller bytes (uniform.go:58-61)."""
        return self._enc.update(bytes(n))

    def sample(self):
        if self.ptr == self.BUF:
            # uniform.go:64-70: XORKeyStream(buf, buf) on a never-cleared buffer (quirk: XOR-accumulates)
            self.buf = bytearray(self._enc.update(bytes(self.buf)))
            self.ptr = 0
        r = int.from_bytes(self.buf[self.ptr:self.ptr + 8], "little")
        self.ptr += 8
        return r

    def sample_many(self, n):
        """n consecutive Sample() values as a numpy uint64 array (same stream positions)."""
        import numpy as np
        out = np.empty(n, dtype=np.uint64)
        i = 0
        while i < n:
            if self.ptr == self.BUF:
                self.buf = bytearray(self._enc.update(bytes(self.buf)))
                self.ptr = 0
            take = min(n - i, (self.BUF - self.ptr) // 8)
            out[i:i + take] = np.frombuffer(bytes(self.buf[self.ptr:self.ptr + 8 * take]), dtype="<u8")
            self.ptr += 8 * take
            i += take
        return out

    def sample_n(self, n):
        bound = MASK64 - MASK64 % n                         # uniform.go:87-95
        while True:
            r = self.sample()
            if r < bound:
                return r % n

    def sample_float(self):
        r = self.sample() % (1 << 52)                       # uniform.go:96-100
        rf = struct.unpack("<d", struct.pack("<Q", r | ((1023 + 52) << 52)))[0]
        return (rf / float(1 << 52)) - 1


class Tape:
    """Deterministic stand-in for crypto/rand.Reader: SHAKE128(seed) byte stream."""

    def __init__(self, seed):
        self._x = hashlib.shake_128(bytes(seed))
        self._pos = 0
        self._cache = b""

    def read(self, n):
        need = self._pos + n
        if need > len(self._cache):
            self._cache = self._x.digest(max(need, 2 * len(self._cache), 1 << 16))
        out = self._cache[self._pos:need]
        self._pos = need
        return out


# ---------------------------------------------------------------------------------------------
# math/csprng/gaussian_*.go
# ---------------------------------------------------------------------------------------------
TAILCUT_CSPRNG = 9
BLOCK = 128
FLOAT_PREC = 52
RN = 3.442619855899


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


def _binary_search(tab, u):
    """slices.BinarySearch: (leftmost index with tab[i] >= u, found)."""
    lo, hi = 0, len(tab)
    while lo < hi:
        mid = (lo + hi) // 2
        if tab[mid] < u:
            lo = mid + 1
        else:
            hi = mid
    return lo, (lo < len(tab) and tab[lo] == u)


class TwinCDTGaussianSampler:
    def __init__(self, std_dev, seed):
        self.base = UniformSampler(seed)
        self.std_dev = std_dev
        self.tables = [compute_cdt(i / BLOCK, std_dev) for i in range(BLOCK)]    # :47-50
        self.tail_hi = int(math.ceil(TAILCUT_CSPRNG * std_dev))
        self.tail_lo = -self.tail_hi

    def sample_with_u(self, center, u):
        """gaussian_twin_cdt.go:77-111 with the uniform draw passed in."""
        c_floor = math.floor(center)
        c_frac = center - c_floor
        c0 = int(math.floor(BLOCK * c_frac)) % BLOCK
        c1 = int(math.ceil(BLOCK * c_frac)) % BLOCK
        v0, ok = _binary_search(self.tables[c0], u)
        if ok:
            v0 -= 1
        v1, ok = _binary_search(self.tables[c1], u)
        if ok:
            v1 -= 1
        if v0 == v1:
            return v0 + int(c_floor) + self.tail_lo
        cdf = 0.0
        norm = math.sqrt(2 * math.pi) * self.std_dev
        for x in range(self.tail_lo, v0 + 1):
            xf = float(x)
            cdf += math.exp(-(xf - c_frac) * (xf - c_frac) / (2 * self.std_dev * self.std_dev)) / norm
        p = float(u) / 18446744073709551616.0
        if p < cdf:
            return v0 + self.tail_lo + int(c_floor)
        return v1 + self.tail_lo + int(c_floor)

    def sample(self, center):
        return self.sample_with_u(center, self.base.sample())


def _ziggurat_tables():
    """gaussian_rounded.go:22-39 (package init)."""
    def normal(x): return math.exp(-0.5 * x * x)
    def normal_integral(x): return math.sqrt(math.pi / 2) * math.erfc(x / math.sqrt(2))
    def normal_inv(x): return math.sqrt(-2 * math.log(x))
    v = RN * normal(RN) + normal_integral(RN)
    xn = [0.0] * BLOCK
    xn[BLOCK - 1] = RN
    for i in range(BLOCK - 2, 0, -1):
        xn[i] = normal_inv(v / xn[i + 1] + normal(xn[i + 1]))
    scale = float(1 << FLOAT_PREC)
    kn = [0] * BLOCK; wn = [0.0] * BLOCK; fn = [0.0] * BLOCK
    for i in range(1, BLOCK):
        kn[i] = int((xn[i - 1] / xn[i]) * scale)
        wn[i] = xn[i] / scale
        fn[i] = normal(xn[i])
    kn[0] = int((RN * normal(RN) / v) * scale)
    wn[0] = (v / normal(RN)) / scale
    return kn, wn, fn


_KN, _WN, _FN = _ziggurat_tables()


class RoundedGaussianSampler:
    def __init__(self, seed):
        self.base = UniformSampler(seed)

    def norm_float(self):
        """gaussian_rounded.go:77-117"""
        while True:
            r = self.base.sample()
            b = r >> 63
            i = r % (1 << 7)
            j = (r >> 7) % (1 << FLOAT_PREC)
            sj = -j if b else j                              # int64((j ^ -b) + b)
            x = float(sj) * _WN[i]
            if j < _KN[i]:
                return x
            if i == 0:
                while True:
                    u = -math.log(self.base.sample_float()) * (1.0 / RN)
                    v = -math.log(self.base.sample_float())
                    if v + v >= u * u:
                        break
                u += RN
                return -u if b == 1 else u
            f0, f1 = _FN[i - 1], _FN[i]
            if self.base.sample_float() * (f0 - f1) < math.exp(-0.5 * x * x) - f1:
                return x

    def sample(self, center, std_dev):
        return int(_go_round(center + self.norm_float() * std_dev))   # :119-125


class COSACSampler:
    def __init__(self, seed_base, seed_rounded):
        self.base = UniformSampler(seed_base)
        self.rounded = RoundedGaussianSampler(seed_rounded)

    def _sample_round(self, c_frac, std_dev):
        """gaussian_cosac.go:23-46"""
        while True:
            y = std_dev * self.rounded.norm_float()
            b = self.base.sample() & 1
            if b == 0:
                y_round = _go_round(y) - 1
                cmp_ = y_round <= 0.5
            else:
                y_round = _go_round(y) + 1
                cmp_ = y_round >= -0.5
            if cmp_:
                r = self.base.sample_float()
                if r < math.exp(-((y_round + c_frac) * (y_round + c_frac) - y * y) / (2 * std_dev * std_dev)):
                    return int(y_round)

    def sample(self, center, std_dev):
        """gaussian_cosac.go:48-57"""
        c_int = _go_round(center)
        c_frac = c_int - center
        r = self.base.sample_float()
        if r < math.exp(-(c_frac * c_frac) / (2 * std_dev * std_dev)) / (math.sqrt(2 * math.pi) * std_dev):
            return int(c_int)
        return self._sample_round(c_frac, std_dev) + int(c_int)


# ---------------------------------------------------------------------------------------------
# [lattigo] ring: primes, sub-rings, NTT, Montgomery forms (third-party; conventions restated)
# ---------------------------------------------------------------------------------------------

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


class SubRing:
    def __init__(self, N, q):
        self.N, self.q = N, q
        self.logN = N.bit_length() - 1
        g = primitive_root(q)
        self.psi = pow(g, (q - 1) // (2 * N), q)            # [lattigo] psi = g^((q-1)/NthRoot)
        self.psi_inv = pow(self.psi, -1, q)
        # RootsForward[brv(j)] = psi^j  (j < N)
        self.roots = [0] * N
        self.roots_inv = [0] * N
        p = pi = 1
        for j in range(N):
            self.roots[brv(j, self.logN)] = p
            self.roots_inv[brv(j, self.logN)] = pi
            p = p * self.psi % q
            pi = pi * self.psi_inv % q
        self.n_inv = pow(N, -1, q)
        self.r = (1 << 64) % q
        self.r_inv = pow(1 << 64, -1, q)

    def ntt(self, a):
        """[lattigo] NTT: Cooley-Tukey, out[i] = a(psi^(2 brv(i)+1)), canonical."""
        a = list(a); q = self.q; N = self.N
        t = N; m = 1
        while m < N:
            t >>= 1
            for i in range(m):
                w = self.roots[m + i]
                j1 = 2 * i * t
                for j in range(j1, j1 + t):
                    v = a[j + t] * w % q
                    u = a[j]
                    a[j] = (u + v) % q
                    a[j + t] = (u - v) % q
            m <<= 1
        return a

    def intt(self, a):
        """[lattigo] INTT: Gentleman-Sande, includes N^-1."""
        a = list(a); q = self.q; N = self.N
        t = 1; m = N
        while m > 1:
            h = m >> 1
            j1 = 0
            for i in range(h):
                w = self.roots_inv[h + i]
                for j in range(j1, j1 + t):
                    u = a[j]; v = a[j + t]
                    a[j] = (u + v) % q
                    a[j + t] = (u - v) * w % q
                j1 += 2 * t
            t <<= 1
            m = h
        return [x * self.n_inv % q for x in a]


class Ring:
    """[lattigo] ring.Ring over the given moduli; polys are [limb][coeff] lists of ints."""

    def __init__(self, N, moduli):
        self.N = N
        self.moduli = list(moduli)
        self.sub = [SubRing(N, q) for q in moduli]
        self.Q = math.prod(moduli)

    def level(self): return len(self.moduli) - 1
    def new_poly(self): return [[0] * self.N for _ in self.moduli]
    def ntt(self, p): return [s.ntt(c) for s, c in zip(self.sub, p)]
    def intt(self, p): return [s.intt(c) for s, c in zip(self.sub, p)]
    def mform(self, p): return [[x * s.r % s.q for x in c] for s, c in zip(self.sub, p)]
    def imform(self, p): return [[x * s.r_inv % s.q for x in c] for s, c in zip(self.sub, p)]
    def add(self, a, b): return [[(x + y) % s.q for x, y in zip(ca, cb)] for s, ca, cb in zip(self.sub, a, b)]
    def sub_(self, a, b): return [[(x - y) % s.q for x, y in zip(ca, cb)] for s, ca, cb in zip(self.sub, a, b)]

    def mul_mont_then_add(self, a, b, c):
        """MulCoeffsMontgomeryThenAdd: c += a*b*2^-64 (returns the new c)."""
        return [[(z + x * y * s.r_inv) % s.q for x, y, z in zip(ca, cb, cc)] for s, ca, cb, cc in zip(self.sub, a, b, c)]

    def mul_mont_then_sub(self, a, b, c):
        return [[(z - x * y * s.r_inv) % s.q for x, y, z in zip(ca, cb, cc)] for s, ca, cb, cc in zip(self.sub, a, b, c)]

    def mul_scalar_then_sub(self, p1, scalar, p2):
        """MulScalarThenSub: p2 = p2 - p1*scalar."""
        return [[(y - x * scalar) % s.q for x, y in zip(c1, c2)] for s, c1, c2 in zip(self.sub, p1, p2)]


def poly_write_to(p):
    """[lattigo] Poly.WriteTo: u64 LE #limbs, then per limb u64 LE length + u64 LE coefficients."""
    out = [struct.pack("<Q", len(p))]
    for c in p:
        out.append(struct.pack("<Q", len(c)))
        out.append(struct.pack("<%dQ" % len(c), *c))
    return b"".join(out)


# ---------------------------------------------------------------------------------------------
# jindo/params.go
# ---------------------------------------------------------------------------------------------
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
    params.ringQ = Ring(params.d, params.q_primes)
    params.ringQOut = Ring(params.d, params.qout_primes)
    return params


def challenge_bound(p):
    return min(p.base, 1 << (120 // p.exp)) // 2           # params.go:358-361


# ---------------------------------------------------------------------------------------------
# jindo/rns.go, utils.go, entities.go
# ---------------------------------------------------------------------------------------------

class RNSReconstructor:
    def __init__(self, ring):
        self.ring = ring
        self.q_half = ring.Q >> 1
        self.gad = []
        for q in ring.moduli:
            qdiv = ring.Q // q
            self.gad.append(qdiv * pow(qdiv, -1, q) % ring.Q)      # rns.go:33-55

    @staticmethod
    def to_balanced(x, q):
        return x - q if x > (q >> 1) else x                  # rns.go:68-73

    def reconstruct(self, p):
        """rns.go:76-105 (fast path when all residues agree as small signed ints)."""
        out = []
        mods = self.ring.moduli
        for i in range(self.ring.N):
            c = self.to_balanced(p[0][i], mods[0])
            if all(self.to_balanced(p[j][i], mods[j]) == c for j in range(1, len(mods))):
                out.append(c)
                continue
            acc = sum(p[j][i] * self.gad[j] for j in range(len(mods))) % self.ring.Q
            if acc >= self.q_half:
                acc -= self.ring.Q
            out.append(acc)
        return out

    def set_big_coeff(self, v):
        return [[x % q for x in v] for q in self.ring.moduli]   # rns.go:108-114 (Euclidean mod)


def div_mod64(x, y):
    """utils.go:12-18 on a little-endian list of u64 (in place); returns the remainder."""
    r = 0
    for i in range(len(x) - 1, -1, -1):
        cur = (r << 64) | x[i]
        x[i] = cur // y
        r = cur % y
    return r


def set_coeff_signed(ring, c):
    """utils.go:49-60 per limb (note: c%q + q == q when c < 0 and c % q == 0 in Go's truncated %)."""
    if c >= 0:
        return [c for _ in ring.moduli]
    return [int(math.fmod(c, q)) + q for q in ring.moduli]


def encode_challenge(params, ring, chal_bytes):
    """utils.go:21-46"""
    p = ring.new_poly()
    c = [int.from_bytes(chal_bytes[:8], "big"), int.from_bytes(chal_bytes[8:16], "big")]
    bound = challenge_bound(params)
    for i in range(params.exp):
        r = div_mod64(c, bound)
        for k, q in enumerate(ring.moduli):
            p[k][i * params.slots] = (q - (bound - r)) if r > bound // 2 else r
    return ring.ntt(ring.mform(p))


class CommitKey:
    def __init__(self, params, crs):
        """entities.go:21-61: poly -> coefficient k -> limb l order of SampleN draws."""
        u = UniformSampler(crs)
        self.crs = bytes(crs)

        def poly(ring):
            p = ring.new_poly()
            for k in range(ring.N):
                for l, q in enumerate(ring.moduli):
                    p[l][k] = u.sample_n(q)
            return p
        self.In = [[poly(params.ringQ) for _ in range(params.rows)] for _ in range(params.in_msis)]
        self.MLWE = [[poly(params.ringQ) for _ in range(params.mlwe)] for _ in range(params.in_msis)]
        self.Out = [[poly(params.ringQOut) for _ in range(params.in_com_dcmp_len)] for _ in range(params.out_msis)]


class Opening:
    def __init__(self, params):
        self.InCommit = [params.ringQOut.new_poly() for _ in range(params.in_com_dcmp_len)]
        self.Encode = [[params.ringQ.new_poly() for _ in range(params.rows)] for _ in range(params.cols + 1)]
        self.MLWE = [[params.ringQ.new_poly() for _ in range(params.mlwe + params.in_msis)] for _ in range(params.cols + 1)]


def commitment_write_raw(params, value):
    """Commitment.WriteRawTo (entities.go:96-100).  Value polys are allocated in ringQ but hold
    ringQOut data (entities.go:85-94): trailing limbs stay zero and are serialized."""
    nl = len(params.ringQ.moduli)
    out = []
    for v in value:
        limbs = [list(c) for c in v] + [[0] * params.d for _ in range(nl - len(v))]
        out.append(poly_write_to(limbs[:max(nl, len(v))]))
    return b"".join(out)


# ---------------------------------------------------------------------------------------------
# field helpers (bignum.Uint semantics on Python ints: canonical values)
# ---------------------------------------------------------------------------------------------

def field_set_random(tape, q):
    """element.go:303-346 SetRandom: k bytes from the reader, top bits cleared, retry while >= q.
    The accepted raw integer is kept *as the Montgomery representation* (no toMont)."""
    bits = q.bit_length()
    k = (bits + 7) // 8
    b = bits % 8 or 8
    L = (bits + 63) // 64
    while True:
        buf = bytearray(tape.read(k))
        buf[k - 1] &= (1 << b) - 1
        raw = int.from_bytes(buf, "little")
        if raw < q:
            R = 1 << (64 * L)
            return raw * pow(R, -1, q) % q       # canonical value whose Montgomery limbs are `raw`


# ---------------------------------------------------------------------------------------------
# jindo/encoder.go
# ---------------------------------------------------------------------------------------------

class Encoder:
    def __init__(self, params, tape):
        self.p = params
        p_big = params.base ** params.exp + 1
        thr = math.pow(2.0, -50) / (float(params.base) * float(params.exp))
        self.delta_inv = []
        for i in range(params.exp):
            # encoder.go:50-67: -(b^i)/p as big.Float of precision bitlen(p), rounded to float64.
            # (exact rational -> nearest double; the 2^-255 relative error of the reference's
            #  iterated big.Float product cannot move a float64 rounding)
            v = float(Fraction(-(params.base ** i), p_big))
            if abs(v) < thr:
                v = 0.0
            self.delta_inv.append(v)
        # sampler construction order: encoder.go:72-74
        self.twin = TwinCDTGaussianSampler(params.ecd_std, tape.read(32))
        self.cosac = COSACSampler(tape.read(32), tape.read(32))
        self.rounded = RoundedGaussianSampler(tape.read(32))
        self.rns = RNSReconstructor(params.ringQ)

    def base_encode(self, v):
        """encoder.go:120-146: v = canonical ints, len(v) <= slots."""
        p = self.p
        if len(v) > p.slots:
            raise ValueError("len(v) > slots")
        L = (p.q_field.bit_length() + 63) // 64
        coeffs = [0] * p.d
        for i, x in enumerate(v):
            limbs = [(x >> (64 * k)) & MASK64 for k in range(L)]
            for j in range(p.exp - 1):
                coeffs[j * p.slots + i] = div_mod64(limbs, p.base)
            coeffs[(p.exp - 1) * p.slots + i] = limbs[0]
        return [list(coeffs) for _ in p.ringQ.moduli]

    def encode(self, v):
        """encoder.go:113-117 (MFormLazy then NTT: the NTT output is canonical)."""
        return self.p.ringQ.ntt(self.p.ringQ.mform(self.base_encode(v)))

    def centre(self, digits0):
        """encoder.go:152-164 fpSample (float64, unfused multiply-add in the reference's order)."""
        p = self.p; N = p.d
        fp = [0.0] * N
        for i in range(p.exp):
            di = self.delta_inv[i]
            if di == 0:
                continue
            dd = N - (i + 1) * p.slots
            for j in range(0, N - dd):
                fp[j + dd] += di * float(digits0[j])
            for jj, j in enumerate(range(N - dd, N)):
                fp[jj] -= di * float(digits0[j])
        return fp

    def rand_encode(self, v, std_dev, g=None, rec=None):
        """encoder.go:149-201.  Returns (poly, g) where g is the sampled integer vector."""
        p = self.p; ring = p.ringQ; N = p.d
        out = self.base_encode(v)
        self.last_u = None
        if g is None:
            fp = self.centre(out[0])
            g = []
            if std_dev == self.twin.std_dev:
                self.last_u = []
            for i in range(N):
                if std_dev == self.twin.std_dev:
                    u = self.twin.base.sample()          # exactly one draw per sample (gaussian_twin_cdt.go:84)
                    self.last_u.append(u)
                    g.append(self.twin.sample_with_u(-fp[i], u))
                else:
                    g.append(self.cosac.sample(-fp[i], std_dev))
        ps = [[(c if c >= 0 else int(math.fmod(c, q)) + q) for c in g] for q in ring.moduli]
        ps = ring.mform(ps)
        shift = ring.new_poly()
        for k, q in enumerate(ring.moduli):
            for i in range(N - p.slots):
                shift[k][i + p.slots] = ps[k][i]
            for ii, i in enumerate(range(N - p.slots, N)):
                shift[k][ii] = q - ps[k][i]                  # may equal q when ps == 0 (as in the reference)
        shift = ring.mul_scalar_then_sub(ps, p.base, shift)
        out = ring.add(ring.mform(out), shift)
        if rec is not None:
            if self.last_u is not None:
                rec["twin_u"].extend(self.last_u)        # TwinCDT rows: the raw uniform draws, call order
            else:
                rec["cosac_g"].extend(g)                 # COSAC rows: the sampled integers, call order
        return ring.ntt(out), g

    def decode(self, poly, count=None):
        """encoder.go:204-219 DecodeTo (returns canonical field values for all slots)."""
        p = self.p
        big = self.rns.reconstruct(poly)
        res = []
        for i in range(p.slots):
            acc = 0
            for j in range(p.exp - 1, -1, -1):
                acc = (acc * p.base + big[j * p.slots + i]) % p.q_field
            res.append(acc)
        return res


# ---------------------------------------------------------------------------------------------
# jindo/prover.go
# ---------------------------------------------------------------------------------------------

def left_vec(params, x):
    q = params.q_field
    skip = pow(x, params.cols * params.slots, q)
    left = [1]
    for _ in range(1, params.rows):
        left.append(left[-1] * skip % q)
    left[params.rows - 1] = x                                   # utils.go:71
    return left


def right_vec(params, x):
    q = params.q_field
    r = [1]
    for _ in range(1, params.cols * params.slots):
        r.append(r[-1] * x % q)
    return r


class Prover:
    def __init__(self, params, crs, tape, commit_key=None):
        """NewProver (prover.go:28-40): encoder samplers first, then uniform/rounded/mlwe seeds."""
        self.p = params
        self.ecd = Encoder(params, tape)
        self.rns_out = RNSReconstructor(params.ringQOut)
        self.ck = commit_key or CommitKey(params, crs)
        self.tape = tape
        self.uniform_seed = tape.read(32)                        # uniformSampler (unused by Commit/Evaluate)
        self.rounded = RoundedGaussianSampler(tape.read(32))
        self.mlwe = TwinCDTGaussianSampler(params.mlwe_std, tape.read(32))
        self.trace = None      # optional dict filled by commit() with every sampled quantity

    # -- prover.go:65-86 ---------------------------------------------------------------------
    def gen_first_last_row(self, v):
        p = self.p; q = p.q_field
        n = p.cols * p.slots
        last = [field_set_random(self.tape, q) for _ in range(n - 1)] + [0]
        first = [v[0]]
        for i in range(1, n):
            first.append(((v[i] if i < len(v) else 0) - last[i - 1]) % q)
        return first, last

    def cut_pipeline(self, ring_in, rns_in, poly, log_cut):
        """prover.go:164-176 / :190-200: IMForm, INTT, CRT, >> cut (floor), re-RNS, MForm, NTT."""
        ro = self.p.ringQOut
        c = ring_in.intt(ring_in.imform(poly))
        big = rns_in.reconstruct(c)
        big = [x >> log_cut for x in big]                        # Python >> on negatives floors, like big.Int.Rsh
        return ro.ntt(ro.mform(self.rns_out.set_big_coeff(big)))

    # -- prover.go:89-177 --------------------------------------------------------------------
    def commit_col(self, i, opening, v, first, last, rec):
        p = self.p; ring = p.ringQ; q = p.q_field
        rs, re = i * p.slots, (i + 1) * p.slots
        if i == p.cols:
            mask = [field_set_random(self.tape, q) for _ in range(p.slots)]
            rec["mask"].append(list(mask))
            opening.Encode[i][0], g = self.ecd.rand_encode(mask, p.mask_blind_std, rec=rec); rec["g"][(i, 0)] = g
            for j in range(1, p.rows - 1):
                if j * p.cols * p.slots > len(v):
                    break
                mask = [field_set_random(self.tape, q) for _ in range(p.slots)]
                rec["mask"].append(list(mask))
                opening.Encode[i][j], g = self.ecd.rand_encode(mask, p.mask_std, rec=rec); rec["g"][(i, j)] = g
            mask = [field_set_random(self.tape, q) for _ in range(p.slots)]
            rec["mask"].append(list(mask))
            opening.Encode[i][p.rows - 1], g = self.ecd.rand_encode(mask, p.mask_std, rec=rec); rec["g"][(i, p.rows - 1)] = g
        else:
            opening.Encode[i][0], g = self.ecd.rand_encode(first[rs:re], p.ecd_blind_std, rec=rec); rec["g"][(i, 0)] = g
            for j in range(1, p.rows - 1):
                s = j * p.cols * p.slots + rs
                e = j * p.cols * p.slots + re
                if s > len(v):
                    break
                opening.Encode[i][j], g = self.ecd.rand_encode(v[s:min(e, len(v))], p.ecd_std, rec=rec); rec["g"][(i, j)] = g
            opening.Encode[i][p.rows - 1], g = self.ecd.rand_encode(last[rs:re], p.ecd_std, rec=rec); rec["g"][(i, p.rows - 1)] = g
        for j in range(p.in_msis + p.mlwe):
            if i == p.cols:
                samples = [self.rounded.sample(0, p.mask_mlwe_std) for _ in range(p.d)]
            else:
                samples = [self.mlwe.sample(0) for _ in range(p.d)]
            rec["mlwe"][(i, j)] = samples
            poly = [[(c if c >= 0 else int(math.fmod(c, qq)) + qq) for c in samples] for qq in ring.moduli]
            opening.MLWE[i][j] = ring.ntt(ring.mform(poly))
        for j in range(p.in_msis):
            com = ring.new_poly()
            for k in range(p.rows):
                com = ring.mul_mont_then_add(self.ck.In[j][k], opening.Encode[i][k], com)
            for k in range(p.mlwe):
                com = ring.mul_mont_then_add(self.ck.MLWE[j][k], opening.MLWE[i][k], com)
            com = ring.add(opening.MLWE[i][p.mlwe + j], com)
            opening.InCommit[i * p.in_msis + j] = self.cut_pipeline(ring, self.ecd.rns, com, p.log_in_cut)

    # -- prover.go:45-62, :180-202 -----------------------------------------------------------
    def commit(self, v):
        p = self.p
        if len(v) > p.rank:
            raise ValueError("len(v) > params.rank")
        opening = Opening(p)
        rec = {"g": {}, "mlwe": {}, "mask": [], "twin_u": [], "cosac_g": []}
        self._rec = rec
        first, last = self.gen_first_last_row(v)
        rec["first"], rec["last"] = first, last
        for i in range(p.cols + 1):
            self.commit_col(i, opening, v, first, last, rec)
        ro = p.ringQOut
        value = []
        for i in range(p.out_msis):
            acc = ro.new_poly()
            for j in range(p.in_com_dcmp_len):
                acc = ro.mul_mont_then_add(self.ck.Out[i][j], opening.InCommit[j], acc)
            value.append(self.cut_pipeline(ro, self.rns_out, acc, p.log_out_cut))
        self.trace = rec
        return value, opening

    # -- prover.go:205-324 -------------------------------------------------------------------
    def evaluate(self, x, vs, coms, opens):
        p = self.p; ring = p.ringQ; ro = p.ringQOut; q = p.q_field
        L = (q.bit_length() + 63) // 64
        if not (len(vs) == len(coms) == len(opens) == p.batch):
            raise ValueError("len(v) != params.batch")
        x_bytes = x.to_bytes(8 * L, "big")                      # Marshal: big-endian canonical (element.go:873-883)
        prefix = self.ck.crs + b"".join(commitment_write_raw(p, c) for c in coms) + x_bytes
        if p.batch > 1:
            batch_bytes = hashlib.shake_128(prefix).digest(16 * p.batch)
            bq = [encode_challenge(p, ring, batch_bytes[16 * i:16 * i + 16]) for i in range(p.batch)]
            bo = [encode_challenge(p, ro, batch_bytes[16 * i:16 * i + 16]) for i in range(p.batch)]
            prefix = prefix + batch_bytes
            ob = Opening(p)
            for i in range(p.batch):
                for j in range(len(ob.InCommit)):
                    ob.InCommit[j] = ro.mul_mont_then_add(opens[i].InCommit[j], bo[i], ob.InCommit[j])
                for j in range(p.cols + 1):
                    for k in range(p.rows):
                        ob.Encode[j][k] = ring.mul_mont_then_add(opens[i].Encode[j][k], bq[i], ob.Encode[j][k])
                    for k in range(p.mlwe + p.in_msis):
                        ob.MLWE[j][k] = ring.mul_mont_then_add(opens[i].MLWE[j][k], bq[i], ob.MLWE[j][k])
        else:
            ob = opens[0]
        pf = {"InCommit": [list(map(list, c)) for c in ob.InCommit]}
        left = [self.ecd.encode([le]) for le in left_vec(p, x)]
        pf["Partial"] = []
        for i in range(p.cols):
            acc = ring.new_poly()
            for j in range(p.rows):
                acc = ring.mul_mont_then_add(left[j], ob.Encode[i][j], acc)
            pf["Partial"].append(acc)
        acc = ring.new_poly()
        for j in range(p.rows):
            acc = ring.mul_mont_then_add(left[j], ob.Encode[p.cols][j], acc)
        pf["PartialMask"] = acc
        tr = prefix + b"".join(poly_write_to(c) for c in pf["Partial"]) + poly_write_to(pf["PartialMask"])
        chal_stream = hashlib.shake_128(tr).digest(16 * p.cols)
        chals = [encode_challenge(p, ring, chal_stream[16 * i:16 * i + 16]) for i in range(p.cols)]
        pf["Encode"] = []
        for i in range(p.rows):
            acc = [list(c) for c in ob.Encode[p.cols][i]]
            for j in range(p.cols):
                acc = ring.mul_mont_then_add(chals[j], ob.Encode[j][i], acc)
            pf["Encode"].append(acc)
        pf["MLWE"] = []
        for i in range(p.mlwe + p.in_msis):
            acc = [list(c) for c in ob.MLWE[p.cols][i]]
            for j in range(p.cols):
                acc = ring.mul_mont_then_add(chals[j], ob.MLWE[j][i], acc)
            pf["MLWE"].append(acc)
        evals = []
        for v in vs:
            z = 0
            for c in reversed(v):
                z = (z * x + c) % q
            evals.append(z)
        pf["_chals"] = chals
        return evals, pf


# ---------------------------------------------------------------------------------------------
# jindo/verifier.go (oracle self-check: restated Verify accepts restated proofs)
# ---------------------------------------------------------------------------------------------

class Verifier:
    def __init__(self, params, crs, tape, commit_key=None):
        self.p = params
        self.ecd = Encoder(params, tape)                         # verifier.go:38 (128 B of the tape)
        self.rns_out = RNSReconstructor(params.ringQOut)
        self.ck = commit_key or CommitKey(params, crs)

    def _norm_ok(self, rns, polys, nm):
        sq = 0
        for poly in polys:
            for c in rns.reconstruct(poly):
                sq += c * c
        return float(math.isqrt(sq)) < nm                        # verifier.go:262-282

    def _mod_up_qout_to_q(self, poly):
        """[lattigo] BasisExtender.ModUpQtoP: centred lift from ringQOut to ringQ."""
        big = self.rns_out.reconstruct(poly)
        return [[x % q for x in big] for q in self.p.ringQ.moduli]

    def verify(self, x, coms, ys, pf):
        p = self.p; ring = p.ringQ; ro = p.ringQOut; q = p.q_field
        L = (q.bit_length() + 63) // 64
        x_bytes = x.to_bytes(8 * L, "big")
        prefix = self.ck.crs + b"".join(commitment_write_raw(p, c) for c in coms) + x_bytes
        bq = bo = None
        if p.batch > 1:
            batch_bytes = hashlib.shake_128(prefix).digest(16 * p.batch)
            bq = [encode_challenge(p, ring, batch_bytes[16 * i:16 * i + 16]) for i in range(p.batch)]
            bo = [encode_challenge(p, ro, batch_bytes[16 * i:16 * i + 16]) for i in range(p.batch)]
            prefix = prefix + batch_bytes
        tr = prefix + b"".join(poly_write_to(c) for c in pf["Partial"]) + poly_write_to(pf["PartialMask"])
        chal_stream = hashlib.shake_128(tr).digest(16 * p.cols)
        chals = [encode_challenge(p, ring, chal_stream[16 * i:16 * i + 16]) for i in range(p.cols)]
        inv = lambda r, c: r.intt(r.imform(c))
        pi_partial = [inv(ring, c) for c in pf["Partial"]]
        pi_encode = [inv(ring, c) for c in pf["Encode"]]
        pi_mlwe = [inv(ring, c) for c in pf["MLWE"]]
        pi_incom = [inv(ro, c) for c in pf["InCommit"]]
        # outer commitment (verifier.go:135-163)
        out_cut = (1 << p.log_out_cut)
        cut = []
        for i in range(p.out_msis):
            if p.batch > 1:
                acc = ro.new_poly()
                for j in range(p.batch):
                    acc = ro.mul_mont_then_add(coms[j][i], bo[j], acc)
            else:
                acc = [list(c) for c in coms[0][i]]
            acc = [[v * out_cut % s.q for v in c] for s, c in zip(ro.sub, acc)]      # MulRNSScalarMontgomery by MForm(2^cut)
            for j in range(p.in_com_dcmp_len):
                acc = ro.mul_mont_then_sub(self.ck.Out[i][j], pf["InCommit"][j], acc)
            cut.append(inv(ro, acc))
        if not self._norm_ok(self.rns_out, pi_incom + cut, p.in_com_dcmp_two_nm):
            return False
        # inner commitment (verifier.go:165-205)
        in_cut = (1 << p.log_in_cut)
        cut = []
        for i in range(p.in_msis):
            acc = ring.new_poly()
            for j in range(p.cols + 1):
                up = ring.ntt(ring.mform(self._mod_up_qout_to_q(pi_incom[j * p.in_msis + i])))
                acc = ring.add(up, acc) if j == p.cols else ring.mul_mont_then_add(up, chals[j], acc)
            acc = [[v * in_cut % s.q for v in c] for s, c in zip(ring.sub, acc)]
            for j in range(p.rows):
                acc = ring.mul_mont_then_sub(self.ck.In[i][j], pf["Encode"][j], acc)
            for j in range(p.mlwe):
                acc = ring.mul_mont_then_sub(self.ck.MLWE[i][j], pf["MLWE"][j], acc)
            acc = ring.sub_(acc, pf["MLWE"][p.mlwe + i])
            cut.append(inv(ring, acc))
        if not self._norm_ok(self.ecd.rns, pi_encode + pi_mlwe + cut, p.res_two_nm):
            return False
        # consistency (verifier.go:207-227)
        test = ring.new_poly()
        left = left_vec(p, x)
        for i in range(p.rows):
            test = ring.mul_mont_then_add(self.ecd.encode([left[i]]), pf["Encode"][i], test)
        for i in range(p.cols):
            test = ring.mul_mont_then_sub(chals[i], pf["Partial"][i], test)
        test = ring.sub_(test, pf["PartialMask"])
        if any(any(c) for c in test):
            return False
        # evaluation (verifier.go:229-260)
        right = right_vec(p, x)
        if p.batch > 1:
            yb = 0
            for i in range(p.batch):
                dcd = self.ecd.decode(inv(ring, bq[i]))
                yb = (yb + dcd[0] * ys[i]) % q
        else:
            yb = ys[0] % q
        t = 0
        for i in range(p.cols):
            dcd = self.ecd.decode(pi_partial[i])
            for j in range(p.slots):
                t = (t + right[i * p.slots + j] * dcd[j]) % q
        return t == yb
