"""oracle/pyref.py -- TEST INFRASTRUCTURE.  Python-integer statements of the
field / bigpoly contracts, used to pin the C oracle (small sizes only).

The reference's own gopter tests compare every field op against math/big mod q
(e.g. examples/mult/zp/element_test.go:946-986); Python ints play that role.
"""

MODULI = {  # q = b^k + 1  (values from the reference's zp packages)
    "zp110": (12640, 8),
    "zp128": (60256, 8),
    "zp220": (13216, 16),
    "zp240": (31432, 16),
    "zp255": (60272, 16),
    "zp440": (13512, 32),
    "zp880": (13694, 64),
}


def modulus(name):
    b, k = MODULI[name]
    return b ** k + 1


def brv(i, bits):
    r = 0
    for _ in range(bits):
        r = (r << 1) | (i & 1)
        i >>= 1
    return r


def find_generator(q, order):
    """math/bigpoly/ntt.go:40-53 / :167-180: smallest x>=2 whose (q-1)/order power has exact order `order`."""
    e = (q - 1) // order
    x = 2
    while True:
        g = pow(x, e, q)
        if pow(g, order // 2, q) != 1:
            return g, x
        x += 1


def ntt_naive(a, q, negacyclic):
    """Forward transform contract: bit-reversed-order evaluations."""
    N = len(a)
    bits = N.bit_length() - 1
    if negacyclic:
        g, _ = find_generator(q, 2 * N)
        pts = [pow(g, 2 * brv(j, bits) + 1, q) for j in range(N)]
    else:
        g, _ = find_generator(q, N)
        pts = [pow(g, brv(j, bits), q) for j in range(N)]
    out = []
    for x in pts:
        acc = 0
        for c in reversed(a):
            acc = (acc * x + c) % q
        out.append(acc)
    return out


def ntt_ref(a, q, negacyclic):
    """Literal restatement of nttInPlaceRef (ntt.go:261-275) on Python ints."""
    N = len(a)
    bits = N.bit_length() - 1
    p = list(a)
    if negacyclic:
        g, _ = find_generator(q, 2 * N)
        tw = [pow(g, brv(i, bits), q) for i in range(N)]
    else:
        g, _ = find_generator(q, N)
        half = [pow(g, brv(i, bits - 1), q) for i in range(N // 2)]
        tw = [0] * N
        m = 1
        while m <= N // 2:
            for i in range(m):
                tw[m + i] = half[i]
            m <<= 1
    t = N
    m = 1
    while m <= N // 2:
        t >>= 1
        for i in range(m):
            j1 = (i * t) << 1
            for j in range(j1, j1 + t):
                v = p[j + t] * tw[m + i] % q
                u = p[j]
                p[j] = (u + v) % q
                p[j + t] = (u - v) % q
        m <<= 1
    return p


def quorem_vanishing(p, N, q):
    """cyclic.go:18-37"""
    rank = len(p)
    quo = [0] * rank
    rem = list(p)
    for i in range(rank - 1, N - 1, -1):
        quo[i - N] = (quo[i - N] + rem[i]) % q
        rem[i - N] = (rem[i - N] + rem[i]) % q
        rem[i] = 0
    return quo, rem
