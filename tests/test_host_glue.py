"""CPU tests of the product-side host helpers (no GPU, no compute through the library)."""
import numpy as np

from oracle import jindo_ref as J, pyref


def test_product_modulus_table_matches_reference_constants(product):
    """bigpoly.MODULI (product side) lists the same b^k + 1 moduli as the oracle's table, which is pinned to the
    reference's zp packages by tests/golden/field_constants.json."""
    assert product.MODULI == pyref.MODULI
    for name in pyref.MODULI:
        assert product.modulus(name) == pyref.modulus(name)


def test_commitment_serialisation_numpy_equals_lists(product):
    """jindo_host.commitment_write_raw: the numpy fast path writes the bytes of the list path, which restates
    Commitment.WriteRawTo (jindo/entities.go:85-100) -- including the zero limbs of polynomials allocated in ringQ."""
    JH = product.jindo_host
    rng = np.random.default_rng(3)
    for name, n, batch in (("zp255", 1024, 1), ("zp220", 2049, 9), ("zp128", 700, 2)):
        p = JH.new_parameters(pyref.modulus(name), n, batch)
        com = np.stack([np.stack([rng.integers(0, pr, size=p.d, dtype=np.uint64) for pr in p.qout_primes]) for _ in range(p.out_msis)])
        a = JH.commitment_write_raw(p, com)
        assert a == JH.commitment_write_raw(p, com.tolist())
        po = J.new_parameters(pyref.modulus(name), n, batch)
        assert a == J.commitment_write_raw(po, com.tolist())


def test_host_parameters_match_oracle(product):
    """jindo_host.new_parameters (product side: a thin wrapper over the compiled rs_jindo_new_parameters, csrc/jindo_params.cpp)
    against the oracle's Python restatement of jindo.NewParameters for the shapes of SURVEY section 8; tests/test_host_params.py
    compares every field, float and table of the C ABI result itself."""
    JH = product.jindo_host
    for name, n, batch in (("zp255", 1 << 22, 1), ("zp128", 16385, 12), ("zp240", 32769, 43), ("zp220", 2049, 9), ("zp255", 1024, 8)):
        a = JH.new_parameters(pyref.modulus(name), n, batch); b = J.new_parameters(pyref.modulus(name), n, batch)
        for f in ("rows", "cols", "slots", "in_msis", "mlwe", "out_msis", "log_in_cut", "log_out_cut", "d", "exp", "base"):
            assert getattr(a, f) == getattr(b, f), (name, f)
        assert list(a.q_primes) == list(b.q_primes) and list(a.qout_primes) == list(b.qout_primes)


def test_twin_cdt_slow_path_lower_bound_is_a_lower_bound(product):
    """The shortcut of encode_kernel's TwinCDT slow path (csrc/jindo.cu, rs_jindo_create): lb[k][v0] must not exceed
    sum_{x = tail_lo .. v0} rho(x - c) / norm for ANY c in [k/16, (k+1)/16) -- the quantity the reference compares u / 2^64
    with (math/csprng/gaussian_twin_cdt.go:95-110).  Restated here term by term and checked on a fine grid of c,
    interval end points included, for the sigma / tail of the parameter sets the tests and the bench use."""
    import math
    JH = product.jindo_host
    for name, n, batch in (("zp255", 1 << 22, 1), ("zp128", 700, 2), ("zp220", 2049, 9)):
        p = JH.new_parameters(pyref.modulus(name), n, batch)
        pd, tables = JH.prover_inputs(p)
        sg, tail_lo, cdt_len = pd["cdt_sigma"], pd["cdt_tail_lo"], len(tables[0])
        den, norm = 2.0 * sg * sg, math.sqrt(2.0 * math.pi) * sg
        for k in range(16):
            a, b = k / 16.0, (k + 1) / 16.0
            neg = sum(math.exp(-((x - b) ** 2) / den) for x in range(tail_lo, 1))
            pos, lb = 0.0, []
            for v0 in range(cdt_len):
                if v0 >= 1:
                    pos += math.exp(-((v0 - a) ** 2) / den)
                lb.append((neg + pos) / norm - 1e-9)
            for c in [a, math.nextafter(b, 0.0)] + [a + (b - a) * t / 7.0 for t in range(1, 7)]:
                run = 0.0
                x = tail_lo
                for v0 in range(cdt_len):
                    while x <= v0:
                        run += math.exp(-((x - c) ** 2) / den)
                        x += 1
                    assert lb[v0] <= run / norm, (name, k, v0, c)
            assert lb[-1] > 0.9                      # and it is tight enough to settle almost every slow sample


def test_mult_circuit_description_matches_the_oracle_compile(product):
    """ringo-snark_b200/circuits.py carries the compiled form of examples/mult's MultCircuit as plain data (what a Go caller
    gets from buckler.Compile); it must be what the restated Compile produces."""
    import oracle
    from oracle import buckler_ref as B, pyref
    q = pyref.modulus("zp128")
    for N in (256, 8192):
        define, n_pw, n_w = B.mult_circuit(N)
        ctx = B.Context(q, N, n_pw, n_w)
        define(ctx, oracle.Field(q))
        terms, coeffs = ctx.terms(ctx.arith_constraints)
        d, commit_rank, batch = product.circuits.mult_circuit(q, N)
        assert d["arith_terms"] == terms and d["arith_coeffs"] == coeffs
        assert (d["w_cnt"], d["pw_cnt"], d["arith_max_rank"]) == (ctx.w_cnt, ctx.pw_cnt, ctx.arith_max_rank)
        assert commit_rank == ctx.commit_rank() and batch == ctx.batch()
        assert d["emb_rank"] == 1 << (max(ctx.arith_max_rank, ctx.sumcheck_max_rank) - 1).bit_length()
        assert [(k, [tuple(p) for p in pairs]) for k, pairs in d["lin"]] == \
            [("ntt", [tuple(p) for p in ctx.lin_constraints[id(chk)]]) for chk in ctx.lin_checkers]
