// pointwise.cu -- element-wise field kernels and the fused Buckler constraint evaluator.
//
// Replaces math/bigpoly/vec.go:9-121 (addVecTo ... mulVecTo), base_op.go:49-171
// (Add/Sub/Neg/ScalarMul/Mul/MulAdd/MulSub/ScalarMulAdd/ScalarMulSub To),
// cyclic.go:18-37 (QuoRemByVanishing), cyclotomic.go:49-82 (autTo/autNTTTo),
// buckler/prover.go:355-379 (evalCircuit), buckler/linear.go:38-41 and poly.go:64-76.
// All of these stream each coefficient once: one thread per element, 16-byte vector
// accesses, grid sized in whole waves of the 148 SMs.
#include "common.cuh"
#include "shoup.cuh"

namespace rs {

enum PwOp { OP_ADD, OP_SUB, OP_NEG, OP_MUL, OP_MULADD, OP_MULSUB, OP_SMUL, OP_SMULADD, OP_SMULSUB };

template <int NL>
__device__ __forceinline__ void ld_elem(uint32_t *x, const uint32_t *p) {
    if (NL % 4 == 0) {
#pragma unroll
        for (int h = 0; h < NL / 4; h++) {
            uint4 v = reinterpret_cast<const uint4 *>(p)[h];
            x[4 * h] = v.x; x[4 * h + 1] = v.y; x[4 * h + 2] = v.z; x[4 * h + 3] = v.w;
        }
    } else {
#pragma unroll
        for (int h = 0; h < NL / 2; h++) {
            uint2 v = reinterpret_cast<const uint2 *>(p)[h];
            x[2 * h] = v.x; x[2 * h + 1] = v.y;
        }
    }
}

template <int NL>
__device__ __forceinline__ void st_elem(uint32_t *p, const uint32_t *x) {
    if (NL % 4 == 0) {
#pragma unroll
        for (int h = 0; h < NL / 4; h++)
            reinterpret_cast<uint4 *>(p)[h] = make_uint4(x[4 * h], x[4 * h + 1], x[4 * h + 2], x[4 * h + 3]);
    } else {
#pragma unroll
        for (int h = 0; h < NL / 2; h++) reinterpret_cast<uint2 *>(p)[h] = make_uint2(x[2 * h], x[2 * h + 1]);
    }
}

template <int NL> struct Scalar { uint32_t v[NL]; };

template <int NL, int QS, int OP>
__global__ void __launch_bounds__(256) pointwise_kernel(uint32_t *out, const uint32_t *a, const uint32_t *b,
                                                        const __grid_constant__ Scalar<NL> c,
                                                        const __grid_constant__ FieldParams<NL> F, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        uint32_t x[NL], y[NL], z[NL];
        ld_elem<NL>(x, a + i * NL);
        if (OP == OP_ADD || OP == OP_SUB || OP == OP_MUL || OP == OP_MULADD || OP == OP_MULSUB) ld_elem<NL>(y, b + i * NL);
        if (OP == OP_ADD) mod_add<NL>(z, x, y, F);
        if (OP == OP_SUB) mod_sub<NL>(z, x, y, F);
        if (OP == OP_NEG) mod_neg<NL>(z, x, F);
        if (OP == OP_MUL || OP == OP_MULADD || OP == OP_MULSUB) mont_mul<NL, QS>(z, x, y, F);
        if (OP == OP_SMUL || OP == OP_SMULADD || OP == OP_SMULSUB) mont_mul<NL, QS>(z, x, c.v, F);
        if (OP == OP_MULADD || OP == OP_SMULADD) { ld_elem<NL>(x, out + i * NL); mod_add<NL>(z, x, z, F); }
        if (OP == OP_MULSUB || OP == OP_SMULSUB) { ld_elem<NL>(x, out + i * NL); mod_sub<NL>(z, x, z, F); }
        st_elem<NL>(out + i * NL, z);
    }
}

static inline unsigned grid_for(size_t n, int threads) {
    size_t blocks = (n + threads - 1) / threads;
    const size_t cap = 148 * 16;   // whole waves of 148 SMs; the kernels grid-stride beyond that
    if (blocks > cap) blocks = cap;
    return (unsigned)(blocks ? blocks : 1);
}

template <int NL, int QS>
static int pointwise_launch(int op, const rs_field *f, uint32_t *out, const uint32_t *a, const uint32_t *b,
                            const uint32_t *c, size_t n) {
    FieldParams<NL> F = make_params<NL>(f);
    Scalar<NL> sc;
    for (int i = 0; i < NL; i++) sc.v[i] = c ? c[i] : 0;
    const unsigned grid = grid_for(n, 256);
    cudaStream_t st = current_stream();
#define RS_PW(OPC) case OPC: pointwise_kernel<NL, QS, OPC><<<grid, 256, 0, st>>>(out, a, b, sc, F, n); break;
    switch (op) {
        RS_PW(OP_ADD) RS_PW(OP_SUB) RS_PW(OP_NEG) RS_PW(OP_MUL) RS_PW(OP_MULADD) RS_PW(OP_MULSUB)
        RS_PW(OP_SMUL) RS_PW(OP_SMULADD) RS_PW(OP_SMULSUB)
        default: return set_error(RS_ERR_ARG, "pointwise: bad op");
    }
#undef RS_PW
    count_launch();
    RS_CUDA_CHECK(cudaGetLastError());
    return RS_OK;
}

// dispatch over the compiled (limbs, modulus shape) combinations
#define RS_FIELD_DISPATCH(f, CALL)                                              \
    switch ((f)->NL) {                                                          \
        case 4: if ((f)->qs >= 1) { CALL(4, 1); } else { CALL(4, 0); } break;   \
        case 8: if ((f)->qs == 2) { CALL(8, 2); } else if ((f)->qs == 1) { CALL(8, 1); } else { CALL(8, 0); } break; \
        case 14: if ((f)->qs == 2) { CALL(14, 2); } else { CALL(14, 0); } break; \
        case 28: if ((f)->qs == 2) { CALL(28, 2); } else { CALL(28, 0); } break; \
        default: return set_error(RS_ERR_UNSUPPORTED, "unsupported limb count %d", (f)->L64); \
    }

int pointwise(int op, const rs_field *f, uint32_t *out, const uint32_t *a, const uint32_t *b, const uint32_t *c, size_t n) {
    if (n == 0) return RS_OK;
#define CALL(NL, QS) return pointwise_launch<NL, QS>(op, f, out, a, b, c, n)
    RS_FIELD_DISPATCH(f, CALL)
#undef CALL
    return RS_OK;
}

// ---------------------------------------------------------------------------
// evalCircuit (buckler/prover.go:355-379), fused: one pass over the witnesses
// ---------------------------------------------------------------------------
// The term list and the constraint coefficients are the same for every coefficient of the polynomial: each CTA stages
// them in shared memory once (a few KB), so the per-coefficient loop reads witnesses from HBM / L1 and nothing else.
template <int NL, int QS>
__global__ void __launch_bounds__(128) eval_circuit_kernel(uint32_t *out, const uint32_t *pw, const uint32_t *w,
                                                           const rs_term *terms, int n_terms, const uint32_t *coeffs, int n_coeffs,
                                                           const __grid_constant__ Scalar<NL> bc,
                                                           const __grid_constant__ FieldParams<NL> F, size_t rank, int staged) {
    extern __shared__ __align__(16) uint32_t ec_smem[];
    const rs_term *T = terms;
    const uint32_t *CO = coeffs;
    if (staged) {
        constexpr int TW = sizeof(rs_term) / 4;
        uint32_t *st = ec_smem, *sc = ec_smem + n_terms * TW;
        for (int i = threadIdx.x; i < n_terms * TW; i += blockDim.x) st[i] = reinterpret_cast<const uint32_t *>(terms)[i];
        for (int i = threadIdx.x; i < n_coeffs * NL; i += blockDim.x) sc[i] = coeffs[i];
        __syncthreads();
        T = reinterpret_cast<const rs_term *>(st);
        CO = sc;
    }
    for (size_t j = blockIdx.x * (size_t)blockDim.x + threadIdx.x; j < rank; j += (size_t)gridDim.x * blockDim.x) {
        uint32_t acc[NL], ev[NL];
#pragma unroll
        for (int l = 0; l < NL; l++) { acc[l] = 0; ev[l] = 0; }
        int cur = n_terms ? T[0].constraint : -1;
        for (int k = 0; k <= n_terms; k++) {
            if (k == n_terms || T[k].constraint != cur) {
                // eval *= batchConst ; pOut += eval   (prover.go:374-375)
                uint32_t t[NL];
                mont_mul<NL, QS>(t, ev, bc.v, F);
                mod_add<NL>(acc, acc, t, F);
#pragma unroll
                for (int l = 0; l < NL; l++) ev[l] = 0;
                if (k == n_terms) break;
                cur = T[k].constraint;
            }
            const rs_term &tm = T[k];
            uint32_t term[NL], fac[NL], t[NL];
#pragma unroll
            for (int l = 0; l < NL; l++) term[l] = CO[(size_t)tm.coeff * NL + l];
            if (tm.pw >= 0) {
                ld_elem<NL>(fac, pw + ((size_t)tm.pw * rank + j) * NL);
                mont_mul<NL, QS>(t, term, fac, F);
#pragma unroll
                for (int l = 0; l < NL; l++) term[l] = t[l];
            }
            const int nw = tm.n_w;
            for (int x = 0; x < nw; x++) {
                ld_elem<NL>(fac, w + ((size_t)tm.w[x] * rank + j) * NL);
                mont_mul<NL, QS>(t, term, fac, F);
#pragma unroll
                for (int l = 0; l < NL; l++) term[l] = t[l];
            }
            mod_add<NL>(ev, ev, term, F);
        }
        st_elem<NL>(out + j * NL, acc);
    }
}

int eval_circuit(const rs_field *f, uint32_t *out, const uint32_t *pw, const uint32_t *w, const rs_term *d_terms,
                 int n_terms, const uint32_t *d_coeffs, int n_coeffs, const uint32_t *bc, size_t rank) {
#define CALL(NL, QS)                                                                                       \
    {                                                                                                      \
        FieldParams<NL> F = make_params<NL>(f);                                                            \
        Scalar<NL> s;                                                                                      \
        for (int i = 0; i < NL; i++) s.v[i] = bc[i];                                                       \
        const size_t smem = (size_t)n_terms * sizeof(rs_term) + (size_t)n_coeffs * NL * 4;                 \
        const int staged = smem <= 40 * 1024;                                                              \
        eval_circuit_kernel<NL, QS><<<grid_for(rank, 128), 128, staged ? smem : 0, current_stream()>>>(out, pw, w, d_terms, n_terms, \
                                                                                  d_coeffs, n_coeffs, s, F, rank, staged);   \
    }
    RS_FIELD_DISPATCH(f, CALL)
#undef CALL
    count_launch();
    RS_CUDA_CHECK(cudaGetLastError());
    return RS_OK;
}

// ---------------------------------------------------------------------------
// QuoRemByVanishing (cyclic.go:18-37).  The sequential fold over i = rank-1 .. N is, per
// residue class r = i mod N, a suffix sum over the blocks k = i / N:
//   S_k = p[kN + r] + S_{k+1}  (S beyond the top = 0);  quo[kN + r] = S_{k+1};
//   rem[r] = S_0; rem[i >= N] = 0.
// (a partial top block, rank not a multiple of N, is handled by bounds checks)
// ---------------------------------------------------------------------------
template <int NL>
__global__ void __launch_bounds__(256) quorem_kernel(uint32_t *quo, uint32_t *rem, const uint32_t *p,
                                                     const __grid_constant__ FieldParams<NL> F, size_t rank, size_t N) {
    for (size_t r = blockIdx.x * (size_t)blockDim.x + threadIdx.x; r < N && r < rank; r += (size_t)gridDim.x * blockDim.x) {
        uint32_t S[NL], x[NL];
#pragma unroll
        for (int l = 0; l < NL; l++) S[l] = 0;
        const size_t nblk = (rank - r + N - 1) / N;   // blocks containing residue r
        for (size_t k = nblk; k-- > 0;) {
            const size_t i = k * N + r;
            st_elem<NL>(quo + i * NL, S);                 // quo[kN + r] = S_{k+1}
            ld_elem<NL>(x, p + i * NL);
            mod_add<NL>(S, S, x, F);
            if (k > 0) {
#pragma unroll
                for (int l = 0; l < NL; l++) x[l] = 0;
                st_elem<NL>(rem + i * NL, x);
            }
        }
        st_elem<NL>(rem + r * NL, S);
    }
    // rank < N leaves nothing to fold; rem = p is handled by the caller
}

int quorem(const rs_field *f, uint32_t *quo, uint32_t *rem, const uint32_t *p, size_t rank, size_t N) {
#define CALL(NL, QS)                                                                                   \
    {                                                                                                  \
        FieldParams<NL> F = make_params<NL>(f);                                                        \
        quorem_kernel<NL><<<grid_for(N, 256), 256, 0, current_stream()>>>(quo, rem, p, F, rank, N);     \
    }
    RS_FIELD_DISPATCH(f, CALL)
#undef CALL
    count_launch();
    RS_CUDA_CHECK(cudaGetLastError());
    return RS_OK;
}

// ---------------------------------------------------------------------------
// automorphism X -> X^idx (cyclotomic.go:49-82), gather form
//   coefficient domain: out[j] = +-p[i] with j = i*idx mod 2N  <=>  i = j*idx^-1 mod 2N
//   NTT domain: out[brv(i)] = p[brv(((2i+1)*idx mod 2N - 1)/2)]
// ---------------------------------------------------------------------------
template <int NL>
__global__ void __launch_bounds__(256) aut_kernel(uint32_t *out, const uint32_t *p,
                                                  const __grid_constant__ FieldParams<NL> F, int logN,
                                                  unsigned long long idx, unsigned long long idx_inv, int is_ntt,
                                                  size_t n_polys) {
    const size_t N = (size_t)1 << logN;
    const unsigned long long mask2 = 2 * N - 1;
    for (size_t g = blockIdx.x * (size_t)blockDim.x + threadIdx.x; g < n_polys * N; g += (size_t)gridDim.x * blockDim.x) {
        const size_t poly = g >> logN, j = g & (N - 1);
        uint32_t x[NL];
        if (!is_ntt) {
            // source i with i*idx = j or j+N (mod 2N)
            unsigned long long i = ((unsigned long long)j * idx_inv) & mask2;
            bool neg = false;
            if (i >= N) { i -= N; neg = true; }   // (i+N)*idx = j + N (idx odd)  =>  sign flips
            ld_elem<NL>(x, p + (poly * N + i) * NL);
            if (neg) { uint32_t y[NL]; mod_neg<NL>(y, x, F);
#pragma unroll
                for (int l = 0; l < NL; l++) x[l] = y[l]; }
        } else {
            const unsigned long long i = logN ? (__brevll((unsigned long long)j) >> (64 - logN)) : 0;   // out index j = brv(i)
            unsigned long long s = (((2 * i + 1) * idx) & mask2);
            s = (s - 1) >> 1;
            const unsigned long long src = logN ? (__brevll(s) >> (64 - logN)) : 0;
            ld_elem<NL>(x, p + (poly * N + src) * NL);
        }
        st_elem<NL>(out + g * NL, x);
    }
}

int aut(const rs_field *f, uint32_t *out, const uint32_t *p, int logN, unsigned long long idx,
        unsigned long long idx_inv, int is_ntt, size_t n_polys) {
#define CALL(NL, QS)                                                                                        \
    {                                                                                                       \
        FieldParams<NL> F = make_params<NL>(f);                                                             \
        aut_kernel<NL><<<grid_for(n_polys << logN, 256), 256, 0, current_stream()>>>(out, p, F, logN, idx,   \
                                                                                     idx_inv, is_ntt, n_polys); \
    }
    RS_FIELD_DISPATCH(f, CALL)
#undef CALL
    count_launch();
    RS_CUDA_CHECK(cudaGetLastError());
    return RS_OK;
}

// out[i] = v[N-1-i] * scale   (buckler/linear.go:38-41)
template <int NL, int QS>
__global__ void __launch_bounds__(256) reverse_scale_kernel(uint32_t *out, const uint32_t *v,
                                                            const __grid_constant__ Scalar<NL> c,
                                                            const __grid_constant__ FieldParams<NL> F, size_t N) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < N; i += (size_t)gridDim.x * blockDim.x) {
        uint32_t x[NL], z[NL];
        ld_elem<NL>(x, v + (N - 1 - i) * NL);
        mont_mul<NL, QS>(z, x, c.v, F);
        st_elem<NL>(out + i * NL, z);
    }
}

int reverse_scale(const rs_field *f, uint32_t *out, const uint32_t *v, const uint32_t *scale, size_t N) {
#define CALL(NL, QS)                                                                                      \
    {                                                                                                     \
        FieldParams<NL> F = make_params<NL>(f);                                                           \
        Scalar<NL> s;                                                                                     \
        for (int i = 0; i < NL; i++) s.v[i] = scale[i];                                                   \
        reverse_scale_kernel<NL, QS><<<grid_for(N, 256), 256, 0, current_stream()>>>(out, v, s, F, N);     \
    }
    RS_FIELD_DISPATCH(f, CALL)
#undef CALL
    count_launch();
    RS_CUDA_CHECK(cudaGetLastError());
    return RS_OK;
}

// ---------------------------------------------------------------------------
// linCheckVec (buckler/prover.go:409-413): out[i] = c^i.  The reference runs N sequential products; here thread i
// computes its power by square-and-multiply (exact field arithmetic: the same canonical elements).
// ---------------------------------------------------------------------------
template <int NL, int QS>
__global__ void __launch_bounds__(128) powers_kernel(uint32_t *out, const __grid_constant__ Scalar<NL> c,
                                                     const __grid_constant__ FieldParams<NL> F, size_t n) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t r[NL], b[NL], t[NL];
#pragma unroll
    for (int l = 0; l < NL; l++) { r[l] = F.one[l]; b[l] = c.v[l]; }
    for (size_t e = i; e; e >>= 1) {
        if (e & 1) {
            mont_mul<NL, QS>(t, r, b, F);
#pragma unroll
            for (int l = 0; l < NL; l++) r[l] = t[l];
        }
        mont_mul<NL, QS>(t, b, b, F);
#pragma unroll
        for (int l = 0; l < NL; l++) b[l] = t[l];
    }
    st_elem<NL>(out + i * NL, r);
}

int powers(const rs_field *f, uint32_t *out, const uint32_t *c, size_t n) {
#define CALL(NL, QS)                                                                                      \
    {                                                                                                     \
        FieldParams<NL> F = make_params<NL>(f);                                                           \
        Scalar<NL> s;                                                                                     \
        for (int i = 0; i < NL; i++) s.v[i] = c[i];                                                       \
        powers_kernel<NL, QS><<<(unsigned)((n + 127) / 128), 128, 0, current_stream()>>>(out, s, F, n);    \
    }
    RS_FIELD_DISPATCH(f, CALL)
#undef CALL
    count_launch();
    RS_CUDA_CHECK(cudaGetLastError());
    return RS_OK;
}

// ---------------------------------------------------------------------------
// Encoder.RandEncodeTo's blinding (buckler/encoder.go:50-54) for a batch of witnesses: c[N] <- r ; c[0] <- c[0] - c[N].
// One thread per witness polynomial.
// ---------------------------------------------------------------------------
template <int NL>
__global__ void blind_kernel(uint32_t *polys, const uint32_t *blind, size_t n, size_t rank, size_t N, const __grid_constant__ FieldParams<NL> F) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t r[NL], c0[NL], d[NL];
    ld_elem<NL>(r, blind + i * NL);
    uint32_t *p = polys + i * rank * NL;
    ld_elem<NL>(c0, p);
    mod_sub<NL>(d, c0, r, F);
    st_elem<NL>(p + N * NL, r);
    st_elem<NL>(p, d);
}

int blind(const rs_field *f, uint32_t *polys, const uint32_t *d_blind, size_t n, size_t rank, size_t N) {
#define CALL(NL, QS)                                                                                      \
    {                                                                                                     \
        FieldParams<NL> F = make_params<NL>(f);                                                           \
        blind_kernel<NL><<<(unsigned)((n + 63) / 64), 64, 0, current_stream()>>>(polys, d_blind, n, rank, N, F); \
    }
    RS_FIELD_DISPATCH(f, CALL)
#undef CALL
    count_launch();
    RS_CUDA_CHECK(cudaGetLastError());
    return RS_OK;
}

// ---------------------------------------------------------------------------
// Poly.Evaluate (poly.go:64-76): Horner is N sequential products on the CPU; here each
// thread Horner-evaluates a contiguous chunk of CH coefficients and the partial values are
// combined as  sum_t part_t * x^(t*CH)  by a second tiny kernel (exact field arithmetic, so
// the result is the same field element).
// ---------------------------------------------------------------------------
template <int NL, int QS>
__global__ void __launch_bounds__(128) horner_chunks_kernel(uint32_t *partial, const uint32_t *coeffs,
                                                            const __grid_constant__ Scalar<NL> x,
                                                            const __grid_constant__ FieldParams<NL> F, size_t n, int CH) {
    const size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const size_t n_chunks = (n + CH - 1) / CH;
    if (t >= n_chunks) return;
    const size_t lo = t * CH, hi = (lo + CH < n) ? lo + CH : n;
    uint32_t acc[NL], c[NL], tmp[NL];
#pragma unroll
    for (int l = 0; l < NL; l++) acc[l] = 0;
    for (size_t i = hi; i-- > lo;) {
        mont_mul<NL, QS>(tmp, acc, x.v, F);
        ld_elem<NL>(c, coeffs + i * NL);
        mod_add<NL>(acc, tmp, c, F);
    }
    st_elem<NL>(partial + t * NL, acc);
}

template <int NL, int QS>
__global__ void horner_combine_kernel(uint32_t *result, const uint32_t *partial,
                                      const __grid_constant__ Scalar<NL> xch,
                                      const __grid_constant__ FieldParams<NL> F, size_t n_chunks) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    uint32_t acc[NL], c[NL], tmp[NL];
#pragma unroll
    for (int l = 0; l < NL; l++) acc[l] = 0;
    for (size_t i = n_chunks; i-- > 0;) {
        mont_mul<NL, QS>(tmp, acc, xch.v, F);
        ld_elem<NL>(c, partial + i * NL);
        mod_add<NL>(acc, tmp, c, F);
    }
    st_elem<NL>(result, acc);
}

int horner(const rs_field *f, uint32_t *d_result, uint32_t *d_partial, const uint32_t *coeffs, size_t n,
           const uint32_t *x, const uint32_t *x_pow_ch, int CH) {
    const size_t n_chunks = (n + CH - 1) / CH;
#define CALL(NL, QS)                                                                                          \
    {                                                                                                         \
        FieldParams<NL> F = make_params<NL>(f);                                                               \
        Scalar<NL> sx, sp;                                                                                    \
        for (int i = 0; i < NL; i++) { sx.v[i] = x[i]; sp.v[i] = x_pow_ch[i]; }                               \
        horner_chunks_kernel<NL, QS><<<(unsigned)((n_chunks + 127) / 128), 128, 0, current_stream()>>>(        \
            d_partial, coeffs, sx, F, n, CH);                                                                 \
        if (d_result) horner_combine_kernel<NL, QS><<<1, 32, 0, current_stream()>>>(d_result, d_partial, sp, F, n_chunks);   \
    }
    RS_FIELD_DISPATCH(f, CALL)
#undef CALL
    count_launch(d_result ? 2 : 1);
    RS_CUDA_CHECK(cudaGetLastError());
    return RS_OK;
}

// ---------------------------------------------------------------------------
// Poly.Evaluate of many polynomials at one point (jindo.Prover.Evaluate's closing loop, jindo/prover.go:318-321: one
// evaluation per committed vector of the batch).  One CTA per polynomial: thread t Horner-evaluates a contiguous chunk
// of CH = ceil(n / 256) coefficients, scales it by x^(t CH) (square-and-multiply on its own), and the 256 terms are
// summed in shared memory.  Exact field arithmetic: the same element as the reference's sequential Horner.
// ---------------------------------------------------------------------------
struct EvalDesc { const uint32_t *coeffs; size_t n; };

template <int NL, int QS>
__device__ __forceinline__ void pow_small(uint32_t *r, const uint32_t *base, size_t e, const FieldParams<NL> &F) {
    uint32_t b[NL], t[NL];
#pragma unroll
    for (int l = 0; l < NL; l++) { r[l] = F.one[l]; b[l] = base[l]; }
    for (; e; e >>= 1) {
        if (e & 1) {
            mont_mul<NL, QS>(t, r, b, F);
#pragma unroll
            for (int l = 0; l < NL; l++) r[l] = t[l];
        }
        if (e > 1) {
            mont_mul<NL, QS>(t, b, b, F);
#pragma unroll
            for (int l = 0; l < NL; l++) b[l] = t[l];
        }
    }
}

// ---------------------------------------------------------------------------
// Twiddle tables of NewCyclicTransformer / NewCyclotomicTransformer (math/bigpoly/ntt.go:56-84, :183-192) generated in HBM:
// the reference fills them with N sequential products; entry k is a fixed power of the generator, so thread k computes it by
// square-and-multiply (exact field arithmetic: the same canonical elements, compared limb for limb with the oracle's tables).
//   negacyclic: tw[brv_logN(i)] = g^i                         cyclic: tw[m + i] = g^(brv_(logN-1)(i)), i < m, m = 1, 2, .., N/2
// ---------------------------------------------------------------------------
template <int NL, int QS>
__global__ void __launch_bounds__(128) twiddle_kernel(uint32_t *tw, uint32_t *twinv, const __grid_constant__ Scalar<NL> g,
                                                      const __grid_constant__ Scalar<NL> ginv, const __grid_constant__ FieldParams<NL> F,
                                                      int logN, int negacyclic) {
    const size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const size_t N = (size_t)1 << logN;
    if (k >= N) return;
    uint32_t r[NL];
    size_t e;
    if (negacyclic) {
        e = (size_t)(__brevll((unsigned long long)k) >> (64 - logN));            // k = brv(i)  <=>  i = brv(k)
    } else {
        if (k == 0) {                                                            // tw[0] is never read: the reference leaves it zero
#pragma unroll
            for (int l = 0; l < NL; l++) r[l] = 0;
            st_elem<NL>(tw, r); st_elem<NL>(twinv, r);
            return;
        }
        const int mb = 63 - __clzll((unsigned long long)k);                      // m = 2^mb <= k
        const size_t i = k - ((size_t)1 << mb);
        e = logN > 1 ? (size_t)(__brevll((unsigned long long)i) >> (64 - (logN - 1))) : 0;
    }
    pow_small<NL, QS>(r, g.v, e, F);
    st_elem<NL>(tw + k * NL, r);
    pow_small<NL, QS>(r, ginv.v, e, F);
    st_elem<NL>(twinv + k * NL, r);
}

template <int NL, int QS>
__global__ void __launch_bounds__(256) evaluate_many_kernel(uint32_t *results, const EvalDesc *descs, const __grid_constant__ Scalar<NL> x,
                                                            const __grid_constant__ FieldParams<NL> F) {
    __shared__ uint32_t sh[256 * NL];
    const EvalDesc D = descs[blockIdx.x];
    const int t = threadIdx.x;
    const size_t CH = (D.n + 255) / 256;
    const size_t lo = (size_t)t * CH, hi = lo + CH < D.n ? lo + CH : D.n;
    uint32_t acc[NL], c[NL], tmp[NL];
#pragma unroll
    for (int l = 0; l < NL; l++) acc[l] = 0;
    if (lo < D.n) {
        for (size_t i = hi; i-- > lo;) {
            mont_mul<NL, QS>(tmp, acc, x.v, F);
            ld_elem<NL>(c, D.coeffs + i * NL);
            mod_add<NL>(acc, tmp, c, F);
        }
        if (t) {
            pow_small<NL, QS>(c, x.v, CH, F);         // x^CH
            pow_small<NL, QS>(tmp, c, (size_t)t, F);  // x^(t CH)
            mont_mul<NL, QS>(c, acc, tmp, F);
#pragma unroll
            for (int l = 0; l < NL; l++) acc[l] = c[l];
        }
    }
#pragma unroll
    for (int l = 0; l < NL; l++) sh[t * NL + l] = acc[l];
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (t < s) {
#pragma unroll
            for (int l = 0; l < NL; l++) { acc[l] = sh[t * NL + l]; c[l] = sh[(t + s) * NL + l]; }
            mod_add<NL>(tmp, acc, c, F);
#pragma unroll
            for (int l = 0; l < NL; l++) sh[t * NL + l] = tmp[l];
        }
        __syncthreads();
    }
    if (t < NL) results[(size_t)blockIdx.x * NL + t] = sh[t];
}

// Shoup pairs of a Montgomery table (shoup.cuh): entry e = w R mod q  ->  (w, floor(w R / q)) = (e R^-1 mod q, e * (-q^-1) mod R)
template <int NL, int QS>
__global__ void __launch_bounds__(128) shoup_table_kernel(uint32_t *sh, const uint32_t *tw, size_t n, const __grid_constant__ Scalar<NL> qinv_full,
                                                          const __grid_constant__ FieldParams<NL> F) {
    const size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (k >= n) return;
    uint32_t e[NL], w[NL], one[NL], E[NL], O[NL], wp[NL];
    ld_elem<NL>(e, tw + k * NL);
#pragma unroll
    for (int l = 0; l < NL; l++) one[l] = l == 0;
    mont_mul<NL, QS>(w, e, one, F);
    tri_lo<NL, false>(E, O, e, qinv_full.v);
    tri_merge<NL>(wp, E, O);
    st_elem<NL>(sh + 2 * k * NL, w);
    st_elem<NL>(sh + (2 * k + 1) * NL, wp);
}

int shoup_tables(const rs_field *f, uint32_t *d_sh, const uint32_t *d_tw, size_t n) {
#define CALL(NL, QS)                                                                                          \
    {                                                                                                         \
        FieldParams<NL> F = make_params<NL>(f);                                                               \
        Scalar<NL> qi;                                                                                        \
        for (int i = 0; i < NL; i++) qi.v[i] = f->qinv_full[i];                                               \
        shoup_table_kernel<NL, QS><<<(unsigned)((n + 127) / 128), 128, 0, current_stream()>>>(d_sh, d_tw, n, qi, F); \
    }
    RS_FIELD_DISPATCH(f, CALL)
#undef CALL
    count_launch();
    RS_CUDA_CHECK(cudaGetLastError());
    return RS_OK;
}

int twiddles(const rs_field *f, uint32_t *d_tw, uint32_t *d_twinv, const uint32_t *g, const uint32_t *ginv, int logN, int negacyclic) {
    const size_t N = (size_t)1 << logN;
#define CALL(NL, QS)                                                                                          \
    {                                                                                                         \
        FieldParams<NL> F = make_params<NL>(f);                                                               \
        Scalar<NL> sg, si;                                                                                    \
        for (int i = 0; i < NL; i++) { sg.v[i] = g[i]; si.v[i] = ginv[i]; }                                   \
        twiddle_kernel<NL, QS><<<(unsigned)((N + 127) / 128), 128, 0, current_stream()>>>(d_tw, d_twinv, sg, si, F, logN, negacyclic); \
    }
    RS_FIELD_DISPATCH(f, CALL)
#undef CALL
    count_launch();
    RS_CUDA_CHECK(cudaGetLastError());
    return RS_OK;
}

int evaluate_many(const rs_field *f, uint32_t *d_results, const void *d_descs, size_t n_polys, const uint32_t *x) {
#define CALL(NL, QS)                                                                                          \
    {                                                                                                         \
        FieldParams<NL> F = make_params<NL>(f);                                                               \
        Scalar<NL> sx;                                                                                        \
        for (int i = 0; i < NL; i++) sx.v[i] = x[i];                                                          \
        evaluate_many_kernel<NL, QS><<<(unsigned)n_polys, 256, 0, current_stream()>>>(d_results, (const EvalDesc *)d_descs, sx, F); \
    }
    RS_FIELD_DISPATCH(f, CALL)
#undef CALL
    count_launch();
    RS_CUDA_CHECK(cudaGetLastError());
    return RS_OK;
}

}  // namespace rs
