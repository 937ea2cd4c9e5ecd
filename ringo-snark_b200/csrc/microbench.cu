// microbench.cu -- measures the integer-pipe roofline denominator used by bench.py:
// sustained 32x32+64 multiply-add (IMAD.WIDE.U32) throughput of the whole chip.
#include "common.cuh"

namespace rs {

template <bool WIDE>
__global__ void __launch_bounds__(256) imad_kernel(unsigned long long *out, uint32_t a, uint32_t b, int iters) {
    unsigned long long acc[8];
    uint32_t lo[8];
#pragma unroll
    for (int k = 0; k < 8; k++) { acc[k] = threadIdx.x + k; lo[k] = threadIdx.x * 3 + k; }
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 16; u++) {
#pragma unroll
            for (int k = 0; k < 8; k++) {
                if (WIDE) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[k]) : "r"(a + k), "r"(b));
                else asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(lo[k]) : "r"(a + k), "r"(b));
            }
        }
    }
    unsigned long long s = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) s += acc[k] + lo[k];
    if (s == 0x1234567ull) out[0] = s;   // keep the chains alive
}

// carry-chained form used by the Montgomery product: mad.lo.cc / madc.hi.cc pairs that ptxas
// fuses into IMAD.WIDE.U32.X (carry in and out through predicates)
__global__ void __launch_bounds__(256) imad_chain_kernel(unsigned long long *out, uint32_t a, uint32_t b, int iters) {
    uint32_t x[8], y[8];
#pragma unroll
    for (int k = 0; k < 8; k++) { x[k] = threadIdx.x + k; y[k] = threadIdx.x * 3 + k; }
    const uint32_t a0 = a, a1 = a + 1, a2 = a + 2, a3 = a + 3;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 16; u++) {
            asm volatile("mad.lo.cc.u32 %0, %1, %2, %0;" : "+r"(x[0]) : "r"(a0), "r"(b));
            asm volatile("madc.hi.cc.u32 %0, %1, %2, %0;" : "+r"(x[1]) : "r"(a0), "r"(b));
            asm volatile("madc.lo.cc.u32 %0, %1, %2, %0;" : "+r"(x[2]) : "r"(a1), "r"(b));
            asm volatile("madc.hi.cc.u32 %0, %1, %2, %0;" : "+r"(x[3]) : "r"(a1), "r"(b));
            asm volatile("madc.lo.cc.u32 %0, %1, %2, %0;" : "+r"(x[4]) : "r"(a2), "r"(b));
            asm volatile("madc.hi.cc.u32 %0, %1, %2, %0;" : "+r"(x[5]) : "r"(a2), "r"(b));
            asm volatile("madc.lo.cc.u32 %0, %1, %2, %0;" : "+r"(x[6]) : "r"(a3), "r"(b));
            asm volatile("madc.hi.u32 %0, %1, %2, %0;" : "+r"(x[7]) : "r"(a3), "r"(b));
            asm volatile("mad.lo.cc.u32 %0, %1, %2, %0;" : "+r"(y[0]) : "r"(a0), "r"(b));
            asm volatile("madc.hi.cc.u32 %0, %1, %2, %0;" : "+r"(y[1]) : "r"(a0), "r"(b));
            asm volatile("madc.lo.cc.u32 %0, %1, %2, %0;" : "+r"(y[2]) : "r"(a1), "r"(b));
            asm volatile("madc.hi.cc.u32 %0, %1, %2, %0;" : "+r"(y[3]) : "r"(a1), "r"(b));
            asm volatile("madc.lo.cc.u32 %0, %1, %2, %0;" : "+r"(y[4]) : "r"(a2), "r"(b));
            asm volatile("madc.hi.cc.u32 %0, %1, %2, %0;" : "+r"(y[5]) : "r"(a2), "r"(b));
            asm volatile("madc.lo.cc.u32 %0, %1, %2, %0;" : "+r"(y[6]) : "r"(a3), "r"(b));
            asm volatile("madc.hi.u32 %0, %1, %2, %0;" : "+r"(y[7]) : "r"(a3), "r"(b));
        }
    }
    unsigned long long s = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) s += x[k] + y[k];
    if (s == 0x1234567ull) out[0] = s;
}

// add-with-carry chains (IADD3.X), 8 limbs
__global__ void __launch_bounds__(256) iadd_chain_kernel(unsigned long long *out, uint32_t a, int iters) {
    uint32_t x[8], y[8];
#pragma unroll
    for (int k = 0; k < 8; k++) { x[k] = threadIdx.x + k; y[k] = threadIdx.x * 3 + k; }
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 16; u++) {
            asm volatile("add.cc.u32 %0, %0, %1;" : "+r"(x[0]) : "r"(a));
#pragma unroll
            for (int k = 1; k < 7; k++) asm volatile("addc.cc.u32 %0, %0, %1;" : "+r"(x[k]) : "r"(a + k));
            asm volatile("addc.u32 %0, %0, %1;" : "+r"(x[7]) : "r"(a));
            asm volatile("add.cc.u32 %0, %0, %1;" : "+r"(y[0]) : "r"(a));
#pragma unroll
            for (int k = 1; k < 7; k++) asm volatile("addc.cc.u32 %0, %0, %1;" : "+r"(y[k]) : "r"(a + k));
            asm volatile("addc.u32 %0, %0, %1;" : "+r"(y[7]) : "r"(a));
        }
    }
    unsigned long long s = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) s += x[k] + y[k];
    if (s == 0x1234567ull) out[0] = s;
}

// mixed stream: independent IMAD.WIDE (fma pipe) and 3-input adds / logic ops (alu pipe), 1:1
__global__ void __launch_bounds__(256) mixed_kernel(unsigned long long *out, uint32_t a, uint32_t b, int iters) {
    unsigned long long acc[8];
    uint32_t x[8];
#pragma unroll
    for (int k = 0; k < 8; k++) { acc[k] = threadIdx.x + k; x[k] = threadIdx.x * 3 + k; }
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 16; u++) {
#pragma unroll
            for (int k = 0; k < 8; k++) {
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[k]) : "r"(a + k), "r"(b));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[k]) : "r"(a), "r"(b + k));
            }
        }
    }
    unsigned long long s = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) s += acc[k] + x[k];
    if (s == 0x1234567ull) out[0] = s;
}

// IMAD.WIDE.U32 with a 64-bit addend and operands that change every iteration (mode 1's multiplicands are loop
// invariant: ptxas hoists the product and the loop degenerates into IADD3 pairs).  The multiply-add is written the
// way the kernels write it (mad.lo.cc + madc.hi on a packed 64-bit register -> ONE IMAD.WIDE.U32 R, a, b, R), and one
// factor is the low word of a neighbouring accumulator, so nothing can be hoisted; 8 independent chains per thread.
__global__ void __launch_bounds__(256) imad_wide_acc_kernel(unsigned long long *out, uint32_t b, int iters) {
    unsigned long long acc[8];
#pragma unroll
    for (int k = 0; k < 8; k++) acc[k] = (unsigned long long)(threadIdx.x * 2654435761u + k) * 0x9e3779b97f4a7c15ull;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 16; u++) {
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const uint32_t f = (uint32_t)acc[(k + 3) & 7];
                asm volatile("{\n\t.reg .u32 lo, hi;\n\tmov.b64 {lo, hi}, %0;\n\tmad.lo.cc.u32 lo, %1, %2, lo;\n\tmadc.hi.u32 hi, %1, %2, hi;\n\tmov.b64 %0, {lo, hi};\n\t}"
                             : "+l"(acc[k]) : "r"(f), "r"(b));
            }
        }
    }
    unsigned long long s = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) s += acc[k];
    if (s == 0x1234567ull) out[0] = s;
}

// ---- FP64 pipe (the roofline denominator of the ntt52 kernels) --------------------------------------------------
// mode 6: independent DFMA chains, operands change every iteration (nothing to hoist)
__global__ void __launch_bounds__(256) dfma_kernel(unsigned long long *out, double b, int iters) {
    double acc[8];
#pragma unroll
    for (int k = 0; k < 8; k++) acc[k] = 1.0 + threadIdx.x * 1e-3 + k;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 16; u++) {
#pragma unroll
            for (int k = 0; k < 8; k++) acc[k] = __fma_rz(acc[(k + 3) & 7], b, acc[k]);
        }
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) s += acc[k];
    if (s == 0.12345) out[0] = (unsigned long long)s;
}

// mode 7: DFMA and IMAD.WIDE.U32 interleaved 1:1 on independent chains: do the FP64 and FMA pipes run concurrently?
__global__ void __launch_bounds__(256) dfma_imad_kernel(unsigned long long *out, double b, uint32_t m, int iters) {
    double acc[8];
    unsigned long long ia[8];
#pragma unroll
    for (int k = 0; k < 8; k++) { acc[k] = 1.0 + threadIdx.x * 1e-3 + k; ia[k] = (unsigned long long)(threadIdx.x * 2654435761u + k) * 0x9e3779b97f4a7c15ull; }
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 16; u++) {
#pragma unroll
            for (int k = 0; k < 8; k++) {
                acc[k] = __fma_rz(acc[(k + 3) & 7], b, acc[k]);
                const uint32_t f = (uint32_t)ia[(k + 3) & 7];
                asm volatile("{\n\t.reg .u32 lo, hi;\n\tmov.b64 {lo, hi}, %0;\n\tmad.lo.cc.u32 lo, %1, %2, lo;\n\tmadc.hi.u32 hi, %1, %2, hi;\n\tmov.b64 %0, {lo, hi};\n\t}"
                             : "+l"(ia[k]) : "r"(f), "r"(m));
            }
        }
    }
    double s = 0; unsigned long long t = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) { s += acc[k]; t += ia[k]; }
    if (s == 0.12345 || t == 0x1234567ull) out[0] = (unsigned long long)s + t;
}

// mode 8: the exact limb-product split of field52.cuh (DFMA.RZ, DADD, DFMA.RZ) with its two 64-bit pattern
// accumulations -- counts limb products
__global__ void __launch_bounds__(256) split52_kernel(unsigned long long *out, double b, int iters) {
    double a[8];
    unsigned long long c0[8], c1[8];
#pragma unroll
    for (int k = 0; k < 8; k++) { a[k] = (double)((threadIdx.x * 2654435761u + k) & 0xfffff) * 4294967296.0 + k; c0[k] = k; c1[k] = 2 * k; }
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 16; u++) {
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const double ph = __fma_rz(a[k], b, 0x1p104);
                const double sub = (0x1p104 + 0x1p52) - ph;
                const double pl = __fma_rz(a[k], b, sub);
                c1[k] += (unsigned long long)__double_as_longlong(ph);
                c0[k] += (unsigned long long)__double_as_longlong(pl);
                // next multiplicand: low 52 bits of the running low column (keeps every product distinct)
                a[k] = __longlong_as_double((long long)((c0[k] & 0xfffffffffffffull) | 0x4330000000000000ull)) - 0x1p52;
            }
        }
    }
    unsigned long long t = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) t += c0[k] + c1[k];
    if (t == 0x1234567ull) out[0] = t;
}

// modes 9-12: one DFMA.RZ stream next to a second stream on another pipe (both counted): do they overlap?
//   9: IADD3 / IADD3.X (64-bit adds, alu pipe)   10: LOP3 (alu pipe)   11: 32-bit IMAD (fma pipe)   12: DADD alone
template <int OTHER>
__global__ void __launch_bounds__(256) dfma_mix_kernel(unsigned long long *out, double b, uint32_t m, int iters) {
    double acc[8];
    unsigned long long ia[8];
    uint32_t x[8];
#pragma unroll
    for (int k = 0; k < 8; k++) { acc[k] = 1.0 + threadIdx.x * 1e-3 + k; ia[k] = (unsigned long long)(threadIdx.x * 2654435761u + k) * 0x9e3779b97f4a7c15ull; x[k] = threadIdx.x * 3 + k; }
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 16; u++) {
#pragma unroll
            for (int k = 0; k < 8; k++) {
                if (OTHER == 12) { acc[k] = __dadd_rn(acc[(k + 3) & 7], acc[k]); continue; }
                acc[k] = __fma_rz(acc[(k + 3) & 7], b, acc[k]);
                if (OTHER == 9) {
                    // one 64-bit add = IADD3 + IADD3.X: counted as ONE op next to the DFMA
                    asm volatile("{\n\t.reg .u32 lo, hi, l2, h2;\n\tmov.b64 {lo, hi}, %0;\n\tmov.b64 {l2, h2}, %1;\n\tadd.cc.u32 lo, lo, l2;\n\taddc.u32 hi, hi, h2;\n\tmov.b64 %0, {lo, hi};\n\t}"
                                 : "+l"(ia[k]) : "l"(ia[(k + 3) & 7]));
                } else if (OTHER == 10) {
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[k]) : "r"(x[(k + 3) & 7]), "r"(m));
                } else if (OTHER == 11) {
                    asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(x[k]) : "r"(x[(k + 3) & 7]), "r"(m));
                }
            }
        }
    }
    double s = 0; unsigned long long t = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) { s += acc[k]; t += ia[k] + x[k]; }
    if (s == 0.12345 || t == 0x1234567ull) out[0] = (unsigned long long)s + t;
}

// modes 100 + 10 P + K: one primary instruction (P = 0: IMAD.WIDE.U32 with a 64-bit addend, P = 1: DFMA.RZ) followed by K
// LOP3 on independent chains.  Reported as groups per second: cycles per group = 32 lanes * 592 sub-partitions * clock / rate.
// Tells how many issue cycles the primary really costs (a pipe-bound instruction leaves K slots free, an issue-bound one not).
template <int P, int K>
__global__ void __launch_bounds__(256) slot_kernel(unsigned long long *out, double b, uint32_t m, int iters) {
    double acc[8];
    unsigned long long ia[8];
    uint32_t x[8][4];
#pragma unroll
    for (int k = 0; k < 8; k++) {
        acc[k] = 1.0 + threadIdx.x * 1e-3 + k; ia[k] = (unsigned long long)(threadIdx.x * 2654435761u + k) * 0x9e3779b97f4a7c15ull;
#pragma unroll
        for (int j = 0; j < 4; j++) x[k][j] = threadIdx.x * 3 + k + 17 * j;
    }
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 16; u++) {
#pragma unroll
            for (int k = 0; k < 8; k++) {
                if (P == 1) acc[k] = __fma_rz(acc[(k + 3) & 7], b, acc[k]);
                else {
                    const uint32_t f = (uint32_t)ia[(k + 3) & 7];
                    asm volatile("{\n\t.reg .u32 lo, hi;\n\tmov.b64 {lo, hi}, %0;\n\tmad.lo.cc.u32 lo, %1, %2, lo;\n\tmadc.hi.u32 hi, %1, %2, hi;\n\tmov.b64 %0, {lo, hi};\n\t}"
                                 : "+l"(ia[k]) : "r"(f), "r"(m));
                }
#pragma unroll
                for (int j = 0; j < K; j++) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[k][j & 3]) : "r"(x[(k + 3) & 7][(j + 1) & 3]), "r"(m));
            }
        }
    }
    double s = 0; unsigned long long t = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) { s += acc[k]; t += ia[k] + x[k][0] + x[k][1] + x[k][2] + x[k][3]; }
    if (s == 0.12345 || t == 0x1234567ull) out[0] = (unsigned long long)s + t;
}

}  // namespace rs

extern "C" {
#pragma GCC visibility push(default)
// Returns sustained G op/s.  wide: 0 = 32-bit IMAD, 1 = IMAD.WIDE.U32 (32x32+64, no carry),
// 2 = carry-chained IMAD.WIDE.U32.X (the Montgomery inner loop), 3 = IADD3.X carry chains,
// 4 = 1:1 mix of independent IMAD.WIDE.U32 and LOP3 (counts both): can the two pipes run concurrently?
// 5 = IMAD.WIDE.U32 with a 64-bit addend and changing operands (what the kernels' multiply-adds are; mode 1's loop-invariant
//     product is hoisted by ptxas, so mode 1 really measures IADD3 pairs)
// 6 = DFMA.RZ alone; 7 = DFMA.RZ and IMAD.WIDE.U32 interleaved 1:1 (counts both); 8 = 52x52-bit limb products through the
//     DFMA.RZ / DADD / DFMA.RZ split of field52.cuh incl. the two 64-bit pattern accumulations (counts limb products)
// 9 / 10 / 11 = DFMA.RZ next to a 64-bit add (IADD3 + IADD3.X, counted as one) / LOP3 / 32-bit IMAD, 1:1, both counted; 12 = DADD
int rs_bench_imad_peak(int wide, double *gops) {
    using namespace rs;
    RS_REQUIRE(gops, RS_ERR_ARG, "rs_bench_imad_peak: null argument");
    unsigned long long *d = nullptr;
    RS_CUDA_CHECK(cudaMalloc((void **)&d, 8));
    const int iters = 4096, blocks = 148 * 8, threads = 256;
    cudaEvent_t e0, e1;
    RS_CUDA_CHECK(cudaEventCreate(&e0));
    RS_CUDA_CHECK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
        RS_CUDA_CHECK(cudaEventRecord(e0, current_stream()));
        if (wide == 1) imad_kernel<true><<<blocks, threads, 0, current_stream()>>>(d, 12345u, 6789u, iters);
        else if (wide == 2) imad_chain_kernel<<<blocks, threads, 0, current_stream()>>>(d, 12345u, 6789u, iters);
        else if (wide == 3) iadd_chain_kernel<<<blocks, threads, 0, current_stream()>>>(d, 12345u, iters);
        else if (wide == 4) mixed_kernel<<<blocks, threads, 0, current_stream()>>>(d, 12345u, 6789u, iters);
        else if (wide == 5) imad_wide_acc_kernel<<<blocks, threads, 0, current_stream()>>>(d, 0x9e3779b1u, iters);
        else if (wide == 6) dfma_kernel<<<blocks, threads, 0, current_stream()>>>(d, 0.999999, iters);
        else if (wide == 7) dfma_imad_kernel<<<blocks, threads, 0, current_stream()>>>(d, 0.999999, 0x9e3779b1u, iters);
        else if (wide == 8) split52_kernel<<<blocks, threads, 0, current_stream()>>>(d, 4503599627370449.0, iters);
        else if (wide == 9) dfma_mix_kernel<9><<<blocks, threads, 0, current_stream()>>>(d, 0.999999, 0x9e3779b1u, iters);
        else if (wide == 10) dfma_mix_kernel<10><<<blocks, threads, 0, current_stream()>>>(d, 0.999999, 0x9e3779b1u, iters);
        else if (wide == 11) dfma_mix_kernel<11><<<blocks, threads, 0, current_stream()>>>(d, 0.999999, 0x9e3779b1u, iters);
        else if (wide == 12) dfma_mix_kernel<12><<<blocks, threads, 0, current_stream()>>>(d, 0.999999, 0x9e3779b1u, iters);
#define RS_SLOT(P, K) else if (wide == 100 + 10 * P + K) slot_kernel<P, K><<<blocks, threads, 0, current_stream()>>>(d, 0.999999, 0x9e3779b1u, iters);
        RS_SLOT(0, 0) RS_SLOT(0, 1) RS_SLOT(0, 2) RS_SLOT(0, 3) RS_SLOT(0, 4) RS_SLOT(0, 5) RS_SLOT(0, 6)
        RS_SLOT(1, 0) RS_SLOT(1, 1) RS_SLOT(1, 2) RS_SLOT(1, 3) RS_SLOT(1, 4)
#undef RS_SLOT
        else imad_kernel<false><<<blocks, threads, 0, current_stream()>>>(d, 12345u, 6789u, iters);
        count_launch();
        RS_CUDA_CHECK(cudaEventRecord(e1, current_stream()));
        RS_CUDA_CHECK(cudaEventSynchronize(e1));
        float ms = 0;
        RS_CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
    // per inner iteration: 8 MADs (modes 0/1), 8 fused wide MADs (mode 2), 16 adds (mode 3)
    const double ops = (double)blocks * threads * (double)iters * 16 * (wide == 3 || wide == 4 || wide == 7 || (wide >= 9 && wide <= 11) ? 16 : 8);
    *gops = ops / (best * 1e-3) / 1e9;
    return RS_OK;
}
#pragma GCC visibility pop
}
