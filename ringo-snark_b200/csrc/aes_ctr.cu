// aes_ctr.cu -- GPU-resident uniform draws for the TwinCDT rows (SURVEY.md 8f "next" row 1).
//
// Reproduces math/csprng/uniform.go bit for bit: AES-256 in CTR mode (Go crypto/cipher.NewCTR:
// 128-bit big-endian counter starting at the IV), keystream consumed 8 bytes little-endian per
// Sample() from an 8 KiB buffer that is refilled with XORKeyStream(buf, buf) on a never-cleared
// buffer (uniform.go:64-70) -- so the n-th 8 KiB chunk handed out is the XOR of keystream chunks 0..n.
// The key / IV are SHA-384(seed) split 32 / 16 bytes by the host (uniform.go:38-46).
//
//   aes_keystream_kernel   one thread per 16-byte block: T-table AES-256 of (IV + block index)
//   xor_scan_*             prefix XOR over chunks, independently for each of the 1024 word offsets of a
//                          chunk, seeded with the sampler's running buffer state (3-phase segmented scan)
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

typedef unsigned long long u64;

namespace rs {

struct AesKey { uint32_t rk[60]; };          // AES-256 expanded key, big-endian words

__constant__ uint8_t c_sbox[256];

__device__ __forceinline__ uint32_t xtime(uint32_t x) { return ((x << 1) ^ (((x >> 7) & 1) * 0x1b)) & 0xff; }

// keystream for blocks [blk0, blk0 + n): out as 16 bytes each.
// The round table Te0 is replicated 32 times in shared memory (word index*32 + lane): every lane reads
// its own bank, so the 224 data-dependent lookups per block are conflict-free; Te1..Te3 are byte
// rotations of Te0 (one PRMT each).
__global__ void __launch_bounds__(256) aes_keystream_kernel(uint4 *__restrict__ out, const __grid_constant__ AesKey K,
                                                            u64 iv_hi, u64 iv_lo, u64 blk0, size_t n) {
    extern __shared__ uint32_t Te[];                 // [256][32]
    __shared__ uint32_t Sb[256];                     // S-box for the last round, one word per entry
    {
        const uint32_t s = c_sbox[threadIdx.x], s2 = xtime(s), s3 = s2 ^ s;
        const uint32_t t = (s2 << 24) | (s << 16) | (s << 8) | s3;       // (02 01 01 03) * S[x], big-endian word
#pragma unroll
        for (int c = 0; c < 32; c++) Te[threadIdx.x * 32 + c] = t;
        Sb[threadIdx.x] = s;
    }
    __syncthreads();
    const uint32_t *T = Te + (threadIdx.x & 31);
#define TE(b) T[(b) << 5]
#define ROT8(x) __byte_perm((x), 0, 0x0321)
#define ROT16(x) __byte_perm((x), 0, 0x1032)
#define ROT24(x) __byte_perm((x), 0, 0x2103)
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        // counter = IV + (blk0 + i) as a 128-bit big-endian integer
        const u64 add = blk0 + i;
        const u64 lo = iv_lo + add, hi = iv_hi + (lo < add);
        uint32_t s0 = (uint32_t)(hi >> 32) ^ K.rk[0], s1 = (uint32_t)hi ^ K.rk[1], s2 = (uint32_t)(lo >> 32) ^ K.rk[2], s3 = (uint32_t)lo ^ K.rk[3];
#pragma unroll
        for (int r = 1; r < 14; r++) {
            const uint32_t t0 = TE(s0 >> 24) ^ ROT8(TE((s1 >> 16) & 0xff)) ^ ROT16(TE((s2 >> 8) & 0xff)) ^ ROT24(TE(s3 & 0xff)) ^ K.rk[4 * r];
            const uint32_t t1 = TE(s1 >> 24) ^ ROT8(TE((s2 >> 16) & 0xff)) ^ ROT16(TE((s3 >> 8) & 0xff)) ^ ROT24(TE(s0 & 0xff)) ^ K.rk[4 * r + 1];
            const uint32_t t2 = TE(s2 >> 24) ^ ROT8(TE((s3 >> 16) & 0xff)) ^ ROT16(TE((s0 >> 8) & 0xff)) ^ ROT24(TE(s1 & 0xff)) ^ K.rk[4 * r + 2];
            const uint32_t t3 = TE(s3 >> 24) ^ ROT8(TE((s0 >> 16) & 0xff)) ^ ROT16(TE((s1 >> 8) & 0xff)) ^ ROT24(TE(s2 & 0xff)) ^ K.rk[4 * r + 3];
            s0 = t0; s1 = t1; s2 = t2; s3 = t3;
        }
        // final round: SubBytes + ShiftRows + AddRoundKey
        const uint32_t o0 = (Sb[s0 >> 24] << 24 | Sb[(s1 >> 16) & 0xff] << 16 | Sb[(s2 >> 8) & 0xff] << 8 | Sb[s3 & 0xff]) ^ K.rk[56];
        const uint32_t o1 = (Sb[s1 >> 24] << 24 | Sb[(s2 >> 16) & 0xff] << 16 | Sb[(s3 >> 8) & 0xff] << 8 | Sb[s0 & 0xff]) ^ K.rk[57];
        const uint32_t o2 = (Sb[s2 >> 24] << 24 | Sb[(s3 >> 16) & 0xff] << 16 | Sb[(s0 >> 8) & 0xff] << 8 | Sb[s1 & 0xff]) ^ K.rk[58];
        const uint32_t o3 = (Sb[s3 >> 24] << 24 | Sb[(s0 >> 16) & 0xff] << 16 | Sb[(s1 >> 8) & 0xff] << 8 | Sb[s2 & 0xff]) ^ K.rk[59];
        // words are big-endian byte strings; memory order = byte order
        out[i] = make_uint4(__byte_perm(o0, 0, 0x0123), __byte_perm(o1, 0, 0x0123), __byte_perm(o2, 0, 0x0123), __byte_perm(o3, 0, 0x0123));
    }
#undef TE
#undef ROT8
#undef ROT16
#undef ROT24
}

// ---------------------------------------------------------------------------------------------
// aes_keystream4_scan_kernel (long draws): the same keystream with the lookups cut to ONE byte-permute + ONE shared load each.
// The kernel above spends ~3.5 ALU instructions per table lookup (byte extract, index scaling, rotation of Te0 into
// Te1..Te3) and is ALU bound; here the four round tables are stored with a row of 256 bytes per byte value,
//     byte offset of (table t, value b, lane l)  =  (t >> 1) * 65536 + b * 256 + (t & 1) * 128 + l * 4
// from a 64 KiB-aligned shared-memory base, so that a lookup address is the lane's base address with its byte 1
// REPLACED by the state byte -- one PRMT of (state word, lane base) -- and the table is chosen by the immediate offset of
// the LDS.  Every lane still reads its own bank.  The last round takes S[b] out of byte 1 / 2 of Te0[b] with three byte-permutes
// per output word, the last of which also puts the word in memory byte order.  128 KiB of tables (+ up to 64 KiB of
// alignment slack): one 1024-thread CTA per SM.
// ---------------------------------------------------------------------------------------------
constexpr int AES4_THREADS = 1024;
constexpr int AES4_SMEM = 3 * 65536;
#define RS_LDS(v, addr, OFF) asm("ld.shared.u32 %0, [%1+" #OFF "];" : "=r"(v) : "r"(addr))
// fills the tables; returns the 64 KiB-aligned base
__device__ __forceinline__ unsigned aes4_setup(uint8_t *raw) {
    const unsigned base = ((unsigned)__cvta_generic_to_shared(raw) + 0xffffu) & ~0xffffu;   // bits 0..15 clear
    for (int e = threadIdx.x; e < 256 * 4 * 32; e += AES4_THREADS) {
        const int lane = e & 31, t = (e >> 5) & 3, b = e >> 7;
        const uint32_t sb = c_sbox[b], s2 = xtime(sb), s3 = s2 ^ sb;
        uint32_t w = (s2 << 24) | (sb << 16) | (sb << 8) | s3;           // Te0[b] = (02 01 01 03) * S[b], big-endian word
        w = t == 0 ? w : (t == 1 ? __byte_perm(w, 0, 0x0321) : (t == 2 ? __byte_perm(w, 0, 0x1032) : __byte_perm(w, 0, 0x2103)));
        const unsigned a = base + (unsigned)(t >> 1) * 65536u + (unsigned)b * 256u + (unsigned)(t & 1) * 128u + (unsigned)lane * 4u;
        asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(w) : "memory");
    }
    __syncthreads();
    return base;
}
// keystream block of counter (hi:lo), as the 16 bytes in memory order; lb = table base + 4 * lane, kl[4] = bswap of rk[56..59]
__device__ __forceinline__ uint4 aes4_block(const AesKey &K, const uint32_t *kl, unsigned lb, u64 hi, u64 lo) {
    uint32_t s0 = (uint32_t)(hi >> 32) ^ K.rk[0], s1 = (uint32_t)hi ^ K.rk[1], s2 = (uint32_t)(lo >> 32) ^ K.rk[2], s3 = (uint32_t)lo ^ K.rk[3];
    // address of (value = byte k of s): lane base with byte 1 replaced
#define A3(s) __byte_perm((s), lb, 0x7634)
#define A2(s) __byte_perm((s), lb, 0x7624)
#define A1(s) __byte_perm((s), lb, 0x7614)
#define A0(s) __byte_perm((s), lb, 0x7604)
#define RS_AES_COL(t, a, b, c, d, rkw)                                   \
    {                                                                   \
        uint32_t v0, v1, v2, v3;                                        \
        RS_LDS(v0, A3(a), 0); RS_LDS(v1, A2(b), 128); RS_LDS(v2, A1(c), 65536); RS_LDS(v3, A0(d), 65664); \
        t = v0 ^ v1 ^ v2 ^ v3 ^ (rkw);                                  \
    }
#pragma unroll
    for (int r = 1; r < 14; r++) {
        uint32_t t0, t1, t2, t3;
        RS_AES_COL(t0, s0, s1, s2, s3, K.rk[4 * r]);
        RS_AES_COL(t1, s1, s2, s3, s0, K.rk[4 * r + 1]);
        RS_AES_COL(t2, s2, s3, s0, s1, K.rk[4 * r + 2]);
        RS_AES_COL(t3, s3, s0, s1, s2, K.rk[4 * r + 3]);
        s0 = t0; s1 = t1; s2 = t2; s3 = t3;
    }
    // final round: S[b] sits in bytes 1 and 2 of Te0[b]
#define RS_AES_LAST(o, a, b, c, d, kw)                                   \
    {                                                                   \
        uint32_t v0, v1, v2, v3;                                        \
        RS_LDS(v0, A3(a), 0); RS_LDS(v1, A2(b), 0); RS_LDS(v2, A1(c), 0); RS_LDS(v3, A0(d), 0); \
        const uint32_t hi2 = __byte_perm(v0, v1, 0x2600), lo2 = __byte_perm(v2, v3, 0x0015); \
        o = __byte_perm(hi2, lo2, 0x4523) ^ (kw);                       \
    }
    uint32_t o0, o1, o2, o3;
    RS_AES_LAST(o0, s0, s1, s2, s3, kl[0]);
    RS_AES_LAST(o1, s1, s2, s3, s0, kl[1]);
    RS_AES_LAST(o2, s2, s3, s0, s1, kl[2]);
    RS_AES_LAST(o3, s3, s0, s1, s2, kl[3]);
    return make_uint4(o0, o1, o2, o3);
#undef A3
#undef A2
#undef A1
#undef A0
#undef RS_AES_COL
#undef RS_AES_LAST
}
#undef RS_LDS

// The keystream with the first two passes of the prefix-XOR scan folded in: a thread owns one 16-byte block offset j of one
// segment of `seg` consecutive 8 KiB chunks, walks the segment's chunks (block index c * 512 + j), keeps the running XOR and
// writes IT instead of the raw keystream; the segment totals go to totals[s][2j .. 2j+1].  What is left for after the prefix
// over the segments is one streaming XOR (xor_prefix_apply).
__global__ void __launch_bounds__(AES4_THREADS, 1) aes_keystream4_scan_kernel(uint4 *__restrict__ out, uint4 *__restrict__ totals,
                                                                              const __grid_constant__ AesKey K, u64 iv_hi, u64 iv_lo,
                                                                              u64 blk0, size_t n_chunks, int seg) {
    extern __shared__ __align__(16) uint8_t aes4_raw[];
    const unsigned lb = aes4_setup(aes4_raw) + (threadIdx.x & 31) * 4;
    const uint32_t kl[4] = {__byte_perm(K.rk[56], 0, 0x0123), __byte_perm(K.rk[57], 0, 0x0123), __byte_perm(K.rk[58], 0, 0x0123),
                            __byte_perm(K.rk[59], 0, 0x0123)};
    const size_t n_seg = (n_chunks + seg - 1) / seg;
    for (size_t pair = blockIdx.x * (size_t)AES4_THREADS + threadIdx.x; pair < n_seg * 512; pair += (size_t)gridDim.x * AES4_THREADS) {
        const size_t s = pair >> 9, j = pair & 511;
        const size_t c0 = s * seg, c1 = (c0 + seg < n_chunks) ? c0 + seg : n_chunks;
        uint4 acc = make_uint4(0, 0, 0, 0);
        for (size_t c = c0; c < c1; c++) {
            const size_t i = c * 512 + j;
            const u64 add = blk0 + i;
            const u64 lo = iv_lo + add, hi = iv_hi + (lo < add);
            const uint4 ks = aes4_block(K, kl, lb, hi, lo);
            acc.x ^= ks.x; acc.y ^= ks.y; acc.z ^= ks.z; acc.w ^= ks.w;
            out[i] = acc;
        }
        totals[pair] = acc;
    }
}
// ks[c][o] ^= pre[c / seg][o]   (pre: the exclusive prefix over segments left by xor_scan_prefix), 16 bytes per thread
__global__ void __launch_bounds__(256) xor_prefix_apply(uint4 *__restrict__ ks, const uint4 *__restrict__ pre, size_t n_chunks, int seg) {
    const size_t n = n_chunks * 512;
    for (size_t i = blockIdx.x * (size_t)256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) {
        const size_t c = i >> 9, j = i & 511;
        const uint4 p = pre[(c / seg) * 512 + j];
        uint4 v = ks[i];
        v.x ^= p.x; v.y ^= p.y; v.z ^= p.z; v.w ^= p.w;
        ks[i] = v;
    }
}

constexpr int CH_WORDS = 1024;              // u64 words per 8 KiB chunk (uniform.go:12 bufSize)

// phase 1: XOR total of each segment of `seg` chunks, per word offset
__global__ void xor_scan_totals(u64 *__restrict__ totals, const u64 *__restrict__ ks, size_t n_chunks, int seg) {
    const int o = blockIdx.x * blockDim.x + threadIdx.x;             // word offset
    const size_t s = blockIdx.y;                                      // segment
    if (o >= CH_WORDS) return;
    const size_t c0 = s * seg, c1 = (c0 + seg < n_chunks) ? c0 + seg : n_chunks;
    u64 acc = 0;
    for (size_t c = c0; c < c1; c++) acc ^= ks[c * CH_WORDS + o];
    totals[s * CH_WORDS + o] = acc;
}
// phase 2: exclusive prefix over segments, seeded with the sampler state; also leaves the new state
__global__ void xor_scan_prefix(u64 *__restrict__ totals, u64 *__restrict__ state, size_t n_seg) {
    const int o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= CH_WORDS) return;
    u64 acc = state[o];
    for (size_t s = 0; s < n_seg; s++) { const u64 t = totals[s * CH_WORDS + o]; totals[s * CH_WORDS + o] = acc; acc ^= t; }
    state[o] = acc;
}
// phase 3: running XOR inside each segment (in place: keystream -> handed-out buffer contents)
__global__ void xor_scan_apply(u64 *__restrict__ ks, const u64 *__restrict__ totals, size_t n_chunks, int seg) {
    const int o = blockIdx.x * blockDim.x + threadIdx.x;
    const size_t s = blockIdx.y;
    if (o >= CH_WORDS) return;
    const size_t c0 = s * seg, c1 = (c0 + seg < n_chunks) ? c0 + seg : n_chunks;
    u64 acc = totals[s * CH_WORDS + o];
    for (size_t c = c0; c < c1; c++) { acc ^= ks[c * CH_WORDS + o]; ks[c * CH_WORDS + o] = acc; }
}

// ---- host: AES-256 key schedule ---------------------------------------------------------------
static const uint8_t h_sbox[256] = {
    0x63,0x7c,0x77,0x7b,0xf2,0x6b,0x6f,0xc5,0x30,0x01,0x67,0x2b,0xfe,0xd7,0xab,0x76,0xca,0x82,0xc9,0x7d,0xfa,0x59,0x47,0xf0,0xad,0xd4,0xa2,0xaf,0x9c,0xa4,0x72,0xc0,
    0xb7,0xfd,0x93,0x26,0x36,0x3f,0xf7,0xcc,0x34,0xa5,0xe5,0xf1,0x71,0xd8,0x31,0x15,0x04,0xc7,0x23,0xc3,0x18,0x96,0x05,0x9a,0x07,0x12,0x80,0xe2,0xeb,0x27,0xb2,0x75,
    0x09,0x83,0x2c,0x1a,0x1b,0x6e,0x5a,0xa0,0x52,0x3b,0xd6,0xb3,0x29,0xe3,0x2f,0x84,0x53,0xd1,0x00,0xed,0x20,0xfc,0xb1,0x5b,0x6a,0xcb,0xbe,0x39,0x4a,0x4c,0x58,0xcf,
    0xd0,0xef,0xaa,0xfb,0x43,0x4d,0x33,0x85,0x45,0xf9,0x02,0x7f,0x50,0x3c,0x9f,0xa8,0x51,0xa3,0x40,0x8f,0x92,0x9d,0x38,0xf5,0xbc,0xb6,0xda,0x21,0x10,0xff,0xf3,0xd2,
    0xcd,0x0c,0x13,0xec,0x5f,0x97,0x44,0x17,0xc4,0xa7,0x7e,0x3d,0x64,0x5d,0x19,0x73,0x60,0x81,0x4f,0xdc,0x22,0x2a,0x90,0x88,0x46,0xee,0xb8,0x14,0xde,0x5e,0x0b,0xdb,
    0xe0,0x32,0x3a,0x0a,0x49,0x06,0x24,0x5c,0xc2,0xd3,0xac,0x62,0x91,0x95,0xe4,0x79,0xe7,0xc8,0x37,0x6d,0x8d,0xd5,0x4e,0xa9,0x6c,0x56,0xf4,0xea,0x65,0x7a,0xae,0x08,
    0xba,0x78,0x25,0x2e,0x1c,0xa6,0xb4,0xc6,0xe8,0xdd,0x74,0x1f,0x4b,0xbd,0x8b,0x8a,0x70,0x3e,0xb5,0x66,0x48,0x03,0xf6,0x0e,0x61,0x35,0x57,0xb9,0x86,0xc1,0x1d,0x9e,
    0xe1,0xf8,0x98,0x11,0x69,0xd9,0x8e,0x94,0x9b,0x1e,0x87,0xe9,0xce,0x55,0x28,0xdf,0x8c,0xa1,0x89,0x0d,0xbf,0xe6,0x42,0x68,0x41,0x99,0x2d,0x0f,0xb0,0x54,0xbb,0x16};

static void aes256_expand(const uint8_t key[32], AesKey &K) {
    for (int i = 0; i < 8; i++) K.rk[i] = (uint32_t)key[4 * i] << 24 | (uint32_t)key[4 * i + 1] << 16 | (uint32_t)key[4 * i + 2] << 8 | key[4 * i + 3];
    uint32_t rcon = 1;
    for (int i = 8; i < 60; i++) {
        uint32_t t = K.rk[i - 1];
        if (i % 8 == 0) {
            t = (t << 8) | (t >> 24);
            t = (uint32_t)h_sbox[t >> 24] << 24 | (uint32_t)h_sbox[(t >> 16) & 0xff] << 16 | (uint32_t)h_sbox[(t >> 8) & 0xff] << 8 | h_sbox[t & 0xff];
            t ^= rcon << 24;
            rcon = ((rcon << 1) ^ ((rcon >> 7) * 0x11b)) & 0xff;
        } else if (i % 8 == 4) {
            t = (uint32_t)h_sbox[t >> 24] << 24 | (uint32_t)h_sbox[(t >> 16) & 0xff] << 16 | (uint32_t)h_sbox[(t >> 8) & 0xff] << 8 | h_sbox[t & 0xff];
        }
        K.rk[i] = K.rk[i - 8] ^ t;
    }
}

// shared with cosac_dev.cu
void aes256_expand_key(const uint8_t key[32], uint32_t rk[60]) {
    AesKey K;
    aes256_expand(key, K);
    memcpy(rk, K.rk, sizeof(K.rk));
}
const uint8_t *aes_sbox_host() { return h_sbox; }

struct UniformStream {
    AesKey key;
    u64 iv_hi, iv_lo;
    u64 pos;             // samples handed out so far
    u64 *d_state[2];     // 1024 words each: contents of the sampler's 8 KiB buffer; d_state[cur] is the live one
    int cur;
    cudaStream_t side;   // generation runs here, so it overlaps the caller's uploads / kernels on the main stream
    // two output buffers: the one handed out last may still be read by the consumer's kernels while the other one is
    // being filled for the next draw (uniform_stream_prefetch)
    u64 *d_buf[2];
    size_t buf_words[2];
    cudaEvent_t ready[2], consumed[2];
    int last;            // buffer handed out by the last draw
    bool two;            // alternate between the two buffers (only once uniform_stream_prefetch has been used: the second
                         // buffer costs as much HBM again -- 537 MB at 2^22 -- and its first use a cudaMalloc)
    bool pre_valid;      // d_buf[last ^ 1] / d_state[cur ^ 1] hold the draws [pre_pos, pre_pos + pre_n) and the state after them
    u64 pre_pos; size_t pre_n;
    u64 *d_tot;          // segment totals of the scan
    size_t tot_words;
};


int uniform_stream_init(UniformStream **out, const uint8_t key[32], const uint8_t iv[16]) {
    UniformStream *s = new UniformStream();
    aes256_expand(key, s->key);
    s->iv_hi = s->iv_lo = 0;
    for (int i = 0; i < 8; i++) { s->iv_hi = (s->iv_hi << 8) | iv[i]; s->iv_lo = (s->iv_lo << 8) | iv[8 + i]; }
    s->pos = 0; s->cur = 0; s->last = 0; s->two = false; s->pre_valid = false; s->pre_pos = 0; s->pre_n = 0;
    RS_CUDA_CHECK(cudaStreamCreateWithFlags(&s->side, cudaStreamNonBlocking));
    for (int k = 0; k < 2; k++) {
        RS_CUDA_CHECK(cudaMalloc((void **)&s->d_state[k], CH_WORDS * 8));
        RS_CUDA_CHECK(cudaMemset(s->d_state[k], 0, CH_WORDS * 8));
        RS_CUDA_CHECK(cudaEventCreateWithFlags(&s->ready[k], cudaEventDisableTiming));
        RS_CUDA_CHECK(cudaEventCreateWithFlags(&s->consumed[k], cudaEventDisableTiming));
        s->d_buf[k] = nullptr; s->buf_words[k] = 0;
    }
    s->d_tot = nullptr; s->tot_words = 0;
    // __constant__ memory is per device: upload once on every device a stream is created on
    if (first_use_on_device((const void *)h_sbox)) RS_CUDA_CHECK(cudaMemcpyToSymbol(c_sbox, h_sbox, 256));
    *out = s;
    return RS_OK;
}
void uniform_stream_free(UniformStream *s) {
    if (!s) return;
    cudaStreamSynchronize(s->side);
    for (int k = 0; k < 2; k++) { cudaFree(s->d_state[k]); cudaFree(s->d_buf[k]); cudaEventDestroy(s->ready[k]); cudaEventDestroy(s->consumed[k]); }
    cudaFree(s->d_tot);
    cudaStreamDestroy(s->side);
    delete s;
}

// Enqueues, on the side stream, the generation of the n Sample() values that follow position `pos` into d_buf[b], from the
// sampler state d_state[cur] into d_state[cur ^ 1] (the live state is not touched); records ready[b].
// Layout of the buffer: [ current 8 KiB buffer of the sampler | every chunk refilled by this call ].
static int generate(UniformStream *s, u64 pos, size_t n, int b) {
    cudaStream_t st = s->side;
    const u64 off = pos % CH_WORDS;
    const u64 c_start = pos / CH_WORDS + (off ? 1 : 0);                     // first chunk not yet refilled
    const u64 c_end = (pos + n + CH_WORDS - 1) / CH_WORDS;
    const size_t n_chunks = c_end > c_start ? (size_t)(c_end - c_start) : 0;
    // long draws: the four-table kernel with the scan's first passes folded in; one (segment, block offset) pair per thread of
    // one wave of 148 x 1024 threads, i.e. 296 segments
    static const bool one_table = getenv("RS_AES_T1") != nullptr;           // the one-table kernel + three-pass scan (A/B switch)
    const bool fused = !one_table && n_chunks * (CH_WORDS / 2) >= (size_t)148 * AES4_THREADS;
    const int seg = fused ? (int)((n_chunks + 295) / 296) : (n_chunks > 4096 ? 256 : (n_chunks > 64 ? 16 : 1));
    const size_t n_seg = n_chunks ? (n_chunks + seg - 1) / seg : 0;
    RS_CUDA_CHECK(cudaStreamWaitEvent(st, s->consumed[b], 0));              // the last consumer of this buffer is done
    if ((n_chunks + 1) * CH_WORDS > s->buf_words[b] || n_seg * CH_WORDS > s->tot_words) {
        RS_CUDA_CHECK(cudaStreamSynchronize(st));
        if ((n_chunks + 1) * CH_WORDS > s->buf_words[b]) {
            RS_CUDA_CHECK(cudaEventSynchronize(s->consumed[b]));
            cudaFree(s->d_buf[b]); s->d_buf[b] = nullptr; s->buf_words[b] = 0;
            RS_CUDA_CHECK(cudaMalloc((void **)&s->d_buf[b], (n_chunks + 1) * CH_WORDS * 8));
            s->buf_words[b] = (n_chunks + 1) * CH_WORDS;
        }
        if (n_seg * CH_WORDS > s->tot_words) {
            cudaFree(s->d_tot); s->d_tot = nullptr; s->tot_words = 0;
            RS_CUDA_CHECK(cudaMalloc((void **)&s->d_tot, n_seg * CH_WORDS * 8));
            s->tot_words = n_seg * CH_WORDS;
        }
    }
    u64 *buf = s->d_buf[b];
    const u64 *state_in = s->d_state[s->cur];
    u64 *state_out = s->d_state[s->cur ^ 1];
    RS_CUDA_CHECK(cudaMemcpyAsync(buf, state_in, CH_WORDS * 8, cudaMemcpyDeviceToDevice, st));
    RS_CUDA_CHECK(cudaMemcpyAsync(state_out, state_in, CH_WORDS * 8, cudaMemcpyDeviceToDevice, st));
    if (n_chunks) {
        u64 *ks = buf + CH_WORDS;
        static const int ctas_per_sm = [] { const char *e = getenv("RS_AES_CTAS"); const int v = e ? atoi(e) : 4; return v >= 1 && v <= 6 ? v : 4; }();
        const size_t n_blocks = n_chunks * (CH_WORDS / 2);
        dim3 grid(CH_WORDS / 256, (unsigned)n_seg);
        if (fused) {
            RS_CUDA_CHECK(ensure_smem_attr((const void *)aes_keystream4_scan_kernel, AES4_SMEM));
            aes_keystream4_scan_kernel<<<148, AES4_THREADS, AES4_SMEM, st>>>((uint4 *)ks, (uint4 *)s->d_tot, s->key, s->iv_hi, s->iv_lo,
                                                                            c_start * (CH_WORDS / 2), n_chunks, seg);
            xor_scan_prefix<<<CH_WORDS / 256, 256, 0, st>>>(s->d_tot, state_out, n_seg);
            xor_prefix_apply<<<148 * 8, 256, 0, st>>>((uint4 *)ks, (const uint4 *)s->d_tot, n_chunks, seg);
            count_launch(3);
        } else {
            aes_keystream_kernel<<<148 * ctas_per_sm, 256, 256 * 32 * 4, st>>>((uint4 *)ks, s->key, s->iv_hi, s->iv_lo, c_start * (CH_WORDS / 2), n_blocks);
            xor_scan_totals<<<grid, 256, 0, st>>>(s->d_tot, ks, n_chunks, seg);
            xor_scan_prefix<<<CH_WORDS / 256, 256, 0, st>>>(s->d_tot, state_out, n_seg);
            xor_scan_apply<<<grid, 256, 0, st>>>(ks, s->d_tot, n_chunks, seg);
            count_launch(4);
        }
        RS_CUDA_CHECK(cudaGetLastError());
    }
    RS_CUDA_CHECK(cudaEventRecord(s->ready[b], st));
    return RS_OK;
}

// Produces the next n Sample() values of the stream in HBM and advances it.  The kernels run on the stream's own side
// stream; the current stream is made to wait for them before this returns.  *d_samples points at sample 0 inside one of
// the stream's two reusable output buffers: call uniform_stream_release on the consuming stream once the kernels reading
// it have been launched.  If the draws were generated ahead by uniform_stream_prefetch they are handed out as they are.
int uniform_stream_draw(UniformStream *s, size_t n, const u64 **d_samples) {
    *d_samples = nullptr;
    if (n == 0) return RS_OK;
    const int b = s->two ? s->last ^ 1 : 0;
    if (!(s->pre_valid && s->pre_pos == s->pos && s->pre_n == n)) {
        const int rc = generate(s, s->pos, n, b);                           // (a stale prefetch in d_buf[b] is simply overwritten)
        if (rc) { s->pre_valid = false; return rc; }
    }
    s->pre_valid = false;
    s->cur ^= 1;                                                            // the state written by generate() becomes live
    RS_CUDA_CHECK(cudaStreamWaitEvent(current_stream(), s->ready[b], 0));
    const u64 off = s->pos % CH_WORDS;
    s->pos += n;
    s->last = b;
    *d_samples = s->d_buf[b] + (off ? off : CH_WORDS);
    return RS_OK;
}

void uniform_stream_release(UniformStream *s) { cudaEventRecord(s->consumed[s->last], current_stream()); }

// Generates the NEXT n draws ahead of time into the idle buffer (the stream's position does not move: this is a cache of
// what the next uniform_stream_draw(n) would compute).  Called after the consumer of the current draws has been launched, so
// the AES kernels of Commit k+1 run next to the mat-vec / cut-off kernels of Commit k instead of in front of its encodes.
int uniform_stream_prefetch(UniformStream *s, size_t n) {
    if (n == 0 || (s->pre_valid && s->pre_pos == s->pos && s->pre_n == n)) return RS_OK;
    s->pre_valid = false;
    s->two = true;
    const int rc = generate(s, s->pos, n, s->last ^ 1);
    if (rc) return rc;
    s->pre_valid = true; s->pre_pos = s->pos; s->pre_n = n;
    return RS_OK;
}

}  // namespace rs
