// bgpt_tc.cuh -- prompt-batch matmul on the 5th-generation tensor cores (tcgen05 + TMEM).
//
//   Y[n][r] = sum_b  fl(d_w[r,b] * d_a[n,b]) * (float) isum[r,n,b]   (+ m_w[r,b] * s_a[n,b])
//   isum[r,n,b] = sum_{e<32} q_w[r,b,e] * q_a[n,b,e]                  exact int32
//
// i.e. ggml's block-quantised mul_mat (ggml.c:11804-12013 with the Q8_0/Q8_1 activation
// conversion of ggml.c:11909-11925) for N token rows at once.  The integer dot of one 32-block
// is exactly ONE tcgen05.mma.kind::i8 instruction (K = 32 bytes): 128 weight rows x 64 token
// rows per CTA tile, int32 accumulators in TMEM, one accumulator per block because every block
// has its own scale pair.  The epilogue reads each accumulator back (tcgen05.ld), converts,
// scales and accumulates in f32 registers in block order.
//
// Parity note: the integer sums are exact and the per-block scale product is the reference's
// single rounded f32 product, but the reference adds the blocks through 8 interleaved lanes
// (bgpt_kernels.cuh "lane order") and a tensor-core instruction cannot expose 4-element partial
// sums -- so this path is NOT bit-identical to the CPU reference, only tolerance-close (f32
// summation order).  It is therefore used only for batches of >= BGPT_TC_MIN_ROWS token rows
// (default 32; the reference's default n_batch = 8 stays on the exact path) and can be switched
// off with BGPT_TC=0.
//
// Pipeline per CTA (256 threads), one K step = a group of 4 blocks, two stages:
//   fill     all threads: weights of the group (re-tiled layout, 16-byte LDG) are unpacked to
//            int8 and written in the canonical K-major no-swizzle UMMA layout (8x16B core
//            matrices); the int8 activation words likewise; scales to shared memory
//   mma      one thread: 4 x tcgen05.mma (M=128, N=64, K=32, accumulate=false) + tcgen05.commit
//   epilogue 8 warps = 4 TMEM lane quadrants x 2 column halves: tcgen05.ld 32x32b.x32,
//            acc[c] = fma(s_w*s_a, (float) d, acc[c])
// fill(g+1) and the MMAs of g+1 are issued before the epilogue of g, so the tensor core works
// while the CUDA cores drain the previous group.
#pragma once
#include "bgpt_kernels.cuh"

#define TC_ROWS 128
#define TC_TOK  64
#define TC_THREADS 256

__device__ __forceinline__ uint32_t tc_smem_u32(const void * p) { return (uint32_t) __cvta_generic_to_shared(p); }

// K-major, no swizzle: element (row, 16-byte chunk kc) of a [rows x 32 B] operand lives at
// (row/8)*256 + kc*128 + (row%8)*16  -> LBO (between K chunks) = 128 B, SBO (between 8-row groups) = 256 B
__device__ __forceinline__ uint64_t tc_make_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t) ((smem_addr >> 4) & 0x3FFF);          // start address
    d |= (uint64_t) ((128u >> 4) & 0x3FFF) << 16;         // leading byte offset
    d |= (uint64_t) ((256u >> 4) & 0x3FFF) << 32;         // stride byte offset
    d |= (uint64_t) 1 << 46;                              // descriptor version (Blackwell)
    return d;                                             // base offset 0, SWIZZLE_NONE
}
// instruction descriptor: D = S32, A = B = S8, both K-major, M = 128, N = TC_TOK
__device__ __forceinline__ uint32_t tc_idesc_i8() {
    return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t) (TC_TOK >> 3) << 17) | ((uint32_t) (TC_ROWS >> 4) << 24);
}
__device__ __forceinline__ void tc_mma_i8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}\n"
        :: "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}
__device__ __forceinline__ void tc_commit(uint32_t mbar_addr) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(mbar_addr) : "memory");
}
__device__ __forceinline__ void tc_mbar_init(uint32_t mbar_addr, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(mbar_addr), "r"(count) : "memory");
}
__device__ __forceinline__ void tc_mbar_wait(uint32_t mbar_addr, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}\n"
        :: "r"(mbar_addr), "r"(parity) : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                   "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                   "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tc_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// exact int32 -> f32 for |v| < 2^22 without the quarter-rate I2F: 1.5*2^23 + v is exact in f32
__device__ __forceinline__ float tc_i2f(uint32_t v) { return __fsub_rn(__int_as_float((int) v + 0x4B400000), 12582912.0f); }

// 4 nibble-codes (one byte each, value 0..15 / 0..31) -> signed int8 bytes of (code - off)
__device__ __forceinline__ uint32_t tc_sub8(uint32_t v)  { const uint32_t t = v ^ 0x08080808u; return t | ((t & 0x08080808u) * 0x1Eu); }
__device__ __forceinline__ uint32_t tc_sub16(uint32_t v) { const uint32_t t = v ^ 0x10101010u; return t | ((t & 0x10101010u) * 0x0Eu); }

struct TcShared {
    uint8_t A[2][4][TC_ROWS * 32];     // [stage][block in group][canonical layout]
    uint8_t B[2][4][TC_TOK * 32];
    float sw[2][4][TC_ROWS];           // weight scales d_w
    float mw[2][4][TC_ROWS];           // weight mins   m_w   (Q4_1 / Q5_1)
    float sa[2][4][TC_TOK];            // activation scales d_a
    float ss[2][4][TC_TOK];            // activation s = d*sum(q) (Q8_1)
    unsigned long long mbar[2];
    uint32_t tmem_base;
};

template <int FMT>
__global__ void __launch_bounds__(TC_THREADS, 1) k_gemm_tc_q(GemvArgs a) {
    extern __shared__ __align__(128) uint8_t tc_smem_raw[];
    TcShared & S = *reinterpret_cast<TcShared *>(tc_smem_raw);
    constexpr bool IS8   = (FMT == BG_Q8_0);
    constexpr bool HASQH = (FMT == BG_Q5_0 || FMT == BG_Q5_1);
    constexpr bool HASM  = (FMT == BG_Q4_1 || FMT == BG_Q5_1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int row0 = blockIdx.x * TC_ROWS;                 // first weight row of this tile (stacked row space)
    const int tok0 = a.tok0 + blockIdx.y * TC_TOK;         // first token row

    if (tid == 0) { tc_mbar_init(tc_smem_u32(&S.mbar[0]), 1); tc_mbar_init(tc_smem_u32(&S.mbar[1]), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(tc_smem_u32(&S.tmem_base)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = S.tmem_base;
    const uint32_t idesc = tc_idesc_i8();

    // ---- fill, split in two so that the global loads of group g+1 travel while group g-1 is
    //      drained: load_group (global -> registers), store_group (unpack, registers -> smem)
    const int frow = tid >> 1, fkc = tid & 1;               // weights: thread (row, kc): kc = 0 -> elements 0..15, 1 -> 16..31
    const uint8_t * wrow;
    {
        int r = row0 + frow; r = r < a.M ? r : a.M - 1;
        const int mat = r / a.rows_per;
        wrow = a.W[mat] + (size_t) (r - mat * a.rows_per) * a.stride;
    }
    const int atk = tid >> 1, akc = tid & 1;                // activations: thread (token, kc), tid < 2*TC_TOK
    const bool a_thread = tid < 2 * TC_TOK;
    const bool a_valid = a_thread && (tok0 + atk) < a.n;
    const uint8_t * arec = a.act + (size_t) (a_valid ? tok0 + atk : 0) * a.act_bytes;
    uint4 wreg[4], areg[4]; uint32_t qhreg[4]; uint2 sreg; float4 asreg;
    auto load_group = [&](int g) {
#pragma unroll
        for (int j = 0; j < 4; j++) {
            wreg[j] = IS8 ? ldg_stream128(wrow + (size_t) ((g * 2 + fkc) * 4 + j) * 16) : ldg_stream128(wrow + (size_t) (g * 4 + j) * 16);
            if (HASQH) qhreg[j] = ldg_stream32(wrow + a.off_qh + g * 16 + j * 4);
        }
        sreg = make_uint2(0, 0);
        if (fkc == 0) sreg = ldg_stream64(wrow + a.off_d + g * 8);
        else if (HASM) sreg = ldg_stream64(wrow + a.off_m + g * 8);
        asreg = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int j = 0; j < 4; j++) areg[j] = make_uint4(0, 0, 0, 0);
        if (a_valid) {
#pragma unroll
            for (int j = 0; j < 4; j++) areg[j] = *(const uint4 *) (arec + (size_t) (g * 8 + akc * 4 + j) * 16);
            if (akc == 0) asreg = *(const float4 *) (arec + a.off_dd + g * 16);
            else if (HASM) asreg = *(const float4 *) (arec + a.off_s + g * 16);
        }
    };
    auto store_group = [&](int s) {
        {
            uint32_t out[4][4];                              // [block i][word j]
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const uint32_t ww[4] = { wreg[j].x, wreg[j].y, wreg[j].z, wreg[j].w };
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    uint32_t v = ww[i];
                    if (!IS8) {
                        v = fkc ? ((ww[i] >> 4) & 0x0F0F0F0Fu) : (ww[i] & 0x0F0F0F0Fu);
                        if (HASQH) { const uint32_t hb = (qhreg[j] >> (8 * i)) & 0xFFu; v |= bg_spread4(fkc ? (hb >> 4) : (hb & 0xFu)); }
                        if (FMT == BG_Q4_0) v = tc_sub8(v);
                        if (FMT == BG_Q5_0) v = tc_sub16(v);
                    }
                    out[i][j] = v;
                }
            }
            const int off = (frow >> 3) * 256 + fkc * 128 + (frow & 7) * 16;
#pragma unroll
            for (int i = 0; i < 4; i++) *(uint4 *) (&S.A[s][i][off]) = make_uint4(out[i][0], out[i][1], out[i][2], out[i][3]);
            float * dst = (fkc == 0) ? &S.sw[s][0][0] : &S.mw[s][0][0];
            if (fkc == 0 || HASM) {
                dst[0 * TC_ROWS + frow] = bg_h2f((uint16_t) (sreg.x & 0xFFFF)); dst[1 * TC_ROWS + frow] = bg_h2f((uint16_t) (sreg.x >> 16));
                dst[2 * TC_ROWS + frow] = bg_h2f((uint16_t) (sreg.y & 0xFFFF)); dst[3 * TC_ROWS + frow] = bg_h2f((uint16_t) (sreg.y >> 16));
            }
        }
        if (a_thread) {
            const int off = (atk >> 3) * 256 + akc * 128 + (atk & 7) * 16;
            *(uint4 *) (&S.B[s][0][off]) = make_uint4(areg[0].x, areg[1].x, areg[2].x, areg[3].x);
            *(uint4 *) (&S.B[s][1][off]) = make_uint4(areg[0].y, areg[1].y, areg[2].y, areg[3].y);
            *(uint4 *) (&S.B[s][2][off]) = make_uint4(areg[0].z, areg[1].z, areg[2].z, areg[3].z);
            *(uint4 *) (&S.B[s][3][off]) = make_uint4(areg[0].w, areg[1].w, areg[2].w, areg[3].w);
            float * dst = (akc == 0) ? &S.sa[s][0][0] : &S.ss[s][0][0];
            if (akc == 0 || HASM) { dst[0 * TC_TOK + atk] = asreg.x; dst[1 * TC_TOK + atk] = asreg.y; dst[2 * TC_TOK + atk] = asreg.z; dst[3 * TC_TOK + atk] = asreg.w; }
        }
    };
    auto issue = [&](int s) {
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int i = 0; i < 4; i++)
                tc_mma_i8(tmem + (uint32_t) ((s * 4 + i) * TC_TOK), tc_make_desc(tc_smem_u32(&S.A[s][i][0])), tc_make_desc(tc_smem_u32(&S.B[s][i][0])), idesc, 0u);
            tc_commit(tc_smem_u32(&S.mbar[s]));
        }
    };

    // ---- epilogue state: this thread owns row (quadrant*32 + lane) and 32 of the 64 token columns
    const int quad = warp & 3, chalf = warp >> 2;
    const int erow = quad * 32 + lane;
    float acc[32], summ[32];
#pragma unroll
    for (int c = 0; c < 32; c++) { acc[c] = 0.0f; summ[c] = 0.0f; }
    auto epilogue = [&](int s, uint32_t parity) {
        tc_mbar_wait(tc_smem_u32(&S.mbar[s]), parity);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        uint32_t v[2][32];
        const uint32_t tbase = tmem + ((uint32_t) (quad * 32) << 16) + (uint32_t) (s * 4 * TC_TOK + chalf * 32);
        tc_ld32(tbase, v[0]);
#pragma unroll
        for (int i = 0; i < 4; i++) {
            tc_ld_wait();                                        // accumulator i is in v[i&1]
            if (i < 3) tc_ld32(tbase + (uint32_t) ((i + 1) * TC_TOK), v[(i + 1) & 1]);   // next one travels during the math
            const float dw = S.sw[s][i][erow];
            const float mwv = HASM ? S.mw[s][i][erow] : 0.0f;
#pragma unroll
            for (int c = 0; c < 32; c++) {
                const float sc = __fmul_rn(dw, S.sa[s][i][chalf * 32 + c]);
                acc[c] = fmaf(sc, tc_i2f(v[i & 1][c]), acc[c]);
                if (HASM) summ[c] = fmaf(mwv, S.ss[s][i][chalf * 32 + c], summ[c]);
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    };

    // ---- main loop over groups of 4 blocks
    const int G = a.G;
    load_group(0);
    store_group(0);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    issue(0);
    if (G > 1) load_group(1);
    for (int g = 1; g < G; g++) {
        const int s = g & 1;
        store_group(s);                                    // stage s was drained by epilogue(g-2)
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        issue(s);
        if (g + 1 < G) load_group(g + 1);                  // in flight during the epilogue below
        epilogue(s ^ 1, (uint32_t) (((g - 1) >> 1) & 1));
        __syncthreads();                                   // scales of stage s^1 are free again
    }
    epilogue((G - 1) & 1, (uint32_t) (((G - 1) >> 1) & 1));

    // ---- write out
    const int r = row0 + erow;
    if (r < a.M) {
#pragma unroll
        for (int c = 0; c < 32; c++) {
            const int n = tok0 + chalf * 32 + c;
            if (n < a.n) bg_epilogue(a.epi, n, r, HASM ? __fadd_rn(acc[c], summ[c]) : acc[c]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512u) : "memory");
}

// ================================================================================================================================
// k_gemm_tc_x -- the same tensor-core matmul, BIT-EXACT: the reference's 8 running sums per row through tcgen05
// ================================================================================================================================
// The reference (AVX2) adds a block's 32 products in eight groups of four: isum_l = sum of elements 4l..4l+3, then
// acc_l = fma(d_w * d_a, (float) isum_l, acc_l) per group l in block order, hsum_float_8 at the end (ggml.c:2518-2541 ...).  One
// tcgen05.mma sums all of K = 32, so k_gemm_tc_q above can only deliver one term per block and drifts from the reference by ~5e-2
// in the logits once activations are re-quantised 24 layers deep.  Here the ACTIVATION operand is expanded instead: token n
// becomes eight columns (n, l), column (n, l) carrying the token's int8 codes in bytes 4l..4l+3 of the block and zeros elsewhere.
// The MMA then produces D[row][(n, l)] = isum_l exactly (int32), i.e. all eight partial sums of 128 rows x 16 tokens per
// instruction (M = 128, N = 128, K = 32), and the epilogue runs the reference's eight fma chains per (row, token) in block
// order.  The tensor pipe does 8x redundant multiplications by zero; it has the room (the step is bound by the f32 epilogue:
// 16 f32-pipe instructions per (row, token, block) -- the arithmetic the bit-exact contract prescribes).
//
// Tile: 128 weight rows x 16 tokens; a group of 4 blocks fills the 512 TMEM columns (4 x 128), single stage: the four MMAs of a
// group take ~0.25 us against ~1.1 us of epilogue, so nothing is gained by splitting TMEM in two.  The zero pattern of the
// activation operand is written once; a group only rewrites the 4-byte words that carry data.
#define TCX_TOK 16
#define TCX_N (TCX_TOK * 8)

struct TcxShared {
    uint8_t A[4][TC_ROWS * 32];        // [block in group][canonical K-major layout]
    uint8_t B[4][TCX_N * 32];          // [block in group][column (n, l)][32 bytes]
    float sw[4][TC_ROWS];              // weight scales d_w
    float mw[4][TC_ROWS];              // weight mins   m_w   (Q4_1 / Q5_1)
    float sa[4][TCX_TOK];              // activation scales d_a
    float ss[4][TCX_TOK];              // activation s = d * sum(q) (Q8_1)
    unsigned long long mbar;
    uint32_t tmem_base;
};
__device__ __forceinline__ uint32_t tcx_idesc_i8() {          // D = S32, A = B = S8, both K-major, M = 128, N = 128
    return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t) (TCX_N >> 3) << 17) | ((uint32_t) (TC_ROWS >> 4) << 24);
}

template <int FMT>
__global__ void __launch_bounds__(TC_THREADS, 1) k_gemm_tc_x(GemvArgs a) {
    extern __shared__ __align__(128) uint8_t tc_smem_raw[];
    TcxShared & S = *reinterpret_cast<TcxShared *>(tc_smem_raw);
    constexpr bool IS8   = (FMT == BG_Q8_0);
    constexpr bool HASQH = (FMT == BG_Q5_0 || FMT == BG_Q5_1);
    constexpr bool HASM  = (FMT == BG_Q4_1 || FMT == BG_Q5_1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int row0 = blockIdx.x * TC_ROWS;                 // first weight row of this tile (stacked row space)
    const int tok0 = a.tok0 + blockIdx.y * TCX_TOK;        // first token row

    if (tid == 0) { tc_mbar_init(tc_smem_u32(&S.mbar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(tc_smem_u32(&S.tmem_base)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // the zero pattern of the expanded activation operand, once
    for (int i = tid; i < (int) (sizeof(S.B) / 16); i += TC_THREADS) ((uint4 *) &S.B[0][0])[i] = make_uint4(0, 0, 0, 0);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = S.tmem_base;
    const uint32_t idesc = tcx_idesc_i8();

    // ---- fill: load_group (global -> registers), store_group (unpack, registers -> shared memory)
    const int frow = tid >> 1, fkc = tid & 1;               // weights: thread (row, kc): kc = 0 -> elements 0..15, 1 -> 16..31
    const uint8_t * wrow;
    {
        int r = row0 + frow; r = r < a.M ? r : a.M - 1;
        const int mat = r / a.rows_per;
        wrow = a.W[mat] + (size_t) (r - mat * a.rows_per) * a.stride;
    }
    const int atk = tid >> 3, al = tid & 7;                 // activations: thread (token, element group l), tid < 8 * TCX_TOK
    const bool a_thread = tid < 8 * TCX_TOK;
    const bool a_valid = a_thread && (tok0 + atk) < a.n;
    const uint8_t * arec = a.act + (size_t) (a_valid ? tok0 + atk : 0) * a.act_bytes;
    uint4 wreg[4], areg; uint32_t qhreg[4]; uint2 sreg; float4 asreg;
    auto load_group = [&](int g) {
#pragma unroll
        for (int j = 0; j < 4; j++) {
            wreg[j] = IS8 ? ldg_stream128(wrow + (size_t) ((g * 2 + fkc) * 4 + j) * 16) : ldg_stream128(wrow + (size_t) (g * 4 + j) * 16);
            if (HASQH) qhreg[j] = ldg_stream32(wrow + a.off_qh + g * 16 + j * 4);
        }
        sreg = make_uint2(0, 0);
        if (fkc == 0) sreg = ldg_stream64(wrow + a.off_d + g * 8);
        else if (HASM) sreg = ldg_stream64(wrow + a.off_m + g * 8);
        asreg = make_float4(0.f, 0.f, 0.f, 0.f);
        areg = make_uint4(0, 0, 0, 0);
        if (a_valid) {
            areg = *(const uint4 *) (arec + (size_t) (g * 8 + al) * 16);          // the codes of element group l for the 4 blocks
            if (al == 0) asreg = *(const float4 *) (arec + a.off_dd + g * 16);
            else if (HASM && al == 1) asreg = *(const float4 *) (arec + a.off_s + g * 16);
        }
    };
    auto store_group = [&]() {
        {
            uint32_t out[4][4];                              // [block i][word j]
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const uint32_t ww[4] = { wreg[j].x, wreg[j].y, wreg[j].z, wreg[j].w };
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    uint32_t v = ww[i];
                    if (!IS8) {
                        v = fkc ? ((ww[i] >> 4) & 0x0F0F0F0Fu) : (ww[i] & 0x0F0F0F0Fu);
                        if (HASQH) { const uint32_t hb = (qhreg[j] >> (8 * i)) & 0xFFu; v |= bg_spread4(fkc ? (hb >> 4) : (hb & 0xFu)); }
                        if (FMT == BG_Q4_0) v = tc_sub8(v);
                        if (FMT == BG_Q5_0) v = tc_sub16(v);
                    }
                    out[i][j] = v;
                }
            }
            const int off = (frow >> 3) * 256 + fkc * 128 + (frow & 7) * 16;
#pragma unroll
            for (int i = 0; i < 4; i++) *(uint4 *) (&S.A[i][off]) = make_uint4(out[i][0], out[i][1], out[i][2], out[i][3]);
            float * dst = (fkc == 0) ? &S.sw[0][0] : &S.mw[0][0];
            if (fkc == 0 || HASM) {
                dst[0 * TC_ROWS + frow] = bg_h2f((uint16_t) (sreg.x & 0xFFFF)); dst[1 * TC_ROWS + frow] = bg_h2f((uint16_t) (sreg.x >> 16));
                dst[2 * TC_ROWS + frow] = bg_h2f((uint16_t) (sreg.y & 0xFFFF)); dst[3 * TC_ROWS + frow] = bg_h2f((uint16_t) (sreg.y >> 16));
            }
        }
        if (a_thread) {
            // column (n, l) = operand row n * 8 + l: 16-byte chunk l >> 2 of its 32 K bytes, word l & 3 -- everything else stays zero
            const int off = atk * 256 + (al >> 2) * 128 + al * 16 + (al & 3) * 4;
            *(uint32_t *) (&S.B[0][off]) = areg.x; *(uint32_t *) (&S.B[1][off]) = areg.y;
            *(uint32_t *) (&S.B[2][off]) = areg.z; *(uint32_t *) (&S.B[3][off]) = areg.w;
            if (al == 0) { S.sa[0][atk] = asreg.x; S.sa[1][atk] = asreg.y; S.sa[2][atk] = asreg.z; S.sa[3][atk] = asreg.w; }
            else if (HASM && al == 1) { S.ss[0][atk] = asreg.x; S.ss[1][atk] = asreg.y; S.ss[2][atk] = asreg.z; S.ss[3][atk] = asreg.w; }
        }
    };

    // ---- epilogue state: this thread owns row (quadrant * 32 + lane) and 8 of the 16 tokens: 8 running sums each
    const int quad = warp & 3, chalf = warp >> 2;
    const int erow = quad * 32 + lane;
    float acc[8][8], summ[8];
#pragma unroll
    for (int t = 0; t < 8; t++) { summ[t] = 0.0f;
#pragma unroll
        for (int l = 0; l < 8; l++) acc[t][l] = 0.0f; }

    const int G = a.G;
    load_group(0);
    for (int g = 0; g < G; g++) {
        store_group();                                     // the previous group's operands and accumulators were drained
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int i = 0; i < 4; i++)
                tc_mma_i8(tmem + (uint32_t) (i * TCX_N), tc_make_desc(tc_smem_u32(&S.A[i][0])), tc_make_desc(tc_smem_u32(&S.B[i][0])), idesc, 0u);
            tc_commit(tc_smem_u32(&S.mbar));
        }
        if (g + 1 < G) load_group(g + 1);                  // in flight during the MMAs and the epilogue
        tc_mbar_wait(tc_smem_u32(&S.mbar), (uint32_t) (g & 1));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // the reference's chains: for every block in order, acc_l = fma(d_w * d_a, (float) isum_l, acc_l); summs = fma(m_w, s_a, summs)
        uint32_t v[2][32];
        const uint32_t tbase = tmem + ((uint32_t) (quad * 32) << 16) + (uint32_t) (chalf * 64);
        tc_ld32(tbase, v[0]);
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const float dw = S.sw[i][erow];
            const float mwv = HASM ? S.mw[i][erow] : 0.0f;
#pragma unroll
            for (int h = 0; h < 2; h++) {                   // 32 columns = 4 tokens x 8 sums per load
                const int k = i * 2 + h;
                tc_ld_wait();
                if (k < 7) tc_ld32(tmem + ((uint32_t) (quad * 32) << 16) + (uint32_t) (((k + 1) >> 1) * TCX_N + chalf * 64 + ((k + 1) & 1) * 32), v[(k + 1) & 1]);
#pragma unroll
                for (int t4 = 0; t4 < 4; t4++) {
                    const int t = h * 4 + t4;
                    const float sc = __fmul_rn(dw, S.sa[i][chalf * 8 + t]);
#pragma unroll
                    for (int l = 0; l < 8; l++) acc[t][l] = fmaf(sc, tc_i2f(v[k & 1][t4 * 8 + l]), acc[t][l]);
                    if (HASM) summ[t] = fmaf(mwv, S.ss[i][chalf * 8 + t], summ[t]);
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();                                   // operands, scales and TMEM are free again
    }

    // ---- hsum_float_8 (+ summs) and the fused epilogue of the matmul
    const int r = row0 + erow;
    if (r < a.M) {
#pragma unroll
        for (int t = 0; t < 8; t++) {
            const int n = tok0 + chalf * 8 + t;
            float x = __fadd_rn(__fadd_rn(__fadd_rn(acc[t][0], acc[t][4]), __fadd_rn(acc[t][2], acc[t][6])),
                                __fadd_rn(__fadd_rn(acc[t][1], acc[t][5]), __fadd_rn(acc[t][3], acc[t][7])));
            if (HASM) x = __fadd_rn(x, summ[t]);
            if (n < a.n) bg_epilogue(a.epi, n, r, x);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512u) : "memory");
}

// ================================================================================================================================
// k_gemm_tc_xf -- the bit-exact matmul on kind::f16: the tensor core hands the partial sums over as f32
// ================================================================================================================================
// k_gemm_tc_x above spends 16 of its ~27 f32-pipe instructions per (row, token, block) turning int32 partial sums into floats.
// The 4-element partial sums are small integers (|isum_l| <= 4 * 127 * 128 < 2^17): computed from fp16 operands with f32
// accumulation they are exact, so the same masked-column scheme on tcgen05.mma.kind::f16 delivers (float) isum_l directly and the
// epilogue is the reference's arithmetic and nothing else (8 fma + 1 mul per (row, token, block)).  K = 16 per instruction: a
// block is two MMAs (elements 0..15 -> sums 0..3, elements 16..31 -> sums 4..7), each over 16 tokens x 4 masked columns (N = 64),
// which also halves the redundant multiplications by zero.
struct TcxfShared {
    uint8_t A[4][2][TC_ROWS * 32];     // [block in group][half of the block][128 rows x 16 fp16, canonical K-major layout]
    uint8_t B[4][2][TCX_TOK * 4 * 32]; // [block][half][column (n, l')][16 fp16]
    float sw[4][TC_ROWS];
    float mw[4][TC_ROWS];
    float sa[4][TCX_TOK];
    float ss[4][TCX_TOK];
    unsigned long long mbar;
    uint32_t tmem_base;
};
__device__ __forceinline__ uint32_t tcxf_idesc() {             // D = F32, A = B = F16, both K-major, M = 128, N = 64
    return (1u << 4) | ((uint32_t) ((TCX_TOK * 4) >> 3) << 17) | ((uint32_t) (TC_ROWS >> 4) << 24);
}
__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}\n"
        :: "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}
// 4 signed bytes -> 4 fp16 values (exact): 0x6400 | (b ^ 0x80) is the fp16 number 1024 + (b + 128)
__device__ __forceinline__ uint2 tc_s8x4_to_h4(uint32_t w) {
    const uint32_t x = w ^ 0x80808080u;
    uint32_t lo = __byte_perm(x, 0x64646464u, 0x5140), hi = __byte_perm(x, 0x64646464u, 0x5342);
    const __half2 off = __floats2half2_rn(1152.0f, 1152.0f);
    __half2 l2 = __hsub2(*reinterpret_cast<__half2 *>(&lo), off), h2 = __hsub2(*reinterpret_cast<__half2 *>(&hi), off);
    return make_uint2(*reinterpret_cast<uint32_t *>(&l2), *reinterpret_cast<uint32_t *>(&h2));
}

template <int FMT>
__global__ void __launch_bounds__(TC_THREADS, 1) k_gemm_tc_xf(GemvArgs a) {
    extern __shared__ __align__(128) uint8_t tc_smem_raw[];
    TcxfShared & S = *reinterpret_cast<TcxfShared *>(tc_smem_raw);
    constexpr bool IS8   = (FMT == BG_Q8_0);
    constexpr bool HASQH = (FMT == BG_Q5_0 || FMT == BG_Q5_1);
    constexpr bool HASM  = (FMT == BG_Q4_1 || FMT == BG_Q5_1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int row0 = blockIdx.x * TC_ROWS;
    const int tok0 = a.tok0 + blockIdx.y * TCX_TOK;

    if (tid == 0) { tc_mbar_init(tc_smem_u32(&S.mbar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(tc_smem_u32(&S.tmem_base)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = tid; i < (int) (sizeof(S.B) / 16); i += TC_THREADS) ((uint4 *) &S.B[0][0][0])[i] = make_uint4(0, 0, 0, 0);   // the zero pattern, once
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = S.tmem_base;
    const uint32_t idesc = tcxf_idesc();

    const int frow = tid >> 1, fkc = tid & 1;               // weights: thread (row, half of the block)
    const uint8_t * wrow;
    {
        int r = row0 + frow; r = r < a.M ? r : a.M - 1;
        const int mat = r / a.rows_per;
        wrow = a.W[mat] + (size_t) (r - mat * a.rows_per) * a.stride;
    }
    const int atk = tid >> 3, al = tid & 7;                 // activations: thread (token, element group l)
    const bool a_thread = tid < 8 * TCX_TOK;
    const bool a_valid = a_thread && (tok0 + atk) < a.n;
    const uint8_t * arec = a.act + (size_t) (a_valid ? tok0 + atk : 0) * a.act_bytes;
    uint4 wreg[4], areg; uint32_t qhreg[4]; uint2 sreg; float4 asreg;
    auto load_group = [&](int g) {
#pragma unroll
        for (int j = 0; j < 4; j++) {
            wreg[j] = IS8 ? ldg_stream128(wrow + (size_t) ((g * 2 + fkc) * 4 + j) * 16) : ldg_stream128(wrow + (size_t) (g * 4 + j) * 16);
            if (HASQH) qhreg[j] = ldg_stream32(wrow + a.off_qh + g * 16 + j * 4);
        }
        sreg = make_uint2(0, 0);
        if (fkc == 0) sreg = ldg_stream64(wrow + a.off_d + g * 8);
        else if (HASM) sreg = ldg_stream64(wrow + a.off_m + g * 8);
        asreg = make_float4(0.f, 0.f, 0.f, 0.f);
        areg = make_uint4(0, 0, 0, 0);
        if (a_valid) {
            areg = *(const uint4 *) (arec + (size_t) (g * 8 + al) * 16);
            if (al == 0) asreg = *(const float4 *) (arec + a.off_dd + g * 16);
            else if (HASM && al == 1) asreg = *(const float4 *) (arec + a.off_s + g * 16);
        }
    };
    auto store_group = [&]() {
        {
            uint32_t out[4][4];                              // [block i][word j]: signed int8 codes of elements 16 fkc + 4j .. + 3
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const uint32_t ww[4] = { wreg[j].x, wreg[j].y, wreg[j].z, wreg[j].w };
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    uint32_t v = ww[i];
                    if (!IS8) {
                        v = fkc ? ((ww[i] >> 4) & 0x0F0F0F0Fu) : (ww[i] & 0x0F0F0F0Fu);
                        if (HASQH) { const uint32_t hb = (qhreg[j] >> (8 * i)) & 0xFFu; v |= bg_spread4(fkc ? (hb >> 4) : (hb & 0xFu)); }
                        if (FMT == BG_Q4_0) v = tc_sub8(v);
                        if (FMT == BG_Q5_0) v = tc_sub16(v);
                    }
                    out[i][j] = v;
                }
            }
            const int off = (frow >> 3) * 256 + (frow & 7) * 16;     // + 128 for the second 8 elements of the 16
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const uint2 h0 = tc_s8x4_to_h4(out[i][0]), h1 = tc_s8x4_to_h4(out[i][1]), h2 = tc_s8x4_to_h4(out[i][2]), h3 = tc_s8x4_to_h4(out[i][3]);
                *(uint4 *) (&S.A[i][fkc][off])       = make_uint4(h0.x, h0.y, h1.x, h1.y);
                *(uint4 *) (&S.A[i][fkc][off + 128]) = make_uint4(h2.x, h2.y, h3.x, h3.y);
            }
            float * dst = (fkc == 0) ? &S.sw[0][0] : &S.mw[0][0];
            if (fkc == 0 || HASM) {
                dst[0 * TC_ROWS + frow] = bg_h2f((uint16_t) (sreg.x & 0xFFFF)); dst[1 * TC_ROWS + frow] = bg_h2f((uint16_t) (sreg.x >> 16));
                dst[2 * TC_ROWS + frow] = bg_h2f((uint16_t) (sreg.y & 0xFFFF)); dst[3 * TC_ROWS + frow] = bg_h2f((uint16_t) (sreg.y >> 16));
            }
        }
        if (a_thread) {
            // column (n, l') of half c = l >> 2, l' = l & 3: operand row n * 4 + l'; its four live fp16 values are elements 4 l' .. 4 l' + 3
            const int c = al >> 2, lp = al & 3, nr = atk * 4 + lp;
            const int off = (nr >> 3) * 256 + (lp >> 1) * 128 + (nr & 7) * 16 + (lp & 1) * 8;
            *(uint2 *) (&S.B[0][c][off]) = tc_s8x4_to_h4(areg.x); *(uint2 *) (&S.B[1][c][off]) = tc_s8x4_to_h4(areg.y);
            *(uint2 *) (&S.B[2][c][off]) = tc_s8x4_to_h4(areg.z); *(uint2 *) (&S.B[3][c][off]) = tc_s8x4_to_h4(areg.w);
            if (al == 0) { S.sa[0][atk] = asreg.x; S.sa[1][atk] = asreg.y; S.sa[2][atk] = asreg.z; S.sa[3][atk] = asreg.w; }
            else if (HASM && al == 1) { S.ss[0][atk] = asreg.x; S.ss[1][atk] = asreg.y; S.ss[2][atk] = asreg.z; S.ss[3][atk] = asreg.w; }
        }
    };

    const int quad = warp & 3, chalf = warp >> 2;           // rows quad * 32 + lane; tokens chalf * 8 .. + 7
    const int erow = quad * 32 + lane;
    float acc[8][8], summ[8];
#pragma unroll
    for (int t = 0; t < 8; t++) { summ[t] = 0.0f;
#pragma unroll
        for (int l = 0; l < 8; l++) acc[t][l] = 0.0f; }

    const int G = a.G;
    load_group(0);
    for (int g = 0; g < G; g++) {
        store_group();
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int c = 0; c < 2; c++)
                    tc_mma_f16(tmem + (uint32_t) (i * TCX_N + c * 64), tc_make_desc(tc_smem_u32(&S.A[i][c][0])), tc_make_desc(tc_smem_u32(&S.B[i][c][0])), idesc, 0u);
            tc_commit(tc_smem_u32(&S.mbar));
        }
        if (g + 1 < G) load_group(g + 1);
        tc_mbar_wait(tc_smem_u32(&S.mbar), (uint32_t) (g & 1));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // TMEM columns of block i: [i * 128 + c * 64 + n * 4 + l'] = (float) isum_{4c + l'} of token n; this thread reads n = chalf * 8 .. + 7
        uint32_t v[2][32];
        const uint32_t lanebase = tmem + ((uint32_t) (quad * 32) << 16) + (uint32_t) (chalf * 32);
        tc_ld32(lanebase, v[0]);
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const float dw = S.sw[i][erow];
            const float mwv = HASM ? S.mw[i][erow] : 0.0f;
            float sc[8];
#pragma unroll
            for (int t = 0; t < 8; t++) sc[t] = __fmul_rn(dw, S.sa[i][chalf * 8 + t]);
#pragma unroll
            for (int c = 0; c < 2; c++) {
                const int k = i * 2 + c;
                tc_ld_wait();
                if (k < 7) tc_ld32(lanebase + (uint32_t) (((k + 1) >> 1) * TCX_N + ((k + 1) & 1) * 64), v[(k + 1) & 1]);
#pragma unroll
                for (int t = 0; t < 8; t++)
#pragma unroll
                    for (int lp = 0; lp < 4; lp++) acc[t][c * 4 + lp] = fmaf(sc[t], __uint_as_float(v[k & 1][t * 4 + lp]), acc[t][c * 4 + lp]);
            }
            if (HASM) {
#pragma unroll
                for (int t = 0; t < 8; t++) summ[t] = fmaf(mwv, S.ss[i][chalf * 8 + t], summ[t]);
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
    }

    const int r = row0 + erow;
    if (r < a.M) {
#pragma unroll
        for (int t = 0; t < 8; t++) {
            const int n = tok0 + chalf * 8 + t;
            float x = __fadd_rn(__fadd_rn(__fadd_rn(acc[t][0], acc[t][4]), __fadd_rn(acc[t][2], acc[t][6])),
                                __fadd_rn(__fadd_rn(acc[t][1], acc[t][5]), __fadd_rn(acc[t][3], acc[t][7])));
            if (HASM) x = __fadd_rn(x, summ[t]);
            if (n < a.n) bg_epilogue(a.epi, n, r, x);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512u) : "memory");
}
