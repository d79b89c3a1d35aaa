// bgpt_tcw.cuh -- warp-specialised, TMA-fed tcgen05 matmuls for prompt batches (the reference's mul_mat, ggml.c:11804-12013, for
// many token rows at once).  Two kernels share one skeleton:
//
//   warp 0   TMA producer   one thread: cp.async.bulk.tensor (128-byte-swizzled operand tiles) + cp.async.bulk (scale planes)
//                           into a ring of shared-memory stages, completion on the stage's `full` mbarrier
//   warp 1   MMA issuer     one thread: tcgen05.mma.kind::f16 from the stage into TMEM, tcgen05.commit -> `empty` (stage free)
//                           and -> `tfull` (accumulators ready); owns the TMEM allocation
//   warps 2..9  epilogue    tcgen05.ld from TMEM (warp % 4 = TMEM lane quadrant, two column halves), the f32 arithmetic, stores;
//                           arrive on `tempty` as soon as a TMEM stage is in registers
//
// persistent over output tiles (static round-robin; CTAs that run at the same time share a weight tile through L2), so the three
// pipelines never drain between tiles.
//
// k_tcw_exact<FMT> -- quantised weights, BIT-EXACT (same scheme as k_gemm_tc_xf, bgpt_tc.cuh: the activation operand is expanded
//   into masked columns so that one MMA delivers the reference's eight 4-element partial sums isum_l of a block as exact f32
//   values; the epilogue runs the reference's chains acc_l = fma(d_w d_a, isum_l, acc_l) in block order, then hsum_float_8:
//   ggml.c:2518-2541, 2824-2857, 3071-3093, 3386-3411, 3597-3618).  What changed against k_gemm_tc_xf: the operands arrive by TMA
//   from pre-built fp16 planes (weight codes decoded ONCE per model into a prompt-operand cache, activations expanded once per
//   matmul by k_tcw_expand) instead of LDG -> unpack -> STS by all threads of every CTA; TMEM is double-buffered (2 x 256 columns
//   = 2 blocks each) so the tensor pipe works while the CUDA cores drain the previous stage; the chains use packed FFMA2
//   (fma.rn.f32x2: two IEEE fmas per instruction, same bits).
//   Tile: 128 weight rows x 16 tokens; a stage = K 64 = two 32-blocks = 4 MMAs (M 128, N 64 = 16 tokens x 4 masked columns, K 16).
//
// k_tcw_f16 -- F16 weights: a K-accumulating UMMA (fp16 x fp16 -> f32 in TMEM, ggml.c:2409-2443 computes the same products and
//   adds them in 32 lanes; the tensor core adds them in its own order, so this path is tolerance-close (north star: <= 1e-3 on the
//   logits), not bit-identical, and serves only evals of f16_tc_min_rows token rows and more).  Tile 128 rows x 128 tokens.
//   The weights are read where they lie: a row's elements are permuted inside the row (bgpt_layout.h) and the activation operand
//   is written in the same permuted order, which a dot product does not see.
#pragma once
#include "bgpt_tc.cuh"
#include <cuda.h>

#define TW_THREADS 320
#define TW_EPI_WARPS 8
#ifndef TW_ROWS
#define TW_ROWS 128
#define TW_BK 64                       // K elements per stage: one 128-byte swizzle atom of fp16
#define TWX_TOK 16
#endif

// ---- exact kernel geometry
#define TWX_BROWS (TWX_TOK * 4)        // rows of the expanded activation operand per tile
#define TWX_STAGES 6
#define TWX_OFF_B  16384
#define TWX_OFF_DW 24576               // f32 [128 rows][2 blocks]
#define TWX_OFF_MW 25600
#define TWX_OFF_DA 26624               // f32 [16 tokens][2 blocks]
#define TWX_OFF_SA 26752
#define TWX_STAGE_BYTES 27648          // multiple of 1024: every stage's operand tiles stay 1024-byte aligned
// ---- f16 kernel geometry
#define TWH_TOK 128
#define TWH_STAGES 5
#define TWH_STAGE_BYTES 32768

__device__ __forceinline__ void tw_mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void tw_mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void tw_mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory"); }
// every wait carries a watchdog: a lost signal ends the launch with a trap (the host sees a launch failure) instead of hanging the GPU
__device__ __forceinline__ void tw_wait(uint32_t bar, uint32_t parity) {
    uint32_t done; unsigned spins = 0; long long t0 = 0;
    for (;;) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (done) return;
        if ((++spins & 0x3FFFu) == 0) { const long long now = clock64(); if (t0 == 0) t0 = now; else if (now - t0 > 4000000000LL) __trap(); }
    }
}
__device__ __forceinline__ void tw_tma_2d(uint32_t dst, const CUtensorMap * map, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 :: "r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
__device__ __forceinline__ void tw_bulk(uint32_t dst, const void * src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// K-major operand tile written by TMA with CU_TENSOR_MAP_SWIZZLE_128B: rows of 128 bytes, 8-row groups of 1024 bytes
__device__ __forceinline__ uint64_t tw_desc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t) ((smem_addr >> 4) & 0x3FFF);          // start address
    d |= (uint64_t) 1 << 16;                              // leading byte offset: unused for a swizzled K-major tile
    d |= (uint64_t) (1024u >> 4) << 32;                   // stride byte offset: between 8-row groups
    d |= (uint64_t) 1 << 46;                              // descriptor version (Blackwell)
    d |= (uint64_t) 2 << 61;                              // SWIZZLE_128B
    return d;
}
__device__ __forceinline__ uint32_t tw_idesc_f16(int n) {  // D = F32, A = B = F16, both K-major, M = 128, N = n
    return (1u << 4) | ((uint32_t) (n >> 3) << 17) | ((uint32_t) (TW_ROWS >> 4) << 24);
}
// two IEEE fmas in one instruction (FFMA2): d0 = fma(a, b0, d0), d1 = fma(a, b1, d1)
__device__ __forceinline__ void tw_fma2(float & d0, float & d1, float a, uint32_t b0, uint32_t b1) {
    uint64_t A, B, C;
    asm("mov.b64 %0, {%1, %2};" : "=l"(A) : "f"(a), "f"(a));
    asm("mov.b64 %0, {%1, %2};" : "=l"(B) : "r"(b0), "r"(b1));
    asm("mov.b64 %0, {%1, %2};" : "=l"(C) : "f"(d0), "f"(d1));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(C) : "l"(A), "l"(B), "l"(C));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(d0), "=f"(d1) : "l"(C));
}

__device__ __forceinline__ void tw_add2(float & d0, float & d1, uint32_t b0, uint32_t b1) {       // d0 += b0, d1 += b1 (two IEEE adds)
    uint64_t B, C;
    asm("mov.b64 %0, {%1, %2};" : "=l"(B) : "r"(b0), "r"(b1));
    asm("mov.b64 %0, {%1, %2};" : "=l"(C) : "f"(d0), "f"(d1));
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(C) : "l"(C), "l"(B));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(d0), "=f"(d1) : "l"(C));
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr) : "memory");
}
template <int NC> __device__ __forceinline__ void tw_ld(uint32_t taddr, uint32_t (&v)[NC]);
template <> __device__ __forceinline__ void tw_ld<32>(uint32_t taddr, uint32_t (&v)[32]) { tc_ld32(taddr, v); }
template <> __device__ __forceinline__ void tw_ld<16>(uint32_t taddr, uint32_t (&v)[16]) { tc_ld16(taddr, v); }

struct TwxArgs {
    CUtensorMap tmA;               // decoded weight codes, fp16 [M][K]
    CUtensorMap tmB;               // expanded activations, fp16 [4 n_pad][K]
    const float * sw; const float * mw;     // weight scales / minima as f32 [K / 64][M][2]
    const float * sa; const float * ss;     // activation scales d / s as f32 [K / 64][n_pad][2]
    int M, n, n_pad, tok0, nkb, n_row_tiles, n_tok_tiles;
    int dbg;                       // timing experiments only (BGPT_TCW_DBG; results are wrong): 1 = read half of the TMEM columns, 2 = no activation-operand loads
    Epi epi;
};

// TPT = tokens per epilogue thread: 8 -> 8 epilogue warps (4 TMEM lane quadrants x 2 token halves), 4 -> 16 epilogue warps
// (x 4 token quarters; half the registers and half the instructions per warp, twice the warps to hide the TMEM / shared-memory latency)
template <int FMT, int TPT>
__global__ void __launch_bounds__(64 + 32 * 4 * (TWX_TOK / TPT), 1) k_tcw_exact(const __grid_constant__ TwxArgs P) {
    extern __shared__ __align__(1024) uint8_t tw_smem[];
    constexpr bool HASM = (FMT == BG_Q4_1 || FMT == BG_Q5_1);
    constexpr int NEW = 4 * (TWX_TOK / TPT);                         // epilogue warps
    constexpr int NC = TPT * 4;                                      // TMEM columns per load: TPT tokens x 4 masked columns
    constexpr uint32_t TX = 16384u + 8192u + 1024u + 128u + (HASM ? 1024u + 128u : 0u);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t raw = tc_smem_u32(tw_smem);
    const uint32_t sbase = (raw + 1023u) & ~1023u;
    uint8_t * gbase = tw_smem + (sbase - raw);                       // generic pointer to the same place
    const uint32_t bars = sbase + TWX_STAGES * TWX_STAGE_BYTES;      // full[S] | empty[S] | tfull[2] | tempty[2] | tmem base
    auto full   = [&](int s) { return bars + 8u * (uint32_t) s; };
    auto empty  = [&](int s) { return bars + 8u * (uint32_t) (TWX_STAGES + s); };
    auto tfull  = [&](int s) { return bars + 8u * (uint32_t) (2 * TWX_STAGES + s); };
    auto tempty = [&](int s) { return bars + 8u * (uint32_t) (2 * TWX_STAGES + 2 + s); };
    volatile uint32_t * tmem_slot = (volatile uint32_t *) (gbase + TWX_STAGES * TWX_STAGE_BYTES + 8 * (2 * TWX_STAGES + 4));

    if (tid == 0) {
        for (int s = 0; s < TWX_STAGES; s++) { tw_mbar_init(full(s), 1); tw_mbar_init(empty(s), 1 + NEW); }
        for (int s = 0; s < 2; s++) { tw_mbar_init(tfull(s), 1); tw_mbar_init(tempty(s), NEW); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(bars + 8u * (2 * TWX_STAGES + 4)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;
    const int n_tiles = P.n_row_tiles * P.n_tok_tiles, nkb = P.nkb;

    if (warp == 0) {
        // ===== TMA producer
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" :: "l"(&P.tmA) : "memory");
            asm volatile("prefetch.tensormap [%0];" :: "l"(&P.tmB) : "memory");
            uint32_t it = 0;
            for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
                const int rt = t / P.n_tok_tiles, tt = t - rt * P.n_tok_tiles;
                for (int kb = 0; kb < nkb; kb++, it++) {
                    const int s = (int) (it % TWX_STAGES); const uint32_t ph = (it / TWX_STAGES) & 1u;
                    tw_wait(empty(s), ph ^ 1u);
                    const uint32_t st = sbase + (uint32_t) s * TWX_STAGE_BYTES;
                    tw_mbar_expect_tx(full(s), (P.dbg & 2) ? TX - 8192u : TX);
                    tw_tma_2d(st, &P.tmA, kb * TW_BK, rt * TW_ROWS, full(s));
                    if (!(P.dbg & 2)) tw_tma_2d(st + TWX_OFF_B, &P.tmB, kb * TW_BK, tt * TWX_BROWS, full(s));
                    tw_bulk(st + TWX_OFF_DW, P.sw + ((size_t) kb * P.M + (size_t) rt * TW_ROWS) * 2, 1024u, full(s));
                    tw_bulk(st + TWX_OFF_DA, P.sa + ((size_t) kb * P.n_pad + (size_t) tt * TWX_TOK) * 2, 128u, full(s));
                    if (HASM) {
                        tw_bulk(st + TWX_OFF_MW, P.mw + ((size_t) kb * P.M + (size_t) rt * TW_ROWS) * 2, 1024u, full(s));
                        tw_bulk(st + TWX_OFF_SA, P.ss + ((size_t) kb * P.n_pad + (size_t) tt * TWX_TOK) * 2, 128u, full(s));
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: stage = two 32-blocks = 4 MMAs (block, half) with their own 64 accumulator columns each, no accumulation
        if (lane == 0) {
            const uint32_t idesc = tw_idesc_f16(TWX_BROWS);
            uint32_t it = 0;
            for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
                for (int kb = 0; kb < nkb; kb++, it++) {
                    const int s = (int) (it % TWX_STAGES); const uint32_t ph = (it / TWX_STAGES) & 1u;
                    const uint32_t ts = it & 1u, tph = (it >> 1) & 1u;
                    tw_wait(tempty(ts), tph ^ 1u);
                    tw_wait(full(s), ph);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t st = sbase + (uint32_t) s * TWX_STAGE_BYTES;
#pragma unroll
                    for (uint32_t q = 0; q < 4; q++)                 // q = block * 2 + half: K elements 16 q .. 16 q + 15 of the stage
                        tc_mma_f16(tmem + ts * 256u + q * 64u, tw_desc_sw128(st + q * 32u), tw_desc_sw128(st + TWX_OFF_B + q * 32u), idesc, 0u);
                    tc_commit(empty(s));
                    tc_commit(tfull(ts));
                }
            }
        }
    } else {
        // ===== epilogue: thread = (weight row quad * 32 + lane, tokens tg * TPT .. + TPT - 1), 8 running sums per token
        const int quad = warp & 3, tg = (warp - 2) >> 2;
        const int erow = quad * 32 + lane;
        uint32_t it = 0;
        for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
            const int rt = t / P.n_tok_tiles, tt = t - rt * P.n_tok_tiles;
            float acc[TPT][8], summ[TPT];
#pragma unroll
            for (int k = 0; k < TPT; k++) { summ[k] = 0.0f;
#pragma unroll
                for (int l = 0; l < 8; l++) acc[k][l] = 0.0f; }
            for (int kb = 0; kb < nkb; kb++, it++) {
                const int s = (int) (it % TWX_STAGES); const uint32_t ph = (it / TWX_STAGES) & 1u;
                const uint32_t ts = it & 1u, tph = (it >> 1) & 1u;
                const uint8_t * st = gbase + (size_t) s * TWX_STAGE_BYTES;
                tw_wait(full(s), ph);                                // the scale planes of the stage (async-proxy writes) are visible
                float sc[2][TPT];
                {
                    const float2 dw = *(const float2 *) (st + TWX_OFF_DW + erow * 8);
                    float2 mwv = make_float2(0.f, 0.f);
                    if (HASM) mwv = *(const float2 *) (st + TWX_OFF_MW + erow * 8);
#pragma unroll
                    for (int k = 0; k < TPT; k++) {
                        const float2 da = *(const float2 *) (st + TWX_OFF_DA + (tg * TPT + k) * 8);
                        sc[0][k] = __fmul_rn(dw.x, da.x); sc[1][k] = __fmul_rn(dw.y, da.y);
                        if (HASM) {
                            const float2 sa = *(const float2 *) (st + TWX_OFF_SA + (tg * TPT + k) * 8);
                            summ[k] = fmaf(mwv.x, sa.x, summ[k]); summ[k] = fmaf(mwv.y, sa.y, summ[k]);
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) tw_mbar_arrive(empty(s));             // this warp is done with the stage's shared memory
                tw_wait(tfull(ts), tph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                // TMEM columns of the stage: [ts * 256 + q * 64 + n * 4 + l'] = (float) isum_{4 (q & 1) + l'} of token n, block q >> 1
                const uint32_t tb = tmem + ((uint32_t) (quad * 32) << 16) + ts * 256u + (uint32_t) (tg * NC);
                uint32_t v[2][NC];
                tw_ld<NC>(tb, v[0]);
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    tc_ld_wait();
                    if (q < 3) { if (!(P.dbg & 1) || q == 2) tw_ld<NC>(tb + (uint32_t) ((q + 1) * 64), v[(q + 1) & 1]); }
                    else {
                        // all four accumulators of the stage are in registers: hand the TMEM stage back before doing the last chains
                        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                        __syncwarp();
                        if (lane == 0) tw_mbar_arrive(tempty(ts));
                    }
                    const int blk = q >> 1, c = q & 1;
#pragma unroll
                    for (int k = 0; k < TPT; k++) {
                        tw_fma2(acc[k][c * 4 + 0], acc[k][c * 4 + 1], sc[blk][k], v[q & 1][k * 4 + 0], v[q & 1][k * 4 + 1]);
                        tw_fma2(acc[k][c * 4 + 2], acc[k][c * 4 + 3], sc[blk][k], v[q & 1][k * 4 + 2], v[q & 1][k * 4 + 3]);
                    }
                }
            }
            // hsum_float_8 (+ summs) and the fused epilogue of the matmul
            const int r = rt * TW_ROWS + erow;
            if (r < P.M) {
                float x[TPT];
#pragma unroll
                for (int k = 0; k < TPT; k++) {
                    x[k] = __fadd_rn(__fadd_rn(__fadd_rn(acc[k][0], acc[k][4]), __fadd_rn(acc[k][2], acc[k][6])),
                                     __fadd_rn(__fadd_rn(acc[k][1], acc[k][5]), __fadd_rn(acc[k][3], acc[k][7])));
                    if (HASM) x[k] = __fadd_rn(x[k], summ[k]);
                }
                bg_epilogue_rows<TPT>(P.epi, P.tok0 + tt * TWX_TOK + tg * TPT, P.n, r, x);
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncwarp();
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512u) : "memory");
}

// ---- prompt-operand cache: device weight rows (bgpt_layout.h planes) -> fp16 codes [M][K] in element order + f32 scale planes
// thread = (row, group g of 4 blocks, piece p of 8): p = half * 4 + j (4/5-bit formats: the nibble half and the word j of the
// group; Q8_0: c * 4 + j).  The thread's 16 codes are elements 16 (p >> 2) + 4 j .. + 3 of the group's four blocks.
struct TwDecodeArgs {
    const uint8_t * W[3]; int rows_per, M, G, stride, off_qh, off_d, off_m;
    __half * out; float * sw; float * mw; int K;
};
template <int FMT>
__global__ void __launch_bounds__(256) k_tcw_decode(TwDecodeArgs a) {
    constexpr bool IS8   = (FMT == BG_Q8_0);
    constexpr bool HASQH = (FMT == BG_Q5_0 || FMT == BG_Q5_1);
    constexpr bool HASM  = (FMT == BG_Q4_1 || FMT == BG_Q5_1);
    const long long idx = (long long) blockIdx.x * blockDim.x + threadIdx.x;
    const int per_row = a.G * 8;
    const int r = (int) (idx / per_row);
    if (r >= a.M) return;
    const int rem = (int) (idx - (long long) r * per_row), g = rem >> 3, p = rem & 7, hi = p >> 2, j = p & 3;
    const int mat = r / a.rows_per;
    const uint8_t * wrow = a.W[mat] + (size_t) (r - mat * a.rows_per) * a.stride;
    const uint4 wq = IS8 ? *(const uint4 *) (wrow + (size_t) ((g * 2 + hi) * 4 + j) * 16) : *(const uint4 *) (wrow + (size_t) (g * 4 + j) * 16);
    uint32_t qh = 0;
    if (HASQH) qh = *(const uint32_t *) (wrow + a.off_qh + g * 16 + j * 4);
    const uint32_t ww[4] = { wq.x, wq.y, wq.z, wq.w };
#pragma unroll
    for (int i = 0; i < 4; i++) {
        uint32_t v = ww[i];
        if (!IS8) {
            v = hi ? ((ww[i] >> 4) & 0x0F0F0F0Fu) : (ww[i] & 0x0F0F0F0Fu);
            if (HASQH) { const uint32_t hb = (qh >> (8 * i)) & 0xFFu; v |= bg_spread4(hi ? (hb >> 4) : (hb & 0xFu)); }
            if (FMT == BG_Q4_0) v = tc_sub8(v);
            if (FMT == BG_Q5_0) v = tc_sub16(v);
        }
        const int b = g * 4 + i;
        if (b * 32 < a.K) *(uint2 *) (a.out + (size_t) r * a.K + b * 32 + hi * 16 + j * 4) = tc_s8x4_to_h4(v);
    }
    if (p == 0) {
        const uint2 dh = *(const uint2 *) (wrow + a.off_d + g * 8);
        const uint16_t dd[4] = { (uint16_t) (dh.x & 0xFFFF), (uint16_t) (dh.x >> 16), (uint16_t) (dh.y & 0xFFFF), (uint16_t) (dh.y >> 16) };
        uint16_t mm[4] = { 0, 0, 0, 0 };
        if (HASM) { const uint2 mh = *(const uint2 *) (wrow + a.off_m + g * 8); mm[0] = (uint16_t) (mh.x & 0xFFFF); mm[1] = (uint16_t) (mh.x >> 16); mm[2] = (uint16_t) (mh.y & 0xFFFF); mm[3] = (uint16_t) (mh.y >> 16); }
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int b = g * 4 + i;
            if (b * 32 >= a.K) continue;
            const size_t o = ((size_t) (b >> 1) * a.M + r) * 2 + (b & 1);
            a.sw[o] = bg_h2f(dd[i]);
            if (HASM) a.mw[o] = bg_h2f(mm[i]);
        }
    }
}

// ---- activation records (k_act) -> the expanded fp16 operand [4 n_pad][K]: row (n, l') holds token n's codes at the elements
// e of every block with (e % 16) / 4 == l' and zeros elsewhere (the zeros are written once, when the buffer is allocated) -- plus
// the f32 planes d / s [K / 64][n_pad][2].  thread = (token, group g, element group l of 8).
struct TwExpandArgs {
    const uint8_t * act; int act_bytes, off_dd, off_s, G, K;
    int n, n_pad, tok0;            // tokens tok0 .. n - 1 of the record array become operand tokens 0 .. ; n_pad - (n - tok0) trailing pad tokens are zeroed
    __half * out; float * sa; float * ss; int hasm;
};
static __global__ void __launch_bounds__(256) k_tcw_expand(TwExpandArgs a) {
    const long long idx = (long long) blockIdx.x * blockDim.x + threadIdx.x;
    const int per_tok = a.G * 8;
    const int tk = (int) (idx / per_tok);
    if (tk >= a.n_pad) return;
    const int rem = (int) (idx - (long long) tk * per_tok), g = rem >> 3, l = rem & 7;
    const bool valid = a.tok0 + tk < a.n;
    const uint8_t * rec = a.act + (size_t) (valid ? a.tok0 + tk : 0) * a.act_bytes;
    uint4 w = make_uint4(0, 0, 0, 0);
    if (valid) w = *(const uint4 *) (rec + (size_t) (g * 8 + l) * 16);
    const uint32_t ww[4] = { w.x, w.y, w.z, w.w };
    __half * row = a.out + ((size_t) tk * 4 + (l & 3)) * a.K;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int b = g * 4 + i;
        if (b * 32 < a.K) *(uint2 *) (row + b * 32 + l * 4) = valid ? tc_s8x4_to_h4(ww[i]) : make_uint2(0u, 0u);
    }
    if (l == 0) {
        float4 d = make_float4(0.f, 0.f, 0.f, 0.f), s = make_float4(0.f, 0.f, 0.f, 0.f);
        if (valid) { d = *(const float4 *) (rec + a.off_dd + g * 16); if (a.hasm) s = *(const float4 *) (rec + a.off_s + g * 16); }
        const float dd[4] = { d.x, d.y, d.z, d.w }, sv[4] = { s.x, s.y, s.z, s.w };
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int b = g * 4 + i;
            if (b * 32 >= a.K) continue;
            const size_t o = ((size_t) (b >> 1) * a.n_pad + tk) * 2 + (b & 1);
            a.sa[o] = dd[i];
            if (a.hasm) a.ss[o] = sv[i];
        }
    }
}

// ================================================================================================================================
// k_tcw_f16 -- F16 weights x fp16-rounded activations, f32 accumulation over all of K in TMEM
// ================================================================================================================================
struct TwhArgs {
    CUtensorMap tmA[3];            // the (up to three stacked) weight matrices where they lie, fp16 [rows_per][K] (K permuted inside the row)
    CUtensorMap tmB;               // activations, fp16 [n - tok0][K] in the same element order
    int rows_per, M, n, tok0, nkb, n_row_tiles, n_tok_tiles;
    Epi epi;
};
// SPLIT: every pipeline stage (K 64 = 4 MMAs) gets its own TMEM accumulator (two, alternating) and the epilogue warps add the
// stage results in registers with round-to-nearest f32 adds while the next stage is multiplied.  The tensor core truncates when it
// accumulates onto a non-zero accumulator; K / 16 such steps in one accumulator measure ~2e-6 (K 1024) .. 5e-6 (K 4096) of the
// largest output, 3 steps per stage + K / 64 rounded adds is the error of an ordinary f32 dot product.  !SPLIT: one accumulator for
// all of K, one TMEM read-out per tile (faster when the step is MMA-bound, less accurate).
template <bool SPLIT>
__global__ void __launch_bounds__(TW_THREADS, 1) k_tcw_f16(const __grid_constant__ TwhArgs P) {
    extern __shared__ __align__(1024) uint8_t tw_smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t raw = tc_smem_u32(tw_smem);
    const uint32_t sbase = (raw + 1023u) & ~1023u;
    uint8_t * gbase = tw_smem + (sbase - raw);
    const uint32_t bars = sbase + TWH_STAGES * TWH_STAGE_BYTES;
    auto full   = [&](int s) { return bars + 8u * (uint32_t) s; };
    auto empty  = [&](int s) { return bars + 8u * (uint32_t) (TWH_STAGES + s); };
    auto tfull  = [&](int s) { return bars + 8u * (uint32_t) (2 * TWH_STAGES + s); };
    auto tempty = [&](int s) { return bars + 8u * (uint32_t) (2 * TWH_STAGES + 2 + s); };
    volatile uint32_t * tmem_slot = (volatile uint32_t *) (gbase + TWH_STAGES * TWH_STAGE_BYTES + 8 * (2 * TWH_STAGES + 4));

    if (tid == 0) {
        for (int s = 0; s < TWH_STAGES; s++) { tw_mbar_init(full(s), 1); tw_mbar_init(empty(s), 1); }
        for (int s = 0; s < 2; s++) { tw_mbar_init(tfull(s), 1); tw_mbar_init(tempty(s), TW_EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(bars + 8u * (2 * TWH_STAGES + 4)), "r"(256u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;
    const int n_tiles = P.n_row_tiles * P.n_tok_tiles, nkb = P.nkb;

    if (warp == 0) {
        if (lane == 0) {
            uint32_t it = 0;
            for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
                const int rt = t / P.n_tok_tiles, tt = t - rt * P.n_tok_tiles;
                const int row0 = rt * TW_ROWS, mat = row0 / P.rows_per;
                const CUtensorMap * mapA = &P.tmA[mat];
                for (int kb = 0; kb < nkb; kb++, it++) {
                    const int s = (int) (it % TWH_STAGES); const uint32_t ph = (it / TWH_STAGES) & 1u;
                    tw_wait(empty(s), ph ^ 1u);
                    const uint32_t st = sbase + (uint32_t) s * TWH_STAGE_BYTES;
                    tw_mbar_expect_tx(full(s), (uint32_t) TWH_STAGE_BYTES);
                    tw_tma_2d(st, mapA, kb * TW_BK, row0 - mat * P.rows_per, full(s));
                    tw_tma_2d(st + 16384u, &P.tmB, kb * TW_BK, tt * TWH_TOK, full(s));
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = tw_idesc_f16(TWH_TOK);
            uint32_t it = 0, tile_it = 0;
            for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, tile_it++) {
                if (!SPLIT) tw_wait(tempty(tile_it & 1u), ((tile_it >> 1) & 1u) ^ 1u);
                for (int kb = 0; kb < nkb; kb++, it++) {
                    const int s = (int) (it % TWH_STAGES); const uint32_t ph = (it / TWH_STAGES) & 1u;
                    const uint32_t ts = SPLIT ? (it & 1u) : (tile_it & 1u);
                    if (SPLIT) tw_wait(tempty(ts), ((it >> 1) & 1u) ^ 1u);
                    tw_wait(full(s), ph);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t st = sbase + (uint32_t) s * TWH_STAGE_BYTES;
#pragma unroll
                    for (uint32_t q = 0; q < 4; q++)
                        tc_mma_f16(tmem + ts * (uint32_t) TWH_TOK, tw_desc_sw128(st + q * 32u), tw_desc_sw128(st + 16384u + q * 32u), idesc,
                                   (q > 0 || (!SPLIT && kb > 0)) ? 1u : 0u);
                    tc_commit(empty(s));
                    if (SPLIT) tc_commit(tfull(ts));
                }
                if (!SPLIT) tc_commit(tfull(tile_it & 1u));
            }
        }
    } else {
        const int quad = warp & 3, half = (warp - 2) >> 2;
        const int erow = quad * 32 + lane;
        uint32_t it = 0, tile_it = 0;
        for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, tile_it++) {
            const int rt = t / P.n_tok_tiles, tt = t - rt * P.n_tok_tiles;
            float acc[2][32];
            if (SPLIT) {
                for (int kb = 0; kb < nkb; kb++, it++) {
                    const uint32_t ts = it & 1u, tph = (it >> 1) & 1u;
                    tw_wait(tfull(ts), tph);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t tb = tmem + ((uint32_t) (quad * 32) << 16) + ts * (uint32_t) TWH_TOK + (uint32_t) (half * 64);
                    uint32_t v[2][32];
                    tc_ld32(tb, v[0]);
                    tc_ld32(tb + 32u, v[1]);
                    tc_ld_wait();
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) tw_mbar_arrive(tempty(ts));
                    if (kb == 0) {
#pragma unroll
                        for (int jj = 0; jj < 2; jj++)
#pragma unroll
                            for (int c = 0; c < 32; c++) acc[jj][c] = __uint_as_float(v[jj][c]);
                    } else {
#pragma unroll
                        for (int jj = 0; jj < 2; jj++)
#pragma unroll
                            for (int c = 0; c < 32; c += 2) tw_add2(acc[jj][c], acc[jj][c + 1], v[jj][c], v[jj][c + 1]);
                    }
                }
            } else {
                const uint32_t ts = tile_it & 1u, tph = (tile_it >> 1) & 1u;
                tw_wait(tfull(ts), tph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t tb = tmem + ((uint32_t) (quad * 32) << 16) + ts * (uint32_t) TWH_TOK + (uint32_t) (half * 64);
                uint32_t v[2][32];
                tc_ld32(tb, v[0]);
                tc_ld32(tb + 32u, v[1]);
                tc_ld_wait();
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) tw_mbar_arrive(tempty(ts));
#pragma unroll
                for (int jj = 0; jj < 2; jj++)
#pragma unroll
                    for (int c = 0; c < 32; c++) acc[jj][c] = __uint_as_float(v[jj][c]);
            }
            const int r = rt * TW_ROWS + erow;
            if (r < P.M) {
                bg_epilogue_rows<32>(P.epi, P.tok0 + tt * TWH_TOK + half * 64, P.n, r, acc[0]);
                bg_epilogue_rows<32>(P.epi, P.tok0 + tt * TWH_TOK + half * 64 + 32, P.n, r, acc[1]);
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncwarp();
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(256u) : "memory");
}

// F16 activation record (f32 values already rounded through fp16, float4 [gg][eh][lane]; bgpt_layout.h) -> fp16 [n][K] in the
// weights' in-row order: position (gg * 32 + lane) * 8 + e  <-  element gg * 256 + e * 32 + lane.  thread = 8 consecutive positions.
static __global__ void __launch_bounds__(256) k_tcw_act_h(const uint8_t * act, int act_bytes, int K, int n, int tok0, __half * out) {
    const long long idx = (long long) blockIdx.x * blockDim.x + threadIdx.x;
    const int per_tok = K / 8;
    const int tk = (int) (idx / per_tok);
    if (tok0 + tk >= n) return;
    const int u = (int) (idx - (long long) tk * per_tok), gg = u >> 5, lane = u & 31;
    const float * rec = (const float *) (act + (size_t) (tok0 + tk) * act_bytes);
    const float4 a = *(const float4 *) (rec + ((gg * 2 + 0) * 32 + lane) * 4), b = *(const float4 *) (rec + ((gg * 2 + 1) * 32 + lane) * 4);
    const __half2 h0 = __floats2half2_rn(a.x, a.y), h1 = __floats2half2_rn(a.z, a.w), h2 = __floats2half2_rn(b.x, b.y), h3 = __floats2half2_rn(b.z, b.w);
    uint4 o;
    o.x = *reinterpret_cast<const uint32_t *>(&h0); o.y = *reinterpret_cast<const uint32_t *>(&h1);
    o.z = *reinterpret_cast<const uint32_t *>(&h2); o.w = *reinterpret_cast<const uint32_t *>(&h3);
    *(uint4 *) (out + (size_t) tk * K + (size_t) u * 8) = o;
}
