// bgpt_skinny.cuh -- fused schedule for SKINNY batches (2..111 token rows) of the quantised formats
// at BioGPT-base layer shapes (d_model 1024, 16 heads of 64, d_ff 4096): the reference's own prompt
// chunking (n_batch = 8, BASELINE configs[2]) and lock-step streams (8 sequences per GPU, configs[3]).
//
// The per-operator schedule (bgpt_kernels.cuh) spends 9 launches per layer, its GEMV keeps 4 threads
// on a row walking K with dependent loads, and its LayerNorm / quantise kernels run on 8 CTAs: 10-17 us
// per kernel, 2.8 ms per 8-row step.  Here a layer is 8 launches, every one of them wide:
//
//   k_sk_ln        : LayerNorm0 + quantise, one token row per CTA                   (rows CTAs)
//   k_sk_mm  qkv   : 3072 stacked rows, one warp per weight row, 8 token rows per warp; epilogue bias,
//                    q scale, KV append; prefetches the cached K/V rows of the layer into L2    (384 CTAs)
//   k_sk_attn      : one (head, row) per CTA of 1024 threads, 16 K rows in flight per warp, transposing
//                    butterfly for the reduce trees, V with 16 rows in flight per thread; the 64 outputs
//                    leave as two quantised blocks of out_proj's activation record  (16 x rows CTAs)
//   k_sk_mm  o     : 4-row token tiles; bias + residual                             (128 x 2 CTAs)
//   k_sk_ln        : LayerNorm1 + quantise
//   k_sk_mm  fc1   : 4096 rows, 8-row tiles; bias -> f32                              (512 CTAs)
//   k_sk_gq        : fp16-table GELU + quantise -> fc2's activation record            (4 x rows CTAs)
//                    (BGPT_SK_FC1_SPLIT=0: 32 consecutive rows per CTA = one block of fc2's input, GELU +
//                    quantise in the matmul's epilogue instead)
//   k_sk_mm  fc2   : K = 4096; bias + residual                                      (128 x 2 CTAs)
//   k_sk_ln + k_sk_mm lm_head (lock-step streams: every row has logits)              (1325 CTAs)
//
// Dot products keep the bits of k_gemv_q and the persistent kernels -- the 8 running sums of
// ggml_vec_dot_q*_q8_*, acc_l = fma(d_w d_a, (float) isum_l, acc_l) in block order, hsum_float_8 --
// with lane = (token, running sum): see k_sk_mm.  The kernels are chained by programmatic dependent
// launch: everything that touches only weights (tile fetch by TMA bulk copy, decode) sits before
// griddepcontrol.wait and overlaps the tail of the previous kernel; one CUDA graph per (rows, mode).
#pragma once
#include "bgpt_kernels.cuh"
#include "bgpt_mega4.cuh"      // mbarrier / bulk-copy helpers

#define SK_D 1024
#define SK_DK 64
#define SK_ANT 1024                                // threads of the attention kernel

enum { SK_EPI_STORE = 0, SK_EPI_QKV = 1, SK_EPI_RESID = 2, SK_EPI_GELUQ = 3 };

struct SkArgs {
    const uint8_t * W[3];          // up to three stacked matrices (q, k, v) of rows_per rows each
    int rows_per, M;               // rows of one matrix, total rows
    int npass;                     // K / 1024
    int stride, off_qh, off_d, off_m;
    const uint8_t * act;           // activation records of the token rows (k_sk_ln, k_sk_attn or a GELU epilogue wrote them)
    int act_bytes, off_n, off_dd, off_s, code_off;      // record layout of the input
    int n, tok0;                   // token rows [tok0, n)
    int rpw;                       // rows per warp: a CTA of nw warps covers nw * rpw consecutive rows
    int pdl_trig;                  // griddepcontrol.launch_dependents: 0 at the start of the kernel, 1 after the dot products
    // epilogue
    int epi;
    const float * bias[3];
    float * out; int ld_out;
    const float * resid; int ld_resid;
    float * kcache; float * vcache; size_t stream_stride; float qscale;
    const DevState * st; int mode;
    const uint16_t * gelu;
    int pf_streams;                // SK_EPI_QKV: streams whose cached K/V rows (n_past positions) are prefetched into the L2 for the attention kernel
    uint8_t * act_out; int out_bytes, out_off_n, out_off_d, out_off_s;    // SK_EPI_GELUQ: record of the next matmul
};

__device__ __forceinline__ void sk_pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void sk_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ---- 8 consecutive lanes (l = lane & 7) hold one 32-element block, lane l its elements 4l..4l+3 ----
// quantize_row_q8_0 / q8_1 (AVX branches, ggml.c:1166-1203, 1403-1450) into the record layout of
// bgpt_layout.h.  All 32 lanes call; `write` selects the lanes whose block is real.
template <int FMT>
__device__ __forceinline__ void sk_quant_block(float4 v, int b, int l, uint8_t * rec, int off_n, int off_d, int off_s, int code_off, bool write) {
    constexpr bool Q81 = (FMT == BG_Q4_1 || FMT == BG_Q5_1);
    float amax = fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w)));
    amax = fmaxf(amax, __shfl_xor_sync(FULLMASK, amax, 1));
    amax = fmaxf(amax, __shfl_xor_sync(FULLMASK, amax, 2));
    amax = fmaxf(amax, __shfl_xor_sync(FULLMASK, amax, 4));
    const float d  = __fdiv_rn(amax, 127.0f);
    const float id = (amax != 0.0f) ? __fdiv_rn(127.0f, amax) : 0.0f;
    const int q0 = __float2int_rn(__fmul_rn(v.x, id)), q1 = __float2int_rn(__fmul_rn(v.y, id));
    const int q2 = __float2int_rn(__fmul_rn(v.z, id)), q3 = __float2int_rn(__fmul_rn(v.w, id));
    const int s4 = q0 + q1 + q2 + q3;
    int stot = s4;
    stot += __shfl_xor_sync(FULLMASK, stot, 1);
    stot += __shfl_xor_sync(FULLMASK, stot, 2);
    stot += __shfl_xor_sync(FULLMASK, stot, 4);
    if (write) {
        const int g = b >> 2, i = b & 3;
        ((uint32_t *) rec)[(g * 8 + l) * 4 + i] = ((uint32_t) q0 & 0xFFu) | (((uint32_t) q1 & 0xFFu) << 8) | (((uint32_t) q2 & 0xFFu) << 16) | (((uint32_t) q3 & 0xFFu) << 24);
        ((int32_t *) (rec + off_n))[(g * 8 + l) * 4 + i] = -code_off * s4;
        if (l == 0) {
            ((float *) (rec + off_d))[b] = Q81 ? d : bg_h2f(bg_f2h(d));
            ((float *) (rec + off_s))[b] = Q81 ? __fmul_rn(d, (float) stot) : 0.0f;
        }
    }
}

// ---- stand-alone LayerNorm + quantise: one token row per CTA of 256 threads (thread t: elements 4t..4t+3 = block t/8, word t%8),
// record to global memory; the matmul kernels then stage the records with one bulk copy.  Doing the LayerNorm ONCE instead of
// in the prologue of each of a matmul's 256-768 CTAs takes ~1000 instructions per warp off their critical path.
struct SkLnArgs {
    const float * xin; int ld_in; const float * lnw; const float * lnb; float eps;
    uint8_t * act; int act_bytes, off_n, off_d, off_s, code_off;
    int pdl_trig;
};
template <int FMT>
__global__ void __launch_bounds__(256, 4) k_sk_ln(const __grid_constant__ SkLnArgs a) {
    __shared__ double sA[8], sB[8];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, row = blockIdx.x;
    const float4 w = *((const float4 *) a.lnw + tid), bb = *((const float4 *) a.lnb + tid);
    if (a.pdl_trig == 0) sk_pdl_launch_dependents();
    sk_pdl_wait();
    float4 v = __ldcg((const float4 *) (a.xin + (size_t) row * a.ld_in) + tid);
    double s = ((double) v.x + (double) v.y) + ((double) v.z + (double) v.w);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(FULLMASK, s, o);
    if (lane == 0) sA[warp] = s;
    __syncthreads();
    s = ((sA[0] + sA[1]) + (sA[2] + sA[3])) + ((sA[4] + sA[5]) + (sA[6] + sA[7]));
    const float mean = (float) (s * (1.0 / SK_D));
    v.x = __fsub_rn(v.x, mean); v.y = __fsub_rn(v.y, mean); v.z = __fsub_rn(v.z, mean); v.w = __fsub_rn(v.w, mean);
    double s2 = ((double) __fmul_rn(v.x, v.x) + (double) __fmul_rn(v.y, v.y)) + ((double) __fmul_rn(v.z, v.z) + (double) __fmul_rn(v.w, v.w));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s2 += __shfl_xor_sync(FULLMASK, s2, o);
    if (lane == 0) sB[warp] = s2;
    __syncthreads();
    s2 = ((sB[0] + sB[1]) + (sB[2] + sB[3])) + ((sB[4] + sB[5]) + (sB[6] + sB[7]));
    if (a.pdl_trig == 1) sk_pdl_launch_dependents();
    const float variance = (float) (s2 * (1.0 / SK_D));
    const float scale = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(variance, a.eps)));
    float4 y;
    y.x = __fadd_rn(__fmul_rn(w.x, __fmul_rn(v.x, scale)), bb.x);
    y.y = __fadd_rn(__fmul_rn(w.y, __fmul_rn(v.y, scale)), bb.y);
    y.z = __fadd_rn(__fmul_rn(w.z, __fmul_rn(v.z, scale)), bb.z);
    y.w = __fadd_rn(__fmul_rn(w.w, __fmul_rn(v.w, scale)), bb.w);
    sk_quant_block<FMT>(y, tid >> 3, tid & 7, a.act + (size_t) row * a.act_bytes, a.off_n, a.off_d, a.off_s, a.code_off, true);
}

// ---- stand-alone GELU + quantise of fc1's output (alternative to the quantising epilogue of k_sk_mm, BGPT_SK_FC1_SPLIT=1):
// grid = (d_ff / 1024, rows), block = 256; thread t: elements 4t..4t+3 of the 1024-wide slice = block t/8, word t%8
struct SkGqArgs {
    const float * hin; int ld_in; const uint16_t * gelu;
    uint8_t * act; int act_bytes, off_n, off_d, off_s, code_off;
    int pdl_trig;
};
template <int FMT>
__global__ void __launch_bounds__(256, 4) k_sk_gq(const __grid_constant__ SkGqArgs a) {
    const int tid = threadIdx.x, row = blockIdx.y;
    if (a.pdl_trig == 0) sk_pdl_launch_dependents();
    sk_pdl_wait();
    const float4 v = __ldcg((const float4 *) (a.hin + (size_t) row * a.ld_in + blockIdx.x * 1024) + tid);
    float4 y;
    y.x = bg_h2f(a.gelu[bg_f2h(v.x)]); y.y = bg_h2f(a.gelu[bg_f2h(v.y)]);
    y.z = bg_h2f(a.gelu[bg_f2h(v.z)]); y.w = bg_h2f(a.gelu[bg_f2h(v.w)]);
    if (a.pdl_trig == 1) sk_pdl_launch_dependents();
    sk_quant_block<FMT>(y, blockIdx.x * 32 + (tid >> 3), tid & 7, a.act + (size_t) row * a.act_bytes, a.off_n, a.off_d, a.off_s, a.code_off, true);
}

// ---------------------------------------------------------------------------------------------
// k_sk_mm: y[tok][row] = dot(W[row], record[tok]) + epilogue, one warp per weight row.
//
// 4-row tiles: lane = (tq = lane / 8, l = lane % 8) owns running sum l of token tq.  8-row tiles: lane = (tq = lane / 4,
// l = lane % 4) owns the PAIR of running sums l and l + 4 of token tq (two chains sharing the scale product).  The lane
// walks its sums from the first block of the row to the last -- integer dot (dp4a), conversion, scale product and the
// fma, in block order; nothing is exchanged between lanes until hsum_float_8 at the end of the row.  The lane's operands
// of four consecutive blocks are one 16-byte word each, by construction of the layouts in bgpt_layout.h:
//   weights  decoded once per CTA to int8 codes [g][sum][i] (Q8_0 rows already are), scales / minima to f32 [blk]
//   records  aq[g][sum][i], an[g][sum][i] (-offset * sum of the 4 codes), ad[blk], as[blk]
// The CTA's weight rows are contiguous in HBM: ONE bulk copy (TMA) brings the whole tile into shared memory, issued
// before griddepcontrol.wait -- the weights are in flight while the previous kernel of the chain is still running --
// and more bring the token records once that kernel has finished.  Lanes of different tokens read the same weight
// word (broadcast), a token's lanes read consecutive 16-byte words of its record, records of an 8-row tile sit 64 bytes apart
// modulo the bank period: no bank conflicts.
// grid = (ceil(M / (nw * rpw)), ceil((n - tok0) / TN)), block = 32 * nw,
// dynamic smem = TN * (act_bytes + (TN == 8 ? 64 : 0)) + nw * rpw * (stride + (Q8_0 ? 0 : K) + K / 32 * 4 * (1 + has minima))
// ---------------------------------------------------------------------------------------------
template <int FMT, int TN>
__global__ void __launch_bounds__(1024, 1) k_sk_mm(const __grid_constant__ SkArgs a) {       // 64 registers: 2 x 512 or 4 x 256 threads per SM as well
    constexpr bool IS8    = (FMT == BG_Q8_0);
    constexpr bool HASQH  = (FMT == BG_Q5_0 || FMT == BG_Q5_1);
    constexpr bool HASM   = (FMT == BG_Q4_1 || FMT == BG_Q5_1);
    constexpr bool HASOFF = (FMT == BG_Q4_0 || FMT == BG_Q5_0);
    extern __shared__ __align__(16) uint8_t sk_smem[];
    __shared__ __align__(16) float s_g[8 * 32];               // SK_EPI_GELUQ: [token][row of the CTA]
    __shared__ __align__(8) uint64_t s_bar[2];                 // weights, records
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nt = blockDim.x, nw = nt >> 5;
    const int rows_cta = nw * a.rpw;
    const int K = a.npass * 1024, nb = K >> 5;
    const int rec_stride = a.act_bytes + (TN == 8 ? 64 : 0);                     // 8-row tiles: two tokens share a quarter-warp, +64 bytes keeps them on different banks
    uint8_t * s_rec = sk_smem;                                                   // [TN][rec_stride] token records
    uint8_t * s_w = s_rec + (size_t) TN * rec_stride;                            // [rows_cta][stride] weight rows as they lie in HBM
    uint8_t * s_dec = s_w + (size_t) rows_cta * a.stride;                        // [rows_cta][K] int8 codes [g][l][i] (Q8_0: the raw rows already are)
    float * s_dw = (float *) (s_dec + (IS8 ? 0 : (size_t) rows_cta * K));        // [rows_cta][nb] block scales as f32
    float * s_mw = s_dw + (size_t) rows_cta * nb;                                // [rows_cta][nb] block minima as f32 (Q4_1 / Q5_1)
    const int tokbase = a.tok0 + blockIdx.y * TN;
    const int rowbase = blockIdx.x * rows_cta;
    // TN = 4: lane = (token tq, running sum l), one chain per lane.  TN = 8: lane = (token tq, pair l): running sums l and l + 4,
    // two chains per lane that share the scale product -- per 8 tokens less than half the instructions of two 4-token tiles.
    const int tq = TN == 8 ? lane >> 2 : lane >> 3, l = TN == 8 ? (lane & 3) : (lane & 7);
    auto mat_of = [&](int row) -> int { return row >= 2 * a.rows_per ? 2 : (row >= a.rows_per ? 1 : 0); };
    // ---- everything that does not depend on the previous kernel: the weight tile, n_past, K/V rows towards the L2
    if (tid == 0) {
        m4_mbar_init(&s_bar[0], 1); m4_mbar_init(&s_bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        int nr = a.M - rowbase; nr = nr > rows_cta ? rows_cta : nr;
        const int mat = mat_of(rowbase);                       // a CTA's rows never straddle two of the stacked matrices
        const uint32_t bytes = (uint32_t) nr * (uint32_t) a.stride;
        m4_mbar_expect(&s_bar[0], bytes);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     :: "r"(m4_s32(s_w)), "l"(a.W[mat] + (size_t) (rowbase - mat * a.rows_per) * a.stride), "r"(bytes), "r"(m4_s32(&s_bar[0])) : "memory");
    }
    int n_past = 0;
    if (a.epi == SK_EPI_QKV) {
        n_past = a.st->n_past;
        if (a.pf_streams > 0 && n_past > 0 && blockIdx.y == 0) {     // 64 lines of 128 bytes per cached position and stream (K row | V row)
            const int per = n_past * 64, totl = per * a.pf_streams;
#pragma unroll 1
            for (int idx = blockIdx.x * nt + tid; idx < totl; idx += gridDim.x * nt) {
                const int sidx = idx / per, rem = idx - sidx * per, t = rem >> 6, jj = rem & 63;
                const float * base = (jj < 32 ? a.kcache : a.vcache) + (size_t) sidx * a.stream_stride + (size_t) t * SK_D + (jj & 31) * 32;
                asm volatile("prefetch.global.L2 [%0];" :: "l"(base));
            }
        }
    }
    if (a.pdl_trig == 0) sk_pdl_launch_dependents();
    __syncthreads();                                           // barriers initialised
    // ---- decode the tile ONCE per CTA (still nothing here depends on the previous kernel): 4/5-bit codes -> int8 words in
    //      the lane order of the dot loop, fp16 scales / minima -> f32.  Every weight is then used by TN token lanes as is.
    m4_mbar_wait(&s_bar[0], 0);
    {
        int nr = a.M - rowbase; nr = nr > rows_cta ? rows_cta : nr;
        if (!IS8) {
            const int upr = nb;                                // (g, j) units per row: one 16-byte weight word each
            for (int u = tid; u < nr * upr; u += nt) {
                const int rl = u / upr, gj = u - rl * upr, g = gj >> 2, jj = gj & 3;
                const uint8_t * wr = s_w + (size_t) rl * a.stride;
                const uint4 w = *(const uint4 *) (wr + gj * 16);
                uint4 lo, hi;
                lo.x = w.x & 0x0F0F0F0Fu; lo.y = w.y & 0x0F0F0F0Fu; lo.z = w.z & 0x0F0F0F0Fu; lo.w = w.w & 0x0F0F0F0Fu;
                hi.x = (w.x >> 4) & 0x0F0F0F0Fu; hi.y = (w.y >> 4) & 0x0F0F0F0Fu; hi.z = (w.z >> 4) & 0x0F0F0F0Fu; hi.w = (w.w >> 4) & 0x0F0F0F0Fu;
                if (HASQH) {
                    const uint32_t qh = *(const uint32_t *) (wr + a.off_qh + gj * 4);
                    lo.x |= bg_spread4(qh & 0xFu);          hi.x |= bg_spread4((qh >> 4) & 0xFu);
                    lo.y |= bg_spread4((qh >> 8) & 0xFu);   hi.y |= bg_spread4((qh >> 12) & 0xFu);
                    lo.z |= bg_spread4((qh >> 16) & 0xFu);  hi.z |= bg_spread4((qh >> 20) & 0xFu);
                    lo.w |= bg_spread4((qh >> 24) & 0xFu);  hi.w |= bg_spread4((qh >> 28) & 0xFu);
                }
                uint8_t * dr = s_dec + (size_t) rl * K;
                *(uint4 *) (dr + (g * 8 + jj) * 16) = lo;
                *(uint4 *) (dr + (g * 8 + jj + 4) * 16) = hi;
            }
        }
        for (int u = tid; u < nr * nb; u += nt) {
            const int rl = u / nb, b = u - rl * nb;
            const uint8_t * wr = s_w + (size_t) rl * a.stride;
            s_dw[u] = bg_h2f(*(const uint16_t *) (wr + a.off_d + b * 2));
            if (HASM) s_mw[u] = bg_h2f(*(const uint16_t *) (wr + a.off_m + b * 2));
        }
    }
    sk_pdl_wait();                                             // from here on the previous kernel's results are visible
    // ---- token records: bulk copies (TMA); 4-row tiles: consecutive records are contiguous, one copy
    {
        int nv = a.n - tokbase; nv = nv > TN ? TN : nv;
        const uint32_t bytes = (uint32_t) nv * (uint32_t) a.act_bytes;
        if (tid == 0) {
            m4_mbar_expect(&s_bar[1], bytes);
            const int ncopy = TN == 8 ? nv : 1;
            const uint32_t each = TN == 8 ? (uint32_t) a.act_bytes : bytes;
            for (int t = 0; t < ncopy; t++)
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             :: "r"(m4_s32(s_rec + (size_t) t * rec_stride)), "l"(a.act + (size_t) (tokbase + t) * a.act_bytes), "r"(each), "r"(m4_s32(&s_bar[1])) : "memory");
        }
        const int v16 = a.act_bytes >> 4;
        for (int i = nv * v16 + tid; i < TN * v16; i += nt) { const int t = i / v16; ((uint4 *) (s_rec + (size_t) t * rec_stride))[i - t * v16] = make_uint4(0, 0, 0, 0); }
        m4_mbar_wait(&s_bar[1], 0);
    }
    __syncthreads();
    const int ng = 8 * a.npass;                                // groups of 4 blocks in a row
    const uint8_t * rec = s_rec + (size_t) tq * rec_stride;
#pragma unroll 1
    for (int i = 0; i < a.rpw; i++) {
        const int rl = i * nw + warp, row = rowbase + rl;
        const uint8_t * wrow = IS8 ? s_w + (size_t) rl * a.stride : s_dec + (size_t) rl * K;   // codes [g][sum][i]
        const float * dwr = s_dw + (size_t) rl * nb, * mwr = s_mw + (size_t) rl * nb;
        const int tok = tokbase + tq;
        const bool owner = l == 0 && row < a.M && tok < a.n;
        // the row owner's bias / residual are in flight while the dots run
        float pbias = 0.0f, presid = 0.0f;
        if (owner) {
            if (a.epi == SK_EPI_QKV) { const int mat = mat_of(row); pbias = a.bias[mat][row - mat * a.rows_per]; }
            else if (a.bias[0]) pbias = a.bias[0][row];
            if (a.epi == SK_EPI_RESID) presid = __ldcg(a.resid + (size_t) tok * a.ld_resid + row);
        }
        float acc0 = 0.0f, acc1 = 0.0f, summ = 0.0f;
#pragma unroll 2
        for (int g = 0; g < ng; g++) {
            const float4 dw = *(const float4 *) (dwr + g * 4);
            const float4 da = *(const float4 *) (rec + a.off_dd + g * 16);
            const float sx = __fmul_rn(dw.x, da.x), sy = __fmul_rn(dw.y, da.y), sz = __fmul_rn(dw.z, da.z), sw = __fmul_rn(dw.w, da.w);
            {
                const uint4 wq = *(const uint4 *) (wrow + (g * 8 + l) * 16);
                const uint4 av = *(const uint4 *) (rec + (g * 8 + l) * 16);
                int4 nv = make_int4(0, 0, 0, 0);
                if (HASOFF) nv = *(const int4 *) (rec + a.off_n + (g * 8 + l) * 16);
                acc0 = fmaf(sx, (float) __dp4a((int) wq.x, (int) av.x, nv.x), acc0);
                acc0 = fmaf(sy, (float) __dp4a((int) wq.y, (int) av.y, nv.y), acc0);
                acc0 = fmaf(sz, (float) __dp4a((int) wq.z, (int) av.z, nv.z), acc0);
                acc0 = fmaf(sw, (float) __dp4a((int) wq.w, (int) av.w, nv.w), acc0);
            }
            if (TN == 8) {
                const uint4 wq = *(const uint4 *) (wrow + (g * 8 + l + 4) * 16);
                const uint4 av = *(const uint4 *) (rec + (g * 8 + l + 4) * 16);
                int4 nv = make_int4(0, 0, 0, 0);
                if (HASOFF) nv = *(const int4 *) (rec + a.off_n + (g * 8 + l + 4) * 16);
                acc1 = fmaf(sx, (float) __dp4a((int) wq.x, (int) av.x, nv.x), acc1);
                acc1 = fmaf(sy, (float) __dp4a((int) wq.y, (int) av.y, nv.y), acc1);
                acc1 = fmaf(sz, (float) __dp4a((int) wq.z, (int) av.z, nv.z), acc1);
                acc1 = fmaf(sw, (float) __dp4a((int) wq.w, (int) av.w, nv.w), acc1);
            }
            if (HASM && l == 0) {
                const float4 mw = *(const float4 *) (mwr + g * 4);
                const float4 sa = *(const float4 *) (rec + a.off_s + g * 16);
                summ = fmaf(mw.x, sa.x, summ); summ = fmaf(mw.y, sa.y, summ); summ = fmaf(mw.z, sa.z, summ); summ = fmaf(mw.w, sa.w, summ);
            }
        }
        // hsum_float_8: (a_l + a_{l+4}), then the pairs 2 apart, then 1 apart (ggml.c:611-617)
        float v = TN == 8 ? __fadd_rn(acc0, acc1) : __fadd_rn(acc0, __shfl_xor_sync(FULLMASK, acc0, 4));
        v = __fadd_rn(v, __shfl_xor_sync(FULLMASK, v, 2));
        v = __fadd_rn(v, __shfl_xor_sync(FULLMASK, v, 1));
        if (HASM) v = __fadd_rn(v, summ);
        if (owner) {
            switch (a.epi) {
            case SK_EPI_STORE:
                a.out[(size_t) tok * a.ld_out + row] = a.bias[0] ? __fadd_rn(pbias, v) : v;
                break;
            case SK_EPI_QKV: {
                const int mat = mat_of(row), rr = row - mat * a.rows_per;
                const float t = __fadd_rn(pbias, v);
                if (mat == 0) a.out[(size_t) tok * a.ld_out + rr] = __fmul_rn(t, a.qscale);
                else {
                    int stream, pos, T; bg_row_info(a.mode, a.n, n_past, tok, stream, pos, T);
                    (mat == 1 ? a.kcache : a.vcache)[(size_t) stream * a.stream_stride + (size_t) pos * a.rows_per + rr] = t;
                }
                break; }
            case SK_EPI_RESID:
                a.out[(size_t) tok * a.ld_out + row] = __fadd_rn(__fadd_rn(v, pbias), presid);
                break;
            default:                                           // GELU input; the table look-ups run as one batch after the loop
                s_g[tq * 32 + rl] = __fadd_rn(pbias, v);
                break;
            }
        }
    }
    if (a.pdl_trig == 1) sk_pdl_launch_dependents();
    if (a.epi == SK_EPI_GELUQ) {                               // the CTA's 32 rows are block blockIdx.x of the next record
        __syncthreads();
        for (int i = tid; i < TN * 32; i += nt) s_g[i] = bg_h2f(a.gelu[bg_f2h(s_g[i])]);
        __syncthreads();
        if (warp < TN) {
            const int tok = tokbase + warp, l8 = lane & 7;
            const float4 v = *(const float4 *) (s_g + warp * 32 + 4 * l8);
            const bool wr = tok < a.n && lane < 8;
            uint8_t * orec = a.act_out + (size_t) (tok < a.n ? tok : 0) * a.out_bytes;
            sk_quant_block<FMT>(v, blockIdx.x, l8, orec, a.out_off_n, a.out_off_d, a.out_off_s, a.code_off, wr);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// k_sk_attn: softmax(K q) V of one (head, token row), un-masked over T positions, d_kv = 64,
// T <= 1024; the 64 outputs are quantised into blocks 2h, 2h+1 of out_proj's activation record.
// Arithmetic order = k_attn / the persistent kernels (ggml_vec_dot_f32 lanes, the xor 16,8,4,1,2
// tree, fp16 exp table, double sum, the as-built scalar tail of the V product).
// grid = (n_head, rows), block = 1024: at T <= 512 every K row and every V row of the head is requested in ONE batch of loads.
// ---------------------------------------------------------------------------------------------
struct SkAttnArgs {
    const float * q; int ld_q;
    const float * kcache; const float * vcache; size_t stream_stride;
    uint8_t * act; int act_bytes, off_n, off_d, off_s, code_off;
    int n, mode; const DevState * st;
    const uint16_t * exp_tab;
};

template <int FMT>
__global__ void __launch_bounds__(SK_ANT, 1) k_sk_attn(const __grid_constant__ SkAttnArgs a) {
    constexpr int NW = SK_ANT / 32;
    __shared__ __align__(16) float sc[1024];
    __shared__ __align__(16) float red[32 * SK_DK];
    __shared__ __align__(16) float tailv[31 * SK_DK];
    __shared__ __align__(16) float s_out[SK_DK];
    __shared__ double sredA[NW];
    __shared__ float sredF[NW];
    const int h = blockIdx.x, row = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    int stream, pos, T; bg_row_info(a.mode, a.n, a.st->n_past, row, stream, pos, T);     // n_past: constant of the whole graph
    const float * Kb = a.kcache + (size_t) stream * a.stream_stride + (size_t) h * SK_DK;
    const float * Vb = a.vcache + (size_t) stream * a.stream_stride + (size_t) h * SK_DK;
    // rows cached by earlier evals can be pulled towards the L2 while the q,k,v projection is still running
    {
        const int told = a.st->n_past;
        for (int t = tid; t < told; t += SK_ANT) {
            asm volatile("prefetch.global.L2 [%0];" :: "l"(Kb + (size_t) t * SK_D));
            asm volatile("prefetch.global.L2 [%0];" :: "l"(Kb + (size_t) t * SK_D + 32));
            asm volatile("prefetch.global.L2 [%0];" :: "l"(Vb + (size_t) t * SK_D));
            asm volatile("prefetch.global.L2 [%0];" :: "l"(Vb + (size_t) t * SK_D + 32));
        }
    }
    sk_pdl_launch_dependents();
    sk_pdl_wait();
    const float q0 = __ldcg(a.q + (size_t) row * a.ld_q + h * SK_DK + lane), q1 = __ldcg(a.q + (size_t) row * a.ld_q + h * SK_DK + 32 + lane);
    // ---- scores: 16 K rows per warp pass
#pragma unroll 1
    for (int tb = warp; tb < T; tb += NW * 16) {
        float kr[16][2];
#pragma unroll
        for (int u = 0; u < 16; u++) {
            const int t = tb + u * NW;
            kr[u][0] = (t < T) ? __ldcg(Kb + (size_t) t * SK_D + lane) : 0.0f;
            kr[u][1] = (t < T) ? __ldcg(Kb + (size_t) t * SK_D + 32 + lane) : 0.0f;
        }
        float s[16];
#pragma unroll
        for (int u = 0; u < 16; u++) { float x = 0.0f; x = fmaf(kr[u][0], q0, x); x = fmaf(kr[u][1], q1, x); s[u] = x; }
        // 16 reduce trees (xor 16, 8, 4, 1, 2: GGML_F32x8_REDUCE) as one transposing butterfly
        float a8[8], a4[4], a2[2];
        const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4, b0 = lane & 1;
#pragma unroll
        for (int i = 0; i < 8; i++) { const float mine = b4 ? s[8 + i] : s[i], send = b4 ? s[i] : s[8 + i]; a8[i] = __fadd_rn(mine, __shfl_xor_sync(FULLMASK, send, 16)); }
#pragma unroll
        for (int i = 0; i < 4; i++) { const float mine = b3 ? a8[4 + i] : a8[i], send = b3 ? a8[i] : a8[4 + i]; a4[i] = __fadd_rn(mine, __shfl_xor_sync(FULLMASK, send, 8)); }
#pragma unroll
        for (int i = 0; i < 2; i++) { const float mine = b2 ? a4[2 + i] : a4[i], send = b2 ? a4[i] : a4[2 + i]; a2[i] = __fadd_rn(mine, __shfl_xor_sync(FULLMASK, send, 4)); }
        const float mine = b0 ? a2[1] : a2[0], send = b0 ? a2[0] : a2[1];
        const float a1 = __fadd_rn(mine, __shfl_xor_sync(FULLMASK, send, 1));
        const float dot = __fadd_rn(a1, __shfl_xor_sync(FULLMASK, a1, 2));
        const int u = ((lane >> 1) & 14) | (lane & 1);              // bits 4,3,2 -> u bits 3,2,1; bit 0 -> u bit 0
        const int t = tb + u * NW;
        if (t < T && !(lane & 2)) sc[t] = dot;
    }
    // V rows of the scalar tail, staged so the tail is not a chain of global loads
    const int np = T & ~31;
    for (int i = tid; i < (T - np) * SK_DK; i += SK_ANT) tailv[i] = __ldcg(Vb + (size_t) (np + i / SK_DK) * SK_D + (i % SK_DK));
    __syncthreads();
    // ---- softmax over sc[0..T): max, fp16-table exp, sum in double (exact for fp16 values), scale
    {
        const float x0 = tid < T ? sc[tid] : -INFINITY;
        float mx = x0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(FULLMASK, mx, o));
        if (lane == 0) sredF[warp] = mx;
        __syncthreads();
        mx = sredF[lane];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(FULLMASK, mx, o));
        float e0 = 0.f;
        if (tid < T) e0 = bg_h2f(a.exp_tab[bg_f2h(__fsub_rn(x0, mx))]);
        double sm = (double) e0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sm += __shfl_xor_sync(FULLMASK, sm, o);
        if (lane == 0) sredA[warp] = sm;
        __syncthreads();
        double tot = sredA[lane];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(FULLMASK, tot, o);
        const float inv = (float) (1.0 / tot);
        if (tid < T) sc[tid] = __fmul_rn(e0, inv);
    }
    __syncthreads();
    // ---- V: thread (r = tid / 32, 2 columns): running sum r over t = r, r + 32, ... < np, 16 rows in flight
    {
        const int vr = tid >> 5, vc = tid & 31;
        const float * vp = Vb + (size_t) vr * SK_D + 2 * vc;
        float2 acc = make_float2(0.f, 0.f);
#pragma unroll 1
        for (int s0 = 0; s0 < np; s0 += 512) {
            float2 vv[16];
#pragma unroll
            for (int k = 0; k < 16; k++) {
                const int t = s0 + 32 * k + vr;
                vv[k] = (t < np) ? __ldcg((const float2 *) (vp + (size_t) (s0 + 32 * k) * SK_D)) : make_float2(0.f, 0.f);
            }
#pragma unroll
            for (int k = 0; k < 16; k++) {
                const int t = s0 + 32 * k + vr;
                if (t < np) {
                    const float pw = sc[t];
                    acc.x = fmaf(vv[k].x, pw, acc.x); acc.y = fmaf(vv[k].y, pw, acc.y);
                }
            }
        }
        *(float2 *) (red + vr * SK_DK + 2 * vc) = acc;
    }
    __syncthreads();
    if (tid < SK_DK) {
        float x0[8];
#pragma unroll
        for (int l8 = 0; l8 < 8; l8++) {
            const float a02 = __fadd_rn(red[(0 * 8 + l8) * SK_DK + tid], red[(2 * 8 + l8) * SK_DK + tid]);
            const float a13 = __fadd_rn(red[(1 * 8 + l8) * SK_DK + tid], red[(3 * 8 + l8) * SK_DK + tid]);
            x0[l8] = __fadd_rn(a02, a13);
        }
        const float t0 = __fadd_rn(x0[0], x0[4]), t1 = __fadd_rn(x0[1], x0[5]);
        const float t2 = __fadd_rn(x0[2], x0[6]), t3 = __fadd_rn(x0[3], x0[7]);
        float sumf = __fadd_rn(__fadd_rn(t0, t1), __fadd_rn(t2, t3));
        const int nv = np + ((T - np) & ~3);
        int t = np;
#pragma unroll 1
        for (; t < nv; t++) sumf = __fadd_rn(sumf, __fmul_rn(tailv[(t - np) * SK_DK + tid], sc[t]));
#pragma unroll 1
        for (; t < T;  t++) sumf = fmaf(tailv[(t - np) * SK_DK + tid], sc[t], sumf);
        s_out[tid] = sumf;
    }
    __syncthreads();
    if (warp == 0) {                                             // lanes 0..15: block lane / 8, word lane % 8
        const int blk = (lane >> 3) & 1, l = lane & 7;
        const float4 v = *(const float4 *) (s_out + blk * 32 + 4 * l);
        sk_quant_block<FMT>(v, h * 2 + blk, l, a.act + (size_t) row * a.act_bytes, a.off_n, a.off_d, a.off_s, a.code_off, lane < 16);
    }
}
