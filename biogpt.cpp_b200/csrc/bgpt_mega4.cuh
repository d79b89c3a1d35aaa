// bgpt_mega4.cuh -- persistent decode kernel, generation 4 (quantised weights, BioGPT-base
// shapes: d_model 1024, 16 heads of 64, d_ff 4096, n_positions <= 1024).
//
// Same arithmetic, same phases as bgpt_mega.cuh (k_mega, "v3"), but the dependent chain -- which
// is all a batch-1 decode step consists of -- is rebuilt around three ideas measured to matter
// (profiles/README.md, r1a): v3 spends 28 us per layer where the chain of L2 round trips
// (~250 cycles each) and block-wide steps needs ~6.
//
//  1. No grid barrier.  A phase's results travel as self-validating 8-byte words
//     { f32/u32 payload, u32 tag } (tag = launch serial, one dedicated buffer per layer and
//     exchange, so a word is valid iff its tag is the current launch's).  The producer's single
//     st.b64 is atomic, the consumer polls its own two words with one 16-byte volatile load: one
//     store flight + one load flight instead of {drain stores, red.release, poll, bar.sync, load}.
//     Words that all 148 CTAs read are written to R = 8 replicas (CTA c reads replica c % 8), so
//     one L2 line is polled by <= 19 CTAs instead of 148.
//  2. Weights never sit on the chain.  Every CTA's weight tile of phase n+3 is in flight while
//     phase n computes: one elected thread issues cp.async.bulk (TMA bulk copy, global -> shared)
//     into a 4-slot ring guarded by mbarriers (expect_tx / complete_tx); rows of a CTA are
//     contiguous, so a tile is 1-2 bulk copies.  The lm_head streams through the same ring.
//  3. Block-wide steps are single-sync: LayerNorm runs on 2 elements per thread in registers
//     (warp shuffle + one 16-entry shared exchange per reduction), the activation quantiser on the
//     same registers; attention pre-loads its K rows into registers BEFORE q arrives and issues
//     the V loads before the softmax, so the L2 latency of the KV cache overlaps the waits.
//
//  phase (per layer)                 CTAs         consumes            publishes
//  P1 LN0 + q,k,v                    all (20/21 rows)   E5[l-1] x        E1 q | k | v (+ KV cache)
//  P2 attention                      32 = (head, 32-col half)  E1       E2 quantised block (R)
//  P3 out_proj + bias + residual     all (6/7 rows)     E2               E3 x1 (R)
//  P4 LN1 + fc1 + bias + GELU        128 (32 rows = one block)  E3       E4 quantised block (R)
//  P5 fc2 + bias + residual          all (6/7 rows)     E4               E5 x (R)
//  final LN + lm_head                all (286/287 rows, 28-row tiles)  E5[L-1]   logits, argmax candidates
//
//  4. The whole kernel is ONE loop over matmul tiles (4 per layer + the lm_head tiles) with a single
//     lexical copy of every stage (poll, LayerNorm+quantise, phase A, phase B, publish, attention):
//     the first v4 build was 136 KB of SASS (v3: 272 KB) against a 32 KB instruction cache and
//     spent ~3000 cycles per phase re-fetching cold code from L2 (I$ stays warm across launches,
//     so a kernel that fits is never cold).
//
// The dot products are bgpt_mega.cuh's phase A (exact integer dp4a per 4-element group) and
// phase B (the 8 running sums in block order, hsum_float_8), reading the weight tile from
// shared memory; every float operation and its order is unchanged, so the logits remain the
// reference's bits (tests/test_gpu_eval.py compares v4 against the oracle and against v3).
#pragma once
#include "bgpt_mega.cuh"

#define M4_NT 512
#define M4_NW (M4_NT / 32)
#define M4_D 1024
#define M4_FF 4096
#define M4_DK 64
#define M4_NH 16
#define M4_R 8            // replicas of an all-to-all exchange buffer
#define M4_NSLOT 4        // weight ring slots (mbarriers)
#define M4_LMRT 32        // lm_head rows per tile
#define M4_NB_D (M4_D / 32)
#define M4_NB_F (M4_FF / 32)
// exchange words per layer
#define M4_E1 0
#define M4_E2 (3 * M4_D)
#define M4_E3 (M4_E2 + M4_R * M4_NB_D * 10)
#define M4_E4 (M4_E3 + M4_R * M4_D)
#define M4_E5 (M4_E4 + M4_R * M4_NB_F * 10)
#define M4_LW (M4_E5 + M4_R * M4_D)

#define M4_MAXL 24        // layer table travels in the kernel parameters (constant bank): no pointer chase, no L2 miss on the chain
struct M4Params {
    MegaParams b;
    MegaLayer layers[M4_MAXL];
    unsigned long long * xch;      // [2][M4_LW] tagged words (layer parity): 0.5 MB, always L2-resident
    unsigned int tag;              // launch serial << 6 (never 0); a word's tag = tag | (layer + 1)
    int prof_cta;                  // CTA whose phase stamps bgpt_cuda_debug_read_prof returns (BGPT_MEGA_PROF_CTA)
    long long * trace;             // BGPT_MEGA_PROF: [nC][prof_n] clock64 stamps of every CTA, then [nC][4] clock calibration
    int prof_n;
    int nslot, slot_bytes;
    unsigned poll_sleep;           // nanosleep between unsuccessful polls (0: spin)
    int sm_w, sm_act0, sm_act1, sm_p, sm_s, sm_m, sm_x, sm_x1, sm_sc, sm_red, sm_tail, sm_total;
};

// ---- tagged words ------------------------------------------------------------------------------
__device__ __forceinline__ void m4_put(unsigned long long * p, uint32_t payload, uint32_t tag) {
    const unsigned long long w = ((unsigned long long) tag << 32) | payload;
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" :: "l"(p), "l"(w) : "memory");
}
__device__ __forceinline__ void m4_put_rep(unsigned long long * p, int rep_stride, uint32_t payload, uint32_t tag) {
#pragma unroll
    for (int r = 0; r < M4_R; r++) m4_put(p + (size_t) r * rep_stride, payload, tag);
}
// two consecutive words (16-byte aligned); spins until both carry the tag
__device__ __forceinline__ void m4_poll2(const unsigned long long * p, uint32_t tag, uint32_t & a, uint32_t & b, unsigned sleep_ns) {
    unsigned long long w0, w1;
    for (;;) {
        asm volatile("ld.volatile.global.v2.u64 {%0,%1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(p) : "memory");
        if ((uint32_t) (w0 >> 32) == tag && (uint32_t) (w1 >> 32) == tag) break;
        if (sleep_ns) __nanosleep(sleep_ns);
    }
    a = (uint32_t) w0; b = (uint32_t) w1;
}

// ---- mbarrier + bulk copy ------------------------------------------------------------------------
__device__ __forceinline__ uint32_t m4_s32(const void * p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void m4_mbar_init(uint64_t * bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(m4_s32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void m4_mbar_expect(uint64_t * bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(m4_s32(bar)), "r"(bytes) : "memory");
}
// weights are read once per token and the set (194+ MB) exceeds the L2: stream them evict-first so they do not push the
// kernel's own instructions, the lookup tables, the KV prefetches and the exchange words out of the L2
__device__ __forceinline__ uint64_t m4_policy_evict_first() {
    uint64_t pol; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol)); return pol;
}
__device__ __forceinline__ void m4_bulk_g2s(void * dst, const void * src, uint32_t bytes, uint64_t * bar, uint64_t pol) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 :: "r"(m4_s32(dst)), "l"(src), "r"(bytes), "r"(m4_s32(bar)), "l"(pol) : "memory");
}
__device__ __forceinline__ void m4_mbar_wait(uint64_t * bar, uint32_t parity) {
    uint32_t done;
    const uint32_t a = m4_s32(bar);
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(a), "r"(parity) : "memory");
    } while (!done);
}

__device__ __forceinline__ double m4_warp_sum_f64(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULLMASK, v, o);
    return v;
}
__device__ __forceinline__ double m4_tree16(const double * s) {
    const double a0 = s[0] + s[1], a1 = s[2] + s[3], a2 = s[4] + s[5], a3 = s[6] + s[7];
    const double a4 = s[8] + s[9], a5 = s[10] + s[11], a6 = s[12] + s[13], a7 = s[14] + s[15];
    return ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}

// ---- LayerNorm + affine + activation quantiser on registers ------------------------------------
// thread t owns elements 2t, 2t+1 of the 1024-wide row.  Same operations as bg_ln_row_pre /
// bg_row_to_record (ggml.c:11403-11420, 1166-1203, 1403-1450); the double sums are combined in a
// different (parallel) order, which only matters when a double rounding lands on a float tie.
// Ends WITHOUT a barrier: the caller synchronises before the record is read.
template <int FMT, bool PROF>
__device__ __forceinline__ void m4_ln_quant(float v0, float v1, float2 lw, float2 lb, float eps,
                                            double * sredA, double * sredB, uint8_t * rec, int off_d, int off_s, long long * stamp) {
    constexpr bool Q81 = (FMT == BG_Q4_1 || FMT == BG_Q5_1);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float mean, variance;
    long long t4 = 0, t6 = 0, t7 = 0, t8 = 0;
    double s = m4_warp_sum_f64((double) v0 + (double) v1);
    if (lane == 0) sredA[warp] = s;
    __syncthreads();
    if (PROF && stamp) t4 = clock64();
    mean = (float) (m4_tree16(sredA) * (1.0 / M4_D));
    const float e0 = __fsub_rn(v0, mean), e1 = __fsub_rn(v1, mean);
    double s2 = m4_warp_sum_f64((double) __fmul_rn(e0, e0) + (double) __fmul_rn(e1, e1));
    if (lane == 0) sredB[warp] = s2;
    if (PROF && stamp) t6 = clock64();
    __syncthreads();
    if (PROF && stamp) t7 = clock64();
    variance = (float) (m4_tree16(sredB) * (1.0 / M4_D));
    const float d0 = __fsub_rn(v0, mean), d1 = __fsub_rn(v1, mean);
    const float scale = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(variance, eps)));
    if (PROF && stamp) t8 = clock64();
    const float y0 = __fadd_rn(__fmul_rn(lw.x, __fmul_rn(d0, scale)), lb.x);
    const float y1 = __fadd_rn(__fmul_rn(lw.y, __fmul_rn(d1, scale)), lb.y);
    // quantise: a 32-element block = 16 consecutive threads
    float amax = fmaxf(fabsf(y0), fabsf(y1));
    amax = fmaxf(amax, __shfl_xor_sync(FULLMASK, amax, 1));
    amax = fmaxf(amax, __shfl_xor_sync(FULLMASK, amax, 2));
    amax = fmaxf(amax, __shfl_xor_sync(FULLMASK, amax, 4));
    amax = fmaxf(amax, __shfl_xor_sync(FULLMASK, amax, 8));
    const float d  = __fdiv_rn(amax, 127.0f);
    const float id = (amax != 0.0f) ? __fdiv_rn(127.0f, amax) : 0.0f;
    const int q0 = __float2int_rn(__fmul_rn(y0, id)), q1 = __float2int_rn(__fmul_rn(y1, id));
    const uint32_t hw = ((uint32_t) q0 & 0xFFu) | (((uint32_t) q1 & 0xFFu) << 8);
    const int ps = q0 + q1;
    const uint32_t hw_p = __shfl_xor_sync(FULLMASK, hw, 1);
    const int b = tid >> 4, g = b >> 2, i = b & 3;
    if ((tid & 1) == 0) {
        const int l = (tid & 15) >> 1;
        ((uint32_t *) rec)[(g * 8 + l) * 4 + i] = hw | (hw_p << 16);
    }
    if (Q81) {
        int stot = ps;
        stot += __shfl_xor_sync(FULLMASK, stot, 1);
        stot += __shfl_xor_sync(FULLMASK, stot, 2);
        stot += __shfl_xor_sync(FULLMASK, stot, 4);
        stot += __shfl_xor_sync(FULLMASK, stot, 8);
        if ((tid & 15) == 0) { ((float *) (rec + off_d))[b] = d; ((float *) (rec + off_s))[b] = __fmul_rn(d, (float) stot); }
    } else if ((tid & 15) == 0) { ((float *) (rec + off_d))[b] = bg_h2f(bg_f2h(d)); ((float *) (rec + off_s))[b] = 0.0f; }
    if (PROF && stamp && tid == 0) { stamp[4] = t4; stamp[6] = t6; stamp[7] = t7; stamp[8] = t8; }
    if (PROF && stamp && tid == 256) stamp[11] = t6;
}

// ---- one finished 32-element block (f32 in shared memory) -> 10 exchange words, R replicas -----
// executed by one full warp; words 0..7 = the int8 codes (word l = elements 4l..4l+3), 8 = d, 9 = s
template <int FMT>
__device__ __forceinline__ void m4_quant_publish(const float * s_v, unsigned long long * dst /*replica 0, block base*/, int rep_stride, uint32_t tag) {
    constexpr bool Q81 = (FMT == BG_Q4_1 || FMT == BG_Q5_1);
    const int lane = threadIdx.x & 31, l = lane & 7;
    const float4 v = *(const float4 *) (s_v + 4 * l);
    float amax = fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w)));
    amax = fmaxf(amax, __shfl_xor_sync(FULLMASK, amax, 1));
    amax = fmaxf(amax, __shfl_xor_sync(FULLMASK, amax, 2));
    amax = fmaxf(amax, __shfl_xor_sync(FULLMASK, amax, 4));
    const float d  = __fdiv_rn(amax, 127.0f);
    const float id = (amax != 0.0f) ? __fdiv_rn(127.0f, amax) : 0.0f;
    const int q0 = __float2int_rn(__fmul_rn(v.x, id)), q1 = __float2int_rn(__fmul_rn(v.y, id));
    const int q2 = __float2int_rn(__fmul_rn(v.z, id)), q3 = __float2int_rn(__fmul_rn(v.w, id));
    int stot = q0 + q1 + q2 + q3;
    stot += __shfl_xor_sync(FULLMASK, stot, 1);
    stot += __shfl_xor_sync(FULLMASK, stot, 2);
    stot += __shfl_xor_sync(FULLMASK, stot, 4);
    const uint32_t word = ((uint32_t) q0 & 0xFFu) | (((uint32_t) q1 & 0xFFu) << 8) | (((uint32_t) q2 & 0xFFu) << 16) | (((uint32_t) q3 & 0xFFu) << 24);
    const uint32_t dbits = __float_as_uint(Q81 ? d : bg_h2f(bg_f2h(d)));
    const uint32_t sbits = __float_as_uint(Q81 ? __fmul_rn(d, (float) stot) : 0.0f);
    // lane k < 10 of every 10-lane group holds word k: lanes 0..29 cover 3 replicas per round
    const int k = lane % 10, rr = lane / 10;
    const uint32_t wk = __shfl_sync(FULLMASK, word, k & 7);
    const uint32_t mine = k < 8 ? wk : (k == 8 ? dbits : sbits);
    if (lane < 30) {
#pragma unroll
        for (int r0 = 0; r0 < M4_R; r0 += 3) { const int r = r0 + rr; if (r < M4_R) m4_put(dst + (size_t) r * rep_stride + k, mine, tag); }
    }
}

// scatter one exchange word (block b, word k) into the activation record in shared memory
__device__ __forceinline__ void m4_scatter_word(uint8_t * rec, int off_d, int off_s, int w, uint32_t v) {
    const int b = w / 10, k = w - b * 10;
    if (k < 8) {
        const int g = b >> 2, i = b & 3;
        ((uint32_t *) rec)[(g * 8 + k) * 4 + i] = v;
    } else if (k == 8) ((float *) (rec + off_d))[b] = __uint_as_float(v);
    else               ((float *) (rec + off_s))[b] = __uint_as_float(v);
}

// ---- matmul over a weight tile in shared memory -----------------------------------------------
struct M4MM { int G, gsh, stride, off_qh, off_d, off_m, off_n, off_dd, off_s; };

// phase B: thread c = 8*row + l walks running sum l of its row in block order; returns the
// finished dot of the row (meaningful in the row's owner thread, l == 0).  rt*8 <= M4_NT.
template <int FMT>
__device__ __forceinline__ float m4_phase_b(const M4MM & D, int rt, const uint8_t * s_act,
                                            const float * s_p, const float * s_s, const float * s_m) {
    constexpr bool HASM = (FMT == BG_Q4_1 || FMT == BG_Q5_1);
    const int nbp = D.G * 4, PS = nbp + 4;
    if ((int) (threadIdx.x & ~31u) >= rt * 8) return 0.0f;      // whole warp idle
    const int c = threadIdx.x;
    const int cc = c < rt * 8 ? c : rt * 8 - 1;
    const int row = cc >> 3, l = cc & 7;
    const float * pp = s_p + (size_t) cc * PS;
    const float * ss = s_s + (size_t) row * nbp;
    float acc = 0.0f, summ = 0.0f;
#pragma unroll 4
    for (int g = 0; g < D.G; g++) {
        const float4 pv = *(const float4 *) (pp + 4 * g);
        const float4 sv = *(const float4 *) (ss + 4 * g);
        acc = fmaf(sv.x, pv.x, acc); acc = fmaf(sv.y, pv.y, acc);
        acc = fmaf(sv.z, pv.z, acc); acc = fmaf(sv.w, pv.w, acc);
        if (HASM && l == 0) {
            const float4 mv = *(const float4 *) (s_m + (size_t) row * nbp + 4 * g);
            const float4 sa = *(const float4 *) (s_act + D.off_s + g * 16);
            summ = fmaf(mv.x, sa.x, summ); summ = fmaf(mv.y, sa.y, summ);
            summ = fmaf(mv.z, sa.z, summ); summ = fmaf(mv.w, sa.w, summ);
        }
    }
    float r = __fadd_rn(acc, __shfl_xor_sync(FULLMASK, acc, 4));
    r = __fadd_rn(r, __shfl_xor_sync(FULLMASK, r, 2));
    r = __fadd_rn(r, __shfl_xor_sync(FULLMASK, r, 1));
    if (HASM) r = __fadd_rn(r, summ);
    return r;
}

// ---- K = 1024 tiles: one warp per row, no shared-memory round trip ------------------------------
// lane = (g = lane >> 2, j = lane & 3): the lane's uint4 holds, for the 4 blocks of group g, the
// 4-byte element groups of running sums j (low nibbles) and j+4 (high nibbles).  Phase A is the
// lane's 8 exact integer dots; phase B -- acc_l = fma(s_b, p_b, acc_l) over the 32 blocks IN ORDER
// -- runs as a relay along g: lane (g, j) continues the sums j and j+4 that lane (g-1, j) hands
// over with one shuffle.  Two rows (warp, warp + 16) are relayed together for ILP.  The finished
// dot (hsum_float_8 order, + summs for Q4_1/Q5_1) lands in lane 28 (g = 7, j = 0).
struct M4Act { uint4 a0, a1; int4 n0, n1; float4 da, sa; };

template <int FMT>
__device__ __forceinline__ void m4_load_act(M4Act & A, const uint8_t * rec, const M4MM & D, int code_off, int g, int j) {
    constexpr bool HASOFF = (FMT == BG_Q4_0 || FMT == BG_Q5_0);
    constexpr bool HASM   = (FMT == BG_Q4_1 || FMT == BG_Q5_1);
    A.a0 = *(const uint4 *) (rec + (g * 8 + j) * 16);
    A.a1 = *(const uint4 *) (rec + (g * 8 + j + 4) * 16);
    A.n0 = make_int4(0, 0, 0, 0); A.n1 = make_int4(0, 0, 0, 0);
    if (HASOFF) {
        A.n0.x = -code_off * __dp4a((int) A.a0.x, 0x01010101, 0); A.n0.y = -code_off * __dp4a((int) A.a0.y, 0x01010101, 0);
        A.n0.z = -code_off * __dp4a((int) A.a0.z, 0x01010101, 0); A.n0.w = -code_off * __dp4a((int) A.a0.w, 0x01010101, 0);
        A.n1.x = -code_off * __dp4a((int) A.a1.x, 0x01010101, 0); A.n1.y = -code_off * __dp4a((int) A.a1.y, 0x01010101, 0);
        A.n1.z = -code_off * __dp4a((int) A.a1.z, 0x01010101, 0); A.n1.w = -code_off * __dp4a((int) A.a1.w, 0x01010101, 0);
    }
    A.da = *(const float4 *) (rec + D.off_dd + g * 16);
    A.sa = make_float4(0.f, 0.f, 0.f, 0.f);
    if (HASM) A.sa = *(const float4 *) (rec + D.off_s + g * 16);
}

// phase A of one row for this lane: 8 products, 4 scales (and 4 mins)
template <int FMT>
__device__ __forceinline__ void m4_row_products(const uint8_t * wrow, const M4MM & D, const M4Act & A, int g, int j, float4 & P0, float4 & P1, float4 & S, float4 & Mv) {
    constexpr bool IS8   = (FMT == BG_Q8_0);
    constexpr bool HASQH = (FMT == BG_Q5_0 || FMT == BG_Q5_1);
    constexpr bool HASM  = (FMT == BG_Q4_1 || FMT == BG_Q5_1);
    const int lane = g * 4 + j;                                   // position of the unit inside the row
    uint32_t lo[4], hi[4];
    if (IS8) {
        const uint4 w0 = *(const uint4 *) (wrow + ((g * 2 + 0) * 4 + j) * 16);
        const uint4 w1 = *(const uint4 *) (wrow + ((g * 2 + 1) * 4 + j) * 16);
        lo[0] = w0.x; lo[1] = w0.y; lo[2] = w0.z; lo[3] = w0.w;
        hi[0] = w1.x; hi[1] = w1.y; hi[2] = w1.z; hi[3] = w1.w;
    } else {
        const uint4 w0 = *(const uint4 *) (wrow + lane * 16);
        uint32_t qh = 0;
        if (HASQH) qh = *(const uint32_t *) (wrow + D.off_qh + lane * 4);
        const uint32_t ww[4] = { w0.x, w0.y, w0.z, w0.w };
#pragma unroll
        for (int i = 0; i < 4; i++) {
            lo[i] = ww[i] & 0x0F0F0F0Fu;
            hi[i] = (ww[i] >> 4) & 0x0F0F0F0Fu;
            if (HASQH) {
                const uint32_t hb = (qh >> (8 * i)) & 0xFFu;
                lo[i] |= bg_spread4(hb & 0xFu);
                hi[i] |= bg_spread4(hb >> 4);
            }
        }
    }
    P0.x = (float) __dp4a((int) lo[0], (int) A.a0.x, A.n0.x); P1.x = (float) __dp4a((int) hi[0], (int) A.a1.x, A.n1.x);
    P0.y = (float) __dp4a((int) lo[1], (int) A.a0.y, A.n0.y); P1.y = (float) __dp4a((int) hi[1], (int) A.a1.y, A.n1.y);
    P0.z = (float) __dp4a((int) lo[2], (int) A.a0.z, A.n0.z); P1.z = (float) __dp4a((int) hi[2], (int) A.a1.z, A.n1.z);
    P0.w = (float) __dp4a((int) lo[3], (int) A.a0.w, A.n0.w); P1.w = (float) __dp4a((int) hi[3], (int) A.a1.w, A.n1.w);
    const uint2 dh = *(const uint2 *) (wrow + D.off_d + g * 8);
    S.x = __fmul_rn(bg_h2f((uint16_t) (dh.x & 0xFFFF)), A.da.x); S.y = __fmul_rn(bg_h2f((uint16_t) (dh.x >> 16)), A.da.y);
    S.z = __fmul_rn(bg_h2f((uint16_t) (dh.y & 0xFFFF)), A.da.z); S.w = __fmul_rn(bg_h2f((uint16_t) (dh.y >> 16)), A.da.w);
    Mv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (HASM) {
        const uint2 mh = *(const uint2 *) (wrow + D.off_m + g * 8);
        Mv.x = bg_h2f((uint16_t) (mh.x & 0xFFFF)); Mv.y = bg_h2f((uint16_t) (mh.x >> 16));
        Mv.z = bg_h2f((uint16_t) (mh.y & 0xFFFF)); Mv.w = bg_h2f((uint16_t) (mh.y >> 16));
    }
}

// rows `ra` and `rb` of the tile (rb may repeat ra); returns their dots in lane 28
template <int FMT>
__device__ __forceinline__ float2 m4_two_rows(const uint8_t * wt, int ra, int rb, const M4MM & D, const M4Act & A) {
    constexpr bool HASM = (FMT == BG_Q4_1 || FMT == BG_Q5_1);
    const int lane = threadIdx.x & 31, g = lane >> 2, j = lane & 3;
    float4 Pa0, Pa1, Sa, Ma, Pb0, Pb1, Sb, Mb;
    m4_row_products<FMT>(wt + (size_t) ra * D.stride, D, A, g, j, Pa0, Pa1, Sa, Ma);
    m4_row_products<FMT>(wt + (size_t) rb * D.stride, D, A, g, j, Pb0, Pb1, Sb, Mb);
    float a0 = 0.f, a1 = 0.f, am = 0.f, b0 = 0.f, b1 = 0.f, bm = 0.f;
#pragma unroll
    for (int step = 0; step < 8; step++) {
        if (step > 0) {
            const float ia0 = __shfl_up_sync(FULLMASK, a0, 4), ia1 = __shfl_up_sync(FULLMASK, a1, 4);
            const float ib0 = __shfl_up_sync(FULLMASK, b0, 4), ib1 = __shfl_up_sync(FULLMASK, b1, 4);
            float iam = 0.f, ibm = 0.f;
            if (HASM) { iam = __shfl_up_sync(FULLMASK, am, 4); ibm = __shfl_up_sync(FULLMASK, bm, 4); }
            if (g == step) { a0 = ia0; a1 = ia1; b0 = ib0; b1 = ib1; am = iam; bm = ibm; }
        }
        if (g == step) {
            a0 = fmaf(Sa.x, Pa0.x, a0); a1 = fmaf(Sa.x, Pa1.x, a1); b0 = fmaf(Sb.x, Pb0.x, b0); b1 = fmaf(Sb.x, Pb1.x, b1);
            a0 = fmaf(Sa.y, Pa0.y, a0); a1 = fmaf(Sa.y, Pa1.y, a1); b0 = fmaf(Sb.y, Pb0.y, b0); b1 = fmaf(Sb.y, Pb1.y, b1);
            a0 = fmaf(Sa.z, Pa0.z, a0); a1 = fmaf(Sa.z, Pa1.z, a1); b0 = fmaf(Sb.z, Pb0.z, b0); b1 = fmaf(Sb.z, Pb1.z, b1);
            a0 = fmaf(Sa.w, Pa0.w, a0); a1 = fmaf(Sa.w, Pa1.w, a1); b0 = fmaf(Sb.w, Pb0.w, b0); b1 = fmaf(Sb.w, Pb1.w, b1);
            if (HASM) {
                am = fmaf(Ma.x, A.sa.x, am); bm = fmaf(Mb.x, A.sa.x, bm);
                am = fmaf(Ma.y, A.sa.y, am); bm = fmaf(Mb.y, A.sa.y, bm);
                am = fmaf(Ma.z, A.sa.z, am); bm = fmaf(Mb.z, A.sa.z, bm);
                am = fmaf(Ma.w, A.sa.w, am); bm = fmaf(Mb.w, A.sa.w, bm);
            }
        }
    }
    // hsum_float_8 in lanes 28..31 (j = 0..3): (acc_j + acc_{j+4}), then +2, then +1
    float ra_ = __fadd_rn(a0, a1), rb_ = __fadd_rn(b0, b1);
    ra_ = __fadd_rn(ra_, __shfl_xor_sync(FULLMASK, ra_, 2)); rb_ = __fadd_rn(rb_, __shfl_xor_sync(FULLMASK, rb_, 2));
    ra_ = __fadd_rn(ra_, __shfl_xor_sync(FULLMASK, ra_, 1)); rb_ = __fadd_rn(rb_, __shfl_xor_sync(FULLMASK, rb_, 1));
    if (HASM) { ra_ = __fadd_rn(ra_, am); rb_ = __fadd_rn(rb_, bm); }
    return make_float2(ra_, rb_);
}

// K = 4096 (fc2): the relay would cross 4 warps per row, so the products go through shared memory
// (phase A, all threads) and 8 threads per row walk the 128 blocks (phase B).  512 = 4 * 128: a
// thread's unit (g, j) = tid & 127 is the same for every row it handles, so its activation words
// stay in registers.
template <int FMT>
__device__ __forceinline__ void m4_phase_a(const M4MM & D, const uint8_t * wt, int rt, const uint8_t * rec, int code_off,
                                           float * s_p, float * s_s, float * s_m) {
    constexpr bool HASM = (FMT == BG_Q4_1 || FMT == BG_Q5_1);
    const int nbp = D.G * 4, PS = nbp + 4;
    const int g = (threadIdx.x & 127) >> 2, j = threadIdx.x & 3;
    M4Act A;
    m4_load_act<FMT>(A, rec, D, code_off, g, j);
#pragma unroll 1
    for (int row = threadIdx.x >> 7; row < rt; row += M4_NT / 128) {
        float4 P0, P1, S, Mv;
        m4_row_products<FMT>(wt + (size_t) row * D.stride, D, A, g, j, P0, P1, S, Mv);
        *(float4 *) (s_p + (size_t) (row * 8 + j) * PS + 4 * g) = P0;
        *(float4 *) (s_p + (size_t) (row * 8 + j + 4) * PS + 4 * g) = P1;
        if (j == 0) *(float4 *) (s_s + (size_t) row * nbp + 4 * g) = S;
        if (HASM && j == 1) *(float4 *) (s_m + (size_t) row * nbp + 4 * g) = Mv;
    }
}

#define M4_PK 12          // stamps per (layer, phase) in the trace
#define M4PROF(ph, k) do { if (PROF && P.trace && threadIdx.x == 0) P.trace[(size_t) blockIdx.x * P.prof_n + (l * 5 + (ph)) * M4_PK + (k)] = clock64(); } while (0)
// (globaltimer edge, clock64) pair: spins until the nanosecond timer ticks, so the pair is exact whatever its resolution
__device__ __forceinline__ void m4_calibrate(long long * out) {
    unsigned long long g0, g1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
    do { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1)); } while (g1 == g0);
    out[0] = (long long) g1; out[1] = clock64();
}

template <int FMT, bool PROF>
__global__ void __launch_bounds__(M4_NT, 1) k_mega4(const __grid_constant__ M4Params P) {
    const MegaParams & p = P.b;
    extern __shared__ __align__(16) uint8_t smem[];
    __shared__ __align__(8) uint64_t mbar[M4_NSLOT];
    __shared__ double sredA[M4_NW], sredB[M4_NW];
    __shared__ float sredF[M4_NW];
    __shared__ __align__(16) float s_blk[32];
    __shared__ __align__(16) float s_q[M4_DK], s_kn[M4_DK], s_vn[32];
    __shared__ float s_cv[M4_NW]; __shared__ int s_ci[M4_NW];
    __shared__ int s_tok;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nC = gridDim.x, cta = blockIdx.x;
    const uint32_t tag0 = P.tag;
    uint8_t * s_w = smem + P.sm_w;
    float * s_p = (float *) (smem + P.sm_p);
    float * s_s = (float *) (smem + P.sm_s);
    float * s_m = (float *) (smem + P.sm_m);
    float * s_x = (float *) (smem + P.sm_x);
    float * s_x1 = (float *) (smem + P.sm_x1);
    float * sc = (float *) (smem + P.sm_sc);
    float * red = (float *) (smem + P.sm_red);
    float * tailv = (float *) (smem + P.sm_tail);

    // rows of this CTA in every phase
    const unsigned uC = (unsigned) nC, uc = (unsigned) cta;        // all products below fit 32 bits
    const int qkv0 = (int) ((uc * 3u * M4_D) / uC), qkv1 = (int) (((uc + 1u) * 3u * M4_D) / uC);
    const int o0 = (int) ((uc * M4_D) / uC), o1 = (int) (((uc + 1u) * M4_D) / uC);
    const bool has_fc1 = cta < M4_NB_F;
    const int v0 = (int) ((uc * (unsigned) p.n_vocab) / uC), v1 = (int) (((uc + 1u) * (unsigned) p.n_vocab) / uC);
    const int n_lm = (v1 - v0 + M4_LMRT - 1) / M4_LMRT;
    const int n_lt = 4 * p.n_layer;                               // layer tiles
    const int n_tiles = n_lt + n_lm;
    const bool is_att = cta < 2 * M4_NH;
    const int att_h = cta >> 1, att_part = cta & 1;
    const int rep = cta % M4_R;

    // ---- weight ring: tile n = layer n>>2, phase n&3 (P1, P3, P4, P5), then the lm_head tiles
    struct TileSrc { const uint8_t * src0, * src1; uint32_t b0, b1; };
    auto describe_tile = [&](int n) -> TileSrc {                 // global pointer loads: call early, fire later
        TileSrc t{nullptr, nullptr, 0u, 0u};
        if (n >= n_tiles) return t;
        if (n < n_lt) {
            const MegaLayer & L = P.layers[n >> 2];
            const int k = n & 3;
            if (k == 0) {
                const int m0 = qkv0 / M4_D, e0 = min(qkv1, (m0 + 1) * M4_D);
                const uint8_t * W0 = m0 == 0 ? L.q_w : (m0 == 1 ? L.k_w : L.v_w);
                t.src0 = W0 + (size_t) (qkv0 - m0 * M4_D) * p.stride_d; t.b0 = (uint32_t) (e0 - qkv0) * p.stride_d;
                if (e0 < qkv1) { t.src1 = (m0 == 0 ? L.k_w : L.v_w); t.b1 = (uint32_t) (qkv1 - e0) * p.stride_d; }
            } else if (k == 2) {
                if (has_fc1) { t.src0 = L.fc1_w + (size_t) cta * 32 * p.stride_d; t.b0 = 32u * p.stride_d; }
            } else {
                const int st = k == 1 ? p.stride_d : p.stride_f;
                t.src0 = (k == 1 ? L.o_w : L.fc2_w) + (size_t) o0 * st; t.b0 = (uint32_t) (o1 - o0) * st;
            }
        } else {
            const int r = v0 + (n - n_lt) * M4_LMRT;
            t.src0 = p.lm_head + (size_t) r * p.stride_d; t.b0 = (uint32_t) min(M4_LMRT, v1 - r) * p.stride_d;
        }
        return t;
    };
    const uint64_t pol_w = m4_policy_evict_first();
    auto fire_tile = [&](int n, const TileSrc & t) {
        if (t.b0 + t.b1 == 0) return;
        const int slot = n % P.nslot;
        uint8_t * dst = s_w + (size_t) slot * P.slot_bytes;
        m4_mbar_expect(&mbar[slot], t.b0 + t.b1);
        m4_bulk_g2s(dst, t.src0, t.b0, &mbar[slot], pol_w);
        if (t.b1) m4_bulk_g2s(dst + t.b0, t.src1, t.b1, &mbar[slot], pol_w);
    };
    constexpr int ISSUER = M4_NT - 32;
    uint32_t wphase = 0;
    if (tid == 0) {
        for (int i = 0; i < P.nslot; i++) m4_mbar_init(&mbar[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == ISSUER) {
#pragma unroll 1
        for (int n = 0; n < P.nslot - 1; n++) fire_tile(n, describe_tile(n));
    }

    // L2 prefetch of the K/V slices an attention CTA reads in layer Ln and of Ln's small f32 vectors
    auto prefetch_layer = [&](int Ln) {
        if (Ln >= p.n_layer) return;
        if (is_att) {
            const float * kc = p.kcache + (size_t) Ln * p.n_positions * M4_D + att_h * M4_DK;
            const float * vc = p.vcache + (size_t) Ln * p.n_positions * M4_D + att_h * M4_DK + att_part * 32;
#pragma unroll 1
            for (int t = tid; t < p.n_past; t += M4_NT) {
                asm volatile("prefetch.global.L2 [%0];" :: "l"(kc + (size_t) t * M4_D));
                asm volatile("prefetch.global.L2 [%0];" :: "l"(kc + (size_t) t * M4_D + 32));
                asm volatile("prefetch.global.L2 [%0];" :: "l"(vc + (size_t) t * M4_D));
            }
        }
        const int who = nC - 1 - cta;                              // 16 vectors of <= 16 KB, one CTA each
        if (who < 16) {
            const float * const * vecs = (const float * const *) &P.layers[Ln].q_b;     // q_b .. fc2_b: 10 consecutive pointers
            if (who < 10) {
                const unsigned bytes = (who == 8 ? M4_FF : M4_D) * 4u;                  // fc1_b is the 9th
#pragma unroll 1
                for (unsigned off = tid * 128u; off < bytes; off += M4_NT * 128u)
                    asm volatile("prefetch.global.L2 [%0];" :: "l"((const uint8_t *) vecs[who] + off));
            }
        }
    };
    prefetch_layer(0);
    if (PROF && P.trace && tid == 0) m4_calibrate(P.trace + (size_t) nC * P.prof_n + 4 * cta);

    // ---- input token: given, or argmax over the candidates the previous launch left
    if (tid < 32) {
        int tok;
        if (p.use_cand == 1) {
            float best = -INFINITY; int bi = 0x7fffffff;
#pragma unroll 1
            for (int i = tid; i < p.n_cand; i += 32) {
                const float v = __ldcg(p.cand_val + i); const int ix = __ldcg(p.cand_idx + i);
                if (v > best || (v == best && ix < bi)) { best = v; bi = ix; }
            }
#pragma unroll 1
            for (int o = 16; o > 0; o >>= 1) {
                const float ov = __shfl_xor_sync(FULLMASK, best, o); const int oi = __shfl_xor_sync(FULLMASK, bi, o);
                if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
            }
            tok = bi == 0x7fffffff ? 0 : bi;
        } else tok = p.use_cand == 2 ? p.tok_imm : __ldcg(p.tok);
        if (tid == 0) {
            s_tok = tok;
            if (cta == 0 && p.log_slot >= 0) p.idlog[p.log_slot] = tok;
        }
    }
    __syncthreads();
    // ---- embedding: every CTA, elements 2t, 2t+1 in registers
    float xa, xb;
    {
        int tok = s_tok; tok = tok < 0 ? 0 : (tok >= p.n_vocab ? p.n_vocab - 1 : tok);
        int prow = p.n_past + 2; prow = prow >= p.n_pos_rows ? p.n_pos_rows - 1 : prow;
        const size_t rb = bg_file_row_bytes(FMT, M4_D);
        const uint8_t * tr = p.embed_tok + rb * (size_t) tok;
        const uint8_t * pr = p.embed_pos + rb * (size_t) prow;
        xa = __fadd_rn(__fmul_rn(bg_dequant_elem(FMT, tr, 2 * tid), p.emb_scale), bg_dequant_elem(FMT, pr, 2 * tid));
        xb = __fadd_rn(__fmul_rn(bg_dequant_elem(FMT, tr, 2 * tid + 1), p.emb_scale), bg_dequant_elem(FMT, pr, 2 * tid + 1));
    }

    const int pos = p.n_past, T = p.n_past + 1;
    float best = -INFINITY; int bi = 0x7fffffff;                  // lm_head argmax of this thread's rows
#pragma unroll 1
    for (int tn = 0; tn < n_tiles; tn++) {
        const bool lm = tn >= n_lt;
        const int kind = lm ? 4 : (tn & 3);                        // 0 P1, 1 P3, 2 P4, 3 P5, 4 lm_head
        const int l = lm ? p.n_layer : (tn >> 2);
        const MegaLayer & L = P.layers[lm ? 0 : l];
        const int lx = lm ? p.n_layer - 1 : l;                     // layer whose exchange buffer (parity) and tag this tile uses
        unsigned long long * X = P.xch + (size_t) (lx & 1) * M4_LW;
        const uint32_t tag = tag0 | (uint32_t) (lx + 1);
        float * kc = p.kcache + (size_t) (lm ? 0 : l) * p.n_positions * M4_D;
        float * vc = p.vcache + (size_t) (lm ? 0 : l) * p.n_positions * M4_D;
        const int phs = kind == 0 ? 0 : kind + 1;                  // profiling slot (1 = attention)
        if (!lm || tn == n_lt) M4PROF(lm ? 0 : phs, 0);
        if (kind == 0) prefetch_layer(l + 1);
        // the ring slot of tile tn-1 is free after this tile's barrier: its next tenant is tile tn-1+nslot
        TileSrc nxt{nullptr, nullptr, 0u, 0u};
        if (tid == ISSUER) nxt = describe_tile(tn - 1 + P.nslot);
        // ---- rows of this tile
        int rbase, rt;
        if (kind == 0)      { rbase = qkv0; rt = qkv1 - qkv0; }
        else if (kind == 2) { rbase = cta * 32; rt = has_fc1 ? 32 : 0; }
        else if (kind == 4) { rbase = v0 + (tn - n_lt) * M4_LMRT; rt = min(M4_LMRT, v1 - rbase); }
        else                { rbase = o0; rt = o1 - o0; }
        const bool Kff = kind == 3;
        M4MM D;
        D.G = Kff ? p.Gf : p.Gd; D.gsh = Kff ? 5 : 3; D.stride = Kff ? p.stride_f : p.stride_d;
        D.off_qh = Kff ? p.offqh_f : p.offqh_d; D.off_d = Kff ? p.offd_f : p.offd_d; D.off_m = Kff ? p.offm_f : p.offm_d;
        D.off_n = Kff ? p.offn_f : p.offn_d; D.off_dd = Kff ? p.offdd_f : p.offdd_d; D.off_s = Kff ? p.offs_f : p.offs_d;
        uint8_t * rec = smem + ((kind & 1) ? P.sm_act1 : P.sm_act0);
        // ---- row owners fetch their bias before anything can stall.  K = 1024 tiles: lane 28 of warp w
        //      finishes rows w and w + 16; fc2: thread 8*row finishes row `row`.
        int row0 = -1, row1 = -1;
        if (Kff) { if (tid < rt * 8 && (tid & 7) == 0) row0 = tid >> 3; }
        else if (lane == 28) { if (warp < rt) row0 = warp; if (warp + M4_NW < rt) row1 = warp + M4_NW; }
        float bias0 = 0.f, bias1 = 0.f;
        if (kind != 4) {
            const float * bp = kind == 1 ? L.o_b : (kind == 2 ? L.fc1_b : L.fc2_b);
            if (row0 >= 0) { const int r = rbase + row0, mat = r >> 10; bias0 = kind == 0 ? (mat == 0 ? L.q_b : mat == 1 ? L.k_b : L.v_b)[r & 1023] : bp[r]; }
            if (row1 >= 0) { const int r = rbase + row1, mat = r >> 10; bias1 = kind == 0 ? (mat == 0 ? L.q_b : mat == 1 ? L.k_b : L.v_b)[r & 1023] : bp[r]; }
        }
        // ---- inputs of the tile -> activation record in shared memory
        if (kind == 0 || kind == 2 || tn == n_lt) {
            const float * lnw = kind == 0 ? L.ln0_w : (kind == 2 ? L.ln1_w : p.lnf_w);
            const float * lnb = kind == 0 ? L.ln0_b : (kind == 2 ? L.ln1_b : p.lnf_b);
            const float2 lw = *(const float2 *) (lnw + 2 * tid), lb = *(const float2 *) (lnb + 2 * tid);
            if (tn > 0) {
                const unsigned long long * src = kind == 2 ? X + M4_E3 : (kind == 0 ? P.xch + (size_t) ((l - 1) & 1) * M4_LW + M4_E5 : X + M4_E5);
                uint32_t a, b;
                m4_poll2(src + (size_t) rep * M4_D + 2 * tid, kind == 0 ? tag - 1 : tag, a, b, P.poll_sleep);
                xa = __uint_as_float(a); xb = __uint_as_float(b);
            }
            M4PROF(phs, 3);
            *(float2 *) ((kind == 2 ? s_x1 : s_x) + 2 * tid) = make_float2(xa, xb);
            if (rt > 0)
                m4_ln_quant<FMT, PROF>(xa, xb, lw, lb, p.eps, sredA, sredB, rec, D.off_dd, D.off_s,
                                       PROF && P.trace ? P.trace + (size_t) cta * P.prof_n + (l * 5 + (lm ? 0 : phs)) * M4_PK : nullptr);
            if (!lm) M4PROF(phs, 5);
        } else if (kind != 4) {
            const int nb10 = (kind == 1 ? M4_NB_D : M4_NB_F) * 10;
            const unsigned long long * src = X + (kind == 1 ? M4_E2 : M4_E4) + (size_t) rep * nb10;
            if (2 * tid < nb10) {                                  // <= 2 units of two words per thread, both loads in flight
                const bool two = 2 * (tid + M4_NT) < nb10;
                const unsigned long long * pa = src + 2 * tid, * pb = src + 2 * (two ? tid + M4_NT : tid);
                unsigned long long a0, a1, b0, b1;
                for (;;) {
                    asm volatile("ld.volatile.global.v2.u64 {%0,%1}, [%2];" : "=l"(a0), "=l"(a1) : "l"(pa) : "memory");
                    asm volatile("ld.volatile.global.v2.u64 {%0,%1}, [%2];" : "=l"(b0), "=l"(b1) : "l"(pb) : "memory");
                    if ((uint32_t) (a0 >> 32) == tag && (uint32_t) (a1 >> 32) == tag && (uint32_t) (b0 >> 32) == tag && (uint32_t) (b1 >> 32) == tag) break;
                }
                m4_scatter_word(rec, D.off_dd, D.off_s, 2 * tid, (uint32_t) a0);
                m4_scatter_word(rec, D.off_dd, D.off_s, 2 * tid + 1, (uint32_t) a1);
                if (two) {
                    m4_scatter_word(rec, D.off_dd, D.off_s, 2 * (tid + M4_NT), (uint32_t) b0);
                    m4_scatter_word(rec, D.off_dd, D.off_s, 2 * (tid + M4_NT) + 1, (uint32_t) b1);
                }
            }
            M4PROF(phs, 3);
        }
        // ---- matmul over the tile
        float dot0 = 0.f, dot1 = 0.f;
        const uint8_t * wt = s_w;
        if (rt > 0) {
            const int slot = tn % P.nslot;
            m4_mbar_wait(&mbar[slot], (wphase >> slot) & 1u);
            wphase ^= 1u << slot;
            wt = s_w + (size_t) slot * P.slot_bytes;
        }
        if (!lm && (kind & 1)) M4PROF(phs, 5);
        __syncthreads();                                           // record complete; previous tile's shared scratch free
        if (tid == ISSUER) fire_tile(tn - 1 + P.nslot, nxt);
        M4PROF(lm ? 0 : phs, 1);
        if (rt > 0) {
            if (!Kff) {
                if (warp < rt) {
                    M4Act A;
                    m4_load_act<FMT>(A, rec, D, p.code_off, lane >> 2, lane & 3);
                    if (!lm) M4PROF(phs, 9);
                    const float2 dd = m4_two_rows<FMT>(wt, warp, warp + M4_NW < rt ? warp + M4_NW : warp, D, A);
                    dot0 = dd.x; dot1 = dd.y;
                }
            } else {
                m4_phase_a<FMT>(D, wt, rt, rec, p.code_off, s_p, s_s, s_m);
                __syncthreads();
                M4PROF(phs, 4);
                dot0 = m4_phase_b<FMT>(D, rt, rec, s_p, s_s, s_m);
            }
        }
        if (!lm) M4PROF(phs, 10);
        // ---- epilogue of the row owners
#pragma unroll 1
        for (int e = 0; e < 2; e++) {
            const int row = e == 0 ? row0 : row1;
            if (row < 0) break;
            const float dotv = e == 0 ? dot0 : dot1, bias = e == 0 ? bias0 : bias1;
            const int r_own = rbase + row;
            if (kind == 0) {
                const int mat = r_own >> 10, rr = r_own & 1023;
                float t = __fadd_rn(bias, dotv);
                if (mat == 0) t = __fmul_rn(t, p.qscale);
                else (mat == 1 ? kc : vc)[(size_t) pos * M4_D + rr] = t;
                s_blk[row] = t;
            } else if (kind == 1) {
                s_blk[row] = __fadd_rn(__fadd_rn(dotv, bias), s_x[r_own]);
            } else if (kind == 2) {
                s_blk[row] = bg_h2f(p.gelu[bg_f2h(__fadd_rn(bias, dotv))]);
            } else if (kind == 3) {
                s_blk[row] = __fadd_rn(__fadd_rn(bias, dotv), s_x1[r_own]);
            } else {
                p.logits[r_own] = dotv;
                if (dotv > best || (dotv == best && r_own < bi)) { best = dotv; bi = r_own; }
            }
        }
        // ---- publish: the CTA's results are gathered in shared memory and ONE warp writes them as consecutive words, so a
        //      polled line receives a few whole-sector writes instead of one partial 8-byte write per row (measured: words
        //      land ~1 us sooner on lines that 19 CTAs are polling)
        if (kind != 4) __syncthreads();
        if (kind == 0 && tid < 32 && lane < rt) m4_put(X + M4_E1 + rbase + lane, __float_as_uint(s_blk[lane]), tag);
        if ((kind == 1 || kind == 3) && tid < 32) {
            const int row = lane & 7;
            if (row < rt) {
                const uint32_t v = __float_as_uint(s_blk[row]);
                unsigned long long * dst = X + (kind == 1 ? M4_E3 : M4_E5) + rbase + row;
#pragma unroll
                for (int r0 = 0; r0 < M4_R; r0 += 4) m4_put(dst + (size_t) (r0 + (lane >> 3)) * M4_D, v, tag);
            }
        }
        if (kind == 2 && rt > 0 && tid < 32) m4_quant_publish<FMT>(s_blk, X + M4_E4 + (size_t) cta * 10, M4_NB_F * 10, tag);
        if (!lm) M4PROF(phs, 2);
        // ================= attention (32 CTAs: head, 32-column half), after P1 =================
        if (kind == 0 && is_att) {
            const int c0 = att_part * 32;
            const float * Kb = kc + att_h * M4_DK;
            const float * Vb = vc + att_h * M4_DK + c0;
            // K rows of the first pass (t = warp + 16u, u < 16): everything but the new row is already cached
            float kr[16][2];
#pragma unroll
            for (int u = 0; u < 16; u++) {
                const int t = warp + u * M4_NW;
                kr[u][0] = (t < T - 1) ? __ldcg(Kb + (size_t) t * M4_D + lane) : 0.0f;
                kr[u][1] = (t < T - 1) ? __ldcg(Kb + (size_t) t * M4_D + 32 + lane) : 0.0f;
            }
            if (tid < 80) {          // q (64 words), k (64), v half (32) of this head
                const int w = 2 * tid;
                const unsigned long long * src = X + M4_E1 + (w < 64 ? att_h * M4_DK + w : (w < 128 ? M4_D + att_h * M4_DK + (w - 64) : 2 * M4_D + att_h * M4_DK + c0 + (w - 128)));
                uint32_t a, b; m4_poll2(src, tag, a, b, P.poll_sleep);
                float * dstp = w < 64 ? s_q + w : (w < 128 ? s_kn + (w - 64) : s_vn + (w - 128));
                dstp[0] = __uint_as_float(a); dstp[1] = __uint_as_float(b);
            }
            __syncthreads();
            M4PROF(1, 3);
            const float q0 = s_q[lane], q1 = s_q[32 + lane];
            const float kn0 = s_kn[lane], kn1 = s_kn[32 + lane];
#pragma unroll 1
            for (int tb = warp; tb < T; tb += M4_NW * 16) {
                if (tb != warp) {
#pragma unroll
                    for (int u = 0; u < 16; u++) {
                        const int t = tb + u * M4_NW;
                        kr[u][0] = (t < T - 1) ? __ldcg(Kb + (size_t) t * M4_D + lane) : 0.0f;
                        kr[u][1] = (t < T - 1) ? __ldcg(Kb + (size_t) t * M4_D + 32 + lane) : 0.0f;
                    }
                }
                float s[16];
#pragma unroll
                for (int u = 0; u < 16; u++) {
                    const bool isnew = (tb + u * M4_NW) == T - 1;
                    float a = 0.0f;
                    a = fmaf(isnew ? kn0 : kr[u][0], q0, a); a = fmaf(isnew ? kn1 : kr[u][1], q1, a);
                    s[u] = a;
                }
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
                const int u = ((lane >> 1) & 14) | (lane & 1);      // bits 4,3,2 -> u bits 3,2,1; bit 0 -> u bit 0
                const int t = tb + u * M4_NW;
                if (t < T && !(lane & 2)) sc[t] = dot;
            }
            // V rows of the first 512 positions: unit (r = t % 32, column pair cp), 16 loads in flight
            // while the softmax runs
            const int np = T & ~31;
            const int vr = tid >> 4, vcp = tid & 15;
            const float * vp = Vb + (size_t) vr * M4_D + 2 * vcp;
            float2 vv[16];
#pragma unroll
            for (int k = 0; k < 16; k++) {
                const int t = 32 * k + vr;
                vv[k] = (t < np && t != T - 1) ? __ldcg((const float2 *) (vp + (size_t) (32 * k) * M4_D)) : make_float2(0.f, 0.f);
            }
#pragma unroll 1
            for (int i = tid; i < (T - np) * 32; i += M4_NT) {
                const int t = np + (i >> 5);
                tailv[i] = (t == T - 1) ? s_vn[i & 31] : __ldcg(Vb + (size_t) t * M4_D + (i & 31));
            }
            __syncthreads();
            // softmax over sc[0..T): max, fp16-table exp, sum in double (exact for fp16 values), scale
            {
                const float x0 = tid < T ? sc[tid] : -INFINITY, x1 = tid + M4_NT < T ? sc[tid + M4_NT] : -INFINITY;
                float mx = fmaxf(x0, x1);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(FULLMASK, mx, o));
                if (lane == 0) sredF[warp] = mx;
                __syncthreads();
                mx = sredF[lane & 15];
#pragma unroll
                for (int o = 8; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(FULLMASK, mx, o));
                float e0 = 0.f, e1 = 0.f;
                if (tid < T) e0 = bg_h2f(p.exp_tab[bg_f2h(__fsub_rn(x0, mx))]);
                if (tid + M4_NT < T) e1 = bg_h2f(p.exp_tab[bg_f2h(__fsub_rn(x1, mx))]);
                const double sm = m4_warp_sum_f64((double) e0 + (double) e1);
                if (lane == 0) sredA[warp] = sm;
                __syncthreads();
                const float inv = (float) (1.0 / m4_tree16(sredA));
                if (tid < T) sc[tid] = __fmul_rn(e0, inv);
                if (tid + M4_NT < T) sc[tid + M4_NT] = __fmul_rn(e1, inv);
            }
            __syncthreads();
            M4PROF(1, 4);
            // V: running sum r of column pair cp over t = r, r+32, ... < np
            {
                const float2 vnew = *(const float2 *) (s_vn + 2 * vcp);
                float2 acc = make_float2(0.f, 0.f);
#pragma unroll 1
                for (int s0 = 0; s0 < np; s0 += 512) {
                    if (s0 > 0) {
#pragma unroll
                        for (int k = 0; k < 16; k++) {
                            const int t = s0 + 32 * k + vr;
                            vv[k] = (t < np && t != T - 1) ? __ldcg((const float2 *) (vp + (size_t) (s0 + 32 * k) * M4_D)) : make_float2(0.f, 0.f);
                        }
                    }
#pragma unroll
                    for (int k = 0; k < 16; k++) {
                        const int t = s0 + 32 * k + vr;
                        if (t < np) {
                            const float2 v = (t == T - 1) ? vnew : vv[k];
                            const float pw = sc[t];
                            acc.x = fmaf(v.x, pw, acc.x); acc.y = fmaf(v.y, pw, acc.y);
                        }
                    }
                }
                *(float2 *) (red + vr * 32 + 2 * vcp) = acc;
            }
            __syncthreads();
            if (tid < 32) {
                float x0[8];
#pragma unroll
                for (int l8 = 0; l8 < 8; l8++) {
                    const float a02 = __fadd_rn(red[(0 * 8 + l8) * 32 + tid], red[(2 * 8 + l8) * 32 + tid]);
                    const float a13 = __fadd_rn(red[(1 * 8 + l8) * 32 + tid], red[(3 * 8 + l8) * 32 + tid]);
                    x0[l8] = __fadd_rn(a02, a13);
                }
                const float t0 = __fadd_rn(x0[0], x0[4]), t1 = __fadd_rn(x0[1], x0[5]);
                const float t2 = __fadd_rn(x0[2], x0[6]), t3 = __fadd_rn(x0[3], x0[7]);
                float sumf = __fadd_rn(__fadd_rn(t0, t1), __fadd_rn(t2, t3));
                const int nv = np + ((T - np) & ~3);
                int t = np;
#pragma unroll 1
                for (; t < nv; t++) sumf = __fadd_rn(sumf, __fmul_rn(tailv[(t - np) * 32 + tid], sc[t]));
#pragma unroll 1
                for (; t < T;  t++) sumf = fmaf(tailv[(t - np) * 32 + tid], sc[t], sumf);
                s_blk[tid] = sumf;
                __syncwarp();
                m4_quant_publish<FMT>(s_blk, X + M4_E2 + (size_t) (att_h * 2 + att_part) * 10, M4_NB_D * 10, tag);
            }
            M4PROF(1, 2);
        }
    }
    // per-CTA argmax candidate (first index wins ties) for the next launch's prologue
#pragma unroll 1
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(FULLMASK, best, o); const int oi = __shfl_xor_sync(FULLMASK, bi, o);
        if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    if (lane == 0) { s_cv[warp] = best; s_ci[warp] = bi; }
    __syncthreads();
    if (tid == 0) {
#pragma unroll 1
        for (int i = 1; i < M4_NW; i++) if (s_cv[i] > best || (s_cv[i] == best && s_ci[i] < bi)) { best = s_cv[i]; bi = s_ci[i]; }
        p.cand_val[cta] = best; p.cand_idx[cta] = bi;
    }
    { const int l = p.n_layer; M4PROF(0, 2); }
    if (PROF && P.trace && tid == 0) m4_calibrate(P.trace + (size_t) nC * P.prof_n + 4 * cta + 2);
}
