// bgpt_mega5.cuh -- persistent decode kernel, generation 5 (quantised weights, BioGPT-base shapes: d_model 1024, 16 heads of 64,
// d_ff 4096, n_positions <= 1024).  One launch per token; 128 CTAs of 512 threads in 32 thread-block CLUSTERS of 4.
//
// Generation 4 (bgpt_mega4.cuh) removed the grid barriers; its trace (profiles/r1_mega4_trace.log) shows what was left of the
// 18 us per layer: five all-to-all exchanges through L2 (~0.9 us each), dot products that were instruction-issue bound (a relay
// of the 8 running sums along the lanes executed 8 predicated steps: 0.9 us for 2 rows per warp), and an attention stage whose
// 32 CTAs each scored every cached position.  Generation 5 changes the decomposition (measurements behind the choices:
// tools/cluster_probe.cu, profiles/r2_cluster_probe.log):
//
//  1. One attention head = one cluster of 4 CTAs (clusters 0..15; 15 clusters of 8 is all a B200 co-schedules at one CTA per SM,
//     33 clusters of 4 fit).  The CTAs of cluster h compute the q, k, v rows of head h (16 rows of each per CTA) and its
//     attention; q and the new k travel to the 4 CTAs of the cluster through DISTRIBUTED SHARED MEMORY (st.shared::cluster +
//     barrier.cluster, 0.24 us) instead of L2 tagged words (~0.9 us); the T scores are computed once (T/4 per CTA; generation
//     4: T per CTA) and all-gathered the same way; every CTA reduces V for its own 16 columns, whose new row it computed itself.
//     Four exchanges through L2 per layer remain (attention output, x1, the GELU blocks, x); they are inherent: every CTA's next
//     dot product needs the whole vector.  128 CTAs instead of 148: an all-to-all exchange is faster with fewer pollers (1.17
//     against 1.68 us per round in the probe) and every stage divides evenly (8 rows, one fc1 block per CTA).
//  2. Dot products in chain order with no relay: lane (row, l) of a warp owns running sum l of one weight row (4 rows per warp)
//     and walks the 32 (128) blocks in order -- dp4a, int -> float, fma -- straight from the shared-memory weight tile; the only
//     cross-lane step is hsum_float_8 (3 shuffles).  ~230 instructions per lane for K = 1024 against ~450 for the relay.
//  3. Every wait carries a watchdog: a lost signal ends the launch with an error code instead of hanging the GPU.
//
//  stage (per layer)           CTAs, rows                         consumes                       publishes
//  P1 LN0 + q,k,v              0..63, 16 + 16 + 16 rows of a head  E5[l-1] x (L2)                 q, k -> cluster (DSMEM); k, v -> KV cache
//  P2 attention                0..63, T/4 scores + 16 V columns    q, k, scores (DSMEM)           E2: 16 f32 attention outputs (L2, R replicas)
//  P3 out_proj + bias + res    all, 8 rows                         E2 (every CTA quantises it)    E3: x1 (L2)
//  P4 LN1 + fc1 + bias + GELU  all, 32 rows = one block            E3                             E4: one quantised block (L2)
//  P5 fc2 + bias + residual    all, 8 rows (K = 4096)              E4                             E5: x (L2)
//  final LN + lm_head          all, 331-332 rows in tiles          E5[L-1]                        logits, per-CTA argmax candidate
//
// Arithmetic: identical to generations 3 and 4 and to the reference (bgpt_cuda.cu, "lane order"); tests/test_gpu_eval.py compares
// the three generations with each other and with the oracle bit for bit.
#pragma once
#include "bgpt_mega4.cuh"
#include "bgpt_topk.cuh"

#define M5_NT 512
#define M5_NW (M5_NT / 32)
#define M5_CL 4               // CTAs per cluster; clusters 0..15 own one attention head each
#define M5_NC 128             // CTAs per launch (32 clusters)
#define M5_HC (M5_NH * M5_CL) // CTAs 0..63 run P1 / P2
#define M5_HR (M5_DK / M5_CL) // q (k, v) rows = attention columns per head CTA: 16
#define M5_R 8                // replicas of an L2 exchange buffer
#define M5_D 1024
#define M5_FF 4096
#define M5_DK 64
#define M5_NH 16
#define M5_NB_F (M5_FF / 32)
// L2 exchange words per layer parity
#define M5_E2 0
#define M5_E3 (M5_E2 + M5_R * M5_D)
#define M5_E4 (M5_E3 + M5_R * M5_D)
#define M5_E5 (M5_E4 + M5_R * M5_NB_F * 10)
#define M5_LW (M5_E5 + M5_R * M5_D)
// F16 weights exchange the 4096 GELU outputs as fp16 values, two per word: 2048 words per replica, more than the 1280 of the
// quantised form.  They live BEHIND the two parity buffers, so the quantised kernels' addresses are unchanged.
#define M5_E4W (M5_FF / 2)
#define M5_E4F(parity) (2 * M5_LW + (parity) * (M5_R * M5_E4W))
#define M5_XCH_WORDS (2 * M5_LW + 2 * M5_R * M5_E4W)
// the fed token (use_cand == 3): M5_R replicas of one tagged word, 128 bytes apart, behind everything else; tag = launch serial << 6
#define M5_TOKX M5_XCH_WORDS
#define M5_XCH_TOTAL (M5_XCH_WORDS + M5_R * 16)
#define M5_TOK_CANCEL (-2)     // the host withdrew the launch
#define M5_TOK_TIMEOUT (-3)    // no token arrived within feed_limit cycles
#define M5_MAXL 24
#define M5_PK 12              // trace stamps per (layer, stage)
// The relay form of fc2 (no shared-memory scratch: BGPT_M5_FC2=0) is ~10 KB of code in the middle of the layer loop that never runs on
// a B200 (the scratch fits).  Built without it the kernels with the most code gain (Q4_1 461 -> 442, Q5_1 463 -> 452 us per token at
// n_past 511), Q5_0 / Q8_0 do not move, and Q4_0 LOSES 2 % (421 -> 430: these instantiations sit at the register budget and every
// rebuild re-rolls ptxas' register assignment, i.e. operand bank conflicts; A/B on one box, profiles/README.md).  So: kept for Q4_0
// only; where it is not built, bgpt_cuda.cu requires the scratch to fit.  -DM5_FC2_RELAY=1 keeps it everywhere.
#ifndef M5_FC2_RELAY
#define M5_FC2_RELAY 0
#endif
#define M5_HAS_FC2_RELAY(fmt) (M5_FC2_RELAY || (fmt) == BG_Q4_0)

struct M5Params {
    MegaParams b;
    MegaLayer layers[M5_MAXL];
    unsigned long long * xch;      // [2][M5_LW] tagged words (layer parity)
    unsigned int tag;              // launch serial << 6 (never 0); a word's tag = tag | (layer + 1)
    int * err;                     // device: 0, or the code of the first wait that timed out (stage << 16 | layer << 8 | 1)
    long long * trace;             // BGPT_MEGA_PROF: [nC][prof_n] clock64 stamps, then [nC][4] clock calibration
    int prof_n;
    int nslot, slot_bytes, lmrt;   // weight ring: slots, bytes per slot, lm_head rows per tile (32 or 64)
    int sm_w, sm_rec0, sm_rec1, sm_x, sm_x1, sm_sc, sm_red, sm_tail, sm_p, sm_s, sm_total;   // sm_p < 0: no fc2 scratch (relay instead)
    // sampler tail (bgpt_cuda_eval_topk): tk_k > 0 -> the CTA that finishes last selects the tk_k largest logits (bgpt_topk.cuh:
    // topk_tail) and writes the result packet {info[4], vals, ids} -- possibly straight into mapped pinned host memory
    int tk_k; unsigned tk_seq;     // packet of serial s: tk_pk + (s & 1) * tk_stride = { info[4], tk_k vals, tk_k ids }
    uint8_t * tk_pk; int tk_stride;
    unsigned * tk_ticket;          // device counter, 0 between launches
    float * tk_full;               // bgpt_cuda_eval: mapped pinned HOST buffer that receives the whole logit row (every CTA its own rows), or NULL
    // chained launch (use_cand == 3): the kernel is queued BEFORE its input token exists; CTA 0 polls the 8-byte word {token, serial} in
    // mapped pinned host memory until the serial is feed_seq and hands the token to the other CTAs through L2.  No token within
    // feed_limit cycles, or the value M5_TOK_CANCEL: the kernel exits without touching the KV cache or the logits (a time-out leaves a
    // packet with info[1] = -2 under the serial tk_seq).
    // (A RESIDENT form -- the kernel loops over positions instead of exiting, 5 us per token less than a queued launch -- was built and
    // measured: any back edge around the layer loop costs the quantised instantiations, which sit at the 128-register budget, 100-280
    // bytes of spills in the hot loop: 370 -> 415-458 us per token.  profiles/README.md, round 2.)
    const unsigned long long * feed; unsigned feed_seq; long long feed_limit;
};

// ---- cluster / DSMEM -------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t m5_cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t m5_mapa(uint32_t addr, uint32_t rank) { uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank)); return r; }
__device__ __forceinline__ void m5_cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void m5_cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void m5_cluster_wait()   { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
// the same 32-bit value to the same shared-memory offset in all M5_CL CTAs of the cluster (visible after the next cluster barrier)
__device__ __forceinline__ void m5_bcast32(const void * local_dst, uint32_t v) {
    const uint32_t la = m4_s32(local_dst);
#pragma unroll
    for (uint32_t r = 0; r < M5_CL; r++) asm volatile("st.shared::cluster.b32 [%0], %1;" :: "r"(m5_mapa(la, r)), "r"(v) : "memory");
}

// ---- waits with a watchdog ---------------------------------------------------------------------------------------------------
// err[0]: the first time-out's code; err[2..3]: the watchdog limit in cycles (written by the host next to the flag)
__device__ __forceinline__ bool m5_give_up(int * err, int code, long long t0) {
    if (*(volatile int *) err != 0) return true;
    const long long limit = *(const long long *) (err + 2);
    if (clock64() - t0 > limit) { atomicCAS(err, 0, code); return true; }
    return false;
}
__device__ __forceinline__ void m5_poll2(const unsigned long long * p, uint32_t tag, uint32_t & a, uint32_t & b, int * err, int code) {
    unsigned long long w0, w1;
    unsigned spins = 0; long long t0 = 0;
    for (;;) {
        asm volatile("ld.volatile.global.v2.u64 {%0,%1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(p) : "memory");
        if ((uint32_t) (w0 >> 32) == tag && (uint32_t) (w1 >> 32) == tag) break;
        if ((++spins & 1023u) == 0) { if (t0 == 0) t0 = clock64(); else if (m5_give_up(err, code, t0)) break; }
    }
    a = (uint32_t) w0; b = (uint32_t) w1;
}
__device__ __forceinline__ void m5_mbar_wait(uint64_t * bar, uint32_t parity, int * err, int code) {
    uint32_t done; unsigned spins = 0; long long t0 = 0;
    const uint32_t a = m4_s32(bar);
    for (;;) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(a), "r"(parity) : "memory");
        if (done) break;
        if ((++spins & 63u) == 0) { if (t0 == 0) t0 = clock64(); else if (m5_give_up(err, code, t0)) break; }
    }
}

// ---- the "prep" stage runs on 4 warps ---------------------------------------------------------------------------------------------
// With 16 warps per SM every instruction that all warps execute costs ~5 issue cycles (4 warps per scheduler), and a LayerNorm over
// 1024 values is almost all per-thread overhead (shuffle rounds, the cross-warp sum, a division and a square root): 2 elements per
// thread on 512 threads measured 2480 cycles.  128 threads x 8 elements issue a quarter of the instructions; the other 12 warps wait
// at the block barrier that follows and cost nothing.
#define M5_PT 128             // threads of the prep stage
#define M5_PS (M5_NB_F + 4)    // padded row of the fc2 product scratch (floats): conflict-free float4 stores
__device__ __forceinline__ void m5_bar_prep() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

// poll 8 consecutive tagged words (64 bytes) until all carry `tag`
__device__ __forceinline__ void m5_poll8(const unsigned long long * p, uint32_t tag, float (&v)[8], int * err, int code) {
    unsigned long long w[8];
    unsigned spins = 0; long long t0 = 0;
    for (;;) {
#pragma unroll
        for (int i = 0; i < 4; i++)
            asm volatile("ld.volatile.global.v2.u64 {%0,%1}, [%2];" : "=l"(w[2 * i]), "=l"(w[2 * i + 1]) : "l"(p + 2 * i) : "memory");
        bool ok = true;
#pragma unroll
        for (int i = 0; i < 8; i++) ok = ok && (uint32_t) (w[i] >> 32) == tag;
        if (ok) break;
        if ((++spins & 1023u) == 0) { if (t0 == 0) t0 = clock64(); else if (m5_give_up(err, code, t0)) break; }
    }
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = __uint_as_float((uint32_t) w[i]);
}

// activation quantiser (ggml.c:1166-1203, 1403-1450: d = amax / 127, id = 127 / amax, round to nearest even; Q8_0 rounds d through
// fp16): thread t < 128 owns elements 8t..8t+7 = the 4-byte groups l = 2 (t & 3), 2 (t & 3) + 1 of block t >> 2.
// Writes the int8 words [g][l][i], the folded code offsets, d and s of the record.  Ends WITHOUT a barrier.
template <int FMT>
__device__ __forceinline__ void m5_quant8(const float (&y)[8], uint8_t * rec, int off_n, int off_d, int off_s, int code_off) {
    constexpr bool Q81 = (FMT == BG_Q4_1 || FMT == BG_Q5_1);
    constexpr bool HASOFF = (FMT == BG_Q4_0 || FMT == BG_Q5_0);
    const int tid = threadIdx.x;
    float amax = fmaxf(fmaxf(fmaxf(fabsf(y[0]), fabsf(y[1])), fmaxf(fabsf(y[2]), fabsf(y[3]))), fmaxf(fmaxf(fabsf(y[4]), fabsf(y[5])), fmaxf(fabsf(y[6]), fabsf(y[7]))));
    amax = fmaxf(amax, __shfl_xor_sync(FULLMASK, amax, 1));
    amax = fmaxf(amax, __shfl_xor_sync(FULLMASK, amax, 2));
    const float d  = __fdiv_rn(amax, 127.0f);
    const float id = (amax != 0.0f) ? __fdiv_rn(127.0f, amax) : 0.0f;
    int q[8];
#pragma unroll
    for (int i = 0; i < 8; i++) q[i] = __float2int_rn(__fmul_rn(y[i], id));
    const uint32_t w0 = ((uint32_t) q[0] & 0xFFu) | (((uint32_t) q[1] & 0xFFu) << 8) | (((uint32_t) q[2] & 0xFFu) << 16) | (((uint32_t) q[3] & 0xFFu) << 24);
    const uint32_t w1 = ((uint32_t) q[4] & 0xFFu) | (((uint32_t) q[5] & 0xFFu) << 8) | (((uint32_t) q[6] & 0xFFu) << 16) | (((uint32_t) q[7] & 0xFFu) << 24);
    const int s0 = (q[0] + q[1]) + (q[2] + q[3]), s1 = (q[4] + q[5]) + (q[6] + q[7]);
    const int b = tid >> 2, g = b >> 2, i = b & 3, l0 = 2 * (tid & 3);
    ((uint32_t *) rec)[(g * 8 + l0) * 4 + i] = w0;
    ((uint32_t *) rec)[(g * 8 + l0 + 1) * 4 + i] = w1;
    if (HASOFF) {
        ((int *) (rec + off_n))[(g * 8 + l0) * 4 + i] = -code_off * s0;
        ((int *) (rec + off_n))[(g * 8 + l0 + 1) * 4 + i] = -code_off * s1;
    }
    if (Q81) {
        int stot = s0 + s1;
        stot += __shfl_xor_sync(FULLMASK, stot, 1);
        stot += __shfl_xor_sync(FULLMASK, stot, 2);
        if ((tid & 3) == 0) { ((float *) (rec + off_d))[b] = d; ((float *) (rec + off_s))[b] = __fmul_rn(d, (float) stot); }
    } else if ((tid & 3) == 0) { ((float *) (rec + off_d))[b] = bg_h2f(bg_f2h(d)); ((float *) (rec + off_s))[b] = 0.0f; }
}

// LayerNorm + affine on registers (ggml.c:11403-11420 then mul, add), 8 elements per thread on 128 threads; the double sums are
// combined in a parallel order.  Two named barriers (prep threads only) inside.
template <bool PROF>
__device__ __forceinline__ void m5_layer_norm8(const float (&v)[8], const float * lnw, const float * lnb, float eps, double * sredA, double * sredB, float (&y)[8],
                                               long long * stamp /* PROF: this CTA's stamps of the current (layer, stage), or nullptr */) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool st = PROF && stamp && tid == 0;
    const float4 lw0 = *(const float4 *) (lnw + 8 * tid), lw1 = *(const float4 *) (lnw + 8 * tid + 4);
    const float4 lb0 = *(const float4 *) (lnb + 8 * tid), lb1 = *(const float4 *) (lnb + 8 * tid + 4);
    double s = (((double) v[0] + (double) v[1]) + ((double) v[2] + (double) v[3])) + (((double) v[4] + (double) v[5]) + ((double) v[6] + (double) v[7]));
    s = m4_warp_sum_f64(s);
    if (lane == 0) sredA[warp] = s;
    if (st) stamp[7] = clock64();
    m5_bar_prep();
    const float mean = (float) (((sredA[0] + sredA[1]) + (sredA[2] + sredA[3])) * (1.0 / M5_D));
    if (st) stamp[8] = clock64();
    float e[8];
#pragma unroll
    for (int i = 0; i < 8; i++) e[i] = __fsub_rn(v[i], mean);
    double s2 = (((double) __fmul_rn(e[0], e[0]) + (double) __fmul_rn(e[1], e[1])) + ((double) __fmul_rn(e[2], e[2]) + (double) __fmul_rn(e[3], e[3]))) +
                (((double) __fmul_rn(e[4], e[4]) + (double) __fmul_rn(e[5], e[5])) + ((double) __fmul_rn(e[6], e[6]) + (double) __fmul_rn(e[7], e[7])));
    s2 = m4_warp_sum_f64(s2);
    if (lane == 0) sredB[warp] = s2;
    if (st) stamp[9] = clock64();
    m5_bar_prep();
    const float variance = (float) (((sredB[0] + sredB[1]) + (sredB[2] + sredB[3])) * (1.0 / M5_D));
    if (st) stamp[11] = clock64();
    const float scale = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(variance, eps)));
    const float lw[8] = { lw0.x, lw0.y, lw0.z, lw0.w, lw1.x, lw1.y, lw1.z, lw1.w }, lb[8] = { lb0.x, lb0.y, lb0.z, lb0.w, lb1.x, lb1.y, lb1.z, lb1.w };
#pragma unroll
    for (int i = 0; i < 8; i++) y[i] = __fadd_rn(__fmul_rn(lw[i], __fmul_rn(e[i], scale)), lb[i]);
}

// ---- one weight row x the activation record, running sum l of the row in this lane -----------------------------------------------
// The reference (AVX2: ggml.c:2518-2541, 2824-2857, 3071-3093, 3386-3411, 3597-3618) keeps 8 running sums per row: sum l takes the
// elements 4l..4l+3 of every 32-block, acc_l = fma(d_w * d_a, (float) isum_l, acc_l) in block order; Q4_1 / Q5_1 add the chain
// summs = fma(m_w, s_a, summs); the result is hsum_float_8 (+ summs).  Lanes 8q..8q+7 of a warp hold the 8 sums of one row; every
// lane of the warp must call (shuffles); the finished dot is in all 8 lanes of the row.
template <int FMT, int G>
__device__ __forceinline__ float m5_row_dot(const uint8_t * wrow, const uint8_t * rec, const M4MM & D) {
    constexpr bool IS8    = (FMT == BG_Q8_0);
    constexpr bool HASQH  = (FMT == BG_Q5_0 || FMT == BG_Q5_1);
    constexpr bool HASM   = (FMT == BG_Q4_1 || FMT == BG_Q5_1);
    constexpr bool HASOFF = (FMT == BG_Q4_0 || FMT == BG_Q5_0);
    const int l = threadIdx.x & 7, j = l & 3, hi = l >> 2, sh = hi * 4;
    float acc = 0.0f, summ = 0.0f;
#pragma unroll 4
    for (int g = 0; g < G; g++) {
        const uint4 wq = IS8 ? *(const uint4 *) (wrow + ((g * 2 + hi) * 4 + j) * 16) : *(const uint4 *) (wrow + (g * 4 + j) * 16);
        uint32_t qh = 0;
        if (HASQH) qh = *(const uint32_t *) (wrow + D.off_qh + (g * 4 + j) * 4);
        const uint4 aw = *(const uint4 *) (rec + (g * 8 + l) * 16);
        int4 an = make_int4(0, 0, 0, 0);
        if (HASOFF) an = *(const int4 *) (rec + D.off_n + (g * 8 + l) * 16);
        const uint2 dh = *(const uint2 *) (wrow + D.off_d + g * 8);
        const float4 da = *(const float4 *) (rec + D.off_dd + g * 16);
        const uint32_t ww[4] = { wq.x, wq.y, wq.z, wq.w }, aa[4] = { aw.x, aw.y, aw.z, aw.w };
        const int nn[4] = { an.x, an.y, an.z, an.w };
        const float dd[4] = { da.x, da.y, da.z, da.w };
        const uint16_t dw[4] = { (uint16_t) (dh.x & 0xFFFF), (uint16_t) (dh.x >> 16), (uint16_t) (dh.y & 0xFFFF), (uint16_t) (dh.y >> 16) };
        uint16_t mw[4] = { 0, 0, 0, 0 }; float ss[4] = { 0.f, 0.f, 0.f, 0.f };
        if (HASM) {
            const uint2 mh = *(const uint2 *) (wrow + D.off_m + g * 8);
            const float4 sa = *(const float4 *) (rec + D.off_s + g * 16);
            mw[0] = (uint16_t) (mh.x & 0xFFFF); mw[1] = (uint16_t) (mh.x >> 16); mw[2] = (uint16_t) (mh.y & 0xFFFF); mw[3] = (uint16_t) (mh.y >> 16);
            ss[0] = sa.x; ss[1] = sa.y; ss[2] = sa.z; ss[3] = sa.w;
        }
#pragma unroll
        for (int i = 0; i < 4; i++) {
            uint32_t code;
            if (IS8) code = ww[i];
            else {
                code = (ww[i] >> sh) & 0x0F0F0F0Fu;
                if (HASQH) code |= bg_spread4((qh >> (8 * i + sh)) & 0xFu);
            }
            const float p = (float) __dp4a((int) code, (int) aa[i], nn[i]);
            const float s = __fmul_rn(bg_h2f(dw[i]), dd[i]);
            acc = fmaf(s, p, acc);
            if (HASM) summ = fmaf(bg_h2f(mw[i]), ss[i], summ);
        }
    }
    float r = __fadd_rn(acc, __shfl_xor_sync(FULLMASK, acc, 4));
    r = __fadd_rn(r, __shfl_xor_sync(FULLMASK, r, 2));
    r = __fadd_rn(r, __shfl_xor_sync(FULLMASK, r, 1));
    if (HASM) r = __fadd_rn(r, summ);
    return r;
}

// ---- one weight row per warp, K split over the four quarter-warps ("relay") ---------------------------------------------------
// For the stages with few rows per CTA (out_proj, fc1, fc2: 8 / 32 / 8 rows) the 4-rows-per-warp mapping above leaves most of the
// SM idle and one lane walks all K/32 blocks.  Here lane (kq = lane >> 3, l = lane & 7) does the integer work of running sum l for
// the blocks of quarter kq only (K/128 blocks), all in registers and all four quarters at once; then the sums are continued in
// block order by handing acc from quarter to quarter (3 shuffles): quarter 0 walks its blocks, quarter 1 continues, ...  The scale
// products d_w * d_a are computed once per block (one lane each) and distributed by shuffles instead of 8 times.  Same operations
// in the same order as m5_row_dot.  The finished dot (hsum_float_8 + summs) is in lanes 24..31.
// SUMM: continue the Q4_1 / Q5_1 `summs` chain in the same lanes (K = 1024); for K = 4096 that would need 64 more registers, so a
// helper warp walks it (m5_summs_chain) and the caller adds the two parts.
template <int FMT, int G, bool SUMM>          // G = K / 128 groups of 4 blocks in the row; a quarter owns G / 4 groups
__device__ __forceinline__ float m5_row_dot_relay(const uint8_t * wrow, const uint8_t * rec, const M4MM & D) {
    constexpr bool IS8    = (FMT == BG_Q8_0);
    constexpr bool HASQH  = (FMT == BG_Q5_0 || FMT == BG_Q5_1);
    constexpr bool HASM   = SUMM && (FMT == BG_Q4_1 || FMT == BG_Q5_1);
    constexpr bool HASOFF = (FMT == BG_Q4_0 || FMT == BG_Q5_0);
    constexpr int GQ = G / 4, NBQ = GQ * 4;                      // groups / blocks per quarter: 2 / 8 (K = 1024), 8 / 32 (K = 4096)
    constexpr int SPL = NBQ / 8;                                 // scale products per lane: 1 or 4
    const int lane = threadIdx.x & 31, kq = lane >> 3, l = lane & 7, j = l & 3, hi = l >> 2, sh = hi * 4;
    // scale products of this quarter's blocks, one lane each: lane l owns blocks kq * NBQ + SPL * l .. + SPL - 1
    float sprod[SPL], mval[SPL], sval[SPL];
#pragma unroll
    for (int e = 0; e < SPL; e++) {
        const int b = kq * NBQ + SPL * l + e;
        sprod[e] = __fmul_rn(bg_h2f(*(const uint16_t *) (wrow + D.off_d + b * 2)), *(const float *) (rec + D.off_dd + b * 4));
        mval[e] = 0.f; sval[e] = 0.f;
        if (HASM) { mval[e] = bg_h2f(*(const uint16_t *) (wrow + D.off_m + b * 2)); sval[e] = *(const float *) (rec + D.off_s + b * 4); }
    }
    // integer dots of this lane's running sum over the quarter's blocks
    float pv[NBQ];
#pragma unroll
    for (int gg = 0; gg < GQ; gg++) {
        const int g = kq * GQ + gg;
        const uint4 wq = IS8 ? *(const uint4 *) (wrow + ((g * 2 + hi) * 4 + j) * 16) : *(const uint4 *) (wrow + (g * 4 + j) * 16);
        uint32_t qh = 0;
        if (HASQH) qh = *(const uint32_t *) (wrow + D.off_qh + (g * 4 + j) * 4);
        const uint4 aw = *(const uint4 *) (rec + (g * 8 + l) * 16);
        int4 an = make_int4(0, 0, 0, 0);
        if (HASOFF) an = *(const int4 *) (rec + D.off_n + (g * 8 + l) * 16);
        const uint32_t ww[4] = { wq.x, wq.y, wq.z, wq.w }, aa[4] = { aw.x, aw.y, aw.z, aw.w };
        const int nn[4] = { an.x, an.y, an.z, an.w };
#pragma unroll
        for (int i = 0; i < 4; i++) {
            uint32_t code;
            if (IS8) code = ww[i];
            else {
                code = (ww[i] >> sh) & 0x0F0F0F0Fu;
                if (HASQH) code |= bg_spread4((qh >> (8 * i + sh)) & 0xFu);
            }
            pv[gg * 4 + i] = (float) __dp4a((int) code, (int) aa[i], nn[i]);
        }
    }
    // every lane collects the scale products of its quarter: block b of the quarter lives in lane (kq, b / SPL), register b % SPL
    float sq[NBQ], mq[HASM ? NBQ : 1], aq[HASM ? NBQ : 1];
#pragma unroll
    for (int b = 0; b < NBQ; b++) {
        const int src = (lane & 24) + b / SPL;
        sq[b] = __shfl_sync(FULLMASK, sprod[b % SPL], src);
        if (HASM) { mq[b] = __shfl_sync(FULLMASK, mval[b % SPL], src); aq[b] = __shfl_sync(FULLMASK, sval[b % SPL], src); }
    }
    // the chains in block order: quarter 0 first, its sums handed to quarter 1, ...
    float acc = 0.0f, summ = 0.0f;
#pragma unroll
    for (int q = 0; q < 4; q++) {
        if (q > 0) {
            const float ia = __shfl_up_sync(FULLMASK, acc, 8);
            float im = 0.f;
            if (HASM) im = __shfl_up_sync(FULLMASK, summ, 8);
            if (kq == q) { acc = ia; summ = im; }
        }
        if (kq == q) {
#pragma unroll
            for (int b = 0; b < NBQ; b++) { acc = fmaf(sq[b], pv[b], acc); if (HASM) summ = fmaf(mq[b], aq[b], summ); }
        }
    }
    float r = __fadd_rn(acc, __shfl_xor_sync(FULLMASK, acc, 4));
    r = __fadd_rn(r, __shfl_xor_sync(FULLMASK, r, 2));
    r = __fadd_rn(r, __shfl_xor_sync(FULLMASK, r, 1));
    if (HASM) r = __fadd_rn(r, summ);
    return r;                                                     // valid in lanes 24..31
}

// the Q4_1 / Q5_1 `summs` chain of one K = 4096 row, summs = fma(m_w[b], s_a[b], summs) over the 128 blocks in order
// (ggml.c:2824-2857, 3386-3411), walked by a helper warp while the row's owner warp does the relay: lane i loads blocks 4i..4i+3;
// the values travel to lane 0 by shuffles and lane 0 runs the 128 dependent fmas.  Returned in every lane.
__device__ __forceinline__ float m5_summs_chain(const uint8_t * wrow, const uint8_t * rec, const M4MM & D) {
    const int lane = threadIdx.x & 31;
    const uint2 mh = *(const uint2 *) (wrow + D.off_m + lane * 8);
    const float4 sa = *(const float4 *) (rec + D.off_s + lane * 16);
    const float m0 = bg_h2f((uint16_t) (mh.x & 0xFFFF)), m1 = bg_h2f((uint16_t) (mh.x >> 16));
    const float m2 = bg_h2f((uint16_t) (mh.y & 0xFFFF)), m3 = bg_h2f((uint16_t) (mh.y >> 16));
    float summ = 0.f;
#pragma unroll
    for (int i = 0; i < 32; i++) {
        summ = fmaf(__shfl_sync(FULLMASK, m0, i), __shfl_sync(FULLMASK, sa.x, i), summ);
        summ = fmaf(__shfl_sync(FULLMASK, m1, i), __shfl_sync(FULLMASK, sa.y, i), summ);
        summ = fmaf(__shfl_sync(FULLMASK, m2, i), __shfl_sync(FULLMASK, sa.z, i), summ);
        summ = fmaf(__shfl_sync(FULLMASK, m3, i), __shfl_sync(FULLMASK, sa.w, i), summ);
    }
    return summ;
}

// scatter one exchange word (block b, word k) into a quantised record, with the folded code offset
__device__ __forceinline__ void m5_scatter_word(uint8_t * rec, int off_n, int off_d, int off_s, int code_off, int w, uint32_t v) {
    const int b = w / 10, k = w - b * 10;
    if (k < 8) {
        const int g = b >> 2, i = b & 3;
        ((uint32_t *) rec)[(g * 8 + k) * 4 + i] = v;
        if (code_off) ((int *) (rec + off_n))[(g * 8 + k) * 4 + i] = -code_off * __dp4a((int) v, 0x01010101, 0);
    } else if (k == 8) ((float *) (rec + off_d))[b] = __uint_as_float(v);
    else               ((float *) (rec + off_s))[b] = __uint_as_float(v);
}

// ---- F16 weights (FMT == BG_F16) ----------------------------------------------------------------------------------------------
// The reference converts the activation row to fp16 and keeps 32 running sums, sum l taking the elements with index % 32 == l in
// order, fma(w, a, sum) in f32 (ggml_vec_dot_f16, ggml.c:2409-2443), then GGML_F32x8_REDUCE.  A weight row lies in HBM / the
// shared-memory tile as [gg][lane][e] fp16 with element = gg * 256 + e * 32 + lane (bgpt_layout.h); the activation record is written
// in the same order, so a lane's 8 weights and 8 activations of a group are one 16-byte load each.
__device__ __forceinline__ int m5_h_pos(int c) { return (((c >> 8) * 32 + (c & 31)) << 3) + ((c >> 5) & 7); }
// prep thread t < 128 owns elements 8t .. 8t + 7
__device__ __forceinline__ void m5_record8_f16(const float (&y)[8], uint8_t * rec) {
    const int t = threadIdx.x;
#pragma unroll
    for (int i = 0; i < 8; i++) ((uint16_t *) rec)[m5_h_pos(8 * t + i)] = bg_f2h(y[i]);
}
// one weight row (NG groups of 256 elements) x the record, by a whole warp; the finished dot is in every lane
template <int NG>
__device__ __forceinline__ float m5_row_dot_f16(const uint8_t * wrow, const uint8_t * rec) {
    const int lane = threadIdx.x & 31;
    float c = 0.0f;
#pragma unroll 4
    for (int g = 0; g < NG; g++) {
        const uint4 w = *(const uint4 *) (wrow + (size_t) (g * 32 + lane) * 16);
        const uint4 a = *(const uint4 *) (rec + (size_t) (g * 32 + lane) * 16);
        const uint32_t ww[4] = { w.x, w.y, w.z, w.w }, aa[4] = { a.x, a.y, a.z, a.w };
#pragma unroll
        for (int i = 0; i < 4; i++) {
            c = fmaf(bg_h2f((uint16_t) (ww[i] & 0xFFFF)), bg_h2f((uint16_t) (aa[i] & 0xFFFF)), c);
            c = fmaf(bg_h2f((uint16_t) (ww[i] >> 16)), bg_h2f((uint16_t) (aa[i] >> 16)), c);
        }
    }
    c = __fadd_rn(c, __shfl_xor_sync(FULLMASK, c, 16));
    c = __fadd_rn(c, __shfl_xor_sync(FULLMASK, c, 8));
    c = __fadd_rn(c, __shfl_xor_sync(FULLMASK, c, 4));
    c = __fadd_rn(c, __shfl_xor_sync(FULLMASK, c, 1));
    c = __fadd_rn(c, __shfl_xor_sync(FULLMASK, c, 2));
    return c;
}

// ---- chained launch: the token arrives after the kernel has started -------------------------------------------------------------
// One thread per CTA.  CTA 0 polls the 8-byte word {token, serial} in mapped pinned HOST memory until the serial is `fseq` (or `limit`
// cycles have passed: M5_TOK_TIMEOUT) and publishes the token as M5_R tagged words in L2; the other CTAs poll their replica.
// Out of line on purpose: the quantised instantiations of k_mega5 sit at the 128-register budget, and code added inline -- even in the
// prologue -- reshuffles the allocation of the layer loop (Q5_0: 420 -> 527 us per token with this block inline).
static __device__ __noinline__ int m5_fetch_token(const unsigned long long * feed, unsigned fseq, long long limit, unsigned long long * xtok,
                                                   uint32_t tag0, int cta, int * err) {
    if (cta == 0) {
        int t = M5_TOK_TIMEOUT;
        const long long t0 = clock64();
        for (;;) {
            unsigned long long w;
            asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(w) : "l"(feed) : "memory");
            if ((uint32_t) (w >> 32) == fseq) { t = (int) (uint32_t) w; break; }
            if (clock64() - t0 > limit) break;
        }
#pragma unroll
        for (int r = 0; r < M5_R; r++) m4_put(xtok + r * 16, (uint32_t) t, tag0);
        return t;
    }
    const unsigned long long * src = xtok + (cta % M5_R) * 16;
    unsigned long long w = 0; unsigned spins = 0; long long t0 = 0;
    for (;;) {
        asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(w) : "l"(src) : "memory");
        if ((uint32_t) (w >> 32) == tag0) break;
        if ((++spins & 1023u) == 0) { if (t0 == 0) t0 = clock64(); else if (m5_give_up(err, 1, t0)) break; }
    }
    return (int) (uint32_t) w;
}

// a withdrawn launch: the ISSUER thread (n_fired > 0) waits for the first-use phase of the ring slots it has fired -- a CTA must not
// exit with bulk copies in flight into its shared memory; CTA 0's thread 0 reports a time-out under the serial the launch owed
static __device__ __noinline__ void m5_withdrawn(uint64_t * mbar, int n_fired, bool is_head, int * err, bool announce, uint8_t * pk, unsigned seq) {
#pragma unroll 1
    for (int n = is_head ? 0 : 1; n < n_fired; n++) m5_mbar_wait(&mbar[n], 0u, err, 2);
    if (announce) {
        int * info = (int *) pk;
        info[0] = 0; info[1] = -2; info[2] = 0;
        __threadfence_system();
        *(volatile unsigned *) (info + 3) = seq;
    }
}
// the sampler tail: whoever takes the last ticket sees every CTA's logits and maximum (fence + atomic on both sides) and selects
// (bgpt_topk.cuh); called by every thread of every CTA
// full != NULL (bgpt_cuda_eval): every CTA first copies ITS lm_head rows to the mapped host buffer -- 128 CTAs x 1.3 KB of posted writes
// in parallel, no D2H copy command behind the kernel -- and releases at system scope, so the serial the last CTA writes orders them all
static __device__ __noinline__ void m5_sampler_tail(unsigned * ticket, const float * logits, int n_vocab, int k, const float * cand_val, uint8_t * scratch,
                                                    uint8_t * pk, const int * err, unsigned seq, float * full) {
    __shared__ int s_last;
    if (full) {
        const unsigned uc = blockIdx.x;
        const int v0 = (int) ((uc * (unsigned) n_vocab) / (unsigned) M5_NC), v1 = (int) (((uc + 1u) * (unsigned) n_vocab) / (unsigned) M5_NC);
        for (int i = v0 + (int) threadIdx.x; i < v1; i += M5_NT) full[i] = __ldcg(logits + i);
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        if (full) asm volatile("fence.acq_rel.sys;" ::: "memory");
        asm volatile("fence.acq_rel.gpu;" ::: "memory");
        s_last = atomicAdd(ticket, 1u) == (unsigned) (M5_NC - 1);
        if (s_last) { asm volatile("fence.acq_rel.gpu;" ::: "memory"); *ticket = 0u; }
    }
    __syncthreads();
    if (s_last) topk_tail<M5_NT>(logits, n_vocab, k, cand_val, M5_NC, scratch, (float *) (pk + 16), (int *) (pk + 16 + (size_t) k * 4), (int *) pk, err, seq);
}

#define M5PROF(ph, k) do { if (PROF && P.trace && threadIdx.x == 0) P.trace[(size_t) blockIdx.x * P.prof_n + (l * 5 + (ph)) * M5_PK + (k)] = clock64(); } while (0)

// TK: the instantiation bgpt_cuda_eval_topk launches -- token feed (use_cand == 3) and the sampler tail.  A template flag, not a
// run-time one: the quantised instantiations sit at the 128-register budget, and the mere presence of that code re-rolls ptxas'
// allocation of the layer loop (one build: Q5_0 434 -> 547 us per token, Q4_1 + 3.6 %, Q4_0 + 1.3 %; the next build, with 10 KB of
// unrelated code removed, had the Q5_0 twin back at 438).  The decode loop keeps the instantiation without it.
template <int FMT, bool PROF, bool TK>
__global__ void __launch_bounds__(M5_NT, 1) k_mega5(const __grid_constant__ M5Params P) {
    const MegaParams & p = P.b;
    extern __shared__ __align__(16) uint8_t smem[];
    __shared__ __align__(8) uint64_t mbar[M4_NSLOT];               // weight ring
    __shared__ double sredA[M5_NW], sredB[M5_NW];
    __shared__ float sredF[M5_NW];
    __shared__ unsigned long long s_isum[4];
    __shared__ __align__(16) float s_blk[32];
    __shared__ __align__(16) float s_q[M5_DK], s_kn[M5_DK];        // q and the new k row of this head: written by the 4 CTAs of the cluster
    __shared__ __align__(16) float s_vn[M5_HR];                    // new v values of this CTA's 16 columns
    __shared__ float s_cv[M5_NW]; __shared__ int s_ci[M5_NW];
    __shared__ int s_tok;
    constexpr bool ISF = (FMT == BG_F16);                          // F16 weights: fp16 records, one row per warp with lane = running sum
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cta = blockIdx.x;
    const bool is_head = cta < M5_HC;                              // clusters 0..15: one attention head each
    const int head = cta >> 2;
    const int rank = (int) m5_cluster_rank();                      // == cta & 3 for a 1-D cluster of 4
    int * const err = P.err;
    uint8_t * s_w = smem + P.sm_w;
    uint8_t * rec0 = smem + P.sm_rec0;
    uint8_t * rec1 = smem + P.sm_rec1;
    float * s_x = (float *) (smem + P.sm_x);
    float * s_x1 = (float *) (smem + P.sm_x1);
    float * sc = (float *) (smem + P.sm_sc);                       // [n_positions] scores / probabilities of this head
    float * red = (float *) (smem + P.sm_red);                     // [32][16]
    float * tailv = (float *) (smem + P.sm_tail);                  // [31][16]
    float * s_p = (float *) (smem + (P.sm_p < 0 ? 0 : P.sm_p));    // fc2: [8 rows][8 sums][M5_PS] block products
    float * s_s = (float *) (smem + (P.sm_s < 0 ? 0 : P.sm_s));    // fc2: [8 rows][128] scale products, then [8][128] minima (Q4_1 / Q5_1)

    const int n_pos = p.n_positions;
    const int o0 = cta * 8;                                        // out_proj / fc2 rows
    const int hrow0 = (head & (M5_NH - 1)) * M5_DK + rank * M5_HR; // first of this CTA's 16 q / k / v rows (= its 16 attention columns)
    const unsigned uC = (unsigned) M5_NC, uc = (unsigned) cta;
    const int v0 = (int) ((uc * (unsigned) p.n_vocab) / uC), v1 = (int) (((uc + 1u) * (unsigned) p.n_vocab) / uC);
    const int n_lm = (v1 - v0 + P.lmrt - 1) / P.lmrt;
    const int n_lt = 4 * p.n_layer;
    const int n_tiles = n_lt + n_lm;
    const int rep = cta % M5_R;

    // ---- weight ring: tile n = layer n>>2, stage n&3 (P1, P3, P4, P5), then the lm_head tiles
    struct TileSrc { const uint8_t * s0, * s1, * s2; uint32_t b0, b1, b2; };
    auto describe_tile = [&](int n) -> TileSrc {
        TileSrc t{nullptr, nullptr, nullptr, 0u, 0u, 0u};
        if (n >= n_tiles) return t;
        if (n < n_lt) {
            const MegaLayer & L = P.layers[n >> 2];
            const int k = n & 3;
            if (k == 0) {
                if (is_head) {
                    const size_t off = (size_t) hrow0 * p.stride_d;
                    t.s0 = L.q_w + off; t.s1 = L.k_w + off; t.s2 = L.v_w + off;
                    t.b0 = t.b1 = t.b2 = (uint32_t) M5_HR * (uint32_t) p.stride_d;
                }
            } else if (k == 1) { t.s0 = L.o_w + (size_t) o0 * p.stride_d; t.b0 = 8u * (uint32_t) p.stride_d; }
            else if (k == 2)   { t.s0 = L.fc1_w + (size_t) cta * 32 * p.stride_d; t.b0 = 32u * (uint32_t) p.stride_d; }
            else               { t.s0 = L.fc2_w + (size_t) o0 * p.stride_f; t.b0 = 8u * (uint32_t) p.stride_f; }
        } else {
            const int r = v0 + (n - n_lt) * P.lmrt;
            t.s0 = p.lm_head + (size_t) r * p.stride_d; t.b0 = (uint32_t) min(P.lmrt, v1 - r) * (uint32_t) p.stride_d;
        }
        return t;
    };
    const uint64_t pol_w = m4_policy_evict_first();
    auto fire_tile = [&](int n, const TileSrc & t) {
        if (t.b0 == 0) return;
        const int slot = n % P.nslot;
        uint8_t * dst = s_w + (size_t) slot * P.slot_bytes;
        m4_mbar_expect(&mbar[slot], t.b0 + t.b1 + t.b2);
        m4_bulk_g2s(dst, t.s0, t.b0, &mbar[slot], pol_w);
        if (t.b1) m4_bulk_g2s(dst + t.b0, t.s1, t.b1, &mbar[slot], pol_w);
        if (t.b2) m4_bulk_g2s(dst + t.b0 + t.b1, t.s2, t.b2, &mbar[slot], pol_w);
    };
    constexpr int ISSUER = M5_NT - 32;
    uint32_t wphase = 0;
    if (tid == 0) {
        for (int i = 0; i < P.nslot; i++) m4_mbar_init(&mbar[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const int pos = p.n_past, T = p.n_past + 1;
    const uint32_t tag0 = P.tag;
    const bool fed = TK && p.use_cand == 3;
    // L2 prefetch of the K/V lines this head reads in layer Ln (2 lines per position and tensor, spread over the cluster) and of
    // Ln's small f32 vectors
    auto prefetch_layer = [&](int Ln) {
        if (Ln >= p.n_layer) return;
        if (is_head) {
            const int pr = rank * 512 + tid;
            const int t = pr >> 1, half = pr & 1;
            if (t < pos) {
                const size_t o = (size_t) Ln * n_pos * M5_D + (size_t) t * M5_D + head * M5_DK + half * 32;
                asm volatile("prefetch.global.L2 [%0];" :: "l"(p.kcache + o));
                asm volatile("prefetch.global.L2 [%0];" :: "l"(p.vcache + o));
            }
        }
        const int who = M5_NC - 1 - cta;                           // 10 vectors of <= 16 KB, one CTA each
        if (who < 10) {
            const float * const * vecs = (const float * const *) &P.layers[Ln].q_b;     // q_b .. fc2_b: 10 consecutive pointers
            const unsigned bytes = (who == 8 ? M5_FF : M5_D) * 4u;                      // fc1_b is the 9th
#pragma unroll 1
            for (unsigned off = tid * 128u; off < bytes; off += M5_NT * 128u)
                asm volatile("prefetch.global.L2 [%0];" :: "l"((const uint8_t *) vecs[who] + off));
        }
    };
    if (PROF && P.trace && tid == 0) m4_calibrate(P.trace + (size_t) M5_NC * P.prof_n + 4 * cta);
    // the first tiles and layer 0's K/V lines are on their way while the token may still be unknown
    if (tid == ISSUER) {
#pragma unroll 1
        for (int n = 0; n < P.nslot - 1; n++) fire_tile(n, describe_tile(n));
    }
    prefetch_layer(0);

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
        } else if (fed) {
            tok = 0;
            if constexpr (TK) if (tid == 0) tok = m5_fetch_token(P.feed, P.feed_seq, P.feed_limit, P.xch + M5_TOKX, tag0, cta, err);
            tok = __shfl_sync(FULLMASK, tok, 0);
        } else tok = p.use_cand == 2 ? p.tok_imm : __ldcg(p.tok);
        if (tid == 0) {
            s_tok = tok;
            if (cta == 0 && p.log_slot >= 0) p.idlog[p.log_slot] = tok;
        }
    }
    // every CTA of the cluster is running (its shared memory exists) before anyone stores into a peer; also publishes s_tok
    m5_cluster_sync();
    if constexpr (TK) if (fed && s_tok < 0) {
        // withdrawn (or no token in time): nothing was written; the weight tiles fired above (tile 0 only by the head CTAs) must land
        // before the CTA may exit
        if (tid == ISSUER || (cta == 0 && tid == 0))
            m5_withdrawn(mbar, tid == ISSUER ? P.nslot - 1 : 0, is_head, err, cta == 0 && tid == 0 && s_tok == M5_TOK_TIMEOUT && P.tk_k > 0,
                         P.tk_pk + (size_t) (P.tk_seq & 1u) * P.tk_stride, P.tk_seq);
        return;
    }
    // ---- embedding: every CTA, 8 elements per prep thread, into s_x (read back by the first tile)
    if (tid < M5_PT) {
        int tok = s_tok; tok = tok < 0 ? 0 : (tok >= p.n_vocab ? p.n_vocab - 1 : tok);
        int prow = pos + 2; prow = prow >= p.n_pos_rows ? p.n_pos_rows - 1 : prow;
        const size_t rb = bg_file_row_bytes(FMT, M5_D);
        const uint8_t * tr = p.embed_tok + rb * (size_t) tok;
        const uint8_t * pr = p.embed_pos + rb * (size_t) prow;
#pragma unroll
        for (int i = 0; i < 8; i++)
            s_x[8 * tid + i] = __fadd_rn(__fmul_rn(bg_dequant_elem(FMT, tr, 8 * tid + i), p.emb_scale), bg_dequant_elem(FMT, pr, 8 * tid + i));
    }

    float best = -INFINITY; int bi = 0x7fffffff;                  // lm_head argmax of this thread's rows
#pragma unroll 1
    for (int tn = 0; tn < n_tiles; tn++) {
        const bool lm = tn >= n_lt;
        const int kind = lm ? 4 : (tn & 3);                        // 0 P1 (+ attention), 1 P3, 2 P4, 3 P5, 4 lm_head
        const int l = lm ? p.n_layer : (tn >> 2);
        const MegaLayer & L = P.layers[lm ? 0 : l];
        const int lx = lm ? p.n_layer - 1 : l;                     // layer whose exchange buffer (parity) and tag this tile uses
        unsigned long long * X = P.xch + (size_t) (lx & 1) * M5_LW;
        const uint32_t tag = tag0 | (uint32_t) (lx + 1);
        const int wcode = ((kind + 1) << 16) | (lx << 8);          // watchdog code of this tile's waits
        float * kc = p.kcache + (size_t) (lm ? 0 : l) * n_pos * M5_D;
        float * vc = p.vcache + (size_t) (lm ? 0 : l) * n_pos * M5_D;
        const int phs = (kind == 0 || lm) ? 0 : kind + 1;         // trace slot (1 = attention; the lm_head tiles use slot 0 of "layer" n_layer)
        if (!lm || tn == n_lt) M5PROF(lm ? 0 : phs, 0);
        if (kind == 0) prefetch_layer(l + 1);
        // the ring slot of tile tn-1 is free after this tile's barrier: its next tenant is tile tn-1+nslot
        TileSrc nxt{nullptr, nullptr, nullptr, 0u, 0u, 0u};
        if (tid == ISSUER) nxt = describe_tile(tn - 1 + P.nslot);
        // ---- rows of this tile: warp w, lanes 8q..8q+7 -> local row 4w + q
        int rbase, rt;
        if (kind == 0)      { rbase = 0; rt = is_head ? 3 * M5_HR : 0; }
        else if (kind == 2) { rbase = cta * 32; rt = 32; }
        else if (kind == 4) { rbase = v0 + (tn - n_lt) * P.lmrt; rt = min(P.lmrt, v1 - rbase); }
        else                { rbase = o0; rt = 8; }
        const bool Kff = kind == 3;
        M4MM D;
        D.G = Kff ? p.Gf : p.Gd; D.gsh = Kff ? 5 : 3; D.stride = Kff ? p.stride_f : p.stride_d;
        D.off_qh = Kff ? p.offqh_f : p.offqh_d; D.off_d = Kff ? p.offd_f : p.offd_d; D.off_m = Kff ? p.offm_f : p.offm_d;
        D.off_n = Kff ? p.offn_f : p.offn_d; D.off_dd = Kff ? p.offdd_f : p.offdd_d; D.off_s = Kff ? p.offs_f : p.offs_d;
        uint8_t * rec = (kind & 1) ? rec1 : rec0;
        // direct mapping (P1, P4, lm_head): warp w, lanes 8q..8q+7 -> local row 4w + q, finished in lane 8q.  Relay mapping (P3, P5): one
        // row per warp, finished in lane 24; the 8 results are gathered and published by warp 0.
        const bool relay = kind == 1 || kind == 3;
        const int myrow = 4 * warp + (lane >> 3);
        const bool owner = !ISF && !relay && (lane & 7) == 0 && myrow < rt;
        // ---- whoever finishes a row fetches its bias before anything can stall
        float bias = 0.f;
        float fbias[3] = { 0.f, 0.f, 0.f };                        // F16: lane 0 of warp w finishes local rows w, w + 16, w + 32
        if (ISF && !relay && lane == 0) {
#pragma unroll
            for (int j = 0; j < 3; j++) {
                const int rr = warp + M5_NW * j;
                if (rr >= rt) break;
                if (kind == 0) { const int mat = rr >> 4; fbias[j] = (mat == 0 ? L.q_b : mat == 1 ? L.k_b : L.v_b)[hrow0 + (rr & 15)]; }
                else if (kind == 2) fbias[j] = L.fc1_b[rbase + rr];
            }
        }
        if (kind == 0) { if (owner) { const int mat = myrow >> 4; bias = (mat == 0 ? L.q_b : mat == 1 ? L.k_b : L.v_b)[hrow0 + (myrow & 15)]; } }
        else if (kind == 2) { if (owner) bias = L.fc1_b[rbase + myrow]; }
        else if (relay) { if (warp == 0 && lane < 8) bias = (kind == 1 ? L.o_b : L.fc2_b)[rbase + lane]; }
        // ---- weights of the tile: ONE thread of a warp that idles during the prep stage waits for the bulk copies (they were issued
        //      three tiles ago); the block barrier below hands the completed phase to everyone
        const int slot = tn % P.nslot;
        const uint8_t * wt = s_w + (size_t) slot * P.slot_bytes;
        if (rt > 0) {
            if (tid == M5_PT) m5_mbar_wait(&mbar[slot], (wphase >> slot) & 1u, err, wcode | 2);
            wphase ^= 1u << slot;
        }
        // ---- inputs of the tile -> activation record in shared memory (4 warps; everyone else goes straight to the barrier)
        if (kind == 0 && !is_head) {
            // clusters 16..31 have no q, k, v rows: they only need x at their 8 out_proj rows (the residual of P3)
            if (tn > 0 && tid < 4) {
                const unsigned long long * src = P.xch + (size_t) ((l - 1) & 1) * M5_LW + M5_E5 + (size_t) rep * M5_D + o0 + 2 * tid;
                uint32_t a, b;
                m5_poll2(src, tag - 1, a, b, err, wcode | 1);
                *(float2 *) (s_x + o0 + 2 * tid) = make_float2(__uint_as_float(a), __uint_as_float(b));
            }
        } else if (kind < 3 || tn == n_lt) {
            if (tid < M5_PT) {
                const bool ln = kind != 1;
                float v[8];
                if (tn > 0) {
                    const unsigned long long * src;
                    uint32_t want = tag;
                    if (kind == 0) { src = P.xch + (size_t) ((l - 1) & 1) * M5_LW + M5_E5; want = tag - 1; }
                    else if (kind == 1) src = X + M5_E2;
                    else if (kind == 2) src = X + M5_E3;
                    else src = X + M5_E5;
                    m5_poll8(src + (size_t) rep * M5_D + 8 * tid, want, v, err, wcode | 1);
                } else {
                    const float4 a = *(const float4 *) (s_x + 8 * tid), b = *(const float4 *) (s_x + 8 * tid + 4);
                    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
                }
                M5PROF(phs, 3);
                if ((kind == 0 && tn > 0) || kind == 2) {
                    float * dstx = kind == 0 ? s_x : s_x1;
                    *(float4 *) (dstx + 8 * tid) = make_float4(v[0], v[1], v[2], v[3]);
                    *(float4 *) (dstx + 8 * tid + 4) = make_float4(v[4], v[5], v[6], v[7]);
                }
                float y[8];
                if (ln) {
                    const float * lnw = kind == 0 ? L.ln0_w : (kind == 2 ? L.ln1_w : p.lnf_w);
                    const float * lnb = kind == 0 ? L.ln0_b : (kind == 2 ? L.ln1_b : p.lnf_b);
                    m5_layer_norm8<PROF>(v, lnw, lnb, p.eps, sredA, sredB, y,
                                         PROF && P.trace ? P.trace + (size_t) cta * P.prof_n + (l * 5 + (lm ? 0 : phs)) * M5_PK : nullptr);
                } else {
#pragma unroll
                    for (int i = 0; i < 8; i++) y[i] = v[i];
                }
                M5PROF(phs, 4);
                if (ISF) m5_record8_f16(y, rec);
                else m5_quant8<FMT>(y, rec, D.off_n, D.off_dd, D.off_s, p.code_off);
                M5PROF(phs, 5);
            }
        } else if (kind == 3 && ISF) {
            // 4096 fp16 GELU outputs, two per word: prep thread t polls the 16 words of elements 32 t .. 32 t + 31 and writes them into the
            // K = 4096 record in lane order
            if (tid < M5_PT) {
                const unsigned long long * src = P.xch + M5_E4F(lx & 1) + (size_t) rep * M5_E4W + (size_t) tid * 16;
                unsigned long long w[16];
                unsigned spins = 0; long long t0 = 0;
                for (;;) {
#pragma unroll
                    for (int i = 0; i < 8; i++)
                        asm volatile("ld.volatile.global.v2.u64 {%0,%1}, [%2];" : "=l"(w[2 * i]), "=l"(w[2 * i + 1]) : "l"(src + 2 * i) : "memory");
                    bool ok = true;
#pragma unroll
                    for (int i = 0; i < 16; i++) ok = ok && (uint32_t) (w[i] >> 32) == tag;
                    if (ok) break;
                    if ((++spins & 1023u) == 0) { if (t0 == 0) t0 = clock64(); else if (m5_give_up(err, wcode | 1, t0)) break; }
                }
#pragma unroll
                for (int j = 0; j < 16; j++) {
                    const uint32_t v = (uint32_t) w[j];
                    ((uint16_t *) rec)[m5_h_pos(32 * tid + 2 * j)] = (uint16_t) (v & 0xFFFF);
                    ((uint16_t *) rec)[m5_h_pos(32 * tid + 2 * j + 1)] = (uint16_t) (v >> 16);
                }
            }
            M5PROF(phs, 3);
        } else if (kind == 3) {
            // 128 blocks x 10 words: prep thread t polls block t (5 loads in flight) and scatters it into the K = 4096 record
            if (tid < M5_PT) {
                const unsigned long long * src = X + M5_E4 + (size_t) rep * (M5_NB_F * 10) + (size_t) tid * 10;
                unsigned long long w[10];
                unsigned spins = 0; long long t0 = 0;
                for (;;) {
#pragma unroll
                    for (int i = 0; i < 5; i++)
                        asm volatile("ld.volatile.global.v2.u64 {%0,%1}, [%2];" : "=l"(w[2 * i]), "=l"(w[2 * i + 1]) : "l"(src + 2 * i) : "memory");
                    bool ok = true;
#pragma unroll
                    for (int i = 0; i < 10; i++) ok = ok && (uint32_t) (w[i] >> 32) == tag;
                    if (ok) break;
                    if ((++spins & 1023u) == 0) { if (t0 == 0) t0 = clock64(); else if (m5_give_up(err, wcode | 1, t0)) break; }
                }
                const int b = tid, g = b >> 2, i = b & 3;
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    const uint32_t v = (uint32_t) w[k];
                    ((uint32_t *) rec)[(g * 8 + k) * 4 + i] = v;
                    if (p.code_off) ((int *) (rec + D.off_n))[(g * 8 + k) * 4 + i] = -p.code_off * __dp4a((int) v, 0x01010101, 0);
                }
                ((float *) (rec + D.off_dd))[b] = __uint_as_float((uint32_t) w[8]);
                ((float *) (rec + D.off_s))[b] = __uint_as_float((uint32_t) w[9]);
            }
            M5PROF(phs, 3);
        }
        M5PROF(phs, 6);
        __syncthreads();                                           // record and weights complete; previous tile's shared scratch free
        if (tid == ISSUER) fire_tile(tn - 1 + P.nslot, nxt);
        M5PROF(lm ? 0 : phs, 1);
        // ---- dot products
        constexpr bool HASMF = (FMT == BG_Q4_1 || FMT == BG_Q5_1);
        float dot = 0.f;
        if constexpr (ISF) {                                      // (if constexpr: a by-reference lambda that merely EXISTS in the quantised instantiation cost it 1.4 %)
            // F16: finishes one row of the direct kinds (P1, P4, lm_head); called by lane 0 of the warp that computed the row's dot
            auto finish_direct = [&](int row_l, float dotv, float b) {
                if (kind == 0) {
                    const int mat = row_l >> 4, idx = row_l & 15;
                    float t = __fadd_rn(b, dotv);
                    if (mat == 0) { t = __fmul_rn(t, p.qscale); m5_bcast32(&s_q[rank * M5_HR + idx], __float_as_uint(t)); }
                    else if (mat == 1) { kc[(size_t) pos * M5_D + hrow0 + idx] = t; m5_bcast32(&s_kn[rank * M5_HR + idx], __float_as_uint(t)); }
                    else { vc[(size_t) pos * M5_D + hrow0 + idx] = t; s_vn[idx] = t; }
                } else if (kind == 2) {
                    s_blk[row_l] = bg_h2f(p.gelu[bg_f2h(__fadd_rn(b, dotv))]);
                } else {
                    const int r_own = rbase + row_l;
                    p.logits[r_own] = dotv;
                    if (dotv > best || (dotv == best && r_own < bi)) { best = dotv; bi = r_own; }
                }
            };
            if (!relay) {
#pragma unroll 1
                for (int j = 0; j < 3; j++) {
                    const int rr = warp + M5_NW * j;
                    if (rr >= rt) break;
                    const float dv = m5_row_dot_f16<M5_D / 256>(wt + (size_t) rr * D.stride, rec);
                    if (lane == 0) finish_direct(rr, dv, fbias[j]);
                }
            } else if (warp < 8) {
                dot = Kff ? m5_row_dot_f16<M5_FF / 256>(wt + (size_t) warp * D.stride, rec) : m5_row_dot_f16<M5_D / 256>(wt + (size_t) warp * D.stride, rec);
            }
        } else if (!relay) {
            if (4 * warp < rt) dot = m5_row_dot<FMT, M5_D / 128>(wt + (size_t) min(myrow, rt - 1) * D.stride, rec, D);
        } else if (kind == 1) {
            if (warp < 8) dot = m5_row_dot_relay<FMT, M5_D / 128, true>(wt + (size_t) warp * D.stride, rec, D);
        } else if (M5_HAS_FC2_RELAY(FMT) && P.sm_p < 0) {
            if (warp < 8) dot = m5_row_dot_relay<FMT, M5_FF / 128, false>(wt + (size_t) warp * D.stride, rec, D);
            else if (HASMF) dot = m5_summs_chain(wt + (size_t) (warp - 8) * D.stride, rec, D);     // the row's summs chain, on an idle warp
        } else {
            // fc2, K = 4096, 8 rows: the chain of 128 dependent fmas per running sum is the floor (512 cycles); everything else is
            // spread over all 16 warps first.  Phase A: thread (g = group of 4 blocks, l) does the integer dots of rows r0, r0+2, r0+4,
            // r0+6 (its activation words stay in registers) and writes (float) isum as float4 per row; 2 scale products per thread.
            {
                constexpr bool IS8 = (FMT == BG_Q8_0), HASQH = (FMT == BG_Q5_0 || FMT == BG_Q5_1), HASOFF = (FMT == BG_Q4_0 || FMT == BG_Q5_0);
                const int l = tid & 7, g = (tid >> 3) & 31, r0 = tid >> 8, j = l & 3, hi = l >> 2, sh = hi * 4;
                const uint4 aw = *(const uint4 *) (rec + (g * 8 + l) * 16);
                int4 an = make_int4(0, 0, 0, 0);
                if (HASOFF) an = *(const int4 *) (rec + D.off_n + (g * 8 + l) * 16);
                const uint32_t aa[4] = { aw.x, aw.y, aw.z, aw.w };
                const int nn[4] = { an.x, an.y, an.z, an.w };
#pragma unroll
                for (int rr = 0; rr < 4; rr++) {
                    const int row = r0 + 2 * rr;
                    const uint8_t * wrow = wt + (size_t) row * D.stride;
                    const uint4 wq = IS8 ? *(const uint4 *) (wrow + ((g * 2 + hi) * 4 + j) * 16) : *(const uint4 *) (wrow + (g * 4 + j) * 16);
                    uint32_t qh = 0;
                    if (HASQH) qh = *(const uint32_t *) (wrow + D.off_qh + (g * 4 + j) * 4);
                    const uint32_t ww[4] = { wq.x, wq.y, wq.z, wq.w };
                    float pv[4];
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        uint32_t code;
                        if (IS8) code = ww[i];
                        else {
                            code = (ww[i] >> sh) & 0x0F0F0F0Fu;
                            if (HASQH) code |= bg_spread4((qh >> (8 * i + sh)) & 0xFu);
                        }
                        pv[i] = (float) __dp4a((int) code, (int) aa[i], nn[i]);
                    }
                    *(float4 *) (s_p + (size_t) (row * 8 + l) * M5_PS + 4 * g) = make_float4(pv[0], pv[1], pv[2], pv[3]);
                }
#pragma unroll
                for (int e = 0; e < 2; e++) {
                    const int b = tid & 127, row = (tid >> 7) + 4 * e;
                    const uint8_t * wrow = wt + (size_t) row * D.stride;
                    s_s[row * M5_NB_F + b] = __fmul_rn(bg_h2f(*(const uint16_t *) (wrow + D.off_d + b * 2)), *(const float *) (rec + D.off_dd + b * 4));
                    if (HASMF) s_s[(8 + row) * M5_NB_F + b] = bg_h2f(*(const uint16_t *) (wrow + D.off_m + b * 2));
                }
            }
            __syncthreads();
            // Phase B: warps 0, 1, lane = (row, running sum): the chains in block order (+ summs), then hsum_float_8
            if (warp < 2) {
                const int row = 4 * warp + (lane >> 3), l = lane & 7;
                const float * pp = s_p + (size_t) (row * 8 + l) * M5_PS;
                const float * ss = s_s + row * M5_NB_F;
                float acc = 0.f, summ = 0.f;
#pragma unroll 8
                for (int g = 0; g < M5_NB_F / 4; g++) {
                    const float4 pv = *(const float4 *) (pp + 4 * g);
                    const float4 sv = *(const float4 *) (ss + 4 * g);
                    acc = fmaf(sv.x, pv.x, acc); acc = fmaf(sv.y, pv.y, acc); acc = fmaf(sv.z, pv.z, acc); acc = fmaf(sv.w, pv.w, acc);
                    if (HASMF) {
                        const float4 mv = *(const float4 *) (ss + 8 * M5_NB_F + 4 * g);
                        const float4 sa = *(const float4 *) (rec + D.off_s + g * 16);
                        summ = fmaf(mv.x, sa.x, summ); summ = fmaf(mv.y, sa.y, summ); summ = fmaf(mv.z, sa.z, summ); summ = fmaf(mv.w, sa.w, summ);
                    }
                }
                float r = __fadd_rn(acc, __shfl_xor_sync(FULLMASK, acc, 4));
                r = __fadd_rn(r, __shfl_xor_sync(FULLMASK, r, 2));
                r = __fadd_rn(r, __shfl_xor_sync(FULLMASK, r, 1));
                if (HASMF) r = __fadd_rn(r, summ);
                dot = r;
            }
        }
        if (!lm) M5PROF(phs, 10);
        // ---- epilogues
        if (kind == 0) {
            if (!ISF && owner) {
                const int mat = myrow >> 4, idx = myrow & 15;
                float t = __fadd_rn(bias, dot);
                if (mat == 0) { t = __fmul_rn(t, p.qscale); m5_bcast32(&s_q[rank * M5_HR + idx], __float_as_uint(t)); }
                else if (mat == 1) { kc[(size_t) pos * M5_D + hrow0 + idx] = t; m5_bcast32(&s_kn[rank * M5_HR + idx], __float_as_uint(t)); }
                else { vc[(size_t) pos * M5_D + hrow0 + idx] = t; s_vn[idx] = t; }
            }
        } else if (kind == 1 || kind == 3) {
            // gather the 8 rows; warp 0 finishes them (bias, residual) and writes 8 rows x 8 replicas as consecutive words
            const bool two_phase = !ISF && kind == 3 && (!M5_HAS_FC2_RELAY(FMT) || P.sm_p >= 0);
            if (two_phase) { if (warp < 2 && (lane & 7) == 0) s_blk[4 * warp + (lane >> 3)] = dot; }
            else if (warp < 8) { if (lane == 24) s_blk[warp] = dot; }
            else if (kind == 3 && HASMF && lane == 0) s_blk[warp] = dot;
            __syncthreads();
            if (warp == 0) {
                float v = 0.f;
                if (lane < 8) {
                    float d = s_blk[lane];
                    if (kind == 3 && HASMF && !two_phase) d = __fadd_rn(d, s_blk[8 + lane]);
                    v = kind == 1 ? __fadd_rn(__fadd_rn(d, bias), s_x[rbase + lane]) : __fadd_rn(__fadd_rn(bias, d), s_x1[rbase + lane]);
                }
                const float mine = __shfl_sync(FULLMASK, v, lane & 7);
                unsigned long long * dst = X + (kind == 1 ? M5_E3 : M5_E5) + rbase + (lane & 7);
                m4_put(dst + (size_t) (lane >> 3) * M5_D, __float_as_uint(mine), tag);
                m4_put(dst + (size_t) ((lane >> 3) + 4) * M5_D, __float_as_uint(mine), tag);
            }
        } else if (kind == 2) {
            if (!ISF && owner) s_blk[myrow] = bg_h2f(p.gelu[bg_f2h(__fadd_rn(bias, dot))]);
            __syncthreads();
            if (ISF) {
                // the CTA's 32 GELU outputs (fp16 values) as 16 words, two per word, to the 8 replicas: 128 words by warp 0
                if (tid < 32) {
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        const int ix = lane + 32 * i, r = ix >> 4, wd = ix & 15;
                        const uint32_t pay = (uint32_t) bg_f2h(s_blk[2 * wd]) | ((uint32_t) bg_f2h(s_blk[2 * wd + 1]) << 16);
                        m4_put(P.xch + M5_E4F(lx & 1) + (size_t) r * M5_E4W + (size_t) cta * 16 + wd, pay, tag);
                    }
                }
            } else if (tid < 32) m4_quant_publish<FMT>(s_blk, X + M5_E4 + (size_t) cta * 10, M5_NB_F * 10, tag);
        } else {
            if (!ISF && owner) {
                const int r_own = rbase + myrow;
                p.logits[r_own] = dot;
                if (dot > best || (dot == best && r_own < bi)) { best = dot; bi = r_own; }
            }
        }
        if (!lm) M5PROF(phs, 2);
        // ================= attention: cluster `head`; this CTA scores T/4 positions and reduces V for its 16 columns =================
        if (kind == 0 && is_head) {
            m5_cluster_arrive();                                   // q, k of this CTA stored into the 4 CTAs of the cluster
            const float * Kb = kc + head * M5_DK;
            const float * Vb = vc + hrow0;
            // K rows of this warp: positions 512 pass + 32 warp + 8 rank + u, u < 8 (all but the new row are in the cache already)
            const int tb0 = 32 * warp + 8 * rank;
            float kr[8][2];
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const int t = tb0 + u;
                kr[u][0] = (t < T - 1) ? __ldcg(Kb + (size_t) t * M5_D + lane) : 0.0f;
                kr[u][1] = (t < T - 1) ? __ldcg(Kb + (size_t) t * M5_D + 32 + lane) : 0.0f;
            }
            m5_cluster_wait();
            M5PROF(1, 3);
            const float q0 = s_q[lane], q1 = s_q[32 + lane];
            const float kn0 = s_kn[lane], kn1 = s_kn[32 + lane];
#pragma unroll 1
            for (int tb = tb0; tb < T; tb += 512) {
                if (tb != tb0) {
#pragma unroll
                    for (int u = 0; u < 8; u++) {
                        const int t = tb + u;
                        kr[u][0] = (t < T - 1) ? __ldcg(Kb + (size_t) t * M5_D + lane) : 0.0f;
                        kr[u][1] = (t < T - 1) ? __ldcg(Kb + (size_t) t * M5_D + 32 + lane) : 0.0f;
                    }
                }
                float s[8];
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    const bool isnew = (tb + u) == T - 1;
                    float a = 0.0f;
                    a = fmaf(isnew ? kn0 : kr[u][0], q0, a); a = fmaf(isnew ? kn1 : kr[u][1], q1, a);
                    s[u] = a;
                }
                // 8 reduce trees (xor 16, 8, 4, 1, 2: GGML_F32x8_REDUCE) -- the first three levels as a transposing butterfly
                float a4[4], a2[2];
                const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
#pragma unroll
                for (int i = 0; i < 4; i++) { const float mine = b4 ? s[4 + i] : s[i], send = b4 ? s[i] : s[4 + i]; a4[i] = __fadd_rn(mine, __shfl_xor_sync(FULLMASK, send, 16)); }
#pragma unroll
                for (int i = 0; i < 2; i++) { const float mine = b3 ? a4[2 + i] : a4[i], send = b3 ? a4[i] : a4[2 + i]; a2[i] = __fadd_rn(mine, __shfl_xor_sync(FULLMASK, send, 8)); }
                const float mine = b2 ? a2[1] : a2[0], send = b2 ? a2[0] : a2[1];
                float dotv = __fadd_rn(mine, __shfl_xor_sync(FULLMASK, send, 4));
                dotv = __fadd_rn(dotv, __shfl_xor_sync(FULLMASK, dotv, 1));
                dotv = __fadd_rn(dotv, __shfl_xor_sync(FULLMASK, dotv, 2));
                const int u = (b4 ? 4 : 0) | (b3 ? 2 : 0) | (b2 ? 1 : 0);
                const int t = tb + u;
                if (t < T && (lane & 3) == 0) m5_bcast32(&sc[t], __float_as_uint(dotv));
            }
            m5_cluster_arrive();                                   // this CTA's scores stored into the 4 CTAs of the cluster
            // V rows: unit (r32 = t % 32, column c), all 512 threads, up to 32 loads in flight while the scores travel
            const int np = T & ~31;
            const int vr = tid >> 4, vcn = tid & 15;
            float vv[32];
#pragma unroll
            for (int k = 0; k < 32; k++) {
                if (32 * k >= np) break;                           // CTA-uniform: nothing is issued for positions that do not exist
                const int t = 32 * k + vr;
                vv[k] = (t != T - 1) ? __ldcg(Vb + (size_t) t * M5_D + vcn) : 0.0f;
            }
            if (np + vr < T - 1) tailv[tid] = __ldcg(Vb + (size_t) (np + vr) * M5_D + vcn);        // tail rows np .. T-2 (at most 30)
            m5_cluster_wait();
            M5PROF(1, 5);
            // softmax over sc[0..T) on 4 warps, 8 scores per thread: max, fp16-table exp, sum, scale (ggml.c:12955-12974).  The reference
            // adds the fp16-valued exponentials in double; they are multiples of 2^-24 not above 1, so the sum is exact in any order --
            // here as integers (two warp-wide redux.add instead of ten 64-bit shuffles), converted to double once.
            if (tid < M5_PT) {
                const float4 xa4 = *(const float4 *) (sc + 8 * tid), xb4 = *(const float4 *) (sc + 8 * tid + 4);
                float x[8] = { xa4.x, xa4.y, xa4.z, xa4.w, xb4.x, xb4.y, xb4.z, xb4.w };
                float mx = -INFINITY;
#pragma unroll
                for (int i = 0; i < 8; i++) { if (8 * tid + i >= T) x[i] = -INFINITY; mx = fmaxf(mx, x[i]); }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(FULLMASK, mx, o));
                if (lane == 0) sredF[warp] = mx;
                m5_bar_prep();
                mx = fmaxf(fmaxf(sredF[0], sredF[1]), fmaxf(sredF[2], sredF[3]));
                float e[8];
                unsigned lo = 0, hi = 0;
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    e[i] = 0.f;
                    if (8 * tid + i < T) {
                        e[i] = bg_h2f(p.exp_tab[bg_f2h(__fsub_rn(x[i], mx))]);
                        const unsigned u = __float2uint_rn(__fmul_rn(e[i], 16777216.0f));       // exact: fp16 values are multiples of 2^-24
                        lo += u & 0xFFFFu; hi += u >> 16;
                    }
                }
                lo = __reduce_add_sync(FULLMASK, lo); hi = __reduce_add_sync(FULLMASK, hi);
                if (lane == 0) s_isum[warp] = ((unsigned long long) hi << 16) + lo;
                m5_bar_prep();
                const unsigned long long tot = (s_isum[0] + s_isum[1]) + (s_isum[2] + s_isum[3]);
                const float inv = (float) (1.0 / ((double) (long long) tot * (1.0 / 16777216.0)));
#pragma unroll
                for (int i = 0; i < 8; i++) if (8 * tid + i < T) sc[8 * tid + i] = __fmul_rn(e[i], inv);
            }
            __syncthreads();                                       // probabilities, s_vn and tailv visible to everyone
            M5PROF(1, 4);
            // V: running sum r32 of column c over t = r32, r32 + 32, ... < np (ggml_vec_dot_f32's 32 lanes, ggml.c:2372-2407)
            {
                const float vnew = s_vn[vcn];
                float acc = 0.f;
#pragma unroll
                for (int k = 0; k < 32; k++) {
                    if (32 * k >= np) break;
                    const int t = 32 * k + vr;
                    acc = fmaf((t == T - 1) ? vnew : vv[k], sc[t], acc);
                }
                red[vr * M5_HR + vcn] = acc;
            }
            __syncthreads();
            if (tid < 32) {
                float sumf = 0.f;
                if (tid < M5_HR) {
                    float x0[8];
#pragma unroll
                    for (int l8 = 0; l8 < 8; l8++) {
                        const float a02 = __fadd_rn(red[(0 * 8 + l8) * M5_HR + tid], red[(2 * 8 + l8) * M5_HR + tid]);
                        const float a13 = __fadd_rn(red[(1 * 8 + l8) * M5_HR + tid], red[(3 * 8 + l8) * M5_HR + tid]);
                        x0[l8] = __fadd_rn(a02, a13);
                    }
                    const float t0 = __fadd_rn(x0[0], x0[4]), t1 = __fadd_rn(x0[1], x0[5]);
                    const float t2 = __fadd_rn(x0[2], x0[6]), t3 = __fadd_rn(x0[3], x0[7]);
                    sumf = __fadd_rn(__fadd_rn(t0, t1), __fadd_rn(t2, t3));
                    // the as-built scalar tail of ggml_vec_dot_f32: products unfused in groups of 4, then <= 3 fused
                    const int nv = np + ((T - np) & ~3);
                    int t = np;
#pragma unroll 4
                    for (; t < nv; t++) sumf = __fadd_rn(sumf, __fmul_rn((t == T - 1) ? s_vn[tid] : tailv[(t - np) * M5_HR + tid], sc[t]));
#pragma unroll 1
                    for (; t < T;  t++) sumf = fmaf((t == T - 1) ? s_vn[tid] : tailv[(t - np) * M5_HR + tid], sc[t], sumf);
                }
                // lane = (replica, column): 16 columns x 8 replicas, 4 stores per lane
                const float mine = __shfl_sync(FULLMASK, sumf, lane & 15);
                unsigned long long * dst = X + M5_E2 + hrow0 + (lane & 15);
#pragma unroll
                for (int r = lane >> 4; r < M5_R; r += 2) m4_put(dst + (size_t) r * M5_D, __float_as_uint(mine), tag);
            }
            M5PROF(1, 2);
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
        for (int i = 1; i < M5_NW; i++) if (s_cv[i] > best || (s_cv[i] == best && s_ci[i] < bi)) { best = s_cv[i]; bi = s_ci[i]; }
        p.cand_val[cta] = best; p.cand_idx[cta] = bi;
    }
    // ---- sampler tail
    if constexpr (TK) if (P.tk_k > 0)
        m5_sampler_tail(P.tk_ticket, p.logits, p.n_vocab, P.tk_k, p.cand_val, s_w, P.tk_pk + (size_t) (P.tk_seq & 1u) * P.tk_stride, err, P.tk_seq, P.tk_full);
    { const int l = p.n_layer; M5PROF(0, 2); }
    if (PROF && P.trace && tid == 0) m4_calibrate(P.trace + (size_t) M5_NC * P.prof_n + 4 * cta + 2);
}
