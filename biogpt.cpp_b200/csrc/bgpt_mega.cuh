// bgpt_mega.cuh -- one persistent kernel per decoded token.
//
// The N=1 decode step of BioGPT is a chain of ~120 dependent small operators:
// as separate launches each costs ~10 us of pure latency (profiles/r1_v0*), 50x more than the
// 30 us the weights need to stream from HBM.  k_mega runs the whole step in ONE launch: a
// resident grid of CTAs (one per SM), every operator a *phase* executed by all CTAs on their
// slice of the rows, phases separated by a grid-wide barrier.
//
//   phase per layer                                          split over CTAs            barrier
//   P1  LN0 + q,k,v projections (+bias, q scale, KV append)   3*d_model rows (8-row tiles)  yes
//   P2  attention; output quantised by its producer          (head, 32-column part)        yes
//   P3  out_proj + bias + residual                            d_model rows                  yes
//   P4  LN1 + fc1 + bias + GELU; output quantised by producer d_ff rows (32-row tiles)      yes
//   P5  fc2 + bias + residual                                 d_model rows                  yes
//   final: LN + lm_head (+ per-CTA argmax candidates)         n_vocab rows                  -
//
// The step is latency-bound, so the kernel is organised around the dependent chain, not around
// bandwidth:
//   * the weights a CTA needs in phase k+1 do not depend on activations: they are loaded into
//     registers BEFORE the barrier that ends phase k, so their HBM latency overlaps the wait;
//   * every CTA recomputes the per-token LayerNorm of its phase from x (4 KB from L2) instead of
//     waiting for one more publish/consume round trip; attention and GELU outputs are quantised
//     by the CTA that produced them (one 32-element block each) and published as ready-made
//     activation records;
//   * cross-CTA data is read with ld.global.cg (L2) after an acquire on per-CTA flags; nothing
//     that another CTA wrote is ever served from L1.
//
// The dot products keep the lane order of bgpt_kernels.cuh but are split for latency:
//   phase A: thread (row, g, j) owns one uint4 = the 4-byte groups of running sums j and j+4 for
//            the blocks 4g..4g+3: exact integer work (dp4a), 8 products p = (float) isum and the
//            4 scales s = d_w*d_a go to shared memory -- fully parallel;
//   phase B: thread (row, l) walks acc = fma(s_b, p_b, acc) over the blocks in order from shared
//            memory -- the only sequential part, 1 FMA per block.
// F16 weights need no integer phase: thread (row, lane) is running-sum lane `lane` of its row.
#pragma once
#include "bgpt_kernels.cuh"

#define MEGA_NT 512          // threads per CTA
#define MEGA_RT 32           // rows per tile
#define MEGA_UMAX 2          // phase-A units per thread whose weights are prefetched into registers

struct MegaLayer {
    const uint8_t *q_w, *k_w, *v_w, *o_w, *fc1_w, *fc2_w;
    const float *q_b, *k_b, *v_b, *o_b, *ln0_w, *ln0_b, *ln1_w, *ln1_b, *fc1_b, *fc2_b;
};

struct MegaParams {
    int d, ff, n_head, dk, n_layer, n_vocab, n_positions, n_pos_rows, wtype;
    float emb_scale, qscale, eps;
    int Gd, stride_d, offqh_d, offd_d, offm_d;          // row layout for K = d_model
    int Gf, stride_f, offqh_f, offd_f, offm_f;          // row layout for K = d_ff
    int actb_d, offn_d, offdd_d, offs_d;                // activation record layout, K = d_model
    int actb_f, offn_f, offdd_f, offs_f;                // K = d_ff
    int code_off;
    const MegaLayer * layers;
    const uint8_t * embed_tok; const uint8_t * embed_pos; const uint8_t * lm_head;
    const float * lnf_w; const float * lnf_b;
    const uint16_t * gelu; const uint16_t * exp_tab;
    float * kcache; float * vcache;                     // [layer][pos][d] of stream 0
    float * x; float * x1; float * q; float * att; float * hff; float * logits;
    uint8_t * rec_att; uint8_t * rec_hff;               // activation records published by their producers
    unsigned long long * flags; unsigned long long epoch0;   // grid barrier counter and its value at launch
    const int * tok;                                    // device: input token id (use_cand == 0)
    int use_cand; float * cand_val; int * cand_idx; int n_cand;   // 1: argmax over the candidates of the previous launch
    int tok_imm;                                        // use_cand == 2: the input token id travels in the kernel parameters
    int * idlog; int log_slot;                          // idlog[log_slot] = input token when log_slot >= 0
    int n_past;
    int attn_parts;                                     // CTAs per head in the attention phase (column split)
    int att_prequant;                                   // 1: attention CTAs publish quantised blocks (d_kv/parts == 32)
    long long * prof;                                   // optional: clock64 stamps of the last CTA, [n_layer+1][5][6]
    int sm_row, sm_act, sm_p, sm_s, sm_m, sm_attn, sm_total;   // shared-memory carve-up (bytes)
};

// ---- grid barrier ---------------------------------------------------------------------------
// One 64-bit counter in L2: arrive = red.release (no return value), wait = thread 0 polls with
// ld.acquire until all gridDim.x CTAs of this epoch arrived.  Measured on B200 against five
// other schemes (tools/barrier_bench.py): 1.25 us per barrier, the fastest; per-CTA flag arrays
// polled by many threads are 3-4x slower (L2 request contention).
// The barrier is split: arrive, then the caller issues the register prefetch of the next
// phase's weights, then wait -- a release must not be stuck behind the arriving thread's own
// outstanding loads, and the loads overlap the wait.
__device__ __forceinline__ unsigned long long mega_ld_acquire(const unsigned long long * p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void mega_barrier_arrive(unsigned long long * ctr) {
    __syncthreads();
    if (threadIdx.x == 0) asm volatile("red.release.gpu.global.add.u64 [%0], 1;" :: "l"(ctr) : "memory");
}
__device__ __forceinline__ void mega_barrier_wait(unsigned long long * ctr, unsigned long long target) {
    if (threadIdx.x == 0) while (mega_ld_acquire(ctr) < target) { }
    __syncthreads();
}

// L2 prefetch of a contiguous byte range (16-byte aligned): issue the HBM read now
__device__ __forceinline__ void mega_prefetch_l2(const void * ptr, size_t bytes64) {
    const unsigned bytes = (unsigned) bytes64;                 // slices are far below 4 GB
    const unsigned n = (bytes + 8191u) >> 13;                  // 8 KB chunks spread over the threads
    for (unsigned i = threadIdx.x; i < n; i += MEGA_NT) {
        const unsigned off = i << 13;
        const unsigned sz = ((bytes - off) < 8192u ? (bytes - off) : 8192u) & ~15u;
        if (sz) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"((const uint8_t *) ptr + off), "r"(sz) : "memory");
    }
}

// L2 prefetch by 128-byte lines through the LSU (cheap to issue, unlike the bulk form): lines
// [0, bytes/128) of `ptr`, thread t takes lines t, t+MEGA_NT, ...
__device__ __forceinline__ void mega_prefetch_lines(const void * ptr, unsigned bytes) {
    for (unsigned off = threadIdx.x * 128u; off < bytes; off += MEGA_NT * 128u)
        asm volatile("prefetch.global.L2 [%0];" :: "l"((const uint8_t *) ptr + off));
}

// rows [r0, r1) of an M-row phase owned by this CTA: balanced contiguous runs of `tile`-row tiles
__device__ __forceinline__ void mega_my_rows(int M, int tile, int & r0, int & r1) {
    const int nt = (M + tile - 1) / tile;
    const int t0 = (int) (((unsigned) blockIdx.x * (unsigned) nt) / gridDim.x);          // nt * gridDim.x < 2^31
    const int t1 = (int) ((((unsigned) blockIdx.x + 1u) * (unsigned) nt) / gridDim.x);
    r0 = t0 * tile; r1 = t1 * tile < M ? t1 * tile : M;
}

// ---- matmul phase descriptor ------------------------------------------------------------------
struct MMDesc {
    const uint8_t * W0; const uint8_t * W1; const uint8_t * W2;   // stacked matrices (q,k,v) or W0 only
    int rows_per;             // rows of one matrix
    int r0, r1;               // this CTA's rows in the stacked row space
    int G, gsh;               // groups per row; log2(G) or -1
    int stride, off_qh, off_d, off_m;
    int off_n, off_dd, off_s; // activation record offsets
    int RT;                   // rows per tile
};
__device__ __forceinline__ MMDesc mega_mm_desc(const MegaParams & p, const uint8_t * W0, const uint8_t * W1, const uint8_t * W2,
                                               int rows_per, int r0, int r1, bool Kff) {
    MMDesc D;
    D.W0 = W0; D.W1 = W1; D.W2 = W2; D.rows_per = rows_per; D.r0 = r0; D.r1 = r1;
    D.G = Kff ? p.Gf : p.Gd; D.gsh = (D.G & (D.G - 1)) == 0 ? 31 - __clz(D.G) : -1;
    D.stride = Kff ? p.stride_f : p.stride_d;
    D.off_qh = Kff ? p.offqh_f : p.offqh_d; D.off_d = Kff ? p.offd_f : p.offd_d; D.off_m = Kff ? p.offm_f : p.offm_d;
    D.off_n = Kff ? p.offn_f : p.offn_d; D.off_dd = Kff ? p.offdd_f : p.offdd_d; D.off_s = Kff ? p.offs_f : p.offs_d;
    int RT = MEGA_RT;
    if (p.sm_s > p.sm_p) { const int cap = (p.sm_s - p.sm_p) / (8 * (4 * D.G + 4) * 4); RT = cap < RT ? cap : RT; RT = RT < 1 ? 1 : RT; }
    D.RT = RT;
    return D;
}
__device__ __forceinline__ const uint8_t * mega_row_ptr(const MMDesc & D, int r) {
    const int mat = (r >= D.rows_per) + (r >= 2 * D.rows_per);
    const uint8_t * W = mat == 0 ? D.W0 : (mat == 1 ? D.W1 : D.W2);
    return W + (size_t) (r - mat * D.rows_per) * D.stride;
}
__device__ __forceinline__ void mega_unit(const MMDesc & D, int u, int & row, int & g, int & j) {
    j = u & 3;
    if (D.gsh >= 0) { g = (u >> 2) & (D.G - 1); row = u >> (2 + D.gsh); }
    else            { g = (u >> 2) % D.G;       row = u / (4 * D.G); }
}

// weights of one phase-A unit, held in registers between the prefetch and the compute
template <int FMT> struct WUnit { uint4 w0, w1; uint32_t qh; uint2 dh, mh; };
template <int FMT> struct WRegs { WUnit<FMT> u[MEGA_UMAX]; };

template <int FMT>
__device__ __forceinline__ void mega_load_unit(WUnit<FMT> & W, const MMDesc & D, int t0, int u) {
    constexpr bool IS8   = (FMT == BG_Q8_0);
    constexpr bool HASQH = (FMT == BG_Q5_0 || FMT == BG_Q5_1);
    constexpr bool HASM  = (FMT == BG_Q4_1 || FMT == BG_Q5_1);
    int row, g, j; mega_unit(D, u, row, g, j);
    const uint8_t * wrow = mega_row_ptr(D, t0 + row);
    if (IS8) {
        W.w0 = ldg_stream128(wrow + (size_t) ((g * 2 + 0) * 4 + j) * 16);
        W.w1 = ldg_stream128(wrow + (size_t) ((g * 2 + 1) * 4 + j) * 16);
    } else {
        W.w0 = ldg_stream128(wrow + (size_t) (g * 4 + j) * 16);
        if (HASQH) W.qh = ldg_stream32(wrow + D.off_qh + g * 16 + j * 4);
    }
    if (j == 0) W.dh = ldg_stream64(wrow + D.off_d + g * 8);
    if (HASM && j == 1) W.mh = ldg_stream64(wrow + D.off_m + g * 8);
}
template <int FMT>
__device__ __forceinline__ void mega_load_units(WRegs<FMT> & R, const MMDesc & D, int t0, int rt) {
    const int total = rt * D.G * 4;
#pragma unroll
    for (int k = 0; k < MEGA_UMAX; k++) {
        const int u = threadIdx.x + k * MEGA_NT;
        if (u < total) mega_load_unit<FMT>(R.u[k], D, t0, u);
    }
}

// phase A for one unit whose weights are in W
template <int FMT>
__device__ __forceinline__ void mega_compute_unit(const WUnit<FMT> & W, const MMDesc & D, int u, const uint8_t * s_act,
                                                  float * s_p, float * s_s, float * s_m) {
    constexpr bool IS8   = (FMT == BG_Q8_0);
    constexpr bool HASQH = (FMT == BG_Q5_0 || FMT == BG_Q5_1);
    constexpr bool HASM  = (FMT == BG_Q4_1 || FMT == BG_Q5_1);
    constexpr bool HASOFF = (FMT == BG_Q4_0 || FMT == BG_Q5_0);
    const int nbp = D.G * 4, PS = nbp + 4;
    int row, g, j; mega_unit(D, u, row, g, j);
    uint32_t lo[4], hi[4];
    if (IS8) {
        lo[0] = W.w0.x; lo[1] = W.w0.y; lo[2] = W.w0.z; lo[3] = W.w0.w;
        hi[0] = W.w1.x; hi[1] = W.w1.y; hi[2] = W.w1.z; hi[3] = W.w1.w;
    } else {
        const uint32_t ww[4] = { W.w0.x, W.w0.y, W.w0.z, W.w0.w };
#pragma unroll
        for (int i = 0; i < 4; i++) {
            lo[i] = ww[i] & 0x0F0F0F0Fu;
            hi[i] = (ww[i] >> 4) & 0x0F0F0F0Fu;
            if (HASQH) {
                const uint32_t hb = (W.qh >> (8 * i)) & 0xFFu;
                lo[i] |= bg_spread4(hb & 0xFu);
                hi[i] |= bg_spread4(hb >> 4);
            }
        }
    }
    const uint4 a0 = *(const uint4 *) (s_act + (g * 8 + j) * 16);
    const uint4 a1 = *(const uint4 *) (s_act + (g * 8 + j + 4) * 16);
    int4 n0 = make_int4(0, 0, 0, 0), n1 = make_int4(0, 0, 0, 0);
    if (HASOFF) {
        n0 = *(const int4 *) (s_act + D.off_n + (g * 8 + j) * 16);
        n1 = *(const int4 *) (s_act + D.off_n + (g * 8 + j + 4) * 16);
    }
    float4 P0, P1;
    P0.x = (float) __dp4a((int) lo[0], (int) a0.x, n0.x); P1.x = (float) __dp4a((int) hi[0], (int) a1.x, n1.x);
    P0.y = (float) __dp4a((int) lo[1], (int) a0.y, n0.y); P1.y = (float) __dp4a((int) hi[1], (int) a1.y, n1.y);
    P0.z = (float) __dp4a((int) lo[2], (int) a0.z, n0.z); P1.z = (float) __dp4a((int) hi[2], (int) a1.z, n1.z);
    P0.w = (float) __dp4a((int) lo[3], (int) a0.w, n0.w); P1.w = (float) __dp4a((int) hi[3], (int) a1.w, n1.w);
    *(float4 *) (s_p + (size_t) (row * 8 + j) * PS + 4 * g) = P0;
    *(float4 *) (s_p + (size_t) (row * 8 + j + 4) * PS + 4 * g) = P1;
    if (j == 0) {
        const float4 da = *(const float4 *) (s_act + D.off_dd + g * 16);
        float4 S;
        S.x = __fmul_rn(bg_h2f((uint16_t) (W.dh.x & 0xFFFF)), da.x); S.y = __fmul_rn(bg_h2f((uint16_t) (W.dh.x >> 16)), da.y);
        S.z = __fmul_rn(bg_h2f((uint16_t) (W.dh.y & 0xFFFF)), da.z); S.w = __fmul_rn(bg_h2f((uint16_t) (W.dh.y >> 16)), da.w);
        *(float4 *) (s_s + (size_t) row * nbp + 4 * g) = S;
    }
    if (HASM && j == 1) {
        float4 Mv;
        Mv.x = bg_h2f((uint16_t) (W.mh.x & 0xFFFF)); Mv.y = bg_h2f((uint16_t) (W.mh.x >> 16));
        Mv.z = bg_h2f((uint16_t) (W.mh.y & 0xFFFF)); Mv.w = bg_h2f((uint16_t) (W.mh.y >> 16));
        *(float4 *) (s_m + (size_t) row * nbp + 4 * g) = Mv;
    }
}

// phase B: the 8 running sums of each of the rt rows, then the row epilogue.  The owner thread
// of row (c >> 3) is thread c = 8*row: it fetched the epilogue's inputs (bias, residual) before
// phase A, so their latency is hidden behind the whole tile.
template <int FMT, class FinF>
__device__ __forceinline__ void mega_phase_b(const MMDesc & D, int t0, int rt, const uint8_t * s_act,
                                             const float * s_p, const float * s_s, const float * s_m, float2 pv0, FinF fin) {
    constexpr bool HASM = (FMT == BG_Q4_1 || FMT == BG_Q5_1);
    const int nbp = D.G * 4, PS = nbp + 4;
    int it = 0;
    for (int c0 = (threadIdx.x >> 5) << 5; c0 < rt * 8; c0 += MEGA_NT, it++) {
        const int c = c0 + (threadIdx.x & 31);
        const bool valid = c < rt * 8;
        const int cc = valid ? c : rt * 8 - 1;
        const int row = cc >> 3, l = cc & 7;
        const bool owner = valid && l == 0;
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
        // hsum_float_8: (a[l+4] + a[l]) then +2, +1
        float r = __fadd_rn(acc, __shfl_xor_sync(FULLMASK, acc, 4));
        r = __fadd_rn(r, __shfl_xor_sync(FULLMASK, r, 2));
        r = __fadd_rn(r, __shfl_xor_sync(FULLMASK, r, 1));
        if (HASM) r = __fadd_rn(r, summ);
        if (owner) fin(t0 + row, r, pv0);
    }
}

// all tiles of a quantised matmul phase; `cur` holds the prefetched first tile
template <int FMT, class PreF, class FinF>
__device__ __forceinline__ void mega_mm_run_q(const MMDesc & D, WRegs<FMT> & cur, const uint8_t * s_act,
                                              float * s_p, float * s_s, float * s_m, PreF pre, FinF fin) {
    for (int t0 = D.r0; t0 < D.r1; t0 += D.RT) {
        const int rt = (D.r1 - t0) < D.RT ? (D.r1 - t0) : D.RT;
        const int total = rt * D.G * 4;
        const bool has_next = t0 + D.RT < D.r1;
        WRegs<FMT> nxt;
        if (has_next) mega_load_units<FMT>(nxt, D, t0 + D.RT, (D.r1 - t0 - D.RT) < D.RT ? (D.r1 - t0 - D.RT) : D.RT);
        // epilogue inputs of the row this thread will finish (rt*8 <= MEGA_NT: one owner row per thread)
        float2 pv0 = make_float2(0.f, 0.f);
        if ((int) threadIdx.x < rt * 8 && (threadIdx.x & 7) == 0) pv0 = pre(t0 + (int) (threadIdx.x >> 3));
#pragma unroll
        for (int k = 0; k < MEGA_UMAX; k++) {
            const int u = threadIdx.x + k * MEGA_NT;
            if (u < total) mega_compute_unit<FMT>(cur.u[k], D, u, s_act, s_p, s_s, s_m);
        }
        for (int u = threadIdx.x + MEGA_UMAX * MEGA_NT; u < total; u += MEGA_NT) {   // shapes with > UMAX units per thread
            WUnit<FMT> W; mega_load_unit<FMT>(W, D, t0, u);
            mega_compute_unit<FMT>(W, D, u, s_act, s_p, s_s, s_m);
        }
        __syncthreads();
        mega_phase_b<FMT>(D, t0, rt, s_act, s_p, s_s, s_m, pv0, fin);
        __syncthreads();
        if (has_next) cur = nxt;
    }
}

// F16 weights: thread (row, lane) is running-sum lane `lane` of its row (weights come from L2:
// every phase's rows were prefetched one layer ahead)
template <class PreF, class FinF>
__device__ __forceinline__ void mega_mm_run_f16(const MMDesc & D, const uint8_t * s_act, PreF pre, FinF fin) {
    const int lane = threadIdx.x & 31;
    for (int r = D.r0 + (threadIdx.x >> 5); r < D.r1; r += MEGA_NT / 32) {
        const uint8_t * wrow = mega_row_ptr(D, r);
        float2 pv0 = make_float2(0.f, 0.f);
        if (lane == 0) pv0 = pre(r);
        float c = 0.0f;
#pragma unroll 4
        for (int g = 0; g < D.G; g++) {
            const uint4 w = ldg_stream128(wrow + (size_t) (g * 32 + lane) * 16);
            const float4 x0 = *(const float4 *) (s_act + (size_t) ((g * 2 + 0) * 32 + lane) * 16);
            const float4 x1 = *(const float4 *) (s_act + (size_t) ((g * 2 + 1) * 32 + lane) * 16);
            c = fmaf(bg_h2f((uint16_t) (w.x & 0xFFFF)), x0.x, c); c = fmaf(bg_h2f((uint16_t) (w.x >> 16)), x0.y, c);
            c = fmaf(bg_h2f((uint16_t) (w.y & 0xFFFF)), x0.z, c); c = fmaf(bg_h2f((uint16_t) (w.y >> 16)), x0.w, c);
            c = fmaf(bg_h2f((uint16_t) (w.z & 0xFFFF)), x1.x, c); c = fmaf(bg_h2f((uint16_t) (w.z >> 16)), x1.y, c);
            c = fmaf(bg_h2f((uint16_t) (w.w & 0xFFFF)), x1.z, c); c = fmaf(bg_h2f((uint16_t) (w.w >> 16)), x1.w, c);
        }
        float v = c;
        v = __fadd_rn(v, __shfl_xor_sync(FULLMASK, v, 16));
        v = __fadd_rn(v, __shfl_xor_sync(FULLMASK, v, 8));
        v = __fadd_rn(v, __shfl_xor_sync(FULLMASK, v, 4));
        v = __fadd_rn(v, __shfl_xor_sync(FULLMASK, v, 1));
        v = __fadd_rn(v, __shfl_xor_sync(FULLMASK, v, 2));
        if (lane == 0) fin(r, v, pv0);
    }
    __syncthreads();
}

// one 32-element block of f32 values in shared memory -> the quantised record piece of block b.
// Same arithmetic as bg_row_to_record; executed by one full warp (lanes 0..7 own 4 elements each).
__device__ __forceinline__ void mega_quant_block(const float * s_v, int b, int wtype, uint8_t * rec, int off_n, int off_d, int off_s, int code_off) {
    const int lane = threadIdx.x & 31, l = lane & 7;
    const int kind = bg_act_kind(wtype);
    const int g = b >> 2, i = b & 3;
    const float4 v = *(const float4 *) (s_v + 4 * l);
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
    if (lane < 8) {
        ((uint32_t *) rec)[(g * 8 + l) * 4 + i] = ((uint32_t) q0 & 0xFFu) | (((uint32_t) q1 & 0xFFu) << 8) | (((uint32_t) q2 & 0xFFu) << 16) | (((uint32_t) q3 & 0xFFu) << 24);
        ((int32_t *) (rec + off_n))[(g * 8 + l) * 4 + i] = -code_off * s4;
    }
    if (lane == 0) {
        if (kind == ACT_Q8_0) { ((float *) (rec + off_d))[b] = bg_h2f(bg_f2h(d)); ((float *) (rec + off_s))[b] = 0.0f; }
        else                  { ((float *) (rec + off_d))[b] = d; ((float *) (rec + off_s))[b] = __fmul_rn(d, (float) stot); }
    }
}
// F16 weights: the "record" is the fp16-rounded value as f32 in the permuted layout
__device__ __forceinline__ void mega_f16_record_elem(float v, int c, uint8_t * rec) {
    const int gg = c >> 8, e = (c & 255) >> 5, lane = c & 31;
    ((float *) rec)[(((gg * 2 + (e >> 2)) * 32 + lane) * 4) + (e & 3)] = bg_h2f(bg_f2h(v));
}

// ---- attention for `ncol` output columns [c0, c0+ncol) of one head ----------------------------
// Scores and softmax are computed for the whole head by every CTA that shares it (K comes from
// L2 after the first reader); the V reduction -- the 32 running sums per column -- is split by
// column, so no CTA ever needs another CTA's partial result.  Results stay in s_out[ncol].
template <int DK>
__device__ __forceinline__ void mega_attention_cols(const float * s_q, const float * Kb, const float * Vb, int ldkv, int T,
                                                    const uint16_t * __restrict__ exp_tab, float * s_f, int Tmax,
                                                    double * sd, float * sm, int c0, int ncol, float * s_out) {
    float * sc  = s_f;                 // [Tmax]
    float * red = s_f + Tmax;          // [32][ncol]
    float * tailv = red + 32 * DK;     // [<=31][ncol]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr int NWARP = MEGA_NT / 32;
    constexpr int NP = DK & ~31;
    constexpr int NV = NP + ((DK - NP) & ~3);
    constexpr int NQ = NP > 0 ? NP / 32 : 1;
    if (NP > 0 && NP == DK) {
        // 32 rows per warp pass (row u -> position tb + u*NWARP).  Each lane first forms its
        // running-sum lane of all 32 dots, then the 32 reduce trees (xor 16, 8, 4, 1, 2 -- the
        // GGML_F32x8_REDUCE order) run as ONE transposing butterfly: at every level a lane keeps
        // half of its rows and trades the other half, 31 shuffles per 32 rows instead of 160,
        // and lane L ends up owning the finished dot of one row.
        float qreg[NQ];
#pragma unroll
        for (int i = 0; i < NQ; i++) qreg[i] = s_q[i * 32 + lane];
        for (int tb = warp; tb < T; tb += NWARP * 32) {
            float s[32];
#pragma unroll
            for (int half = 0; half < 2; half++) {
                float kv[16][NQ];
#pragma unroll
                for (int u = 0; u < 16; u++)
#pragma unroll
                    for (int i = 0; i < NQ; i++) {
                        const int t = tb + (half * 16 + u) * NWARP;
                        kv[u][i] = (t < T) ? __ldcg(Kb + (size_t) t * ldkv + i * 32 + lane) : 0.0f;
                    }
#pragma unroll
                for (int u = 0; u < 16; u++) {
                    float a = 0.0f;
#pragma unroll
                    for (int i = 0; i < NQ; i++) a = fmaf(kv[u][i], qreg[i], a);
                    s[half * 16 + u] = a;
                }
            }
            float a16[16], a8[8], a4[4], a2[2];
            const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4, b0 = lane & 1, b1 = lane & 2;
#pragma unroll
            for (int i = 0; i < 16; i++) { const float mine = b4 ? s[16 + i] : s[i], send = b4 ? s[i] : s[16 + i]; a16[i] = __fadd_rn(mine, __shfl_xor_sync(FULLMASK, send, 16)); }
#pragma unroll
            for (int i = 0; i < 8; i++)  { const float mine = b3 ? a16[8 + i] : a16[i], send = b3 ? a16[i] : a16[8 + i]; a8[i] = __fadd_rn(mine, __shfl_xor_sync(FULLMASK, send, 8)); }
#pragma unroll
            for (int i = 0; i < 4; i++)  { const float mine = b2 ? a8[4 + i] : a8[i], send = b2 ? a8[i] : a8[4 + i]; a4[i] = __fadd_rn(mine, __shfl_xor_sync(FULLMASK, send, 4)); }
#pragma unroll
            for (int i = 0; i < 2; i++)  { const float mine = b0 ? a4[2 + i] : a4[i], send = b0 ? a4[i] : a4[2 + i]; a2[i] = __fadd_rn(mine, __shfl_xor_sync(FULLMASK, send, 1)); }
            const float mine = b1 ? a2[1] : a2[0], send = b1 ? a2[0] : a2[1];
            const float dot = __fadd_rn(mine, __shfl_xor_sync(FULLMASK, send, 2));
            // row owned by this lane: bit4..bit2 from the lane, then bit1 <- lane bit0, bit0 <- lane bit1
            const int u = (lane & 28) | ((lane & 1) << 1) | ((lane & 2) >> 1);
            const int t = tb + u * NWARP;
            if (t < T) sc[t] = dot;
        }
    } else {
        // generic head sizes (d_kv not a multiple of 32): one row at a time, scalar tail on lane 0
        float qreg[NQ];
#pragma unroll
        for (int i = 0; i < NP / 32; i++) qreg[i] = s_q[i * 32 + lane];
        for (int t = warp; t < T; t += NWARP) {
            const float * kr = Kb + (size_t) t * ldkv;
            float s = 0.0f;
#pragma unroll
            for (int i = 0; i < NP / 32; i++) s = fmaf(__ldcg(kr + i * 32 + lane), qreg[i], s);
            if (NP > 0) {
                s = __fadd_rn(s, __shfl_xor_sync(FULLMASK, s, 16));
                s = __fadd_rn(s, __shfl_xor_sync(FULLMASK, s, 8));
                s = __fadd_rn(s, __shfl_xor_sync(FULLMASK, s, 4));
                s = __fadd_rn(s, __shfl_xor_sync(FULLMASK, s, 1));
                s = __fadd_rn(s, __shfl_xor_sync(FULLMASK, s, 2));
            }
            if (lane == 0) {
#pragma unroll
                for (int i = NP; i < NV; i++) s = __fadd_rn(s, __fmul_rn(__ldcg(kr + i), s_q[i]));
#pragma unroll
                for (int i = NV; i < DK; i++) s = fmaf(__ldcg(kr + i), s_q[i], s);
                sc[t] = s;
            }
        }
    }
    __syncthreads();
    float mx = -INFINITY;
    for (int t = tid; t < T; t += MEGA_NT) mx = fmaxf(mx, sc[t]);
    mx = bg_block_max_f32(mx, sm);
    double sum = 0.0;
    for (int t = tid; t < T; t += MEGA_NT) {
        const float v = bg_h2f(exp_tab[bg_f2h(__fsub_rn(sc[t], mx))]);
        sc[t] = v;
        sum += (double) v;
    }
    sum = bg_block_sum_f64(sum, sd);
    const float inv = (float) (1.0 / sum);
    for (int t = tid; t < T; t += MEGA_NT) sc[t] = __fmul_rn(sc[t], inv);
    __syncthreads();
    // ---- V: unit (r, col), r = t % 32
    const int np = T & ~31;
    for (int i = tid; i < (T - np) * ncol; i += MEGA_NT) tailv[i] = __ldcg(Vb + (size_t) (np + i / ncol) * ldkv + c0 + (i % ncol));
    if ((ncol & 3) == 0 && (Tmax & 3) == 0 && (c0 & 3) == 0 && (ldkv & 3) == 0) {
        // unit (r, 4 columns): one float4 load per step, 16 steps in flight -> one round trip per 512 positions
        const int ncg = ncol >> 2;
        for (int u = tid; u < 32 * ncg; u += MEGA_NT) {
            const int cg = u % ncg, r = u / ncg;
            const float * vp = Vb + (size_t) r * ldkv + c0 + 4 * cg;
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int s0 = 0; s0 < np; s0 += 16 * 32) {
                float4 v[16];
#pragma unroll
                for (int k = 0; k < 16; k++)
                    v[k] = (s0 + 32 * k < np) ? __ldcg((const float4 *) (vp + (size_t) (s0 + 32 * k) * ldkv)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int k = 0; k < 16; k++) if (s0 + 32 * k < np) {
                    const float pw = sc[s0 + 32 * k + r];
                    acc.x = fmaf(v[k].x, pw, acc.x); acc.y = fmaf(v[k].y, pw, acc.y);
                    acc.z = fmaf(v[k].z, pw, acc.z); acc.w = fmaf(v[k].w, pw, acc.w);
                }
            }
            *(float4 *) (red + r * ncol + 4 * cg) = acc;
        }
    } else {
        for (int u = tid; u < 32 * ncol; u += MEGA_NT) {
            const int col = u % ncol, r = u / ncol;
            const float * vp = Vb + (size_t) r * ldkv + c0 + col;
            float acc = 0.0f;
            for (int s0 = 0; s0 < np; s0 += 32) acc = fmaf(__ldcg(vp + (size_t) s0 * ldkv), sc[s0 + r], acc);
            red[r * ncol + col] = acc;
        }
    }
    __syncthreads();
    if (tid < ncol) {
        float x0[8];
#pragma unroll
        for (int l = 0; l < 8; l++) {
            const float a02 = __fadd_rn(red[(0 * 8 + l) * ncol + tid], red[(2 * 8 + l) * ncol + tid]);
            const float a13 = __fadd_rn(red[(1 * 8 + l) * ncol + tid], red[(3 * 8 + l) * ncol + tid]);
            x0[l] = __fadd_rn(a02, a13);
        }
        const float t0 = __fadd_rn(x0[0], x0[4]), t1 = __fadd_rn(x0[1], x0[5]);
        const float t2 = __fadd_rn(x0[2], x0[6]), t3 = __fadd_rn(x0[3], x0[7]);
        float sumf = __fadd_rn(__fadd_rn(t0, t1), __fadd_rn(t2, t3));
        const int nv = np + ((T - np) & ~3);
        int t = np;
        for (; t < nv; t++) sumf = __fadd_rn(sumf, __fmul_rn(tailv[(t - np) * ncol + tid], sc[t]));
        for (; t < T;  t++) sumf = fmaf(tailv[(t - np) * ncol + tid], sc[t], sumf);
        s_out[tid] = sumf;
    }
    __syncthreads();
}

#define PROF(ph, k) do { if (p.prof && blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) p.prof[(l * 5 + (ph)) * 6 + (k)] = clock64(); } while (0)

// L2-prefetch everything layer `Ln` will read in this CTA's slices (weights for F16 / safety,
// and the small f32 vectors: LayerNorm weights, biases)
struct MegaRows { int qkv0, qkv1, o0, o1, f10, f11, v0, v1; };
__device__ __forceinline__ void mega_prefetch_layer(const MegaParams & p, const MegaLayer & Ln, const MegaRows & R, bool weights) {
    const int d = p.d, ff = p.ff;
    if (weights) {
        for (int mat = 0; mat < 3; mat++) {
            const int a = max(R.qkv0, mat * d) - mat * d, b = min(R.qkv1, (mat + 1) * d) - mat * d;
            if (b > a) mega_prefetch_l2((mat == 0 ? Ln.q_w : mat == 1 ? Ln.k_w : Ln.v_w) + (size_t) a * p.stride_d, (size_t) (b - a) * p.stride_d);
        }
        mega_prefetch_l2(Ln.o_w + (size_t) R.o0 * p.stride_d, (size_t) (R.o1 - R.o0) * p.stride_d);
        mega_prefetch_l2(Ln.fc2_w + (size_t) R.o0 * p.stride_f, (size_t) (R.o1 - R.o0) * p.stride_f);
        mega_prefetch_l2(Ln.fc1_w + (size_t) R.f10 * p.stride_d, (size_t) (R.f11 - R.f10) * p.stride_d);
    }
    // 53 KB of f32 vectors per layer, each asked for by one CTA
    const unsigned db = (unsigned) d * 4u, fb = (unsigned) ff * 4u;
    switch ((gridDim.x - 1 - blockIdx.x)) {
        case 0: mega_prefetch_lines(Ln.ln0_w, db); mega_prefetch_lines(Ln.ln0_b, db); break;
        case 1: mega_prefetch_lines(Ln.ln1_w, db); mega_prefetch_lines(Ln.ln1_b, db); break;
        case 2: mega_prefetch_lines(Ln.q_b, db); mega_prefetch_lines(Ln.k_b, db); mega_prefetch_lines(Ln.v_b, db); break;
        case 3: mega_prefetch_lines(Ln.o_b, db); mega_prefetch_lines(Ln.fc2_b, db); break;
        case 4: mega_prefetch_lines(Ln.fc1_b, fb); break;
        default: break;
    }
}

template <int FMT, int DK>
__global__ void __launch_bounds__(MEGA_NT, 1) k_mega(MegaParams p) {
    extern __shared__ __align__(16) uint8_t smem[];
    __shared__ double sd[32];
    __shared__ float smx[32];
    __shared__ __align__(16) float s_blk[128];
    __shared__ int s_tok;
    constexpr bool ISF16 = (FMT == BG_F16);
    float * s_row = (float *) (smem + p.sm_row);
    uint8_t * s_act = smem + p.sm_act;
    float * s_p = (float *) (smem + p.sm_p);
    float * s_s = (float *) (smem + p.sm_s);
    float * s_m = (float *) (smem + p.sm_m);
    float * s_attn = (float *) (smem + p.sm_attn);
    const int tid = threadIdx.x, d = p.d, ff = p.ff;
    unsigned long long epoch = p.epoch0;

    // rows of this CTA in every phase (static for the whole kernel)
    MegaRows R;
    mega_my_rows(3 * d, 8, R.qkv0, R.qkv1);
    mega_my_rows(d, 8, R.o0, R.o1);
    mega_my_rows(ff, 32, R.f10, R.f11);
    mega_my_rows(p.n_vocab, 8, R.v0, R.v1);
    const int qkv0 = R.qkv0, qkv1 = R.qkv1, o0 = R.o0, o1 = R.o1, f10 = R.f10, f11 = R.f11, v0 = R.v0, v1 = R.v1;

    const bool ln_fast = d <= 4 * MEGA_NT;     // LayerNorm parameters fit 4 registers per thread
    float lw[4], lb[4];
    auto load_ln = [&](const float * w, const float * b) {
#pragma unroll
        for (int i = 0; i < 4; i++) { const int c = tid + i * MEGA_NT; lw[i] = (ln_fast && c < d) ? w[c] : 0.f; lb[i] = (ln_fast && c < d) ? b[c] : 0.f; }
    };
    WRegs<FMT> wr;     // weights of the NEXT matmul phase, loaded one phase ahead
    {   // first phase of the first layer
        const MegaLayer L0 = p.layers[0];
        if (!ISF16) { const MMDesc D = mega_mm_desc(p, L0.q_w, L0.k_w, L0.v_w, d, qkv0, qkv1, false); mega_load_units<FMT>(wr, D, D.r0, min(D.RT, D.r1 - D.r0)); }
        mega_prefetch_layer(p, L0, R, ISF16);
    }

    // ---- input token: given, or argmax over the candidates the previous launch left
    if (tid < 32) {
        int tok;
        if (p.use_cand == 1) {
            float best = -INFINITY; int bi = 0x7fffffff;
            for (int i = tid; i < p.n_cand; i += 32) {
                const float v = __ldcg(p.cand_val + i); const int ix = __ldcg(p.cand_idx + i);
                if (v > best || (v == best && ix < bi)) { best = v; bi = ix; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ov = __shfl_xor_sync(FULLMASK, best, o); const int oi = __shfl_xor_sync(FULLMASK, bi, o);
                if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
            }
            tok = bi == 0x7fffffff ? 0 : bi;
        } else tok = p.use_cand == 2 ? p.tok_imm : __ldcg(p.tok);
        if (tid == 0) {
            s_tok = tok;
            if (blockIdx.x == 0 && p.log_slot >= 0) p.idlog[p.log_slot] = tok;
        }
    }
    __syncthreads();
    // ---- embedding (every CTA, into shared memory; CTA 0 publishes x for the residual adds)
    {
        int tok = s_tok; tok = tok < 0 ? 0 : (tok >= p.n_vocab ? p.n_vocab - 1 : tok);
        int prow = p.n_past + 2; prow = prow >= p.n_pos_rows ? p.n_pos_rows - 1 : prow;
        const size_t rb = bg_file_row_bytes(p.wtype, d);
        const uint8_t * tr = p.embed_tok + rb * (size_t) tok;
        const uint8_t * pr = p.embed_pos + rb * (size_t) prow;
        for (int c = tid; c < d; c += MEGA_NT) {
            const float a = __fmul_rn(bg_dequant_elem(p.wtype, tr, c), p.emb_scale);
            const float v = __fadd_rn(a, bg_dequant_elem(p.wtype, pr, c));
            s_row[c] = v;
            if (blockIdx.x == 0) p.x[c] = v;
        }
        __syncthreads();
    }

    for (int l = 0; l < p.n_layer; l++) {
        const MegaLayer L = p.layers[l];
        float * kc = p.kcache + (size_t) l * p.n_positions * d;
        float * vc = p.vcache + (size_t) l * p.n_positions * d;
        const int pos = p.n_past, T = p.n_past + 1;
        // ================= P1: LN0 + q,k,v =================
        PROF(0, 0);
        if (l + 1 < p.n_layer) mega_prefetch_layer(p, p.layers[l + 1], R, ISF16);
        else if (ISF16) mega_prefetch_l2(p.lm_head + (size_t) v0 * p.stride_d, (size_t) (v1 - v0) * p.stride_d);
        {   // K/V rows this CTA will read in P2: start them towards L2 now (128-byte lines of its head slice)
            constexpr int LPR = (DK * 4 + 127) / 128;          // lines per row slice
            for (int w = blockIdx.x; w < p.n_head * p.attn_parts; w += gridDim.x) {
                const int h = w / p.attn_parts;
                for (int t = tid; t < T - 1; t += MEGA_NT) {
                    const float * ks = kc + (size_t) t * d + (size_t) h * DK, * vs = vc + (size_t) t * d + (size_t) h * DK;
#pragma unroll
                    for (int ln = 0; ln < LPR; ln++) {
                        asm volatile("prefetch.global.L2 [%0];" :: "l"(ks + ln * 32));
                        asm volatile("prefetch.global.L2 [%0];" :: "l"(vs + ln * 32));
                    }
                }
            }
        }
        load_ln(L.ln0_w, L.ln0_b);
        PROF(0, 5);
        if (l > 0) { for (int c = tid; c < d; c += MEGA_NT) s_row[c] = __ldcg(p.x + c); __syncthreads(); }
        PROF(0, 3);
        if (ln_fast) bg_ln_row_pre<MEGA_NT, 4>(s_row, d, lw, lb, p.eps, sd); else bg_ln_row(s_row, d, L.ln0_w, L.ln0_b, p.eps, sd);
        PROF(0, 4);
        bg_row_to_record(s_row, d, p.wtype, s_act, p.actb_d, p.offn_d, p.offdd_d, p.offs_d, p.code_off);
        __syncthreads();
        PROF(0, 1);
        {
            const MMDesc D = mega_mm_desc(p, L.q_w, L.k_w, L.v_w, d, qkv0, qkv1, false);
            auto pre = [&](int r) { const int mat = (r >= d) + (r >= 2 * d); return make_float2((mat == 0 ? L.q_b : mat == 1 ? L.k_b : L.v_b)[r - mat * d], 0.f); };
            auto fin = [&](int r, float v, float2 pv) {
                const int mat = (r >= d) + (r >= 2 * d), rr = r - mat * d;
                const float t = __fadd_rn(pv.x, v);
                if (mat == 0) p.q[rr] = __fmul_rn(t, p.qscale);
                else (mat == 1 ? kc : vc)[(size_t) pos * d + rr] = t;
            };
            if (ISF16) mega_mm_run_f16(D, s_act, pre, fin);
            else mega_mm_run_q<FMT>(D, wr, s_act, s_p, s_s, s_m, pre, fin);
        }
        const MMDesc Do = mega_mm_desc(p, L.o_w, L.o_w, L.o_w, d, o0, o1, false);
        PROF(0, 2);
        epoch += gridDim.x; mega_barrier_arrive(p.flags); mega_barrier_wait(p.flags, epoch);
        // ================= P2: attention =================
        PROF(1, 0);
        for (int w = blockIdx.x; w < p.n_head * p.attn_parts; w += gridDim.x) {
            const int h = w / p.attn_parts, part = w - h * p.attn_parts;
            const int ncol = DK / p.attn_parts, c0 = part * ncol;
            float * s_q = s_row;                       // d floats free to reuse here
            for (int c = tid; c < DK; c += MEGA_NT) s_q[c] = __ldcg(p.q + h * DK + c);
            __syncthreads();
            mega_attention_cols<DK>(s_q, kc + (size_t) h * DK, vc + (size_t) h * DK, d, T, p.exp_tab,
                                    s_attn, p.n_positions, sd, smx, c0, ncol, s_blk);
            if (p.att_prequant) {                      // ncol == 32: this CTA owns one activation block of out_proj's input
                if (tid < 32) {
                    const int c = h * DK + c0 + tid;
                    if (ISF16) mega_f16_record_elem(s_blk[tid], c, p.rec_att);
                    else mega_quant_block(s_blk, c >> 5, p.wtype, p.rec_att, p.offn_d, p.offdd_d, p.offs_d, p.code_off);
                }
            } else if (tid < ncol) p.att[h * DK + c0 + tid] = s_blk[tid];
            __syncthreads();
        }
        PROF(1, 2);
        epoch += gridDim.x; mega_barrier_arrive(p.flags);
        if (!ISF16) mega_load_units<FMT>(wr, Do, Do.r0, min(Do.RT, Do.r1 - Do.r0));   // P3's weights travel while we wait
        mega_barrier_wait(p.flags, epoch);
        // ================= P3: out_proj + bias + residual =================
        PROF(2, 0);
        if (p.att_prequant) {
            for (int i = tid; i < p.actb_d / 16; i += MEGA_NT) ((uint4 *) s_act)[i] = __ldcg((const uint4 *) p.rec_att + i);
        } else {
            for (int c = tid; c < d; c += MEGA_NT) s_row[c] = __ldcg(p.att + c);
            __syncthreads();
            bg_row_to_record(s_row, d, p.wtype, s_act, p.actb_d, p.offn_d, p.offdd_d, p.offs_d, p.code_off);
        }
        __syncthreads();
        PROF(2, 1);
        {
            auto pre = [&](int r) { return make_float2(L.o_b[r], __ldcg(p.x + r)); };
            auto fin = [&](int r, float v, float2 pv) { p.x1[r] = __fadd_rn(__fadd_rn(v, pv.x), pv.y); };
            if (ISF16) mega_mm_run_f16(Do, s_act, pre, fin);
            else mega_mm_run_q<FMT>(Do, wr, s_act, s_p, s_s, s_m, pre, fin);
        }
        const MMDesc D1 = mega_mm_desc(p, L.fc1_w, L.fc1_w, L.fc1_w, ff, f10, f11, false);
        PROF(2, 2);
        epoch += gridDim.x; mega_barrier_arrive(p.flags);
        if (!ISF16) mega_load_units<FMT>(wr, D1, D1.r0, min(D1.RT, D1.r1 - D1.r0));
        mega_barrier_wait(p.flags, epoch);
        // ================= P4: LN1 + fc1 + bias + GELU (+ quantise own 32-row blocks) =================
        PROF(3, 0);
        load_ln(L.ln1_w, L.ln1_b);
        for (int c = tid; c < d; c += MEGA_NT) s_row[c] = __ldcg(p.x1 + c);
        __syncthreads();
        PROF(3, 3);
        if (ln_fast) bg_ln_row_pre<MEGA_NT, 4>(s_row, d, lw, lb, p.eps, sd); else bg_ln_row(s_row, d, L.ln1_w, L.ln1_b, p.eps, sd);
        PROF(3, 4);
        bg_row_to_record(s_row, d, p.wtype, s_act, p.actb_d, p.offn_d, p.offdd_d, p.offs_d, p.code_off);
        __syncthreads();
        PROF(3, 1);
        {
            float * s_h = s_row;                       // LN row no longer needed: GELU outputs of this CTA's rows
            auto pre = [&](int r) { return make_float2(L.fc1_b[r], 0.f); };
            auto fin = [&](int r, float v, float2 pv) { s_h[r - f10] = bg_h2f(p.gelu[bg_f2h(__fadd_rn(pv.x, v))]); };
            if (ISF16) mega_mm_run_f16(D1, s_act, pre, fin);
            else mega_mm_run_q<FMT>(D1, wr, s_act, s_p, s_s, s_m, pre, fin);
            // f11 - f10 is a multiple of 32: one warp per finished block
            for (int b = tid >> 5; b < (f11 - f10) >> 5; b += MEGA_NT / 32) {
                const int c = f10 + b * 32 + (tid & 31);
                if (ISF16) mega_f16_record_elem(s_h[b * 32 + (tid & 31)], c, p.rec_hff);
                else mega_quant_block(s_h + b * 32, c >> 5, p.wtype, p.rec_hff, p.offn_f, p.offdd_f, p.offs_f, p.code_off);
            }
        }
        const MMDesc D2 = mega_mm_desc(p, L.fc2_w, L.fc2_w, L.fc2_w, d, o0, o1, true);
        PROF(3, 2);
        epoch += gridDim.x; mega_barrier_arrive(p.flags);
        if (!ISF16) mega_load_units<FMT>(wr, D2, D2.r0, min(D2.RT, D2.r1 - D2.r0));
        mega_barrier_wait(p.flags, epoch);
        // ================= P5: fc2 + bias + residual =================
        PROF(4, 0);
        for (int i = tid; i < p.actb_f / 16; i += MEGA_NT) ((uint4 *) s_act)[i] = __ldcg((const uint4 *) p.rec_hff + i);
        __syncthreads();
        PROF(4, 1);
        {
            auto pre = [&](int r) { return make_float2(L.fc2_b[r], __ldcg(p.x1 + r)); };
            auto fin = [&](int r, float v, float2 pv) { p.x[r] = __fadd_rn(__fadd_rn(pv.x, v), pv.y); };
            if (ISF16) mega_mm_run_f16(D2, s_act, pre, fin);
            else mega_mm_run_q<FMT>(D2, wr, s_act, s_p, s_s, s_m, pre, fin);
        }
        PROF(4, 2);
        epoch += gridDim.x; mega_barrier_arrive(p.flags);
        if (!ISF16) {    // next matmul phase: P1 of the next layer, or the first lm_head tile
            if (l + 1 < p.n_layer) {
                const MegaLayer Ln = p.layers[l + 1];
                const MMDesc Dn = mega_mm_desc(p, Ln.q_w, Ln.k_w, Ln.v_w, d, qkv0, qkv1, false);
                mega_load_units<FMT>(wr, Dn, Dn.r0, min(Dn.RT, Dn.r1 - Dn.r0));
            } else {
                const MMDesc Dn = mega_mm_desc(p, p.lm_head, p.lm_head, p.lm_head, p.n_vocab, v0, v1, false);
                mega_load_units<FMT>(wr, Dn, Dn.r0, min(Dn.RT, Dn.r1 - Dn.r0));
            }
        }
        mega_barrier_wait(p.flags, epoch);
    }
    // ================= final LayerNorm + lm_head =================
    { const int l = p.n_layer; PROF(0, 0); }
    load_ln(p.lnf_w, p.lnf_b);
    for (int c = tid; c < d; c += MEGA_NT) s_row[c] = __ldcg(p.x + c);
    __syncthreads();
    if (ln_fast) bg_ln_row_pre<MEGA_NT, 4>(s_row, d, lw, lb, p.eps, sd); else bg_ln_row(s_row, d, p.lnf_w, p.lnf_b, p.eps, sd);
    bg_row_to_record(s_row, d, p.wtype, s_act, p.actb_d, p.offn_d, p.offdd_d, p.offs_d, p.code_off);
    __syncthreads();
    { const int l = p.n_layer; PROF(0, 1); }
    {
        const MMDesc D = mega_mm_desc(p, p.lm_head, p.lm_head, p.lm_head, p.n_vocab, v0, v1, false);
        float best = -INFINITY; int bi = 0x7fffffff;
        auto pre = [&](int) { return make_float2(0.f, 0.f); };
        auto fin = [&](int r, float v, float2) {
            p.logits[r] = v;
            if (v > best || (v == best && r < bi)) { best = v; bi = r; }
        };
        if (ISF16) mega_mm_run_f16(D, s_act, pre, fin);
        else mega_mm_run_q<FMT>(D, wr, s_act, s_p, s_s, s_m, pre, fin);
        // per-CTA argmax candidate (first index wins ties) for the next launch's prologue
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(FULLMASK, best, o); const int oi = __shfl_xor_sync(FULLMASK, bi, o);
            if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
        }
        int * si = (int *) (smx + 16);      // smx: 32 floats; [0,16) values, [16,32) indices
        __syncthreads();
        if ((tid & 31) == 0) { smx[tid >> 5] = best; si[tid >> 5] = bi; }
        __syncthreads();
        if (tid == 0) {
            for (int i = 1; i < MEGA_NT / 32; i++) if (smx[i] > best || (smx[i] == best && si[i] < bi)) { best = smx[i]; bi = si[i]; }
            p.cand_val[blockIdx.x] = best; p.cand_idx[blockIdx.x] = bi;
        }
    }
    { const int l = p.n_layer; PROF(0, 2); }
}

// reduces the candidates of the last launch of a greedy loop into the id log
static __global__ void k_mega_pick(const float * __restrict__ cand_val, const int * __restrict__ cand_idx, int n_cand,
                            int * __restrict__ idlog, int slot, int * __restrict__ next_tok) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        float best = -INFINITY; int bi = 0x7fffffff;
        for (int i = 0; i < n_cand; i++) { const float v = cand_val[i]; const int ix = cand_idx[i]; if (v > best || (v == best && ix < bi)) { best = v; bi = ix; } }
        if (bi == 0x7fffffff) bi = 0;
        idlog[slot] = bi;
        if (next_tok) next_tok[0] = bi;
    }
}
