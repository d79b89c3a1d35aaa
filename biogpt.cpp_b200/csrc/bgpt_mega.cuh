// bgpt_mega.cuh -- one persistent kernel per decoded token.
//
// The N=1 decode step of BioGPT is a chain of ~120 dependent small operators:
// as separate launches each costs ~10 us of pure latency (profiles/r1_v0*), 50x more than the
// 30 us the weights need to stream from HBM.  k_mega runs the whole step in ONE launch: a
// resident grid of CTAs (one per SM), every operator a *phase* executed by all CTAs on their
// slice of the rows, phases separated by a grid-wide barrier (an atomic counter in L2).
//
//   phase per layer                                 rows split over CTAs        barrier after
//   LN0 + q,k,v projections (+bias, q scale, KV append)   3*d_model                  yes
//   attention (one head per CTA)                          n_head                     yes
//   out_proj + bias + residual                            d_model                    yes
//   LN1 + fc1 + bias + GELU                               d_ff                       yes
//   fc2 + bias + residual                                 d_model                    yes
//   final: LN + lm_head (+ per-CTA argmax candidates)     n_vocab                    -
//
// Every CTA recomputes the cheap per-token prologue of its phase (LayerNorm, activation
// quantisation: <= 16 KB read from L2) instead of waiting for another CTA to publish it.
//
// The dot products keep the lane order of bgpt_kernels.cuh but are organised for latency, not
// for one-thread-per-sum: within a tile of <= 32 rows,
//   phase A: thread (row, g, j) loads one uint4 = the 4-byte groups of sums j and j+4 for the
//            blocks 4g..4g+3, does the exact integer work (dp4a) and writes the 8 products
//            p = (float) isum and the 4 scales s = d_w*d_a to shared memory -- fully parallel;
//   phase B: thread (row, l) walks its running sum acc = fma(s_b, p_b, acc) over the blocks in
//            order from shared memory -- the only sequential part, 1 FMA per block.
// F16 weights need no integer phase: thread (row, lane) is the running sum lane of its row.
#pragma once
#include "bgpt_kernels.cuh"

#define MEGA_NT 512          // threads per CTA
#define MEGA_RT 32           // rows per tile

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
    unsigned long long * bar; unsigned long long bar_base;
    const int * tok;                                    // device: input token id (use_cand == 0)
    int use_cand; float * cand_val; int * cand_idx; int n_cand;   // argmax candidates of the previous launch
    int * idlog; int log_slot;                          // idlog[log_slot] = input token when log_slot >= 0
    int n_past;
    int attn_parts;                                     // CTAs per head in the attention phase (column split)
    long long * prof;                                   // optional: clock64 stamps of CTA 0, [n_layer][5][3] (+3 for lm_head)
    // shared-memory carve-up (bytes, computed on the host)
    int sm_row, sm_act, sm_p, sm_s, sm_m, sm_attn, sm_total;
};

// ---- grid barrier: all CTAs of the (co-resident) grid --------------------------------------
__device__ __forceinline__ unsigned long long mega_ld_acquire(const unsigned long long * p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// Arrivals are counted on bar[0]; the last arriver publishes the generation on bar[16] (a
// different 128-byte line), which is the only word the waiters poll -- the atomics and the
// polling loads never fight over the same L2 line.
__device__ __forceinline__ void mega_grid_barrier(unsigned long long * bar, unsigned long long target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned long long old = atomicAdd(bar, 1ULL);
        if (old + 1 == target) {
            asm volatile("st.release.gpu.global.u64 [%0], %1;" :: "l"(bar + 16), "l"(target) : "memory");
        } else {
            while (mega_ld_acquire(bar + 16) < target) { }
        }
        __threadfence();
    }
    __syncthreads();
}

// L2 prefetch of a contiguous byte range (weights of an upcoming phase): the HBM read is issued
// now, the later ld.global finds the lines in L2
__device__ __forceinline__ void mega_prefetch_l2(const uint8_t * ptr, size_t bytes) {
    // chunks of 16 KB spread over the threads of the CTA
    const size_t CH = 16384;
    const size_t n = (bytes + CH - 1) / CH;
    for (size_t i = threadIdx.x; i < n; i += MEGA_NT) {
        const size_t off = i * CH;
        const unsigned sz = (unsigned) ((bytes - off) < CH ? (bytes - off) : CH) & ~15u;
        if (sz) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(ptr + off), "r"(sz) : "memory");
    }
}

// rows [r0, r1) of an M-row phase owned by this CTA: contiguous chunks of 8-row tiles
__device__ __forceinline__ void mega_my_rows(int M, int & r0, int & r1) {
    const int nt = (M + 7) >> 3;
    const int t0 = (int) (((long long) blockIdx.x * nt) / gridDim.x);
    const int t1 = (int) (((long long) (blockIdx.x + 1) * nt) / gridDim.x);
    r0 = t0 * 8; r1 = t1 * 8 < M ? t1 * 8 : M;
}

// ---- tile of RT rows, block-quantised formats ---------------------------------------------
// s_rowptr[r] = start of the r-th row (device layout).  epi(tile_row, value) by one thread per row.
template <int FMT, class EpiF>
__device__ __forceinline__ void mega_tile_q(const uint8_t * const * s_rowptr, int RT, int G, int off_qh, int off_d, int off_m,
                                            const uint8_t * s_act, int off_n, int off_dd, int off_s,
                                            float * s_p, float * s_s, float * s_m, EpiF epi) {
    constexpr bool IS8   = (FMT == BG_Q8_0);
    constexpr bool HASQH = (FMT == BG_Q5_0 || FMT == BG_Q5_1);
    constexpr bool HASM  = (FMT == BG_Q4_1 || FMT == BG_Q5_1);
    constexpr bool HASOFF = (FMT == BG_Q4_0 || FMT == BG_Q5_0);
    const int nbp = G * 4, PS = nbp + 4;
    // ---- phase A
    const bool gpow2 = (G & (G - 1)) == 0;
    const int gsh = 31 - __clz(G);
    for (int u = threadIdx.x; u < RT * G * 4; u += MEGA_NT) {
        const int j = u & 3;
        const int g = gpow2 ? ((u >> 2) & (G - 1)) : ((u >> 2) % G);
        const int row = gpow2 ? (u >> (2 + gsh)) : (u / (4 * G));
        const uint8_t * wrow = s_rowptr[row];
        uint32_t lo[4], hi[4];
        if (IS8) {
            const uint4 w0 = ldg_stream128(wrow + (size_t) ((g * 2 + 0) * 4 + j) * 16);
            const uint4 w1 = ldg_stream128(wrow + (size_t) ((g * 2 + 1) * 4 + j) * 16);
            lo[0] = w0.x; lo[1] = w0.y; lo[2] = w0.z; lo[3] = w0.w;
            hi[0] = w1.x; hi[1] = w1.y; hi[2] = w1.z; hi[3] = w1.w;
        } else {
            const uint4 w = ldg_stream128(wrow + (size_t) (g * 4 + j) * 16);
            const uint32_t ww[4] = { w.x, w.y, w.z, w.w };
            uint32_t qh = 0;
            if (HASQH) qh = ldg_stream32(wrow + off_qh + g * 16 + j * 4);
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
        const uint4 a0 = *(const uint4 *) (s_act + (g * 8 + j) * 16);
        const uint4 a1 = *(const uint4 *) (s_act + (g * 8 + j + 4) * 16);
        int4 n0 = make_int4(0, 0, 0, 0), n1 = make_int4(0, 0, 0, 0);
        if (HASOFF) {
            n0 = *(const int4 *) (s_act + off_n + (g * 8 + j) * 16);
            n1 = *(const int4 *) (s_act + off_n + (g * 8 + j + 4) * 16);
        }
        float4 P0, P1;
        P0.x = (float) __dp4a((int) lo[0], (int) a0.x, n0.x); P1.x = (float) __dp4a((int) hi[0], (int) a1.x, n1.x);
        P0.y = (float) __dp4a((int) lo[1], (int) a0.y, n0.y); P1.y = (float) __dp4a((int) hi[1], (int) a1.y, n1.y);
        P0.z = (float) __dp4a((int) lo[2], (int) a0.z, n0.z); P1.z = (float) __dp4a((int) hi[2], (int) a1.z, n1.z);
        P0.w = (float) __dp4a((int) lo[3], (int) a0.w, n0.w); P1.w = (float) __dp4a((int) hi[3], (int) a1.w, n1.w);
        *(float4 *) (s_p + (size_t) (row * 8 + j) * PS + 4 * g) = P0;
        *(float4 *) (s_p + (size_t) (row * 8 + j + 4) * PS + 4 * g) = P1;
        if (j == 0) {
            const uint2 dh = ldg_stream64(wrow + off_d + g * 8);
            const float4 da = *(const float4 *) (s_act + off_dd + g * 16);
            float4 S;
            S.x = __fmul_rn(bg_h2f((uint16_t) (dh.x & 0xFFFF)), da.x); S.y = __fmul_rn(bg_h2f((uint16_t) (dh.x >> 16)), da.y);
            S.z = __fmul_rn(bg_h2f((uint16_t) (dh.y & 0xFFFF)), da.z); S.w = __fmul_rn(bg_h2f((uint16_t) (dh.y >> 16)), da.w);
            *(float4 *) (s_s + (size_t) row * nbp + 4 * g) = S;
        }
        if (HASM && j == 1) {
            const uint2 mh = ldg_stream64(wrow + off_m + g * 8);
            float4 Mv;
            Mv.x = bg_h2f((uint16_t) (mh.x & 0xFFFF)); Mv.y = bg_h2f((uint16_t) (mh.x >> 16));
            Mv.z = bg_h2f((uint16_t) (mh.y & 0xFFFF)); Mv.w = bg_h2f((uint16_t) (mh.y >> 16));
            *(float4 *) (s_m + (size_t) row * nbp + 4 * g) = Mv;
        }
    }
    __syncthreads();
    // ---- phase B: 8 running sums per row, whole warps participate (shuffles)
    for (int c0 = (threadIdx.x >> 5) << 5; c0 < RT * 8; c0 += MEGA_NT) {
        const int c = c0 + (threadIdx.x & 31);
        const bool valid = c < RT * 8;
        const int cc = valid ? c : RT * 8 - 1;
        const int row = cc >> 3, l = cc & 7;
        const float * pp = s_p + (size_t) cc * PS;
        const float * ss = s_s + (size_t) row * nbp;
        float acc = 0.0f, summ = 0.0f;
        for (int g = 0; g < G; g++) {
            const float4 pv = *(const float4 *) (pp + 4 * g);
            const float4 sv = *(const float4 *) (ss + 4 * g);
            acc = fmaf(sv.x, pv.x, acc); acc = fmaf(sv.y, pv.y, acc);
            acc = fmaf(sv.z, pv.z, acc); acc = fmaf(sv.w, pv.w, acc);
            if (HASM && l == 0) {
                const float4 mv = *(const float4 *) (s_m + (size_t) row * nbp + 4 * g);
                const float4 sa = *(const float4 *) (s_act + off_s + g * 16);
                summ = fmaf(mv.x, sa.x, summ); summ = fmaf(mv.y, sa.y, summ);
                summ = fmaf(mv.z, sa.z, summ); summ = fmaf(mv.w, sa.w, summ);
            }
        }
        // hsum_float_8: (a[l+4] + a[l]) then +2, +1
        float r = __fadd_rn(acc, __shfl_xor_sync(FULLMASK, acc, 4));
        r = __fadd_rn(r, __shfl_xor_sync(FULLMASK, r, 2));
        r = __fadd_rn(r, __shfl_xor_sync(FULLMASK, r, 1));
        if (HASM) r = __fadd_rn(r, summ);
        if (valid && l == 0) epi(row, r);
    }
    __syncthreads();
}

// ---- tile of RT rows, F16 weights: thread (row, lane) is running-sum lane `lane` of its row
template <class EpiF>
__device__ __forceinline__ void mega_tile_f16(const uint8_t * const * s_rowptr, int RT, int G, const uint8_t * s_act, EpiF epi) {
    const int lane = threadIdx.x & 31;
    for (int row = threadIdx.x >> 5; row < RT; row += MEGA_NT / 32) {
        const uint8_t * wrow = s_rowptr[row];
        float c = 0.0f;
#pragma unroll 4
        for (int g = 0; g < G; g++) {
            const uint4 w = ldg_stream128(wrow + (size_t) (g * 32 + lane) * 16);
            const float4 x0 = *(const float4 *) (s_act + (size_t) ((g * 2 + 0) * 32 + lane) * 16);
            const float4 x1 = *(const float4 *) (s_act + (size_t) ((g * 2 + 1) * 32 + lane) * 16);
            c = fmaf(bg_h2f((uint16_t) (w.x & 0xFFFF)), x0.x, c); c = fmaf(bg_h2f((uint16_t) (w.x >> 16)), x0.y, c);
            c = fmaf(bg_h2f((uint16_t) (w.y & 0xFFFF)), x0.z, c); c = fmaf(bg_h2f((uint16_t) (w.y >> 16)), x0.w, c);
            c = fmaf(bg_h2f((uint16_t) (w.z & 0xFFFF)), x1.x, c); c = fmaf(bg_h2f((uint16_t) (w.z >> 16)), x1.y, c);
            c = fmaf(bg_h2f((uint16_t) (w.w & 0xFFFF)), x1.z, c); c = fmaf(bg_h2f((uint16_t) (w.w >> 16)), x1.w, c);
        }
        float r = c;
        r = __fadd_rn(r, __shfl_xor_sync(FULLMASK, r, 16));
        r = __fadd_rn(r, __shfl_xor_sync(FULLMASK, r, 8));
        r = __fadd_rn(r, __shfl_xor_sync(FULLMASK, r, 4));
        r = __fadd_rn(r, __shfl_xor_sync(FULLMASK, r, 1));
        r = __fadd_rn(r, __shfl_xor_sync(FULLMASK, r, 2));
        if (lane == 0) epi(row, r);
    }
    __syncthreads();
}

// ---- attention for `ncol` output columns [c0, c0+ncol) of one head ----------------------------
// Scores and softmax are computed for the whole head by every CTA that shares it (K comes from
// L2 after the first reader); the V reduction -- the 32 running sums per column -- is split by
// column, so no CTA ever needs another CTA's partial result.
template <int DK>
__device__ __forceinline__ void mega_attention_cols(const float * s_q, const float * Kb, const float * Vb, int ldkv, int T,
                                                    const uint16_t * __restrict__ exp_tab, float * s_f, int Tmax,
                                                    double * sd, float * sm, int c0, int ncol, float * out) {
    float * sc  = s_f;                 // [Tmax]
    float * red = s_f + Tmax;          // [32][ncol]
    float * tailv = red + 32 * DK;     // [<=31][ncol]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr int NWARP = MEGA_NT / 32;
    constexpr int NP = DK & ~31;
    constexpr int NV = NP + ((DK - NP) & ~3);
    constexpr int NQ = NP > 0 ? NP / 32 : 1;
    constexpr int U = 8;               // rows in flight per warp
    float qreg[NQ];
#pragma unroll
    for (int i = 0; i < NP / 32; i++) qreg[i] = s_q[i * 32 + lane];
    for (int t0 = warp * U; t0 < T; t0 += NWARP * U) {
        float kv[U][NQ];
#pragma unroll
        for (int u = 0; u < U; u++)
#pragma unroll
            for (int i = 0; i < NP / 32; i++)
                kv[u][i] = (t0 + u < T) ? __ldcg(Kb + (size_t) (t0 + u) * ldkv + i * 32 + lane) : 0.0f;
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int t = t0 + u;
            if (t >= T) break;                      // warp-uniform
            float s = 0.0f;
#pragma unroll
            for (int i = 0; i < NP / 32; i++) s = fmaf(kv[u][i], qreg[i], s);
            if (NP > 0) {
                s = __fadd_rn(s, __shfl_xor_sync(FULLMASK, s, 16));
                s = __fadd_rn(s, __shfl_xor_sync(FULLMASK, s, 8));
                s = __fadd_rn(s, __shfl_xor_sync(FULLMASK, s, 4));
                s = __fadd_rn(s, __shfl_xor_sync(FULLMASK, s, 1));
                s = __fadd_rn(s, __shfl_xor_sync(FULLMASK, s, 2));
            }
            if (lane == 0) {
                const float * kr = Kb + (size_t) t * ldkv;
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
    for (int u = tid; u < 32 * ncol; u += MEGA_NT) {
        const int col = u % ncol, r = u / ncol;
        const float * vp = Vb + (size_t) r * ldkv + c0 + col;
        float acc = 0.0f;
        int s0 = 0;
        for (; s0 + 8 * 32 <= np; s0 += 8 * 32) {
            float v[8];
#pragma unroll
            for (int k = 0; k < 8; k++) v[k] = __ldcg(vp + (size_t) (s0 + 32 * k) * ldkv);
#pragma unroll
            for (int k = 0; k < 8; k++) acc = fmaf(v[k], sc[s0 + 32 * k + r], acc);
        }
        for (; s0 < np; s0 += 32) acc = fmaf(__ldcg(vp + (size_t) s0 * ldkv), sc[s0 + r], acc);
        red[r * ncol + col] = acc;
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
        out[c0 + tid] = sumf;
    }
    __syncthreads();
}

// one matmul phase: rows [r0, r1) of the stacked matrices W[0..nmat) (rows_per rows each)
template <int FMT, class EpiF>
__device__ __forceinline__ void mega_matmul(const MegaParams & p, const uint8_t * const W[3], int rows_per, int r0, int r1,
                                            bool Kff, const uint8_t * s_act, const uint8_t ** s_rowptr,
                                            float * s_p, float * s_s, float * s_m, EpiF epi) {
    const int G = Kff ? p.Gf : p.Gd, stride = Kff ? p.stride_f : p.stride_d;
    const int off_qh = Kff ? p.offqh_f : p.offqh_d, off_d = Kff ? p.offd_f : p.offd_d, off_m = Kff ? p.offm_f : p.offm_d;
    const int off_n = Kff ? p.offn_f : p.offn_d, off_dd = Kff ? p.offdd_f : p.offdd_d, off_s = Kff ? p.offs_f : p.offs_d;
    // rows per tile: bounded by the scratch (sm_p holds RT*8*(4G+4) floats)
    int RT = MEGA_RT;
    if (FMT != BG_F16) { const int cap = p.sm_p / (8 * (4 * G + 4) * 4); RT = cap < RT ? cap : RT; RT = RT < 1 ? 1 : RT; }
    for (int t0 = r0; t0 < r1; t0 += RT) {
        const int rt = (r1 - t0) < RT ? (r1 - t0) : RT;
        if ((int) threadIdx.x < rt) {
            const int r = t0 + threadIdx.x, mat = r / rows_per;
            s_rowptr[threadIdx.x] = W[mat] + (size_t) (r - mat * rows_per) * stride;
        }
        __syncthreads();
        auto epi_row = [&](int tile_row, float v) { epi(t0 + tile_row, v); };
        if (FMT == BG_F16) mega_tile_f16(s_rowptr, rt, G, s_act, epi_row);
        else mega_tile_q<FMT>(s_rowptr, rt, G, off_qh, off_d, off_m, s_act, off_n, off_dd, off_s, s_p, s_s, s_m, epi_row);
    }
}

#define PROF(ph, k) do { if (p.prof && blockIdx.x == 0 && threadIdx.x == 0) p.prof[(l * 5 + (ph)) * 3 + (k)] = clock64(); } while (0)

template <int FMT, int DK>
__global__ void __launch_bounds__(MEGA_NT, 1) k_mega(MegaParams p) {
    extern __shared__ __align__(16) uint8_t smem[];
    __shared__ double sd[32];
    __shared__ float smx[32];
    __shared__ const uint8_t * s_rowptr[MEGA_RT];
    __shared__ int s_tok;
    float * s_row = (float *) (smem + p.sm_row);
    uint8_t * s_act = smem + p.sm_act;
    float * s_p = (float *) (smem + p.sm_p);
    float * s_s = (float *) (smem + p.sm_s);
    float * s_m = (float *) (smem + p.sm_m);
    float * s_attn = (float *) (smem + p.sm_attn);
    const int tid = threadIdx.x, d = p.d, ff = p.ff;
    unsigned long long bar_target = p.bar_base;
    const unsigned long long nct = gridDim.x;

    // ---- input token: given, or argmax over the candidates the previous launch left
    if (tid < 32) {
        int tok;
        if (p.use_cand) {
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
        } else tok = __ldcg(p.tok);
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
        // ================= phase 1: LN0 + q,k,v =================
        PROF(0, 0);
        {   // start the HBM reads of the NEXT layer's weights (or of lm_head) into L2 now
            int r0, r1;
            if (l + 1 < p.n_layer) {
                const MegaLayer Ln = p.layers[l + 1];
                mega_my_rows(3 * d, r0, r1);
                for (int mat = 0; mat < 3; mat++) {
                    const int a = max(r0, mat * d) - mat * d, b = min(r1, (mat + 1) * d) - mat * d;
                    if (b > a) mega_prefetch_l2((mat == 0 ? Ln.q_w : mat == 1 ? Ln.k_w : Ln.v_w) + (size_t) a * p.stride_d, (size_t) (b - a) * p.stride_d);
                }
                mega_my_rows(d, r0, r1);
                mega_prefetch_l2(Ln.o_w + (size_t) r0 * p.stride_d, (size_t) (r1 - r0) * p.stride_d);
                mega_prefetch_l2(Ln.fc2_w + (size_t) r0 * p.stride_f, (size_t) (r1 - r0) * p.stride_f);
                mega_my_rows(ff, r0, r1);
                mega_prefetch_l2(Ln.fc1_w + (size_t) r0 * p.stride_d, (size_t) (r1 - r0) * p.stride_d);
            } else {
                mega_my_rows(p.n_vocab, r0, r1);
                mega_prefetch_l2(p.lm_head + (size_t) r0 * p.stride_d, (size_t) (r1 - r0) * p.stride_d);
            }
        }
        if (l > 0) { for (int c = tid; c < d; c += MEGA_NT) s_row[c] = __ldcg(p.x + c); __syncthreads(); }
        bg_ln_row(s_row, d, L.ln0_w, L.ln0_b, p.eps, sd);
        bg_row_to_record(s_row, d, p.wtype, s_act, p.actb_d, p.offn_d, p.offdd_d, p.offs_d, p.code_off);
        __syncthreads();
        PROF(0, 1);
        {
            int r0, r1; mega_my_rows(3 * d, r0, r1);
            const uint8_t * W[3] = { L.q_w, L.k_w, L.v_w };
            const int pos = p.n_past;
            mega_matmul<FMT>(p, W, d, r0, r1, false, s_act, s_rowptr, s_p, s_s, s_m, [&](int r, float v) {
                const int mat = r / d, rr = r - mat * d;
                const float * b = mat == 0 ? L.q_b : (mat == 1 ? L.k_b : L.v_b);
                const float t = __fadd_rn(b[rr], v);
                if (mat == 0) p.q[rr] = __fmul_rn(t, p.qscale);
                else (mat == 1 ? kc : vc)[(size_t) pos * d + rr] = t;
            });
        }
        PROF(0, 2);
        bar_target += nct; mega_grid_barrier(p.bar, bar_target);
        // ================= phase 2: attention, attn_parts CTAs per head (column split) =================
        PROF(1, 0);
        for (int w = blockIdx.x; w < p.n_head * p.attn_parts; w += gridDim.x) {
            const int h = w / p.attn_parts, part = w - h * p.attn_parts;
            const int ncol = DK / p.attn_parts;
            float * s_q = s_row;                       // d floats free to reuse here
            for (int c = tid; c < DK; c += MEGA_NT) s_q[c] = __ldcg(p.q + h * DK + c);
            __syncthreads();
            mega_attention_cols<DK>(s_q, kc + (size_t) h * DK, vc + (size_t) h * DK, d, p.n_past + 1, p.exp_tab,
                                    s_attn, p.n_positions, sd, smx, part * ncol, ncol, p.att + (size_t) h * DK);
        }
        PROF(1, 2);
        bar_target += nct; mega_grid_barrier(p.bar, bar_target);
        // ================= phase 3: out_proj + bias + residual =================
        PROF(2, 0);
        for (int c = tid; c < d; c += MEGA_NT) s_row[c] = __ldcg(p.att + c);
        __syncthreads();
        bg_row_to_record(s_row, d, p.wtype, s_act, p.actb_d, p.offn_d, p.offdd_d, p.offs_d, p.code_off);
        __syncthreads();
        PROF(2, 1);
        {
            int r0, r1; mega_my_rows(d, r0, r1);
            const uint8_t * W[3] = { L.o_w, L.o_w, L.o_w };
            mega_matmul<FMT>(p, W, d, r0, r1, false, s_act, s_rowptr, s_p, s_s, s_m, [&](int r, float v) {
                const float t = __fadd_rn(v, L.o_b[r]);
                p.x1[r] = __fadd_rn(t, __ldcg(p.x + r));
            });
        }
        PROF(2, 2);
        bar_target += nct; mega_grid_barrier(p.bar, bar_target);
        // ================= phase 4: LN1 + fc1 + bias + GELU =================
        PROF(3, 0);
        for (int c = tid; c < d; c += MEGA_NT) s_row[c] = __ldcg(p.x1 + c);
        __syncthreads();
        bg_ln_row(s_row, d, L.ln1_w, L.ln1_b, p.eps, sd);
        bg_row_to_record(s_row, d, p.wtype, s_act, p.actb_d, p.offn_d, p.offdd_d, p.offs_d, p.code_off);
        __syncthreads();
        PROF(3, 1);
        {
            int r0, r1; mega_my_rows(ff, r0, r1);
            const uint8_t * W[3] = { L.fc1_w, L.fc1_w, L.fc1_w };
            mega_matmul<FMT>(p, W, ff, r0, r1, false, s_act, s_rowptr, s_p, s_s, s_m, [&](int r, float v) {
                const float t = __fadd_rn(L.fc1_b[r], v);
                p.hff[r] = bg_h2f(p.gelu[bg_f2h(t)]);
            });
        }
        PROF(3, 2);
        bar_target += nct; mega_grid_barrier(p.bar, bar_target);
        // ================= phase 5: fc2 + bias + residual =================
        PROF(4, 0);
        for (int c = tid; c < ff; c += MEGA_NT) s_row[c] = __ldcg(p.hff + c);
        __syncthreads();
        bg_row_to_record(s_row, ff, p.wtype, s_act, p.actb_f, p.offn_f, p.offdd_f, p.offs_f, p.code_off);
        __syncthreads();
        PROF(4, 1);
        {
            int r0, r1; mega_my_rows(d, r0, r1);
            const uint8_t * W[3] = { L.fc2_w, L.fc2_w, L.fc2_w };
            mega_matmul<FMT>(p, W, d, r0, r1, true, s_act, s_rowptr, s_p, s_s, s_m, [&](int r, float v) {
                const float t = __fadd_rn(L.fc2_b[r], v);
                p.x[r] = __fadd_rn(t, __ldcg(p.x1 + r));
            });
        }
        PROF(4, 2);
        bar_target += nct; mega_grid_barrier(p.bar, bar_target);
    }
    // ================= final LayerNorm + lm_head =================
    { const int l = p.n_layer; PROF(0, 0); }
    for (int c = tid; c < d; c += MEGA_NT) s_row[c] = __ldcg(p.x + c);
    __syncthreads();
    bg_ln_row(s_row, d, p.lnf_w, p.lnf_b, p.eps, sd);
    bg_row_to_record(s_row, d, p.wtype, s_act, p.actb_d, p.offn_d, p.offdd_d, p.offs_d, p.code_off);
    __syncthreads();
    { const int l = p.n_layer; PROF(0, 1); }
    {
        int r0, r1; mega_my_rows(p.n_vocab, r0, r1);
        const uint8_t * W[3] = { p.lm_head, p.lm_head, p.lm_head };
        float best = -INFINITY; int bi = 0x7fffffff;
        mega_matmul<FMT>(p, W, p.n_vocab, r0, r1, false, s_act, s_rowptr, s_p, s_s, s_m, [&](int r, float v) {
            p.logits[r] = v;
            if (v > best || (v == best && r < bi)) { best = v; bi = r; }
        });
        // per-CTA argmax candidate (first index wins ties) for the next launch's prologue
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(FULLMASK, best, o); const int oi = __shfl_xor_sync(FULLMASK, bi, o);
            if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
        }
        int * si = (int *) (smx + 16);      // smx: 32 floats; [0,16) values, [16,32) indices
        if ((tid & 31) == 0) { smx[tid >> 5] = best; si[tid >> 5] = bi; }
        __syncthreads();
        if (tid == 0) {
            for (int i = 1; i < MEGA_NT / 32; i++) if (smx[i] > best || (smx[i] == best && si[i] < bi)) { best = smx[i]; bi = si[i]; }
            p.cand_val[blockIdx.x] = best; p.cand_idx[blockIdx.x] = bi;
        }
    }
    { const int l = p.n_layer; PROF(0, 2); }
    // first layer of the NEXT token: its weights are static, start pulling them into L2
    {
        int r0, r1;
        const MegaLayer Ln = p.layers[0];
        mega_my_rows(3 * d, r0, r1);
        for (int mat = 0; mat < 3; mat++) {
            const int a = max(r0, mat * d) - mat * d, b = min(r1, (mat + 1) * d) - mat * d;
            if (b > a) mega_prefetch_l2((mat == 0 ? Ln.q_w : mat == 1 ? Ln.k_w : Ln.v_w) + (size_t) a * p.stride_d, (size_t) (b - a) * p.stride_d);
        }
        mega_my_rows(d, r0, r1);
        mega_prefetch_l2(Ln.o_w + (size_t) r0 * p.stride_d, (size_t) (r1 - r0) * p.stride_d);
        mega_prefetch_l2(Ln.fc2_w + (size_t) r0 * p.stride_f, (size_t) (r1 - r0) * p.stride_f);
        mega_my_rows(ff, r0, r1);
        mega_prefetch_l2(Ln.fc1_w + (size_t) r0 * p.stride_d, (size_t) (r1 - r0) * p.stride_d);
    }
}

// reduces the candidates of the last launch of a greedy loop into the id log
__global__ void k_mega_pick(const float * __restrict__ cand_val, const int * __restrict__ cand_idx, int n_cand,
                            int * __restrict__ idlog, int slot, int * __restrict__ next_tok) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        float best = -INFINITY; int bi = 0x7fffffff;
        for (int i = 0; i < n_cand; i++) { const float v = cand_val[i]; const int ix = cand_idx[i]; if (v > best || (v == best && ix < bi)) { best = v; bi = ix; } }
        if (bi == 0x7fffffff) bi = 0;
        idlog[slot] = bi;
        if (next_tok) next_tok[0] = bi;
    }
}
