// bgpt_kernels.cuh -- hand-written sm_100a kernels of the `biogpt_eval` hot path.
//
// Arithmetic contract ("lane order").  Every kernel below reproduces the floating-point
// operation ORDER of the reference's AVX2 CPU path, so the logits are the reference's logits
// bit for bit, not merely close:
//   * quantised dots keep the 8 running sums of ggml_vec_dot_q*_q8_* (ggml.c:2518-2541,
//     2824-2857, 3071-3093, 3386-3411, 3597-3618): sum l takes elements 4l..4l+3 of each
//     block, acc_l = fma(d_w*d_a, (float) isum_l, acc_l) in block order, then hsum_float_8
//     (ggml.c:611-617); the Q4_1/Q5_1 `summs` chain is fma(m_w, s_a, summs).
//   * F16/F32 dots keep the 4x8 accumulators of ggml_vec_dot_f16/f32 (ggml.c:2372-2443) and
//     the GGML_F32x8_REDUCE tree (ggml.c:1981-1999): lane = element % 32, reduced with
//     xor-shuffles 16, 8, 4, 1, 2.
//   * LayerNorm and softmax accumulate in double like ggml.c:11403-11420, 12955-12974; the
//     fp16 GELU/exp tables are the host-built tables of ggml.c:4620-4640.
// The file is compiled with -fmad=false: a multiply-add is fused only where it is written as
// fmaf(), exactly where the reference binary fuses it.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include "bgpt_layout.h"

#define BG_WARP 32
#define FULLMASK 0xffffffffu

__device__ __forceinline__ float bg_h2f(uint16_t h) { return __half2float(__ushort_as_half(h)); }
__device__ __forceinline__ uint16_t bg_f2h(float f) { return __half_as_ushort(__float2half_rn(f)); }

// streaming 128-bit weight load: read-only path, do not pollute L1 (each byte is used once)
__device__ __forceinline__ uint4 ldg_stream128(const void * p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ uint2 ldg_stream64(const void * p) {
    uint2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ uint32_t ldg_stream32(const void * p) {
    uint32_t r;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}

// ---------------------------------------------------------------------------------------------
// per-eval state that lives on the device, so that a captured CUDA graph can be replayed for
// any position: kernels read n_past through this pointer.
// ---------------------------------------------------------------------------------------------
struct DevState {
    int n_past;     // positions already in the KV cache
    int step;       // greedy-decode step counter (index into the id log)
    int pad0, pad1;
};
// batch row -> (stream, position, attention length).  mode 0: prompt rows of stream 0 at
// n_past+row, all attending to n_past+n positions (NO causal mask, biogpt.cpp:741-744).
// mode 1: lock-step streams, row = stream, one token each at n_past.
__device__ __forceinline__ void bg_row_info(int mode, int n, int n_past, int row, int & stream, int & pos, int & T) {
    if (mode == 0) { stream = 0; pos = n_past + row; T = n_past + n; }
    else           { stream = row; pos = n_past;     T = n_past + 1; }
}

// ---------------------------------------------------------------------------------------------
// embedding: get_rows(embed_tokens) * sqrt(d_model) + get_rows(embed_pos, n_past+i+2)
// (biogpt.cpp:663-686; dequantize_row_q*, ggml.c:1536-1646).  Embedding tables stay in the
// file's AoS layout: two rows per token are read, nothing streams.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float bg_dequant_elem(int type, const uint8_t * row, int c) {
    if (type == BG_F32) return ((const float *) row)[c];
    if (type == BG_F16) return bg_h2f(((const uint16_t *) row)[c]);
    const int b = c >> 5, e = c & 31, j = e & 15;
    const uint8_t * blk = row + (size_t) b * (type == BG_Q4_0 ? 18 : type == BG_Q4_1 ? 20 : type == BG_Q5_0 ? 22 : type == BG_Q5_1 ? 24 : 34);
    const float d = bg_h2f(*(const uint16_t *) blk);
    if (type == BG_Q8_0) return __fmul_rn((float) (int) (int8_t) blk[2 + e], d);
    if (type == BG_Q4_0) { const int q = (e < 16) ? (blk[2 + j] & 0x0F) : (blk[2 + j] >> 4); return __fmul_rn((float) (q - 8), d); }
    if (type == BG_Q4_1) {
        const float m = bg_h2f(*(const uint16_t *) (blk + 2));
        const int q = (e < 16) ? (blk[4 + j] & 0x0F) : (blk[4 + j] >> 4);
        return __fadd_rn(__fmul_rn((float) q, d), m);
    }
    if (type == BG_Q5_0) {
        const uint32_t qh = (uint32_t) blk[2] | ((uint32_t) blk[3] << 8) | ((uint32_t) blk[4] << 16) | ((uint32_t) blk[5] << 24);
        const int nib = (e < 16) ? (blk[6 + j] & 0x0F) : (blk[6 + j] >> 4);
        const int q = nib | (int) (((qh >> e) & 1u) << 4);
        return __fmul_rn((float) (q - 16), d);
    }
    { // Q5_1
        const float m = bg_h2f(*(const uint16_t *) (blk + 2));
        const uint32_t qh = (uint32_t) blk[4] | ((uint32_t) blk[5] << 8) | ((uint32_t) blk[6] << 16) | ((uint32_t) blk[7] << 24);
        const int nib = (e < 16) ? (blk[8 + j] & 0x0F) : (blk[8 + j] >> 4);
        const int q = nib | (int) (((qh >> e) & 1u) << 4);
        return __fadd_rn(__fmul_rn((float) q, d), m);
    }
}

static __global__ void k_embed(const uint8_t * __restrict__ tokW, const uint8_t * __restrict__ posW, int type,
                        const int * __restrict__ toks, const DevState * __restrict__ st, int mode, int n,
                        int d, int n_vocab, int n_pos_rows, float scale, float * __restrict__ x) {
    const int row = blockIdx.x;
    int stream, pos, T; bg_row_info(mode, n, st->n_past, row, stream, pos, T);
    int tok = toks[row]; tok = tok < 0 ? 0 : (tok >= n_vocab ? n_vocab - 1 : tok);
    int prow = pos + 2; prow = prow >= n_pos_rows ? n_pos_rows - 1 : prow;
    const size_t rb = (size_t) d / (bg_is_quant(type) ? 32 : 1) * (type == BG_F32 ? 4 : type == BG_F16 ? 2 : type == BG_Q4_0 ? 18 : type == BG_Q4_1 ? 20 : type == BG_Q5_0 ? 22 : type == BG_Q5_1 ? 24 : 34);
    const uint8_t * tr = tokW + rb * (size_t) tok;
    const uint8_t * pr = posW + rb * (size_t) prow;
    for (int c = threadIdx.x; c < d; c += blockDim.x) {
        const float a = __fmul_rn(bg_dequant_elem(type, tr, c), scale);
        x[(size_t) row * d + c] = __fadd_rn(a, bg_dequant_elem(type, pr, c));
    }
}

static __global__ void k_dequant_rows(const uint8_t * __restrict__ W, int type, int K, float * __restrict__ y) {
    const size_t rb = (size_t) K / (bg_is_quant(type) ? 32 : 1) * (type == BG_F32 ? 4 : type == BG_F16 ? 2 : type == BG_Q4_0 ? 18 : type == BG_Q4_1 ? 20 : type == BG_Q5_0 ? 22 : type == BG_Q5_1 ? 24 : 34);
    const uint8_t * r = W + rb * blockIdx.x;
    for (int c = threadIdx.x; c < K; c += blockDim.x) y[(size_t) blockIdx.x * K + c] = bg_dequant_elem(type, r, c);
}

// ---------------------------------------------------------------------------------------------
// block reductions (256 threads)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double bg_block_sum_f64(double v, double * scratch /*>= 32*/) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULLMASK, v, o);
    const int w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) scratch[w] = v;
    __syncthreads();
    double t = 0.0;
    for (int i = 0; i < nw; i++) t += scratch[i];
    return t;
}
__device__ __forceinline__ float bg_block_max_f32(float v, float * scratch /*>= 32*/) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(FULLMASK, v, o));
    const int w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) scratch[w] = v;
    __syncthreads();
    float t = scratch[0];
    for (int i = 1; i < nw; i++) t = fmaxf(t, scratch[i]);
    return t;
}

// ---------------------------------------------------------------------------------------------
// k_act: (optional LayerNorm + affine) then conversion of one token row to the activation
// record its consumer matmul needs -- the work mul_mat's INIT phase does on the CPU
// (ggml.c:11909-11925), fused behind the producer.
//   LayerNorm: ggml_compute_forward_norm_f32, ggml.c:11377-11426, then w*y, +b as separate
//              roundings (biogpt.cpp:693-700).
//   Q8_0: quantize_row_q8_0 AVX branch ggml.c:1166-1203 (d=amax/127 -> fp16, id=127/amax, RNE)
//   Q8_1: quantize_row_q8_1 AVX2 branch ggml.c:1403-1450 (f32 d, s = d*sum(q))
//   F16 : ggml_fp32_to_fp16_row, ggml.c:493-510
// grid = token rows, block = 256, dynamic smem = K floats.
// ---------------------------------------------------------------------------------------------
struct ActArgs {
    const float * in; int ld_in;         // [rows][ld_in]
    const float * lnw; const float * lnb; int do_ln; float eps;
    int K; int wtype;                    // consumer weight type -> record kind
    uint8_t * act; int act_bytes;        // record base and stride
    int off_n, off_d, off_s, code_off;
    float * f32_out; int ld_out;         // optional copy of the f32 row (taps / unit tests)
};

// LayerNorm (+affine) of the K floats in shared memory `srow`, in place.  All threads of the CTA.
__device__ __forceinline__ void bg_ln_row(float * srow, int K, const float * __restrict__ lnw, const float * __restrict__ lnb,
                                          float eps, double * sd /*>= 32 doubles*/) {
    const int tid = threadIdx.x;
    double s = 0.0;
    for (int c = tid; c < K; c += blockDim.x) s += (double) srow[c];
    s = bg_block_sum_f64(s, sd);
    // s / K: for K a power of two the product with the exact reciprocal is the same double
    const bool kpow2 = (K & (K - 1)) == 0;
    const double invK = 1.0 / (double) K;
    const float mean = (float) (kpow2 ? s * invK : s / K);
    double s2 = 0.0;
    for (int c = tid; c < K; c += blockDim.x) {
        const float v = __fsub_rn(srow[c], mean);
        srow[c] = v;
        s2 += (double) __fmul_rn(v, v);
    }
    s2 = bg_block_sum_f64(s2, sd);
    const float variance = (float) (kpow2 ? s2 * invK : s2 / K);
    const float scale = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(variance, eps)));
    for (int c = tid; c < K; c += blockDim.x) {
        float y = __fmul_rn(srow[c], scale);
        if (lnw) y = __fmul_rn(lnw[c], y);
        if (lnb) y = __fadd_rn(y, lnb[c]);
        srow[c] = y;
    }
    __syncthreads();
}

// Same LayerNorm for a CTA of NT threads with K <= EPT*NT: the affine parameters of this thread's
// elements (c = tid + i*NT) were loaded into registers earlier, so no global load sits between
// the reductions and the output.
template <int NT, int EPT>
__device__ __forceinline__ void bg_ln_row_pre(float * srow, int K, const float (&lw)[EPT], const float (&lb)[EPT], float eps, double * sd) {
    const int tid = threadIdx.x;
    float xv[EPT];
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < EPT; i++) { const int c = tid + i * NT; xv[i] = c < K ? srow[c] : 0.0f; if (c < K) s += (double) xv[i]; }
    s = bg_block_sum_f64(s, sd);
    const bool kpow2 = (K & (K - 1)) == 0;
    const double invK = 1.0 / (double) K;
    const float mean = (float) (kpow2 ? s * invK : s / K);
    double s2 = 0.0;
#pragma unroll
    for (int i = 0; i < EPT; i++) { const int c = tid + i * NT; xv[i] = __fsub_rn(xv[i], mean); if (c < K) s2 += (double) __fmul_rn(xv[i], xv[i]); }
    s2 = bg_block_sum_f64(s2, sd);
    const float variance = (float) (kpow2 ? s2 * invK : s2 / K);
    const float scale = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(variance, eps)));
#pragma unroll
    for (int i = 0; i < EPT; i++) {
        const int c = tid + i * NT;
        if (c < K) srow[c] = __fadd_rn(__fmul_rn(lw[i], __fmul_rn(xv[i], scale)), lb[i]);
    }
    __syncthreads();
}

// f32 row in shared memory -> activation record (global or shared).  All threads of the CTA;
// the caller synchronises before consuming the record.
__device__ __forceinline__ void bg_row_to_record(const float * srow, int K, int wtype, uint8_t * rec, int act_bytes,
                                                 int off_n, int off_d, int off_s, int code_off) {
    const int tid = threadIdx.x;
    const int kind = bg_act_kind(wtype);
    if (kind == ACT_F32) {
        float * o = (float *) rec;
        const int Kp = act_bytes / 4;
        for (int i = tid; i < Kp; i += blockDim.x) {     // i = ((gg*32 + lane)*4 + e)
            const int e = i & 3, lane = (i >> 2) & 31, gg = i >> 7;
            const int c = gg * 128 + e * 32 + lane;
            o[i] = c < K ? srow[c] : 0.0f;
        }
        return;
    }
    if (kind == ACT_F16) {
        float * o = (float *) rec;
        const int Kp = act_bytes / 4;
        for (int i = tid; i < Kp; i += blockDim.x) {     // i = (((gg*2 + eh)*32 + lane)*4 + ee)
            const int ee = i & 3, lane = (i >> 2) & 31, eh = (i >> 7) & 1, gg = i >> 8;
            const int c = gg * 256 + (eh * 4 + ee) * 32 + lane;
            o[i] = c < K ? bg_h2f(bg_f2h(srow[c])) : 0.0f;
        }
        return;
    }
    // quantised records: 8 threads per 32-element block, thread l owns elements 4l..4l+3 (= one
    // 32-bit word of the record); the block maximum needs 3 shuffle levels
    const int nb = K >> 5, nbp = ((nb + 3) >> 2) << 2;
    uint32_t * aq = (uint32_t *) rec;
    int32_t  * an = (int32_t *) (rec + off_n);
    float    * ad = (float *) (rec + off_d);
    float    * as = (float *) (rec + off_s);
    for (int idx = tid; idx < nbp * 8; idx += blockDim.x) {      // nbp*8 is a multiple of 32: warps stay whole
        const int b = idx >> 3, l = idx & 7;
        const int g = b >> 2, i = b & 3;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (b < nb) v = *(const float4 *) (srow + b * 32 + 4 * l);
        float amax = fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w)));
        amax = fmaxf(amax, __shfl_xor_sync(FULLMASK, amax, 1));
        amax = fmaxf(amax, __shfl_xor_sync(FULLMASK, amax, 2));
        amax = fmaxf(amax, __shfl_xor_sync(FULLMASK, amax, 4));
        const float d  = __fdiv_rn(amax, 127.0f);
        const float id = (amax != 0.0f) ? __fdiv_rn(127.0f, amax) : 0.0f;
        const int q0 = __float2int_rn(__fmul_rn(v.x, id)), q1 = __float2int_rn(__fmul_rn(v.y, id));
        const int q2 = __float2int_rn(__fmul_rn(v.z, id)), q3 = __float2int_rn(__fmul_rn(v.w, id));
        const int s4 = q0 + q1 + q2 + q3;
        aq[(g * 8 + l) * 4 + i] = ((uint32_t) q0 & 0xFFu) | (((uint32_t) q1 & 0xFFu) << 8) | (((uint32_t) q2 & 0xFFu) << 16) | (((uint32_t) q3 & 0xFFu) << 24);
        an[(g * 8 + l) * 4 + i] = -code_off * s4;
        if (kind == ACT_Q8_0) { if (l == 0) { ad[b] = bg_h2f(bg_f2h(d)); as[b] = 0.0f; } }
        else {
            int stot = s4;
            stot += __shfl_xor_sync(FULLMASK, stot, 1);
            stot += __shfl_xor_sync(FULLMASK, stot, 2);
            stot += __shfl_xor_sync(FULLMASK, stot, 4);
            if (l == 0) { ad[b] = d; as[b] = __fmul_rn(d, (float) stot); }
        }
    }
}

static __global__ void __launch_bounds__(256) k_act(ActArgs a) {
    extern __shared__ __align__(16) float srow[];
    __shared__ double sd[32];
    const int row = blockIdx.x, tid = threadIdx.x, K = a.K;
    const float * in = a.in + (size_t) row * a.ld_in;
    for (int c = tid; c < K; c += blockDim.x) srow[c] = in[c];
    __syncthreads();
    if (a.do_ln) bg_ln_row(srow, K, a.lnw, a.lnb, a.eps, sd);
    if (a.f32_out) for (int c = tid; c < K; c += blockDim.x) a.f32_out[(size_t) row * a.ld_out + c] = srow[c];
    if (!a.act) return;
    bg_row_to_record(srow, K, a.wtype, a.act + (size_t) row * a.act_bytes, a.act_bytes, a.off_n, a.off_d, a.off_s, a.code_off);
}

// ---------------------------------------------------------------------------------------------
// GEMV / skinny GEMM epilogues (all the bias / scale / residual / GELU / KV-append nodes of
// biogpt.cpp:705-727, 767-772, 783-795 folded behind the dot products)
// ---------------------------------------------------------------------------------------------
enum { EPI_STORE = 0, EPI_QKV = 1, EPI_RESID = 2, EPI_GELU = 3 };
struct Epi {
    int kind;
    const float * bias[3];
    float * out; int ld_out;
    const float * resid; int ld_resid;
    float * kcache; float * vcache;      // layer base of stream 0
    size_t stream_stride;                // floats between streams' caches
    int d; float qscale;
    const DevState * st; int mode; int n;
    const uint16_t * gelu;
};
__device__ __forceinline__ void bg_epilogue(const Epi & e, int tok, int r, float v) {
    switch (e.kind) {
    case EPI_STORE:
        e.out[(size_t) tok * e.ld_out + r] = e.bias[0] ? __fadd_rn(e.bias[0][r], v) : v;
        break;
    case EPI_QKV: {
        const int mat = r / e.d, rr = r - mat * e.d;
        const float t = __fadd_rn(e.bias[mat][rr], v);
        if (mat == 0) { e.out[(size_t) tok * e.ld_out + rr] = __fmul_rn(t, e.qscale); }
        else {
            int stream, pos, T; bg_row_info(e.mode, e.n, e.st->n_past, tok, stream, pos, T);
            float * c = (mat == 1 ? e.kcache : e.vcache) + (size_t) stream * e.stream_stride + (size_t) pos * e.d + rr;
            *c = t;
        }
        break; }
    case EPI_RESID: {
        const float t = __fadd_rn(v, e.bias[0][r]);
        e.out[(size_t) tok * e.ld_out + r] = __fadd_rn(t, e.resid[(size_t) tok * e.ld_resid + r]);
        break; }
    case EPI_GELU: {
        const float t = __fadd_rn(e.bias[0][r], v);
        e.out[(size_t) tok * e.ld_out + r] = bg_h2f(e.gelu[bg_f2h(t)]);
        break; }
    }
}

// the same epilogues for NT consecutive token rows n0 .. n0 + NT - 1 of one weight row r (the tensor-core kernels: a thread owns a
// row and a run of tokens): everything that is loaded (bias, residual, GELU table) is loaded before the first store, so the loads
// of the NT tokens travel together instead of one dependent round trip per token
template <int NT>
__device__ __forceinline__ void bg_epilogue_rows(const Epi & e, int n0, int nmax, int r, const float (&v)[NT]) {
    switch (e.kind) {
    case EPI_STORE: {
        const bool hb = e.bias[0] != nullptr;
        const float b = hb ? e.bias[0][r] : 0.0f;
#pragma unroll
        for (int t = 0; t < NT; t++) if (n0 + t < nmax) e.out[(size_t) (n0 + t) * e.ld_out + r] = hb ? __fadd_rn(b, v[t]) : v[t];
        break; }
    case EPI_QKV: {
        const int mat = r / e.d, rr = r - mat * e.d;
        const float b = e.bias[mat][rr];
        if (mat == 0) {
#pragma unroll
            for (int t = 0; t < NT; t++) if (n0 + t < nmax) e.out[(size_t) (n0 + t) * e.ld_out + rr] = __fmul_rn(__fadd_rn(b, v[t]), e.qscale);
        } else {
            const int n_past = e.st->n_past;
            float * base = mat == 1 ? e.kcache : e.vcache;
#pragma unroll
            for (int t = 0; t < NT; t++) {
                if (n0 + t >= nmax) continue;
                int stream, pos, T; bg_row_info(e.mode, e.n, n_past, n0 + t, stream, pos, T);
                base[(size_t) stream * e.stream_stride + (size_t) pos * e.d + rr] = __fadd_rn(b, v[t]);
            }
        }
        break; }
    case EPI_RESID: {
        const float b = e.bias[0][r];
        float res[NT];
#pragma unroll
        for (int t = 0; t < NT; t++) res[t] = n0 + t < nmax ? e.resid[(size_t) (n0 + t) * e.ld_resid + r] : 0.0f;
#pragma unroll
        for (int t = 0; t < NT; t++) if (n0 + t < nmax) e.out[(size_t) (n0 + t) * e.ld_out + r] = __fadd_rn(__fadd_rn(v[t], b), res[t]);
        break; }
    case EPI_GELU: {
        const float b = e.bias[0][r];
        uint16_t g[NT];
#pragma unroll
        for (int t = 0; t < NT; t++) g[t] = e.gelu[bg_f2h(__fadd_rn(b, v[t]))];
#pragma unroll
        for (int t = 0; t < NT; t++) if (n0 + t < nmax) e.out[(size_t) (n0 + t) * e.ld_out + r] = bg_h2f(g[t]);
        break; }
    }
}

struct GemvArgs {
    const uint8_t * W[3];     // up to three stacked matrices (q,k,v) of rows_per rows each
    int rows_per;             // rows of one matrix
    int M;                    // total rows = rows_per * (#matrices)
    int G;                    // groups per row
    int stride;               // device row stride (bytes)
    int off_qh, off_d, off_m;
    const uint8_t * act; int act_bytes; int off_n, off_dd, off_s;
    int n;                    // token rows
    int tok0;                 // first token row handled (lm_head: last row only)
    Epi epi;
};

// expand 4 bits (bit k -> byte k) to 0x10 per set bit
__device__ __forceinline__ uint32_t bg_spread4(uint32_t nib) { return (nib * 0x02040810u) & 0x10101010u; }

// ---------------------------------------------------------------------------------------------
// k_gemv_q: y[tok][row] = dot(W[row], act[tok]) for the five block-quantised formats.
// 4 threads per row (thread j: running sums j and j+4), 8 rows per warp, TN token rows per
// CTA pass with their activation records staged in shared memory.  grid.x strides over row
// tiles, grid.y over token tiles.
// ---------------------------------------------------------------------------------------------
template <int FMT, int TN>
__global__ void __launch_bounds__(256) k_gemv_q(GemvArgs a) {
    extern __shared__ uint4 s_act4[];
    constexpr bool IS8   = (FMT == BG_Q8_0);
    constexpr bool HASQH = (FMT == BG_Q5_0 || FMT == BG_Q5_1);
    constexpr bool HASM  = (FMT == BG_Q4_1 || FMT == BG_Q5_1);
    constexpr bool HASOFF = (FMT == BG_Q4_0 || FMT == BG_Q5_0);
    const int tid = threadIdx.x;
    const int tokbase = a.tok0 + blockIdx.y * TN;
    {   // stage TN activation records (zeros for rows past n)
        const int v16 = a.act_bytes >> 4;
        for (int i = tid; i < TN * v16; i += blockDim.x) {
            const int t = i / v16, o = i - t * v16;
            uint4 z = make_uint4(0, 0, 0, 0);
            if (tokbase + t < a.n) z = *(const uint4 *) (a.act + (size_t) (tokbase + t) * a.act_bytes + (size_t) o * 16);
            s_act4[i] = z;
        }
    }
    __syncthreads();
    const uint8_t * s_act = (const uint8_t *) s_act4;
    const int warp = tid >> 5, lane = tid & 31, nw = blockDim.x >> 5;
    const int rr = lane >> 2, j = lane & 3;
    for (int rowbase = (blockIdx.x * nw + warp) * 8; rowbase < a.M; rowbase += gridDim.x * nw * 8) {
        const int row = rowbase + rr;
        const int rowc = row < a.M ? row : a.M - 1;
        const int mat = rowc / a.rows_per;
        const uint8_t * wrow = a.W[mat] + (size_t) (rowc - mat * a.rows_per) * a.stride;
        float acc0[TN], acc1[TN], summ[TN];
#pragma unroll
        for (int t = 0; t < TN; t++) { acc0[t] = 0.0f; acc1[t] = 0.0f; summ[t] = 0.0f; }
#pragma unroll 2
        for (int g = 0; g < a.G; g++) {
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
                if (HASQH) qh = ldg_stream32(wrow + a.off_qh + g * 16 + j * 4);
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
            const uint2 dh = ldg_stream64(wrow + a.off_d + g * 8);
            float dw[4];
            dw[0] = bg_h2f((uint16_t) (dh.x & 0xFFFF)); dw[1] = bg_h2f((uint16_t) (dh.x >> 16));
            dw[2] = bg_h2f((uint16_t) (dh.y & 0xFFFF)); dw[3] = bg_h2f((uint16_t) (dh.y >> 16));
            float mw[4] = { 0.f, 0.f, 0.f, 0.f };
            if (HASM) {
                const uint2 mh = ldg_stream64(wrow + a.off_m + g * 8);
                mw[0] = bg_h2f((uint16_t) (mh.x & 0xFFFF)); mw[1] = bg_h2f((uint16_t) (mh.x >> 16));
                mw[2] = bg_h2f((uint16_t) (mh.y & 0xFFFF)); mw[3] = bg_h2f((uint16_t) (mh.y >> 16));
            }
#pragma unroll
            for (int t = 0; t < TN; t++) {
                const uint8_t * rec = s_act + (size_t) t * a.act_bytes;
                const uint4 a0 = *(const uint4 *) (rec + (g * 8 + j) * 16);
                const uint4 a1 = *(const uint4 *) (rec + (g * 8 + j + 4) * 16);
                int4 n0 = make_int4(0, 0, 0, 0), n1 = make_int4(0, 0, 0, 0);
                if (HASOFF) {
                    n0 = *(const int4 *) (rec + a.off_n + (g * 8 + j) * 16);
                    n1 = *(const int4 *) (rec + a.off_n + (g * 8 + j + 4) * 16);
                }
                const float4 da = *(const float4 *) (rec + a.off_dd + g * 16);
                float4 sa = make_float4(0.f, 0.f, 0.f, 0.f);
                if (HASM) sa = *(const float4 *) (rec + a.off_s + g * 16);
                const uint32_t a0w[4] = { a0.x, a0.y, a0.z, a0.w }, a1w[4] = { a1.x, a1.y, a1.z, a1.w };
                const int n0w[4] = { n0.x, n0.y, n0.z, n0.w }, n1w[4] = { n1.x, n1.y, n1.z, n1.w };
                const float daw[4] = { da.x, da.y, da.z, da.w }, saw[4] = { sa.x, sa.y, sa.z, sa.w };
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const float s  = __fmul_rn(dw[i], daw[i]);
                    const float p0 = (float) __dp4a((int) lo[i], (int) a0w[i], n0w[i]);
                    const float p1 = (float) __dp4a((int) hi[i], (int) a1w[i], n1w[i]);
                    acc0[t] = fmaf(s, p0, acc0[t]);
                    acc1[t] = fmaf(s, p1, acc1[t]);
                    if (HASM) summ[t] = fmaf(mw[i], saw[i], summ[t]);
                }
            }
        }
        // hsum_float_8: ((a0+a4)+(a2+a6)) + ((a1+a5)+(a3+a7))
#pragma unroll
        for (int t = 0; t < TN; t++) {
            float r = __fadd_rn(acc1[t], acc0[t]);
            r = __fadd_rn(r, __shfl_xor_sync(FULLMASK, r, 2));
            r = __fadd_rn(r, __shfl_xor_sync(FULLMASK, r, 1));
            if (HASM) r = __fadd_rn(r, summ[t]);
            if (j == 0 && row < a.M && tokbase + t < a.n) bg_epilogue(a.epi, tokbase + t, row, r);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// k_gemv_f: F16 / F32 weights.  One warp per row, lane = running sum (element % 32).
// F16 activations were rounded through fp16 by k_act and are held as f32, so
// fma(h2f(w), a, acc) is the reference's GGML_F16_VEC_FMA (ggml.c:2421-2431).
// ---------------------------------------------------------------------------------------------
template <int FMT, int TN>
__global__ void __launch_bounds__(256) k_gemv_f(GemvArgs a) {
    extern __shared__ uint4 s_act4[];
    const int tid = threadIdx.x;
    const int tokbase = a.tok0 + blockIdx.y * TN;
    {
        const int v16 = a.act_bytes >> 4;
        for (int i = tid; i < TN * v16; i += blockDim.x) {
            const int t = i / v16, o = i - t * v16;
            uint4 z = make_uint4(0, 0, 0, 0);
            if (tokbase + t < a.n) z = *(const uint4 *) (a.act + (size_t) (tokbase + t) * a.act_bytes + (size_t) o * 16);
            s_act4[i] = z;
        }
    }
    __syncthreads();
    const uint8_t * s_act = (const uint8_t *) s_act4;
    const int warp = tid >> 5, lane = tid & 31, nw = blockDim.x >> 5;
    for (int row = blockIdx.x * nw + warp; row < a.M; row += gridDim.x * nw) {
        const int mat = row / a.rows_per;
        const uint8_t * wrow = a.W[mat] + (size_t) (row - mat * a.rows_per) * a.stride;
        float acc[TN];
#pragma unroll
        for (int t = 0; t < TN; t++) acc[t] = 0.0f;
#pragma unroll 2
        for (int g = 0; g < a.G; g++) {
            const uint4 w = ldg_stream128(wrow + (size_t) (g * 32 + lane) * 16);
            if (FMT == BG_F16) {
                float wf[8];
                const uint32_t ww[4] = { w.x, w.y, w.z, w.w };
#pragma unroll
                for (int i = 0; i < 4; i++) { wf[2 * i] = bg_h2f((uint16_t) (ww[i] & 0xFFFF)); wf[2 * i + 1] = bg_h2f((uint16_t) (ww[i] >> 16)); }
#pragma unroll
                for (int t = 0; t < TN; t++) {
                    const uint8_t * rec = s_act + (size_t) t * a.act_bytes;
                    const float4 x0 = *(const float4 *) (rec + (size_t) ((g * 2 + 0) * 32 + lane) * 16);
                    const float4 x1 = *(const float4 *) (rec + (size_t) ((g * 2 + 1) * 32 + lane) * 16);
                    float c = acc[t];
                    c = fmaf(wf[0], x0.x, c); c = fmaf(wf[1], x0.y, c); c = fmaf(wf[2], x0.z, c); c = fmaf(wf[3], x0.w, c);
                    c = fmaf(wf[4], x1.x, c); c = fmaf(wf[5], x1.y, c); c = fmaf(wf[6], x1.z, c); c = fmaf(wf[7], x1.w, c);
                    acc[t] = c;
                }
            } else {
#pragma unroll
                for (int t = 0; t < TN; t++) {
                    const uint8_t * rec = s_act + (size_t) t * a.act_bytes;
                    const float4 x0 = *(const float4 *) (rec + (size_t) (g * 32 + lane) * 16);
                    float c = acc[t];
                    c = fmaf(__uint_as_float(w.x), x0.x, c); c = fmaf(__uint_as_float(w.y), x0.y, c);
                    c = fmaf(__uint_as_float(w.z), x0.z, c); c = fmaf(__uint_as_float(w.w), x0.w, c);
                    acc[t] = c;
                }
            }
        }
        // GGML_F32x8_REDUCE: (s0+s2), (s1+s3), sum, then lanes l+4, pairs, pairs
#pragma unroll
        for (int t = 0; t < TN; t++) {
            float r = acc[t];
            r = __fadd_rn(r, __shfl_xor_sync(FULLMASK, r, 16));
            r = __fadd_rn(r, __shfl_xor_sync(FULLMASK, r, 8));
            r = __fadd_rn(r, __shfl_xor_sync(FULLMASK, r, 4));
            r = __fadd_rn(r, __shfl_xor_sync(FULLMASK, r, 1));
            r = __fadd_rn(r, __shfl_xor_sync(FULLMASK, r, 2));
            if (lane == 0 && tokbase + t < a.n) bg_epilogue(a.epi, tokbase + t, row, r);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// k_attn: softmax(K q) V for one (head, token row), un-masked over T positions.
//   scores : ggml_vec_dot_f32(d_kv, K[t], q)            (mul_mat on F32 src0, ggml.c:2372-2407)
//   softmax: max, fp16-table exp of fp16(x-max), double sum (ggml.c:12914-12983)
//   output : ggml_vec_dot_f32(T, V_trans[c], p): 32 running sums by t%32, the reduce tree, then
//            the as-built tail of the reference binary (unfused in groups of 4, fused last <=3)
// grid = (n_head, rows), block = 256, dynamic smem = (Tmax + 64*DK) floats
// ---------------------------------------------------------------------------------------------
struct AttnArgs {
    const float * q; int ld_q;
    const float * kcache; const float * vcache; size_t stream_stride;
    float * out; int ld_out;
    int d; int n; int mode; const DevState * st;
    const uint16_t * exp_tab;
    int Tmax;
};

// one (head, row) of attention by the whole CTA of NT threads; `s_f` = (Tmax + 32*DK) floats of
// shared memory.  KV reads bypass L1 (__ldcg): rows appended by other CTAs earlier in a
// persistent kernel must never be served from a stale L1 line.
template <int DK, int NT>
__device__ __forceinline__ void bg_attention_head(const float * __restrict__ q /*DK*/, const float * Kb, const float * Vb, int ldkv, int T,
                                                  const uint16_t * __restrict__ exp_tab, float * s_f, int Tmax,
                                                  double * sd /*32*/, float * sm /*32*/, float * out /*DK*/) {
    float * sc  = s_f;                 // [Tmax]
    float * red = s_f + Tmax;          // [32][DK]
    float * tailv = red + 32 * DK;     // [<=31][DK] V rows of the scalar tail, staged so the tail is not a chain of global loads
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr int NWARP = NT / 32;
    constexpr int NP = DK & ~31;                 // vectorised part of the d_kv-long dot
    constexpr int NV = NP + ((DK - NP) & ~3);
    // ---- scores
    float qreg[(NP > 0 ? NP / 32 : 1)];
#pragma unroll
    for (int i = 0; i < NP / 32; i++) qreg[i] = q[i * 32 + lane];
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
            for (int i = NP; i < NV; i++) s = __fadd_rn(s, __fmul_rn(__ldcg(kr + i), q[i]));
#pragma unroll
            for (int i = NV; i < DK; i++) s = fmaf(__ldcg(kr + i), q[i], s);
            sc[t] = s;
        }
    }
    __syncthreads();
    // ---- softmax
    float mx = -INFINITY;
    for (int t = tid; t < T; t += NT) mx = fmaxf(mx, sc[t]);
    mx = bg_block_max_f32(mx, sm);
    double sum = 0.0;
    for (int t = tid; t < T; t += NT) {
        const float v = bg_h2f(exp_tab[bg_f2h(__fsub_rn(sc[t], mx))]);
        sc[t] = v;
        sum += (double) v;
    }
    sum = bg_block_sum_f64(sum, sd);
    const float inv = (float) (1.0 / sum);
    for (int t = tid; t < T; t += NT) sc[t] = __fmul_rn(sc[t], inv);
    __syncthreads();
    // ---- output: 32 running sums per column
    constexpr int NG = NT / DK;           // thread groups
    constexpr int CH = (32 + NG - 1) / NG; // running sums per thread
    const int c = tid % DK, grp = tid / DK;
    const int np = T & ~31;
    for (int i = tid; i < (T - np) * DK; i += NT) tailv[i] = __ldcg(Vb + (size_t) (np + i / DK) * ldkv + (i % DK));
    {
        float acc[CH];
#pragma unroll
        for (int u = 0; u < CH; u++) acc[u] = 0.0f;
        for (int s0 = 0; s0 < np; s0 += 32) {
#pragma unroll
            for (int u = 0; u < CH; u++) {
                const int r = grp * CH + u;
                if (r < 32) { const int t = s0 + r; acc[u] = fmaf(__ldcg(Vb + (size_t) t * ldkv + c), sc[t], acc[u]); }
            }
        }
#pragma unroll
        for (int u = 0; u < CH; u++) { const int r = grp * CH + u; if (r < 32) red[r * DK + c] = acc[u]; }
    }
    __syncthreads();
    if (tid < DK) {
        float x0[8];
#pragma unroll
        for (int l = 0; l < 8; l++) {
            const float a02 = __fadd_rn(red[(0 * 8 + l) * DK + tid], red[(2 * 8 + l) * DK + tid]);
            const float a13 = __fadd_rn(red[(1 * 8 + l) * DK + tid], red[(3 * 8 + l) * DK + tid]);
            x0[l] = __fadd_rn(a02, a13);
        }
        const float t0 = __fadd_rn(x0[0], x0[4]), t1 = __fadd_rn(x0[1], x0[5]);
        const float t2 = __fadd_rn(x0[2], x0[6]), t3 = __fadd_rn(x0[3], x0[7]);
        float sumf = __fadd_rn(__fadd_rn(t0, t1), __fadd_rn(t2, t3));
        const int nv = np + ((T - np) & ~3);
        int t = np;
        for (; t < nv; t++) sumf = __fadd_rn(sumf, __fmul_rn(tailv[(t - np) * DK + tid], sc[t]));
        for (; t < T;  t++) sumf = fmaf(tailv[(t - np) * DK + tid], sc[t], sumf);
        out[tid] = sumf;
    }
    __syncthreads();
}

template <int DK>
__global__ void __launch_bounds__(256) k_attn(AttnArgs a) {
    extern __shared__ float s_f[];
    __shared__ double sd[32];
    __shared__ float sm[32];
    const int h = blockIdx.x, row = blockIdx.y;
    int stream, pos, T; bg_row_info(a.mode, a.n, a.st->n_past, row, stream, pos, T);
    const float * Kb = a.kcache + (size_t) stream * a.stream_stride + (size_t) h * DK;
    const float * Vb = a.vcache + (size_t) stream * a.stream_stride + (size_t) h * DK;
    bg_attention_head<DK, 256>(a.q + (size_t) row * a.ld_q + (size_t) h * DK, Kb, Vb, a.d, T, a.exp_tab, s_f, a.Tmax, sd, sm,
                               a.out + (size_t) row * a.ld_out + (size_t) h * DK);
}

// ---------------------------------------------------------------------------------------------
// k_attn_tile: un-masked attention for a TILE of AT_R query rows of one head (prompt batches).
// All rows of a prompt batch see the same T = n_past + n positions of the same K/V, so one CTA
// streams K and V once for AT_R queries instead of once per query.  Same arithmetic order as
// k_attn: scores by the transposing butterfly (32 key rows per warp pass, xor 16,8,4,1,2 tree),
// softmax per row with the fp16 exp table and a double sum, V reduction with lane = running sum
// (t % 32) for 4 columns x AT_R queries, all-reduced in the same tree, then the scalar tail.
// grid = (n_head, ceil(n / AT_R)), block = 256, dynamic smem = AT_R * (Tpad + 64) floats. d_kv = 64.
// ---------------------------------------------------------------------------------------------
#define AT_R 16
static __global__ void __launch_bounds__(256) k_attn_tile(AttnArgs a) {
    constexpr int DK = 64, NW = 8;
    extern __shared__ __align__(16) float s_at[];
    const int h = blockIdx.x, r0 = blockIdx.y * AT_R, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int T = a.st->n_past + a.n;                       // mode 0 only
    const int Tpad = (T + 31) & ~31;
    float * sq = s_at;                                      // [AT_R][64]
    float * sc = s_at + AT_R * DK;                          // [AT_R][Tpad]
    const float * Kb = a.kcache + (size_t) h * DK;
    const float * Vb = a.vcache + (size_t) h * DK;
    for (int i = tid; i < AT_R * DK; i += 256) {
        int row = r0 + i / DK; row = row < a.n ? row : a.n - 1;
        sq[i] = a.q[(size_t) row * a.ld_q + (size_t) h * DK + (i % DK)];
    }
    __syncthreads();
    // ---- scores
    for (int tb = warp; tb < T; tb += NW * 32) {
        float k0[32], k1[32];
#pragma unroll
        for (int u = 0; u < 32; u++) {
            const int t = tb + u * NW;
            k0[u] = t < T ? __ldcg(Kb + (size_t) t * a.d + lane) : 0.0f;
            k1[u] = t < T ? __ldcg(Kb + (size_t) t * a.d + 32 + lane) : 0.0f;
        }
        const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4, b0 = lane & 1, b1 = lane & 2;
        const int urow = (lane & 28) | ((lane & 1) << 1) | ((lane & 2) >> 1);
        const int tl = tb + urow * NW;
#pragma unroll 1
        for (int r = 0; r < AT_R; r++) {
            const float q0 = sq[r * DK + lane], q1 = sq[r * DK + 32 + lane];
            float s[32], a16[16], a8[8], a4[4], a2[2];
#pragma unroll
            for (int u = 0; u < 32; u++) s[u] = fmaf(k1[u], q1, fmaf(k0[u], q0, 0.0f));
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
            if (tl < T) sc[r * Tpad + tl] = dot;
        }
    }
    __syncthreads();
    // ---- softmax, one warp per query row
    for (int r = warp; r < AT_R; r += NW) {
        float * row = sc + r * Tpad;
        float mx = -INFINITY;
        for (int t = lane; t < T; t += 32) mx = fmaxf(mx, row[t]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(FULLMASK, mx, o));
        double sum = 0.0;
        for (int t = lane; t < T; t += 32) {
            const float v = bg_h2f(a.exp_tab[bg_f2h(__fsub_rn(row[t], mx))]);
            row[t] = v; sum += (double) v;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(FULLMASK, sum, o);
        const float inv = (float) (1.0 / sum);
        for (int t = lane; t < T; t += 32) row[t] = __fmul_rn(row[t], inv);
    }
    __syncthreads();
    // ---- V: warp = 4 output columns, lane = running sum (t % 32), AT_R queries at once
    const int np = T & ~31, nv = np + ((T - np) & ~3);
    for (int cg = warp; cg < DK / 4; cg += NW) {
        float4 acc[AT_R];
#pragma unroll
        for (int r = 0; r < AT_R; r++) acc[r] = make_float4(0.f, 0.f, 0.f, 0.f);
        const float * vp = Vb + cg * 4;
        for (int s0 = 0; s0 < np; s0 += 64) {
            const int t0 = s0 + lane, t1 = s0 + 32 + lane;
            const float4 v0 = __ldcg((const float4 *) (vp + (size_t) t0 * a.d));
            float4 v1 = make_float4(0.f, 0.f, 0.f, 0.f);
            const bool two = s0 + 32 < np;
            if (two) v1 = __ldcg((const float4 *) (vp + (size_t) t1 * a.d));
#pragma unroll
            for (int r = 0; r < AT_R; r++) {
                const float p0 = sc[r * Tpad + t0];
                acc[r].x = fmaf(v0.x, p0, acc[r].x); acc[r].y = fmaf(v0.y, p0, acc[r].y);
                acc[r].z = fmaf(v0.z, p0, acc[r].z); acc[r].w = fmaf(v0.w, p0, acc[r].w);
            }
            if (two) {
#pragma unroll
                for (int r = 0; r < AT_R; r++) {
                    const float p1 = sc[r * Tpad + t1];
                    acc[r].x = fmaf(v1.x, p1, acc[r].x); acc[r].y = fmaf(v1.y, p1, acc[r].y);
                    acc[r].z = fmaf(v1.z, p1, acc[r].z); acc[r].w = fmaf(v1.w, p1, acc[r].w);
                }
            }
        }
#pragma unroll
        for (int r = 0; r < AT_R; r++) {
#pragma unroll
            for (int k = 0; k < 5; k++) {
                const int o = k == 0 ? 16 : k == 1 ? 8 : k == 2 ? 4 : k == 3 ? 1 : 2;
                acc[r].x = __fadd_rn(acc[r].x, __shfl_xor_sync(FULLMASK, acc[r].x, o));
                acc[r].y = __fadd_rn(acc[r].y, __shfl_xor_sync(FULLMASK, acc[r].y, o));
                acc[r].z = __fadd_rn(acc[r].z, __shfl_xor_sync(FULLMASK, acc[r].z, o));
                acc[r].w = __fadd_rn(acc[r].w, __shfl_xor_sync(FULLMASK, acc[r].w, o));
            }
        }
        for (int t = np; t < T; t++) {                       // scalar tail, identical on every lane
            const float4 v = __ldcg((const float4 *) (vp + (size_t) t * a.d));
            const bool fused = t >= nv;
#pragma unroll
            for (int r = 0; r < AT_R; r++) {
                const float pw = sc[r * Tpad + t];
                if (fused) {
                    acc[r].x = fmaf(v.x, pw, acc[r].x); acc[r].y = fmaf(v.y, pw, acc[r].y);
                    acc[r].z = fmaf(v.z, pw, acc[r].z); acc[r].w = fmaf(v.w, pw, acc[r].w);
                } else {
                    acc[r].x = __fadd_rn(acc[r].x, __fmul_rn(v.x, pw)); acc[r].y = __fadd_rn(acc[r].y, __fmul_rn(v.y, pw));
                    acc[r].z = __fadd_rn(acc[r].z, __fmul_rn(v.z, pw)); acc[r].w = __fadd_rn(acc[r].w, __fmul_rn(v.w, pw));
                }
            }
        }
#pragma unroll
        for (int r = 0; r < AT_R; r++)
            if (lane == r && r0 + r < a.n) *(float4 *) (a.out + (size_t) (r0 + r) * a.ld_out + (size_t) h * DK + cg * 4) = acc[r];
    }
}

// ---------------------------------------------------------------------------------------------
// k_attn_tile2: the same un-masked attention tile with the reference's arithmetic mapped so that no score needs a shuffle.
//   scores: THREAD = key.  The reference's dot of one (query, key) pair is 32 lanes of fma(k[32+i], q[32+i], fma(k[i], q[i], 0))
//           reduced by the tree 16, 8, 4, 1, 2 (ggml_vec_dot_f32 + GGML_F32x8_REDUCE, ggml.c:2372-2407, 1981-1999).  k_attn_tile
//           spreads the 32 lanes over a warp (32 keys in flight, a transposing butterfly: ~190 instructions per 32 scores); here
//           one thread holds its key's 64 values in registers and evaluates the 32 lanes AND the tree itself, as packed f32x2
//           operations (two IEEE operations per instruction, same bits): 16 + 16 FFMA2, 14 FADD2, 3 FADD = 49 instructions per
//           score and thread, the query values broadcast from shared memory.
//   V     : as k_attn_tile (warp = 4 output columns, lane = running sum t % 32) with FFMA2 for the (x, y) / (z, w) halves.
// Same bits as k_attn_tile and k_attn (tests/test_gpu_ops.py::test_attention, the whole-eval tests).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t bg_pack2(float a, float b) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void bg_unpack2(uint64_t v, float & a, float & b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t bg_fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ uint64_t bg_add2(uint64_t a, uint64_t b) { uint64_t d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }

static __global__ void __launch_bounds__(256) k_attn_tile2(AttnArgs a) {
    constexpr int DK = 64, NW = 8;
    extern __shared__ __align__(16) float s_at[];
    const int h = blockIdx.x, r0 = blockIdx.y * AT_R, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int T = a.st->n_past + a.n;                       // mode 0 only
    const int Tpad = (T + 31) & ~31;
    float * sq = s_at;                                      // [AT_R][64]
    float * sc = s_at + AT_R * DK;                          // [AT_R][Tpad]
    const float * Kb = a.kcache + (size_t) h * DK;
    const float * Vb = a.vcache + (size_t) h * DK;
    for (int i = tid; i < AT_R * DK; i += 256) {
        int row = r0 + i / DK; row = row < a.n ? row : a.n - 1;
        sq[i] = a.q[(size_t) row * a.ld_q + (size_t) h * DK + (i % DK)];
    }
    __syncthreads();
    // ---- scores: thread = key
    for (int t = tid; t < T; t += 256) {
        uint64_t k0[16], k1[16];                            // pairs (k[2i], k[2i+1]) and (k[32+2i], k[32+2i+1])
        const float4 * kp = (const float4 *) (Kb + (size_t) t * a.d);
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const float4 f = __ldcg(kp + j), g = __ldcg(kp + 8 + j);
            k0[2 * j] = bg_pack2(f.x, f.y); k0[2 * j + 1] = bg_pack2(f.z, f.w);
            k1[2 * j] = bg_pack2(g.x, g.y); k1[2 * j + 1] = bg_pack2(g.z, g.w);
        }
        const uint64_t zero2 = bg_pack2(0.0f, 0.0f);
#pragma unroll 2
        for (int r = 0; r < AT_R; r++) {
            const float4 * qp = (const float4 *) (sq + r * DK);
            uint64_t v[16];
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const float4 q0 = qp[j], q1 = qp[8 + j];      // broadcast: every thread of the CTA reads the same query
                v[2 * j]     = bg_fma2(k1[2 * j],     bg_pack2(q1.x, q1.y), bg_fma2(k0[2 * j],     bg_pack2(q0.x, q0.y), zero2));
                v[2 * j + 1] = bg_fma2(k1[2 * j + 1], bg_pack2(q1.z, q1.w), bg_fma2(k0[2 * j + 1], bg_pack2(q0.z, q0.w), zero2));
            }
            uint64_t a8[8], a4[4], a2[2];
#pragma unroll
            for (int i = 0; i < 8; i++) a8[i] = bg_add2(v[i], v[i + 8]);          // lanes i, i + 16
#pragma unroll
            for (int i = 0; i < 4; i++) a4[i] = bg_add2(a8[i], a8[i + 4]);        // + 8
#pragma unroll
            for (int i = 0; i < 2; i++) a2[i] = bg_add2(a4[i], a4[i + 2]);        // + 4: (c0, c1), (c2, c3)
            float c0, c1, c2, c3;
            bg_unpack2(a2[0], c0, c1); bg_unpack2(a2[1], c2, c3);
            sc[r * Tpad + t] = __fadd_rn(__fadd_rn(c0, c1), __fadd_rn(c2, c3));   // xor 1, then xor 2
        }
    }
    __syncthreads();
    // ---- softmax, one warp per query row
    for (int r = warp; r < AT_R; r += NW) {
        float * row = sc + r * Tpad;
        float mx = -INFINITY;
        for (int t = lane; t < T; t += 32) mx = fmaxf(mx, row[t]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(FULLMASK, mx, o));
        double sum = 0.0;
        for (int t = lane; t < T; t += 32) {
            const float v = bg_h2f(a.exp_tab[bg_f2h(__fsub_rn(row[t], mx))]);
            row[t] = v; sum += (double) v;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(FULLMASK, sum, o);
        const float inv = (float) (1.0 / sum);
        for (int t = lane; t < T; t += 32) row[t] = __fmul_rn(row[t], inv);
    }
    __syncthreads();
    // ---- V: warp = 4 output columns, lane = running sum (t % 32), AT_R queries at once
    const int np = T & ~31, nv = np + ((T - np) & ~3);
    for (int cg = warp; cg < DK / 4; cg += NW) {
        uint64_t axy[AT_R], azw[AT_R];
#pragma unroll
        for (int r = 0; r < AT_R; r++) { axy[r] = bg_pack2(0.f, 0.f); azw[r] = bg_pack2(0.f, 0.f); }
        const float * vp = Vb + cg * 4;
        for (int s0 = 0; s0 < np; s0 += 64) {
            const int t0 = s0 + lane, t1 = s0 + 32 + lane;
            const float4 v0 = __ldcg((const float4 *) (vp + (size_t) t0 * a.d));
            float4 v1 = make_float4(0.f, 0.f, 0.f, 0.f);
            const bool two = s0 + 32 < np;
            if (two) v1 = __ldcg((const float4 *) (vp + (size_t) t1 * a.d));
            const uint64_t v0xy = bg_pack2(v0.x, v0.y), v0zw = bg_pack2(v0.z, v0.w);
#pragma unroll
            for (int r = 0; r < AT_R; r++) {
                const float p0 = sc[r * Tpad + t0];
                const uint64_t pp = bg_pack2(p0, p0);
                axy[r] = bg_fma2(v0xy, pp, axy[r]); azw[r] = bg_fma2(v0zw, pp, azw[r]);
            }
            if (two) {
                const uint64_t v1xy = bg_pack2(v1.x, v1.y), v1zw = bg_pack2(v1.z, v1.w);
#pragma unroll
                for (int r = 0; r < AT_R; r++) {
                    const float p1 = sc[r * Tpad + t1];
                    const uint64_t pp = bg_pack2(p1, p1);
                    axy[r] = bg_fma2(v1xy, pp, axy[r]); azw[r] = bg_fma2(v1zw, pp, azw[r]);
                }
            }
        }
        float4 acc[AT_R];
#pragma unroll
        for (int r = 0; r < AT_R; r++) { bg_unpack2(axy[r], acc[r].x, acc[r].y); bg_unpack2(azw[r], acc[r].z, acc[r].w); }
#pragma unroll
        for (int r = 0; r < AT_R; r++) {
#pragma unroll
            for (int k = 0; k < 5; k++) {
                const int o = k == 0 ? 16 : k == 1 ? 8 : k == 2 ? 4 : k == 3 ? 1 : 2;
                acc[r].x = __fadd_rn(acc[r].x, __shfl_xor_sync(FULLMASK, acc[r].x, o));
                acc[r].y = __fadd_rn(acc[r].y, __shfl_xor_sync(FULLMASK, acc[r].y, o));
                acc[r].z = __fadd_rn(acc[r].z, __shfl_xor_sync(FULLMASK, acc[r].z, o));
                acc[r].w = __fadd_rn(acc[r].w, __shfl_xor_sync(FULLMASK, acc[r].w, o));
            }
        }
        for (int t = np; t < T; t++) {                       // scalar tail, identical on every lane
            const float4 v = __ldcg((const float4 *) (vp + (size_t) t * a.d));
            const bool fused = t >= nv;
#pragma unroll
            for (int r = 0; r < AT_R; r++) {
                const float pw = sc[r * Tpad + t];
                if (fused) {
                    acc[r].x = fmaf(v.x, pw, acc[r].x); acc[r].y = fmaf(v.y, pw, acc[r].y);
                    acc[r].z = fmaf(v.z, pw, acc[r].z); acc[r].w = fmaf(v.w, pw, acc[r].w);
                } else {
                    acc[r].x = __fadd_rn(acc[r].x, __fmul_rn(v.x, pw)); acc[r].y = __fadd_rn(acc[r].y, __fmul_rn(v.y, pw));
                    acc[r].z = __fadd_rn(acc[r].z, __fmul_rn(v.z, pw)); acc[r].w = __fadd_rn(acc[r].w, __fmul_rn(v.w, pw));
                }
            }
        }
#pragma unroll
        for (int r = 0; r < AT_R; r++)
            if (lane == r && r0 + r < a.n) *(float4 *) (a.out + (size_t) (r0 + r) * a.ld_out + (size_t) h * DK + cg * 4) = acc[r];
    }
}

// fp16-table GELU as a stand-alone op (unit tests); the eval fuses it into the fc1 epilogue
static __global__ void k_gelu(const float * __restrict__ x, float * __restrict__ y, int n, const uint16_t * __restrict__ tab) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] = bg_h2f(tab[bg_f2h(x[i])]);
}

// convert an activation record back to the reference's block_q8_0 / block_q8_1 / fp16 bytes
static __global__ void k_act_export(const uint8_t * __restrict__ rec, int wtype, int K, int off_d, int off_s, uint8_t * __restrict__ out) {
    const int kind = bg_act_kind(wtype);
    const int nb = K >> 5;
    if (kind == ACT_F16) {
        const float * f = (const float *) rec;
        for (int c = threadIdx.x; c < K; c += blockDim.x) {
            const int gg = c / 256, e = (c % 256) / 32, lane = c % 32;
            ((uint16_t *) out)[c] = bg_f2h(f[(((gg * 2 + (e >> 2)) * 32 + lane) * 4) + (e & 3)]);
        }
        return;
    }
    if (kind == ACT_F32) {
        const float * f = (const float *) rec;
        for (int c = threadIdx.x; c < K; c += blockDim.x) {
            const int gg = c / 128, e = (c % 128) / 32, lane = c % 32;
            ((float *) out)[c] = f[(gg * 32 + lane) * 4 + e];
        }
        return;
    }
    const uint32_t * aq = (const uint32_t *) rec;
    const float * ad = (const float *) (rec + off_d), * as = (const float *) (rec + off_s);
    for (int b = threadIdx.x; b < nb; b += blockDim.x) {
        const int g = b >> 2, i = b & 3;
        if (kind == ACT_Q8_0) {          // block_q8_0: fp16 d, 32 x int8 (2-byte aligned)
            uint16_t * o16 = (uint16_t *) (out + (size_t) b * 34);
            o16[0] = bg_f2h(ad[b]);
            for (int l = 0; l < 8; l++) {
                const uint32_t w = aq[(g * 8 + l) * 4 + i];
                o16[1 + 2 * l] = (uint16_t) (w & 0xFFFFu);
                o16[2 + 2 * l] = (uint16_t) (w >> 16);
            }
        } else {                          // block_q8_1: f32 d, f32 s, 32 x int8 (8-byte aligned)
            uint32_t * o32 = (uint32_t *) (out + (size_t) b * 40);
            o32[0] = __float_as_uint(ad[b]);
            o32[1] = __float_as_uint(as[b]);
            for (int l = 0; l < 8; l++) o32[2 + l] = aq[(g * 8 + l) * 4 + i];
        }
    }
}

// greedy sampling on the device: argmax (first index wins), log the id, feed it back and
// advance n_past -- lets a whole decode loop run without the host (bgpt_cuda_decode_greedy)
static __global__ void __launch_bounds__(1024) k_argmax_advance(const float * __restrict__ logits, int n_vocab,
                                                         int * __restrict__ next_tok, int * __restrict__ id_log,
                                                         DevState * st, int n_advance) {
    __shared__ float sv[32]; __shared__ int si[32];
    float best = -INFINITY; int bi = 0x7fffffff;
    for (int i = threadIdx.x; i < n_vocab; i += blockDim.x) {
        const float v = logits[i];
        if (v > best || (v == best && i < bi)) { best = v; bi = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(FULLMASK, best, o); const int oi = __shfl_xor_sync(FULLMASK, bi, o);
        if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = best; si[threadIdx.x >> 5] = bi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int) (blockDim.x >> 5); w++)
            if (sv[w] > best || (sv[w] == best && si[w] < bi)) { best = sv[w]; bi = si[w]; }
        if (bi == 0x7fffffff) bi = 0;
        next_tok[0] = bi;
        id_log[st->step] = bi;
        st->step += 1;
        st->n_past += n_advance;
    }
}


// the same for S lock-step streams: block s takes the argmax of logit row s (first index wins), feeds it back as stream s's next
// token and logs it at id_log[step * S + s]; k_streams_advance then moves the step counter and n_past (one tiny launch: the S
// blocks of k_argmax_rows read st->step, so it must not change under them)
static __global__ void __launch_bounds__(1024) k_argmax_rows(const float * __restrict__ logits, int n_vocab, int * __restrict__ next_tok,
                                                             int * __restrict__ id_log, const DevState * __restrict__ st, int n_streams) {
    __shared__ float sv[32]; __shared__ int si[32];
    const float * row = logits + (size_t) blockIdx.x * n_vocab;
    float best = -INFINITY; int bi = 0x7fffffff;
    for (int i = threadIdx.x; i < n_vocab; i += blockDim.x) {
        const float v = row[i];
        if (v > best || (v == best && i < bi)) { best = v; bi = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(FULLMASK, best, o); const int oi = __shfl_xor_sync(FULLMASK, bi, o);
        if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = best; si[threadIdx.x >> 5] = bi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int) (blockDim.x >> 5); w++)
            if (sv[w] > best || (sv[w] == best && si[w] < bi)) { best = sv[w]; bi = si[w]; }
        if (bi == 0x7fffffff) bi = 0;
        next_tok[blockIdx.x] = bi;
        id_log[(size_t) st->step * n_streams + blockIdx.x] = bi;
    }
}
static __global__ void k_streams_advance(DevState * st) { st->step += 1; st->n_past += 1; }
