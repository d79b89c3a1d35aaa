// bgpt_quant.cuh -- f32 -> Q4_0 / Q4_1 / Q5_0 / Q5_1 / Q8_0 weight quantiser on the device.
//
// Bit-identical to the reference's `quantize_row_q*_reference` (ggml.c:892-1094) as the `quantize` tool runs them
// (ggml_quantize_q*_0/1 -> *_reference; examples/quantize, biogpt.cpp:459-621), including what its build does to
// the arithmetic: `x*id + 8.5f` / `(x-min)*id + 0.5f` are contracted to one fma (gcc -O3 -mfma, -ffp-contract=fast),
// `d` and `m` are stored as fp16 (round to nearest even) and the scale used for the codes is the UNROUNDED f32 `d`.
// Output is the file (AoS) block layout: {fp16 d, [fp16 m], [u32 qh], 16 nibble bytes} / {fp16 d, 32 int8}.
//
// HBM-bound: 128 B read + 18..34 B written per 32 weights.  A warp takes 32 consecutive blocks: coalesced 16-byte loads,
// a padded shared-memory transpose so that one thread owns one block (the reference's scalar loop order, trivially),
// and the 32 output blocks leave as coalesced 32-bit words.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "bgpt_layout.h"

#define BQ_THREADS 256

__device__ __forceinline__ uint16_t bq_f2h(float f) { return __half_as_ushort(__float2half_rn(f)); }

template <int FMT>
__global__ void __launch_bounds__(BQ_THREADS) k_quantize_weights(const float * __restrict__ x, long long nblocks, uint8_t * __restrict__ out) {
    constexpr int BS = FMT == BG_Q4_0 ? 18 : FMT == BG_Q4_1 ? 20 : FMT == BG_Q5_0 ? 22 : FMT == BG_Q5_1 ? 24 : 34;
    __shared__ __align__(16) uint8_t stage[BQ_THREADS / 32][32 * BS];
    __shared__ float tile[BQ_THREADS / 32][32 * 33];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long wstride = (long long) gridDim.x * (BQ_THREADS / 32);
    for (long long wb = (long long) blockIdx.x * (BQ_THREADS / 32) + warp; wb * 32 < nblocks; wb += wstride) {
        const long long b = wb * 32 + lane;
        // the warp's 32 blocks are 4 KB of consecutive floats: 8 fully coalesced 16-byte loads per lane, transposed through
        // shared memory (row stride 33 words: conflict-free both ways) so that every lane ends up with one whole block
        const long long left_f = (nblocks - wb * 32) * 32;        // floats of this warp's span that exist
        const float4 * src = (const float4 *) (x + wb * 1024);
        float4 t[8];
#pragma unroll
        for (int k = 0; k < 8; k++) t[k] = (long long) (k * 32 + lane) * 4 < left_f ? __ldcs(src + k * 32 + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
        float * tin = tile[warp];
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const int e = (k * 32 + lane) * 4, blk = e >> 5, j = e & 31;
            float * dstp = tin + blk * 33 + j;
            dstp[0] = t[k].x; dstp[1] = t[k].y; dstp[2] = t[k].z; dstp[3] = t[k].w;
        }
        __syncwarp();
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; j++) v[j] = tin[lane * 33 + j];
        (void) b;
        uint8_t * o = stage[warp] + lane * BS;
        uint16_t * o16 = (uint16_t *) o;                           // BS is even: every block starts 2-byte aligned
        uint32_t q[32];
        if (FMT == BG_Q4_0 || FMT == BG_Q5_0) {
            // ggml.c:896-907 / 976-987: the value of largest magnitude, sign kept, first occurrence wins
            float amax = 0.f, mx = 0.f;
#pragma unroll
            for (int j = 0; j < 32; j++) if (amax < fabsf(v[j])) { amax = fabsf(v[j]); mx = v[j]; }
            const float d = __fdiv_rn(mx, FMT == BG_Q4_0 ? -8.0f : -16.0f);
            const float id = d != 0.f ? __fdiv_rn(1.0f, d) : 0.f;
            const float off = FMT == BG_Q4_0 ? 8.5f : 16.5f;
            const int top = FMT == BG_Q4_0 ? 15 : 31;
#pragma unroll
            for (int j = 0; j < 32; j++) { const int t = (int) (int8_t) (int) fmaf(v[j], id, off); q[j] = (uint32_t) min(top, t) & 0xFFu; }
            o16[0] = bq_f2h(d);
        } else if (FMT == BG_Q4_1 || FMT == BG_Q5_1) {
            // ggml.c:935-952 / 1022-1039
            float mn = v[0], mx = v[0];
#pragma unroll
            for (int j = 1; j < 32; j++) { mn = fminf(mn, v[j]); mx = fmaxf(mx, v[j]); }
            const float d = __fdiv_rn(__fsub_rn(mx, mn), FMT == BG_Q4_1 ? 15.0f : 31.0f);
            const float id = d != 0.f ? __fdiv_rn(1.0f, d) : 0.f;
#pragma unroll
            for (int j = 0; j < 32; j++) {
                const float t = fmaf(__fsub_rn(v[j], mn), id, 0.5f);
                q[j] = FMT == BG_Q4_1 ? ((uint32_t) min(15, (int) (int8_t) (int) t) & 0xFFu) : ((uint32_t) (int) t & 0xFFu);
            }
            o16[0] = bq_f2h(d); o16[1] = bq_f2h(mn);
        } else {
            // ggml.c:1062-1080: d = amax / 127, codes = roundf(x * id) (half away from zero)
            float amax = 0.f;
#pragma unroll
            for (int j = 0; j < 32; j++) amax = fmaxf(amax, fabsf(v[j]));
            const float d = __fdiv_rn(amax, 127.0f);
            const float id = d != 0.f ? __fdiv_rn(1.0f, d) : 0.f;
#pragma unroll
            for (int j = 0; j < 32; j++) q[j] = (uint32_t) (int) roundf(__fmul_rn(v[j], id)) & 0xFFu;
            o16[0] = bq_f2h(d);
        }
        if (FMT == BG_Q8_0) {
#pragma unroll
            for (int j = 0; j < 16; j++) o16[1 + j] = (uint16_t) (q[2 * j] | (q[2 * j + 1] << 8));
        } else {
            constexpr int QOFF = FMT == BG_Q4_0 ? 1 : FMT == BG_Q4_1 ? 2 : FMT == BG_Q5_0 ? 3 : 4;    // in 16-bit units
            if (FMT == BG_Q5_0 || FMT == BG_Q5_1) {
                uint32_t qh = 0;
#pragma unroll
                for (int j = 0; j < 32; j++) qh |= ((q[j] >> 4) & 1u) << j;           // element j -> bit j (ggml.c:1000-1001)
                o16[QOFF - 2] = (uint16_t) (qh & 0xFFFFu); o16[QOFF - 1] = (uint16_t) (qh >> 16);
            }
#pragma unroll
            for (int j = 0; j < 8; j++) {                                              // byte j = elem j | elem j+16 << 4
                const uint32_t b0 = (q[2 * j] & 0xFu) | ((q[2 * j + 16] & 0xFu) << 4);
                const uint32_t b1 = (q[2 * j + 1] & 0xFu) | ((q[2 * j + 17] & 0xFu) << 4);
                o16[QOFF + j] = (uint16_t) (b0 | (b1 << 8));
            }
        }
        __syncwarp();
        // 32 blocks = 32*BS bytes, a multiple of 64: coalesced 32-bit words (the last warp of the tensor may be short)
        const long long left = nblocks - wb * 32;
        const int words = (int) ((left < 32 ? left : 32) * BS / 4);
        const int tail = (int) ((left < 32 ? left : 32) * BS) - words * 4;
        uint32_t * dst = (uint32_t *) (out + wb * 32 * BS);
        const uint32_t * sw = (const uint32_t *) stage[warp];
        for (int i = lane; i < words; i += 32) dst[i] = sw[i];
        if (lane < tail) out[wb * 32 * BS + words * 4 + lane] = stage[warp][words * 4 + lane];
        __syncwarp();
    }
}
