// bgpt_layout.h -- device layouts of weight rows and activation records (host + device).
//
// The `.bin` file stores a matrix row as K/32 consecutive AoS blocks (block_q4_0 ... block_q8_0,
// ggml.c:844-889) of 18/20/22/24/34 bytes -- none of which is a multiple of 16, so no block
// but the first of a row can be fetched with an aligned 128-bit load.  On upload every row is
// re-tiled into planes (quants | high bits | scales | mins).  The bytes are the same bytes,
// permuted inside the row; the row stride equals the file's row size for every K that is a
// multiple of 256 (all BioGPT shapes), so HBM traffic per row is the file-format traffic.
//
// The permutation is chosen for the exact-order dot product (see bgpt_cuda.cu, "lane order"):
// the reference's AVX2 kernels keep 8 running float sums per row, sum l accumulating the
// elements 4l..4l+3 of every block in block order.  Here 4 threads share a row; thread j owns
// the two sums j and j+4.  One aligned uint4 load hands a thread the 4-byte element groups it
// owns for 4 consecutive blocks:
//
//   Q4_x/Q5_x quants : [g = blk/4][j = 0..3][i = blk%4] 32-bit words; word = qs[4j..4j+3] of the
//                      block (low nibbles -> sum j, high nibbles -> sum j+4)
//   Q5_x high bits   : [g][j][i] bytes; low nibble = qh bits 4j..4j+3, high nibble = bits
//                      16+4j..16+4j+3
//   Q8_0 quants      : [g][c = 0..1][j][i] words; word = qs[16c+4j .. 16c+4j+3] (c=0 -> sum j,
//                      c=1 -> sum j+4)
//   scales / mins    : [blk] fp16
//   F16 / F32        : the reference keeps 32 running sums (element i -> sum i%32); one warp owns
//                      a row, lane = sum.  [gg][lane][e] with element = gg*32*E + e*32 + lane,
//                      E = 8 (F16) or 4 (F32) so that a lane's load is one uint4.
//
// Rows whose block count is not a multiple of 4 (only synthetic test shapes) are padded with
// zero blocks: fma(0, 0, acc) leaves every running sum unchanged.
#pragma once
#include <stdint.h>
#include <stddef.h>
#include <string.h>

#ifdef __CUDACC__
#define BG_HD __host__ __device__
#else
#define BG_HD
#endif

enum {
    BG_F32 = 0, BG_F16 = 1, BG_Q4_0 = 2, BG_Q4_1 = 3, BG_Q5_0 = 6, BG_Q5_1 = 7, BG_Q8_0 = 8,
};

// activation record kinds (what mul_mat converts src1 to, ggml.c:11909-11925)
enum { ACT_F32 = 0, ACT_F16 = 1, ACT_Q8_0 = 2, ACT_Q8_1 = 3 };

BG_HD static inline int bg_is_quant(int t) { return t == BG_Q4_0 || t == BG_Q4_1 || t == BG_Q5_0 || t == BG_Q5_1 || t == BG_Q8_0; }
BG_HD static inline int bg_type_ok(int t) { return t == BG_F32 || t == BG_F16 || bg_is_quant(t); }
BG_HD static inline int bg_file_block_bytes(int t) {
    switch (t) { case BG_F32: return 4; case BG_F16: return 2; case BG_Q4_0: return 18; case BG_Q4_1: return 20;
                 case BG_Q5_0: return 22; case BG_Q5_1: return 24; case BG_Q8_0: return 34; }
    return 0;
}
BG_HD static inline int bg_file_block_elems(int t) { return bg_is_quant(t) ? 32 : 1; }
BG_HD static inline size_t bg_file_row_bytes(int t, int K) { return (size_t) K / bg_file_block_elems(t) * bg_file_block_bytes(t); }
BG_HD static inline int bg_act_kind(int t) {
    switch (t) { case BG_F32: return ACT_F32; case BG_F16: return ACT_F16;
                 case BG_Q4_0: case BG_Q5_0: case BG_Q8_0: return ACT_Q8_0;
                 case BG_Q4_1: case BG_Q5_1: return ACT_Q8_1; }
    return -1;
}
// integer offset folded into the activation side for the symmetric formats: sum((q-off)*a) =
// dp4a(q, a) - off*sum(a)
BG_HD static inline int bg_code_offset(int t) { return t == BG_Q4_0 ? 8 : (t == BG_Q5_0 ? 16 : 0); }

struct RowLayout {
    int type, K;
    int nb;       // 32-element blocks in the file row
    int G;        // groups: of 4 blocks (quantised), of 256 (F16) / 128 (F32) elements
    int off_qh;   // byte offsets of the planes inside a device row (-1: absent)
    int off_d;
    int off_m;
    int stride;   // device row stride in bytes (multiple of 16)
};

static inline RowLayout bg_row_layout(int type, int K) {
    RowLayout L; L.type = type; L.K = K; L.nb = K / 32; L.off_qh = L.off_d = L.off_m = -1;
    if (type == BG_F16) { L.G = (K + 255) / 256; L.stride = L.G * 512; return L; }
    if (type == BG_F32) { L.G = (K + 127) / 128; L.stride = L.G * 512; return L; }
    L.G = (L.nb + 3) / 4;
    int o = L.G * (type == BG_Q8_0 ? 128 : 64);
    if (type == BG_Q5_0 || type == BG_Q5_1) { L.off_qh = o; o += L.G * 16; }
    L.off_d = o; o += L.G * 8;
    if (type == BG_Q4_1 || type == BG_Q5_1) { L.off_m = o; o += L.G * 8; }
    L.stride = (o + 15) & ~15;
    return L;
}

// Activation record of one token row for a matmul with reduction length K.
//   quantised: words aq[g][l=0..7][i] (bytes 4l..4l+3 of block 4g+i), offsets an[g][l][i]
//              (= -code_offset * sum of those 4 bytes), scales ad[blk] (f32), as[blk] (f32, Q8_1)
//   F16      : f32 values (already rounded through fp16) as float4 [gg][eh=0..1][lane]
//   F32      : f32 values as float4 [gg][lane]
struct ActLayout {
    int kind, K, nb, G;
    int off_n, off_d, off_s;
    int bytes;    // per token row, multiple of 16
};
static inline ActLayout bg_act_layout(int wtype, int K) {
    ActLayout A; A.kind = bg_act_kind(wtype); A.K = K; A.nb = K / 32; A.off_n = A.off_d = A.off_s = 0;
    if (A.kind == ACT_F16) { A.G = (K + 255) / 256; A.bytes = A.G * 1024; return A; }
    if (A.kind == ACT_F32) { A.G = (K + 127) / 128; A.bytes = A.G * 512; return A; }
    A.G = (A.nb + 3) / 4;
    A.off_n = A.G * 128; A.off_d = A.G * 256; A.off_s = A.G * 272; A.bytes = A.G * 288;
    return A;
}

#ifndef __CUDACC_RTC__
// host: file row (AoS blocks) -> device row.  `dst` must be zero-initialised (padding).
static inline void bg_repack_row(const RowLayout & L, const uint8_t * src, uint8_t * dst) {
    const int t = L.type;
    if (t == BG_F16) {
        const uint16_t * s = (const uint16_t *) src; uint16_t * d = (uint16_t *) dst;
        for (int i = 0; i < L.K; i++) { const int gg = i / 256, e = (i % 256) / 32, lane = i % 32; d[(gg * 32 + lane) * 8 + e] = s[i]; }
        return;
    }
    if (t == BG_F32) {
        const uint32_t * s = (const uint32_t *) src; uint32_t * d = (uint32_t *) dst;
        for (int i = 0; i < L.K; i++) { const int gg = i / 128, e = (i % 128) / 32, lane = i % 32; d[(gg * 32 + lane) * 4 + e] = s[i]; }
        return;
    }
    const int bs = bg_file_block_bytes(t);
    for (int b = 0; b < L.nb; b++) {
        const uint8_t * blk = src + (size_t) b * bs;
        const int g = b / 4, i = b % 4;
        int o = 0;
        memcpy(dst + L.off_d + b * 2, blk + o, 2); o += 2;
        if (L.off_m >= 0) { memcpy(dst + L.off_m + b * 2, blk + o, 2); o += 2; }
        if (L.off_qh >= 0) {
            uint32_t qh; memcpy(&qh, blk + o, 4); o += 4;
            for (int j = 0; j < 4; j++)
                dst[L.off_qh + g * 16 + j * 4 + i] = (uint8_t) (((qh >> (4 * j)) & 0xF) | (((qh >> (16 + 4 * j)) & 0xF) << 4));
        }
        const uint8_t * qs = blk + o;
        if (t == BG_Q8_0) {
            for (int c = 0; c < 2; c++) for (int j = 0; j < 4; j++)
                memcpy(dst + ((g * 2 + c) * 4 + j) * 16 + i * 4, qs + 16 * c + 4 * j, 4);
        } else {
            for (int j = 0; j < 4; j++) memcpy(dst + (g * 4 + j) * 16 + i * 4, qs + 4 * j, 4);
        }
    }
}
#endif
