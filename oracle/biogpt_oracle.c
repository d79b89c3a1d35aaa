/* oracle/biogpt_oracle.c -- TEST INFRASTRUCTURE ONLY (see biogpt_oracle.h).
 *
 * Plain-C restatement of the reference's CPU forward pass, op by op, as the reference
 * executes it in its AVX2 build (the build oracle/Makefile produces and the one its own
 * CMake picks on any AVX2 host).  Scalar code; every fused multiply-add of the reference
 * binary is an explicit fmaf() here and the file is compiled with -ffp-contract=off, so the
 * rounding sequence is spelled out rather than left to the compiler.
 *
 * "Lane" below always means one of the 8 float lanes of an AVX register: the reference
 * keeps 8 (quantised dots) or 4x8 (f16/f32 dots) running sums and only adds them together at
 * the end, in a fixed tree.  Restating that tree is what makes this oracle bit-identical to
 * the reference instead of merely close.
 *
 * Reference map (all under /root/reference):
 *   file parsing                 biogpt.cpp:27-453
 *   graph wiring                 biogpt.cpp:624-810
 *   activation quantisers        ggml/src/ggml.c:1166-1249 (q8_0), 1403-1494 (q8_1), 493-510 (f16)
 *   weight dequantisers          ggml/src/ggml.c:1536-1646
 *   dot products                 ggml/src/ggml.c:2372-2443 (f32,f16), 2518-2541 (q4_0), 2824-2857 (q4_1),
 *                                3071-3093 (q5_0), 3386-3411 (q5_1), 3597-3618 (q8_0); helpers 611-690
 *   mul_mat driver               ggml/src/ggml.c:11804-12013
 *   norm / soft_max / gelu       ggml/src/ggml.c:11377-11426 / 12914-12983 / 3842-3861, tables 4620-4640
 *   get_rows / scale / add       ggml/src/ggml.c:12524-12628 / 12300-12341 / 9286-9357
 */
#include "biogpt_oracle.h"

#include <immintrin.h>
#include <math.h>
#include <pthread.h>
#include <unistd.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

enum { T_F32 = 0, T_F16 = 1, T_Q4_0 = 2, T_Q4_1 = 3, T_Q5_0 = 6, T_Q5_1 = 7, T_Q8_0 = 8, T_Q8_1 = 9 };
#define QK 32
#define NORM_EPS 1e-5f /* biogpt.cpp:24 */

/* ---- fp16 <-> fp32 (IEEE, round-to-nearest-even; ggml uses F16C for both directions) ---- */
static inline float    h2f(uint16_t h) { return _cvtsh_ss(h); }
static inline uint16_t f2h(float f)    { return _cvtss_sh(f, 0); }

static size_t type_size(int t) {
    switch (t) {
        case T_F32: return 4;  case T_F16: return 2;
        case T_Q4_0: return 18; case T_Q4_1: return 20; case T_Q5_0: return 22;
        case T_Q5_1: return 24; case T_Q8_0: return 34; case T_Q8_1: return 40;
    }
    return 0;
}
static int blck_size(int t) { return (t == T_F32 || t == T_F16) ? 1 : QK; }
static size_t row_size(int t, int k) { return (size_t) k / blck_size(t) * type_size(t); }
/* ggml.c type_traits[].vec_dot_type */
static int vec_dot_type(int t) {
    switch (t) {
        case T_F32: return T_F32; case T_F16: return T_F16;
        case T_Q4_0: case T_Q5_0: case T_Q8_0: return T_Q8_0;
        case T_Q4_1: case T_Q5_1: return T_Q8_1;
    }
    return -1;
}
/* ggml_ftype -> ggml_type, ggml.c:4500-4523 */
static int ftype_to_type(int f) {
    switch (f) { case 0: return T_F32; case 1: return T_F16; case 2: return T_Q4_0; case 3: return T_Q4_1;
                 case 7: return T_Q8_0; case 8: return T_Q5_0; case 9: return T_Q5_1; }
    return -1;
}

/* ---- lookup tables, ggml.c:4620-4640 ------------------------------------------------------ */
static uint16_t g_gelu[65536];
static uint16_t g_exp[65536];
static int g_tables_ready = 0;

/* ggml_gelu_f32 (ggml.c:3842-3844) as gcc -O3 -mfma compiles it: the inner `1 + A*x*x` is
 * contracted into one fma, everything else is separate roundings (checked against all 65536
 * entries of the reference's table in tests/test_oracle_vs_ref.py). */
static float gelu_f32(float x) {
    const float A = 0.044715f, S = 0.79788456080286535587989211986876f;
    const float inner = fmaf(A * x, x, 1.0f);
    return (0.5f * x) * (1.0f + tanhf((S * x) * inner));
}
static void init_tables(void) {
    if (g_tables_ready) return;
    for (int i = 0; i < 65536; i++) {
        const float f = h2f((uint16_t) i);
        g_gelu[i] = f2h(gelu_f32(f));
        g_exp[i]  = f2h(expf(f));
    }
    g_tables_ready = 1;
}
void bo_tables(uint16_t * gelu, uint16_t * ex) {
    init_tables();
    memcpy(gelu, g_gelu, sizeof g_gelu);
    memcpy(ex, g_exp, sizeof g_exp);
}

/* ---- activation quantisers (what mul_mat applies to src1) -------------------------------- */
/* quantize_row_q8_0, AVX branch ggml.c:1166-1203: d = amax/127 stored as fp16, multiply by
 * id = 127/amax, round to nearest even. */
void bo_quantize_row_q8_0(const float * x, void * vy, int k) {
    uint8_t * y = (uint8_t *) vy;
    for (int i = 0; i < k / QK; i++, x += QK, y += 34) {
        float amax = 0.0f;
        for (int j = 0; j < QK; j++) { const float a = fabsf(x[j]); if (a > amax) amax = a; }
        const float d  = amax / 127.f;
        const float id = (amax != 0.0f) ? 127.f / amax : 0.0f;
        const uint16_t dh = f2h(d);
        memcpy(y, &dh, 2);
        for (int j = 0; j < QK; j++) y[2 + j] = (uint8_t) (int8_t) (int) nearbyintf(x[j] * id);
    }
}
/* quantize_row_q8_1, AVX2 branch ggml.c:1403-1450: d kept as f32, s = d * sum(q). */
void bo_quantize_row_q8_1(const float * x, void * vy, int k) {
    uint8_t * y = (uint8_t *) vy;
    for (int i = 0; i < k / QK; i++, x += QK, y += 40) {
        float amax = 0.0f;
        for (int j = 0; j < QK; j++) { const float a = fabsf(x[j]); if (a > amax) amax = a; }
        const float d  = amax / 127.f;
        const float id = (amax != 0.0f) ? 127.f / amax : 0.0f;
        int sum = 0;
        for (int j = 0; j < QK; j++) {
            const int q = (int) nearbyintf(x[j] * id);
            y[8 + j] = (uint8_t) (int8_t) q;
            sum += q;
        }
        const float s = d * (float) sum;
        memcpy(y, &d, 4);
        memcpy(y + 4, &s, 4);
    }
}
/* ggml_fp32_to_fp16_row, ggml.c:493-510 */
void bo_fp32_to_fp16_row(const float * x, uint16_t * y, int k) {
    for (int i = 0; i < k; i++) y[i] = f2h(x[i]);
}

/* ---- weight block unpacking: 32 integer codes of one block, element order -------------- */
/* returns codes q[0..31] (already offset for the symmetric formats), scale d and min m */
static inline void unpack_block(int t, const uint8_t * b, int * q, float * d, float * m) {
    uint16_t h; memcpy(&h, b, 2); *d = h2f(h); *m = 0.0f;
    const uint8_t * qs; uint32_t qh = 0;
    switch (t) {
        case T_Q4_0: qs = b + 2;
            for (int j = 0; j < 16; j++) { q[j] = (qs[j] & 0x0F) - 8; q[j + 16] = (qs[j] >> 4) - 8; } break;
        case T_Q4_1: memcpy(&h, b + 2, 2); *m = h2f(h); qs = b + 4;
            for (int j = 0; j < 16; j++) { q[j] = (qs[j] & 0x0F); q[j + 16] = (qs[j] >> 4); } break;
        case T_Q5_0: memcpy(&qh, b + 2, 4); qs = b + 6;
            for (int j = 0; j < 16; j++) {
                q[j]      = ((qs[j] & 0x0F) | (((qh >> j) & 1) << 4)) - 16;
                q[j + 16] = ((qs[j] >> 4)   | (((qh >> (j + 16)) & 1) << 4)) - 16;
            } break;
        case T_Q5_1: memcpy(&h, b + 2, 2); *m = h2f(h); memcpy(&qh, b + 4, 4); qs = b + 8;
            for (int j = 0; j < 16; j++) {
                q[j]      = (qs[j] & 0x0F) | (((qh >> j) & 1) << 4);
                q[j + 16] = (qs[j] >> 4)   | (((qh >> (j + 16)) & 1) << 4);
            } break;
        case T_Q8_0:
            for (int j = 0; j < 32; j++) q[j] = (int8_t) b[2 + j];
            break;
    }
}

/* dequantize_row_q*, ggml.c:1536-1646 (x*d is exact in f32: <=6-bit code x 11-bit scale) */
void bo_dequantize_row(int t, const void * vx, float * y, int k) {
    const uint8_t * x = (const uint8_t *) vx;
    if (t == T_F32) { memcpy(y, x, (size_t) k * 4); return; }
    if (t == T_F16) { const uint16_t * hx = (const uint16_t *) x; for (int i = 0; i < k; i++) y[i] = h2f(hx[i]); return; }
    const size_t bs = type_size(t);
    for (int i = 0; i < k / QK; i++) {
        int q[32]; float d, m;
        unpack_block(t, x + i * bs, q, &d, &m);
        const int has_m = (t == T_Q4_1 || t == T_Q5_1);
        for (int j = 0; j < 32; j++) y[i * QK + j] = has_m ? (float) q[j] * d + m : (float) q[j] * d;
    }
}

/* ---- dot products ---------------------------------------------------------------------------- */
/* hsum_float_8, ggml.c:611-617: ((a0+a4)+(a2+a6)) + ((a1+a5)+(a3+a7)) */
static inline float hsum8(const float * a) {
    const float r0 = a[4] + a[0], r1 = a[5] + a[1], r2 = a[6] + a[2], r3 = a[7] + a[3];
    const float s0 = r0 + r2, s1 = r1 + r3;
    return s0 + s1;
}
/* GGML_F32x8_REDUCE, ggml.c:1981-1999: 4 registers -> 1, then lo+hi, hadd, hadd */
static inline float reduce4x8(float s[4][8]) {
    float x0[8], x1[8];
    for (int l = 0; l < 8; l++) { x0[l] = s[0][l] + s[2][l]; x1[l] = s[1][l] + s[3][l]; }
    for (int l = 0; l < 8; l++) x0[l] = x0[l] + x1[l];
    const float t0 = x0[0] + x0[4], t1 = x0[1] + x0[5], t2 = x0[2] + x0[6], t3 = x0[3] + x0[7];
    return (t0 + t1) + (t2 + t3);
}

/* ggml_vec_dot_f32, ggml.c:2372-2407, AVX2 build.  The tail loop `sumf += x[i]*y[i]` is what
 * gcc 13 -O3 makes of it (objdump of the reference build): products are formed 8-wide, then
 * 4-wide, UNFUSED (vmulps) and added one at a time in index order; only the last <=3 elements
 * go through a scalar fma.  See DESIGN.md "as-built tail of ggml_vec_dot_f32". */
float bo_vec_dot_f32(int n, const float * x, const float * y) {
    const int np = n & ~31;
    float s[4][8] = {{0}};
    for (int i = 0; i < np; i += 32)
        for (int j = 0; j < 4; j++)
            for (int l = 0; l < 8; l++)
                s[j][l] = fmaf(x[i + j * 8 + l], y[i + j * 8 + l], s[j][l]);
    float sumf = reduce4x8(s);
    const int nv = np + ((n - np) & ~3); /* end of the vectorised (unfused) part of the tail */
    int i = np;
    for (; i < nv; i++) { const float p = x[i] * y[i]; sumf = sumf + p; }
    for (; i < n;  i++) sumf = fmaf(x[i], y[i], sumf);
    return sumf;
}
/* ggml_vec_dot_f16, ggml.c:2409-2443 (K is always a multiple of 32 here; the scalar tail of
 * the reference accumulates in double and is restated for completeness) */
static float vec_dot_f16(int n, const uint16_t * x, const uint16_t * y) {
    const int np = n & ~31;
    float s[4][8] = {{0}};
    for (int i = 0; i < np; i += 32)
        for (int j = 0; j < 4; j++)
            for (int l = 0; l < 8; l++)
                s[j][l] = fmaf(h2f(x[i + j * 8 + l]), h2f(y[i + j * 8 + l]), s[j][l]);
    double sumf = (double) reduce4x8(s);
    for (int i = np; i < n; i++) sumf += (double) (h2f(x[i]) * h2f(y[i]));
    return (float) sumf;
}
/* the five quantised dots: per block, 8 integer lane sums (4 consecutive elements each) are
 * converted to float and fma'd with the combined scale into 8 running lanes */
static float vec_dot_q(int t, int n, const uint8_t * x, const uint8_t * y) {
    const int nb = n / QK;
    const size_t bs = type_size(t);
    const int q81 = (vec_dot_type(t) == T_Q8_1);
    float acc[8] = {0};
    float summs = 0.0f;
    for (int i = 0; i < nb; i++) {
        int q[32]; float dx, mx;
        unpack_block(t, x + i * bs, q, &dx, &mx);
        const uint8_t * yb = y + (size_t) i * (q81 ? 40 : 34);
        float dy, sy = 0.0f;
        const int8_t * yq;
        if (q81) { memcpy(&dy, yb, 4); memcpy(&sy, yb + 4, 4); yq = (const int8_t *) (yb + 8); }
        else     { uint16_t h; memcpy(&h, yb, 2); dy = h2f(h); yq = (const int8_t *) (yb + 2); }
        const float d = dx * dy;
        if (q81) summs = fmaf(mx, sy, summs); /* vfmadd231ss in the reference build */
        for (int l = 0; l < 8; l++) {
            int isum = 0;
            for (int e = 0; e < 4; e++) isum += q[4 * l + e] * yq[4 * l + e];
            acc[l] = fmaf(d, (float) isum, acc[l]);
        }
    }
    const float r = hsum8(acc);
    return q81 ? r + summs : r;
}
float bo_vec_dot(int t, int n, const void * x, const void * y) {
    if (t == T_F32) return bo_vec_dot_f32(n, (const float *) x, (const float *) y);
    if (t == T_F16) return vec_dot_f16(n, (const uint16_t *) x, (const uint16_t *) y);
    return vec_dot_q(t, n, (const uint8_t *) x, (const uint8_t *) y);
}

typedef struct { int t; const void * w; const uint8_t * act; const float * x; float * y;
                 int k, rows, n; size_t wrs, ars; int r0, r1; } mm_job;
static void * mm_worker(void * vj) {
    const mm_job * j = (const mm_job *) vj;
    for (int i = 0; i < j->n; i++) {
        const void * a = j->act ? (const void *) (j->act + j->ars * i) : (const void *) (j->x + (size_t) i * j->k);
        for (int r = j->r0; r < j->r1; r++)
            j->y[(size_t) i * j->rows + r] = bo_vec_dot(j->t, j->k, (const uint8_t *) j->w + j->wrs * r, a);
    }
    return NULL;
}
static int bo_threads(void) {
    const char * e = getenv("BO_THREADS");
    int n = e ? atoi(e) : (int) sysconf(_SC_NPROCESSORS_ONLN);
    return n < 1 ? 1 : n;
}

/* ggml_compute_forward_mul_mat, ggml.c:11804-12013: src1 rows are converted to the weight
 * type's vec_dot_type once (INIT phase), then every output is one vec_dot.
 * y[n][rows] = W[rows][k] . x[n][k] */
void bo_mul_mat(int t, const void * w, const float * x, float * y, int k, int rows, int n) {
    const int vt = vec_dot_type(t);
    const size_t wrs = row_size(t, k), ars = row_size(vt, k);
    uint8_t * act = NULL;
    if (vt != T_F32) {
        act = (uint8_t *) malloc(ars * (size_t) n);
        for (int i = 0; i < n; i++) {
            if (vt == T_Q8_0) bo_quantize_row_q8_0(x + (size_t) i * k, act + ars * i, k);
            if (vt == T_Q8_1) bo_quantize_row_q8_1(x + (size_t) i * k, act + ars * i, k);
            if (vt == T_F16)  bo_fp32_to_fp16_row(x + (size_t) i * k, (uint16_t *) (act + ars * i), k);
        }
    }
    /* every output element is one independent vec_dot, so splitting rows over threads
     * (as the reference does, ggml.c:11941-11954) cannot change any result */
    mm_job jobs[64];
    pthread_t th[64];
    int nt = bo_threads();
    if (nt > 64) nt = 64;
    if ((size_t) rows * k * n < (1u << 16)) nt = 1;
    for (int j = 0; j < nt; j++) {
        mm_job jb = { t, w, act, x, y, k, rows, n, wrs, ars, (int) ((long long) rows * j / nt), (int) ((long long) rows * (j + 1) / nt) };
        jobs[j] = jb;
        if (j > 0) pthread_create(&th[j], NULL, mm_worker, &jobs[j]);
    }
    mm_worker(&jobs[0]);
    for (int j = 1; j < nt; j++) pthread_join(th[j], NULL);
    free(act);
}

/* ---- row ops ---------------------------------------------------------------------------------- */
/* ggml_compute_forward_norm_f32, ggml.c:11377-11426 (no affine) */
void bo_norm(const float * x, float * y, int nc, float eps) {
    double sum = 0.0;
    for (int i = 0; i < nc; i++) sum += (double) x[i];
    const float mean = (float) (sum / nc);
    double sum2 = 0.0;
    for (int i = 0; i < nc; i++) { const float v = x[i] - mean; y[i] = v; sum2 += (double) (v * v); }
    const float variance = (float) (sum2 / nc);
    const float scale = 1.0f / sqrtf(variance + eps);
    for (int i = 0; i < nc; i++) y[i] = y[i] * scale;
}
/* ggml_compute_forward_soft_max_f32, ggml.c:12914-12983 */
void bo_soft_max(const float * x, float * y, int nc) {
    init_tables();
    float max = -INFINITY;
    for (int i = 0; i < nc; i++) if (x[i] > max) max = x[i];
    double sum = 0.0;
    for (int i = 0; i < nc; i++) {
        if (x[i] == -INFINITY) { y[i] = 0.0f; continue; }
        const float val = h2f(g_exp[f2h(x[i] - max)]);
        sum += (double) val;
        y[i] = val;
    }
    const float inv = (float) (1.0 / sum);
    for (int i = 0; i < nc; i++) y[i] = y[i] * inv;
}
/* ggml_vec_gelu_f32 with GGML_GELU_FP16, ggml.c:3853-3861 */
void bo_gelu(const float * x, float * y, int n) {
    init_tables();
    for (int i = 0; i < n; i++) y[i] = h2f(g_gelu[f2h(x[i])]);
}

/* ---- model -------------------------------------------------------------------------------------- */
typedef struct { int type; int ne0, ne1; uint8_t * data; } bo_tensor;
typedef struct {
    bo_tensor q_w, k_w, v_w, o_w, q_b, k_b, v_b, o_b, ln0_w, ln0_b, ln1_w, ln1_b, fc0_w, fc0_b, fc1_w, fc1_b;
} bo_layer;
struct bo_model {
    int32_t n_vocab, n_layer, n_head, n_positions, d_ff, d_model, ftype;
    int wtype;
    bo_tensor embed_tokens, embed_pos, ln_w, ln_b, lm_head;
    bo_layer * layers;
    float * mem_k, * mem_v; /* [n_layer][n_positions][d_model], biogpt.cpp:324-335 */
    bo_taps taps; int have_taps;
};

static bo_tensor * find_tensor(bo_model * m, const char * name) {
    if (!strcmp(name, "biogpt.embed_tokens.weight"))    return &m->embed_tokens;
    if (!strcmp(name, "biogpt.embed_positions.weight")) return &m->embed_pos;
    if (!strcmp(name, "biogpt.layer_norm.weight"))      return &m->ln_w;
    if (!strcmp(name, "biogpt.layer_norm.bias"))        return &m->ln_b;
    if (!strcmp(name, "output_projection.weight"))      return &m->lm_head;
    int li; char rest[128];
    if (sscanf(name, "biogpt.layers.%d.%127s", &li, rest) != 2 || li < 0 || li >= m->n_layer) return NULL;
    bo_layer * L = &m->layers[li];
    static const char * names[] = {
        "self_attn.q_proj.weight", "self_attn.k_proj.weight", "self_attn.v_proj.weight", "self_attn.out_proj.weight",
        "self_attn.q_proj.bias", "self_attn.k_proj.bias", "self_attn.v_proj.bias", "self_attn.out_proj.bias",
        "self_attn_layer_norm.weight", "self_attn_layer_norm.bias", "final_layer_norm.weight", "final_layer_norm.bias",
        "fc1.weight", "fc1.bias", "fc2.weight", "fc2.bias" };
    bo_tensor * slots[] = { &L->q_w, &L->k_w, &L->v_w, &L->o_w, &L->q_b, &L->k_b, &L->v_b, &L->o_b,
        &L->ln0_w, &L->ln0_b, &L->ln1_w, &L->ln1_b, &L->fc0_w, &L->fc0_b, &L->fc1_w, &L->fc1_b };
    for (int i = 0; i < 16; i++) if (!strcmp(rest, names[i])) return slots[i];
    return NULL;
}

bo_model * bo_load(const char * path) {
    init_tables();
    FILE * f = fopen(path, "rb");
    if (!f) { fprintf(stderr, "bo_load: cannot open %s\n", path); return NULL; }
    uint32_t magic = 0;
    if (fread(&magic, 4, 1, f) != 1 || magic != 0x67676d6c) { fprintf(stderr, "bo_load: bad magic\n"); fclose(f); return NULL; }
    bo_model * m = (bo_model *) calloc(1, sizeof *m);
    int32_t hp[7];
    if (fread(hp, 4, 7, f) != 7) goto fail;
    m->n_vocab = hp[0]; m->n_layer = hp[1]; m->n_head = hp[2]; m->n_positions = hp[3];
    m->d_ff = hp[4]; m->d_model = hp[5]; m->ftype = hp[6];
    m->wtype = ftype_to_type(m->ftype);
    if (m->wtype < 0) { fprintf(stderr, "bo_load: bad ftype %d\n", m->ftype); goto fail; }
    for (int pass = 0; pass < 2; pass++) { /* vocab, merges: skipped, the hot path never reads them */
        int32_t n; if (fread(&n, 4, 1, f) != 1) goto fail;
        for (int i = 0; i < n; i++) { uint32_t len; if (fread(&len, 4, 1, f) != 1) goto fail; fseek(f, len, SEEK_CUR); }
    }
    m->layers = (bo_layer *) calloc((size_t) m->n_layer, sizeof(bo_layer));
    int n_loaded = 0;
    for (;;) {
        int32_t h3[3];
        if (fread(h3, 4, 3, f) != 3) break;
        int32_t ne[2] = {1, 1};
        for (int i = 0; i < h3[0]; i++) if (fread(&ne[i], 4, 1, f) != 1) goto fail;
        char name[256] = {0};
        if (h3[1] >= 256 || fread(name, 1, (size_t) h3[1], f) != (size_t) h3[1]) goto fail;
        bo_tensor * t = find_tensor(m, name);
        if (!t) { fprintf(stderr, "bo_load: unknown tensor '%s'\n", name); goto fail; }
        t->type = h3[2]; t->ne0 = ne[0]; t->ne1 = ne[1];
        const size_t nbytes = row_size(t->type, ne[0]) * (size_t) ne[1];
        t->data = (uint8_t *) malloc(nbytes);
        if (fread(t->data, 1, nbytes, f) != nbytes) { fprintf(stderr, "bo_load: short read in '%s'\n", name); goto fail; }
        n_loaded++;
    }
    fclose(f);
    if (n_loaded != 5 + 16 * m->n_layer) { fprintf(stderr, "bo_load: %d tensors, expected %d\n", n_loaded, 5 + 16 * m->n_layer); bo_free(m); return NULL; }
    const size_t nkv = (size_t) m->n_layer * m->n_positions * m->d_model;
    m->mem_k = (float *) calloc(nkv, 4);
    m->mem_v = (float *) calloc(nkv, 4);
    return m;
fail:
    fclose(f);
    bo_free(m);
    return NULL;
}
static void free_t(bo_tensor * t) { free(t->data); t->data = NULL; }
void bo_free(bo_model * m) {
    if (!m) return;
    free_t(&m->embed_tokens); free_t(&m->embed_pos); free_t(&m->ln_w); free_t(&m->ln_b); free_t(&m->lm_head);
    for (int i = 0; m->layers && i < m->n_layer; i++) {
        bo_tensor * s = (bo_tensor *) &m->layers[i];
        for (int j = 0; j < 16; j++) free_t(&s[j]);
    }
    free(m->layers); free(m->mem_k); free(m->mem_v); free(m);
}
void bo_hparams(const bo_model * m, int32_t * o) {
    o[0] = m->n_vocab; o[1] = m->n_layer; o[2] = m->n_head; o[3] = m->n_positions; o[4] = m->d_ff; o[5] = m->d_model; o[6] = m->ftype;
}
void bo_reset(bo_model * m) {
    const size_t nkv = (size_t) m->n_layer * m->n_positions * m->d_model;
    memset(m->mem_k, 0, nkv * 4); memset(m->mem_v, 0, nkv * 4);
}
void bo_set_taps(bo_model * m, const bo_taps * t) {
    if (t) { m->taps = *t; m->have_taps = 1; } else m->have_taps = 0;
}
#define TAP(field, src, count) do { if (m->have_taps && m->taps.field) memcpy(m->taps.field, (src), (size_t) (count) * 4); } while (0)

/* norm + `w (.) y + b` as three separate graph nodes (mul then add, never fused), biogpt.cpp:693-700 */
static void ln_affine(const float * x, float * y, const bo_tensor * w, const bo_tensor * b, int n, int d) {
    const float * wf = (const float *) w->data, * bf = (const float *) b->data;
    for (int i = 0; i < n; i++) {
        bo_norm(x + (size_t) i * d, y + (size_t) i * d, d, NORM_EPS);
        for (int c = 0; c < d; c++) { const float t = wf[c] * y[(size_t) i * d + c]; y[(size_t) i * d + c] = t + bf[c]; }
    }
}
static void linear(const bo_tensor * w, const bo_tensor * b, const float * x, float * y, int n) {
    const int k = w->ne0, rows = w->ne1;
    bo_mul_mat(w->type, w->data, x, y, k, rows, n);
    if (b) { const float * bf = (const float *) b->data;
        for (int i = 0; i < n; i++) for (int r = 0; r < rows; r++) y[(size_t) i * rows + r] = bf[r] + y[(size_t) i * rows + r]; }
}

/* biogpt_graph + biogpt_eval, biogpt.cpp:624-847 */
int bo_eval(bo_model * m, const int32_t * tokens, int N, int n_past, float * logits_out) {
    const int d = m->d_model, nh = m->n_head, dk = d / nh, ff = m->d_ff, T = n_past + N;
    if (N < 1 || T > m->n_positions) return 1;
    float * x    = (float *) malloc((size_t) N * d * 4);
    float * cur  = (float *) malloc((size_t) N * d * 4);
    float * q    = (float *) malloc((size_t) N * d * 4);
    float * kc   = (float *) malloc((size_t) N * d * 4);
    float * vc   = (float *) malloc((size_t) N * d * 4);
    float * att  = (float *) malloc((size_t) N * d * 4);
    float * x1   = (float *) malloc((size_t) N * d * 4);
    float * hff  = (float *) malloc((size_t) N * ff * 4);
    float * tmp  = (float *) malloc((size_t) (d > ff ? d : ff) * 4);
    float * sc   = (float *) malloc((size_t) T * 4);
    float * vt   = (float *) malloc((size_t) T * 4);

    /* embeddings: get_rows (dequantised) * sqrt(d_model) + get_rows(pos, n_past+i+2), biogpt.cpp:663-686 */
    const float emb_scale = sqrtf((float) d);
    for (int i = 0; i < N; i++) {
        const bo_tensor * et = &m->embed_tokens, * ep = &m->embed_pos;
        bo_dequantize_row(et->type, et->data + row_size(et->type, d) * (size_t) tokens[i], tmp, d);
        for (int c = 0; c < d; c++) x[(size_t) i * d + c] = tmp[c] * emb_scale;
        bo_dequantize_row(ep->type, ep->data + row_size(ep->type, d) * (size_t) (n_past + i + 2), tmp, d);
        for (int c = 0; c < d; c++) x[(size_t) i * d + c] = x[(size_t) i * d + c] + tmp[c];
    }
    TAP(embed, x, N * d);

    for (int l = 0; l < m->n_layer; l++) {
        const bo_layer * L = &m->layers[l];
        ln_affine(x, cur, &L->ln0_w, &L->ln0_b, N, d);
        if (l == 0) TAP(layer0_ln, cur, N * d);

        linear(&L->q_w, &L->q_b, cur, q, N);
        for (size_t i = 0; i < (size_t) N * d; i++) q[i] = q[i] * (1.0f / sqrtf((float) dk)); /* biogpt.cpp:681,710 */
        linear(&L->k_w, &L->k_b, cur, kc, N);
        linear(&L->v_w, &L->v_b, cur, vc, N);
        if (l == 0) TAP(layer0_q, q, N * d);

        float * K = m->mem_k + ((size_t) l * m->n_positions) * d;
        float * V = m->mem_v + ((size_t) l * m->n_positions) * d;
        memcpy(K + (size_t) n_past * d, kc, (size_t) N * d * 4); /* biogpt.cpp:721-727 */
        memcpy(V + (size_t) n_past * d, vc, (size_t) N * d * 4);

        /* no causal mask: every query row sees all T = n_past+N positions, biogpt.cpp:730-764 */
        for (int h = 0; h < nh; h++)
            for (int i = 0; i < N; i++) {
                for (int t = 0; t < T; t++)
                    sc[t] = bo_vec_dot_f32(dk, K + (size_t) t * d + h * dk, q + (size_t) i * d + h * dk);
                bo_soft_max(sc, sc, T);
                for (int c = 0; c < dk; c++) {
                    for (int t = 0; t < T; t++) vt[t] = V[(size_t) t * d + h * dk + c]; /* V_trans row */
                    att[(size_t) i * d + h * dk + c] = bo_vec_dot_f32(T, vt, sc);
                }
            }
        if (l == 0) TAP(layer0_att, att, N * d);

        linear(&L->o_w, NULL, att, cur, N);
        { const float * bf = (const float *) L->o_b.data; /* add(cur, repeat(b)) then add(cur, inpL) */
          for (int i = 0; i < N; i++) for (int c = 0; c < d; c++) {
              const float t = cur[(size_t) i * d + c] + bf[c];
              x1[(size_t) i * d + c] = t + x[(size_t) i * d + c]; } }

        ln_affine(x1, cur, &L->ln1_w, &L->ln1_b, N, d);
        linear(&L->fc0_w, &L->fc0_b, cur, hff, N);
        bo_gelu(hff, hff, N * ff);
        linear(&L->fc1_w, &L->fc1_b, hff, cur, N);
        for (size_t i = 0; i < (size_t) N * d; i++) x[i] = cur[i] + x1[i];
        if (l == 0) TAP(layer0_out, x, N * d);
    }

    ln_affine(x, cur, &m->ln_w, &m->ln_b, N, d);
    TAP(final_ln, cur, N * d);
    /* the reference computes all N rows and returns the last (biogpt.cpp:803,844); rows are independent */
    bo_mul_mat(m->lm_head.type, m->lm_head.data, cur + (size_t) (N - 1) * d, logits_out, d, m->n_vocab, 1);

    free(x); free(cur); free(q); free(kc); free(vc); free(att); free(x1); free(hff); free(tmp); free(sc); free(vt);
    return 0;
}
