/* oracle/biogpt_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C CPU restatement of the reference's `biogpt_eval` hot path (see biogpt_oracle.c for
 * the per-function reference citations).  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load this library; the product
 * (biogpt.cpp_b200/) never links or calls it.
 *
 * Parity status: PINNED -- checked bit-for-bit against the unmodified reference built by
 * oracle/Makefile (oracle/_ref/libbiogpt_ref.so) in tests/test_oracle_vs_ref.py, and against
 * the committed fixtures in tests/golden/ (generated from that same reference build).
 */
#ifndef BIOGPT_ORACLE_H
#define BIOGPT_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct bo_model bo_model;

bo_model * bo_load(const char * path);          /* NULL on error (message on stderr) */
void       bo_free(bo_model * m);
void       bo_hparams(const bo_model * m, int32_t * out7); /* n_vocab,n_layer,n_head,n_positions,d_ff,d_model,ftype */
void       bo_reset(bo_model * m);              /* zero the KV cache */

/* one biogpt_eval: n tokens at n_past; logits_out gets n_vocab floats (last row). */
int        bo_eval(bo_model * m, const int32_t * tokens, int n, int n_past, float * logits_out);

/* optional taps, filled by the next bo_eval when non-NULL (row-major [n][width]) */
typedef struct {
    float * embed;      /* [n][d_model]   token*32 + position embedding          */
    float * layer0_ln;  /* [n][d_model]   first LayerNorm+affine output          */
    float * layer0_q;   /* [n][d_model]   scaled q of layer 0                    */
    float * layer0_att; /* [n][d_model]   merged attention output of layer 0     */
    float * layer0_out; /* [n][d_model]   layer-0 output (after both residuals)  */
    float * final_ln;   /* [n][d_model]   final LayerNorm+affine output          */
} bo_taps;
void       bo_set_taps(bo_model * m, const bo_taps * taps);

/* unit-level restatements (same arithmetic the eval uses) */
void  bo_quantize_row_q8_0(const float * x, void * y, int k);
void  bo_quantize_row_q8_1(const float * x, void * y, int k);
void  bo_fp32_to_fp16_row(const float * x, uint16_t * y, int k);
void  bo_dequantize_row(int ggml_type, const void * x, float * y, int k);
float bo_vec_dot(int ggml_type, int n, const void * x, const void * y); /* y in the type's vec_dot_type */
float bo_vec_dot_f32(int n, const float * x, const float * y);
void  bo_mul_mat(int ggml_type, const void * w, const float * x, float * y, int k, int rows, int n);
void  bo_norm(const float * x, float * y, int nc, float eps);
void  bo_soft_max(const float * x, float * y, int nc);
void  bo_gelu(const float * x, float * y, int n);
void  bo_tables(uint16_t * gelu65536, uint16_t * exp65536);

#ifdef __cplusplus
}
#endif
#endif
