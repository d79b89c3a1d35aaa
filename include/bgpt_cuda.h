/* include/bgpt_cuda.h -- the drop-in boundary of the B200 `biogpt_eval` hot path.
 *
 * A plain C ABI (extern "C", pointers and sizes only, no C++ / torch types) exported by
 * biogpt.cpp_b200/csrc/libbgpt_cuda.so.  It is what the reference's C++ model API binds to
 * when its ggml CPU graph is replaced by the device engine:
 *
 *   reference interface (under /root/reference)              replaced by
 *   -------------------------------------------------------  ---------------------------------
 *   biogpt_model_load: ggml_init / ggml_backend_alloc_buffer  bgpt_cuda_model_create
 *     biogpt.cpp:215-242, KV cache biogpt.cpp:324-358
 *   biogpt_model_load: tensor loop, the non-CPU upload hook   bgpt_cuda_upload_tensor
 *     ggml_backend_tensor_set, biogpt.cpp:369-434 (421-427)
 *   ggml fp16 GELU / exp tables, ggml.c:4620-4640             bgpt_cuda_set_tables
 *   "all tensors present" check, biogpt.cpp:442-447           bgpt_cuda_model_finalize
 *   biogpt_eval + biogpt_graph + ggml_backend_graph_compute   bgpt_cuda_eval
 *     biogpt.cpp:812-847, 624-810
 *   ggml_free / ggml_backend_buffer_free / ggml_backend_free  bgpt_cuda_model_free
 *     examples/main/main.cpp:164-169
 *
 * The host side (biogpt.cpp_b200/host/, C++) keeps the reference's own signatures
 * (biogpt_model_load / biogpt_eval / biogpt_sample_top_k_top_p, biogpt.h:128-172) and calls
 * only the functions below.  Every function returns 0 on success or a negative BGPT_E_* code;
 * bgpt_cuda_last_error() describes the last failure on the calling thread.  Nothing here ever
 * falls back to the CPU: without a CUDA device every entry point fails with BGPT_E_CUDA.
 */
#ifndef BGPT_CUDA_H
#define BGPT_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BGPT_OK            0
#define BGPT_E_CUDA      (-1)  /* CUDA runtime error / no device */
#define BGPT_E_ARG       (-2)  /* bad argument (shape, type, range) */
#define BGPT_E_STATE     (-3)  /* call order (e.g. eval before finalize) */
#define BGPT_E_UNSUPPORTED (-4)

/* ggml_type codes used in the `.bin` tensor headers (ggml.h:306-314) */
#define BGPT_TYPE_F32   0
#define BGPT_TYPE_F16   1
#define BGPT_TYPE_Q4_0  2
#define BGPT_TYPE_Q4_1  3
#define BGPT_TYPE_Q5_0  6
#define BGPT_TYPE_Q5_1  7
#define BGPT_TYPE_Q8_0  8

typedef struct bgpt_model bgpt_model;

/* ---- library / device ------------------------------------------------------------------ */
const char * bgpt_cuda_last_error(void);
int          bgpt_cuda_device_count(void);               /* <0 on error */
const char * bgpt_cuda_version(void);

/* ---- model life cycle ------------------------------------------------------------------- */
/* hparams7 = { n_vocab, n_layer, n_head, n_positions, d_ff, d_model, ftype } exactly as stored
 * in the file header (biogpt.cpp:54-60).  Allocates the F32 KV cache
 * [n_layer][n_positions][d_model] x2 (biogpt.cpp:324-335) and the activation arena for up to
 * `max_batch` tokens per eval (the reference's n_batch; evals with more tokens re-size it). */
bgpt_model * bgpt_cuda_model_create(const int32_t hparams7[7], int device, int max_batch);

/* One tensor of the file, `data` = the raw bytes that follow its header in the `.bin`
 * (nbytes is checked against ne0*ne1 and the type, like biogpt.cpp:412-417). Names are the
 * loader's lookup keys (biogpt.cpp:258-317). Matrices are re-tiled on upload (same bytes,
 * permuted inside each row) so that device loads are aligned 128-bit; see DESIGN.md. */
int bgpt_cuda_upload_tensor(bgpt_model * m, const char * name, int ggml_type,
                            int64_t ne0, int64_t ne1, const void * data, size_t nbytes);

/* The two 65536-entry fp16 lookup tables of ggml (GELU and exp, ggml.c:4620-4640), built by
 * the caller with the host libm so that they are the tables the reference would build on
 * this machine. */
int bgpt_cuda_set_tables(bgpt_model * m, const uint16_t * gelu_f16, const uint16_t * exp_f16);
/* Host helper (no GPU involved): fills both tables with the host libm, the way ggml_init
 * does (ggml.c:4620-4640). */
void bgpt_host_build_tables(uint16_t * gelu_f16, uint16_t * exp_f16);

/* Checks that all 4 + 18*n_layer tensors and the tables arrived; after this the model is
 * immutable and bgpt_cuda_eval may be called. */
int bgpt_cuda_model_finalize(bgpt_model * m);

void bgpt_cuda_model_free(bgpt_model * m);

/* ---- the hot path ------------------------------------------------------------------------- */
/* biogpt_eval (biogpt.cpp:812-847): `n` tokens at positions [n_past, n_past+n) of stream 0;
 * `logits_out` (HOST) receives the n_vocab logits of the last token.  Attention is
 * un-masked over all n_past+n positions exactly like the reference graph
 * (biogpt.cpp:741-744). Host->device copy of the ids and device->host copy of the logits
 * are part of the call.  Single-token steps on the generation-5 decode kernel (n == 1, the loop of
 * examples/main/main.cpp:93-151) are served like bgpt_cuda_eval_topk's: the kernel writes the logit row into
 * mapped pinned host memory itself and the launch for position n_past + 1 is chained (bgpt_cuda_set_chain);
 * bgpt_cuda_last_eval_ms is 0 for those calls. */
int bgpt_cuda_eval(bgpt_model * m, const int32_t * tokens, int n, int n_past, float * logits_out);

/* Same step with everything resident in HBM: token ids are read from `d_tokens` (device
 * pointer) and the logits stay in the model's device buffer (bgpt_cuda_logits_device).
 * Asynchronous on the model's stream; used by the bench to time the kernels alone. */
int bgpt_cuda_eval_device(bgpt_model * m, const int32_t * d_tokens, int n, int n_past);
const float * bgpt_cuda_logits_device(bgpt_model * m);
int bgpt_cuda_synchronize(bgpt_model * m);

/* biogpt_eval + the K largest logits of the returned row, selected on the device (csrc/bgpt_topk.cuh), for
 * biogpt_sample_top_k_top_p (biogpt.cpp:908-980): only the top_k (logit, id) pairs influence the reference's draw, so 8 K + 8
 * bytes cross PCIe per token instead of 4 n_vocab.  vals / ids (HOST, K entries) are sorted by logit descending; *n_out =
 * min(K, n_vocab).  *exact = 0 when equal values make std::partial_sort's selection or order ambiguous -- then the whole logit row
 * is copied to logits_fallback (HOST, n_vocab floats; may be NULL) and the caller runs the reference's sampler on it, so the
 * drawn id is the reference's in every case.  K <= 128. */
int bgpt_cuda_eval_topk(bgpt_model * m, const int32_t * tokens, int n, int n_past, int k,
                        float * vals, int32_t * ids, int * n_out, int * exact, float * logits_fallback);
/* Chained launches of bgpt_cuda_eval_topk and of single-token bgpt_cuda_eval (generation-5 decode kernel; default on, BGPT_CHAIN=0 turns
 * it off): a sampling loop
 * (examples/main/main.cpp:93-151) asks for position p + 1 right after position p, so the call for p also queues the kernel of p + 1
 * behind the one it waits for.  That kernel starts the moment its predecessor ends, fetches its first weights and polls an 8-byte
 * word in mapped pinned host memory; the next call only writes the sampled token id there -- no launch and no front-end latency on
 * the token-to-token path.  Any other call on the model withdraws the queued kernel first (it exits without touching the KV cache);
 * a kernel that saw no token within BGPT_CHAIN_WAIT_US (default 2000) gives up and the position is evaluated by a fresh launch.
 * bgpt_cuda_last_eval_ms is 0 for chained calls.  on: 1 / 0, -1 = the BGPT_CHAIN default. */
int bgpt_cuda_set_chain(bgpt_model * m, int on);

/* Greedy decode entirely on the device: starting from `first_token` at position n_past,
 * run `n_steps` evals of one token each, feeding argmax(logits) back in (first index wins
 * ties, like std::partial_sort with top_k = 1 in biogpt_sample_top_k_top_p,
 * biogpt.cpp:908-980).  ids_out (HOST, n_steps entries) receives the sampled ids.
 * `ms_out` (optional) receives the device time of the loop measured with CUDA events. */
int bgpt_cuda_decode_greedy(bgpt_model * m, int32_t first_token, int n_past, int n_steps,
                            int32_t * ids_out, float * ms_out);

/* Which schedule evaluates single-token steps (n == 1): 1 = one persistent kernel per token, newest generation available
 * (default: generation 5 -- thread-block clusters, one attention head per cluster, csrc/bgpt_mega5.cuh -- for quantised
 * weights at BioGPT-base shapes, else generation 3), 3 = generation 4 (tagged-word exchange without clusters,
 * csrc/bgpt_mega4.cuh), 2 = generation 3 (grid barriers, csrc/bgpt_mega.cuh; every format and shape), 0 = one kernel per
 * fused operator.  All produce identical bits; tests compare them.  BGPT_MEGA_V=3|4 caps the generation at load time. */
int bgpt_cuda_set_decode_path(bgpt_model * m, int path);
int bgpt_cuda_get_decode_path(const bgpt_model * m);
/* Which schedule evaluates skinny batches (2 <= n < 128 token rows: prompt chunks of the reference's
 * n_batch = 8 and lock-step streams): 1 (default) = the fused schedule of csrc/bgpt_skinny.cuh (quantised
 * weights at BioGPT-base layer shapes; 8 launches per layer chained by programmatic dependent launch);
 * 2 = the persistent multi-row kernel of csrc/bgpt_rows.cuh for 2..8 rows -- ONE launch per eval, stage
 * boundaries are counters in L2 -- and the fused schedule beyond (opt-in: it measures 15-25 % slower than
 * the fused schedule, profiles/README.md); 0 = one kernel per fused operator.
 * Identical bits; tests compare them.  get: 2 / 1 / 0 = what an n_rows-row eval would run on. */
int bgpt_cuda_set_batch_path(bgpt_model * m, int path);
int bgpt_cuda_get_batch_path(const bgpt_model * m, int n_rows);
/* Which schedule an eval of n_rows token rows takes on this model: 3 = persistent decode kernel (n_rows == 1), 5 = persistent
 * multi-row kernel (2..8 rows), 1 = fused skinny-batch schedule, 0 = per-operator schedule with the exact-order SIMT matmul -- these give the reference's bits
 * -- as does 4 = per-operator schedule with the bit-exact tcgen05 matmul (k_tcw_exact / k_gemm_tc_xf; quantised evals of 128+ rows);
 * these are the only ones used by default -- 7 = F16 weights, opt-in (bgpt_cuda_set_f16_tc_min_rows): per-operator schedule with
 * the K-accumulating tcgen05 matmul (k_tcw_f16; tolerance-close); 2 = per-operator schedule with the one-term-per-block tcgen05 matmul of csrc/bgpt_tc.cuh
 * (exact integer block dots but one f32 term per block: logits drift by ~5e-2, see tests/test_gpu_eval.py), which is OFF
 * unless enabled with bgpt_cuda_set_tc_min_rows / BGPT_TC_MIN_ROWS. */
int bgpt_cuda_get_eval_path(const bgpt_model * m, int n_rows);
/* the bit-exact tcgen05 matmul (csrc/bgpt_tc.cuh, k_gemm_tc_x: the reference's 8 running sums per row through masked activation
 * columns) serves quantised evals of `rows`+ token rows on the per-operator schedule (default 128; 0 = off: the skinny-batch
 * schedule then serves every batch size).  bgpt_cuda_get_eval_path reports 4 for it; results are the reference's bits. */
int bgpt_cuda_set_tcx_min_rows(bgpt_model * m, int rows);
/* 1 (default): that bit-exact matmul runs as the warp-specialised, TMA-fed kernel of csrc/bgpt_tcw.cuh (k_tcw_exact: a TMA
 * producer warp, an MMA warp, eight epilogue warps, double-buffered TMEM; operands = the model's prompt-operand cache + the
 * eval's expanded activations) wherever rows % 128 == 0 and K % 64 == 0 (all BioGPT shapes); 0: k_gemm_tc_xf.  Same bits.
 * BGPT_TCW=0 at load does the same. */
int bgpt_cuda_set_tcw(bgpt_model * m, int on);
/* F16 weights, OPT-IN (default 0 = off: every F16 eval is bit-exact): evals of `rows`+ token rows run their matmuls on the
 * K-accumulating tcgen05 kernel k_tcw_f16 (csrc/bgpt_tcw.cuh; fp16 x fp16 -> f32 in TMEM, fed by TMA from the weights where they
 * lie; 1024 prompt tokens in one eval: 138 -> 12.6 ms).  The tensor core adds the reference's products in its own order: each
 * matmul is within 3e-7 (relative to its largest output) of the reference, the error of any f32 re-ordering, and the logits of
 * the 2-layer test model within 8e-4 -- but 24 layers of fp16 re-rounding amplify it to 1.9e-3 on the full-size synthetic model
 * (same argmax and top-5), outside the north star's 1e-3 gate for f16, hence opt-in.  bgpt_cuda_get_eval_path reports 7 for it.
 * BGPT_F16_TC_MIN_ROWS at load does the same. */
int bgpt_cuda_set_f16_tc_min_rows(bgpt_model * m, int rows);
/* opt in to the tolerance-close integer tcgen05 matmul for quantised evals of `rows`+ token rows (0 = off, the default) */
int bgpt_cuda_set_tc_min_rows(bgpt_model * m, int rows);
/* debug: copy one of the eval arena's buffers as the last eval left it (0 x, 1 x1, 2 q, 3 the d_model-wide activation
 * records, 4 the d_ff-wide activation records; `rows` token rows) to HOST memory; returns the bytes copied, -1 on error.
 * tools/skinny_check.py uses it to localise a mismatch between the two batch schedules. */
long long bgpt_cuda_debug_read_buffer(bgpt_model * m, int which, int rows, void * out, long long cap_bytes);
/* 5, 4 or 3: the persistent-kernel generation single-token steps run on; 0: per-operator kernels */
int bgpt_cuda_decode_kernel_generation(const bgpt_model * m);
/* debug (env BGPT_MEGA_PROF=1 at load): per-phase clock64 stamps of CTA 0 of the last
 * persistent-kernel launch; returns the number of entries copied (0 when profiling is off). */
int bgpt_cuda_debug_read_prof(bgpt_model * m, long long * out, int cap);
/* debug, generation-4 kernel: the stamps of every CTA, [n_cta][per_cta], followed by [n_cta][4] =
 * {globaltimer ns, clock64} pairs taken at the start and the end of the launch, which put the
 * per-SM clocks on one time axis.  Returns the number of entries copied (0: off / cap too small). */
int bgpt_cuda_debug_read_trace(bgpt_model * m, long long * out, int cap, int * n_cta, int * per_cta);
/* debug (BGPT_MEGA_PROF=1): clock64 stamps of CTA 0 in the last multi-row launch, [n_layer + 1][8 stages]; [n_layer][7] = start */
int bgpt_cuda_debug_read_rows_trace(bgpt_model * m, long long * out, int cap);
/* f32 -> Q4_0/Q4_1/Q5_0/Q5_1/Q8_0 blocks in the file layout, on the device; bit-identical to the reference's
 * quantize_row_q*_reference as its `quantize` tool runs them (ggml.c:892-1094, biogpt.cpp:459-621).  `type` is a
 * ggml_type; n (a multiple of 32) host floats in, n/32 blocks out. */
int bgpt_cuda_op_quantize_weights(int type, const float * x, long long n, uint8_t * out);

/* multi-stream state: `n_streams` independent sequences, each with its own KV cache
 * (SURVEY 8(d) config 4).  Stream 0 always exists. */
int bgpt_cuda_set_streams(bgpt_model * m, int n_streams);
/* lock-step decode: one token per stream at the same n_past; logits_out = [n_streams][n_vocab]
 * on the HOST (may be NULL to leave them on the device). */
int bgpt_cuda_eval_streams(bgpt_model * m, const int32_t * tokens, int n_streams, int n_past,
                           float * logits_out);
/* bgpt_cuda_eval_streams + the K largest logits of every stream's row, selected on the device (bgpt_cuda_eval_topk's three-launch
 * selection over all rows; biogpt.cpp:908-980): vals / ids [n_streams][k], n_out / exact [n_streams] (HOST).  A row with exact = 0
 * (equal values make std::partial_sort's choice ambiguous) also gets its full logit row in logits_fallback[row] ([n_streams][n_vocab],
 * may be NULL).  K <= 128. */
int bgpt_cuda_eval_streams_topk(bgpt_model * m, const int32_t * tokens, int n_streams, int n_past, int k,
                                float * vals, int32_t * ids, int * n_out, int * exact, float * logits_fallback);

/* greedy decode of n_streams lock-step streams entirely on the device: n_steps forward passes over the n_streams rows, per-row
 * argmax fed back on the device, one synchronisation at the end.  ids_out (HOST) = [n_steps][n_streams]; ms_out = device time. */
int bgpt_cuda_decode_greedy_streams(bgpt_model * m, const int32_t * first_tokens, int n_streams, int n_past, int n_steps,
                                    int32_t * ids_out, float * ms_out);

/* ---- introspection ---------------------------------------------------------------------- */
void   bgpt_cuda_hparams(const bgpt_model * m, int32_t out7[7]);
size_t bgpt_cuda_weight_bytes(const bgpt_model * m);     /* device bytes of all uploaded tensors */
/* kernels launched by the library since load (the bench reports launches per step) */
uint64_t bgpt_cuda_launch_count(const bgpt_model * m);
/* device time of the last bgpt_cuda_eval / eval_streams call, CUDA events on its stream */
float  bgpt_cuda_last_eval_ms(const bgpt_model * m);
/* debug taps of the last eval, [n][d_model] floats each, HOST pointers, any may be NULL:
 * which = 0 embed (token*sqrt(d)+pos), 1 layer-0 scaled q, 2 layer-0 merged attention,
 * 3 layer-0 output, 4 input of the final LayerNorm.  Must be armed before the eval. */
int    bgpt_cuda_set_taps(bgpt_model * m, float * const taps5[5]);

/* ---- unit-level operators (same arithmetic the eval uses; parity tests call these) ------- */
/* y[n][rows] = W[rows][k] . x[n][k]; `w` in file layout for `ggml_type`
 * (ggml_compute_forward_mul_mat, ggml.c:11804-12013).  All pointers HOST. */
int bgpt_cuda_op_mul_mat(int ggml_type, const void * w, const float * x, float * y,
                         int k, int rows, int n);
/* the same product through the BIT-EXACT tcgen05 tensor-core kernel used for prompt batches of 128+ rows (csrc/bgpt_tc.cuh,
 * k_gemm_tc_x; quantised types only): identical bits to bgpt_cuda_op_mul_mat */
int bgpt_cuda_op_mul_mat_tcx(int ggml_type, const void * w, const float * x, float * y,
                             int k, int rows, int n);
/* the same product through the warp-specialised TMA-fed kernels of csrc/bgpt_tcw.cuh: quantised types -> k_tcw_exact (identical
 * bits to bgpt_cuda_op_mul_mat), F16 -> k_tcw_f16 (tolerance-close).  rows % 128 == 0, k % 64 == 0 (F16: k % 256 == 0). */
int bgpt_cuda_op_mul_mat_tcw(int ggml_type, const void * w, const float * x, float * y,
                             int k, int rows, int n);
/* the same product through the opt-in one-term-per-block tcgen05 kernel (k_gemm_tc_q).  Exact integer block dots, f32 block
 * accumulation in block order: close to, not bit-identical with, the CPU reference (see the file header). */
int bgpt_cuda_op_mul_mat_tc(int ggml_type, const void * w, const float * x, float * y,
                            int k, int rows, int n);
/* the device top-k selection of bgpt_cuda_eval_topk applied to a HOST logit row: the k largest (value, index) pairs by value
 * descending; *exact = 0 when equal values make std::partial_sort's choice or order ambiguous (biogpt.cpp:908-980) */
int bgpt_cuda_op_topk(const float * logits, int n, int k, float * vals, int32_t * ids, int * n_out, int * exact);
/* activation quantisers applied to src1 by mul_mat (ggml.c:1166-1249, 1403-1494, 493-510):
 * out = k/32 blocks of block_q8_0 (34 B) / block_q8_1 (40 B) for the weight type's
 * vec_dot_type, or k fp16 values for F16 weights. */
int bgpt_cuda_op_quantize_act(int weight_type, const float * x, void * out, int k);
/* LayerNorm + affine (ggml.c:11377-11426 then mul, add; biogpt.cpp:693-700); w or b may be
 * NULL for the bare norm. */
int bgpt_cuda_op_norm(const float * x, const float * w, const float * b, float * y,
                      int rows, int nc, float eps);
/* one attention call: q [n][d_model] (already scaled), caches k,v [T][d_model] with
 * T = n_past + n; out [n][d_model]  (biogpt.cpp:730-764) */
int bgpt_cuda_op_attention(const float * q, const float * k, const float * v, float * out,
                           int n, int n_past, int d_model, int n_head,
                           const uint16_t * exp_f16);
/* y = fp16-table GELU (ggml.c:3853-3861) */
int bgpt_cuda_op_gelu(const float * x, float * y, int n, const uint16_t * gelu_f16);
/* dequantise `rows` rows of k elements (get_rows, ggml.c:12524-12628) */
int bgpt_cuda_op_dequantize(int ggml_type, const void * w, float * y, int k, int rows);

#ifdef __cplusplus
}
#endif
#endif /* BGPT_CUDA_H */
