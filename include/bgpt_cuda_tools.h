/* bgpt_cuda_tools.h -- micro-benchmarks used while designing the kernels (tools/*.py).  They are NOT part of the product:
 * libbgpt_cuda.so does not contain them; `make -C biogpt.cpp_b200/csrc tools` builds libbgpt_cuda_tools.so (the same
 * sources with -DBGPT_BENCH_TOOLS), which exports the product ABI of bgpt_cuda.h plus the entry points below. */
#ifndef BGPT_CUDA_TOOLS_H
#define BGPT_CUDA_TOOLS_H
#include "bgpt_cuda.h"
#ifdef __cplusplus
extern "C" {
#endif
/* microseconds per grid-wide barrier / all-to-all exchange for the candidate implementations in csrc/bgpt_barbench.cuh
 * (one CTA per SM, `iters` back-to-back rounds) */
int bgpt_cuda_debug_barrier_bench(int variant, int iters, int with_load, float * us_per_barrier);
/* cycles per iteration of a loop with `kb` KB of straight-line code on every SM at once */
int bgpt_cuda_debug_icache_bench(int kb, int iters, int nwarps, float * cycles_per_iter);
/* microseconds per launch of the device weight quantiser over n synthetic weights (device resident) */
int bgpt_cuda_debug_quantize_bench(int type, long long n, int iters, float * us_per_launch);
/* milliseconds per matmul y[n][rows] = W[rows][k].x[n][k] (synthetic data, device resident, `iters` back-to-back
 * launches); path 0 = exact-order SIMT kernels, 1 = tcgen05 */
int bgpt_cuda_debug_gemm_bench(int ggml_type, int k, int rows, int n, int iters, int path, float * ms_out);
#ifdef __cplusplus
}
#endif
#endif
