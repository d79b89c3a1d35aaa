// bgpt_tu.h -- kernel entry points by translation unit.  The persistent decode kernels, the skinny-batch schedule and the
// tensor-core matmul are each instantiated in their own .cu file (tu_*.cu) so that the library builds in parallel and a
// change to one schedule recompiles one file; bgpt_cuda.cu launches them through these pointers (cudaLaunchKernelExC /
// cudaLaunchCooperativeKernel).  nullptr = no instantiation for that format / shape.
#pragma once
const void * bgpt_k_mega_fn(int wtype, int dk);                 // tu_mega3.cu : k_mega<FMT, DK>
const void * bgpt_k_mega_pick_fn();                             // tu_mega3.cu : k_mega_pick
const void * bgpt_k_mega4_fn(int wtype, bool prof);             // tu_mega4.cu : k_mega4<FMT, PROF>
const void * bgpt_k_mega5_fn(int wtype, bool prof, bool tk);    // tu_mega5.cu : k_mega5<FMT, PROF, TK> (tk: token feed + sampler tail)
const void * bgpt_k_rows_fn(int wtype);                          // tu_rows.cu  : k_rows<FMT>
const void * bgpt_k_sk_mm_fn(int wtype, int TN);                // tu_skinny.cu: k_sk_mm<FMT, TN>
const void * bgpt_k_sk_ln_fn(int wtype);
const void * bgpt_k_sk_gq_fn(int wtype);
const void * bgpt_k_sk_attn_fn(int wtype);
const void * bgpt_k_gemm_tc_fn(int wtype);                      // tu_tc.cu    : k_gemm_tc_q<FMT>  (one f32 term per block: tolerance-close)
const void * bgpt_k_gemm_tcxf_fn(int wtype);                    // tu_tc.cu    : k_gemm_tc_xf<FMT> (bit-exact, kind::f16: partial sums arrive as f32)
const void * bgpt_k_gemm_tcx_fn(int wtype);                     // tu_tc.cu    : k_gemm_tc_x<FMT>  (8 running sums per row: bit-exact)
// tu_tcw.cu : warp-specialised, TMA-fed tcgen05 matmuls (bgpt_tcw.cuh) and their host-side launchers
struct GemvArgs; struct Epi;
#ifndef TW_ROWS
#define TW_ROWS 128            // weight rows per tile
#define TW_BK 64               // K elements per pipeline stage
#define TWX_TOK 16             // tokens per tile of the bit-exact kernel
#endif
bool bgpt_tcw_available();                                       // the driver exports cuTensorMapEncodeTiled
cudaError_t bgpt_tcw_decode(int wtype, cudaStream_t s, const GemvArgs & g, int K, void * out16, float * sw, float * mw);
cudaError_t bgpt_tcw_expand(cudaStream_t s, const GemvArgs & g, int K, int n_pad, void * out16, float * sa, float * ss, int hasm);
cudaError_t bgpt_tcw_gemm_exact(int wtype, cudaStream_t s, const void * a16, const float * sw, const float * mw, const void * b16,
                                const float * sa, const float * ss, int M, int K, int n, int tok0, int n_pad, const Epi & epi, int n_sm);
cudaError_t bgpt_tcw_act_h(cudaStream_t s, const GemvArgs & g, int K, void * out16);
cudaError_t bgpt_tcw_gemm_f16(cudaStream_t s, const GemvArgs & g, int K, const void * b16, int n_sm);
