// bgpt_tu.h -- kernel entry points by translation unit.  The persistent decode kernels, the skinny-batch schedule and the
// tensor-core matmul are each instantiated in their own .cu file (tu_*.cu) so that the library builds in parallel and a
// change to one schedule recompiles one file; bgpt_cuda.cu launches them through these pointers (cudaLaunchKernelExC /
// cudaLaunchCooperativeKernel).  nullptr = no instantiation for that format / shape.
#pragma once
const void * bgpt_k_mega_fn(int wtype, int dk);                 // tu_mega3.cu : k_mega<FMT, DK>
const void * bgpt_k_mega_pick_fn();                             // tu_mega3.cu : k_mega_pick
const void * bgpt_k_mega4_fn(int wtype, bool prof);             // tu_mega4.cu : k_mega4<FMT, PROF>
const void * bgpt_k_mega5_fn(int wtype, bool prof);             // tu_mega5.cu : k_mega5<FMT, PROF>
const void * bgpt_k_rows_fn(int wtype);                          // tu_rows.cu  : k_rows<FMT>
const void * bgpt_k_sk_mm_fn(int wtype, int TN);                // tu_skinny.cu: k_sk_mm<FMT, TN>
const void * bgpt_k_sk_ln_fn(int wtype);
const void * bgpt_k_sk_gq_fn(int wtype);
const void * bgpt_k_sk_attn_fn(int wtype);
const void * bgpt_k_gemm_tc_fn(int wtype);                      // tu_tc.cu    : k_gemm_tc_q<FMT>  (one f32 term per block: tolerance-close)
const void * bgpt_k_gemm_tcxf_fn(int wtype);                    // tu_tc.cu    : k_gemm_tc_xf<FMT> (bit-exact, kind::f16: partial sums arrive as f32)
const void * bgpt_k_gemm_tcx_fn(int wtype);                     // tu_tc.cu    : k_gemm_tc_x<FMT>  (8 running sums per row: bit-exact)
