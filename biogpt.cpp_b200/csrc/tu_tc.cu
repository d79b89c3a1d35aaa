// tu_tc.cu -- instantiations of the tcgen05 prompt-batch matmul (bgpt_tc.cuh)
#include "bgpt_tc.cuh"
#include "bgpt_tu.h"

const void * bgpt_k_gemm_tc_fn(int wtype) {
    switch (wtype) {
        case BG_Q4_0: return (const void *) k_gemm_tc_q<BG_Q4_0>; case BG_Q4_1: return (const void *) k_gemm_tc_q<BG_Q4_1>;
        case BG_Q5_0: return (const void *) k_gemm_tc_q<BG_Q5_0>; case BG_Q5_1: return (const void *) k_gemm_tc_q<BG_Q5_1>;
        case BG_Q8_0: return (const void *) k_gemm_tc_q<BG_Q8_0>;
    }
    return nullptr;
}

const void * bgpt_k_gemm_tcx_fn(int wtype) {
    switch (wtype) {
        case BG_Q4_0: return (const void *) k_gemm_tc_x<BG_Q4_0>; case BG_Q4_1: return (const void *) k_gemm_tc_x<BG_Q4_1>;
        case BG_Q5_0: return (const void *) k_gemm_tc_x<BG_Q5_0>; case BG_Q5_1: return (const void *) k_gemm_tc_x<BG_Q5_1>;
        case BG_Q8_0: return (const void *) k_gemm_tc_x<BG_Q8_0>;
    }
    return nullptr;
}

const void * bgpt_k_gemm_tcxf_fn(int wtype) {
    switch (wtype) {
        case BG_Q4_0: return (const void *) k_gemm_tc_xf<BG_Q4_0>; case BG_Q4_1: return (const void *) k_gemm_tc_xf<BG_Q4_1>;
        case BG_Q5_0: return (const void *) k_gemm_tc_xf<BG_Q5_0>; case BG_Q5_1: return (const void *) k_gemm_tc_xf<BG_Q5_1>;
        case BG_Q8_0: return (const void *) k_gemm_tc_xf<BG_Q8_0>;
    }
    return nullptr;
}
