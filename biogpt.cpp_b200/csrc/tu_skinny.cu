// tu_skinny.cu -- instantiations of the fused skinny-batch schedule (bgpt_skinny.cuh)
#include "bgpt_skinny.cuh"
#include "bgpt_tu.h"

template <int FMT> static const void * sk_mm_fn(int TN) { return TN == 4 ? (const void *) k_sk_mm<FMT, 4> : (const void *) k_sk_mm<FMT, 8>; }
const void * bgpt_k_sk_mm_fn(int wtype, int TN) {
    switch (wtype) {
        case BG_Q4_0: return sk_mm_fn<BG_Q4_0>(TN); case BG_Q4_1: return sk_mm_fn<BG_Q4_1>(TN); case BG_Q5_0: return sk_mm_fn<BG_Q5_0>(TN);
        case BG_Q5_1: return sk_mm_fn<BG_Q5_1>(TN); case BG_Q8_0: return sk_mm_fn<BG_Q8_0>(TN);
    }
    return nullptr;
}
const void * bgpt_k_sk_ln_fn(int wtype) {
    switch (wtype) {
        case BG_Q4_0: return (const void *) k_sk_ln<BG_Q4_0>; case BG_Q4_1: return (const void *) k_sk_ln<BG_Q4_1>;
        case BG_Q5_0: return (const void *) k_sk_ln<BG_Q5_0>; case BG_Q5_1: return (const void *) k_sk_ln<BG_Q5_1>;
        case BG_Q8_0: return (const void *) k_sk_ln<BG_Q8_0>;
    }
    return nullptr;
}
const void * bgpt_k_sk_gq_fn(int wtype) {
    switch (wtype) {
        case BG_Q4_0: return (const void *) k_sk_gq<BG_Q4_0>; case BG_Q4_1: return (const void *) k_sk_gq<BG_Q4_1>;
        case BG_Q5_0: return (const void *) k_sk_gq<BG_Q5_0>; case BG_Q5_1: return (const void *) k_sk_gq<BG_Q5_1>;
        case BG_Q8_0: return (const void *) k_sk_gq<BG_Q8_0>;
    }
    return nullptr;
}
const void * bgpt_k_sk_attn_fn(int wtype) {
    switch (wtype) {
        case BG_Q4_0: return (const void *) k_sk_attn<BG_Q4_0>; case BG_Q4_1: return (const void *) k_sk_attn<BG_Q4_1>;
        case BG_Q5_0: return (const void *) k_sk_attn<BG_Q5_0>; case BG_Q5_1: return (const void *) k_sk_attn<BG_Q5_1>;
        case BG_Q8_0: return (const void *) k_sk_attn<BG_Q8_0>;
    }
    return nullptr;
}
