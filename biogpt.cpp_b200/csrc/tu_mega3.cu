// tu_mega3.cu -- instantiations of the generation-3 persistent decode kernel (bgpt_mega.cuh)
#include "bgpt_mega.cuh"
#include "bgpt_tu.h"

template <int FMT> static const void * mega_fn_dk(int dk) {
    switch (dk) {   // head dims the persistent kernel is instantiated for; others use the per-op kernels
        case 16:  return (const void *) k_mega<FMT, 16>;
        case 64:  return (const void *) k_mega<FMT, 64>;
    }
    return nullptr;
}
const void * bgpt_k_mega_fn(int wtype, int dk) {
    switch (wtype) {
        case BG_Q4_0: return mega_fn_dk<BG_Q4_0>(dk);
        case BG_Q4_1: return mega_fn_dk<BG_Q4_1>(dk);
        case BG_Q5_0: return mega_fn_dk<BG_Q5_0>(dk);
        case BG_Q5_1: return mega_fn_dk<BG_Q5_1>(dk);
        case BG_Q8_0: return mega_fn_dk<BG_Q8_0>(dk);
        case BG_F16:  return mega_fn_dk<BG_F16>(dk);
    }
    return nullptr;   // F32 weights: per-op kernels only
}
const void * bgpt_k_mega_pick_fn() { return (const void *) k_mega_pick; }
