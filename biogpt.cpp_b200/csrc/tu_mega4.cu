// tu_mega4.cu -- instantiations of the generation-4 persistent decode kernel (bgpt_mega4.cuh).
// prof: the instantiation with clock stamps (BGPT_MEGA_PROF); the production kernel carries none of that code --
// the per-layer loop has to stay inside the SM's instruction cache (profiles/README.md)
#include "bgpt_mega4.cuh"
#include "bgpt_tu.h"

const void * bgpt_k_mega4_fn(int wtype, bool prof) {
    switch (wtype) {
        case BG_Q4_0: return prof ? (const void *) k_mega4<BG_Q4_0, true> : (const void *) k_mega4<BG_Q4_0, false>;
        case BG_Q4_1: return prof ? (const void *) k_mega4<BG_Q4_1, true> : (const void *) k_mega4<BG_Q4_1, false>;
        case BG_Q5_0: return prof ? (const void *) k_mega4<BG_Q5_0, true> : (const void *) k_mega4<BG_Q5_0, false>;
        case BG_Q5_1: return prof ? (const void *) k_mega4<BG_Q5_1, true> : (const void *) k_mega4<BG_Q5_1, false>;
        case BG_Q8_0: return prof ? (const void *) k_mega4<BG_Q8_0, true> : (const void *) k_mega4<BG_Q8_0, false>;
    }
    return nullptr;
}
