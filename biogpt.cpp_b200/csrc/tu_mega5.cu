// tu_mega5.cu -- instantiations of the generation-5 persistent decode kernel (bgpt_mega5.cuh): thread-block clusters of 8,
// one attention head per cluster.  prof: the instantiation with clock stamps (BGPT_MEGA_PROF).
#include "bgpt_mega5.cuh"
#include "bgpt_tu.h"

const void * bgpt_k_mega5_fn(int wtype, bool prof, bool tk) {
    if (tk) {                                                     // the sampler's instantiation (no trace variant)
        switch (wtype) {
            case BG_Q4_0: return (const void *) k_mega5<BG_Q4_0, false, true>;
            case BG_Q4_1: return (const void *) k_mega5<BG_Q4_1, false, true>;
            case BG_Q5_0: return (const void *) k_mega5<BG_Q5_0, false, true>;
            case BG_Q5_1: return (const void *) k_mega5<BG_Q5_1, false, true>;
            case BG_Q8_0: return (const void *) k_mega5<BG_Q8_0, false, true>;
            case BG_F16:  return (const void *) k_mega5<BG_F16, false, true>;
        }
        return nullptr;
    }
    switch (wtype) {
        case BG_Q4_0: return prof ? (const void *) k_mega5<BG_Q4_0, true, false> : (const void *) k_mega5<BG_Q4_0, false, false>;
        case BG_Q4_1: return prof ? (const void *) k_mega5<BG_Q4_1, true, false> : (const void *) k_mega5<BG_Q4_1, false, false>;
        case BG_Q5_0: return prof ? (const void *) k_mega5<BG_Q5_0, true, false> : (const void *) k_mega5<BG_Q5_0, false, false>;
        case BG_Q5_1: return prof ? (const void *) k_mega5<BG_Q5_1, true, false> : (const void *) k_mega5<BG_Q5_1, false, false>;
        case BG_Q8_0: return prof ? (const void *) k_mega5<BG_Q8_0, true, false> : (const void *) k_mega5<BG_Q8_0, false, false>;
        case BG_F16:  return prof ? (const void *) k_mega5<BG_F16, true, false>  : (const void *) k_mega5<BG_F16, false, false>;
    }
    return nullptr;
}
