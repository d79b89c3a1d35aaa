// tu_rows.cu -- instantiations of the persistent multi-row kernel (bgpt_rows.cuh): 2..8 token rows per eval in one launch
#include "bgpt_rows.cuh"
#include "bgpt_tu.h"

const void * bgpt_k_rows_fn(int wtype) {
    switch (wtype) {
        case BG_Q4_0: return (const void *) k_rows<BG_Q4_0>;
        case BG_Q4_1: return (const void *) k_rows<BG_Q4_1>;
        case BG_Q5_0: return (const void *) k_rows<BG_Q5_0>;
        case BG_Q5_1: return (const void *) k_rows<BG_Q5_1>;
        case BG_Q8_0: return (const void *) k_rows<BG_Q8_0>;
    }
    return nullptr;
}
