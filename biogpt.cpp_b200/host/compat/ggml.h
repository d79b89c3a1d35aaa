// compat/ggml.h -- the sliver of ggml's public surface that the reference's own front ends
// (/root/reference/examples/main/main.cpp, examples/quantize/quantize.cpp) and its biogpt.h touch.
// This is NOT ggml: there is no tensor library behind these names.  The B200 build keeps the
// weights and the whole forward pass on the device (csrc/); the opaque handles below only carry
// what the reference's calling convention needs so that its main.cpp compiles and links unchanged
// (SURVEY 8(b)).  Reference declarations: ggml/include/ggml/ggml.h:306-345, 658, 1871-1872.
#pragma once
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

// tensor storage types and file types: the numeric values are the on-disk contract
enum ggml_type {
    GGML_TYPE_F32 = 0, GGML_TYPE_F16 = 1, GGML_TYPE_Q4_0 = 2, GGML_TYPE_Q4_1 = 3,
    GGML_TYPE_Q5_0 = 6, GGML_TYPE_Q5_1 = 7, GGML_TYPE_Q8_0 = 8, GGML_TYPE_Q8_1 = 9,
    GGML_TYPE_COUNT = 19,
};
enum ggml_ftype {
    GGML_FTYPE_UNKNOWN = -1, GGML_FTYPE_ALL_F32 = 0, GGML_FTYPE_MOSTLY_F16 = 1,
    GGML_FTYPE_MOSTLY_Q4_0 = 2, GGML_FTYPE_MOSTLY_Q4_1 = 3, GGML_FTYPE_MOSTLY_Q4_1_SOME_F16 = 4,
    GGML_FTYPE_MOSTLY_Q8_0 = 7, GGML_FTYPE_MOSTLY_Q5_0 = 8, GGML_FTYPE_MOSTLY_Q5_1 = 9,
    GGML_FTYPE_MOSTLY_Q2_K = 10, GGML_FTYPE_MOSTLY_Q3_K = 11, GGML_FTYPE_MOSTLY_Q4_K = 12,
    GGML_FTYPE_MOSTLY_Q5_K = 13, GGML_FTYPE_MOSTLY_Q6_K = 14,
};

typedef uint16_t ggml_fp16_t;

// handle of one model tensor: where it lives on the device is the engine's business
struct ggml_tensor {
    enum ggml_type type;
    int64_t ne[4];
    char    name[64];
};
struct ggml_context;   // owns the device engine of one loaded model
struct ggml_cgraph;    // the per-eval op graph of the reference; a placeholder here

void    ggml_time_init(void);
int64_t ggml_time_us(void);
int64_t ggml_time_ms(void);
void    ggml_free(struct ggml_context * ctx);

enum ggml_type ggml_ftype_to_ggml_type(enum ggml_ftype ftype);
const char *   ggml_type_name(enum ggml_type type);
int            ggml_blck_size(enum ggml_type type);
size_t         ggml_type_size(enum ggml_type type);
bool           ggml_is_quantized(enum ggml_type type);
float          ggml_fp16_to_fp32(ggml_fp16_t h);
ggml_fp16_t    ggml_fp32_to_fp16(float f);

#ifdef __cplusplus
}
#endif
