// compat/ggml-alloc.h -- the four allocator entry points main.cpp calls around its "measure pass"
// (/root/reference/examples/main/main.cpp:51-70; reference: ggml/include/ggml/ggml-alloc.h:11-29).
// On the device build the activation arena is sized by the engine from n_batch; these calls keep
// the reference's protocol (measure -> size -> allocate -> pass to every eval) and report the
// arena size the engine will use.
#pragma once
#include "ggml.h"

#ifdef __cplusplus
extern "C" {
#endif

struct ggml_allocr;
struct ggml_backend_buffer;

struct ggml_allocr * ggml_allocr_new_measure(size_t alignment);
struct ggml_allocr * ggml_allocr_new_from_buffer(struct ggml_backend_buffer * buffer);
size_t               ggml_allocr_alloc_graph(struct ggml_allocr * alloc, struct ggml_cgraph * graph);
void                 ggml_allocr_free(struct ggml_allocr * alloc);
bool                 ggml_allocr_is_measure(struct ggml_allocr * alloc);

#ifdef __cplusplus
}
#endif
