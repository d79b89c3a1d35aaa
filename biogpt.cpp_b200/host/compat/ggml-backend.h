// compat/ggml-backend.h -- backend / buffer handles of the reference's calling convention
// (ggml/include/ggml/ggml-backend.h:14-59, 98-113).  One backend exists: the B200 engine.  There
// is no CPU backend and no dispatch.
#pragma once
#include "ggml.h"
#include "ggml-alloc.h"

#ifdef __cplusplus
extern "C" {
#endif

struct ggml_backend;
typedef struct ggml_backend * ggml_backend_t;
typedef struct ggml_backend_buffer * ggml_backend_buffer_t;

ggml_backend_t        ggml_backend_b200_init(int device);        // fails (NULL) without a CUDA device
const char *          ggml_backend_name(ggml_backend_t backend);
void                  ggml_backend_free(ggml_backend_t backend);
size_t                ggml_backend_get_alignment(ggml_backend_t backend);
ggml_backend_buffer_t ggml_backend_alloc_buffer(ggml_backend_t backend, size_t size);
void                  ggml_backend_buffer_free(ggml_backend_buffer_t buffer);
size_t                ggml_backend_buffer_get_size(ggml_backend_buffer_t buffer);

#ifdef __cplusplus
}
#endif
