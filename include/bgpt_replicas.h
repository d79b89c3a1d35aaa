/* bgpt_replicas.h -- multi-GPU replica driver (C ABI, implemented in biogpt.cpp_b200/host/libbiogpt_b200.so; no torch, no NCCL).
 *
 * The reference serves one prompt per process (examples/main/main.cpp:93-151); BioGPT-base (<= 0.8 GB) fits any GPU many times
 * over and independent prompt streams share nothing, so scaling out is replication: every device holds a full copy
 * of the weights and the KV caches of its streams, stream s lives on device s % n_devices, one host thread drives each device,
 * and the streams of a device are decoded in lock step so its weights are read once per step (bgpt_cuda_eval_streams /
 * bgpt_cuda_decode_greedy_streams).  There is no inter-GPU dependency and no collective.
 *
 * Every stream computes exactly what a single-stream biogpt_eval computes (tests/test_replica_driver.py). */
#ifndef BGPT_REPLICAS_H
#define BGPT_REPLICAS_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct bgpt_replicas bgpt_replicas;

/* load `model_path` (ggml .bin, biogpt.cpp:27-453) on n_devices GPUs (0 = all visible) for n_streams sequences in total;
 * NULL on failure (bgpt_replicas_last_error) */
bgpt_replicas * bgpt_replicas_open(const char * model_path, int n_devices, int n_streams);
void bgpt_replicas_close(bgpt_replicas * r);
int  bgpt_replicas_devices(const bgpt_replicas * r);
int  bgpt_replicas_streams(const bgpt_replicas * r);
int  bgpt_replicas_n_vocab(const bgpt_replicas * r);
int  bgpt_replicas_device_of(const bgpt_replicas * r, int stream);          /* s % n_devices */
/* one lock-step token per stream at position n_past: tokens[n_streams] in, logits_out[n_streams][n_vocab] out (HOST buffers) */
int  bgpt_replicas_eval(bgpt_replicas * r, const int32_t * tokens, int n_past, float * logits_out);
/* the same step for a sampler: per stream the k largest (logit, id) pairs by logit descending, selected on the device
 * (bgpt_cuda_eval_streams_topk; biogpt.cpp:908-980) -- vals / ids [n_streams][k], n_out / exact [n_streams].  A stream with exact = 0
 * (equal logits make std::partial_sort's choice ambiguous) also gets its full row in logits_fallback[stream] ([n_streams][n_vocab],
 * may be NULL).  8 k + 16 bytes per stream cross PCIe instead of 4 n_vocab. */
int  bgpt_replicas_eval_topk(bgpt_replicas * r, const int32_t * tokens, int n_past, int k, float * vals, int32_t * ids, int * n_out, int * exact,
                             float * logits_fallback);
/* greedy continuation of every stream: first_tokens[n_streams] at n_past, n_steps tokens each; the devices run independently and
 * entirely device-side (argmax fed back on the GPU).  ids_out[n_steps][n_streams]; device_ms[n_devices] = CUDA-event time of each
 * device's loop (may be NULL).  Returns 0 or the first failing device's status. */
int  bgpt_replicas_decode_greedy(bgpt_replicas * r, const int32_t * first_tokens, int n_past, int n_steps,
                                 int32_t * ids_out, float * device_ms);
const char * bgpt_replicas_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
