// biogpt.h -- the reference's model API, re-declared for the B200 build.
//
// Source compatibility contract: /root/reference/examples/main/main.cpp and
// examples/quantize/quantize.cpp compile against THIS header and link against
// libbiogpt_b200.so without a single edit.  That fixes every name, field and signature below
// (reference: /root/reference/biogpt.h:13-172); what sits behind them is different:
//
//   * a loaded biogpt_model owns no host tensors.  `ctx` is the handle of a device engine
//     (csrc/libbgpt_cuda.so, include/bgpt_cuda.h) holding the re-tiled weights, the F32 KV cache
//     and the activation arena in HBM; the ggml_tensor pointers are descriptors (name/type/shape).
//   * biogpt_eval runs the fused CUDA schedule (one persistent kernel per decoded token, one kernel
//     per fused operator for prompt batches); n_threads is accepted and ignored.
//   * biogpt_graph / the ggml_allocr "measure pass" are kept as a protocol; the allocator reports
//     the size of the device arena for n_batch tokens.
//
// There is no CPU backend: biogpt_model_load fails if no CUDA device is present.
#pragma once

#include <algorithm>
#include <cstdint>
#include <fstream>
#include <iostream>
#include <map>
#include <random>
#include <string>
#include <thread>
#include <vector>

#include "bpe.h"
#include "ggml-backend.h"

// 'ggml' as a multi-character constant, the value convert.py writes (0x67676d6c)
#define BIOGPT_FILE_MAGIC   'ggml'

// raw little-endian POD I/O helpers used by the front ends when they copy file headers
template <typename T> static void read_safe(std::ifstream & infile, T & dest)   { infile.read((char *) &dest, sizeof(T)); }
template <typename T> static void write_safe(std::ofstream & outfile, T & dest) { outfile.write((char *) &dest, sizeof(T)); }

// ---- vocabulary -------------------------------------------------------------------------------
struct biogpt_vocab {
    using id    = int32_t;
    using token = std::string;

    int n_vocab  = 42384;
    int n_merges = 40000;

    std::map<token, id> token_to_id;
    std::map<id, token> id_to_token;
    std::map<word_pair, int> bpe_ranks;     // merge -> priority (lower merges first)
};
typedef std::vector<biogpt_vocab::id> token_sequence;

// ---- hyper-parameters (defaults = BioGPT-base; the file header overrides them) ---------------
struct biogpt_hparams {
    int32_t n_vocab     = 42384;
    int32_t n_merges    = 40000;
    int32_t d_ff        = 4096;
    int32_t d_model     = 1024;
    int32_t n_layer     = 24;
    int32_t n_head      = 16;
    int32_t n_positions = 1024;
    int32_t ftype       = 0;
};

// ---- weight handles ---------------------------------------------------------------------------
struct biogpt_layer_decoder {
    struct ggml_tensor * q_proj_w; struct ggml_tensor * k_proj_w; struct ggml_tensor * v_proj_w; struct ggml_tensor * o_proj_w;
    struct ggml_tensor * q_proj_b; struct ggml_tensor * k_proj_b; struct ggml_tensor * v_proj_b; struct ggml_tensor * o_proj_b;
    struct ggml_tensor * ln_0_w;   struct ggml_tensor * ln_1_w;   struct ggml_tensor * ln_0_b;   struct ggml_tensor * ln_1_b;
    struct ggml_tensor * fc_0_w;   struct ggml_tensor * fc_0_b;   struct ggml_tensor * fc_1_w;   struct ggml_tensor * fc_1_b;
};

struct biogpt_model {
    biogpt_hparams hparams;

    struct ggml_tensor * embed_tokens;
    struct ggml_tensor * embed_pos;
    struct ggml_tensor * ln_w;
    struct ggml_tensor * ln_b;
    struct ggml_tensor * lm_head;
    struct ggml_tensor * memory_k;          // F32 [n_layer * n_positions * d_model], resident in HBM
    struct ggml_tensor * memory_v;

    std::vector<biogpt_layer_decoder> layers_decoder;

    struct ggml_context * ctx;              // the device engine; released by ggml_free(ctx)
    std::map<std::string, struct ggml_tensor *> tensors;
    int n_loaded;

    ggml_backend_t backend = NULL;          // NULL on entry -> the loader opens the B200 backend
    ggml_backend_buffer_t buffer_w;
    ggml_backend_buffer_t buffer_kv;
};

// ---- command line -----------------------------------------------------------------------------
struct biogpt_params {
    int32_t seed      = -1;
    int32_t n_threads = std::min(4, (int32_t) std::thread::hardware_concurrency());   // unused on the GPU
    int32_t n_predict = 200;

    int32_t top_k = 40;
    float   top_p = 0.9f;
    float   temp  = 0.9f;

    uint8_t verbosity = 0;
    int32_t n_batch = 8;                    // prompt tokens per eval (un-masked attention inside a batch!)

    std::string model = "../ggml_weights/ggml-model.bin";
    std::string prompt;
    std::string lang;
};

// ---- API ----------------------------------------------------------------------------------------
// parses the `.bin`, uploads every tensor to the device; false + message on stderr on any error
bool biogpt_model_load(const std::string & fname, biogpt_model & model, biogpt_vocab & vocab, const uint8_t verbosity);

// stream-rewrites the tensors of an f32/f16 file as `ftype` blocks (header already copied by the caller)
void biogpt_model_quantize_internal(std::ifstream & fin, std::ofstream & fout, const ggml_ftype ftype);

// protocol placeholder for the reference's graph builder (see ggml-alloc.h)
struct ggml_cgraph * biogpt_graph(const biogpt_model & model, struct ggml_allocr * allocr, const token_sequence & embed_inp, const int n_past);

// one forward step: embed_inp at positions [n_past, n_past+N); logits of the LAST token (n_vocab floats)
bool biogpt_eval(const biogpt_model & model, const token_sequence & embed_inp, std::vector<float> & logits,
                 struct ggml_allocr * allocr, const int n_past, const int n_threads);

token_sequence gpt_tokenize(biogpt_vocab & vocab, const std::string & text, const std::string & lang);
std::string    gpt_decode(std::vector<std::string> & tokens, const std::string & lang);

biogpt_vocab::id biogpt_sample_top_k_top_p(const biogpt_vocab & vocab, const float * logits, int top_k, double top_p, double temp, std::mt19937 & rng);

// EXTENSION (not in the reference): biogpt_eval immediately followed by biogpt_sample_top_k_top_p on its logits, with the top_k
// selection done on the device -- the returned id is the one the two reference calls would return (same RNG draw), but only
// top_k (logit, id) pairs cross PCIe instead of the n_vocab logits.  Falls back to the two calls when
// top_k > 128 or when equal logits make std::partial_sort's choice ambiguous.  Returns -1 on error.
biogpt_vocab::id biogpt_eval_sample(const biogpt_model & model, const biogpt_vocab & vocab, const token_sequence & embed_inp,
                                    const int n_past, int top_k, double top_p, double temp, std::mt19937 & rng);

bool biogpt_params_parse(int argc, char ** argv, biogpt_params & params);
void biogpt_print_usage(char ** argv, const biogpt_params & params);
