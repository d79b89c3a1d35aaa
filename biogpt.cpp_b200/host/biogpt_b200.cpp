// biogpt_b200.cpp -- host side of the B200 build: the reference's C++ model API
// (biogpt.h) implemented on top of the extern "C" device engine (include/bgpt_cuda.h).
//
//   biogpt_model_load        reference biogpt.cpp:27-453   file walk is the same contract; tensors go
//                                                          to HBM through bgpt_cuda_upload_tensor
//   biogpt_eval              reference biogpt.cpp:812-847  -> bgpt_cuda_eval
//   biogpt_graph + allocr    reference biogpt.cpp:624-810, main.cpp:51-70  protocol only
//   biogpt_sample_top_k_top_p reference biogpt.cpp:908-980 same libstdc++ RNG draw, host side
//   gpt_tokenize / gpt_decode reference biogpt.cpp:850-906
//   biogpt_model_quantize_internal reference biogpt.cpp:459-621, block codecs ggml.c:892-1094
//   biogpt_params_parse / biogpt_print_usage  reference biogpt.cpp:982-1040
//
// Error behaviour mirrors the reference: load returns false with a message on stderr, nothing
// throws on the eval path, the quantizer throws std::runtime_error.
#include "biogpt.h"
#include "mosestokenizer.h"
#include "../../include/bgpt_cuda.h"

#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#if defined(__AVX2__)
#include <immintrin.h>
#endif
#include <regex>
#include <sstream>
#include <stdexcept>

// ------------------------------------------------------------------------------------------------
// opaque handles of the compat headers
// ------------------------------------------------------------------------------------------------
struct ggml_context {
    bgpt_model * engine = nullptr;
    std::vector<ggml_tensor *> descriptors;       // owned
};
struct ggml_backend { int device = 0; };
struct ggml_backend_buffer { size_t size = 0; };
struct ggml_cgraph { int n_tokens = 0; int n_past = 0; size_t arena_bytes = 0; };
struct ggml_allocr { bool measure = false; size_t last_size = 0; };

// the device engine behind a loaded model (replicas.cpp drives it directly for multi-stream calls)
bgpt_model * bgpt_host_engine_of(const biogpt_model & model) { return model.ctx ? model.ctx->engine : nullptr; }

static std::chrono::steady_clock::time_point g_t0;

extern "C" {

void ggml_time_init(void) { g_t0 = std::chrono::steady_clock::now(); }
int64_t ggml_time_us(void) { return std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - g_t0).count(); }
int64_t ggml_time_ms(void) { return ggml_time_us() / 1000; }

void ggml_free(struct ggml_context * ctx) {
    if (!ctx) return;
    bgpt_cuda_model_free(ctx->engine);
    for (ggml_tensor * t : ctx->descriptors) delete t;
    delete ctx;
}

enum ggml_type ggml_ftype_to_ggml_type(enum ggml_ftype ftype) {
    switch (ftype) {
        case GGML_FTYPE_ALL_F32:     return GGML_TYPE_F32;
        case GGML_FTYPE_MOSTLY_F16:  return GGML_TYPE_F16;
        case GGML_FTYPE_MOSTLY_Q4_0: return GGML_TYPE_Q4_0;
        case GGML_FTYPE_MOSTLY_Q4_1: return GGML_TYPE_Q4_1;
        case GGML_FTYPE_MOSTLY_Q5_0: return GGML_TYPE_Q5_0;
        case GGML_FTYPE_MOSTLY_Q5_1: return GGML_TYPE_Q5_1;
        case GGML_FTYPE_MOSTLY_Q8_0: return GGML_TYPE_Q8_0;
        default: return GGML_TYPE_COUNT;
    }
}
const char * ggml_type_name(enum ggml_type t) {
    switch (t) {
        case GGML_TYPE_F32: return "f32"; case GGML_TYPE_F16: return "f16"; case GGML_TYPE_Q4_0: return "q4_0";
        case GGML_TYPE_Q4_1: return "q4_1"; case GGML_TYPE_Q5_0: return "q5_0"; case GGML_TYPE_Q5_1: return "q5_1";
        case GGML_TYPE_Q8_0: return "q8_0"; case GGML_TYPE_Q8_1: return "q8_1"; default: return "?";
    }
}
int ggml_blck_size(enum ggml_type t) { return (t == GGML_TYPE_F32 || t == GGML_TYPE_F16) ? 1 : 32; }
size_t ggml_type_size(enum ggml_type t) {
    switch (t) {
        case GGML_TYPE_F32: return 4; case GGML_TYPE_F16: return 2; case GGML_TYPE_Q4_0: return 18; case GGML_TYPE_Q4_1: return 20;
        case GGML_TYPE_Q5_0: return 22; case GGML_TYPE_Q5_1: return 24; case GGML_TYPE_Q8_0: return 34; case GGML_TYPE_Q8_1: return 40;
        default: return 0;
    }
}
bool ggml_is_quantized(enum ggml_type t) { return ggml_blck_size(t) == 32; }

// IEEE binary16 <-> binary32, round to nearest even (software; identical to F16C results)
float ggml_fp16_to_fp32(ggml_fp16_t h) {
    const uint32_t sign = (uint32_t) (h & 0x8000u) << 16;
    uint32_t exp = (h >> 10) & 0x1Fu, man = h & 0x3FFu, bits;
    if (exp == 0) {
        if (man == 0) bits = sign;
        else { int e = -1; do { man <<= 1; e++; } while (!(man & 0x400u)); bits = sign | ((uint32_t) (127 - 15 - e) << 23) | ((man & 0x3FFu) << 13); }
    } else if (exp == 31) bits = sign | 0x7F800000u | (man << 13);
    else bits = sign | ((exp + 112u) << 23) | (man << 13);
    float f; memcpy(&f, &bits, 4); return f;
}
ggml_fp16_t ggml_fp32_to_fp16(float f) {
    uint32_t x; memcpy(&x, &f, 4);
    const uint32_t sign = (x >> 16) & 0x8000u;
    const uint32_t ax = x & 0x7FFFFFFFu;
    if (ax >= 0x7F800000u) return (ggml_fp16_t) (sign | 0x7C00u | (ax > 0x7F800000u ? 0x200u | ((ax >> 13) & 0x3FFu) : 0u));
    if (ax >= 0x477FF000u) return (ggml_fp16_t) (sign | 0x7C00u);                 // rounds to infinity
    if (ax < 0x33000001u) return (ggml_fp16_t) sign;                               // rounds to zero
    int e = (int) (ax >> 23) - 127;
    uint32_t man = (ax & 0x7FFFFFu) | 0x800000u;
    int shift; uint32_t base;
    if (e < -14) { shift = 13 + (-14 - e); base = 0; }                             // subnormal half
    else { shift = 13; base = (uint32_t) (e + 15) << 10; man &= 0x7FFFFFu; }
    uint32_t q = man >> shift;
    const uint32_t rem = man & ((1u << shift) - 1u), half = 1u << (shift - 1);
    if (rem > half || (rem == half && (q & 1u))) q++;
    return (ggml_fp16_t) (sign | (base + q));
}

struct ggml_allocr * ggml_allocr_new_measure(size_t) { ggml_allocr * a = new ggml_allocr(); a->measure = true; return a; }
struct ggml_allocr * ggml_allocr_new_from_buffer(struct ggml_backend_buffer * b) { ggml_allocr * a = new ggml_allocr(); a->last_size = b ? b->size : 0; return a; }
size_t ggml_allocr_alloc_graph(struct ggml_allocr * a, struct ggml_cgraph * g) { if (a && g) a->last_size = g->arena_bytes; return g ? g->arena_bytes : 0; }
void ggml_allocr_free(struct ggml_allocr * a) { delete a; }
bool ggml_allocr_is_measure(struct ggml_allocr * a) { return a && a->measure; }

ggml_backend_t ggml_backend_b200_init(int device) {
    const int n = bgpt_cuda_device_count();
    if (n <= 0 || device < 0 || device >= n) return nullptr;
    ggml_backend * b = new ggml_backend(); b->device = device; return b;
}
const char * ggml_backend_name(ggml_backend_t) { return "B200"; }
void ggml_backend_free(ggml_backend_t b) { delete b; }
size_t ggml_backend_get_alignment(ggml_backend_t) { return 256; }
ggml_backend_buffer_t ggml_backend_alloc_buffer(ggml_backend_t, size_t size) { ggml_backend_buffer * b = new ggml_backend_buffer(); b->size = size; return b; }
void ggml_backend_buffer_free(ggml_backend_buffer_t b) { delete b; }
size_t ggml_backend_buffer_get_size(ggml_backend_buffer_t b) { return b ? b->size : 0; }

}  // extern "C"

// ------------------------------------------------------------------------------------------------
// loader
// ------------------------------------------------------------------------------------------------
static ggml_tensor * new_descriptor(ggml_context * ctx, const std::string & name, ggml_type type, int64_t ne0, int64_t ne1) {
    ggml_tensor * t = new ggml_tensor();
    t->type = type; t->ne[0] = ne0; t->ne[1] = ne1; t->ne[2] = t->ne[3] = 1;
    snprintf(t->name, sizeof t->name, "%s", name.c_str());
    ctx->descriptors.push_back(t);
    return t;
}

bool biogpt_model_load(const std::string & fname, biogpt_model & model, biogpt_vocab & vocab, const uint8_t verbosity) {
    fprintf(stderr, "%s: loading model from '%s'\n", __func__, fname.c_str());
    std::ifstream in(fname, std::ios::binary);
    if (!in) { fprintf(stderr, "%s: failed to open '%s'\n", __func__, fname.c_str()); return false; }

    uint32_t magic = 0;
    read_safe(in, magic);
    if (magic != BIOGPT_FILE_MAGIC) { fprintf(stderr, "%s: invalid model file '%s' (bad magic)\n", __func__, fname.c_str()); return false; }

    biogpt_hparams & hp = model.hparams;
    read_safe(in, hp.n_vocab); read_safe(in, hp.n_layer); read_safe(in, hp.n_head); read_safe(in, hp.n_positions);
    read_safe(in, hp.d_ff);    read_safe(in, hp.d_model); read_safe(in, hp.ftype);
    fprintf(stderr, "%s: n_vocab       = %d\n", __func__, hp.n_vocab);
    fprintf(stderr, "%s: d_ff          = %d\n", __func__, hp.d_ff);
    fprintf(stderr, "%s: d_model       = %d\n", __func__, hp.d_model);
    fprintf(stderr, "%s: n_positions   = %d\n", __func__, hp.n_positions);
    fprintf(stderr, "%s: n_head        = %d\n", __func__, hp.n_head);
    fprintf(stderr, "%s: n_layer       = %d\n", __func__, hp.n_layer);
    fprintf(stderr, "%s: ftype         = %d\n", __func__, hp.ftype);

    // vocabulary: i32 count, then { u32 len, bytes }
    int32_t n_vocab = 0;
    read_safe(in, n_vocab);
    if (n_vocab != hp.n_vocab) {
        fprintf(stderr, "%s: invalid model file '%s' (bad vocab size %d != %d)\n", __func__, fname.c_str(), n_vocab, hp.n_vocab);
        return false;
    }
    // a corrupt or truncated file must end in `return false`, not in a multi-GB resize or a bad_alloc: sizes read from the file
    // are checked before they are used
    const uint32_t MAX_STR = 1u << 16;
    if (hp.n_vocab <= 0 || hp.n_layer <= 0 || hp.n_head <= 0 || hp.n_positions <= 0 || hp.d_ff <= 0 || hp.d_model <= 0) {
        fprintf(stderr, "%s: invalid model file '%s' (bad hyper-parameters)\n", __func__, fname.c_str());
        return false;
    }
    std::string word;
    for (int i = 0; i < n_vocab; i++) {
        uint32_t len = 0; read_safe(in, len);
        if (!in.good() || len > MAX_STR) { fprintf(stderr, "%s: invalid model file '%s' (bad vocabulary entry %d)\n", __func__, fname.c_str(), i); return false; }
        word.resize(len);
        if (len) in.read(&word[0], len);
        if (!in.good()) { fprintf(stderr, "%s: model file '%s' is truncated in the vocabulary\n", __func__, fname.c_str()); return false; }
        vocab.token_to_id[word] = i;
        vocab.id_to_token[i] = word;
    }
    vocab.n_vocab = hp.n_vocab;

    // merges: i32 count (must be hparams.n_merges), then { u32 len, "left right" }
    int32_t n_merges = 0;
    read_safe(in, n_merges);
    if (n_merges != hp.n_merges) {
        fprintf(stderr, "%s: invalid model file '%s' (bad merge size %d != %d)\n", __func__, fname.c_str(), n_merges, hp.n_merges);
        return false;
    }
    word_pair last_pair;
    for (int i = 0; i < n_merges; i++) {
        uint32_t len = 0; read_safe(in, len);
        if (!in.good() || len > MAX_STR) { fprintf(stderr, "%s: invalid model file '%s' (bad merge entry %d)\n", __func__, fname.c_str(), i); return false; }
        if (len) {
            word.resize(len); in.read(&word[0], len);
            if (!in.good()) { fprintf(stderr, "%s: model file '%s' is truncated in the merges\n", __func__, fname.c_str()); return false; }
            std::stringstream ss(word);
            ss >> last_pair.first >> last_pair.second;
        }
        vocab.bpe_ranks[last_pair] = i;      // an empty entry re-ranks the previous pair, like the reference
    }
    vocab.n_merges = hp.n_merges;

    const ggml_type wtype = ggml_ftype_to_ggml_type((ggml_ftype) hp.ftype);
    if (wtype == GGML_TYPE_COUNT) {
        fprintf(stderr, "%s: invalid model file '%s' (bad ftype value %d)\n", __func__, fname.c_str(), hp.ftype);
        return false;
    }

    // the one backend: the B200 engine
    if (model.backend == NULL) {
        int device = 0;
        if (const char * e = getenv("BIOGPT_CUDA_DEVICE")) device = atoi(e);
        fprintf(stderr, "%s: using B200 CUDA backend (device %d)\n", __func__, device);
        model.backend = ggml_backend_b200_init(device);
    }
    if (!model.backend) {
        fprintf(stderr, "%s: no CUDA device available: %s (this build has no CPU path)\n", __func__, bgpt_cuda_last_error());
        return false;
    }
    int n_batch = 8;
    if (const char * e = getenv("BIOGPT_MAX_BATCH")) n_batch = std::max(1, atoi(e));
    const int32_t hp7[7] = { hp.n_vocab, hp.n_layer, hp.n_head, hp.n_positions, hp.d_ff, hp.d_model, hp.ftype };
    bgpt_model * engine = bgpt_cuda_model_create(hp7, model.backend->device, n_batch);
    if (!engine) { fprintf(stderr, "%s: %s\n", __func__, bgpt_cuda_last_error()); return false; }
    ggml_context * ctx = new ggml_context();
    ctx->engine = engine;
    model.ctx = ctx;
    model.layers_decoder.resize(hp.n_layer);

    // tensor records: { i32 n_dims, i32 name_len, i32 type, i32 ne[n_dims], name, raw bytes }
    size_t total = 0;
    model.n_loaded = 0;
    std::vector<char> raw;
    while (true) {
        int32_t n_dims = 0, name_len = 0, ttype = 0;
        read_safe(in, n_dims); read_safe(in, name_len); read_safe(in, ttype);
        if (in.eof()) break;
        if (n_dims < 1 || n_dims > 2 || name_len <= 0 || name_len > 255) { fprintf(stderr, "%s: corrupt tensor header in '%s'\n", __func__, fname.c_str()); ggml_free(ctx); return false; }
        int32_t ne[2] = { 1, 1 };
        for (int i = 0; i < n_dims; i++) read_safe(in, ne[i]);
        std::string name(name_len, 0);
        in.read(&name[0], name_len);
        if (!in.good() || ne[0] <= 0 || ne[1] <= 0 || (int64_t) ne[0] * ne[1] > ((int64_t) 1 << 31)) {
            fprintf(stderr, "%s: tensor '%s' has a bad shape [%d, %d] in '%s'\n", __func__, name.c_str(), ne[0], ne[1], fname.c_str()); ggml_free(ctx); return false;
        }
        if (ggml_type_size((ggml_type) ttype) == 0 || ttype == GGML_TYPE_Q8_1) { fprintf(stderr, "%s: tensor '%s' has unknown type %d\n", __func__, name.c_str(), ttype); ggml_free(ctx); return false; }
        const size_t nbytes = (size_t) ne[0] * ne[1] / ggml_blck_size((ggml_type) ttype) * ggml_type_size((ggml_type) ttype);
        raw.resize(nbytes);
        in.read(raw.data(), nbytes);
        if ((size_t) in.gcount() != nbytes) { fprintf(stderr, "%s: tensor '%s' is truncated in model file\n", __func__, name.c_str()); ggml_free(ctx); return false; }
        if (bgpt_cuda_upload_tensor(engine, name.c_str(), ttype, ne[0], ne[1], raw.data(), nbytes) != BGPT_OK) {
            fprintf(stderr, "%s: %s\n", __func__, bgpt_cuda_last_error()); ggml_free(ctx); return false;
        }
        model.tensors[name] = new_descriptor(ctx, name, (ggml_type) ttype, ne[0], ne[1]);
        if (verbosity > 0) printf("%48s - [%5d, %5d], type = %6s, %6.2f MB\n", name.c_str(), ne[0], ne[1], ggml_type_name((ggml_type) ttype), nbytes / 1024.0 / 1024.0);
        total += nbytes;
        model.n_loaded++;
    }
    fprintf(stderr, "%s: model size = %8.2f MB\n", __func__, total / 1024.0 / 1024.0);

    const int expected = 5 + 16 * hp.n_layer;
    if (model.n_loaded != expected) {
        fprintf(stderr, "%s: not all tensors loaded from model file - expected %d, got %d\n", __func__, expected, model.n_loaded);
        ggml_free(ctx); return false;
    }
    // descriptors under the reference's field names
    auto T = [&](const std::string & n) { auto it = model.tensors.find(n); return it == model.tensors.end() ? (ggml_tensor *) nullptr : it->second; };
    model.embed_tokens = T("biogpt.embed_tokens.weight"); model.embed_pos = T("biogpt.embed_positions.weight");
    model.ln_w = T("biogpt.layer_norm.weight"); model.ln_b = T("biogpt.layer_norm.bias"); model.lm_head = T("output_projection.weight");
    for (int i = 0; i < hp.n_layer; i++) {
        const std::string p = "biogpt.layers." + std::to_string(i) + ".";
        biogpt_layer_decoder & L = model.layers_decoder[i];
        L.q_proj_w = T(p + "self_attn.q_proj.weight"); L.k_proj_w = T(p + "self_attn.k_proj.weight");
        L.v_proj_w = T(p + "self_attn.v_proj.weight"); L.o_proj_w = T(p + "self_attn.out_proj.weight");
        L.q_proj_b = T(p + "self_attn.q_proj.bias");   L.k_proj_b = T(p + "self_attn.k_proj.bias");
        L.v_proj_b = T(p + "self_attn.v_proj.bias");   L.o_proj_b = T(p + "self_attn.out_proj.bias");
        L.ln_0_w = T(p + "self_attn_layer_norm.weight"); L.ln_0_b = T(p + "self_attn_layer_norm.bias");
        L.ln_1_w = T(p + "final_layer_norm.weight");     L.ln_1_b = T(p + "final_layer_norm.bias");
        L.fc_0_w = T(p + "fc1.weight"); L.fc_0_b = T(p + "fc1.bias"); L.fc_1_w = T(p + "fc2.weight"); L.fc_1_b = T(p + "fc2.bias");
    }
    const int64_t nkv = (int64_t) hp.n_layer * hp.n_positions * hp.d_model;
    model.memory_k = new_descriptor(ctx, "memory_k", GGML_TYPE_F32, nkv, 1);
    model.memory_v = new_descriptor(ctx, "memory_v", GGML_TYPE_F32, nkv, 1);
    model.buffer_w  = ggml_backend_alloc_buffer(model.backend, bgpt_cuda_weight_bytes(engine));
    model.buffer_kv = ggml_backend_alloc_buffer(model.backend, (size_t) nkv * 4 * 2);

    // ggml's fp16 lookup tables, built with this machine's libm exactly as ggml_init would
    std::vector<uint16_t> gelu(65536), ex(65536);
    bgpt_host_build_tables(gelu.data(), ex.data());
    if (bgpt_cuda_set_tables(engine, gelu.data(), ex.data()) != BGPT_OK || bgpt_cuda_model_finalize(engine) != BGPT_OK) {
        fprintf(stderr, "%s: %s\n", __func__, bgpt_cuda_last_error());
        ggml_backend_buffer_free(model.buffer_w); ggml_backend_buffer_free(model.buffer_kv);
        ggml_free(ctx); return false;
    }
    return true;
}

// ------------------------------------------------------------------------------------------------
// graph placeholder + eval
// ------------------------------------------------------------------------------------------------
struct ggml_cgraph * biogpt_graph(const biogpt_model & model, struct ggml_allocr *, const token_sequence & embed_inp, const int n_past) {
    static ggml_cgraph g;          // the reference's builder is not re-entrant either (static arena)
    const biogpt_hparams & hp = model.hparams;
    g.n_tokens = (int) embed_inp.size(); g.n_past = n_past;
    // device arena for N token rows: x, x1, q, att (d_model), hff (d_ff), logits (n_vocab), f32
    g.arena_bytes = (size_t) g.n_tokens * ((size_t) 4 * hp.d_model + hp.d_ff + hp.n_vocab) * sizeof(float);
    return &g;
}

bool biogpt_eval(const biogpt_model & model, const token_sequence & embed_inp, std::vector<float> & logits,
                 struct ggml_allocr *, const int n_past, const int /*n_threads*/) {
    if (!model.ctx || !model.ctx->engine || embed_inp.empty()) return false;
    logits.resize(model.hparams.n_vocab);
    const int rc = bgpt_cuda_eval(model.ctx->engine, embed_inp.data(), (int) embed_inp.size(), n_past, logits.data());
    if (rc != BGPT_OK) { fprintf(stderr, "%s: %s\n", __func__, bgpt_cuda_last_error()); return false; }
    return true;
}

// ------------------------------------------------------------------------------------------------
// text in / out
// ------------------------------------------------------------------------------------------------
token_sequence gpt_tokenize(biogpt_vocab & vocab, const std::string & text, const std::string & lang) {
    token_sequence ids;
    ids.push_back(2);                                   // "</s>" opens the sequence
    for (const std::string & w : moses_tokenize(text, lang)) {
        std::stringstream pieces(bpe(w, vocab.bpe_ranks));
        std::string piece;
        while (pieces >> piece) {
            auto it = vocab.token_to_id.find(piece);
            if (it != vocab.token_to_id.end()) ids.push_back(it->second);
            else fprintf(stderr, "%s: unknown token '%s'\n", __func__, piece.c_str());
        }
    }
    return ids;
}

std::string gpt_decode(std::vector<std::string> & tokens, const std::string & lang) {
    std::string joined;
    for (std::string & t : tokens) {
        std::string s;
        for (char c : t) if (c != ' ') s += c;          // BPE pieces carry no inner spaces
        size_t pos;
        while ((pos = s.find("</w>")) != std::string::npos) s.replace(pos, 4, " ");
        while ((pos = s.find("</s>")) != std::string::npos) s.replace(pos, 4, " ");
        t = s;
        joined += s;
    }
    std::vector<std::string> words;
    std::stringstream ss(joined);
    for (std::string w; ss >> w; ) words.push_back(w);
    return moses_detokenize(words, lang);
}

// ------------------------------------------------------------------------------------------------
// sampler: identical draw to the reference for the same rng state (std::discrete_distribution)
// ------------------------------------------------------------------------------------------------
// softmax over the sorted candidates, top-p cut and the draw: the tail of the reference function, shared by both entry points
static biogpt_vocab::id draw_from_sorted(std::vector<std::pair<double, biogpt_vocab::id>> & cand, int top_k, double top_p, std::mt19937 & rng) {
    double maxl = -INFINITY;
    for (const auto & c : cand) maxl = std::max(maxl, c.first);
    std::vector<double> probs;
    probs.reserve(cand.size());
    double sum = 0.0;
    for (const auto & c : cand) { const double p = exp(c.first - maxl); probs.push_back(p); sum += p; }
    for (double & p : probs) p /= sum;

    if (top_p < 1.0f) {
        double cum = 0.0f;
        for (int i = 0; i < top_k; i++) {
            cum += probs[i];
            if (cum >= top_p) { top_k = i + 1; probs.resize(top_k); cand.resize(top_k); break; }
        }
        cum = 1.0 / cum;
        for (double & p : probs) p *= cum;
    }
    std::discrete_distribution<> dist(probs.begin(), probs.end());
    return cand[dist(rng)].second;
}

// The reference's selection, as written: n_vocab (logit / temp, id) pairs in doubles, std::partial_sort (biogpt.cpp:908-940).  210-380 us
// for 42384 logits on the bench host -- as long as the whole forward pass on the GPU.
static biogpt_vocab::id sample_reference_order(int n_logits, const float * logits, int top_k, double top_p, double temp, std::mt19937 & rng) {
    std::vector<std::pair<double, biogpt_vocab::id>> cand;
    cand.reserve(n_logits);
    const double inv_temp = 1.0 / temp;
    for (int i = 0; i < n_logits; i++) cand.emplace_back(logits[i] * inv_temp, i);
    std::partial_sort(cand.begin(), cand.begin() + top_k, cand.end(),
                      [](const std::pair<double, biogpt_vocab::id> & a, const std::pair<double, biogpt_vocab::id> & b) { return a.first > b.first; });
    cand.resize(top_k);
    return draw_from_sorted(cand, top_k, top_p, rng);
}

// The same pairs without sorting the vocabulary.  For temp > 0, logit -> logit * (1 / temp) in double is strictly increasing on
// floats (a float's 24 bits times a double keeps neighbours apart), so the top_k pairs by scaled value are the top_k logits: one
// pass keeps the top_k + 1 largest floats in a min-heap -- an AVX2 compare of 8 logits against the heap's minimum rejects almost
// every group at once -- and only the survivors are scaled.  std::partial_sort leaves the choice and order among EQUAL values
// unspecified: if two neighbours among the top_k + 1 scaled values are equal (or a logit is NaN, or temp is not a positive finite
// number), the reference's own code above decides, so the drawn id is the reference's in every case
// (tests/test_host_lib.py: test_sampler_draw_identical_to_reference, test_sampler_fast_selection_*).  ~10 us instead of 210-380.
biogpt_vocab::id biogpt_sample_top_k_top_p(const biogpt_vocab & vocab, const float * logits, int top_k, double top_p, double temp, std::mt19937 & rng) {
    const int n_logits = (int) vocab.id_to_token.size();
    top_k = std::max(1, std::min(top_k, n_logits));
    const int kc = top_k + 1;
    if (!(temp > 0.0) || !std::isfinite(temp) || kc > 512 || kc > n_logits) return sample_reference_order(n_logits, logits, top_k, top_p, temp, rng);
    struct Ent { float v; int i; };
    Ent heap[512];
    int hn = 0;
    auto cmp = [](const Ent & a, const Ent & b) { return a.v > b.v; };          // std::*_heap with this comparator: heap[0] is the minimum
    bool has_nan = false;
    // A lower bound of the kc-th largest logit first, so that the selecting pass below rejects almost everything with one compare:
    // the kc-th largest of the maxima of 256-element chunks (kc chunks hold a value >= it).  Without it the heap is rebuilt
    // ~ kc ln(n / kc) times (300 x 50 ns for 42384 logits) while its minimum climbs.
    float bound = -INFINITY;
    const int CH = 256;
    int nch = 0;
    static thread_local std::vector<float> cmax, csel;
#if defined(__AVX2__)
    if (n_logits / CH >= kc) {
        nch = n_logits / CH;
        cmax.resize(nch);
        __m256 unord = _mm256_setzero_ps();
        for (int c = 0; c < nch; c++) {
            const float * p = logits + (size_t) c * CH;
            __m256 m0 = _mm256_loadu_ps(p), m1 = _mm256_loadu_ps(p + 8);
            __m256 u0 = _mm256_cmp_ps(m0, m1, _CMP_UNORD_Q);
            for (int j = 16; j < CH; j += 16) {
                const __m256 a = _mm256_loadu_ps(p + j), b2 = _mm256_loadu_ps(p + j + 8);
                u0 = _mm256_or_ps(u0, _mm256_cmp_ps(a, b2, _CMP_UNORD_Q));
                m0 = _mm256_max_ps(m0, a); m1 = _mm256_max_ps(m1, b2);
            }
            unord = _mm256_or_ps(unord, u0);
            m0 = _mm256_max_ps(m0, m1);
            __m128 h = _mm_max_ps(_mm256_castps256_ps128(m0), _mm256_extractf128_ps(m0, 1));
            h = _mm_max_ps(h, _mm_movehl_ps(h, h));
            h = _mm_max_ss(h, _mm_shuffle_ps(h, h, 1));
            cmax[c] = _mm_cvtss_f32(h);
        }
        has_nan = _mm256_movemask_ps(unord) != 0;
        csel = cmax;
        std::nth_element(csel.begin(), csel.begin() + (kc - 1), csel.end(), [](float a, float b) { return a > b; });
        bound = csel[kc - 1];
    }
#endif
    // the heap starts with the first kc logits that reach the bound (at least kc do); chunks whose maximum is below the bound are not
    // read again, everything else below the bound is rejected by one vector compare per 8 logits
    auto offer = [&](int j) {
        const float x = logits[j];
        if (hn < kc) { heap[hn++] = Ent{ x, j }; if (hn == kc) std::make_heap(heap, heap + hn, cmp); }
        else if (x > heap[0].v) { std::pop_heap(heap, heap + kc, cmp); heap[kc - 1] = Ent{ x, j }; std::push_heap(heap, heap + kc, cmp); }
    };
    auto scan = [&](int lo, int hi) {
        int i = lo;
#if defined(__AVX2__)
        for (; i + 8 <= hi; i += 8) {
            const __m256 x = _mm256_loadu_ps(logits + i);
            int m = hn < kc ? _mm256_movemask_ps(_mm256_cmp_ps(x, _mm256_set1_ps(bound), _CMP_GE_OQ))
                            : _mm256_movemask_ps(_mm256_cmp_ps(x, _mm256_set1_ps(heap[0].v), _CMP_GT_OQ));
            while (m) { const int b = __builtin_ctz(m); m &= m - 1; offer(i + b); }
        }
#endif
        for (; i < hi; i++) { const float x = logits[i]; if (hn < kc ? x >= bound : x > heap[0].v) offer(i); }
    };
    for (int c = 0; c < nch; c++) if (cmax[c] >= bound) scan(c * CH, (c + 1) * CH);
    for (int i = nch * CH; i < n_logits; i++) has_nan = has_nan || logits[i] != logits[i];
    scan(nch * CH, n_logits);
    if (hn < kc) return sample_reference_order(n_logits, logits, top_k, top_p, temp, rng);      // (only with NaNs around)
    if (has_nan) return sample_reference_order(n_logits, logits, top_k, top_p, temp, rng);
    std::sort(heap, heap + kc, [](const Ent & a, const Ent & b) { return a.v > b.v; });
    const double inv_temp = 1.0 / temp;
    std::vector<std::pair<double, biogpt_vocab::id>> cand;
    cand.reserve(kc);
    for (int j = 0; j < kc; j++) cand.emplace_back(heap[j].v * inv_temp, heap[j].i);
    for (int j = 0; j + 1 < kc; j++)
        if (!(cand[j].first > cand[j + 1].first)) return sample_reference_order(n_logits, logits, top_k, top_p, temp, rng);
    cand.resize(top_k);
    return draw_from_sorted(cand, top_k, top_p, rng);
}

// test door (host_capi.cpp): the reference's selection as written, on a vocabulary of any size
biogpt_vocab::id bgpt_sample_reference_order(int n_logits, const float * logits, int top_k, double top_p, double temp, std::mt19937 & rng) {
    return sample_reference_order(n_logits, logits, std::max(1, std::min(top_k, n_logits)), top_p, temp, rng);
}

biogpt_vocab::id biogpt_eval_sample(const biogpt_model & model, const biogpt_vocab & vocab, const token_sequence & embed_inp,
                                    const int n_past, int top_k, double top_p, double temp, std::mt19937 & rng) {
    if (!model.ctx || !model.ctx->engine || embed_inp.empty()) return -1;
    const int n_vocab = model.hparams.n_vocab;
    top_k = std::max(1, std::min(top_k, n_vocab));
    static thread_local std::vector<float> full;
    if (top_k > 128 || !(temp > 0.0)) {                            // outside the device selector's range: the reference's two calls
        if (!biogpt_eval(model, embed_inp, full, nullptr, n_past, 0)) return -1;
        return biogpt_sample_top_k_top_p(vocab, full.data(), top_k, top_p, temp, rng);
    }
    float vals[128]; int32_t ids[128]; int n_out = 0, exact = 0;
    full.resize(n_vocab);
    const int rc = bgpt_cuda_eval_topk(model.ctx->engine, embed_inp.data(), (int) embed_inp.size(), n_past, top_k, vals, ids, &n_out, &exact, full.data());
    if (rc != BGPT_OK) { fprintf(stderr, "%s: %s\n", __func__, bgpt_cuda_last_error()); return -1; }
    if (!exact) return biogpt_sample_top_k_top_p(vocab, full.data(), top_k, top_p, temp, rng);      // tied logits: the full path decides
    // the device returned the top_k pairs by logit descending; logit * (1 / temp) in double keeps that order (temp > 0)
    std::vector<std::pair<double, biogpt_vocab::id>> cand;
    cand.reserve(n_out);
    const double inv_temp = 1.0 / temp;
    for (int i = 0; i < n_out; i++) cand.emplace_back(vals[i] * inv_temp, ids[i]);
    return draw_from_sorted(cand, n_out, top_p, rng);
}

// ------------------------------------------------------------------------------------------------
// quantize tool: f32/f16 `.bin` tensors -> block formats (weight quantisers of ggml.c:892-1094)
// ------------------------------------------------------------------------------------------------
namespace {
inline void put_h(uint8_t * p, float f) { const ggml_fp16_t h = ggml_fp32_to_fp16(f); memcpy(p, &h, 2); }

// x: 32 floats -> one block at `out`; arithmetic order follows quantize_row_q*_reference as
// compiled in the reference build (x*id + c is one fused multiply-add there)
void quant_block(ggml_type t, const float * x, uint8_t * out) {
    if (t == GGML_TYPE_Q8_0) {
        float amax = 0.f;
        for (int j = 0; j < 32; j++) amax = std::max(amax, fabsf(x[j]));
        const float d = amax / 127.f, id = d ? 1.f / d : 0.f;
        put_h(out, d);
        for (int j = 0; j < 32; j++) out[2 + j] = (uint8_t) (int8_t) roundf(x[j] * id);
        return;
    }
    const bool sym = (t == GGML_TYPE_Q4_0 || t == GGML_TYPE_Q5_0);
    const bool five = (t == GGML_TYPE_Q5_0 || t == GGML_TYPE_Q5_1);
    const int levels = five ? 31 : 15;
    float d, base, add;
    if (sym) {          // d = (signed value of largest magnitude) / -(levels+1)/2
        float amax = 0.f, vmax = 0.f;
        for (int j = 0; j < 32; j++) if (amax < fabsf(x[j])) { amax = fabsf(x[j]); vmax = x[j]; }
        d = vmax / (five ? -16.f : -8.f); base = 0.f; add = five ? 16.5f : 8.5f;
    } else {            // d = (max - min) / levels, codes relative to min
        float mn = x[0], mx = x[0];
        for (int j = 1; j < 32; j++) { mn = std::min(mn, x[j]); mx = std::max(mx, x[j]); }
        d = (mx - mn) / (float) levels; base = mn; add = 0.5f;
    }
    const float id = d ? 1.f / d : 0.f;
    uint8_t q[32];
    for (int j = 0; j < 32; j++) {
        const float v = fmaf(x[j] - base, id, add);
        const int c = (t == GGML_TYPE_Q5_1) ? (int) (uint8_t) v : std::min(levels, (int) (int8_t) v);
        q[j] = (uint8_t) c;
    }
    uint8_t * o = out;
    put_h(o, d); o += 2;
    if (!sym) { put_h(o, base); o += 2; }
    if (five) {
        uint32_t qh = 0;
        for (int j = 0; j < 16; j++) { qh |= (uint32_t) ((q[j] >> 4) & 1) << j; qh |= (uint32_t) ((q[j + 16] >> 4) & 1) << (j + 16); }
        memcpy(o, &qh, 4); o += 4;
    }
    for (int j = 0; j < 16; j++) o[j] = (uint8_t) ((q[j] & 0x0F) | ((q[j + 16] & 0x0F) << 4));
}
}  // namespace

void biogpt_model_quantize_internal(std::ifstream & fin, std::ofstream & fout, const ggml_ftype ftype) {
    const ggml_type qtype = ggml_ftype_to_ggml_type(ftype);
    if (qtype == GGML_TYPE_COUNT || !ggml_is_quantized(qtype)) {
        fprintf(stderr, "%s: invalid model type %d\n", __func__, (int) ftype);
        throw std::runtime_error("invalid model type");
    }
    size_t bytes_in = 0, bytes_out = 0;
    std::vector<float> f32;
    std::vector<uint8_t> raw, packed;
    while (true) {
        int32_t n_dims = 0, name_len = 0, ttype = 0;
        read_safe(fin, n_dims); read_safe(fin, name_len); read_safe(fin, ttype);
        if (fin.eof()) break;
        int32_t ne[2] = { 1, 1 };
        int64_t n = 1;
        for (int i = 0; i < n_dims; i++) { read_safe(fin, ne[i]); n *= ne[i]; }
        std::string name(name_len, 0);
        fin.read(&name[0], name_len);
        printf("%64s - [%5d, %5d], type = %6s ", name.c_str(), ne[0], ne[1], ggml_type_name((ggml_type) ttype));
        const bool quantize = name.find("weight") != std::string::npos && ne[1] != 1;     // every 2-D "weight"
        if (quantize) {
            if (ttype != GGML_TYPE_F32 && ttype != GGML_TYPE_F16) throw std::runtime_error("unsupported ttype for integer quantization");
            if (ne[0] % 32) throw std::runtime_error("row length is not a multiple of the block size");
            f32.resize(n);
            if (ttype == GGML_TYPE_F16) {
                raw.resize(n * 2); fin.read((char *) raw.data(), n * 2);
                for (int64_t i = 0; i < n; i++) { ggml_fp16_t h; memcpy(&h, &raw[2 * i], 2); f32[i] = ggml_fp16_to_fp32(h); }
            } else fin.read((char *) f32.data(), n * 4);
            ttype = qtype;
        } else {
            raw.resize(n * (ttype == GGML_TYPE_F32 ? 4 : 2));
            fin.read((char *) raw.data(), raw.size());
        }
        write_safe(fout, n_dims); write_safe(fout, name_len); write_safe(fout, ttype);
        for (int i = 0; i < n_dims; i++) write_safe(fout, ne[i]);
        fout.write(name.data(), name_len);
        if (quantize) {
            const size_t bs = ggml_type_size(qtype);
            packed.resize((size_t) n / 32 * bs);
            for (int64_t b = 0; b < n / 32; b++) quant_block(qtype, f32.data() + 32 * b, packed.data() + bs * b);
            fout.write((const char *) packed.data(), packed.size());
            printf("size = %8.2f MB -> %8.2f MB\n", n * 4 / 1024.0 / 1024.0, packed.size() / 1024.0 / 1024.0);
            bytes_in += n * 4; bytes_out += packed.size();
        } else {
            fout.write((const char *) raw.data(), raw.size());
            printf("size = %8.3f MB\n", raw.size() / 1024.0 / 1024.0);
            bytes_in += raw.size(); bytes_out += raw.size();
        }
    }
    printf("%s: model size  = %8.2f MB\n", __func__, bytes_in / 1024.0 / 1024.0);
    printf("%s: quant size  = %8.2f MB | ftype = %d (%s)\n", __func__, bytes_out / 1024.0 / 1024.0, (int) ftype, ggml_type_name(qtype));
}

// ------------------------------------------------------------------------------------------------
// command line (same flags and quirks as the reference: -l writes the prompt, unknown flags exit 0)
// ------------------------------------------------------------------------------------------------
bool biogpt_params_parse(int argc, char ** argv, biogpt_params & params) {
    auto next = [&](int & i) -> const char * {
        if (i + 1 >= argc) { fprintf(stderr, "error: missing value for %s\n", argv[i]); biogpt_print_usage(argv, params); exit(0); }
        return argv[++i];
    };
    for (int i = 1; i < argc; i++) {
        const std::string a = argv[i];
        if      (a == "-s" || a == "--seed")       params.seed = std::stoi(next(i));
        else if (a == "-t" || a == "--threads")    params.n_threads = std::stoi(next(i));
        else if (a == "-p" || a == "--prompt")     params.prompt = next(i);
        else if (a == "-l" || a == "--lang")       params.prompt = next(i);        // reference quirk, biogpt.cpp:992-993
        else if (a == "-n" || a == "--n_predict")  params.n_predict = std::stoi(next(i));
        else if (a == "-v" || a == "--verbosity")  params.verbosity = (uint8_t) std::stoi(next(i));
        else if (a == "--top_k")                   params.top_k = std::stoi(next(i));
        else if (a == "--top_p")                   params.top_p = std::stof(next(i));
        else if (a == "--temp")                    params.temp = std::stof(next(i));
        else if (a == "-b" || a == "--batch_size") params.n_batch = std::stoi(next(i));
        else if (a == "-m" || a == "--model")      params.model = next(i);
        else if (a == "-h" || a == "--help")       { biogpt_print_usage(argv, params); exit(0); }
        else { fprintf(stderr, "error: unknown argument: %s\n", a.c_str()); biogpt_print_usage(argv, params); exit(0); }
    }
    return true;
}

void biogpt_print_usage(char ** argv, const biogpt_params & params) {
    fprintf(stderr, "usage: %s [options]\n\noptions:\n", argv[0]);
    fprintf(stderr, "  -h, --help            show this help message and exit\n");
    fprintf(stderr, "  -s SEED, --seed SEED  RNG seed (default: -1)\n");
    fprintf(stderr, "  -t N, --threads N     accepted for compatibility; the forward pass runs on the GPU (default: %d)\n", params.n_threads);
    fprintf(stderr, "  -p PROMPT, --prompt PROMPT\n                        prompt to start generation with (default: random)\n");
    fprintf(stderr, "  -l LANG               language of the prompt          (default: %s)\n", params.lang.c_str());
    fprintf(stderr, "  -n N, --n_predict N   number of tokens to predict (default: %d)\n", params.n_predict);
    fprintf(stderr, "  -v V, --verbosity V   verbosity level (default: %d)\n", params.verbosity);
    fprintf(stderr, "  --top_k N             top-k sampling  (default: %d)\n", params.top_k);
    fprintf(stderr, "  --top_p N             top-p sampling  (default: %.1f)\n", params.top_p);
    fprintf(stderr, "  --temp N              temperature     (default: %.1f)\n", params.temp);
    fprintf(stderr, "  -b N, --batch_size N  batch size for prompt processing (default: %d)\n", params.n_batch);
    fprintf(stderr, "  -m FNAME, --model FNAME\n                        model path (default: %s)\n\n", params.model.c_str());
}
