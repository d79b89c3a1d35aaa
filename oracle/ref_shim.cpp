// oracle/ref_shim.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// A thin extern "C" door onto the UNMODIFIED reference, compiled in place from
// /root/reference (see oracle/Makefile; nothing is copied into this repo).  It lets
// the python tests and bench.py's cpu_baseline / --impl reference legs call the
// reference's own public API:
//
//   biogpt_model_load            /root/reference/biogpt.cpp:27
//   biogpt_graph (measure pass)  /root/reference/examples/main/main.cpp:51-70
//   biogpt_eval                  /root/reference/biogpt.cpp:812
//   biogpt_sample_top_k_top_p    /root/reference/biogpt.cpp:908
//   biogpt_model_quantize_internal + the header copy done by
//   examples/quantize/quantize.cpp:8-135
//   ggml_internal_get_type_traits (block codecs, vec_dot)  ggml.h:2108
//
// Everything below is glue written for this repo; the arithmetic is the reference's.

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <random>
#include <string>
#include <vector>

#include "ggml.h"
#include "ggml-alloc.h"
#include "ggml-backend.h"
#include "biogpt.h"

namespace {

struct ref_handle {
    biogpt_model model;
    biogpt_vocab vocab;
    ggml_backend_buffer_t buf_compute = nullptr;
    ggml_allocr * allocr = nullptr;
    std::vector<float> logits;
};

// same worst-case measure pass as examples/main/main.cpp:51-70
void make_allocr(ref_handle * h, int n_batch) {
    size_t align = ggml_backend_get_alignment(h->model.backend);
    ggml_allocr * m = ggml_allocr_new_measure(align);
    int n_tokens = std::min(h->model.hparams.n_positions, n_batch);
    int n_past   = h->model.hparams.n_positions - n_tokens;
    ggml_cgraph * gf = biogpt_graph(h->model, m, token_sequence(n_tokens, 0), n_past);
    size_t mem = ggml_allocr_alloc_graph(m, gf);
    ggml_allocr_free(m);
    h->buf_compute = ggml_backend_alloc_buffer(h->model.backend, mem);
    h->allocr      = ggml_allocr_new_from_buffer(h->buf_compute);
}

}  // namespace

extern "C" {

void * ref_load(const char * path, int n_batch) {
    ggml_time_init();
    ref_handle * h = new ref_handle();
    if (!biogpt_model_load(path, h->model, h->vocab, 0)) {
        delete h;
        return nullptr;
    }
    make_allocr(h, n_batch > 0 ? n_batch : 8);
    return h;
}

void ref_hparams(void * vh, int32_t * out7) {
    ref_handle * h = (ref_handle *) vh;
    const biogpt_hparams & p = h->model.hparams;
    out7[0] = p.n_vocab; out7[1] = p.n_layer; out7[2] = p.n_head; out7[3] = p.n_positions;
    out7[4] = p.d_ff;    out7[5] = p.d_model; out7[6] = p.ftype;
}

// one biogpt_eval; logits_out receives n_vocab floats (last row, as the reference returns)
int ref_eval(void * vh, const int32_t * tokens, int n, int n_past, int n_threads, float * logits_out) {
    ref_handle * h = (ref_handle *) vh;
    token_sequence seq(tokens, tokens + n);
    if (!biogpt_eval(h->model, seq, h->logits, h->allocr, n_past, n_threads)) return 1;
    if (logits_out) memcpy(logits_out, h->logits.data(), h->logits.size() * sizeof(float));
    return 0;
}

// wall-clock microseconds of `reps` back-to-back evals of the same (tokens, n_past)
int64_t ref_time_eval(void * vh, const int32_t * tokens, int n, int n_past, int n_threads, int reps) {
    ref_handle * h = (ref_handle *) vh;
    token_sequence seq(tokens, tokens + n);
    const int64_t t0 = ggml_time_us();
    for (int i = 0; i < reps; i++) biogpt_eval(h->model, seq, h->logits, h->allocr, n_past, n_threads);
    return ggml_time_us() - t0;
}

// greedy / seeded sampling with the reference sampler
int ref_sample(void * vh, const float * logits, int top_k, double top_p, double temp, uint32_t seed) {
    ref_handle * h = (ref_handle *) vh;
    std::mt19937 rng(seed);
    return biogpt_sample_top_k_top_p(h->vocab, logits, top_k, top_p, temp, rng);
}

void ref_free(void * vh) {
    ref_handle * h = (ref_handle *) vh;
    if (!h) return;
    ggml_allocr_free(h->allocr);
    ggml_free(h->model.ctx);
    ggml_backend_buffer_free(h->model.buffer_w);
    ggml_backend_buffer_free(h->model.buffer_kv);
    ggml_backend_buffer_free(h->buf_compute);
    ggml_backend_free(h->model.backend);
    delete h;
}

// the quantize tool's flow: copy header/vocab/merges, rewrite ftype, quantise tensors
int ref_quantize(const char * fin_path, const char * fout_path, int ftype) {
    std::ifstream fin(fin_path, std::ios::binary);
    std::ofstream fout(fout_path, std::ios::binary);
    if (!fin || !fout) return 1;
    uint32_t magic; read_safe(fin, magic);
    if (magic != BIOGPT_FILE_MAGIC) return 2;
    write_safe(fout, magic);
    int32_t hp[7];
    for (int i = 0; i < 7; i++) read_safe(fin, hp[i]);
    hp[6] = ftype;
    for (int i = 0; i < 7; i++) write_safe(fout, hp[i]);
    for (int pass = 0; pass < 2; pass++) {   // vocab, then merges
        int32_t n; read_safe(fin, n); write_safe(fout, n);
        std::vector<char> tmp;
        for (int i = 0; i < n; i++) {
            uint32_t len; read_safe(fin, len); write_safe(fout, len);
            if (len) { tmp.resize(len); fin.read(tmp.data(), len); fout.write(tmp.data(), len); }
        }
    }
    try {
        // silence the per-tensor printf of the reference
        FILE * saved = stdout; (void) saved;
        fflush(stdout);
        biogpt_model_quantize_internal(fin, fout, (ggml_ftype) ftype);
    } catch (const std::exception & e) {
        fprintf(stderr, "ref_quantize: %s\n", e.what());
        return 3;
    }
    return 0;
}

// ---- block codec / dot-product pins (ggml type traits) -------------------------------

size_t ref_type_size(int type)  { return ggml_type_size((ggml_type) type); }
int    ref_blck_size(int type)  { return ggml_blck_size((ggml_type) type); }
int    ref_vec_dot_type(int type) { ggml_init_params p = {1024, nullptr, false}; ggml_free(ggml_init(p));
                                    return (int) ggml_internal_get_type_traits((ggml_type) type).vec_dot_type; }

// the quantiser mul_mat uses on activations (AVX2 build of the reference)
void ref_from_float(int type, const float * x, void * y, int k) {
    ggml_internal_get_type_traits((ggml_type) type).from_float(x, y, k);
}
// the deterministic quantiser the quantize tool uses on weights
void ref_from_float_reference(int type, const float * x, void * y, int k) {
    ggml_internal_get_type_traits((ggml_type) type).from_float_reference(x, y, k);
}
void ref_to_float(int type, const void * x, float * y, int k) {
    ggml_init_params p = {1024, nullptr, false}; ggml_free(ggml_init(p));   // builds the fp16 tables
    ggml_internal_get_type_traits((ggml_type) type).to_float(x, y, k);
}
float ref_vec_dot(int type, int n, const void * x, const void * y) {
    ggml_init_params p = {1024, nullptr, false}; ggml_free(ggml_init(p));
    float s = 0.0f;
    ggml_internal_get_type_traits((ggml_type) type).vec_dot(n, &s, x, y);
    return s;
}

// gelu / softmax / norm through tiny one-op graphs, so the fp16 tables and the
// double-precision reductions of the reference can be probed directly
static void run_unary(int op, const float * x, float * y, int nc, int nr, float eps) {
    size_t mem = (size_t) nc * nr * 4 * 4 + (1u << 20);
    ggml_init_params p = {mem, nullptr, false};
    ggml_context * ctx = ggml_init(p);
    ggml_tensor * a = ggml_new_tensor_2d(ctx, GGML_TYPE_F32, nc, nr);
    memcpy(a->data, x, (size_t) nc * nr * 4);
    ggml_tensor * r = nullptr;
    if (op == 0) r = ggml_gelu(ctx, a);
    if (op == 1) r = ggml_soft_max(ctx, a);
    if (op == 2) r = ggml_norm(ctx, a, eps);
    ggml_cgraph gf = ggml_build_forward(r);
    ggml_graph_compute_with_ctx(ctx, &gf, 1);
    memcpy(y, r->data, (size_t) nc * nr * 4);
    ggml_free(ctx);
}
void ref_gelu(const float * x, float * y, int n)                   { run_unary(0, x, y, n, 1, 0.f); }
void ref_soft_max(const float * x, float * y, int nc, int nr)      { run_unary(1, x, y, nc, nr, 0.f); }
void ref_norm(const float * x, float * y, int nc, int nr, float e) { run_unary(2, x, y, nc, nr, e); }

// dst[ne01, ne11] = mul_mat(src0 (type, [k, ne01]), src1 (f32, [k, ne11]))
void ref_mul_mat(int type, const void * w, const float * x, float * y, int k, int ne01, int ne11, int n_threads) {
    size_t wbytes = (size_t) k * ne01 * ggml_type_size((ggml_type) type) / ggml_blck_size((ggml_type) type);
    size_t mem = wbytes + (size_t) k * ne11 * 4 * 3 + (size_t) ne01 * ne11 * 4 + (8u << 20);
    ggml_init_params p = {mem, nullptr, false};
    ggml_context * ctx = ggml_init(p);
    ggml_tensor * a = ggml_new_tensor_2d(ctx, (ggml_type) type, k, ne01);
    ggml_tensor * b = ggml_new_tensor_2d(ctx, GGML_TYPE_F32, k, ne11);
    memcpy(a->data, w, wbytes);
    memcpy(b->data, x, (size_t) k * ne11 * 4);
    ggml_tensor * r = ggml_mul_mat(ctx, a, b);
    ggml_cgraph gf = ggml_build_forward(r);
    ggml_graph_compute_with_ctx(ctx, &gf, n_threads);
    memcpy(y, r->data, (size_t) ne01 * ne11 * 4);
    ggml_free(ctx);
}

}  // extern "C"
