// host_capi.cpp -- extern "C" doors onto the C++ host API for the python tests (ctypes cannot
// call functions that take std::string / std::vector).  Test plumbing only; the product entry
// points are the C++ functions of biogpt.h.
#include "biogpt.h"
#include "mosestokenizer.h"

#include <cstring>
#include <chrono>
#include "../../include/bgpt_cuda.h"

namespace {
int join(const std::vector<std::string> & v, char * out, int cap) {
    std::string s;
    for (size_t i = 0; i < v.size(); i++) { if (i) s += '\n'; s += v[i]; }
    if ((int) s.size() + 1 > cap) return -1;
    memcpy(out, s.c_str(), s.size() + 1);
    return (int) v.size();
}
std::vector<std::string> split_lines(const char * s) {
    std::vector<std::string> v; std::string cur;
    for (const char * p = s; *p; p++) { if (*p == '\n') { v.push_back(cur); cur.clear(); } else cur += *p; }
    if (!cur.empty() || !v.empty()) v.push_back(cur);
    return v;
}
struct Session { biogpt_model model; biogpt_vocab vocab; ggml_allocr * allocr = nullptr; ggml_backend_buffer_t buf = nullptr; std::vector<float> logits; };
}

void bgpt_text_class_mask(int which, uint32_t out[8]);
biogpt_vocab::id bgpt_sample_reference_order(int n_logits, const float * logits, int top_k, double top_p, double temp, std::mt19937 & rng);

extern "C" {

void bgpt_host_text_class_mask(int which, uint32_t * out8) { bgpt_text_class_mask(which, out8); }
int bgpt_host_moses_tokenize(const char * text, char * out, int cap) { return join(moses_tokenize(text, "en"), out, cap); }
int bgpt_host_moses_detokenize(const char * tokens_nl, char * out, int cap) {
    std::vector<std::string> v = split_lines(tokens_nl);
    const std::string s = moses_detokenize(v, "en");
    if ((int) s.size() + 1 > cap) return -1;
    memcpy(out, s.c_str(), s.size() + 1);
    return (int) s.size();
}
// merges: "left right" lines in priority order
int bgpt_host_bpe(const char * word, const char * merges_nl, char * out, int cap) {
    std::map<word_pair, int> ranks; int r = 0;
    for (const std::string & line : split_lines(merges_nl)) {
        const size_t sp = line.find(' ');
        if (sp != std::string::npos) ranks[word_pair(line.substr(0, sp), line.substr(sp + 1))] = r++;
    }
    const std::string s = bpe(word, ranks);
    if ((int) s.size() + 1 > cap) return -1;
    memcpy(out, s.c_str(), s.size() + 1);
    return (int) s.size();
}
int bgpt_host_sample(const float * logits, int n_vocab, int top_k, double top_p, double temp, uint32_t seed) {
    biogpt_vocab vocab;
    for (int i = 0; i < n_vocab; i++) vocab.id_to_token[i] = "";
    std::mt19937 rng(seed);
    return biogpt_sample_top_k_top_p(vocab, logits, top_k, top_p, temp, rng);
}

// which = 0: biogpt_sample_top_k_top_p (one-pass selection); 1: the reference's partial_sort over the whole vocabulary.  vocab_cache: the
// id -> token map of n_vocab empty strings is built once per size (the sampler only asks for its size)
int bgpt_host_sample_n(const float * logits, int n_vocab, int top_k, double top_p, double temp, uint32_t seed, int which) {
    static biogpt_vocab vocab;
    if ((int) vocab.id_to_token.size() != n_vocab) { vocab.id_to_token.clear(); for (int i = 0; i < n_vocab; i++) vocab.id_to_token[i] = ""; }
    std::mt19937 rng(seed);
    return which ? bgpt_sample_reference_order(n_vocab, logits, top_k, top_p, temp, rng) : biogpt_sample_top_k_top_p(vocab, logits, top_k, top_p, temp, rng);
}

// the flow of examples/main/main.cpp:29-70: load, measure pass, allocator
void * bgpt_host_open(const char * path, int n_batch) {
    Session * s = new Session();
    if (!biogpt_model_load(path, s->model, s->vocab, 0)) { delete s; return nullptr; }
    ggml_allocr * m = ggml_allocr_new_measure(ggml_backend_get_alignment(s->model.backend));
    const int n_tokens = std::min(s->model.hparams.n_positions, n_batch);
    ggml_cgraph * gf = biogpt_graph(s->model, m, token_sequence(n_tokens, 0), s->model.hparams.n_positions - n_tokens);
    const size_t mem = ggml_allocr_alloc_graph(m, gf);
    ggml_allocr_free(m);
    s->buf = ggml_backend_alloc_buffer(s->model.backend, mem);
    s->allocr = ggml_allocr_new_from_buffer(s->buf);
    return s;
}
int bgpt_host_n_vocab(void * h) { return ((Session *) h)->model.hparams.n_vocab; }
int bgpt_host_eval(void * h, const int32_t * tokens, int n, int n_past, float * logits_out) {
    Session * s = (Session *) h;
    if (!biogpt_eval(s->model, token_sequence(tokens, tokens + n), s->logits, s->allocr, n_past, 4)) return 1;
    memcpy(logits_out, s->logits.data(), s->logits.size() * sizeof(float));
    return 0;
}
// biogpt_eval_sample (device top-k + host draw): the id biogpt_eval + biogpt_sample_top_k_top_p would return
int bgpt_host_eval_sample(void * h, const int32_t * tokens, int n, int n_past, int top_k, double top_p, double temp, uint32_t seed) {
    Session * s = (Session *) h;
    std::mt19937 rng(seed);
    return biogpt_eval_sample(s->model, s->vocab, token_sequence(tokens, tokens + n), n_past, top_k, top_p, temp, rng);
}
int bgpt_host_tokenize(void * h, const char * text, int32_t * out, int cap) {
    Session * s = (Session *) h;
    const token_sequence ids = gpt_tokenize(s->vocab, text, "en");
    if ((int) ids.size() > cap) return -1;
    for (size_t i = 0; i < ids.size(); i++) out[i] = ids[i];
    return (int) ids.size();
}
// The token-by-token loop of examples/main/main.cpp:93-151 on an engine handle, for bench.py's end-to-end leg and tools/e2e_bench.py:
// ONE bgpt_cuda_eval_topk call per token with host buffers (token id in, top_k (logit, id) pairs out), the host picks the next token
// from the returned pairs -- the best one, so the ids can be compared with the device-resident greedy loop; equal logits among the
// pairs (exact = 0) are decided on the full row the call then delivers.  *wall_s = time inside the loop (std::chrono::steady_clock).
int bgpt_host_sampling_loop(bgpt_model * engine, int first_token, int n_past, int n_steps, int top_k, int32_t * ids_out, double * wall_s) {
    if (!engine || !ids_out || n_steps < 1 || top_k < 1 || top_k > 128) return -1;
    int32_t hp[7]; bgpt_cuda_hparams(engine, hp);
    std::vector<float> full((size_t) hp[0]);
    float vals[128]; int32_t ids[128]; int n_out = 0, exact = 0;
    int32_t tok = first_token;
    const auto t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < n_steps; i++) {
        const int rc = bgpt_cuda_eval_topk(engine, &tok, 1, n_past + i, top_k, vals, ids, &n_out, &exact, full.data());
        if (rc != BGPT_OK) return rc < 0 ? rc : -rc;
        if (exact && n_out > 0) tok = ids[0];
        else { int best = 0; for (int j = 1; j < hp[0]; j++) if (full[j] > full[best]) best = j; tok = best; }
        ids_out[i] = tok;
    }
    if (wall_s) *wall_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return 0;
}
// examples/main/main.cpp:93-151 as written -- biogpt_eval (the whole logit row comes back to the host), then biogpt_sample_top_k_top_p
// on it -- one token at a time, for bench.py's "unmodified front end" line.  ids_out[n_steps]; *wall_s = time inside the loop.
int bgpt_host_main_loop(void * h, int first_token, int n_past, int n_steps, int top_k, double top_p, double temp, uint32_t seed,
                        int32_t * ids_out, double * wall_s) {
    Session * s = (Session *) h;
    if (!s || !ids_out || n_steps < 1) return -1;
    std::mt19937 rng(seed);
    token_sequence embd(1, first_token);
    const auto t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < n_steps; i++) {
        if (!biogpt_eval(s->model, embd, s->logits, s->allocr, n_past + i, 4)) return 1;
        const biogpt_vocab::id id = biogpt_sample_top_k_top_p(s->vocab, s->logits.data() + (s->logits.size() - s->model.hparams.n_vocab), top_k, top_p, temp, rng);
        ids_out[i] = id;
        embd[0] = id;
    }
    if (wall_s) *wall_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return 0;
}
void bgpt_host_close(void * h) {
    Session * s = (Session *) h;
    if (!s) return;
    ggml_allocr_free(s->allocr);
    ggml_free(s->model.ctx);
    ggml_backend_buffer_free(s->model.buffer_w); ggml_backend_buffer_free(s->model.buffer_kv); ggml_backend_buffer_free(s->buf);
    ggml_backend_free(s->model.backend);
    delete s;
}

}  // extern "C"
