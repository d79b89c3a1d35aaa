// replicas.cpp -- include/bgpt_replicas.h: one host thread and one engine per GPU, stream s on device s % G.
//
// The reference's serving loop is a single prompt in a single process (examples/main/main.cpp:93-151).  Independent streams share
// nothing, so the multi-GPU form is replication: each worker thread loads the model on its device through the reference's own
// loader API (biogpt_model_load with model.backend preset), keeps the KV caches of its streams there and decodes them in lock
// step.  The threads only meet at the boundaries of a call; inside bgpt_replicas_decode_greedy every device runs its whole loop
// without looking at the others.
#include "biogpt.h"
#include "../../include/bgpt_cuda.h"
#include "../../include/bgpt_replicas.h"

#include <condition_variable>
#include <cstring>
#include <memory>
#include <mutex>
#include <thread>

extern "C" ggml_backend_t ggml_backend_b200_init(int device);

namespace {
thread_local std::string g_rep_err;

enum Cmd { CMD_NONE = 0, CMD_LOAD, CMD_EVAL, CMD_EVAL_TOPK, CMD_DECODE, CMD_QUIT };

struct Worker {
    int device = 0;
    std::vector<int> streams;                  // global stream ids living on this device, ascending
    biogpt_model model;
    biogpt_vocab vocab;
    bool loaded = false;
    // mailbox
    std::thread th;
    std::mutex mu;
    std::condition_variable cv;
    Cmd cmd = CMD_NONE;
    bool done = true;
    // arguments / results of the current command
    std::string path;
    const int32_t * tokens = nullptr; int n_past = 0, n_steps = 0, top_k = 0; bool want_full = false;
    std::vector<float> vals_local; std::vector<int32_t> tid_local; std::vector<int> nout_local, exact_local;
    std::vector<int32_t> tok_local, ids_local;
    std::vector<float> logits_local;
    int rc = 0; std::string err; float ms = 0.f;
};
}  // namespace

// the engine handle lives in the opaque ggml_context of biogpt_b200.cpp
bgpt_model * bgpt_host_engine_of(const biogpt_model & model);

namespace {
void run(Worker * w) {
    for (;;) {
        Cmd c;
        {
            std::unique_lock<std::mutex> lk(w->mu);
            w->cv.wait(lk, [&] { return w->cmd != CMD_NONE; });
            c = w->cmd;
        }
        w->rc = 0; w->err.clear();
        const int nl = (int) w->streams.size();
        if (c == CMD_LOAD) {
            w->model.backend = ggml_backend_b200_init(w->device);
            if (!w->model.backend || !biogpt_model_load(w->path, w->model, w->vocab, 0)) { w->rc = -1; w->err = "load failed on device " + std::to_string(w->device) + ": " + bgpt_cuda_last_error(); }
            else {
                w->loaded = true;
                if (nl > 0 && bgpt_cuda_set_streams(bgpt_host_engine_of(w->model), nl) != BGPT_OK) { w->rc = -1; w->err = bgpt_cuda_last_error(); }
            }
        } else if (c == CMD_EVAL && nl > 0) {
            w->tok_local.resize(nl);
            for (int i = 0; i < nl; i++) w->tok_local[i] = w->tokens[w->streams[i]];
            w->logits_local.resize((size_t) nl * w->model.hparams.n_vocab);
            w->rc = bgpt_cuda_eval_streams(bgpt_host_engine_of(w->model), w->tok_local.data(), nl, w->n_past, w->logits_local.data());
            if (w->rc) w->err = bgpt_cuda_last_error();
        } else if (c == CMD_EVAL_TOPK && nl > 0) {
            const int k = w->top_k;
            w->tok_local.resize(nl);
            for (int i = 0; i < nl; i++) w->tok_local[i] = w->tokens[w->streams[i]];
            w->vals_local.resize((size_t) nl * k); w->tid_local.resize((size_t) nl * k); w->nout_local.resize(nl); w->exact_local.resize(nl);
            if (w->want_full) w->logits_local.resize((size_t) nl * w->model.hparams.n_vocab);
            w->rc = bgpt_cuda_eval_streams_topk(bgpt_host_engine_of(w->model), w->tok_local.data(), nl, w->n_past, k, w->vals_local.data(), w->tid_local.data(),
                                                w->nout_local.data(), w->exact_local.data(), w->want_full ? w->logits_local.data() : nullptr);
            if (w->rc) w->err = bgpt_cuda_last_error();
        } else if (c == CMD_DECODE && nl > 0) {
            w->tok_local.resize(nl);
            for (int i = 0; i < nl; i++) w->tok_local[i] = w->tokens[w->streams[i]];
            w->ids_local.resize((size_t) nl * w->n_steps);
            w->rc = bgpt_cuda_decode_greedy_streams(bgpt_host_engine_of(w->model), w->tok_local.data(), nl, w->n_past, w->n_steps, w->ids_local.data(), &w->ms);
            if (w->rc) w->err = bgpt_cuda_last_error();
        } else if (c == CMD_QUIT) {
            if (w->loaded) {
                ggml_free(w->model.ctx);
                ggml_backend_buffer_free(w->model.buffer_w); ggml_backend_buffer_free(w->model.buffer_kv);
                ggml_backend_free(w->model.backend);
            }
        }
        {
            std::lock_guard<std::mutex> lk(w->mu);
            w->cmd = CMD_NONE; w->done = true;
        }
        w->cv.notify_all();
        if (c == CMD_QUIT) return;
    }
}
void post(Worker * w, Cmd c) {
    { std::lock_guard<std::mutex> lk(w->mu); w->cmd = c; w->done = false; }
    w->cv.notify_all();
}
void wait(Worker * w) {
    std::unique_lock<std::mutex> lk(w->mu);
    w->cv.wait(lk, [&] { return w->done; });
}
}  // namespace

struct bgpt_replicas {
    std::vector<std::unique_ptr<Worker>> workers;
    int n_streams = 0, n_vocab = 0;
    int all(Cmd c) {                           // post to every device, then wait for all: the devices run concurrently
        for (auto & w : workers) post(w.get(), c);
        int rc = 0;
        for (auto & w : workers) { wait(w.get()); if (w->rc && !rc) { rc = w->rc; g_rep_err = w->err; } }
        return rc;
    }
};

extern "C" {

const char * bgpt_replicas_last_error(void) { return g_rep_err.c_str(); }

bgpt_replicas * bgpt_replicas_open(const char * model_path, int n_devices, int n_streams) {
    if (!model_path || n_streams < 1 || n_devices < 0) { g_rep_err = "replicas_open: bad arguments"; return nullptr; }
    const int visible = bgpt_cuda_device_count();
    if (visible <= 0) { g_rep_err = std::string("replicas_open: no CUDA device (this library has no CPU path): ") + bgpt_cuda_last_error(); return nullptr; }
    if (n_devices == 0 || n_devices > visible) n_devices = visible;
    if (n_devices > n_streams) n_devices = n_streams;
    bgpt_replicas * r = new bgpt_replicas();
    r->n_streams = n_streams;
    for (int g = 0; g < n_devices; g++) {
        std::unique_ptr<Worker> w(new Worker());
        w->device = g; w->path = model_path;
        for (int s = g; s < n_streams; s += n_devices) w->streams.push_back(s);
        w->th = std::thread(run, w.get());
        r->workers.push_back(std::move(w));
    }
    if (r->all(CMD_LOAD) != 0) { bgpt_replicas_close(r); return nullptr; }
    r->n_vocab = r->workers[0]->model.hparams.n_vocab;
    return r;
}

void bgpt_replicas_close(bgpt_replicas * r) {
    if (!r) return;
    const std::string keep = g_rep_err;
    for (auto & w : r->workers) post(w.get(), CMD_QUIT);
    for (auto & w : r->workers) if (w->th.joinable()) w->th.join();
    delete r;
    g_rep_err = keep;
}

int bgpt_replicas_devices(const bgpt_replicas * r) { return r ? (int) r->workers.size() : 0; }
int bgpt_replicas_streams(const bgpt_replicas * r) { return r ? r->n_streams : 0; }
int bgpt_replicas_n_vocab(const bgpt_replicas * r) { return r ? r->n_vocab : 0; }
int bgpt_replicas_device_of(const bgpt_replicas * r, int stream) { return r && stream >= 0 && stream < r->n_streams ? stream % (int) r->workers.size() : -1; }

int bgpt_replicas_eval(bgpt_replicas * r, const int32_t * tokens, int n_past, float * logits_out) {
    if (!r || !tokens || !logits_out) { g_rep_err = "replicas_eval: bad arguments"; return -1; }
    for (auto & w : r->workers) { w->tokens = tokens; w->n_past = n_past; }
    const int rc = r->all(CMD_EVAL);
    if (rc) return rc;
    for (auto & w : r->workers)
        for (size_t i = 0; i < w->streams.size(); i++)
            memcpy(logits_out + (size_t) w->streams[i] * r->n_vocab, w->logits_local.data() + i * r->n_vocab, (size_t) r->n_vocab * sizeof(float));
    return 0;
}

int bgpt_replicas_eval_topk(bgpt_replicas * r, const int32_t * tokens, int n_past, int k, float * vals, int32_t * ids, int * n_out, int * exact,
                            float * logits_fallback) {
    if (!r || !tokens || !vals || !ids || !n_out || !exact || k < 1) { g_rep_err = "replicas_eval_topk: bad arguments"; return -1; }
    for (auto & w : r->workers) { w->tokens = tokens; w->n_past = n_past; w->top_k = k; w->want_full = logits_fallback != nullptr; }
    const int rc = r->all(CMD_EVAL_TOPK);
    if (rc) return rc;
    for (auto & w : r->workers)
        for (size_t i = 0; i < w->streams.size(); i++) {
            const size_t sg = (size_t) w->streams[i];
            n_out[sg] = w->nout_local[i]; exact[sg] = w->exact_local[i];
            memcpy(vals + sg * k, w->vals_local.data() + i * k, (size_t) k * sizeof(float));
            memcpy(ids + sg * k, w->tid_local.data() + i * k, (size_t) k * sizeof(int32_t));
            if (!w->exact_local[i] && logits_fallback)
                memcpy(logits_fallback + sg * r->n_vocab, w->logits_local.data() + i * r->n_vocab, (size_t) r->n_vocab * sizeof(float));
        }
    return 0;
}

int bgpt_replicas_decode_greedy(bgpt_replicas * r, const int32_t * first_tokens, int n_past, int n_steps, int32_t * ids_out, float * device_ms) {
    if (!r || !first_tokens || !ids_out || n_steps < 1) { g_rep_err = "replicas_decode_greedy: bad arguments"; return -1; }
    for (auto & w : r->workers) { w->tokens = first_tokens; w->n_past = n_past; w->n_steps = n_steps; }
    const int rc = r->all(CMD_DECODE);
    if (rc) return rc;
    for (size_t g = 0; g < r->workers.size(); g++) {
        Worker * w = r->workers[g].get();
        const size_t nl = w->streams.size();
        for (int t = 0; t < n_steps; t++)
            for (size_t i = 0; i < nl; i++) ids_out[(size_t) t * r->n_streams + w->streams[i]] = w->ids_local[(size_t) t * nl + i];
        if (device_ms) device_ms[g] = w->ms;
    }
    return 0;
}

}  // extern "C"
