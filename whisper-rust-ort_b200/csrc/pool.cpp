// pool.cpp — one handle that keeps several batches in flight on a GPU (wb_pool_*).
//
// A decode is a chain of small dependent kernels that leaves most SMs idle, so one GPU reaches its throughput only
// with several independent batches in flight.  A wb_ctx is one such batch slot: its calls block
// the calling thread.  The pool owns S contexts on one device (weights shared through the WeightStore) and S worker
// threads; the caller — the reference's single-threaded file loop, main.rs:1161-1210 — submits batches without
// blocking and collects them by ticket, so a single host thread (a Rust `fn main`) drives the whole GPU.
//
// Ownership: the caller's pcm / offsets / output buffers must stay valid until wb_pool_wait returns for that ticket;
// prompt and suppress lists are copied at submit.  Tickets are handed out in submit order and may be waited for in
// any order, each exactly once.
#include <condition_variable>
#include <deque>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "common.h"
#include "../../include/whisper_b200.h"

struct wb_pool {
    struct Job {
        const float* pcm; const int64_t* offsets; int n_files;
        std::vector<int64_t> prompt, suppress, begin_suppress;
        int max_new; int64_t eot;
        int64_t* tokens_out; int32_t* lens_out; int32_t* file_idx_out; int cap_chunks;
        int n_chunks = 0, status = WB_OK;
        std::string err;
        bool done = false;
    };
    std::vector<wb_ctx*> slots;
    std::vector<std::thread> workers;
    std::mutex m;
    std::condition_variable cv_work, cv_done, cv_room;
    std::deque<int> queue;                 // tickets waiting for a slot
    std::map<int, Job> jobs;               // submitted and not yet collected
    int next_ticket = 0;
    int max_pending = 0;                   // submit blocks while this many tickets wait for a slot
    bool stop = false;

    void run(int slot) {
        for (;;) {
            Job* job = nullptr;
            {
                std::unique_lock<std::mutex> lk(m);
                cv_work.wait(lk, [&] { return stop || !queue.empty(); });
                if (queue.empty()) return;                     // stop requested and nothing left to do
                job = &jobs[queue.front()];                    // std::map nodes are stable under insertion
                queue.pop_front();
                cv_room.notify_one();
            }
            const int rc = wb_transcribe_batch(slots[slot], job->pcm, job->offsets, job->n_files, job->prompt.data(), (int)job->prompt.size(),
                                               job->max_new, job->eot, job->suppress.data(), (int)job->suppress.size(),
                                               job->begin_suppress.data(), (int)job->begin_suppress.size(), job->tokens_out, job->lens_out,
                                               job->file_idx_out, job->cap_chunks, &job->n_chunks);
            std::string err = rc == WB_OK ? std::string() : std::string(wb_last_error());   // this worker's thread-local text
            {
                std::lock_guard<std::mutex> lk(m);
                job->status = rc;
                job->err.swap(err);
                job->done = true;
            }
            cv_done.notify_all();
        }
    }
};

extern "C" {

int wb_pool_create(wb_pool** out, int device, const wb_model_cfg* cfg, const char* weights_path, int n_slots) {
    if (!out || !cfg || n_slots < 1 || n_slots > 64) { wb_set_error("wb_pool_create: bad arguments (1 <= n_slots <= 64)"); return WB_EINVAL; }
    *out = nullptr;
    wb_pool* p = new wb_pool();
    for (int i = 0; i < n_slots; ++i) {
        wb_ctx* c = nullptr;
        const int rc = wb_create(&c, device, cfg, weights_path);          // later contexts reuse the uploaded weights
        if (rc != WB_OK) {
            for (wb_ctx* q : p->slots) wb_destroy(q);
            delete p;
            return rc;                                                     // wb_last_error() holds wb_create's message
        }
        wb_set_load_hint(c, n_slots);                                      // one slot: latency-oriented kernels; several: throughput-oriented
        p->slots.push_back(c);
    }
    p->max_pending = n_slots;
    for (int i = 0; i < n_slots; ++i) p->workers.emplace_back([p, i] { p->run(i); });
    *out = p;
    return WB_OK;
}

void wb_pool_destroy(wb_pool* p) {
    if (!p) return;
    {
        std::lock_guard<std::mutex> lk(p->m);
        p->stop = true;                                                    // queued work is still completed
    }
    p->cv_work.notify_all();
    for (auto& t : p->workers) t.join();
    for (wb_ctx* c : p->slots) wb_destroy(c);
    delete p;
}

int wb_pool_slots(const wb_pool* p) { return p ? (int)p->slots.size() : 0; }

int wb_pool_submit(wb_pool* p, const float* pcm, const int64_t* offsets, int n_files, const int64_t* prompt, int prompt_len,
                   int max_new_tokens, int64_t eot, const int64_t* suppress, int n_suppress, const int64_t* begin_suppress,
                   int n_begin_suppress, int64_t* tokens_out, int32_t* lens_out, int32_t* file_idx_out, int cap_chunks) {
    if (!p || !pcm || !offsets || n_files < 1 || !prompt || prompt_len < 1 || !tokens_out || !lens_out || n_suppress < 0 ||
        n_begin_suppress < 0 || (n_suppress > 0 && !suppress) || (n_begin_suppress > 0 && !begin_suppress)) {
        wb_set_error("wb_pool_submit: bad arguments");
        return WB_EINVAL;
    }
    std::unique_lock<std::mutex> lk(p->m);
    if (p->stop) { wb_set_error("wb_pool_submit: the pool is shutting down"); return WB_ESTATE; }
    p->cv_room.wait(lk, [&] { return (int)p->queue.size() < p->max_pending; });
    const int ticket = p->next_ticket++;
    if (p->next_ticket < 0) p->next_ticket = 0;
    wb_pool::Job& j = p->jobs[ticket];
    j.pcm = pcm; j.offsets = offsets; j.n_files = n_files;
    j.prompt.assign(prompt, prompt + prompt_len);
    if (n_suppress) j.suppress.assign(suppress, suppress + n_suppress);
    if (n_begin_suppress) j.begin_suppress.assign(begin_suppress, begin_suppress + n_begin_suppress);
    j.max_new = max_new_tokens; j.eot = eot;
    j.tokens_out = tokens_out; j.lens_out = lens_out; j.file_idx_out = file_idx_out; j.cap_chunks = cap_chunks;
    p->queue.push_back(ticket);
    lk.unlock();
    p->cv_work.notify_one();
    return ticket;
}

int wb_pool_wait(wb_pool* p, int ticket, int* n_chunks_out) {
    if (!p) { wb_set_error("wb_pool_wait: null pool"); return WB_EINVAL; }
    std::unique_lock<std::mutex> lk(p->m);
    auto it = p->jobs.find(ticket);
    if (it == p->jobs.end()) { wb_set_error("wb_pool_wait: unknown ticket (never submitted, or already collected)"); return WB_EINVAL; }
    p->cv_done.wait(lk, [&] { return it->second.done; });
    const int rc = it->second.status;
    if (n_chunks_out) *n_chunks_out = it->second.n_chunks;
    if (rc != WB_OK) wb_set_error(it->second.err);                         // the worker's message, on the caller's thread
    p->jobs.erase(it);
    return rc;
}

}  // extern "C"
