// coalesce.hpp -- seam-call coalescing ("group commit") for the host-buffer extension seam.
//
// Many executor threads call the JNI seam concurrently, each with a small batch (-bSWExtSize
// reads, often only a few thousand tasks -- too few to fill 148 SMs, and each call would pay its
// own launches, copies and synchronisation).  The coalescer merges calls that are pending at the
// same time into ONE device submission: callers copy their bytes into a shared pinned staging
// buffer in parallel, one worker thread issues a single H2D, one multi-call launch sequence and a
// single D2H for the whole group, and the callers copy their own replies out in parallel.
// No timer is involved: a group closes as soon as a worker is free (classic group commit), so an
// isolated call pays no added latency, and under load groups grow by themselves.
//
// The class is a template over the executor so that tests can drive the very same queueing code
// with a host executor; the product instantiates it with the CUDA executor in csbwa_api.cu.
#pragma once
#include <stdint.h>
#include <string.h>
#include <stdio.h>
#include <stdlib.h>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

namespace csw {

struct CoCall {             // same layout as ExtCall (ext_kernels.cuh)
    long long in_off;
    int in_bytes;
    int n_tasks;
    long long out_off;      // in shorts
    int task_base;
    int pad;
};

// Executor concept:
//   uint8_t *in_staging(int slot);  int16_t *out_staging(int slot);      (pinned, capacity below)
//   int run(int slot, const CoCall *calls, int n_calls, size_t span_bytes, int n_tasks);
//       -> transfers staging[0, span_bytes), runs all calls, fills out_staging; returns 0 or <0
template <class Exec>
class Coalescer {
public:
    struct Limits {
        size_t max_bytes;   // staging capacity per group (input side)
        int max_tasks;      // tasks per group
        int max_calls;      // calls per group
    };

    Coalescer(Exec *ex, int n_slots, int n_workers, Limits lim)
        : ex_(ex), lim_(lim), groups_(n_slots), stop_(false)
    {
        for (int i = 0; i < n_slots; ++i) groups_[i].slot = i;
        for (int i = 0; i < n_workers; ++i) workers_.emplace_back([this] { worker(); });
    }
    ~Coalescer()
    {
        {
            std::lock_guard<std::mutex> lk(mu_);
            stop_ = true;
        }
        cv_work_.notify_all();
        for (auto &t : workers_) t.join();
        const char *e = getenv("CSBWA_CO_TIMING");
        if (e && e[0] == '1' && n_calls_ > 0) {
            const double c = 1e-3 / (double)n_calls_, g = 1e-3 / (double)(n_groups_ > 0 ? n_groups_ : 1);
            fprintf(stderr, "[csbwa coalescer] calls %lld groups %lld | per call us: wait-slot %.1f copy-in %.1f wait-done %.1f copy-out %.1f"
                            " | per group us: wait-copies %.1f run %.1f\n", n_calls_, n_groups_, t_slot_ * c, t_in_ * c, t_done_ * c,
                    t_out_ * c, t_close_ * g, t_run_ * g);
        }
    }

    // bytes reserved at the start of the staging buffer: the call table, then {n_calls, n_tasks}
    // of the group at header_off() (so the device can size the launch sequence by itself)
    size_t header_off() const { return (size_t)lim_.max_calls * sizeof(CoCall); }
    size_t table_bytes() const { return (header_off() + 16 + 255) & ~(size_t)255; }

    // true if a call of this size can be coalesced at all
    bool fits(int in_bytes, int n_tasks) const
    {
        return (size_t)in_bytes + table_bytes() + 256 <= lim_.max_bytes && n_tasks <= lim_.max_tasks;
    }

    // Blocking: returns the executor's status for the group this call travelled in.
    int submit(const uint8_t *in, int in_bytes, int16_t *out, int n_tasks)
    {
        const long long t0 = now_ns();
        std::unique_lock<std::mutex> lk(mu_);
        Group *g = nullptr;
        for (;;) {
            g = find_open(in, in_bytes, n_tasks);
            if (g) break;
            g = find_free();
            if (g) { open_group(g, in); break; }
            cv_free_.wait(lk);
        }
        CoCall c;
        c.in_off = (long long)g->bytes;
        c.in_bytes = in_bytes;
        c.n_tasks = n_tasks;
        c.out_off = (long long)10 * g->tasks;
        c.task_base = g->tasks;
        c.pad = 0;
        g->calls.push_back(c);
        g->bytes += ((size_t)in_bytes + 255) & ~(size_t)255;
        g->tasks += n_tasks;
        g->copying++;
        const unsigned my_gen = g->gen;
        cv_work_.notify_one();
        lk.unlock();
        const long long t1 = now_ns();
        memcpy(ex_->in_staging(g->slot) + c.in_off, in, (size_t)in_bytes);   // parallel across callers
        const long long t2 = now_ns();
        lk.lock();
        if (--g->copying == 0) cv_work_.notify_all();
        g->cv_done.wait(lk, [&] { return g->gen == my_gen && g->state == DONE; });
        const int rc = g->rc;
        lk.unlock();
        const long long t3 = now_ns();
        if (rc == 0) memcpy(out, ex_->out_staging(g->slot) + c.out_off, (size_t)n_tasks * 20);
        lk.lock();
        if (--g->readers == 0) {
            g->state = FREE;
            g->gen++;
            cv_free_.notify_all();
        }
        t_slot_ += t1 - t0; t_in_ += t2 - t1; t_done_ += t3 - t2; t_out_ += now_ns() - t3;
        return rc;
    }

    // counters (for stats / tests)
    long long groups_run() const { return n_groups_; }
    long long calls_run() const { return n_calls_; }

private:
    enum State { FREE, OPEN, CLOSED, DONE };
    struct Group {
        int slot = 0;
        State state = FREE;
        unsigned gen = 0;
        std::vector<CoCall> calls;
        size_t bytes = 0;
        int tasks = 0;
        int copying = 0;
        int readers = 0;
        int rc = 0;
        uint8_t key[28];
        std::condition_variable cv_done;
    };

    static void make_key(const uint8_t *in, uint8_t *key)
    {
        memcpy(key, in, 8);             // options (bytes 0..7)
        memcpy(key + 8, in + 12, 20);   // bytes 12..31 (optional extension); 8..11 = taskNum excluded
    }
    Group *find_open(const uint8_t *in, int in_bytes, int n_tasks)
    {
        uint8_t key[28];
        make_key(in, key);
        for (auto &g : groups_)
            if (g.state == OPEN && memcmp(g.key, key, 28) == 0 &&
                g.bytes + (size_t)in_bytes + 256 <= lim_.max_bytes && g.tasks + n_tasks <= lim_.max_tasks &&
                (int)g.calls.size() < lim_.max_calls)
                return &g;
        return nullptr;
    }
    Group *find_free()
    {
        for (auto &g : groups_) if (g.state == FREE) return &g;
        return nullptr;
    }
    void open_group(Group *g, const uint8_t *in)
    {
        g->state = OPEN;
        g->calls.clear();
        g->bytes = table_bytes();
        g->tasks = 0;
        g->copying = 0;
        g->rc = 0;
        make_key(in, g->key);
    }

    void worker()
    {
        std::unique_lock<std::mutex> lk(mu_);
        for (;;) {
            Group *g = nullptr;
            cv_work_.wait(lk, [&] {
                if (stop_) return true;
                for (auto &x : groups_) if (x.state == OPEN && !x.calls.empty()) { g = &x; return true; }
                return false;
            });
            if (stop_ && !g) return;
            g->state = CLOSED;                              // group commit: no more joiners
            const long long t0 = now_ns();
            cv_work_.wait(lk, [&] { return g->copying == 0; });
            std::vector<CoCall> calls = g->calls;
            const size_t span = g->bytes;
            const int tasks = g->tasks;
            lk.unlock();
            const long long t1 = now_ns();
            memcpy(ex_->in_staging(g->slot), calls.data(), calls.size() * sizeof(CoCall));
            const int32_t dyn[4] = {(int32_t)calls.size(), tasks, 0, 0};
            memcpy(ex_->in_staging(g->slot) + header_off(), dyn, sizeof dyn);
            const int rc = ex_->run(g->slot, calls.data(), (int)calls.size(), span, tasks);
            lk.lock();
            t_close_ += t1 - t0; t_run_ += now_ns() - t1;
            n_groups_++;
            n_calls_ += (long long)calls.size();
            g->rc = rc;
            g->readers = (int)calls.size();
            g->state = DONE;
            g->cv_done.notify_all();
        }
    }

    Exec *ex_;
    Limits lim_;
    std::mutex mu_;
    std::condition_variable cv_work_, cv_free_;
    std::vector<Group> groups_;
    std::vector<std::thread> workers_;
    bool stop_;
    long long n_groups_ = 0, n_calls_ = 0;
    // phase accumulators in ns (under mu_); printed at destruction with CSBWA_CO_TIMING=1
    long long t_slot_ = 0, t_in_ = 0, t_done_ = 0, t_out_ = 0, t_close_ = 0, t_run_ = 0;
    static long long now_ns()
    {
        return std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count();
    }
};

} // namespace csw
