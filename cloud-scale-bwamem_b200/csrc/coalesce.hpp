// coalesce.hpp -- seam-call coalescing ("group commit") for the host-buffer extension seam.
//
// Many executor threads call the JNI seam concurrently, each with a small batch (-bSWExtSize
// reads, often only a few thousand tasks -- too few to fill 148 SMs, and each call would pay its
// own launches, copies and synchronisation).  The coalescer merges calls that are pending at the
// same time into ONE device submission.  No timer is involved: a group closes as soon as a
// submission permit is free (classic group commit), so an isolated call pays no added latency, and
// under load groups grow by themselves.
//
// Host cost per call is the point of this design (8 GPUs share the cores of one box):
//   * ONE pump thread per GPU launches the groups and detects their completion; it is the only
//     thread of the seam that talks to the CUDA driver, so there is no contention for the driver's
//     context lock.  Completion is a flag the device writes into pinned host memory: polling it
//     costs a load, not a driver call.
//   * a call whose buffers are pinned (csbwa_host_alloc / csbwa_host_register) is served with NO
//     host copy at all: the group's table carries the device-visible addresses of the caller's
//     buffers, the device gathers the wire bytes from them and scatters the replies into them.
//   * other callers write their bytes straight into the group's pinned staging through a `fill`
//     callback (memcpy for C hosts, GetByteArrayRegion for the JNI glue: one copy, not two) and
//     read their replies from it through `drain`.
//   * callers sleep on a futex word per group; nobody re-acquires the queue mutex to learn that
//     the group is done.
//
// The class is a template over the executor so that tests can drive the very same queueing code
// with a host executor; the product instantiates it with the CUDA executor in api_extend.cu.
#pragma once
#include <stdint.h>
#include <string.h>
#include <stdio.h>
#include <stdlib.h>
#include <limits.h>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>
#if defined(__linux__)
#include <linux/futex.h>
#include <sys/prctl.h>
#include <sys/resource.h>
#include <sys/syscall.h>
#include <unistd.h>
#endif

namespace csw {

struct CoCall {             // same layout as ExtCall (ext_kernels.cuh)
    long long in_off;       // byte offset of the call's wire bytes inside the group's device input region
    int in_bytes;
    int n_tasks;
    long long out_off;      // in shorts
    int task_base;
    int pad;
};
// Where the device finds a call's wire bytes and puts its replies: device-visible addresses of either
// the caller's own pinned buffers or the group's staging.  One entry per call, after the call table.
struct CoExt {
    unsigned long long src;     // 16-byte aligned
    unsigned long long dst;     // 4-byte aligned
    int unit_base;              // running sum of 16-byte units of the preceding calls
    int n_units;                // ceil(in_bytes / 16)
    int pad[2];
};

// One seam call as the coalescer sees it.
struct CoRequest {
    const uint8_t *hdr;         // the 32 header bytes of the wire buffer (options = group key)
    int in_bytes, n_tasks;
    const void *src_dev;        // device-visible address of the whole wire buffer, or null -> fill
    void (*fill)(void *user, uint8_t *dst, int in_bytes);          // write the wire bytes to dst (pinned staging)
    void *dst_dev;              // device-visible address of the reply array, or null -> drain
    void (*drain)(void *user, const int16_t *src, int n_shorts);   // read 10 * n_tasks shorts from src (pinned staging)
    void *user;
};

// Executor concept:
//   uint8_t *in_staging(int slot);  int16_t *out_staging(int slot);          host addresses (pinned, capacity = limits)
//   unsigned long long in_staging_dev(int slot), out_staging_dev(int slot);   the same, as the device sees them
//   int launch(int slot, int n_calls, size_t span_bytes, int n_tasks, int n_units, unsigned gen, int others);
//       (others: groups already on the device when this one starts)
//       asynchronous: the table {CoCall[max_calls], dyn[4], CoExt[max_calls]} is already in in_staging
//   int poll(int slot, unsigned gen);      0 = still running, 1 = finished, < 0 = the submission failed
//   int finish(int slot, int n_calls, int n_tasks, uint8_t *call_bad);
//       after poll() != 0: group status (0 or < 0); call_bad[c] != 0 <=> call c carried a bad record
//   const char *detail(int slot);          message of the last failure of this slot
template <class Exec>
class Coalescer {
public:
    struct Limits {
        size_t max_bytes;   // staging capacity per group (input side)
        int max_tasks;      // tasks per group
        int max_calls;      // calls per group
        int reply_shorts = 10;   // 16-bit units of reply per task (10 = the extension seam's ExtRet record)
    };

    // max_inflight: groups on the device at once (<= n_slots - 1: one buffer keeps accepting calls)
    Coalescer(Exec *ex, int n_slots, int max_inflight, Limits lim)
        : ex_(ex), lim_(lim), groups_(n_slots), stop_(false)
    {
        max_inflight_ = max_inflight < 1 ? 1 : (max_inflight > n_slots - 1 && n_slots > 1 ? n_slots - 1 : max_inflight);
        for (int i = 0; i < n_slots; ++i) {
            groups_[i].slot = i;
            groups_[i].call_bad.assign((size_t)lim.max_calls, 0);
        }
        pump_ = std::thread([this] { pump(); });
    }
    ~Coalescer()
    {
        {
            std::lock_guard<std::mutex> lk(mu_);
            stop_ = true;
        }
        cv_work_.notify_all();
        pump_.join();
        const char *e = getenv("CSBWA_CO_TIMING");
        if (e && e[0] == '1' && n_calls_ > 0) {
            const double c = 1e-3 / (double)n_calls_, g = 1e-3 / (double)(n_groups_ > 0 ? n_groups_ : 1);
            fprintf(stderr, "[csbwa coalescer] calls %lld (%lld zero-copy) groups %lld | per call us: wait-slot %.1f copy-in %.1f wait-done %.1f copy-out %.1f"
                            " | per group us: close->launch %.1f launch call %.1f on device %.1f\n", n_calls_, n_zero_copy_.load(), n_groups_,
                    t_slot_.load() * c, t_in_.load() * c, t_done_.load() * c, t_out_.load() * c, t_close_ * g, t_launch_ * g, t_run_ * g);
        }
    }

    // layout of the start of the staging buffer: the call table, {n_calls, n_tasks, n_units, gen} of the group,
    // then the CoExt table (so the device can size and route everything by itself)
    size_t header_off() const { return (size_t)lim_.max_calls * sizeof(CoCall); }
    size_t ext_off() const { return header_off() + 16; }
    size_t table_bytes() const { return (ext_off() + (size_t)lim_.max_calls * sizeof(CoExt) + 255) & ~(size_t)255; }

    // true if a call of this size can be coalesced at all
    bool fits(int in_bytes, int n_tasks) const
    {
        return (size_t)in_bytes + table_bytes() + 256 <= lim_.max_bytes && n_tasks <= lim_.max_tasks;
    }

    // Blocking.  Returns 0, or the negative status of this call (bad record in THIS call) / of its group.
    // detail (nullable): receives the executor's message when the group failed as a whole.
    int submit(const CoRequest &rq, char *detail = nullptr, size_t detail_cap = 0)
    {
        const long long t0 = now_ns();
        std::unique_lock<std::mutex> lk(mu_);
        Group *g = nullptr;
        for (;;) {
            g = find_open(rq.hdr, rq.in_bytes, rq.n_tasks);
            if (g) break;
            g = find_free();
            if (g) { open_group(g, rq.hdr); break; }
            ++free_waiters_;
            cv_free_.wait(lk);
            --free_waiters_;
        }
        const int my = (int)g->calls.size();
        CoCall c;
        c.in_off = (long long)g->bytes;
        c.in_bytes = rq.in_bytes;
        c.n_tasks = rq.n_tasks;
        c.out_off = (long long)lim_.reply_shorts * g->tasks;
        c.task_base = g->tasks;
        c.pad = 0;
        CoExt x;
        x.src = rq.src_dev ? (unsigned long long)(uintptr_t)rq.src_dev : ex_->in_staging_dev(g->slot) + (unsigned long long)c.in_off;
        x.dst = rq.dst_dev ? (unsigned long long)(uintptr_t)rq.dst_dev : ex_->out_staging_dev(g->slot) + (unsigned long long)c.out_off * 2;
        x.unit_base = g->units;
        x.n_units = (rq.in_bytes + 15) >> 4;
        x.pad[0] = x.pad[1] = 0;
        g->calls.push_back(c);
        g->ext.push_back(x);
        g->bytes += ((size_t)rq.in_bytes + 255) & ~(size_t)255;
        g->tasks += rq.n_tasks;
        g->units += x.n_units;
        const unsigned my_gen = g->gen;
        const bool copy_in = rq.src_dev == nullptr;
        if (copy_in) g->copying++;
        const bool wake = pump_idle_;
        lk.unlock();
        if (wake) cv_work_.notify_one();
        const long long t1 = now_ns();
        if (copy_in) {
            rq.fill(rq.user, ex_->in_staging(g->slot) + c.in_off, rq.in_bytes);       // parallel across callers
            lk.lock();
            const bool last = --g->copying == 0 && g->state == CLOSED;
            lk.unlock();
            if (last) cv_work_.notify_one();
        }
        const long long t2 = now_ns();
        for (;;) {                                                                    // sleep until the group is done
            const unsigned seen = g->done_gen.load(std::memory_order_acquire);
            if (seen == my_gen) break;
            futex_wait(&g->done_gen, seen);
        }
        int rc = g->rc;
        if (rc < 0 && detail && detail_cap > 0) snprintf(detail, detail_cap, "%s", g->detail);
        if (rc == 0 && g->call_bad[(size_t)my]) rc = g->bad_rc;
        const long long t3 = now_ns();
        if (rc == 0 && !rq.dst_dev) rq.drain(rq.user, ex_->out_staging(g->slot) + c.out_off, lim_.reply_shorts * rq.n_tasks);
        if (g->readers.fetch_sub(1, std::memory_order_acq_rel) == 1) {
            lk.lock();
            g->state = FREE;
            g->gen++;
            const bool fw = free_waiters_ > 0;
            lk.unlock();
            if (fw) cv_free_.notify_all();
        }
        t_slot_ += t1 - t0; t_in_ += t2 - t1; t_done_ += t3 - t2; t_out_ += now_ns() - t3;
        if (!copy_in && rq.dst_dev) ++n_zero_copy_;
        return rc;
    }

    // status a call gets when one of ITS records was bad (the other calls of the group are unaffected)
    void set_bad_call_status(int rc) { bad_rc_default_ = rc; }

    // counters (for stats / tests)
    long long groups_run() const { return n_groups_; }
    long long calls_run() const { return n_calls_; }

private:
    enum State { FREE, OPEN, CLOSED, RUNNING, DONE };
    struct Group {
        int slot = 0;
        State state = FREE;
        unsigned gen = 1;                       // generation of the group currently using this buffer
        std::atomic<unsigned> done_gen{0};      // futex word: generation whose results are ready
        std::vector<CoCall> calls;
        std::vector<CoExt> ext;
        std::vector<uint8_t> call_bad;
        size_t bytes = 0;
        int tasks = 0;
        int units = 0;
        int copying = 0;
        std::atomic<int> readers{0};
        int rc = 0, bad_rc = 0;
        long long t_closed = 0, t_launched = 0;
        uint8_t key[28];
        char detail[160];
    };

    static void futex_wait(std::atomic<unsigned> *w, unsigned seen)
    {
#if defined(__linux__)
        syscall(SYS_futex, (unsigned *)w, FUTEX_WAIT_PRIVATE, seen, nullptr, nullptr, 0);
#else
        (void)seen;
        std::this_thread::yield();
#endif
    }
    static void futex_wake_all(std::atomic<unsigned> *w)
    {
#if defined(__linux__)
        syscall(SYS_futex, (unsigned *)w, FUTEX_WAKE_PRIVATE, INT_MAX, nullptr, nullptr, 0);
#else
        (void)w;
#endif
    }

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
        g->ext.clear();
        g->bytes = table_bytes();
        g->tasks = 0;
        g->units = 0;
        g->copying = 0;
        g->rc = 0;
        g->bad_rc = bad_rc_default_;
        make_key(in, g->key);
    }

    // results of group g are final: publish them and wake its callers (mu_ not held)
    void complete(Group *g, int rc)
    {
        g->rc = rc;
        g->detail[0] = 0;
        if (rc < 0) snprintf(g->detail, sizeof g->detail, "%s", ex_->detail(g->slot));
        g->readers.store((int)g->calls.size(), std::memory_order_relaxed);
        {
            std::lock_guard<std::mutex> lk(mu_);
            g->state = DONE;
            --inflight_;
            n_groups_++;
            n_calls_ += (long long)g->calls.size();
            if (g->t_launched) t_run_ += now_ns() - g->t_launched;
        }
        g->done_gen.store(g->gen, std::memory_order_release);
        futex_wake_all(&g->done_gen);
    }

    void pump()
    {
#if defined(__linux__)
        prctl(PR_SET_TIMERSLACK, 1000UL, 0, 0, 0);             // 1 us: the naps below are 20 us
        { const char *e = getenv("CSBWA_CO_NAP_US"); if (e && atoi(e) > 0 && atoi(e) <= 1000) nap_us_ = atoi(e); }
        // The pump is the one thread every caller of this GPU waits on, and it sleeps most of the time: when the callers
        // outnumber the cores (8 GPUs on 32 vCPUs: 64 callers per 4 cores) it should not queue behind them.  Needs
        // CAP_SYS_NICE; silently stays at the default priority without it.  CSBWA_PUMP_NICE=0 keeps the default.
        {
            const char *e = getenv("CSBWA_PUMP_NICE");
            const int nice_v = e ? atoi(e) : -10;
            if (nice_v != 0) (void)setpriority(PRIO_PROCESS, (id_t)syscall(SYS_gettid), nice_v);
        }
#endif
        std::unique_lock<std::mutex> lk(mu_);
        for (;;) {
            bool busy = false;
            // 1. groups that finished on the device
            for (auto &g : groups_) {
                if (g.state != RUNNING) continue;
                const int st = ex_->poll(g.slot, g.gen);
                if (st == 0) { busy = true; continue; }
                lk.unlock();
                int rc = st;
                if (st > 0) rc = ex_->finish(g.slot, (int)g.calls.size(), g.tasks, g.call_bad.data());
                complete(&g, rc);
                lk.lock();
            }
            // 2. commit: close open groups while permits are free
            for (auto &g : groups_) {
                if (g.state == OPEN && !g.calls.empty() && inflight_ < max_inflight_) {
                    g.state = CLOSED;
                    g.t_closed = now_ns();
                    ++inflight_;
                }
            }
            // 3. launch closed groups whose callers have finished writing their bytes
            for (auto &g : groups_) {
                if (g.state != CLOSED) continue;
                if (g.copying > 0) { busy = true; continue; }
                g.state = RUNNING;
                int others = 0;                             // groups on the device when this one starts
                for (auto &q : groups_) if (&q != &g && q.state == RUNNING) ++others;
                lk.unlock();
                const long long t1 = now_ns();
                uint8_t *h = ex_->in_staging(g.slot);
                memcpy(h, g.calls.data(), g.calls.size() * sizeof(CoCall));
                const int32_t dyn[4] = {(int32_t)g.calls.size(), g.tasks, g.units, (int32_t)g.gen};
                memcpy(h + header_off(), dyn, sizeof dyn);
                memcpy(h + ext_off(), g.ext.data(), g.ext.size() * sizeof(CoExt));
                memset(g.call_bad.data(), 0, g.call_bad.size());
                const int rc = ex_->launch(g.slot, (int)g.calls.size(), g.bytes, g.tasks, g.units, g.gen, others);
                g.t_launched = now_ns();
                t_close_ += t1 - g.t_closed; t_launch_ += g.t_launched - t1;
                if (rc < 0) { g.t_launched = 0; complete(&g, rc); }
                else busy = true;
                lk.lock();
            }
            if (stop_) {
                bool pending = false;
                for (auto &g : groups_) if (g.state == OPEN || g.state == CLOSED || g.state == RUNNING) pending = true;
                if (!pending) return;
            }
            // 4. wait: short naps while something is in flight, else until a caller arrives
            bool open_work = false;
            for (auto &g : groups_) if (g.state == OPEN && !g.calls.empty() && inflight_ < max_inflight_) open_work = true;
            if (open_work) continue;
            if (busy) {
                // one group in flight = a caller (or two) waiting on exactly that group: poll it closely
                cv_work_.wait_for(lk, std::chrono::microseconds(inflight_ <= 1 ? (nap_us_ < 5 ? nap_us_ : 5) : nap_us_));
            } else {
                pump_idle_ = true;
                cv_work_.wait(lk, [&] {
                    if (stop_) return true;
                    for (auto &g : groups_) if (g.state == OPEN && !g.calls.empty()) return true;
                    return false;
                });
                pump_idle_ = false;
            }
        }
    }

    Exec *ex_;
    Limits lim_;
    std::mutex mu_;
    std::condition_variable cv_work_, cv_free_;
    std::vector<Group> groups_;
    std::thread pump_;
    bool stop_;
    bool pump_idle_ = false;
    int inflight_ = 0, max_inflight_ = 1, free_waiters_ = 0;
    int bad_rc_default_ = -3;
    int nap_us_ = 20;
    long long n_groups_ = 0, n_calls_ = 0;
    // phase accumulators in ns; printed at destruction with CSBWA_CO_TIMING=1
    std::atomic<long long> t_slot_{0}, t_in_{0}, t_done_{0}, t_out_{0}, n_zero_copy_{0};
    long long t_close_ = 0, t_launch_ = 0, t_run_ = 0;
    static long long now_ns()
    {
        return std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count();
    }
};

} // namespace csw
