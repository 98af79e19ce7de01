// tests/stress/co_stress.cpp -- TEST INFRASTRUCTURE: the product's Coalescer<> (csrc/coalesce.hpp) under a host executor,
// many caller threads, small limits so that slots recycle constantly.  Built with -fsanitize=thread by
// tests/test_coalesce.py::test_coalescer_thread_sanitizer (skipped when the toolchain has no TSan runtime).
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>
#include "../../cloud-scale-bwamem_b200/csrc/coalesce.hpp"

using namespace csw;

struct Exec {
    std::vector<std::vector<uint8_t>> hin;
    std::vector<std::vector<int16_t>> hout;
    std::vector<std::thread> th;
    std::vector<std::atomic<unsigned>> done;
    size_t hdr_off, ext_off;
    Exec(int n, size_t bytes, int tasks, int calls) : hin(n, std::vector<uint8_t>(bytes)), hout(n, std::vector<int16_t>((size_t)tasks * 10 + 32)), th(n), done(n)
    {
        for (auto &d : done) d.store(0);
        hdr_off = (size_t)calls * sizeof(CoCall); ext_off = hdr_off + 16;
    }
    ~Exec() { for (auto &t : th) if (t.joinable()) t.join(); }
    uint8_t *in_staging(int s) { return hin[s].data(); }
    int16_t *out_staging(int s) { return hout[s].data(); }
    unsigned long long in_staging_dev(int s) { return (unsigned long long)(uintptr_t)hin[s].data(); }
    unsigned long long out_staging_dev(int s) { return (unsigned long long)(uintptr_t)hout[s].data(); }
    const char *detail(int) { return ""; }
    int launch(int s, int n_calls, size_t, int, int, unsigned gen, int)
    {
        if (th[s].joinable()) th[s].join();
        th[s] = std::thread([this, s, n_calls, gen] {
            const CoCall *tab = (const CoCall *)hin[s].data();
            const CoExt *ext = (const CoExt *)(hin[s].data() + ext_off);
            for (int c = 0; c < n_calls; ++c) {            // reply = first byte of the call's payload + task index
                const uint8_t *src = (const uint8_t *)(uintptr_t)ext[c].src;
                int16_t *dst = (int16_t *)(uintptr_t)ext[c].dst;
                for (int k = 0; k < 10 * tab[c].n_tasks; ++k) dst[k] = (int16_t)(src[32] + k);
            }
            done[s].store(gen, std::memory_order_release);
        });
        return 0;
    }
    int poll(int s, unsigned gen) { return done[s].load(std::memory_order_acquire) == gen ? 1 : 0; }
    int finish(int, int, int, uint8_t *) { return 0; }
};

struct User { const uint8_t *in; int16_t *out; };
static void fill(void *u, uint8_t *dst, int n) { memcpy(dst, ((User *)u)->in, (size_t)n); }
static void drain(void *u, const int16_t *src, int n) { memcpy(((User *)u)->out, src, (size_t)n * 2); }

int main()
{
    const int n_slots = 3, n_threads = 12, reps = 400;
    Exec ex(n_slots, 1 << 16, 256, 8);
    Coalescer<Exec>::Limits lim{(size_t)1 << 16, 256, 8};
    std::atomic<int> bad{0};
    {
        Coalescer<Exec> co(&ex, n_slots, 2, lim);
        std::vector<std::thread> th;
        for (int t = 0; t < n_threads; ++t)
            th.emplace_back([&, t] {
                std::vector<uint8_t> in(64 + 16 * (size_t)(t % 5));
                for (int r = 0; r < reps; ++r) {
                    const int n_tasks = 1 + (t + r) % 7;
                    memset(in.data(), 0, in.size());
                    in[32] = (uint8_t)(t * 7 + r);
                    std::vector<int16_t> out((size_t)10 * n_tasks, -1);
                    User u{in.data(), out.data()};
                    CoRequest rq;
                    rq.hdr = in.data(); rq.in_bytes = (int)in.size(); rq.n_tasks = n_tasks;
                    const bool zc = (r % 3) == 0;
                    rq.src_dev = zc ? in.data() : nullptr; rq.dst_dev = zc ? out.data() : nullptr;
                    rq.fill = fill; rq.drain = drain; rq.user = &u;
                    if (co.submit(rq) != 0) ++bad;
                    for (int k = 0; k < 10 * n_tasks; ++k) if (out[(size_t)k] != (int16_t)((uint8_t)(t * 7 + r) + k)) { ++bad; break; }
                }
            });
        for (auto &x : th) x.join();
        printf("calls %lld groups %lld bad %d\n", co.calls_run(), co.groups_run(), bad.load());
    }
    return bad.load() ? 1 : 0;
}
