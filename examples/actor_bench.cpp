// The batch-draining actor (gcra_actor_*, include/gcra_b200.h) under per-request callers.
//   1. the reference's actor tests (throttlecrab-server/src/actor_tests.rs:8-70) through the C ABI
//   2. T producer threads, each a blocking caller issuing one request at a time (what a connection handler of the
//      server does, actor.rs:68-82): requests/s and per-call latency percentiles; beside it the same calls made one
//      by one without the actor (one GPU round trip per request)
// Build: g++ -std=c++17 -O2 -pthread -Iinclude examples/actor_bench.cpp -Lthrottlecrab_b200 -lgcra_b200 -Wl,-rpath,...
// Usage: actor_bench [threads=32] [requests_per_thread=20000] [keys_per_thread=1000]
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "gcra_b200.h"

#define CHECK(c) do { if (!(c)) { std::printf("FAILED %s:%d %s\n", __FILE__, __LINE__, #c); std::exit(1); } } while (0)

static int64_t now_ns() {
    return std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::system_clock::now().time_since_epoch()).count();
}
static double mono() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int main(int argc, char **argv) {
    const int T = argc > 1 ? std::atoi(argv[1]) : 32;
    const int R = argc > 2 ? std::atoi(argv[2]) : 20000;
    const int K = argc > 3 ? std::atoi(argv[3]) : 1000;
    gcra_config cfg;
    std::memset(&cfg, 0, sizeof(cfg));
    cfg.capacity = 1000000;
    cfg.store_kind = GCRA_STORE_PERIODIC;
    cfg.p0 = 60;
    cfg.created_ns = now_ns();
    cfg.max_batch = 1 << 16;
    gcra_engine *h = nullptr;
    CHECK(gcra_create(&cfg, &h) == GCRA_OK);
    gcra_actor *a = nullptr;
    CHECK(gcra_actor_create(h, 100, 0, &a) == GCRA_OK);               // buffer 100 as in the reference's tests

    {   // actor_tests.rs:8-31 test_basic_rate_limiting
        gcra_result r;
        CHECK(gcra_actor_throttle(a, "test", 4, 5, 10, 60, 1, now_ns(), &r) == GCRA_OK);
        CHECK(r.allowed && r.remaining == 4);
    }
    {   // actor_tests.rs:33-70 test_concurrent_requests: 20 concurrent callers, burst 10 -> exactly 10 allowed
        std::atomic<int> allowed{0};
        const int64_t ts = now_ns();
        std::vector<std::thread> th;
        for (int i = 0; i < 20; i++)
            th.emplace_back([&] {
                gcra_result r;
                CHECK(gcra_actor_throttle(a, "concurrent_test", 15, 10, 10, 60, 1, ts, &r) == GCRA_OK);
                if (r.allowed) allowed++;
            });
        for (auto &t : th) t.join();
        CHECK(allowed.load() == 10);
    }
    {   // errors travel back to their caller (rate_limiter.rs:111-117)
        gcra_result r;
        CHECK(gcra_actor_throttle(a, "neg", 3, 10, 10, 60, -1, now_ns(), &r) == GCRA_NEGATIVE_QUANTITY);
        CHECK(gcra_actor_throttle(a, "bad", 3, 0, 10, 60, 1, now_ns(), &r) == GCRA_INVALID_RATE_LIMIT);
    }
    gcra_actor_destroy(a);

    // ---- load: T blocking callers
    CHECK(gcra_actor_create(h, 100000, 0, &a) == GCRA_OK);
    std::vector<std::vector<float>> lat(T);
    std::atomic<long> allowed_total{0};
    const double t0 = mono();
    {
        std::vector<std::thread> th;
        for (int t = 0; t < T; t++)
            th.emplace_back([&, t] {
                lat[t].reserve(R);
                char key[48];
                long ok = 0;
                for (int i = 0; i < R; i++) {
                    const int len = std::snprintf(key, sizeof(key), "user:%d:%d", t, i % K);
                    gcra_result r;
                    const double a0 = mono();
                    const int st = gcra_actor_throttle(a, key, (uint64_t)len, 100, 1000, 60, 1, now_ns(), &r);
                    lat[t].push_back((float)((mono() - a0) * 1e6));
                    if (st == GCRA_OK && r.allowed) ok++;
                }
                allowed_total += ok;
            });
        for (auto &t : th) t.join();
    }
    const double sec = mono() - t0;
    uint64_t st[3];
    gcra_actor_stats(a, st);
    gcra_actor_destroy(a);
    std::vector<float> all;
    for (auto &v : lat) all.insert(all.end(), v.begin(), v.end());
    std::sort(all.begin(), all.end());
    auto pct = [&](double p) { return all[(size_t)std::min<double>(all.size() - 1, p * all.size())]; };

    // ---- the same calls without the actor: one blocking gcra_rate_limit (one GPU round trip) per request
    const int N1 = 2000;
    const double s0 = mono();
    for (int i = 0; i < N1; i++) {
        char key[48];
        const int len = std::snprintf(key, sizeof(key), "solo:%d", i % K);
        gcra_result r;
        gcra_rate_limit(h, key, (uint64_t)len, 100, 1000, 60, 1, now_ns(), &r);
    }
    const double solo = N1 / (mono() - s0);
    gcra_destroy(h);
    std::printf("{\"actor_bench\": {\"threads\": %d, \"requests\": %ld, \"seconds\": %.3f, \"requests_per_s\": %.0f, "
                "\"latency_us\": {\"p50\": %.1f, \"p90\": %.1f, \"p99\": %.1f, \"p999\": %.1f, \"max\": %.1f}, "
                "\"batches\": %llu, \"mean_batch\": %.1f, \"largest_batch\": %llu, \"allowed\": %ld, "
                "\"without_actor_requests_per_s\": %.0f}}\n",
                T, (long)T * R, sec, (double)T * R / sec, pct(0.50), pct(0.90), pct(0.99), pct(0.999), all.back(),
                (unsigned long long)st[0], (double)st[1] / (double)std::max<uint64_t>(st[0], 1), (unsigned long long)st[2],
                allowed_total.load(), solo);
    std::printf("actor ok\n");
    return 0;
}
