// A few of the reference's known-answer tests through the C++ mirror (include/throttlecrab_b200.hpp):
// core/tests.rs:17-33 (burst capacity), :94-118 (quantity), :121-145 (errors), store_test_suite.rs:113-170 (TTL).
// Build: g++ -std=c++17 -Iinclude examples/cpp_known_answers.cpp -Lthrottlecrab_b200 -lgcra_b200 -Wl,-rpath,...
#include <cstdio>
#include <cstdlib>

#include "throttlecrab_b200.hpp"

using namespace throttlecrab;
#define CHECK(c) do { if (!(c)) { std::printf("FAILED %s:%d %s\n", __FILE__, __LINE__, #c); std::exit(1); } } while (0)

int main() {
    const SystemTime now = std::chrono::system_clock::now();
    {
        PeriodicStore store(1000);
        RateLimiter limiter(store);
        for (int i = 0; i < 5; i++) {                                   // core/tests.rs:22-26
            auto [allowed, r] = limiter.rate_limit("burst_test", 5, 10, 60, 1, now);
            CHECK(allowed && r.remaining == 5 - (i + 1) && r.limit == 5);
        }
        auto [allowed, r] = limiter.rate_limit("burst_test", 5, 10, 60, 1, now);   // :29-32
        CHECK(!allowed && r.remaining == 0 && r.retry_after >= std::chrono::seconds(1));
        auto a1 = limiter.rate_limit("quantity_test", 10, 10, 60, 5, now);         // :98-117
        auto a2 = limiter.rate_limit("quantity_test", 10, 10, 60, 6, now);
        auto a3 = limiter.rate_limit("quantity_test", 10, 10, 60, 5, now);
        CHECK(a1.first && a1.second.remaining == 5 && !a2.first && a2.second.remaining == 5 && a3.first && a3.second.remaining == 0);
        bool threw = false;
        try { limiter.rate_limit("negative_test", 10, 10, 60, -1, now); }          // :121-127
        catch (const CellError &e) { threw = e.kind == CellError::NegativeQuantity && e.quantity == -1; }
        CHECK(threw);
        threw = false;
        try { limiter.rate_limit("test", 0, 10, 60, 1, now); }                     // :135
        catch (const CellError &e) { threw = e.kind == CellError::InvalidRateLimit; }
        CHECK(threw);
    }
    {
        AdaptiveStore store(100);                                                  // store_test_suite.rs:113-170
        const Duration ttl = std::chrono::seconds(60);
        CHECK(store.set_if_not_exists_with_ttl("key1", 100, ttl, now));
        CHECK(store.get("key1", now + std::chrono::seconds(59)) == std::optional<int64_t>(100));
        const SystemTime expired = now + std::chrono::seconds(61);
        CHECK(!store.get("key1", expired).has_value());
        CHECK(!store.compare_and_swap_with_ttl("key1", 100, 200, ttl, expired));
        CHECK(store.set_if_not_exists_with_ttl("key1", 300, ttl, expired));
        CHECK(store.get("key1", expired) == std::optional<int64_t>(300));
        CHECK(store.len() == 1);
    }
    std::printf("cpp_known_answers ok\n");
    return 0;
}
