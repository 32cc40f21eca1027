// throttlecrab_b200.hpp -- C++ host-side mirror of the part of throttlecrab's library API that sits on
// the GCRA hot path, implemented over the C ABI in gcra_b200.h.  Header-only; link libgcra_b200.so.
//
//   reference (Rust)                                                 here (namespace throttlecrab)
//   enum CellError                     core/mod.rs:48-56             struct CellError (kind + message), thrown
//   struct RateLimitResult             rate_limiter.rs:12-22         struct RateLimitResult
//   trait Store                        core/store/mod.rs:85-133      class Store (pure virtual, same 3 methods)
//   AdaptiveStore/PeriodicStore/ProbabilisticStore (+with_capacity)  classes of the same names over the GPU table
//   RateLimiter<S>::new / rate_limit   rate_limiter.rs:56,102-110    class RateLimiter: rate_limit(...) + rate_limit_batch
//
// All state and every decision live on the GPU; nothing here computes a GCRA decision.
#pragma once
#include <chrono>
#include <cstdint>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "gcra_b200.h"

namespace throttlecrab {

using SystemTime = std::chrono::system_clock::time_point;
using Duration = std::chrono::nanoseconds;

inline int64_t to_ns(SystemTime t) {   // now.duration_since(UNIX_EPOCH).as_nanos() as i64 (rate_limiter.rs:126-127)
    return std::chrono::duration_cast<std::chrono::nanoseconds>(t.time_since_epoch()).count();
}

struct CellError : std::runtime_error {   // core/mod.rs:48-56
    enum Kind { NegativeQuantity = 1, InvalidRateLimit = 2, Internal = 3 } kind;
    int64_t quantity;
    CellError(Kind k, const std::string &msg, int64_t q = 0) : std::runtime_error(msg), kind(k), quantity(q) {}
};

struct RateLimitResult {   // rate_limiter.rs:12-22
    int64_t limit;
    int64_t remaining;
    Duration reset_after;
    Duration retry_after;
};

class Store {   // core/store/mod.rs:85-133
  public:
    virtual ~Store() = default;
    virtual bool compare_and_swap_with_ttl(const std::string &key, int64_t old_v, int64_t new_v, Duration ttl, SystemTime now) = 0;
    virtual std::optional<int64_t> get(const std::string &key, SystemTime now) = 0;
    virtual bool set_if_not_exists_with_ttl(const std::string &key, int64_t value, Duration ttl, SystemTime now) = 0;
};

class GpuStore : public Store {
  public:
    GpuStore(int kind, uint64_t capacity, int device = 0, uint64_t p0 = 0, uint64_t p1 = 0, uint64_t p2 = 0,
             uint32_t max_batch = 0) {
        gcra_config cfg{};
        cfg.capacity = capacity; cfg.device = device; cfg.store_kind = kind;
        cfg.p0 = p0; cfg.p1 = p1; cfg.p2 = p2;
        cfg.created_ns = to_ns(std::chrono::system_clock::now());   // SystemTime::now() in the constructors
        cfg.max_batch = max_batch;
        if (gcra_create(&cfg, &h_) != GCRA_OK || !h_)
            throw CellError(CellError::Internal, "gcra_create failed: no usable CUDA device (there is no CPU path)");
    }
    ~GpuStore() override { gcra_destroy(h_); }
    GpuStore(const GpuStore &) = delete;
    GpuStore &operator=(const GpuStore &) = delete;

    std::optional<int64_t> get(const std::string &key, SystemTime now) override {
        int64_t v = 0; uint8_t f = 0;
        check(gcra_store_get(h_, key.data(), key.size(), to_ns(now), &v, &f));
        return f ? std::optional<int64_t>(v) : std::nullopt;
    }
    bool compare_and_swap_with_ttl(const std::string &key, int64_t old_v, int64_t new_v, Duration ttl, SystemTime now) override {
        uint8_t ok = 0;
        check(gcra_store_cas(h_, key.data(), key.size(), old_v, new_v, (uint64_t)ttl.count(), to_ns(now), &ok));
        return ok != 0;
    }
    bool set_if_not_exists_with_ttl(const std::string &key, int64_t value, Duration ttl, SystemTime now) override {
        uint8_t ok = 0;
        check(gcra_store_set_nx(h_, key.data(), key.size(), value, (uint64_t)ttl.count(), to_ns(now), &ok));
        return ok != 0;
    }
    uint64_t len() { return gcra_len(h_); }                      // periodic.rs:113-116
    bool is_empty() { return len() == 0; }
    uint64_t sweep(SystemTime now) { uint64_t r = 0; check(gcra_sweep(h_, to_ns(now), &r)); return r; }
    gcra_engine *handle() { return h_; }

  protected:
    void check(int32_t rc) { if (rc != GCRA_OK) throw CellError(CellError::Internal, gcra_last_error(h_)); }
    gcra_engine *h_ = nullptr;
};

struct AdaptiveStore : GpuStore {        // adaptive_cleanup.rs:78-136
    explicit AdaptiveStore(uint64_t capacity = 1000, int device = 0, uint64_t min_interval_s = 0,
                           uint64_t max_interval_s = 0, uint64_t max_operations = 0)
        : GpuStore(GCRA_STORE_ADAPTIVE, capacity, device, min_interval_s, max_interval_s, max_operations) {}
    static AdaptiveStore with_capacity(uint64_t c) { return AdaptiveStore(c); }
};
struct PeriodicStore : GpuStore {        // periodic.rs:73-111
    explicit PeriodicStore(uint64_t capacity = 1000, int device = 0, uint64_t cleanup_interval_s = 0)
        : GpuStore(GCRA_STORE_PERIODIC, capacity, device, cleanup_interval_s) {}
};
struct ProbabilisticStore : GpuStore {   // probabilistic.rs:73-108
    explicit ProbabilisticStore(uint64_t capacity = 1000, int device = 0, uint64_t cleanup_modulo = 0)
        : GpuStore(GCRA_STORE_PROBABILISTIC, capacity, device, cleanup_modulo) {}
};

class RateLimiter {   // rate_limiter.rs:42-58
  public:
    explicit RateLimiter(GpuStore &store) : store_(store) {}

    // rate_limiter.rs:102-110
    std::pair<bool, RateLimitResult> rate_limit(const std::string &key, int64_t max_burst, int64_t count_per_period,
                                                int64_t period, int64_t quantity, SystemTime now) {
        gcra_result r{};
        int32_t st = gcra_rate_limit(store_.handle(), key.data(), key.size(), max_burst, count_per_period, period,
                                     quantity, to_ns(now), &r);
        if (st == GCRA_NEGATIVE_QUANTITY) throw CellError(CellError::NegativeQuantity, "negative quantity", quantity);
        if (st == GCRA_INVALID_RATE_LIMIT) throw CellError(CellError::InvalidRateLimit, "invalid rate limit parameters");
        if (st != GCRA_OK) throw CellError(CellError::Internal, gcra_last_error(store_.handle()));
        return {r.allowed != 0, RateLimitResult{max_burst, r.remaining, Duration(r.reset_after_ns), Duration(r.retry_after_ns)}};
    }

    // n requests, results as if applied in index order (what actor.rs:217-236 does one message at a time)
    void rate_limit_batch(const std::vector<gcra_request> &req, std::vector<gcra_result> &res) {
        res.resize(req.size());
        if (gcra_rate_limit_batch(store_.handle(), req.size(), req.data(), res.data()) != GCRA_OK)
            throw CellError(CellError::Internal, gcra_last_error(store_.handle()));
    }
    static uint64_t hash_key(const std::string &key) { return gcra_hash_key(key.data(), key.size()); }

  private:
    GpuStore &store_;
};

}  // namespace throttlecrab
