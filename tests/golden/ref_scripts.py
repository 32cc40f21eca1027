"""Known-answer scripts transcribed (as data) from the reference's own tests.

Every scenario cites the reference test it restates.  A step is
    (key, max_burst, count_per_period, period, quantity, t_ns, expect)
with `t_ns` relative to the scenario's base time and `expect` holding only what
the reference test asserts:
    err        -> expected status code (1 NegativeQuantity, 2 InvalidRateLimit); "ok" = any Ok(..)
    allowed    -> bool
    remaining  -> exact value        remaining_lt / remaining_gt / remaining_in=(lo,hi)
    retry_s_gt0 -> retry_after.as_secs() > 0
    reset_s / retry_s -> whole seconds (server types.rs:93-94 truncation)
Each scenario starts from an empty store.  These scripts are run against the
oracle (tests/test_oracle_known_answers.py, CPU) and against the CUDA engine
through the C ABI (tests/test_gpu_known_answers.py, GPU).
"""

S = 1_000_000_000
MS = 1_000_000
I64_MAX = 9223372036854775807

A = dict(allowed=True)
D = dict(allowed=False)


def _a(rem):
    return dict(allowed=True, remaining=rem)


def _d(rem):
    return dict(allowed=False, remaining=rem)


SCENARIOS = {}

# throttlecrab/src/core/tests.rs:5-14
SCENARIOS["core_basic_rate_limiting"] = [("test", 5, 10, 60, 1, 0, _a(4))]

# core/tests.rs:17-33
SCENARIOS["core_burst_capacity"] = (
    [("burst_test", 5, 10, 60, 1, 0, _a(5 - (i + 1))) for i in range(5)]
    + [("burst_test", 5, 10, 60, 1, 0, dict(allowed=False, remaining=0, retry_s_gt0=True))])

# core/tests.rs:36-62
SCENARIOS["core_rate_replenishment"] = [
    ("replenish_test", 2, 60, 60, 1, 0, A), ("replenish_test", 2, 60, 60, 1, 0, A),
    ("replenish_test", 2, 60, 60, 1, 0, D), ("replenish_test", 2, 60, 60, 1, 1 * S, A)]

# core/tests.rs:65-91
SCENARIOS["core_different_keys"] = [
    ("key1", 2, 2, 60, 1, 0, A), ("key2", 2, 2, 60, 1, 0, A), ("key1", 2, 2, 60, 1, 0, A),
    ("key1", 2, 2, 60, 1, 0, D), ("key2", 2, 2, 60, 1, 0, A), ("key2", 2, 2, 60, 1, 0, D)]

# core/tests.rs:94-118
SCENARIOS["core_quantity_parameter"] = [
    ("quantity_test", 10, 10, 60, 5, 0, _a(5)), ("quantity_test", 10, 10, 60, 6, 0, _d(5)),
    ("quantity_test", 10, 10, 60, 5, 0, _a(0))]

# core/tests.rs:121-127
SCENARIOS["core_negative_quantity_error"] = [("negative_test", 10, 10, 60, -1, 0, dict(err=1))]

# core/tests.rs:130-145
SCENARIOS["core_invalid_parameters"] = [
    ("test", 0, 10, 60, 1, 0, dict(err=2)), ("test", 10, 0, 60, 1, 0, dict(err=2)),
    ("test", 10, 10, 0, 1, 0, dict(err=2))]

# core/tests.rs:148-160
SCENARIOS["core_large_quantity_overflow_protection"] = [
    ("overflow_test", 10, 10, 60, I64_MAX // 2, 0, D)]

# core/tests.rs:163-176
SCENARIOS["core_saturating_arithmetic"] = [
    ("saturate_test", I64_MAX // 1000, 100, 60, 1, 0, dict(err="ok")),
    ("saturate_test2", 10, I64_MAX // 1000, 60, 1, 0, dict(err="ok"))]

# core/tests.rs:179-296
SCENARIOS["core_remaining_count_accuracy"] = (
    [("remaining_test", 5, 10, 60, 1, 0, _a(4))]
    + [("remaining_test", 5, 10, 60, 1, 0, _a(5 - i)) for i in range(2, 6)]
    + [("remaining_test", 5, 10, 60, 1, 0, dict(allowed=False, remaining=0, retry_s_gt0=True)),
       ("remaining_test", 5, 10, 60, 1, 6 * S, _a(0)),
       ("remaining_test", 5, 10, 60, 1, 6 * S, _d(0)),
       ("quantity_remaining", 5, 10, 60, 3, 0, _a(2)),
       ("quantity_remaining", 5, 10, 60, 3, 0, _d(2)),
       ("quantity_remaining", 5, 10, 60, 2, 0, _a(0)),
       ("high_rate", 10, 600, 60, 1, 0, _a(9))]
    + [("high_rate", 10, 600, 60, 1, 0, {}) for _ in range(9)]
    + [("high_rate", 10, 600, 60, 1, 1 * S, dict(allowed=True, remaining_lt=10))])

# core/tests.rs:299-347 (the same script on Periodic / Adaptive / Probabilistic stores)
SCENARIOS["core_remaining_count_all_stores"] = (
    [("test_key", 3, 6, 60, 1, 0, _a(3 - i)) for i in range(1, 4)]
    + [("test_key", 3, 6, 60, 1, 0, _d(0)), ("test_key", 3, 6, 60, 1, 10 * S, _a(0))])

# core/tests.rs:350-412
SCENARIOS["core_edge_cases_zero_remaining"] = [
    ("exact_timing", 2, 120, 60, 1, 0, _a(1)), ("exact_timing", 2, 120, 60, 1, 0, _a(0)),
    ("exact_timing", 2, 120, 60, 1, 500 * MS, _a(0)),
    ("zero_period", 10, 10, 0, 1, 0, dict(err=2)),
    ("fractional", 3, 7, 60, 1, 0, _a(2)),
    ("fractional", 3, 7, 60, 1, 0, {}), ("fractional", 3, 7, 60, 1, 0, {}),
    ("fractional", 3, 7, 60, 1, 8 * S, D), ("fractional", 3, 7, 60, 1, 9 * S, _a(0)),
    ("max_burst", I64_MAX // 1000, 100, 60, 1, 0, dict(allowed=True, remaining_gt=0))]

# core/tests.rs:415-500
_grad = []
for _ms, _avail, _rem in [(500, 1, 0), (1000, 2, 1), (1500, 3, 2), (2000, 4, 3), (2500, 5, 4)]:
    _k = "gradual_replenish_%d" % _ms
    _grad += [(_k, 5, 120, 60, 1, 0, {}) for _ in range(5)]
    _grad += [(_k, 5, 120, 60, 1, _ms * MS, _a(_rem))]
SCENARIOS["core_quantity_variations_and_replenishment"] = (
    [("multi_quantity", 10, 60, 60, 5, 0, _a(5)), ("multi_quantity", 10, 60, 60, 6, 0, _d(5)),
     ("multi_quantity", 10, 60, 60, 5, 0, _a(0)), ("multi_quantity", 10, 60, 60, 2, 3 * S, _a(1))]
    + [("gradual_replenish", 5, 120, 60, 1, 0, {}) for _ in range(5)] + _grad)

# core/tests.rs:503-601
_frac = []
for _ms, _rem in [(600, 0), (1200, 1), (1800, 2), (2400, 3), (3000, 4)]:
    _k = "fractional_accumulation_%d" % _ms
    _frac += [(_k, 5, 100, 60, 1, 0, {}) for _ in range(5)]
    _frac += [(_k, 5, 100, 60, 1, _ms * MS, _a(_rem))]
SCENARIOS["core_complex_replenishment_scenarios"] = (
    [("partial_burst", 8, 240, 60, 6, 0, _a(2)), ("partial_burst", 8, 240, 60, 1, 500 * MS, _a(3)),
     ("partial_burst", 8, 240, 60, 1, 1500 * MS, _a(6))]
    + [("slow_replenish", 3, 6, 60, 1, 0, {}) for _ in range(3)]
    + [("slow_replenish", 3, 6, 60, 1, 5 * S, D), ("slow_replenish", 3, 6, 60, 1, 10 * S, _a(0)),
       ("slow_replenish", 3, 6, 60, 1, 20 * S, _a(0))]
    + [("fractional_accumulation", 5, 100, 60, 1, 0, {}) for _ in range(5)] + _frac)

# core/tests.rs:604-655
SCENARIOS["core_quantity_edge_cases"] = [
    ("zero_quantity", 10, 100, 60, 0, 0, _a(10)), ("neg_quantity", 10, 100, 60, -5, 0, dict(err=1)),
    ("large_quantity", 5, 100, 60, 10, 0, _d(5)), ("exact_burst", 10, 100, 60, 10, 0, _a(0)),
    ("large_quantity_replenish", 20, 600, 60, 15, 0, _a(5)),
    ("large_quantity_replenish", 20, 600, 60, 12, 1 * S, _a(3)),
    ("large_quantity_replenish", 20, 600, 60, 5, 1 * S, _d(3))]

# core/tests.rs:658-694
SCENARIOS["core_rapid_time_changes"] = (
    [("time_jump", 3, 10, 60, 1, 0, A), ("time_jump", 3, 10, 60, 1, -5 * S, dict(err="ok")),
     ("time_jump", 3, 10, 60, 1, 10 * S, A)]
    + [("time_jitter", 10, 10, 60, 1, (i if i % 2 == 0 else -i) * S, dict(err="ok")) for i in range(5)])

# throttlecrab/src/core/store/store_test_suite.rs:543-598
SCENARIOS["store_rate_limiting_all_stores"] = (
    [("test_key", 5, 10, 3600, 1, 0, _a(5 - i - 1)) for i in range(5)]
    + [("test_key", 5, 10, 3600, 1, 0, D), ("test_key", 5, 10, 3600, 1, 360 * S, _a(0))])

# throttlecrab-server/src/transport/redis_test.rs:116-144 (second-granularity answers)
SCENARIOS["redis_throttle_allowed"] = [
    ("test_key", 10, 100, 60, 1, 0, dict(allowed=True, remaining=9, reset_s=5, retry_s=0))]
SCENARIOS["redis_throttle_with_quantity"] = [
    ("test_key2", 10, 100, 60, 5, 0, dict(allowed=True, remaining=5, reset_s=7, retry_s=0))]
# redis_test.rs:271-304
SCENARIOS["redis_throttle_exhaustion"] = [
    ("exhaustion_test", 3, 100, 60, 1, 0, _a(2)), ("exhaustion_test", 3, 100, 60, 1, 0, _a(1)),
    ("exhaustion_test", 3, 100, 60, 1, 0, _a(0)), ("exhaustion_test", 3, 100, 60, 1, 0, _d(0))]
# redis_test.rs:306-330
SCENARIOS["redis_multiple_keys"] = (
    [(k, 5, 100, 60, 1, 0, _a(4)) for k in ("user:123", "user:456", "api:endpoint")]
    + [(k, 5, 100, 60, 1, 0, _a(3)) for k in ("user:123", "user:456", "api:endpoint")])
# redis_test.rs:332-381
SCENARIOS["redis_different_limits_same_key"] = [
    ("dynamic_limit_key", 10, 100, 60, 1, 0, _a(9)),
    ("dynamic_limit_key", 5, 100, 60, 1, 0, dict(allowed=True, remaining_in=(0, 5)))]
# redis_test.rs:383-395
SCENARIOS["redis_large_quantity"] = [("large_quantity_key", 10, 100, 60, 15, 0, _d(10))]
# redis_test.rs:491-502
SCENARIOS["redis_zero_quantity"] = [("zero_quantity_key", 10, 100, 60, 0, 0, _a(10))]
# redis_test.rs:677-717
SCENARIOS["redis_boundary_values"] = [
    ("boundary_key", I64_MAX, I64_MAX, I64_MAX, 1, 0, A), ("tiny_key", 1, 1, 1, 1, 0, _a(0))]
# redis_test.rs:763-805
SCENARIOS["redis_very_long_key"] = [
    ("x" * 1000, 10, 100, 60, 1, 0, _a(9)), ("x" * 1000, 10, 100, 60, 1, 0, _a(8))]
# throttlecrab-server/src/actor_tests.rs:8-31 (first request: remaining = burst - 1) and
# :33-70 / transport/grpc.rs:244-295 (burst N => exactly N allowed out of more)
SCENARIOS["actor_first_request"] = [("test_key", 5, 10, 60, 1, 0, _a(4))]
SCENARIOS["actor_exactly_burst_allowed"] = (
    [("concurrent", 10, 100, 60, 1, 0, A) for _ in range(10)]
    + [("concurrent", 10, 100, 60, 1, 0, D) for _ in range(10)])

# throttlecrab/src/core/rate/tests.rs:41-48: (count, period) -> emission interval ns
RATE_VECTORS = [((10, 60), 6 * S), ((30, 60), 2 * S)]


def check_step(expect, status, allowed, remaining, reset_ns, retry_ns):
    """Assert one step's outputs against what the reference test asserts."""
    err = expect.get("err")
    if err == "ok":
        assert status == 0
    elif err is not None:
        assert status == err, (status, err)
        return
    else:
        assert status == 0, status
    if "allowed" in expect:
        assert allowed == expect["allowed"], (allowed, expect)
    if "remaining" in expect:
        assert remaining == expect["remaining"], (remaining, expect)
    if "remaining_lt" in expect:
        assert remaining < expect["remaining_lt"]
    if "remaining_gt" in expect:
        assert remaining > expect["remaining_gt"]
    if "remaining_in" in expect:
        lo, hi = expect["remaining_in"]
        assert lo <= remaining <= hi
    if expect.get("retry_s_gt0"):
        assert retry_ns // S > 0
    if "reset_s" in expect:
        assert reset_ns // S == expect["reset_s"]
    if "retry_s" in expect:
        assert retry_ns // S == expect["retry_s"]
