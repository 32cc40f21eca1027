"""The batch-draining actor mirror against the reference's actor tests and RESP known answers
(throttlecrab-server/src/actor_tests.rs:8-70, transport/redis_test.rs:116-144)."""
import threading
import time

import pytest

import throttlecrab_b200 as tc
from throttlecrab_b200.actor import RateLimiterHandle, ThrottleRequest

pytestmark = pytest.mark.gpu
NOW = 1_700_000_000 * 10**9


def _handle():
    return RateLimiterHandle(tc.PeriodicStore(capacity=1000, created_ns=NOW, max_batch=4096, p0=60), buffer_size=100)


def test_basic_rate_limiting():                     # actor_tests.rs:8-31
    h = _handle()
    resp = h.throttle(ThrottleRequest("test", 5, 10, 60, 1, NOW)).result(timeout=30)
    assert resp.allowed and resp.limit == 5 and resp.remaining == 4
    h.shutdown()


def test_concurrent_requests():                     # actor_tests.rs:33-70
    h = _handle()
    futs, lock = [], threading.Lock()

    def send():
        f = h.throttle(ThrottleRequest("concurrent_test", 10, 10, 60, 1, NOW))
        with lock:
            futs.append(f)
    ts = [threading.Thread(target=send) for _ in range(20)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert sum(f.result(timeout=30).allowed for f in futs) == 10
    h.shutdown()


def test_second_granularity_and_errors():           # redis_test.rs:116-144, actor.rs:252
    h = _handle()
    r = h.throttle(ThrottleRequest("test_key", 10, 100, 60, 1, NOW)).result(timeout=30)
    assert (r.allowed, r.limit, r.remaining, r.reset_after, r.retry_after) == (True, 10, 9, 5, 0)
    r = h.throttle(ThrottleRequest("test_key2", 10, 100, 60, 5, NOW)).result(timeout=30)
    assert (r.remaining, r.reset_after) == (5, 7)
    with pytest.raises(tc.NegativeQuantity):
        h.throttle(ThrottleRequest("k", 10, 100, 60, -5, NOW)).result(timeout=30)
    with pytest.raises(tc.InvalidRateLimit):
        h.throttle(ThrottleRequest("k", 0, 100, 60, 1, NOW)).result(timeout=30)
    h.shutdown()


def test_many_callers_are_batched_in_order():
    """2 000 queued requests on one key are drained in a few batches and applied in arrival order."""
    h = _handle()
    futs = [h.throttle(ThrottleRequest("burst", 100, 1000, 60, 1, NOW)) for _ in range(2000)]
    res = [f.result(timeout=60) for f in futs]
    assert [r.allowed for r in res] == [True] * 100 + [False] * 1900
    assert [r.remaining for r in res[:100]] == list(range(99, -1, -1))
    assert h.batches < 200
    h.shutdown()
