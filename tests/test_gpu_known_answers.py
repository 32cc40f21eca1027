"""The CUDA engine, through the C ABI, against the reference's known-answer tests and the
nanosecond-exact vectors -- the same data the oracle is pinned with."""
import pytest

import throttlecrab_b200 as tc
from golden import ns_vectors, ref_scripts

pytestmark = pytest.mark.gpu
T0 = ns_vectors.T0
STORES = [tc.PeriodicStore, tc.ProbabilisticStore, tc.AdaptiveStore]


def _call(lim, key, b, c, p, q, now):
    try:
        allowed, r = lim.rate_limit(key, b, c, p, q, now)
        return 0, allowed, r.remaining, r.reset_after, r.retry_after
    except tc.NegativeQuantity:
        return 1, False, 0, 0, 0
    except tc.InvalidRateLimit:
        return 2, False, 0, 0, 0
    except tc.Internal:
        return 3, False, 0, 0, 0


@pytest.mark.parametrize("store_cls", STORES, ids=lambda c: c.__name__)
@pytest.mark.parametrize("name", sorted(ref_scripts.SCENARIOS))
def test_reference_scripts(name, store_cls):
    lim = tc.RateLimiter(store_cls(capacity=1000, created_ns=T0, max_batch=4096))
    for (key, b, c, p, q, t, expect) in ref_scripts.SCENARIOS[name]:
        ref_scripts.check_step(expect, *_call(lim, key, b, c, p, q, T0 + t))
    lim.store.close()


@pytest.mark.parametrize("name", sorted(ns_vectors.VECTORS))
def test_ns_vectors(name):
    st = tc.AdaptiveStore(capacity=1000, created_ns=T0, max_batch=4096)
    lim = tc.RateLimiter(st)
    for row in ns_vectors.VECTORS[name]:
        key, b, c, p, q, t, status, allowed, rem, reset, retry, tat_rel, exp_rel = row
        out = _call(lim, key, b, c, p, q, T0 + t)
        assert out[0] == status
        if status == 0:
            assert out[1:] == (bool(allowed), rem, reset, retry), (name, row, out)
        ent = st.peek(tc.hash_key(key))
        if tat_rel is None:
            assert ent is None
        else:
            assert ent == (T0 + tat_rel, ns_vectors.sat_expiry(exp_rel)), (name, row, ent)
    st.close()


def test_limit_is_echoed():
    lim = tc.RateLimiter(tc.PeriodicStore(capacity=100, created_ns=T0, max_batch=4096))
    _, r = lim.rate_limit("k", 2**63 - 1, 2**63 - 1, 2**63 - 1, 1, T0)   # redis_test.rs:677-697
    assert r.limit == 2**63 - 1 and r.remaining == 4294967294


def test_derive_matches_rate_tests():
    for (count, period), ei in ref_scripts.RATE_VECTORS:       # rate/tests.rs:41-48
        assert tc.derive_params(1, count, period)[:2] == (0, ei)
