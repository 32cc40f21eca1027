"""The CUDA table through the Store-shaped C ABI against the reference's Store contract and
sweep tests (store_test_suite.rs, cleanup_test.rs)."""
import pytest

import throttlecrab_b200 as tc
from store_contract import CONTRACT, NOW, S

pytestmark = pytest.mark.gpu
STORES = [tc.PeriodicStore, tc.ProbabilisticStore, tc.AdaptiveStore]


@pytest.mark.parametrize("store_cls", STORES, ids=lambda c: c.__name__)
@pytest.mark.parametrize("fn", CONTRACT, ids=lambda f: f.__name__)
def test_store_contract(fn, store_cls):
    st = store_cls(capacity=100, created_ns=NOW, max_batch=4096)
    fn(st)
    st.close()


def test_cleanup_actually_happens():           # cleanup_test.rs:8-41 (also grows 100 -> 1000 keys)
    st = tc.PeriodicStore(capacity=100, created_ns=NOW, max_batch=4096)
    for i in range(1000):
        st.set_if_not_exists_with_ttl("key_%d" % i, i, 1 * S, NOW)
    assert st.len() == 1000
    fut = NOW + 61 * S
    st.set_if_not_exists_with_ttl("trigger", 999, 60 * S, fut)
    assert st.len() < 50
    assert st.get("trigger", fut) is not None
    assert st.stats()["grows"] >= 1


def test_cleanup_with_memory_pressure():       # cleanup_test.rs:44-83
    st = tc.PeriodicStore(capacity=100, created_ns=NOW, max_batch=4096)
    for i in range(500):
        st.set_if_not_exists_with_ttl("key_%d" % i, i, (1 if i % 2 == 0 else 3600) * S, NOW)
    later = NOW + 61 * S
    st.set_if_not_exists_with_ttl("trigger", 999, 60 * S, later)
    assert 200 < st.len() < 300
    for i in range(1, 100, 2):
        assert st.get("key_%d" % i, later) is not None


def test_no_cleanup_without_triggers():        # cleanup_test.rs:86-107
    st = tc.PeriodicStore(capacity=100, created_ns=NOW, max_batch=4096)
    for i in range(100):
        st.set_if_not_exists_with_ttl("key_%d" % i, i, 3600 * S, NOW)
    for i in range(10):
        st.get("key_%d" % i, NOW)
    assert st.len() == 100
    assert st.stats()["sweeps"] == 0


@pytest.mark.parametrize("k1_path", ["auto"], indirect=True)
def test_seeded_key_identity(k1_path, tmp_path):
    """An engine with a hash seed identifies string keys with SipHash-2-4 under that seed (include/gcra_b200.h, "key
    identity"): same decisions, other identities than the unkeyed hash, the seed survives a snapshot."""
    import traces
    seed = (0x1122334455667788, 0x99aabbccddeeff01)
    st = tc.PeriodicStore(capacity=1000, created_ns=traces.T0, hash_seed=seed)
    lim = tc.RateLimiter(st)
    for i in range(5):
        ok, r = lim.rate_limit("burst_test", 5, 10, 60, 1, traces.T0)          # core/tests.rs:17-33
        assert ok and r.remaining == 4 - i
    ok, r = lim.rate_limit("burst_test", 5, 10, 60, 1, traces.T0)
    assert not ok
    assert st.hash_seed() == seed
    assert st.hash_key("burst_test") != tc.hash_key("burst_test")
    assert st.peek(st.hash_key("burst_test")) is not None and st.peek(tc.hash_key("burst_test")) is None
    assert st.get("burst_test", traces.T0) is not None
    path = str(tmp_path / "seeded.snap")
    st.save(path)
    st2 = tc.PeriodicStore(capacity=1000, created_ns=traces.T0)               # no seed of its own
    st2.load(path)
    assert st2.hash_seed() == seed
    ok, _ = tc.RateLimiter(st2).rate_limit("burst_test", 5, 10, 60, 1, traces.T0)
    assert not ok                                                             # the same key, still exhausted
    rnd = tc.PeriodicStore(capacity=1000, created_ns=traces.T0, flags=8)      # GCRA_FLAG_RANDOM_SEED
    assert rnd.hash_seed() != (0, 0)
    for s in (st, st2, rnd):
        s.close()


def test_top_denied_keys_from_kernel_outputs():
    """metrics.rs:162-173 (top denied keys), fed from the batches' own request / result rows on the device: equal to a
    host-side count of the denied rows per key hash."""
    import numpy as np
    import traces
    from gpu_util import engine_requests
    n_keys = 5000
    req = traces.config4(n_keys=n_keys, n_ticks=4, tick_size=1 << 15, hot=50)
    ereq = engine_requests(req)
    st = tc.ManualStore(capacity=n_keys, created_ns=traces.T0, max_batch=1 << 15)
    st.track_denied(20000)                                   # room for every key: exact counts
    lim = tc.RateLimiter(st)
    res = np.empty(len(req), tc.RES_DTYPE)
    for a in range(0, len(req), 1 << 15):
        lim.rate_limit_batch(ereq[a:a + (1 << 15)], out=res[a:a + (1 << 15)])
    denied = (res["status"] == 0) & (res["allowed"] == 0)
    keys, cnt = np.unique(ereq["key_hash"][denied], return_counts=True)
    order = np.lexsort((keys, -cnt))
    want = [(int(keys[i]), int(cnt[i])) for i in order[:25]]
    got, dropped = st.top_denied(25)
    assert dropped == 0 and got == want and want[0][1] > 100
    assert sum(c for _, c in st.top_denied(20000)[0]) == int(denied.sum())
    st.close()
