"""Oracle stores against the reference's Store contract and sweep tests."""
import pytest

import oracle
from store_contract import CONTRACT, NOW, S

KINDS = [oracle.PERIODIC, oracle.PROBABILISTIC, oracle.ADAPTIVE]


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("fn", CONTRACT, ids=lambda f: f.__name__)
def test_store_contract(fn, kind):
    fn(oracle.OracleStore(kind, capacity=100, created_ns=NOW))


def test_cleanup_actually_happens():           # store/cleanup_test.rs:8-41
    st = oracle.OracleStore(oracle.PERIODIC, capacity=100, created_ns=NOW)
    for i in range(1000):
        st.set_if_not_exists_with_ttl("key_%d" % i, i, 1 * S, NOW)
    assert st.len() == 1000
    fut = NOW + 61 * S
    st.set_if_not_exists_with_ttl("trigger", 999, 60 * S, fut)
    assert st.len() < 50
    assert st.get("trigger", fut) is not None


def test_cleanup_with_memory_pressure():       # cleanup_test.rs:44-83
    st = oracle.OracleStore(oracle.PERIODIC, capacity=100, created_ns=NOW)
    for i in range(500):
        st.set_if_not_exists_with_ttl("key_%d" % i, i, (1 if i % 2 == 0 else 3600) * S, NOW)
    later = NOW + 61 * S
    st.set_if_not_exists_with_ttl("trigger", 999, 60 * S, later)
    assert 200 < st.len() < 300
    for i in range(1, 100, 2):
        assert st.get("key_%d" % i, later) is not None


def test_no_cleanup_without_triggers():        # cleanup_test.rs:86-107
    st = oracle.OracleStore(oracle.PERIODIC, capacity=100, created_ns=NOW)
    for i in range(100):
        st.set_if_not_exists_with_ttl("key_%d" % i, i, 3600 * S, NOW)
    for i in range(10):
        st.get("key_%d" % i, NOW)
    assert st.len() == 100 and st.expired_count() == 0


def test_sweep_is_unobservable():
    """results are identical with and without sweeping (get hides expired entries)."""
    import numpy as np
    rng = np.random.default_rng(7)
    n = 20000
    req = np.zeros(n, oracle.REQ_DTYPE)
    req["key"] = rng.integers(0, 300, n)
    req["max_burst"], req["count_per_period"], req["period"] = 5, 10, 60
    req["quantity"] = rng.choice([0, 1, 2, 5], n)
    req["now_ns"] = NOW + np.cumsum(rng.choice([0, 1000, 10**9, 10**11], n))
    outs = []
    for kind, p in ((oracle.PERIODIC, 1), (oracle.PROBABILISTIC, 7), (oracle.ADAPTIVE, 0),
                    (oracle.PERIODIC, 10**9)):
        st = oracle.OracleStore(kind, capacity=100, created_ns=NOW, p0=p)
        outs.append((st.replay(req).tobytes(), st.sweeps()))
    assert len({o[0] for o in outs}) == 1
    assert outs[0][1] > 0 and outs[3][1] == 0
