"""Parity at BASELINE.json's full sizes (10 M resident keys, ticks of 2^20 requests).

The oracle cannot replay 10 M string keys x millions of requests in seconds, so these tests use
size-independent properties of the domain:
  * per-key independence (core/tests.rs:65-91): the rows of any subset of keys, taken in trace order,
    must equal the oracle's replay of just those rows -- checked for the HOTTEST keys (tens of
    thousands of requests per tick each) and for a random sample of cold keys, bit-exact;
  * batch-split invariance: results do not depend on how the trace is cut into batches
    ("as if applied in index order");
  * placement invariance: a table of a different geometry (other capacity => other buckets, other
    sort digits) gives identical results;
  * counters: allowed + denied + errors == requests; len == number of keys ever allowed.
"""
import numpy as np
import pytest

import oracle
import throttlecrab_b200 as tc
import traces
from gpu_util import first_mismatch

pytestmark = pytest.mark.gpu
N_KEYS = 10_000_000
TICK = 1 << 20


def _engine_rows(trace, key_hash_of):
    req = np.empty(len(trace), tc.REQ_DTYPE)
    req["key_hash"] = key_hash_of[trace["key"].astype(np.int64)]
    for f in ("max_burst", "count_per_period", "period", "quantity", "now_ns"):
        req[f] = trace[f]
    return req


@pytest.fixture(scope="module")
def key_hash_of():
    h = tc.hash_key_ids(np.arange(N_KEYS, dtype=np.uint64))
    assert len(np.unique(h)) == N_KEYS          # the synthetic universe is collision-free
    return h


def _run(trace, key_hash_of, batch, capacity=N_KEYS):
    st = tc.ManualStore(capacity=capacity, created_ns=traces.T0, max_batch=TICK)
    lim = tc.RateLimiter(st)
    ereq = _engine_rows(trace, key_hash_of)
    res = np.empty(len(trace), tc.RES_DTYPE)
    for a in range(0, len(trace), batch):
        lim.rate_limit_batch(ereq[a:a + batch], out=res[a:a + batch])
    return res, st


def _check_subset(trace, res, keys):
    mask = np.isin(trace["key"], keys)
    sub = trace[mask]
    want = oracle.OracleStore(oracle.PERIODIC, capacity=len(keys), created_ns=traces.T0, p0=10**9).replay(sub)
    bad = first_mismatch(want, res[mask], sub)
    assert bad is None, bad
    return int(mask.sum())


@pytest.mark.parametrize("config", ["zipf", "hot100"])
def test_full_size_sampled_keys_match_oracle(config, key_hash_of):
    n_ticks = 3
    if config == "zipf":      # BASELINE configs[1]
        body = traces.config2(n_keys=N_KEYS, n_ticks=n_ticks, tick_size=TICK)
    else:                     # BASELINE configs[3]
        body = traces.config4(n_keys=N_KEYS, n_ticks=n_ticks, tick_size=TICK)
    trace = np.concatenate([traces.warm_pass(N_KEYS), body])
    res, st = _run(trace, key_hash_of, TICK)
    # hottest keys of the body + a random sample of everything else
    uniq, cnt = np.unique(body["key"], return_counts=True)
    hot = uniq[np.argsort(cnt)[-40:]]
    rng = np.random.default_rng(11)
    cold = rng.choice(N_KEYS, 3000, replace=False).astype(np.uint64)
    n_hot = _check_subset(trace, res, hot)
    n_cold = _check_subset(trace, res, cold)
    assert n_hot > 300_000 and n_cold > 3000
    s = st.stats()
    assert s["allowed"] + s["denied"] + s["errors"] == len(trace)
    assert s["allowed"] == int(res["allowed"].sum())
    assert s["len"] == N_KEYS                     # the warm pass stored every key once
    st.close()


def test_full_size_batch_split_and_placement_invariance(key_hash_of):
    body = traces.config2(n_keys=N_KEYS, n_ticks=2, tick_size=TICK)
    trace = np.concatenate([traces.warm_pass(N_KEYS), body])
    a, st_a = _run(trace, key_hash_of, TICK)
    st_a.close()
    b, st_b = _run(trace, key_hash_of, 100_003)                       # ragged batches
    st_b.close()
    assert a.tobytes() == b.tobytes()
    c, st_c = _run(trace, key_hash_of, TICK, capacity=3 * N_KEYS)     # other table geometry
    assert st_c.stats()["table_slots"] != 2**25
    st_c.close()
    assert a.tobytes() == c.tobytes()


def test_full_size_sweep_round_trip(key_hash_of):
    """insert 10 M keys -> sweep far in the future -> empty -> same trace gives the same answers."""
    trace = traces.warm_pass(N_KEYS)
    st = tc.ManualStore(capacity=N_KEYS, created_ns=traces.T0, max_batch=TICK)
    lim = tc.RateLimiter(st)
    ereq = _engine_rows(trace, key_hash_of)
    first = np.empty(len(trace), tc.RES_DTYPE)
    for a in range(0, len(trace), TICK):
        lim.rate_limit_batch(ereq[a:a + TICK], out=first[a:a + TICK])
    assert st.len() == N_KEYS
    # policies of the warm pass expire at T0 + {0, .5, 5.4, 5.94, 17.1, 24, 356, 594} s
    removed = st.sweep(traces.T0 + 6 * 10**9)
    assert removed == N_KEYS // 8 * 4 and st.len() == N_KEYS - removed
    assert st.sweep(traces.T0 + 6 * 10**9) == 0                    # idempotent
    assert st.sweep(traces.T0 + 10**12) == N_KEYS - removed and st.len() == 0
    assert st.stats()["occupied_slots"] == N_KEYS      # swept keys keep their slots until a purge is needed
    again = np.empty(len(trace), tc.RES_DTYPE)
    for a in range(0, len(trace), TICK):
        lim.rate_limit_batch(ereq[a:a + TICK], out=again[a:a + TICK])
    assert first.tobytes() == again.tobytes()
    st.close()
