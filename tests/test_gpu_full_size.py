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


def test_full_size_full_trace_matches_oracle(key_hash_of):
    """BASELINE configs[1], the WHOLE trace: 10 M-key warm pass + 3 Zipf ticks of 2^20 requests, every row against ONE
    oracle store (the reference's semantics: one request at a time, rate_limiter.rs:102-250)."""
    body = traces.config2(n_keys=N_KEYS, n_ticks=3, tick_size=TICK)
    trace = np.concatenate([traces.warm_pass(N_KEYS), body])
    res, st = _run(trace, key_hash_of, TICK)
    st.close()
    want = oracle.OracleStore(oracle.ADAPTIVE, capacity=N_KEYS, created_ns=traces.T0).replay(trace)
    bad = first_mismatch(want, res, trace)
    assert bad is None, bad


@pytest.mark.parametrize("k1_path", ["auto", "sort"], indirect=True)
def test_single_key_run_of_a_million_requests(k1_path):
    """ONE key, 2^20 requests in ONE batch: varying policy, quantity and clock, so that the run holds thousands of
    state changes and long saturated stretches (in the index-order pipeline: marks, residue, giant-run kernels;
    in the sort pipeline: the cluster kernel's finite-state rounds)."""
    n = TICK
    rng = np.random.default_rng(5)
    req = np.zeros(n, traces.REQ_DTYPE)
    req["key"] = 7
    # long stretches of one policy (saturation) with bursts of mixed ones
    pol = np.repeat(rng.integers(0, 8, n // 4096 + 1), 4096)[:n]
    mixed = rng.random(n) < 0.02
    pol[mixed] = rng.integers(0, 8, int(mixed.sum()))
    traces.fill_policy(req, pol)
    req["quantity"] = rng.choice([0, 1, 1, 1, 1, 2, 5], n)
    req["now_ns"] = traces.T0 + np.cumsum(rng.choice([0, 0, 0, 1000, 1_000_000, 50_000_000], n))
    st = tc.ManualStore(capacity=1000, created_ns=traces.T0, max_batch=TICK)
    lim = tc.RateLimiter(st)
    ereq = np.empty(n, tc.REQ_DTYPE)
    ereq["key_hash"] = tc.hash_key_ids(req["key"])
    for f in ("max_burst", "count_per_period", "period", "quantity", "now_ns"):
        ereq[f] = req[f]
    res = lim.rate_limit_batch(ereq)
    st.close()
    want = oracle.OracleStore(oracle.PERIODIC, capacity=100, created_ns=traces.T0, p0=10**9).replay(req)
    bad = first_mismatch(want, res, req)
    assert bad is None, bad
    assert 1000 < int(want["allowed"].sum()) < n - 1000


@pytest.mark.parametrize("k1_path", ["auto"], indirect=True)
def test_config3_100m_keys_sampled_parity_and_sweep(k1_path):
    """BASELINE configs[2]: 100 M resident keys (table of 2^28 slots, 10 GB), uniform ticks of 2^20 requests.
    Per-key independence: 20 000 sampled keys (their warm-pass row and every tick row) bit-exact against the oracle;
    then the sweep at three clocks removes exactly the entries whose expiry has passed (adaptive_cleanup.rs:176-182)."""
    n_keys, n_ticks = 100_000_000, 2
    st = tc.ManualStore(capacity=n_keys, created_ns=traces.T0, max_batch=TICK)
    lim = tc.RateLimiter(st)
    rng = np.random.default_rng(17)
    sample = np.unique(rng.integers(0, n_keys, 20_000).astype(np.uint64))
    sub_rows, sub_res = [], []
    req = np.empty(TICK, tc.REQ_DTYPE)
    for a in range(0, n_keys, TICK):                       # warm pass: one q=1 request per key at T0
        ids = np.arange(a, min(a + TICK, n_keys), dtype=np.uint64)
        w = np.zeros(len(ids), traces.REQ_DTYPE)
        w["key"] = ids
        traces.fill_policy(w, (ids % np.uint64(8)).astype(np.int64))
        w["quantity"] = 1
        w["now_ns"] = traces.T0
        r = req[:len(ids)]
        r["key_hash"] = tc.hash_key_ids(ids)
        for f in ("max_burst", "count_per_period", "period", "quantity", "now_ns"):
            r[f] = w[f]
        out = lim.rate_limit_batch(r)
        m = np.isin(ids, sample)
        sub_rows.append(w[m])
        sub_res.append(out[m].copy())
    assert st.len() == n_keys
    body = traces.config3(n_keys=n_keys, n_ticks=n_ticks, tick_size=TICK)
    for t in range(n_ticks):
        sl = body[t * TICK:(t + 1) * TICK]
        req["key_hash"] = tc.hash_key_ids(sl["key"])
        for f in ("max_burst", "count_per_period", "period", "quantity", "now_ns"):
            req[f] = sl[f]
        out = lim.rate_limit_batch(req)
        m = np.isin(sl["key"], sample)
        sub_rows.append(sl[m])
        sub_res.append(out[m].copy())
    sub = np.concatenate(sub_rows)
    got = np.concatenate(sub_res)
    want = oracle.OracleStore(oracle.PERIODIC, capacity=len(sample) * 2, created_ns=traces.T0, p0=10**9).replay(sub)
    bad = first_mismatch(want, got, sub)
    assert bad is None, bad
    assert len(sub) > len(sample)
    s = st.stats()
    assert s["allowed"] + s["denied"] + s["errors"] == n_keys + n_ticks * TICK
    # expiry of a key touched only by the warm pass = T0 + dvt(policy): P6 0 s, P5 0.5 s, P2 5.4 s, P0 5.94 s, ...
    # keys touched by a tick moved on; count what the oracle semantics say is expired at T0 + 1 s among the sample,
    # and for the whole table compare with the engine's own len bookkeeping
    before = st.len()
    removed = st.sweep(traces.T0 + 1_000_000_000)
    assert 0.20 * n_keys < removed < 0.26 * n_keys          # P6 and P5 keys (2/8), minus those refreshed by a tick
    assert st.len() == before - removed
    assert st.sweep(traces.T0 + 1_000_000_000) == 0
    rest = st.sweep(traces.T0 + 10**13)
    # what is left never expires: a zero-quantity request on a fresh max_burst = 1 key stores a wrapped TTL
    # (rate_limiter.rs:179-183) -- about 2 % of the 2^21 tick requests on the 1/8 of keys with policy P6
    left = st.len()
    assert rest == before - removed - left and 3000 < left < 8000
    st.close()
