"""Synthetic request traces shared by the parity tests and bench.py.

Counter-based SplitMix64 streams (seed 42), T0 = 1.7e18 ns, key i <-> ASCII "k:<i>", the 8-entry
policy table taken from the reference's own tests/benches, quantity mix 1/2/5/0 = 90/5/3/2 %.
Pure numpy; no dependency on the engine or the oracle (key hashes are filled in by the caller).
"""
import numpy as np

T0 = 1_700_000_000 * 1_000_000_000
SEED = 42

REQ_DTYPE = np.dtype([("key", "<u8"), ("max_burst", "<i8"), ("count_per_period", "<i8"),
                      ("period", "<i8"), ("quantity", "<i8"), ("now_ns", "<i8")])

# (max_burst, count_per_period, period s) with the reference location each comes from
POLICIES = np.array([
    (100, 1000, 60),    # P0 throttlecrab-server/benches/store_performance.rs:26-28
    (100, 1000, 3600),  # P1 throttlecrab-server/examples/store_comparison.rs:17
    (10, 100, 60),      # P2 transport/redis_test.rs:120
    (5, 10, 60),        # P3 throttlecrab/src/core/tests.rs:10
    (3, 7, 60),         # P4 core/tests.rs:383
    (2, 120, 60),       # P5 core/tests.rs:357
    (1, 3, 3600),       # P6 throttlecrab/src/lib.rs:101
    (100, 10, 60),      # P7 integration-tests/src/perf_test_multi_transport.rs:70-72
], dtype=np.int64)


def splitmix64(x):
    x = np.asarray(x, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = x + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def stream(config_id, lane, n, start=0):
    """n uniform u64 values of stream (config_id, lane), counter-based."""
    with np.errstate(over="ignore"):
        idx = np.arange(start, start + n, dtype=np.uint64)
        base = splitmix64(np.uint64(SEED) * np.uint64(0x100000001B3) + np.uint64(config_id * 1000 + lane))
        return splitmix64(idx * np.uint64(0x9E3779B97F4A7C15) + base)


def unit(u):
    return (u >> np.uint64(11)).astype(np.float64) * (1.0 / (1 << 53))


def quantities(u):
    """1 (90 %), 2 (5 %), 5 (3 %), 0 (2 %)"""
    f = unit(u)
    q = np.ones(len(u), np.int64)
    q[f >= 0.90] = 2
    q[f >= 0.95] = 5
    q[f >= 0.98] = 0
    return q


def fill_policy(req, pol_idx):
    p = POLICIES[pol_idx]
    req["max_burst"], req["count_per_period"], req["period"] = p[:, 0], p[:, 1], p[:, 2]


def config1(n=100_000, keys=1000, policy_of_key=None):
    """Correctness replay: uniform keys, per-request clock, 1 % foreign policy, 100 invalid rows."""
    cid = 1
    req = np.zeros(n, REQ_DTYPE)
    key = (stream(cid, 0, n) % np.uint64(keys)).astype(np.uint64)
    req["key"] = key
    pol = (key % np.uint64(8)).astype(np.int64) if policy_of_key is None else policy_of_key[key.astype(np.int64)]
    other = unit(stream(cid, 1, n)) < 0.01
    pol = np.where(other, (stream(cid, 2, n) % np.uint64(8)).astype(np.int64), pol)
    fill_policy(req, pol)
    req["quantity"] = quantities(stream(cid, 3, n))
    # clock: 0 (50 %), 1 us..1 ms (30 %), 0.1..10 s (19 %), 1..2 h (1 %)
    f = unit(stream(cid, 4, n))
    g = unit(stream(cid, 5, n))
    dt = np.zeros(n, np.int64)
    m = (f >= 0.5) & (f < 0.8)
    dt[m] = (1_000 + g[m] * (1_000_000 - 1_000)).astype(np.int64)
    m = (f >= 0.8) & (f < 0.99)
    dt[m] = (100_000_000 + g[m] * (10_000_000_000 - 100_000_000)).astype(np.int64)
    m = f >= 0.99
    dt[m] = (3_600_000_000_000 + g[m] * 3_600_000_000_000).astype(np.int64)
    req["now_ns"] = T0 + np.cumsum(dt)
    # 100 invalid rows: q < 0, zero / negative parameters
    bad = (stream(cid, 6, 100) % np.uint64(n)).astype(np.int64)
    for j, i in enumerate(bad):
        kind = j % 4
        if kind == 0:
            req["quantity"][i] = -1 - j
        elif kind == 1:
            req["max_burst"][i] = 0
        elif kind == 2:
            req["count_per_period"][i] = -5
        else:
            req["period"][i] = 0
    return req


_CDF_CACHE = {}


def zipf_ranks(u, n_keys, s=1.0):
    """Inverse-CDF Zipf(s) over ranks 1..n_keys on a harmonic table (float64 cumsum)."""
    cdf = _CDF_CACHE.get((n_keys, s))
    if cdf is None:
        w = 1.0 / np.power(np.arange(1, n_keys + 1, dtype=np.float64), s)
        cdf = np.cumsum(w)
        cdf /= cdf[-1]
        _CDF_CACHE.clear()
        _CDF_CACHE[(n_keys, s)] = cdf
    return np.searchsorted(cdf, unit(u), side="left").astype(np.uint64)   # 0-based rank


def rank_to_key(rank, n_keys):
    """Fixed permutation of 0..n_keys-1 (multiplicative, odd multiplier coprime to n_keys)."""
    mult = 2654435761
    while np.gcd(mult, n_keys) != 1:
        mult += 2
    return (rank.astype(np.uint64) * np.uint64(mult) + np.uint64(12345)) % np.uint64(n_keys)


def ticks(config_id, n_keys, n_ticks, tick_size, key_fn, start_tick=0, tick_ns=1_000_000):
    """Generic tick trace: one `now` per tick advancing 1 ms, policy = key & 7 (fixed per key)."""
    n = n_ticks * tick_size
    start = start_tick * tick_size
    req = np.zeros(n, REQ_DTYPE)
    key = key_fn(stream(config_id, 0, n, start), stream(config_id, 1, n, start))
    req["key"] = key
    fill_policy(req, (key % np.uint64(8)).astype(np.int64))
    req["quantity"] = quantities(stream(config_id, 3, n, start))
    t = (np.arange(n, dtype=np.int64) // tick_size + start_tick + 1) * tick_ns
    req["now_ns"] = T0 + t
    return req


def config2(n_keys=10_000_000, n_ticks=64, tick_size=1 << 20, start_tick=0, _cache={}):
    """Zipf-1.0 over n_keys, ticks of tick_size requests."""
    def kf(u0, u1):
        return rank_to_key(zipf_ranks(u0, n_keys), n_keys)
    return ticks(2, n_keys, n_ticks, tick_size, kf, start_tick)


def config3(n_keys=100_000_000, n_ticks=4, tick_size=1 << 20, start_tick=0):
    """Uniform keys."""
    def kf(u0, u1):
        return u0 % np.uint64(n_keys)
    return ticks(3, n_keys, n_ticks, tick_size, kf, start_tick)


def config4(n_keys=10_000_000, n_ticks=16, tick_size=1 << 20, start_tick=0, hot=100):
    """50 % of the traffic on a fixed top-`hot` set, 50 % uniform over the rest."""
    def kf(u0, u1):
        hotkeys = rank_to_key(np.arange(hot, dtype=np.uint64), n_keys)
        is_hot = unit(u1) < 0.5
        cold = u0 % np.uint64(n_keys)
        return np.where(is_hot, hotkeys[(u0 % np.uint64(hot)).astype(np.int64)], cold)
    return ticks(4, n_keys, n_ticks, tick_size, kf, start_tick)


def warm_pass(n_keys, now_ns=T0):
    """One q=1 request per key (pre-inserts every key)."""
    req = np.zeros(n_keys, REQ_DTYPE)
    key = np.arange(n_keys, dtype=np.uint64)
    req["key"] = key
    fill_policy(req, (key % np.uint64(8)).astype(np.int64))
    req["quantity"] = 1
    req["now_ns"] = now_ns
    return req


def config2_rank_slice(n_keys, tick_size, first_tick, n_ticks, rank, world):
    """Rank `rank`'s slice of global Zipf ticks: global tick t has world*tick_size requests, rank r
    owns rows [r*tick_size, (r+1)*tick_size) of it (global index order = rank order inside a tick)."""
    out = np.zeros(n_ticks * tick_size, REQ_DTYPE)
    for j in range(n_ticks):
        t = first_tick + j
        start = (t * world + rank) * tick_size
        key = rank_to_key(zipf_ranks(stream(2, 0, tick_size, start), n_keys), n_keys)
        sl = out[j * tick_size:(j + 1) * tick_size]
        sl["key"] = key
        fill_policy(sl, (key % np.uint64(8)).astype(np.int64))
        sl["quantity"] = quantities(stream(2, 3, tick_size, start))
        sl["now_ns"] = T0 + (t + 1) * 1_000_000
    return out
