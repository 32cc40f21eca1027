"""Differential replay: the CUDA engine (batched, through the C ABI) against the CPU oracle on the
same seeded traces -- bit-exact on (status, allowed, remaining, reset_after_ns, retry_after_ns)
and on the final table state."""
import numpy as np
import pytest

import oracle
import throttlecrab_b200 as tc
import traces
from gpu_util import (assert_universe_collision_free, engine_requests, first_mismatch, replay_both)

pytestmark = pytest.mark.gpu


def _check_tables(st_o, st_e, key_ids, hashes):
    for k in key_ids:
        eo = st_o.entry("k:%d" % k)
        ee = st_e.peek(int(hashes[k]))
        assert eo == ee, (k, eo, ee)


@pytest.mark.parametrize("batch", [4096, 1, 37, 100_000])
def test_config1_correctness_replay(batch):
    """BASELINE.json configs[0]: 1 K keys, 100 K requests, per-request clock."""
    n = 100_000 if batch != 1 else 300
    req = traces.config1(n=100_000)[:n]
    hashes = assert_universe_collision_free(1000)
    res_o, res_e, st_o, st_e = replay_both(req, capacity=1000, batch=batch)
    assert first_mismatch(res_o, res_e, req) is None, first_mismatch(res_o, res_e, req)
    assert (res_o["status"] != 0).sum() >= (90 if n == 100_000 else 0)
    _check_tables(st_o, st_e, range(0, 1000, 7 if batch != 4096 else 1), hashes)
    assert st_e.len() == st_o.len()


@pytest.mark.parametrize("store_cls,okind", [(tc.PeriodicStore, oracle.PERIODIC),
                                             (tc.ProbabilisticStore, oracle.PROBABILISTIC),
                                             (tc.AdaptiveStore, oracle.ADAPTIVE)])
def test_config1_with_sweeping_stores(store_cls, okind):
    """Same replay while both sides sweep on their own policy: results must not change."""
    req = traces.config1(n=100_000)
    res_o, res_e, st_o, st_e = replay_both(req, capacity=1000, batch=4096, store_cls=store_cls,
                                           oracle_kind=okind, oracle_p0=0)
    assert first_mismatch(res_o, res_e, req) is None, first_mismatch(res_o, res_e, req)
    assert st_e.stats()["sweeps"] > 0 and st_o.sweeps() > 0


def test_hot_key_contention():
    """BASELINE.json configs[3] shape at oracle-friendly size: half the traffic on 100 keys,
    thousands of same-key requests inside one batch (order-exact)."""
    n_keys = 50_000
    req = traces.config4(n_keys=n_keys, n_ticks=6, tick_size=1 << 17)
    assert_universe_collision_free(n_keys)
    res_o, res_e, _, _ = replay_both(req, capacity=n_keys, batch=1 << 17)
    assert first_mismatch(res_o, res_e, req) is None, first_mismatch(res_o, res_e, req)
    assert res_o["allowed"].sum() > 0 and (res_o["allowed"] == 0).sum() > 0


def test_zipf_ticks():
    """BASELINE.json configs[1] shape at oracle-friendly size: Zipf-1.0 ticks after a warm pass."""
    n_keys = 200_000
    warm = traces.warm_pass(n_keys)
    req = np.concatenate([warm, traces.config2(n_keys=n_keys, n_ticks=8, tick_size=1 << 17)])
    assert_universe_collision_free(n_keys)
    res_o, res_e, st_o, st_e = replay_both(req, capacity=n_keys, batch=1 << 17)
    assert first_mismatch(res_o, res_e, req) is None, first_mismatch(res_o, res_e, req)
    assert st_e.len() == st_o.len()


def test_single_key_long_run_mixed_parameters():
    """One key, 40 000 requests in one batch with varying quantity, policy and clock: the run spans
    >1000 warp chunks and every request may change the state."""
    n = 40_000
    rng = np.random.default_rng(3)
    req = np.zeros(n, traces.REQ_DTYPE)
    req["key"] = 5
    traces.fill_policy(req, rng.integers(0, 8, n))
    req["quantity"] = rng.choice([0, 1, 1, 1, 2, 5], n)
    req["now_ns"] = traces.T0 + np.cumsum(rng.choice([0, 0, 1000, 50_000_000, 2_000_000_000], n))
    res_o, res_e, _, _ = replay_both(req, capacity=100, batch=n)
    assert first_mismatch(res_o, res_e, req) is None, first_mismatch(res_o, res_e, req)


def test_extreme_parameters_batch():
    """Saturation / truncation corners of rate_limiter.rs:120-238 in one batch, several keys."""
    I = 2**63 - 1
    rows = [(0, I, I, I, 1), (1, 1, 1, 1, 1), (2, I // 1000, 100, 60, 1), (3, 10, I // 1000, 60, 1),
            (4, 10, 10, 60, I // 2), (4, 10, 10, 60, 1), (5, 1, 1, 1, 0), (5, 1, 1, 1, 1),
            (6, 2**32, 1, I, 1), (7, 2**32 + 1, 7, 60, 3), (8, 5, 1, I, 1), (8, 5, 1, I, 1),
            (9, 3, I, 1, 1), (9, 3, I, 1, 1), (10, 10, 10, 60, -1), (11, 0, 10, 60, 1),
            (12, 10, -1, 60, 1), (13, 10, 10, -7, 1), (0, I, I, I, I), (0, I, I, I, 0)]
    req = np.zeros(len(rows) * 3, traces.REQ_DTYPE)
    for rep in range(3):
        for j, (k, b, c, p, q) in enumerate(rows):
            i = rep * len(rows) + j
            req[i] = (k, b, c, p, q, traces.T0 + rep * 5_000_000_000)
    res_o, res_e, _, _ = replay_both(req, capacity=100, batch=len(req))
    assert first_mismatch(res_o, res_e, req) is None, first_mismatch(res_o, res_e, req)
    assert set(np.unique(res_o["status"])) == {0, 1, 2, 3}


def test_empty_and_ragged_batches():
    st = tc.ManualStore(capacity=100, created_ns=traces.T0, max_batch=4096)
    lim = tc.RateLimiter(st)
    assert len(lim.rate_limit_batch(np.zeros(0, tc.REQ_DTYPE))) == 0
    req = traces.config1(n=10_000)
    ereq = engine_requests(req)
    sto = oracle.OracleStore(oracle.PERIODIC, capacity=100, created_ns=traces.T0, p0=10**9)
    res_o = sto.replay(req)
    res_e = np.empty(len(req), tc.RES_DTYPE)
    a = 0
    sizes = [1, 31, 32, 33, 255, 256, 257, 1023, 1025, 4096, 5000]   # 5000 > max_batch: chunked inside
    while a < len(req):
        m = sizes[(a * 7) % len(sizes)]
        lim.rate_limit_batch(ereq[a:a + m], out=res_e[a:a + m])
        a += m
    assert first_mismatch(res_o, res_e, req) is None, first_mismatch(res_o, res_e, req)


def test_compact_requests_match_full_requests():
    """gcra_rate_limit_batch16 (policy table + per-call now) == the 48-byte path on the same trace."""
    n_keys = 20_000
    req = traces.config2(n_keys=n_keys, n_ticks=4, tick_size=1 << 15)
    ereq = engine_requests(req)
    a = tc.RateLimiter(tc.ManualStore(capacity=n_keys, created_ns=traces.T0, max_batch=1 << 15))
    b = tc.RateLimiter(tc.ManualStore(capacity=n_keys, created_ns=traces.T0, max_batch=1 << 15))
    pol = np.zeros(8, tc.POLICY_DTYPE)
    pol["max_burst"], pol["count_per_period"], pol["period"] = traces.POLICIES.T
    b.set_policies(pol)
    for t in range(4):
        sl = slice(t << 15, (t + 1) << 15)
        ra = a.rate_limit_batch(ereq[sl])
        r16 = np.zeros(1 << 15, tc.REQ16_DTYPE)
        r16["key_hash"] = ereq["key_hash"][sl]
        r16["quantity"] = req["quantity"][sl]
        r16["policy"] = req["key"][sl] % 8
        rb = b.rate_limit_batch16(r16, int(req["now_ns"][sl][0]))
        assert ra.tobytes() == rb.tobytes()


def test_ring_matches_blocking_calls():
    n_keys = 20_000
    req = traces.config2(n_keys=n_keys, n_ticks=8, tick_size=1 << 14)
    ereq = engine_requests(req)
    a = tc.RateLimiter(tc.ManualStore(capacity=n_keys, created_ns=traces.T0, max_batch=1 << 14))
    b = tc.RateLimiter(tc.ManualStore(capacity=n_keys, created_ns=traces.T0, max_batch=1 << 14))
    ring = tc.Ring(b, slots=3, slot_capacity=1 << 14)
    want = a.rate_limit_batch(ereq)
    got = np.empty_like(want)
    T = 1 << 14
    for t in range(8 + 3):
        if t >= 3:
            s = (t - 3) % 3
            ring.wait(s)
            got[(t - 3) * T:(t - 2) * T] = ring.res[s][:T]
        if t < 8:
            s = t % 3
            ring.req[s][:T] = ereq[t * T:(t + 1) * T]
            ring.submit(s, T)
    assert want.tobytes() == got.tobytes()


def test_sweep_removes_exactly_the_expired():
    n_keys = 30_000
    st = tc.ManualStore(capacity=n_keys, created_ns=traces.T0, max_batch=1 << 15)
    lim = tc.RateLimiter(st)
    warm = engine_requests(traces.warm_pass(n_keys))
    lim.rate_limit_batch(warm)
    assert st.len() == n_keys
    sto = oracle.OracleStore(oracle.PERIODIC, capacity=n_keys, created_ns=traces.T0, p0=10**9)
    sto.replay(traces.warm_pass(n_keys))
    for dt in (0, 1_000_000_000, 30_000_000_000, 4_000_000_000_000):
        removed = st.sweep(traces.T0 + dt)
        assert removed == sto.force_sweep(traces.T0 + dt)
        assert st.len() == sto.len()
    assert st.len() == 0 or st.len() < n_keys


@pytest.mark.parametrize("batch", [64, 4096])
def test_tight_table_stash_and_growth(batch):
    """GCRA_FLAG_TIGHT_TABLE: a 64-slot table filled to 15/16 and grown repeatedly while 20 000 keys
    arrive -- second-choice buckets, stash probing, tombstones after sweeps and rehash all on the path."""
    n_keys = 20_000
    rng = np.random.default_rng(5)
    n = 120_000
    req = np.zeros(n, traces.REQ_DTYPE)
    # keys arrive progressively so the table grows many times while old keys stay hot
    req["key"] = (rng.integers(0, n_keys, n) * np.linspace(0.01, 1.0, n)).astype(np.uint64)
    traces.fill_policy(req, (req["key"] % np.uint64(8)).astype(np.int64))
    req["quantity"] = rng.choice([0, 1, 1, 1, 2, 5], n)
    req["now_ns"] = traces.T0 + np.cumsum(rng.choice([0, 1000, 2_000_000, 700_000_000], n))
    st_o = oracle.OracleStore(oracle.PERIODIC, capacity=16, created_ns=traces.T0, p0=10**9)
    res_o = st_o.replay(req)
    st_e = tc.ManualStore(capacity=16, created_ns=traces.T0, max_batch=4096, flags=1)
    lim = tc.RateLimiter(st_e)
    ereq = engine_requests(req)
    res_e = np.empty(n, tc.RES_DTYPE)
    swept = 0
    every = max(1, 6000 // batch)
    for a in range(0, n, batch):
        lim.rate_limit_batch(ereq[a:a + batch], out=res_e[a:a + batch])
        if (a // batch) % every == every - 1:      # sweeps leave holes and stash tombstones behind
            swept += st_e.sweep(int(req["now_ns"][min(a + batch, n) - 1]))
    assert first_mismatch(res_o, res_e, req) is None, first_mismatch(res_o, res_e, req)
    s = st_e.stats()
    assert s["grows"] >= 1 and swept > 0
    assert s["purges"] >= 1                       # swept keys' slots were reclaimed before growing again
    if batch == 64:
        assert s["stash_entries"] > 0             # keys really live in the stash
    assert s["table_slots"] < 4 * n_keys          # stayed tight: load well above the production 0.5


def test_pipelined_and_mixed_submission_match_blocking():
    """gcra_rate_limit_batch_device_pipelined (front half of batch i+1 overlapping the decide kernels of
    batch i), mixed with in-stream batches, gives bit-identical results to blocking host batches."""
    import torch
    n_keys, T, n_ticks = 30_000, 1 << 14, 12
    req = traces.config4(n_keys=n_keys, n_ticks=n_ticks, tick_size=T, hot=50)
    ereq = engine_requests(req)
    a = tc.RateLimiter(tc.ManualStore(capacity=n_keys, created_ns=traces.T0, max_batch=T))
    want = a.rate_limit_batch(ereq)
    b = tc.RateLimiter(tc.ManualStore(capacity=n_keys, created_ns=traces.T0, max_batch=T))
    dev = torch.device("cuda", 0)
    st = torch.cuda.Stream(dev)
    with torch.cuda.stream(st):
        d_req = torch.from_numpy(ereq.view(np.uint8).reshape(n_ticks, T * 48)).to(dev)
        d_res = torch.zeros((n_ticks, T * 32), dtype=torch.uint8, device=dev)
        st.synchronize()
        for t in range(n_ticks):
            if t % 5 == 3:      # an in-stream batch in the middle of pipelined ones
                b.rate_limit_batch_device(T, d_req[t].data_ptr(), d_res[t].data_ptr(), st.cuda_stream)
            else:
                b.submit_device(T, d_req[t].data_ptr(), d_res[t].data_ptr(), st.cuda_stream)
        b.join(st.cuda_stream)
        st.synchronize()
    got = d_res.cpu().numpy().view(tc.RES_DTYPE).reshape(-1)
    assert first_mismatch(want, got, req) is None, first_mismatch(want, got, req)


def test_snapshot_restore_continues_identically(tmp_path):
    """gcra_snapshot_save / _load: an engine restarted from a snapshot (even one created with another
    capacity) continues with exactly the answers of the oracle that never stopped."""
    n_keys = 8000
    req = traces.config1(n=60_000, keys=n_keys)
    ereq = engine_requests(req)
    want = oracle.OracleStore(oracle.PERIODIC, capacity=n_keys, created_ns=traces.T0, p0=10**9).replay(req)
    a = tc.RateLimiter(tc.ManualStore(capacity=n_keys, created_ns=traces.T0, max_batch=4096))
    got = np.empty(len(req), tc.RES_DTYPE)
    half = 30_000
    for s in range(0, half, 4096):
        e = min(s + 4096, half)
        a.rate_limit_batch(ereq[s:e], out=got[s:e])
    path = str(tmp_path / "table.gcra")
    a.store.save(path)
    len_before = a.store.len()
    a.store.close()
    b = tc.RateLimiter(tc.ManualStore(capacity=100, created_ns=traces.T0, max_batch=4096))   # other geometry
    b.store.load(path)
    assert b.store.len() == len_before
    for s in range(half, len(req), 4096):
        e = min(s + 4096, len(req))
        b.rate_limit_batch(ereq[s:e], out=got[s:e])
    assert first_mismatch(want, got, req) is None, first_mismatch(want, got, req)
