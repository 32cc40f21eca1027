"""K2 (expired-key sweep) on BASELINE.json configs[2]: 100 M keys uniform, sweep at several expired
fractions; achieved HBM GB/s against MEASURED_PEAKS.json.  Prints one JSON line per sweep.

Algorithmic bytes: 16 B per table slot scanned (tat + burst offset) + 16 B per evicted
entry (the pair reset; keys are reclaimed lazily by purge_kernel)."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import throttlecrab_b200 as tc  # noqa: E402
import traces  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--keys", type=int, default=100_000_000)
ap.add_argument("--reps", type=int, default=5)
args = ap.parse_args()
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    peak = 6650.0

B = 1 << 20
st = tc.ManualStore(capacity=args.keys, created_ns=traces.T0, max_batch=B)
lim = tc.RateLimiter(st)


def fill():
    for a in range(0, args.keys, B):
        ids = np.arange(a, min(a + B, args.keys), dtype=np.uint64)
        req = np.empty(len(ids), tc.REQ_DTYPE)
        req["key_hash"] = tc.hash_key_ids(ids)
        p = traces.POLICIES[(ids % 8).astype(np.int64)]
        req["max_burst"], req["count_per_period"], req["period"] = p[:, 0], p[:, 1], p[:, 2]
        req["quantity"] = 1
        req["now_ns"] = traces.T0
        lim.rate_limit_batch(req)


fill()
slots = st.stats()["table_slots"]
# expiry = T0 + dvt per policy: P6 0 s, P5 0.5 s, P2 5.4 s, P0 5.94 s, P4 17.1 s, P3 24 s, P1 356 s, P7 594 s
plan = [(-1, "0 %"), (-1, "0 % (repeat, nothing expired)"), (1_000_000_000, "25 %"),
        (10_000_000_000, "25 % more (50 % total)"), (700_000_000_000, "remaining 50 %")]
for dt, label in plan:
    before = st.len()
    removed = st.sweep(traces.T0 + dt)
    ms = st.last_sweep_ms()
    alg = 16.0 * slots + 16.0 * removed
    print(json.dumps({"kernel": "sweep_kernel", "keys": args.keys, "table_slots": slots, "live_before": before,
                      "expired": label, "removed": removed, "ms": ms, "achieved_GBps": alg / ms / 1e6,
                      "peak_GBps": peak, "frac": alg / ms / 1e6 / peak}), flush=True)
# steady-state scan of an all-empty table (nothing to evict), several repetitions
for _ in range(args.reps):
    st.sweep(traces.T0 + 800_000_000_000)
    ms = st.last_sweep_ms()
    print(json.dumps({"kernel": "sweep_kernel", "expired": "empty table", "ms": ms,
                      "achieved_GBps": 16.0 * slots / ms / 1e6, "frac": 16.0 * slots / ms / 1e6 / peak}), flush=True)
