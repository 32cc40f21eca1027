"""Latency of small calls through the C ABI (host buffers): one RateLimiter::rate_limit call, and batches of
16 / 64 / 255 / 256 / 4096 requests.  GCRA_NO_SMALL=1 disables the single-CTA small-batch kernel (A/B)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import throttlecrab_b200 as tc  # noqa: E402
import traces  # noqa: E402

st = tc.ManualStore(capacity=100_000, created_ns=traces.T0, max_batch=4096)
lim = tc.RateLimiter(st)
req = traces.config1(n=200_000, keys=50_000)
ereq = np.empty(len(req), tc.REQ_DTYPE)
ereq["key_hash"] = tc.hash_key_ids(req["key"])
for f in ("max_burst", "count_per_period", "period", "quantity", "now_ns"):
    ereq[f] = req[f]
out = {"no_small": os.environ.get("GCRA_NO_SMALL", "0")}
for i in range(200):
    lim.rate_limit("warm%d" % i, 10, 100, 60, 1, traces.T0)
t = time.perf_counter()
for i in range(2000):
    lim.rate_limit("k%d" % (i % 500), 10, 100, 60, 1, traces.T0 + i)
out["single_call_us"] = (time.perf_counter() - t) / 2000 * 1e6
pos = 0
for n in (16, 64, 255, 256, 4096):
    reps = 300
    res = np.empty(n, tc.RES_DTYPE)
    t = time.perf_counter()
    for _ in range(reps):
        lim.rate_limit_batch(ereq[pos:pos + n], out=res)
        pos = (pos + n) % (len(ereq) - 4096)
    out["batch_%d_us" % n] = (time.perf_counter() - t) / reps * 1e6
print(json.dumps(out))
