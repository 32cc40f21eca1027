"""Child process of tests/test_gpu_sharded.py::test_peer_memory_pipeline_loopback_on_one_gpu: `world` engines in one
process on cuda:0, the peer-memory pipeline (gcra_p2p_*) between them, results against ONE oracle store."""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle  # noqa: E402
import throttlecrab_b200 as tc  # noqa: E402
import traces  # noqa: E402
from throttlecrab_b200.sharded import connect_local  # noqa: E402

world, tick, hot = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
n_keys, n_ticks = 40_000, 6
dev = torch.device("cuda", 0)
key_hash_of = tc.hash_key_ids(np.arange(n_keys, dtype=np.uint64))
lims = [tc.RateLimiter(tc.ManualStore(capacity=4_000_000, device=0, created_ns=traces.T0, max_batch=4 * tick))
        for _ in range(world)]
connect_local(lims, tick)
glob = traces.config4(n_keys=n_keys, n_ticks=n_ticks, tick_size=tick * world, hot=hot)
stream = torch.cuda.Stream(dev)
torch.cuda.set_stream(stream)
ress = [[None] * n_ticks for _ in range(world)]
keep = []
for t in range(n_ticks):
    for r in range(world):
        sl = glob[t * tick * world:(t + 1) * tick * world][r * tick:(r + 1) * tick]
        e = np.empty(tick, tc.REQ_DTYPE)
        e["key_hash"] = key_hash_of[sl["key"].astype(np.int64)]
        for f in ("max_burst", "count_per_period", "period", "quantity", "now_ns"):
            e[f] = sl[f]
        q = torch.from_numpy(e.view(np.uint8)).to(dev)
        ress[r][t] = torch.zeros(tick * 32, dtype=torch.uint8, device=dev)
        keep.append(q)
        lims[r].store._check(lims[r]._L.gcra_p2p_submit_route(lims[r]._h, tick, q.data_ptr(), stream.cuda_stream))
    for r in range(world):        # every engine's route is enqueued before the first kernel that waits for it
        lims[r].store._check(lims[r]._L.gcra_p2p_submit_finish(lims[r]._h, ress[r][t].data_ptr()))
for lim in lims:
    lim.store._check(lim._L.gcra_p2p_join(lim._h, stream.cuda_stream))
torch.cuda.synchronize()
for lim in lims:
    e = C.c_uint32()
    lim.store._check(lim._L.gcra_p2p_error(lim._h, C.byref(e)))
    assert e.value == 0, "a wait gave up"
want = oracle.OracleStore(oracle.PERIODIC, capacity=n_keys, created_ns=traces.T0, p0=10**9).replay(glob)
got = np.empty(len(glob), oracle.RES_DTYPE)
for r in range(world):
    for t in range(n_ticks):
        a = t * tick * world + r * tick
        got[a:a + tick] = ress[r][t].cpu().numpy().view(oracle.RES_DTYPE)
bad = np.nonzero((got.view(np.uint8).reshape(len(got), -1) != want.view(np.uint8).reshape(len(want), -1)).any(axis=1))[0]
assert len(bad) == 0, (len(bad), int(bad[0]), want[bad[0]], got[bad[0]], glob[bad[0]])
assert want["allowed"].sum() > 0 and (want["allowed"] == 0).sum() > 0
for lim in lims:
    lim.store.close()
print("loopback ok: %d rows, %d engines" % (len(glob), world))
