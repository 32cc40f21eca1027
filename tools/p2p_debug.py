"""Debug helper: loop-back peer-memory pipeline on one GPU, mismatch statistics against the oracle."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle, traces
import throttlecrab_b200 as tc
from throttlecrab_b200.sharded import connect_local
world, tick, hot, n_keys, n_ticks = int(sys.argv[1]), int(sys.argv[2]), 20, 40_000, int(sys.argv[3])
dev = torch.device("cuda", 0)
kh = tc.hash_key_ids(np.arange(n_keys, dtype=np.uint64))
lims = [tc.RateLimiter(tc.ManualStore(capacity=4_000_000, device=0, created_ns=traces.T0, max_batch=4 * tick)) for _ in range(world)]
connect_local(lims, tick)
glob = traces.config4(n_keys=n_keys, n_ticks=n_ticks, tick_size=tick * world, hot=hot)
stream = torch.cuda.Stream(dev); torch.cuda.set_stream(stream)
ress = [[None] * n_ticks for _ in range(world)]; keep = []
for t in range(n_ticks):
    for r in range(world):
        sl = glob[t * tick * world:(t + 1) * tick * world][r * tick:(r + 1) * tick]
        e = np.empty(tick, tc.REQ_DTYPE); e["key_hash"] = kh[sl["key"].astype(np.int64)]
        for f in ("max_burst", "count_per_period", "period", "quantity", "now_ns"): e[f] = sl[f]
        q = torch.from_numpy(e.view(np.uint8)).to(dev); keep.append(q)
        ress[r][t] = torch.zeros(tick * 32, dtype=torch.uint8, device=dev)
        lims[r].store._check(lims[r]._L.gcra_p2p_submit_route(lims[r]._h, tick, q.data_ptr(), stream.cuda_stream))
    for r in range(world):
        lims[r].store._check(lims[r]._L.gcra_p2p_submit_finish(lims[r]._h, ress[r][t].data_ptr()))
for lim in lims: lim.store._check(lim._L.gcra_p2p_join(lim._h, stream.cuda_stream))
torch.cuda.synchronize()
want = oracle.OracleStore(oracle.PERIODIC, capacity=n_keys, created_ns=traces.T0, p0=10**9).replay(glob)
got = np.empty(len(glob), oracle.RES_DTYPE)
for r in range(world):
    for t in range(n_ticks):
        a = t * tick * world + r * tick
        got[a:a + tick] = ress[r][t].cpu().numpy().view(oracle.RES_DTYPE)
bad = (got.view(np.uint8).reshape(len(got), -1) != want.view(np.uint8).reshape(len(want), -1)).any(axis=1)
L = lims[0]._L
own = np.array([L.gcra_owner_of(int(h), world) for h in kh[glob["key"].astype(np.int64)][:200000]])
idx = np.arange(len(glob)); tk = idx // (tick * world); rk = (idx % (tick * world)) // tick
import ctypes as C
for lim in lims:
    e = C.c_uint32(); lim._L.gcra_p2p_error(lim._h, C.byref(e)); print("error flag", e.value)
print("mismatches", bad.sum(), "of", len(bad))
for t in range(n_ticks):
    for r in range(world):
        m = (tk == t) & (rk == r)
        print("tick", t, "sender", r, "bad", int(bad[m].sum()), "of", int(m.sum()))
m = np.nonzero(bad)[0][:8]
for i in m: print(i, "owner", own[i] if i < len(own) else "?", "want", want[i], "got", got[i], "req", glob[i])
zero = (got.view(np.uint8).reshape(len(got), -1) == 0).all(axis=1)
print("all-zero result rows:", int(zero.sum()), "bad&zero", int((bad & zero).sum()))
nb = min(len(own), len(bad))
for o in range(world): print("owner", o, "bad", int(bad[:nb][own[:nb] == o].sum()), "of", int((own[:nb] == o).sum()))
