"""Per-stage device time of the peer-memory sharded tick (run under torchrun, N >= 2): ticks are run ONE AT A TIME
(submit, join, barrier) with the library's stage events on, so the stages do not overlap and every number is a plain
stage duration.  Rank 0 prints one JSON line with the per-rank medians; the pipelined per-tick time is printed beside
it for comparison.  Usage: torchrun --nproc-per-node N tools/sharded_probe.py [keys_per_gpu=10000000] [ticks=12]"""
import ctypes as C
import json
import os
import sys

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import throttlecrab_b200 as tc  # noqa: E402
import traces  # noqa: E402
from throttlecrab_b200.sharded import PeerShardedLimiter  # noqa: E402

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
TICK = 1 << 20
n_local = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
NT = int(sys.argv[2]) if len(sys.argv) > 2 else 12
n_keys = n_local * world
kh = tc.hash_key_ids(np.arange(n_keys, dtype=np.uint64))
store = tc.ManualStore(capacity=n_local, device=lr, created_ns=traces.T0, max_batch=2 * TICK)
lim = tc.RateLimiter(store)
sh = PeerShardedLimiter(lim, dist, dev)
st = torch.cuda.Stream(dev)
torch.cuda.set_stream(st)


def rows(tr):
    e = np.empty(len(tr), tc.REQ_DTYPE)
    e["key_hash"] = kh[tr["key"].astype(np.int64)]
    for f in ("max_burst", "count_per_period", "period", "quantity", "now_ns"):
        e[f] = tr[f]
    return e


# warm pass: rank r inserts keys [r*n_local, (r+1)*n_local) through the sharded path
wres = torch.empty(TICK * 32, dtype=torch.uint8, device=dev)
for a in range(0, n_local, TICK):
    ids = np.arange(rank * n_local + a, rank * n_local + min(a + TICK, n_local), dtype=np.uint64)
    w = np.zeros(TICK, traces.REQ_DTYPE)
    w["key"][:len(ids)] = ids
    w["key"][len(ids):] = ids[-1]
    traces.fill_policy(w, (w["key"] % np.uint64(8)).astype(np.int64))
    w["quantity"] = 1
    w["now_ns"] = traces.T0
    sh.step(torch.from_numpy(rows(w).view(np.uint8)).to(dev), wres)
torch.cuda.synchronize()
tr = traces.config2_rank_slice(n_keys, TICK, 0, NT, rank, world)
d_req = torch.from_numpy(rows(tr).view(np.uint8).reshape(NT, TICK * 48)).to(dev)
d_res = torch.empty((NT, TICK * 32), dtype=torch.uint8, device=dev)
L, h = lim._L, lim._h
# ---- serial ticks with stage events
store._check(L.gcra_p2p_set_timing(h, 1))
stages = []
for t in range(NT // 2):
    torch.cuda.synchronize(); dist.barrier()
    sh.step(d_req[t], d_res[t])
    torch.cuda.synchronize()
    out = (C.c_float * 5)()
    store._check(L.gcra_p2p_last_tick_ms(h, C.byref(out)))
    stages.append([float(x) for x in out])
store._check(L.gcra_p2p_set_timing(h, 0))
# ---- pipelined ticks
torch.cuda.synchronize(); dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(st)
for t in range(NT // 2, NT):
    sh.submit(d_req[t], d_res[t])
sh.finish()
e1.record(st)
torch.cuda.synchronize()
pipe = torch.tensor([e0.elapsed_time(e1) / (NT - NT // 2)], device=dev)
dist.all_reduce(pipe, op=dist.ReduceOp.MAX)
med = np.median(np.array(stages[2:] if len(stages) > 3 else stages), axis=0)
stats = store.stats()
mine = {"rank": rank, "route": med[0], "wait_rows": med[1], "engine": med[2], "wait_results": med[3], "unpermute": med[4],
        "serial_total": float(med.sum()), "residue_fraction": stats["residue_rows"] / max(stats["residue_batches"], 1) / TICK}
allr = [None] * world if rank == 0 else None
dist.gather_object(mine, allr, dst=0)
if rank == 0:
    print(json.dumps({"sharded_probe": {"world": world, "keys_per_gpu": n_local, "tick": TICK,
                                        "pipelined_ms_per_tick_max_over_ranks": float(pipe.item()),
                                        "serial_stage_ms_median_per_rank": allr}}), flush=True)
dist.barrier()
dist.destroy_process_group()
