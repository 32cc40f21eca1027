"""Per-phase timing of one sharded tick (run under torchrun, N>=2): where does the time go?"""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import throttlecrab_b200 as tc  # noqa: E402
import traces  # noqa: E402
from throttlecrab_b200.sharded import ShardedLimiter, REQ_B, RES_B  # noqa: E402

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
TICK = 1 << 20
n_keys = 2_000_000 * world
kh = tc.hash_key_ids(np.arange(n_keys, dtype=np.uint64))
lim = tc.RateLimiter(tc.ManualStore(capacity=n_keys // world, device=lr, created_ns=traces.T0, max_batch=2 * TICK))
sh = ShardedLimiter(lim, dist, dev)
st = torch.cuda.Stream(dev)
torch.cuda.set_stream(st)
tr = traces.config2_rank_slice(n_keys, TICK, 0, 12, rank, world)
e = np.empty(len(tr), tc.REQ_DTYPE)
e["key_hash"] = kh[tr["key"].astype(np.int64)]
for f in ("max_burst", "count_per_period", "period", "quantity", "now_ns"):
    e[f] = tr[f]
d_req = torch.from_numpy(e.view(np.uint8).reshape(12, TICK * 48)).to(dev)
d_res = torch.empty((12, TICK * 32), dtype=torch.uint8, device=dev)
slot = sh.slots[0]
W = world


def timed(name, fn, acc):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    r = fn()
    b.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    acc.setdefault(name, []).append((a.elapsed_time(b), (t1 - t0) * 1e3, (t2 - t0) * 1e3))
    return r


acc = {}
for i in range(12):
    n = TICK
    timed("partition", lambda: sh.ops.partition(n, d_req[i], W, slot.routed, slot.src_index, slot.counts, st.cuda_stream), acc)

    def counts():
        send = slot.counts[:W].to(torch.int64)
        recv = torch.empty_like(send)
        dist.all_to_all_single(recv, send)
        return send.tolist(), recv.tolist()
    send_l, recv_l = timed("counts_a2a+tolist", counts, acc)
    n_recv = sum(recv_l)
    timed("req_a2a", lambda: dist.all_to_all_single(slot.recv_req[:n_recv * REQ_B], slot.routed[:n * REQ_B],
                                                    output_split_sizes=[c * REQ_B for c in recv_l],
                                                    input_split_sizes=[c * REQ_B for c in send_l]), acc)
    timed("decide", lambda: sh.ops.decide(n_recv, slot.recv_req, slot.recv_res, st.cuda_stream), acc)
    timed("res_a2a", lambda: dist.all_to_all_single(slot.back_res[:n * RES_B], slot.recv_res[:n_recv * RES_B],
                                                    output_split_sizes=[c * RES_B for c in send_l],
                                                    input_split_sizes=[c * RES_B for c in recv_l]), acc)
    timed("unpermute", lambda: sh.ops.unpermute(n, slot.back_res, slot.src_index, d_res[i], st.cuda_stream), acc)
if rank == 0:
    print("phase: gpu_ms / host_enqueue_ms / host_until_done_ms   (median of last 8 ticks), n_recv=%d" % n_recv)
    for k, v in acc.items():
        m = np.median(np.array(v[4:]), axis=0)
        print("%-20s %7.3f %7.3f %7.3f" % (k, m[0], m[1], m[2]))
dist.destroy_process_group()
