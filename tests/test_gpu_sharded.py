"""Two real GPUs: the hash-sharded engine (partition kernel -> NCCL all-to-all -> decide -> all-to-all
-> unpermute, pipelined) must give exactly the answers of ONE oracle store fed the union trace in global
index order.  Skipped on boxes with fewer than 2 GPUs (run with `gpurun --gpus 2`)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_keys, n_ticks, tick, ret, native=False):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    import throttlecrab_b200 as tc
    import traces
    from throttlecrab_b200.sharded import NativeShardedLimiter, ShardedLimiter
    key_hash_of = tc.hash_key_ids(np.arange(n_keys, dtype=np.uint64))
    lim = tc.RateLimiter(tc.ManualStore(capacity=n_keys, device=rank, created_ns=traces.T0, max_batch=2 * tick))
    sh = NativeShardedLimiter(lim, dist, dev) if native else ShardedLimiter(lim, dist, dev)
    glob = traces.config4(n_keys=n_keys, n_ticks=n_ticks, tick_size=tick * world, hot=20)
    stream = torch.cuda.Stream(dev)
    torch.cuda.set_stream(stream)
    reqs, ress = [], []
    for t in range(n_ticks):
        sl = glob[t * tick * world:(t + 1) * tick * world][rank * tick:(rank + 1) * tick]
        e = np.empty(tick, tc.REQ_DTYPE)
        e["key_hash"] = key_hash_of[sl["key"].astype(np.int64)]
        for f in ("max_burst", "count_per_period", "period", "quantity", "now_ns"):
            e[f] = sl[f]
        reqs.append(torch.from_numpy(e.view(np.uint8)).to(dev))
        ress.append(torch.empty(tick * 32, dtype=torch.uint8, device=dev))
    for t in range(n_ticks):
        sh.submit(reqs[t], ress[t])            # pipelined across ticks
    sh.finish()
    torch.cuda.synchronize()
    ret[rank] = np.concatenate([r.cpu().numpy() for r in ress]).tobytes()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(600)
@pytest.mark.parametrize("native", [False, True], ids=["torch_distributed", "native_nccl"])
def test_two_gpu_sharded_matches_single_oracle(native):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import oracle
    import traces
    world, n_keys, n_ticks, tick = 2, 50_000, 6, 1 << 16
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), n_keys, n_ticks, tick, ret, native), nprocs=world, join=True)
    glob = traces.config4(n_keys=n_keys, n_ticks=n_ticks, tick_size=tick * world, hot=20)
    want = oracle.OracleStore(oracle.PERIODIC, capacity=n_keys, created_ns=traces.T0, p0=10**9).replay(glob)
    got = np.empty(len(glob), oracle.RES_DTYPE)
    for r in range(world):
        pr = np.frombuffer(ret[r], oracle.RES_DTYPE).reshape(n_ticks, tick)
        for t in range(n_ticks):
            a = t * tick * world + r * tick
            got[a:a + tick] = pr[t]
    assert got.tobytes() == want.tobytes()
