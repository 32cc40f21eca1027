"""Hash-sharded multi-GPU engine: one process per GPU, key space partitioned by owner shard.

The reference is single-process; its docs recommend "client-side sharding by key"
(README.md:247-249).  Here every rank ingests an arbitrary slice of a tick, a stable partition
kernel (gcra_route_partition) groups the slice by owner, one all-to-all moves each group to its
owning GPU, the owner decides locally (gcra_rate_limit_batch_device), a second all-to-all returns the
results and gcra_route_unpermute puts them back in input order.

Ordering rule: inside a tick, requests are applied in GLOBAL index order, where rank r's slice
precedes rank r+1's.  The stable partition keeps slice order inside every group and all-to-all
concatenates the groups in source-rank order, so the owner sees each key's requests in that order.

torch is plumbing only: device buffers, the NCCL process group and stream ordering.
"""
import ctypes as C

import torch

from . import _native

REQ_B, RES_B = 48, 32


class CudaOps:
    """The three device steps, all through the C ABI."""

    def __init__(self, limiter):
        self.lim = limiter
        self.L, self.h = limiter._L, limiter._h

    def _check(self, rc):
        self.lim.store._check(rc)

    def partition(self, n, req, world, routed, src_index, counts, stream):
        self._check(self.L.gcra_route_partition(self.h, n, req.data_ptr(), world, routed.data_ptr(),
                                                src_index.data_ptr(), counts.data_ptr(), stream))

    def decide(self, n, req, res, stream):
        self._check(self.L.gcra_rate_limit_batch_device(self.h, n, req.data_ptr(), res.data_ptr(), stream))

    def unpermute(self, n, routed_res, src_index, res, stream):
        self._check(self.L.gcra_route_unpermute(self.h, n, routed_res.data_ptr(), src_index.data_ptr(),
                                                res.data_ptr(), stream))


class ShardedLimiter:
    def __init__(self, limiter, dist, device, ops=None, max_rows=None):
        self.dist, self.dev = dist, device
        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        self.ops = ops if ops is not None else CudaOps(limiter)
        self.max_rows = max_rows or limiter.store.max_batch
        u8 = dict(dtype=torch.uint8, device=device)
        self.routed = torch.empty(self.max_rows * REQ_B, **u8)
        self.recv_req = torch.empty(self.max_rows * REQ_B, **u8)
        self.recv_res = torch.empty(self.max_rows * RES_B, **u8)
        self.back_res = torch.empty(self.max_rows * RES_B, **u8)
        self.src_index = torch.empty(self.max_rows, dtype=torch.int32, device=device)
        self.counts = torch.zeros(16, dtype=torch.int32, device=device)
        self.last_recv_rows = 0

    def _stream(self):
        return torch.cuda.current_stream(self.dev).cuda_stream if self.dev.type == "cuda" else None

    def step(self, d_req, d_res):
        """d_req: uint8 tensor of n*48 bytes (gcra_request rows); d_res: uint8 tensor of n*32 bytes."""
        n = d_req.numel() // REQ_B
        W, dist, st = self.world, self.dist, self._stream()
        self.ops.partition(n, d_req, W, self.routed, self.src_index, self.counts, st)
        send = self.counts[:W].to(torch.int64)
        recv = torch.empty_like(send)
        dist.all_to_all_single(recv, send)                       # how many rows every peer sends me
        send_l, recv_l = send.tolist(), recv.tolist()            # one host sync per tick
        n_recv = sum(recv_l)
        if n_recv > self.max_rows:
            raise RuntimeError("shard received %d rows > max_batch %d" % (n_recv, self.max_rows))
        self.last_recv_rows = n_recv
        dist.all_to_all_single(self.recv_req[:n_recv * REQ_B], self.routed[:n * REQ_B],
                               output_split_sizes=[c * REQ_B for c in recv_l],
                               input_split_sizes=[c * REQ_B for c in send_l])
        self.ops.decide(n_recv, self.recv_req, self.recv_res, st)
        dist.all_to_all_single(self.back_res[:n * RES_B], self.recv_res[:n_recv * RES_B],
                               output_split_sizes=[c * RES_B for c in send_l],
                               input_split_sizes=[c * RES_B for c in recv_l])
        self.ops.unpermute(n, self.back_res, self.src_index, d_res, st)
        return n_recv
