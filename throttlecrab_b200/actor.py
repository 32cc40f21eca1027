"""Batch-draining actor: the server-side seam of the hot path (SURVEY §8f "next" row #1).

The reference serialises every request through one actor task that handles ONE message per `recv`
(throttlecrab-server/src/actor.rs:217-236, handle at :68-82).  This mirror keeps the contract --
requests are applied strictly in arrival order by a single owner of the limiter, callers get a
`ThrottleResponse` back (types.rs:19-31, seconds floored as in types.rs:87-97) -- but drains the whole
queue per wake-up and hands it to the engine as ONE batch (`gcra_rate_limit_batch`), which is what lets
a GPU engine sit behind per-request callers.  Host-side glue only: no decision is made here.
"""
import queue
import threading
import time
from concurrent.futures import Future
from dataclasses import dataclass

import numpy as np

from . import (REQ_DTYPE, RES_DTYPE, NS, RateLimiter, hash_key, Internal, InvalidRateLimit,
               NegativeQuantity, NEGATIVE_QUANTITY, INVALID_RATE_LIMIT, OK)


@dataclass
class ThrottleRequest:          # throttlecrab-server/src/types.rs:31-45
    key: str
    max_burst: int
    count_per_period: int
    period: int
    quantity: int
    timestamp: int              # ns since the epoch (SystemTime)


@dataclass
class ThrottleResponse:         # types.rs:19-31
    allowed: bool
    limit: int
    remaining: int
    reset_after: int            # whole seconds (types.rs:93)
    retry_after: int            # whole seconds (types.rs:94)


class RateLimiterHandle:
    """`RateLimiterHandle::throttle` (actor.rs:68-82): thread-safe, returns a Future."""

    def __init__(self, store, buffer_size=100_000, max_batch=None):
        self._lim = RateLimiter(store)
        self._q = queue.Queue(maxsize=buffer_size)          # bounded channel, --buffer-size
        self._max = max_batch or store.max_batch
        self._stop = False
        self.batches = 0
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()

    def throttle(self, request):
        fut = Future()
        if self._stop:
            fut.set_exception(Internal("Rate limiter actor has shut down"))     # actor.rs:77
            return fut
        self._q.put((request, fut))
        return fut

    def shutdown(self):
        self._stop = True
        self._q.put(None)
        self._t.join()

    def _run(self):                                         # run_actor, actor.rs:217-236
        while True:
            item = self._q.get()
            if item is None:
                return
            batch = [item]
            while len(batch) < self._max:                   # recv_many: take what is already queued
                try:
                    nxt = self._q.get_nowait()
                except queue.Empty:
                    break
                if nxt is None:
                    self._stop = True
                    break
                batch.append(nxt)
            self._handle(batch)
            if self._stop:
                return

    def _handle(self, batch):                               # handle_throttle, actor.rs:238-255
        req = np.empty(len(batch), REQ_DTYPE)
        for i, (r, _) in enumerate(batch):
            req[i] = (hash_key(r.key), r.max_burst, r.count_per_period, r.period, r.quantity, r.timestamp)
        try:
            res = self._lim.rate_limit_batch(req)
        except Exception as ex:                              # engine error: fail every caller of the batch
            for _, fut in batch:
                fut.set_exception(ex)
            return
        self.batches += 1
        for (r, fut), o in zip(batch, res):
            st = int(o["status"])
            if st == OK:
                fut.set_result(ThrottleResponse(bool(o["allowed"]), r.max_burst, int(o["remaining"]),
                                                int(o["reset_after_ns"]) // NS, int(o["retry_after_ns"]) // NS))
            elif st == NEGATIVE_QUANTITY:
                fut.set_exception(NegativeQuantity(r.quantity))
            elif st == INVALID_RATE_LIMIT:
                fut.set_exception(InvalidRateLimit())
            else:
                fut.set_exception(Internal("rate limiter internal error"))
