#!/usr/bin/env python
"""bench.py -- GCRA decisions/sec on B200 (BASELINE.json metric), one JSON line.

A "step" is one tick: one pass of the hot path (ingest -> order -> decide) over one batch of
2^20 synthetic requests.

  N=1   BASELINE.json configs[1]: 10 M resident keys, Zipf-1.0 request stream (tests/traces.py
        config2), one `now` per tick advancing 1 ms, after a warm pass that inserts every key.
  N>1   configs[4] shape: the key space (10 M keys per GPU) is hash-sharded across the N engines;
        every rank ingests its own 2^20-request slice of the global tick, routes each request to
        the owning shard (stable partition kernel + NCCL all-to-all), decides locally, and routes
        the results back.  Weak scaling.

value      kernel-only: requests already resident in HBM, K steps back to back on the stream
e2e        the same K ticks through the C-ABI pinned host ring (gcra_ring_*): H2D of every tick's
           requests, kernels, D2H of every tick's results inside the timed region
roofline   K1 (ingest+order+decide launches of one tick), algorithmic bytes / CUDA-event time
cpu_baseline  the CPU oracle (C++ restatement of the reference; the reference is Rust and cannot be
           built here), single thread = the reference's design point, on a bounded sample

`--impl reference` times that CPU restatement on all host cores (hash-sharded stores) instead.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import traces  # noqa: E402

TICK = 1 << 20
KEYS_PER_GPU = 10_000_000
METRIC = "gcra_decisions_per_sec"
UNIT = "decisions/s"


def measured_peak_gbs():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu=0):
        self.gpu, self.rows, self.proc = gpu, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for nm, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(mx)) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def build_requests(tc, key_hash_of, trace):
    req = np.empty(len(trace), tc.REQ_DTYPE)
    req["key_hash"] = key_hash_of[trace["key"].astype(np.int64)]
    for f in ("max_burst", "count_per_period", "period", "quantity", "now_ns"):
        req[f] = trace[f]
    return req


def cpu_baseline_run(n_keys, n_ticks, threads, start_tick=0):
    """Oracle (C++ restatement of the reference AdaptiveStore path) on host cores: warm pass over
    every key, then n_ticks Zipf ticks timed.  threads>1 = independent hash-sharded stores."""
    import oracle
    warm = traces.warm_pass(n_keys)
    tr = traces.config2(n_keys=n_keys, n_ticks=n_ticks, tick_size=TICK, start_tick=start_tick)
    stores = [oracle.OracleStore(oracle.ADAPTIVE, capacity=max(n_keys // threads, 1000), created_ns=traces.T0)
              for _ in range(threads)]
    oracle.replay_sharded(stores, warm)                      # untimed: table population
    _, sec = oracle.replay_sharded(stores, tr)               # timed: decision loops only
    return len(tr) / sec, sec


def run_reference(args, rank, world):
    """`--impl reference`: the CPU restatement on all host cores, same metric/config."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n_keys = KEYS_PER_GPU * max(args.gpus, 1)
    sample_ticks = 4
    # bounded sample: the table holds n_keys/4 keys so that population stays within ~1 minute of CPU
    sample_keys = min(n_keys, 2_500_000)
    vals = []
    t_all = time.time()
    for s in range(args.warmup + args.steps):
        v, sec = cpu_baseline_run(sample_keys, sample_ticks, cores, start_tick=s * sample_ticks)
        if s >= args.warmup:
            vals.append(v)
        if time.time() - t_all > 240:
            break
    value = float(np.median(vals)) if vals else 0.0
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * TICK / value if value else None,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64",
        "data": "synthetic",
        "config": {"workload": "zipf1.0 ticks of 2^20 requests (tests/traces.py config2)",
                   "keys": sample_keys, "tick": TICK, "note": "bounded CPU sample of the 10M-key workload"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "%d hash-sharded AdaptiveStore restatements (C++ oracle; the Rust reference "
                                   "cannot be built here), %d keys warm pass + %d ticks per step"
                                   % (cores, sample_keys, sample_ticks)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=16)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--keys", type=int, default=KEYS_PER_GPU)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import throttlecrab_b200 as tc
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        # the all-to-alls are pairwise sends over NVLink/NVSwitch: give each peer pair more channels
        os.environ.setdefault("NCCL_MIN_P2P_NCHANNELS", "8")
        dist.init_process_group("nccl", device_id=dev)

    W, K = max(args.warmup, 3), args.steps
    n_local_keys = args.keys
    n_keys = n_local_keys * world
    peak, peak_kind = measured_peak_gbs()

    # ---------------------------------------------------------------- synthetic workload
    t0 = time.time()
    key_hash_of = tc.hash_key_ids(np.arange(n_keys, dtype=np.uint64))
    store = tc.ManualStore(capacity=n_local_keys, device=local_rank, created_ns=traces.T0,
                           max_batch=TICK if world == 1 else 2 * TICK)
    lim = tc.RateLimiter(store)
    if world == 1:
        # warm pass: every key inserted once (BASELINE configs[1]: keys resident)
        warm = build_requests(tc, key_hash_of, traces.warm_pass(n_keys))
        for a in range(0, n_keys, TICK):
            lim.rate_limit_batch(warm[a:a + TICK])
        del warm
        tr = traces.config2(n_keys=n_keys, n_ticks=W + K, tick_size=TICK)
        ticks = build_requests(tc, key_hash_of, tr)
        del tr
    else:
        from throttlecrab_b200.sharded import NativeShardedLimiter, ShardedLimiter
        # default: the native pipeline (one C call per tick, NCCL inside the library); GCRA_SHARD_PY=1 selects
        # the torch.distributed implementation of the same stages
        native_shard = os.environ.get("GCRA_SHARD_PY", "0") != "1"
        sh = NativeShardedLimiter(lim, dist, dev) if native_shard else ShardedLimiter(lim, dist, dev)
        # warm pass through the sharded path: rank r submits keys [r*10M, (r+1)*10M), owners insert them
        stream0 = torch.cuda.Stream(dev)
        torch.cuda.set_stream(stream0)
        wres = torch.empty(TICK * 32, dtype=torch.uint8, device=dev)
        for a in range(0, n_local_keys, TICK):
            ids = np.arange(rank * n_local_keys + a, rank * n_local_keys + min(a + TICK, n_local_keys), dtype=np.uint64)
            w = np.zeros(TICK, traces.REQ_DTYPE)          # padded with copies of the last key (harmless)
            w["key"][:len(ids)] = ids
            w["key"][len(ids):] = ids[-1]
            traces.fill_policy(w, (w["key"] % np.uint64(8)).astype(np.int64))
            w["quantity"] = 1
            w["now_ns"] = traces.T0
            wreq = torch.from_numpy(build_requests(tc, key_hash_of, w).view(np.uint8)).to(dev)
            sh.step(wreq, wres)
        torch.cuda.synchronize()
        # every rank generates ITS slice of each global Zipf tick (same generator as N=1)
        tr = traces.config2_rank_slice(n_keys, TICK, 0, W + K, rank, world)
        ticks = build_requests(tc, key_hash_of, tr)
        del tr
    gen_s = time.time() - t0

    d_req = torch.from_numpy(ticks.view(np.uint8).reshape(W + K, TICK * 48)).to(dev)
    d_res = torch.empty((W + K, TICK * 32), dtype=torch.uint8, device=dev)
    # an explicit (non-default) stream: handle 0 would mean "the engine's own stream" to the C ABI
    stream = torch.cuda.Stream(dev)
    torch.cuda.set_stream(stream)

    def step(i):
        if world == 1:
            # pipelined submission: ingest+order of tick i+1 overlap the decide kernels of tick i
            lim.submit_device(TICK, d_req[i].data_ptr(), d_res[i].data_ptr(), stream.cuda_stream)
        else:
            sh.submit(d_req[i], d_res[i])       # pipelined: routing of tick i+1 overlaps deciding tick i

    # ---------------------------------------------------------------- kernel-only (value)
    for i in range(W):
        step(i)
    if world > 1:
        sh.finish()
    else:
        lim.join(stream.cuda_stream)
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = store.launch_count()
    stats0 = store.stats()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k1_ms = []
    torch.cuda.synchronize()
    step_ev = [torch.cuda.Event(enable_timing=True) for _ in range(K + 1)]
    ev0.record(stream)
    step_ev[0].record(stream)
    for i in range(W, W + K):
        step(i)
        step_ev[i - W + 1].record(stream)
    if world > 1:
        sh.finish()
    else:
        lim.join(stream.cuda_stream)
    ev1.record(stream)
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    total_ms = ev0.elapsed_time(ev1)
    launches = store.launch_count() - launches0
    step_ms = [step_ev[i].elapsed_time(step_ev[i + 1]) for i in range(K)]
    if dist:
        tmax = torch.tensor([total_ms], device=dev)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        total_ms = float(tmax.item())
    value = world * K * TICK / (total_ms * 1e-3)

    # per-kernel times of single ticks (library-side CUDA events on the launching stream)
    res_np = d_res[W:W + K].cpu().numpy().view(tc.RES_DTYPE).reshape(K, TICK)
    n_allowed = int(res_np["allowed"].sum())
    n_ok = int((res_np["status"] == 0).sum())
    phases = None
    if world == 1:
        # phase split of ONE tick run serially on the stream (library-side CUDA events); the timed
        # region above pipelines consecutive ticks, so its per-tick time is below this total
        extra = torch.empty(TICK * 32, dtype=torch.uint8, device=dev)
        if os.environ.get("GCRA_DBG"):          # timing experiments only: this serial tick's results are wrong
            store._L.gcra_debug_set(store._h, int(os.environ["GCRA_DBG"]))
        lim.rate_limit_batch_device(TICK, d_req[W].data_ptr(), extra.data_ptr(), stream.cuda_stream)
        torch.cuda.synchronize()
        phases = store.last_kernel_ms()
        try:
            phase_detail = store.last_kernel_ms_detail()
        except Exception:
            phase_detail = None
    # (the clock sampler keeps running through the e2e region: the kernel-only region lasts only a few ms)

    # ---------------------------------------------------------------- e2e through the pinned ring
    e2e = None
    if world == 1 and not args.no_e2e:
        store2 = tc.ManualStore(capacity=n_local_keys, device=local_rank, created_ns=traces.T0, max_batch=TICK)
        lim2 = tc.RateLimiter(store2)
        warm = build_requests(tc, key_hash_of, traces.warm_pass(n_keys))
        for a in range(0, n_keys, TICK):
            lim2.rate_limit_batch(warm[a:a + TICK])
        del warm
        KE = min(K, 32)          # every step owns a pinned 80-MiB slot: cap the pinned memory at ~3 GB
        ring = tc.Ring(lim2, slots=W + KE, slot_capacity=TICK)
        for i in range(W + KE):
            ring.req[i][:] = ticks[i * TICK:(i + 1) * TICK]      # requests sit in pinned host memory
        for i in range(W):
            ring.submit(i, TICK)
        for i in range(W):
            ring.wait(i)
        store2.sync()
        t_a = time.perf_counter()
        for i in range(W, W + KE):
            ring.submit(i, TICK)
        for i in range(W, W + KE):
            ring.wait(i)
        t_b = time.perf_counter()
        e2e_val = KE * TICK / (t_b - t_a)
        got = np.concatenate([ring.res[i] for i in range(W, W + KE)])
        same = got.tobytes() == res_np[:KE].reshape(-1).tobytes()
        e2e = {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": TICK * 48, "d2h_bytes_per_step": TICK * 32,
               "api": "gcra_ring_submit/gcra_ring_wait, pinned host ring, 48-byte requests", "steps": KE,
               "matches_kernel_only_results": bool(same)}
        del ring
        # extra: compact 16-byte requests (policy table + per-call now), same ticks, same results
        ring16 = None
        try:
            store3 = tc.ManualStore(capacity=n_local_keys, device=local_rank, created_ns=traces.T0, max_batch=TICK)
            lim3 = tc.RateLimiter(store3)
            pol = np.zeros(8, tc.POLICY_DTYPE)
            pol["max_burst"], pol["count_per_period"], pol["period"] = traces.POLICIES.T
            lim3.set_policies(pol)
            warm = build_requests(tc, key_hash_of, traces.warm_pass(n_keys))
            for a in range(0, n_keys, TICK):
                lim3.rate_limit_batch(warm[a:a + TICK])
            del warm
            ring16 = tc.Ring(lim3, slots=W + KE, slot_capacity=TICK, compact=True)
            for i in range(W + KE):
                sl = ticks[i * TICK:(i + 1) * TICK]
                r16 = ring16.req[i]
                r16["key_hash"] = sl["key_hash"]
                r16["quantity"] = sl["quantity"]
                pidx = np.zeros(TICK, np.uint32)
                for j, p in enumerate(traces.POLICIES):
                    m = (sl["max_burst"] == p[0]) & (sl["count_per_period"] == p[1]) & (sl["period"] == p[2])
                    pidx[m] = j
                r16["policy"] = pidx
            nows = [int(ticks["now_ns"][i * TICK]) for i in range(W + KE)]
            for i in range(W):
                ring16.submit(i, TICK, nows[i])
            for i in range(W):
                ring16.wait(i)
            store3.sync()
            t_a = time.perf_counter()
            for i in range(W, W + KE):
                ring16.submit(i, TICK, nows[i])
            for i in range(W, W + KE):
                ring16.wait(i)
            t_b = time.perf_counter()
            got16 = np.concatenate([ring16.res[i] for i in range(W, W + KE)])
            e2e["compact_requests"] = {"value": KE * TICK / (t_b - t_a), "unit": UNIT, "h2d_bytes_per_step": TICK * 16,
                                       "d2h_bytes_per_step": TICK * 32,
                                       "matches_kernel_only_results": bool(got16.tobytes() == res_np[:KE].reshape(-1).tobytes())}
            del ring16
            store3.close()
        except Exception as ex:      # the extra must never cost the main line
            e2e["compact_requests"] = {"error": repr(ex)}
        store2.close()

    if world > 1 and not args.no_e2e:
        # e2e at N GPUs: every rank copies its tick slice from pinned host memory, routes, decides,
        # routes back and copies the results to pinned host memory (time continues after the timed ticks)
        tr2 = traces.config2_rank_slice(n_keys, TICK, W + K, K, rank, world)
        h_req = torch.from_numpy(build_requests(tc, key_hash_of, tr2).view(np.uint8).reshape(K, TICK * 48)).pin_memory()
        h_res = torch.empty((K, TICK * 32), dtype=torch.uint8).pin_memory()
        torch.cuda.synchronize()
        dist.barrier()
        dqs = [torch.empty(TICK * 48, dtype=torch.uint8, device=dev) for _ in range(K)]
        drs = [torch.empty(TICK * 32, dtype=torch.uint8, device=dev) for _ in range(K)]
        s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        sh.finish()
        if not native_shard:
            sh.pop_returned()
        torch.cuda.synchronize()
        dist.barrier()
        t_a = time.perf_counter()
        copied = 0
        for i in range(K):
            with torch.cuda.stream(s_in):                     # H2D of tick i overlaps earlier ticks
                dqs[i].copy_(h_req[i], non_blocking=True)
            if native_shard:
                sh.submit(dqs[i], drs[i], ready_stream=s_in.cuda_stream)
                if i >= 1:                                    # tick i-1's way back was issued inside submit(i)
                    sh.wait_tick(1, s_out)
                    with torch.cuda.stream(s_out):
                        h_res[copied].copy_(drs[copied], non_blocking=True)
                    copied += 1
            else:
                stream.wait_stream(s_in)
                sh.submit(dqs[i], drs[i])
                for ev in sh.pop_returned():                  # D2H of every tick whose results are on the way
                    s_out.wait_event(ev)
                    with torch.cuda.stream(s_out):
                        h_res[copied].copy_(drs[copied], non_blocking=True)
                    copied += 1
        sh.finish()
        if native_shard:
            s_out.wait_stream(stream)
            with torch.cuda.stream(s_out):
                while copied < K:
                    h_res[copied].copy_(drs[copied], non_blocking=True)
                    copied += 1
        else:
            for ev in sh.pop_returned():
                s_out.wait_event(ev)
                with torch.cuda.stream(s_out):
                    h_res[copied].copy_(drs[copied], non_blocking=True)
                copied += 1
        stream.wait_stream(s_out)
        torch.cuda.synchronize()
        dist.barrier()
        t_b = time.perf_counter()
        tt = torch.tensor([t_b - t_a], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e = {"value": world * K * TICK / float(tt.item()), "unit": UNIT, "h2d_bytes_per_step": TICK * 48,
               "d2h_bytes_per_step": TICK * 32,
               "api": "per rank: pinned host -> H2D -> ShardedLimiter.submit (partition, all-to-all, decide, all-to-all, "
                      "unpermute) -> D2H to pinned host, copies on their own streams"}

    clocks = sampler.stop() if rank == 0 else None

    # ---------------------------------------------------------------- CPU baseline (rank 0, N=1)
    cpu = None
    if world == 1 and rank == 0 and not args.no_cpu:
        v, sec = cpu_baseline_run(n_keys, 4, 1)
        cpu = {"value": v, "unit": UNIT, "cores": 1, "kind": "port",
               "sample": "C++ restatement of throttlecrab AdaptiveStore + RateLimiter (Rust toolchain "
                         "unavailable), string keys, 1 thread: %d-key warm pass (untimed) + 4 Zipf ticks "
                         "of 2^20 requests timed (%.1f s)" % (n_keys, sec)}

    if rank != 0:
        if dist:
            dist.destroy_process_group()
        return
    alg_bytes = 112.0 * n_allowed + 96.0 * (K * TICK - n_allowed)
    roof = None
    if world == 1:
        t_k1 = total_ms * 1e-3
        ach = alg_bytes / t_k1 / 1e9
        st_k = store.stats()
        roof = {"bound": "hbm", "kernel": "K1 = probe + note | decide (batch order) + resolve | sorted residue; the three stages of consecutive ticks run on three streams",
                "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "peak_kind": peak_kind,
                # measured DRAM bytes of one tick's K1 kernels (dram__bytes_read.sum + dram__bytes_write.sum from the
                # committed `ncu --set full` capture profiles/r01d_ncu_full_k1_raw.csv: ingest 108.7 MB + decide
                # 105.2 + hot-run kernels 14.9 + 21.4 + 3 x sort_scatter 9.8; sort hist/rowscan were not captured)
                "traffic": 279.8e6, "traffic_note": "bytes per tick, from profiles/r01d_ncu_full_k1_raw.csv",
                "algorithmic_bytes_per_tick": alg_bytes / K,
                "algorithmic_bytes_per_decision": {"allowed": 112, "denied": 96},
                "serial_tick_phase_ms": {"total": phases[0], "probe+note": phases[1], "decide+resolve": phases[2],
                                         "sorted_residue": phases[3]},
                "serial_tick_detail_ms": phase_detail,
                # requests that went through the sorted tail, over the batches of the timed region whose count
                # had reached the host when the last one was submitted
                "residue_fraction": ((st_k["residue_rows"] - stats0["residue_rows"])
                                     / max(st_k["residue_batches"] - stats0["residue_batches"], 1)) / TICK,
                "pipeline_drains": st_k["drains"] - stats0["drains"], "index_batches": st_k["index_batches"]}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int64", "data": "synthetic",
        "config": {"workload": ("10M keys, Zipf-1.0 request stream, ticks of 2^20 requests (BASELINE configs[1])"
                                if world == 1 else
                                "%dM keys hash-sharded over %d GPUs, Zipf-1.0 stream, 2^20 requests per GPU per tick, "
                                "stable partition + NCCL all-to-all routing (BASELINE configs[4] shape)"
                                % (n_keys // 1_000_000, world)),
                   "sharded_pipeline": (None if world == 1 else ("native (gcra_shard_submit)" if native_shard else "torch.distributed")),
                   "keys": n_keys, "tick": TICK, "request_bytes": 48, "result_bytes": 32,
                   "l2": "no flush: table %.2f GB and a distinct 80 MB tick per step exceed the 126 MB L2"
                         % (store.stats()["table_slots"] * 32 / 1e9),
                   "allowed_fraction": n_allowed / max(n_ok, 1), "gen_seconds": round(gen_s, 1)},
        "host_enqueue_ms_per_step": {"min": min(step_ms), "median": float(np.median(step_ms)), "max": max(step_ms)},
        "roofline": roof, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
    }
    print(json.dumps(line), flush=True)
    if dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
