#!/usr/bin/env python
"""bench.py — headline benchmark of the hot path (BASELINE.json): queries/s on a 10M-key ASCII gSize=3 index,
batch 4096, top-100 (+ index-build keys/s), through the engine's C ABI.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--keys 10000000] [--batch 4096] [--limit 100]
    python bench.py --impl reference ...      # the reference's own CPU implementation (oracle/_ref) on host cores

A step = one batch of `--batch` queries answered against the whole index.
  value      queries/s with the batch already resident in HBM (device part only, CUDA events on the engine's stream)
  e2e        queries/s through ssb_search_batch() with HOST char** queries: packing, H2D, kernels, D2H, marshalling
  roofline   k_fast_search: algorithmic bytes (4 B per posting of every distinct query gram + 8 B per gram lookup +
             8 B per result) / its CUDA-event duration, against the measured HBM copy bandwidth
  N > 1      key-range shards (one process per GPU, torchrun): every rank answers the full batch on its shard, one
             NCCL all-gather of fixed-size records, merge kernel on every rank; total index size fixed => "strong"
Every step uses a different query batch and the posting table (>= 360 MB at 10M keys; ~750 MB touched per batch) is
larger than L2, so no explicit L2 flush is needed.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "queries/sec (10M-key index, batch 4096, top-100) + index-build keys/sec"


def read_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """samples nvidia-smi clocks / throttle reasons while the timed region runs"""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        self.gpu, self.proc, self.lines = gpu, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.t.join(timeout=2)
        sm, mx, reasons = [], None, set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def reference_leg(args, nkeys_sample, nq_sample, threads, steps=1, warmup=0):
    """the reference's own CPU implementation (oracle/_ref/libref.so: indexN + score) on the host cores.
    Bounded sample: a prefix of the same vocabulary and the first queries of the same batches."""
    from oracle import loader
    from stringsearchlib_b200 import synth
    if not loader.have_ref():
        raise RuntimeError("oracle/_ref missing")
    R = loader.RefLib()
    keys = synth.ascii_keys(nkeys_sample)
    t0 = time.perf_counter()
    h = R.indexN(keys, 1, None)
    t_build = time.perf_counter() - t0
    L = R.L

    def run_batch(q, threads=threads):
        ptrs = q.pointers()
        nq = len(q)
        addr = ctypes.cast(ptrs, ctypes.POINTER(ctypes.c_char_p))
        counts = [0] * threads

        def work(t):
            res, sc = ctypes.POINTER(ctypes.c_char_p)(), ctypes.POINTER(ctypes.c_float)()
            c = 0
            for i in range(t, nq, threads):
                n = L.score(h, addr[i], ctypes.byref(res), ctypes.byref(sc), ctypes.c_float(args.threshold), args.limit)
                c += n
                L.release(h, res, sc)
            counts[t] = c

        th = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
        t0 = time.perf_counter()
        for t in th:
            t.start()
        for t in th:
            t.join()
        return time.perf_counter() - t0, sum(counts)

    times = []
    for s in range(warmup + steps):
        q = synth.ascii_queries(nq_sample, nkeys_sample, first_query=s * args.batch)
        dt, _ = run_batch(q)
        if s >= warmup:
            times.append(dt)
    # (i) one caller thread next to (ii) all host threads — the first queries of the same sample
    n1 = max(8, nq_sample // max(threads, 1))
    dt1, _ = run_batch(synth.ascii_queries(n1, nkeys_sample), threads=1)
    R.dispose(h)
    qps = nq_sample * len(times) / sum(times)
    return dict(qps=qps, qps_one_thread=n1 / dt1, build_keys_per_s=nkeys_sample / t_build, build_s=t_build, ms_per_step=1e3 * sum(times) / len(times))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--keys", type=int, default=10_000_000)
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--limit", type=int, default=100)
    ap.add_argument("--threshold", type=float, default=0.0)
    ap.add_argument("--cpu-keys", type=int, default=1_000_000, help="vocabulary prefix indexed by the CPU reference leg")
    ap.add_argument("--cpu-queries", type=int, default=256, help="queries per step of the CPU reference leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-micro", action="store_true", help="(kept for older scripts; no effect)")
    ap.add_argument("--profile-exchange", action="store_true", help="N>1: print CUDA-event times of the exchange phases (stderr)")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "p2p-barrier", "nccl"],
                    help="N>1: p2p = export/merge kernels storing into NVLink peer memory with a device-side epoch hand-over, no "
                         "collective at all (default); p2p-barrier = the same stores with two 4-byte all-reduces as barriers and "
                         "the results replicated on every rank; nccl = NCCL all-to-all + all-gather")
    ap.add_argument("--layout", choices=("shards", "replicas"), default="shards",
                    help="N > 1: shards = the index split in key ranges, every rank answers the whole batch on its shard, per-query top-k "
                         "merge across ranks (BASELINE.json's configuration, the default); replicas = every rank holds the whole index "
                         "and answers 1/N of the batch, no exchange (what a deployment would do with an index that fits one GPU)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cores = os.cpu_count() or 1
    workload = f"C4: {args.keys}-key synthetic ASCII vocabulary (6-16 letters), rowSize=1, gSize=3, batch {args.batch}, " \
               f"threshold {args.threshold:g}, top-{args.limit}"
    config = {"workload": workload, "keys": args.keys, "batch": args.batch, "limit": args.limit, "threshold": args.threshold,
              "gSize": 3, "rowSize": 1,
              "sharding": f"key-range x{world}; per-query top-k merge, exchange={args.exchange}; e2e: every rank returns the merged "
                          f"results of its 1/{world} slice of the batch to its host" if world > 1 else "single shard",
              "l2": "inputs larger than L2: new query batch every step, ~184 KB of postings per query"}
    replicas = world > 1 and args.layout == "replicas"
    if replicas:
        if args.batch % world:
            raise SystemExit("--layout replicas needs a batch divisible by the number of GPUs")
        config["sharding"] = f"replicated index x{world}: every rank answers its 1/{world} slice of the batch on the whole index, no exchange"

    if args.impl == "reference":
        if rank != 0:
            return
        threads = cores
        r = reference_leg(args, args.cpu_keys, args.cpu_queries, threads, steps=max(args.steps, 1), warmup=min(args.warmup, 1))
        sample = f"unmodified reference (oracle/_ref/libref.so): indexN over the first {args.cpu_keys} keys of the vocabulary, " \
                 f"{args.cpu_queries} queries of each batch per step through score()/release(), {threads} caller threads"
        line = {"impl": "reference", "metric": METRIC, "value": r["qps"], "unit": "queries/s", "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "u32", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": r["qps"], "unit": "queries/s", "cores": threads, "kind": "reference", "sample": sample},
                "index_build_keys_per_s": r["build_keys_per_s"], "one_thread_value": r["qps_one_thread"],
                "e2e": {"value": r["qps"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    import torch
    import stringsearchlib_b200 as ssb
    from stringsearchlib_b200 import synth
    from stringsearchlib_b200 import api
    from stringsearchlib_b200.api import RECORD_DTYPE, merge_records

    torch.cuda.set_device(local_rank)
    ssb.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- index build through the char** entry point (host pointers, like an FFI caller)
    lo = 0 if replicas else args.keys * rank // world
    hi = args.keys if replicas else args.keys * (rank + 1) // world
    keys = synth.ascii_keys(hi - lo, first=lo)
    ptrs = keys.pointers()
    # first build = cold (CUDA context, memory pool growth, pinned staging allocation); it is reported separately and the
    # index-build rate is the steady state of the following builds, like any other warmed-up step of this benchmark
    barrier()
    t0 = time.perf_counter()
    ix = ssb.Index.from_pointers(ptrs, len(keys), 1, None, 3, False, row_base=lo)
    t_build_cold = time.perf_counter() - t0
    ix.close()
    t_builds = []
    for _ in range(3):
        barrier()
        t0 = time.perf_counter()
        ix = ssb.Index.from_pointers(ptrs, len(keys), 1, None, 3, False, row_base=lo)   # synchronous: returns a searchable index
        t_builds.append(time.perf_counter() - t0)
        if _ < 2:
            ix.close()
    t_build = float(np.median(t_builds))
    bms = ix.build_ms()
    info = ix.info()
    if world > 1:
        tb = torch.tensor([t_build], device="cuda")
        dist.all_reduce(tb, op=dist.ReduceOp.MAX)
        t_build = float(tb.item())
    # getSize / getLibSize of the whole index: distinct terms and grams over all shards (SURVEY.md §8e "Global counters")
    global_sizes = None
    if world > 1 and not replicas:
        barrier()
        t0 = time.perf_counter()
        gs = api.global_sizes(ix, dist)
        barrier()
        global_sizes = {"getSize": gs[0], "getLibSize": gs[1], "seconds": time.perf_counter() - t0,
                        "how": "16-byte digests per term / gram, all-gather, distinct count per residue class on every rank, all-reduce"}

    nsteps = args.warmup + args.steps
    nb = args.batch // world if replicas else args.batch          # queries this rank answers per step
    batches = [synth.ascii_queries(nb, args.keys, first_query=s * args.batch + (rank * nb if replicas else 0)) for s in range(nsteps)]
    qptrs = [b.pointers() for b in batches]
    rec_bytes = RECORD_DTYPE.itemsize
    sliced = world > 1 and args.batch % world == 0 and not replicas
    use_p2p = sliced and args.exchange in ("p2p", "p2p-barrier")
    use_signal = sliced and args.exchange == "p2p"
    if world > 1 and not replicas:
        part_q = args.batch // world if sliced else 0
        host_rec = torch.empty((args.batch, args.limit, rec_bytes), dtype=torch.uint8, pin_memory=True)
        host_cnt = torch.empty(args.batch, dtype=torch.int32, pin_memory=True)
        flag = torch.zeros(1, device="cuda")
    if use_p2p:
        # result / receive buffers are cudaMalloc'ed by the engine and opened in every peer process through CUDA IPC
        recv_ptr = api.device_alloc(world * part_q * args.limit * rec_bytes)
        recv2_ptr = api.device_alloc(world * part_q * args.limit * rec_bytes)      # exchanges alternate between two receive buffers
        flags_ptr = api.device_alloc(64)
        api.device_memset(flags_ptr, 0, 64)
        merged_ptr = api.device_alloc(args.batch * args.limit * rec_bytes)
        counts_ptr = api.device_alloc(args.batch * 4)
        handles = [None] * world
        dist.all_gather_object(handles, tuple(api.ipc_get_handle(p) for p in (recv_ptr, merged_ptr, counts_ptr, recv2_ptr, flags_ptr)))
        try:
            peer_recv = [recv_ptr if r == rank else api.ipc_open_handle(handles[r][0]) for r in range(world)]
            peer_merged = [merged_ptr if r == rank else api.ipc_open_handle(handles[r][1]) for r in range(world)]
            peer_counts = [counts_ptr if r == rank else api.ipc_open_handle(handles[r][2]) for r in range(world)]
            peer_recv2 = [recv2_ptr if r == rank else api.ipc_open_handle(handles[r][3]) for r in range(world)]
            peer_flags = [flags_ptr if r == rank else api.ipc_open_handle(handles[r][4]) for r in range(world)]
            ok = torch.ones(1, device="cuda")
        except RuntimeError:
            ok = torch.zeros(1, device="cuda")
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        use_p2p = bool(ok.item() > 0)
    if world > 1 and not use_p2p and not replicas:
        mine = torch.empty((args.batch, args.limit, rec_bytes), dtype=torch.uint8, device="cuda")
        merged = torch.empty((args.batch, args.limit, rec_bytes), dtype=torch.uint8, device="cuda")
        mcount = torch.empty(args.batch, dtype=torch.int32, device="cuda")
        if sliced:
            recv = torch.empty((world, part_q, args.limit, rec_bytes), dtype=torch.uint8, device="cuda")
            part = torch.empty((part_q, args.limit, rec_bytes), dtype=torch.uint8, device="cuda")
            pcount = torch.empty(part_q, dtype=torch.int32, device="cuda")
        else:
            gathered = torch.empty((world, args.batch, args.limit, rec_bytes), dtype=torch.uint8, device="cuda")

    xprof = []
    epoch = [0]
    if world > 1:
        xev0, xev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def exchange_timed():
        """exchange() bracketed by CUDA events on the stream it is enqueued on; returns device milliseconds"""
        xev0.record()
        exchange()
        return xev0.elapsed_time(xev1)

    def exchange():
        """per-query top-k merge across the key-range shards; every rank merges a 1/N slice of the queries.
        p2p : the export kernel stores each record into the merging rank's buffer over NVLink and then signals the exchange's
              epoch into every rank's flags; the merge kernel waits for the flags of all shards and keeps its slice of final
              records on its own rank. No collective, no host round trip between the two kernels.
        p2p-barrier: the same stores, two 4-byte all-reduces as cross-rank barriers, merged slices stored into every rank.
        nccl: all-to-all of the records, merge kernel, all-gather of the merged slices."""
        if use_signal:
            epoch[0] += 1
            e = epoch[0]
            ix.export_records_p2p_signal(args.batch, args.limit, peer_recv2 if e & 1 else peer_recv, rank, peer_flags, e)
            api.merge_records_p2p_wait(recv2_ptr if e & 1 else recv_ptr, world, part_q, args.limit, [merged_ptr], [counts_ptr], rank,
                                       flags_ptr, e)
            xev1.record()
            torch.cuda.synchronize()
            return
        if use_p2p:
            # everything below is enqueued on the default stream: kernel -> 4-byte all-reduce (cross-rank barrier: all
            # exports have landed) -> kernel -> all-reduce (all merged slices have landed); one host sync at the end
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)] if args.profile_exchange else None
            if ev: ev[0].record()
            ix.export_records_p2p(args.batch, args.limit, peer_recv, rank)
            if ev: ev[1].record()
            dist.all_reduce(flag)
            if ev: ev[2].record()
            api.merge_records_p2p(recv_ptr, world, part_q, args.limit, peer_merged, peer_counts, rank)
            if ev: ev[3].record()
            dist.all_reduce(flag)
            if ev: ev[4].record()
            xev1.record()
            torch.cuda.synchronize()
            if ev and rank == 0:
                xprof.append([ev[i].elapsed_time(ev[i + 1]) for i in range(4)])
            return
        ix.export_records(args.batch, args.limit, mine.data_ptr())
        if sliced:
            dist.all_to_all_single(recv, mine)
            merge_records(recv.data_ptr(), world, part_q, args.limit, part.data_ptr(), pcount.data_ptr())
            dist.all_gather_into_tensor(merged, part)
            dist.all_gather_into_tensor(mcount, pcount)
        else:
            dist.all_gather_into_tensor(gathered, mine)
            merge_records(gathered.data_ptr(), world, args.batch, args.limit, merged.data_ptr(), mcount.data_ptr())
        xev1.record()
        torch.cuda.synchronize()

    def results_to_host():
        """every rank delivers the merged results of the query slice it merged (one process per GPU: the caller of rank r gets
        the answers to queries [r * B/N, (r+1) * B/N)); without slicing rank-local copies of the whole batch."""
        if sliced:
            q0 = rank * part_q
            if use_p2p:
                api.memcpy_d2h(host_rec.numpy()[q0:q0 + part_q], merged_ptr + q0 * args.limit * rec_bytes)
                api.memcpy_d2h(host_cnt.numpy()[q0:q0 + part_q], counts_ptr + q0 * 4)
            else:
                host_rec[q0:q0 + part_q].copy_(part, non_blocking=True)
                host_cnt[q0:q0 + part_q].copy_(pcount, non_blocking=True)
                torch.cuda.synchronize()
        elif use_p2p:
            api.memcpy_d2h(host_rec.numpy(), merged_ptr)
            api.memcpy_d2h(host_cnt.numpy(), counts_ptr)
        else:
            host_rec.copy_(merged, non_blocking=True)
            host_cnt.copy_(mcount, non_blocking=True)
            torch.cuda.synchronize()

    dev_ms, fast_ms, e2e_s, step_s, postings, results, exch_ms = [], [], [], [], [], [], []
    last_st = {}
    sampler = ClockSampler(local_rank)
    launches0 = 0
    h2d = d2h = 0
    for s in range(nsteps):
        if s == 0:
            sampler.start()          # clocks are sampled from the first warm-up step on (the same load as the timed region)
        if s == args.warmup:
            barrier()
            launches0 = ssb.kernel_launches()
        timed = s >= args.warmup
        # (1) end to end through the C ABI with host buffers
        barrier()
        t0 = time.perf_counter()
        if world == 1 or replicas:
            bp = ix.search_batch_raw(qptrs[s], nb, args.threshold, args.limit)   # H2D + kernels + D2H inside
            n_results = int(bp.contents.offsets[nb])
            ix.free_raw(bp)
        else:
            # sharded: H2D + kernels on every rank, cross-shard merge, then ONE D2H of the merged records
            ix.search_batch_device(qptrs[s], args.batch, args.threshold, args.limit)
            exchange()
            results_to_host()
            n_results = int(host_cnt[rank * part_q:(rank + 1) * part_q].sum().item()) if sliced else int(host_cnt.sum().item())
        t1 = time.perf_counter()
        # (2) device part only, the same batch resident in HBM
        barrier()
        t2 = time.perf_counter()
        ms, msk, st = ix.search_resident(nb, args.threshold, args.limit)
        if world > 1 and not replicas:
            xms = exchange_timed()
            ms += xms
            if timed:
                exch_ms.append(xms)          # device time of the step = search (engine-stream events) + exchange (default-stream events)
        torch.cuda.synchronize()
        t3 = time.perf_counter()
        if timed:
            e2e_s.append(t1 - t0)
            step_s.append(t3 - t2)
            dev_ms.append(ms)
            fast_ms.append(msk)
            postings.append(st["postings"])
            last_st = st
            results.append(n_results)
            assert st["nFast"] == nb and st["nDense"] == 0, st
            h2d = int(batches[s].offsets[-1]) + 4 * (nb + 1)
            d2h = 4 * (nb + 1) + 24 * n_results if (world == 1 or replicas) else args.batch * args.limit * rec_bytes + 4 * args.batch
    barrier()
    launches = ssb.kernel_launches() - launches0
    # the timed region of a few steps lasts only milliseconds; keep the identical load running (untimed) until nvidia-smi
    # has produced enough 50 ms samples for a meaningful clock / throttle record
    t_extra = time.perf_counter()
    extra_steps = 0
    while len(sampler.lines) < 8 and time.perf_counter() - t_extra < 2.0:
        ix.search_resident(nb, args.threshold, args.limit)
        extra_steps += 1
    clocks = sampler.stop()
    clocks["note"] = f"sampled every 50 ms from the first warm-up step; {extra_steps} extra untimed steps of the same load to reach {len(sampler.lines)} samples"

    def agg_max(x):
        if world == 1:
            return x
        t = torch.tensor(x, device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.tolist()

    step_s, e2e_s, dev_ms, fast_ms = agg_max(step_s), agg_max(e2e_s), agg_max(dev_ms), agg_max(fast_ms)
    exch_ms = agg_max(exch_ms) if exch_ms else []
    if xprof and rank == 0:
        m = np.median(np.array(xprof), axis=0)
        sys.stderr.write(f"[exchange ms] export={m[0]:.3f} barrier1={m[1]:.3f} merge={m[2]:.3f} barrier2={m[3]:.3f}\n")
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    # device time of a step, CUDA events on the streams the work is enqueued on, max over ranks per step:
    # search (engine stream) [+ exchange (default stream) for N > 1]; the wall-clocked step is reported beside it
    step_ms = float(np.mean(dev_ms))
    value = args.batch / (step_ms / 1e3)
    e2e = args.batch / float(np.mean(e2e_s))
    peak, peak_src = read_peaks()
    nlists_est = 8.0 * 9                                         # 8 B x ~9 gram lookups per query (negligible next to postings)
    algo_bytes = float(np.mean(postings)) * 4.0 + nb * nlists_est + 8.0 * float(np.mean(results))
    k_ms = float(np.mean(fast_ms))
    achieved = algo_bytes / (k_ms / 1e3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    # (the ncu capture is of the default single-GPU workload: other sizes / shard counts have no traffic figure)
    if os.path.exists(tp) and world == 1 and args.keys == 10_000_000 and args.batch == 4096 and args.limit == 100 and args.threshold == 0.0:
        try:
            with open(tp) as f:
                traffic = json.load(f).get(("k_sieve_search" if last_st.get("nSieve", 0) * 2 > nb else "k_fast_search") + "_dram_bytes_per_launch")
        except (OSError, ValueError):
            traffic = None
    line = {"metric": METRIC, "value": value, "unit": "queries/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": step_ms, "wall_ms_per_step": 1e3 * float(np.mean(step_s)), "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u32",
            "data": "synthetic", "config": config, "clocks": clocks,
            "e2e": {"value": e2e, "unit": "queries/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": 1e3 * float(np.mean(e2e_s)),
                    "delivery": "N=1: the search kernel stores each finished query straight into the pinned ssb_batch arrays (device-to-host "
                                "over PCIe, overlapped with the kernel) when every query fills its limit; otherwise compaction + copies"},
            "gpu_launches": int(launches),   # kernels launched in the timed region (each step runs the path twice: e2e + resident)
            "queries_by_kernel": {"k_sieve_search": last_st.get("nSieve", 0),
                                  "k_fast_search": last_st.get("nFast", 0) - last_st.get("nSieve", 0),
                                  "dense": last_st.get("nDense", 0)},
            "roofline": {"bound": "hbm", "kernel": "k_sieve_search" if last_st.get("nSieve", 0) * 2 > nb else "k_fast_search",
                         "timed": "k_sieve_search + k_fast_search launches of one batch (CUDA events on the engine's stream)",
                         "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": algo_bytes, "kernel_ms": k_ms},
            "index_build": {"keys_per_s": args.keys / t_build, "seconds": t_build, "cold_first_build_seconds": t_build_cold,
                            "entry": "ssb_index_create(char**): host packing + H2D + device build + key-table D2H", "rank0_ms": bms,
                            "device_only_keys_per_s": (hi - lo) / max((bms["total"] - bms["pack"]) / 1e3, 1e-9),
                            "terms": info["T"], "grams": info["G"], "postings": info["P"], "device_bytes": info["deviceBytes"]},
            "host_cores": cores}
    if traffic:
        # DRAM bytes the kernel really moves per launch (ncu capture of the same command, profiles/traffic.json) over the live kernel time:
        # how busy HBM is, next to `frac`, which counts only the bytes the reference's algorithm needs
        line["roofline"]["traffic_gbs"] = traffic / (k_ms / 1e3) / 1e9
        line["roofline"]["traffic_frac"] = line["roofline"]["traffic_gbs"] / peak
    if global_sizes:
        line["index_build"]["global_sizes"] = global_sizes
    if exch_ms:
        line["exchange_ms_per_step"] = float(np.mean(exch_ms))      # device time of export + cross-shard merge (max over ranks), part of ms_per_step
    if world == 1 and not args.no_cpu_baseline:
        try:
            cb = reference_leg(args, args.cpu_keys, min(args.cpu_queries, args.batch), cores, steps=1, warmup=0)
            line["cpu_baseline"] = {"value": cb["qps"], "unit": "queries/s", "cores": cores, "kind": "reference",
                                    "sample": f"unmodified reference (oracle/_ref): indexN over the first {args.cpu_keys} keys, "
                                              f"{min(args.cpu_queries, args.batch)} queries via score()/release(), {cores} threads",
                                    "index_build_keys_per_s": cb["build_keys_per_s"], "one_thread_value": cb["qps_one_thread"]}
        except Exception as e:  # the reference .so is test infrastructure; its absence must not hide the GPU numbers
            line["cpu_baseline"] = {"value": None, "unit": "queries/s", "cores": cores, "kind": "reference", "sample": f"unavailable: {e}"}
    print(json.dumps(line))
    ix.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
