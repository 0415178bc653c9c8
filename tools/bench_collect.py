"""BASELINE.json config 3: hazard-set collection (collect_surrogate_collisions + collect_poly_collisions into one
collector, lbf/benches/ci_bench.rs:56-70) on the full shirts and trousers layouts, 10M queries. Prints one JSON line per
workload: device-resident and end-to-end queries/s, the FP32 (non-FMA) roofline fraction from the algorithmic flops the
instrumented oracle counted for exactly this batch (workloads/<name>.flops.json, tools/count_flops.py), the CPU restatement on
all host cores, and the number of hazard-set mismatches over ALL queries of the batch (must be 0).

    python tools/bench_collect.py [--queries 10000000]
    python tools/bench_collect.py --pool N      # ONE process, the batch sharded over N GPUs by jg_pool_collect_multi
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/bench_collect.py --gpus N        # one process per GPU, rank r resolves the slice [r n/N, (r+1) n/N)

One jg_collect_multi call = ONE k_collect_grid launch for all item types of the (slice of the) batch. Strong scaling: the
batch is fixed, times are the max over ranks; every rank checks its own slice against the oracle (all queries are compared).
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import jagua_rs_b200 as jg  # noqa: E402
from jagua_rs_b200 import multigpu  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--queries", type=int, default=10_000_000)
ap.add_argument("--workloads", default="shirts_full,trousers_full")
ap.add_argument("--gpus", type=int, default=1)
ap.add_argument("--pool", type=int, default=0)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--no-check", action="store_true")
args = ap.parse_args()
rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local_rank = int(os.environ.get("LOCAL_RANK", "0"))
dist = None
if world > 1:
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

import orc  # noqa: E402  (checker + CPU baseline only)
from helpers import oracle_world  # noqa: E402

orc.build()
cores = max(1, (os.cpu_count() or 1) // max(world, 1))
for name in args.workloads.split(","):
    wl = jg.load_workload(name)
    n_items = len(wl.items)
    sizes = jg.split_batch(args.queries, n_items)
    full_offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    total = int(full_offs[-1])
    lo, hi, seg_ids, offs = multigpu.shard_segments(full_offs, rank, world)
    # this rank's slice of the seeded batch (segment k of the slice = part of item seg_ids[k]'s queries)
    parts = []
    for k, item_id in enumerate(seg_ids):
        a = max(int(full_offs[item_id]), lo) - int(full_offs[item_id])
        parts.append(jg.make_queries(wl, item_id, sizes[item_id], seed=0)[a:a + int(offs[k + 1] - offs[k])])
    t = np.ascontiguousarray(np.concatenate(parts)) if parts else np.zeros((0, 4), np.float32)
    n = hi - lo
    bcast_ms = None
    if args.pool > 0:
        host = jg.build_engine(wl, device=None)
        eng = jg.EnginePool(host, list(range(args.pool)))
        eng.upload_items(wl.items)
        eng.sync_layout()
        words = host.collector_words()
        path = "pool"
    else:
        if world > 1:
            if rank == 0:
                eng = jg.build_engine(wl, device=local_rank)
            else:
                eng = jg.CDEngine(wl.cde_bbox, wl.quadtree_depth, wl.cd_threshold, wl.n_ff_poles, wl.n_ff_piers, device=local_rank)
                for i, it in enumerate(wl.items):
                    eng.upload_item(i, jg.ItemShape(it.verts, it.poi, it.poles, it.piers))
            bc = multigpu.LayoutBroadcaster(eng, dist, rank, world)
            bc.broadcast(root=0)
            bcast_ms = bc.info["last_ms"]
            n_slots = torch.tensor([eng.slot_capacity()], device="cuda")
            dist.broadcast(n_slots, 0)
            words = max(1, (int(n_slots.item()) + 31) // 32)
        else:
            eng = jg.build_engine(wl, device=local_rank)
            eng.sync_layout()
            words = eng.collector_words()
        path = eng.detect_path()
    ht = jg.PinnedBuffer((n, 4), np.float32)
    ht.array[:] = t
    ho = jg.PinnedBuffer((n, words), np.uint32)
    dev_ms = None
    launches = None
    if args.pool == 0:
        d_t = jg.DeviceBuffer(eng, n * 16)
        d_t.upload(t)
        d_o = jg.DeviceBuffer(eng, n * 4 * words)
        eng.collect_multi_device(seg_ids, offs, d_t.ptr, d_o.ptr, words)  # warm-up
        eng.sync()
        l0 = eng.stats()["launches"]
        best = 1e30
        for _ in range(args.reps):
            eng.timer_start()
            eng.collect_multi_device(seg_ids, offs, d_t.ptr, d_o.ptr, words)
            best = min(best, eng.timer_stop())
        dev_ms = best
        launches = (eng.stats()["launches"] - l0) // args.reps
    eng.collect_multi(seg_ids, offs, ht.array, words, out=ho.array, wait=False)  # warm-up
    eng.sync()
    if dist is not None:
        dist.barrier()
    e2e_s = 1e30
    for _ in range(args.reps):
        t0 = time.perf_counter()
        eng.collect_multi(seg_ids, offs, ht.array, words, out=ho.array, wait=False)
        eng.sync()
        e2e_s = min(e2e_s, time.perf_counter() - t0)
    mism = checked = 0
    cpu_s = 0.0
    if not args.no_check and n > 0:
        w = oracle_world(orc, wl)
        t0 = time.perf_counter()
        exp = w.collect_multi(seg_ids, offs, t, words, nthreads=cores)
        cpu_s = time.perf_counter() - t0
        mism += int((exp != ho.array).any(axis=1).sum())
        if args.pool == 0:
            mism += int((d_o.download((n, words), np.uint32) != exp).any(axis=1).sum())
        checked = n
    if dist is not None:
        tt = torch.tensor([dev_ms, e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dev_ms, e2e_s = float(tt[0]), float(tt[1])
        cc = torch.tensor([mism, checked], dtype=torch.int64, device="cuda")
        dist.all_reduce(cc, op=dist.ReduceOp.SUM)
        mism, checked = int(cc[0]), int(cc[1])
    if rank == 0:
        roofline = None
        try:
            with open(os.path.join(ROOT, "workloads", f"{name}.flops.json")) as f:
                fl = json.load(f)
            if fl["queries"] == total and dev_ms is not None:
                peak = None
                if args.pool == 0:
                    peak, _ = eng.fp32_nofma_peak()
                ach = fl["collect_flops_total"] / (dev_ms * 1e-3) / 1e12 / world
                roofline = {"bound": "fp32_nofma", "achieved": ach, "peak": peak / 1e3 if peak else None, "unit": "TFLOP/s per GPU",
                            "frac": ach / (peak / 1e3) if peak else None, "flops_per_query": fl["collect_flops_per_query"],
                            "flops_source": f"workloads/{name}.flops.json (instrumented oracle, this exact batch)",
                            "peak_source": "measured live: jg_fp32_nofma_peak"}
        except (OSError, KeyError):
            pass
        line = {"metric": "collision queries/sec (hazard-set collection)", "workload": name, "queries": total, "n_gpus": max(world, args.pool, 1),
                "mode": "pool (one process)" if args.pool else ("one process per GPU" if world > 1 else "one GPU"), "scaling": "strong",
                "value": total / (dev_ms * 1e-3) if dev_ms else None, "unit": "queries/s", "ms_total": dev_ms, "kernel_launches_per_batch": launches,
                "e2e": {"value": total / e2e_s, "ms_total": e2e_s * 1e3, "h2d_bytes": 16 * total, "d2h_bytes": 4 * words * total},
                "words_per_query": words, "path": path, "roofline": roofline, "layout_broadcast_ms": bcast_ms,
                "cpu_baseline": {"value": (n / cpu_s) if cpu_s else None, "cores": cores, "kind": "port",
                                 "sample": f"rank 0's slice ({n} queries), oracle -O3 -march=native, persistent pool"},
                "checked_queries": checked, "mismatches_vs_oracle": mism}
        print(json.dumps(line), flush=True)
    ht.free(), ho.free()
    eng.close()
    if args.pool > 0:
        host.close()
if dist is not None:
    dist.destroy_process_group()
