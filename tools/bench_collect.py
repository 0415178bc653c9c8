"""BASELINE.json config 3: hazard-set collection (collect_surrogate_collisions + collect_poly_collisions into one
collector, lbf/benches/ci_bench.rs:56-70) on the full shirts and trousers layouts. Prints one JSON line per
workload: device-resident and end-to-end queries/s, the CPU restatement on all host cores, and the number of
hazard-set mismatches on the compared sample (must be 0).

    python tools/bench_collect.py [--queries 10000000] [--check 400000]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/bench_collect.py --gpus N        # the 10M batch sharded over N GPUs (strong scaling, contiguous slices)

N > 1: one process per GPU, the layout is flattened on rank 0 and broadcast (NCCL), every rank resolves the contiguous
slice [r * n / N, (r + 1) * n / N) of every item's queries; no collective on the query path; times are the max over ranks.
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

ap = argparse.ArgumentParser()
ap.add_argument("--queries", type=int, default=10_000_000)
ap.add_argument("--check", type=int, default=400_000)
ap.add_argument("--workloads", default="shirts_full,trousers_full")
ap.add_argument("--gpus", type=int, default=1)
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
    from jagua_rs_b200 import multigpu

import orc  # noqa: E402  (checker + CPU baseline only)
from helpers import oracle_world  # noqa: E402

orc.build()
for name in args.workloads.split(","):
    wl = jg.load_workload(name)
    if world > 1:
        if rank == 0:
            cde = jg.build_engine(wl, device=local_rank)
            blob = cde.flatten()
        else:
            cde = jg.CDEngine(wl.cde_bbox, wl.quadtree_depth, wl.cd_threshold, wl.n_ff_poles, wl.n_ff_piers, device=local_rank)
            for i, it in enumerate(wl.items):
                cde.upload_item(i, jg.ItemShape(it.verts, it.poi, it.poles, it.piers))
            blob = None
        d_blob = multigpu.broadcast_blob(blob, dist, src=0, device=f"cuda:{local_rank}")
        torch.cuda.synchronize()
        cde.upload_layout_device(d_blob.data_ptr(), d_blob.numel(), epoch=1)
        n_slots = torch.tensor([cde.slot_capacity()], device="cuda")
        dist.broadcast(n_slots, 0)
        words = max(1, (int(n_slots.item()) + 31) // 32)
    else:
        cde = jg.build_engine(wl, device=local_rank)
        cde.sync_layout()
        words = cde.collector_words()
    w = oracle_world(orc, wl) if rank == 0 else None
    sizes = jg.split_batch(args.queries, len(wl.items))
    dev_ms = e2e_s = 0.0
    mism = checked = 0
    cpu_s = cpu_q = 0
    cores = os.cpu_count() or 1
    for item_id, n_all in enumerate(sizes):
        lo, hi = (0, n_all) if world == 1 else multigpu.shard_bounds(n_all, rank, world)
        n = hi - lo
        t = np.ascontiguousarray(jg.make_queries(wl, item_id, n_all, seed=0)[lo:hi])
        ht = jg.PinnedBuffer((n, 4), np.float32)
        ht.array[:] = t
        ho = jg.PinnedBuffer((n, words), np.uint32)
        d_t = jg.DeviceBuffer(cde, n * 16)
        d_t.upload(t)
        d_o = jg.DeviceBuffer(cde, n * 4 * words)
        cde.collect_batch_device(item_id, d_t.ptr, n, d_o.ptr, words)  # warm-up
        cde.sync()
        cde.timer_start()
        cde.collect_batch_device(item_id, d_t.ptr, n, d_o.ptr, words)
        dev_ms += cde.timer_stop()
        cde.collect_batch_async(item_id, ht.array, ho.array)  # warm-up
        cde.sync()
        t0 = time.perf_counter()
        cde.collect_batch_async(item_id, ht.array, ho.array)
        cde.sync()
        e2e_s += time.perf_counter() - t0
        if rank == 0:
            m = min(n, max(args.check // len(sizes), 1))
            t0 = time.perf_counter()
            exp = w.collect_batch(item_id, t[:m], words, nthreads=cores)
            cpu_s += time.perf_counter() - t0
            cpu_q += m
            mism += int((exp != ho.array[:m]).any(axis=1).sum())
            mism += int((d_o.download((n, words), np.uint32)[:m] != exp).any(axis=1).sum())
            checked += m
        for b in (ht, ho):
            b.free()
        d_t.free(), d_o.free()
    total = sum(sizes)
    if dist is not None:
        tt = torch.tensor([dev_ms, e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dev_ms, e2e_s = float(tt[0]), float(tt[1])
    if rank == 0:
      print(json.dumps({"metric": "collision queries/sec (hazard-set collection)", "workload": name, "queries": total, "n_gpus": world, "scaling": "strong",
                      "value": total / (dev_ms * 1e-3), "unit": "queries/s", "ms_total": dev_ms,
                      "e2e": {"value": total / e2e_s, "h2d_bytes": 16 * total, "d2h_bytes": 4 * words * total},
                      "words_per_query": words, "path": cde.detect_path(),
                      "cpu_baseline": {"value": cpu_q / cpu_s, "cores": cores, "kind": "port", "sample": f"{cpu_q} queries"},
                      "checked_queries": checked, "mismatches_vs_oracle": mism}))
    cde.close()
if dist is not None:
    dist.destroy_process_group()
