"""BASELINE.json config 3: hazard-set collection (collect_surrogate_collisions + collect_poly_collisions into one
collector, lbf/benches/ci_bench.rs:56-70) on the full shirts and trousers layouts. Prints one JSON line per
workload: device-resident and end-to-end queries/s, the CPU restatement on all host cores, and the number of
hazard-set mismatches on the compared sample (must be 0).

    python tools/bench_collect.py [--queries 10000000] [--check 400000]
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
args = ap.parse_args()

import orc  # noqa: E402  (checker + CPU baseline only)
from helpers import oracle_world  # noqa: E402

orc.build()
for name in args.workloads.split(","):
    wl = jg.load_workload(name)
    cde = jg.build_engine(wl, device=0)
    cde.sync_layout()
    w = oracle_world(orc, wl)
    words = cde.collector_words()
    sizes = jg.split_batch(args.queries, len(wl.items))
    dev_ms = e2e_s = 0.0
    mism = checked = 0
    cpu_s = cpu_q = 0
    cores = os.cpu_count() or 1
    for item_id, n in enumerate(sizes):
        t = jg.make_queries(wl, item_id, n, seed=0)
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
    print(json.dumps({"metric": "collision queries/sec (hazard-set collection)", "workload": name, "queries": total,
                      "value": total / (dev_ms * 1e-3), "unit": "queries/s", "ms_total": dev_ms,
                      "e2e": {"value": total / e2e_s, "h2d_bytes": 16 * total, "d2h_bytes": 4 * words * total},
                      "words_per_query": words, "path": cde.detect_path(),
                      "cpu_baseline": {"value": cpu_q / cpu_s, "cores": cores, "kind": "port", "sample": f"{cpu_q} queries"},
                      "checked_queries": checked, "mismatches_vs_oracle": mism}))
    cde.close()
