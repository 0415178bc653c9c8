"""Detect-mode throughput on any workload fixture (BASELINE.json config 4 and the other parity layouts): every
item type is queried under the HazKeyFilter::from_irrelevant_qzones filter of its own min_quality
(hazards/filter.rs:21-41, callers lbf_bpp.rs:125-127; items without one use NoFilter). One detect_multi call per
filter. Prints one JSON line per workload: device-resident and end-to-end queries/s, the verdict split, the CPU
restatement on all host cores and the mismatches over ALL queries of the batch (must be 0; --check N compares the first
N queries per item type only).

    python tools/bench_detect.py [--workloads baldacci1_bin,baldacci3_bin,gardeyn1c_half,gardeyn4c_half]
                                 [--queries 4194304] [--check 0]
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
ap.add_argument("--workloads", default="baldacci1_bin,baldacci3_bin,gardeyn1c_half,gardeyn4c_half")
ap.add_argument("--queries", type=int, default=1 << 22)
ap.add_argument("--check", type=int, default=0, help="0 = compare every query of the batch with the oracle (default)")
args = ap.parse_args()

import orc  # noqa: E402  (checker + CPU baseline only)
from helpers import oracle_world  # noqa: E402

orc.build()
cores = os.cpu_count() or 1
for name in args.workloads.split(","):
    wl = jg.load_workload(name)
    cde = jg.build_engine(wl, device=0)
    cde.sync_layout()
    w = oracle_world(orc, wl)
    words = cde.collector_words()
    # groups of item types sharing a filter
    groups = {}
    for i, it in enumerate(wl.items):
        groups.setdefault(it.min_quality, []).append(i)
    fbits = {}
    for fid, q in enumerate(sorted(groups)):
        if q >= 0:
            fbits[q] = (fid, cde.filter_from_irrelevant_qzones(q))
            cde.upload_filter(fid, fbits[q][1])
    per_item = max(args.queries // len(wl.items), 1)
    dev_ms = e2e_s = 0.0
    split = dict(surrogate=0, poly=0, free=0, ref_panic=0)
    mism = checked = 0
    cpu_s = 0.0
    total = 0
    for q, items in sorted(groups.items()):
        fid, bits = fbits.get(q, (-1, None))
        offs = np.arange(len(items) + 1, dtype=np.int64) * per_item
        n = int(offs[-1])
        ht = jg.PinnedBuffer((n, 4), np.float32)
        for k, i in enumerate(items):
            ht.array[offs[k]:offs[k + 1]] = jg.make_queries(wl, i, per_item, seed=5)
        ho = jg.PinnedBuffer((n,), np.uint8)
        d_t = jg.DeviceBuffer(cde, n * 16)
        d_t.upload(ht.array)
        d_o = jg.DeviceBuffer(cde, n)
        for _ in range(2):
            cde.detect_multi_device(items, offs, d_t.ptr, d_o.ptr, filter_id=fid)
        cde.sync()
        cde.timer_start()
        cde.detect_multi_device(items, offs, d_t.ptr, d_o.ptr, filter_id=fid)
        dev_ms += cde.timer_stop()
        cde.detect_multi(items, offs, ht.array, filter_id=fid, out=ho.array, wait=False)
        cde.sync()
        t0 = time.perf_counter()
        cde.detect_multi(items, offs, ht.array, filter_id=fid, out=ho.array, wait=False)
        cde.sync()
        e2e_s += time.perf_counter() - t0
        got_dev = d_o.download((n,), np.uint8)
        mism += int((got_dev != ho.array).sum())
        split["surrogate"] += int((got_dev == 1).sum())
        split["poly"] += int((got_dev == 2).sum())
        split["free"] += int((got_dev == 0).sum())
        split["ref_panic"] += int((got_dev == 0x80).sum())
        if args.check <= 0:  # 100 % of the batch, one pass of the oracle's thread pool over all segments
            t0 = time.perf_counter()
            exp = w.detect_multi(items, offs, ht.array, filter_bits=bits, nthreads=cores)
            cpu_s += time.perf_counter() - t0
            mism += int((exp != got_dev).sum())
            checked += n
        else:
            m = min(per_item, max(args.check // len(wl.items), 1))
            for k, i in enumerate(items):
                tq = np.ascontiguousarray(ht.array[offs[k]:offs[k] + m])
                t0 = time.perf_counter()
                exp = w.detect_batch(i, tq, filter_bits=bits, nthreads=cores)
                cpu_s += time.perf_counter() - t0
                mism += int((exp != got_dev[offs[k]:offs[k] + m]).sum())
                checked += m
        total += n
        ht.free(), ho.free(), d_t.free(), d_o.free()
    roofline = None
    try:  # algorithmic flops of exactly this batch (tools/count_flops_config4.py) against the non-FMA FP32 peak measured live
        with open(os.path.join(ROOT, "workloads", f"{name}.detect_filtered.flops.json")) as f:
            fl = json.load(f)
        if fl["queries"] == total:
            peak, _ = cde.fp32_nofma_peak()
            ach = fl["detect_flops_total"] / (dev_ms * 1e-3) / 1e12
            roofline = {"bound": "fp32_nofma", "achieved": ach, "peak": peak / 1e3, "unit": "TFLOP/s", "frac": ach / (peak / 1e3),
                        "flops_per_query": fl["detect_flops_per_query"],
                        "flops_source": f"workloads/{name}.detect_filtered.flops.json (instrumented oracle, this exact batch)",
                        "peak_source": "measured live: jg_fp32_nofma_peak"}
    except (OSError, KeyError):
        pass
    print(json.dumps({"metric": "collision queries/sec (detect, per-item quality-zone filters)", "workload": name, "queries": total,
                      "roofline": roofline,
                      "value": total / (dev_ms * 1e-3), "unit": "queries/s", "ms_total": dev_ms,
                      "e2e": {"value": total / e2e_s, "h2d_bytes": 16 * total, "d2h_bytes": total},
                      "filters": {str(q): int(np.unpackbits(b.view(np.uint8)).sum()) for q, (_, b) in fbits.items()},
                      "hazards": len(wl.hazards), "item_types": len(wl.items), "max_item_vertices": int(max(len(it.verts) for it in wl.items)),
                      "max_hazard_vertices": int(max(len(h.verts) for h in wl.hazards)), "quadtree_depth": wl.quadtree_depth,
                      "path": cde.detect_path(), "verdict_split": split,
                      "cpu_baseline": {"value": checked / cpu_s, "cores": cores, "kind": "port", "sample": f"{checked} queries"},
                      "checked_queries": checked, "mismatches_vs_oracle": mism}))
    cde.close()
