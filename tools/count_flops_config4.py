"""Algorithmic flops of the config-4 detect batches of tools/bench_detect.py (every item type queried under the
HazKeyFilter::from_irrelevant_qzones filter of its own min_quality, seed 5), counted by the instrumented oracle
(liboracle_count.so) with the per-primitive table of SURVEY.md section 8(d). Written next to the workload as
workloads/<name>.detect_filtered.flops.json: the numerator of bench_detect.py's roofline.

    python tools/count_flops_config4.py [workload ...] [--queries 4194304]
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import orc  # noqa: E402
import jagua_rs_b200 as jg  # noqa: E402
from helpers import oracle_world  # noqa: E402

args = [a for a in sys.argv[1:] if not a.startswith("--")]
total = int(sys.argv[sys.argv.index("--queries") + 1]) if "--queries" in sys.argv else 1 << 22
names = args or ["baldacci1_bin", "baldacci3_bin", "gardeyn1c_half", "gardeyn4c_half"]
for name in names:
    wl = jg.load_workload(name)
    w = oracle_world(orc, wl, count=True)
    words = max(1, (w.haz_capacity() + 31) // 32)
    per_item = max(total // len(wl.items), 1)  # (bench_detect.py's split)
    flops = 0
    hist = np.zeros(256, np.int64)
    for i, it in enumerate(wl.items):
        t = jg.make_queries(wl, i, per_item, seed=5)
        bits = w.filter_qzones(it.min_quality, words) if it.min_quality >= 0 else None
        v = w.detect_batch(i, t, filter_bits=bits, nthreads=os.cpu_count() or 1)
        flops += w.last_flops
        hist += np.bincount(v, minlength=256)
    n = per_item * len(wl.items)
    out = dict(workload=name, queries=n, seed=5, per_item=per_item, detect_flops_total=int(flops), detect_flops_per_query=flops / n,
               verdicts=dict(surrogate=int(hist[1]), poly=int(hist[2]), free=int(hist[0]), ref_panic=int(hist[0x80])))
    with open(os.path.join(ROOT, "workloads", f"{name}.detect_filtered.flops.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out))
