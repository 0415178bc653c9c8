"""Counts the ALGORITHMIC flops of a bench batch with the instrumented oracle (liboracle_count.so):
the sum, over the primitives the reference algorithm executes for each query with its early exits, of
the fixed per-primitive table of SURVEY.md §8(d) / DESIGN.md. The result is committed next to the
workload (workloads/<name>.flops.json) and is the numerator of bench.py's roofline.

    python tools/count_flops.py swim_half 1048576 [seed]
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


def main():
    name, total = sys.argv[1], int(sys.argv[2])
    seed = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    wl = jg.load_workload(name)
    w = oracle_world(orc, wl, count=True)
    sizes = jg.split_batch(total, len(wl.items))
    out = dict(workload=name, queries=total, seed=seed, items=[])
    tot_d = tot_c = 0
    hist = np.zeros(256, np.int64)
    for i, n in enumerate(sizes):
        t = jg.make_queries(wl, i, n, seed)
        v = w.detect_batch(i, t, nthreads=os.cpu_count() or 1)  # (thread-local counters, summed)
        fd = w.last_flops
        w.collect_batch(i, t, max(1, (w.haz_capacity() + 31) // 32), nthreads=os.cpu_count() or 1)
        fc = w.last_flops
        hist += np.bincount(v, minlength=256)
        out["items"].append(dict(item=i, n_verts=len(wl.items[i].verts), queries=n, detect_flops=fd, collect_flops=fc))
        tot_d += fd
        tot_c += fc
    out["detect_flops_total"] = tot_d
    out["detect_flops_per_query"] = tot_d / total
    out["collect_flops_total"] = tot_c
    out["collect_flops_per_query"] = tot_c / total
    out["verdicts"] = dict(surrogate=int(hist[1]), poly=int(hist[2]), free=int(hist[0]), ref_panic=int(hist[0x80]))
    path = os.path.join(ROOT, "workloads", f"{name}.flops.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps({k: v for k, v in out.items() if k != "items"}))


if __name__ == "__main__":
    main()
