"""Detect throughput with two generated piers per item (fail_fast/piers.rs) on the grid and on the walk kernels.

    python tools/pier_probe.py
"""
import dataclasses
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import jagua_rs_b200 as jg
wl = jg.load_workload("swim_half")
items = [dataclasses.replace(it, piers=jg.generate_piers(it.verts, it.poles[:wl.n_ff_poles], 2, 24, 10, 24)) for it in wl.items]
wlp = dataclasses.replace(wl, items=items, n_ff_piers=2)
cde = jg.build_engine(wlp, device=0); cde.sync_layout()
n = 1 << 20
ids = list(range(len(wlp.items)))
sizes = jg.split_batch(n, len(ids)); offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
t = np.concatenate([jg.make_queries(wlp, i, m, seed=0) for i, m in enumerate(sizes)])
dt_ = jg.DeviceBuffer(cde, n * 16); dt_.upload(t); do = jg.DeviceBuffer(cde, n)
for walk in (False, True):
    cde.set_detect_path(walk)
    for _ in range(5): cde.detect_multi_device(ids, offs, dt_.ptr, do.ptr)
    cde.sync(); cde.timer_start()
    for _ in range(20): cde.detect_multi_device(ids, offs, dt_.ptr, do.ptr)
    ms = cde.timer_stop() / 20
    v = do.download((n,), np.uint8)
    print("walk" if walk else "grid", f"{ms*1e3:.1f} us  {n/ms/1e6:.2f} G q/s", np.bincount(v, minlength=3)[:3])
