"""What would spatially coherent warps buy? The bench batch as is (uniform random order inside each item segment) vs the
same queries sorted, per item segment, by the finest grid cell of their translation (and by rotation). Device-resident."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import jagua_rs_b200 as jg

wl = jg.load_workload(sys.argv[1] if len(sys.argv) > 1 else "swim_half")
cde = jg.build_engine(wl, device=0)
cde.sync_layout()
n = 1 << 20
sizes = jg.split_batch(n, len(wl.items))
offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
ids = list(range(len(wl.items)))
b = wl.cde_bbox
base = [jg.make_queries(wl, i, m, seed=0) for i, m in enumerate(sizes)]


def sort_by_cell(t, g):
    ci = np.clip(((t[:, 2] - b[0]) / (b[2] - b[0]) * g).astype(np.int64), 0, g - 1)
    cj = np.clip(((t[:, 3] - b[1]) / (b[3] - b[1]) * g).astype(np.int64), 0, g - 1)
    rot = (t[:, 0] < 0).astype(np.int64)
    return t[np.lexsort((ci, cj, rot))]


ref = None
for name, prep in (("random order", lambda t: t), ("sorted by 32x32 cell", lambda t: sort_by_cell(t, 32)),
                   ("sorted by 8x8 cell", lambda t: sort_by_cell(t, 8))):
    t = np.ascontiguousarray(np.concatenate([prep(x) for x in base]))
    d_t = jg.DeviceBuffer(cde, n * 16)
    d_t.upload(t)
    d_o = jg.DeviceBuffer(cde, n)
    for _ in range(3):
        cde.detect_multi_device(ids, offs, d_t.ptr, d_o.ptr)
    cde.sync()
    cde.timer_start()
    for _ in range(10):
        cde.detect_multi_device(ids, offs, d_t.ptr, d_o.ptr)
    ms = cde.timer_stop() / 10
    v = d_o.download((n,), np.uint8)
    print(f"{name}: {ms * 1e3:.1f} us per 1M-query batch ({n / ms / 1e6:.2f} G q/s); verdict counts {np.bincount(v, minlength=3)[:3]}")
    d_t.free(), d_o.free()
