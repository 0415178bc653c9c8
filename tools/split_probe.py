"""Device-resident detect time with the containment tests inline (k_s2_grid) and in their own kernel (k_s2_contain),
per batch size (CUDA events on the compute stream, warm L2).

    python tools/split_probe.py [sizes...]
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import jagua_rs_b200 as jg

wl = jg.load_workload(os.environ.get("JG_WORKLOAD", "swim_half"))
cde = jg.build_engine(wl, device=0)
cde.sync_layout()
ids = list(range(len(wl.items)))
SIZES = [int(a) for a in sys.argv[1:]] or [1 << 16, 1 << 18, 1 << 20, 1 << 21, 1 << 22]
for n in SIZES:
    sizes = jg.split_batch(n, len(wl.items))
    offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    t = np.concatenate([jg.make_queries(wl, i, m, seed=0) for i, m in enumerate(sizes)])
    dt_ = jg.DeviceBuffer(cde, n * 16)
    dt_.upload(t)
    do = jg.DeviceBuffer(cde, n)
    res = {}
    for name, split_min in (("inline", 1 << 40), ("split", 0)):
        cde.set_split_min(split_min)
        for _ in range(5):
            cde.detect_multi_device(ids, offs, dt_.ptr, do.ptr)
        cde.sync()
        reps = 20
        cde.timer_start()
        for _ in range(reps):
            cde.detect_multi_device(ids, offs, dt_.ptr, do.ptr)
        res[name] = cde.timer_stop() / reps
    print(f"n={n:8d}  inline {res['inline'] * 1e3:8.1f} us   split {res['split'] * 1e3:8.1f} us")
    dt_.free(), do.free()
