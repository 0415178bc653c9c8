"""Device-resident detect time vs batch size (CUDA events on the compute stream, warm L2).

    [JG_SWEEP_WORKLOAD=gardeyn4c_half] python tools/size_sweep.py [n_queries ...]
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import jagua_rs_b200 as jg

wl = jg.load_workload(os.environ.get("JG_SWEEP_WORKLOAD", "swim_half"))
cde = jg.build_engine(wl, device=0)
cde.sync_layout()
ids = list(range(len(wl.items)))
SIZES = [int(a) for a in sys.argv[1:]] or [1 << 14, 1 << 15, 1 << 16, 1 << 17, 1 << 18, 1 << 19, 1 << 20, 1 << 22]
for n in SIZES:
    sizes = jg.split_batch(n, len(wl.items))
    offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    t = np.concatenate([jg.make_queries(wl, i, m, seed=0) for i, m in enumerate(sizes)])
    dt_ = jg.DeviceBuffer(cde, n * 16)
    dt_.upload(t)
    do = jg.DeviceBuffer(cde, n)
    for _ in range(5):
        cde.detect_multi_device(ids, offs, dt_.ptr, do.ptr)
    cde.sync()
    reps = 20
    cde.timer_start()
    for _ in range(reps):
        cde.detect_multi_device(ids, offs, dt_.ptr, do.ptr)
    ms = cde.timer_stop() / reps
    print(f"n={n:8d}  {ms * 1e3:8.1f} us/batch  {n / ms / 1e6:8.2f} G q/s")
    dt_.free(), do.free()
