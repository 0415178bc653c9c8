"""GPU sweep of the host-batcher parameters (chunk size x in-flight chunks) for the bench workload.

    python tools/sweep_e2e.py            # prints ms per 1M-query step for each setting
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import jagua_rs_b200 as jg

wl = jg.load_workload("swim_half")
cde = jg.build_engine(wl, device=0)
cde.sync_layout()
n = 1 << 20
sizes = jg.split_batch(n, len(wl.items))
offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
host_t = jg.PinnedBuffer((n, 4), np.float32)
for i, m in enumerate(sizes):
    host_t.array[offs[i]:offs[i + 1]] = jg.make_queries(wl, i, m, seed=0)
host_out = jg.PinnedBuffer((n,), np.uint8)
ids = list(range(len(wl.items)))
for chunk in (1 << 16, 1 << 17, 1 << 18):
    for streams in (3, 6):
        cde.set_batching(chunk, streams)
        for _ in range(5):
            cde.detect_multi(ids, offs, host_t.array, out=host_out.array, wait=False)
            cde.sync()
        t0 = time.perf_counter()
        for _ in range(20):
            cde.detect_multi(ids, offs, host_t.array, out=host_out.array, wait=False)
            cde.sync()
        dt = (time.perf_counter() - t0) / 20
        print(f"chunk {chunk:7d} streams {streams}: {dt * 1e3:.3f} ms/step  {n / dt / 1e9:.2f} G q/s")
