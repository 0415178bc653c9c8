"""e2e throughput with several host batches in flight (detect_multi(wait=False) back to back, one sync at the end)
vs one batch at a time, over the batcher settings. 16 B/query composed records and the 9 B/query compact form."""
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
host_r = jg.PinnedBuffer((n,), np.uint8)
host_xy = jg.PinnedBuffer((n, 2), np.float32)
for i, m in enumerate(sizes):
    host_t.array[offs[i]:offs[i + 1]] = jg.make_queries(wl, i, m, seed=0)
    cde.set_item_rotations(i, wl.items[i].rotations)
    r, t = jg.make_queries_indexed(wl, i, m, seed=0)
    host_r.array[offs[i]:offs[i + 1]] = r
    host_xy.array[offs[i]:offs[i + 1]] = t
outs = [jg.PinnedBuffer((n,), np.uint8) for _ in range(4)]
ids = list(range(len(wl.items)))
K = 40
for chunk, streams in ((1 << 18, 3), (1 << 18, 4), (1 << 18, 6), (1 << 18, 8), (1 << 19, 4), (1 << 17, 8), (1 << 20, 2), (1 << 20, 3)):
    cde.set_batching(chunk, streams)
    res = []
    for compact in (False, True):
        def enq(i):
            if compact:
                cde.detect_multi_indexed(ids, offs, host_r.array, host_xy.array, out=outs[i % 4].array, wait=False)
            else:
                cde.detect_multi(ids, offs, host_t.array, out=outs[i % 4].array, wait=False)
        for i in range(5):
            enq(i)
        cde.sync()
        t0 = time.perf_counter()
        for i in range(K):
            enq(i)
            cde.sync()
        serial = (time.perf_counter() - t0) / K
        t0 = time.perf_counter()
        for i in range(K):
            enq(i)
        cde.sync()
        piped = (time.perf_counter() - t0) / K
        res.append((serial, piped))
    print(f"chunk {chunk:8d} streams {streams}: 16B serial {res[0][0]*1e3:.3f} ms piped {res[0][1]*1e3:.3f} ms ({n/res[0][1]/1e9:.2f} G q/s) | "
          f"9B serial {res[1][0]*1e3:.3f} ms piped {res[1][1]*1e3:.3f} ms ({n/res[1][1]/1e9:.2f} G q/s)")
