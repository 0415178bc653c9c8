"""Device timeline of the pieces of one host batch (JG_TRACE=1): where the e2e step's time goes.

    JG_TRACE=1 python tools/trace_e2e.py [chunk_queries (0 = the library's schedule)] [n_streams] [indexed]
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import jagua_rs_b200 as jg

chunk = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 18
streams = int(sys.argv[2]) if len(sys.argv) > 2 else 3
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
indexed = len(sys.argv) > 3 and sys.argv[3] == "indexed"
if indexed:
    h_r = jg.PinnedBuffer((n,), np.uint8)
    h_t = jg.PinnedBuffer((n, 2), np.float32)
    for i, m in enumerate(sizes):
        cde.set_item_rotations(i, wl.items[i].rotations)
        r, t = jg.make_queries_indexed(wl, i, m, seed=0)
        h_r.array[offs[i]:offs[i + 1]] = r
        h_t.array[offs[i]:offs[i + 1]] = t
cde.set_batching(chunk, streams)
for it in range(6):
    t0 = time.perf_counter()
    if indexed:
        cde.detect_multi_indexed(ids, offs, h_r.array, h_t.array, out=host_out.array, wait=False)
    else:
        cde.detect_multi(ids, offs, host_t.array, out=host_out.array, wait=False)
    t1 = time.perf_counter()
    cde.sync()
    t2 = time.perf_counter()
    print(f"step {it}: enqueue {(t1 - t0) * 1e6:.0f} us, total {(t2 - t0) * 1e6:.0f} us", file=sys.stderr)
