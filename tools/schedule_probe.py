"""Host-buffer detect path, 9 B/query indexed form: one batch at a time and batches in flight under fixed piece sizes
(jg_engine_set_batching(chunk, streams)) and under the per-batch schedule (chunk = 0).

    python tools/schedule_probe.py [n_queries ...]
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
ids = list(range(len(wl.items)))
SIZES = [int(a) for a in sys.argv[1:]] or [1 << 18, 1 << 19, 1 << 20, 1 << 21, 1 << 22]
for n in SIZES:
    sizes = jg.split_batch(n, len(wl.items))
    offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    ridx = jg.PinnedBuffer((n,), np.uint8)
    txy = jg.PinnedBuffer((n, 2), np.float32)
    for k, (i, m) in enumerate(zip(ids, sizes)):
        r, t = jg.make_queries_indexed(wl, i, m, seed=0)
        ridx.array[offs[k]:offs[k + 1]] = r
        txy.array[offs[k]:offs[k + 1]] = t
    for i in ids:
        cde.set_item_rotations(i, wl.items[i].rotations)
    outs = [jg.PinnedBuffer((n,), np.uint8) for _ in range(4)]
    ref = None
    for label, chunk, streams in (("fixed 256Ki x4", 1 << 18, 4), ("fixed 1Mi x3", 1 << 20, 3), ("auto x3", 0, 3), ("auto x4", 0, 4)):
        cde.set_batching(chunk, streams)
        res = {}
        for pipelined in (False, True):
            for i in range(4):
                cde.detect_multi_indexed(ids, offs, ridx.array, txy.array, out=outs[i % 4].array, wait=False)
                cde.sync()
            reps = 20
            t0 = time.perf_counter()
            for i in range(reps):
                cde.detect_multi_indexed(ids, offs, ridx.array, txy.array, out=outs[i % 4].array, wait=False)
                if not pipelined:
                    cde.sync()
            cde.sync()
            res[pipelined] = (time.perf_counter() - t0) / reps
        if ref is None:
            ref = outs[0].array.copy()
        same = all(np.array_equal(o.array, ref) for o in outs)
        print(f"n={n:8d} {label:15s} one at a time {res[False] * 1e3:7.3f} ms ({n / res[False] / 1e9:5.2f} G q/s)   in flight {res[True] * 1e3:7.3f} ms ({n / res[True] / 1e9:5.2f} G q/s)  same verdicts {same}")
    for b in outs + [ridx, txy]:
        b.free()
