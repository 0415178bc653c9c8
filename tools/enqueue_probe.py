import os, sys, time
import numpy as np
sys.path.insert(0, '/root/repo')
import jagua_rs_b200 as jg
wl = jg.load_workload("swim_half"); cde = jg.build_engine(wl, device=0); cde.sync_layout()
n = 1 << 20
sizes = jg.split_batch(n, len(wl.items)); offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
host_t = jg.PinnedBuffer((n, 4), np.float32)
for i, m in enumerate(sizes): host_t.array[offs[i]:offs[i + 1]] = jg.make_queries(wl, i, m, seed=0)
host_out = jg.PinnedBuffer((n,), np.uint8); ids = list(range(len(wl.items)))
for chunk, streams in ((1 << 16, 16), (1 << 17, 8), (1 << 18, 4), (1<<19, 3)):
    cde.set_batching(chunk, streams)
    for _ in range(5):
        cde.detect_multi(ids, offs, host_t.array, out=host_out.array, wait=False); cde.sync()
    te = tt = 0
    for _ in range(20):
        t0 = time.perf_counter(); cde.detect_multi(ids, offs, host_t.array, out=host_out.array, wait=False); t1 = time.perf_counter(); cde.sync(); t2 = time.perf_counter()
        te += t1 - t0; tt += t2 - t0
    print(f"chunk {chunk} streams {streams}: enqueue {te/20*1e3:.3f} ms, total {tt/20*1e3:.3f} ms")
