"""Feasibility probe: the detect kernels reading the queries straight from pinned host memory (UVA pointer
passed to the *_device entry point) vs from device memory. ncu-friendly (few launches)."""
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
ids = list(range(len(wl.items)))
dev_t = jg.DeviceBuffer(cde, n * 16)
dev_t.upload(host_t.array)
dev_out = jg.DeviceBuffer(cde, n)
dev_out2 = jg.DeviceBuffer(cde, n)
for name, ptr, out in (("device", dev_t.ptr, dev_out), ("pinned host (zero-copy)", host_t.ptr, dev_out2)):
    for _ in range(3):
        cde.detect_multi_device(ids, offs, ptr, out.ptr)
    cde.sync()
    cde.timer_start()
    for _ in range(10):
        cde.detect_multi_device(ids, offs, ptr, out.ptr)
    ms = cde.timer_stop() / 10
    print(f"{name}: {ms * 1e3:.1f} us per 1M-query batch")
a, b = dev_out.download((n,), np.uint8), dev_out2.download((n,), np.uint8)
print("same verdicts:", np.array_equal(a, b))
