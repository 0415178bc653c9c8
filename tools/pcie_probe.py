"""Measures pinned host<->device copy rates through the C ABI's memcpy helpers (the transfer ceiling of
the e2e path: 16 B up + 1 B down per detect query).

    python tools/pcie_probe.py
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import jagua_rs_b200 as jg

wl = jg.load_workload("swim_half")
cde = jg.build_engine(wl, device=0)
for mb in (1, 4, 16, 64):
    n = mb << 20
    h = jg.PinnedBuffer((n,), np.uint8)
    d = jg.DeviceBuffer(cde, n)
    for _ in range(3):
        d.upload(h.array)
    t0 = time.perf_counter()
    reps = 20
    for _ in range(reps):
        d.upload(h.array)
    dt = (time.perf_counter() - t0) / reps
    t0 = time.perf_counter()
    for _ in range(reps):
        d.download((n,), np.uint8)
    dt2 = (time.perf_counter() - t0) / reps
    print(f"{mb:3d} MiB  H2D {dt * 1e6:8.1f} us  {n / dt / 1e9:6.1f} GB/s   D2H(+alloc) {dt2 * 1e6:8.1f} us")
