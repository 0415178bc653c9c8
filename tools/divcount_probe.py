"""Experiment (build with JG_EXTRA_NVCC=-DJG_DIVCOUNT): how many IEEE divisions the detect kernels of detect.cu
execute per site class on the headline batch (0: infinite / NaN divisor, 1: possible underflow to -0, 2: the half of
edge_intersection_half).

    JG_EXTRA_NVCC=-DJG_DIVCOUNT python jagua-rs_b200/build.py --force && python tools/divcount_probe.py
"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import jagua_rs_b200 as jg

wl = jg.load_workload(sys.argv[1] if len(sys.argv) > 1 else "swim_half")
cde = jg.build_engine(wl, device=0)
cde.sync_layout()
n = 1 << 20
ids = list(range(len(wl.items)))
sizes = jg.split_batch(n, len(wl.items))
offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
t = np.concatenate([jg.make_queries(wl, i, m, seed=0) for i, m in enumerate(sizes)])
dt_ = jg.DeviceBuffer(cde, n * 16)
dt_.upload(t)
do = jg.DeviceBuffer(cde, n)
out = (C.c_ulonglong * 4)()
cde._lib.jg_debug_divcount(out, 1)
cde.detect_multi_device(ids, offs, dt_.ptr, do.ptr)
cde.sync()
cde._lib.jg_debug_divcount(out, 1)
print("divisions per 1M-query step by site class:", list(out), "survivors:", cde.stats()["survivors"])
