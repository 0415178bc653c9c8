"""H2D bandwidth from ordinary pinned memory vs write-combined pinned memory, one GPU or all GPUs of the box at once
(one engine + one host thread per GPU): is the aggregate host-to-device ceiling of the 8-GPU boxes (~165 GB/s measured
in round 1) a snooping limit that write-combined buffers avoid?

    python tools/wc_probe.py [n_gpus]
"""
import ctypes as C
import os
import sys
import threading
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import jagua_rs_b200 as jg

n_gpus = int(sys.argv[1]) if len(sys.argv) > 1 else 1
MB = 64
lib = jg.load_library()
cfg = jg.JgConfig(5, 16, 2, 0)
engs = []
for d in range(n_gpus):
    e = C.c_void_p()
    assert lib.jg_engine_create(d, C.byref(cfg), C.byref(e)) == 0
    engs.append(e)


def run(wc, devs, reps=20, d2h=False):
    bufs, dptrs = [], []
    for d in devs:
        b = jg.PinnedBuffer((MB << 20,), np.uint8, write_combined=wc)
        b.array[:] = 1
        p = C.c_void_p()
        assert lib.jg_device_alloc(engs[d], MB << 20, C.byref(p)) == 0
        bufs.append(b), dptrs.append(p)
    times = [0.0] * len(devs)

    def work(k):
        d = devs[k]
        fn = lib.jg_memcpy_d2h if d2h else lib.jg_memcpy_h2d
        a, b = (bufs[k].ptr, dptrs[k]) if d2h else (dptrs[k], bufs[k].ptr)
        for _ in range(3):
            fn(engs[d], a, b, MB << 20)
        bar.wait()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn(engs[d], a, b, MB << 20)
        times[k] = time.perf_counter() - t0

    bar = threading.Barrier(len(devs))
    th = [threading.Thread(target=work, args=(k,)) for k in range(len(devs))]
    for t in th:
        t.start()
    for t in th:
        t.join()
    for d, p in zip(devs, dptrs):
        lib.jg_device_free(engs[d], p)
    agg = len(devs) * reps * MB / 1024 / max(times)
    return agg


for devs in ([0], list(range(min(2, n_gpus))), list(range(min(4, n_gpus))), list(range(n_gpus))):
    if len(devs) > 1 and len(devs) > n_gpus:
        continue
    for wc in (False, True):
        print(f"devices {devs}  {'write-combined' if wc else 'pinned        '}  H2D {run(wc, devs):7.1f} GB/s aggregate" +
              ("" if wc else f"   D2H {run(False, devs, d2h=True):7.1f} GB/s"), flush=True)
    if len(devs) == n_gpus:
        break
