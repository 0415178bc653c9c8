"""Layout-update latency (deregister + register one placed item, re-flatten, upload, rebuild the per-filter
summaries) and collect-mode throughput on the fixtures of BASELINE.json configs 2-3.

    python tools/update_probe.py
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import jagua_rs_b200 as jg

for name in ("swim_half", "shirts_full", "trousers_full"):
    wl = jg.load_workload(name)
    cde = jg.build_engine(wl, device=0)
    cde.sync_layout()
    placed = [(k, h) for k, h in zip(cde.hazard_keys, wl.hazards) if h.kind == 0]
    t_host = t_sync = 0.0
    reps = 50
    for r in range(reps):
        k, h = placed[r % len(placed)]
        t0 = time.perf_counter()
        cde.deregister_hazard_by_key(k)
        nk = cde.register_hazard(h.kind, h.verts, h.poi, id=h.id, quality=h.quality, pk=h.pk, dynamic=h.dynamic)
        t1 = time.perf_counter()
        cde.sync_layout()
        cde.sync()
        t2 = time.perf_counter()
        placed[r % len(placed)] = (nk, h)
        t_host += t1 - t0
        t_sync += t2 - t1
    print(f"{name}: remove+place on the host mirror {t_host / reps * 1e6:.0f} us, flatten+upload+summaries {t_sync / reps * 1e6:.0f} us")
    # the same update issued while a 1M-query host batch of the same engine is in flight: time until sync_layout
    # returns (stream-ordered upload: the host does not wait for the batch; JG_SYNC_UPLOAD=1 restores the blocking form)
    nq = 1 << 20
    sizes = jg.split_batch(nq, len(wl.items))
    offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    ht = jg.PinnedBuffer((nq, 4), np.float32)
    for i, m in enumerate(sizes):
        ht.array[offs[i]:offs[i + 1]] = jg.make_queries(wl, i, m, seed=0)
    ho = jg.PinnedBuffer((nq,), np.uint8)
    ids = list(range(len(wl.items)))
    t_upd = t_all = 0.0
    for r in range(20):
        k, h = placed[r % len(placed)]
        cde.detect_multi(ids, offs, ht.array, out=ho.array, wait=False)
        t0 = time.perf_counter()
        cde.deregister_hazard_by_key(k)
        nk = cde.register_hazard(h.kind, h.verts, h.poi, id=h.id, quality=h.quality, pk=h.pk, dynamic=h.dynamic)
        cde.sync_layout()
        t1 = time.perf_counter()
        cde.detect_multi(ids, offs, ht.array, out=ho.array, wait=False)
        cde.sync()
        t2 = time.perf_counter()
        placed[r % len(placed)] = (nk, h)
        t_upd += t1 - t0
        t_all += t2 - t0
    print(f"{name}: update behind an in-flight 1M batch: host returns after {t_upd / 20 * 1e6:.0f} us; update + next 1M batch done after {t_all / 20 * 1e6:.0f} us")
    ht.free(), ho.free()
    # collect throughput, device-resident
    n = 1 << 18
    words = cde.collector_words()
    tot_ms, tot_q = 0.0, 0
    for item_id in range(len(wl.items)):
        t = jg.make_queries(wl, item_id, n, seed=0)
        d_t = jg.DeviceBuffer(cde, n * 16)
        d_t.upload(t)
        d_o = jg.DeviceBuffer(cde, n * 4 * words)
        for _ in range(2):
            cde.collect_batch_device(item_id, d_t.ptr, n, d_o.ptr, words)
        cde.sync()
        cde.timer_start()
        for _ in range(3):
            cde.collect_batch_device(item_id, d_t.ptr, n, d_o.ptr, words)
        tot_ms += cde.timer_stop() / 3
        tot_q += n
        d_t.free(), d_o.free()
    print(f"{name}: collect (surrogate + poly), {len(wl.items)} items x {n} queries: {tot_q / tot_ms / 1e6:.3f} G q/s device-resident, {words} words/query")
    cde.close()
