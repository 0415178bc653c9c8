"""One collect batch per item on a workload, device-resident (for ncu)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import jagua_rs_b200 as jg
wl = jg.load_workload(sys.argv[1] if len(sys.argv) > 1 else "shirts_full")
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 20
cde = jg.build_engine(wl, device=0); cde.sync_layout()
words = cde.collector_words()
tot = 0.0
for item_id in range(len(wl.items)):
    t = jg.make_queries(wl, item_id, n, seed=0)
    d_t = jg.DeviceBuffer(cde, n * 16); d_t.upload(t)
    d_o = jg.DeviceBuffer(cde, n * 4 * words)
    cde.collect_batch_device(item_id, d_t.ptr, n, d_o.ptr, words); cde.sync()
    cde.timer_start(); cde.collect_batch_device(item_id, d_t.ptr, n, d_o.ptr, words); ms = cde.timer_stop()
    tot += ms
    print(f"item {item_id}: {ms*1e3:.1f} us for {n} queries ({n/ms/1e6:.2f} G q/s), {len(wl.items[item_id].verts)} vertices")
    d_t.free(); d_o.free()
print(f"total {tot:.3f} ms, {len(wl.items)*n/tot/1e6:.3f} G q/s")
