"""Writes tests/golden/<workload>.npz: seeded queries with the ORACLE's verdicts and hazard sets.

These are regression pins of the oracle (and the fixed answers the GPU path must reproduce), NOT
reference outputs: the Rust reference cannot run in this image. tools/ref_dump.rs replays the same
files against the real crate on a machine that has cargo.

    python tools/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import orc  # noqa: E402
import jagua_rs_b200 as jg  # noqa: E402
from helpers import ALL_WORKLOADS, oracle_world  # noqa: E402

N_PER_ITEM = 256
SEED = 11


def main():
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    for name in ALL_WORKLOADS:
        if os.path.exists(os.path.join(out_dir, name + ".npz")) and "--all" not in sys.argv:
            continue  # existing pins stay byte-identical; --all regenerates everything
        wl = jg.load_workload(name)
        w = oracle_world(orc, wl)
        words = max(1, (w.haz_capacity() + 31) // 32)
        ts, det, det_poly, col, filt_det, filt_bits = [], [], [], [], [], []
        for i, it in enumerate(wl.items):
            t = jg.make_queries(wl, i, N_PER_ITEM, SEED)
            ts.append(t)
            det.append(w.detect_batch(i, t))
            det_poly.append(w.detect_batch(i, t, poly_only=True))
            col.append(w.collect_batch(i, t, words))
            # HazKeyFilter::from_irrelevant_qzones for the item's min_quality (lbf_spp.rs:54-57); items
            # without a min_quality get a filter that hides the first dynamic hazard (self-filter shape
            # of Layout::is_feasible, layout.rs:128-136)
            if it.min_quality >= 0:
                fb = w.filter_qzones(it.min_quality, words)
            else:
                fb = np.zeros(words, np.uint32)
                dyn = [k for k, h in enumerate(wl.hazards) if h.dynamic]
                if dyn:
                    s = dyn[i % len(dyn)]
                    fb[s >> 5] |= np.uint32(1 << (s & 31))
            filt_bits.append(fb)
            filt_det.append(w.detect_batch(i, t, filter_bits=fb))
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), transforms=np.stack(ts), detect=np.stack(det),
                            detect_poly_only=np.stack(det_poly), collect=np.stack(col), filter_bits=np.stack(filt_bits),
                            detect_filtered=np.stack(filt_det), seed=np.array(SEED))
        d = np.stack(det)
        print(name, d.shape, "surrogate", int((d == 1).sum()), "poly", int((d == 2).sum()), "free", int((d == 0).sum()),
              "panic(filtered)", int((np.stack(filt_det) == 0x80).sum()))


if __name__ == "__main__":
    main()
