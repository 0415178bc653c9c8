"""Exports a self-contained parity case for tools/ref_dump.rs (the dumper that runs the REAL jagua-rs crate on a machine
that has cargo): instance path, CDE configuration, strip width, the placements of an LBF layout in placement order,
which of them are removed again, seeded queries per item as (rotation, tx, ty), and what the CPU oracle answers -
verdict bytes (bit0 surrogate, bit1 poly) and colliding hazard sets as indices into the placement list (-1 = Exterior).

    python tools/export_ref_case.py [swim.json] [--samples 5000] [--queries 2000] [--out tests/golden/ref_case_swim.json]

Floats are written as the shortest decimal that round-trips the f32 value.
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import orc  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("instance", nargs="?", default="swim.json")
ap.add_argument("--samples", type=int, default=5000)
ap.add_argument("--queries", type=int, default=2000)
ap.add_argument("--out", default=None)
args = ap.parse_args()


def f32s(a):
    return [float(np.format_float_positional(np.float32(x), unique=True, trim="0")) for x in np.ravel(a)]


inst = orc.load_instance_json(os.path.join(ROOT, "assets", args.instance))
w = orc.world_from_instance(inst)
w.lbf_solve_spp(args.samples, 0.2, 0)
pks, ids, dts = w.placed()
placements = [dict(item_id=int(i), d_transf=f32s(dt)) for i, dt in zip(ids, dts)]
removed = list(range(1, len(pks), 2))
for k in removed:
    w.remove_item(int(pks[k]))
# hazard slot -> placement index (or -1 for the exterior)
slot_of = {}
for h in w.hazards():
    if h["kind"] == 1:
        slot_of[h["slot"]] = -1
    elif h["kind"] == 0:
        slot_of[h["slot"]] = int(np.nonzero(pks == np.uint64(h["pk"]))[0][0])
words = max(1, (w.haz_capacity() + 31) // 32)
b = w.cde_bbox()
queries = []
for i in range(w.n_items()):
    it = w.item(i)
    rng = np.random.Generator(np.random.Philox(key=[17, i]))
    u = rng.random((args.queries, 2), dtype=np.float32)
    tx = (b[0] + (b[2] - b[0]) * u[:, 0]).astype(np.float32)
    ty = (b[1] + (b[3] - b[1]) * u[:, 1]).astype(np.float32)
    rots = it["rotations"]
    rot = (np.zeros(args.queries, np.float32) if rots is not None and len(rots) == 0 else
           (rng.random(args.queries, dtype=np.float32) * np.float32(2 * np.pi)).astype(np.float32) if rots is None else
           np.asarray(rots, np.float32)[rng.integers(0, len(rots), args.queries)])
    dt = np.stack([rot, tx, ty], 1).astype(np.float32)
    t4 = orc.compose(dt)
    det = w.detect_batch(i, t4)
    col = w.collect_batch(i, t4, words)
    sets = []
    for row in col:
        s = sorted(slot_of[32 * k + bit] for k in range(words) for bit in range(32) if (int(row[k]) >> bit) & 1)
        sets.append(s)
    queries.append(dict(item_id=i, d_transfs=[f32s(r) for r in dt], expect_detect=[int(v) for v in det], expect_collect=sets))
case = dict(instance=os.path.join("assets", args.instance), strip_width=f32s([w.strip_width()])[0],
            cde=dict(quadtree_depth=5, cd_threshold=16, n_ff_poles=2, n_ff_piers=0,
                     n_pole_limits=[[100, 0.0], [20, 0.75], [10, 0.90]]),
            shape_modification="none (poly_simpl_tolerance = None, min_item_separation = None)",
            placements=placements, removed=removed, queries=queries,
            produced_by="oracle/ (CPU restatement), tools/export_ref_case.py")
out = args.out or os.path.join(ROOT, "tests", "golden", "ref_case_" + os.path.splitext(args.instance)[0] + ".json")
with open(out, "w") as f:
    json.dump(case, f, separators=(",", ":"))
print(out, os.path.getsize(out), "bytes;", len(placements), "placements,", len(removed), "removed,", sum(len(q["d_transfs"]) for q in queries), "queries")
