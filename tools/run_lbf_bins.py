"""LBF bin packing of a BPP instance (bins with holes and quality zones, items with min_quality) on the engine
(jagua-rs_b200/lbf.py::BatchedLBFBins, after lbf/src/opt/lbf_bpp.rs), with an oracle feasibility check of every layout.

    python tools/run_lbf_bins.py [assets/baldacci1.json] [--samples 20000] [--no-check]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import jagua_rs_b200 as jg  # noqa: E402
from jagua_rs_b200 import lbf  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("instance", nargs="?", default=os.path.join(ROOT, "assets", "baldacci1.json"))
ap.add_argument("--samples", type=int, default=20000)
ap.add_argument("--seed", type=int, default=0)
ap.add_argument("--no-check", action="store_true")
ap.add_argument("--rotation-lattice", type=int, default=256, help="angles of the device-side lattice for items that rotate continuously (0: sample them on the host)")
args = ap.parse_args()
with open(args.instance) as f:
    inst = json.load(f)
solver = lbf.BatchedLBFBins(inst, device=0, seed=args.seed, rotation_lattice=args.rotation_lattice)
res = solver.solve(n_samples=args.samples)
line = {"instance": os.path.basename(args.instance), "n_samples_per_search": args.samples, "placed": res.placed, "unplaced": res.unplaced,
        "bins_used": [lay.bin_id for lay in res.layouts], "items_per_bin": [len(lay.placements) for lay in res.layouts],
        "seconds": res.seconds, "samples_evaluated": res.samples, "samples_per_second": res.samples / res.seconds}
if not args.no_check:
    for p in (os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
        sys.path.insert(0, p)
    import orc
    orc.build()
    ok = []
    for lay in res.layouts:
        w = orc.world_from_instance(inst, bin_index=lay.bin_id)
        for p in lay.placements:
            w.place_item(p.item_id, float(p.d_transf[0]), float(p.d_transf[1]), float(p.d_transf[2]))
        ok.append(bool(w.is_feasible()))
    line["oracle_says_feasible"] = ok
print(json.dumps(line))
solver.close()
