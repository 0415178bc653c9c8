"""BASELINE.json config 1 on the GPU engine: LBF strip packing of an instance file with the batched search of
jagua-rs_b200/lbf.py, next to the CPU restatement of the reference's sequential LBF (oracle/, same sample budget) and
an oracle feasibility check of the GPU solution (every placed item re-placed in an oracle layout, Layout::is_feasible).

    python tools/run_lbf.py [assets/swim.json] [--samples 5000] [--no-check] [--solution out.json]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import jagua_rs_b200 as jg  # noqa: E402
from jagua_rs_b200 import lbf  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("instance", nargs="?", default=os.path.join(ROOT, "assets", "swim.json"))
ap.add_argument("--samples", type=int, default=5000)
ap.add_argument("--ls-frac", type=float, default=0.2)
ap.add_argument("--seed", type=int, default=0)
ap.add_argument("--no-check", action="store_true")
ap.add_argument("--solution", default=None, help="write the solution as the reference's ExtSPSolution JSON (probs/spp/io/ext_repr.rs)")
ap.add_argument("--shape-defaults", action="store_true", help="the lbf defaults for item shapes: poly_simpl_tolerance 0.001, narrow_concavity_cutoff_ratio 0.01")
ap.add_argument("--rotation-lattice", type=int, default=256, help="angles of the device-side lattice for items that rotate continuously (0: sample them on the host)")
ap.add_argument("--host-sampling", action="store_true", help="draw the samples with numpy on the host instead of on the device")
args = ap.parse_args()

with open(args.instance) as f:
    inst = json.load(f)
solver = lbf.BatchedLBF(inst, device=0, seed=args.seed, device_sampling=not args.host_sampling, rotation_lattice=args.rotation_lattice,
                        simplify_tolerance=0.001 if args.shape_defaults else None,
                        narrow_concavity_cutoff_ratio=0.01 if args.shape_defaults else None)
res = solver.solve(n_samples=args.samples, ls_frac=args.ls_frac)
stats = solver.cde.stats()
line = {"instance": os.path.basename(args.instance), "n_samples_per_placement": args.samples, "placed": len(res.placements),
        "item_vertices": int(sum(len(it.shape.verts) for it in solver.items)), "strip_width": res.strip_width, "strip_height": res.strip_height, "density": res.density, "seconds": res.seconds,
        "samples_evaluated": res.samples, "strip_extensions": res.strip_extensions,
        "samples_per_second": res.samples / res.seconds,
        "sampling": "host (numpy Philox)" if args.host_sampling else "device (k_search_sample) for items with discrete rotations"}
if not args.no_check:
    for p in (os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
        sys.path.insert(0, p)
    import orc
    orc.build()
    w = orc.world_from_instance(inst)
    w.change_strip_width(res.strip_width)
    for p in res.placements:
        w.place_item(p.item_id, float(p.d_transf[0]), float(p.d_transf[1]), float(p.d_transf[2]))
    line["oracle_says_feasible"] = bool(w.is_feasible())
    w2 = orc.world_from_instance(inst)
    t0 = time.perf_counter()
    n = w2.lbf_solve_spp(args.samples, args.ls_frac, args.seed)
    dt = time.perf_counter() - t0
    line["cpu_lbf"] = {"placed": int(n), "strip_width": w2.strip_width(), "seconds": dt, "kind": "port (sequential search.rs restatement, 1 thread)",
                       "density": solver.total_area / (w2.strip_width() * float(res.strip_height))}
if args.solution:
    with open(args.solution, "w") as f:
        json.dump(jg.export.export_spp_solution(res, solver.items), f, indent=1)
    line["solution_file"] = args.solution
print(json.dumps(line))
solver.close()
