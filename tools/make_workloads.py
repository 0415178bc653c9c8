"""Generates the workload fixtures under workloads/ (committed) from the instance files in assets/.

Runs once in the build container. It uses the CPU oracle's importer and LBF filler to *produce inputs*
(item shapes with poles, a filled layout); nothing here is on the product path, and the fixtures are
plain data: CDE bbox, hazards (kind, ids, vertices, poi) in slot order, items (vertices, poi, poles,
allowed rotations). bench.py and the tests rebuild both engines (GPU product, CPU oracle) from them.

    python tools/make_workloads.py [name ...]
"""
from __future__ import annotations

import importlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import orc  # noqa: E402

jg = importlib.import_module("jagua_rs_b200")
from jagua_rs_b200 import Workload  # noqa: E402
from jagua_rs_b200.workload import WHazard, WItem, WORKLOAD_DIR  # noqa: E402

ASSETS = os.path.join(ROOT, "assets")


def snapshot(name: str, w: orc.OracleWorld) -> Workload:
    cfg = w.config
    wl = Workload(name, cfg["quadtree_depth"], cfg["cd_threshold"], cfg["n_ff_poles"], cfg["n_ff_piers"], w.cde_bbox())
    for i in range(w.n_items()):
        it = w.item(i)
        wl.items.append(WItem(it["verts"], it["poi"], it["poles"], it["piers"], it["rotations"], it["min_quality"]))
    for h in w.hazards():  # slot order; static hazards first
        wl.hazards.append(WHazard(h["kind"], h["id"], h["quality"], h["pk"], h["dynamic"], h["verts"], h["poi"]))
    return wl


def spp(name, asset, keep="all", n_samples=5000, seed=0, **cfg):
    inst = orc.load_instance_json(os.path.join(ASSETS, asset))
    w = orc.world_from_instance(inst, **cfg)
    t = time.time()
    n = w.lbf_solve_spp(n_samples, 0.2, seed)
    pks, ids, _ = w.placed()
    if keep == "half":  # lbf/benches/util.rs:27-79 removes items from a full LBF layout
        for k in range(1, len(pks), 2):
            w.remove_item(int(pks[k]))
    assert w.is_feasible() and w.check_invariants() == 0 and w.matches_fresh() == 0
    print(f"{name}: placed {n} (kept {w.hazards().__len__() - 1}), strip width {w.strip_width():.1f}, lbf {time.time() - t:.1f}s")
    return snapshot(name, w)


def bpp(name, asset, bin_index=0, n_samples=1000, seed=0, max_items=40, **cfg):
    inst = orc.load_instance_json(os.path.join(ASSETS, asset))
    w = orc.world_from_instance(inst, bin_index=bin_index, **cfg)
    t = time.time()
    n = w.lbf_fill_bin(n_samples, 0.2, seed, max_items)
    assert w.check_invariants() == 0 and w.matches_fresh() == 0
    print(f"{name}: placed {n}, hazards {len(w.hazards())}, lbf {time.time() - t:.1f}s")
    return snapshot(name, w)


RECIPES = {
    # config 2 (headline): half-filled swim layout, LBFConfig defaults (assets/config_lbf.json)
    "swim_half": lambda: spp("swim_half", "swim.json", keep="half"),
    # config 3: full shirts / trousers layouts
    "shirts_full": lambda: spp("shirts_full", "shirts.json"),
    "trousers_full": lambda: spp("trousers_full", "trousers.json"),
    # config 4: bin with holes + inferior-quality zones, items with min_quality (filters)
    "baldacci1_bin": lambda: bpp("baldacci1_bin", "baldacci1.json", bin_index=0),
    # config 5 sources / small parity cases
    "albano_half": lambda: spp("albano_half", "albano.json", keep="half", n_samples=1000),
    "jakobs1_full": lambda: spp("jakobs1_full", "jakobs1.json", n_samples=1000),
    # other quadtree depths the reference's integration tests run (lbf/tests/tests.rs:18)
    "swim_half_d0": lambda: spp("swim_half_d0", "swim.json", keep="half", n_samples=500, quadtree_depth=0),
    "swim_half_d3": lambda: spp("swim_half_d3", "swim.json", keep="half", n_samples=1000, quadtree_depth=3),
    "shirts_full_d10": lambda: spp("shirts_full_d10", "shirts.json", n_samples=1000, quadtree_depth=10),
    # continuous rotation, many vertices
    "gardeyn1c_half": lambda: spp("gardeyn1c_half", "gardeyn1_c.json", keep="half", n_samples=500),
}


def main():
    os.makedirs(WORKLOAD_DIR, exist_ok=True)
    names = sys.argv[1:] or list(RECIPES)
    for n in names:
        wl = RECIPES[n]()
        wl.save(os.path.join(WORKLOAD_DIR, n + ".npz"))


if __name__ == "__main__":
    main()
