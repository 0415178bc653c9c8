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
REF_ASSETS = "/root/reference/assets"  # build container only: larger instance files are read where they lie


def asset_path(asset):
    p = os.path.join(ASSETS, asset)
    return p if os.path.exists(p) else os.path.join(REF_ASSETS, asset)


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
    inst = orc.load_instance_json(asset_path(asset))
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
    inst = orc.load_instance_json(asset_path(asset))
    w = orc.world_from_instance(inst, bin_index=bin_index, **cfg)
    t = time.time()
    n = w.lbf_fill_bin(n_samples, 0.2, seed, max_items)
    assert w.check_invariants() == 0 and w.matches_fresh() == 0
    print(f"{name}: placed {n}, hazards {len(w.hazards())}, lbf {time.time() - t:.1f}s")
    return snapshot(name, w)


def offset_polygon(pts, d, miter_limit=2.0):
    """Own polygon offsetter (NOT geo-buffer, which the reference calls in shape_modification.rs:348; results
    are not bit-equal to it - the fixtures record the offset polygons, so oracle and GPU see the same inputs).
    Every edge of the CCW polygon moves outward by d (inward for d < 0); consecutive offset edges meet at their
    intersection (miter); corners sharper than miter_limit * |d| are bevelled."""
    p = np.asarray(pts, np.float64).reshape(-1, 2)
    if len(p) > 1 and np.allclose(p[0], p[-1]):
        p = p[:-1]
    a2 = np.sum(p[:, 0] * np.roll(p[:, 1], -1) - np.roll(p[:, 0], -1) * p[:, 1])
    if a2 < 0:
        p = p[::-1]
    n = len(p)
    out = []
    for i in range(n):
        a, b, c = p[i - 1], p[i], p[(i + 1) % n]
        e0, e1 = b - a, c - b
        e0, e1 = e0 / np.linalg.norm(e0), e1 / np.linalg.norm(e1)
        n0, n1 = np.array([e0[1], -e0[0]]), np.array([e1[1], -e1[0]])  # outward normals of a CCW polygon
        denom = 1.0 + float(n0 @ n1)
        if denom < 1e-9:
            out += [b + d * n0, b + d * n1]
            continue
        m = (n0 + n1) / denom  # b + d*m is the intersection of the two offset lines
        if np.linalg.norm(m) > miter_limit:
            out += [b + d * (n0 + e0 * np.tan(0.25 * np.arccos(np.clip(n0 @ n1, -1, 1)))),
                    b + d * (n1 - e1 * np.tan(0.25 * np.arccos(np.clip(n0 @ n1, -1, 1))))]
        else:
            out.append(b + d * m)
    q = np.array(out)

    def simple(poly):
        m = len(poly)

        def ccw(a, b, c):
            return (c[1] - a[1]) * (b[0] - a[0]) - (b[1] - a[1]) * (c[0] - a[0])
        for i in range(m):
            for j in range(i + 2, m):
                if i == 0 and j == m - 1:
                    continue
                a, b, c, dd = poly[i], poly[(i + 1) % m], poly[j], poly[(j + 1) % m]
                if ccw(a, b, c) * ccw(a, b, dd) < 0 and ccw(c, dd, a) * ccw(c, dd, b) < 0:
                    return False
        return True
    assert simple(q), "offset polygon self-intersects: reduce the separation"
    return q.astype(np.float32)


def mspp(prefix, asset, sep_frac=0.01, width_frac=0.75, n_samples=1000, seed=0, max_strips=6, **cfg):
    """BASELINE.json config 5 inputs: multi-strip packing with a minimum separation distance. No MSPP asset or
    optimizer exists in the reference tree (probs/mspp holds entities + io only), so strips are synthesised:
    ExtMSPStrip{height = strip_height, max_width = width_frac x the 100 %-density width}; items are inflated by
    sep / 2 and every strip deflated by sep / 2 (what original_shape.rs:23-51 does with the configured
    min_item_separation) by the offsetter above; LBF fills strip after strip with the demand that is left."""
    inst = orc.load_instance_json(asset_path(asset))
    items = [it for it in sorted(inst["items"], key=lambda it: it["id"]) if it.get("demand", 1) > 0]
    H = float(inst["strip_height"])
    sep = sep_frac * H
    probe = orc.world_from_instance(inst, **cfg)
    w100 = probe.strip_width()
    max_w = width_frac * w100
    shapes = [offset_polygon(orc._shape_points(it["shape"])[0], sep / 2) for it in items]
    left = [int(it.get("demand", 1)) for it in items]
    d = sep / 2
    rect = np.array([[d, d], [max_w - d, d], [max_w - d, H - d], [d, H - d]], np.float32)
    out = []
    for k in range(max_strips):
        if sum(left) == 0:
            break
        w = orc.OracleWorld(**cfg)
        for it, pts, dem in zip(items, shapes, left):
            w.add_item(pts, it.get("allowed_orientations"), it.get("min_quality"), dem)
        w.set_container(rect)
        n = w.lbf_fill_bin(n_samples, 0.2, seed + k)
        _, ids, _ = w.placed()
        for i in ids:
            left[int(i)] -= 1
        assert n > 0 and w.check_invariants() == 0 and w.matches_fresh() == 0
        print(f"{prefix}_ms{k}: strip {max_w:.1f} x {H:.1f}, separation {sep:.3f}, placed {n}, demand left {sum(left)}")
        out.append(snapshot(f"{prefix}_ms{k}", w))
    assert sum(left) == 0, "demand left after max_strips"
    return out


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
    # config 5: multi-strip + minimum separation (synthesised strips, own offsetter); one fixture per strip
    "jakobs1_ms": lambda: mspp("jakobs1", "jakobs1.json"),
    "jakobs2_ms": lambda: mspp("jakobs2", "jakobs2.json"),
    "albano_ms": lambda: mspp("albano", "albano.json"),
    # config 4: second bin with holes / zones (1884-vertex outer boundary)
    "baldacci3_bin": lambda: bpp("baldacci3_bin", "baldacci3.json", bin_index=0),
    "gardeyn4c_half": lambda: spp("gardeyn4c_half", "gardeyn4_c.json", keep="half", n_samples=300),
    # continuous rotation, many vertices
    "gardeyn1c_half": lambda: spp("gardeyn1c_half", "gardeyn1_c.json", keep="half", n_samples=500),
}


def main():
    os.makedirs(WORKLOAD_DIR, exist_ok=True)
    names = sys.argv[1:] or list(RECIPES)
    for n in names:
        wl = RECIPES[n]()
        for one in (wl if isinstance(wl, list) else [wl]):
            one.save(os.path.join(WORKLOAD_DIR, one.name + ".npz"))


if __name__ == "__main__":
    main()
