"""jg_shape_simplify / jg_item_import_simplified (product C++) against oracle/simplify.py (an independent float32
restatement of geometry/shape_modification.rs:48-330): same vertices, bit for bit, on the items and bins of the
instance files; plus the properties the reference asserts (shape_modification_valid, shape_modification.rs:504-531)."""
import glob
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
ASSETS = sorted(f for f in glob.glob(os.path.join(ROOT, "assets", "*.json")) if "config" not in os.path.basename(f))
MAX_ORACLE_VERTS = 70  # the checker is pure Python (O(n^2) per elimination round)


def shoelace(v):
    v = np.asarray(v, np.float64)
    return 0.5 * float(np.sum(v[:, 0] * np.roll(v[:, 1], -1) - np.roll(v[:, 0], -1) * v[:, 1]))


@pytest.mark.parametrize("path", ASSETS, ids=[os.path.basename(p) for p in ASSETS])
def test_items_simplify_like_the_oracle(jg, path):
    import simplify as osimp

    with open(path) as f:
        inst = json.load(f)
    n_checked = n_reduced = 0
    for it in inst["items"]:
        pts, _ = jg.instance.shape_points(it["shape"])
        base, _, _ = jg.import_item(pts)  # cleaned, counter-clockwise, centred: what simplify_shape receives
        for tol in (0.001, 0.02):
            got, poi = jg.simplify_shape(base.verts, "inflate", tol)
            a0, a1 = shoelace(base.verts), shoelace(got)
            assert len(got) <= len(base.verts) and len(got) >= 3
            assert a1 >= a0 * (1 - 1e-5) and abs(a1 - a0) / a0 <= tol * 1.001  # inflate: superset, within the tolerance
            assert poi[2] > 0
            n_reduced += int(len(got) < len(base.verts))
            if len(base.verts) <= MAX_ORACLE_VERTS:
                exp = osimp.simplify(base.verts, "inflate", tol)
                assert np.array_equal(got, exp), f"item {it['id']} tol {tol}: {len(got)} vs {len(exp)} vertices"
                n_checked += 1
    assert n_checked > 0 or all(len(jg.instance.shape_points(it["shape"])[0]) > MAX_ORACLE_VERTS for it in inst["items"])
    # the importer entry point = centring, simplification, SPolygon::new, surrogate on the simplified shape
    it = inst["items"][0]
    pts, _ = jg.instance.shape_points(it["shape"])
    base, pre0, area0 = jg.import_item(pts)
    simp, pre1, area1 = jg.import_item(pts, simplify_tolerance=0.02)
    v, poi = jg.simplify_shape(base.verts, "inflate", 0.02)
    assert np.array_equal(simp.verts, v) and np.array_equal(simp.poi, poi) and np.array_equal(pre0, pre1) and area0 == area1
    assert len(simp.poles) >= 2 and np.array_equal(simp.poles[0], simp.poi)


def test_containers_deflate_like_the_oracle(jg):
    import simplify as osimp

    with open(os.path.join(ROOT, "assets", "baldacci1.json")) as f:
        inst = json.load(f)
    n = 0
    for b in inst["bins"][:6]:
        outer, holes = jg.instance.shape_points(b["shape"])
        verts, _ = jg.import_shape(outer)
        for tol in (0.001, 0.01):
            got, _ = jg.simplify_shape(verts, "deflate", tol)
            a0, a1 = shoelace(verts), shoelace(got)
            assert a1 <= a0 * (1 + 1e-5) and abs(a1 - a0) / a0 <= tol * 1.001  # deflate: subset
            if len(verts) <= MAX_ORACLE_VERTS + 30:
                assert np.array_equal(got, osimp.simplify(verts, "deflate", tol))
                n += 1
        for h in holes[:3]:  # holes are inflated (io/import.rs:172-181)
            hv, _ = jg.import_shape(h)
            got, _ = jg.simplify_shape(hv, "inflate", 0.01)
            if len(hv) <= MAX_ORACLE_VERTS:
                assert np.array_equal(got, osimp.simplify(hv, "inflate", 0.01))
    assert n > 0


def test_simplify_rejects_bad_input(jg):
    with pytest.raises(jg.JaguaCudaError):
        jg.simplify_shape(np.array([[0, 0], [0, 1], [1, 1], [1, 0]], np.float32))  # clockwise: not an SPolygon vertex list
