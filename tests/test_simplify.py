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


def _comb(n_teeth=5, gap=0.02, depth=1.0, width=1.0):
    """a comb: n_teeth teeth separated by narrow slots (concavities `gap` wide, `depth` deep), counter-clockwise"""
    pts = [(0.0, 0.0)]
    x = 0.0
    for k in range(n_teeth):
        pts += [(x + width, 0.0)] if k == n_teeth - 1 else [(x + width, 0.0), (x + width, depth * 0.9), (x + width + gap, depth * 0.9), (x + width + gap, 0.0)]
        x += width + gap
    pts = [(0.0, -1.0), (x, -1.0)] + [(px, py) for px, py in reversed([(p[0], -p[1]) for p in pts[1:]])]
    v = np.array(pts, np.float32)
    if np.sum(v[:, 0] * np.roll(v[:, 1], -1) - np.roll(v[:, 0], -1) * v[:, 1]) < 0:
        v = v[::-1].copy()
    return v


def test_close_narrow_concavities_like_the_oracle(jg):
    import simplify as osimp

    comb = _comb()
    base, _ = jg.import_shape(comb)
    got, _ = jg.close_narrow_concavities(base, "inflate", 0.05)
    exp = osimp.close_narrow_concavities(base, "inflate", 0.05)
    assert np.array_equal(got, exp)
    assert len(got) < len(base) and shoelace(got) > shoelace(base)  # slots closed: fewer vertices, larger area
    same, _ = jg.close_narrow_concavities(base, "inflate", 0.001)      # mouths wider than the cutoff: untouched
    assert np.array_equal(same, base) and np.array_equal(osimp.close_narrow_concavities(base, "inflate", 0.001), base)
    # items of the instance files with the lbf default (narrow_concavity_cutoff_ratio 0.01)
    n = 0
    for name in ("swim.json", "trousers.json", "jakobs1.json"):
        with open(os.path.join(ROOT, "assets", name)) as f:
            inst = json.load(f)
        for it in inst["items"]:
            pts, _ = jg.instance.shape_points(it["shape"])
            b, _, _ = jg.import_item(pts)
            if len(b.verts) > 24:
                continue
            g, _ = jg.close_narrow_concavities(b.verts, "inflate", 0.01)
            assert np.array_equal(g, osimp.close_narrow_concavities(b.verts, "inflate", 0.01))
            n += 1
    assert n >= 10


def test_default_lbf_preprocessing_pipeline(jg):
    """original_shape.rs:23-51 with the lbf defaults: simplify(0.001) -> close_narrow_concavities(0.01) -> simplify(0.0001),
    each step through SPolygon::new; the importer entry point must equal the three calls chained by hand."""
    import simplify as osimp

    with open(os.path.join(ROOT, "assets", "swim.json")) as f:
        inst = json.load(f)
    for it in inst["items"]:
        pts, _ = jg.instance.shape_points(it["shape"])
        base, pre, area = jg.import_item(pts)
        full, pre2, area2 = jg.import_item(pts, simplify_tolerance=0.001, narrow_concavity_cutoff_ratio=0.01)
        v1, _ = jg.simplify_shape(base.verts, "inflate", 0.001)
        v2, _ = jg.close_narrow_concavities(v1, "inflate", 0.01)
        v3, poi3 = jg.simplify_shape(v2, "inflate", np.float32(0.001) / np.float32(10.0))
        assert np.array_equal(full.verts, v3) and np.array_equal(full.poi, poi3)
        assert np.array_equal(pre, pre2) and area == area2 and len(full.verts) <= len(base.verts)
        if len(base.verts) <= 40:
            o = osimp.simplify(osimp.close_narrow_concavities(osimp.simplify(base.verts, "inflate", 0.001), "inflate", 0.01),
                               "inflate", np.float32(0.001) / np.float32(10.0))
            assert np.array_equal(full.verts, o)
