"""jg_shape_offset (the repo's own offsetter standing in for geo-buffer in offset_shape, shape_modification.rs:330-372):
agreement with the float64 numpy version the config-5 fixtures were made with (tools/make_workloads.py), and the
properties minimum-separation needs - an inflated item contains every point within `distance` of the original (so two
inflated items that do not collide are at least 2 x distance apart), a deflated container lies inside the original."""
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


def items_of(jg, name):
    with open(os.path.join(ROOT, "assets", name)) as f:
        inst = json.load(f)
    return [jg.import_item(jg.instance.shape_points(it["shape"])[0])[0] for it in inst["items"]]


def seg_dist(p, a, b):
    ab, ap = b - a, p - a
    t = np.clip((ap @ ab) / max(ab @ ab, 1e-30), 0.0, 1.0)
    return float(np.linalg.norm(p - (a + t * ab)))


def dist_to_polygon(p, poly):
    return min(seg_dist(p, poly[i], poly[(i + 1) % len(poly)]) for i in range(len(poly)))


@pytest.mark.parametrize("name", ["jakobs1.json", "albano.json", "swim.json"])
def test_offset_matches_the_fixture_generator(jg, orc, name):
    from make_workloads import offset_polygon

    n = 0
    for sh in items_of(jg, name)[:10]:
        v = sh.verts.astype(np.float64)
        diam = float(np.sqrt(((v[:, None] - v[None]) ** 2).sum(-1).max()))
        d = 0.01 * diam
        try:
            exp = offset_polygon(sh.verts, d)
        except AssertionError:
            with pytest.raises(jg.JaguaCudaError):
                jg.offset_shape(sh.verts, "inflate", d)
            continue
        got, poi = jg.offset_shape(sh.verts, "inflate", d)
        # same corners up to the importer's cleaning (approx-equal neighbours dropped) and the start vertex
        assert len(got) <= len(exp) and poi[2] > 0
        for g in got:
            assert np.abs(exp.astype(np.float64) - g.astype(np.float64)).sum(1).min() <= 1e-5 * max(1.0, diam)
        n += 1
    assert n >= 3


@pytest.mark.parametrize("name", ["jakobs2.json", "shirts.json"])
def test_inflated_shape_contains_the_rounded_offset(jg, orc, name):
    rng = np.random.default_rng(0)
    lib = orc.lib()
    for sh in items_of(jg, name)[:6]:
        v = sh.verts.astype(np.float64)
        diam = float(np.sqrt(((v[:, None] - v[None]) ** 2).sum(-1).max()))
        d = 0.02 * diam
        try:
            got, _ = jg.offset_shape(sh.verts, "inflate", d)
        except jg.JaguaCudaError:
            continue
        g32 = np.ascontiguousarray(got, np.float32)
        g = got.astype(np.float64)
        # the original polygon's vertices and points at distance 0.999 d around them lie inside the inflated polygon
        for p in v:
            for ang in rng.uniform(0, 2 * np.pi, 6):
                q = p + 0.999 * d * np.array([np.cos(ang), np.sin(ang)])
                assert lib.orc_pip(g32, len(g32), float(q[0]), float(q[1])) == 1
        # and no boundary point of the inflated polygon is closer than d to the original boundary
        for i in range(len(g)):
            for t in (0.0, 0.5):
                q = g[i] + t * (g[(i + 1) % len(g)] - g[i])
                assert dist_to_polygon(q, v) >= d * (1 - 1e-4)


def test_deflated_container_lies_inside(jg, orc):
    lib = orc.lib()
    outer = np.array([[0, 0], [100, 0], [100, 60], [55, 60], [50, 30], [45, 60], [0, 60]], np.float32)
    got, poi = jg.offset_shape(outer, "deflate", 2.0)
    o32 = np.ascontiguousarray(outer)
    for p in got:
        assert lib.orc_pip(o32, len(o32), float(p[0]), float(p[1])) == 1
        assert dist_to_polygon(p.astype(np.float64), outer.astype(np.float64)) >= 2.0 * (1 - 1e-4)
    same, _ = jg.offset_shape(outer, "inflate", 0.0)
    assert np.array_equal(np.sort(same, axis=0), np.sort(outer, axis=0))
    with pytest.raises(jg.JaguaCudaError):
        jg.offset_shape(outer, "deflate", 40.0)   # the notch closes the strip: self-intersection
    with pytest.raises(jg.JaguaCudaError):
        jg.offset_shape(outer[::-1].copy(), "inflate", 1.0)  # clockwise input


def test_import_item_with_offset(jg):
    """original_shape.rs:23-51 order: centre, offset (Inflate), simplify; Item::area() stays the original shape's."""
    with open(os.path.join(ROOT, "assets", "jakobs1.json")) as f:
        inst = json.load(f)
    for it in inst["items"][:8]:
        pts, _ = jg.instance.shape_points(it["shape"])
        a, pre_a, area_a = jg.import_item(pts)
        b, pre_b, area_b = jg.import_item(pts, offset=0.25)
        assert np.array_equal(pre_a, pre_b) and area_a == area_b
        # the inscribed circle grows by about d (the pole of inaccessibility is a quadtree approximation, pole.rs:49-76)
        assert abs(float(b.poi[2]) - float(a.poi[2]) - 0.25) <= 0.1 * max(1.0, float(a.poi[2]))
        exp, _ = jg.offset_shape(a.verts, "inflate", 0.25)
        assert np.array_equal(b.verts, exp)
        c, _, _ = jg.import_item(pts, offset=0.25, simplify_tolerance=0.001, narrow_concavity_cutoff_ratio=0.01)
        assert 3 <= len(c.verts) <= len(b.verts) and len(c.poles) >= 2
