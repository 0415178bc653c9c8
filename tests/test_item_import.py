"""jg_item_import (product host C++: importer + pole of inaccessibility + surrogate poles) vs the oracle's
restatement of io/import.rs + fail_fast/pole.rs on every item of every instance file in assets/: cleaned and
centred vertices, poi, all surrogate poles, pre-transform and area must be bit-identical."""
import glob
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ASSETS = sorted(f for f in glob.glob(os.path.join(ROOT, "assets", "*.json")) if "config" not in os.path.basename(f))


@pytest.mark.parametrize("path", ASSETS, ids=[os.path.basename(p) for p in ASSETS])
def test_import_matches_oracle(jg, orc, path):
    inst = orc.load_instance_json(path)
    items = [it for it in sorted(inst["items"], key=lambda it: it["id"]) if it.get("demand", 1) > 0]
    w = orc.OracleWorld()
    for it in items:
        pts, _ = orc._shape_points(it["shape"])
        o = w.item(w.add_item(pts, it.get("allowed_orientations"), it.get("min_quality"), it.get("demand", 1)))
        shape, pre, area = jg.import_item(pts)
        assert np.array_equal(shape.verts, o["verts"])
        assert np.array_equal(shape.poi, o["poi"])
        assert np.array_equal(shape.poles, o["poles"]) and len(shape.poles) >= 2
        assert np.array_equal(pre, o["pre_transform"]) and np.float32(area) == np.float32(o["area"])


def test_import_cleans_and_orients(jg, orc):
    # clockwise square with a closing duplicate and an approx-equal neighbour (io/import.rs:198-260)
    pts = np.array([[0, 0], [0, 2], [2, 2], [2, 2.0000001], [2, 0], [0, 0]], np.float32)
    shape, pre, area = jg.import_item(pts)
    assert len(shape.verts) == 4 and area == pytest.approx(4.0)
    x, y = shape.verts[:, 0], shape.verts[:, 1]
    assert float(np.sum(x * np.roll(y, -1) - np.roll(x, -1) * y)) > 0  # counter-clockwise
    assert np.allclose(shape.verts.mean(0), 0.0, atol=1e-6) and np.allclose(pre, [-1.0, -1.0])
    w = orc.OracleWorld()
    o = w.item(w.add_item(pts))
    assert np.array_equal(shape.verts, o["verts"]) and np.array_equal(shape.poles, o["poles"])


@pytest.mark.parametrize("pts", [[[0, 0], [1, 0]], [[0, 0], [1, 1], [2, 2]], [[0, 0], [1, 0], [0, 0], [1, 1], [0, 1]]])
def test_degenerate_polygons_are_errors(jg, pts):
    with pytest.raises(jg.JaguaCudaError):  # SPolygon::new bails (simple_polygon.rs:41-52)
        jg.import_item(np.array(pts, np.float32))
