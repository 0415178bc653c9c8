"""Solution export (jagua-rs_b200/export.py after io/export.rs, io/ext_repr.rs:84-121, probs/{spp,bpp}/io/export.rs): the
external transformation applied to an item's ORIGINAL shape puts it where the internal transformation puts the centred
shape; the JSON carries the reference's field names; rotations are degrees."""
import json
import os
from types import SimpleNamespace

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def apply(dt, pts):
    r, tx, ty = [float(v) for v in dt]
    c, s = np.cos(r), np.sin(r)
    p = np.asarray(pts, np.float64)
    return np.stack([c * p[:, 0] - s * p[:, 1] + tx, s * p[:, 0] + c * p[:, 1] + ty], 1)


@pytest.mark.parametrize("name", ["swim.json", "shirts.json", "baldacci1.json"])
def test_external_transformation_places_the_original_shape(jg, name):
    inst = jg.instance.load_instance(os.path.join(ROOT, "assets", name))
    raw = json.load(open(os.path.join(ROOT, "assets", name)))
    recs = sorted((it for it in raw["items"] if it.get("demand", 1) > 0), key=lambda it: it["id"])
    rng = np.random.default_rng(3)
    for it, rec in list(zip(inst.items, recs))[:8]:
        orig, _ = jg.instance.shape_points(rec["shape"])
        if np.array_equal(orig[0], orig[-1]):
            orig = orig[:-1]
        for _ in range(5):
            dt = np.array([rng.uniform(-3.1, 3.1), rng.uniform(-500, 500), rng.uniform(-500, 500)], np.float32)
            ext = jg.export.int_to_ext_transformation(dt, (0.0, it.pre_transform[0], it.pre_transform[1]))
            centred = apply((0.0, it.pre_transform[0], it.pre_transform[1]), orig)
            want, got = apply(dt, centred), apply(ext, orig)
            scale = max(1.0, float(np.abs(want).max()))
            assert np.abs(want - got).max() <= 2e-5 * scale
            e = jg.export.ext_transformation(ext)
            assert abs(e["rotation"] - np.degrees(float(ext[0]))) <= 1e-4 * max(1.0, abs(e["rotation"]))
            back = np.float32(e["rotation"]) * (np.float32(np.pi) / np.float32(180.0))  # ExtTransformation -> DTransformation
            assert abs(float(back) - float(ext[0])) <= 1e-6


def test_identity_and_pure_translation(jg):
    ext = jg.export.int_to_ext_transformation((0.0, 10.0, -4.0), (0.0, 1.5, 2.5))
    assert np.array_equal(ext, np.array([0.0, 11.5, -1.5], np.float32))
    # a quarter turn: the pre-translation is rotated before the internal translation is added
    ext = jg.export.int_to_ext_transformation((np.float32(np.pi / 2), 0.0, 0.0), (0.0, 1.0, 0.0))
    assert abs(float(ext[0]) - np.pi / 2) < 1e-6 and abs(float(ext[1])) < 1e-6 and abs(float(ext[2]) - 1.0) < 1e-6


def test_spp_and_bpp_solution_records(jg):
    inst = jg.instance.load_instance(os.path.join(ROOT, "assets", "swim.json"))
    pl = [SimpleNamespace(item_id=0, d_transf=np.array([0.0, 100.0, 200.0], np.float32), pk=1),
          SimpleNamespace(item_id=3, d_transf=np.array([np.pi, 900.0, 300.0], np.float32), pk=2)]
    res = SimpleNamespace(strip_width=7000.0, strip_height=inst.strip_height, placements=pl, seconds=1.7)
    sol = jg.export.export_spp_solution(res, inst.items)
    assert set(sol) == {"strip_width", "layout", "density", "run_time_sec"} and sol["run_time_sec"] == 1
    assert set(sol["layout"]) == {"container_id", "placed_items", "density"}
    assert [p["item_id"] for p in sol["layout"]["placed_items"]] == [0, 3]
    assert set(sol["layout"]["placed_items"][0]["transformation"]) == {"rotation", "translation"}
    # sinf(pi_f32) = -8.7e-8: atan2f gives -pi, as in the reference
    assert abs(abs(sol["layout"]["placed_items"][1]["transformation"]["rotation"]) - 180.0) < 1e-3
    want = (inst.items[0].area + inst.items[3].area) / (7000.0 * inst.strip_height)
    assert abs(sol["density"] - want) <= 1e-6 and sol["density"] == sol["layout"]["density"]
    json.dumps(sol)
    lay = SimpleNamespace(bin_id=2, placements=pl)
    bp = jg.export.export_bpp_solution([lay], inst.items, {2: {"area": 4.0e7, "cost": 5}}, run_time_sec=3.9)
    assert set(bp) == {"cost", "layouts", "density", "run_time_sec"} and bp["cost"] == 5 and bp["run_time_sec"] == 3
    assert bp["layouts"][0]["container_id"] == 2 and abs(bp["density"] - bp["layouts"][0]["density"]) < 1e-7
