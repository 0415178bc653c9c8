"""jg_item_generate_piers (product C++) against oracle/piers.py (an independent numpy-float32 restatement of
geometry/fail_fast/piers.rs:25-226): the same piers, bit for bit; plus the properties the construction guarantees (piers
lie inside the shape; each selection does not increase the loss) and the reference's own use of piers: an item with
generated piers keeps the verdicts of detect_surrogate_collision consistent with its polygon (a pier inside the shape can
only collide where the shape does)."""
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def items_of(jg, name, n_ff_poles=2):
    with open(os.path.join(ROOT, "assets", name)) as f:
        inst = json.load(f)
    out = []
    for it in inst["items"]:
        pts, _ = jg.instance.shape_points(it["shape"])
        shape, _, _ = jg.import_item(pts, n_ff_poles=n_ff_poles)
        out.append(shape)
    return out


# (instance, item indices, n_piers, rays_per_angle, n_angles, n_points_per_dimension)
CASES = [
    ("swim.json", [0, 3, 7], 3, 24, 10, 24),
    ("shirts.json", [0, 4], 2, 20, 8, 20),
    ("albano.json", [1, 5], 4, 16, 12, 30),
    ("swim.json", [5], 2, 10, 4, 10),      # the reference's debug-build constants (piers.rs:16-18)
]


@pytest.mark.parametrize("name,idx,n,rpa,na,npd", CASES, ids=[f"{c[0]}-{c[3]}x{c[4]}x{c[5]}" for c in CASES])
def test_piers_match_the_oracle(jg, name, idx, n, rpa, na, npd):
    import piers as opiers

    shapes = items_of(jg, name)
    for i in idx:
        sh = shapes[i]
        got = jg.generate_piers(sh.verts, sh.poles[:2], n, rpa, na, npd)
        exp = opiers.generate_piers(sh.verts, sh.poles[:2], n, rpa, na, npd)
        assert got.shape == (n, 4)
        assert np.array_equal(got.view(np.uint32), exp.view(np.uint32)), (name, i, got, exp)


def test_piers_release_constants_one_item(jg):
    """The reference's release constants (200 rays x 90 angles, 100 x 100 grid) on one swim item."""
    import piers as opiers

    sh = items_of(jg, "swim.json")[2]
    got = jg.generate_piers(sh.verts, sh.poles[:2], 2)
    exp = opiers.generate_piers(sh.verts, sh.poles[:2], 2)
    assert np.array_equal(got.view(np.uint32), exp.view(np.uint32)), (got, exp)


def test_piers_lie_inside_the_shape(jg, orc):
    for name in ("swim.json", "trousers.json"):
        for sh in items_of(jg, name)[:4]:
            piers = jg.generate_piers(sh.verts, sh.poles[:2], 3, 30, 12, 30)
            v = np.ascontiguousarray(sh.verts, np.float32)
            for sx, sy, ex, ey in piers:
                assert (sx, sy) != (ex, ey)
                for t in np.linspace(0.0, 1.0, 9):  # clipped to the border and trimmed by 0.1 %: strictly inside
                    x, y = np.float32(sx + t * (ex - sx)), np.float32(sy + t * (ey - sy))
                    assert orc.lib().orc_pip(v, len(v), float(x), float(y)) == 1


def test_no_piers_and_bad_input(jg):
    sh = items_of(jg, "swim.json")[0]
    assert jg.generate_piers(sh.verts, sh.poles[:2], 0).shape == (0, 4)
    with pytest.raises(jg.JaguaCudaError):
        jg.generate_piers(sh.verts[::-1].copy(), sh.poles[:2], 1, 10, 4, 10)  # clockwise: not an imported shape
    shape, _, _ = jg.import_item(jg.instance.shape_points(json.load(open(os.path.join(ROOT, "assets", "swim.json")))["items"][0]["shape"])[0],
                                 n_ff_piers=0)
    assert shape.piers.shape == (0, 4)
