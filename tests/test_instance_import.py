"""Product-side importer (jagua-rs_b200/instance.py over jg_item_import / jg_shape_import) vs the oracle's restatement
of io/import.rs + entities/container.rs: containers with holes and quality zones give the same CDE bbox, the same static
hazards in the same slot order and the same quadtree, word for word."""
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_bins_with_holes_and_zones_match_oracle(jg, orc):
    d = orc.load_instance_json(os.path.join(ROOT, "assets", "baldacci1.json"))
    inst = jg.instance.load_instance(d)
    assert len(inst.bins) == 12 and any(it.min_quality is not None for it in inst.items)
    kinds_seen = set()
    for b in range(len(inst.bins)):
        bbox, mine = jg.instance.container_hazards(inst.bins[b])
        w = orc.world_from_instance(d, bin_index=b)
        theirs = w.hazards()
        assert np.array_equal(bbox, w.cde_bbox()) and len(mine) == len(theirs)
        for h, m in zip(theirs, mine):
            assert (h["kind"], h["id"], h["quality"]) == (m[0], m[1], m[2])
            assert np.array_equal(h["verts"], m[3]) and np.array_equal(h["poi"], m[4])
            kinds_seen.add(m[0])
        cde = jg.instance.container_engine(inst, b, device=None)
        assert np.array_equal(cde.tree_dump(), w.tree_dump())
        assert cde.hkey_exterior() == w.L.orc_hkey_exterior(w.w)
        cde.close()
    assert kinds_seen == {jg.HAZ_EXTERIOR, jg.HAZ_HOLE, jg.HAZ_QZONE}


@pytest.mark.parametrize("asset", ["swim.json", "shirts.json", "gardeyn1_c.json"])
def test_items_rotations_and_demands(jg, orc, asset):
    d = orc.load_instance_json(os.path.join(ROOT, "assets", asset))
    inst = jg.instance.load_instance(d)
    w = orc.world_from_instance(d)
    assert len(inst.items) == w.n_items() and inst.strip_height == pytest.approx(float(d["strip_height"]))
    for it in inst.items:
        o = w.item(it.id)
        assert it.demand == o["demand"]
        if o["rotations"] is None:
            assert it.rotations is None
        else:
            assert np.array_equal(it.rotations, o["rotations"])  # f32::to_radians, bit for bit
        assert np.array_equal(it.shape.poles, o["poles"])
