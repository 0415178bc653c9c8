"""The parity case exported for tools/ref_dump.rs (tests/golden/ref_case_swim.json: placements, removals, seeded queries
and the oracle's answers) replayed through the PRODUCT path only: instance file -> jg_item_import -> strip container ->
jg_cde_place_item / jg_cde_remove_item -> detect / collect kernels. The same file is what the Rust dumper replays against
the real crate on a machine that has cargo, so the three implementations meet on one set of expectations."""
import json
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASE = os.path.join(ROOT, "tests", "golden", "ref_case_swim.json")


def build_product_layout(jg, case, device):
    inst = jg.instance.load_instance(os.path.join(ROOT, case["instance"]))
    w, h = np.float32(case["strip_width"]), np.float32(inst.strip_height)
    d = np.float32(abs(w - h) / np.float32(2.0))
    bbox = np.array([0, -d, w, h + d] if h < w else [-d, 0, w + d, h], np.float32)  # rect.rs:93-109 inflate_to_square
    c = case["cde"]
    cde = jg.CDEngine(bbox, c["quadtree_depth"], c["cd_threshold"], c["n_ff_poles"], c["n_ff_piers"], device=device)
    verts, poi = jg.polygon_poi(np.array([[0, 0], [w, 0], [w, h], [0, h]], np.float32))
    ext = cde.register_hazard(jg.HAZ_EXTERIOR, verts, poi, dynamic=False)
    slot_of = {ext & 0xFFFFFFFF: -1}
    keys = []
    for idx, p in enumerate(case["placements"]):
        it = inst.items[p["item_id"]]
        t4 = jg.compose(np.array([p["d_transf"]], np.float32))[0]
        keys.append(cde.place_item(it.id, it.shape, t4, pk=idx + 1))
    for r in case["removed"]:
        cde.remove_item(r + 1)
    removed = set(case["removed"])
    for idx, k in enumerate(keys):
        if idx not in removed:
            slot_of[k & 0xFFFFFFFF] = idx
    if device is not None:
        for it in inst.items:
            cde.upload_item(it.id, it.shape)
    return inst, cde, slot_of


def test_case_layout_matches_oracle_tree(jg, orc):
    """host side only: the product's quadtree for the case's layout is the oracle's, word for word"""
    with open(CASE) as f:
        case = json.load(f)
    inst, cde, _ = build_product_layout(jg, case, device=None)
    d = orc.load_instance_json(os.path.join(ROOT, case["instance"]))
    w = orc.world_from_instance(d)
    w.change_strip_width(case["strip_width"])
    pks = [w.place_item(p["item_id"], *[float(np.float32(v)) for v in p["d_transf"]]) for p in case["placements"]]
    for r in case["removed"]:
        w.remove_item(pks[r])
    assert np.array_equal(cde.tree_dump(), w.tree_dump())
    cde.close()


@pytest.mark.gpu
def test_gpu_answers_the_exported_case(jg):
    with open(CASE) as f:
        case = json.load(f)
    inst, cde, slot_of = build_product_layout(jg, case, device=0)
    words = cde.collector_words()
    n = 0
    for q in case["queries"]:
        t4 = jg.compose(np.array(q["d_transfs"], np.float32))
        assert np.array_equal(cde.detect_batch(q["item_id"], t4), np.array(q["expect_detect"], np.uint8))
        bits = cde.collect_batch(q["item_id"], t4, words)
        for row, want in zip(bits, q["expect_collect"]):
            got = sorted(slot_of[32 * k + b] for k in range(words) for b in range(32) if (int(row[k]) >> b) & 1)
            assert got == want
            n += 1
    assert n == sum(len(q["d_transfs"]) for q in case["queries"]) and n >= 1000
    cde.close()
