"""The parity cases exported for tools/ref_dump.rs (tests/golden/ref_case_*.json, schema 2, tools/export_ref_cases.py:
placements, removals, query sets with their filters and the oracle's answers) replayed through the PRODUCT path only:
instance file -> jg_item_import (+ jg_item_generate_piers) -> strip container or bin -> jg_cde_place_item /
jg_cde_remove_item -> filters -> detect / collect kernels. The same files are what the Rust dumper replays against the real
crate on a machine that has cargo, so the three implementations meet on one set of expectations:
quadtree depths 0 / 3 / 5 / 10 (lbf/tests/tests.rs:18), surrogate piers, continuous rotation, a bin with holes and
inferior-quality zones under from_irrelevant_qzones filters, and the self-key filter of Layout::is_feasible."""
import glob
import json
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "ref_case_*.json")))


def _name(path):
    return os.path.basename(path)[len("ref_case_"):-len(".json")]


def build_product_layout(jg, case, device):
    """-> (instance, engine, {hazard slot: case id}, [HazKey per placement])"""
    inst = jg.instance.load_instance(os.path.join(ROOT, case["instance"]))
    c = case["cde"]
    cfg = dict(quadtree_depth=c["quadtree_depth"], cd_threshold=c["cd_threshold"], n_ff_poles=c["n_ff_poles"], n_ff_piers=c["n_ff_piers"])
    slot_of = {}
    if case["kind"] == "spp":
        w, h = np.float32(case["strip_width"]), np.float32(inst.strip_height)
        dx = dy = np.float32(0.0)
        if h < w:
            dy = np.float32((w - h) / np.float32(2.0))
        elif w < h:
            dx = np.float32((h - w) / np.float32(2.0))
        bbox = np.array([np.float32(0.0) - dx, np.float32(0.0) - dy, w + dx, h + dy], np.float32)  # rect.rs:93-109 inflate_to_square
        cde = jg.CDEngine(bbox, device=device, **cfg)
        verts, poi = jg.polygon_poi(np.array([[0, 0], [w, 0], [w, h], [0, h]], np.float32))
        ext = cde.register_hazard(jg.HAZ_EXTERIOR, verts, poi, dynamic=False)
        slot_of[ext & 0xFFFFFFFF] = -1
    else:
        cde = jg.instance.container_engine(inst, bin_index=case["bin_index"], device=None, **cfg)
        if device is not None:  # (container_engine with a device also uploads the plain item shapes; piers are added below)
            cde.close()
            cde = jg.instance.container_engine(inst, bin_index=case["bin_index"], device=device, **cfg)
        _, hazards = jg.instance.container_hazards(inst.bins[case["bin_index"]])
        for key, (kind, idx, q, v, poi) in zip(cde.hazard_keys, hazards):
            slot_of[key & 0xFFFFFFFF] = -1 if kind == jg.HAZ_EXTERIOR else (1000000 + idx if kind == jg.HAZ_HOLE else 2000000 + idx)
    shapes = {}
    for it in inst.items:
        sh = it.shape
        if c["n_ff_piers"] > 0:  # SPSurrogate::new (sp_surrogate.rs:40-44): piers against the first n_ff_poles poles
            sh = jg.ItemShape(sh.verts, sh.poi, sh.poles, jg.generate_piers(sh.verts, sh.poles[:c["n_ff_poles"]], c["n_ff_piers"]))
        shapes[it.id] = sh
    keys = []
    for idx, p in enumerate(case["placements"]):
        t4 = jg.compose(np.array([p["d_transf"]], np.float32))[0]
        keys.append(cde.place_item(p["item_id"], shapes[p["item_id"]], t4, pk=idx + 1))
    for r in case["removed"]:
        cde.remove_item(r + 1)
    removed = set(case["removed"])
    for idx, k in enumerate(keys):
        if idx not in removed:
            slot_of[k & 0xFFFFFFFF] = idx
    if device is not None:
        for iid, sh in shapes.items():
            cde.upload_item(iid, sh)
    return inst, cde, slot_of, keys


def build_oracle_layout(orc, case):
    d = orc.load_instance_json(os.path.join(ROOT, case["instance"]))
    c = case["cde"]
    w = orc.world_from_instance(d, bin_index=case.get("bin_index", 0), quadtree_depth=c["quadtree_depth"], cd_threshold=c["cd_threshold"],
                                n_ff_poles=c["n_ff_poles"], n_ff_piers=c["n_ff_piers"])
    if case["kind"] == "spp":
        w.change_strip_width(case["strip_width"])
    pks = [w.place_item(p["item_id"], *[float(np.float32(v)) for v in p["d_transf"]]) for p in case["placements"]]
    for r in case["removed"]:
        w.remove_item(pks[r])
    return w


@pytest.mark.parametrize("path", CASES, ids=_name)
def test_case_layout_matches_oracle_tree(jg, orc, path):
    """host side only: the product's quadtree for the case's layout is the oracle's, word for word"""
    with open(path) as f:
        case = json.load(f)
    assert case["schema"] == 2
    inst, cde, _, _ = build_product_layout(jg, case, device=None)
    w = build_oracle_layout(orc, case)
    assert np.array_equal(cde.tree_dump(), w.tree_dump())
    cde.close()


def test_there_are_cases_for_every_mechanism():
    names = {_name(p) for p in CASES}
    assert {"swim_d0", "swim_d3", "swim_d5", "swim_d10", "shirts_piers", "gardeyn1c", "baldacci1_bin0", "swim_selfkey"} <= names
    kinds = set()
    for p in CASES:
        with open(p) as f:
            case = json.load(f)
        kinds |= {q["filter"]["type"] for q in case["query_sets"]}
    assert kinds == {"none", "qzones", "self"}


@pytest.mark.gpu
@pytest.mark.parametrize("path", CASES, ids=_name)
def test_gpu_answers_the_exported_case(jg, path):
    with open(path) as f:
        case = json.load(f)
    inst, cde, slot_of, keys = build_product_layout(jg, case, device=0)
    words = cde.collector_words()
    n = 0
    try:
        for fid, q in enumerate(case["query_sets"]):
            t4 = jg.compose(np.array(q["d_transfs"], np.float32))
            flt = q["filter"]
            filter_id = -1
            if flt["type"] == "qzones":  # HazKeyFilter::from_irrelevant_qzones (hazards/filter.rs:21-41)
                cde.upload_filter(fid, cde.filter_from_irrelevant_qzones(flt["required_quality"]))
                filter_id = fid
            elif flt["type"] == "self":  # `impl HazardFilter for HazKey` (filter.rs:70-76), by versioned key
                cde.upload_filter_keys(fid, [keys[flt["placement"]]])
                filter_id = fid
            got = cde.detect_batch(q["item_id"], t4, filter_id=filter_id)
            exp = np.array(q["expect_detect"], np.uint8)
            assert np.array_equal(got, exp), f"{_name(path)} set {fid}: {(got != exp).sum()} verdict mismatches"
            bits = cde.collect_batch(q["item_id"], t4, words)
            for row, want in zip(bits, q["expect_collect"]):
                have = sorted(slot_of[32 * k + b] for k in range(words) for b in range(32) if (int(row[k]) >> b) & 1)
                assert have == want
                n += 1
        assert n == sum(len(q["d_transfs"]) for q in case["query_sets"]) and n >= 1000
    finally:
        cde.close()
