"""Exports the parity cases of tests/golden/ref_case_*.json (schema 2) for tools/ref_dump.rs - the dumper that replays them
against the REAL jagua-rs crate on a machine that has cargo - and for tests/test_ref_case.py, which replays the same files
through the product path (C ABI -> sm_100a kernels). One `cargo run` pins everything the hot path does:

  swim_d5 / swim_d0 / swim_d3 / swim_d10   strip layout, quadtree depths of lbf/tests/tests.rs:18, NoFilter
  shirts_piers                             n_ff_piers = 2 (fail_fast/piers.rs:25-226: the reference generates its own piers)
  gardeyn1c                                continuous rotation, 176-vertex items
  baldacci1_bin0                           bin with holes and inferior-quality zones, HazKeyFilter::from_irrelevant_qzones per
                                           item.min_quality (hazards/filter.rs:21-41; lbf_bpp.rs:125-127), N3 status bit
  swim_selfkey                             placed items queried at / around their own placement under the self-key filter
                                           (`impl HazardFilter for HazKey`, filter.rs:70-76; Layout::is_feasible, layout.rs:128-136)

A case: instance path, CDE configuration, the container (strip width or bin index), placements in order, removals, and
query sets {item, filter, (rotation, tx, ty) rows, expected verdict bytes, expected hazard sets}. Verdict byte: bit0
detect_surrogate_collision, bit1 detect_poly_collision (evaluated when bit0 == 0), bit7 = the reference panics
(`unreachable!`, cd_engine.rs:179). Hazard ids: -1 Exterior, k >= 0 the k-th placement, 1000000 + idx Hole, 2000000 + idx
InferiorQualityZone (idx = position in the container's zone list, as entities/container.rs:104-131 numbers them).
The expectations are the CPU oracle's (oracle/), so a mismatch reported by ref_dump.rs is a finding about the oracle.

    python tools/export_ref_cases.py [--only name] [--queries 400]

Floats are written as the shortest decimal that round-trips the f32 value.
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import orc  # noqa: E402
import piers as opiers  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--only", default=None)
ap.add_argument("--queries", type=int, default=400)
args = ap.parse_args()


def f32s(a):
    return [float(np.format_float_positional(np.float32(x), unique=True, trim="0")) for x in np.ravel(a)]


def hazard_ids(w, pks):
    """hazard slot -> case id"""
    ids = {}
    n_hole = n_zone = 0
    for h in sorted(w.hazards(), key=lambda h: h["slot"]):
        if h["kind"] == 1:
            ids[h["slot"]] = -1
        elif h["kind"] == 0:
            ids[h["slot"]] = int(np.nonzero(pks == np.uint64(h["pk"]))[0][0])
        elif h["kind"] == 2:
            ids[h["slot"]] = 1000000 + int(h["id"])
        else:
            ids[h["slot"]] = 2000000 + int(h["id"])
    return ids


def sets_of(col, ids, words):
    out = []
    for row in col:
        out.append(sorted(ids[32 * k + bit] for k in range(words) for bit in range(32) if (int(row[k]) >> bit) & 1))
    return out


def sample_dts(w, i, n, seed):
    it = w.item(i)
    b = w.cde_bbox()
    rng = np.random.Generator(np.random.Philox(key=[seed, i]))
    u = rng.random((n, 2), dtype=np.float32)
    tx = (b[0] + (b[2] - b[0]) * u[:, 0]).astype(np.float32)
    ty = (b[1] + (b[3] - b[1]) * u[:, 1]).astype(np.float32)
    rots = it["rotations"]
    rot = (np.zeros(n, np.float32) if rots is not None and len(rots) == 0 else
           (rng.random(n, dtype=np.float32) * np.float32(2 * np.pi)).astype(np.float32) if rots is None else
           np.asarray(rots, np.float32)[rng.integers(0, len(rots), n)])
    return np.stack([rot, tx, ty], 1).astype(np.float32)


def query_set(w, ids, words, item_id, dts, flt=None, filter_bits=None):
    t4 = orc.compose(dts)
    det = w.detect_batch(item_id, t4, filter_bits=filter_bits)
    col = w.collect_batch(item_id, t4, words)  # (the collect functions take no filter: cd_engine.rs:261-310)
    return dict(item_id=int(item_id), filter=flt or {"type": "none"}, d_transfs=[f32s(r) for r in dts],
                expect_detect=[int(v) for v in det], expect_collect=sets_of(col, ids, words))


def make_case(name, instance, cfg, n_q, kind="spp", bin_index=0, samples=3000, piers=0, mode="uniform"):
    inst = orc.load_instance_json(os.path.join(ROOT, "assets", instance))
    w = orc.world_from_instance(inst, bin_index=bin_index, quadtree_depth=cfg["quadtree_depth"], cd_threshold=cfg["cd_threshold"],
                                n_ff_poles=cfg["n_ff_poles"], n_ff_piers=cfg["n_ff_piers"])
    if piers:
        for i in range(w.n_items()):
            it = w.item(i)
            w.set_item_piers(i, opiers.generate_piers(it["verts"], it["poles"][:cfg["n_ff_poles"]], piers))
    if kind == "spp":
        w.lbf_solve_spp(samples, 0.2, 0)
    else:
        w.lbf_fill_bin(samples, 0.2, 0)
    pks, item_ids, dts = w.placed()
    placements = [dict(item_id=int(i), d_transf=f32s(dt)) for i, dt in zip(item_ids, dts)]
    removed = list(range(1, len(pks), 2)) if mode != "selfkey" else list(range(2, len(pks), 3))
    for k in removed:
        w.remove_item(int(pks[k]))
    ids = hazard_ids(w, pks)
    words = max(1, (w.haz_capacity() + 31) // 32)
    qsets = []
    if mode == "uniform":
        for i in range(w.n_items()):
            dt = sample_dts(w, i, n_q, 17)
            mq = w.item(i)["min_quality"]
            if kind == "bpp" and mq >= 0:
                # lbf_bpp.rs:125-127: the search for an item with a min_quality runs under from_irrelevant_qzones(min_quality)
                bits = w.filter_qzones(mq, words)
                qsets.append(query_set(w, ids, words, i, dt, {"type": "qzones", "required_quality": int(mq)}, bits))
            else:
                qsets.append(query_set(w, ids, words, i, dt))
        if kind == "bpp":  # every required quality seen in the instance, on the first item that has one: more N3 coverage
            quals = sorted({w.item(i)["min_quality"] for i in range(w.n_items()) if w.item(i)["min_quality"] >= 0} | {0, 1, 5, 10})
            i0 = next((i for i in range(w.n_items()) if w.item(i)["min_quality"] >= 0), 0)
            for q in quals:
                bits = w.filter_qzones(q, words)
                qsets.append(query_set(w, ids, words, i0, sample_dts(w, i0, n_q, 100 + q), {"type": "qzones", "required_quality": int(q)}, bits))
    else:  # self-key filter: a kept placement queried at (and around) its own transformation with its own hazard filtered out
        kept = [k for k in range(len(pks)) if k not in removed][:8]
        slot_of_pk = {int(h["pk"]): h["slot"] for h in w.hazards() if h["kind"] == 0}
        rng = np.random.default_rng(23)
        scale = float(w.cde_bbox()[2] - w.cde_bbox()[0])
        for k in kept:
            base = np.asarray(dts[k], np.float32)
            n = n_q
            jitter = rng.normal(0, 2e-3 * scale, (n, 2))
            jitter[: n // 4] = 0.0
            jitter[n // 4: n // 2] *= 0.01
            d = np.concatenate([np.full((n, 1), base[0], np.float32), (base[1:][None, :] + jitter).astype(np.float32)], 1).astype(np.float32)
            slot = slot_of_pk[int(pks[k])]
            bits = np.zeros(words, np.uint32)
            bits[slot >> 5] |= np.uint32(1 << (slot & 31))
            qsets.append(query_set(w, ids, words, int(item_ids[k]), d, {"type": "self", "placement": int(k)}, bits))
    case = dict(schema=2, name=name, instance=os.path.join("assets", instance), kind=kind, cde=dict(cfg, n_pole_limits=[[100, 0.0], [20, 0.75], [10, 0.90]]),
                shape_modification="none (poly_simpl_tolerance = None, min_item_separation = None)",
                placements=placements, removed=removed, query_sets=qsets,
                produced_by="oracle/ (CPU restatement), tools/export_ref_cases.py")
    if kind == "spp":
        case["strip_width"] = f32s([w.strip_width()])[0]
    else:
        case["bin_index"] = bin_index
    out = os.path.join(ROOT, "tests", "golden", f"ref_case_{name}.json")
    with open(out, "w") as f:
        json.dump(case, f, separators=(",", ":"))
    n_tot = sum(len(q["d_transfs"]) for q in qsets)
    n_panic = sum(sum(1 for v in q["expect_detect"] if v & 0x80) for q in qsets)
    n_free = sum(sum(1 for v in q["expect_detect"] if v == 0) for q in qsets)
    print(f"{out}: {os.path.getsize(out)} bytes; {len(placements)} placements, {len(removed)} removed, {len(qsets)} query sets, "
          f"{n_tot} queries ({n_free} free, {n_panic} reference panics)")


D = dict(quadtree_depth=5, cd_threshold=16, n_ff_poles=2, n_ff_piers=0)
CASES = {
    "swim_d5": lambda: make_case("swim_d5", "swim.json", D, args.queries),
    "swim_d0": lambda: make_case("swim_d0", "swim.json", dict(D, quadtree_depth=0), args.queries // 2),
    "swim_d3": lambda: make_case("swim_d3", "swim.json", dict(D, quadtree_depth=3), args.queries // 2),
    "swim_d10": lambda: make_case("swim_d10", "swim.json", dict(D, quadtree_depth=10), args.queries // 2),
    "shirts_piers": lambda: make_case("shirts_piers", "shirts.json", dict(D, n_ff_piers=2), args.queries, piers=2),
    "gardeyn1c": lambda: make_case("gardeyn1c", "gardeyn1_c.json", D, args.queries // 2, samples=1500),
    "baldacci1_bin0": lambda: make_case("baldacci1_bin0", "baldacci1.json", D, args.queries // 2, kind="bpp", bin_index=0, samples=1500),
    "swim_selfkey": lambda: make_case("swim_selfkey", "swim.json", D, args.queries, mode="selfkey"),
}
for name, fn in CASES.items():
    if args.only is None or args.only == name:
        fn()
