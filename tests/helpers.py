"""Shared test helpers: rebuild the CPU oracle's world from a workload fixture (same hazard order,
hence the same slot numbering as the product's host mirror)."""
import numpy as np

ALL_WORKLOADS = ["swim_half", "shirts_full", "trousers_full", "baldacci1_bin", "albano_half", "jakobs1_full",
                 "swim_half_d0", "swim_half_d3", "shirts_full_d10", "gardeyn1c_half",
                 # config 4: second bin (1884-vertex outer boundary, holes, zones); continuous rotation, up to 1044 vertices
                 "baldacci3_bin", "gardeyn4c_half",
                 # config 5: synthesised multi-strip layouts with a minimum separation (first / last strip of each)
                 "jakobs1_ms0", "jakobs1_ms3", "jakobs2_ms1", "albano_ms0", "albano_ms2"]
MSPP_STRIPS = {"jakobs1": 4, "jakobs2": 4, "albano": 3}


def oracle_world(orc, wl, count=False):
    w = orc.OracleWorld(wl.quadtree_depth, wl.cd_threshold, wl.n_ff_poles, wl.n_ff_piers, count=count)
    for it in wl.items:
        w.add_item_raw(it.verts, it.poi, it.poles, it.piers, it.rotations, it.min_quality)
    statics = [h for h in wl.hazards if not h.dynamic]
    w.set_cde_raw(wl.cde_bbox, [dict(kind=h.kind, id=h.id, quality=h.quality, verts=h.verts, poi=h.poi) for h in statics])
    keys = []
    for h in wl.hazards:
        if h.dynamic:
            keys.append(w.register_hazard_raw(h.kind, h.verts, h.poi, h.id, h.quality, h.pk, True))
    w.dynamic_keys = keys
    return w


def verdict_summary(v):
    v = np.asarray(v)
    return dict(n=len(v), surrogate=int((v == 1).sum()), poly=int((v == 2).sum()), free=int((v == 0).sum()),
                panic=int((v == 0x80).sum()))


def place_remove_sequence(jg, wl, cde, w, n_ops=24, seed=0, after_each=None):
    """Drives the product's host mirror and the oracle through the same Layout::place_item / remove_item
    sequence (entities/layout.rs:40-74): all fixture placements are removed first, then items are placed at
    seeded transforms and removed again in random order. Returns the PItemKeys still placed."""
    rng = np.random.default_rng(seed)
    for k, ok, h in zip(cde.hazard_keys, _all_keys(cde, w, wl), wl.hazards):
        if h.dynamic:
            cde.deregister_hazard_by_key(k)
            w.deregister_hazard(ok)
    placed = []
    b = wl.cde_bbox
    for op in range(n_ops):
        if placed and rng.random() < 0.35:
            pk = placed.pop(int(rng.integers(len(placed))))
            cde.remove_item(pk)
            w.remove_item(pk)
        else:
            i = int(rng.integers(len(wl.items)))
            it = wl.items[i]
            rot = 0.0 if it.rotations is not None and len(it.rotations) == 0 else (
                float(rng.random() * 6.28) if it.rotations is None else float(it.rotations[int(rng.integers(len(it.rotations)))]))
            dt = np.array([[rot, b[0] + (b[2] - b[0]) * rng.random(), b[1] + (b[3] - b[1]) * rng.random()]], np.float32)
            pk = w.place_item(i, float(dt[0, 0]), float(dt[0, 1]), float(dt[0, 2]))
            hk = cde.place_item(i, jg.ItemShape(it.verts, it.poi, it.poles, it.piers), jg.compose(dt)[0], pk)
            assert hk != 0
            placed.append(pk)
        if after_each is not None:
            after_each(op)
    return placed


def _all_keys(cde, w, wl):
    """oracle HazKeys in fixture order (static hazards have no oracle-side key list: None)."""
    it = iter(w.dynamic_keys)
    return [next(it) if h.dynamic else None for h in wl.hazards]
