"""Shared test helpers: rebuild the CPU oracle's world from a workload fixture (same hazard order,
hence the same slot numbering as the product's host mirror)."""
import numpy as np

ALL_WORKLOADS = ["swim_half", "shirts_full", "trousers_full", "baldacci1_bin", "albano_half", "jakobs1_full",
                 "swim_half_d0", "swim_half_d3", "shirts_full_d10", "gardeyn1c_half"]


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
