"""CPU tests of the update-path additions of round 2 (host mirror behind jg_cde_*), against the oracle:
Layout::swap_container (entities/layout.rs:40-50), CDEngine::deregister_hazard_by_entity (cd_engine.rs:72-85),
CDEngine::clone, and the haz_version section that carries HazKey versions to the device."""
import ctypes as C
import os
import struct

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _strip_engine(jg, bbox_w, h, device=None):
    """Container::base_cde of a strip (probs/spp/entities/strip.rs:44-61): rectangle (0, 0, w, h), CDE bbox = its square hull."""
    w = np.float32(bbox_w)
    h = np.float32(h)
    dx = dy = np.float32(0.0)
    if h < w:
        dy = np.float32((w - h) / np.float32(2.0))
    elif w < h:
        dx = np.float32((h - w) / np.float32(2.0))
    bbox = np.array([np.float32(0.0) - dx, np.float32(0.0) - dy, w + dx, h + dy], np.float32)
    cde = jg.CDEngine(bbox, device=device)
    verts, poi = jg.polygon_poi(np.array([[0, 0], [w, 0], [w, h], [0, h]], np.float32))
    cde.register_hazard(jg.HAZ_EXTERIOR, verts, poi, dynamic=False)
    return cde


@pytest.mark.parametrize("inst_name", ["swim", "shirts", "albano"])
def test_swap_container_matches_oracle(jg, orc, inst_name):
    """change_strip_width (probs/spp/entities/problem.rs:27-37) = Layout::swap_container: the quadtree after the swap -
    dynamic hazards re-registered in snapshot order into a clone of the new base CDE - is the oracle's word for word,
    HazKeys included; further place / remove operations keep agreeing."""
    inst = orc.load_instance_json(os.path.join(ROOT, "assets", f"{inst_name}.json"))
    w = orc.world_from_instance(inst)
    width, height = w.strip_width() * 1.6, float(np.float32(inst["strip_height"]))
    w.change_strip_width(width)  # (the oracle's own swap, on an empty layout)
    cde = _strip_engine(jg, width, height)
    assert np.array_equal(cde.tree_dump(), w.tree_dump())
    rng = np.random.default_rng(5)
    items = [w.item(i) for i in range(w.n_items())]
    placed = []
    for k in range(14):
        i = int(rng.integers(len(items)))
        rots = items[i]["rotations"]
        rot = float(rots[int(rng.integers(len(rots)))]) if rots is not None and len(rots) else 0.0
        dt = np.array([[rot, rng.random() * width, rng.random() * height]], np.float32)
        pk = w.place_item(i, float(dt[0, 0]), float(dt[0, 1]), float(dt[0, 2]))
        cde.place_item(i, jg.ItemShape(items[i]["verts"], items[i]["poi"], items[i]["poles"], items[i]["piers"]), jg.compose(dt)[0], pk)
        placed.append(pk)
    for pk in (placed.pop(3), placed.pop(7)):  # holes in the slot map: the re-registration order is the slot order
        w.remove_item(pk)
        cde.remove_item(pk)
    assert np.array_equal(cde.tree_dump(), w.tree_dump())
    for new_w in (width * 1.1, width * 0.97, width * 1.3):
        w.change_strip_width(float(np.float32(new_w)))
        base = _strip_engine(jg, np.float32(new_w), height)
        e0 = cde.epoch()
        cde.swap_container(base)
        base.close()
        assert cde.epoch() > e0
        assert np.array_equal(cde.tree_dump(), w.tree_dump()), f"trees differ after the swap to width {new_w}"
        assert w.check_invariants() == 0
        assert sorted(int(h["key"]) for h in w.hazards()) == sorted(_product_keys(jg, cde))
    # the swapped engine keeps working: same keys for the next placement, same tree
    dt = np.array([[0.0, width * 0.5, height * 0.5]], np.float32)
    pk = w.place_item(0, 0.0, float(dt[0, 1]), float(dt[0, 2]))
    hk = cde.place_item(0, jg.ItemShape(items[0]["verts"], items[0]["poi"], items[0]["poles"], items[0]["piers"]), jg.compose(dt)[0], pk)
    assert hk == cde.haz_key_from_pi_key(pk)
    assert np.array_equal(cde.tree_dump(), w.tree_dump())
    cde.close()


def _product_keys(jg, cde):
    blob = cde.flatten()
    hdr = struct.unpack_from("<12I4f12I", blob.tobytes(), 0)
    n_slots, off_meta, off_ver = hdr[6], hdr[22], hdr[26]
    hm = np.frombuffer(blob, np.int32, 4 * n_slots, off_meta).reshape(-1, 4)
    hv = np.frombuffer(blob, np.uint32, n_slots, off_ver)
    return [int(s) | (int(hv[s]) << 32) for s in range(n_slots) if hm[s, 2] > 0]


def test_deregister_by_entity_matches_by_key(jg, orc):
    from helpers import oracle_world

    wl = jg.load_workload("baldacci1_bin")
    a = jg.build_engine(wl, device=None)
    b = jg.build_engine(wl, device=None)
    w = oracle_world(orc, wl)
    dyn = [(k, h) for k, h in zip(a.hazard_keys, wl.hazards) if h.dynamic]
    for (k, h), ok in list(zip(dyn, w.dynamic_keys))[::2]:
        a.deregister_hazard_by_key(k)
        b.deregister_hazard_by_entity(h.kind, id=h.id, quality=h.quality, pk=h.pk)
        w.deregister_hazard(ok)
        assert np.array_equal(a.tree_dump(), b.tree_dump()) and np.array_equal(a.tree_dump(), w.tree_dump())
    with pytest.raises(jg.JaguaCudaError):  # cd_engine.rs:78 "Cannot deregister hazard that is not registered"
        b.deregister_hazard_by_entity(jg.HAZ_PLACED_ITEM, id=12345, pk=0x999900000001)
    zone = [h for h in wl.hazards if h.kind == jg.HAZ_QZONE][0]
    with pytest.raises(jg.JaguaCudaError):  # same kind and id but another quality is another entity
        b.deregister_hazard_by_entity(zone.kind, id=zone.id, quality=zone.quality + 1, pk=zone.pk)
    b.deregister_hazard_by_entity(zone.kind, id=zone.id, quality=zone.quality, pk=zone.pk)
    assert b.n_hazards() == a.n_hazards() - 1
    a.close(), b.close()


def test_clone_is_independent(jg):
    wl = jg.load_workload("swim_half")
    a = jg.build_engine(wl, device=None)
    lib = jg.load_library()
    c = C.c_void_p()
    assert lib.jg_cde_clone(a._cde, C.byref(c)) == jg.JG_OK
    n = lib.jg_cde_tree_dump(c, None, 0)
    before = a.tree_dump()
    assert n == len(before)
    k = [k for k, h in zip(a.hazard_keys, wl.hazards) if h.dynamic][0]
    a.deregister_hazard_by_key(k)
    buf = np.zeros(n, np.uint32)
    lib.jg_cde_tree_dump(c, buf.ctypes.data, n)
    assert np.array_equal(buf, before) and not np.array_equal(a.tree_dump(), before)
    assert lib.jg_cde_epoch(c) + 1 == a.epoch()
    lib.jg_cde_destroy(c)
    a.close()


def test_pool_needs_a_gpu_or_reports_cuda_error(jg):
    """jg_pool_create without a usable device fails loudly (no CPU fallback); bad arguments are JG_ERR_INVALID."""
    lib = jg.load_library()
    cfg = jg.JgConfig(5, 16, 2, 0)
    out = C.c_void_p()
    devs = (C.c_int * 2)(0, 0)
    assert lib.jg_pool_create(devs, 2, C.byref(cfg), C.byref(out)) == jg.JG_ERR_INVALID  # duplicate device
    assert lib.jg_pool_create(devs, 0, C.byref(cfg), C.byref(out)) == jg.JG_ERR_INVALID
    assert lib.jg_pool_size(None) == 0 and lib.jg_pool_sync(None) == jg.JG_ERR_INVALID


def test_lattice_points_are_plain_f32_arithmetic(jg):
    """jg_lattice_points: x0 + (float)i * dx with one rounding per operation - numpy float32 arithmetic reproduces it."""
    rng = np.random.default_rng(1)
    lat = jg.JgLattice.over_box(np.array([-3.25, 10.5, 5748.75, 5762.5], np.float32))
    ij = rng.integers(0, 65536, (5000, 2)).astype(np.uint16)
    ij[:4] = [[0, 0], [65535, 65535], [0, 65535], [1, 1]]
    got = jg.lattice_points(lat, ij)
    x = np.float32(lat.x0) + ij[:, 0].astype(np.float32) * np.float32(lat.dx)
    y = np.float32(lat.y0) + ij[:, 1].astype(np.float32) * np.float32(lat.dy)
    assert np.array_equal(got[:, 0], x.astype(np.float32)) and np.array_equal(got[:, 1], y.astype(np.float32))
    assert got[:, 0].max() < 5748.75 and got[:, 1].max() < 5762.5  # the lattice stays inside the half-open box


# ---- incremental device image (jg_cde_image_sync) --------------------------------------------------------------
def _decode_tree(blob):
    """canonical pre-order dump (the format of jg_cde_tree_dump) of a layout blob, packed or position-stable"""
    h = struct.unpack_from("<12I4f12I", blob.tobytes(), 0)
    n_nodes, n_entries, n_edges, root_level = h[3], h[4], h[5], h[8]
    o = h[16:27]
    nm = np.frombuffer(blob, np.int32, 4 * n_nodes, o[0]).reshape(-1, 4)
    nb = np.frombuffer(blob, np.uint32, 4 * n_nodes, o[1]).reshape(-1, 4)
    na = np.frombuffer(blob, np.uint32, n_nodes, o[2])
    em = np.frombuffer(blob, np.int32, 4 * n_entries, o[3]).reshape(-1, 4)
    ef = np.frombuffer(blob, np.uint32, 4 * n_entries, o[4]).reshape(-1, 4)
    ed = np.frombuffer(blob, np.uint32, 4 * n_edges, o[5]).reshape(-1, 4)
    out = []

    def rec(n, level):
        cb, eb, cnt, _ = (int(v) for v in nm[n])
        n_ent = (cnt & 0xFFFF) + ((cnt >> 16) & 0xFFFF)
        out.extend([level, *[int(v) for v in nb[n]], 1 if cb >= 0 else 0, n_ent, int(na[n])])
        ents = sorted(range(eb, eb + n_ent), key=lambda i: (-int(em[i, 3]), int(em[i, 0])))
        for i in ents:
            out.extend([int(em[i, 0]), int(em[i, 3])])
            if em[i, 3] == 1:
                out.extend(int(v) for v in ef[i])
                out.append(int(em[i, 2]))
                for k in range(int(em[i, 1]), int(em[i, 1]) + int(em[i, 2])):
                    out.extend(int(v) for v in ed[k])
            else:
                out.extend([0] * 5)
        if cb >= 0:
            for q in range(4):
                rec(cb + q, level - 1)

    rec(0, root_level)
    return np.array(out, np.uint32)


@pytest.mark.parametrize("name", ["swim_half", "shirts_full", "baldacci1_bin", "swim_half_d0", "shirts_full_d10"])
def test_incremental_image_tracks_the_tree(jg, orc, name):
    """The position-stable image after every place / remove / restore: (1) it decodes to the same
    quadtree as the host mirror and the oracle, word for word; (2) a consumer that applies only the reported byte ranges
    to its previous copy holds exactly the image (nothing that changed goes unreported); (3) an update re-serialises a
    fraction of the image, not all of it."""
    from helpers import oracle_world, place_remove_sequence

    wl = jg.load_workload(name)
    cde = jg.build_engine(wl, device=None)
    w = oracle_world(orc, wl)
    shadow = {"blob": None, "lineage": 0, "seq": 0}
    stats = {"patched": 0, "full": 0, "bytes": 0, "total": 0}

    def sync(_op=None):
        blob, lin, seq, ranges = cde.image_sync()
        if lin != shadow["lineage"]:
            assert len(ranges) == 0  # a new lineage is taken whole
            shadow["blob"] = blob.copy()
            stats["full"] += 1
        else:
            assert seq == shadow["seq"] + 1 and len(ranges) > 0
            assert np.all(ranges[:, 0] % 4 == 0) and np.all(ranges[:, 1] % 4 == 0) and np.all(np.diff(ranges[:, 0]) > 0)
            for off, ln in ranges:
                shadow["blob"][off:off + ln] = blob[off:off + ln]
            stats["patched"] += 1
            stats["bytes"] += int(ranges[:, 1].sum())
            stats["total"] += len(blob)
        shadow["lineage"], shadow["seq"] = lin, seq
        assert np.array_equal(shadow["blob"], blob), "a changed byte was not reported"
        dump = cde.tree_dump()
        assert np.array_equal(_decode_tree(blob), dump)
        assert np.array_equal(dump, w.tree_dump())
        # the grid premise, kept node by node for the image, equals the whole-tree check of the packed blob
        assert blob.view(np.uint32)[11] == cde.flatten().view(np.uint32)[11]

    sync()
    place_remove_sequence(jg, wl, cde, w, n_ops=40, seed=11, after_each=sync)
    snap_p, snap_o = cde.save(), w.cde_save()
    rng = np.random.default_rng(12)
    b = wl.cde_bbox
    for k in range(5):  # a few more placements, then back to the snapshot
        i = int(rng.integers(len(wl.items)))
        it = wl.items[i]
        rot = 0.0 if it.rotations is not None and len(it.rotations) == 0 else (
            float(rng.random() * 6.28) if it.rotations is None else float(it.rotations[int(rng.integers(len(it.rotations)))]))
        dt = np.array([[rot, b[0] + (b[2] - b[0]) * rng.random(), b[1] + (b[3] - b[1]) * rng.random()]], np.float32)
        pk = w.place_item(i, float(dt[0, 0]), float(dt[0, 1]), float(dt[0, 2]))
        cde.place_item(i, jg.ItemShape(it.verts, it.poi, it.poles, it.piers), jg.compose(dt)[0], pk)
    sync()
    cde.restore(snap_p)
    w.cde_restore(snap_o)
    sync()
    assert np.array_equal(_decode_tree(cde.flatten()), cde.tree_dump())  # (the packed form decodes to the same tree)
    assert stats["patched"] >= 30, stats
    if wl.quadtree_depth >= 3:
        assert stats["bytes"] < 0.35 * stats["total"], stats  # updates are local
    cde.close()


@pytest.mark.parametrize("n,n_dev,min_shard", [(0, 4, 32768), (1, 8, 32768), (32768, 8, 32768), (32769, 8, 32768), (1 << 20, 8, 32768),
                                               (1_000_003, 3, 1), (100_000_007, 8, 1 << 15), (200_000, 8, 1 << 15)])
def test_pool_sharding_rule(jg, n, n_dev, min_shard):
    """jg_pool_shard_bounds (the rule jg_pool_* cut a host batch by): contiguous, disjoint, complete, sizes differ by at most one
    among the active devices, and a batch is not spread thinner than min_shard queries per device."""
    import ctypes as C

    lib = jg.load_library()
    cuts = []
    for k in range(n_dev):
        lo, hi = C.c_int64(-1), C.c_int64(-1)
        assert lib.jg_pool_shard_bounds(n, k, n_dev, min_shard, C.byref(lo), C.byref(hi)) == 0
        cuts.append((lo.value, hi.value))
    assert cuts[0][0] == 0 and cuts[-1][1] == n
    assert all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))
    sizes = [b - a for a, b in cuts if b > a]
    if n > 0:
        assert max(sizes) - min(sizes) <= 1
        assert len(sizes) == min(n_dev, max(1, -(-n // min_shard)))
    assert lib.jg_pool_shard_bounds(10, n_dev, n_dev, min_shard, C.byref(C.c_int64()), C.byref(C.c_int64())) != 0  # k out of range
