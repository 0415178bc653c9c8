"""GPU (C ABI -> sm_100a kernels) vs the CPU oracle on the same seeded inputs: verdict bytes and
order-normalised hazard bitsets must be bit-identical for every query."""
import numpy as np
import pytest

from helpers import ALL_WORKLOADS, oracle_world, place_remove_sequence, verdict_summary

pytestmark = pytest.mark.gpu

N_PER_ITEM = 4096


@pytest.fixture(scope="module")
def engines(jg, orc):
    cache = {}

    def get(name):
        if name not in cache:
            wl = jg.load_workload(name)
            cache[name] = (wl, jg.build_engine(wl, device=0), oracle_world(orc, wl))
        return cache[name]

    yield get
    for _, cde, _ in cache.values():
        cde.close()


@pytest.mark.parametrize("name", ALL_WORKLOADS)
def test_detect_matches_oracle(jg, engines, name):
    wl, cde, w = engines(name)
    total = dict(n=0, surrogate=0, poly=0, free=0, panic=0)
    for item_id in range(len(wl.items)):
        t = jg.make_queries(wl, item_id, N_PER_ITEM, seed=1)
        got = cde.detect_batch(item_id, t)
        exp = w.detect_batch(item_id, t)
        bad = np.nonzero(got != exp)[0]
        assert len(bad) == 0, f"{name} item {item_id}: {len(bad)} mismatches, first {bad[:5]} got {got[bad[:5]]} exp {exp[bad[:5]]}"
        for k, v in verdict_summary(got).items():
            total[k] += v
    # the batch exercises every stage (uniform samples over the CDE square of a narrow, nearly full strip may hold
    # no free placement at all: the synthesised multi-strip layouts only need the polygon stage to be reached)
    assert total["poly"] > 0 and (total["free"] > 0 or "_ms" in name), total


@pytest.mark.parametrize("name", ALL_WORKLOADS)
def test_detect_poly_only_matches_oracle(jg, engines, name):
    wl, cde, w = engines(name)
    for item_id in range(len(wl.items)):
        t = jg.make_queries(wl, item_id, 2048, seed=2)
        got = cde.detect_batch(item_id, t, poly_only=True)
        exp = w.detect_batch(item_id, t, poly_only=True)
        assert np.array_equal(got, exp), f"{name} item {item_id}: {(got != exp).sum()} mismatches"


@pytest.mark.parametrize("name", ALL_WORKLOADS)
def test_collect_matches_oracle(jg, engines, name):
    wl, cde, w = engines(name)
    words = cde.collector_words()
    n_sets = 0
    for item_id in range(len(wl.items)):
        t = jg.make_queries(wl, item_id, 2048, seed=3)
        got = cde.collect_batch(item_id, t, words)
        exp = w.collect_batch(item_id, t, words)
        bad = np.nonzero((got != exp).any(axis=1))[0]
        assert len(bad) == 0, f"{name} item {item_id}: {len(bad)} mismatches, first {bad[:3]} got {got[bad[:3]]} exp {exp[bad[:3]]}"
        n_sets += int((got != 0).any(axis=1).sum())
    assert n_sets > 0


@pytest.mark.parametrize("group", [4, 8, 16, 32])
def test_group_size_does_not_change_results(jg, engines, group):
    wl, cde, w = engines("swim_half")
    cde.set_groups(group, group)
    try:
        for item_id in (0, 8, 9):
            t = jg.make_queries(wl, item_id, 4096, seed=4)
            assert np.array_equal(cde.detect_batch(item_id, t), w.detect_batch(item_id, t))
            words = cde.collector_words()
            assert np.array_equal(cde.collect_batch(item_id, t, words), w.collect_batch(item_id, t, words))
    finally:
        cde.set_groups(0, 0)


@pytest.mark.parametrize("name", ["swim_half", "baldacci1_bin", "gardeyn1c_half", "shirts_full_d10"])
def test_multi_segment_batch_matches_oracle(jg, engines, name):
    """All item types in one launch sequence (jg_detect_multi), ragged and empty segments included."""
    wl, cde, w = engines(name)
    rng = np.random.default_rng(3)
    ids = list(range(len(wl.items))) + [0, 0]
    sizes = [int(x) for x in rng.integers(0, 3000, len(ids))]
    sizes[1] = 0      # empty segment
    sizes[-1] = 1     # single query
    ts = [jg.make_queries(wl, i, n, seed=21 + k) for k, (i, n) in enumerate(zip(ids, sizes))]
    offs = np.concatenate([[0], np.cumsum(sizes)])
    t = np.concatenate(ts)
    exp = np.concatenate([w.detect_batch(i, q) for i, q in zip(ids, ts)])
    got = cde.detect_multi(ids, offs, t)
    assert np.array_equal(got, exp), f"{(got != exp).sum()} mismatches"
    got2 = cde.detect_multi(ids, offs, t, poly_only=True)
    exp2 = np.concatenate([w.detect_batch(i, q, poly_only=True) for i, q in zip(ids, ts)])
    assert np.array_equal(got2, exp2)


def test_many_segments_and_small_chunks(jg, engines):
    """More segments than one launch holds (JG_MAX_SEG = 32) and chunks that cut segments."""
    wl, cde, w = engines("swim_half")
    ids = [k % len(wl.items) for k in range(75)]
    sizes = [37 + 13 * (k % 7) for k in range(75)]
    ts = [jg.make_queries(wl, i, n, seed=100 + k) for k, (i, n) in enumerate(zip(ids, sizes))]
    offs = np.concatenate([[0], np.cumsum(sizes)])
    t = np.concatenate(ts)
    exp = np.concatenate([w.detect_batch(i, q) for i, q in zip(ids, ts)])
    assert np.array_equal(cde.detect_multi(ids, offs, t), exp)
    cde.set_batching(1000, 2)
    try:
        assert np.array_equal(cde.detect_multi(ids, offs, t), exp)
        one = jg.make_queries(wl, 3, 5000, seed=5)
        assert np.array_equal(cde.detect_batch(3, one), w.detect_batch(3, one))
    finally:
        cde.set_batching(0, 4)  # back to the per-batch schedule (the default)


def test_empty_batch(jg, engines):
    wl, cde, w = engines("swim_half")
    assert len(cde.detect_batch(0, np.zeros((0, 4), np.float32))) == 0
    assert cde.collect_batch(0, np.zeros((0, 4), np.float32)).shape[0] == 0


GRID_WORKLOADS = [n for n in ALL_WORKLOADS if n != "shirts_full_d10"]  # depth <= 5: served by the grid kernels


@pytest.mark.parametrize("name", GRID_WORKLOADS)
def test_grid_and_walk_paths_agree_with_oracle(jg, engines, name):
    """The grid kernels (depth <= 5) and the quadtree-walk kernels are two implementations of the same
    reference semantics: both must match the oracle, with and without the surrogate stage."""
    wl, cde, w = engines(name)
    assert cde.detect_path() == "grid"
    try:
        for item_id in range(len(wl.items)):
            t = jg.make_queries(wl, item_id, 3000, seed=11)
            exp = w.detect_batch(item_id, t)
            exp_p = w.detect_batch(item_id, t, poly_only=True)
            for walk in (False, True):
                cde.set_detect_path(walk)
                assert cde.detect_path() == ("walk" if walk else "grid")
                got = cde.detect_batch(item_id, t)
                assert np.array_equal(got, exp), f"{name} item {item_id} walk={walk}: {(got != exp).sum()} mismatches"
                assert np.array_equal(cde.detect_batch(item_id, t, poly_only=True), exp_p)
                words = cde.collector_words()
                tc = t[:1500]
                gc = cde.collect_batch(item_id, tc, words)
                ec = w.collect_batch(item_id, tc, words)
                badc = np.nonzero((gc != ec).any(axis=1))[0]
                assert len(badc) == 0, f"{name} item {item_id} walk={walk}: collect {len(badc)} mismatches, first {badc[:3]} got {gc[badc[:3]]} exp {ec[badc[:3]]}"
                assert np.array_equal(cde.collect_batch(item_id, tc, words, poly_only=True), w.collect_batch(item_id, tc, words, poly_only=True))
    finally:
        cde.set_detect_path(False)


def test_deep_quadtree_uses_walk_path(jg, engines):
    wl, cde, w = engines("shirts_full_d10")
    assert cde.detect_path() == "walk"


def test_grid_path_near_item_placements(jg, engines):
    """Queries concentrated where edges graze hazards: translations of placed items jittered by a few
    units, so that most survivors reach the leaf edge tests and the candidate verification."""
    wl, cde, w = engines("swim_half")
    rng = np.random.default_rng(5)
    placed = [h for h in wl.hazards if h.kind == 0]
    for item_id in range(len(wl.items)):
        c = np.stack([h.verts.mean(0) for h in placed])[rng.integers(0, len(placed), 4000)]
        jit = rng.normal(0, 40.0, c.shape)
        rot = np.asarray(wl.items[item_id].rotations, np.float32)[rng.integers(0, 2, len(c))]
        t = jg.compose(np.concatenate([rot[:, None], (c + jit).astype(np.float32)], 1).astype(np.float32))
        exp = w.detect_batch(item_id, t, poly_only=True)
        got = cde.detect_batch(item_id, t, poly_only=True)
        assert np.array_equal(got, exp), f"item {item_id}: {(got != exp).sum()} mismatches"


@pytest.mark.parametrize("name", ["swim_half", "shirts_full", "shirts_full_d10"])
def test_indexed_queries_match_composed_queries(jg, engines, name):
    """The compact (rotation index, translation) form resolves to the same verdicts as the composed
    (cos, sin, tx, ty) form of the same queries, on the grid and on the walk kernels."""
    wl, cde, w = engines(name)
    ids = list(range(len(wl.items)))
    for i in ids:
        cde.set_item_rotations(i, wl.items[i].rotations)
    sizes = [1500 + 37 * i for i in ids]
    offs = np.concatenate([[0], np.cumsum(sizes)])
    comp = np.concatenate([jg.make_queries(wl, i, n, seed=9) for i, n in zip(ids, sizes)])
    pairs = [jg.make_queries_indexed(wl, i, n, seed=9) for i, n in zip(ids, sizes)]
    ridx = np.ascontiguousarray(np.concatenate([p[0] for p in pairs]))
    txy = np.ascontiguousarray(np.concatenate([p[1] for p in pairs]))
    exp = np.concatenate([w.detect_batch(i, jg.make_queries(wl, i, n, seed=9)) for i, n in zip(ids, sizes)])
    assert np.array_equal(cde.detect_multi(ids, offs, comp), exp)
    assert np.array_equal(cde.detect_multi_indexed(ids, offs, ridx, txy), exp)
    assert np.array_equal(cde.detect_multi_indexed(ids, offs, ridx, txy, poly_only=True),
                          cde.detect_multi(ids, offs, comp, poly_only=True))


def test_indexed_queries_need_registered_rotations(jg):
    wl = jg.load_workload("albano_half")
    cde = jg.build_engine(wl, device=0)
    try:
        with pytest.raises(jg.JaguaCudaError):
            cde.detect_multi_indexed([0], [0, 4], np.zeros(4, np.uint8), np.zeros((4, 2), np.float32))
    finally:
        cde.close()


def test_fused_streaming_path_matches_oracle(jg, engines):
    """JG_PATH_FUSED: one kernel per batch that follows the H2D copy front (host batches >= 128 Ki queries are
    streamed; smaller ones and device-resident ones run the same kernel with everything resident)."""
    wl, cde, w = engines("swim_half")
    ids = list(range(len(wl.items)))
    sizes = jg.split_batch(150_000, len(ids))
    offs = np.concatenate([[0], np.cumsum(sizes)])
    ts = [jg.make_queries(wl, i, n, seed=31) for i, n in zip(ids, sizes)]
    exp = np.concatenate([w.detect_batch(i, q) for i, q in zip(ids, ts)])
    t = jg.PinnedBuffer((int(offs[-1]), 4), np.float32)
    t.array[:] = np.concatenate(ts)
    out = jg.PinnedBuffer((int(offs[-1]),), np.uint8)
    cde.set_detect_path(fused=True)
    try:
        assert cde.detect_path() == "fused"
        for _ in range(3):  # streamed (>= 128 Ki queries), repeated to reuse the buffers
            out.array[:] = 0xFF
            cde.detect_multi(ids, offs, t.array, out=out.array)
            assert np.array_equal(out.array, exp)
        small = jg.make_queries(wl, 3, 5000, seed=32)
        assert np.array_equal(cde.detect_batch(3, small), w.detect_batch(3, small))
    finally:
        cde.set_detect_path()
    assert cde.detect_path() == "grid"
    t.free(), out.free()


@pytest.mark.parametrize("name", ["swim_half", "shirts_full", "baldacci1_bin", "swim_half_d3"])
def test_queries_aligned_with_cell_boundaries(jg, engines, name):
    """Adversarial placements for the grid kernels: translations that put an item vertex (or a pole centre, or a
    pole's extreme point) exactly on a quadtree cell boundary, where every inclusive / strict comparison of the
    reference (rect.rs:282-288, qt_traits.rs:36-44, circle.rs:115-128) decides the verdict."""
    wl, cde, w = engines(name)
    rng = np.random.default_rng(17)
    b = wl.cde_bbox.astype(np.float32)
    G = 1 << wl.quadtree_depth
    xs, ys = np.zeros(G + 1, np.float32), np.zeros(G + 1, np.float32)
    xs[0], xs[G], ys[0], ys[G] = b[0], b[2], b[1], b[3]
    step = G
    while step >= 2:  # nested midpoints, rect.rs:248-253
        for a in range(0, G, step):
            xs[a + step // 2] = (xs[a] + xs[a + step]) / np.float32(2)
            ys[a + step // 2] = (ys[a] + ys[a + step]) / np.float32(2)
        step //= 2
    for item_id, it in enumerate(wl.items):
        n = 3000
        rots = np.zeros(n, np.float32) if it.rotations is None or len(it.rotations) == 0 else \
            np.asarray(it.rotations, np.float32)[rng.integers(0, len(it.rotations), n)]
        cs = jg.compose(np.stack([rots, np.zeros(n, np.float32), np.zeros(n, np.float32)], 1))
        # anchor points of the item in its own frame: vertices, pole centres, pole extreme points
        anchors = [it.verts]
        for p in it.poles[:wl.n_ff_poles]:
            anchors.append(np.array([[p[0], p[1]], [p[0] + p[2], p[1]], [p[0] - p[2], p[1]], [p[0], p[1] + p[2]], [p[0], p[1] - p[2]]], np.float32))
        anchors = np.concatenate(anchors).astype(np.float32)
        a = anchors[rng.integers(0, len(anchors), n)]
        ax = cs[:, 0] * a[:, 0] - cs[:, 1] * a[:, 1]  # rotated anchor (float32, like the device would)
        ay = cs[:, 1] * a[:, 0] + cs[:, 0] * a[:, 1]
        tx = xs[rng.integers(0, G + 1, n)] - ax.astype(np.float32)
        ty = ys[rng.integers(0, G + 1, n)] - ay.astype(np.float32)
        free_axis = rng.integers(0, 3, n)  # align x only, y only, or both
        tx = np.where(free_axis == 1, (b[0] + (b[2] - b[0]) * rng.random(n)).astype(np.float32), tx).astype(np.float32)
        ty = np.where(free_axis == 0, (b[1] + (b[3] - b[1]) * rng.random(n)).astype(np.float32), ty).astype(np.float32)
        t = np.ascontiguousarray(np.stack([cs[:, 0], cs[:, 1], tx, ty], 1).astype(np.float32))
        for poly_only in (False, True):
            exp = w.detect_batch(item_id, t, poly_only=poly_only)
            got = cde.detect_batch(item_id, t, poly_only=poly_only)
            bad = np.nonzero(got != exp)[0]
            assert len(bad) == 0, f"{name} item {item_id} poly_only={poly_only}: {len(bad)} mismatches, first {bad[:5]}"
            words = cde.collector_words()
            gc, ec = cde.collect_batch(item_id, t[:1000], words, poly_only=poly_only), w.collect_batch(item_id, t[:1000], words, poly_only=poly_only)
            assert np.array_equal(gc, ec), f"{name} item {item_id} collect poly_only={poly_only}: {(gc != ec).any(axis=1).sum()} mismatches"


def test_quality_zone_filters_match_oracle(jg, engines):
    """BASELINE.json config 4: bin with holes and inferior-quality zones; every item is queried under the
    HazKeyFilter::from_irrelevant_qzones filter of its min_quality (hazards/filter.rs:21-41; callers
    lbf_bpp.rs:125-127), on the grid and on the walk kernels. Verdict byte incl. the N3 status bit."""
    wl, cde, w = engines("baldacci1_bin")
    words = cde.collector_words()
    filters = {}
    for fid, q in enumerate((0, 1, 2, 5)):
        bits = cde.filter_from_irrelevant_qzones(q)
        assert np.array_equal(bits, w.filter_qzones(q, words))
        cde.upload_filter(fid, bits)
        filters[fid] = bits
    n_filtered_diff = 0
    try:
        for walk in (False, True):
            cde.set_detect_path(walk)
            for item_id in range(0, len(wl.items), 3):
                t = jg.make_queries(wl, item_id, 3000, seed=41)
                base = w.detect_batch(item_id, t)
                for fid, bits in filters.items():
                    exp = w.detect_batch(item_id, t, filter_bits=bits)
                    got = cde.detect_batch(item_id, t, filter_id=fid)
                    assert np.array_equal(got, exp), f"item {item_id} filter {fid} walk={walk}: {(got != exp).sum()} mismatches"
                    expp = w.detect_batch(item_id, t, filter_bits=bits, poly_only=True)
                    gotp = cde.detect_batch(item_id, t, filter_id=fid, poly_only=True)
                    assert np.array_equal(gotp, expp)
                    n_filtered_diff += int((exp != base).sum())
    finally:
        cde.set_detect_path(False)
    assert n_filtered_diff > 0  # the filters really change verdicts


@pytest.mark.parametrize("name", ["swim_half", "shirts_full"])
def test_self_key_filter_matches_oracle(jg, engines, name):
    """`impl HazardFilter for HazKey` (filter.rs:70-76), used by Layout::is_feasible (layout.rs:128-136): a placed
    item queried at (and around) its own placement with its own hazard filtered out."""
    wl, cde, w = engines(name)
    rng = np.random.default_rng(23)
    placed = [(k, h) for k, h in zip(cde.hazard_keys, wl.hazards) if h.kind == 0]
    n_free = 0
    for fid, (key, h) in enumerate(placed[:6]):
        bits = cde.filter_from_keys([key])
        cde.upload_filter(100 + fid, bits)
        item_id = h.id
        it = wl.items[item_id]
        # recover the placement: the hazard's vertices are the item's under one allowed rotation + a translation
        best = None
        for r in np.asarray(it.rotations, np.float32):
            cs = jg.compose(np.array([[r, 0, 0]], np.float32))[0]
            rx = cs[0] * it.verts[:, 0] - cs[1] * it.verts[:, 1]
            ry = cs[1] * it.verts[:, 0] + cs[0] * it.verts[:, 1]
            off = h.verts[0] - np.array([rx[0], ry[0]], np.float32)
            res = float(np.abs(np.stack([rx, ry], 1) + off - h.verts).max()) if len(h.verts) == len(it.verts) else 1e30
            if best is None or res < best[0]:
                best = (res, r, off)
        scale = float(wl.cde_bbox[2] - wl.cde_bbox[0])
        assert best[0] < 1e-3 * scale, "placement not recovered"
        n = 2000
        rots = np.full(n, best[1], np.float32)
        jitter = rng.normal(0, 2e-3 * scale, (n, 2))
        jitter[: n // 4] = 0.0  # the exact placement (up to f32 rounding of the recovered offset)
        jitter[n // 4: n // 2] *= 0.01
        txy = (best[2][None, :] + jitter).astype(np.float32)
        t = jg.compose(np.concatenate([rots[:, None], txy], 1).astype(np.float32))
        exp = w.detect_batch(item_id, t, filter_bits=bits)
        got = cde.detect_batch(item_id, t, filter_id=100 + fid)
        assert np.array_equal(got, exp), f"{name} hazard {fid}: {(got != exp).sum()} mismatches"
        n_free += int((exp == 0).sum())
    assert n_free > 0


def test_filtered_entire_hazard_reports_reference_panic(jg, orc):
    """with the virtual root's Entire hazard filtered out and no edge hit, the reference reaches
    `unreachable!` (cd_engine.rs:179); the verdict byte must carry the status bit, never a silent verdict.
    Synthetic fixture: a square container whose left half is one inferior-quality zone, a small square item."""
    def rect(x0, y0, x1, y1):
        return np.array([[x0, y0], [x1, y0], [x1, y1], [x0, y1]], np.float32)

    wl = jg.Workload("n3_synthetic", 5, 16, 2, 0, np.array([0, 0, 1024, 1024], np.float32))
    wl.hazards.append(jg.workload.WHazard(1, 0, 0, 0, False, rect(0, 0, 1024, 1024), np.array([512, 512, 512], np.float32)))
    wl.hazards.append(jg.workload.WHazard(3, 0, 3, 0, False, rect(0, 0, 512, 1024), np.array([256, 512, 256], np.float32)))
    item = rect(-6, -6, 6, 6)
    wl.items.append(jg.workload.WItem(item, np.array([0, 0, 6], np.float32), np.array([[0, 0, 6], [3, 3, 3]], np.float32),
                                      np.zeros((0, 4), np.float32), np.array([0.0], np.float32), 5))
    cde = jg.build_engine(wl, device=0)
    w = oracle_world(orc, wl)
    try:
        zone_key = cde.hazard_keys[1]
        bits = cde.filter_from_keys([zone_key])
        assert np.array_equal(cde.filter_from_irrelevant_qzones(5), bits)  # quality 3 < required 5 -> irrelevant
        cde.upload_filter(0, bits)
        rng = np.random.default_rng(3)
        n = 20000
        txy = (rng.random((n, 2)) * 1024).astype(np.float32)
        t = jg.compose(np.concatenate([np.zeros((n, 1), np.float32), txy], 1).astype(np.float32))
        seen = 0
        for walk in (False, True):
            cde.set_detect_path(walk)
            for poly_only in (False, True):
                exp = w.detect_batch(0, t, filter_bits=bits, poly_only=poly_only)
                got = cde.detect_batch(0, t, filter_id=0, poly_only=poly_only)
                assert np.array_equal(got, exp), f"walk={walk} poly_only={poly_only}: {(got != exp).sum()} mismatches"
                seen += int((exp == 0x80).sum())
            # without the filter the zone is a plain collision
            assert np.array_equal(cde.detect_batch(0, t), w.detect_batch(0, t))
        assert seen > 0, "the fixture never reaches the reference's unreachable!()"
    finally:
        cde.close()


@pytest.mark.parametrize("name", ["swim_half", "jakobs1_ms1", "albano_ms0", "baldacci1_bin"])
def test_queries_follow_layout_updates(jg, orc, name):
    """Layout::place_item / remove_item between batches (config 5's update cadence): after every few updates the
    re-uploaded layout answers detect and collect batches exactly like the oracle's incrementally updated CDE."""
    wl = jg.load_workload(name)
    cde = jg.build_engine(wl, device=0)
    w = oracle_world(orc, wl)
    words_seen = set()

    def check(op):
        if op % 5 != 4:
            return
        words = cde.collector_words()
        words_seen.add(words)
        for item_id in (0, len(wl.items) // 2, len(wl.items) - 1):
            t = jg.make_queries(wl, item_id, 3000, seed=100 + op)
            assert np.array_equal(cde.detect_batch(item_id, t), w.detect_batch(item_id, t)), f"{name}: detect differs after update {op}"
            assert np.array_equal(cde.collect_batch(item_id, t, words), w.collect_batch(item_id, t, words)), f"{name}: collect differs after update {op}"

    left = place_remove_sequence(jg, wl, cde, w, n_ops=40, seed=11, after_each=check)
    assert len(left) > 0 and cde.stats()["layout_uploads"] >= 8
    cde.close()


@pytest.mark.parametrize("asset,samples", [("jakobs1.json", 3000), ("swim.json", 4000)])
def test_batched_lbf_solution_is_feasible(jg, orc, asset, samples):
    """the batched LBF search on the engine (jagua-rs_b200/lbf.py). Judged on the solution, not bit
    for bit: every item is placed, the oracle's Layout::is_feasible accepts the layout, and the strip is not much wider than
    what the sequential CPU restatement of lbf_spp.rs reaches with the same sample budget."""
    import json
    import os

    from jagua_rs_b200 import lbf

    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "assets", asset)
    with open(path) as f:
        inst = json.load(f)
    solver = lbf.BatchedLBF(inst, device=0, seed=1)
    res = solver.solve(n_samples=samples)
    solver.close()
    assert len(res.placements) == sum(it.get("demand", 1) for it in inst["items"])
    w = orc.world_from_instance(inst)
    w.change_strip_width(res.strip_width)
    for p in res.placements:
        w.place_item(p.item_id, float(p.d_transf[0]), float(p.d_transf[1]), float(p.d_transf[2]))
    assert w.is_feasible()
    w2 = orc.world_from_instance(inst)
    w2.lbf_solve_spp(samples, 0.2, 1)
    assert res.strip_width < 1.25 * w2.strip_width(), (res.strip_width, w2.strip_width())


@pytest.mark.parametrize("name", ["swim_half", "shirts_full"])
def test_search_round_matches_host_evaluation(jg, engines, name):
    """jg_search_round (samples drawn on the device, detect kernels, loss reduction on the device) against the same
    round evaluated through the ordinary path: the device's samples are fetched, composed on the host, resolved by
    detect_batch AND by the oracle, and the minimal LBF loss over the free ones is recomputed in f32 on the host."""
    wl, cde, w = engines(name)
    b = wl.cde_bbox
    for item_id in (0, len(wl.items) - 1):
        it = wl.items[item_id]
        rots = it.rotations if it.rotations is not None and len(it.rotations) else np.zeros(1, np.float32)
        cde.set_item_rotations(item_id, rots)
        cs = jg.compose(np.stack([rots, np.zeros_like(rots), np.zeros_like(rots)], 1))[:, :2]
        n = 100_000
        rounds = [dict(box=b, seed=5, counter=1), dict(box=[b[0], b[1], (b[0] + b[2]) / 2, b[3]], seed=5, counter=2, rot_idx=0)]
        first = cde.search_round(item_id, n, **rounds[0])
        if first["index"] >= 0:
            rounds.append(dict(mean=(first["tx"], first["ty"]), sd=0.01 * float(b[2] - b[0]), rot_idx=first["rot_idx"], seed=5, counter=3))
        for kw in rounds:
            got = cde.search_round(item_id, n, **kw)
            ridx, txy, verdicts = cde.search_samples(n)
            assert ridx.max() < len(rots)
            if "box" in kw:
                bx = kw["box"]
                assert txy[:, 0].min() >= bx[0] and txy[:, 0].max() <= bx[2] and txy[:, 1].min() >= bx[1] and txy[:, 1].max() <= bx[3]
                if kw.get("rot_idx", -1) < 0 and len(rots) > 1:
                    assert len(np.unique(ridx)) == len(rots)
            else:
                assert np.all(ridx == kw["rot_idx"]) and abs(float(txy[:, 0].mean()) - kw["mean"][0]) < 0.05 * kw["sd"] * 10
            t4 = np.ascontiguousarray(np.concatenate([cs[ridx], txy], 1).astype(np.float32))
            assert np.array_equal(verdicts, cde.detect_batch(item_id, t4))
            assert np.array_equal(verdicts, w.detect_batch(item_id, t4))
            free = np.nonzero(verdicts == 0)[0]
            assert got["n_free"] == len(free)
            if len(free) == 0:
                assert got["index"] == -1
                continue
            v = it.verts
            c, s = t4[free, 0][:, None], t4[free, 1][:, None]
            x = ((c * v[None, :, 0] + (-s) * v[None, :, 1]).astype(np.float32) + t4[free, 2][:, None]).max(1)
            y = ((s * v[None, :, 0] + c * v[None, :, 1]).astype(np.float32) + t4[free, 3][:, None]).max(1)
            cost = (x * np.float32(10.0) + y).astype(np.float32)
            assert np.float32(got["loss"]) == cost.min()
            k = got["index"]
            assert verdicts[k] == 0 and cost[np.searchsorted(free, k)] == cost.min()
            assert (got["rot_idx"], np.float32(got["tx"]), np.float32(got["ty"])) == (int(ridx[k]), txy[k, 0], txy[k, 1])


@pytest.mark.parametrize("name", ["swim_half", "baldacci1_bin", "gardeyn1c_half", "albano_ms0"])
def test_split_containment_kernel_matches_inline_path(jg, engines, name):
    """Large launches run detect_containment_collision in k_s2_contain (teams of four lanes per survivor) instead of
    inside k_s2_grid; forced here for small batches: same verdicts as the oracle, with and without filters."""
    wl, cde, w = engines(name)
    fbits = cde.filter_from_irrelevant_qzones(2)
    cde.upload_filter(7, fbits)
    try:
        cde.set_split_min(0)
        cde.detect_batch(0, jg.make_queries(wl, 0, 1000, seed=1))  # (a fresh engine's first call also uploads the layout)
        l0 = cde.stats()["launches"]
        cde.detect_batch(0, jg.make_queries(wl, 0, 1000, seed=1))
        assert cde.stats()["launches"] - l0 == 3  # k_s1_grid, k_s2_grid<split>, k_s2_contain
        n_contained = 0
        for item_id in range(0, len(wl.items), max(1, len(wl.items) // 4)):
            t = jg.make_queries(wl, item_id, 6000, seed=77)
            exp = w.detect_batch(item_id, t)
            assert np.array_equal(cde.detect_batch(item_id, t), exp)
            assert np.array_equal(cde.detect_batch(item_id, t, poly_only=True), w.detect_batch(item_id, t, poly_only=True))
            assert np.array_equal(cde.detect_batch(item_id, t, filter_id=7), w.detect_batch(item_id, t, filter_bits=fbits))
            n_contained += int((exp == 0).sum())
        assert n_contained > 0 or name == "baldacci1_bin"  # (uniform samples over that bin are almost never free)
    finally:
        cde.set_split_min(5 << 17)  # back to the default


@pytest.mark.parametrize("n", [3000, 200_000, 1_300_000])
def test_collect_group_sizes_match_oracle(jg, engines, n):
    """k_collect_grid resolves 4, 8 or 16 queries per warp group depending on the launch size (small launches are bound by
    a warp's chain through one group): the three regimes against the oracle, on the layout with the most hazards."""
    import os

    wl, cde, w = engines("shirts_full")
    words = cde.collector_words()
    item_id = len(wl.items) - 1
    t = jg.make_queries(wl, item_id, n, seed=123)
    got = cde.collect_batch(item_id, t, words)
    exp = w.collect_batch(item_id, t, words, nthreads=min(16, os.cpu_count() or 1))
    bad = np.nonzero((got != exp).any(axis=1))[0]
    assert len(bad) == 0, f"n={n}: {len(bad)} hazard-set mismatches, first {bad[:3]}"


@pytest.mark.parametrize("name", ["swim_half", "trousers_full"])
def test_generated_piers_on_the_fail_fast_path(jg, orc, name):
    """Items whose surrogate carries piers generated by jg_item_generate_piers (fail_fast/piers.rs; n_ff_piers = 2):
    detect_surrogate_collision tests the poles, then the piers (cd_engine.rs:204-223), collect_surrogate_collisions
    likewise (cd_engine.rs:296-310). Batches with piers stay on the grid kernels (k_s1_grid<.., PIERS> and the pier stage
    of k_collect_grid: an edge entity from the root); the walk kernels are checked beside them.
    Verdict bytes and hazard sets against the oracle; a query the pole-only surrogate reports colliding stays colliding,
    and piers move queries from the polygon stage to the surrogate stage."""
    import dataclasses

    wl = jg.load_workload(name)
    base_cde = jg.build_engine(wl, device=0)
    items = []
    for it in wl.items:
        piers = jg.generate_piers(it.verts, it.poles[:wl.n_ff_poles], 2, 24, 10, 24)
        items.append(dataclasses.replace(it, piers=piers))
    wlp = dataclasses.replace(wl, items=items, n_ff_piers=2)
    cde = jg.build_engine(wlp, device=0)
    w = oracle_world(orc, wlp)
    try:
        words = cde.collector_words()
        moved = 0
        for item_id in range(0, len(wlp.items), max(1, len(wlp.items) // 5)):
            t = jg.make_queries(wlp, item_id, 20000, seed=31)
            got, exp = cde.detect_batch(item_id, t), w.detect_batch(item_id, t)
            assert cde.detect_path() == "grid"
            assert np.array_equal(got, exp), f"{name} item {item_id}: {(got != exp).sum()} mismatches"
            fb = cde.filter_from_irrelevant_qzones(2)
            cde.upload_filter(3, fb)
            assert np.array_equal(cde.detect_batch(item_id, t, filter_id=3), w.detect_batch(item_id, t, filter_bits=fb))
            cde.set_detect_path(True)
            try:
                assert np.array_equal(cde.detect_batch(item_id, t), exp)
                assert np.array_equal(cde.collect_batch(item_id, t[:4000], words), w.collect_batch(item_id, t[:4000], words))
            finally:
                cde.set_detect_path(False)
            plain = base_cde.detect_batch(item_id, t)
            assert not ((plain != 0) & (got == 0)).any() and not ((got == 2) & (plain == 1)).any()
            moved += int(((got == 1) & (plain == 2)).sum())
            gc, ec = cde.collect_batch(item_id, t[:4000], words), w.collect_batch(item_id, t[:4000], words)
            assert np.array_equal(gc, ec)
        assert moved > 0  # the piers caught collisions the two poles missed
    finally:
        cde.close()
        base_cde.close()


@pytest.mark.parametrize("name", ["swim_half", "baldacci1_bin", "shirts_full"])
def test_dependent_launch_matches_stream_order(jg, engines, name):
    """Stage 2 launched with programmatic stream serialization (its CTAs move in while stage 1 drains and wait in
    griddepcontrol.wait): same verdicts as the oracle and as the ordinary stream order, single batches, multi-segment
    batches through host buffers (pieces on several streams) and the split containment path."""
    wl, cde, w = engines(name)
    fbits = cde.filter_from_irrelevant_qzones(2)
    cde.upload_filter(7, fbits)
    was = cde._lib.jg_dependent_launch()
    try:
        cde.set_dependent_launch(True)
        assert cde._lib.jg_dependent_launch() == 1
        for item_id in range(0, len(wl.items), max(1, len(wl.items) // 3)):
            for n in (1, 33, 20000):
                t = jg.make_queries(wl, item_id, n, seed=5 + n)
                exp = w.detect_batch(item_id, t)
                assert np.array_equal(cde.detect_batch(item_id, t), exp)
                assert np.array_equal(cde.detect_batch(item_id, t, filter_id=7), w.detect_batch(item_id, t, filter_bits=fbits))
        ids = list(range(len(wl.items)))
        ts = [jg.make_queries(wl, i, 30000, seed=900 + i) for i in ids]
        exp = np.concatenate([w.detect_batch(i, t) for i, t in zip(ids, ts)])
        offs = np.arange(len(ids) + 1, dtype=np.int64) * 30000
        tall = np.ascontiguousarray(np.concatenate(ts))
        for rep in range(3):
            assert np.array_equal(cde.detect_multi(ids, offs, tall), exp)
        cde.set_split_min(0)
        assert np.array_equal(cde.detect_multi(ids, offs, tall), exp)
        cde.set_dependent_launch(False)
        assert np.array_equal(cde.detect_multi(ids, offs, tall), exp)
    finally:
        cde.set_split_min(5 << 17)  # back to the default
        cde.set_dependent_launch(bool(was))


@pytest.mark.parametrize("name", ["swim_half", "jakobs1_ms1"])
def test_layout_update_behind_batches_in_flight(jg, orc, name):
    """Stream-ordered layout uploads: a batch enqueued before an update is resolved on the old layout, a batch enqueued
    after it on the new one, with no host synchronisation in between (fences only order the host's reads)."""
    wl = jg.load_workload(name)
    cde = jg.build_engine(wl, device=0)
    w = oracle_world(orc, wl)
    cde.sync_layout()
    n_items = len(wl.items)
    n = 40_000
    sizes = jg.split_batch(n, n_items)
    offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    ids = list(range(n_items))
    t = jg.PinnedBuffer((n, 4), np.float32)
    for i, m in enumerate(sizes):
        t.array[offs[i]:offs[i + 1]] = jg.make_queries(wl, i, m, seed=9)
    outs = [jg.PinnedBuffer((n,), np.uint8) for _ in range(8)]
    rng = np.random.default_rng(4)
    b = wl.cde_bbox
    expected, fences = [], []
    uploads0 = cde.stats()["layout_uploads"]
    for step in range(8):
        cde.detect_multi(ids, offs, t.array, out=outs[step].array, wait=False)  # in flight while the layout changes
        fences.append(cde.fence())
        expected.append(np.concatenate([w.detect_batch(i, np.ascontiguousarray(t.array[offs[i]:offs[i + 1]])) for i in range(n_items)]))
        i = int(rng.integers(n_items))
        it = wl.items[i]
        rot = 0.0 if it.rotations is None or len(it.rotations) == 0 else float(it.rotations[0])
        dt = np.array([[rot, b[0] + (b[2] - b[0]) * rng.random(), b[1] + (b[3] - b[1]) * rng.random()]], np.float32)
        pk = w.place_item(i, float(dt[0, 0]), float(dt[0, 1]), float(dt[0, 2]))
        cde.place_item(i, jg.ItemShape(it.verts, it.poi, it.poles, it.piers), jg.compose(dt)[0], pk)
        cde.sync_layout()  # returns without waiting for the batch above
    for step in (3, 0, 7):  # fences can be waited for in any order
        cde.wait_fence(fences[step])
        assert np.array_equal(outs[step].array, expected[step]), f"step {step}"
    cde.sync()
    for step in range(8):
        assert np.array_equal(outs[step].array, expected[step]), f"step {step}"
    assert len({e.tobytes() for e in expected}) > 1  # the updates really changed verdicts
    assert cde.stats()["layout_uploads"] - uploads0 == 8
    cde.close()


def test_batched_lbf_bins_solution_is_feasible(jg, orc):
    """lbf_bpp.rs on the engine (bins with holes and quality zones, items searched under the from_irrelevant_qzones
    filter of their min_quality): every layout is re-placed in an oracle layout of the same bin and must be feasible
    (Layout::is_feasible checks every placed item against holes, zones below its min_quality and the other items)."""
    import json
    import os

    from jagua_rs_b200 import lbf

    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "assets", "baldacci1.json")
    with open(path) as f:
        inst = json.load(f)
    solver = lbf.BatchedLBFBins(inst, device=0, seed=2)
    res = solver.solve(n_samples=4000, max_items=30)
    assert res.placed == 30 and len(res.layouts) >= 2
    assert any(solver.items[p.item_id].min_quality is not None for lay in res.layouts for p in lay.placements)
    for lay in res.layouts:
        w = orc.world_from_instance(inst, bin_index=lay.bin_id)
        for p in lay.placements:
            w.place_item(p.item_id, float(p.d_transf[0]), float(p.d_transf[1]), float(p.d_transf[2]))
        assert w.is_feasible(), f"bin {lay.bin_id}"
    solver.close()


@pytest.mark.parametrize("name", ["swim_half", "shirts_full_d10"])
def test_non_finite_queries_terminate_and_do_not_disturb_others(jg, engines, name):
    """The reference cannot even build such a Transformation (NotNan panics); the engine must still terminate on them and
    every ordinary query of the same batch must keep its verdict and hazard set (grid and walk kernels)."""
    wl, cde, w = engines(name)
    words = cde.collector_words()
    for item_id in (0, len(wl.items) - 1):
        t = jg.make_queries(wl, item_id, 4096, seed=123)
        exp = w.detect_batch(item_id, t)
        expc = w.collect_batch(item_id, t, words)
        bad = t.copy()
        rows = np.arange(0, len(t), 7)
        vals = [np.nan, np.inf, -np.inf, 3.0e38, -3.0e38, 1.0e-45]
        for k, r in enumerate(rows):
            bad[r, (k // len(vals)) % 4] = vals[k % len(vals)]
        keep = np.ones(len(t), bool)
        keep[rows] = False
        got = cde.detect_batch(item_id, bad)
        assert np.array_equal(got[keep], exp[keep])
        gotc = cde.collect_batch(item_id, bad, words)
        assert np.array_equal(gotc[keep], expc[keep])
