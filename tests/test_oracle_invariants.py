"""The reference ships no golden vectors for collision queries; what it does ship are structural
invariants (debug assertions, util/assertions.rs) exercised by lbf/tests/tests.rs. They are restated here
against the oracle, together with primitive-level properties and a brute-force cross-check."""
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ASSETS = os.path.join(ROOT, "assets")


def test_approx_eq_is_float_cmp_default_margin(orc):
    L = orc.lib()
    f = np.float32
    assert L.orc_approx_eq(1.0, 1.0)
    assert L.orc_approx_eq(f(1.0), f(1.0) + f(1.1920929e-07))            # |a-b| <= f32::EPSILON
    assert L.orc_approx_eq(f(1.0), f(1.0) + f(3 * 1.1920929e-07))         # 3 ulps <= 4
    assert not L.orc_approx_eq(f(1.0), f(1.0) + f(6 * 1.1920929e-07))     # 6 ulps, |a-b| > eps
    big = f(1.0e6)
    assert L.orc_approx_eq(big, np.nextafter(np.nextafter(big, f(np.inf)), f(np.inf)))   # 2 ulps
    x = big
    for _ in range(5):
        x = np.nextafter(x, f(np.inf))
    assert not L.orc_approx_eq(big, x)                                    # 5 ulps, |a-b| = 0.3 > eps
    assert L.orc_approx_eq(f(0.0), f(-0.0))
    assert not L.orc_approx_eq(f(1.0), f(-1.0))


def test_edge_quadrants_is_superset_of_individual_tests(orc):
    """qt_traits.rs:87-102: every quadrant that collides individually is flagged (N1: may flag more)."""
    L = orc.lib()
    rng = np.random.default_rng(5)
    r = np.array([0, 0, 8, 8], np.float32)
    strict = 0
    for _ in range(20000):
        e = rng.uniform(-4, 12, 4).astype(np.float32)
        if rng.random() < 0.3:  # snap to the grid lines to provoke the boundary cases
            e = np.round(e / 4) * 4
            e = e.astype(np.float32)
        if e[0] == e[2] and e[1] == e[3]:
            continue
        m, ind = L.orc_edge_quadrants(e, r), L.orc_edge_quadrants_individual(e, r)
        assert (ind & ~m) == 0, (e, m, ind)
        strict += m != ind
    assert strict >= 0


def test_point_in_polygon_vertex_rule(orc):
    L = orc.lib()
    sq = np.array([[0, 0], [4, 0], [4, 4], [0, 4]], np.float32)
    assert L.orc_pip(sq, 4, 2.0, 2.0) and not L.orc_pip(sq, 4, 5.0, 2.0)
    dia = np.array([[0, 2], [2, 0], [4, 2], [2, 4]], np.float32)  # ray through the vertex (4, 2)
    assert L.orc_pip(dia, 4, 2.0, 2.0) and L.orc_pip(dia, 4, 1.0, 2.0)
    assert not L.orc_pip(dia, 4, 0.5, 0.5)


@pytest.mark.parametrize("asset", ["swim.json", "shirts.json", "trousers.json", "albano.json", "jakobs1.json", "baldacci1.json"])
@pytest.mark.parametrize("depth", [0, 3, 10])
def test_solve_remove_save_restore_keeps_invariants(orc, asset, depth):
    """lbf/tests/tests.rs:20-131 (n_samples 100, depths [0, 3, 10]) with the debug assertions as checks."""
    inst = orc.load_instance_json(os.path.join(ASSETS, asset))
    w = orc.world_from_instance(inst, quadtree_depth=depth)
    fill = (lambda: w.lbf_solve_spp(100, 0.2, 0)) if "strip_height" in inst else (lambda: w.lbf_fill_bin(100, 0.2, 0))

    def check(feasible=True):
        assert w.check_invariants() == 0        # sortedness, n_active_edges, no dangling hazards, children rule
        assert w.matches_fresh() == 0           # assertions.rs:91-104 layout_qt_matches_fresh_qt
        if feasible:
            assert w.is_feasible()

    n = fill()
    assert n > 0
    check()
    pks, _, _ = w.placed()
    rng = np.random.default_rng(0)
    for pk in rng.choice(pks, size=min(5, len(pks)), replace=False):
        w.remove_item(int(pk))
        check()
    snap = w.save()
    n_saved = len(w.placed()[0])
    pks, _, _ = w.placed()
    for pk in pks[:3]:
        w.remove_item(int(pk))
    check()
    w.restore(snap)
    assert len(w.placed()[0]) == n_saved
    check()


@pytest.mark.parametrize("name", ["swim_half", "shirts_full", "baldacci1_bin"])
def test_quadtree_agrees_with_brute_force(jg, orc, name):
    """Depth 0 (every edge against every hazard edge) is itself a reference configuration
    (lbf/tests/tests.rs:18); the quadtree at the workload's depth must give the same verdicts and sets."""
    from helpers import oracle_world
    wl = jg.load_workload(name)
    w = oracle_world(orc, wl)
    wl0 = jg.load_workload(name)
    wl0.quadtree_depth = 0
    w0 = oracle_world(orc, wl0)
    words = max(1, (w.haz_capacity() + 31) // 32)
    for item_id in range(len(wl.items)):
        t = jg.make_queries(wl, item_id, 3000, seed=9)
        a, b = w.detect_batch(item_id, t, poly_only=True), w0.detect_batch(item_id, t, poly_only=True)
        assert np.array_equal(a, b), f"item {item_id}: {(a != b).sum()} verdict differences vs brute force"
        ca, cb = w.collect_batch(item_id, t, words, poly_only=True), w0.collect_batch(item_id, t, words, poly_only=True)
        assert np.array_equal(ca, cb), f"item {item_id}: {(ca != cb).any(axis=1).sum()} set differences vs brute force"
