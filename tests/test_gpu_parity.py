"""GPU (C ABI -> sm_100a kernels) vs the CPU oracle on the same seeded inputs: verdict bytes and
order-normalised hazard bitsets must be bit-identical for every query."""
import numpy as np
import pytest

from helpers import ALL_WORKLOADS, oracle_world, verdict_summary

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
    assert total["free"] > 0 and total["poly"] > 0, total  # the batch exercises every stage


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
