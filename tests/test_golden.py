"""Committed golden vectors (tests/golden/, written by tools/make_golden.py from the oracle; regression
pins, not reference outputs — see DESIGN.md §6): the oracle must keep reproducing them on CPU, and the
GPU path must reproduce them through the C ABI."""
import os

import numpy as np
import pytest

from helpers import ALL_WORKLOADS, oracle_world

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("name", ALL_WORKLOADS)
def test_oracle_reproduces_golden(jg, orc, name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    wl = jg.load_workload(name)
    w = oracle_world(orc, wl)
    words = g["collect"].shape[2]
    for i in range(len(wl.items)):
        t = g["transforms"][i]
        assert np.array_equal(jg.make_queries(wl, i, len(t), int(g["seed"])), t)  # the query generator is pinned too
        assert np.array_equal(w.detect_batch(i, t), g["detect"][i])
        assert np.array_equal(w.detect_batch(i, t, poly_only=True), g["detect_poly_only"][i])
        assert np.array_equal(w.collect_batch(i, t, words), g["collect"][i])
        assert np.array_equal(w.detect_batch(i, t, filter_bits=g["filter_bits"][i]), g["detect_filtered"][i])


@pytest.mark.gpu
@pytest.mark.parametrize("name", ALL_WORKLOADS)
def test_gpu_reproduces_golden(jg, name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    wl = jg.load_workload(name)
    cde = jg.build_engine(wl, device=0)
    words = g["collect"].shape[2]
    try:
        for i in range(len(wl.items)):
            t = np.ascontiguousarray(g["transforms"][i])
            assert np.array_equal(cde.detect_batch(i, t), g["detect"][i])
            assert np.array_equal(cde.detect_batch(i, t, poly_only=True), g["detect_poly_only"][i])
            assert np.array_equal(cde.collect_batch(i, t, words), g["collect"][i])
            cde.upload_filter(i, g["filter_bits"][i])
            assert np.array_equal(cde.detect_batch(i, t, filter_id=i), g["detect_filtered"][i])
    finally:
        cde.close()
