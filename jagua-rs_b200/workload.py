"""Workload fixtures: a layout (CDE bbox + hazards), the instance's item shapes and seeded query
batches, in the shape the reference benches use them (lbf/benches/ci_bench.rs:27-167,
lbf/benches/util.rs:18-83): items imported from an instance file, a layout produced by LBF with some
items removed, queries sampled uniformly over the CDE bbox with a rotation from the item's allowed set.

A fixture is an .npz written by tools/make_workloads.py (which runs the importer + LBF filler once,
in the build container) so that bench.py and the GPU tests need neither the reference tree nor any
CPU implementation of the path to build their inputs.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

WORKLOAD_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "workloads")


@dataclass
class WItem:
    verts: np.ndarray      # (V, 2)
    poi: np.ndarray        # (3,) x, y, r
    poles: np.ndarray      # (P, 3)
    piers: np.ndarray      # (Q, 4)
    rotations: Optional[np.ndarray]  # radians; None = continuous
    min_quality: int = -1


@dataclass
class WHazard:
    kind: int
    id: int
    quality: int
    pk: int
    dynamic: bool
    verts: np.ndarray      # (V, 2) already in layout coordinates
    poi: np.ndarray        # (3,)


@dataclass
class Workload:
    name: str
    quadtree_depth: int
    cd_threshold: int
    n_ff_poles: int
    n_ff_piers: int
    cde_bbox: np.ndarray
    items: List[WItem] = field(default_factory=list)
    hazards: List[WHazard] = field(default_factory=list)

    def save(self, path: str):
        def cat(arrs, width):
            arrs = [np.asarray(a, np.float32).reshape(-1, width) for a in arrs]
            return np.concatenate(arrs, 0) if arrs else np.zeros((0, width), np.float32)

        np.savez_compressed(
            path,
            name=np.array(self.name),
            config=np.array([self.quadtree_depth, self.cd_threshold, self.n_ff_poles, self.n_ff_piers], np.int32),
            cde_bbox=np.asarray(self.cde_bbox, np.float32),
            item_nverts=np.array([len(i.verts) for i in self.items], np.int32),
            item_verts=cat([i.verts for i in self.items], 2),
            item_poi=cat([i.poi for i in self.items], 3),
            item_npoles=np.array([len(i.poles) for i in self.items], np.int32),
            item_poles=cat([i.poles for i in self.items], 3),
            item_npiers=np.array([len(i.piers) for i in self.items], np.int32),
            item_piers=cat([i.piers for i in self.items], 4),
            item_nrot=np.array([-1 if i.rotations is None else len(i.rotations) for i in self.items], np.int32),
            item_rots=np.concatenate([np.asarray(i.rotations, np.float32) for i in self.items if i.rotations is not None]
                                     + [np.zeros(0, np.float32)]),
            item_min_quality=np.array([i.min_quality for i in self.items], np.int32),
            haz_meta=np.array([[h.kind, h.id, h.quality, int(h.dynamic), len(h.verts)] for h in self.hazards], np.int64).reshape(-1, 5),
            haz_pk=np.array([h.pk for h in self.hazards], np.uint64),
            haz_verts=cat([h.verts for h in self.hazards], 2),
            haz_poi=cat([h.poi for h in self.hazards], 3),
        )


def load_workload(name_or_path: str) -> Workload:
    path = name_or_path if os.path.exists(name_or_path) else os.path.join(WORKLOAD_DIR, name_or_path + ".npz")
    z = np.load(path)
    cfg = z["config"]
    w = Workload(str(z["name"]), int(cfg[0]), int(cfg[1]), int(cfg[2]), int(cfg[3]), z["cde_bbox"].astype(np.float32))
    vo = po = qo = ro = 0
    for i in range(len(z["item_nverts"])):
        nv, npol, npier, nrot = int(z["item_nverts"][i]), int(z["item_npoles"][i]), int(z["item_npiers"][i]), int(z["item_nrot"][i])
        rots = None if nrot < 0 else z["item_rots"][ro:ro + nrot].copy()
        w.items.append(WItem(z["item_verts"][vo:vo + nv].copy(), z["item_poi"][i].copy(), z["item_poles"][po:po + npol].copy(),
                             z["item_piers"][qo:qo + npier].copy(), rots, int(z["item_min_quality"][i])))
        vo, po, qo, ro = vo + nv, po + npol, qo + npier, ro + max(nrot, 0)
    vo = 0
    for i, m in enumerate(z["haz_meta"]):
        nv = int(m[4])
        w.hazards.append(WHazard(int(m[0]), int(m[1]), int(m[2]), int(z["haz_pk"][i]), bool(m[3]),
                                 z["haz_verts"][vo:vo + nv].copy(), z["haz_poi"][i].copy()))
        vo += nv
    return w


def build_engine(w: Workload, device: Optional[int] = 0):
    """CDEngine::new + register_hazard for every hazard of the fixture, items uploaded."""
    from . import CDEngine, ItemShape

    cde = CDEngine(w.cde_bbox, w.quadtree_depth, w.cd_threshold, w.n_ff_poles, w.n_ff_piers, device=device)
    keys = []
    for h in w.hazards:
        keys.append(cde.register_hazard(h.kind, h.verts, h.poi, id=h.id, quality=h.quality, pk=h.pk, dynamic=h.dynamic))
    if device is not None:
        for i, it in enumerate(w.items):
            cde.upload_item(i, ItemShape(it.verts, it.poi, it.poles, it.piers))
    cde.hazard_keys = keys
    return cde


def make_queries_indexed(w: Workload, item_id: int, n: int, seed: int = 0):
    """The same n seeded queries as make_queries in their compact form: (rot_idx u8, txy (n, 2) f32). Only for
    items with a discrete set of allowed orientations."""
    rng = np.random.Generator(np.random.Philox(key=[seed, item_id]))
    it = w.items[item_id]
    b = w.cde_bbox.astype(np.float32)
    u = rng.random((n, 2), dtype=np.float32)
    tx = (b[0] + (b[2] - b[0]) * u[:, 0]).astype(np.float32)
    ty = (b[1] + (b[3] - b[1]) * u[:, 1]).astype(np.float32)
    assert it.rotations is not None and len(it.rotations) > 0, "item has no discrete rotation set"
    idx = rng.integers(0, len(it.rotations), n)
    return idx.astype(np.uint8), np.ascontiguousarray(np.stack([tx, ty], 1))


def make_queries_lattice(w: Workload, item_id: int, n: int, seed: int = 0, steps: int = 65536):
    """n seeded queries of one item on the 65536 x 65536 sampling lattice over the CDE bbox (the 5-byte query form):
    (lattice, rot_idx u8, ij (n, 2) u16). The same uniform distribution as make_queries, quantised; the translations
    are lattice_points(lattice, ij)."""
    from . import JgLattice

    rng = np.random.Generator(np.random.Philox(key=[seed, item_id]))
    it = w.items[item_id]
    assert it.rotations is not None and len(it.rotations) > 0, "item has no discrete rotation set"
    lat = JgLattice.over_box(w.cde_bbox, steps)
    ij = rng.integers(0, steps, (n, 2)).astype(np.uint16)
    idx = rng.integers(0, len(it.rotations), n)
    return lat, idx.astype(np.uint8), np.ascontiguousarray(ij)


def compose_lattice_queries(w: Workload, item_id: int, lattice, rot_idx: np.ndarray, ij: np.ndarray, compose_fn=None) -> np.ndarray:
    """The (cos, sin, tx, ty) records a lattice batch stands for (checker side of the 5-byte form)."""
    from . import compose, lattice_points

    it = w.items[item_id]
    txy = lattice_points(lattice, ij)
    rot = np.asarray(it.rotations, np.float32)[np.minimum(rot_idx, len(it.rotations) - 1)]
    return (compose_fn or compose)(np.stack([rot, txy[:, 0], txy[:, 1]], 1))


def make_queries(w: Workload, item_id: int, n: int, seed: int = 0, compose_fn=None) -> np.ndarray:
    """n seeded (cos, sin, tx, ty) queries for one item: translation uniform over the CDE bbox
    (UniformRectSampler::new(cde.bbox(), item), lbf/src/samplers/uniform_rect_sampler.rs:18-36;
    lbf/benches/ci_bench.rs:55,106), rotation uniform over the item's allowed orientations
    (rotation_distr.rs:32-49). Counter-based PRNG (Philox), not the reference's SmallRng stream."""
    if compose_fn is None:  # (bench.py's reference arm passes the oracle's compose: it must not load the product library)
        from . import compose as compose_fn

    rng = np.random.Generator(np.random.Philox(key=[seed, item_id]))
    it = w.items[item_id]
    b = w.cde_bbox.astype(np.float32)
    u = rng.random((n, 2), dtype=np.float32)
    tx = (b[0] + (b[2] - b[0]) * u[:, 0]).astype(np.float32)
    ty = (b[1] + (b[3] - b[1]) * u[:, 1]).astype(np.float32)
    if it.rotations is None:
        rot = (rng.random(n, dtype=np.float32) * np.float32(2.0 * np.pi)).astype(np.float32)
    elif len(it.rotations) == 0:
        rot = np.zeros(n, np.float32)
    else:
        rot = np.asarray(it.rotations, np.float32)[rng.integers(0, len(it.rotations), n)]
    return compose_fn(np.stack([rot, tx, ty], 1))


def split_batch(total: int, n_items: int) -> List[int]:
    """Per-item query counts of a `total`-query batch spread evenly over the item types."""
    base, rem = divmod(int(total), int(n_items))
    return [base + (1 if i < rem else 0) for i in range(n_items)]
