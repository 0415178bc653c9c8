"""Batched left-bottom-fill on top of the engine: the caller side of the hot path.

Mirrors the reference's LBF strip-packing optimizer (lbf/src/opt/lbf_spp.rs:37-123, lbf/src/opt/search.rs:17-115,
lbf/src/opt/loss.rs:12-52, lbf/src/samplers/{uniform_rect_sampler,ls_sampler}.rs) with one structural change: the
reference evaluates its samples one by one, each decision (worth_testing, bbox tightening, mean shift) depending on
the previous sample; here every decision point takes a whole BATCH of samples through `CDEngine.detect_batch`
(surrogate + polygon stage on the GPU) and keeps the feasible sample of minimal loss. The sample order therefore
differs from the reference and results are compared on solution quality (strip width / density, feasibility checked
by the oracle in tests/), not bit for bit:

  * uniform phase: `uniform_rounds` batches over the CDE bbox, the sampling box tightened to x_max <= cost / 10
    after every round (LBFLoss::tighten_sample_bbox);
  * local search: `ls_rounds` batches drawn from normal distributions around the best sample, the standard
    deviations decaying from 1 % to 0.05 % of the bbox (translation) and 2 deg to 0.5 deg (continuous rotation)
    along init * (end / init)^progress (LSSampler::decay_stddev), the mean shifted to every improvement;
  * items with a discrete set of allowed rotations (or none) run every round entirely on the device
    (`CDEngine.search_round`: Philox sampling kernel, detect kernels, loss reduction kernel; 32 bytes come back per
    round); the (cos, sin) of a sample must come from the host libm to stay bit-compatible with the reference's
    `Transformation`s, so items that rotate continuously are searched on a lattice of `rotation_lattice` (256) evenly
    spaced angles composed once on the host - every sample is still an exact transformation, only the sampling
    distribution of the rotation is discretised - or, with `rotation_lattice=0`, sampled on the host as before;
  * no feasible sample: the strip grows by 10 % (change_strip_width) - a new CDEngine, all placements re-registered;
  * at the end the strip is fitted to the right-most item * 1.00001 (SPProblem::fit_strip).

Items come from an instance file through jg_item_import (importer cleaning, centring, surrogate poles, product C++),
placements go through jg_cde_place_item, layouts through the stream-ordered upload. Nothing here touches oracle/.
"""
from __future__ import annotations

import json
import time
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

from . import HAZ_EXTERIOR, CDEngine, compose, polygon_poi
from .instance import InstItem as LbfItem, import_items

X_MULTIPLIER = np.float32(10.0)  # lbf/src/opt/loss.rs:5


@dataclass
class Placement:
    item_id: int
    d_transf: np.ndarray  # (rotation, tx, ty) f32
    pk: int


@dataclass
class LbfResult:
    strip_width: float
    strip_height: float
    density: float
    placements: List[Placement] = field(default_factory=list)
    samples: int = 0
    seconds: float = 0.0
    strip_extensions: int = 0


class BatchedLBF:
    def __init__(self, instance: dict, device: int = 0, quadtree_depth: int = 5, cd_threshold: int = 16, n_ff_poles: int = 2,
                 seed: int = 0, device_sampling: bool = True, rotation_lattice: int = 256, simplify_tolerance=None,
                 narrow_concavity_cutoff_ratio=None):
        self.rotation_lattice = int(rotation_lattice)
        self.items = import_items(instance["items"], simplify_tolerance, narrow_concavity_cutoff_ratio)
        self.height = np.float32(instance["strip_height"])
        area = np.float32(0.0)
        for it in self.items:  # probs/spp/io/import.rs:40-48: f32 sum of area * demand
            area = np.float32(area + np.float32(np.float32(it.area) * np.float32(it.demand)))
        self.total_area = float(area)
        self.width = np.float32(area / self.height)
        self.cfg = dict(quadtree_depth=quadtree_depth, cd_threshold=cd_threshold, n_ff_poles=n_ff_poles, n_ff_piers=0)
        self.device = device
        self.rng = np.random.Generator(np.random.Philox(key=[seed, 0xF4]))
        self.seed = int(seed)
        self.device_sampling = device_sampling
        self._round = 0
        self.placements: List[Placement] = []
        self.cde: Optional[CDEngine] = None
        self.samples = 0
        self.extensions = 0
        self._next_pk = 1
        self.filter_id = -1  # HazardFilter of the current search (-1 = NoFilter)
        self._build_engine()

    # -- container (probs/spp/entities/strip.rs:44-61: rectangle (0, 0, width, height); CDE bbox = its square hull) ----
    def _build_engine(self):
        """The strip's CDE. First call: a new engine; afterwards (change_strip_width, fit_strip: problem.rs:27-57)
        Layout::swap_container - the host mirror becomes a clone of the new strip's base CDE with the placed items
        re-registered; the device engine, the uploaded items and their rotation tables stay as they are."""
        w, h = np.float32(self.width), self.height
        rect = np.array([[0, 0], [w, 0], [w, h], [0, h]], np.float32)
        dx = dy = np.float32(0.0)
        if h < w:
            dy = np.float32((w - h) / np.float32(2.0))
        elif w < h:
            dx = np.float32((h - w) / np.float32(2.0))
        self.bbox = np.array([np.float32(0.0) - dx, np.float32(0.0) - dy, w + dx, h + dy], np.float32)  # rect.rs:93-109 inflate_to_square
        verts, poi = polygon_poi(rect)
        if self.cde is not None:
            base = CDEngine(self.bbox, device=None, **self.cfg)  # Container::base_cde of the new strip
            base.register_hazard(HAZ_EXTERIOR, verts, poi, dynamic=False)
            self.cde.swap_container(base)
            base.close()
            return
        self.cde = CDEngine(self.bbox, device=self.device, **self.cfg)
        self.cde.register_hazard(HAZ_EXTERIOR, verts, poi, dynamic=False)
        for it in self.items:
            self.cde.upload_item(it.id, it.shape)
            rots = self._device_rotations(it)
            if rots is not None:  # search rounds can run on the device
                self.cde.set_item_rotations(it.id, rots)
        for p in self.placements:
            self.cde.place_item(p.item_id, self.items[p.item_id].shape, compose(p.d_transf.reshape(1, 3))[0], p.pk)

    def _device_rotations(self, item):
        """The rotation set registered with the engine for device-side search rounds (None: host sampling)."""
        if item.rotations is not None:
            return item.rotations if len(item.rotations) else np.zeros(1, np.float32)
        if self.rotation_lattice > 0:
            n = min(self.rotation_lattice, 256)
            return (np.arange(n, dtype=np.float64) * (2.0 * np.pi / n)).astype(np.float32)
        return None

    # -- samplers ------------------------------------------------------------------------------------------------
    def _rotations(self, item: LbfItem, n: int) -> np.ndarray:
        if item.rotations is None:
            return (self.rng.random(n, dtype=np.float32) * np.float32(2.0 * np.pi)).astype(np.float32)
        if len(item.rotations) == 0:
            return np.zeros(n, np.float32)
        return item.rotations[self.rng.integers(0, len(item.rotations), n)]

    def _uniform(self, item: LbfItem, box: np.ndarray, n: int) -> np.ndarray:
        u = self.rng.random((n, 2), dtype=np.float32)
        x = (box[0] + (box[2] - box[0]) * u[:, 0]).astype(np.float32)
        y = (box[1] + (box[3] - box[1]) * u[:, 1]).astype(np.float32)
        return np.stack([self._rotations(item, n), x, y], 1).astype(np.float32)

    def _normal(self, item: LbfItem, mean: np.ndarray, sd_t: float, sd_r: float, n: int) -> np.ndarray:
        x = (mean[1] + sd_t * self.rng.standard_normal(n, dtype=np.float32)).astype(np.float32)
        y = (mean[2] + sd_t * self.rng.standard_normal(n, dtype=np.float32)).astype(np.float32)
        if item.rotations is None:  # NormalRotDistr: normal around the reference rotation
            r = (mean[0] + sd_r * self.rng.standard_normal(n, dtype=np.float32)).astype(np.float32)
        else:  # discrete / none: the rotation of the reference transformation is kept
            r = np.full(n, mean[0], np.float32)
        return np.stack([r, x, y], 1).astype(np.float32)

    # -- one batch through the engine: feasible sample of minimal loss -----------------------------------------------
    def _best_of(self, item: LbfItem, dts: np.ndarray, below: Optional[float]):
        t4 = compose(dts)
        verdicts = self.cde.detect_batch(item.id, t4, filter_id=self.filter_id)
        self.samples += len(dts)
        free = np.nonzero(verdicts == 0)[0]
        if len(free) == 0:
            return None
        v = item.shape.verts
        c, s, tx, ty = (t4[free, k][:, None] for k in range(4))
        # the transformed shape's bbox (SPolygon::transform_from + generate_bounding_box), f32 like the kernels
        x = (c * v[None, :, 0] + (-s) * v[None, :, 1]).astype(np.float32) + tx
        y = (s * v[None, :, 0] + c * v[None, :, 1]).astype(np.float32) + ty
        cost = x.max(1) * X_MULTIPLIER + y.max(1)  # LBFLoss::cost
        k = int(np.argmin(cost))
        if below is not None and not cost[k] < below:
            return None
        return dts[free[k]].copy(), float(cost[k])

    def _search_on_device(self, item: LbfItem, n_uni: int, n_ls: int, uniform_rounds: int, ls_rounds: int):
        rots = self._device_rotations(item)
        box = self.bbox.copy()
        best = None  # (d_transf, loss, rotation index)

        def take(got):
            nonlocal best
            if got["index"] >= 0 and (best is None or got["loss"] < best[1]):
                best = (np.array([rots[got["rot_idx"]], got["tx"], got["ty"]], np.float32), got["loss"], got["rot_idx"])
                return True
            return False

        per = max(n_uni // uniform_rounds, 1)
        for _ in range(uniform_rounds):
            self._round += 1
            self.samples += per
            if take(self.cde.search_round(item.id, per, box=box, seed=self.seed, counter=self._round, filter_id=self.filter_id)):
                box[2] = min(box[2], np.float32(best[1]) / X_MULTIPLIER)  # tighten_sample_bbox
                if not box[2] > box[0]:
                    break
        if best is None:
            return None
        max_dim = float(max(self.bbox[2] - self.bbox[0], self.bbox[3] - self.bbox[1]))
        sd0, sd1 = 0.01 * max_dim, 0.0005 * max_dim  # ls_sampler.rs:12-16
        per = max(n_ls // max(ls_rounds, 1), 1)
        for r in range(ls_rounds):
            self._round += 1
            self.samples += per
            sd = sd0 * (sd1 / sd0) ** (r / max(ls_rounds, 1))
            # items on a rotation lattice also try the neighbouring angles (the reference's LSSampler perturbs the rotation)
            ri = best[2] if item.rotations is not None else (best[2] + (0, 1, -1)[r % 3]) % len(rots)
            take(self.cde.search_round(item.id, per, mean=(best[0][1], best[0][2]), sd=sd, rot_idx=ri, seed=self.seed,
                                       counter=self._round, filter_id=self.filter_id))
        return best[0], best[1]

    def search(self, item: LbfItem, n_samples: int, ls_frac: float, uniform_rounds: int, ls_rounds: int):
        """search.rs:17-106 with batches at the decision points."""
        n_ls = int(n_samples * ls_frac)
        n_uni = n_samples - n_ls
        if self.device_sampling and self._device_rotations(item) is not None:
            return self._search_on_device(item, n_uni, n_ls, uniform_rounds, ls_rounds)
        box = self.bbox.copy()
        best = None
        per = max(n_uni // uniform_rounds, 1)
        for _ in range(uniform_rounds):
            got = self._best_of(item, self._uniform(item, box, per), None if best is None else best[1])
            if got is not None:
                best = got
                box[2] = min(box[2], np.float32(best[1]) / X_MULTIPLIER)  # tighten_sample_bbox
                if not box[2] > box[0]:
                    break
        if best is None:
            return None
        max_dim = float(max(self.bbox[2] - self.bbox[0], self.bbox[3] - self.bbox[1]))
        sd_t0, sd_t1 = 0.01 * max_dim, 0.0005 * max_dim  # ls_sampler.rs:12-16
        sd_r0, sd_r1 = 2.0 * np.pi / 180.0, 0.5 * np.pi / 180.0
        per = max(n_ls // max(ls_rounds, 1), 1)
        for r in range(ls_rounds):
            pct = r / max(ls_rounds, 1)
            sd_t, sd_r = sd_t0 * (sd_t1 / sd_t0) ** pct, sd_r0 * (sd_r1 / sd_r0) ** pct
            got = self._best_of(item, self._normal(item, best[0], sd_t, sd_r, per), best[1])
            if got is not None:
                best = got
        return best

    def solve(self, n_samples: int = 5000, ls_frac: float = 0.2, uniform_rounds: int = 4, ls_rounds: int = 8,
              max_items: int = 1 << 30) -> LbfResult:
        t0 = time.perf_counter()
        order = sorted(self.items, key=lambda it: -it.diameter)  # search.rs:108-115 item_placement_order
        placed = 0
        for item in order:
            demand = item.demand
            guard = 0
            while demand > 0 and placed < max_items:
                got = self.search(item, n_samples, ls_frac, uniform_rounds, ls_rounds)
                if got is None:
                    guard += 1
                    if guard > 100:
                        raise RuntimeError("strip width is running away: does every item fit in the strip's height?")
                    self.width = np.float32(self.width * np.float32(1.1))
                    self.extensions += 1
                    self._build_engine()
                    continue
                dt = got[0]
                pk = self._next_pk
                self._next_pk += 1
                self.cde.place_item(item.id, item.shape, compose(dt.reshape(1, 3))[0], pk)
                self.placements.append(Placement(item.id, dt, pk))
                demand -= 1
                placed += 1
        # problem.rs:39-57 fit_strip
        x_max = -np.inf
        for p in self.placements:
            t4 = compose(p.d_transf.reshape(1, 3))[0]
            v = self.items[p.item_id].shape.verts
            x = (t4[0] * v[:, 0] + (-t4[1]) * v[:, 1]).astype(np.float32) + t4[2]
            x_max = max(x_max, float(x.max()))
        if self.placements:
            self.width = np.float32(np.float32(x_max) * np.float32(1.00001))
            self._build_engine()
        used = sum(self.items[p.item_id].area for p in self.placements)
        return LbfResult(float(self.width), float(self.height), used / float(self.width * self.height), list(self.placements),
                         self.samples, time.perf_counter() - t0, self.extensions)

    def close(self):
        if self.cde is not None:
            self.cde.close()
            self.cde = None


@dataclass
class BinLayout:
    bin_id: int
    cde: CDEngine
    bbox: np.ndarray
    placements: List[Placement] = field(default_factory=list)


@dataclass
class BinResult:
    layouts: List[BinLayout]
    placed: int
    unplaced: int
    samples: int
    seconds: float


class BatchedLBFBins(BatchedLBF):
    """lbf_bpp.rs:39-137 on the engine: items by descending diameter; every copy is searched for in the open layouts in
    the order they were opened, then in the first closed bin type that still has stock (a layout is opened only when a
    placement is found there). Items with a min_quality are searched under HazKeyFilter::from_irrelevant_qzones
    (filter.rs:21-41), one filter id per quality, uploaded per layout."""

    def __init__(self, instance: dict, device: int = 0, quadtree_depth: int = 5, cd_threshold: int = 16, n_ff_poles: int = 2,
                 seed: int = 0, device_sampling: bool = True, rotation_lattice: int = 256):
        from .instance import load_instance

        self.rotation_lattice = int(rotation_lattice)

        self.inst = load_instance(instance)
        self.items = self.inst.items
        self.cfg = dict(quadtree_depth=quadtree_depth, cd_threshold=cd_threshold, n_ff_poles=n_ff_poles, n_ff_piers=0)
        self.device = device
        self.rng = np.random.Generator(np.random.Philox(key=[seed, 0xB1]))
        self.seed = int(seed)
        self.device_sampling = device_sampling
        self._round = 0
        self.samples = 0
        self._next_pk = 1
        self.filter_id = -1
        self.cde = None
        self.bbox = None
        self.layouts: List[BinLayout] = []
        self.stock = [int(b.get("stock", 1)) for b in self.inst.bins]
        self.qualities = sorted({it.min_quality for it in self.items if it.min_quality is not None})

    def _open(self, bin_id: int) -> BinLayout:
        from .instance import container_engine, container_hazards

        cde = container_engine(self.inst, bin_id, device=self.device, **self.cfg)
        bbox, _ = container_hazards(self.inst.bins[bin_id])
        for it in self.items:
            rots = self._device_rotations(it)
            if rots is not None:
                cde.set_item_rotations(it.id, rots)
        for q in self.qualities:  # padded so that the bitset keeps covering the slot map as items are placed
            bits = np.zeros(64, np.uint32)
            fb = cde.filter_from_irrelevant_qzones(q)
            bits[:len(fb)] = fb
            cde.upload_filter(int(q), bits)
        return BinLayout(bin_id, cde, bbox)

    def _search_in(self, lay: BinLayout, item, n_samples, ls_frac, uniform_rounds, ls_rounds):
        self.cde, self.bbox = lay.cde, lay.bbox
        self.filter_id = -1 if item.min_quality is None else int(item.min_quality)
        return self.search(item, n_samples, ls_frac, uniform_rounds, ls_rounds)

    def solve(self, n_samples: int = 5000, ls_frac: float = 0.2, uniform_rounds: int = 4, ls_rounds: int = 8,
              max_items: int = 1 << 30) -> BinResult:
        t0 = time.perf_counter()
        placed = unplaced = 0
        for item in sorted(self.items, key=lambda it: -it.diameter):
            demand = item.demand
            while demand > 0 and placed < max_items:
                where = None
                for lay in self.layouts:  # open layouts first
                    got = self._search_in(lay, item, n_samples, ls_frac, uniform_rounds, ls_rounds)
                    if got is not None:
                        where = (lay, got)
                        break
                if where is None:
                    for bin_id, qty in enumerate(self.stock):  # then closed bins with stock
                        if qty <= 0:
                            continue
                        lay = self._open(bin_id)
                        got = self._search_in(lay, item, n_samples, ls_frac, uniform_rounds, ls_rounds)
                        if got is not None:
                            self.stock[bin_id] -= 1
                            self.layouts.append(lay)
                            where = (lay, got)
                            break
                        lay.cde.close()
                if where is None:
                    unplaced += demand  # items of this type do not fit anywhere (lbf_bpp.rs:70)
                    break
                lay, (dt, _) = where
                pk = self._next_pk
                self._next_pk += 1
                lay.cde.place_item(item.id, item.shape, compose(dt.reshape(1, 3))[0], pk)
                lay.placements.append(Placement(item.id, dt, pk))
                demand -= 1
                placed += 1
        self.cde = None
        return BinResult(self.layouts, placed, unplaced, self.samples, time.perf_counter() - t0)

    def close(self):
        for lay in self.layouts:
            lay.cde.close()
        self.layouts = []
        self.cde = None


def solve_file(path: str, **kw) -> LbfResult:
    with open(path) as f:
        inst = json.load(f)
    solve_kw = {k: kw.pop(k) for k in list(kw) if k in ("n_samples", "ls_frac", "uniform_rounds", "ls_rounds", "max_items")}
    lbf = BatchedLBF(inst, **kw)
    try:
        return lbf.solve(**solve_kw)
    finally:
        lbf.close()
