"""ORACLE — TEST INFRASTRUCTURE ONLY. PARITY UNPINNED.

ctypes binding over oracle/_build/liboracle.so (the CPU restatement of jagua-rs's CDE, see
oracle/geom.hpp). Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this.
"""
from __future__ import annotations

import ctypes as C
import json
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS = {}

f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")
u64p = np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS")
i64p = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


def cpu_tag() -> str:
    """Tag of this machine's CPU feature flags: the oracle is built -march=native (like the reference's CI,
    -C target-cpu=native), so a library built on another CPU must not be loaded here."""
    import hashlib
    flags = ""
    try:
        with open("/proc/cpuinfo") as f:
            for ln in f:
                if ln.startswith("flags"):
                    flags = " ".join(sorted(ln.split(":", 1)[1].split()))
                    break
    except OSError:
        pass
    return hashlib.sha1(flags.encode()).hexdigest()[:10]


def _paths():
    tag = cpu_tag()
    return tag, os.path.join(_HERE, "_build", f"liboracle.{tag}.so"), os.path.join(_HERE, "_build", f"liboracle_count.{tag}.so")


def build(force: bool = False) -> None:
    """Compile the oracle with g++ (-O3 -march=native -ffp-contract=off). Building the checker is not using it."""
    tag, out, out_count = _paths()
    srcs = [os.path.join(_HERE, f) for f in ("capi.cpp", "geom.hpp", "spoly.hpp", "cde.hpp", "entities.hpp", "Makefile")]
    if not force and all(os.path.exists(o) and all(os.path.getmtime(o) >= os.path.getmtime(s) for s in srcs) for o in (out, out_count)):
        return
    subprocess.run(["make", "-C", _HERE, "-s", f"TAG={tag}"] + (["-B"] if force else []), check=True)


def lib(count: bool = False):
    key = "count" if count else "plain"
    if key in _LIBS:
        return _LIBS[key]
    build()
    path = _paths()[2 if count else 1]
    L = C.CDLL(path)
    vp = C.c_void_p
    sig = {
        "orc_world_new": (vp, [C.c_int] * 4),
        "orc_world_free": (None, [vp]),
        "orc_last_error": (C.c_char_p, [vp]),
        "orc_counting_enabled": (C.c_int, []),
        "orc_add_item": (C.c_int, [vp, f32p, C.c_int, f32p, C.c_int, C.c_int, C.c_int]),
        "orc_n_items": (C.c_int, [vp]),
        "orc_item_n_vertices": (C.c_int, [vp, C.c_int]),
        "orc_item_n_poles": (C.c_int, [vp, C.c_int]),
        "orc_item_n_piers": (C.c_int, [vp, C.c_int]),
        "orc_item_n_rotations": (C.c_int, [vp, C.c_int]),
        "orc_item_get": (None, [vp, C.c_int, f32p, f32p, f32p, f32p, f32p]),
        "orc_item_set_piers": (None, [vp, C.c_int, f32p, C.c_int]),
        "orc_add_item_raw": (C.c_int, [vp, f32p, C.c_int, f32p, f32p, C.c_int, f32p, C.c_int, f32p, C.c_int, C.c_int]),
        "orc_set_strip": (C.c_int, [vp, C.c_float, C.c_float]),
        "orc_strip_width": (C.c_float, [vp]),
        "orc_container_begin": (None, [vp, f32p, C.c_int]),
        "orc_container_add_zone": (C.c_int, [vp, C.c_int, f32p, C.c_int]),
        "orc_container_add_hole": (C.c_int, [vp, f32p, C.c_int]),
        "orc_container_finish": (C.c_int, [vp]),
        "orc_cde_begin": (None, [vp, f32p]),
        "orc_cde_add_static": (None, [vp, C.c_int, C.c_int64, C.c_int64, f32p, C.c_int, f32p]),
        "orc_cde_finish": (C.c_int, [vp]),
        "orc_cde_bbox": (None, [vp, f32p]),
        "orc_register_hazard_raw": (C.c_uint64, [vp, C.c_int, C.c_int64, C.c_int64, C.c_uint64, f32p, C.c_int, f32p, C.c_int]),
        "orc_deregister_hazard": (C.c_int, [vp, C.c_uint64]),
        "orc_cde_save": (C.c_int, [vp]),
        "orc_cde_restore": (C.c_int, [vp, C.c_int]),
        "orc_place_item": (C.c_uint64, [vp, C.c_int, C.c_float, C.c_float, C.c_float]),
        "orc_remove_item": (C.c_int, [vp, C.c_uint64]),
        "orc_n_placed": (C.c_int, [vp]),
        "orc_placed_get": (None, [vp, u64p, i32p, f32p]),
        "orc_is_feasible": (C.c_int, [vp]),
        "orc_change_strip_width": (C.c_int, [vp, C.c_float]),
        "orc_save": (C.c_int, [vp]),
        "orc_restore": (C.c_int, [vp, C.c_int]),
        "orc_n_hazards": (C.c_int, [vp]),
        "orc_haz_capacity": (C.c_int, [vp]),
        "orc_hazards_get": (None, [vp, u64p, i64p, f32p, u64p]),
        "orc_hazard_vertices": (C.c_int, [vp, C.c_uint64, f32p]),
        "orc_hkey_exterior": (C.c_uint64, [vp]),
        "orc_filter_qzones": (C.c_int, [vp, C.c_int, u32p, C.c_int]),
        "orc_tree_dump": (C.c_int64, [vp, u32p, C.c_int64]),
        "orc_check_invariants": (C.c_int, [vp]),
        "orc_matches_fresh": (C.c_int, [vp]),
        "orc_lbf_solve_spp": (C.c_int, [vp, C.c_int, C.c_float, C.c_uint64, C.c_int]),
        "orc_lbf_fill_bin": (C.c_int, [vp, C.c_int, C.c_float, C.c_uint64, C.c_int]),
        "orc_detect_batch": (C.c_uint64, [vp, C.c_int, f32p, C.c_int64, C.c_void_p, C.c_int, u8p, C.c_int, C.c_int]),
        "orc_collect_batch": (C.c_uint64, [vp, C.c_int, f32p, C.c_int64, u32p, C.c_int, C.c_int, C.c_int]),
        "orc_detect_multi": (C.c_uint64, [vp, C.c_int, i32p, i64p, f32p, C.c_void_p, C.c_int, u8p, C.c_int, C.c_int]),
        "orc_collect_multi": (C.c_uint64, [vp, C.c_int, i32p, i64p, f32p, C.c_void_p, u32p, C.c_int, C.c_int, C.c_int]),
        "orc_compose_batch": (None, [f32p, C.c_int64, f32p]),
        "orc_approx_eq": (C.c_int, [C.c_float, C.c_float]),
        "orc_edge_edge": (C.c_int, [f32p, f32p]),
        "orc_rect_edge": (C.c_int, [f32p, f32p]),
        "orc_circle_rect": (C.c_int, [f32p, f32p]),
        "orc_circle_edge": (C.c_int, [f32p, f32p]),
        "orc_rect_relation": (C.c_int, [f32p, f32p, C.c_int]),
        "orc_edge_quadrants": (C.c_int, [f32p, f32p]),
        "orc_edge_quadrants_individual": (C.c_int, [f32p, f32p]),
        "orc_pip": (C.c_int, [f32p, C.c_int, C.c_float, C.c_float]),
        "orc_transform_points": (None, [f32p, f32p, C.c_int, f32p]),
        "orc_import_polygon": (C.c_int, [f32p, C.c_int, f32p, f32p]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _LIBS[key] = L
    return L


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


KIND_PLACED_ITEM, KIND_EXTERIOR, KIND_HOLE, KIND_QZONE = 0, 1, 2, 3


class OracleWorld:
    """One layout (container + CDE + placed items) plus the instance's items."""

    def __init__(self, quadtree_depth=5, cd_threshold=16, n_ff_poles=2, n_ff_piers=0, count=False):
        self.L = lib(count)
        self.w = self.L.orc_world_new(quadtree_depth, cd_threshold, n_ff_poles, n_ff_piers)
        self.config = dict(quadtree_depth=quadtree_depth, cd_threshold=cd_threshold,
                           n_ff_poles=n_ff_poles, n_ff_piers=n_ff_piers)

    def __del__(self):
        try:
            self.L.orc_world_free(self.w)
        except Exception:
            pass

    def _chk(self, rc, bad=-1):
        if rc == bad:
            raise RuntimeError(self.L.orc_last_error(self.w).decode())
        return rc

    # -- items
    def add_item(self, pts, orientations_deg=None, min_quality=None, demand=1) -> int:
        pts = _f32(pts).reshape(-1, 2)
        if orientations_deg is None:
            o, n = np.zeros(1, np.float32), -1
        else:
            o, n = _f32(list(orientations_deg) + [0.0]), len(orientations_deg)
        mq = -1 if min_quality is None else int(min_quality)
        return self._chk(self.L.orc_add_item(self.w, pts, len(pts), o, n, mq, int(demand)))

    def add_item_raw(self, verts, poi, poles, piers, rotations, min_quality=-1) -> int:
        verts = _f32(verts).reshape(-1, 2)
        poles = _f32(poles).reshape(-1, 3)
        piers = _f32(piers).reshape(-1, 4)
        if rotations is None:
            r, n = np.zeros(1, np.float32), -1
        else:
            r, n = _f32(list(rotations) + [0.0]), len(rotations)
        return self._chk(self.L.orc_add_item_raw(self.w, verts, len(verts), _f32(poi), poles if len(poles) else np.zeros((1, 3), np.float32),
                                                 len(poles), piers if len(piers) else np.zeros((1, 4), np.float32), len(piers), r, n, int(min_quality)))

    def n_items(self) -> int:
        return self.L.orc_n_items(self.w)

    def item(self, i) -> dict:
        nv, npol, npier = (self.L.orc_item_n_vertices(self.w, i), self.L.orc_item_n_poles(self.w, i),
                           self.L.orc_item_n_piers(self.w, i))
        nrot = self.L.orc_item_n_rotations(self.w, i)
        verts = np.zeros((nv, 2), np.float32)
        poles = np.zeros((max(npol, 1), 3), np.float32)
        piers = np.zeros((max(npier, 1), 4), np.float32)
        rots = np.zeros(max(nrot, 1), np.float32)
        info = np.zeros(13, np.float32)
        self.L.orc_item_get(self.w, i, verts, poles, piers, rots, info)
        return dict(verts=verts, poles=poles[:npol], piers=piers[:npier],
                    rotations=None if nrot < 0 else rots[:nrot], poi=info[0:3].copy(), area=float(info[3]),
                    diameter=float(info[4]), bbox=info[5:9].copy(), pre_transform=info[9:11].copy(),
                    min_quality=int(info[11]), demand=int(info[12]))

    def set_item_piers(self, i, piers):
        piers = _f32(piers).reshape(-1, 4)
        self.L.orc_item_set_piers(self.w, i, piers if len(piers) else np.zeros((1, 4), np.float32), len(piers))

    # -- container
    def set_strip(self, width, height):
        self._chk(self.L.orc_set_strip(self.w, float(width), float(height)))

    def strip_width(self) -> float:
        return float(self.L.orc_strip_width(self.w))

    def set_container(self, outer, holes=(), zones=()):
        outer = _f32(outer).reshape(-1, 2)
        self.L.orc_container_begin(self.w, outer, len(outer))
        for q, pts in zones:
            pts = _f32(pts).reshape(-1, 2)
            self._chk(self.L.orc_container_add_zone(self.w, int(q), pts, len(pts)))
        for pts in holes:
            pts = _f32(pts).reshape(-1, 2)
            self._chk(self.L.orc_container_add_hole(self.w, pts, len(pts)))
        self._chk(self.L.orc_container_finish(self.w))

    def set_cde_raw(self, bbox, static_hazards):
        """static_hazards: list of dict(kind,id,quality,verts,poi); first must be the Exterior."""
        self.L.orc_cde_begin(self.w, _f32(bbox))
        for h in static_hazards:
            v = _f32(h["verts"]).reshape(-1, 2)
            self.L.orc_cde_add_static(self.w, int(h["kind"]), int(h["id"]), int(h["quality"]), v, len(v), _f32(h["poi"]))
        self._chk(self.L.orc_cde_finish(self.w))

    def register_hazard_raw(self, kind, verts, poi, id=0, quality=0, pk=0, dynamic=True) -> int:
        v = _f32(verts).reshape(-1, 2)
        p = np.zeros(3, np.float32)
        p[:len(np.ravel(poi))] = np.ravel(poi)[:3]
        return self._chk(self.L.orc_register_hazard_raw(self.w, int(kind), int(id), int(quality), int(pk), v, len(v), p,
                                                        1 if dynamic else 0), 0)

    def deregister_hazard(self, key):
        self._chk(self.L.orc_deregister_hazard(self.w, int(key)))

    def cde_save(self) -> int:
        return self.L.orc_cde_save(self.w)

    def cde_restore(self, snap):
        self._chk(self.L.orc_cde_restore(self.w, int(snap)))

    def cde_bbox(self) -> np.ndarray:
        b = np.zeros(4, np.float32)
        self.L.orc_cde_bbox(self.w, b)
        return b

    # -- layout
    def place_item(self, item_id, rot, tx, ty) -> int:
        return self._chk(self.L.orc_place_item(self.w, int(item_id), float(rot), float(tx), float(ty)), 0)

    def remove_item(self, pk):
        self._chk(self.L.orc_remove_item(self.w, int(pk)))

    def placed(self):
        n = self.L.orc_n_placed(self.w)
        pks = np.zeros(max(n, 1), np.uint64)
        ids = np.zeros(max(n, 1), np.int32)
        dts = np.zeros((max(n, 1), 3), np.float32)
        self.L.orc_placed_get(self.w, pks, ids, dts)
        return pks[:n], ids[:n], dts[:n]

    def is_feasible(self) -> bool:
        return bool(self._chk(self.L.orc_is_feasible(self.w)))

    def change_strip_width(self, w):
        self._chk(self.L.orc_change_strip_width(self.w, float(w)))

    def save(self) -> int:
        return self.L.orc_save(self.w)

    def restore(self, snap):
        self._chk(self.L.orc_restore(self.w, int(snap)))

    # -- hazards
    def hazards(self):
        n = self.L.orc_n_hazards(self.w)
        keys = np.zeros(n, np.uint64)
        meta = np.zeros((n, 6), np.int64)
        fl = np.zeros((n, 10), np.float32)
        pks = np.zeros(n, np.uint64)
        self.L.orc_hazards_get(self.w, keys, meta, fl, pks)
        out = []
        for i in range(n):
            v = np.zeros((int(meta[i, 5]), 2), np.float32)
            self.L.orc_hazard_vertices(self.w, int(keys[i]), v)
            out.append(dict(key=int(keys[i]), slot=int(keys[i]) & 0xFFFFFFFF, kind=int(meta[i, 0]), id=int(meta[i, 1]),
                            quality=int(meta[i, 2]), scope=int(meta[i, 3]), dynamic=bool(meta[i, 4]), verts=v,
                            poi=fl[i, 0:3].copy(), bbox=fl[i, 3:7].copy(), dt=fl[i, 7:10].copy(), pk=int(pks[i])))
        return out

    def haz_capacity(self) -> int:
        return self.L.orc_haz_capacity(self.w)

    def filter_qzones(self, required_quality, words) -> np.ndarray:
        bits = np.zeros(words, np.uint32)
        self.L.orc_filter_qzones(self.w, int(required_quality), bits, words)
        return bits

    def tree_dump(self) -> np.ndarray:
        n = self.L.orc_tree_dump(self.w, np.zeros(1, np.uint32), 0)
        buf = np.zeros(n, np.uint32)
        self.L.orc_tree_dump(self.w, buf, n)
        return buf

    def check_invariants(self) -> int:
        return self.L.orc_check_invariants(self.w)

    def matches_fresh(self) -> int:
        return self._chk(self.L.orc_matches_fresh(self.w))

    # -- fillers
    def lbf_solve_spp(self, n_samples=5000, ls_frac=0.2, seed=0, max_items=1 << 30) -> int:
        return self._chk(self.L.orc_lbf_solve_spp(self.w, n_samples, ls_frac, seed, max_items))

    def lbf_fill_bin(self, n_samples=5000, ls_frac=0.2, seed=0, max_items=1 << 30) -> int:
        return self._chk(self.L.orc_lbf_fill_bin(self.w, n_samples, ls_frac, seed, max_items))

    # -- queries
    def detect_batch(self, item_id, t4, filter_bits=None, nthreads=1, poly_only=False):
        t4 = _f32(t4).reshape(-1, 4)
        out = np.zeros(len(t4), np.uint8)
        if filter_bits is None:
            fb, fw = None, 0
        else:
            filter_bits = np.ascontiguousarray(filter_bits, dtype=np.uint32)
            fb, fw = filter_bits.ctypes.data_as(C.c_void_p), len(filter_bits)
        flops = self.L.orc_detect_batch(self.w, int(item_id), t4, len(t4), fb, fw, out, int(nthreads), 1 if poly_only else 0)
        self.last_flops = int(flops)
        return out

    def collect_batch(self, item_id, t4, words, nthreads=1, poly_only=False):
        t4 = _f32(t4).reshape(-1, 4)
        out = np.zeros((len(t4), words), np.uint32)
        flops = self.L.orc_collect_batch(self.w, int(item_id), t4, len(t4), out, int(words), int(nthreads), 1 if poly_only else 0)
        self.last_flops = int(flops)
        return out


    def detect_multi(self, item_ids, seg_offsets, t4, filter_bits=None, nthreads=1, poly_only=False, out=None):
        """One parallel pass over a segmented batch (segment k = queries seg_offsets[k]..seg_offsets[k+1] of item_ids[k])."""
        t4 = _f32(t4).reshape(-1, 4)
        ids = np.ascontiguousarray(item_ids, dtype=np.int32)
        offs = np.ascontiguousarray(seg_offsets, dtype=np.int64)
        assert len(offs) == len(ids) + 1 and int(offs[-1]) == len(t4)
        if out is None:
            out = np.zeros(len(t4), np.uint8)
        if filter_bits is None:
            fb, fw = None, 0
        else:
            filter_bits = np.ascontiguousarray(filter_bits, dtype=np.uint32)
            fb, fw = filter_bits.ctypes.data_as(C.c_void_p), len(filter_bits)
        self.last_flops = int(self.L.orc_detect_multi(self.w, len(ids), ids, offs, t4, fb, fw, out, int(nthreads), 1 if poly_only else 0))
        return out

    def collect_multi(self, item_ids, seg_offsets, t4, words, seed=None, nthreads=1, poly_only=False):
        """Segmented collect batch; `seed` (n x words): collectors passed in non-empty."""
        t4 = _f32(t4).reshape(-1, 4)
        ids = np.ascontiguousarray(item_ids, dtype=np.int32)
        offs = np.ascontiguousarray(seg_offsets, dtype=np.int64)
        assert len(offs) == len(ids) + 1 and int(offs[-1]) == len(t4)
        out = np.zeros((len(t4), words), np.uint32)
        sp = None
        if seed is not None:
            seed = np.ascontiguousarray(seed, dtype=np.uint32)
            assert seed.shape == out.shape
            sp = seed.ctypes.data_as(C.c_void_p)
        self.last_flops = int(self.L.orc_collect_multi(self.w, len(ids), ids, offs, t4, sp, out, int(words), int(nthreads), 1 if poly_only else 0))
        return out


def compose(dts) -> np.ndarray:
    """(rotation, tx, ty) -> (cos, sin, tx, ty) with the host libm, as the reference composes."""
    dts = _f32(dts).reshape(-1, 3)
    out = np.zeros((len(dts), 4), np.float32)
    lib().orc_compose_batch(dts, len(dts), out)
    return out


# ---- reference instance JSON (io/ext_repr.rs + probs/*/io/ext_repr.rs) -------------------------
def _shape_points(shape: dict):
    """ExtShape -> (outer points, [hole points]) (io/ext_repr.rs:34-52)."""
    t, d = shape["type"], shape["data"]
    if t == "simple_polygon":
        return np.array(d, dtype=np.float64).astype(np.float32), []
    if t == "polygon":
        return (np.array(d["outer"], dtype=np.float64).astype(np.float32),
                [np.array(h, dtype=np.float64).astype(np.float32) for h in d.get("inner", [])])
    if t == "rectangle":
        x0, y0 = np.float32(d["x_min"]), np.float32(d["y_min"])
        x1, y1 = np.float32(x0 + np.float32(d["width"])), np.float32(y0 + np.float32(d["height"]))
        return np.array([[x0, y0], [x1, y0], [x1, y1], [x0, y1]], np.float32), []
    raise ValueError(f"unsupported shape type {t}")


def load_instance_json(path: str) -> dict:
    """Parse a reference instance file (serde f32 semantics: decimal -> f64 -> f32)."""
    with open(path) as f:
        return json.load(f)


def world_from_instance(inst: dict, bin_index: int = 0, **cfg) -> OracleWorld:
    """Import items (+ strip or one bin) the way io/import.rs + probs/{spp,bpp}/io/import.rs do."""
    w = OracleWorld(**cfg)
    items = sorted(inst["items"], key=lambda it: it["id"])
    items = [it for it in items if it.get("demand", 1) > 0]
    for it in items:
        pts, _ = _shape_points(it["shape"])
        w.add_item(pts, it.get("allowed_orientations"), it.get("min_quality"), it.get("demand", 1))
    if "strip_height" in inst:
        h = np.float32(inst["strip_height"])
        area = np.float32(0.0)
        for i, it in enumerate(items):  # probs/spp/io/import.rs:40-48 (f32 sum of area * demand)
            area = np.float32(area + np.float32(np.float32(_orig_area(w, i)) * np.float32(it["demand"])))
        w.set_strip(float(np.float32(area / h)), float(h))
    else:
        b = sorted(inst["bins"], key=lambda b: b["id"])[bin_index]
        outer, holes = _shape_points(b["shape"])
        zones = [(z["quality"], _shape_points(z["shape"])[0]) for z in b.get("zones", [])]
        w.set_container(outer, holes, zones)
    return w


def _orig_area(w: OracleWorld, i: int) -> float:
    # Item::area() = shape_orig.area; centring is a pure translation and area is not recomputed
    return w.item(i)["area"]
