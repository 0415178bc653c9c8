"""jagua-rs_b200 — B200-native batched collision queries behind jagua-rs's CDEngine interface.

Host-side mirror (Python over the C ABI in include/jagua_cuda.h) of the reference's Rust API for the
collision-query path: `CDEngine` with `register_hazard`, `deregister_hazard_by_key`, `save`,
`restore`, and batched versions of `detect_surrogate_collision` + `detect_poly_collision`
(`detect_batch`) and `collect_surrogate_collisions` + `collect_poly_collisions` (`collect_batch`)
— jagua-rs/src/collision_detection/cd_engine.rs:33-342.

All compute runs in libjagua_cuda.so (hand-written sm_100a kernels). There is no CPU fallback: if the
library is missing or no GPU is present, the query calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np

from . import build as _build

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

HAZ_PLACED_ITEM, HAZ_EXTERIOR, HAZ_HOLE, HAZ_QZONE = 0, 1, 2, 3
VERDICT_SURROGATE, VERDICT_POLY, VERDICT_REF_PANIC = 1, 2, 0x80
MODE_DEFAULT, MODE_POLY_ONLY = 0, 1

JG_OK, JG_ERR_CUDA, JG_ERR_INVALID, JG_ERR_NO_LAYOUT, JG_ERR_CAPACITY, JG_ERR_REFERENCE_PANIC = 0, -1, -2, -3, -4, -5


class JgConfig(C.Structure):
    _fields_ = [("quadtree_depth", C.c_int32), ("cd_threshold", C.c_int32), ("n_ff_poles", C.c_int32),
                ("n_ff_piers", C.c_int32)]


class JaguaCudaError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"jagua_cuda error {code}: {msg}")
        self.code = code


_vp = C.c_void_p
_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")

# every symbol include/jagua_cuda.h declares: (restype, argtypes)
C_ABI = {
    "jg_engine_create": (C.c_int, [C.c_int, C.POINTER(JgConfig), C.POINTER(_vp)]),
    "jg_engine_destroy": (None, [_vp]),
    "jg_last_error": (C.c_char_p, [_vp]),
    "jg_layout_upload": (C.c_int, [_vp, _vp, C.c_size_t, C.c_uint64]),
    "jg_layout_upload_device": (C.c_int, [_vp, _vp, C.c_size_t, C.c_uint64]),
    "jg_layout_patch": (C.c_int, [_vp, _vp, C.c_size_t, C.c_uint64, C.c_uint64, _vp, C.c_int, C.c_uint64]),
    "jg_cde_image_sync": (C.c_int, [_vp, C.POINTER(_vp), C.POINTER(C.c_size_t), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(_vp),
                                   C.POINTER(C.c_int)]),
    "jg_engine_layout_epoch": (C.c_uint64, [_vp]),
    "jg_nccl_unique_id": (C.c_int, [_u8p]),
    "jg_engine_comm_init": (C.c_int, [_vp, _u8p, C.c_int, C.c_int, C.c_size_t]),
    "jg_engine_comm_destroy": (C.c_int, [_vp]),
    "jg_layout_broadcast": (C.c_int, [_vp, C.c_int, _vp, C.c_size_t, C.c_uint64]),
    "jg_cde_broadcast": (C.c_int, [_vp, _vp, C.c_int, C.POINTER(C.c_uint64)]),
    "jg_item_upload": (C.c_int, [_vp, C.c_int, _f32p, C.c_int, _f32p, _f32p, C.c_int, _f32p, C.c_int]),
    "jg_filter_upload": (C.c_int, [_vp, C.c_int, _u32p, C.c_int]),
    "jg_filter_upload_keys": (C.c_int, [_vp, C.c_int, _vp, C.c_int]),
    "jg_detect_multi_lattice": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp, _vp, _vp, C.c_int, C.c_int, _vp]),
    "jg_detect_multi_lattice_async": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp, _vp, _vp, C.c_int, C.c_int, _vp]),
    "jg_lattice_points": (None, [_vp, _vp, C.c_int64, _vp]),
    "jg_collect_multi": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp, C.c_int, _vp, _vp, C.c_int]),
    "jg_collect_multi_async": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp, C.c_int, _vp, _vp, C.c_int]),
    "jg_collect_multi_device": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp, C.c_int, _vp, _vp, C.c_int]),
    "jg_host_alloc_wc": (C.c_int, [C.c_size_t, C.POINTER(_vp)]),
    "jg_cde_deregister_hazard_by_entity": (C.c_int, [_vp, C.c_int, C.c_int64, C.c_int64, C.c_uint64]),
    "jg_cde_swap_container": (C.c_int, [_vp, _vp]),
    "jg_cde_clone": (C.c_int, [_vp, C.POINTER(_vp)]),
    "jg_cde_epoch": (C.c_uint64, [_vp]),
    "jg_pool_create": (C.c_int, [_vp, C.c_int, C.POINTER(JgConfig), C.POINTER(_vp)]),
    "jg_pool_destroy": (None, [_vp]),
    "jg_pool_size": (C.c_int, [_vp]),
    "jg_pool_engine": (_vp, [_vp, C.c_int]),
    "jg_pool_last_error": (C.c_char_p, [_vp]),
    "jg_pool_set_min_shard": (C.c_int, [_vp, C.c_int64]),
    "jg_pool_item_upload": (C.c_int, [_vp, C.c_int, _f32p, C.c_int, _f32p, _f32p, C.c_int, _f32p, C.c_int]),
    "jg_pool_item_set_rotations": (C.c_int, [_vp, C.c_int, _vp, C.c_int]),
    "jg_pool_filter_upload": (C.c_int, [_vp, C.c_int, _u32p, C.c_int]),
    "jg_pool_layout_upload": (C.c_int, [_vp, _vp, C.c_size_t, C.c_uint64]),
    "jg_pool_cde_sync": (C.c_int, [_vp, _vp, C.POINTER(C.c_uint64)]),
    "jg_pool_detect_multi": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp, C.c_int, C.c_int, _vp]),
    "jg_pool_detect_multi_async": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp, C.c_int, C.c_int, _vp]),
    "jg_pool_detect_multi_indexed": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp, _vp, C.c_int, C.c_int, _vp]),
    "jg_pool_detect_multi_indexed_async": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp, _vp, C.c_int, C.c_int, _vp]),
    "jg_pool_detect_multi_lattice": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp, _vp, _vp, C.c_int, C.c_int, _vp]),
    "jg_pool_detect_multi_lattice_async": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp, _vp, _vp, C.c_int, C.c_int, _vp]),
    "jg_pool_collect_multi": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp, C.c_int, _vp, _vp, C.c_int]),
    "jg_pool_collect_multi_async": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp, C.c_int, _vp, _vp, C.c_int]),
    "jg_pool_sync": (C.c_int, [_vp]),
    "jg_detect_batch": (C.c_int, [_vp, C.c_int, _vp, C.c_int64, C.c_int, C.c_int, _vp]),
    "jg_detect_batch_async": (C.c_int, [_vp, C.c_int, _vp, C.c_int64, C.c_int, C.c_int, _vp]),
    "jg_collect_batch_async": (C.c_int, [_vp, C.c_int, _vp, C.c_int64, C.c_int, _vp, C.c_int]),
    "jg_detect_multi": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp, C.c_int, C.c_int, _vp]),
    "jg_detect_multi_async": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp, C.c_int, C.c_int, _vp]),
    "jg_detect_multi_device": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp, C.c_int, C.c_int, _vp]),
    "jg_item_set_rotations": (C.c_int, [_vp, C.c_int, _vp, C.c_int]),
    "jg_detect_multi_indexed": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp, _vp, C.c_int, C.c_int, _vp]),
    "jg_detect_multi_indexed_async": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp, _vp, C.c_int, C.c_int, _vp]),
    "jg_detect_batch_device": (C.c_int, [_vp, C.c_int, _vp, C.c_int64, C.c_int, C.c_int, _vp]),
    "jg_collect_batch": (C.c_int, [_vp, C.c_int, _vp, C.c_int64, C.c_int, _vp, C.c_int]),
    "jg_collect_batch_device": (C.c_int, [_vp, C.c_int, _vp, C.c_int64, C.c_int, _vp, C.c_int]),
    "jg_search_round": (C.c_int, [_vp, C.c_int, _vp, C.c_int64, C.c_int, _vp]),
    "jg_search_samples": (C.c_int, [_vp, C.c_int64, _vp, _vp, _vp]),
    "jg_sync": (C.c_int, [_vp]),
    "jg_fence_record": (C.c_int, [_vp, C.POINTER(C.c_uint64)]),
    "jg_fence_wait": (C.c_int, [_vp, C.c_uint64]),
    "jg_engine_stats": (C.c_int, [_vp, C.POINTER(C.c_uint64)]),
    "jg_engine_last_kernel_ms": (C.c_float, [_vp]),
    "jg_timer_start": (C.c_int, [_vp]),
    "jg_timer_stop": (C.c_int, [_vp, C.POINTER(C.c_float)]),
    "jg_engine_set_batching": (C.c_int, [_vp, C.c_int64, C.c_int]),
    "jg_timer_start_gated": (C.c_int, [_vp]),
    "jg_pool_shard_bounds": (C.c_int, [C.c_int64, C.c_int, C.c_int, C.c_int64, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "jg_engine_set_groups": (C.c_int, [_vp, C.c_int, C.c_int]),
    "jg_engine_set_split_min": (C.c_int, [_vp, C.c_int64]),
    "jg_set_dependent_launch": (C.c_int, [C.c_int]),
    "jg_dependent_launch": (C.c_int, []),
    "jg_engine_set_detect_path": (C.c_int, [_vp, C.c_int]),
    "jg_engine_detect_path": (C.c_int, [_vp]),
    "jg_host_alloc": (C.c_int, [C.c_size_t, C.POINTER(_vp)]),
    "jg_host_free": (C.c_int, [_vp]),
    "jg_device_alloc": (C.c_int, [_vp, C.c_size_t, C.POINTER(_vp)]),
    "jg_device_free": (C.c_int, [_vp, _vp]),
    "jg_memcpy_h2d": (C.c_int, [_vp, _vp, _vp, C.c_size_t]),
    "jg_memcpy_d2h": (C.c_int, [_vp, _vp, _vp, C.c_size_t]),
    "jg_compose_batch": (None, [_f32p, C.c_int64, _f32p, C.c_int]),
    "jg_fp32_nofma_peak": (C.c_int, [_vp, C.POINTER(C.c_double), C.POINTER(C.c_float)]),
    "jg_flush_l2": (C.c_int, [_vp]),
    "jg_item_import_modified": (C.c_int, [_f32p, C.c_int, _vp, C.c_float, C.c_float, C.c_float, _f32p, C.c_int, C.POINTER(C.c_int), _f32p,
                                          _f32p, C.c_int, C.POINTER(C.c_int), _f32p, C.POINTER(C.c_float)]),
    "jg_shape_offset": (C.c_int, [_f32p, C.c_int, C.c_int, C.c_float, C.c_float, _f32p, C.c_int, C.POINTER(C.c_int), _f32p]),
    "jg_item_generate_piers": (C.c_int, [_f32p, C.c_int, _vp, C.c_int, C.c_int, _vp, _f32p]),
    "jg_item_import": (C.c_int, [_f32p, C.c_int, _vp, _f32p, C.POINTER(C.c_int), _f32p, _f32p, C.c_int, C.POINTER(C.c_int), _f32p,
                                 C.POINTER(C.c_float)]),
    "jg_item_import_simplified": (C.c_int, [_f32p, C.c_int, _vp, C.c_float, C.c_float, _f32p, C.POINTER(C.c_int), _f32p, _f32p, C.c_int,
                                            C.POINTER(C.c_int), _f32p, C.POINTER(C.c_float)]),
    "jg_shape_close_concavities": (C.c_int, [_f32p, C.c_int, C.c_int, C.c_float, _f32p, C.POINTER(C.c_int), _f32p]),
    "jg_shape_simplify": (C.c_int, [_f32p, C.c_int, C.c_int, C.c_float, _f32p, C.POINTER(C.c_int), _f32p]),
    "jg_polygon_poi": (C.c_int, [_f32p, C.c_int, _f32p, _vp]),
    "jg_shape_import": (C.c_int, [_f32p, C.c_int, _f32p, C.POINTER(C.c_int), _f32p]),
    "jg_cde_create": (C.c_int, [_f32p, C.POINTER(JgConfig), C.POINTER(_vp)]),
    "jg_cde_destroy": (None, [_vp]),
    "jg_cde_last_error": (C.c_char_p, [_vp]),
    "jg_cde_register_hazard": (C.c_int, [_vp, C.c_int, C.c_int64, C.c_int64, C.c_uint64, _f32p, C.c_int, _f32p, C.c_int,
                                         C.POINTER(C.c_uint64)]),
    "jg_cde_place_item": (C.c_int, [_vp, C.c_int64, C.c_uint64, _f32p, C.c_int, _f32p, _f32p, C.POINTER(C.c_uint64)]),
    "jg_cde_remove_item": (C.c_int, [_vp, C.c_uint64]),
    "jg_cde_deregister_hazard": (C.c_int, [_vp, C.c_uint64]),
    "jg_cde_hkey_from_pk": (C.c_int, [_vp, C.c_uint64, C.POINTER(C.c_uint64)]),
    "jg_cde_save": (C.c_int, [_vp, C.POINTER(C.c_int)]),
    "jg_cde_restore": (C.c_int, [_vp, C.c_int]),
    "jg_cde_slot_capacity": (C.c_int, [_vp]),
    "jg_cde_n_hazards": (C.c_int, [_vp]),
    "jg_cde_hkey_exterior": (C.c_uint64, [_vp]),
    "jg_cde_filter_qzones": (C.c_int, [_vp, C.c_int, _u32p, C.c_int]),
    "jg_cde_flatten": (C.c_int64, [_vp, _vp, C.c_int64]),
    "jg_cde_sync_to_engine": (C.c_int, [_vp, _vp, C.POINTER(C.c_uint64)]),
    "jg_cde_tree_dump": (C.c_int64, [_vp, _vp, C.c_int64]),
}


def lib_path() -> str:
    return _build.LIB


def load_library(build_if_missing: bool = True):
    """dlopen libjagua_cuda.so (building it with nvcc when absent) and type every C-ABI symbol."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(_build.LIB):
        if not build_if_missing:
            raise JaguaCudaError(JG_ERR_INVALID, f"{_build.LIB} is missing; run __graft_entry__.build()")
        _build.build()
    lib = C.CDLL(_build.LIB)
    for name, (res, args) in C_ABI.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _LIB = lib
    return lib


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def compose(d_transfs, n_threads: int = 1) -> np.ndarray:
    """DTransformation (rotation, tx, ty) -> (cos, sin, tx, ty) with the host libm
    (geometry/transformation.rs:131-139 rot_transl_m; f32::sin_cos)."""
    d = _f32(d_transfs).reshape(-1, 3)
    out = np.empty((len(d), 4), np.float32)
    load_library().jg_compose_batch(d, len(d), out, int(n_threads))
    return out


@dataclass
class ItemShape:
    """The part of an Item the CDE queries read: SPolygon vertices + poi centre + SPSurrogate
    (entities/item.rs:12-53, geometry/primitives/simple_polygon.rs:24-37, fail_fast/sp_surrogate.rs:17-30)."""
    verts: np.ndarray            # (V, 2) CCW, centred
    poi: np.ndarray              # (x, y[, r])
    poles: np.ndarray            # (P, 3) x, y, r
    piers: np.ndarray = None     # (Q, 4)


class JgSearchRound(C.Structure):
    _fields_ = [("mode", C.c_int), ("rot_idx", C.c_int), ("box", C.c_float * 4), ("mean", C.c_float * 2), ("sd", C.c_float),
                ("seed", C.c_uint64), ("counter", C.c_uint64)]


class JgSearchBest(C.Structure):
    _fields_ = [("n_free", C.c_int64), ("index", C.c_int64), ("loss", C.c_float), ("tx", C.c_float), ("ty", C.c_float),
                ("rot_idx", C.c_int)]


class JgSurrogateConfig(C.Structure):
    _fields_ = [("n_pole_limits", C.c_int * 3), ("coverage_limits", C.c_float * 3), ("n_ff_poles", C.c_int), ("n_ff_piers", C.c_int)]


def simplify_shape(points, mode: str = "inflate", max_area_change_ratio: float = 0.001):
    """simplify_shape (geometry/shape_modification.rs:48-144) on a clean counter-clockwise vertex list: (vertices, poi)."""
    lib = load_library()
    pts = _f32(points).reshape(-1, 2)
    verts = np.zeros_like(pts)
    poi = np.zeros(3, np.float32)
    nv = C.c_int(0)
    rc = lib.jg_shape_simplify(pts, len(pts), {"inflate": 0, "deflate": 1}[mode], float(max_area_change_ratio), verts, C.byref(nv), poi)
    if rc != JG_OK:
        raise JaguaCudaError(rc, "jg_shape_simplify: bad input or a degenerate replacement edge (the reference panics)")
    return verts[:nv.value].copy(), poi


def close_narrow_concavities(points, mode: str = "inflate", max_distance_ratio: float = 0.01):
    """close_narrow_concavities (geometry/shape_modification.rs:374-502) on a clean counter-clockwise vertex list."""
    lib = load_library()
    pts = _f32(points).reshape(-1, 2)
    verts = np.zeros_like(pts)
    poi = np.zeros(3, np.float32)
    nv = C.c_int(0)
    rc = lib.jg_shape_close_concavities(pts, len(pts), {"inflate": 0, "deflate": 1}[mode], float(max_distance_ratio), verts, C.byref(nv), poi)
    if rc != JG_OK:
        raise JaguaCudaError(rc, "jg_shape_close_concavities: bad input or a degenerate shape (the reference panics)")
    return verts[:nv.value].copy(), poi


def offset_shape(points, mode: str = "inflate", distance: float = 0.0, miter_limit: float = 2.0):
    """offset_shape (geometry/shape_modification.rs:330-372) with the repo's own miter offsetter (not geo-buffer: not
    bit-equal to the reference): (vertices, poi) of the clean counter-clockwise polygon moved by `distance`."""
    lib = load_library()
    pts = _f32(points).reshape(-1, 2)
    verts = np.zeros((2 * len(pts), 2), np.float32)
    poi = np.zeros(3, np.float32)
    nv = C.c_int(0)
    rc = lib.jg_shape_offset(pts, len(pts), {"inflate": 0, "deflate": 1}[mode], float(distance), float(miter_limit), verts, len(verts),
                             C.byref(nv), poi)
    if rc != JG_OK:
        raise JaguaCudaError(rc, "jg_shape_offset: bad input, or the offset polygon intersects itself (reduce the separation)")
    return verts[:nv.value].copy(), poi


class JgPierConfig(C.Structure):
    _fields_ = [("rays_per_angle", C.c_int), ("n_angles", C.c_int), ("n_points_per_dimension", C.c_int)]


def generate_piers(verts, poles, n_piers: int, rays_per_angle: int = 0, n_angles: int = 0, n_points_per_dimension: int = 0):
    """generate_piers (geometry/fail_fast/piers.rs:25-226): n_piers segments (sx, sy, ex, ey) inside the imported shape
    `verts` that cover what `poles` (the fail-fast poles: sp_surrogate.rs:40-44) leave unrepresented. Zero sizes select
    the reference's release constants (200 rays per angle, 90 angles, 100 x 100 grid points)."""
    lib = load_library()
    v = _f32(verts).reshape(-1, 2)
    pl = _f32(poles).reshape(-1, 3)
    out = np.zeros((int(n_piers), 4), np.float32)
    cfg = JgPierConfig(int(rays_per_angle), int(n_angles), int(n_points_per_dimension))
    rc = lib.jg_item_generate_piers(v, len(v), pl.ctypes.data if len(pl) else None, len(pl), int(n_piers), C.addressof(cfg), out)
    if rc != JG_OK:
        raise JaguaCudaError(rc, "jg_item_generate_piers: bad input, or a state in which the reference panics")
    return out


def import_item(points, n_pole_limits=((100, 0.0), (20, 0.75), (10, 0.90)), n_ff_poles: int = 2, simplify_tolerance=None,
                narrow_concavity_cutoff_ratio=None, n_ff_piers: int = 0, offset: float = 0.0):
    """Importer::import_item + Item::new's surrogate for a simple polygon (io/import.rs:47-98, fail_fast/pole.rs:12-135)
    through jg_item_import: returns (ItemShape, pre_transform (tx, ty), area). offset > 0 = ShapeModifyConfig::offset (half
    the minimum separation; the repo's own offsetter, not bit-equal to geo-buffer); n_ff_piers > 0 adds generate_piers against
    the first n_ff_poles poles (sp_surrogate.rs:40-44)."""
    lib = load_library()
    pts = _f32(points).reshape(-1, 2)
    cfg = JgSurrogateConfig((C.c_int * 3)(*[int(a) for a, _ in n_pole_limits]), (C.c_float * 3)(*[float(b) for _, b in n_pole_limits]),
                            int(n_ff_poles), 0)
    verts = np.zeros((2 * len(pts), 2), np.float32)
    poles = np.zeros((1000, 3), np.float32)
    poi, pre = np.zeros(3, np.float32), np.zeros(2, np.float32)
    nv, npol, area = C.c_int(0), C.c_int(0), C.c_float(0.0)
    if offset > 0.0:
        rc = lib.jg_item_import_modified(pts, len(pts), C.addressof(cfg), float(offset),
                                         -1.0 if simplify_tolerance is None else float(simplify_tolerance),
                                         -1.0 if narrow_concavity_cutoff_ratio is None else float(narrow_concavity_cutoff_ratio),
                                         verts, len(verts), C.byref(nv), poi, poles, len(poles), C.byref(npol), pre, C.byref(area))
    elif simplify_tolerance is None:
        rc = lib.jg_item_import(pts, len(pts), C.addressof(cfg), verts, C.byref(nv), poi, poles, len(poles), C.byref(npol), pre, C.byref(area))
    else:  # Importer's poly_simpl_tolerance (lbf default 0.001)
        rc = lib.jg_item_import_simplified(pts, len(pts), C.addressof(cfg), float(simplify_tolerance),
                                           -1.0 if narrow_concavity_cutoff_ratio is None else float(narrow_concavity_cutoff_ratio),
                                           verts, C.byref(nv), poi, poles, len(poles), C.byref(npol), pre, C.byref(area))
    if rc != JG_OK:
        raise JaguaCudaError(rc, "jg_item_import: degenerate polygon (fewer than 3 distinct vertices, duplicate points or no area)")
    piers = np.zeros((0, 4), np.float32)
    if n_ff_piers > 0:
        piers = generate_piers(verts[:nv.value], poles[:min(int(n_ff_poles), npol.value)], n_ff_piers)
    return ItemShape(verts[:nv.value].copy(), poi, poles[:npol.value].copy(), piers), pre, float(area.value)


def polygon_poi(points):
    """SPolygon::new for a shape registered as it is: (counter-clockwise vertices, poi (x, y, r))."""
    lib = load_library()
    pts = _f32(points).reshape(-1, 2)
    poi = np.zeros(3, np.float32)
    ccw = np.zeros_like(pts)
    rc = lib.jg_polygon_poi(pts, len(pts), poi, ccw.ctypes.data)
    if rc != JG_OK:
        raise JaguaCudaError(rc, "jg_polygon_poi: degenerate polygon")
    return ccw, poi


def import_shape(points):
    """import_simple_polygon + SPolygon::new for a container boundary / hole / quality zone: (vertices, poi)."""
    lib = load_library()
    pts = _f32(points).reshape(-1, 2)
    verts = np.zeros_like(pts)
    poi = np.zeros(3, np.float32)
    nv = C.c_int(0)
    rc = lib.jg_shape_import(pts, len(pts), verts, C.byref(nv), poi)
    if rc != JG_OK:
        raise JaguaCudaError(rc, "jg_shape_import: degenerate polygon")
    return verts[:nv.value].copy(), poi


class JgLattice(C.Structure):
    """Sampling lattice of the 5-byte query form: translation = (x0 + i * dx, y0 + j * dy), i, j in 0..65535."""
    _fields_ = [("x0", C.c_float), ("y0", C.c_float), ("dx", C.c_float), ("dy", C.c_float)]

    @staticmethod
    def over_box(box, steps: int = 65536) -> "JgLattice":
        """`steps` points per axis over [x_min, x_max) x [y_min, y_max) (UniformRectSampler's box, quantised)."""
        b = np.asarray(box, np.float32)
        return JgLattice(float(b[0]), float(b[1]), float((b[2] - b[0]) / np.float32(steps)), float((b[3] - b[1]) / np.float32(steps)))


def lattice_points(lattice: JgLattice, ij: np.ndarray) -> np.ndarray:
    """The f32 translations the device derives from lattice coordinates (jg_lattice_points, the host twin)."""
    ij = np.ascontiguousarray(ij, dtype=np.uint16).reshape(-1, 2)
    out = np.empty((len(ij), 2), np.float32)
    load_library().jg_lattice_points(C.addressof(lattice), ij.ctypes.data, len(ij), out.ctypes.data)
    return out


class PinnedBuffer:
    """Page-locked host array (cudaHostAlloc) so that H2D/D2H copies overlap the kernels. write_combined=True: for
    buffers the host only writes (query records)."""

    def __init__(self, shape, dtype, write_combined: bool = False):
        self._lib = load_library()
        self.nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        p = _vp()
        rc = (self._lib.jg_host_alloc_wc if write_combined else self._lib.jg_host_alloc)(max(self.nbytes, 1), C.byref(p))
        if rc != JG_OK:
            raise JaguaCudaError(rc, "cudaHostAlloc failed (no GPU?)")
        self.ptr = p.value
        buf = (C.c_char * max(self.nbytes, 1)).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def free(self):
        if self.ptr:
            self.array = None
            self._lib.jg_host_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class DeviceBuffer:
    def __init__(self, cde: "CDEngine", nbytes: int):
        self._cde = cde
        self.nbytes = int(nbytes)
        p = _vp()
        cde._chk(cde._lib.jg_device_alloc(cde._eng, max(self.nbytes, 1), C.byref(p)))
        self.ptr = p.value

    def upload(self, arr: np.ndarray):
        arr = np.ascontiguousarray(arr)
        assert arr.nbytes <= self.nbytes
        self._cde._chk(self._cde._lib.jg_memcpy_h2d(self._cde._eng, self.ptr, arr.ctypes.data, arr.nbytes))

    def download(self, shape, dtype) -> np.ndarray:
        out = np.empty(shape, dtype)
        assert out.nbytes <= self.nbytes
        self._cde._chk(self._cde._lib.jg_memcpy_d2h(self._cde._eng, out.ctypes.data, self.ptr, out.nbytes))
        return out

    def free(self):
        if self.ptr and self._cde._eng:
            self._cde._lib.jg_device_free(self._cde._eng, self.ptr)
        self.ptr = None


class CDEngine:
    """Mirror of jagua-rs's CDEngine (cd_engine.rs:22-342) whose queries run batched on one B200.

    The update path (register / deregister / save / restore) runs in the C++ host mirror and the
    flattened quadtree is re-uploaded lazily before the next query. `device=None` creates the host
    half only (quadtree maintenance + flatten; usable without a GPU)."""

    def __init__(self, bbox, quadtree_depth: int = 5, cd_threshold: int = 16, n_ff_poles: int = 2, n_ff_piers: int = 0,
                 device: Optional[int] = 0):
        self._lib = load_library()
        self.config = JgConfig(int(quadtree_depth), int(cd_threshold), int(n_ff_poles), int(n_ff_piers))
        self._cde = _vp()
        self._eng = _vp()
        b = _f32(bbox)
        rc = self._lib.jg_cde_create(b, C.byref(self.config), C.byref(self._cde))
        if rc != JG_OK:
            raise JaguaCudaError(rc, "jg_cde_create: invalid bbox / config")
        self._dirty = True
        self.device = device
        if device is not None:
            rc = self._lib.jg_engine_create(int(device), C.byref(self.config), C.byref(self._eng))
            if rc != JG_OK:
                self._lib.jg_cde_destroy(self._cde)
                self._cde = _vp()
                raise JaguaCudaError(rc, "jg_engine_create failed: no usable CUDA device (this engine has no CPU fallback)")
        self._items = {}
        self._filters = {}

    # -- lifetime ---------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_eng", None) and self._eng.value:
            self._lib.jg_engine_destroy(self._eng)
            self._eng = _vp()
        if getattr(self, "_cde", None) and self._cde.value:
            self._lib.jg_cde_destroy(self._cde)
            self._cde = _vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc: int) -> int:
        if rc < 0:
            msg = self._lib.jg_last_error(self._eng).decode() if self._eng.value else ""
            if not msg and self._cde.value:
                msg = self._lib.jg_cde_last_error(self._cde).decode()
            raise JaguaCudaError(rc, msg)
        return rc

    def _need_engine(self):
        if not self._eng.value:
            raise JaguaCudaError(JG_ERR_CUDA, "this CDEngine was created without a device; queries need a B200")

    # -- update path (cd_engine.rs:59-149) -----------------------------------------------------
    def register_hazard(self, kind: int, verts, poi, *, id: int = 0, quality: int = 0, pk: int = 0, dynamic: bool = True) -> int:
        """CDEngine::register_hazard. Returns the HazKey (slot | version << 32)."""
        v = _f32(verts).reshape(-1, 2)
        key = C.c_uint64(0)
        rc = self._lib.jg_cde_register_hazard(self._cde, int(kind), int(id), int(quality), int(pk), v, len(v),
                                              _f32(poi)[:2].copy(), 1 if dynamic else 0, C.byref(key))
        self._dirty = True  # (also when the mirror reports a reference panic: the hazard is in the tree by then)
        if rc < 0:
            raise JaguaCudaError(rc, self._lib.jg_cde_last_error(self._cde).decode())
        return int(key.value)

    def place_item(self, item_id: int, shape: "ItemShape", transform4, pk: int) -> int:
        """Layout::place_item (entities/layout.rs:40-60): registers the item's shape under the composed
        transformation (cos, sin, tx, ty) as a PlacedItem hazard with PItemKey `pk`. Returns the HazKey."""
        v = _f32(shape.verts).reshape(-1, 2)
        key = C.c_uint64(0)
        rc = self._lib.jg_cde_place_item(self._cde, int(item_id), int(pk), v, len(v), _f32(shape.poi)[:2].copy(),
                                         _f32(transform4).reshape(4).copy(), C.byref(key))
        self._dirty = True
        if rc < 0:
            raise JaguaCudaError(rc, self._lib.jg_cde_last_error(self._cde).decode())
        return int(key.value)

    def remove_item(self, pk: int):
        """Layout::remove_item (entities/layout.rs:62-74)."""
        rc = self._lib.jg_cde_remove_item(self._cde, int(pk))
        if rc < 0:
            raise JaguaCudaError(rc, f"remove_item: no placed item with key {pk}")
        self._dirty = True

    def deregister_hazard_by_key(self, hkey: int):
        rc = self._lib.jg_cde_deregister_hazard(self._cde, int(hkey))
        if rc < 0:
            raise JaguaCudaError(rc, self._lib.jg_cde_last_error(self._cde).decode())
        self._dirty = True

    def deregister_hazard_by_entity(self, kind: int, *, id: int = 0, quality: int = 0, pk: int = 0):
        """CDEngine::deregister_hazard_by_entity (cd_engine.rs:72-85)."""
        rc = self._lib.jg_cde_deregister_hazard_by_entity(self._cde, int(kind), int(id), int(quality), int(pk))
        if rc < 0:
            raise JaguaCudaError(rc, self._lib.jg_cde_last_error(self._cde).decode())
        self._dirty = True

    def swap_container(self, base: "CDEngine"):
        """Layout::swap_container (entities/layout.rs:40-50): `base` is the new container's base CDE (bbox, config, static
        hazards); this engine becomes a clone of it with its own dynamic hazards re-registered."""
        rc = self._lib.jg_cde_swap_container(self._cde, base._cde)
        self._dirty = True
        if rc < 0:
            raise JaguaCudaError(rc, self._lib.jg_cde_last_error(self._cde).decode())

    def epoch(self) -> int:
        return int(self._lib.jg_cde_epoch(self._cde))

    def haz_key_from_pi_key(self, pk: int) -> Optional[int]:
        key = C.c_uint64(0)
        self._lib.jg_cde_hkey_from_pk(self._cde, int(pk), C.byref(key))
        return int(key.value) or None

    def save(self) -> int:
        sid = C.c_int(0)
        self._lib.jg_cde_save(self._cde, C.byref(sid))
        return sid.value

    def restore(self, snapshot: int):
        rc = self._lib.jg_cde_restore(self._cde, int(snapshot))
        if rc < 0:
            raise JaguaCudaError(rc, self._lib.jg_cde_last_error(self._cde).decode())
        self._dirty = True

    def slot_capacity(self) -> int:
        return self._lib.jg_cde_slot_capacity(self._cde)

    def n_hazards(self) -> int:
        return self._lib.jg_cde_n_hazards(self._cde)

    def hkey_exterior(self) -> int:
        return int(self._lib.jg_cde_hkey_exterior(self._cde))

    def collector_words(self) -> int:
        return max(1, (self.slot_capacity() + 31) // 32)

    def filter_from_irrelevant_qzones(self, required_quality: int) -> np.ndarray:
        """HazKeyFilter::from_irrelevant_qzones (hazards/filter.rs:21-41) as a slot bitset."""
        bits = np.zeros(self.collector_words(), np.uint32)
        self._lib.jg_cde_filter_qzones(self._cde, int(required_quality), bits, len(bits))
        return bits

    def filter_from_keys(self, hkeys: Sequence[int]) -> np.ndarray:
        """HazKeyFilter::from_keys / `impl HazardFilter for HazKey` (filter.rs:43-76)."""
        bits = np.zeros(self.collector_words(), np.uint32)
        for k in hkeys:
            s = int(k) & 0xFFFFFFFF
            bits[s >> 5] |= np.uint32(1 << (s & 31))
        return bits

    def flatten(self) -> np.ndarray:
        n = self._lib.jg_cde_flatten(self._cde, None, 0)
        if n < 0:
            raise JaguaCudaError(int(n), self._lib.jg_cde_last_error(self._cde).decode())
        buf = np.zeros(n, np.uint8)
        self._lib.jg_cde_flatten(self._cde, buf.ctypes.data, n)
        return buf

    def image_sync(self):
        """jg_cde_image_sync: (image bytes as a uint8 array copy, lineage, seq, ranges as an (n, 2) uint32 array of (offset, length))."""
        img, nbytes, lin, seq, rng, n = _vp(), C.c_size_t(0), C.c_uint64(0), C.c_uint64(0), _vp(), C.c_int(0)
        rc = self._lib.jg_cde_image_sync(self._cde, C.byref(img), C.byref(nbytes), C.byref(lin), C.byref(seq), C.byref(rng), C.byref(n))
        if rc < 0:
            raise JaguaCudaError(rc, self._lib.jg_cde_last_error(self._cde).decode())
        blob = np.frombuffer((C.c_char * nbytes.value).from_address(img.value), np.uint8).copy()
        ranges = (np.frombuffer((C.c_uint32 * (2 * n.value)).from_address(rng.value), np.uint32).reshape(-1, 2).copy()
                  if n.value else np.zeros((0, 2), np.uint32))
        return blob, int(lin.value), int(seq.value), ranges

    def tree_dump(self) -> np.ndarray:
        n = self._lib.jg_cde_tree_dump(self._cde, None, 0)
        buf = np.zeros(n, np.uint32)
        self._lib.jg_cde_tree_dump(self._cde, buf.ctypes.data, n)
        return buf

    # -- device residency ---------------------------------------------------------------------
    def upload_item(self, item_id: int, shape: ItemShape):
        self._need_engine()
        v = _f32(shape.verts).reshape(-1, 2)
        poles = _f32(shape.poles).reshape(-1, 3)
        piers = _f32(shape.piers if shape.piers is not None else np.zeros((0, 4))).reshape(-1, 4)
        self._chk(self._lib.jg_item_upload(self._eng, int(item_id), v, len(v), _f32(shape.poi)[:2].copy(),
                                           poles if len(poles) else np.zeros((1, 3), np.float32), len(poles),
                                           piers if len(piers) else np.zeros((1, 4), np.float32), len(piers)))
        self._items[int(item_id)] = len(v)

    def upload_filter(self, filter_id: int, bits: np.ndarray):
        self._need_engine()
        bits = np.ascontiguousarray(bits, dtype=np.uint32)
        self._chk(self._lib.jg_filter_upload(self._eng, int(filter_id), bits, len(bits)))
        self._filters[int(filter_id)] = len(bits)

    def upload_filter_keys(self, filter_id: int, hkeys: Sequence[int]):
        """HazKeyFilter by versioned HazKeys (jg_filter_upload_keys): survives layout changes - a key whose slot was freed
        or reused matches nothing, as in the reference."""
        self._need_engine()
        k = np.ascontiguousarray(np.asarray(list(hkeys), dtype=np.uint64))
        self._chk(self._lib.jg_filter_upload_keys(self._eng, int(filter_id), k.ctypes.data if len(k) else None, len(k)))
        self._filters[int(filter_id)] = -1

    def sync_layout(self, force: bool = False) -> int:
        """Flatten the host quadtree and upload it if it changed since the last upload."""
        self._need_engine()
        if self._dirty or force:
            ep = C.c_uint64(0)
            self._chk(self._lib.jg_cde_sync_to_engine(self._cde, self._eng, C.byref(ep)))
            self._dirty = False
        return int(self._lib.jg_engine_layout_epoch(self._eng))

    def upload_layout_blob(self, blob: np.ndarray, epoch: int = 0):
        """jg_layout_upload of an externally produced blob (e.g. received from rank 0)."""
        self._need_engine()
        blob = np.ascontiguousarray(blob, dtype=np.uint8)
        self._chk(self._lib.jg_layout_upload(self._eng, blob.ctypes.data, blob.nbytes, int(epoch)))
        self._dirty = False

    def upload_layout_device(self, d_ptr: int, nbytes: int, epoch: int = 0):
        self._need_engine()
        self._chk(self._lib.jg_layout_upload_device(self._eng, int(d_ptr), int(nbytes), int(epoch)))
        self._dirty = False

    # -- queries ----------------------------------------------------------------------------
    def detect_batch(self, item_id: int, transforms, filter_id: int = -1, poly_only: bool = False,
                     out: Optional[np.ndarray] = None) -> np.ndarray:
        """detect_surrogate_collision then detect_poly_collision (cd_engine.rs:160-223) for every
        (cos, sin, tx, ty) row. Host buffers in, verdict bytes out (see JG_VERDICT_* bits)."""
        self.sync_layout()
        t = transforms if isinstance(transforms, np.ndarray) and transforms.dtype == np.float32 and transforms.flags.c_contiguous else _f32(transforms)
        n = t.size // 4
        if out is None:
            out = np.empty(n, np.uint8)
        self._chk(self._lib.jg_detect_batch(self._eng, int(item_id), t.ctypes.data, n, int(filter_id),
                                            MODE_POLY_ONLY if poly_only else MODE_DEFAULT, out.ctypes.data))
        return out

    def collect_batch(self, item_id: int, transforms, words: Optional[int] = None, poly_only: bool = False,
                      out: Optional[np.ndarray] = None) -> np.ndarray:
        """collect_surrogate_collisions + collect_poly_collisions (cd_engine.rs:261-310) into one
        bitset over hazard slots per query."""
        self.sync_layout()
        t = transforms if isinstance(transforms, np.ndarray) and transforms.dtype == np.float32 and transforms.flags.c_contiguous else _f32(transforms)
        n = t.size // 4
        words = int(words or self.collector_words())
        if out is None:
            out = np.empty((n, words), np.uint32)
        self._chk(self._lib.jg_collect_batch(self._eng, int(item_id), t.ctypes.data, n,
                                             MODE_POLY_ONLY if poly_only else MODE_DEFAULT, out.ctypes.data, words))
        return out

    def detect_batch_async(self, item_id: int, transforms: np.ndarray, out: np.ndarray, filter_id: int = -1, poly_only: bool = False):
        """detect_batch without the final wait: `transforms` (float32, C-contiguous) and `out` must stay
        alive — and should be pinned (PinnedBuffer) — until sync()."""
        self.sync_layout()
        assert transforms.dtype == np.float32 and transforms.flags.c_contiguous and out.dtype == np.uint8
        self._chk(self._lib.jg_detect_batch_async(self._eng, int(item_id), transforms.ctypes.data, transforms.size // 4,
                                                  int(filter_id), MODE_POLY_ONLY if poly_only else MODE_DEFAULT, out.ctypes.data))

    def collect_batch_async(self, item_id: int, transforms: np.ndarray, out: np.ndarray, poly_only: bool = False):
        self.sync_layout()
        assert transforms.dtype == np.float32 and transforms.flags.c_contiguous and out.dtype == np.uint32 and out.ndim == 2
        self._chk(self._lib.jg_collect_batch_async(self._eng, int(item_id), transforms.ctypes.data, transforms.size // 4,
                                                   MODE_POLY_ONLY if poly_only else MODE_DEFAULT, out.ctypes.data, out.shape[1]))

    @staticmethod
    def _segments(item_ids, seg_offsets):
        ids = np.ascontiguousarray(item_ids, dtype=np.int32)
        offs = np.ascontiguousarray(seg_offsets, dtype=np.int64)
        assert len(offs) == len(ids) + 1 and offs[0] == 0
        return ids, offs

    def detect_multi(self, item_ids, seg_offsets, transforms: np.ndarray, filter_id: int = -1, poly_only: bool = False,
                     out: Optional[np.ndarray] = None, wait: bool = True) -> np.ndarray:
        """detect for a batch of segments: queries seg_offsets[k]..seg_offsets[k+1] place item item_ids[k].
        One launch sequence for all segments. With wait=False the call returns after enqueueing
        (transforms / out must stay alive, ideally pinned, until sync())."""
        self.sync_layout()
        ids, offs = self._segments(item_ids, seg_offsets)
        t = transforms if isinstance(transforms, np.ndarray) and transforms.dtype == np.float32 and transforms.flags.c_contiguous else _f32(transforms)
        assert t.size // 4 == int(offs[-1])
        if out is None:
            out = np.empty(int(offs[-1]), np.uint8)
        fn = self._lib.jg_detect_multi if wait else self._lib.jg_detect_multi_async
        self._chk(fn(self._eng, len(ids), ids.ctypes.data, offs.ctypes.data, t.ctypes.data, int(filter_id),
                     MODE_POLY_ONLY if poly_only else MODE_DEFAULT, out.ctypes.data))
        return out

    def collect_multi(self, item_ids, seg_offsets, transforms: np.ndarray, words: Optional[int] = None, poly_only: bool = False,
                      seed: Optional[np.ndarray] = None, out: Optional[np.ndarray] = None, wait: bool = True) -> np.ndarray:
        """collect for a batch of segments in ONE launch (jg_collect_multi). `seed` (n x words) pre-seeds the collectors:
        hazards already present are skipped by the traversal and stay in the result (collector.contains_key)."""
        self.sync_layout()
        ids, offs = self._segments(item_ids, seg_offsets)
        t = transforms if isinstance(transforms, np.ndarray) and transforms.dtype == np.float32 and transforms.flags.c_contiguous else _f32(transforms)
        n = int(offs[-1])
        assert t.size // 4 == n
        words = int(words or self.collector_words())
        if out is None:
            out = np.empty((n, words), np.uint32)
        if seed is not None:
            seed = np.ascontiguousarray(seed, dtype=np.uint32)
            assert seed.shape == (n, words)
        fn = self._lib.jg_collect_multi if wait else self._lib.jg_collect_multi_async
        self._chk(fn(self._eng, len(ids), ids.ctypes.data, offs.ctypes.data, t.ctypes.data, MODE_POLY_ONLY if poly_only else MODE_DEFAULT,
                     seed.ctypes.data if seed is not None else None, out.ctypes.data, words))
        return out

    def collect_multi_device(self, item_ids, seg_offsets, d_transforms: int, d_out_bits: int, words: int, d_seed: int = 0,
                             poly_only: bool = False):
        self.sync_layout()
        ids, offs = self._segments(item_ids, seg_offsets)
        self._chk(self._lib.jg_collect_multi_device(self._eng, len(ids), ids.ctypes.data, offs.ctypes.data, int(d_transforms),
                                                    MODE_POLY_ONLY if poly_only else MODE_DEFAULT, int(d_seed) or None, int(d_out_bits), int(words)))

    def detect_multi_lattice(self, item_ids, seg_offsets, lattice: JgLattice, rot_idx: np.ndarray, ij: np.ndarray, filter_id: int = -1,
                             poly_only: bool = False, out: Optional[np.ndarray] = None, wait: bool = True) -> np.ndarray:
        """detect_multi with the queries as (rotation index u8, lattice coordinates 2 x u16) = 5 B/query."""
        self.sync_layout()
        ids, offs = self._segments(item_ids, seg_offsets)
        assert rot_idx.dtype == np.uint8 and rot_idx.flags.c_contiguous and ij.dtype == np.uint16 and ij.flags.c_contiguous
        assert rot_idx.size == int(offs[-1]) and ij.size == 2 * int(offs[-1])
        if out is None:
            out = np.empty(int(offs[-1]), np.uint8)
        self._lat_keep = lattice
        fn = self._lib.jg_detect_multi_lattice if wait else self._lib.jg_detect_multi_lattice_async
        self._chk(fn(self._eng, len(ids), ids.ctypes.data, offs.ctypes.data, C.addressof(lattice), rot_idx.ctypes.data, ij.ctypes.data,
                     int(filter_id), MODE_POLY_ONLY if poly_only else MODE_DEFAULT, out.ctypes.data))
        return out

    def set_item_rotations(self, item_id: int, rotations):
        """Registers an item's allowed orientations (radians) for the compact query form. (cos, sin) are
        composed on the host by the same libm call as every other transformation (compose())."""
        self._need_engine()
        rots = np.asarray(rotations, np.float32).reshape(-1)
        cs = np.ascontiguousarray(compose(np.stack([rots, np.zeros_like(rots), np.zeros_like(rots)], 1))[:, :2])
        self._chk(self._lib.jg_item_set_rotations(self._eng, int(item_id), cs.ctypes.data, len(rots)))

    def detect_multi_indexed(self, item_ids, seg_offsets, rot_idx: np.ndarray, txy: np.ndarray, filter_id: int = -1,
                             poly_only: bool = False, out: Optional[np.ndarray] = None, wait: bool = True) -> np.ndarray:
        """detect_multi with the queries as (index into the item's allowed rotations: u8, translation: 2 x f32)."""
        self.sync_layout()
        ids, offs = self._segments(item_ids, seg_offsets)
        assert rot_idx.dtype == np.uint8 and rot_idx.flags.c_contiguous and txy.dtype == np.float32 and txy.flags.c_contiguous
        assert rot_idx.size == int(offs[-1]) and txy.size == 2 * int(offs[-1])
        if out is None:
            out = np.empty(int(offs[-1]), np.uint8)
        fn = self._lib.jg_detect_multi_indexed if wait else self._lib.jg_detect_multi_indexed_async
        self._chk(fn(self._eng, len(ids), ids.ctypes.data, offs.ctypes.data, rot_idx.ctypes.data, txy.ctypes.data, int(filter_id),
                     MODE_POLY_ONLY if poly_only else MODE_DEFAULT, out.ctypes.data))
        return out

    def detect_multi_device(self, item_ids, seg_offsets, d_transforms: int, d_out: int, filter_id: int = -1, poly_only: bool = False):
        self.sync_layout()
        ids, offs = self._segments(item_ids, seg_offsets)
        self._chk(self._lib.jg_detect_multi_device(self._eng, len(ids), ids.ctypes.data, offs.ctypes.data, int(d_transforms),
                                                   int(filter_id), MODE_POLY_ONLY if poly_only else MODE_DEFAULT, int(d_out)))

    def detect_batch_device(self, item_id: int, d_transforms: int, n: int, d_out: int, filter_id: int = -1,
                            poly_only: bool = False):
        """Same as detect_batch with device pointers; asynchronous (call sync())."""
        self.sync_layout()
        self._chk(self._lib.jg_detect_batch_device(self._eng, int(item_id), int(d_transforms), int(n), int(filter_id),
                                                   MODE_POLY_ONLY if poly_only else MODE_DEFAULT, int(d_out)))

    def collect_batch_device(self, item_id: int, d_transforms: int, n: int, d_out_bits: int, words: int, poly_only: bool = False):
        self.sync_layout()
        self._chk(self._lib.jg_collect_batch_device(self._eng, int(item_id), int(d_transforms), int(n),
                                                    MODE_POLY_ONLY if poly_only else MODE_DEFAULT, int(d_out_bits), int(words)))

    def search_round(self, item_id: int, n: int, *, box=None, mean=None, sd: float = 0.0, rot_idx: int = -1, seed: int = 0,
                     counter: int = 0, filter_id: int = -1):
        """One batched decision point of the LBF search, samples drawn on the device (jg_search_round): uniform over `box`
        when `mean` is None, else normal(mean, sd) around the reference translation with rotation `rot_idx`. Returns
        dict(n_free, index, loss, rot_idx, tx, ty); index == -1 when no sample is free."""
        self.sync_layout()
        r = JgSearchRound()
        r.mode = 0 if mean is None else 1
        r.rot_idx = int(rot_idx)
        if box is not None:
            r.box = (C.c_float * 4)(*[float(v) for v in box])
        if mean is not None:
            r.mean = (C.c_float * 2)(float(mean[0]), float(mean[1]))
        r.sd = float(sd)
        r.seed, r.counter = int(seed), int(counter)
        b = JgSearchBest()
        self._chk(self._lib.jg_search_round(self._eng, int(item_id), C.addressof(r), int(n), int(filter_id), C.addressof(b)))
        return dict(n_free=int(b.n_free), index=int(b.index), loss=float(b.loss), rot_idx=int(b.rot_idx), tx=float(b.tx), ty=float(b.ty))

    def search_samples(self, n: int):
        """(rot_idx u8, txy (n, 2) f32, verdict bytes) of the last search_round (tests)."""
        ridx, txy, v = np.zeros(n, np.uint8), np.zeros((n, 2), np.float32), np.zeros(n, np.uint8)
        self._chk(self._lib.jg_search_samples(self._eng, int(n), ridx.ctypes.data, txy.ctypes.data, v.ctypes.data))
        return ridx, txy, v

    def fence(self) -> int:
        """Marks everything enqueued so far; wait_fence(token) blocks until that much is done (later batches keep running)."""
        tok = C.c_uint64(0)
        self._chk(self._lib.jg_fence_record(self._eng, C.byref(tok)))
        return int(tok.value)

    def wait_fence(self, token: int):
        self._chk(self._lib.jg_fence_wait(self._eng, int(token)))

    def sync(self):
        self._need_engine()
        self._chk(self._lib.jg_sync(self._eng))

    def last_kernel_ms(self) -> float:
        return float(self._lib.jg_engine_last_kernel_ms(self._eng))

    def timer_start(self, gated: bool = False):
        """CUDA-event stopwatch on the compute stream; gated: the stream is held until timer_stop, so the launches enqueued in
        between run back to back on the device (the host's enqueue time is not measured)."""
        self._chk((self._lib.jg_timer_start_gated if gated else self._lib.jg_timer_start)(self._eng))

    def timer_stop(self) -> float:
        ms = C.c_float(0)
        self._chk(self._lib.jg_timer_stop(self._eng, C.byref(ms)))
        return float(ms.value)

    def stats(self) -> dict:
        a = (C.c_uint64 * 8)()
        self._lib.jg_engine_stats(self._eng, a)
        return dict(launches=int(a[0]), queries=int(a[1]), h2d_bytes=int(a[2]), d2h_bytes=int(a[3]),
                    layout_uploads=int(a[4]), survivors=int(a[5]), fallbacks=int(a[6]), patched_uploads=int(a[7]))

    def set_batching(self, chunk_queries: int, n_streams: int):
        self._chk(self._lib.jg_engine_set_batching(self._eng, int(chunk_queries), int(n_streams)))

    def set_split_min(self, min_queries: int):
        """Launch size from which the containment tests run in their own kernel (0 = always)."""
        self._chk(self._lib.jg_engine_set_split_min(self._eng, int(min_queries)))

    def set_dependent_launch(self, on: bool):
        """Stage 2 of the grid path as a programmatic dependent launch behind stage 1 (process-wide)."""
        self._chk(self._lib.jg_set_dependent_launch(1 if on else 0))

    def set_groups(self, detect_group: int = 0, collect_group: int = 0):
        self._chk(self._lib.jg_engine_set_groups(self._eng, int(detect_group), int(collect_group)))

    def set_detect_path(self, walk: bool = False, fused: bool = False):
        """JG_PATH_WALK keeps the quadtree-walk kernels even where the grid kernels apply; JG_PATH_FUSED selects
        the experimental single-kernel, copy-following form of the grid path. Same results in every case."""
        self._chk(self._lib.jg_engine_set_detect_path(self._eng, 1 if walk else (2 if fused else 0)))

    def detect_path(self) -> str:
        return ("walk", "grid", "fused")[self._lib.jg_engine_detect_path(self._eng)]

    def fp32_nofma_peak(self):
        g, ms = C.c_double(0), C.c_float(0)
        self._chk(self._lib.jg_fp32_nofma_peak(self._eng, C.byref(g), C.byref(ms)))
        return g.value, ms.value

    def flush_l2(self):
        self._chk(self._lib.jg_flush_l2(self._eng))


from .pool import EnginePool  # noqa: E402
from . import instance  # noqa: E402
from . import export  # noqa: E402
from . import workload  # noqa: E402
from .workload import (Workload, load_workload, build_engine, make_queries, make_queries_indexed, make_queries_lattice,  # noqa: E402
                       compose_lattice_queries, split_batch)
