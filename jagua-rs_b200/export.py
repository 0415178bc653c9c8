"""Solution export in the reference's external representation (the JSON `lbf` writes next to its SVGs).

    jagua-rs/src/io/export.rs:6-45                 export_layout_snapshot, int_to_ext_transformation
    jagua-rs/src/io/ext_repr.rs:84-121             ExtLayout, ExtPlacedItem, ExtTransformation (rotation in degrees)
    jagua-rs/src/geometry/transformation.rs:47-51,83-88,131-139,164-172   rotate_translate, decompose, rot_transl_m, dot_prod
    jagua-rs/src/probs/spp/io/{export.rs,ext_repr.rs}   ExtSPSolution {strip_width, layout, density, run_time_sec}
    jagua-rs/src/probs/bpp/io/{export.rs,ext_repr.rs}   ExtBPSolution {cost, layouts, density, run_time_sec}

A placement lives in the engine as a DTransformation of the *centred* item shape (io/import.rs:214-217); the external
transformation applies to the item's original shape: pre-transform first, internal transformation second, composed as
3x3 f32 matrices and decomposed with atan2f. Everything rounds to f32 like the reference; sin / cos / atan2 come from the
platform libm (what f32::sin_cos and f32::atan2 lower to on linux-gnu).
"""
from __future__ import annotations

import ctypes
import ctypes.util
from typing import Iterable, List, Optional

import numpy as np

F = np.float32
_libm = ctypes.CDLL(ctypes.util.find_library("m"))
for _fn, _n in (("sinf", 1), ("cosf", 1), ("atan2f", 2)):
    getattr(_libm, _fn).restype = ctypes.c_float
    getattr(_libm, _fn).argtypes = [ctypes.c_float] * _n
PIS_IN_180 = F(57.2957795130823208767981548141051703)  # f32::to_degrees multiplies by this constant


def _rot_transl_m(rot, tx, ty):
    sn, cs = F(_libm.sinf(float(rot))), F(_libm.cosf(float(rot)))
    return [[cs, F(-sn), F(tx)], [sn, cs, F(ty)], [F(0.0), F(0.0), F(1.0)]]


def _dot_prod(l, r):
    return [[F(F(F(l[i][0] * r[0][j]) + F(l[i][1] * r[1][j])) + F(l[i][2] * r[2][j])) for j in range(3)] for i in range(3)]


def int_to_ext_transformation(int_dt, pre_dt):
    """(rotation rad, tx, ty) of the centred shape + the item's pre-transform (rotation, tx, ty) -> the DTransformation of
    the original shape (io/export.rs:31-45): Transformation::empty().transform_from_decomposed(pre).transform_from_decomposed(int).decompose()."""
    m = [[F(1.0), F(0.0), F(0.0)], [F(0.0), F(1.0), F(0.0)], [F(0.0), F(0.0), F(1.0)]]
    with np.errstate(all="ignore"):
        m = _dot_prod(_rot_transl_m(*[F(v) for v in pre_dt]), m)
        m = _dot_prod(_rot_transl_m(*[F(v) for v in int_dt]), m)
    angle = F(_libm.atan2f(float(m[1][0]), float(m[0][0])))
    return np.array([angle, m[0][2], m[1][2]], F)


def ext_transformation(dt) -> dict:
    """ExtTransformation::from(DTransformation) (io/ext_repr.rs:101-108): rotation in degrees."""
    return {"rotation": float(F(F(dt[0]) * PIS_IN_180)), "translation": [float(F(dt[1])), float(F(dt[2]))]}


def export_layout(container_id: int, placements: Iterable, items: List, container_area: float) -> dict:
    """export_layout_snapshot (io/export.rs:6-29). placements: objects with .item_id and .d_transf; items: the imported
    items (InstItem / LbfItem: .pre_transform = (tx, ty) of the centring translation, .area = the original shape's)."""
    placed, area = [], F(0.0)
    for p in placements:
        it = items[p.item_id]
        pre = np.asarray(getattr(it, "pre_transform", getattr(it, "pre", (0.0, 0.0))), F).reshape(-1)
        ext = int_to_ext_transformation(p.d_transf, (F(0.0), pre[-2], pre[-1]))
        placed.append({"item_id": int(p.item_id), "transformation": ext_transformation(ext)})
        area = F(area + F(it.area))  # Layout::placed_item_area: an f32 sum in placement order
    return {"container_id": int(container_id), "placed_items": placed, "density": float(F(area / F(container_area)))}


def export_spp_solution(result, items: List, run_time_sec: Optional[float] = None) -> dict:
    """ExtSPSolution (probs/spp/io/export.rs) from a BatchedLBF result: the strip is container 0, its area width x height."""
    lay = export_layout(0, result.placements, items, float(F(F(result.strip_width) * F(result.strip_height))))
    return {"strip_width": float(F(result.strip_width)), "layout": lay, "density": lay["density"],
            "run_time_sec": int(result.seconds if run_time_sec is None else run_time_sec)}


def export_bpp_solution(layouts: Iterable, items: List, bins_by_id: dict, run_time_sec: float = 0.0) -> dict:
    """ExtBPSolution (probs/bpp/io/export.rs). layouts: objects with .bin_id and .placements; bins_by_id[bin_id] =
    {"area": original container area, "cost": bin cost}."""
    outs, cost, used, avail = [], 0, F(0.0), F(0.0)
    for lay in layouts:
        b = bins_by_id[lay.bin_id]
        e = export_layout(lay.bin_id, lay.placements, items, b["area"])
        outs.append(e)
        cost += int(b.get("cost", 0))
        for p in lay.placements:
            used = F(used + F(items[p.item_id].area))
        avail = F(avail + F(b["area"]))
    return {"cost": cost, "layouts": outs, "density": float(F(used / avail)) if avail > 0 else 0.0, "run_time_sec": int(run_time_sec)}
