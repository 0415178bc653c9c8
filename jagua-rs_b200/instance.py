"""Instance files of the reference (io/ext_repr.rs:6-121, probs/{spp,bpp}/io/ext_repr.rs) on the product side: what the
importer turns them into before any collision query runs (io/import.rs:47-195, entities/container.rs:28-133).

Shapes are parsed with serde's f32 semantics (decimal -> f64 -> f32). Items go through jg_item_import[_simplified],
container boundaries / holes / quality zones through jg_shape_import (product C++). Item simplification and narrow-
concavity closing (geometry/shape_modification.rs) are optional, as in the reference's Importer; offsetting (geo-buffer)
and the containers' own modification are not applied. `container_engine` registers the static hazards in the order
Container::new does - Exterior, then per quality 0..9 the zones in input order, the holes of a `polygon` outer shape
appended to quality 0 (io/import.rs:166-167) - so HazKeys (slot numbers) match the reference's.
"""
from __future__ import annotations

import json
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

N_QUALITIES = 10  # entities/container.rs:88


@dataclass
class InstItem:
    id: int
    shape: object                    # ItemShape (centred vertices, poi, surrogate poles)
    rotations: Optional[np.ndarray]  # radians (f32); None = continuous; empty = no rotation
    demand: int
    min_quality: Optional[int]
    area: float
    diameter: float
    pre_transform: np.ndarray


@dataclass
class Instance:
    name: str
    items: List[InstItem] = field(default_factory=list)
    strip_height: Optional[float] = None
    bins: List[dict] = field(default_factory=list)  # raw ExtContainer records, sorted by id


def shape_points(shape: dict):
    """ExtShape -> (outer points f32, [hole points]) (io/ext_repr.rs:34-52)."""
    t, d = shape["type"], shape["data"]
    if t == "simple_polygon":
        return np.array(d, np.float64).astype(np.float32), []
    if t == "polygon":
        return (np.array(d["outer"], np.float64).astype(np.float32),
                [np.array(h, np.float64).astype(np.float32) for h in d.get("inner", [])])
    if t == "rectangle":
        x0, y0 = np.float32(d["x_min"]), np.float32(d["y_min"])
        x1, y1 = np.float32(x0 + np.float32(d["width"])), np.float32(y0 + np.float32(d["height"]))
        return np.array([[x0, y0], [x1, y0], [x1, y1], [x0, y1]], np.float32), []
    raise ValueError(f"unsupported shape type {t}")


def import_items(records, simplify_tolerance=None, narrow_concavity_cutoff_ratio=None) -> List[InstItem]:
    """Importer::import_item for every record with demand > 0, ids renumbered densely in id order. simplify_tolerance /
    narrow_concavity_cutoff_ratio = Importer's shape modification (lbf defaults 0.001 / 0.01; None = unmodified shapes)."""
    from . import import_item

    items = []
    for it in sorted(records, key=lambda it: it["id"]):
        if it.get("demand", 1) <= 0:
            continue
        pts, _ = shape_points(it["shape"])  # holes of a `polygon` item are ignored, as the reference warns
        shape, pre, area = import_item(pts, simplify_tolerance=simplify_tolerance, narrow_concavity_cutoff_ratio=narrow_concavity_cutoff_ratio)
        ao = it.get("allowed_orientations")
        if ao is None:
            rots = None  # RotationRange::Continuous
        elif len(ao) == 0 or (len(ao) == 1 and ao[0] == 0.0):
            rots = np.zeros(0, np.float32)  # RotationRange::None
        else:  # f32::to_radians = deg * (PI / 180) in f32 (io/import.rs:80-89)
            rots = (np.array(ao, np.float64).astype(np.float32) * (np.float32(np.pi) / np.float32(180.0))).astype(np.float32)
        v = shape.verts.astype(np.float64)
        diam = float(np.sqrt(((v[:, None, :] - v[None, :, :]) ** 2).sum(-1).max()))
        items.append(InstItem(len(items), shape, rots, int(it.get("demand", 1)), it.get("min_quality"), area, diam, pre))
    return items


def load_instance(path_or_dict, simplify_tolerance=None, narrow_concavity_cutoff_ratio=None) -> Instance:
    if isinstance(path_or_dict, dict):
        d = path_or_dict
    else:
        with open(path_or_dict) as f:
            d = json.load(f)
    inst = Instance(str(d.get("name", "")), import_items(d["items"], simplify_tolerance, narrow_concavity_cutoff_ratio))
    if "strip_height" in d:
        inst.strip_height = float(np.float32(d["strip_height"]))
    inst.bins = sorted(d.get("bins", []), key=lambda b: b["id"])
    return inst


def container_hazards(ext_container: dict):
    """import_container + Container::new: (cde bbox, [(kind, idx, quality, vertices, poi)]) with the Exterior first."""
    from . import HAZ_EXTERIOR, HAZ_HOLE, HAZ_QZONE, import_shape, polygon_poi

    def one(pts, is_rect):
        return (lambda r: (r[0], r[1]))(polygon_poi(pts)) if is_rect else import_shape(pts)

    outer_pts, hole_pts = shape_points(ext_container["shape"])
    outer, outer_poi = one(outer_pts, ext_container["shape"]["type"] == "rectangle")
    x0, y0, x1, y1 = outer[:, 0].min(), outer[:, 1].min(), outer[:, 0].max(), outer[:, 1].max()
    w, h = np.float32(x1 - x0), np.float32(y1 - y0)
    dx = dy = np.float32(0.0)
    if h < w:
        dy = np.float32((w - h) / np.float32(2.0))
    elif w < h:
        dx = np.float32((h - w) / np.float32(2.0))
    bbox = np.array([x0 - dx, y0 - dy, x1 + dx, y1 + dy], np.float32)  # rect.rs:93-109 inflate_to_square
    hazards = [(HAZ_EXTERIOR, 0, 0, outer, outer_poi)]
    zones = ext_container.get("zones", [])
    assert all(z["quality"] < N_QUALITIES for z in zones)
    for q in range(N_QUALITIES):
        shapes = [one(shape_points(z["shape"])[0], z["shape"]["type"] == "rectangle") for z in zones if z["quality"] == q]
        if q == 0:
            shapes += [import_shape(hp) for hp in hole_pts]
        for idx, (v, poi) in enumerate(shapes):
            hazards.append((HAZ_HOLE if q == 0 else HAZ_QZONE, idx, q, v, poi))
    return bbox, hazards


def container_engine(inst: Instance, bin_index: int = 0, device: Optional[int] = 0, **cfg):
    """A CDEngine holding the base CDE of one bin (Container::new, entities/container.rs:57-70) with the instance's
    items uploaded."""
    from . import CDEngine

    bbox, hazards = container_hazards(inst.bins[bin_index])
    cde = CDEngine(bbox, device=device, **cfg)
    cde.hazard_keys = [cde.register_hazard(kind, v, poi, id=idx, quality=q, dynamic=False) for kind, idx, q, v, poi in hazards]
    if device is not None:
        for it in inst.items:
            cde.upload_item(it.id, it.shape)
    return cde
