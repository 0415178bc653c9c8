"""TEST INFRASTRUCTURE - the checker for jg_item_generate_piers, never imported by the product.

A restatement of the reference's surrogate-pier generation in numpy float32 (every operation rounds to f32, no fused
multiply-add), written from the reference source and independent of the product's C++:

    jagua-rs/src/geometry/fail_fast/piers.rs:25-99     generate_piers
    :101-127 generate_ray_transformations   :130-157 clip   :159-178 generate_unrepresented_point_grid
    :180-202 loss_function   :204-226 min_distances_to_rays / min_distances_to_poles
    jagua-rs/src/geometry/transformation.rs:36-39,114-172   rotate, rot_m, transl_m, dot_prod
    jagua-rs/src/geometry/primitives/point.rs:10-31,43-47   TRANSFORM_FORMULA, distance_to
    jagua-rs/src/geometry/primitives/edge.rs:38-45,63-87,153-159,180-228   scale, closest_point_on_edge, sq_distance_to,
                                                                            edge_intersection with calc_loc = true
    jagua-rs/src/geometry/primitives/circle.rs:130-147      Circle: CollidesWith<Point>, DistanceTo<Point>
    jagua-rs/src/geometry/primitives/rect.rs:93-109         inflate_to_square
    jagua-rs/src/geometry/primitives/simple_polygon.rs:173-192   centroid

Third-party arithmetic: ndarray 0.17 `Array::linspace(a, b, n)` is restated as a + ((b - a) / (n - 1)) * i; `f32::sin_cos`
is taken from the platform libm (sinf / cosf through ctypes). Parity status: unpinned against the Rust crate (no toolchain
here); it pins the product against an independent second reading of the same source.
"""
import ctypes
import ctypes.util

import numpy as np

import simplify as S  # oracle/ is on sys.path (tests/conftest.py), modules are imported flat

F = np.float32
_libm = ctypes.CDLL(ctypes.util.find_library("m"))
_libm.sinf.restype = ctypes.c_float
_libm.sinf.argtypes = [ctypes.c_float]
_libm.cosf.restype = ctypes.c_float
_libm.cosf.argtypes = [ctypes.c_float]


def linspace(a, b, n):
    a, b = F(a), F(b)
    step = F(F(b - a) / F(n - 1)) if n > 1 else F(0.0)
    return [F(a + F(step * F(i))) for i in range(n)]


def dot_prod(l, r):
    o = [[F(0)] * 3 for _ in range(3)]
    for i in range(3):
        for j in range(3):
            o[i][j] = F(F(F(l[i][0] * r[0][j]) + F(l[i][1] * r[1][j])) + F(l[i][2] * r[2][j]))
    return o


def transform_point(t, p):
    x, y = p
    return (F(F(F(t[0][0] * x) + F(t[0][1] * y)) + F(t[0][2] * F(1.0))),
            F(F(F(t[1][0] * x) + F(t[1][1] * y)) + F(t[1][2] * F(1.0))))


def centroid(pts, area):
    cx = cy = F(0.0)
    n = len(pts)
    for i in range(n):
        j = 0 if i == n - 1 else i + 1
        (xi, yi), (xj, yj) = pts[i], pts[j]
        cr = F(F(xi * yj) - F(xj * yi))
        cx = F(cx + F(F(xi + xj) * cr))
        cy = F(cy + F(F(yi + yj) * cr))
    d = F(F(6.0) * area)
    return F(cx / d), F(cy / d)


def edge_distances(seg, px, py):
    """Edge::distance_to for one segment against arrays of points (edge.rs:63-87, 153-159)."""
    (x1, y1), (x2, y2) = seg
    a, b = (px - x1).astype(F), (py - y1).astype(F)
    c, d = F(x2 - x1), F(y2 - y1)
    dot = (a * c).astype(F) + (b * d).astype(F)
    len_sq = F(F(c * c) + F(d * d))
    with np.errstate(all="ignore"):
        param = (dot / len_sq).astype(F) if len_sq != 0 else np.full_like(px, F(-1.0))
    xx = np.where(param < 0, x1, np.where(param > 1, x2, (x1 + (param * c).astype(F)).astype(F))).astype(F)
    yy = np.where(param < 0, y1, np.where(param > 1, y2, (y1 + (param * d).astype(F)).astype(F))).astype(F)
    dx, dy = (px - xx).astype(F), (py - yy).astype(F)
    return np.sqrt(((dx * dx).astype(F) + (dy * dy).astype(F)).astype(F)).astype(F)


def edge_distances_many(segs, px, py):
    """Edge::distance_to for R segments x P points -> (R, P) float32."""
    x1, y1, x2, y2 = (segs[:, k:k + 1].astype(F) for k in range(4))
    a, b = (px[None, :] - x1).astype(F), (py[None, :] - y1).astype(F)
    c, d = (x2 - x1).astype(F), (y2 - y1).astype(F)
    dot = ((a * c).astype(F) + (b * d).astype(F)).astype(F)
    len_sq = ((c * c).astype(F) + (d * d).astype(F)).astype(F)
    with np.errstate(all="ignore"):
        param = np.where(len_sq != 0, (dot / len_sq).astype(F), F(-1.0)).astype(F)
    xx = np.where(param < 0, x1, np.where(param > 1, x2, (x1 + (param * c).astype(F)).astype(F))).astype(F)
    yy = np.where(param < 0, y1, np.where(param > 1, y2, (y1 + (param * d).astype(F)).astype(F))).astype(F)
    dx, dy = (px[None, :] - xx).astype(F), (py[None, :] - yy).astype(F)
    return np.sqrt(((dx * dx).astype(F) + (dy * dy).astype(F)).astype(F)).astype(F)


def collides_at(e1, e2):
    """edge.rs:180-228 with calc_loc = true: the location lies on e1."""
    (x1, y1), (x2, y2) = e1
    (x3, y3), (x4, y4) = e2
    if max(min(x1, x2), min(x3, x4)) > min(max(x1, x2), max(x3, x4)):
        return None
    if max(min(y1, y2), min(y3, y4)) > min(max(y1, y2), max(y3, y4)):
        return None
    t_nom = F(F(F(x2 - x4) * F(y4 - y3)) - F(F(y2 - y4) * F(x4 - x3)))
    den = F(F(F(x2 - x1) * F(y4 - y3)) - F(F(y2 - y1) * F(x4 - x3)))
    u_nom = F(F(F(x2 - x4) * F(y2 - y1)) - F(F(y2 - y4) * F(x2 - x1)))
    if den == 0:
        return None
    with np.errstate(all="ignore"):
        t, u = F(t_nom / den), F(u_nom / den)
    if not (0 <= t <= 1 and 0 <= u <= 1):
        return None
    return (F(x2 + F(t * F(x1 - x2))), F(y2 + F(t * F(y1 - y2))))


def generate_piers(verts, poles, n, rays_per_angle=200, n_angles=90, n_points_per_dimension=100):
    """-> (n, 4) float32 (sx, sy, ex, ey). verts: counter-clockwise (n_verts, 2); poles: (n_poles, 3) (x, y, r)."""
    if n == 0:
        return np.zeros((0, 4), F)
    pts = [(F(x), F(y)) for x, y in np.asarray(verts, F).reshape(-1, 2)]
    poles = [(F(x), F(y), F(r)) for x, y, r in np.asarray(poles, F).reshape(-1, 3)]
    nv = len(pts)
    area = S.area(pts)
    xs, ys = [p[0] for p in pts], [p[1] for p in pts]
    x_min, y_min, x_max, y_max = min(xs), min(ys), max(xs), max(ys)
    w, h = F(x_max - x_min), F(y_max - y_min)
    dx = dy = F(0.0)
    if h < w:
        dy = F(F(w - h) / F(2.0))
    elif w < h:
        dx = F(F(h - w) / F(2.0))
    sq = (F(x_min - dx), F(y_min - dy), F(x_max + dx), F(y_max + dy))
    sq_w, sq_h = F(sq[2] - sq[0]), F(sq[3] - sq[1])
    cx, cy = centroid(pts, area)
    base = ((cx, F(cy - F(F(2.0) * sq_h))), (cx, F(cy + F(F(2.0) * sq_h))))
    assert base[0] != base[1]
    dxr = F(sq_w / F(rays_per_angle))
    angles = linspace(0.0, np.float32(np.pi), n_angles + 1)[:n_angles]
    edges = [(pts[k], pts[0 if k == nv - 1 else k + 1]) for k in range(nv)]
    clipped = []
    one, zero = F(1.0), F(0.0)
    for ang in angles:
        sn, cs = F(_libm.sinf(float(ang))), F(_libm.cosf(float(ang)))
        rot = [[cs, F(-sn), zero], [sn, cs, zero], [zero, zero, one]]
        for i in range(rays_per_angle):
            x = F(sq[0] + F(dxr * F(i)))
            t = dot_prod(rot, [[one, zero, x], [zero, one, zero], [zero, zero, one]])
            ray = (transform_point(t, base[0]), transform_point(t, base[1]))
            assert not S.point_in_polygon(pts, ray[0]) and not S.point_in_polygon(pts, ray[1])
            # cheap exact pre-filter: the range rejects of edge_intersection against the whole polygon's bbox
            if max(min(ray[0][0], ray[1][0]), x_min) > min(max(ray[0][0], ray[1][0]), x_max) or \
               max(min(ray[0][1], ray[1][1]), y_min) > min(max(ray[0][1], ray[1][1]), y_max):
                continue
            hits = []
            for e in edges:
                at = collides_at(e, ray)
                if at is not None:
                    ddx, ddy = F(ray[0][0] - at[0]), F(ray[0][1] - at[1])
                    hits.append((F(np.sqrt(F(F(ddx * ddx) + F(ddy * ddy)))), at))
            hits.sort(key=lambda kv: kv[0])  # stable, like sorted_by_key
            for k in range(0, len(hits) - 1, 2):
                s, e = hits[k][1], hits[k + 1][1]
                if s == e:
                    continue
                ddx, ddy = F(e[0] - s[0]), F(e[1] - s[1])
                f1 = F(F(0.999) - F(1.0))
                clipped.append((F(s[0] - F(F(ddx * f1) / F(2.0))), F(s[1] - F(F(ddy * f1) / F(2.0))),
                                F(e[0] + F(F(ddx * f1) / F(2.0))), F(e[1] + F(F(ddy * f1) / F(2.0)))))
    clipped = np.array(clipped, F).reshape(-1, 4)
    grid = []
    for x in linspace(sq[0], sq[2], n_points_per_dimension):
        for y in linspace(sq[1], sq[3], n_points_per_dimension):
            if not S.point_in_polygon(pts, (x, y)):
                continue
            if any(F(F(F(x - px) * F(x - px)) + F(F(y - py) * F(y - py))) <= F(pr * pr) for px, py, pr in poles):
                continue
            grid.append((x, y))
    gx = np.array([g[0] for g in grid], F)
    gy = np.array([g[1] for g in grid], F)
    radius = F(F(0.10) * sq_w)
    bw, bh = w, h
    forfeit = F(np.sqrt(F(F(bw * bw) * F(bh * bh))))
    d_poles = np.full(len(grid), forfeit, F)
    for px, py, pr in poles:
        ddx, ddy = (gx - px).astype(F), (gy - py).astype(F)
        sqd = ((ddx * ddx).astype(F) + (ddy * ddy).astype(F)).astype(F)
        dist = np.where(sqd < F(pr * pr), F(0.0), (np.sqrt(sqd).astype(F) - pr).astype(F)).astype(F)
        d_poles = np.minimum(d_poles, dist)
    selected = []
    for _ in range(n):
        d_rays = np.full(len(grid), forfeit, F)
        for sgm in selected:
            d_rays = np.minimum(d_rays, edge_distances(((sgm[0], sgm[1]), (sgm[2], sgm[3])), gx, gy))
        assert len(clipped) > 0, "no ray found"
        losses = np.empty(len(clipped), F)
        for r0 in range(0, len(clipped), 1024):
            dn = edge_distances_many(clipped[r0:r0 + 1024], gx, gy)
            dr = np.minimum(d_rays[None, :], dn)
            d = np.where(dr < radius, np.minimum(d_poles[None, :], dr), d_poles[None, :]).astype(F)
            sq_d = (d * d).astype(F)
            # Iterator::sum: a sequential f32 fold (np.cumsum accumulates in order, in the array's dtype)
            losses[r0:r0 + 1024] = np.cumsum(sq_d, axis=1, dtype=F)[:, -1] if sq_d.shape[1] else F(0.0)
        assert not np.isnan(losses).any()
        selected.append(tuple(clipped[int(np.argmin(losses))]))  # argmin: the first minimum, like min_by_key
    return np.array(selected, F).reshape(-1, 4)
