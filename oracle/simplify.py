"""TEST INFRASTRUCTURE - the checker for jg_shape_simplify / jg_item_import_simplified, never imported by the product.

A restatement of the reference's polygon simplification in numpy float32 scalars (every operation rounds to f32, no
fused multiply-add), written from the reference source and independent of the product's C++:

    jagua-rs/src/geometry/shape_modification.rs:48-144   simplify_shape
    :146-174 calculate_area_delta   :176-208 candidate_is_valid   :218-234 execute_candidate
    :236-280 replacing_vertex_convex_convex_candidate, calculate_intersection_in_front   :293-329 Corner, CornerType
    jagua-rs/src/geometry/primitives/edge.rs:180-228     edge_intersection (Edge: CollidesWith<Edge>)
    jagua-rs/src/geometry/primitives/simple_polygon.rs:135-149 calculate_area

Parity status: unpinned against the Rust crate (no toolchain here); it pins the product against an independent second
reading of the same source.
"""
import numpy as np

F = np.float32
ZERO, ONE, TWO = F(0.0), F(1.0), F(2.0)


def area(pts):
    """simple_polygon.rs:135-149 (shoelace, counter-clockwise positive)"""
    sigma = ZERO
    n = len(pts)
    for i in range(n):
        j = (i + 1) % n
        (xi, yi), (xj, yj) = pts[i], pts[j]
        sigma = F(sigma + F(F(yi + yj) * F(xi - xj)))
    return F(F(0.5) * sigma)


def edges_collide(e1, e2):
    """edge.rs:180-228 with calc_loc = false"""
    (x1, y1), (x2, y2) = e1
    (x3, y3), (x4, y4) = e2
    if max(min(x1, x2), min(x3, x4)) > min(max(x1, x2), max(x3, x4)):
        return False
    if max(min(y1, y2), min(y3, y4)) > min(max(y1, y2), max(y3, y4)):
        return False
    t_nom = F(F(F(x2 - x4) * F(y4 - y3)) - F(F(y2 - y4) * F(x4 - x3)))
    t_den = F(F(F(x2 - x1) * F(y4 - y3)) - F(F(y2 - y1) * F(x4 - x3)))
    u_nom = F(F(F(x2 - x4) * F(y2 - y1)) - F(F(y2 - y4) * F(x2 - x1)))
    u_den = t_den
    if t_den == ZERO or u_den == ZERO:
        return False
    with np.errstate(all="ignore"):
        t, u = F(t_nom / t_den), F(u_nom / u_den)
    return bool(ZERO <= t <= ONE and ZERO <= u <= ONE)


def intersection_in_front(l1, l2):
    """shape_modification.rs:249-280"""
    (x1, y1), (x2, y2) = l1
    (x3, y3), (x4, y4) = l2
    t_nom = F(F(F(x2 - x4) * F(y4 - y3)) - F(F(y2 - y4) * F(x4 - x3)))
    t_den = F(F(F(x2 - x1) * F(y4 - y3)) - F(F(y2 - y1) * F(x4 - x3)))
    u_nom = F(F(F(x2 - x4) * F(y2 - y1)) - F(F(y2 - y4) * F(x2 - x1)))
    u_den = F(F(F(x2 - x1) * F(y4 - y3)) - F(F(y2 - y1) * F(x4 - x3)))
    with np.errstate(all="ignore"):
        t = F(t_nom / t_den) if t_den != ZERO else ZERO
        u = F(u_nom / u_den) if u_den != ZERO else ZERO
    if t < ZERO and u < ZERO:
        return (F(x2 + F(t * F(x1 - x2))), F(y2 + F(t * F(y1 - y2))))
    return None


def corner_type(p1, p2, p3):
    """shape_modification.rs:312-329: left-hand side of p1 -> p2 -> p3"""
    ax, ay = F(p2[0] - p1[0]), F(p2[1] - p1[1])
    bx, by = F(p3[0] - p1[0]), F(p3[1] - p1[1])
    cross = F(F(ax * by) - F(ay * bx))
    return "concave" if cross < ZERO else ("collinear" if cross == ZERO else "convex")


def replacing_vertex(shape, c1, c2):
    """shape_modification.rs:236-247"""
    assert c1[2] == c2[1] and c1[1] == c2[0]
    return intersection_in_front((shape[c1[0]], shape[c1[1]]), (shape[c2[2]], shape[c2[1]]))


def tri(p0, p1, p2):
    (x0, y0), (x1, y1), (x2, y2) = p0, p1, p2
    a = F(x0 * y1)
    a = F(a + F(x1 * y2))
    a = F(a + F(x2 * y0))
    a = F(a - F(x0 * y2))
    a = F(a - F(x1 * y0))
    a = F(a - F(x2 * y1))
    return abs(F(a / TWO))


def area_delta(shape, cand):
    """shape_modification.rs:146-174 (None = InvalidCandidate)"""
    kind = cand[0]
    if kind == "collinear":
        return ZERO
    if kind == "concave":
        c = cand[1]
        return tri(shape[c[0]], shape[c[1]], shape[c[2]])
    c1, c2 = cand[1], cand[2]
    r = replacing_vertex(shape, c1, c2)
    if r is None:
        return None
    return tri(shape[c1[1]], r, shape[c2[1]])


def edge_list(points):
    n = len(points)
    return [(points[i], points[(i + 1) % n]) for i in range(n)]


def is_valid(shape, cand):
    """shape_modification.rs:176-208 (raises where the reference unwraps a degenerate Edge)"""
    kind = cand[0]
    if kind == "collinear":
        return True
    if kind == "concave":
        c = cand[1]
        if shape[c[0]] == shape[c[2]]:
            raise ValueError("degenerate edge")
        new_edge = (shape[c[0]], shape[c[2]])
        affected = [shape[c[0]], shape[c[1]], shape[c[2]]]
        return all(not edges_collide(l, new_edge) for l in edge_list(shape) if l[0] not in affected and l[1] not in affected)
    c1, c2 = cand[1], cand[2]
    r = replacing_vertex(shape, c1, c2)
    if r is None:
        return False
    if shape[c1[0]] == r or r == shape[c2[2]]:
        raise ValueError("degenerate edge")
    e1, e2 = (shape[c1[0]], r), (r, shape[c2[2]])
    affected = [shape[c1[1]], shape[c1[0]], shape[c2[1]], shape[c2[2]]]
    return all(not edges_collide(l, e1) and not edges_collide(l, e2)
               for l in edge_list(shape) if l[0] not in affected and l[1] not in affected)


def execute(shape, cand):
    """shape_modification.rs:218-234"""
    pts = list(shape)
    if cand[0] in ("collinear", "concave"):
        del pts[cand[1][1]]
        return pts
    c1, c2 = cand[1], cand[2]
    r = replacing_vertex(shape, c1, c2)
    del pts[c1[1]]
    other = c2[1] - 1 if c1[1] < c2[1] else c2[1]
    del pts[other]
    pts.insert(other, r)
    return pts


def simplify(points, mode="inflate", max_area_change_ratio=0.001):
    """shape_modification.rs:48-144 up to (not including) the final SPolygon::new. points: counter-clockwise (x, y)."""
    ref = [(F(x), F(y)) for x, y in np.asarray(points, np.float32)]
    original_area = area(ref)
    max_ratio = F(max_area_change_ratio)
    for _ in range(len(ref)):
        n = len(ref)
        if n < 4:
            break
        corners = [((i - 1) % n, i, (i + 1) % n) for i in range(n)]
        if mode == "deflate":
            corners = [(c[2], c[1], c[0]) for c in reversed(corners)]
        cands = []
        prev = corners[-1]
        prev_type = corner_type(*[ref[k] for k in prev])
        for c in corners:
            ct = corner_type(*[ref[k] for k in c])
            if ct == "concave":
                cands.append(("concave", c))
            elif ct == "collinear":
                cands.append(("collinear", c))
            elif prev_type == "convex":
                cands.append(("convex_convex", prev, c))
            prev, prev_type = c, ct
        keyed = []
        for c in cands:
            d = area_delta(ref, c)
            keyed.append(float("inf") if d is None else float(d))
        order = sorted(range(len(cands)), key=lambda k: (np.isnan(keyed[k]), keyed[k]))  # stable; OrderedFloat puts NaN last
        best = next((cands[k] for k in order if is_valid(ref, cands[k])), None)
        if best is None:
            break
        new = execute(ref, best)
        with np.errstate(all="ignore"):
            delta = F(abs(F(area(new) - original_area)) / original_area)
        if delta <= max_ratio:
            ref = new
        else:
            break
    return np.array(ref, np.float32).reshape(-1, 2)


# ---- close_narrow_concavities (shape_modification.rs:374-502) and what it needs ------------------------------------------
def approx_eq(a, b):
    """util/fpa.rs over float-cmp 0.10 approx_eq!(f32): epsilon = f32::EPSILON, ulps = 4"""
    a, b = F(a), F(b)
    if a == b:
        return True
    with np.errstate(all="ignore"):
        if abs(F(a - b)) <= F(1.1920929e-07):
            return True
    ia, ib = int(np.array(a).view(np.int32)), int(np.array(b).view(np.int32))
    d = (ia - ib + 2 ** 31) % 2 ** 32 - 2 ** 31  # wrapping_sub on the i32 views
    ad = (2 ** 31 - 1) if d == -2 ** 31 else abs(d)
    return ad <= 4


def point_in_polygon(pts, p):
    """simple_polygon.rs:255-292 (ray casting with the FPA vertex rule)"""
    xs, ys = [q[0] for q in pts], [q[1] for q in pts]
    x_min, x_max, y_min, y_max = min(xs), max(xs), min(ys), max(ys)
    px, py = p
    if not (x_min <= px <= x_max and y_min <= py <= y_max):
        return False
    ray = ((px, py), (F(x_max + F(x_max - x_min)), py))
    n = 0
    for s, e in edge_list(pts):
        def gt(a, b):
            return (not approx_eq(a, b)) and a > b

        def lt(a, b):
            return (not approx_eq(a, b)) and a < b
        if (approx_eq(s[1], py) and gt(s[0], px)) or (approx_eq(e[1], py) and gt(e[0], px)):
            if lt(s[1], py) or lt(e[1], py):
                n += 1
        elif edges_collide(ray, (s, e)):
            n += 1
    return n % 2 == 1


def diameter(pts):
    """simple_polygon.rs:105-120 over convex_hull.rs:30-63 (monotone chain on the points sorted by x, stable)"""
    def cross(a, b, c):
        return F(F(F(b[0] - a[0]) * F(c[1] - a[1])) - F(F(b[1] - a[1]) * F(c[0] - a[0])))

    def grow(h, nxt):
        while len(h) >= 2 and cross(h[-2], h[-1], nxt) <= ZERO:
            h.pop()
        h.append(nxt)
        return h
    srt = sorted(pts, key=lambda q: q[0])
    lower, upper = [], []
    for q in srt:
        grow(lower, q)
    for q in reversed(srt):
        grow(upper, q)
    hull = lower[:-1] + upper[:-1]
    best = F(-1.0)
    for i in range(len(hull)):
        for j in range(i + 1, len(hull)):
            dx, dy = F(hull[i][0] - hull[j][0]), F(hull[i][1] - hull[j][1])
            best = max(best, F(F(dx * dx) + F(dy * dy)))
    return F(np.sqrt(best))


def close_narrow_concavities(points, mode="inflate", max_distance_ratio=0.01):
    """shape_modification.rs:374-502 (each round ends in SPolygon::new: orientation normalised)"""
    v = [(F(x), F(y)) for x, y in np.asarray(points, np.float32)]
    ratio = F(max_distance_ratio)
    factor = F(0.9999)
    for _ in range(len(v)):
        n = len(v)
        diam = diameter(v)

        def elim(i, j):
            return j - i - 1 if j > i else n - i + j - 1
        best = None
        for i in range(n):
            for j in range(n):
                if i == j or (i + 1) % n == j or (j + 1) % n == i:
                    continue
                (sx, sy), (ex, ey) = v[i], v[j]
                dx, dy = F(ex - sx), F(ey - sy)
                fm1 = F(factor - ONE)
                sx = F(sx - F(F(dx * fm1) / TWO))
                sy = F(sy - F(F(dy * fm1) / TWO))
                ex = F(ex + F(F(dx * fm1) / TWO))
                ey = F(ey + F(F(dy * fm1) / TWO))
                lx, ly = F(sx - ex), F(sy - ey)
                if F(np.sqrt(F(F(lx * lx) + F(ly * ly)))) > F(ratio * diam):
                    continue
                s_in, e_in = point_in_polygon(v, (sx, sy)), point_in_polygon(v, (ex, ey))
                if mode == "inflate" and (s_in or e_in):
                    continue
                if mode == "deflate" and not (s_in and e_in):
                    continue
                c_edge = ((sx, sy), (ex, ey))
                if any(edges_collide(e, c_edge) for e in edge_list(v)):
                    continue
                sub = v[i:j] if j > i else v[i:] + v[:j]
                if area(sub) >= ZERO:
                    continue
                if best is None or elim(i, j) > elim(*best):
                    best = (i, j)
        if best is None:
            break
        i, j = best
        start, end = i + 1, j - 1
        ref = list(v)
        if start <= end:
            del ref[start:end + 1]
        else:
            if start < n:
                del ref[start:]
            if end >= 0:
                del ref[0:end + 1]
        if area(ref) < ZERO:
            ref.reverse()
        v = ref
    return np.array(v, np.float32).reshape(-1, 2)
