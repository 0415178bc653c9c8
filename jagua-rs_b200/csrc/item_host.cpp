// Host-side preparation of an item for the batched CDE: what the reference does once per item type before any
// query runs, and whose results cross the C ABI with jg_item_upload. Host C++ because the reference's is compiled
// host code too; f32 throughout, compiled -ffp-contract=off like every other predicate.
//
// Mirrors (paths under jagua-rs/src):
//   io/import.rs:198-260                       import_simple_polygon, eliminate_degenerate_vertices, centering_transformation
//   geometry/primitives/simple_polygon.rs:40-71,122-192   SPolygon::new, generate_bounding_box, calculate_area, calculate_poi, centroid
//   geometry/original_shape.rs:23-51           convert_to_internal with shape modification disabled (offset / simplification
//                                              are the "next" row f3 of SURVEY section 8)
//   geometry/fail_fast/pole.rs:12-135          generate_surrogate_poles, compute_pole, POINode
//   geometry/fail_fast/sp_surrogate.rs:33-61   SPSurrogate::new (poles; piers: jg_item_generate_piers, on request)
//   geometry/fail_fast/piers.rs:25-226         generate_piers
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <deque>
#include <utility>
#include <vector>

#include "../../include/jagua_cuda.h"
#include "predicates.cuh"

namespace jg {
namespace {

struct Pt {
    float x, y;
};

// simple_polygon.rs:135-149: shoelace, counter-clockwise positive
float signed_area(const std::vector<Pt>& p) {
    float sigma = 0.0f;
    const size_t n = p.size();
    for (size_t i = 0; i < n; i++) {
        const size_t j = (i + 1) % n;
        sigma += (p[i].y + p[j].y) * (p[i].x - p[j].x);
    }
    return 0.5f * sigma;
}

Rect bbox_of(const std::vector<Pt>& p) {  // simple_polygon.rs:122-133
    float x0 = 3.402823466e+38f, y0 = 3.402823466e+38f, x1 = -3.402823466e+38f, y1 = -3.402823466e+38f;
    for (const Pt& q : p) {
        x0 = fminf(x0, q.x), y0 = fminf(y0, q.y);
        x1 = fmaxf(x1, q.x), y1 = fmaxf(y1, q.y);
    }
    return Rect{x0, y0, x1, y1};
}

// edge.rs:63-87 closest_point_on_edge + :153-159 distance_to
float edge_distance(const Pt& s, const Pt& e, float x, float y) {
    const float a = x - s.x, b = y - s.y, c = e.x - s.x, d = e.y - s.y;
    const float dot = a * c + b * d;
    const float len_sq = c * c + d * d;
    float param = -1.0f;
    if (len_sq != 0.0f) param = dot / len_sq;
    float xx, yy;
    if (param < 0.0f) xx = s.x, yy = s.y;
    else if (param > 1.0f) xx = e.x, yy = e.y;
    else xx = s.x + param * c, yy = s.y + param * d;
    const float dx = x - xx, dy = y - yy;
    return sqrtf(dx * dx + dy * dy);
}

struct Shape {
    std::vector<Pt> v;
    Rect bbox;
    float area;
};
struct Pole {
    float x, y, r;
};

struct PoiNode {  // pole.rs:81-135
    int level;
    Rect bbox;
    float radius, distance;
};

PoiNode make_node(const Rect& bb, int level, const Shape& sh, const std::vector<Pole>& poles) {
    PoiNode n;
    n.level = level;
    n.bbox = bb;
    const float dx = bb.x_max - bb.x_min, dy = bb.y_max - bb.y_min;
    n.radius = sqrtf(dx * dx + dy * dy) / 2.0f;  // rect.rs:259-263 diameter() / 2
    const float cx = (bb.x_min + bb.x_max) / 2.0f, cy = (bb.y_min + bb.y_max) / 2.0f;  // rect.rs:248-253
    bool inside = poly_point<false>(&sh.v[0].x, (int)sh.v.size(), sh.bbox, cx, cy, Xf{1.0f, 0.0f, 0.0f, 0.0f});
    if (inside)
        for (const Pole& p : poles)  // circle.rs:130-134
            if ((cx - p.x) * (cx - p.x) + (cy - p.y) * (cy - p.y) <= p.r * p.r) { inside = false; break; }
    float border = 3.402823466e+38f;
    const size_t nv = sh.v.size();
    for (size_t i = 0; i < nv; i++) border = fminf(border, edge_distance(sh.v[i], sh.v[i + 1 == nv ? 0 : i + 1], cx, cy));
    for (const Pole& p : poles) {  // circle.rs:154-163 separation_distance(point).1
        const float dc = sqrtf((cx - p.x) * (cx - p.x) + (cy - p.y) * (cy - p.y));
        border = fminf(border, dc <= p.r ? p.r - dc : dc - p.r);
    }
    n.distance = inside ? border : -border;
    return n;
}

// pole.rs:49-76: breadth-first refinement of a depth-10 quadtree over the squared bbox
bool compute_pole(const Shape& sh, const std::vector<Pole>& poles, Pole& out) {
    Rect sq = sh.bbox;  // rect.rs:93-109 inflate_to_square
    {
        const float w = sq.x_max - sq.x_min, h = sq.y_max - sq.y_min;
        float dx = 0.0f, dy = 0.0f;
        if (h < w) dy = (w - h) / 2.0f;
        else if (w < h) dx = (h - w) / 2.0f;
        sq = Rect{sq.x_min - dx, sq.y_min - dy, sq.x_max + dx, sq.y_max + dy};
    }
    std::deque<PoiNode> queue;
    queue.push_back(make_node(sq, 10, sh, poles));
    bool have = false;
    float best_r = 0.0f;
    while (!queue.empty()) {
        const PoiNode n = queue.front();
        queue.pop_front();
        if (n.distance > best_r) {
            have = true;
            best_r = n.distance;
            out = Pole{(n.bbox.x_min + n.bbox.x_max) / 2.0f, (n.bbox.y_min + n.bbox.y_max) / 2.0f, n.distance};
        }
        if (n.radius + n.distance > best_r && n.level > 0)
            for (int q = 0; q < 4; q++) queue.push_back(make_node(rect_quadrant(n.bbox, q), n.level - 1, sh, poles));
    }
    return have;
}

}  // namespace
}  // namespace jg

using namespace jg;

// ---- geometry/shape_modification.rs:48-330 simplify_shape --------------------------------------------------------
// Removes one vertex (concave / collinear corner) or replaces two consecutive convex corners by the intersection of
// their outer edges, always the valid candidate of least area change, until the accumulated area change would exceed
// max_ratio of the original area or fewer than 4 vertices remain. Inflate: the result is a superset of the input;
// Deflate: a subset (corners are walked in reverse and flipped).
enum CornerKind { CK_CONCAVE, CK_CONVEX, CK_COLLINEAR };
struct Corner3 {
    int a, b, c;  // left-hand side of a -> b -> c
};
struct Cand {
    int kind;     // 0 concave, 1 convex-convex, 2 collinear
    Corner3 c1, c2;
};

static CornerKind corner_kind(const std::vector<Pt>& p, const Corner3& c) {  // shape_modification.rs:312-329
    const Pt p1 = p[c.a], p2 = p[c.b], p3 = p[c.c];
    const float ux = p2.x - p1.x, uy = p2.y - p1.y, vx = p3.x - p1.x, vy = p3.y - p1.y;
    const float cross = ux * vy - uy * vx;
    return cross < 0.0f ? CK_CONCAVE : (cross == 0.0f ? CK_COLLINEAR : CK_CONVEX);
}

// shape_modification.rs:249-280 calculate_intersection_in_front
static bool intersection_in_front(const Pt& a1, const Pt& a2, const Pt& b1, const Pt& b2, Pt& out) {
    const float x1 = a1.x, y1 = a1.y, x2 = a2.x, y2 = a2.y, x3 = b1.x, y3 = b1.y, x4 = b2.x, y4 = b2.y;
    const float t_nom = (x2 - x4) * (y4 - y3) - (y2 - y4) * (x4 - x3);
    const float t_denom = (x2 - x1) * (y4 - y3) - (y2 - y1) * (x4 - x3);
    const float u_nom = (x2 - x4) * (y2 - y1) - (y2 - y4) * (x2 - x1);
    const float u_denom = (x2 - x1) * (y4 - y3) - (y2 - y1) * (x4 - x3);
    const float t = t_denom != 0.0f ? t_nom / t_denom : 0.0f;
    const float u = u_denom != 0.0f ? u_nom / u_denom : 0.0f;
    if (t < 0.0f && u < 0.0f) {
        out = Pt{x2 + t * (x1 - x2), y2 + t * (y1 - y2)};
        return true;
    }
    return false;
}
// shape_modification.rs:236-247: edge_prev = c1.a -> c1.b, edge_next = c2.c -> c2.b
static bool replacing_vertex(const std::vector<Pt>& p, const Corner3& c1, const Corner3& c2, Pt& out) {
    return intersection_in_front(p[c1.a], p[c1.b], p[c2.c], p[c2.b], out);
}
static float tri_area_abs(const Pt& p0, const Pt& p1, const Pt& p2) {  // shape_modification.rs:155,168
    const float a = (p0.x * p1.y + p1.x * p2.y + p2.x * p0.y - p0.x * p2.y - p1.x * p0.y - p2.x * p1.y) / 2.0f;
    return fabsf(a);
}
static float cand_area_delta(const std::vector<Pt>& p, const Cand& c) {  // :146-174; infinity for an invalid candidate
    if (c.kind == 2) return 0.0f;
    if (c.kind == 0) return tri_area_abs(p[c.c1.a], p[c.c1.b], p[c.c1.c]);
    Pt r;
    if (!replacing_vertex(p, c.c1, c.c2, r)) return INFINITY;
    return tri_area_abs(p[c.c1.b], r, p[c.c2.b]);
}
static bool same_pt(const Pt& a, const Pt& b) { return a.x == b.x && a.y == b.y; }
static bool edge_hits(const Pt& s, const Pt& e, const Pt& ns, const Pt& ne) {  // Edge::collides_with(Edge), edge.rs:162-170
    return seg_seg(Seg{s.x, s.y, e.x, e.y}, Seg{ns.x, ns.y, ne.x, ne.y});
}
// shape_modification.rs:176-208; *panic = the reference would unwrap a degenerate Edge
static bool cand_valid(const std::vector<Pt>& p, const Cand& c, bool* panic) {
    const size_t n = p.size();
    if (c.kind == 2) return true;
    if (c.kind == 0) {
        const Pt ns = p[c.c1.a], ne = p[c.c1.c];
        if (same_pt(ns, ne)) { *panic = true; return false; }
        const Pt aff[3] = {p[c.c1.a], p[c.c1.b], p[c.c1.c]};
        for (size_t i = 0; i < n; i++) {
            const Pt s = p[i], e = p[(i + 1) % n];
            bool skip = false;
            for (const Pt& a : aff) skip = skip || same_pt(a, s) || same_pt(a, e);
            if (!skip && edge_hits(s, e, ns, ne)) return false;
        }
        return true;
    }
    Pt r;
    if (!replacing_vertex(p, c.c1, c.c2, r)) return false;
    if (same_pt(p[c.c1.a], r) || same_pt(r, p[c.c2.c])) { *panic = true; return false; }
    const Pt aff[4] = {p[c.c1.b], p[c.c1.a], p[c.c2.b], p[c.c2.c]};
    for (size_t i = 0; i < n; i++) {
        const Pt s = p[i], e = p[(i + 1) % n];
        bool skip = false;
        for (const Pt& a : aff) skip = skip || same_pt(a, s) || same_pt(a, e);
        if (!skip && (edge_hits(s, e, p[c.c1.a], r) || edge_hits(s, e, r, p[c.c2.c]))) return false;
    }
    return true;
}
static std::vector<Pt> cand_execute(const std::vector<Pt>& p, const Cand& c) {  // :218-234
    std::vector<Pt> q = p;
    if (c.kind != 1) {
        q.erase(q.begin() + c.c1.b);
        return q;
    }
    Pt r;
    replacing_vertex(p, c.c1, c.c2, r);
    q.erase(q.begin() + c.c1.b);
    const int other = c.c1.b < c.c2.b ? c.c2.b - 1 : c.c2.b;
    q.erase(q.begin() + other);
    q.insert(q.begin() + other, r);
    return q;
}
// returns false when the reference would panic (degenerate replacement edge)
static bool simplify_points(std::vector<Pt>& pts, float original_area, bool deflate, float max_ratio) {
    const size_t rounds = pts.size();
    for (size_t round = 0; round < rounds; round++) {
        const int n = (int)pts.size();
        if (n < 4) break;
        std::vector<Corner3> corners((size_t)n);
        for (int i = 0; i < n; i++) corners[(size_t)i] = Corner3{(i - 1 + n) % n, i, (i + 1) % n};
        if (deflate) {
            std::reverse(corners.begin(), corners.end());
            for (Corner3& c : corners) std::swap(c.a, c.c);
        }
        std::vector<Cand> cands;
        Corner3 prev = corners.back();
        CornerKind prev_kind = corner_kind(pts, prev);
        for (const Corner3& c : corners) {
            const CornerKind k = corner_kind(pts, c);
            if (k == CK_CONCAVE) cands.push_back(Cand{0, c, c});
            else if (k == CK_COLLINEAR) cands.push_back(Cand{2, c, c});
            else if (prev_kind == CK_CONVEX) cands.push_back(Cand{1, prev, c});
            prev = c;
            prev_kind = k;
        }
        // stable sort by area change (OrderedFloat: NaN last), first valid candidate wins
        std::vector<std::pair<float, size_t>> order;
        for (size_t i = 0; i < cands.size(); i++) order.push_back({cand_area_delta(pts, cands[i]), i});
        std::stable_sort(order.begin(), order.end(), [](const std::pair<float, size_t>& a, const std::pair<float, size_t>& b) {
            const bool an = a.first != a.first, bn = b.first != b.first;
            if (an || bn) return !an && bn;
            return a.first < b.first;
        });
        const Cand* best = nullptr;
        for (const auto& o : order) {
            bool panic = false;
            if (cand_valid(pts, cands[o.second], &panic)) { best = &cands[o.second]; break; }
            if (panic) return false;
        }
        if (!best) break;
        std::vector<Pt> next = cand_execute(pts, *best);
        const float delta = fabsf(signed_area(next) - original_area) / original_area;
        if (!(delta <= max_ratio)) break;  // area change too significant
        pts.swap(next);
    }
    return true;
}

// ---- geometry/convex_hull.rs:30-63 + simple_polygon.rs:105-120 calculate_diameter ---------------------------------
static float hull_cross(const Pt& a, const Pt& b, const Pt& c) { return (b.x - a.x) * (c.y - a.y) - (b.y - a.y) * (c.x - a.x); }
static void hull_grow(std::vector<Pt>& h, const Pt& next) {
    while (h.size() >= 2 && hull_cross(h[h.size() - 2], h[h.size() - 1], next) <= 0.0f) h.pop_back();
    h.push_back(next);
}
static float polygon_diameter(std::vector<Pt> pts) {
    std::stable_sort(pts.begin(), pts.end(), [](const Pt& a, const Pt& b) { return a.x < b.x; });  // sort_by_key(OrderedFloat(x))
    std::vector<Pt> lower, upper;
    for (size_t i = 0; i < pts.size(); i++) hull_grow(lower, pts[i]);
    for (size_t i = pts.size(); i-- > 0;) hull_grow(upper, pts[i]);
    upper.pop_back();
    lower.pop_back();
    lower.insert(lower.end(), upper.begin(), upper.end());
    float best = -1.0f;
    for (size_t i = 0; i < lower.size(); i++)
        for (size_t j = i + 1; j < lower.size(); j++) {
            const float dx = lower[i].x - lower[j].x, dy = lower[i].y - lower[j].y;
            best = fmaxf(best, dx * dx + dy * dy);
        }
    return sqrtf(best);
}

// ---- geometry/shape_modification.rs:374-502 close_narrow_concavities -----------------------------------------------
// Replaces a concavity whose mouth (v[i] -> v[j], shrunk by 0.9999) is shorter than max_ratio x the diameter by that
// straight edge, always the valid candidate that eliminates most vertices, until none is left.
static bool close_concavities(std::vector<Pt>& v, bool deflate, float max_ratio) {
    const size_t rounds = v.size();
    for (size_t round = 0; round < rounds; round++) {
        const int n = (int)v.size();
        Shape sh;
        sh.v = v;
        sh.bbox = bbox_of(v);
        sh.area = signed_area(v);
        const float diameter = polygon_diameter(v);
        auto elim = [&](int i, int j) { return j > i ? j - i - 1 : n - i + j - 1; };
        auto inside = [&](float x, float y) { return poly_point<false>(&sh.v[0].x, n, sh.bbox, x, y, Xf{1.0f, 0.0f, 0.0f, 0.0f}); };
        int bi = -1, bj = -1;
        for (int i = 0; i < n; i++)
            for (int j = 0; j < n; j++) {
                if (i == j || (i + 1) % n == j || (j + 1) % n == i) continue;  // adjacent points
                if (same_pt(v[i], v[j])) return false;                          // Edge::try_new would fail
                // Edge::scale(0.9999) (edge.rs:39-46)
                const float factor = 0.9999f;
                const float dx = v[j].x - v[i].x, dy = v[j].y - v[i].y;
                float sx = v[i].x, sy = v[i].y, ex = v[j].x, ey = v[j].y;
                sx -= dx * (factor - 1.0f) / 2.0f;
                sy -= dy * (factor - 1.0f) / 2.0f;
                ex += dx * (factor - 1.0f) / 2.0f;
                ey += dy * (factor - 1.0f) / 2.0f;
                const float lx = sx - ex, ly = sy - ey;
                if (sqrtf(lx * lx + ly * ly) > max_ratio * diameter) continue;  // mouth too wide
                const bool s_in = inside(sx, sy), e_in = inside(ex, ey);
                if (!deflate && (s_in || e_in)) continue;
                if (deflate && !(s_in && e_in)) continue;
                bool hits = false;
                for (int k = 0; k < n && !hits; k++) hits = edge_hits(v[k], v[(k + 1) % n], Pt{sx, sy}, Pt{ex, ey});
                if (hits) continue;
                std::vector<Pt> sub;
                if (j > i) sub.assign(v.begin() + i, v.begin() + j);
                else {
                    sub.assign(v.begin() + i, v.end());
                    sub.insert(sub.end(), v.begin(), v.begin() + j);
                }
                if (signed_area(sub) >= 0.0f) continue;
                if (bi < 0 || elim(i, j) > elim(bi, bj)) bi = i, bj = j;
            }
        if (bi < 0) break;
        std::vector<Pt> next;
        if (bi + 1 <= bj - 1) {  // j does not wrap around
            for (int k = 0; k < n; k++)
                if (k < bi + 1 || k > bj - 1) next.push_back(v[k]);
        } else {
            for (int k = 0; k < n; k++) {
                const bool drop = (bi + 1 < n && k >= bi + 1) || (bj - 1 >= 0 && k <= bj - 1);
                if (!drop) next.push_back(v[k]);
            }
        }
        // SPolygon::new (simple_polygon.rs:40-71)
        if (next.size() < 3) return false;
        const float a = signed_area(next);
        if (a == 0.0f) return false;
        if (a < 0.0f) std::reverse(next.begin(), next.end());
        v.swap(next);
    }
    return true;
}

// SPolygon::new's pole of inaccessibility (simple_polygon.rs:40-71,151-171) for a shape that is used as it is
// (container outer boundary, holes, zones): same cleaning-free path as the reference - the caller passes the
// final vertex list; orientation is normalised to counter-clockwise first, as SPolygon::new does.
extern "C" int jg_polygon_poi(const float* pts_xy, int n_pts, float poi_xyr[3], float* verts_ccw_xy) {
    if (!pts_xy || n_pts < 3 || !poi_xyr) return JG_ERR_INVALID;
    Shape sh;
    sh.v.resize((size_t)n_pts);
    for (int i = 0; i < n_pts; i++) sh.v[i] = Pt{pts_xy[2 * i], pts_xy[2 * i + 1]};
    sh.area = signed_area(sh.v);
    if (sh.area == 0.0f) return JG_ERR_INVALID;
    if (sh.area < 0.0f) {
        std::vector<Pt> r(sh.v.rbegin(), sh.v.rend());
        sh.v.swap(r);
        sh.area = -sh.area;
    }
    sh.bbox = bbox_of(sh.v);
    Pole poi;
    if (!compute_pole(sh, {}, poi)) return JG_ERR_REFERENCE_PANIC;
    poi_xyr[0] = poi.x, poi_xyr[1] = poi.y, poi_xyr[2] = poi.r;
    if (verts_ccw_xy)
        for (int i = 0; i < n_pts; i++) verts_ccw_xy[2 * i] = sh.v[i].x, verts_ccw_xy[2 * i + 1] = sh.v[i].y;
    return JG_OK;
}

// io/import.rs:198-211,236-260 import_simple_polygon: strips a closing duplicate and approx-equal consecutive vertices;
// false when fewer than 3 vertices remain or a non-consecutive duplicate exists (the reference bails)
static bool clean_points(std::vector<Pt>& p) {
    if (p.size() > 1 && p.front().x == p.back().x && p.front().y == p.back().y) p.pop_back();
    std::vector<size_t> drop;
    const size_t n = p.size();
    for (size_t i = 0; i < n; i++) {
        const size_t j = (i + 1) % n;
        if (approx_eq(p[i].x, p[j].x) && approx_eq(p[i].y, p[j].y)) drop.push_back(i);
    }
    for (size_t k = drop.size(); k-- > 0;)  // io/import.rs:250-259: descending index order
        if (drop[k] < p.size()) p.erase(p.begin() + (long)drop[k]);
    if (p.size() < 3) return false;
    for (size_t i = 0; i < p.size(); i++)  // simple_polygon.rs:45-47: no duplicate points
        for (size_t j = i + 1; j < p.size(); j++)
            if (p[i].x == p[j].x && p[i].y == p[j].y) return false;
    return true;
}

// import_simple_polygon + SPolygon::new for a container boundary, hole or quality zone (io/import.rs:100-195: their
// pre-transform is empty and shape modification is disabled): cleaned, counter-clockwise vertices + poi.
extern "C" int jg_shape_import(const float* pts_xy, int n_pts, float* verts_xy, int* n_verts, float poi_xyr[3]) {
    if (!pts_xy || n_pts < 3 || !verts_xy || !n_verts || !poi_xyr) return JG_ERR_INVALID;
    std::vector<Pt> p((size_t)n_pts);
    for (int i = 0; i < n_pts; i++) p[i] = Pt{pts_xy[2 * i], pts_xy[2 * i + 1]};
    if (!clean_points(p)) return JG_ERR_INVALID;
    std::vector<float> flat(2 * p.size());
    for (size_t i = 0; i < p.size(); i++) flat[2 * i] = p[i].x, flat[2 * i + 1] = p[i].y;
    *n_verts = (int)p.size();
    return jg_polygon_poi(flat.data(), (int)p.size(), poi_xyr, verts_xy);
}

// simplify_shape (geometry/shape_modification.rs:48-144) on a clean counter-clockwise vertex list: mode 0 = Inflate (items,
// holes, zones), 1 = Deflate (container boundaries). The result goes through SPolygon::new again (orientation, poi).
extern "C" int jg_shape_simplify(const float* pts_xy, int n_pts, int mode, float max_area_change_ratio, float* verts_xy, int* n_verts,
                                 float poi_xyr[3]) {
    if (!pts_xy || n_pts < 3 || !verts_xy || !n_verts || !poi_xyr || (mode != 0 && mode != 1)) return JG_ERR_INVALID;
    std::vector<Pt> p((size_t)n_pts);
    for (int i = 0; i < n_pts; i++) p[i] = Pt{pts_xy[2 * i], pts_xy[2 * i + 1]};
    const float area = signed_area(p);
    if (!(area > 0.0f)) return JG_ERR_INVALID;  // SPolygon vertices are counter-clockwise
    if (!simplify_points(p, area, mode == 1, max_area_change_ratio)) return JG_ERR_REFERENCE_PANIC;
    std::vector<float> flat(2 * p.size());
    for (size_t i = 0; i < p.size(); i++) flat[2 * i] = p[i].x, flat[2 * i + 1] = p[i].y;
    *n_verts = (int)p.size();
    return jg_polygon_poi(flat.data(), (int)p.size(), poi_xyr, verts_xy);
}

// close_narrow_concavities (geometry/shape_modification.rs:374-502) on a clean counter-clockwise vertex list, followed by
// SPolygon::new (poi). mode 0 Inflate, 1 Deflate.
extern "C" int jg_shape_close_concavities(const float* pts_xy, int n_pts, int mode, float max_distance_ratio, float* verts_xy,
                                          int* n_verts, float poi_xyr[3]) {
    if (!pts_xy || n_pts < 3 || !verts_xy || !n_verts || !poi_xyr || (mode != 0 && mode != 1)) return JG_ERR_INVALID;
    std::vector<Pt> p((size_t)n_pts);
    for (int i = 0; i < n_pts; i++) p[i] = Pt{pts_xy[2 * i], pts_xy[2 * i + 1]};
    if (!(signed_area(p) > 0.0f)) return JG_ERR_INVALID;
    if (!close_concavities(p, mode == 1, max_distance_ratio)) return JG_ERR_REFERENCE_PANIC;
    std::vector<float> flat(2 * p.size());
    for (size_t i = 0; i < p.size(); i++) flat[2 * i] = p[i].x, flat[2 * i + 1] = p[i].y;
    *n_verts = (int)p.size();
    return jg_polygon_poi(flat.data(), (int)p.size(), poi_xyr, verts_xy);
}

// offset_shape (geometry/shape_modification.rs:330-372) - the repo's OWN offsetter, not geo-buffer 0.2 (a third-party crate
// that is not in the reference tree): results are NOT bit-equal to the reference's, and every consumer records the offset
// polygons it used. The reference buffers with round joins; here every edge of the counter-clockwise polygon moves by d along
// its outward normal (mode 0 Inflate: outwards, 1 Deflate: inwards), consecutive offset edges meet at their intersection
// (miter), and corners whose miter would reach further than miter_limit * |d| from the vertex are cut by two points on the
// tangents of the round join - so an inflated shape contains the reference's rounded one. Double arithmetic like the
// reference's call (f32 -> f64 -> f32). Self-intersecting results are refused (JG_ERR_REFERENCE_PANIC): the separation is too
// large for the shape's concavities. The result goes through the importer's cleaning and SPolygon::new like the reference's.
extern "C" int jg_shape_offset(const float* pts_xy, int n_pts, int mode, float distance, float miter_limit, float* verts_xy,
                               int verts_cap, int* n_verts, float poi_xyr[3]) {
    if (!pts_xy || n_pts < 3 || !verts_xy || !n_verts || !poi_xyr || (mode != 0 && mode != 1) || !(distance >= 0.0f)) return JG_ERR_INVALID;
    if (!(miter_limit > 1.0f)) miter_limit = 2.0f;
    std::vector<Pt> p0((size_t)n_pts);
    for (int i = 0; i < n_pts; i++) p0[(size_t)i] = Pt{pts_xy[2 * i], pts_xy[2 * i + 1]};
    if (!(signed_area(p0) > 0.0f)) return JG_ERR_INVALID;
    const double d = mode == 1 ? -(double)distance : (double)distance;
    const int n = n_pts;
    std::vector<double> ox, oy;
    std::vector<int> first((size_t)n), last((size_t)n);  // the offset points of corner i
    for (int i = 0; i < n; i++) {
        first[(size_t)i] = (int)ox.size();
        if (i > 0) last[(size_t)(i - 1)] = (int)ox.size() - 1;
        const Pt a = p0[(size_t)((i + n - 1) % n)], b = p0[(size_t)i], c = p0[(size_t)((i + 1) % n)];
        double e0x = (double)b.x - a.x, e0y = (double)b.y - a.y, e1x = (double)c.x - b.x, e1y = (double)c.y - b.y;
        const double l0 = std::sqrt(e0x * e0x + e0y * e0y), l1 = std::sqrt(e1x * e1x + e1y * e1y);
        if (!(l0 > 0.0) || !(l1 > 0.0)) return JG_ERR_INVALID;
        e0x /= l0, e0y /= l0, e1x /= l1, e1y /= l1;
        const double n0x = e0y, n0y = -e0x, n1x = e1y, n1y = -e1x;  // outward normals of a counter-clockwise polygon
        const double dotn = n0x * n1x + n0y * n1y;
        const double denom = 1.0 + dotn;
        if (denom < 1e-9) {  // a spike: the two offset edges are anti-parallel
            ox.push_back(b.x + d * n0x), oy.push_back(b.y + d * n0y);
            ox.push_back(b.x + d * n1x), oy.push_back(b.y + d * n1y);
            continue;
        }
        const double mx = (n0x + n1x) / denom, my = (n0y + n1y) / denom;  // b + d * m: intersection of the offset lines
        if (std::sqrt(mx * mx + my * my) > (double)miter_limit) {
            const double t = std::tan(0.25 * std::acos(std::min(1.0, std::max(-1.0, dotn))));
            ox.push_back(b.x + d * (n0x + e0x * t)), oy.push_back(b.y + d * (n0y + e0y * t));
            ox.push_back(b.x + d * (n1x - e1x * t)), oy.push_back(b.y + d * (n1y - e1y * t));
        } else {
            ox.push_back(b.x + d * mx), oy.push_back(b.y + d * my);
        }
    }
    const int m = (int)ox.size();
    last[(size_t)(n - 1)] = m - 1;
    for (int i = 0; i < n; i++) {  // every original edge must survive with its direction (else d exceeds the local feature size)
        const int u = last[(size_t)i], w = first[(size_t)((i + 1) % n)];
        const Pt b = p0[(size_t)i], c = p0[(size_t)((i + 1) % n)];
        if (!((ox[(size_t)w] - ox[(size_t)u]) * ((double)c.x - b.x) + (oy[(size_t)w] - oy[(size_t)u]) * ((double)c.y - b.y) > 0.0))
            return JG_ERR_REFERENCE_PANIC;
    }
    auto ccw = [&](int a, int b, int c) { return (oy[c] - oy[a]) * (ox[b] - ox[a]) - (oy[b] - oy[a]) * (ox[c] - ox[a]); };
    for (int i = 0; i < m; i++)
        for (int j = i + 2; j < m; j++) {
            if (i == 0 && j == m - 1) continue;
            const int i1 = (i + 1) % m, j1 = (j + 1) % m;
            if (ccw(i, i1, j) * ccw(i, i1, j1) < 0.0 && ccw(j, j1, i) * ccw(j, j1, i1) < 0.0) return JG_ERR_REFERENCE_PANIC;
        }
    double a2 = 0.0;  // an offset that turns the polygon inside out (deflated past its width) has no proper crossing
    for (int i = 0; i < m; i++) a2 += ox[(size_t)i] * oy[(size_t)((i + 1) % m)] - ox[(size_t)((i + 1) % m)] * oy[(size_t)i];
    if (!(a2 > 0.0)) return JG_ERR_REFERENCE_PANIC;  // "Offset resulted in an empty polygon" (shape_modification.rs:351)
    std::vector<float> flat((size_t)(2 * m));
    for (int i = 0; i < m; i++) flat[(size_t)(2 * i)] = (float)ox[(size_t)i], flat[(size_t)(2 * i + 1)] = (float)oy[(size_t)i];
    if (m > verts_cap) return JG_ERR_INVALID;
    return jg_shape_import(flat.data(), m, verts_xy, n_verts, poi_xyr);  // import_simple_polygon + SPolygon::new
}

static int item_import_impl(const float* pts_xy, int n_pts, const JgSurrogateConfig* scfg, float* verts_xy, int* n_verts,
                            float poi_xyr[3], float* poles_xyr, int poles_cap, int* n_poles, float pre_transform_xy[2],
                            float* area, float simplify_tolerance, float concavity_ratio, float offset = 0.0f) {
    if (!pts_xy || n_pts < 3 || !scfg || !verts_xy || !n_verts || !poi_xyr || !n_poles || !pre_transform_xy) return JG_ERR_INVALID;
    std::vector<Pt> p((size_t)n_pts);
    for (int i = 0; i < n_pts; i++) p[i] = Pt{pts_xy[2 * i], pts_xy[2 * i + 1]};
    if (!clean_points(p)) return JG_ERR_INVALID;
    Shape sh;
    sh.v = p;
    sh.area = signed_area(sh.v);
    if (sh.area == 0.0f) return JG_ERR_INVALID;
    if (sh.area < 0.0f) {  // simple_polygon.rs:49-57: edges are kept counter-clockwise
        std::vector<Pt> r(sh.v.rbegin(), sh.v.rend());
        sh.v.swap(r);
        sh.area = -sh.area;
    }
    sh.bbox = bbox_of(sh.v);
    Pole poi;
    if (!compute_pole(sh, {}, poi)) return JG_ERR_REFERENCE_PANIC;  // simple_polygon.rs:151-171 calculate_poi
    // simple_polygon.rs:173-192 centroid, io/import.rs:214-217 centering_transformation
    float cx = 0.0f, cy = 0.0f;
    const size_t nv = sh.v.size();
    for (size_t i = 0; i < nv; i++) {
        const size_t j = i + 1 == nv ? 0 : i + 1;
        const float cr = sh.v[i].x * sh.v[j].y - sh.v[j].x * sh.v[i].y;
        cx += (sh.v[i].x + sh.v[j].x) * cr;
        cy += (sh.v[i].y + sh.v[j].y) * cr;
    }
    cx /= 6.0f * sh.area;
    cy /= 6.0f * sh.area;
    // original_shape.rs:25 transform_clone(pre_transform.compose()): rotation 0 -> (cos, sin) = (1, 0)
    const Xf pre{1.0f, 0.0f, -cx, -cy};
    for (Pt& q : sh.v) xf_point(pre, q.x, q.y, q.x, q.y);
    xf_point(pre, poi.x, poi.y, poi.x, poi.y);
    sh.bbox = bbox_of(sh.v);
    const float area_orig = sh.area;  // Item::area() stays the original shape's
    if (offset > 0.0f) {
        // original_shape.rs:27-32: offset_shape in Inflate mode (ShapeModifyConfig::offset, half the minimum separation),
        // with the repo's own offsetter (jg_shape_offset: not geo-buffer, not bit-equal to the reference)
        std::vector<float> in(2 * sh.v.size()), outv(4 * sh.v.size());
        for (size_t i = 0; i < sh.v.size(); i++) in[2 * i] = sh.v[i].x, in[2 * i + 1] = sh.v[i].y;
        int nv2 = 0;
        float poi2[3];
        const int rc = jg_shape_offset(in.data(), (int)sh.v.size(), 0, offset, 2.0f, outv.data(), (int)(2 * sh.v.size()), &nv2, poi2);
        if (rc != JG_OK) return rc;
        sh.v.resize((size_t)nv2);
        for (int i = 0; i < nv2; i++) sh.v[(size_t)i] = Pt{outv[(size_t)(2 * i)], outv[(size_t)(2 * i + 1)]};
        sh.area = signed_area(sh.v);
        if (!(sh.area > 0.0f)) return JG_ERR_REFERENCE_PANIC;
        sh.bbox = bbox_of(sh.v);
        poi = Pole{poi2[0], poi2[1], poi2[2]};
    }
    if (simplify_tolerance >= 0.0f) {
        // original_shape.rs:33-45: simplify the transformed shape (Inflate for items); simplify_shape ends in
        // SPolygon::new, so area, bbox and the pole of inaccessibility are recomputed from the resulting points
        if (!simplify_points(sh.v, sh.area, false, simplify_tolerance)) return JG_ERR_REFERENCE_PANIC;
        if (concavity_ratio >= 0.0f) {
            // original_shape.rs:38-44: close narrow concavities, then simplify once more with a tenth of the tolerance
            // (each step ends in SPolygon::new: the second simplification measures against the area at its start)
            if (!close_concavities(sh.v, false, concavity_ratio)) return JG_ERR_REFERENCE_PANIC;
            const float a2 = signed_area(sh.v);
            if (!(a2 > 0.0f) || !simplify_points(sh.v, a2, false, simplify_tolerance / 10.0f)) return JG_ERR_REFERENCE_PANIC;
        }
        sh.area = signed_area(sh.v);
        if (!(sh.area > 0.0f)) return JG_ERR_REFERENCE_PANIC;
        sh.bbox = bbox_of(sh.v);
        if (!compute_pole(sh, {}, poi)) return JG_ERR_REFERENCE_PANIC;
    }
    // pole.rs:12-47 generate_surrogate_poles
    std::vector<Pole> poles{poi};
    const float PI_F = 3.14159265358979323846f;
    float total = poi.r * poi.r * PI_F;
    for (;;) {
        Pole next;
        if (!compute_pole(sh, poles, next)) return JG_ERR_REFERENCE_PANIC;
        total += next.r * next.r * PI_F;
        poles.push_back(next);
        const float coverage = total / sh.area;
        int limit = -1;
        for (int k = 0; k < 3; k++)
            if (coverage > scfg->coverage_limits[k] && (limit < 0 || scfg->n_pole_limits[k] < limit)) limit = scfg->n_pole_limits[k];
        if (limit >= 0 && (int)poles.size() >= limit) break;
        if (poles.size() >= 1000) return JG_ERR_REFERENCE_PANIC;  // pole.rs:40-43 assert
    }
    *n_verts = (int)sh.v.size();
    for (size_t i = 0; i < sh.v.size(); i++) verts_xy[2 * i] = sh.v[i].x, verts_xy[2 * i + 1] = sh.v[i].y;
    poi_xyr[0] = poi.x, poi_xyr[1] = poi.y, poi_xyr[2] = poi.r;
    *n_poles = (int)poles.size();
    if (poles_xyr)
        for (int i = 0; i < (int)poles.size() && i < poles_cap; i++)
            poles_xyr[3 * i] = poles[i].x, poles_xyr[3 * i + 1] = poles[i].y, poles_xyr[3 * i + 2] = poles[i].r;
    pre_transform_xy[0] = -cx, pre_transform_xy[1] = -cy;
    if (area) *area = area_orig;
    return JG_OK;
}

// ---- geometry/fail_fast/piers.rs:25-226 generate_piers ------------------------------------------------------------
// Piers are segments inside the shape that cover what the fail-fast poles leave unrepresented. A fan of rays (a
// vertical segment through the centroid, translated along x and rotated about the origin) is clipped to the shape;
// n times over, the clipped ray that minimises the summed squared distance of the unrepresented grid points to their
// nearest pole / selected pier / the candidate is selected. f32 throughout, every loop in the reference's order
// (the sums are sequential folds, min_by_key keeps the first minimum, sorted_by_key is stable).
// Third-party arithmetic: ndarray 0.17 `Array::linspace(a, b, n)` is restated as a + ((b - a) / (n - 1)) * i (unpinned,
// like float-cmp); `f32::sin_cos` is glibc sinf / cosf.
namespace jg {
namespace {
struct M3 {
    float m[3][3];
};
M3 dot_prod(const M3& l, const M3& r) {  // transformation.rs:150-172
    M3 o;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) o.m[i][j] = l.m[i][0] * r.m[0][j] + l.m[i][1] * r.m[1][j] + l.m[i][2] * r.m[2][j];
    return o;
}
Pt m3_point(const M3& t, const Pt& p) {  // point.rs:10-31 TRANSFORM_FORMULA
    return Pt{t.m[0][0] * p.x + t.m[0][1] * p.y + t.m[0][2] * 1.0f, t.m[1][0] * p.x + t.m[1][1] * p.y + t.m[1][2] * 1.0f};
}
struct SegPt {
    Pt s, e;
};
// edge.rs:180-228 edge_intersection(e1, e2, true): the location is x2 + t * (x1 - x2) on e1
bool edge_collides_at(const SegPt& e1, const SegPt& e2, Pt& out) {
    const float x1 = e1.s.x, y1 = e1.s.y, x2 = e1.e.x, y2 = e1.e.y;
    const float x3 = e2.s.x, y3 = e2.s.y, x4 = e2.e.x, y4 = e2.e.y;
    if (fmaxf(fminf(x1, x2), fminf(x3, x4)) > fminf(fmaxf(x1, x2), fmaxf(x3, x4))) return false;
    if (fmaxf(fminf(y1, y2), fminf(y3, y4)) > fminf(fmaxf(y1, y2), fmaxf(y3, y4))) return false;
    const float t_nom = (x2 - x4) * (y4 - y3) - (y2 - y4) * (x4 - x3);
    const float denom = (x2 - x1) * (y4 - y3) - (y2 - y1) * (x4 - x3);
    const float u_nom = (x2 - x4) * (y2 - y1) - (y2 - y4) * (x2 - x1);
    if (denom == 0.0f) return false;
    const float t = t_nom / denom, u = u_nom / denom;
    if (!(t >= 0.0f && t <= 1.0f && u >= 0.0f && u <= 1.0f)) return false;
    out = Pt{x2 + t * (x1 - x2), y2 + t * (y1 - y2)};
    return true;
}
float pt_distance(const Pt& a, const Pt& b) {  // point.rs:43-47
    const float dx = a.x - b.x, dy = a.y - b.y;
    return sqrtf(dx * dx + dy * dy);
}
float circle_distance(const Pole& c, const Pt& p) {  // circle.rs:136-147
    const float dx = p.x - c.x, dy = p.y - c.y;
    const float sq = dx * dx + dy * dy;
    return sq < c.r * c.r ? 0.0f : sqrtf(sq) - c.r;
}
std::vector<float> linspace(float a, float b, int n) {
    std::vector<float> v((size_t)n);
    const float step = n > 1 ? (b - a) / (float)(n - 1) : 0.0f;
    for (int i = 0; i < n; i++) v[(size_t)i] = a + step * (float)i;
    return v;
}
}  // namespace
}  // namespace jg

extern "C" int jg_item_generate_piers(const float* verts_xy, int n_verts, const float* poles_xyr, int n_poles, int n_piers,
                                      const JgPierConfig* cfg, float* piers_xyxy) {
    if (!verts_xy || n_verts < 3 || n_piers < 0 || n_poles < 0 || (n_poles > 0 && !poles_xyr) || (n_piers > 0 && !piers_xyxy))
        return JG_ERR_INVALID;
    if (n_piers == 0) return JG_OK;  // piers.rs:26-28
    const int rays_per_angle = cfg && cfg->rays_per_angle > 0 ? cfg->rays_per_angle : 200;  // piers.rs:16-18, release values
    const int n_angles = cfg && cfg->n_angles > 0 ? cfg->n_angles : 90;
    const int n_grid = cfg && cfg->n_points_per_dimension > 0 ? cfg->n_points_per_dimension : 100;
    const float CLIPPING_TRIM = 0.999f, ACTION_RADIUS_RATIO = 0.10f;
    Shape sh;
    sh.v.resize((size_t)n_verts);
    for (int i = 0; i < n_verts; i++) sh.v[(size_t)i] = Pt{verts_xy[2 * i], verts_xy[2 * i + 1]};
    sh.area = signed_area(sh.v);
    if (!(sh.area > 0.0f)) return JG_ERR_INVALID;  // SPolygon::new keeps the vertices counter-clockwise
    sh.bbox = bbox_of(sh.v);
    std::vector<Pole> poles((size_t)n_poles);
    for (int i = 0; i < n_poles; i++) poles[(size_t)i] = Pole{poles_xyr[3 * i], poles_xyr[3 * i + 1], poles_xyr[3 * i + 2]};
    const size_t nv = sh.v.size();
    const Rect bbox = sh.bbox;
    Rect sq = bbox;  // rect.rs:93-109 inflate_to_square
    {
        const float w = sq.x_max - sq.x_min, h = sq.y_max - sq.y_min;
        float dx = 0.0f, dy = 0.0f;
        if (h < w) dy = (w - h) / 2.0f;
        else if (w < h) dx = (h - w) / 2.0f;
        sq = Rect{sq.x_min - dx, sq.y_min - dy, sq.x_max + dx, sq.y_max + dy};
    }
    const float sq_w = sq.x_max - sq.x_min, sq_h = sq.y_max - sq.y_min;
    float cx = 0.0f, cy = 0.0f;  // simple_polygon.rs:173-192 centroid
    for (size_t i = 0; i < nv; i++) {
        const size_t j = i + 1 == nv ? 0 : i + 1;
        const float cr = sh.v[i].x * sh.v[j].y - sh.v[j].x * sh.v[i].y;
        cx += (sh.v[i].x + sh.v[j].x) * cr;
        cy += (sh.v[i].y + sh.v[j].y) * cr;
    }
    cx /= 6.0f * sh.area;
    cy /= 6.0f * sh.area;
    const SegPt base_ray{Pt{cx, cy - 2.0f * sq_h}, Pt{cx, cy + 2.0f * sq_h}};  // piers.rs:36-40
    if (base_ray.s.x == base_ray.e.x && base_ray.s.y == base_ray.e.y) return JG_ERR_REFERENCE_PANIC;  // Edge::try_new
    // piers.rs:101-127 generate_ray_transformations: translation (x, 0), then rotation (angle-major order)
    const float dx_ray = sq_w / (float)rays_per_angle;
    const std::vector<float> angles = linspace(0.0f, 3.14159265358979323846f, n_angles + 1);
    std::vector<SegPt> clipped;
    std::vector<std::pair<float, Pt>> hits;
    for (int a = 0; a < n_angles; a++) {
        const float sn = sinf(angles[(size_t)a]), cs = cosf(angles[(size_t)a]);
        const M3 rot{{{cs, -sn, 0.0f}, {sn, cs, 0.0f}, {0.0f, 0.0f, 1.0f}}};
        for (int i = 0; i < rays_per_angle; i++) {
            const float x = sq.x_min + dx_ray * (float)i;
            const M3 tr{{{1.0f, 0.0f, x}, {0.0f, 1.0f, 0.0f}, {0.0f, 0.0f, 1.0f}}};
            const M3 t = dot_prod(rot, tr);  // Transformation::rotate (transformation.rs:36-39)
            const SegPt ray{m3_point(t, base_ray.s), m3_point(t, base_ray.e)};
            // piers.rs:130-157 clip
            const Xf id{1.0f, 0.0f, 0.0f, 0.0f};
            if (poly_point<false>(&sh.v[0].x, (int)nv, sh.bbox, ray.s.x, ray.s.y, id) ||
                poly_point<false>(&sh.v[0].x, (int)nv, sh.bbox, ray.e.x, ray.e.y, id))
                return JG_ERR_REFERENCE_PANIC;  // piers.rs:132 assert!
            hits.clear();
            for (size_t k = 0; k < nv; k++) {
                const SegPt edge{sh.v[k], sh.v[k + 1 == nv ? 0 : k + 1]};
                Pt at;
                if (edge_collides_at(edge, ray, at)) hits.push_back({pt_distance(ray.s, at), at});
            }
            for (const auto& h : hits)
                if (h.first != h.first) return JG_ERR_REFERENCE_PANIC;  // NotNan::new(..).unwrap()
            std::stable_sort(hits.begin(), hits.end(), [](const std::pair<float, Pt>& l, const std::pair<float, Pt>& r) { return l.first < r.first; });
            for (size_t k = 0; k + 1 < hits.size(); k += 2) {
                const Pt s = hits[k].second, e = hits[k + 1].second;
                if (s.x == e.x && s.y == e.y) continue;
                // Edge::scale (edge.rs:38-45)
                const float ddx = e.x - s.x, ddy = e.y - s.y;
                const float f1 = CLIPPING_TRIM - 1.0f;
                SegPt c{s, e};
                c.s.x -= ddx * f1 / 2.0f;
                c.s.y -= ddy * f1 / 2.0f;
                c.e.x += ddx * f1 / 2.0f;
                c.e.y += ddy * f1 / 2.0f;
                clipped.push_back(c);
            }
        }
    }
    // piers.rs:159-178 generate_unrepresented_point_grid (x-major)
    std::vector<Pt> grid;
    {
        const std::vector<float> xr = linspace(sq.x_min, sq.x_max, n_grid), yr = linspace(sq.y_min, sq.y_max, n_grid);
        const Xf id{1.0f, 0.0f, 0.0f, 0.0f};
        for (float x : xr)
            for (float y : yr) {
                if (!poly_point<false>(&sh.v[0].x, (int)nv, sh.bbox, x, y, id)) continue;
                bool in_pole = false;
                for (const Pole& c : poles) {  // circle.rs:130-134
                    const float ddx = x - c.x, ddy = y - c.y;
                    if (ddx * ddx + ddy * ddy <= c.r * c.r) { in_pole = true; break; }
                }
                if (!in_pole) grid.push_back(Pt{x, y});
            }
    }
    const float radius_of_ray_influence = ACTION_RADIUS_RATIO * sq_w;
    const float bw = bbox.x_max - bbox.x_min, bh = bbox.y_max - bbox.y_min;
    const float forfeit = sqrtf((bw * bw) * (bh * bh));  // piers.rs:58 (sic: a product)
    std::vector<float> d_poles(grid.size()), d_rays(grid.size());
    for (size_t g = 0; g < grid.size(); g++) {  // piers.rs:216-226
        float m = forfeit;
        for (const Pole& c : poles) m = fminf(m, circle_distance(c, grid[g]));
        d_poles[g] = m;
    }
    std::vector<SegPt> selected;
    for (int it = 0; it < n_piers; it++) {
        for (size_t g = 0; g < grid.size(); g++) {  // piers.rs:204-214
            float m = forfeit;
            for (const SegPt& r : selected) m = fminf(m, edge_distance(r.s, r.e, grid[g].x, grid[g].y));
            d_rays[g] = m;
        }
        if (clipped.empty()) return JG_ERR_REFERENCE_PANIC;  // piers.rs:92 bail!("no ray found")
        size_t best = 0;
        float best_loss = 0.0f;
        for (size_t r = 0; r < clipped.size(); r++) {  // piers.rs:180-202 loss_function
            float loss = 0.0f;
            for (size_t g = 0; g < grid.size(); g++) {
                const float dn = edge_distance(clipped[r].s, clipped[r].e, grid[g].x, grid[g].y);
                const float dr = fminf(d_rays[g], dn);
                const float d = dr < radius_of_ray_influence ? fminf(d_poles[g], dr) : d_poles[g];
                loss += d * d;
            }
            if (loss != loss) return JG_ERR_REFERENCE_PANIC;  // NotNan::new(x).unwrap()
            if (r == 0 || loss < best_loss) best = r, best_loss = loss;  // min_by_key: the first minimum
        }
        selected.push_back(clipped[best]);
    }
    for (int i = 0; i < n_piers; i++) {
        piers_xyxy[4 * i] = selected[(size_t)i].s.x, piers_xyxy[4 * i + 1] = selected[(size_t)i].s.y;
        piers_xyxy[4 * i + 2] = selected[(size_t)i].e.x, piers_xyxy[4 * i + 3] = selected[(size_t)i].e.y;
    }
    return JG_OK;
}

extern "C" int jg_item_import(const float* pts_xy, int n_pts, const JgSurrogateConfig* scfg, float* verts_xy, int* n_verts,
                              float poi_xyr[3], float* poles_xyr, int poles_cap, int* n_poles, float pre_transform_xy[2],
                              float* area) {
    return item_import_impl(pts_xy, n_pts, scfg, verts_xy, n_verts, poi_xyr, poles_xyr, poles_cap, n_poles, pre_transform_xy, area, -1.0f, -1.0f);
}
// ... with Importer's poly_simpl_tolerance (io/import.rs:31-45; lbf default 0.001): the centred shape is simplified in
// Inflate mode before the surrogate is generated; narrow_concavity_cutoff_ratio >= 0 (lbf default 0.01) adds
// close_narrow_concavities and the second simplification pass (original_shape.rs:38-44); < 0 = None.
extern "C" int jg_item_import_simplified(const float* pts_xy, int n_pts, const JgSurrogateConfig* scfg, float simplify_tolerance,
                                         float narrow_concavity_cutoff_ratio, float* verts_xy, int* n_verts, float poi_xyr[3],
                                         float* poles_xyr, int poles_cap, int* n_poles, float pre_transform_xy[2], float* area) {
    if (!(simplify_tolerance >= 0.0f)) return JG_ERR_INVALID;
    return item_import_impl(pts_xy, n_pts, scfg, verts_xy, n_verts, poi_xyr, poles_xyr, poles_cap, n_poles, pre_transform_xy, area,
                            simplify_tolerance, narrow_concavity_cutoff_ratio);
}
// ... with the whole ShapeModifyConfig (geometry/shape_modification.rs:22-46): offset > 0 inflates the centred shape first
// (jg_shape_offset), simplify_tolerance >= 0 / narrow_concavity_cutoff_ratio >= 0 as above (< 0 = None).
extern "C" int jg_item_import_modified(const float* pts_xy, int n_pts, const JgSurrogateConfig* scfg, float offset,
                                       float simplify_tolerance, float narrow_concavity_cutoff_ratio, float* verts_xy, int verts_cap,
                                       int* n_verts, float poi_xyr[3], float* poles_xyr, int poles_cap, int* n_poles,
                                       float pre_transform_xy[2], float* area) {
    if (!(offset >= 0.0f) || verts_cap < 2 * n_pts) return JG_ERR_INVALID;
    return item_import_impl(pts_xy, n_pts, scfg, verts_xy, n_verts, poi_xyr, poles_xyr, poles_cap, n_poles, pre_transform_xy, area,
                            simplify_tolerance, narrow_concavity_cutoff_ratio, offset);
}
