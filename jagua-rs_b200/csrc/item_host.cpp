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
//   geometry/fail_fast/sp_surrogate.rs:33-61   SPSurrogate::new (poles; piers only when n_ff_piers > 0: not generated here)
#include <cmath>
#include <cstdint>
#include <deque>
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

extern "C" int jg_item_import(const float* pts_xy, int n_pts, const JgSurrogateConfig* scfg, float* verts_xy, int* n_verts,
                              float poi_xyr[3], float* poles_xyr, int poles_cap, int* n_poles, float pre_transform_xy[2],
                              float* area) {
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
    *n_verts = (int)nv;
    for (size_t i = 0; i < nv; i++) verts_xy[2 * i] = sh.v[i].x, verts_xy[2 * i + 1] = sh.v[i].y;
    poi_xyr[0] = poi.x, poi_xyr[1] = poi.y, poi_xyr[2] = poi.r;
    *n_poles = (int)poles.size();
    if (poles_xyr)
        for (int i = 0; i < (int)poles.size() && i < poles_cap; i++)
            poles_xyr[3 * i] = poles[i].x, poles_xyr[3 * i + 1] = poles[i].y, poles_xyr[3 * i + 2] = poles[i].r;
    pre_transform_xy[0] = -cx, pre_transform_xy[1] = -cy;
    if (area) *area = sh.area;
    return JG_OK;
}
