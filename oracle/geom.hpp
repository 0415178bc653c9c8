// ORACLE — TEST INFRASTRUCTURE ONLY. PARITY UNPINNED.
//
// CPU restatement of jagua-rs's f32 geometry primitives (reference v0.6.4). Nothing under oracle/
// is product code: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
// reference legs may load it. The reference (Rust) cannot be compiled in this image and ships no
// golden vectors for collision queries, so this restatement is NOT pinned against reference
// outputs ("parity unpinned"); it is pinned only by the structural invariants the reference's own
// debug assertions state (tests/test_oracle_invariants.py) and by a brute-force cross-check.
//
// Every function cites the reference file:line (relative to /root/reference/jagua-rs/src) it
// follows. Build with -ffp-contract=off -fno-fast-math: Rust never contracts a*b+c into an FMA.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>

namespace orc {

// Algorithmic flop counters. Only bumped when ORC_COUNT is defined so
// the timed CPU baseline is not slowed down by the bookkeeping.
#ifdef ORC_COUNT
extern thread_local uint64_t g_flops;
#define ORC_FLOPS(n) (::orc::g_flops += (n))
#else
#define ORC_FLOPS(n) ((void)0)
#endif

enum class GeoPosition { Exterior, Interior };                           // geometry/geo_enums.rs:2-6
enum class GeoRelation { Intersecting, Enclosed, Surrounding, Disjoint };  // geo_enums.rs:8-20

// util/fpa.rs:18-31 over float-cmp 0.10 `approx_eq!(f32, a, b)` with the default F32Margin
// {epsilon: f32::EPSILON, ulps: 4}. float-cmp is a third-party crate that is not vendored in the
// reference tree; this is a restatement of its published algorithm:
//   a == b || |a-b| <= epsilon || saturating_abs(bits(a) as i32 wrapping_sub bits(b) as i32) <= ulps
inline bool approx_eq(float a, float b) {
    ORC_FLOPS(4);
    if (a == b) return true;
    float eps = std::fabs(a - b);
    if (eps <= std::numeric_limits<float>::epsilon()) return true;
    int32_t ai, bi;
    std::memcpy(&ai, &a, 4);
    std::memcpy(&bi, &b, 4);
    int32_t diff = (int32_t)((uint32_t)ai - (uint32_t)bi);  // wrapping_sub
    int32_t ad = diff < 0 ? (diff == INT32_MIN ? INT32_MAX : -diff) : diff;
    return ad <= 4;
}
// FPA ordering: Equal when approx-equal, else exact partial_cmp (fpa.rs:24-31)
inline bool fpa_eq(float a, float b) { return approx_eq(a, b); }
inline bool fpa_lt(float a, float b) { return !approx_eq(a, b) && a < b; }
inline bool fpa_gt(float a, float b) { return !approx_eq(a, b) && a > b; }
inline bool fpa_le(float a, float b) { return approx_eq(a, b) || a < b; }
inline bool fpa_ge(float a, float b) { return approx_eq(a, b) || a > b; }

struct Point {
    float x, y;
    bool operator==(const Point& o) const { return x == o.x && y == o.y; }
};

// geometry/transformation.rs:11-13 — 3x3 row-major matrix
struct Transformation {
    float m[3][3];
    static Transformation empty() { return {{{1, 0, 0}, {0, 1, 0}, {0, 0, 1}}}; }
};

// geometry/d_transformation.rs:11-16
struct DTransformation {
    float rotation, tx, ty;
    bool operator==(const DTransformation& o) const {
        return rotation == o.rotation && tx == o.tx && ty == o.ty;
    }
};

// transformation.rs:131-139 rot_transl_m; f32::sin_cos lowers to libm sinf/cosf on linux-gnu.
inline Transformation compose(const DTransformation& dt) {
    float s = sinf(dt.rotation), c = cosf(dt.rotation);
    return {{{c, -s, dt.tx}, {s, c, dt.ty}, {0, 0, 1}}};
}
// The boundary format of a query: the composed (cos, sin, tx, ty); m01 = -sin is exact.
inline Transformation from_cstt(float c, float s, float tx, float ty) {
    return {{{c, -s, tx}, {s, c, ty}, {0, 0, 1}}};
}

// primitives/point.rs:26-31 TRANSFORM_FORMULA: (m00*x + m01*y) + m02*1.0, non-fused
inline Point transform_point(Point p, const Transformation& t) {
    ORC_FLOPS(8);
    float tx = t.m[0][0] * p.x + t.m[0][1] * p.y + t.m[0][2] * 1.0f;
    float ty = t.m[1][0] * p.x + t.m[1][1] * p.y + t.m[1][2] * 1.0f;
    return {tx, ty};
}

// point.rs:49-52
inline float sq_distance(Point a, Point b) {
    return (a.x - b.x) * (a.x - b.x) + (a.y - b.y) * (a.y - b.y);
}
inline float distance(Point a, Point b) { return std::sqrt(sq_distance(a, b)); }

struct Edge;
struct Rect {
    float x_min, y_min, x_max, y_max;
    bool operator==(const Rect& o) const {
        return x_min == o.x_min && y_min == o.y_min && x_max == o.x_max && y_max == o.y_max;
    }
    float width() const { return x_max - x_min; }    // rect.rs:208-210
    float height() const { return y_max - y_min; }   // rect.rs:212-214
    Point centroid() const {                         // rect.rs:248-253
        return {(x_min + x_max) / 2.0f, (y_min + y_max) / 2.0f};
    }
    float diameter() const {                         // rect.rs:259-263
        float dx = x_max - x_min, dy = y_max - y_min;
        return std::sqrt(dx * dx + dy * dy);
    }
    // rect.rs:154-161 — NE, NW, SW, SE
    void corners(Point c[4]) const {
        c[0] = {x_max, y_max};
        c[1] = {x_min, y_max};
        c[2] = {x_min, y_min};
        c[3] = {x_max, y_min};
    }
    // rect.rs:35-41 (the try_new validity check is a panic path in the reference)
    static Rect from_diagonal_corners(Point a, Point b) {
        return {std::fmin(a.x, b.x), std::fmin(a.y, b.y), std::fmax(a.x, b.x), std::fmax(a.y, b.y)};
    }
    // rect.rs:141-151
    void quadrants(Rect q[4]) const {
        Point mid = centroid();
        Point c[4];
        corners(c);
        for (int i = 0; i < 4; i++) q[i] = from_diagonal_corners(c[i], mid);
    }
    // rect.rs:93-109
    Rect inflate_to_square() const {
        float w = x_max - x_min, h = y_max - y_min, dx = 0.0f, dy = 0.0f;
        if (h < w) dy = (w - h) / 2.0f;
        else if (w < h) dx = (h - w) / 2.0f;
        return {x_min - dx, y_min - dy, x_max + dx, y_max + dy};
    }
    // rect.rs:266-272
    bool collides_with(const Rect& o) const {
        ORC_FLOPS(6);
        return std::fmax(x_min, o.x_min) <= std::fmin(x_max, o.x_max) &&
               std::fmax(y_min, o.y_min) <= std::fmin(y_max, o.y_max);
    }
    // rect.rs:274-280
    bool almost_collides_with(const Rect& o) const {
        return fpa_le(std::fmax(x_min, o.x_min), std::fmin(x_max, o.x_max)) &&
               fpa_le(std::fmax(y_min, o.y_min), std::fmin(y_max, o.y_max));
    }
    // rect.rs:282-288
    bool collides_with(const Point& p) const {
        return p.x >= x_min && p.x <= x_max && p.y >= y_min && p.y <= y_max;
    }
    // rect.rs:46-65
    GeoRelation relation_to(const Rect& o) const {
        if (!collides_with(o)) return GeoRelation::Disjoint;
        ORC_FLOPS(8);
        if (x_min <= o.x_min && y_min <= o.y_min && x_max >= o.x_max && y_max >= o.y_max)
            return GeoRelation::Surrounding;
        if (x_min >= o.x_min && y_min >= o.y_min && x_max <= o.x_max && y_max <= o.y_max)
            return GeoRelation::Enclosed;
        return GeoRelation::Intersecting;
    }
    // rect.rs:70-89
    GeoRelation almost_relation_to(const Rect& o) const {
        if (!almost_collides_with(o)) return GeoRelation::Disjoint;
        if (fpa_le(x_min, o.x_min) && fpa_le(y_min, o.y_min) && fpa_ge(x_max, o.x_max) &&
            fpa_ge(y_max, o.y_max))
            return GeoRelation::Surrounding;
        if (fpa_ge(x_min, o.x_min) && fpa_ge(y_min, o.y_min) && fpa_le(x_max, o.x_max) &&
            fpa_le(y_max, o.y_max))
            return GeoRelation::Enclosed;
        return GeoRelation::Intersecting;
    }
    inline bool collides_with(const Edge& e) const;
};

struct Edge {
    Point start, end;
    bool operator==(const Edge& o) const { return start == o.start && end == o.end; }
    float x_min() const { return std::fmin(start.x, end.x); }  // edge.rs:89-103
    float y_min() const { return std::fmin(start.y, end.y); }
    float x_max() const { return std::fmax(start.x, end.x); }
    float y_max() const { return std::fmax(start.y, end.y); }

    // edge.rs:63-87
    Point closest_point_on_edge(Point p) const {
        float x1 = start.x, y1 = start.y, x2 = end.x, y2 = end.y;
        float a = p.x - x1, b = p.y - y1, c = x2 - x1, d = y2 - y1;
        float dot = a * c + b * d;
        float len_sq = c * c + d * d;
        float param = -1.0f;
        if (len_sq != 0.0f) param = dot / len_sq;
        if (param < 0.0f) return {x1, y1};
        if (param > 1.0f) return {x2, y2};
        return {x1 + param * c, y1 + param * d};
    }
    // edge.rs:153-159
    float sq_distance_to(Point p) const {
        ORC_FLOPS(25);
        Point c = closest_point_on_edge(p);
        float dx = p.x - c.x, dy = p.y - c.y;
        return dx * dx + dy * dy;
    }
    float distance_to(Point p) const { return std::sqrt(sq_distance_to(p)); }  // edge.rs:148-150
    Edge transformed(const Transformation& t) const {                          // edge.rs:126-134
        return {transform_point(start, t), transform_point(end, t)};
    }
};

// edge.rs:180-228 edge_intersection(e1, e2, calc_loc = false) as used by CollidesWith<Edge>
inline bool edge_edge(const Edge& e1, const Edge& e2) {
    float x1 = e1.start.x, y1 = e1.start.y, x2 = e1.end.x, y2 = e1.end.y;
    float x3 = e2.start.x, y3 = e2.start.y, x4 = e2.end.x, y4 = e2.end.y;
    {
        float x_min_e1 = std::fmin(x1, x2), x_max_e1 = std::fmax(x1, x2);
        float y_min_e1 = std::fmin(y1, y2), y_max_e1 = std::fmax(y1, y2);
        float x_min_e2 = std::fmin(x3, x4), x_max_e2 = std::fmax(x3, x4);
        float y_min_e2 = std::fmin(y3, y4), y_max_e2 = std::fmax(y3, y4);
        bool x_no = std::fmax(x_min_e1, x_min_e2) > std::fmin(x_max_e1, x_max_e2);
        bool y_no = std::fmax(y_min_e1, y_min_e2) > std::fmin(y_max_e1, y_max_e2);
        ORC_FLOPS(14);
        if (x_no || y_no) return false;
    }
    ORC_FLOPS(29);
    float t_nom = (x2 - x4) * (y4 - y3) - (y2 - y4) * (x4 - x3);
    float t_denom = (x2 - x1) * (y4 - y3) - (y2 - y1) * (x4 - x3);
    float u_nom = (x2 - x4) * (y2 - y1) - (y2 - y4) * (x2 - x1);
    float u_denom = (x2 - x1) * (y4 - y3) - (y2 - y1) * (x4 - x3);
    if (t_denom == 0.0f || u_denom == 0.0f) return false;
    float t = t_nom / t_denom, u = u_nom / u_denom;
    return t >= 0.0f && t <= 1.0f && u >= 0.0f && u <= 1.0f;
}

// rect.rs:301-345
inline bool Rect::collides_with(const Edge& e) const {
    float ex_min = e.x_min(), ex_max = e.x_max(), ey_min = e.y_min(), ey_max = e.y_max();
    bool x_no = std::fmax(ex_min, x_min) > std::fmin(ex_max, x_max);
    bool y_no = std::fmax(ey_min, y_min) > std::fmin(ey_max, y_max);
    ORC_FLOPS(10);
    if (x_no || y_no) return false;
    ORC_FLOPS(8);
    if (collides_with(e.start) || collides_with(e.end)) return true;
    ORC_FLOPS(30);
    float s_x = e.start.x, s_y = e.start.y, e_x = e.end.x, e_y = e.end.y;
    float edge_dx = e_x - s_x, edge_dy = e_y - s_y;
    Point c[4];
    corners(c);
    float sides[4];
    for (int i = 0; i < 4; i++) sides[i] = (c[i].x - s_x) * edge_dy - (c[i].y - s_y) * edge_dx;
    bool all_pos = true, all_neg = true;
    for (int i = 0; i < 4; i++) {
        all_pos = all_pos && sides[i] > 0.0f;
        all_neg = all_neg && sides[i] < 0.0f;
    }
    return !(all_pos || all_neg);
}

struct Circle {
    Point center;
    float radius;
    float area() const { return radius * radius * 3.14159265358979323846f; }  // circle.rs:60-62
    // circle.rs:109-113
    bool collides_with(const Edge& e) const { return e.sq_distance_to(center) <= radius * radius; }
    // circle.rs:115-128
    bool collides_with(const Rect& r) const {
        ORC_FLOPS(11);
        float nx = std::fmax(r.x_min, std::fmin(center.x, r.x_max));
        float ny = std::fmax(r.y_min, std::fmin(center.y, r.y_max));
        return (nx - center.x) * (nx - center.x) + (ny - center.y) * (ny - center.y) <=
               radius * radius;
    }
    // circle.rs:130-134 (point.sq_distance_to(center))
    bool collides_with(const Point& p) const { return sq_distance(p, center) <= radius * radius; }
    // circle.rs:154-163 — .1 of separation_distance(point)
    float separation_distance_abs(const Point& p) const {
        float d = std::sqrt((p.x - center.x) * (p.x - center.x) + (p.y - center.y) * (p.y - center.y));
        return d <= radius ? radius - d : d - radius;
    }
    Circle transformed(const Transformation& t) const {  // circle.rs:79-85: radius unchanged
        return {transform_point(center, t), radius};
    }
};

inline bool collides(const Edge& a, const Edge& b) { return edge_edge(a, b); }
inline bool collides(const Edge& a, const Rect& r) { return r.collides_with(a); }  // edge.rs:172-177
inline bool collides(const Circle& c, const Edge& e) { return c.collides_with(e); }
inline bool collides(const Circle& c, const Rect& r) { return c.collides_with(r); }

}  // namespace orc
