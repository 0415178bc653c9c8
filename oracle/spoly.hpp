// ORACLE — TEST INFRASTRUCTURE ONLY. PARITY UNPINNED (see geom.hpp header).
// Restatement of geometry/primitives/simple_polygon.rs, geometry/convex_hull.rs,
// geometry/fail_fast/{pole,sp_surrogate}.rs and the polygon part of io/import.rs.
#pragma once
#include <algorithm>
#include <deque>
#include <stdexcept>
#include <string>

#include "geom.hpp"

namespace orc {

struct SPSurrogateConfig {  // fail_fast/sp_surrogate.rs:120-131
    int n_pole_limits_n[3] = {100, 20, 10};
    float n_pole_limits_cov[3] = {0.0f, 0.75f, 0.90f};
    int n_ff_poles = 2;
    int n_ff_piers = 0;
};

struct SPSurrogate {  // sp_surrogate.rs:17-30 (convex hull fields are not on the query path)
    std::vector<Circle> poles;
    std::vector<Edge> piers;
    SPSurrogateConfig config;
    // sp_surrogate.rs:55-61
    int n_ff_poles() const { return config.n_ff_poles; }
};

struct SPolygon {
    std::vector<Point> vertices;
    Rect bbox;
    float area;
    float diameter;
    Circle poi;
    bool has_surrogate = false;
    SPSurrogate surrogate;

    int n_vertices() const { return (int)vertices.size(); }
    // simple_polygon.rs:86-93
    Edge edge(int i) const {
        int j = (i == n_vertices() - 1) ? 0 : i + 1;
        return {vertices[i], vertices[j]};
    }

    // simple_polygon.rs:122-133
    static Rect generate_bounding_box(const std::vector<Point>& pts) {
        float x_min = std::numeric_limits<float>::max(), y_min = x_min;
        float x_max = std::numeric_limits<float>::lowest(), y_max = x_max;
        for (const Point& p : pts) {
            ORC_FLOPS(4);
            x_min = std::fmin(x_min, p.x);
            y_min = std::fmin(y_min, p.y);
            x_max = std::fmax(x_max, p.x);
            y_max = std::fmax(y_max, p.y);
        }
        return {x_min, y_min, x_max, y_max};
    }
    // simple_polygon.rs:137-150 (shoelace; CCW positive)
    static float calculate_area(const std::vector<Point>& pts) {
        float sigma = 0.0f;
        size_t n = pts.size();
        for (size_t i = 0; i < n; i++) {
            size_t j = (i + 1) % n;
            sigma += (pts[i].y + pts[j].y) * (pts[i].x - pts[j].x);
        }
        return 0.5f * sigma;
    }
    // simple_polygon.rs:173-192
    Point centroid() const {
        float c_x = 0.0f, c_y = 0.0f;
        int n = n_vertices();
        for (int i = 0; i < n; i++) {
            int j = (i == n - 1) ? 0 : i + 1;
            float x_i = vertices[i].x, y_i = vertices[i].y, x_j = vertices[j].x, y_j = vertices[j].y;
            c_x += (x_i + x_j) * (x_i * y_j - x_j * y_i);
            c_y += (y_i + y_j) * (x_i * y_j - x_j * y_i);
        }
        c_x /= 6.0f * area;
        c_y /= 6.0f * area;
        return {c_x, c_y};
    }

    // simple_polygon.rs:255-292 — point in polygon by ray casting with the FPA vertex rule
    bool collides_with(const Point& point) const {
        if (!bbox.collides_with(point)) return false;
        ORC_FLOPS(20);
        Point point_outside{bbox.x_max + bbox.width(), point.y};
        Edge ray{point, point_outside};
        int n_intersections = 0;
        int n = n_vertices();
        for (int i = 0; i < n; i++) {
            Edge e = edge(i);
            float s_x = e.start.x, s_y = e.start.y, e_x = e.end.x, e_y = e.end.y;
            float p_x = point.x, p_y = point.y;
            if ((fpa_eq(s_y, p_y) && fpa_gt(s_x, p_x)) || (fpa_eq(e_y, p_y) && fpa_gt(e_x, p_x))) {
                if (fpa_lt(s_y, p_y) || fpa_lt(e_y, p_y)) n_intersections += 1;
            } else if (edge_edge(ray, e)) {
                n_intersections += 1;
            }
        }
        return n_intersections % 2 == 1;
    }

    // simple_polygon.rs:195-224 Transformable::transform (in place; area/diameter untouched)
    void transform(const Transformation& t) {
        for (Point& p : vertices) p = transform_point(p, t);
        poi.center = transform_point(poi.center, t);
        if (has_surrogate) {  // sp_surrogate.rs:64-87
            for (Circle& c : surrogate.poles) c.center = transform_point(c.center, t);
            for (Edge& e : surrogate.piers) e = e.transformed(t);
        }
        bbox = generate_bounding_box(vertices);
    }
    // simple_polygon.rs:226-253 TransformableFrom (buffer has no surrogate in the callers)
    void transform_from(const SPolygon& ref, const Transformation& t) {
        size_t n = std::min(vertices.size(), ref.vertices.size());  // zip
        for (size_t i = 0; i < n; i++) vertices[i] = transform_point(ref.vertices[i], t);
        poi.center = transform_point(ref.poi.center, t);
        poi.radius = poi.radius;
        bbox = generate_bounding_box(vertices);
    }
};

// ---- geometry/convex_hull.rs:30-63 -------------------------------------------------------
inline float hull_cross(Point a, Point b, Point c) {
    return (b.x - a.x) * (c.y - a.y) - (b.y - a.y) * (c.x - a.x);
}
inline void grow_convex_hull(std::vector<Point>& h, Point next) {
    while (h.size() >= 2 && hull_cross(h[h.size() - 2], h[h.size() - 1], next) <= 0.0f) h.pop_back();
    h.push_back(next);
}
inline std::vector<Point> convex_hull_from_points(std::vector<Point> pts) {
    std::stable_sort(pts.begin(), pts.end(), [](const Point& a, const Point& b) { return a.x < b.x; });
    std::vector<Point> lower, upper;
    for (size_t i = 0; i < pts.size(); i++) grow_convex_hull(lower, pts[i]);
    for (size_t i = pts.size(); i-- > 0;) grow_convex_hull(upper, pts[i]);
    upper.pop_back();
    lower.pop_back();
    lower.insert(lower.end(), upper.begin(), upper.end());
    return lower;
}
// simple_polygon.rs:107-120
inline float calculate_diameter(const std::vector<Point>& pts) {
    std::vector<Point> ch = convex_hull_from_points(pts);
    float best = -1.0f;
    for (size_t i = 0; i < ch.size(); i++)
        for (size_t j = i + 1; j < ch.size(); j++) best = std::fmax(best, sq_distance(ch[i], ch[j]));
    if (best < 0.0f) throw std::runtime_error("convex hull is empty");
    return std::sqrt(best);
}

// ---- geometry/fail_fast/pole.rs ----------------------------------------------------------
struct POINode {  // pole.rs:81-135
    int level;
    Rect bbox;
    float radius;
    float distance;
    POINode(Rect bb, int lvl, const SPolygon& poly, const std::vector<Circle>& poles)
        : level(lvl), bbox(bb) {
        radius = bbox.diameter() / 2.0f;
        Point c = bbox.centroid();
        bool centroid_inside = poly.collides_with(c);
        if (centroid_inside)
            for (const Circle& p : poles)
                if (p.collides_with(c)) {
                    centroid_inside = false;
                    break;
                }
        float d = std::numeric_limits<float>::max();
        for (int i = 0; i < poly.n_vertices(); i++) d = std::fmin(d, poly.edge(i).distance_to(c));
        for (const Circle& p : poles) d = std::fmin(d, p.separation_distance_abs(c));
        distance = centroid_inside ? d : -d;
    }
    float distance_upperbound() const { return radius + distance; }
};
constexpr int MAX_POI_TREE_DEPTH = 10;  // pole.rs:79

// pole.rs:52-77
inline Circle compute_pole(const SPolygon& shape, const std::vector<Circle>& poles) {
    Rect square_bbox = shape.bbox.inflate_to_square();
    std::deque<POINode> queue;
    queue.emplace_back(square_bbox, MAX_POI_TREE_DEPTH, shape, poles);
    bool has_best = false;
    Circle best{{0, 0}, 0.0f};
    while (!queue.empty()) {
        POINode node = queue.front();
        queue.pop_front();
        float best_d = has_best ? best.radius : 0.0f;
        if (node.distance > best_d) {
            best = {node.bbox.centroid(), node.distance};
            has_best = true;
        }
        best_d = has_best ? best.radius : 0.0f;
        if (node.distance_upperbound() > best_d && node.level != 0) {
            Rect q[4];
            node.bbox.quadrants(q);
            for (int i = 0; i < 4; i++) queue.emplace_back(q[i], node.level - 1, shape, poles);
        }
    }
    if (!has_best) throw std::runtime_error("no pole found");
    return best;
}

// pole.rs:12-48
inline std::vector<Circle> generate_surrogate_poles(const SPolygon& shape, const SPSurrogateConfig& cfg) {
    std::vector<Circle> all_poles{shape.poi};
    float total_pole_area = shape.poi.area();
    for (;;) {
        Circle next = compute_pole(shape, all_poles);
        total_pole_area += next.area();
        all_poles.push_back(next);
        float current_coverage = total_pole_area / shape.area;
        bool have = false;
        int active = 0;
        for (int k = 0; k < 3; k++)
            if (current_coverage > cfg.n_pole_limits_cov[k]) {
                if (!have || cfg.n_pole_limits_n[k] < active) active = cfg.n_pole_limits_n[k];
                have = true;
            }
        if (have && (int)all_poles.size() >= active) break;
        if (all_poles.size() >= 1000) throw std::runtime_error("more than 1000 poles generated");
    }
    return all_poles;
}

// simple_polygon.rs:41-71 SPolygon::new
inline SPolygon make_spolygon(std::vector<Point> points) {
    if (points.size() < 3) throw std::runtime_error("simple polygon must have at least 3 points");
    for (size_t i = 0; i < points.size(); i++)
        for (size_t j = i + 1; j < points.size(); j++)
            if (points[i] == points[j]) throw std::runtime_error("simple polygon has duplicate points");
    float area = SPolygon::calculate_area(points);
    if (area == 0.0f) throw std::runtime_error("simple polygon has no area");
    if (area < 0.0f) {
        std::reverse(points.begin(), points.end());
        area = -area;
    }
    SPolygon sp;
    sp.diameter = calculate_diameter(points);
    sp.bbox = SPolygon::generate_bounding_box(points);
    sp.vertices = std::move(points);
    sp.area = area;
    // simple_polygon.rs:152-171 calculate_poi: pole of the polygon itself with no prior poles
    sp.poi = {{std::numeric_limits<float>::max(), std::numeric_limits<float>::max()},
              std::numeric_limits<float>::max()};
    sp.poi = compute_pole(sp, {});
    return sp;
}

// simple_polygon.rs:340-343 From<Rect>
inline SPolygon spolygon_from_rect(const Rect& r) {
    return make_spolygon({{r.x_min, r.y_min}, {r.x_max, r.y_min}, {r.x_max, r.y_max}, {r.x_min, r.y_max}});
}

// simple_polygon.rs:73-80 + sp_surrogate.rs:33-53 (piers are pre-processing, not generated here:
// n_ff_piers == 0 at the reference's default config; tests inject piers directly)
inline void generate_surrogate(SPolygon& sp, const SPSurrogateConfig& cfg) {
    sp.surrogate.poles = generate_surrogate_poles(sp, cfg);
    sp.surrogate.piers.clear();
    sp.surrogate.config = cfg;
    sp.has_surrogate = true;
}

// io/import.rs:236-260
inline void eliminate_degenerate_vertices(std::vector<Point>& points) {
    std::vector<size_t> to_remove;
    size_t n = points.size();
    for (size_t i = 0; i < n; i++) {
        size_t j = (i + 1) % n;
        if (approx_eq(points[i].x, points[j].x) && approx_eq(points[i].y, points[j].y))
            to_remove.push_back(i);
    }
    std::sort(to_remove.begin(), to_remove.end(), [](size_t a, size_t b) { return a > b; });
    for (size_t idx : to_remove)
        if (idx < points.size()) points.erase(points.begin() + idx);
}
// io/import.rs:198-211
inline SPolygon import_simple_polygon(std::vector<Point> points) {
    if (points.size() > 1 && points.front() == points.back()) points.pop_back();
    eliminate_degenerate_vertices(points);
    return make_spolygon(std::move(points));  // make_spolygon bails on any remaining duplicates
}

}  // namespace orc
