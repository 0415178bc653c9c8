// f32 geometric predicates of the collision-query path, usable from host (quadtree maintenance)
// and device (query kernels). Every function states the jagua-rs function whose arithmetic it must
// reproduce bit for bit (paths relative to /root/reference/jagua-rs/src).
//
// Numeric contract: Rust never contracts a*b+c, so this header must be compiled with
//   device: -fmad=false -prec-div=true -prec-sqrt=true -ftz=false
//   host:   -ffp-contract=off -fno-fast-math
// f32::min / f32::max are IEEE minNum/maxNum = fminf/fmaxf (inputs are never NaN here).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

#if defined(__CUDACC__)
#define JG_HD __host__ __device__ __forceinline__
#else
#define JG_HD inline
#endif

namespace jg {

struct Rect {
    float x_min, y_min, x_max, y_max;
};
struct Seg {  // Edge: start -> end
    float sx, sy, ex, ey;
};
struct Circ {
    float cx, cy, r;
};
// Query transformation as it crosses the boundary: (cos, sin, tx, ty) of
// geometry/transformation.rs:131-139 rot_transl_m = [[cos,-sin,tx],[sin,cos,ty],[0,0,1]].
struct Xf {
    float c, s, tx, ty;
};

enum : int { REL_INTERSECTING = 0, REL_ENCLOSED = 1, REL_SURROUNDING = 2, REL_DISJOINT = 3 };

JG_HD uint32_t f32_bits(float f) {
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    uint32_t u;
    std::memcpy(&u, &f, 4);
    return u;
#endif
}

// geometry/primitives/point.rs:26-31 TRANSFORM_FORMULA: (m00*x + m01*y) + m02*1.0, three roundings.
JG_HD void xf_point(const Xf& t, float x, float y, float& ox, float& oy) {
    const float m01 = -t.s;  // exact negation
    ox = (t.c * x + m01 * y) + t.tx * 1.0f;
    oy = (t.s * x + t.c * y) + t.ty * 1.0f;
}

// util/fpa.rs:18-22 over float-cmp 0.10 approx_eq!(f32) with the default F32Margin
// {epsilon: f32::EPSILON, ulps: 4} (third-party crate; restated from its published algorithm).
JG_HD bool approx_eq(float a, float b) {
    if (a == b) return true;
    const float diff = fabsf(a - b);
    // Far apart: neither margin can hold. Exact, not a heuristic: the epsilon test needs diff <= 1.2e-7, and four ulps of
    // same-signed values are at most 4 * 2^-23 = 4.8e-7 of the larger magnitude (opposite signs never pass the ulps test:
    // their i32 views differ by more than 2^31 - 4). NaN and infinite differences fall through to the full test.
    if (diff > 1e-6f * fmaxf(fabsf(a), fabsf(b)) + 1e-6f) return false;
    if (diff <= 1.1920929e-07f) return true;
    int32_t d = (int32_t)(f32_bits(a) - f32_bits(b));  // wrapping_sub on the i32 views
    int32_t ad = d < 0 ? (d == INT32_MIN ? INT32_MAX : -d) : d;
    return ad <= 4;
}
// util/fpa.rs:24-31: Equal when approx-equal, otherwise the exact ordering. approx_eq is a pure function, so the cheap
// exact comparison goes first (the same boolean, most calls never reach the ulp arithmetic).
JG_HD bool fpa_le(float a, float b) { return a <= b || approx_eq(a, b); }
JG_HD bool fpa_ge(float a, float b) { return a >= b || approx_eq(a, b); }
JG_HD bool fpa_lt(float a, float b) { return a < b && !approx_eq(a, b); }
JG_HD bool fpa_gt(float a, float b) { return a > b && !approx_eq(a, b); }

// primitives/rect.rs:266-272
JG_HD bool rect_rect(const Rect& a, const Rect& b) {
    return fmaxf(a.x_min, b.x_min) <= fminf(a.x_max, b.x_max) &&
           fmaxf(a.y_min, b.y_min) <= fminf(a.y_max, b.y_max);
}
// primitives/rect.rs:46-65 relation_to
JG_HD int rect_relation(const Rect& a, const Rect& b) {
    if (!rect_rect(a, b)) return REL_DISJOINT;
    if (a.x_min <= b.x_min && a.y_min <= b.y_min && a.x_max >= b.x_max && a.y_max >= b.y_max)
        return REL_SURROUNDING;
    if (a.x_min >= b.x_min && a.y_min >= b.y_min && a.x_max <= b.x_max && a.y_max <= b.y_max)
        return REL_ENCLOSED;
    return REL_INTERSECTING;
}
// primitives/rect.rs:70-89 almost_relation_to (+ :274-280 almost_collides_with)
JG_HD int rect_almost_relation(const Rect& a, const Rect& b) {
    if (!(fpa_le(fmaxf(a.x_min, b.x_min), fminf(a.x_max, b.x_max)) &&
          fpa_le(fmaxf(a.y_min, b.y_min), fminf(a.y_max, b.y_max))))
        return REL_DISJOINT;
    if (fpa_le(a.x_min, b.x_min) && fpa_le(a.y_min, b.y_min) && fpa_ge(a.x_max, b.x_max) &&
        fpa_ge(a.y_max, b.y_max))
        return REL_SURROUNDING;
    if (fpa_ge(a.x_min, b.x_min) && fpa_ge(a.y_min, b.y_min) && fpa_le(a.x_max, b.x_max) &&
        fpa_le(a.y_max, b.y_max))
        return REL_ENCLOSED;
    return REL_INTERSECTING;
}
// primitives/rect.rs:282-288
JG_HD bool rect_point(const Rect& r, float x, float y) {
    return x >= r.x_min && x <= r.x_max && y >= r.y_min && y <= r.y_max;
}
// primitives/rect.rs:301-345 Rect ∩ Edge
JG_HD bool rect_seg(const Rect& r, const Seg& e) {
    const float ex_min = fminf(e.sx, e.ex), ex_max = fmaxf(e.sx, e.ex);
    const float ey_min = fminf(e.sy, e.ey), ey_max = fmaxf(e.sy, e.ey);
    const bool x_no = fmaxf(ex_min, r.x_min) > fminf(ex_max, r.x_max);
    const bool y_no = fmaxf(ey_min, r.y_min) > fminf(ey_max, r.y_max);
    if (x_no || y_no) return false;
    if (rect_point(r, e.sx, e.sy) || rect_point(r, e.ex, e.ey)) return true;
    const float dx = e.ex - e.sx, dy = e.ey - e.sy;
    // corners NE, NW, SW, SE (rect.rs:154-161)
    const float s0 = (r.x_max - e.sx) * dy - (r.y_max - e.sy) * dx;
    const float s1 = (r.x_min - e.sx) * dy - (r.y_max - e.sy) * dx;
    const float s2 = (r.x_min - e.sx) * dy - (r.y_min - e.sy) * dx;
    const float s3 = (r.x_max - e.sx) * dy - (r.y_min - e.sy) * dx;
    const bool all_pos = s0 > 0.0f && s1 > 0.0f && s2 > 0.0f && s3 > 0.0f;
    const bool all_neg = s0 < 0.0f && s1 < 0.0f && s2 < 0.0f && s3 < 0.0f;
    return !(all_pos || all_neg);
}

// fl(a / d) in [0, 1] for finite a and d != 0, WITHOUT the division whenever the answer is already
// determined. Exact, not approximate (round-to-nearest-even, denormals kept):
//   * |a| > |d|  ->  |a/d| >= 1/(1 - 2^-24) > 1 + 2^-24, which rounds to >= 1 + 2^-23: the quotient is
//     > 1 (or < -1); conversely a/d <= 1 rounds to <= 1 by monotonicity.
//   * a == 0 or equal signs, |a| <= |d|  ->  0 <= a/d <= 1, and -0 compares >= 0 like the reference's
//     `(0.0..=1.0).contains(&t)`.
//   * opposite signs  ->  negative, unless the quotient underflows to -0 (|a/d| <= 2^-150), which again
//     compares >= 0. That needs |a| < |d| * 1e-30 by a wide margin; only then is the division performed.
// The IEEE division itself is needed only on rare branches (and expands to ~20 instructions wherever it is
// inlined): on the device it lives out of line, which takes ~550 instructions out of the stage-2 kernel, whose
// speed is set by instruction fetch (3.6 K SASS instructions against a 32 KB L1.5 instruction cache).
#if defined(JG_DIVCOUNT) && defined(__CUDACC__)
static __device__ unsigned long long g_divcount[4];  // experiments (-DJG_DIVCOUNT): thread-level divisions per site class, per translation unit
#endif
#if defined(JG_DIVCOUNT) && defined(__CUDA_ARCH__)
#define JG_DIV_SITE(k) atomicAdd(&g_divcount[k], 1ull)
#else
#define JG_DIV_SITE(k) ((void)0)
#endif
#if defined(__CUDA_ARCH__)
static __device__ __noinline__ float div_rare(float a, float d) { return a / d; }
#else
static inline float div_rare(float a, float d) { return a / d; }
#endif
JG_HD bool unit_quotient(float a, float d) {
    const float aa = fabsf(a), ad = fabsf(d);
    if (!(ad <= 3.402823466e+38f)) {  // infinite / NaN divisor (coordinates overflowed): no shortcut
        JG_DIV_SITE(0);
        const float q = div_rare(a, d);
        return q >= 0.0f && q <= 1.0f;
    }
    if (!(aa <= ad)) return false;  // also NaN / infinite numerators
    if (a == 0.0f || (a < 0.0f) == (d < 0.0f)) return true;
    if (aa < ad * 1e-30f) {
        JG_DIV_SITE(1);
        const float q = div_rare(a, d);
        return q >= 0.0f && q <= 1.0f;
    }
    return false;
}

// primitives/edge.rs:180-228 edge_intersection(e1, e2, false): range reject with strict '>',
// then t,u by IEEE division, hit iff both in [0,1]; parallel (denominator == 0) is "no".
JG_HD bool seg_seg(const Seg& e1, const Seg& e2) {
    const float x1 = e1.sx, y1 = e1.sy, x2 = e1.ex, y2 = e1.ey;
    const float x3 = e2.sx, y3 = e2.sy, x4 = e2.ex, y4 = e2.ey;
    const bool x_no = fmaxf(fminf(x1, x2), fminf(x3, x4)) > fminf(fmaxf(x1, x2), fmaxf(x3, x4));
    const bool y_no = fmaxf(fminf(y1, y2), fminf(y3, y4)) > fminf(fmaxf(y1, y2), fmaxf(y3, y4));
    if (x_no || y_no) return false;
    const float t_nom = (x2 - x4) * (y4 - y3) - (y2 - y4) * (x4 - x3);
    const float denom = (x2 - x1) * (y4 - y3) - (y2 - y1) * (x4 - x3);  // t_denom == u_denom
    const float u_nom = (x2 - x4) * (y2 - y1) - (y2 - y4) * (x2 - x1);
    if (denom == 0.0f) return false;
    return unit_quotient(t_nom, denom) && unit_quotient(u_nom, denom);  // t, u in [0, 1]
}

// collision_detection/quadtree/qt_traits.rs:144-177 edge_intersection_half: same t,u without the
// range reject. Returns 0 = no hit, 1 = first half of e2 (u > 0.5), 2 = second half, 3 = halfway.
JG_HD int seg_seg_half(const Seg& e1, const Seg& e2) {
    const float x1 = e1.sx, y1 = e1.sy, x2 = e1.ex, y2 = e1.ey;
    const float x3 = e2.sx, y3 = e2.sy, x4 = e2.ex, y4 = e2.ey;
    const float t_nom = (x2 - x4) * (y4 - y3) - (y2 - y4) * (x4 - x3);
    const float denom = (x2 - x1) * (y4 - y3) - (y2 - y1) * (x4 - x3);
    const float u_nom = (x2 - x4) * (y2 - y1) - (y2 - y4) * (x2 - x1);
    if (denom == 0.0f) return 0;
    if (!(unit_quotient(t_nom, denom) && unit_quotient(u_nom, denom))) return 0;
    JG_DIV_SITE(2);
    const float u = div_rare(u_nom, denom);  // the half is decided on the rounded quotient, as in the reference
    return u > 0.5f ? 1 : (u < 0.5f ? 2 : 3);
}

// primitives/rect.rs:141-151 quadrants(): child i = from_diagonal_corners(corner_i, centroid),
// corners NE, NW, SW, SE; centroid = ((x_min+x_max)/2, (y_min+y_max)/2) (rect.rs:248-253).
JG_HD Rect rect_quadrant(const Rect& r, int i) {
    const float mx = (r.x_min + r.x_max) / 2.0f, my = (r.y_min + r.y_max) / 2.0f;
    const float cx = (i == 0 || i == 3) ? r.x_max : r.x_min;
    const float cy = (i == 0 || i == 1) ? r.y_max : r.y_min;
    return Rect{fminf(cx, mx), fminf(cy, my), fmaxf(cx, mx), fmaxf(cy, my)};
}

// collision_detection/quadtree/qt_traits.rs:19-106 Edge::collides_with_quadrants. Returns a 4-bit
// mask (bit i = quadrant i of r collides). NOT equivalent to four rect_seg calls: half_intersect
// (qt_traits.rs:110-140) overwrites already-rejected quadrants and runs in a fixed order.
JG_HD unsigned seg_quadrants(const Seg& e, const Rect& r) {
    const float ex_min = fminf(e.sx, e.ex), ex_max = fmaxf(e.sx, e.ex);
    const float ey_min = fminf(e.sy, e.ey), ey_max = fmaxf(e.sy, e.ey);
    unsigned yes = 0, undet = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const Rect q = rect_quadrant(r, i);
        const bool x_no = fmaxf(ex_min, q.x_min) > fminf(ex_max, q.x_max);
        const bool y_no = fmaxf(ey_min, q.y_min) > fminf(ey_max, q.y_max);
        if (x_no || y_no) continue;  // Some(false)
        if (rect_point(q, e.sx, e.sy) || rect_point(q, e.ex, e.ey)) yes |= 1u << i;  // Some(true)
        else undet |= 1u << i;                                                       // None
    }
    if (undet == 0) return yes;
    const float cx = (r.x_min + r.x_max) / 2.0f, cy = (r.y_min + r.y_max) / 2.0f;
    // rect.rs:164-184 sides(): top NE->NW, left NW->SW, bottom SW->SE, right SE->NE
    const Seg left{r.x_min, r.y_max, r.x_min, r.y_min};
    const Seg right{r.x_max, r.y_min, r.x_max, r.y_max};
    const Seg top{r.x_max, r.y_max, r.x_min, r.y_max};
    const Seg bottom{r.x_min, r.y_min, r.x_max, r.y_min};
    const Seg h_bisect{r.x_min, cy, r.x_max, cy};
    const Seg v_bisect{cx, r.y_min, cx, r.y_max};
    // (side, first-half quadrants, second-half quadrants), in the reference's call order
#define JG_HALF(side, fst, sec)                              \
    if (undet & ((fst) | (sec))) {                           \
        const int h = seg_seg_half(e, side);                 \
        if (h & 1) { yes |= (fst); undet &= ~(fst); }        \
        if (h & 2) { yes |= (sec); undet &= ~(sec); }        \
    }
    JG_HALF(left, 2u, 4u)
    JG_HALF(right, 8u, 1u)
    JG_HALF(top, 1u, 2u)
    JG_HALF(bottom, 4u, 8u)
    JG_HALF(h_bisect, 2u | 4u, 1u | 8u)
    JG_HALF(v_bisect, 4u | 8u, 1u | 2u)
#undef JG_HALF
    return yes;
}

// The same function with its two unrolled parts as loops: a fifth of the instructions (one seg_seg_half body instead of
// six), a few more executed. For kernels whose hot loop must stay inside the instruction cache and which verify paths
// rarely (k_collect_grid). Same operations on the same values in the same order: bit-identical masks.
JG_HD unsigned seg_quadrants_rolled(const Seg& e, const Rect& r) {
    const float ex_min = fminf(e.sx, e.ex), ex_max = fmaxf(e.sx, e.ex);
    const float ey_min = fminf(e.sy, e.ey), ey_max = fmaxf(e.sy, e.ey);
    unsigned yes = 0, undet = 0;
#pragma unroll 1
    for (int i = 0; i < 4; i++) {
        const Rect q = rect_quadrant(r, i);
        const bool x_no = fmaxf(ex_min, q.x_min) > fminf(ex_max, q.x_max);
        const bool y_no = fmaxf(ey_min, q.y_min) > fminf(ey_max, q.y_max);
        if (x_no || y_no) continue;
        if (rect_point(q, e.sx, e.sy) || rect_point(q, e.ex, e.ey)) yes |= 1u << i;
        else undet |= 1u << i;
    }
    if (undet == 0) return yes;
    const float cx = (r.x_min + r.x_max) / 2.0f, cy = (r.y_min + r.y_max) / 2.0f;
#pragma unroll 1
    for (int k = 0; k < 6; k++) {  // left, right, top, bottom, horizontal bisector, vertical bisector (qt_traits.rs:60-84)
        Seg side;
        unsigned fst, sec;
        switch (k) {
            case 0: side = Seg{r.x_min, r.y_max, r.x_min, r.y_min}; fst = 2u; sec = 4u; break;
            case 1: side = Seg{r.x_max, r.y_min, r.x_max, r.y_max}; fst = 8u; sec = 1u; break;
            case 2: side = Seg{r.x_max, r.y_max, r.x_min, r.y_max}; fst = 1u; sec = 2u; break;
            case 3: side = Seg{r.x_min, r.y_min, r.x_max, r.y_min}; fst = 4u; sec = 8u; break;
            case 4: side = Seg{r.x_min, cy, r.x_max, cy}; fst = 2u | 4u; sec = 1u | 8u; break;
            default: side = Seg{cx, r.y_min, cx, r.y_max}; fst = 4u | 8u; sec = 1u | 2u; break;
        }
        if (!(undet & (fst | sec))) continue;
        const int h = seg_seg_half(e, side);
        if (h & 1) { yes |= fst; undet &= ~fst; }
        if (h & 2) { yes |= sec; undet &= ~sec; }
    }
    return yes;
}

// primitives/circle.rs:115-128 Circle ∩ Rect
JG_HD bool circ_rect(const Circ& c, const Rect& r) {
    const float nx = fmaxf(r.x_min, fminf(c.cx, r.x_max));
    const float ny = fmaxf(r.y_min, fminf(c.cy, r.y_max));
    return (nx - c.cx) * (nx - c.cx) + (ny - c.cy) * (ny - c.cy) <= c.r * c.r;
}
// primitives/circle.rs:109-113 via edge.rs:63-87 closest_point_on_edge + :153-159 sq_distance_to.
// param = dot / len_sq is only divided out when it is used (0 < dot <= len_sq): dot <= 0 gives a
// quotient <= +0 whose three possible treatments (param < 0, or param = +-0 in the interpolation) all
// yield the start point; dot > len_sq > 0 gives a rounded quotient > 1 (see unit_quotient); len_sq == 0
// is the reference's param = -1.
JG_HD bool circ_seg(const Circ& c, const Seg& e) {
    const float a = c.cx - e.sx, b = c.cy - e.sy, cc = e.ex - e.sx, d = e.ey - e.sy;
    const float dot = a * cc + b * d;
    const float len_sq = cc * cc + d * d;
    float xx, yy;
    if (len_sq == 0.0f || !(dot > 0.0f)) { xx = e.sx; yy = e.sy; }
    else if (dot > len_sq) { xx = e.ex; yy = e.ey; }
    else {
        const float param = dot / len_sq;
        xx = e.sx + param * cc;
        yy = e.sy + param * d;
    }
    const float dx = c.cx - xx, dy = c.cy - yy;
    return dx * dx + dy * dy <= c.r * c.r;
}
// qt_traits.rs:10-13 default collides_with_quadrants (Circle): four Circle ∩ Rect tests
JG_HD unsigned circ_quadrants(const Circ& c, const Rect& r) {
    unsigned m = 0;
#pragma unroll
    for (int i = 0; i < 4; i++)
        if (circ_rect(c, rect_quadrant(r, i))) m |= 1u << i;
    return m;
}

// Overloads so traversal code can be written once per entity kind (QTQueryable).
JG_HD bool ent_rect(const Seg& e, const Rect& r) { return rect_seg(r, e); }  // edge.rs:172-177
JG_HD bool ent_rect(const Circ& c, const Rect& r) { return circ_rect(c, r); }
JG_HD bool ent_seg(const Seg& e, const Seg& h) { return seg_seg(e, h); }     // e1 = query, e2 = hazard
JG_HD bool ent_seg(const Circ& c, const Seg& h) { return circ_seg(c, h); }
JG_HD unsigned ent_quadrants(const Seg& e, const Rect& r) { return seg_quadrants(e, r); }
JG_HD unsigned ent_quadrants(const Circ& c, const Rect& r) { return circ_quadrants(c, r); }

// primitives/simple_polygon.rs:255-292 point-in-polygon by ray casting with the FPA vertex rule.
// verts: n points (x,y interleaved); an optional transformation is applied on the fly (used for the
// query item, whose transformed vertices are never materialised on the device).
template <bool XF>
JG_HD bool poly_point(const float* verts, int n, const Rect& bbox, float px, float py, const Xf& t) {
    if (!rect_point(bbox, px, py)) return false;
    const Seg ray{px, py, bbox.x_max + (bbox.x_max - bbox.x_min), py};
    int n_int = 0;
    float sx = verts[0], sy = verts[1];
    if (XF) xf_point(t, verts[0], verts[1], sx, sy);
    const float fx = sx, fy = sy;
    for (int i = 0; i < n; i++) {
        float ex, ey;
        if (i == n - 1) { ex = fx; ey = fy; }
        else if (XF) xf_point(t, verts[2 * i + 2], verts[2 * i + 3], ex, ey);
        else { ex = verts[2 * i + 2]; ey = verts[2 * i + 3]; }
        if ((approx_eq(sy, py) && fpa_gt(sx, px)) || (approx_eq(ey, py) && fpa_gt(ex, px))) {
            if (fpa_lt(sy, py) || fpa_lt(ey, py)) n_int++;
        } else if (seg_seg(ray, Seg{sx, sy, ex, ey})) {
            n_int++;
        }
        sx = ex; sy = ey;
    }
    return (n_int & 1) == 1;
}

}  // namespace jg
