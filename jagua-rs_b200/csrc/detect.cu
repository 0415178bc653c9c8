// detect pipeline (sm_100a): detect_surrogate_collision + detect_poly_collision for a batch made of up
// to JG_MAX_SEG segments (one segment = one item's queries). Two persistent kernels per batch:
//
//   k_s1_surr   surrogate stage (cd_engine.rs:204-223). A warp takes 32 consecutive queries; lane = query.
//               Poles (then piers) are resolved one entity index at a time, all 32 queries together.
//   k_s2_poly   polygon stage (cd_engine.rs:160-197) for the survivors of stage 1. A warp takes a GROUP of
//               up to 8 survivors; their transformed bboxes, virtual roots (cd_engine.rs:314-327) and the
//               flat list of all their edges are set up once, then the edges are resolved together; the
//               containment tests (cd_engine.rs:174-193) of the survivors without an edge hit follow.
//
// Both kernels run QTNode::collides (qt_node.rs:99-147) as a WAVEFRONT instead of a per-lane recursion:
// the unit of work is a pair (entity, node). Pairs live on two per-warp LIFO stacks in shared memory
// ("inner": the node has children and the quadrant test decides which are entered; "leaf": the node's
// Partial hazards are tested edge by edge). Every step pops up to 32 pairs of ONE kind, so all lanes run
// the same predicate, and pushes the surviving (entity, child) pairs with ballot-compacted stores. A
// collision verdict is an OR over the visited nodes, so the visiting order is free; pairs of an
// already-decided query are dropped when popped.
//
// Two exact shortcuts keep most entities off the stacks (proofs in DESIGN.md §4):
//   * seg_chain   an edge whose endpoints lie inside a node and strictly on one side of both bisectors gets
//                 exactly that quadrant from Edge::collides_with_quadrants (qt_traits.rs:19-106: the other
//                 three are range-rejected, the fourth holds an endpoint) -> 4 compares instead of the test.
//   * circ_chain  a circle whose centre is inside the node and further than r(1+1e-5) from both bisectors
//                 collides with exactly the centre's quadrant (circle.rs:115-128 in f32, margin >> rounding).
//   The children's hazard summary ("kind": none / partial / entire, per filter; k_node_kinds) is read from
//   the parent's pair, so an Entire child (qt_node.rs:108) decides the query without being pushed.
//
// Node arrays (meta, bbox, kind) are staged into shared memory once per CTA when they fit; entries, hazard
// edges and hazard shapes are read through L1. If a stack could overflow (never seen; the capacity follows
// the tree depth) the warp falls back to the register-only walk of kernels.cuh for that work item.
#include <cstdlib>
#include "kernels.cuh"
#include "launch.hpp"

namespace jg {

#define S_THREADS 256
#define S_WARPS (S_THREADS / 32)
#define LEAF_CAP 160  // a leaf step runs as soon as 32 pairs wait; one inner step adds at most 128
#define FULL 0xFFFFFFFFu
#define S2_MAX_GROUP 8
#define JG_SPLIT_MIN_QUERIES (5 << 17)  // launches from this size on (640 Ki; tools/split_probe.py, profiles/r2ai_split_probe.txt: the
                                       // cross-over moved down from 1.5 M once stage 2 ran convergent rounds) run the containment tests in k_s2_contain

enum : int { KIND_NONE = 0, KIND_INNER = 1, KIND_ENTIRE = 2, KIND_LEAF = 3 };
enum : int { CH_DEAD = 0, CH_HIT = 1, CH_LEAF = 2, CH_INNER = 3 };

// ---- per-filter node summary ------------------------------------------------------------------------
// kinds[i + 3] = what QTHazardVec::strongest (qt_hazard_vec.rs:60-63) yields at node i under the filter:
// none / Entire / Partial with children (KIND_INNER) / Partial without children (KIND_LEAF). The +3 offset
// makes the four kinds of a child block (first child = 1 + 4k) one aligned 32-bit word.
__global__ void k_node_kinds(const DevLayout L, const uint32_t* __restrict__ filt, uint8_t* __restrict__ kinds) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= L.n_nodes) return;
    const int4 m = L.node_meta[i];
    const int n_entire = m.z & 0xFFFF, n_partial = (m.z >> 16) & 0xFFFF;
    int kind = KIND_NONE;
    if (!filt) {
        kind = n_entire ? KIND_ENTIRE : (n_partial ? KIND_INNER : KIND_NONE);
    } else {
        for (int j = 0; j < n_entire && kind == KIND_NONE; j++)
            if (!filt_bit(filt, L.entry_meta[m.y + j].x)) kind = KIND_ENTIRE;
        for (int j = 0; j < n_partial && kind == KIND_NONE; j++)
            if (!filt_bit(filt, L.entry_meta[m.y + n_entire + j].x)) kind = KIND_INNER;
    }
    if (kind == KIND_INNER && m.x < 0) kind = KIND_LEAF;
    kinds[i + 3] = (uint8_t)kind;
}

// ---- node arrays: shared memory when they fit, else the HBM copies ----------------------------------
struct NodeView {
    const int4* meta;
    const float4* bbox;
    const uint8_t* kind;  // kind[i + 3]
    __device__ __forceinline__ int kind_of(int node) const { return kind[node + 3]; }
    __device__ __forceinline__ uint32_t child_kinds(int child_base) const { return *(const uint32_t*)(kind + child_base + 3); }
};

__host__ __device__ inline size_t kinds_bytes(int n_nodes) { return ((size_t)n_nodes + 3 + 15) & ~(size_t)15; }
__host__ __device__ inline size_t node_smem_bytes(int n_nodes) { return (size_t)n_nodes * 32 + kinds_bytes(n_nodes); }

// carves [SegTable][node arrays] from the front of dynamic shared memory and fills them (whole CTA);
// returns the first free byte (16-byte aligned)
__device__ __forceinline__ unsigned char* stage_common(unsigned char* smem, const SegTable& gtab, const DevLayout& L,
                                                       const uint8_t* g_kinds, int nodes_in_smem, const SegTable*& tab,
                                                       NodeView& nv) {
    SegTable* t = (SegTable*)smem;
    for (int i = threadIdx.x; i < (int)(sizeof(SegTable) / 4); i += blockDim.x) ((int*)t)[i] = ((const int*)&gtab)[i];
    tab = t;
    unsigned char* p = smem + ((sizeof(SegTable) + 15) & ~(size_t)15);
    if (nodes_in_smem) {
        const int n = L.n_nodes;
        int4* s_meta = (int4*)p;
        float4* s_bbox = (float4*)(p + (size_t)n * 16);
        uint4* s_kind = (uint4*)(p + (size_t)n * 32);
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            s_meta[i] = __ldg(&L.node_meta[i]);
            s_bbox[i] = __ldg(&L.node_bbox[i]);
        }
        const int kw = (int)(kinds_bytes(n) / 16);
        for (int i = threadIdx.x; i < kw; i += blockDim.x) s_kind[i] = __ldg((const uint4*)g_kinds + i);
        nv.meta = s_meta;
        nv.bbox = s_bbox;
        nv.kind = (const uint8_t*)s_kind;
        p += node_smem_bytes(n);
    } else {
        nv.meta = L.node_meta;
        nv.bbox = L.node_bbox;
        nv.kind = g_kinds;
    }
    __syncthreads();
    return p;
}

// ---- warp stacks ----------------------------------------------------------------------------------
__device__ __forceinline__ unsigned lanemask_lt() {
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}
// every lane may push one value; order among lanes is kept
__device__ __forceinline__ void wpush(uint32_t* buf, int& n, bool want, uint32_t val) {
    const unsigned b = __ballot_sync(FULL, want);
    if (want) buf[n + __popc(b & lanemask_lt())] = val;
    n += __popc(b);
}

// ---- the exact shortcuts ----------------------------------------------------------------------------
// Descends from `node` (kind `kind`) while Edge::collides_with_quadrants provably flags exactly one child.
// Precondition: both endpoints inside bbox(node), inclusive. Returns CH_*; `node` = where it stopped.
__device__ __forceinline__ int seg_chain(const NodeView& nv, int& node, int kind, const Seg& s) {
    for (;;) {
        if (kind == KIND_ENTIRE) return CH_HIT;   // qt_node.rs:108
        if (kind == KIND_NONE) return CH_DEAD;    // qt_node.rs:105
        if (kind == KIND_LEAF) return CH_LEAF;    // qt_node.rs:128-143
        const float4 b = nv.bbox[node];
        const float mx = (b.x + b.z) / 2.0f, my = (b.y + b.w) / 2.0f;  // rect.rs:248-253 centroid
        const bool xr = s.sx > mx && s.ex > mx, xl = s.sx < mx && s.ex < mx;
        const bool yu = s.sy > my && s.ey > my, yd = s.sy < my && s.ey < my;
        if (!((xr || xl) && (yu || yd))) return CH_INNER;
        const int q = yu ? (xr ? 0 : 1) : (xl ? 2 : 3);  // NE, NW, SW, SE (rect.rs:141-161)
        node = nv.meta[node].x + q;
        kind = nv.kind_of(node);
    }
}
// Same for a circle, from the root. The margin 1e-5 covers the <= 3 roundings of circle.rs:115-128.
__device__ __forceinline__ int circ_chain(const NodeView& nv, int& node, const Circ& c) {
    node = 0;
    int kind = nv.kind_of(0);
    const float rk = c.r * 1.00001f;
    bool ok = c.r > 1e-15f && c.r < 1e15f;
    {
        const float4 b = nv.bbox[0];
        ok = ok && c.cx >= b.x && c.cx <= b.z && c.cy >= b.y && c.cy <= b.w;
    }
    for (;;) {
        if (kind == KIND_ENTIRE) return CH_HIT;
        if (kind == KIND_NONE) return CH_DEAD;
        if (kind == KIND_LEAF) return CH_LEAF;
        const float4 b = nv.bbox[node];
        const float mx = (b.x + b.z) / 2.0f, my = (b.y + b.w) / 2.0f;
        const float dx = c.cx - mx, dy = c.cy - my;
        if (!(ok && fabsf(dx) > rk && fabsf(dy) > rk)) return CH_INNER;
        const int q = dy > 0.0f ? (dx > 0.0f ? 0 : 1) : (dx < 0.0f ? 2 : 3);
        node = nv.meta[node].x + q;
        kind = nv.kind_of(node);
    }
}

// QTHazPartial::collides_with (qt_partial_hazard.rs:67-72) over the relevant Partial entries of a node
// without children (qt_node.rs:128-143)
template <class Ent, bool FILT>
__device__ __forceinline__ bool leaf_collides(const DevLayout& L, const int4 m, const Ent& ent, const uint32_t* filt) {
    const int n_entire = m.z & 0xFFFF, n_partial = (m.z >> 16) & 0xFFFF;
    const int e0 = m.y + n_entire;
    for (int i = 0; i < n_partial; i++) {
        const int4 em = __ldg(&L.entry_meta[e0 + i]);
        if (FILT && filt_bit(filt, em.x)) continue;
        if (!ent_rect(ent, ld_rect(&L.entry_ff[e0 + i]))) continue;
        for (int k = 0; k < em.z; k++)
            if (ent_seg(ent, ld_seg(&L.edges[em.y + k]))) return true;
    }
    return false;
}

// visit of an inner pair: which children are entered (pin), which are leaves to test (plf), or a hit
template <class Ent>
__device__ __forceinline__ bool inner_visit(const NodeView& nv, int node, const Ent& ent, int& child_base, unsigned& pin,
                                            unsigned& plf) {
    const unsigned mask = ent_quadrants(ent, ld_rect_any(&nv.bbox[node]));
    child_base = nv.meta[node].x;
    const uint32_t ck = nv.child_kinds(child_base);
    bool hit = false;
#pragma unroll
    for (int q = 0; q < 4; q++) {
        if (!((mask >> q) & 1u)) continue;
        const int kq = (ck >> (8 * q)) & 0xFF;
        hit = hit || kq == KIND_ENTIRE;
        if (kq == KIND_INNER) pin |= 1u << q;
        if (kq == KIND_LEAF) plf |= 1u << q;
    }
    return hit;
}

__device__ __forceinline__ int seg_of_query(const SegTable& tab, int q) {
    int k = 0;
    for (int j = 1; j < tab.n_seg; j++) k += (q >= tab.seg[j].begin) ? 1 : 0;
    return k;
}
// survivor entries: query index in the low 26 bits, segment in the high 6
__device__ __forceinline__ int sv_pack(int q, int k) { return q | (k << 26); }

// edge e of the item under xf (simple_polygon.rs:86-93 edge_iter over the transformed vertices)
__device__ __forceinline__ Seg item_edge(const SegDesc& sd, const Xf& xf, int e) {
    const float2 a = __ldg(&sd.verts[e]);
    const float2 b = __ldg(&sd.verts[e == sd.n_verts - 1 ? 0 : e + 1]);
    Seg sg;
    xf_point(xf, a.x, a.y, sg.sx, sg.sy);
    xf_point(xf, b.x, b.y, sg.ex, sg.ey);
    return sg;
}

// ---------------------------------------------------------------------------------------------------
// stage 1: surrogate
// ---------------------------------------------------------------------------------------------------
template <bool FILT>
__global__ void __launch_bounds__(S_THREADS)
k_s1_surr(const DevLayout L, const __grid_constant__ SegTable gtab, const uint8_t* __restrict__ g_kinds, int nodes_in_smem,
          int cap_in, const XfSrc xforms, const uint32_t* __restrict__ filt, uint8_t* __restrict__ out,
          DetectCtl* __restrict__ ctl, int* __restrict__ sv_entry) {
    extern __shared__ __align__(16) unsigned char smem[];
    const SegTable* tabp;
    NodeView nv;
    unsigned char* p = stage_common(smem, gtab, L, g_kinds, nodes_in_smem, tabp, nv);
    const SegTable& tab = *tabp;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t per_warp = 32 * sizeof(float4) + ((size_t)cap_in + LEAF_CAP + 4) * sizeof(uint32_t);
    unsigned char* wp = p + (size_t)warp * per_warp;
    float4* s_ent = (float4*)wp;                      // entity of each lane's query for the current round
    uint32_t* s_inner = (uint32_t*)(wp + 32 * sizeof(float4));
    uint32_t* s_leaf = s_inner + cap_in;
    unsigned* s_hit = (unsigned*)(s_leaf + LEAF_CAP);  // bit = lane whose query is decided
    const int n = tab.n_queries;

    for (;;) {
        int base = 0;
        if (lane == 0) base = atomicAdd(&ctl->s1_next, 32);
        base = __shfl_sync(FULL, base, 0);
        if (base >= n) break;
        const int q = base + lane;
        const bool valid = q < n;
        Xf xf{1.0f, 0.0f, 0.0f, 0.0f};
        int k = 0, n_poles = 0, n_ent = 0;
        if (valid) {
            k = seg_of_query(tab, q);
            const float4 t4 = load_xf(xforms, tab.seg[k], q);
            xf = Xf{t4.x, t4.y, t4.z, t4.w};
            n_poles = tab.seg[k].n_poles;
            n_ent = n_poles + tab.seg[k].n_piers;
        }
        bool hit = false;
        const int max_ent = __reduce_max_sync(FULL, n_ent);
        for (int j = 0; j < max_ent; j++) {
            // cd_engine.rs:210-221: poles in order, then piers; the first colliding one decides
            const bool act = valid && !hit && j < n_ent;
            if (!__any_sync(FULL, act)) break;
            const bool is_seg = act && j >= n_poles;
            const unsigned segmask = __ballot_sync(FULL, is_seg);
            int code = CH_DEAD, node = 0;
            if (act) {
                const SegDesc& sd = tab.seg[k];
                if (!is_seg) {  // circle.rs:79-85: centre transformed, radius unchanged
                    const float4 pl = __ldg(&sd.poles[j]);
                    Circ c;
                    xf_point(xf, pl.x, pl.y, c.cx, c.cy);
                    c.r = pl.z;
                    s_ent[lane] = make_float4(c.cx, c.cy, c.r, 0.0f);
                    code = circ_chain(nv, node, c);
                } else {  // edge.rs:126-134
                    const float4 pr = __ldg(&sd.piers[j - n_poles]);
                    Seg s;
                    xf_point(xf, pr.x, pr.y, s.sx, s.sy);
                    xf_point(xf, pr.z, pr.w, s.ex, s.ey);
                    s_ent[lane] = make_float4(s.sx, s.sy, s.ex, s.ey);
                    const int kr = nv.kind_of(0);  // no shortcut: a pier may leave the root's bbox
                    code = kr == KIND_ENTIRE ? CH_HIT : kr == KIND_INNER ? CH_INNER : kr == KIND_LEAF ? CH_LEAF : CH_DEAD;
                }
            }
            hit = hit || code == CH_HIT;
            if (lane == 0) *s_hit = 0u;
            int n_in = 0, n_lf = 0;
            wpush(s_inner, n_in, code == CH_INNER, ((uint32_t)node << 5) | (uint32_t)lane);
            wpush(s_leaf, n_lf, code == CH_LEAF, ((uint32_t)node << 5) | (uint32_t)lane);
            bool overflow = false;
            for (;;) {
                __syncwarp();
                const unsigned hm = *(volatile unsigned*)s_hit;
                bool lanehit = false;
                int slot = 0;
                if (n_lf >= 32 || (n_in == 0 && n_lf > 0)) {
                    const int cnt = n_lf < 32 ? n_lf : 32;
                    n_lf -= cnt;
                    if (lane < cnt) {
                        const uint32_t pr = s_leaf[n_lf + lane];
                        slot = pr & 31;
                        if (!((hm >> slot) & 1u)) {
                            const int4 m = nv.meta[pr >> 5];
                            const float4 e4 = s_ent[slot];
                            lanehit = ((segmask >> slot) & 1u)
                                          ? leaf_collides<Seg, FILT>(L, m, Seg{e4.x, e4.y, e4.z, e4.w}, filt)
                                          : leaf_collides<Circ, FILT>(L, m, Circ{e4.x, e4.y, e4.z}, filt);
                        }
                    }
                } else if (n_in > 0) {
                    const int cnt = n_in < 32 ? n_in : 32;
                    if (n_in + 3 * cnt > cap_in) { overflow = true; break; }
                    n_in -= cnt;
                    unsigned pin = 0, plf = 0;
                    int cb = 0;
                    if (lane < cnt) {
                        const uint32_t pr = s_inner[n_in + lane];
                        slot = pr & 31;
                        if (!((hm >> slot) & 1u)) {
                            const float4 e4 = s_ent[slot];
                            lanehit = ((segmask >> slot) & 1u)
                                          ? inner_visit<Seg>(nv, pr >> 5, Seg{e4.x, e4.y, e4.z, e4.w}, cb, pin, plf)
                                          : inner_visit<Circ>(nv, pr >> 5, Circ{e4.x, e4.y, e4.z}, cb, pin, plf);
                            if (lanehit) pin = plf = 0;
                        }
                    }
                    __syncwarp();  // all pops done before the pushes overwrite the popped range
#pragma unroll
                    for (int qd = 0; qd < 4; qd++) {
                        const uint32_t v = ((uint32_t)(cb + qd) << 5) | (uint32_t)slot;
                        wpush(s_inner, n_in, (pin >> qd) & 1u, v);
                        wpush(s_leaf, n_lf, (plf >> qd) & 1u, v);
                    }
                } else {
                    break;
                }
                if (__any_sync(FULL, lanehit)) {
                    if (lanehit) atomicOr(s_hit, 1u << slot);
                }
            }
            __syncwarp();
            if (overflow) {  // register-only walk of this round's entity (kernels.cuh), never seen in practice
                if (lane == 0) atomicAdd(&ctl->fallbacks, 1);
                if (act && !hit) {
                    const float4 e4 = s_ent[lane];
                    hit = is_seg ? qt_collides<Seg, FILT>(L, 0, Seg{e4.x, e4.y, e4.z, e4.w}, filt)
                                 : qt_collides<Circ, FILT>(L, 0, Circ{e4.x, e4.y, e4.z}, filt);
                }
            } else {
                hit = hit || ((*(volatile unsigned*)s_hit >> lane) & 1u);
            }
            __syncwarp();
        }
        if (valid) out[q] = hit ? (uint8_t)JG_VERDICT_SURROGATE : (uint8_t)0;
        // survivors keep the query order inside the warp's 32
        const unsigned sm = __ballot_sync(FULL, valid && !hit);
        if (sm) {
            int sb = 0;
            if (lane == 0) sb = atomicAdd(&ctl->n_surv, __popc(sm));
            sb = __shfl_sync(FULL, sb, 0);
            if (valid && !hit) sv_entry[sb + __popc(sm & lanemask_lt())] = sv_pack(q, k);
        }
    }
}

// JG_MODE_POLY_ONLY: no surrogate stage, every query is a survivor (lbf/benches/quadtree_bench.rs:73-124)
__global__ void __launch_bounds__(S_THREADS)
k_s1_all_survive(const __grid_constant__ SegTable tab, DetectCtl* __restrict__ ctl, int* __restrict__ sv_entry) {
    const int n = tab.n_queries;
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < n; q += gridDim.x * blockDim.x) {
        int k = 0;
        for (int j = 1; j < tab.n_seg; j++) k += (q >= tab.seg[j].begin) ? 1 : 0;
        sv_entry[q] = sv_pack(q, k);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) ctl->n_surv = n;
}

// ---------------------------------------------------------------------------------------------------
// stage 2: polygon
// ---------------------------------------------------------------------------------------------------
struct S2Slots {  // per-warp state of the survivors of the current group
    float4 xf[S2_MAX_GROUP];
    float4 bb[S2_MAX_GROUP];
    int q[S2_MAX_GROUP];
    int seg[S2_MAX_GROUP];
    int vroot[S2_MAX_GROUP];
    int vk[S2_MAX_GROUP];
    int ebeg[S2_MAX_GROUP + 1];  // flat edge index ranges of the slots
    unsigned hit;                // bit = slot with a colliding edge
    int pad[2];
};

// CDEngine::get_virtual_root (cd_engine.rs:314-327) on the staged node arrays
__device__ __forceinline__ int virtual_root_nv(const NodeView& nv, const Rect& bb) {
    int node = 0;
    for (;;) {
        const int cb = nv.meta[node].x;
        if (cb < 0) return node;
        int nxt = -1;
#pragma unroll
        for (int i = 0; i < 4; i++)
            if (nxt < 0 && rect_relation(ld_rect_any(&nv.bbox[cb + i]), bb) == REL_SURROUNDING) nxt = cb + i;
        if (nxt < 0) return node;
        node = nxt;
    }
}

template <bool FILT>
__global__ void __launch_bounds__(S_THREADS)
k_s2_poly(const DevLayout L, const __grid_constant__ SegTable gtab, const uint8_t* __restrict__ g_kinds, int nodes_in_smem,
          int cap_in, int gbits, const XfSrc xforms, const uint32_t* __restrict__ filt,
          uint8_t* __restrict__ out, DetectCtl* __restrict__ ctl, const int* __restrict__ sv_entry) {
    extern __shared__ __align__(16) unsigned char smem[];
    const SegTable* tabp;
    NodeView nv;
    unsigned char* p = stage_common(smem, gtab, L, g_kinds, nodes_in_smem, tabp, nv);
    const SegTable& tab = *tabp;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t per_warp = sizeof(S2Slots) + ((size_t)cap_in + LEAF_CAP) * sizeof(uint32_t);
    unsigned char* wp = p + (size_t)warp * per_warp;
    S2Slots& W = *(S2Slots*)wp;
    uint32_t* s_inner = (uint32_t*)(wp + sizeof(S2Slots));
    uint32_t* s_leaf = s_inner + cap_in;
    const int G = 1 << gbits;
    const int ebits = 12 + gbits;  // pair = node << ebits | edge << gbits | slot
    const uint32_t emask = (1u << ebits) - 1u;
    const int total = ctl->n_surv;

    for (;;) {
        int gbase = 0;
        if (lane == 0) gbase = atomicAdd(&ctl->s2_next, G);
        gbase = __shfl_sync(FULL, gbase, 0);
        if (gbase >= total) break;
        const int ng = (total - gbase) < G ? (total - gbase) : G;
        if (lane < ng) {
            const int ent = sv_entry[gbase + lane];
            const int q = ent & 0x3FFFFFF;
            W.q[lane] = q;
            W.seg[lane] = (int)((unsigned)ent >> 26);
            W.xf[lane] = load_xf(xforms, tab.seg[(unsigned)ent >> 26], q);
        }
        __syncwarp();
        // SPolygon::transform_from + generate_bounding_box (simple_polygon.rs:226-253, 122-133): min/max folds
        // are exact in any order, so lanes fold strided vertices and a butterfly joins them
        for (int s = 0; s < ng; s++) {
            const SegDesc& sd = tab.seg[W.seg[s]];
            const float4 t4 = W.xf[s];
            const Xf xf{t4.x, t4.y, t4.z, t4.w};
            float x0 = 3.402823466e+38f, y0 = 3.402823466e+38f, x1 = -3.402823466e+38f, y1 = -3.402823466e+38f;
            for (int j = lane; j < sd.n_verts; j += 32) {
                const float2 v = __ldg(&sd.verts[j]);
                float x, y;
                xf_point(xf, v.x, v.y, x, y);
                x0 = fminf(x0, x), y0 = fminf(y0, y), x1 = fmaxf(x1, x), y1 = fmaxf(y1, y);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                x0 = fminf(x0, __shfl_xor_sync(FULL, x0, o));
                y0 = fminf(y0, __shfl_xor_sync(FULL, y0, o));
                x1 = fmaxf(x1, __shfl_xor_sync(FULL, x1, o));
                y1 = fmaxf(y1, __shfl_xor_sync(FULL, y1, o));
            }
            if (lane == 0) W.bb[s] = make_float4(x0, y0, x1, y1);
        }
        __syncwarp();
        // lane s classifies survivor s
        int verdict = -1, nE = 0;
        if (lane < ng) {
            const float4 b4 = W.bb[lane];
            const Rect bb{b4.x, b4.y, b4.z, b4.w};
            if (rect_relation(L.cde_bbox, bb) != REL_SURROUNDING) {
                verdict = JG_VERDICT_POLY;  // cd_engine.rs:161-163
            } else {
                const int vroot = virtual_root_nv(nv, bb);
                const int vk = nv.kind_of(vroot);
                W.vroot[lane] = vroot;
                W.vk[lane] = vk;
                if (vk == KIND_ENTIRE) verdict = JG_VERDICT_POLY;  // every edge sees it (qt_node.rs:108)
                else if (vk != KIND_NONE) nE = tab.seg[W.seg[lane]].n_verts;
            }
        }
        int incl = nE;
#pragma unroll
        for (int o = 1; o < S2_MAX_GROUP; o <<= 1) {
            const int t = __shfl_up_sync(FULL, incl, o);
            if (lane >= o) incl += t;
        }
        const int E = __shfl_sync(FULL, incl, S2_MAX_GROUP - 1);
        if (lane < S2_MAX_GROUP) W.ebeg[lane + 1] = incl;
        if (lane == 0) {
            W.ebeg[0] = 0;
            W.hit = 0u;
        }
        const unsigned need = __ballot_sync(FULL, nE > 0);

        int n_in = 0, n_lf = 0, f_next = 0;
        bool overflow = false;
        for (;;) {
            __syncwarp();
            const unsigned hm = *(volatile unsigned*)&W.hit;
            if ((hm & need) == need) break;  // every survivor with edges has a colliding one
            bool lanehit = false;
            int slot = 0;
            if (n_lf >= 32 || (n_lf > 0 && n_in == 0 && f_next >= E)) {
                const int cnt = n_lf < 32 ? n_lf : 32;
                n_lf -= cnt;
                if (lane < cnt) {
                    const uint32_t pr = s_leaf[n_lf + lane];
                    slot = pr & (G - 1);
                    if (!((hm >> slot) & 1u)) {
                        const float4 t4 = W.xf[slot];
                        const Seg sg = item_edge(tab.seg[W.seg[slot]], Xf{t4.x, t4.y, t4.z, t4.w}, (pr & emask) >> gbits);
                        lanehit = leaf_collides<Seg, FILT>(L, nv.meta[pr >> ebits], sg, filt);
                    }
                }
            } else if (n_in < 32 && f_next < E) {
                // next 32 edges of the group: shortcut descent from the virtual root
                const int f = f_next + lane;
                f_next += 32;
                int code = CH_DEAD, node = 0;
                uint32_t ent = 0;
                if (f < E) {
                    for (int j = 1; j < G; j++) slot += (W.ebeg[j] <= f) ? 1 : 0;
                    if (!((hm >> slot) & 1u)) {
                        const int e = f - W.ebeg[slot];
                        const float4 t4 = W.xf[slot];
                        const Seg sg = item_edge(tab.seg[W.seg[slot]], Xf{t4.x, t4.y, t4.z, t4.w}, e);
                        node = W.vroot[slot];
                        code = seg_chain(nv, node, W.vk[slot], sg);
                        ent = ((uint32_t)e << gbits) | (uint32_t)slot;
                        lanehit = code == CH_HIT;
                    }
                }
                wpush(s_inner, n_in, code == CH_INNER, ((uint32_t)node << ebits) | ent);
                wpush(s_leaf, n_lf, code == CH_LEAF, ((uint32_t)node << ebits) | ent);
            } else if (n_in > 0) {
                const int cnt = n_in < 32 ? n_in : 32;
                if (n_in + 3 * cnt > cap_in) { overflow = true; break; }
                n_in -= cnt;
                unsigned pin = 0, plf = 0;
                int cb = 0;
                uint32_t ent = 0;
                if (lane < cnt) {
                    const uint32_t pr = s_inner[n_in + lane];
                    ent = pr & emask;
                    slot = ent & (G - 1);
                    if (!((hm >> slot) & 1u)) {
                        const float4 t4 = W.xf[slot];
                        const Seg sg = item_edge(tab.seg[W.seg[slot]], Xf{t4.x, t4.y, t4.z, t4.w}, ent >> gbits);
                        lanehit = inner_visit<Seg>(nv, pr >> ebits, sg, cb, pin, plf);
                        if (lanehit) pin = plf = 0;
                    }
                }
                __syncwarp();  // all pops done before the pushes overwrite the popped range
#pragma unroll
                for (int qd = 0; qd < 4; qd++) {
                    const uint32_t v = ((uint32_t)(cb + qd) << ebits) | ent;
                    wpush(s_inner, n_in, (pin >> qd) & 1u, v);
                    wpush(s_leaf, n_lf, (plf >> qd) & 1u, v);
                }
            } else {
                break;
            }
            if (__any_sync(FULL, lanehit)) {
                if (lanehit) atomicOr(&W.hit, 1u << slot);
            }
        }
        __syncwarp();
        unsigned hm = *(volatile unsigned*)&W.hit;
        if (overflow) {  // register-only walks (kernels.cuh) for the undecided survivors; never seen in practice
            if (lane == 0) atomicAdd(&ctl->fallbacks, 1);
            for (int s = 0; s < ng; s++) {
                if (!((need >> s) & 1u) || ((hm >> s) & 1u)) continue;
                const SegDesc& sd = tab.seg[W.seg[s]];
                const float4 t4 = W.xf[s];
                bool h = false;
                for (int e = lane; e < sd.n_verts && !h; e += 32)
                    h = qt_collides<Seg, FILT>(L, W.vroot[s], item_edge(sd, Xf{t4.x, t4.y, t4.z, t4.w}, e), filt);
                if (__any_sync(FULL, h)) hm |= 1u << s;
            }
        }
        // cd_engine.rs:174-193: survivors without an edge hit -> `unreachable!` or containment tests
        for (int s = 0; s < ng; s++) {
            if (__shfl_sync(FULL, verdict, s) >= 0) continue;
            int v = 0;
            if ((hm >> s) & 1u) {
                v = JG_VERDICT_POLY;
            } else {
                const int4 m = nv.meta[W.vroot[s]];
                const int n_entire = m.z & 0xFFFF, n_partial = (m.z >> 16) & 0xFFFF;
                if (n_entire > 0) {
                    v = JG_VERDICT_REF_PANIC;  // cd_engine.rs:179, reachable only under a filter
                } else {
                    const SegDesc& sd = tab.seg[W.seg[s]];
                    const float4 t4 = W.xf[s], b4 = W.bb[s];
                    bool c = false;
                    for (int j = lane; j < n_partial && !c; j += 32) {
                        const int hs = __ldg(&L.entry_meta[m.y + j]).x;
                        if (FILT && filt_bit(filt, hs)) continue;
                        c = containment_collision(L, sd.verts, sd.n_verts, sd.poi_x, sd.poi_y, Xf{t4.x, t4.y, t4.z, t4.w},
                                                  Rect{b4.x, b4.y, b4.z, b4.w}, hs);
                    }
                    if (__any_sync(FULL, c)) v = JG_VERDICT_POLY;
                }
            }
            if (lane == s) verdict = v;
        }
        if (lane < ng) out[W.q[lane]] = (uint8_t)verdict;
        __syncwarp();
    }
}

}  // namespace jg
#include "detect_grid.cuh"
#include "collect_grid.cuh"
namespace jg {

// ---------------------------------------------------------------------------------------------------
// occupancy (and the opt-in to more than 48 KB of dynamic shared memory, a PER-DEVICE function attribute) per device and
// kernel slot: an in-process pool launches the same kernels on every GPU of the box
#define JG_OCC_DEVS 64
#define JG_OCC_SLOTS 32
static int g_occ[JG_OCC_DEVS][JG_OCC_SLOTS];
static size_t g_occ_smem[JG_OCC_DEVS][JG_OCC_SLOTS];
#ifdef JG_DIVCOUNT
extern "C" int jg_debug_divcount(unsigned long long out[4], int reset) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out, g_divcount, sizeof(unsigned long long) * 4);
    if (reset) { unsigned long long z[4] = {0, 0, 0, 0}; cudaMemcpyToSymbol(g_divcount, z, sizeof(z)); }
    return 0;
}
#endif
// jg_set_dependent_launch / JG_PDL: stage 2 of the grid path launched with programmatic stream serialization
static int g_dependent_launch = [] { const char* v = getenv("JG_PDL"); return v ? atoi(v) : 1; }();
void set_dependent_launch(int on) { g_dependent_launch = on ? 1 : 0; }
int dependent_launch() { return g_dependent_launch; }
template <class K>
static int persistent_grid(K kernel, int slot, int n_sms, size_t smem, int threads = S_THREADS) {
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= JG_OCC_DEVS - 1;
    if (g_occ[dev][slot] == 0 || g_occ_smem[dev][slot] != smem) {
        if (smem > 48 * 1024) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        int b = 1;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, kernel, threads, smem);
        g_occ_smem[dev][slot] = smem;
        g_occ[dev][slot] = b < 1 ? 1 : b;
    }
    static const int cap = [] { const char* v = getenv("JG_CTAS_PER_SM"); return v ? atoi(v) : 0; }();  // experiments
    const int b = (cap > 0 && cap < g_occ[dev][slot]) ? cap : g_occ[dev][slot];
    return n_sms * b;
}

// the per-(layout, filter) summary buffer: [kinds][GridDev]
size_t detect_kinds_bytes(int n_nodes) { return kinds_bytes(n_nodes) + ((sizeof(GridDev) + 15) & ~(size_t)15) + sizeof(GridCollect); }
static const GridCollect* cgrid_of(const uint8_t* d_kinds, int n_nodes) {
    return (const GridCollect*)(d_kinds + kinds_bytes(n_nodes) + ((sizeof(GridDev) + 15) & ~(size_t)15));
}
static const GridDev* grid_of(const uint8_t* d_kinds, int n_nodes) { return (const GridDev*)(d_kinds + kinds_bytes(n_nodes)); }

cudaError_t launch_node_kinds(const DevLayout& L, int depth, const uint32_t* d_filt, uint8_t* d_kinds, cudaStream_t st) {
    cudaMemsetAsync(d_kinds, 0, kinds_bytes(L.n_nodes), st);
    k_node_kinds<<<(L.n_nodes + 255) / 256, 256, 0, st>>>(L, d_filt, d_kinds);
    k_grid_build<<<1, 1024, 0, st>>>(L, depth, d_kinds, (GridDev*)(d_kinds + kinds_bytes(L.n_nodes)));
    if (!d_filt)  // the collect grid does not depend on a filter: built with the NoFilter summary only
        k_cgrid_build<<<1, 1024, 0, st>>>(L, grid_of(d_kinds, L.n_nodes), (GridCollect*)cgrid_of(d_kinds, L.n_nodes));
    return cudaGetLastError();
}
const int* detect_grid_ok_ptr(const uint8_t* d_kinds, int n_nodes) { return &grid_of(d_kinds, n_nodes)->ok; }
// 1 when the grid kernels can serve this layout (k_grid_build's verdict); synchronises the stream
cudaError_t detect_grid_usable(const uint8_t* d_kinds, int n_nodes, cudaStream_t st, int* ok) {
    *ok = 0;
    cudaError_t err = cudaMemcpyAsync(ok, &grid_of(d_kinds, n_nodes)->ok, sizeof(int), cudaMemcpyDeviceToHost, st);
    if (err != cudaSuccess) return err;
    return cudaStreamSynchronize(st);
}

cudaError_t launch_detect(const DevLayout& L, const DetectPlan& plan, const SegTable& tab, const XfSrc& d_xf,
                          const uint32_t* d_filt, const uint8_t* d_kinds, int mode, uint8_t* d_out, const DetectScratch& sc,
                          int n_sms, cudaStream_t st, int* n_launches) {
    const int n = tab.n_queries;
    cudaMemsetAsync(sc.ctl, 0, sizeof(DetectCtl), st);
    const size_t common = ((sizeof(SegTable) + 15) & ~(size_t)15) + (plan.nodes_in_smem ? node_smem_bytes(L.n_nodes) : 0);
    const size_t smem1 = common + S_WARPS * (32 * sizeof(float4) + ((size_t)plan.cap_in + LEAF_CAP + 4) * sizeof(uint32_t));
    const size_t smem2 = common + S_WARPS * (sizeof(S2Slots) + ((size_t)plan.cap_in + LEAF_CAP) * sizeof(uint32_t));
    const int want1 = (n + S_THREADS - 1) / S_THREADS;                               // 32 queries per warp batch
    const bool grid = plan.use_grid != 0;
    bool piers = false;  // items whose surrogate carries piers: k_s1_grid<.., true> tests them after the poles
    for (int k = 0; k < tab.n_seg; k++) piers = piers || tab.seg[k].n_piers != 0;
    if (grid) {
        const GridDev* gd = grid_of(d_kinds, L.n_nodes);
        const size_t gcommon = ((sizeof(SegTable) + 15) & ~(size_t)15) + ((sizeof(GridDev) + 15) & ~(size_t)15);
        const size_t gs1 = gcommon + S_WARPS * G1_WARP_BYTES;
        const size_t gs2 = gcommon + S_WARPS * sizeof(G2Slots);
        const int want2g = (n + S_WARPS * S2_MAX_GROUP - 1) / (S_WARPS * S2_MAX_GROUP);
        if (mode & JG_MODE_POLY_ONLY) {
            const int gs = want1 < n_sms * 8 ? want1 : n_sms * 8;
            k_s1_all_survive<<<gs, S_THREADS, 0, st>>>(tab, sc.ctl, sc.sv_entry);
        } else if (piers && d_filt) {
            const int g = persistent_grid(k_s1_grid<true, true>, 16, n_sms, gs1);
            k_s1_grid<true, true><<<want1 < g ? want1 : g, S_THREADS, gs1, st>>>(L, tab, gd, d_xf, d_filt, d_out, sc.ctl, sc.sv_entry);
        } else if (piers) {
            const int g = persistent_grid(k_s1_grid<false, true>, 17, n_sms, gs1);
            k_s1_grid<false, true><<<want1 < g ? want1 : g, S_THREADS, gs1, st>>>(L, tab, gd, d_xf, d_filt, d_out, sc.ctl, sc.sv_entry);
        } else if (d_filt) {
            const int g = persistent_grid(k_s1_grid<true>, 4, n_sms, gs1);
            k_s1_grid<true><<<want1 < g ? want1 : g, S_THREADS, gs1, st>>>(L, tab, gd, d_xf, d_filt, d_out, sc.ctl, sc.sv_entry);
        } else {
            const int g = persistent_grid(k_s1_grid<false>, 5, n_sms, gs1);
            k_s1_grid<false><<<want1 < g ? want1 : g, S_THREADS, gs1, st>>>(L, tab, gd, d_xf, d_filt, d_out, sc.ctl, sc.sv_entry);
        }
        // Large launches hand the containment tests to their own kernel (k_s2_contain): stage 2 is bound by instruction
        // fetch and runs faster without that code (and the team code of k_s2_contain needs 8.7 M warp instructions per 1M
        // queries where the inline, warp-cooperative form needs 22 M), which pays for the third launch (a dependent launch)
        // from 640 Ki queries on (1 Mi: 198.3 -> 190.3 us, 4 Mi: 619.6 -> 564.0 us, profiles/r2ai_split_probe.txt); smaller
        // launches keep the containment tests inline.
        const bool split = (long long)n >= plan.split_min && sc.ct_meta != nullptr;
        const size_t gs3 = ((sizeof(SegTable) + 15) & ~(size_t)15) + 16;
        const int want3 = (n + S_WARPS * CT_GROUP - 1) / (S_WARPS * CT_GROUP);
#define JG_LAUNCH_S2(FILT_, SPLIT_, SLOT_)                                                                                     \
    {                                                                                                                          \
        const int g = persistent_grid(k_s2_grid<FILT_, SPLIT_>, SLOT_, n_sms, gs2);                                            \
        if (g_dependent_launch && !(mode & JG_MODE_POLY_ONLY)) {                                                               \
            /* programmatic dependent launch: k_s1_grid releases its dependents at once (griddepcontrol.launch_dependents), */ \
            /* so stage-2 CTAs move in while stage 1 drains, stage their tables and wait in griddepcontrol.wait */             \
            cudaLaunchConfig_t lc = {};                                                                                        \
            lc.gridDim = dim3((unsigned)(want2g < g ? want2g : g));                                                            \
            lc.blockDim = dim3(S_THREADS);                                                                                     \
            lc.dynamicSmemBytes = gs2;                                                                                         \
            lc.stream = st;                                                                                                    \
            cudaLaunchAttribute at[1];                                                                                         \
            at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;                                                     \
            at[0].val.programmaticStreamSerializationAllowed = 1;                                                              \
            lc.attrs = at;                                                                                                     \
            lc.numAttrs = 1;                                                                                                   \
            cudaLaunchKernelEx(&lc, k_s2_grid<FILT_, SPLIT_>, L, tab, gd, d_kinds, d_xf, d_filt, d_out, sc.ctl, sc.sv_entry,   \
                               sc.ct_meta, sc.ct_bb, sc.ct_xf);                                                                \
        } else                                                                                                                 \
            k_s2_grid<FILT_, SPLIT_><<<want2g < g ? want2g : g, S_THREADS, gs2, st>>>(L, tab, gd, d_kinds, d_xf, d_filt, d_out, \
                                                                                     sc.ctl, sc.sv_entry, sc.ct_meta,          \
                                                                                     sc.ct_bb, sc.ct_xf);                      \
    }
        // BASELINE north_star item 3: the leaf entries + hazard edges of the layout staged into shared memory by TMA bulk copies
        // (cp.async.bulk + mbarrier), one 1024-thread CTA per SM so that the SM holds one copy. On whenever the sections fit
        // (swim: 100 KB); JG_STAGE_S2=0 restores the L1 path. Same-box A/B (profiles/r2k_size_sweep_stage*.txt): 64 Ki queries
        // 54.7 -> 52.6 us, 256 Ki 93.2 -> 89.0, 1 Mi 214 -> 209, 4 Mi 643 -> 635; stage 2 alone 137 -> 131 us, long-scoreboard
        // stalls 1.43 -> 1.00 per issue (profiles/r2j_ncu_key_metrics_s2_*.txt).
        static const int stage_env = [] { const char* v = getenv("JG_STAGE_S2"); return v ? atoi(v) : 1; }();
        const size_t stage_bytes = 16 + (size_t)L.n_entries * 32 + (size_t)L.n_edges * 16;
        const size_t gs2s = gcommon + stage_bytes + (S2_STAGE_THREADS / 32) * sizeof(G2Slots);
        if (stage_env && gs2s <= 220 * 1024) {
            const int want2s = (n + (S2_STAGE_THREADS / 32) * S2_MAX_GROUP - 1) / ((S2_STAGE_THREADS / 32) * S2_MAX_GROUP);
#define JG_LAUNCH_S2S(FILT_, SPLIT_, SLOT_)                                                                                    \
    {                                                                                                                          \
        const int g = persistent_grid(k_s2_grid<FILT_, SPLIT_, true>, SLOT_, n_sms, gs2s, S2_STAGE_THREADS);                   \
        cudaLaunchConfig_t lc = {};                                                                                            \
        lc.gridDim = dim3((unsigned)(want2s < g ? want2s : g));                                                                \
        lc.blockDim = dim3(S2_STAGE_THREADS);                                                                                  \
        lc.dynamicSmemBytes = gs2s;                                                                                            \
        lc.stream = st;                                                                                                        \
        cudaLaunchAttribute at[1];                                                                                             \
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;                                                         \
        at[0].val.programmaticStreamSerializationAllowed = 1;                                                                  \
        lc.attrs = at;                                                                                                         \
        lc.numAttrs = (g_dependent_launch && !(mode & JG_MODE_POLY_ONLY)) ? 1 : 0;                                             \
        cudaLaunchKernelEx(&lc, k_s2_grid<FILT_, SPLIT_, true>, L, tab, gd, d_kinds, d_xf, d_filt, d_out, sc.ctl, sc.sv_entry, \
                           sc.ct_meta, sc.ct_bb, sc.ct_xf);                                                                    \
    }
            if (d_filt && split) JG_LAUNCH_S2S(true, true, 21)
            else if (d_filt) JG_LAUNCH_S2S(true, false, 19)
            else if (split) JG_LAUNCH_S2S(false, true, 22)
            else JG_LAUNCH_S2S(false, false, 20)
#undef JG_LAUNCH_S2S
        } else if (d_filt && split) JG_LAUNCH_S2(true, true, 6)
        else if (d_filt) JG_LAUNCH_S2(true, false, 14)
        else if (split) JG_LAUNCH_S2(false, true, 7)
        else JG_LAUNCH_S2(false, false, 15)
#undef JG_LAUNCH_S2
#define JG_LAUNCH_S3(FILT_, SLOT_)                                                                                             \
    {                                                                                                                          \
        const int g3 = persistent_grid(k_s2_contain<FILT_>, SLOT_, n_sms, gs3);                                                \
        cudaLaunchConfig_t lc = {};                                                                                            \
        lc.gridDim = dim3((unsigned)(want3 < g3 ? want3 : g3));                                                                \
        lc.blockDim = dim3(S_THREADS);                                                                                         \
        lc.dynamicSmemBytes = gs3;                                                                                             \
        lc.stream = st;                                                                                                        \
        cudaLaunchAttribute at[1];                                                                                             \
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;                                                         \
        at[0].val.programmaticStreamSerializationAllowed = 1;                                                                  \
        lc.attrs = at;                                                                                                         \
        lc.numAttrs = g_dependent_launch ? 1 : 0;                                                                              \
        const int4* cm_ = sc.ct_meta;                                                                                          \
        const float4 *cb_ = sc.ct_bb, *cx_ = sc.ct_xf;                                                                         \
        cudaLaunchKernelEx(&lc, k_s2_contain<FILT_>, L, tab, d_filt, d_out, sc.ctl, cm_, cb_, cx_);                            \
    }
        if (split) {
            if (d_filt) JG_LAUNCH_S3(true, 12)
            else JG_LAUNCH_S3(false, 13)
        }
#undef JG_LAUNCH_S3
        if (n_launches) *n_launches += split ? 3 : 2;
        return cudaGetLastError();
    }
    const int want2 = (n + S_WARPS * (1 << plan.gbits) - 1) / (S_WARPS * (1 << plan.gbits));  // upper bound
    if (mode & JG_MODE_POLY_ONLY) {
        const int gs = want1 < n_sms * 8 ? want1 : n_sms * 8;
        k_s1_all_survive<<<gs, S_THREADS, 0, st>>>(tab, sc.ctl, sc.sv_entry);
    } else if (d_filt) {
        const int g = persistent_grid(k_s1_surr<true>, 0, n_sms, smem1);
        k_s1_surr<true><<<want1 < g ? want1 : g, S_THREADS, smem1, st>>>(L, tab, d_kinds, plan.nodes_in_smem, plan.cap_in, d_xf,
                                                                       d_filt, d_out, sc.ctl, sc.sv_entry);
    } else {
        const int g = persistent_grid(k_s1_surr<false>, 1, n_sms, smem1);
        k_s1_surr<false><<<want1 < g ? want1 : g, S_THREADS, smem1, st>>>(L, tab, d_kinds, plan.nodes_in_smem, plan.cap_in, d_xf,
                                                                        d_filt, d_out, sc.ctl, sc.sv_entry);
    }
    if (d_filt) {
        const int g = persistent_grid(k_s2_poly<true>, 2, n_sms, smem2);
        k_s2_poly<true><<<want2 < g ? want2 : g, S_THREADS, smem2, st>>>(L, tab, d_kinds, plan.nodes_in_smem, plan.cap_in,
                                                                       plan.gbits, d_xf, d_filt, d_out, sc.ctl, sc.sv_entry);
    } else {
        const int g = persistent_grid(k_s2_poly<false>, 3, n_sms, smem2);
        k_s2_poly<false><<<want2 < g ? want2 : g, S_THREADS, smem2, st>>>(L, tab, d_kinds, plan.nodes_in_smem, plan.cap_in,
                                                                        plan.gbits, d_xf, d_filt, d_out, sc.ctl, sc.sv_entry);
    }
    if (n_launches) *n_launches += 2;
    return cudaGetLastError();
}

// collect on the grid (k_collect_grid); d_kinds = the NoFilter summary buffer, d_next = a zeroed device int
bool collect_grid_applicable(const DetectPlan& plan, const SegTable& tab, int words) {
    if (!plan.use_grid || words > CG_MAX_WORDS) return false;
    for (int k = 0; k < tab.n_seg; k++)
        if (tab.seg[k].n_poles > CG_MAX_POLES || tab.seg[k].n_piers > CG_MAX_PIERS || tab.seg[k].n_verts > CG_VERT_CAP) return false;
    return true;
}
cudaError_t launch_collect_grid(const DevLayout& L, SegTable& tab, const uint8_t* d_kinds, const float4* d_xf, int mode,
                                const uint32_t* d_seed, uint32_t* d_out_bits, int words, int* d_next, int n_sms, cudaStream_t st,
                                int* n_launches) {
    cudaMemsetAsync(d_next, 0, 2 * sizeof(int), st);  // work counters of the two kernels (DetectCtl::s1_next, s2_next)
    bool piers = false;
    int max_v = 1;
    for (int k = 0; k < tab.n_seg; k++) {
        piers = piers || tab.seg[k].n_piers > 0;
        if (tab.seg[k].n_verts > max_v) max_v = tab.seg[k].n_verts;
    }
    const size_t smem = ((sizeof(SegTable) + 15) & ~(size_t)15) + ((sizeof(GridDev) + 15) & ~(size_t)15) +
                        CG_WARPS * (piers ? sizeof(CGWarp<true>) : sizeof(CGWarp<false>));
    const int g = piers ? persistent_grid(k_collect_grid<true>, 11, n_sms, smem, CG_THREADS) : persistent_grid(k_collect_grid<false>, 10, n_sms, smem, CG_THREADS);
    // queries per group: 8 when every warp has several groups to do; 4 / 2 for small launches (see the kernel)
    const long long per_warp = (long long)tab.n_queries / ((long long)g * CG_WARPS * CG_GROUP);
    int GS = per_warp >= 4 ? CG_GROUP : per_warp >= 1 ? CG_GROUP / 2 : CG_GROUP / 4;
    if (GS > CG_VERT_CAP / max_v) GS = CG_VERT_CAP / max_v;  // the group's transformed vertices live in the warp's shared-memory slice
    if (GS < 1) GS = 1;
    int gshift = 0;
    while ((1 << gshift) < GS) gshift++;
    int n_groups = 0;
    for (int k = 0; k < tab.n_seg; k++) {
        const int len = (k + 1 < tab.n_seg ? tab.seg[k + 1].begin : tab.n_queries) - tab.seg[k].begin;
        tab.seg[k].gbeg = n_groups;
        n_groups += (len + GS - 1) / GS;
    }
    const long long want = ((long long)n_groups + CG_WARPS - 1) / CG_WARPS;
    const int blocks = (int)(want < g ? want : g);
    if (piers)
        k_collect_grid<true><<<blocks, CG_THREADS, smem, st>>>(L, tab, grid_of(d_kinds, L.n_nodes), cgrid_of(d_kinds, L.n_nodes), d_xf, mode, d_seed,
                                                              d_out_bits, words, d_next, GS, gshift, n_groups);
    else
        k_collect_grid<false><<<blocks, CG_THREADS, smem, st>>>(L, tab, grid_of(d_kinds, L.n_nodes), cgrid_of(d_kinds, L.n_nodes), d_xf, mode, d_seed,
                                                               d_out_bits, words, d_next, GS, gshift, n_groups);
    {   // cd_engine.rs:274-288 in a kernel of its own (see k_collect_contain): groups of CC_GROUP queries, collectors completed in place
        int n2 = 0;
        for (int k = 0; k < tab.n_seg; k++) {
            const int len = (k + 1 < tab.n_seg ? tab.seg[k + 1].begin : tab.n_queries) - tab.seg[k].begin;
            tab.seg[k].gbeg = n2;
            n2 += (len + CC_GROUP - 1) / CC_GROUP;
        }
        const size_t smem2 = ((sizeof(SegTable) + 15) & ~(size_t)15) + ((sizeof(GridDev) + 15) & ~(size_t)15) + S_WARPS * sizeof(CCSlots);
        const int g2 = persistent_grid(k_collect_contain, 18, n_sms, smem2);
        const long long want2 = ((long long)n2 + S_WARPS - 1) / S_WARPS;
        k_collect_contain<<<(int)(want2 < g2 ? want2 : g2), S_THREADS, smem2, st>>>(L, tab, grid_of(d_kinds, L.n_nodes), d_xf, d_out_bits, words,
                                                                                 d_next + 1, n2);
        if (n_launches) *n_launches += 1;
    }
    if (n_launches) *n_launches += 1;
    return cudaGetLastError();
}

bool detect_fused_applicable(const DetectPlan& plan, const SegTable& tab, int mode) {
    if (!plan.use_grid || (mode & JG_MODE_POLY_ONLY)) return false;
    for (int k = 0; k < tab.n_seg; k++)
        if (tab.seg[k].n_piers != 0) return false;
    return true;
}

cudaError_t launch_detect_fused(const DevLayout& L, const SegTable& tab, const XfSrc& d_xf, const uint32_t* d_filt,
                                const uint8_t* d_kinds, uint8_t* d_out, const DetectScratch& sc, int n_sms, int avail_all,
                                cudaStream_t st, int* n_launches) {
    const GridDev* gd = grid_of(d_kinds, L.n_nodes);
    const size_t smem = ((sizeof(SegTable) + 15) & ~(size_t)15) + ((sizeof(GridDev) + 15) & ~(size_t)15) + S_WARPS * GF_WARP_BYTES;
    const int want = (tab.n_queries + S_THREADS - 1) / S_THREADS;
    if (d_filt) {
        const int g = persistent_grid(k_detect_fused<true>, 8, n_sms, smem);
        k_detect_fused<true><<<want < g ? want : g, S_THREADS, smem, st>>>(L, tab, gd, d_kinds, d_xf, d_filt, d_out, sc.ctl, sc.sv_entry, avail_all);
    } else {
        const int g = persistent_grid(k_detect_fused<false>, 9, n_sms, smem);
        k_detect_fused<false><<<want < g ? want : g, S_THREADS, smem, st>>>(L, tab, gd, d_kinds, d_xf, d_filt, d_out, sc.ctl, sc.sv_entry, avail_all);
    }
    if (n_launches) *n_launches += 1;
    return cudaGetLastError();
}

// Chooses where the node arrays live and how the stacks are sized for a layout.
DetectPlan make_detect_plan(int n_nodes, int depth) {
    DetectPlan p;
    p.cap_in = 96 * (depth > 0 ? depth : 1) + 128;
    // pair word of stage 2: node << (12 + gbits); shrink the group when the tree has many nodes
    p.gbits = 3;
    while (p.gbits > 0 && ((uint64_t)n_nodes >> (32 - 12 - p.gbits)) != 0) p.gbits--;
    p.nodes_in_smem = node_smem_bytes(n_nodes) <= 72 * 1024 ? 1 : 0;
    p.use_grid = 0;  // set by the engine after k_grid_build has run (detect_grid_usable)
    p.split_min = JG_SPLIT_MIN_QUERIES;
    return p;
}

}  // namespace jg
