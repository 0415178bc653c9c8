// sm_100a query kernels of the batched CDE: device-side views of the flattened layout and the
// traversal / predicate code shared by the detect and collect kernels.
// Compile with -fmad=false -prec-div=true -prec-sqrt=true -ftz=false (see predicates.cuh).
#pragma once
#include <cuda_runtime.h>

#include "layout_soa.h"
#include "predicates.cuh"

namespace jg {

// Pointers into the device copy of the blob (layout_soa.h) + the scalars the kernels need.
struct DevLayout {
    const int4* node_meta;
    const float4* node_bbox;
    const uint32_t* node_nactive;
    const int4* entry_meta;
    const float4* entry_ff;
    const float4* edges;
    const int4* haz_meta;
    const float4* haz_bbox;
    const float4* haz_poi;
    const float2* haz_verts;
    Rect cde_bbox;
    int n_nodes, n_slots, cd_threshold, exterior_slot;
    int n_entries, n_edges;
};

// One item's collision shape in device memory (SPolygon vertices + poi centre + ff surrogate).
struct DevItem {
    const float2* verts;
    int n_verts;
    float poi_x, poi_y;
    const float4* poles;  // (x, y, r, 0) — the first n_ff poles only
    int n_poles;
    const float4* piers;  // (sx, sy, ex, ey)
    int n_piers;
};

#define JG_MAX_DEPTH 15  // 4 pending bits per level in one 64-bit register

__device__ __forceinline__ Rect ld_rect(const float4* p) {
    const float4 v = __ldg(p);
    return Rect{v.x, v.y, v.z, v.w};
}
// same for a pointer that may be shared memory (staged node arrays)
__device__ __forceinline__ Rect ld_rect_any(const float4* p) {
    const float4 v = *p;
    return Rect{v.x, v.y, v.z, v.w};
}
__device__ __forceinline__ Seg ld_seg(const float4* p) {
    const float4 v = __ldg(p);
    return Seg{v.x, v.y, v.z, v.w};
}
__device__ __forceinline__ bool filt_bit(const uint32_t* filt, int slot) {
    return (__ldg(filt + (slot >> 5)) >> (slot & 31)) & 1u;
}

// Depth-first walk state: the current node plus, per level below the start node, the 4-bit mask of
// sibling quadrants that still have to be visited. Children blocks start at flat indices 1 + 4k
// (breadth-first numbering), so a sibling is reached without touching the parent.
struct Walk {
    int node;
    int depth;
    unsigned long long pending;
    __device__ __forceinline__ void start(int n) { node = n; depth = 0; pending = 0ull; }
    // enter the first quadrant of `mask` (non-zero) below the current node, remember the others
    __device__ __forceinline__ void descend(int child_base, unsigned mask) {
        const int q = __ffs(mask) - 1;
        mask &= mask - 1;
        pending |= (unsigned long long)mask << (4 * depth);
        node = child_base + q;
        depth++;
    }
    // advance to the next pending node; false when the walk is finished
    __device__ __forceinline__ bool next(const DevLayout& L) {
        while (depth > 0) {
            const int sh = 4 * (depth - 1);
            const unsigned pm = (unsigned)(pending >> sh) & 0xFu;
            if (pm) {
                const int q = __ffs(pm) - 1;
                pending &= ~((unsigned long long)(pm & (0u - pm)) << sh);
                node = ((node - 1) & ~3) + 1 + q;
                return true;
            }
            node = __ldg(&L.node_meta[node]).w;  // parent
            depth--;
        }
        return false;
    }
};

// QTNode::collides (qt_node.rs:99-147) from `start`: true iff some hazard collides with `ent`.
//   strongest (qt_hazard_vec.rs:60-63) = first non-filtered entry of the Entire-first vector:
//   none -> nothing here; Entire -> hit; Partial -> children flagged by collides_with_quadrants, or,
//   without children, QTHazPartial::collides_with (qt_partial_hazard.rs:67-72) per relevant entry.
template <class Ent, bool FILT>
__device__ bool qt_collides(const DevLayout& L, int start, const Ent& ent, const uint32_t* filt) {
    Walk w;
    w.start(start);
    do {
        const int4 m = __ldg(&L.node_meta[w.node]);
        const int n_entire = m.z & 0xFFFF, n_partial = (m.z >> 16) & 0xFFFF;
        int first_partial = 0;  // index (within the partial range) of the first relevant entry
        int kind;               // 0 none, 1 partial, 2 entire
        if (!FILT) {
            kind = n_entire ? 2 : (n_partial ? 1 : 0);
        } else {
            kind = 0;
            for (int i = 0; i < n_entire; i++)
                if (!filt_bit(filt, __ldg(&L.entry_meta[m.y + i]).x)) { kind = 2; break; }
            if (kind == 0)
                for (int i = 0; i < n_partial; i++)
                    if (!filt_bit(filt, __ldg(&L.entry_meta[m.y + n_entire + i]).x)) { kind = 1; first_partial = i; break; }
        }
        if (kind == 2) return true;
        if (kind == 1) {
            if (m.x >= 0) {
                const unsigned mask = ent_quadrants(ent, ld_rect(&L.node_bbox[w.node]));
                if (mask) { w.descend(m.x, mask); continue; }
            } else {
                const int e0 = m.y + n_entire;
                for (int i = first_partial; i < n_partial; i++) {
                    const int4 em = __ldg(&L.entry_meta[e0 + i]);
                    if (FILT && filt_bit(filt, em.x)) continue;
                    if (!ent_rect(ent, ld_rect(&L.entry_ff[e0 + i]))) continue;
                    for (int k = 0; k < em.z; k++)
                        if (ent_seg(ent, ld_seg(&L.edges[em.y + k]))) return true;
                }
            }
        }
        if (!w.next(L)) return false;
    } while (true);
}

// QTNode::collect_collisions (qt_node.rs:151-191) from `start` into a bitset over hazard slots.
// `bits` is shared by the lanes of one query (shared memory); already-set hazards are skipped like
// `collector.contains_key` — which only saves work, the resulting set is a union either way.
template <class Ent>
__device__ void qt_collect(const DevLayout& L, int start, const Ent& ent, uint32_t* bits) {
    Walk w;
    w.start(start);
    do {
        const int4 m = __ldg(&L.node_meta[w.node]);
        const bool perform_cd_now = __ldg(&L.node_nactive[w.node]) <= (uint32_t)L.cd_threshold;
        if (m.x >= 0 && !perform_cd_now) {
            const unsigned mask = ent_quadrants(ent, ld_rect(&L.node_bbox[w.node]));
            if (mask) { w.descend(m.x, mask); continue; }
        } else {
            const int n_all = (m.z & 0xFFFF) + ((m.z >> 16) & 0xFFFF);
            for (int i = 0; i < n_all; i++) {
                const int4 em = __ldg(&L.entry_meta[m.y + i]);
                const uint32_t bit = 1u << (em.x & 31);
                volatile uint32_t* word = bits + (em.x >> 5);
                if (*word & bit) continue;
                bool hit = em.w == 2;
                if (!hit && ent_rect(ent, ld_rect(&L.entry_ff[m.y + i])))
                    for (int k = 0; k < em.z && !hit; k++) hit = ent_seg(ent, ld_seg(&L.edges[em.y + k]));
                if (hit) atomicOr((uint32_t*)word, bit);
            }
        }
        if (!w.next(L)) return;
    } while (true);
}

// CDEngine::get_virtual_root (cd_engine.rs:314-327): lowest node whose bbox surrounds `bb`; at each
// level the FIRST child (NE, NW, SW, SE) with relation Surrounding is taken.
__device__ __forceinline__ int virtual_root(const DevLayout& L, const Rect& bb) {
    int node = 0;
    for (;;) {
        const int cb = __ldg(&L.node_meta[node]).x;
        if (cb < 0) return node;
        int nxt = -1;
#pragma unroll
        for (int i = 0; i < 4; i++)
            if (nxt < 0 && rect_relation(ld_rect(&L.node_bbox[cb + i]), bb) == REL_SURROUNDING) nxt = cb + i;
        if (nxt < 0) return node;
        node = nxt;
    }
}

// CDEngine::detect_containment_collision (cd_engine.rs:231-255) between the query shape (item under
// xf, bbox `sb`) and the hazard in `slot`.
__device__ inline bool containment_collision(const DevLayout& L, const float2* iverts, int n_iverts, float ipoi_x,
                                             float ipoi_y, const Xf& xf, const Rect& sb, int slot) {
    const int4 hm = __ldg(&L.haz_meta[slot]);
    const Rect hb = ld_rect(&L.haz_bbox[slot]);
    const int rel = rect_almost_relation(hb, sb);
    bool contained = false;
    if (rel == REL_SURROUNDING) {
        float px, py;
        xf_point(xf, ipoi_x, ipoi_y, px, py);
        contained = poly_point<false>((const float*)(L.haz_verts + hm.y), hm.z, hb, px, py, xf);
    } else if (rel == REL_ENCLOSED) {
        const float4 hp = __ldg(&L.haz_poi[slot]);
        contained = poly_point<true>((const float*)iverts, n_iverts, sb, hp.x, hp.y, xf);
    }
    return hm.x == 1 ? contained : !contained;  // hazard.rs:55-62 scope: Interior / Exterior
}

}  // namespace jg
