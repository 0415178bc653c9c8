// Grid form of the collect path (sm_100a): collect_surrogate_collisions + collect_poly_collisions
// (cd_engine.rs:261-310) into one collector per query (lbf/benches/ci_bench.rs:56-70), for quadtrees of depth
// <= 5 and items without surrogate piers. Included by detect.cu after detect_grid.cuh.
//
// QTNode::collect_collisions (qt_node.rs:151-191) descends while a node has children AND more active edges
// than cd_threshold; the first node where that fails is a COLLECT TERMINAL, and there every hazard of the node
// is examined: Entire -> inserted, Partial -> inserted iff QTHazPartial::collides_with. The collector is a set,
// so a query's result is the union, over its entities (poles from the root, edges from the virtual root) and
// over the collect terminals each entity reaches, of the hazards that pass. As in detect_grid.cuh:
//   * a circle reaches a terminal iff circ_rect passes for the terminal's own bbox (monotone under nesting);
//   * an edge's candidate terminals are those covering the cells that meet its delta-inflated bbox inside the
//     virtual root; a hazard is inserted only after the reference's path to that terminal has been verified
//     (an endpoint inside the terminal's bbox, or path_reaches). Terminals above the virtual root collapse onto
//     the virtual root itself: active-edge counts only shrink towards the leaves (every child entry holds a
//     subset of its parent entry's edges, qt_hazard.rs:91-109), so the virtual root is then terminal too.
// Afterwards the Exterior-by-bbox rule (cd_engine.rs:262-264) and the containment tests of the virtual root's
// Partial hazards that are not yet in the set (cd_engine.rs:274-288).
#pragma once

namespace jg {

struct GridCollect {
    int ok, pad[3];
    uint32_t rows[GRID_G];                 // finest cells whose collect terminal holds at least one hazard entry
    uint32_t colfirst[GRID_G];             // ... cells in the first column / the first row of their terminal: a window visits a
    uint32_t rowfirst[GRID_G];             //     terminal at its first cell inside the window, found with row-mask arithmetic
    uint16_t cell_node[GRID_G * GRID_G];   // collect terminal covering the cell
    uint8_t cell_level[GRID_G * GRID_G];   // ... and its level
};

__global__ void __launch_bounds__(1024) k_cgrid_build(const DevLayout L, const GridDev* __restrict__ g, GridCollect* __restrict__ cg) {
    const int tid = threadIdx.x;
    for (int i = tid; i < GRID_G; i += blockDim.x) cg->rows[i] = 0u, cg->colfirst[i] = 0u, cg->rowfirst[i] = 0u;
    __syncthreads();
    if (!g->ok) {
        if (tid == 0) cg->ok = 0;
        return;
    }
    const int D = g->D, G = g->G;
    for (int c = tid; c < G * G; c += blockDim.x) {
        const int i = c % G, j = c / G;
        int node = 0, level = 0;
        for (;;) {
            const int4 m = L.node_meta[node];
            const bool perform_cd_now = L.node_nactive[node] <= (uint32_t)L.cd_threshold;  // qt_node.rs:157
            if (m.x < 0 || perform_cd_now || level >= D) break;
            level++;
            const int ci = i >> (D - level), cj = j >> (D - level);
            node = m.x + quad_of(ci & 1, cj & 1);
        }
        const int4 m = L.node_meta[node];
        cg->cell_node[j * GRID_G + i] = (uint16_t)node;
        cg->cell_level[j * GRID_G + i] = (uint8_t)level;
        if ((m.z & 0xFFFF) + ((m.z >> 16) & 0xFFFF) > 0) atomicOr(&cg->rows[j], 1u << i);
        const int sh = D - level;
        if (((i >> sh) << sh) == i) atomicOr(&cg->colfirst[j], 1u << i);
        if (((j >> sh) << sh) == j) atomicOr(&cg->rowfirst[j], 1u << i);
    }
    if (tid == 0) cg->ok = 1;
}

#define CG_MAX_WORDS 8
#ifndef CG_GROUP
#define CG_GROUP 16  // queries a warp resolves together: two poles per query fill the 32 lanes of the pole stage
#endif
struct CGSlots {  // per-warp state of the queries of the current group
    float4 xf[CG_GROUP];
    float4 bb[CG_GROUP];
    int vnode[CG_GROUP];
    int vcell[CG_GROUP];  // level << 16 | row << 8 | column of the virtual root at its level
    int vterm[CG_GROUP];  // 1: the virtual root itself is the collect terminal of everything below it
    uint32_t bits[CG_GROUP][CG_MAX_WORDS];
};

__device__ __forceinline__ bool cg_has(const uint32_t* bits, int slot) { return (((volatile const uint32_t*)bits)[slot >> 5] >> (slot & 31)) & 1u; }
__device__ __forceinline__ void cg_set(uint32_t* bits, int slot) { atomicOr(bits + (slot >> 5), 1u << (slot & 31)); }

// every hazard entry of collect terminal `node` against one entity; `reach` decides lazily (once) whether the
// reference's walk gets to this terminal at all
template <class Ent, class Reach>
__device__ __forceinline__ void cg_terminal(const DevLayout& L, int node, const Ent& ent, uint32_t* bits, Reach reach) {
    const int4 m = __ldg(&L.node_meta[node]);
    const int n_all = (m.z & 0xFFFF) + ((m.z >> 16) & 0xFFFF);
    int reached = -1;  // unknown
    for (int i = 0; i < n_all; i++) {
        const int4 em = __ldg(&L.entry_meta[m.y + i]);
        if (cg_has(bits, em.x)) continue;  // collector.contains_key: skipping only saves work
        bool hit = em.w == 2;              // Entire
        if (!hit && ent_rect(ent, ld_rect(&L.entry_ff[m.y + i])))
            for (int k = 0; k < em.z && !hit; k++) hit = ent_seg(ent, ld_seg(&L.edges[em.y + k]));
        if (!hit) continue;
        if (reached < 0) reached = reach() ? 1 : 0;
        if (!reached) return;
        cg_set(bits, em.x);
    }
}

// CDEngine::detect_containment_collision (cd_engine.rs:231-255) for up to 32 hazards (lane l: slot hs or -1);
// returns the ballot of lanes whose hazard collides. Point-in-polygon tests run one hazard at a time, warp-wide.
__device__ __forceinline__ unsigned containment_mask_warp(const DevLayout& L, const float2* iverts, int n_iverts, float ipoi_x,
                                                          float ipoi_y, const Xf& xf, const Rect& sb, int hs, int lane) {
    int rel = REL_DISJOINT, scope_interior = 1;
    if (hs >= 0) {
        scope_interior = __ldg(&L.haz_meta[hs]).x == 1 ? 1 : 0;
        rel = rect_almost_relation(ld_rect(&L.haz_bbox[hs]), sb);
    }
    const bool pip = hs >= 0 && (rel == REL_SURROUNDING || rel == REL_ENCLOSED);
    unsigned res = __ballot_sync(FULL, hs >= 0 && !pip && !scope_interior);  // not contained + Exterior scope
    unsigned todo = __ballot_sync(FULL, pip);
    while (todo) {
        const int src = __ffs(todo) - 1;
        todo &= todo - 1;
        const int h = __shfl_sync(FULL, hs, src), r = __shfl_sync(FULL, rel, src), si = __shfl_sync(FULL, scope_interior, src);
        const int4 hm = __ldg(&L.haz_meta[h]);
        const Rect hb = ld_rect(&L.haz_bbox[h]);
        bool contained;
        if (r == REL_SURROUNDING) {
            float px, py;
            xf_point(xf, ipoi_x, ipoi_y, px, py);
            contained = poly_point_warp<false>(L.haz_verts + hm.y, hm.z, hb, px, py, xf, lane);
        } else {
            const float4 hp = __ldg(&L.haz_poi[h]);
            contained = poly_point_warp<true>(iverts, n_iverts, sb, hp.x, hp.y, xf, lane);
        }
        if (si ? contained : !contained) res |= 1u << src;
    }
    return res;
}

// PIERS: the pier stage is compiled in only for items that carry piers (the kernel is sensitive to its code size: with the
// stage compiled in and unused, the shirts batch ran 8 % slower)
#ifndef CG_MIN_BLOCKS
#define CG_MIN_BLOCKS 5  // 48 registers: shirts 10M 0.74 -> 0.756 G queries/s, trousers 1.02 -> 1.05
#endif
template <bool PIERS>
__global__ void __launch_bounds__(S_THREADS, CG_MIN_BLOCKS)
k_collect_grid(const DevLayout L, const __grid_constant__ SegTable gtab, const GridDev* __restrict__ ggrid, const GridCollect* __restrict__ cg,
               const float4* __restrict__ xforms, int mode, const uint32_t* __restrict__ seed_bits, uint32_t* __restrict__ out_bits, int words,
               int* __restrict__ next, int GS, int n_groups) {
    extern __shared__ __align__(16) unsigned char smem[];
    SegTable* t = (SegTable*)smem;
    for (int i = threadIdx.x; i < (int)(sizeof(SegTable) / 4); i += blockDim.x) ((int*)t)[i] = ((const int*)&gtab)[i];
    unsigned char* p = smem + ((sizeof(SegTable) + 15) & ~(size_t)15);
    const GridDev& g = *stage_grid(p, ggrid);
    __syncthreads();
    const SegTable& tab = *t;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    CGSlots& W = *(CGSlots*)(p + (size_t)warp * sizeof(CGSlots));
    const int D = g.D, G = g.G;
    // GS = queries per group (chosen by the launcher): 16 when every warp has several groups to do (best lane use); fewer for
    // small launches, where the length of a warp's dependent chain through one group, not the issue rate, sets the kernel's
    // duration. A group never straddles two segments: segment k owns the groups [seg[k].gbeg, seg[k + 1].gbeg).
    const int quad = lane >> 2, sub = lane & 3;
    int seg_hint = 0;

    for (;;) {
        int gi = 0;
        if (lane == 0) gi = atomicAdd(next, 1);
        gi = __shfl_sync(FULL, gi, 0);
        if (gi >= n_groups) break;
        while (seg_hint + 1 < tab.n_seg && gi >= tab.seg[seg_hint + 1].gbeg) seg_hint++;  // groups are claimed in increasing order
        const SegDesc& it = tab.seg[seg_hint];
        const int seg_end = seg_hint + 1 < tab.n_seg ? tab.seg[seg_hint + 1].begin : tab.n_queries;
        const int qbase = it.begin + (gi - it.gbeg) * GS;
        const int ng = (seg_end - qbase) < GS ? (seg_end - qbase) : GS;
        const int V = it.n_verts;
        const float inv_v = 1.0f / (float)V;  // f / V for f < 16 V without the integer division: (f + 0.5) / V is >= 0.5 / V away
                                              // from an integer, far beyond the rounding of two f32 operations on values < 2^17
        if (lane < ng) W.xf[lane] = __ldg(&xforms[qbase + lane]);
        // collectors start empty, or from the caller's seed: a hazard already present is skipped everywhere below and stays set
        for (int i = lane; i < GS * CG_MAX_WORDS; i += 32) {
            const int s = i / CG_MAX_WORDS, wd = i - s * CG_MAX_WORDS;
            (&W.bits[0][0])[i] = (seed_bits && s < ng && wd < words) ? __ldg(&seed_bits[(size_t)(qbase + s) * words + wd]) : 0u;
        }
        __syncwarp();

        // ---- surrogate poles from the root (cd_engine.rs:296-310); lane = (query, pole)
        if (!(mode & JG_MODE_POLY_ONLY)) {
            const int n_ent = ng * it.n_poles;
            for (int e0 = 0; e0 < n_ent; e0 += 32) {
                const int e = e0 + lane;
                if (e < n_ent) {
                    const int s = e / it.n_poles, pj = e - s * it.n_poles;
                    const float4 t4 = W.xf[s];
                    const float4 pl = __ldg(&it.poles[pj]);
                    Circ c;
                    xf_point(Xf{t4.x, t4.y, t4.z, t4.w}, pl.x, pl.y, c.cx, c.cy);
                    c.r = pl.z;
                    const PoleWin w = pole_window(g, c);
                    const uint32_t cm = col_mask(w.i0, w.i1);
                    for (int j = w.j0; j <= w.j1; j++) {
                        // one visit per terminal and circle: at the terminal's first cell inside the window, i.e. the cells
                        // with (i == ti || i == i0) && (j == tj || j == j0)
                        uint32_t bits = cg->rows[j] & cm & (cg->colfirst[j] | (1u << w.i0));
                        if (j != w.j0) bits &= cg->rowfirst[j];
                        while (bits) {
                            const int i = __ffs(bits) - 1;
                            bits &= bits - 1;
                            const int lt = cg->cell_level[j * GRID_G + i], sh = D - lt;
                            const int ti = (i >> sh) << sh, tj = (j >> sh) << sh;  // first finest cell of the terminal
                            const Rect tb{g.xs[ti], g.ys[tj], g.xs[ti + (1 << sh)], g.ys[tj + (1 << sh)]};
                            if (!circ_rect(c, tb)) continue;  // circle.rs:115-128: reached iff its own bbox passes
                            cg_terminal<Circ>(L, cg->cell_node[j * GRID_G + i], c, W.bits[s], [] { return true; });
                        }
                    }
                }
            }
            __syncwarp();
            // ---- surrogate piers from the root (cd_engine.rs:305-308), after the poles: lane = (query, pier), an edge entity
            // whose walk starts at the root: the window is not clipped and the path is verified from level 0
            const int n_pent = PIERS ? ng * it.n_piers : 0;
            for (int e0 = 0; e0 < n_pent; e0 += 32) {
                const int e = e0 + lane;
                if (e < n_pent) {
                    const int s = e / it.n_piers, pj = e - s * it.n_piers;
                    const float4 t4 = W.xf[s];
                    const Xf xf{t4.x, t4.y, t4.z, t4.w};
                    const float4 pr = __ldg(&it.piers[pj]);
                    Seg sg;
                    xf_point(xf, pr.x, pr.y, sg.sx, sg.sy);
                    xf_point(xf, pr.z, pr.w, sg.ex, sg.ey);
                    uint32_t* bits = W.bits[s];
                    const float amax = fmaxf(fmaxf(fabsf(sg.sx), fabsf(sg.ex)), fmaxf(fabsf(sg.sy), fabsf(sg.ey)));
                    const float delta = fmaxf(g.delta, amax * (1.0f / 8192.0f));
                    const int i0 = cell_lo(g.xs, G, fminf(sg.sx, sg.ex) - delta, g.x0, g.inv_cw);
                    const int i1 = cell_hi(g.xs, G, fmaxf(sg.sx, sg.ex) + delta, g.x0, g.inv_cw);
                    const int j0 = cell_lo(g.ys, G, fminf(sg.sy, sg.ey) - delta, g.y0, g.inv_ch);
                    const int j1 = cell_hi(g.ys, G, fmaxf(sg.sy, sg.ey) + delta, g.y0, g.inv_ch);
                    const uint32_t cm = col_mask(i0, i1);
                    for (int j = j0; j <= j1; j++) {
                        uint32_t rb = cm ? cg->rows[j] & cm & (cg->colfirst[j] | (1u << i0)) : 0u;
                        if (j != j0) rb &= cg->rowfirst[j];
                        while (rb) {
                            const int i = __ffs(rb) - 1;
                            rb &= rb - 1;
                            cg_terminal<Seg>(L, cg->cell_node[j * GRID_G + i], sg, bits, [&] {
                                const int lt = cg->cell_level[j * GRID_G + i], sh = D - lt;
                                const int ti = (i >> sh) << sh, tj = (j >> sh) << sh;
                                const Rect tb{g.xs[ti], g.ys[tj], g.xs[ti + (1 << sh)], g.ys[tj + (1 << sh)]};
                                return rect_point(tb, sg.sx, sg.sy) || rect_point(tb, sg.ex, sg.ey) || path_reaches(g, sg, 0, lt, i, j);
                            });
                        }
                    }
                }
            }
            __syncwarp();
        }

        // ---- polygon: bbox (quad = query, eight queries per pass), Exterior rule, virtual root
        for (int half = 0; half < ng; half += 8) {
            const int qs = half + quad;
            float x0 = 3.402823466e+38f, y0 = 3.402823466e+38f, x1 = -3.402823466e+38f, y1 = -3.402823466e+38f;
            if (qs < ng) {
                const float4 t4 = W.xf[qs];
                const Xf xf{t4.x, t4.y, t4.z, t4.w};
                for (int j = sub; j < V; j += 4) {
                    const float2 v = __ldg(&it.verts[j]);
                    float x, y;
                    xf_point(xf, v.x, v.y, x, y);
                    x0 = fminf(x0, x), y0 = fminf(y0, y), x1 = fmaxf(x1, x), y1 = fmaxf(y1, y);
                }
            }
#pragma unroll
            for (int o = 2; o > 0; o >>= 1) {
                x0 = fminf(x0, __shfl_xor_sync(FULL, x0, o));
                y0 = fminf(y0, __shfl_xor_sync(FULL, y0, o));
                x1 = fmaxf(x1, __shfl_xor_sync(FULL, x1, o));
                y1 = fmaxf(y1, __shfl_xor_sync(FULL, y1, o));
            }
            const Rect bb{x0, y0, x1, y1};
            if (qs < ng && sub == 0) {
                W.bb[qs] = make_float4(x0, y0, x1, y1);
                if (rect_relation(L.cde_bbox, bb) != REL_SURROUNDING) cg_set(W.bits[qs], L.exterior_slot);  // cd_engine.rs:262-264
            }
            bool walking = qs < ng;
            int level = 0, ci = 0, cj = 0, vnode = 0;
            for (;;) {  // get_virtual_root (cd_engine.rs:314-327): lane = child quadrant, first surrounding child wins
                bool sur = false;
                const int nl = level + 1, sh = D - nl, w = 1 << nl, base = mip_off(nl);
                if (walking && level < D && g.mip_node[base + (2 * cj) * w + 2 * ci] != 0xFFFF) {
                    const int xi = 2 * ci + ((sub == 0 || sub == 3) ? 1 : 0), yj = 2 * cj + ((sub == 0 || sub == 1) ? 1 : 0);
                    const Rect cb{g.xs[xi << sh], g.ys[yj << sh], g.xs[(xi + 1) << sh], g.ys[(yj + 1) << sh]};
                    sur = rect_relation(cb, bb) == REL_SURROUNDING;
                }
                const unsigned all = __ballot_sync(FULL, sur);
                if (!all) break;
                const unsigned mine = (all >> (quad * 4)) & 0xFu;
                if (mine) {
                    const int found = __ffs(mine) - 1;
                    ci = 2 * ci + ((found == 0 || found == 3) ? 1 : 0);
                    cj = 2 * cj + ((found == 0 || found == 1) ? 1 : 0);
                    level = nl;
                    vnode = g.mip_node[base + cj * w + ci];
                } else {
                    walking = false;
                }
            }
            if (qs < ng && sub == 0) {
                W.vnode[qs] = vnode;
                W.vcell[qs] = (level << 16) | (cj << 8) | ci;
                const int sh = D - level;
                W.vterm[qs] = cg->cell_level[(cj << sh) * GRID_G + (ci << sh)] <= level ? 1 : 0;
            }
        }
        __syncwarp();

        // ---- edges from the virtual root; lane = edge of the group's flat edge list
        const int E = ng * V;
        for (int f0 = 0; f0 < E; f0 += 32) {
            const int f = f0 + lane;
            if (f < E) {
                const int s = (int)(((float)f + 0.5f) * inv_v), e = f - s * V;
                const float4 t4 = W.xf[s];
                const Xf xf{t4.x, t4.y, t4.z, t4.w};
                const float2 a = __ldg(&it.verts[e]), b = __ldg(&it.verts[e == V - 1 ? 0 : e + 1]);
                Seg sg;
                xf_point(xf, a.x, a.y, sg.sx, sg.sy);
                xf_point(xf, b.x, b.y, sg.ex, sg.ey);
                uint32_t* bits = W.bits[s];
                if (W.vterm[s]) {
                    cg_terminal<Seg>(L, W.vnode[s], sg, bits, [] { return true; });  // the walk starts (and ends) there
                } else {
                    const int vc = W.vcell[s];
                    const int l0 = vc >> 16, s0 = D - l0;
                    // the superset argument of detect_grid.cuh scales with the coordinates involved
                    const float amax = fmaxf(fmaxf(fabsf(sg.sx), fabsf(sg.ex)), fmaxf(fabsf(sg.sy), fabsf(sg.ey)));
                    const float delta = fmaxf(g.delta, amax * (1.0f / 8192.0f));
                    int i0 = cell_lo(g.xs, G, fminf(sg.sx, sg.ex) - delta, g.x0, g.inv_cw);
                    int i1 = cell_hi(g.xs, G, fmaxf(sg.sx, sg.ex) + delta, g.x0, g.inv_cw);
                    int j0 = cell_lo(g.ys, G, fminf(sg.sy, sg.ey) - delta, g.y0, g.inv_ch);
                    int j1 = cell_hi(g.ys, G, fmaxf(sg.sy, sg.ey) + delta, g.y0, g.inv_ch);
                    const int vi0 = (vc & 255) << s0, vj0 = ((vc >> 8) & 255) << s0;
                    i0 = max(i0, vi0), i1 = min(i1, vi0 + (1 << s0) - 1);
                    j0 = max(j0, vj0), j1 = min(j1, vj0 + (1 << s0) - 1);
                    const uint32_t cm = col_mask(i0, i1);
                    for (int j = j0; j <= j1; j++) {
                        // one visit per terminal and edge (see the poles above)
                        uint32_t rb = cm ? cg->rows[j] & cm & (cg->colfirst[j] | (1u << i0)) : 0u;
                        if (j != j0) rb &= cg->rowfirst[j];
                        while (rb) {
                            const int i = __ffs(rb) - 1;
                            rb &= rb - 1;
                            cg_terminal<Seg>(L, cg->cell_node[j * GRID_G + i], sg, bits, [&] {
                                // only asked once an entry of the terminal is hit: its level and box are not needed before
                                const int lt = cg->cell_level[j * GRID_G + i], sh = D - lt;
                                const int ti = (i >> sh) << sh, tj = (j >> sh) << sh;
                                const Rect tb{g.xs[ti], g.ys[tj], g.xs[ti + (1 << sh)], g.ys[tj + (1 << sh)]};
                                return rect_point(tb, sg.sx, sg.sy) || rect_point(tb, sg.ex, sg.ey) || path_reaches(g, sg, l0, lt, i, j);
                            });
                        }
                    }
                }
            }
        }
        __syncwarp();

        // ---- containment for the virtual root's Partial hazards not collected yet (cd_engine.rs:274-288)
        for (int s = 0; s < ng; s++) {
            const int4 m = __ldg(&L.node_meta[W.vnode[s]]);
            const int n_entire = m.z & 0xFFFF, n_partial = (m.z >> 16) & 0xFFFF;
            const float4 t4 = W.xf[s], b4 = W.bb[s];
            for (int j0 = 0; j0 < n_partial; j0 += 32) {
                const int j = j0 + lane;
                int hs = -1;
                if (j < n_partial) {
                    hs = __ldg(&L.entry_meta[m.y + n_entire + j]).x;
                    if (cg_has(W.bits[s], hs)) hs = -1;
                }
                const unsigned hitm = containment_mask_warp(L, it.verts, V, it.poi_x, it.poi_y, Xf{t4.x, t4.y, t4.z, t4.w},
                                                            Rect{b4.x, b4.y, b4.z, b4.w}, hs, lane);
                if ((hitm >> lane) & 1u) cg_set(W.bits[s], hs);
            }
        }
        __syncwarp();
        for (int i = lane; i < ng * words; i += 32) {
            const int s = i / words, wd = i - s * words;
            out_bits[(size_t)(qbase + s) * words + wd] = ((volatile uint32_t*)W.bits[s])[wd];
        }
        __syncwarp();
    }
}

}  // namespace jg
