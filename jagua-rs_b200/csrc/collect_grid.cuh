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
#ifndef CG_TRIPLES
#define CG_TRIPLES 1  // lane = (query, terminal, entry) in the resolve step (0: lane = (query, terminal), for A/B)
#endif
#ifndef CG_GROUP
#define CG_GROUP 8  // queries a warp resolves together
#endif
// Shared memory is the scarce resource of this kernel: the layout's entries and edges live in L1, which shares the SM's 256 KB
// with it (52 KB per 256-thread CTA at 4 CTAs per SM left L1 a 52 % hit rate and the kernel latency-bound - profiles/r2b_*;
// with 2.2 KB per warp and one staged grid summary per 512 threads the hit rate is back at 99 %).
#define CG_THREADS 512
#define CG_WARPS (CG_THREADS / 32)
#define CG_VERT_CAP 96   // transformed vertices a warp keeps in shared memory for its group; longer polygons are transformed on the fly
#define CG_MAX_POLES 4   // fail-fast poles / piers per item on this path (the reference's defaults: 2 / 0); more: walk kernel
#define CG_MAX_PIERS 4
#define CG_EMIT 4        // (query, terminal) pairs a lane lists per pass
#define CG_LIST_CAP (32 * CG_EMIT + CG_GROUP)
struct CGSlots {  // per-warp state of the queries of the current group
    float4 xf[CG_GROUP];
    float4 bb[CG_GROUP];
    int vnode[CG_GROUP];
    int vcell[CG_GROUP];  // level << 16 | row << 8 | column of the virtual root at its level; bit 30: the virtual root itself is
                          // the collect terminal of everything below it
    float dq[CG_GROUP];   // candidate inflation of the query's edges (see detect_grid.cuh; scaled to the coordinates involved)
    uint32_t bits[CG_GROUP][CG_MAX_WORDS];
    float4 pole[CG_GROUP][CG_MAX_POLES];  // transformed fail-fast poles (cx, cy, r, -)
    float2 tv[CG_VERT_CAP];               // transformed vertices, query s at s * V
    uint16_t list[CG_LIST_CAP];           // (query, terminal) pairs of the current pass
#if CG_TRIPLES
    int4 pp[32];        // per pair of the current 32: {first entry, candidate edges, edges ending inside, packed(query, vpair, poles, levels, cell)}
    int pfx[36];        // prefix sums of the pairs' entry counts
#endif
};
struct CGPierSlots {  // ... of the instantiation for items whose surrogate carries piers
    float4 pier[CG_GROUP][CG_MAX_PIERS];  // transformed fail-fast piers (sx, sy, ex, ey)
};
template <bool PIERS>
struct CGWarp : CGSlots {};
template <>
struct CGWarp<true> : CGSlots, CGPierSlots {};
#define CG_VTERM (1 << 30)
#define CG_PAIR_VNODE 0x4000u  // the pair (query, its virtual root as terminal): edges only, reached by definition

__device__ __forceinline__ bool cg_has(const uint32_t* bits, int slot) { return (((volatile const uint32_t*)bits)[slot >> 5] >> (slot & 31)) & 1u; }
__device__ __forceinline__ void cg_set(uint32_t* bits, int slot) { atomicOr(bits + (slot >> 5), 1u << (slot & 31)); }

// ---------------------------------------------------------------------------------------------------
// k_collect_grid. A warp resolves a group of GS queries of one item:
//   1. transformed vertices (kept in shared memory), bbox, Exterior-by-bbox rule, virtual root (quad = query);
//   2. transformed fail-fast poles / piers;
//   3. LANE = (QUERY, COLLECT TERMINAL): every terminal with hazard entries that meets the query's window is listed once
//      per query (window = the delta-inflated bbox joined with the pole / pier windows; the first-cell masks of
//      k_cgrid_build visit a terminal at its first cell inside the window), 32 pairs are resolved at a time. A pair walks the
//      terminal's entries ONCE for all entities of the query: per entry, first the poles that reach the terminal
//      (circ_rect on its own box - exact, detect_grid.cuh), then the edges whose inflated box meets the terminal (the
//      candidate superset of detect_grid.cuh, as a bit mask built once per pair), the query's bbox against the entry's
//      ff_bbox as an exact early reject (an edge outside the bbox's range is range-rejected by rect.rs:301-345 too), and
//      - only for an (entry, edge) that WOULD insert - the reference's path to the terminal (an endpoint inside its box,
//      or path_reaches; cached per edge). The collector is a set, so the order in which (entity, terminal, entry) triples
//      are examined is free; `contains_key` only skips work.
//      The virtual root as terminal of everything below it (active-edge counts shrink towards the leaves) is one more
//      pair: all edges, reached by definition.
//   4. containment tests of the virtual root's Partial hazards not yet collected; results out.
// (Round 1 ran lane = item edge with per-lane loops over terminals and entries: 9 active lanes per instruction and no
// reconvergence between the rounds of the edge loop; profiles/r2a_*.)
// ---------------------------------------------------------------------------------------------------
#ifndef CG_MIN_BLOCKS
#define CG_MIN_BLOCKS 2
#endif
template <bool PIERS>
__global__ void __launch_bounds__(CG_THREADS, CG_MIN_BLOCKS)
k_collect_grid(const DevLayout L, const __grid_constant__ SegTable gtab, const GridDev* __restrict__ ggrid, const GridCollect* __restrict__ cg,
               const float4* __restrict__ xforms, int mode, const uint32_t* __restrict__ seed_bits, uint32_t* __restrict__ out_bits, int words,
               int* __restrict__ next, int GS, int gshift, int n_groups) {
    extern __shared__ __align__(16) unsigned char smem[];
    SegTable* t = (SegTable*)smem;
    for (int i = threadIdx.x; i < (int)(sizeof(SegTable) / 4); i += blockDim.x) ((int*)t)[i] = ((const int*)&gtab)[i];
    unsigned char* p = smem + ((sizeof(SegTable) + 15) & ~(size_t)15);
    const GridDev& g = *stage_grid(p, ggrid);
    __syncthreads();
    const SegTable& tab = *t;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    CGWarp<PIERS>& W = *(CGWarp<PIERS>*)(p + (size_t)warp * sizeof(CGWarp<PIERS>));
    const int D = g.D, G = g.G;
    // GS = queries per group (chosen by the launcher: 16 when every warp has several groups to do and the vertices fit; fewer
    // for small launches, where the length of a warp's dependent chain through one group sets the kernel's duration). A group
    // never straddles two segments: segment k owns the groups [seg[k].gbeg, seg[k + 1].gbeg).
    const int quad = lane >> 2, sub = lane & 3;
    // lane = (query slot, part) in the listing step: the `parts` lanes of a query take its window's rows in turn
    const int ls = lane & ((1 << gshift) - 1), part = lane >> gshift, parts = 32 >> gshift;
    const bool surr = !(mode & JG_MODE_POLY_ONLY);
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
        const int np = surr ? it.n_poles : 0, npr = (PIERS && surr) ? it.n_piers : 0;
        const float2* __restrict__ iverts = it.verts;
        if (lane < ng) W.xf[lane] = __ldg(&xforms[qbase + lane]);
        // collectors start empty, or from the caller's seed: a hazard already present is skipped everywhere below and stays set
#pragma unroll 1
        for (int i = lane; i < GS * CG_MAX_WORDS; i += 32) {
            const int s = i / CG_MAX_WORDS, wd = i - s * CG_MAX_WORDS;
            (&W.bits[0][0])[i] = (seed_bits && s < ng && wd < words) ? __ldg(&seed_bits[(size_t)(qbase + s) * words + wd]) : 0u;
        }
        __syncwarp();

        // ---- 1. SPolygon::transform_from + bbox (simple_polygon.rs:226-253, 122-133), Exterior rule, virtual root: quad = query
        for (int half = 0; half < ng; half += 8) {
            const int qs = half + quad;
            float x0 = 3.402823466e+38f, y0 = 3.402823466e+38f, x1 = -3.402823466e+38f, y1 = -3.402823466e+38f;
            if (qs < ng) {
                const float4 t4 = W.xf[qs];
                const Xf xf{t4.x, t4.y, t4.z, t4.w};
#pragma unroll 1
                for (int j = sub; j < V; j += 4) {
                    const float2 v = __ldg(&iverts[j]);
                    float x, y;
                    xf_point(xf, v.x, v.y, x, y);
                    W.tv[qs * V + j] = make_float2(x, y);
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
                // the superset argument of detect_grid.cuh scales with the coordinates involved
                const float amax = fmaxf(fmaxf(fabsf(x0), fabsf(x1)), fmaxf(fabsf(y0), fabsf(y1)));
                W.dq[qs] = fmaxf(g.delta, amax * (1.0f / 8192.0f));
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
                const int sh = D - level;
                W.vcell[qs] = (level << 16) | (cj << 8) | ci | (cg->cell_level[(cj << sh) * GRID_G + (ci << sh)] <= level ? CG_VTERM : 0);
            }
        }
        // ---- 2. surrogate entities (circle.rs:79-85: centre transformed, radius unchanged; edge.rs:126-144: both end points)
#pragma unroll 1
        for (int e = lane; e < ng * np; e += 32) {
            const int s = e / np, pj = e - s * np;
            const float4 t4 = W.xf[s];
            const float4 pl = __ldg(&it.poles[pj]);
            float cx, cy;
            xf_point(Xf{t4.x, t4.y, t4.z, t4.w}, pl.x, pl.y, cx, cy);
            W.pole[s][pj] = make_float4(cx, cy, pl.z, 0.0f);
        }
        if constexpr (PIERS)
            for (int e = lane; e < ng * npr; e += 32) {
                const int s = e / (npr > 0 ? npr : 1), pj = e - s * npr;
                const float4 t4 = W.xf[s];
                const Xf xf{t4.x, t4.y, t4.z, t4.w};
                const float4 pr = __ldg(&it.piers[pj]);
                Seg sg;
                xf_point(xf, pr.x, pr.y, sg.sx, sg.sy);
                xf_point(xf, pr.z, pr.w, sg.ex, sg.ey);
                W.pier[s][pj] = make_float4(sg.sx, sg.sy, sg.ex, sg.ey);
            }
        __syncwarp();
        // ---- 3. (query, terminal) pairs. Listing: this lane's query window and its cursor (row j, remaining cells rb)
        int wi0 = 0, wj0 = 0, wj1 = -1, j = 0;
        uint32_t cm = 0u, rb = 0u;
        bool alive = false;
        if (ls < ng) {
            const float4 b4 = W.bb[ls];
            const float dq = W.dq[ls];
            float lx = b4.x - dq, ly = b4.y - dq, hx = b4.z + dq, hy = b4.w + dq;
#pragma unroll 1
            for (int pj = 0; pj < np; pj++) {  // poles walk from the root: their windows are not clipped to the virtual root
                const float4 c4 = W.pole[ls][pj];
                const float rk = c4.z * 1.001f + g.slack;
                lx = fminf(lx, c4.x - rk), ly = fminf(ly, c4.y - rk), hx = fmaxf(hx, c4.x + rk), hy = fmaxf(hy, c4.y + rk);
            }
            if constexpr (PIERS)
                for (int pj = 0; pj < npr; pj++) {
                    const float4 s4 = W.pier[ls][pj];
                    const float amax = fmaxf(fmaxf(fabsf(s4.x), fabsf(s4.z)), fmaxf(fabsf(s4.y), fabsf(s4.w)));
                    const float dp = fmaxf(g.delta, amax * (1.0f / 8192.0f));
                    lx = fminf(lx, fminf(s4.x, s4.z) - dp), ly = fminf(ly, fminf(s4.y, s4.w) - dp);
                    hx = fmaxf(hx, fmaxf(s4.x, s4.z) + dp), hy = fmaxf(hy, fmaxf(s4.y, s4.w) + dp);
                }
            // a window entirely beside the grid holds no cell any entity can reach
            if (!(hx < g.xs[0] || lx > g.xs[G] || hy < g.ys[0] || ly > g.ys[G])) {
                wi0 = cell_lo(g.xs, G, lx, g.x0, g.inv_cw);
                const int wi1 = cell_hi(g.xs, G, hx, g.x0, g.inv_cw);
                wj0 = cell_lo(g.ys, G, ly, g.y0, g.inv_ch);
                wj1 = cell_hi(g.ys, G, hy, g.y0, g.inv_ch);
                cm = col_mask(wi0, wi1);
                j = wj0 + part;
                alive = cm != 0u && j <= wj1;
                // one visit per terminal: at its first cell inside the window, i.e. the cells with (i == ti || i == i0) && (j == tj || j == j0)
                if (alive) {
                    rb = cg->rows[j] & cm & (cg->colfirst[j] | (1u << wi0));
                    if (j != wj0) rb &= cg->rowfirst[j];
                }
            }
        }
        bool first = true;
        for (;;) {  // passes: list up to CG_EMIT pairs per lane, resolve them 32 at a time
            __syncwarp();  // (the previous pass is done with the list)
            int cnt = 0;
            {   // the virtual root as terminal of everything below it: one pair per such query, first pass only
                const bool v = first && part == 0 && ls < ng && (W.vcell[ls] & CG_VTERM);
                const unsigned bm = __ballot_sync(FULL, v);
                if (v) W.list[__popc(bm & lanemask_lt())] = (uint16_t)((uint32_t)ls | CG_PAIR_VNODE);
                cnt = __popc(bm);
            }
#pragma unroll 1
            for (int u = 0; u < CG_EMIT; u++) {  // every lane lists its next terminal; the warp packs them (ballot + prefix count)
                uint32_t item = 0u;
                bool has = false;
                while (alive) {
                    if (rb) {
                        const int i = __ffs(rb) - 1;
                        rb &= rb - 1;
                        item = (uint32_t)ls | ((uint32_t)(j * GRID_G + i) << 4);
                        has = true;
                        break;
                    }
                    j += parts;
                    if (j > wj1) { alive = false; break; }
                    rb = cg->rows[j] & cm & (cg->colfirst[j] | (1u << wi0)) & cg->rowfirst[j];  // (j != wj0 here)
                }
                const unsigned bm = __ballot_sync(FULL, has);
                if (!bm) break;
                if (has) W.list[cnt + __popc(bm & lanemask_lt())] = (uint16_t)item;
                cnt += __popc(bm);
            }
            __syncwarp();
            const bool more = __any_sync(FULL, alive);
            first = false;
            for (int p0 = 0; p0 < cnt; p0 += 32) {
                if constexpr (PIERS || !CG_TRIPLES) {
                if (p0 + lane < cnt) do {
                    const uint32_t pr = (uint32_t)W.list[p0 + lane];
                    const int s = (int)(pr & 15u);
                    const bool vpair = (pr & CG_PAIR_VNODE) != 0u;
                    uint32_t* bits = W.bits[s];
                    const float2* tv = W.tv + s * V;  // transformed vertices of query s
                    const int cell = (int)((pr >> 4) & 1023u);
                    const int ci = cell & (GRID_G - 1), cj = cell >> 5;
                    const int node = vpair ? W.vnode[s] : cg->cell_node[cell];
                    const int4 m = __ldg(&L.node_meta[node]);
                    const int n_all = (m.z & 0xFFFF) + ((m.z >> 16) & 0xFFFF);
                    {   // every hazard of this terminal collected already (through other terminals, or seeded)? nothing to do here
                        bool any_new = false;
#pragma unroll 1
                        for (int i = 0; i < n_all && !any_new; i++) any_new = !cg_has(bits, __ldg(&L.entry_meta[m.y + i].x));
                        if (!any_new) break;
                    }
                    int lt = 0, l0 = 0;
                    Rect tb{0.0f, 0.0f, 0.0f, 0.0f};
                    bool edge_ok = true;
                    unsigned pm = 0u;  // poles that reach this terminal
                    if (!vpair) {
                        lt = cg->cell_level[cell];
                        const int sh = D - lt;
                        const int ti = (ci >> sh) << sh, tj = (cj >> sh) << sh;  // first finest cell of the terminal
                        tb = Rect{g.xs[ti], g.ys[tj], g.xs[ti + (1 << sh)], g.ys[tj + (1 << sh)]};
                        // edges walk from the virtual root (cd_engine.rs:266-272): only terminals strictly below it, inside it
                        const int vc = W.vcell[s];
                        l0 = (vc >> 16) & 255;
                        const int s0 = D - l0;
                        edge_ok = lt > l0 && (ci >> s0) == (vc & 255) && (cj >> s0) == ((vc >> 8) & 255);
#pragma unroll 1
                        for (int pj = 0; pj < np; pj++) {  // circle.rs:115-128: a circle reaches a node iff its own box passes
                            const float4 c4 = W.pole[s][pj];
                            if (circ_rect(Circ{c4.x, c4.y, c4.z}, tb)) pm |= 1u << pj;
                        }
                    }
                    const float4 b4 = W.bb[s];
                    const Rect qbb{b4.x, b4.y, b4.z, b4.w};
                    const float dq = W.dq[s];
                    for (int e0 = 0; e0 < V; e0 += 32) {  // (one trip for polygons of up to 32 edges)
                        const int ne32 = (V - e0) < 32 ? (V - e0) : 32;
                        // candidate edges of this terminal: the edge's delta-inflated box meets the terminal's box (the candidate
                        // superset of detect_grid.cuh); epm: edges with an end point inside the terminal's box - reached for sure
                        unsigned cand = 0u, epm = 0u;
                        if (vpair) {
                            cand = ne32 == 32 ? 0xFFFFFFFFu : ((1u << ne32) - 1u);
                        } else if (edge_ok) {
                            float2 a = tv[e0];
                            bool a_in = rect_point(tb, a.x, a.y);
#pragma unroll 2
                            for (int k = 0; k < ne32; k++) {
                                const int e = e0 + k;
                                const float2 b = tv[e == V - 1 ? 0 : e + 1];
                                const bool b_in = rect_point(tb, b.x, b.y);
                                const bool out = fminf(a.x, b.x) - dq > tb.x_max || fmaxf(a.x, b.x) + dq < tb.x_min ||
                                                 fminf(a.y, b.y) - dq > tb.y_max || fmaxf(a.y, b.y) + dq < tb.y_min;
                                if (!out) cand |= 1u << k;
                                if (a_in || b_in) epm |= 1u << k;
                                a = b, a_in = b_in;
                            }
                        }
                        const unsigned pm_now = e0 == 0 ? pm : 0u;
                        if (!cand && !pm_now && !(PIERS && npr && e0 == 0 && !vpair)) continue;
                        unsigned rk = epm, rv = epm;  // per edge: reach known / reached (an endpoint inside: qt_traits.rs:36-44)
#pragma unroll 1
                        for (int i = 0; i < n_all; i++) {
                            const int4 em = __ldg(&L.entry_meta[m.y + i]);
                            if (cg_has(bits, em.x)) continue;  // collector.contains_key: skipping only saves work
                            const bool entire = em.w == 2;
                            bool ins = false;
                            Rect ff{0.0f, 0.0f, 0.0f, 0.0f};
                            if (!entire) ff = ld_rect(&L.entry_ff[m.y + i]);
                            // poles (cd_engine.rs:296-304): Entire -> inserted by any pole that reaches the terminal
                            if (entire) ins = pm_now != 0u;
                            else
                                for (unsigned pmm = pm_now; pmm && !ins; pmm &= pmm - 1) {
                                    const float4 c4 = W.pole[s][__ffs(pmm) - 1];
                                    const Circ c{c4.x, c4.y, c4.z};
                                    if (!circ_rect(c, ff)) continue;  // QTHazPartial::collides_with (qt_partial_hazard.rs:67-72)
#pragma unroll 1
                                    for (int k = 0; k < em.z && !ins; k++) ins = circ_seg(c, ld_seg(&L.edges[em.y + k]));
                                }
                            // edges: the bbox of the shape against ff_bbox first (exact: every edge lies inside the bbox)
                            if (entire && !ins) ins = (cand & epm) != 0u;  // some edge ends inside the terminal: reached, inserted
                            if (!ins && cand && (entire || rect_rect(qbb, ff))) {
                                for (unsigned cc = cand; cc && !ins; cc &= cc - 1) {
                                    const int k = __ffs(cc) - 1, e = e0 + k;
                                    const float2 a = tv[e], b = tv[e == V - 1 ? 0 : e + 1];
                                    const Seg sg{a.x, a.y, b.x, b.y};
                                    bool hit = entire;
                                    if (!hit && rect_seg(ff, sg)) {
#pragma unroll 1
                                        for (int h = 0; h < em.z && !hit; h++) hit = seg_seg(sg, ld_seg(&L.edges[em.y + h]));
                                    }
                                    if (!hit) continue;
                                    // would insert: is the terminal on the reference's walk of this edge? (cached per edge)
                                    if (!vpair && !((rk >> k) & 1u)) {
                                        const bool r = path_reaches<true>(g, sg, l0, lt, ci, cj);  // (no endpoint inside: epm)
                                        rk |= 1u << k;
                                        if (r) rv |= 1u << k;
                                    }
                                    ins = vpair || ((rv >> k) & 1u);
                                }
                            }
                            if constexpr (PIERS) {  // piers: edge entities from the ROOT (cd_engine.rs:305-308)
                                for (int pj = 0; pj < npr && !ins && !vpair && e0 == 0; pj++) {
                                    const float4 s4 = W.pier[s][pj];
                                    const Seg sg{s4.x, s4.y, s4.z, s4.w};
                                    const float amax = fmaxf(fmaxf(fabsf(sg.sx), fabsf(sg.ex)), fmaxf(fabsf(sg.sy), fabsf(sg.ey)));
                                    const float dp = fmaxf(g.delta, amax * (1.0f / 8192.0f));
                                    if (fminf(sg.sx, sg.ex) - dp > tb.x_max || fmaxf(sg.sx, sg.ex) + dp < tb.x_min ||
                                        fminf(sg.sy, sg.ey) - dp > tb.y_max || fmaxf(sg.sy, sg.ey) + dp < tb.y_min)
                                        continue;
                                    bool hit = entire;
                                    if (!hit && rect_seg(ff, sg))
                                        for (int h = 0; h < em.z && !hit; h++) hit = seg_seg(sg, ld_seg(&L.edges[em.y + h]));
                                    if (hit) ins = rect_point(tb, sg.sx, sg.sy) || rect_point(tb, sg.ex, sg.ey) || path_reaches<true>(g, sg, 0, lt, ci, cj);
                                }
                            }
                            if (ins) cg_set(bits, em.x);
                        }
                    }
                } while (0);
                } else {
                // LANE = (QUERY, TERMINAL, ENTRY). Step a, lane = pair: the terminal's box, the poles that reach it, the
                // candidate edges (once per pair), written to shared memory with a prefix sum of the pairs' entry counts. Step
                // b, lane = entry of some pair (found by a search in the prefix sums): that entry against the pair's poles and
                // candidate edges. The collector is a set: any order of the (entity, terminal, entry) triples gives the same
                // result, `contains_key` only skips work. With lane = pair the per-entry work (ff_bbox pre-rejects per pole
                // and candidate edge, hazard-edge loops) ran one entry after the other at 2 - 6 active lanes.
                bool active = false, vpair = false, edge_ok = true;
                int s = 0, my = 0, n_all = 0, lt = 0, l0 = 0, ci = 0, cj = 0;
                unsigned pm = 0u;  // poles that reach this terminal
                Rect tb{0.0f, 0.0f, 0.0f, 0.0f};
                if (p0 + lane < cnt) {
                    const uint32_t pr = (uint32_t)W.list[p0 + lane];
                    s = (int)(pr & 15u);
                    vpair = (pr & CG_PAIR_VNODE) != 0u;
                    const int cell = (int)((pr >> 4) & 1023u);
                    ci = cell & (GRID_G - 1), cj = cell >> 5;
                    const int4 m = __ldg(&L.node_meta[vpair ? W.vnode[s] : cg->cell_node[cell]]);
                    my = m.y;
                    n_all = (m.z & 0xFFFF) + ((m.z >> 16) & 0xFFFF);
                    // every hazard of this terminal collected already (through other terminals, or seeded)? nothing to do here
#pragma unroll 1
                    for (int i = 0; i < n_all && !active; i++) active = !cg_has(W.bits[s], __ldg(&L.entry_meta[my + i].x));
                    if (active && !vpair) {
                        lt = cg->cell_level[cell];
                        const int sh = D - lt;
                        const int ti = (ci >> sh) << sh, tj = (cj >> sh) << sh;  // first finest cell of the terminal
                        tb = Rect{g.xs[ti], g.ys[tj], g.xs[ti + (1 << sh)], g.ys[tj + (1 << sh)]};
                        // edges walk from the virtual root (cd_engine.rs:266-272): only terminals strictly below it, inside it
                        const int vc = W.vcell[s];
                        l0 = (vc >> 16) & 255;
                        const int s0 = D - l0;
                        edge_ok = lt > l0 && (ci >> s0) == (vc & 255) && (cj >> s0) == ((vc >> 8) & 255);
#pragma unroll 1
                        for (int pj = 0; pj < np; pj++) {  // circle.rs:115-128: a circle reaches a node iff its own box passes
                            const float4 c4 = W.pole[s][pj];
                            if (circ_rect(Circ{c4.x, c4.y, c4.z}, tb)) pm |= 1u << pj;
                        }
                    }
                }
                for (int e0 = 0; e0 < V; e0 += 32) {  // (one trip for polygons of up to 32 edges)
                    const int ne32 = (V - e0) < 32 ? (V - e0) : 32;
                    // candidate edges of this terminal: the edge's delta-inflated box meets the terminal's box (the candidate
                    // superset of detect_grid.cuh); epm: edges with an end point inside the terminal's box - reached for sure
                    unsigned cand = 0u, epm = 0u;
                    if (active) {
                        if (vpair) {
                            cand = ne32 == 32 ? 0xFFFFFFFFu : ((1u << ne32) - 1u);
                        } else if (edge_ok) {
                            const float2* tv = W.tv + s * V;
                            const float dq = W.dq[s];
                            float2 a = tv[e0];
                            bool a_in = rect_point(tb, a.x, a.y);
#pragma unroll 2
                            for (int k = 0; k < ne32; k++) {
                                const int e = e0 + k;
                                const float2 b = tv[e == V - 1 ? 0 : e + 1];
                                const bool b_in = rect_point(tb, b.x, b.y);
                                const bool out = fminf(a.x, b.x) - dq > tb.x_max || fmaxf(a.x, b.x) + dq < tb.x_min ||
                                                 fminf(a.y, b.y) - dq > tb.y_max || fmaxf(a.y, b.y) + dq < tb.y_min;
                                if (!out) cand |= 1u << k;
                                if (a_in || b_in) epm |= 1u << k;
                                a = b, a_in = b_in;
                            }
                        }
                    }
                    const unsigned pm_now = e0 == 0 ? pm : 0u;
                    const int n_mine = (active && (cand != 0u || pm_now != 0u)) ? n_all : 0;
                    int incl = n_mine;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const int tt = __shfl_up_sync(FULL, incl, o);
                        if (lane >= o) incl += tt;
                    }
                    const int T = __shfl_sync(FULL, incl, 31);
                    __syncwarp();  // (the previous trip is done with pp / pfx)
                    W.pp[lane] = make_int4(my, (int)cand, (int)epm,
                                           s | (vpair ? 16 : 0) | ((int)pm_now << 5) | (lt << 9) | (l0 << 12) | (ci << 15) | (cj << 20));
                    W.pfx[lane + 1] = incl;
                    if (lane == 0) W.pfx[0] = 0;
                    __syncwarp();
                    for (int t0 = 0; t0 < T; t0 += 32) {
                        const int t = t0 + lane;
                        if (t < T) {
                            int p = 0;  // the pair of triple t: the last one whose prefix is <= t
#pragma unroll
                            for (int o = 16; o > 0; o >>= 1)
                                if (W.pfx[p + o] <= t) p += o;
                            const int4 q4 = W.pp[p];
                            const int i = t - W.pfx[p];
                            const int4 em = __ldg(&L.entry_meta[q4.x + i]);
                            const int qs = q4.w & 15;
                            uint32_t* bits = W.bits[qs];
                            if (!cg_has(bits, em.x)) {  // collector.contains_key: skipping only saves work
                                const bool vp = (q4.w & 16) != 0;
                                const unsigned pmq = ((unsigned)q4.w >> 5) & 15u, cnd = (unsigned)q4.y, epq = (unsigned)q4.z;
                                const float2* tv = W.tv + qs * V;
                                const bool entire = em.w == 2;
                                bool ins = false;
                                Rect ff{0.0f, 0.0f, 0.0f, 0.0f};
                                if (!entire) ff = ld_rect(&L.entry_ff[q4.x + i]);
                                // poles (cd_engine.rs:296-304): Entire -> inserted by any pole that reaches the terminal
                                if (entire) ins = pmq != 0u;
                                else
                                    for (unsigned pmm = pmq; pmm && !ins; pmm &= pmm - 1) {
                                        const float4 c4 = W.pole[qs][__ffs(pmm) - 1];
                                        const Circ c{c4.x, c4.y, c4.z};
                                        if (!circ_rect(c, ff)) continue;  // QTHazPartial::collides_with (qt_partial_hazard.rs:67-72)
#pragma unroll 1
                                        for (int k = 0; k < em.z && !ins; k++) ins = circ_seg(c, ld_seg(&L.edges[em.y + k]));
                                    }
                                // edges: the bbox of the shape against ff_bbox first (exact: every edge lies inside the bbox)
                                if (entire && !ins) ins = (cnd & epq) != 0u;  // some edge ends inside the terminal: reached, inserted
                                if (!ins && cnd) {
                                    const float4 b4 = W.bb[qs];
                                    if (entire || rect_rect(Rect{b4.x, b4.y, b4.z, b4.w}, ff)) {
                                        for (unsigned cc = cnd; cc && !ins; cc &= cc - 1) {
                                            const int k = __ffs(cc) - 1, e = e0 + k;
                                            const float2 a = tv[e], b = tv[e == V - 1 ? 0 : e + 1];
                                            const Seg sg{a.x, a.y, b.x, b.y};
                                            bool hit = entire;
                                            if (!hit && rect_seg(ff, sg)) {
#pragma unroll 1
                                                for (int h = 0; h < em.z && !hit; h++) hit = seg_seg(sg, ld_seg(&L.edges[em.y + h]));
                                            }
                                            if (!hit) continue;
                                            // would insert: is the terminal on the reference's walk of this edge?
                                            ins = vp || ((epq >> k) & 1u) ||
                                                  path_reaches<true>(g, sg, (q4.w >> 12) & 7, (q4.w >> 9) & 7, (q4.w >> 15) & 31, (q4.w >> 20) & 31);
                                        }
                                    }
                                }
                                if (ins) cg_set(bits, em.x);
                            }
                        }
                        __syncwarp();
                    }
                }
                }
                __syncwarp();
            }
            if (!more) break;
        }
        __syncwarp();

        // ---- 4. (containment: k_collect_contain, launched behind this kernel) results out
        for (int i = lane; i < ng * words; i += 32) {
            const int s = i / words, wd = i - s * words;
            out_bits[(size_t)(qbase + s) * words + wd] = ((volatile uint32_t*)W.bits[s])[wd];
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------------
// k_collect_contain: cd_engine.rs:274-288 - for every Partial hazard of the virtual root that the edge and surrogate walks
// did not collect, CDEngine::detect_containment_collision (cd_engine.rs:231-255). A kernel of its own behind
// k_collect_grid (whose collectors it reads and completes in place): the point-in-polygon bodies and the FPA comparisons are
// a fifth of the collect path's instructions, and k_collect_grid - like stage 2 of the detect path - is bound by instruction
// fetch once its hot loop no longer fits the instruction cache (4.8 K instructions: 12 warps stalled on fetch per issue,
// profiles/r2c_*). The bbox and the virtual root of a query are recomputed here (a few per cent of the work) instead of being
// handed over through memory. Lane = (query, hazard) over the flat list of the group's candidate hazards; the rare
// point-in-polygon tests run one at a time with the whole warp (lanes over polygon edges).
// ---------------------------------------------------------------------------------------------------
#define CC_GROUP 8
struct CCSlots {
    float4 xf[CC_GROUP];
    float4 bb[CC_GROUP];
    int ent0[CC_GROUP];      // first Partial entry of the query's virtual root
    int ebeg[CC_GROUP + 1];  // flat index ranges of the queries' candidate hazards
    uint32_t bits[CC_GROUP][CG_MAX_WORDS];
};
__global__ void __launch_bounds__(S_THREADS, 4)
k_collect_contain(const DevLayout L, const __grid_constant__ SegTable gtab, const GridDev* __restrict__ ggrid,
                  const float4* __restrict__ xforms, uint32_t* __restrict__ out_bits, int words, int* __restrict__ next, int n_groups) {
    extern __shared__ __align__(16) unsigned char smem[];
    SegTable* t = (SegTable*)smem;
    for (int i = threadIdx.x; i < (int)(sizeof(SegTable) / 4); i += blockDim.x) ((int*)t)[i] = ((const int*)&gtab)[i];
    unsigned char* p = smem + ((sizeof(SegTable) + 15) & ~(size_t)15);
    const GridDev& g = *stage_grid(p, ggrid);
    __syncthreads();
    const SegTable& tab = *t;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    CCSlots& W = *(CCSlots*)(p + (size_t)warp * sizeof(CCSlots));
    const int D = g.D;
    const int quad = lane >> 2, sub = lane & 3;
    int seg_hint = 0;
    for (;;) {
        int gi = 0;
        if (lane == 0) gi = atomicAdd(next, 1);
        gi = __shfl_sync(FULL, gi, 0);
        if (gi >= n_groups) break;
        while (seg_hint + 1 < tab.n_seg && gi >= tab.seg[seg_hint + 1].gbeg) seg_hint++;
        const SegDesc& it = tab.seg[seg_hint];
        const int seg_end = seg_hint + 1 < tab.n_seg ? tab.seg[seg_hint + 1].begin : tab.n_queries;
        const int qbase = it.begin + (gi - it.gbeg) * CC_GROUP;
        const int ng = (seg_end - qbase) < CC_GROUP ? (seg_end - qbase) : CC_GROUP;
        const int V = it.n_verts;
        const float2* __restrict__ iverts = it.verts;
        if (lane < ng) W.xf[lane] = __ldg(&xforms[qbase + lane]);
        for (int i = lane; i < CC_GROUP * CG_MAX_WORDS; i += 32) {
            const int s = i / CG_MAX_WORDS, wd = i - s * CG_MAX_WORDS;
            (&W.bits[0][0])[i] = (s < ng && wd < words) ? __ldcg(&out_bits[(size_t)(qbase + s) * words + wd]) : 0u;
        }
        __syncwarp();
        // bbox (simple_polygon.rs:122-133) and virtual root (cd_engine.rs:314-327) again: quad = query, lane = child quadrant
        int n_cand = 0;
        {
            const int qs = quad;
            float x0 = 3.402823466e+38f, y0 = 3.402823466e+38f, x1 = -3.402823466e+38f, y1 = -3.402823466e+38f;
            if (qs < ng) {
                const float4 t4 = W.xf[qs];
                const Xf xf{t4.x, t4.y, t4.z, t4.w};
                for (int j = sub; j < V; j += 4) {
                    const float2 v = __ldg(&iverts[j]);
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
            bool walking = qs < ng;
            int level = 0, ci = 0, cj = 0, vnode = 0;
            for (;;) {
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
                const int4 m = __ldg(&L.node_meta[vnode]);
                W.bb[qs] = make_float4(x0, y0, x1, y1);
                W.ent0[qs] = m.y + (m.z & 0xFFFF);
                n_cand = (m.z >> 16) & 0xFFFF;
            }
        }
        // flat list of (query, Partial hazard of its virtual root): inclusive scan over the eight quad leaders
        int incl = __shfl_sync(FULL, n_cand, (lane & 7) * 4);  // lane k < 8 holds query k's count
        if (lane >= CC_GROUP) incl = 0;
#pragma unroll
        for (int o = 1; o < CC_GROUP; o <<= 1) {
            const int tt = __shfl_up_sync(FULL, incl, o);
            if (lane >= o) incl += tt;
        }
        if (lane < CC_GROUP) W.ebeg[lane + 1] = incl;
        if (lane == 0) W.ebeg[0] = 0;
        const int E = __shfl_sync(FULL, incl, CC_GROUP - 1);
        __syncwarp();
        for (int f0 = 0; f0 < E; f0 += 32) {
            const int f = f0 + lane;
            int hs = -1, s = 0, rel = REL_DISJOINT, scope_interior = 1;
            if (f < E) {
                if (W.ebeg[4] <= f) s = 4;
                if (W.ebeg[s + 2] <= f) s += 2;
                if (W.ebeg[s + 1] <= f) s += 1;
                hs = __ldg(&L.entry_meta[W.ent0[s] + (f - W.ebeg[s])]).x;
                if (cg_has(W.bits[s], hs)) hs = -1;  // cd_engine.rs:279
            }
            if (hs >= 0) {
                const float4 b4 = W.bb[s];
                scope_interior = __ldg(&L.haz_meta[hs]).x == 1 ? 1 : 0;
                rel = rect_almost_relation(ld_rect(&L.haz_bbox[hs]), Rect{b4.x, b4.y, b4.z, b4.w});
            }
            const bool pip = hs >= 0 && (rel == REL_SURROUNDING || rel == REL_ENCLOSED);
            // no point-in-polygon needed: contained = false -> collision iff the hazard's scope is Exterior (hazard.rs:55-62)
            if (hs >= 0 && !pip && !scope_interior) cg_set(W.bits[s], hs);
            unsigned todo = __ballot_sync(FULL, pip);
            while (todo) {
                const int src = __ffs(todo) - 1;
                todo &= todo - 1;
                const int h = __shfl_sync(FULL, hs, src), r = __shfl_sync(FULL, rel, src), si = __shfl_sync(FULL, scope_interior, src);
                const int ss = __shfl_sync(FULL, s, src);
                const float4 t4 = W.xf[ss], b4 = W.bb[ss];
                const Xf xf{t4.x, t4.y, t4.z, t4.w};
                bool contained;
                if (r == REL_SURROUNDING) {  // the hazard's shape against the item's (transformed) poi
                    const int4 hm = __ldg(&L.haz_meta[h]);
                    float px, py;
                    xf_point(xf, it.poi_x, it.poi_y, px, py);
                    contained = poly_point_warp<false>(L.haz_verts + hm.y, hm.z, ld_rect(&L.haz_bbox[h]), px, py, xf, lane);
                } else {  // the item's transformed shape against the hazard's poi
                    const float4 hp = __ldg(&L.haz_poi[h]);
                    contained = poly_point_warp<true>(iverts, V, Rect{b4.x, b4.y, b4.z, b4.w}, hp.x, hp.y, xf, lane);
                }
                if (lane == src && (si ? contained : !contained)) cg_set(W.bits[ss], h);
            }
            __syncwarp();
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
