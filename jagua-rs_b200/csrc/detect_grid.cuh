// Grid form of the detect pipeline (sm_100a) for quadtrees of depth <= 5 without surrogate piers
// (CDEConfig defaults: depth 5, lbf/src/config.rs:33-41). Included by detect.cu.
//
// QTNode::collides (qt_node.rs:99-147) is a depth-first walk, but its verdict is an OR over the TERMINAL
// nodes it reaches: nodes whose strongest hazard (qt_hazard_vec.rs:60-63) is Entire (-> hit, qt_node.rs:108)
// or Partial without children (-> edge tests, qt_node.rs:128-143). With depth D <= 5 the terminals tile a
// 2^D x 2^D grid of finest cells whose boundaries xs[], ys[] are the nested midpoints of rect.rs:141-151 /
// :248-253 (k_grid_build recomputes them and checks every node bbox against them bit for bit). Per
// (layout, filter) the grid keeps, per row of finest cells, the bit masks of cells covered by an Entire
// terminal and of leaf cells, plus per cell its terminal's level and leaf entry range. Both stages then go
// straight to the terminals that can matter instead of walking the inner nodes:
//
//  circles (surrogate poles, cd_engine.rs:204-223). Circle-Rect (circle.rs:115-128) is monotone under
//      nesting in f32: clamping to a larger interval never moves the clamped point away, and fl(-), fl(x*x),
//      fl(+) and <= are monotone. Hence a node is reached iff its own bbox test passes (all ancestors then
//      pass too), and a coarse Entire terminal passes iff one of its finest cells does (the cell holding the
//      clamped point yields the same clamped point). So: hit iff some Entire-flagged finest cell passes
//      circ_rect (per row only the flagged cells nearest to the centre column need testing, again by
//      monotonicity) or some leaf cell passes circ_rect and QTHazPartial::collides_with. EXACT, no walk.
//
//  edges (polygon stage, cd_engine.rs:160-197). Edge::collides_with_quadrants (qt_traits.rs:19-106) is not
//      monotone, so terminals are CANDIDATES that are verified: a terminal T matters only if it
//      would produce a hit (Entire, or a leaf whose edge tests succeed - both independent of the path), and
//      then T is reached by the reference iff every ancestor below the virtual root flags the quadrant
//      towards T: at most D seg_quadrants evaluations, run only for would-be hits. Candidates = finest cells
//      meeting the edge's bbox inflated by delta, clipped to the virtual root's cells. That is a superset of
//      the reached terminals: a quadrant is flagged either because it holds an endpoint, or because
//      edge_intersection_half (qt_traits.rs:144-177) reports a crossing with a side/bisector half lying on
//      that quadrant's boundary. Sides and bisectors are axis-aligned, so one product of t_nom and of denom is
//      exactly zero and t in [0,1] places the side's coordinate within 5 ulp-relative of the edge's own
//      range, while the crossing's position along the side is off by < 2^-20 of the coordinate scale M;
//      delta = M * 2^-13 exceeds that bound by two orders of magnitude.
#pragma once

namespace jg {

#define GRID_MAX_D 5
#define GRID_G 32            // row stride of the per-cell arrays (finest cells per side <= 32)
#define GRID_MIP_NODES 1365  // 1 + 4 + ... + 4^5

struct GridDev {
    int ok;       // 1 = usable (depth <= 5, node boxes match the grid, leaf records fit)
    int D, G, n_nodes;
    float delta;  // candidate inflation for edges
    float x0, y0, inv_cw, inv_ch;
    float slack;  // absolute slack added to circle windows
    int pad[2];
    float xs[GRID_G + 4], ys[GRID_G + 4];
    uint32_t ent_rows[GRID_G], leaf_rows[GRID_G];
    uint32_t cell_leaf[GRID_G * GRID_G];   // leaf cells: first Partial entry << 8 | n_partial
    uint16_t mip_node[GRID_MIP_NODES + 3];  // node index at (level, j, i), 0xFFFF = no such node
    uint8_t cell_level[GRID_G * GRID_G];   // level (0 = root) of the terminal covering the cell
};

__host__ __device__ inline int mip_off(int level) { return ((1 << (2 * level)) - 1) / 3; }
// quadrant index (NE, NW, SW, SE; rect.rs:141-161) of a child from the low bits of its column / row
__host__ __device__ inline int quad_of(int dx, int dy) { return dy ? (dx ? 0 : 1) : (dx ? 3 : 2); }

// arithmetic guess of the cell holding coordinate v (see cell_lo / cell_hi)
__device__ __forceinline__ int cell_guess(float v, float b0, float inv) { return __float2int_rd((v - b0) * inv); }

// One CTA of 1024 threads per (layout, filter): fills GridDev from the node arrays and the kinds summary.
__global__ void __launch_bounds__(1024) k_grid_build(const DevLayout L, int depth, const uint8_t* __restrict__ kinds,
                                                      GridDev* __restrict__ g) {
    __shared__ int s_ok;
    const int tid = threadIdx.x;
    const int D = depth, G = 1 << (depth <= GRID_MAX_D ? depth : 0);
    if (tid == 0) {
        s_ok = (depth >= 0 && depth <= GRID_MAX_D) ? 1 : 0;
        g->D = D;
        g->G = G;
        g->n_nodes = L.n_nodes;
        const float4 b = L.node_bbox[0];
        g->xs[0] = b.x, g->xs[G] = b.z, g->ys[0] = b.y, g->ys[G] = b.w;
        for (int step = G; step >= 2; step >>= 1)
            for (int a = 0; a < G; a += step) {
                g->xs[a + step / 2] = (g->xs[a] + g->xs[a + step]) / 2.0f;  // rect.rs:248-253 centroid
                g->ys[a + step / 2] = (g->ys[a] + g->ys[a + step]) / 2.0f;
            }
        const float M = fmaxf(fmaxf(fabsf(b.x), fabsf(b.z)), fmaxf(fabsf(b.y), fabsf(b.w)));
        g->delta = M * (1.0f / 8192.0f);
        g->slack = M * (1.0f / 1048576.0f);
        g->x0 = b.x, g->y0 = b.y;
        g->inv_cw = (float)G / (b.z - b.x);
        g->inv_ch = (float)G / (b.w - b.y);
        if (!(b.z > b.x && b.w > b.y)) s_ok = 0;
        for (int k = 0; k <= G && s_ok; k++) {  // the premise of cell_lo / cell_hi
            const int gx = cell_guess(g->xs[k], g->x0, g->inv_cw), gy = cell_guess(g->ys[k], g->y0, g->inv_ch);
            if (gx < k - 1 || gx > k || gy < k - 1 || gy > k) s_ok = 0;
            if (k > 0 && !(g->xs[k] > g->xs[k - 1] && g->ys[k] > g->ys[k - 1])) s_ok = 0;
        }
    }
    for (int i = tid; i < GRID_G; i += blockDim.x) g->ent_rows[i] = 0u, g->leaf_rows[i] = 0u;
    for (int i = tid; i < GRID_G * GRID_G; i += blockDim.x) g->cell_leaf[i] = 0u, g->cell_level[i] = 0;
    for (int i = tid; i < GRID_MIP_NODES + 3; i += blockDim.x) g->mip_node[i] = 0xFFFF;
    __syncthreads();
    if (!s_ok) {
        if (tid == 0) g->ok = 0;
        return;
    }
    // every node: (level, column, row) from its chain of parents; its bbox must be the grid's
    for (int n = tid; n < L.n_nodes; n += blockDim.x) {
        if (L.node_meta[n].w == -2) continue;  // a record of a freed child block (incremental image): not part of the tree
        int level = 0, ci = 0, cj = 0, cur = n;
        bool bad = n > 0xFFFE;
        while (cur > 0 && level <= D) {
            const int q = (cur - 1) & 3;  // children blocks start at 1 + 4k
            ci |= ((q == 0 || q == 3) ? 1 : 0) << level;
            cj |= ((q == 0 || q == 1) ? 1 : 0) << level;
            level++;
            const int par = L.node_meta[cur].w;
            if (par < 0 || L.node_meta[par].x != ((cur - 1) & ~3) + 1) { bad = true; break; }
            cur = par;
        }
        if (level > D || cur != 0) bad = true;
        if (!bad) {
            const int s = D - level;
            const float4 b = L.node_bbox[n];
            if (!(b.x == g->xs[ci << s] && b.z == g->xs[(ci + 1) << s] && b.y == g->ys[cj << s] && b.w == g->ys[(cj + 1) << s])) bad = true;
            if (!bad) g->mip_node[mip_off(level) + cj * (1 << level) + ci] = (uint16_t)n;
        }
        if (bad) atomicExch(&s_ok, 0);
    }
    __syncthreads();
    // every finest cell: the terminal the reference's descent from the root ends in
    for (int c = tid; c < G * G; c += blockDim.x) {
        const int i = c % G, j = c / G;
        int node = 0, level = 0;
        while (kinds[node + 3] == 1 /*KIND_INNER*/ && level < D) {
            level++;
            const int ci = i >> (D - level), cj = j >> (D - level);
            node = L.node_meta[node].x + quad_of(ci & 1, cj & 1);
        }
        const int k = kinds[node + 3];
        g->cell_level[j * GRID_G + i] = (uint8_t)level;
        if (k == 2 /*KIND_ENTIRE*/) atomicOr(&g->ent_rows[j], 1u << i);
        if (k == 3 /*KIND_LEAF*/) {
            const int4 m = L.node_meta[node];
            const int n_entire = m.z & 0xFFFF, n_partial = (m.z >> 16) & 0xFFFF;
            if (level != D || n_partial > 255 || m.y + n_entire >= (1 << 24)) atomicExch(&s_ok, 0);
            g->cell_leaf[j * GRID_G + i] = ((uint32_t)(m.y + n_entire) << 8) | (uint32_t)(n_partial & 255);
            atomicOr(&g->leaf_rows[j], 1u << i);
        }
        if (k == 1) atomicExch(&s_ok, 0);  // an inner node at the finest level cannot exist
    }
    __syncthreads();
    if (tid == 0) g->ok = s_ok;
}

// ---- staged copy -------------------------------------------------------------------------------------
__device__ __forceinline__ const GridDev* stage_grid(unsigned char*& p, const GridDev* g) {
    uint4* dst = (uint4*)p;
    const uint4* src = (const uint4*)g;
    const int n16 = (int)((sizeof(GridDev) + 15) / 16);
    for (int i = threadIdx.x; i < n16; i += blockDim.x) dst[i] = __ldg(src + i);
    p += (size_t)n16 * 16;
    return (const GridDev*)dst;
}

// first cell whose upper boundary is >= v / last cell whose lower boundary is <= v (clamped to the grid).
// The arithmetic guess floor((v - b0) * inv) is within one cell of the answer - k_grid_build checks, for every
// boundary b[k], that the guess there is k - 1 or k; the guess is monotone in v, so inside cell k it is
// k - 1, k or k + 1 - and one exact comparison against each neighbouring boundary settles it.
__device__ __forceinline__ int cell_lo(const float* b, int G, float v, float b0, float inv) {
    const int g = min(max(cell_guess(v, b0, inv), 0), G - 1);
    return g - ((g > 0 && b[g] >= v) ? 1 : 0) + ((g < G - 1 && b[g + 1] < v) ? 1 : 0);
}
__device__ __forceinline__ int cell_hi(const float* b, int G, float v, float b0, float inv) {
    const int g = min(max(cell_guess(v, b0, inv), 0), G - 1);
    return g + ((g < G - 1 && b[g + 1] <= v) ? 1 : 0) - ((g > 0 && b[g] > v) ? 1 : 0);
}
__device__ __forceinline__ uint32_t col_mask(int lo, int hi) {  // bits lo..hi, 0 <= lo, hi <= 31
    return hi < lo ? 0u : ((0xFFFFFFFFu >> (31 - hi)) & (0xFFFFFFFFu << lo));
}

// QTHazPartial::collides_with (qt_partial_hazard.rs:67-72) over the relevant Partial entries of a leaf cell
// SM: the entry and edge sections were staged into shared memory (k_s2_grid<.., STAGE>): plain loads instead of ld.global.nc
template <class Ent, bool FILT, bool SM = false>
__device__ __forceinline__ bool leaf_cell_collides(const DevLayout& L, uint32_t rec, const Ent& ent, const uint32_t* filt) {
    const int e0 = (int)(rec >> 8), n_partial = (int)(rec & 255u);
    for (int i = 0; i < n_partial; i++) {
        const int4 em = SM ? L.entry_meta[e0 + i] : __ldg(&L.entry_meta[e0 + i]);
        if (FILT && filt_bit(filt, em.x)) continue;
        if (!ent_rect(ent, SM ? ld_rect_any(&L.entry_ff[e0 + i]) : ld_rect(&L.entry_ff[e0 + i]))) continue;
        for (int k = 0; k < em.z; k++) {
            const float4 v = SM ? L.edges[em.y + k] : __ldg(&L.edges[em.y + k]);
            if (ent_seg(ent, Seg{v.x, v.y, v.z, v.w})) return true;
        }
    }
    return false;
}

// ---- TMA bulk staging (cp.async.bulk + mbarrier, sm_90+): one elected thread copies whole sections of the layout blob into
// the CTA's shared memory; every thread waits on the barrier's phase. BASELINE north_star item (3) "hazard edges staged through
// TMA or shared memory"; data shape: qt_partial_hazard.rs:7-12 (edges + ff_bbox per QTHazPartial).
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
                 "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!ok);
}

// Does the circle pass circ_rect (circle.rs:115-128) for some Entire-flagged finest cell? Rows of the
// conservative window; per row the flagged cells nearest to the centre's column on either side.
__device__ __forceinline__ bool circ_touches_entire(const GridDev& g, const Circ& c, int i0, int i1, int j0, int j1) {
    const int G = g.G;
    const uint32_t cm = col_mask(i0, i1);
    // column holding the centre (any column whose closed range holds cx; outermost column if outside)
    const int ic = cell_lo(g.xs, G, c.cx, g.x0, g.inv_cw);
    const uint32_t below = (1u << ic) - 1u;
    const float r2 = c.r * c.r;
    bool hit = false;
    for (int j = j0; j <= j1; j++) {  // branch-free body: the lanes of a warp run nearly the same row counts
        const uint32_t m = g.ent_rows[j] & cm;
        const float ny = fmaxf(g.ys[j], fminf(c.cy, g.ys[j + 1]));
        const float dy2 = (ny - c.cy) * (ny - c.cy);
        const uint32_t right = m >> ic, left = m & below;
        const int ir = right ? ic + __ffs(right) - 1 : ic;
        const int il = left ? 31 - __clz(left) : ic;
        const float nxr = fmaxf(g.xs[ir], fminf(c.cx, g.xs[ir + 1]));
        const float nxl = fmaxf(g.xs[il], fminf(c.cx, g.xs[il + 1]));
        const bool hr = (nxr - c.cx) * (nxr - c.cx) + dy2 <= r2, hl = (nxl - c.cx) * (nxl - c.cx) + dy2 <= r2;
        hit = hit || (right != 0u && hr) || (left != 0u && hl);
    }
    return hit;
}

#define G1_PAIR_CAP 192  // leaf pairs waiting per warp: drained below 32 before every push round of <= 4 per lane
#define G1_PUSH 4
#define G1_SURV_CAP 40  // local survivors: fewer than 8 left over + up to 32 new ones
#define G1_WARP_BYTES (32 * sizeof(float4) + (G1_PAIR_CAP + 64 + 4 + G1_SURV_CAP) * sizeof(uint32_t))
#define GF_WARP_BYTES (G1_WARP_BYTES + ((sizeof(G2Slots) + 15) & ~(size_t)15))

// ---------------------------------------------------------------------------------------------------
// stage 1 on the grid: surrogate poles. Lane = query.
//   phase A (every query): do the poles touch an Entire terminal? Uniform, cheap, decides ~60 % of a uniform
//           batch. Undecided queries go to a warp-local queue.
//   phase B (32 queued queries at a time, so the lanes are dense again): the leaf cells inside each pole's
//           window that pass circ_rect become (query, cell) pairs, resolved 32 at a time.
// ---------------------------------------------------------------------------------------------------
struct PoleWin {
    int i0, i1, j0, j1;
};
__device__ __forceinline__ PoleWin pole_window(const GridDev& g, const Circ& c) {
    const int G = g.G;
    const float rk = c.r * 1.001f + g.slack;
    PoleWin w;
    w.i0 = cell_lo(g.xs, G, c.cx - rk, g.x0, g.inv_cw);
    w.i1 = cell_hi(g.xs, G, c.cx + rk, g.x0, g.inv_cw);
    w.j0 = cell_lo(g.ys, G, c.cy - rk, g.y0, g.inv_ch);
    w.j1 = cell_hi(g.ys, G, c.cy + rk, g.y0, g.inv_ch);
    // a window entirely beside the grid holds no cell the circle can reach
    if (c.cx + rk < g.xs[0] || c.cx - rk > g.xs[G] || c.cy + rk < g.ys[0] || c.cy - rk > g.ys[G]) w.i1 = -1, w.j1 = -1;
    return w;
}

// what every warp-level routine of the grid kernels needs
struct GCtx {
    const DevLayout& L;   // (k_s2_grid<.., STAGE>: entry_meta / entry_ff / edges point into shared memory - leaf tests only)
    const SegTable& tab;
    const GridDev& g;
    const XfSrc& xforms;
    const uint32_t* filt;
    const uint8_t* g_kinds;
    uint8_t* out;
    DetectCtl* ctl;
    int* sv_entry;
    int4* ct_meta = nullptr;   // containment list (k_s2_grid -> k_s2_contain): {query | segment, first entry, n_partial, -}
    float4* ct_bb = nullptr;   // ... the transformed shape's bbox
    float4* ct_xf = nullptr;   // ... and the query's transformation
    const int4* g_entry_meta = nullptr;  // the entry section in global memory (for __ldg readers next to a staged L)
};
struct G1Warp {  // stage-1 state of a warp (shared memory + two counters)
    float4* s_circ;    // phase B: the current pole of each lane's query
    uint32_t* s_pair;  // (query lane, leaf cell) pairs waiting for their edge tests
    int* s_pend;       // queries waiting for phase B
    unsigned* s_hit;
    int* s_cnt;
    int* s_surv;       // fused kernel: this warp's survivors waiting for stage 2
    int n_pend, seg_hint, n_sv;
    __device__ __forceinline__ void bind(unsigned char* wp) {
        s_circ = (float4*)wp;
        s_pair = (uint32_t*)(wp + 32 * sizeof(float4));
        s_pend = (int*)(s_pair + G1_PAIR_CAP);
        s_hit = (unsigned*)(s_pend + 64);
        s_cnt = (int*)(s_hit + 1);
        s_surv = s_cnt + 3;
        n_pend = 0, seg_hint = 0, n_sv = 0;
    }
};

// phase A on the 32 queries starting at `base` (base < n_queries)
template <bool FILT>
__device__ __forceinline__ void g1_phase_a(const GCtx& c, G1Warp& w, int base, int lane) {
    const int q = base + lane;
    const bool valid = q < c.tab.n_queries;
    bool hit = false;
    // batches are claimed in increasing order: the segment of `base` is at or after the last one seen
    while (w.seg_hint + 1 < c.tab.n_seg && base >= c.tab.seg[w.seg_hint + 1].begin) w.seg_hint++;
    int k = w.seg_hint;
    if (valid) {
        while (k + 1 < c.tab.n_seg && q >= c.tab.seg[k + 1].begin) k++;
        const SegDesc& sd = c.tab.seg[k];
        const float4 t4 = load_xf(c.xforms, sd, q);
        const Xf xf{t4.x, t4.y, t4.z, t4.w};
        for (int j = 0; j < sd.n_poles && !hit; j++) {  // circle.rs:79-85: centre transformed, radius unchanged
            const float4 pl = __ldg(&sd.poles[j]);
            Circ cc;
            xf_point(xf, pl.x, pl.y, cc.cx, cc.cy);
            cc.r = pl.z;
            const PoleWin pw = pole_window(c.g, cc);
            hit = circ_touches_entire(c.g, cc, pw.i0, pw.i1, pw.j0, pw.j1);
        }
        if (hit) c.out[q] = (uint8_t)JG_VERDICT_SURROGATE;
    }
    const unsigned um = __ballot_sync(FULL, valid && !hit);
    if (valid && !hit) w.s_pend[w.n_pend + __popc(um & lanemask_lt())] = sv_pack(q, k);
    w.n_pend += __popc(um);
    __syncwarp();
}

// Surrogate piers on the grid (cd_engine.rs:216-221: after the poles, from the ROOT): defined next to path_reaches below.
template <bool FILT>
__device__ bool seg_collides_from_root(const GCtx& c, const Seg& sg);

// phase B on up to 32 queued queries; survivors go to the survivor list (FUSED: published in order). PIERS: items whose
// surrogate carries piers test them after the poles, lane = query (piers are few and off in the reference's defaults).
template <bool FILT, bool FUSED, bool PIERS = false>
__device__ __forceinline__ void g1_phase_b(const GCtx& c, G1Warp& w, int lane) {
    const int take = w.n_pend < 32 ? w.n_pend : 32;
    w.n_pend -= take;
    const bool valid = lane < take;
    const int qk = valid ? w.s_pend[w.n_pend + lane] : 0;
    const int q = qk & 0x3FFFFFF, k = (int)((unsigned)qk >> 26);
    Xf xf{1.0f, 0.0f, 0.0f, 0.0f};
    int n_poles = 0;
    if (valid) {
        const float4 t4 = load_xf(c.xforms, c.tab.seg[k], q);
        xf = Xf{t4.x, t4.y, t4.z, t4.w};
        n_poles = c.tab.seg[k].n_poles;
    }
    bool hit = false;
    const int max_poles = __reduce_max_sync(FULL, n_poles);
#ifdef G1_SKIP_B
    const int max_poles_dbg = 0;
#else
    const int max_poles_dbg = max_poles;
#endif
    for (int j = 0; j < max_poles_dbg; j++) {  // cd_engine.rs:210-215: any colliding pole decides
        bool act = valid && !hit && j < n_poles;
        if (!__any_sync(FULL, act)) break;
        Circ cc{0.0f, 0.0f, 0.0f};
        PoleWin pw{0, -1, 0, -1};
        if (act) {
            const float4 pl = __ldg(&c.tab.seg[k].poles[j]);
            xf_point(xf, pl.x, pl.y, cc.cx, cc.cy);
            cc.r = pl.z;
            w.s_circ[lane] = make_float4(cc.cx, cc.cy, cc.r, 0.0f);
            pw = pole_window(c.g, cc);
            act = pw.j1 >= pw.j0 && pw.i1 >= pw.i0;
        }
        if (!__any_sync(FULL, act)) continue;
        if (lane == 0) *w.s_hit = 0u, *w.s_cnt = 0;
        __syncwarp();
        const uint32_t cm = col_mask(pw.i0, pw.i1);
        const float r2 = cc.r * cc.r;
        int row = pw.j0;
        uint32_t bits = 0u;
        float dy2 = 0.0f;
        bool more = act;
        if (more) {
            bits = c.g.leaf_rows[row] & cm;
            const float ny = fmaxf(c.g.ys[row], fminf(cc.cy, c.g.ys[row + 1]));
            dy2 = (ny - cc.cy) * (ny - cc.cy);
        }
        for (;;) {
            if (more) {  // generation round: up to G1_PUSH cells per lane
                int pushed = 0;
                uint32_t mine[G1_PUSH];
                while (pushed < G1_PUSH) {
                    if (!bits) {
                        if (row >= pw.j1) { more = false; break; }
                        row++;
                        bits = c.g.leaf_rows[row] & cm;
                        const float ny = fmaxf(c.g.ys[row], fminf(cc.cy, c.g.ys[row + 1]));
                        dy2 = (ny - cc.cy) * (ny - cc.cy);
                        continue;
                    }
                    const int i = __ffs(bits) - 1;
                    bits &= bits - 1;
                    const float nx = fmaxf(c.g.xs[i], fminf(cc.cx, c.g.xs[i + 1]));
                    if ((nx - cc.cx) * (nx - cc.cx) + dy2 <= r2) mine[pushed++] = ((uint32_t)(row * GRID_G + i) << 5) | (uint32_t)lane;
                }
                if (pushed) {
                    const int o = atomicAdd(w.s_cnt, pushed);
#pragma unroll
                    for (int u = 0; u < G1_PUSH; u++)
                        if (u < pushed) w.s_pair[o + u] = mine[u];
                }
            }
            __syncwarp();
            const bool any_more = __any_sync(FULL, more);
            int cnt = *(volatile int*)w.s_cnt;
            __syncwarp();
            while (cnt >= 32 || (!any_more && cnt > 0)) {  // drain
                const int tk = cnt < 32 ? cnt : 32;
                cnt -= tk;
                int ol = 0;
                uint32_t rec = 0u;
                float4 e4 = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                if (lane < tk) {
                    const uint32_t pr = w.s_pair[cnt + lane];
                    ol = pr & 31;
                    if (!((*(volatile unsigned*)w.s_hit >> ol) & 1u)) {
                        e4 = w.s_circ[ol];
                        rec = c.g.cell_leaf[pr >> 5];
                    }
                }
                const bool lanehit = rec != 0u && leaf_cell_collides<Circ, FILT>(c.L, rec, Circ{e4.x, e4.y, e4.z}, c.filt);
                if (lanehit) atomicOr(w.s_hit, 1u << ol);
                __syncwarp();
            }
            if (lane == 0) *w.s_cnt = cnt;
            __syncwarp();
            if ((*(volatile unsigned*)w.s_hit >> lane) & 1u) hit = true, more = false;
            if (!any_more) break;
            __syncwarp();
        }
        __syncwarp();
    }
    if (PIERS && valid && !hit) {
        const SegDesc& sd = c.tab.seg[k];
        for (int j = 0; j < sd.n_piers && !hit; j++) {  // edge.rs:126-144: both end points transformed
            const float4 pr = __ldg(&sd.piers[j]);
            Seg sg;
            xf_point(xf, pr.x, pr.y, sg.sx, sg.sy);
            xf_point(xf, pr.z, pr.w, sg.ex, sg.ey);
            hit = seg_collides_from_root<FILT>(c, sg);
        }
    }
    __syncwarp();
    if (valid) c.out[q] = hit ? (uint8_t)JG_VERDICT_SURROGATE : (uint8_t)0;
    const unsigned sm = __ballot_sync(FULL, valid && !hit);
    if (sm) {
        if (FUSED) {  // the warp resolves its own survivors (k_detect_fused)
            if (valid && !hit) w.s_surv[w.n_sv + __popc(sm & lanemask_lt())] = sv_pack(q, k);
            w.n_sv += __popc(sm);
        } else {
            int sb = 0;
            if (lane == 0) sb = atomicAdd(&c.ctl->n_surv, __popc(sm));
            sb = __shfl_sync(FULL, sb, 0);
            if (valid && !hit) c.sv_entry[sb + __popc(sm & lanemask_lt())] = sv_pack(q, k);
        }
    }
    __syncwarp();
}

#ifndef G1_MIN_BLOCKS
#define G1_MIN_BLOCKS 5  // 48 registers: 40 resident warps per SM for the latency-bound stage 1 (4 CTAs: 214 us per 1M queries, 5: 212, 6: 226)
#endif
#ifndef G2_MIN_BLOCKS
#define G2_MIN_BLOCKS 4
#endif
template <bool FILT, bool PIERS = false>
__global__ void __launch_bounds__(S_THREADS, G1_MIN_BLOCKS)
k_s1_grid(const DevLayout L, const __grid_constant__ SegTable gtab, const GridDev* __restrict__ ggrid,
          const XfSrc xforms, const uint32_t* __restrict__ filt, uint8_t* __restrict__ out,
          DetectCtl* __restrict__ ctl, int* __restrict__ sv_entry) {
    // a stage-2 grid launched with programmatic stream serialization may take the SM slots this grid's CTAs free as they
    // exit (every CTA of this persistent grid is resident from the start, so nothing of it can be locked out); it reads
    // nothing of ours before its griddepcontrol.wait. A no-op for ordinary launches.
    asm volatile("griddepcontrol.launch_dependents;");
    extern __shared__ __align__(16) unsigned char smem[];
    SegTable* t = (SegTable*)smem;
    for (int i = threadIdx.x; i < (int)(sizeof(SegTable) / 4); i += blockDim.x) ((int*)t)[i] = ((const int*)&gtab)[i];
    unsigned char* p = smem + ((sizeof(SegTable) + 15) & ~(size_t)15);
    const GridDev& g = *stage_grid(p, ggrid);
    __syncthreads();
    const GCtx c{L, *t, g, xforms, filt, nullptr, out, ctl, sv_entry};
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    G1Warp w;
    w.bind(p + (size_t)warp * G1_WARP_BYTES);
    const int n = c.tab.n_queries;
    bool feeding = true;
    for (;;) {
        if (feeding && w.n_pend < 32) {
            int base = 0;
            if (lane == 0) base = atomicAdd(&ctl->s1_next, 32);
            base = __shfl_sync(FULL, base, 0);
            if (base >= n) feeding = false;
            else g1_phase_a<FILT>(c, w, base, lane);
            continue;
        }
        if (w.n_pend == 0) break;
        g1_phase_b<FILT, false, PIERS>(c, w, lane);
    }
}

// SPolygon::collides_with(Point) (simple_polygon.rs:255-292) by a whole warp: lanes take polygon edges, the
// crossing count is the sum of the per-edge contributions (its parity does not depend on the order).
template <bool XF>
__device__ __forceinline__ bool poly_point_warp(const float2* verts, int n, const Rect& bbox, float px, float py, const Xf& t,
                                                int lane) {
    if (!rect_point(bbox, px, py)) return false;
    const Seg ray{px, py, bbox.x_max + (bbox.x_max - bbox.x_min), py};
    int n_int = 0;
    for (int b = 0; b < n; b += 32) {
        const int i = b + lane;
        bool cross = false;
        if (i < n) {
            const float2 a = __ldg(&verts[i]), bb = __ldg(&verts[i == n - 1 ? 0 : i + 1]);
            float sx = a.x, sy = a.y, ex = bb.x, ey = bb.y;
            if (XF) {
                xf_point(t, a.x, a.y, sx, sy);
                xf_point(t, bb.x, bb.y, ex, ey);
            }
            if ((approx_eq(sy, py) && fpa_gt(sx, px)) || (approx_eq(ey, py) && fpa_gt(ex, px))) cross = fpa_lt(sy, py) || fpa_lt(ey, py);
            else cross = seg_seg(ray, Seg{sx, sy, ex, ey});
        }
        n_int += __popc(__ballot_sync(FULL, cross));
    }
    return (n_int & 1) == 1;
}

// CDEngine::detect_containment_collision (cd_engine.rs:231-255) for up to 32 hazards at once: lane l holds
// hazard slot hs (or -1). The bbox relations are evaluated per lane; the (rare) point-in-polygon tests run one
// hazard at a time with the whole warp. Returns the OR over the lanes' hazards, uniform across the warp.
__device__ __forceinline__ bool containment_collision_warp(const DevLayout& L, const float2* iverts, int n_iverts, float ipoi_x,
                                                           float ipoi_y, const Xf& xf, const Rect& sb, int hs, int lane) {
    int rel = REL_DISJOINT, scope_interior = 1;
    if (hs >= 0) {
        scope_interior = __ldg(&L.haz_meta[hs]).x == 1 ? 1 : 0;
        rel = rect_almost_relation(ld_rect(&L.haz_bbox[hs]), sb);
    }
    const bool pip = hs >= 0 && (rel == REL_SURROUNDING || rel == REL_ENCLOSED);
    // no point-in-polygon needed: contained = false -> collision iff the hazard's scope is Exterior (hazard.rs:55-62)
    bool c = hs >= 0 && !pip && !scope_interior;
    unsigned todo = __ballot_sync(FULL, pip);
    if (__any_sync(FULL, c)) return true;
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
        if (si ? contained : !contained) return true;
    }
    return false;
}

// ---------------------------------------------------------------------------------------------------
// stage 2 on the grid: polygon edges of the survivors
// ---------------------------------------------------------------------------------------------------
struct G2Slots {  // per-warp state of the survivors of the current group
    float4 xf[S2_MAX_GROUP];
    float4 bb[S2_MAX_GROUP];
    int q[S2_MAX_GROUP];
    int seg[S2_MAX_GROUP];
    int vnode[S2_MAX_GROUP];  // virtual root: node index
    int vcell[S2_MAX_GROUP];  // ... and level << 16 | row << 8 | column at its level
    int ebeg[S2_MAX_GROUP + 1];
    unsigned hit;
    int pad[2];
    int q_verdict[S2_MAX_GROUP];  // classification of each survivor (-1 = undecided) and its edge count
    int q_ne[S2_MAX_GROUP];
};

// CDEngine::get_virtual_root (cd_engine.rs:314-327) on the grid: at each level the FIRST child (NE, NW, SW,
// SE) whose bbox surrounds `bb` (rect.rs:46-65); children exist iff the mip map has them.
__device__ __forceinline__ int virtual_root_grid(const GridDev& g, const Rect& bb, int& node) {
    int level = 0, ci = 0, cj = 0;
    node = 0;
    while (level < g.D) {
        const int nl = level + 1, s = g.D - nl, w = 1 << nl;
        const int base = mip_off(nl);
        if (g.mip_node[base + (2 * cj) * w + 2 * ci] == 0xFFFF) break;  // no children
        int found = -1;
#pragma unroll
        for (int qd = 0; qd < 4; qd++) {
            const int dx = (qd == 0 || qd == 3) ? 1 : 0, dy = (qd == 0 || qd == 1) ? 1 : 0;
            const int xi = 2 * ci + dx, yj = 2 * cj + dy;
            const Rect cb{g.xs[xi << s], g.ys[yj << s], g.xs[(xi + 1) << s], g.ys[(yj + 1) << s]};
            if (found < 0 && rect_relation(cb, bb) == REL_SURROUNDING) found = qd;
        }
        if (found < 0) break;
        ci = 2 * ci + ((found == 0 || found == 3) ? 1 : 0);
        cj = 2 * cj + ((found == 0 || found == 1) ? 1 : 0);
        level = nl;
        node = g.mip_node[base + cj * w + ci];
    }
    return (level << 16) | (cj << 8) | ci;
}

// Is the terminal of finest cell (i, j), at level lt, reached by QTNode::collides from the virtual root
// (level l0)? Every ancestor from the virtual root down must flag the quadrant towards the cell
// (qt_node.rs:111-126 with qt_traits.rs:19-106).
template <bool ROLLED = false>
__device__ __forceinline__ bool path_reaches(const GridDev& g, const Seg& sg, int l0, int lt, int i, int j) {
    const int D = g.D;
    for (int l = l0; l < lt; l++) {
        const int s = D - l;
        const int ci = i >> s, cj = j >> s;
        // the child towards the cell holds an endpoint: flagged for sure (qt_traits.rs:36-44: `contains` -> Some(true), and
        // half_intersect only ever adds quadrants). The child's box is the grid's (k_grid_build checked every node box against
        // the nested midpoints that Rect::quadrants computes, rect.rs:141-151), so this is the reference's own comparison.
        const int s1 = s - 1, ki = i >> s1, kj = j >> s1;
        const Rect kb{g.xs[ki << s1], g.ys[kj << s1], g.xs[(ki + 1) << s1], g.ys[(kj + 1) << s1]};
        if (rect_point(kb, sg.sx, sg.sy) || rect_point(kb, sg.ex, sg.ey)) continue;
        const Rect nb{g.xs[ci << s], g.ys[cj << s], g.xs[(ci + 1) << s], g.ys[(cj + 1) << s]};
        const unsigned mask = ROLLED ? seg_quadrants_rolled(sg, nb) : seg_quadrants(sg, nb);
        const int qd = quad_of(ki & 1, kj & 1);
        if (!((mask >> qd) & 1u)) return false;
    }
    return true;
}

// QTNode::collides (qt_node.rs:99-147) for an edge entity from the root, on the grid: the candidate cells of the edge's
// window, would-be hits (Entire cells, leaf cells whose edge tests succeed) verified against the reference's descent - the
// scheme of stage 2 (section 4.1 of DESIGN.md) with the root as the starting node. One lane, one edge.
template <bool FILT>
__device__ bool seg_collides_from_root(const GCtx& c, const Seg& sg) {
    const GridDev& g = c.g;
    const int G = g.G;
    // the superset argument scales with the coordinates involved (a pier may stick out of the CDE's box)
    const float amax = fmaxf(fmaxf(fabsf(sg.sx), fabsf(sg.ex)), fmaxf(fabsf(sg.sy), fabsf(sg.ey)));
    const float delta = fmaxf(g.delta, amax * (1.0f / 8192.0f));
    const int i0 = cell_lo(g.xs, G, fminf(sg.sx, sg.ex) - delta, g.x0, g.inv_cw);
    const int i1 = cell_hi(g.xs, G, fmaxf(sg.sx, sg.ex) + delta, g.x0, g.inv_cw);
    const int j0 = cell_lo(g.ys, G, fminf(sg.sy, sg.ey) - delta, g.y0, g.inv_ch);
    const int j1 = cell_hi(g.ys, G, fmaxf(sg.sy, sg.ey) + delta, g.y0, g.inv_ch);
    const uint32_t cm = col_mask(i0, i1);
    for (int j = j0; j <= j1; j++) {
        const uint32_t er = g.ent_rows[j] & cm;
        uint32_t bits = er | (g.leaf_rows[j] & cm);
        while (bits) {
            const int i = __ffs(bits) - 1;
            bits &= bits - 1;
            if (!(((er >> i) & 1u) || leaf_cell_collides<Seg, FILT>(c.L, g.cell_leaf[j * GRID_G + i], sg, c.filt))) continue;
            const Rect cb{g.xs[i], g.ys[j], g.xs[i + 1], g.ys[j + 1]};
            if (rect_point(cb, sg.sx, sg.sy) || rect_point(cb, sg.ex, sg.ey) || path_reaches<true>(g, sg, 0, g.cell_level[j * GRID_G + i], i, j))
                return true;
        }
    }
    return false;
}

// one group of `ng` survivors (ents[0 .. ng): survivor-list entries in global memory, or in the warp's own shared
// memory when LOCAL) resolved by a warp; GS = slots in use (>= ng)
#ifndef JG_S2_BURST
#define JG_S2_BURST 16
#endif
#ifndef JG_S2_BURST_MIN_RUN
#define JG_S2_BURST_MIN_RUN 6  // mean hazard edges per leaf entry above which a round runs JG_S2_BURST tests
#endif
#ifndef JG_S2_REFILL
#define JG_S2_REFILL 8
#endif
template <bool FILT, bool LOCAL, bool SPLIT, bool STAGE = false>
__device__ __forceinline__ void g2_group(const GCtx& c, G2Slots& W, const int* ents, int ng, int GS, int lane) {
    const int D = c.g.D, G = c.g.G;
    if (lane < ng) {
        const int ent = LOCAL ? ents[lane] : __ldcg(&ents[lane]);
        const int q = ent & 0x3FFFFFF;
        W.q[lane] = q;
        W.seg[lane] = (int)((unsigned)ent >> 26);
        W.xf[lane] = load_xf(c.xforms, c.tab.seg[(unsigned)ent >> 26], q);
    }
    __syncwarp();
    // SPolygon::transform_from + generate_bounding_box (simple_polygon.rs:226-253, 122-133): min/max folds are
    // exact in any order. Four lanes per survivor (quad = survivor) for ordinary items, the whole warp per
    // survivor for long vertex lists.
    const int qs = lane >> 2, sub = lane & 3;
    int nv_mine = 0;
    if (qs < ng) nv_mine = c.tab.seg[W.seg[qs]].n_verts;
    if (__reduce_max_sync(FULL, nv_mine) <= 128) {
        float x0 = 3.402823466e+38f, y0 = 3.402823466e+38f, x1 = -3.402823466e+38f, y1 = -3.402823466e+38f;
        if (qs < ng) {
            const SegDesc& sd = c.tab.seg[W.seg[qs]];
            const float4 t4 = W.xf[qs];
            const Xf xf{t4.x, t4.y, t4.z, t4.w};
            for (int j = sub; j < nv_mine; j += 4) {
                const float2 v = __ldg(&sd.verts[j]);
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
        if (qs < ng && sub == 0) W.bb[qs] = make_float4(x0, y0, x1, y1);
    } else {
        for (int s = 0; s < ng; s++) {
            const SegDesc& sd = c.tab.seg[W.seg[s]];
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
    }
    __syncwarp();
    // cd_engine.rs:161-163 + get_virtual_root (cd_engine.rs:314-327): quad = survivor, lane = child quadrant
    // (NE, NW, SW, SE), so "the first child that surrounds the bbox" is the lowest set bit of the quad's ballot
    {
        Rect bb{0.0f, 0.0f, 0.0f, 0.0f};
        bool walking = false;
        if (qs < ng) {
            const float4 b4 = W.bb[qs];
            bb = Rect{b4.x, b4.y, b4.z, b4.w};
            walking = rect_relation(c.L.cde_bbox, bb) == REL_SURROUNDING;
        }
        const bool inside = walking;
        int level = 0, ci = 0, cj = 0, vnode = 0;
        for (;;) {
            bool sur = false;
            const int nl = level + 1, sh = D - nl, w = 1 << nl, base = mip_off(nl);
            if (walking && level < D && c.g.mip_node[base + (2 * cj) * w + 2 * ci] != 0xFFFF) {
                const int xi = 2 * ci + ((sub == 0 || sub == 3) ? 1 : 0), yj = 2 * cj + ((sub == 0 || sub == 1) ? 1 : 0);
                const Rect cb{c.g.xs[xi << sh], c.g.ys[yj << sh], c.g.xs[(xi + 1) << sh], c.g.ys[(yj + 1) << sh]};
                sur = rect_relation(cb, bb) == REL_SURROUNDING;
            }
            const unsigned all = __ballot_sync(FULL, sur);
            if (!all) break;
            const unsigned mine = (all >> (qs * 4)) & 0xFu;
            if (mine) {
                const int found = __ffs(mine) - 1;
                ci = 2 * ci + ((found == 0 || found == 3) ? 1 : 0);
                cj = 2 * cj + ((found == 0 || found == 1) ? 1 : 0);
                level = nl;
                vnode = c.g.mip_node[base + cj * w + ci];
            } else {
                walking = false;
            }
        }
        if (qs < ng && sub == 0) {
            int v = -1, ne = 0;
            if (!inside) {
                v = JG_VERDICT_POLY;
            } else {
                const int vk = __ldg(&c.g_kinds[vnode + 3]);
                W.vnode[qs] = vnode;
                W.vcell[qs] = (level << 16) | (cj << 8) | ci;
                if (vk == KIND_ENTIRE) v = JG_VERDICT_POLY;  // every edge sees it (qt_node.rs:108)
                else if (vk != KIND_NONE) ne = c.tab.seg[W.seg[qs]].n_verts;
            }
            W.q_verdict[qs] = v;
            W.q_ne[qs] = ne;
        }
    }
    __syncwarp();
    int verdict = -1, nE = 0;
    if (lane < ng) verdict = W.q_verdict[lane], nE = W.q_ne[lane];
    int incl = nE;
#pragma unroll
    for (int o = 1; o < S2_MAX_GROUP; o <<= 1) {
        const int tt = __shfl_up_sync(FULL, incl, o);
        if (lane >= o) incl += tt;
    }
    const int E = __shfl_sync(FULL, incl, S2_MAX_GROUP - 1);
    if (lane < S2_MAX_GROUP) W.ebeg[lane + 1] = incl;
    if (lane == 0) {
        W.ebeg[0] = 0;
        W.hit = 0u;
    }
    const unsigned need = __ballot_sync(FULL, nE > 0);
    __syncwarp();

    // Convergent rounds WITH REFILL over the group's flat edge list. Every lane is a small state machine over (item edge ->
    // candidate cell -> relevant Partial entry -> hazard edge). A round (1) hands the next edges of the list to the idle lanes
    // - once at least JG_S2_REFILL lanes are idle, so that the edge set-up (transform, four cell look-ups) runs densely too -,
    // (2) moves each lane on to its next PENDING EDGE TEST (cheap scanning: row words, entry boxes), (3) runs seg_seg for all
    // lanes that have one, side by side, and (4) lets the lanes whose test (or Entire cell) produced a would-be hit verify
    // that the reference's walk reaches the cell. Per edge the same tests in the same order as the nested loops
    // (qt_partial_hazard.rs:67-72: first colliding edge of the first colliding entry); a survivor's remaining edges are
    // skipped as soon as one of its edges collides (the verdict is an OR over the edges, cd_engine.rs:164-173).
    {
        int fnext = 0;  // next edge of the list to hand out (warp-uniform)
        const int burst = c.L.n_edges > JG_S2_BURST_MIN_RUN * c.L.n_entries ? JG_S2_BURST : 1;  // by the layout's mean edge run
        int slot = 0;
        bool scanning = false;
        Seg sg{0.0f, 0.0f, 0.0f, 0.0f};
        int l0 = 0, j = 0, j1 = -1;
        uint32_t cm = 0u, bits = 0u, er = 0u;
        int hi = -1, ci = 0;   // would-be hit: column (the row is j); column of the cell being tested
        int li = 0, ln = 0;    // entries of the current leaf cell still to look at
        int hk = 0, hn = 0;    // hazard edges of the current entry still to test
        for (;;) {
            const unsigned hm0 = *(volatile unsigned*)&W.hit;
            if ((hm0 & need) == need) break;  // every survivor with edges already has a colliding one
            if ((hm0 >> slot) & 1u) scanning = false, hk = hn, li = ln;
            const unsigned idle = __ballot_sync(FULL, !scanning);
            if (fnext < E && (__popc(idle) >= JG_S2_REFILL || idle == FULL)) {
                const int f = fnext + __popc(idle & lanemask_lt());
                fnext += __popc(idle);
                if (!scanning && f < E) {
                    // the last slot whose first edge is <= f: ebeg[0 .. 8] is non-decreasing (slots beyond ng hold E)
                    slot = 0;
                    if (W.ebeg[4] <= f) slot = 4;
                    if (W.ebeg[slot + 2] <= f) slot += 2;
                    if (W.ebeg[slot + 1] <= f) slot += 1;
                    if (!((hm0 >> slot) & 1u)) {
                        const float4 t4 = W.xf[slot];
                        sg = item_edge(c.tab.seg[W.seg[slot]], Xf{t4.x, t4.y, t4.z, t4.w}, f - W.ebeg[slot]);
                        const int vc = W.vcell[slot];
                        l0 = vc >> 16;
                        const int s0 = D - l0;
                        // candidate cells: the edge's bbox inflated by delta, clipped to the virtual root's cells
                        int i0 = cell_lo(c.g.xs, G, fminf(sg.sx, sg.ex) - c.g.delta, c.g.x0, c.g.inv_cw);
                        int i1 = cell_hi(c.g.xs, G, fmaxf(sg.sx, sg.ex) + c.g.delta, c.g.x0, c.g.inv_cw);
                        j = cell_lo(c.g.ys, G, fminf(sg.sy, sg.ey) - c.g.delta, c.g.y0, c.g.inv_ch);
                        j1 = cell_hi(c.g.ys, G, fmaxf(sg.sy, sg.ey) + c.g.delta, c.g.y0, c.g.inv_ch);
                        const int vi0 = (vc & 255) << s0, vj0 = ((vc >> 8) & 255) << s0;
                        i0 = max(i0, vi0), i1 = min(i1, vi0 + (1 << s0) - 1);
                        j = max(j, vj0), j1 = min(j1, vj0 + (1 << s0) - 1);
                        cm = col_mask(i0, i1);
                        scanning = j <= j1 && cm != 0u;
                        li = ln = hk = hn = 0;
                        if (scanning) {
                            er = c.g.ent_rows[j] & cm;
                            bits = er | (c.g.leaf_rows[j] & cm);
                        }
                    }
                }
            } else if (idle == FULL) {
                break;  // the list is handed out and every lane is through its edge
            }
            if (hk >= hn) {
                while (scanning) {
                    if (li < ln) {
                        const int4 em = STAGE ? c.L.entry_meta[li] : __ldg(&c.L.entry_meta[li]);
                        const Rect ff = STAGE ? ld_rect_any(&c.L.entry_ff[li]) : ld_rect(&c.L.entry_ff[li]);
                        li++;
                        if (FILT && filt_bit(c.filt, em.x)) continue;
                        if (!rect_seg(ff, sg)) continue;
                        hk = em.y, hn = em.y + em.z;
                        if (hk < hn) break;
                        continue;
                    }
                    if (!bits) {
                        if (j >= j1) { scanning = false; break; }
                        j++;
                        er = c.g.ent_rows[j] & cm;
                        bits = er | (c.g.leaf_rows[j] & cm);
                        continue;
                    }
                    ci = __ffs(bits) - 1;
                    bits &= bits - 1;
                    if ((er >> ci) & 1u) { hi = ci; break; }
                    const uint32_t rec = c.g.cell_leaf[j * GRID_G + ci];
                    li = (int)(rec >> 8), ln = li + (int)(rec & 255u);
                }
            }
            if (hi < 0 && scanning && hk < hn) {
                // (up to `burst` tests per round: entries with long edge runs - hazards of a thousand vertices in depth-5
                // leaves, 121 edges per entry - would otherwise pay the round's bookkeeping per edge; gardeyn4_c, 4.2 M queries:
                // 0.86 G queries/s with one test per round, 1.05 G with 8, 1.12 G with 16; swim, 3.4 edges per entry, would lose
                // 1.7 % to a burst of 4 and keeps one)
#pragma unroll 1
                for (int u = 0; u < burst && hk < hn; u++) {
                    const float4 v = STAGE ? c.L.edges[hk] : __ldg(&c.L.edges[hk]);
                    hk++;
                    if (seg_seg(sg, Seg{v.x, v.y, v.z, v.w})) hi = ci, hk = hn, li = ln;  // the cell collides: leave it
                }
            }
            if (!__any_sync(FULL, hi >= 0)) continue;
            if (hi >= 0) {
                const Rect cb{c.g.xs[hi], c.g.ys[j], c.g.xs[hi + 1], c.g.ys[j + 1]};
                if (rect_point(cb, sg.sx, sg.sy) || rect_point(cb, sg.ex, sg.ey) ||
                    path_reaches<!SPLIT>(c.g, sg, l0, c.g.cell_level[j * GRID_G + hi], hi, j))
                    atomicOr(&W.hit, 1u << slot);
                hi = -1;
            }
            __syncwarp();
        }
    }
    __syncwarp();
    const unsigned hm = *(volatile unsigned*)&W.hit;
    if (SPLIT) {
        // cd_engine.rs:174-193 continues in k_s2_contain: survivors without an edge hit whose virtual root holds Partial
        // hazards are handed over (their verdict stays 0 unless a containment collision is found there). Keeping the
        // containment code out of this kernel takes a fifth off its instruction-cache footprint.
        bool handover = false;
        int ct_entries = 0;  // first entry of the virtual root (all Partial: n_entire == 0)
        if (lane < ng && verdict < 0) {
            if ((hm >> lane) & 1u) {
                verdict = JG_VERDICT_POLY;
            } else {
                const int4 m = __ldg(&c.L.node_meta[W.vnode[lane]]);
                const int n_partial = (m.z >> 16) & 0xFFFF;
                if ((m.z & 0xFFFF) > 0) verdict = JG_VERDICT_REF_PANIC;  // cd_engine.rs:179
                else {
                    verdict = 0;
                    handover = n_partial > 0;
                    ct_entries = m.y;
                    W.q_ne[lane] = n_partial;  // (the edge list is done with: the slot carries the entry count now)
                }
            }
        }
        const unsigned hv = __ballot_sync(FULL, handover);
        if (hv) {
            int cb = 0;
            if (lane == 0) cb = atomicAdd(&c.ctl->n_ct, __popc(hv));
            cb = __shfl_sync(FULL, cb, 0);
            if (handover) {
                const int k = cb + __popc(hv & lanemask_lt());
                c.ct_meta[k] = make_int4(sv_pack(W.q[lane], W.seg[lane]), ct_entries, W.q_ne[lane], 0);
                c.ct_bb[k] = W.bb[lane];
                c.ct_xf[k] = W.xf[lane];
            }
        }
    } else {
    // cd_engine.rs:174-193: survivors without an edge hit -> `unreachable!` or containment tests
    for (int s = 0; s < ng; s++) {
        if (__shfl_sync(FULL, verdict, s) >= 0) continue;
        int v = 0;
        if ((hm >> s) & 1u) {
            v = JG_VERDICT_POLY;
        } else {
            const int4 m = __ldg(&c.L.node_meta[W.vnode[s]]);
            const int n_entire = m.z & 0xFFFF, n_partial = (m.z >> 16) & 0xFFFF;
            if (n_entire > 0) {
                v = JG_VERDICT_REF_PANIC;  // cd_engine.rs:179, reachable only under a filter
            } else {
                const SegDesc& sd = c.tab.seg[W.seg[s]];
                const float4 t4 = W.xf[s], b4 = W.bb[s];
                bool col = false;
                for (int j0 = 0; j0 < n_partial && !col; j0 += 32) {
                    const int j = j0 + lane;
                    int hs = -1;
                    if (j < n_partial) {
                        hs = __ldg(&c.g_entry_meta[m.y + j]).x;
                        if (FILT && filt_bit(c.filt, hs)) hs = -1;
                    }
                    col = containment_collision_warp(c.L, sd.verts, sd.n_verts, sd.poi_x, sd.poi_y, Xf{t4.x, t4.y, t4.z, t4.w},
                                                   Rect{b4.x, b4.y, b4.z, b4.w}, hs, lane);
                }
                if (__any_sync(FULL, col)) v = JG_VERDICT_POLY;
            }
        }
        if (lane == s) verdict = v;
    }
    }
    if (lane < ng) c.out[W.q[lane]] = (uint8_t)verdict;
    __syncwarp();
}

#define S2_STAGE_THREADS 1024  // STAGE: one CTA per SM, so that the SM holds ONE copy of the staged sections
template <bool FILT, bool SPLIT, bool STAGE = false>
// (the split variant, without the containment code, gains from a fifth CTA per SM at 48 registers - 4M queries 646 -> 637 us;
// the inline variant loses: 1M queries 212 -> 222 us)
__global__ void __launch_bounds__(STAGE ? S2_STAGE_THREADS : S_THREADS, STAGE ? 1 : (SPLIT ? G2_MIN_BLOCKS + 1 : G2_MIN_BLOCKS))
k_s2_grid(const DevLayout L, const __grid_constant__ SegTable gtab, const GridDev* __restrict__ ggrid,
          const uint8_t* __restrict__ g_kinds, const XfSrc xforms, const uint32_t* __restrict__ filt,
          uint8_t* __restrict__ out, DetectCtl* __restrict__ ctl, int* __restrict__ sv_entry,
          int4* __restrict__ ct_meta, float4* __restrict__ ct_bb, float4* __restrict__ ct_xf) {
    extern __shared__ __align__(16) unsigned char smem[];
    SegTable* t = (SegTable*)smem;
    for (int i = threadIdx.x; i < (int)(sizeof(SegTable) / 4); i += blockDim.x) ((int*)t)[i] = ((const int*)&gtab)[i];
    unsigned char* p = smem + ((sizeof(SegTable) + 15) & ~(size_t)15);
    const GridDev& g = *stage_grid(p, ggrid);
    DevLayout Ls = L;
    if (STAGE) {
        // the leaf entries and hazard edges of the whole layout (swim: 100 KB) by TMA bulk copies into this CTA's shared
        // memory: issued by one thread, completion counted in bytes on an mbarrier (the layout belongs to an earlier upload,
        // so this too runs ahead of griddepcontrol.wait)
        uint64_t* bar = (uint64_t*)p;
        p += 16;
        const uint32_t b_em = (uint32_t)L.n_entries * 16u, b_ed = (uint32_t)L.n_edges * 16u;
        int4* s_em = (int4*)p;
        float4* s_ff = (float4*)(p + b_em);
        float4* s_ed = (float4*)(p + 2 * (size_t)b_em);
        p += 2 * (size_t)b_em + b_ed;
        if (threadIdx.x == 0) mbar_init(bar, 1);
        __syncthreads();
        if (threadIdx.x == 0) {
            mbar_expect_tx(bar, 2u * b_em + b_ed);
            bulk_g2s(s_em, L.entry_meta, b_em, bar);
            bulk_g2s(s_ff, L.entry_ff, b_em, bar);
            bulk_g2s(s_ed, L.edges, b_ed, bar);
        }
        mbar_wait(bar, 0u);
        Ls.entry_meta = s_em, Ls.entry_ff = s_ff, Ls.edges = s_ed;
    }
    __syncthreads();
    // dependent launch: everything above reads launch parameters and the grid summary of an earlier upload; the survivor
    // list, its count and the verdict bytes belong to stage 1, complete and visible only after this wait
    asm volatile("griddepcontrol.wait;" ::: "memory");
    // ... and k_s2_contain, when it follows as a dependent launch, may move in as this (fully resident) grid drains
    if (SPLIT) asm volatile("griddepcontrol.launch_dependents;");
    const GCtx c{STAGE ? Ls : L, *t, g, xforms, filt, g_kinds, out, ctl, sv_entry, ct_meta, ct_bb, ct_xf, L.entry_meta};
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    G2Slots& W = *(G2Slots*)(p + (size_t)warp * sizeof(G2Slots));
    const int total = ctl->n_surv;
    // survivors per group: 8 when every warp has plenty to do (best lane use); fewer when the launch is small,
    // where the length of a warp's dependent chain, not the issue rate, sets the kernel's duration
    const int per_warp = total / (int)(gridDim.x * (blockDim.x >> 5));
    const int GS = per_warp >= 16 ? 8 : per_warp >= 6 ? 4 : per_warp >= 2 ? 2 : 1;
    for (;;) {
        int gbase = 0;
        if (lane == 0) gbase = atomicAdd(&ctl->s2_next, GS);
        gbase = __shfl_sync(FULL, gbase, 0);
        if (gbase >= total) break;
        g2_group<FILT, false, SPLIT, STAGE>(c, W, sv_entry + gbase, (total - gbase) < GS ? (total - gbase) : GS, GS, lane);
    }
}

// ---------------------------------------------------------------------------------------------------
// containment pass: CDEngine::detect_containment_collision (cd_engine.rs:231-255) against every relevant Partial
// hazard of the virtual root (cd_engine.rs:182-192), for the survivors k_s2_grid handed over. A team of four lanes owns
// one survivor (8 survivors per warp, all teams in lockstep): the team's lanes take the hazards four at a time for the
// bbox relation (rect.rs:70-89), and walk a polygon with stride 4 for the rare point-in-polygon tests
// (simple_polygon.rs:255-292; the crossing count of a task is the popcount of its team's ballot bits). The verdict of a
// survivor is an OR over its hazards, so the order is free.
// ---------------------------------------------------------------------------------------------------
#define CT_GROUP 8
#ifndef CT_MIN_BLOCKS
#define CT_MIN_BLOCKS 4
#endif
template <bool FILT>
__global__ void __launch_bounds__(S_THREADS, CT_MIN_BLOCKS)
k_s2_contain(const DevLayout L, const __grid_constant__ SegTable gtab, const uint32_t* __restrict__ filt,
             uint8_t* __restrict__ out, DetectCtl* __restrict__ ctl, const int4* __restrict__ ct_meta,
             const float4* __restrict__ ct_bb, const float4* __restrict__ ct_xf) {
    extern __shared__ __align__(16) unsigned char smem[];
    SegTable* t = (SegTable*)smem;
    for (int i = threadIdx.x; i < (int)(sizeof(SegTable) / 4); i += blockDim.x) ((int*)t)[i] = ((const int*)&gtab)[i];
    __syncthreads();
    // dependent launch behind k_s2_grid<.., true>: the hand-over list and its count are complete after this wait
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const int total = ctl->n_ct;
    if (total == 0) return;
    const SegTable& tab = *t;
    const int lane = threadIdx.x & 31, team = lane >> 2, sub = lane & 3;
    const unsigned team_bits = 0xFu << (team * 4);
    for (;;) {
        int gbase = 0;
        if (lane == 0) gbase = atomicAdd(&ctl->ct_next, CT_GROUP);
        gbase = __shfl_sync(FULL, gbase, 0);
        if (gbase >= total) break;
        const bool valid = gbase + team < total;
        int q = 0, np = 0, eb = 0;
        const SegDesc* sd = &tab.seg[0];
        Xf xf{1.0f, 0.0f, 0.0f, 0.0f};
        Rect sb{0.0f, 0.0f, 0.0f, 0.0f};
        if (valid) {
            const int4 cm = __ldcg(&ct_meta[gbase + team]);
            const float4 t4 = __ldcg(&ct_xf[gbase + team]), b4 = __ldcg(&ct_bb[gbase + team]);
            q = cm.x & 0x3FFFFFF;
            sd = &tab.seg[(unsigned)cm.x >> 26];
            xf = Xf{t4.x, t4.y, t4.z, t4.w};
            sb = Rect{b4.x, b4.y, b4.z, b4.w};
            eb = cm.y;  // n_entire == 0 here (k_s2_grid reports the other case itself): Partial entries start the node's list
            np = cm.z;
        }
        bool collided = false;
        const int max_np = __reduce_max_sync(FULL, np);
        for (int j0 = 0; j0 < max_np; j0 += 4) {
            const int j = j0 + sub;
            int hs = -1, rel = REL_DISJOINT, scope_interior = 1;
            if (j < np && !collided) {
                hs = __ldg(&L.entry_meta[eb + j]).x;
                if (FILT && filt_bit(filt, hs)) hs = -1;
            }
            if (hs >= 0) {
                scope_interior = __ldg(&L.haz_meta[hs]).x == 1 ? 1 : 0;
                rel = rect_almost_relation(ld_rect(&L.haz_bbox[hs]), sb);
            }
            const bool pip = hs >= 0 && (rel == REL_SURROUNDING || rel == REL_ENCLOSED);
            // no point-in-polygon needed: contained = false -> collision iff the hazard's scope is Exterior (hazard.rs:55-62)
            const unsigned quick = __ballot_sync(FULL, hs >= 0 && !pip && !scope_interior);
            if (quick & team_bits) collided = true;
            const unsigned pips = __ballot_sync(FULL, pip);
            for (int k = 0; k < 4; k++) {
                if (!(pips & (0x11111111u << k))) continue;  // no team has a task in this position
                const int owner = team * 4 + k;
                const int h = __shfl_sync(FULL, hs, owner), r = __shfl_sync(FULL, rel, owner), si = __shfl_sync(FULL, scope_interior, owner);
                const bool act = ((pips >> owner) & 1u) && !collided;
                const float2* verts = nullptr;
                int n = 0;
                bool use_xf = false;
                Rect pb{0.0f, 0.0f, 0.0f, 0.0f};
                float px = 0.0f, py = 0.0f;
                if (act) {
                    if (r == REL_SURROUNDING) {  // the hazard's shape against the item's (transformed) poi
                        const int4 hmeta = __ldg(&L.haz_meta[h]);
                        verts = L.haz_verts + hmeta.y, n = hmeta.z;
                        pb = ld_rect(&L.haz_bbox[h]);
                        xf_point(xf, sd->poi_x, sd->poi_y, px, py);
                    } else {  // the item's transformed shape against the hazard's poi
                        const float4 hp = __ldg(&L.haz_poi[h]);
                        verts = sd->verts, n = sd->n_verts, use_xf = true;
                        pb = sb;
                        px = hp.x, py = hp.y;
                    }
                    if (!rect_point(pb, px, py)) n = 0;  // outside the bbox: not contained (simple_polygon.rs:258-262)
                }
                const Seg ray{px, py, pb.x_max + (pb.x_max - pb.x_min), py};
                const int trips = __reduce_max_sync(FULL, act ? (n + 3) >> 2 : 0);
                int n_int = 0;
                for (int b = 0; b < trips; b++) {
                    const int i = b * 4 + sub;
                    bool cross = false;
                    if (act && i < n) {
                        const float2 a = __ldg(&verts[i]), bb2 = __ldg(&verts[i == n - 1 ? 0 : i + 1]);
                        float sx = a.x, sy = a.y, ex = bb2.x, ey = bb2.y;
                        if (use_xf) {
                            xf_point(xf, a.x, a.y, sx, sy);
                            xf_point(xf, bb2.x, bb2.y, ex, ey);
                        }
                        if ((approx_eq(sy, py) && fpa_gt(sx, px)) || (approx_eq(ey, py) && fpa_gt(ex, px))) cross = fpa_lt(sy, py) || fpa_lt(ey, py);
                        else cross = seg_seg(ray, Seg{sx, sy, ex, ey});
                    }
                    n_int += __popc(__ballot_sync(FULL, cross) & team_bits);
                }
                const bool contained = (n_int & 1) == 1;
                if (act && (si ? contained : !contained)) collided = true;
            }
        }
        if (valid && sub == 0 && collided) out[q] = (uint8_t)JG_VERDICT_POLY;
    }
}

// ---------------------------------------------------------------------------------------------------
// fused, streaming form: ONE persistent kernel in which every warp alternates between the two stages.
//   * stage-2 groups are taken first whenever 8 published survivors wait, so the polygon stage overlaps the
//     surrogate stage instead of following it (no second launch, one ramp-up, one tail);
//   * queries may still be in flight over PCIe: a warp processes its batch of 32 once ctl->avail covers it. The
//     host batcher copies the batch in chunks and bumps `avail` after each chunk (a 4-byte copy ordered behind
//     the chunk on the same stream), so compute follows the copy front. avail_all != 0: everything is resident.
//   * every spin is bounded (JG_SPIN_LIMIT clock ticks without progress set ctl->error and the warp leaves), so a
//     lost copy cannot hang the device.
// ---------------------------------------------------------------------------------------------------
#define JG_SPIN_LIMIT (4000000000ll)  // ~2 s at 1.9 GHz
template <bool FILT>
__global__ void __launch_bounds__(S_THREADS)
k_detect_fused(const DevLayout L, const __grid_constant__ SegTable gtab, const GridDev* __restrict__ ggrid,
               const uint8_t* __restrict__ g_kinds, const XfSrc xforms, const uint32_t* __restrict__ filt,
               uint8_t* __restrict__ out, DetectCtl* __restrict__ ctl, int* __restrict__ sv_entry, int avail_all) {
    extern __shared__ __align__(16) unsigned char smem[];
    SegTable* t = (SegTable*)smem;
    for (int i = threadIdx.x; i < (int)(sizeof(SegTable) / 4); i += blockDim.x) ((int*)t)[i] = ((const int*)&gtab)[i];
    unsigned char* p = smem + ((sizeof(SegTable) + 15) & ~(size_t)15);
    const GridDev& g = *stage_grid(p, ggrid);
    __syncthreads();
    const GCtx c{L, *t, g, xforms, filt, g_kinds, out, ctl, sv_entry, nullptr, nullptr, nullptr, L.entry_meta};
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char* wp = p + (size_t)warp * GF_WARP_BYTES;
    G1Warp w;
    w.bind(wp);
    G2Slots& W = *(G2Slots*)(wp + G1_WARP_BYTES);
    const int n = c.tab.n_queries, n_warps = gridDim.x * S_WARPS;
    const int GS = S2_MAX_GROUP;
    // One call site per routine (each is several thousand instructions): the loop is a small state machine.
    //   MAIN      stage-1 batches as their queries arrive; the warp's own survivors 8 at a time
    //   BARRIER   leftovers (< 8 per warp) handed to the global list; wait until every warp has done so
    //   LEFTOVER  the global leftovers in full groups again
    enum { MAIN = 0, BARRIER = 1, LEFTOVER = 2 };
    int phase = MAIN, total = 0, base = 0;
    bool s1_open = true, have_batch = false;
    long long idle_since = -1;
    for (;;) {
        int grp_off = -1, grp_n = 0;
        bool run_b = false, run_a = false, idle = false;
        if (phase == MAIN) {
            if (w.n_sv >= GS) {
                w.n_sv -= GS;
                grp_off = w.n_sv, grp_n = GS;
            } else if (w.n_pend >= 32) {
                run_b = true;
            } else {
                if (!have_batch && s1_open) {
                    if (lane == 0) base = atomicAdd(&ctl->s1_next, 32);
                    base = __shfl_sync(FULL, base, 0);
                    if (base >= n) s1_open = false;
                    else have_batch = true;
                }
                if (have_batch) {
                    const int need = (base + 32) < n ? (base + 32) : n;
                    int av = n;
                    if (!avail_all) {
                        if (lane == 0) av = *(volatile int*)&ctl->avail;
                        av = __shfl_sync(FULL, av, 0);
                    }
                    if (av >= need) run_a = true;
                    else if (w.n_pend > 0) run_b = true;  // the batch has not arrived yet: use the time
                    else idle = true;
                } else if (w.n_pend > 0) {
                    run_b = true;  // drain
                } else {
                    if (w.n_sv > 0) {  // hand in the leftovers
                        int sb = 0;
                        if (lane == 0) sb = atomicAdd(&ctl->n_res, w.n_sv);
                        sb = __shfl_sync(FULL, sb, 0);
                        if (lane < w.n_sv) sv_entry[sb + lane] = w.s_surv[lane];
                        __threadfence();
                        w.n_sv = 0;
                    }
                    __syncwarp();
                    if (lane == 0) atomicAdd(&ctl->done_warps, 1);
                    phase = BARRIER;
                    idle_since = -1;
                    continue;
                }
            }
        } else if (phase == BARRIER) {
            int dn = 0;
            if (lane == 0) dn = *(volatile int*)&ctl->done_warps;
            dn = __shfl_sync(FULL, dn, 0);
            if (dn == n_warps) {
                __threadfence();
                total = *(volatile int*)&ctl->n_res;
                phase = LEFTOVER;
                idle_since = -1;
                continue;
            }
            idle = true;
        } else {
            int gbase = 0;
            if (lane == 0) gbase = atomicAdd(&ctl->s2_next, GS);
            gbase = __shfl_sync(FULL, gbase, 0);
            if (gbase >= total) break;
            grp_n = (total - gbase) < GS ? (total - gbase) : GS;
            if (lane < grp_n) w.s_surv[lane] = __ldcg(&sv_entry[gbase + lane]);
            __syncwarp();
            grp_off = 0;
        }
        if (grp_off >= 0) {
            g2_group<FILT, true, false>(c, W, w.s_surv + grp_off, grp_n, GS, lane);
            idle_since = -1;
        } else if (run_b) {
            g1_phase_b<FILT, true>(c, w, lane);
            idle_since = -1;
        } else if (run_a) {
            g1_phase_a<FILT>(c, w, base, lane);
            have_batch = false;
            idle_since = -1;
        } else if (idle) {  // wait a little; give up after JG_SPIN_LIMIT ticks without progress
            const long long now = clock64();
            if (idle_since < 0) idle_since = now;
            if (now - idle_since > JG_SPIN_LIMIT) {
                if (lane == 0) atomicExch(&ctl->error, 1);
                return;
            }
            __nanosleep(100);
        }
    }
}

}  // namespace jg
