// detect pipeline (sm_100a): detect_surrogate_collision + detect_poly_collision for a batch made of up
// to JG_MAX_SEG segments (one segment = one item's queries), four launches per batch:
//
//   k_s1_poles   persistent warps, one lane = one query, lanes REFILL from a shared work counter as soon
//                as their query is decided, so a warp never idles behind its slowest walk. Surrogate
//                stage (cd_engine.rs:204-223): poles then piers from the root. Undecided queries are
//                appended to their segment's survivor list.
//   k_s1b_bbox   one lane per survivor, uniform work: transformed bbox (simple_polygon.rs:226-253),
//                `bbox.relation_to(shape.bbox) != Surrounding -> collision` (cd_engine.rs:161-163),
//                virtual root (cd_engine.rs:314-327).
//   k_s2_edges   persistent warps over the flat list of (survivor, edge) pairs, same refill scheme;
//                neighbouring lanes hold neighbouring edges of one polygon, i.e. the same region of the
//                tree. QTNode::collides (qt_node.rs:99-147) from the virtual root; a hit stores the
//                verdict, edges of an already-decided query are skipped.
//   k_s3_contain one lane per still-undecided survivor: the `unreachable!` of cd_engine.rs:179 (bit 7) or
//                detect_containment_collision over the virtual root's relevant Partial entries.
//
// A verdict is an OR over independent per-edge walks, so the order in which lanes pick up work is never
// observable; every predicate is the reference's arithmetic (predicates.cuh).
#include "kernels.cuh"
#include "launch.hpp"

namespace jg {

#define S_THREADS 256
#define S1_CHUNK 128   // queries fetched per atomic by a warp
#define S2_CHUNK 256   // (survivor, edge) pairs fetched per atomic by a warp
#define REFILL_MIN 8   // refill once this many lanes of a warp are idle (or none is active)

// One node visit of QTNode::collides. Returns 0 = walk continues, 1 = hit, 2 = walk finished, no hit.
template <class Ent, bool FILT>
__device__ __forceinline__ int walk_step(const DevLayout& L, Walk& w, const Ent& ent, const uint32_t* filt) {
    const int4 m = __ldg(&L.node_meta[w.node]);
    const int n_entire = m.z & 0xFFFF, n_partial = (m.z >> 16) & 0xFFFF;
    int first_partial = 0, kind;
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
    if (kind == 2) return 1;
    if (kind == 1) {
        if (m.x >= 0) {
            const unsigned mask = ent_quadrants(ent, ld_rect(&L.node_bbox[w.node]));
            if (mask) { w.descend(m.x, mask); return 0; }
        } else {
            const int e0 = m.y + n_entire;
            for (int i = first_partial; i < n_partial; i++) {
                const int4 em = __ldg(&L.entry_meta[e0 + i]);
                if (FILT && filt_bit(filt, em.x)) continue;
                if (!ent_rect(ent, ld_rect(&L.entry_ff[e0 + i]))) continue;
                for (int k = 0; k < em.z; k++)
                    if (ent_seg(ent, ld_seg(&L.edges[em.y + k]))) return 1;
            }
        }
    }
    return w.next(L) ? 0 : 2;
}

// Warp-level work distribution: the warp owns a range [next, end) of work indices fetched CHUNK at a
// time from a global counter; idle lanes take consecutive indices from it.
struct WarpFeed {
    int next, end;
    bool dry;  // the global counter is exhausted
    __device__ __forceinline__ void init() { next = end = 0; dry = false; }
    // gives lane `lane` an index (or -1). `want` = this lane is idle. Warp-uniform control flow.
    template <int CHUNK>
    __device__ __forceinline__ int take(bool want, int lane, int* counter, int total) {
        const unsigned need = __ballot_sync(0xFFFFFFFFu, want);
        if (next >= end && !dry) {
            int b = 0;
            if (lane == 0) b = atomicAdd(counter, CHUNK);
            b = __shfl_sync(0xFFFFFFFFu, b, 0);
            next = b;
            end = b + CHUNK < total ? b + CHUNK : total;
            if (b >= total) { dry = true; next = end = 0; }
        }
        const int avail = end - next;
        const int rank = __popc(need & ((1u << lane) - 1u));
        const int cnt = __popc(need);
        const int idx = (want && rank < avail) ? next + rank : -1;
        next += cnt < avail ? cnt : avail;
        return idx;
    }
};

__device__ __forceinline__ int seg_of_query(const SegTable& tab, int q) {
    int k = 0;
    for (int j = 1; j < tab.n_seg; j++) k += (q >= tab.seg[j].begin) ? 1 : 0;
    return k;
}

// ---------------------------------------------------------------------------------------------
template <bool FILT>
__global__ void __launch_bounds__(S_THREADS)
k_s1_poles(const DevLayout L, const __grid_constant__ SegTable gtab, const float4* __restrict__ xforms,
           const uint32_t* __restrict__ filt, uint8_t* __restrict__ out, DetectCtl* __restrict__ ctl, Survivors sv) {
    __shared__ SegTable tab;
    for (int i = threadIdx.x; i < (int)(sizeof(SegTable) / 4); i += blockDim.x) ((int*)&tab)[i] = ((const int*)&gtab)[i];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int n = tab.n_queries;

    WarpFeed feed;
    feed.init();
    bool active = false;
    int q = 0, k = 0, ent_i = 0, n_ent = 0;
    Xf xf{1, 0, 0, 0};
    Circ c{0, 0, 0};
    Seg s{0, 0, 0, 0};
    bool is_circ = true;
    Walk w;
    w.start(0);

    for (;;) {
        const unsigned idle = __ballot_sync(0xFFFFFFFFu, !active);
        if (idle == 0xFFFFFFFFu && feed.dry && feed.next >= feed.end) break;
        if (__popc(idle) >= REFILL_MIN || idle == 0xFFFFFFFFu) {
            const int nq = feed.take<S1_CHUNK>(!active, lane, &ctl->s1_next, n);
            if (nq >= 0) {
                q = nq;
                k = seg_of_query(tab, q);
                const float4 t4 = __ldg(&xforms[q]);
                xf = Xf{t4.x, t4.y, t4.z, t4.w};
                ent_i = -1;
                n_ent = tab.seg[k].n_poles + tab.seg[k].n_piers;
                active = true;
                w.depth = -1;  // "entity finished": pick the next one below
            }
        }
        if (active) {
            int r = 2;
            if (w.depth >= 0) r = is_circ ? walk_step<Circ, FILT>(L, w, c, filt) : walk_step<Seg, FILT>(L, w, s, filt);
            if (r == 1) {  // cd_engine.rs:212,218: first colliding pole / pier decides
                out[q] = JG_VERDICT_SURROGATE;
                active = false;
            } else if (r == 2) {
                ent_i++;
                if (ent_i < n_ent) {
                    const SegDesc& sd = tab.seg[k];
                    if (ent_i < sd.n_poles) {  // circle.rs:79-85: centre transformed, radius unchanged
                        const float4 pl = __ldg(&sd.poles[ent_i]);
                        xf_point(xf, pl.x, pl.y, c.cx, c.cy);
                        c.r = pl.z;
                        is_circ = true;
                    } else {  // edge.rs:126-134
                        const float4 pr = __ldg(&sd.piers[ent_i - sd.n_poles]);
                        xf_point(xf, pr.x, pr.y, s.sx, s.sy);
                        xf_point(xf, pr.z, pr.w, s.ex, s.ey);
                        is_circ = false;
                    }
                    w.start(0);  // surrogate walks start at the ROOT (cd_engine.rs:212)
                } else {
                    active = false;  // no surrogate collision -> survivor of segment k
                    const int pos = atomicAdd(&ctl->n_surv[k], 1);
                    sv.query[tab.seg[k].begin + pos] = q;
                }
            }
        }
    }
}

// JG_MODE_POLY_ONLY: no surrogate stage, every query is a survivor (lbf/benches/quadtree_bench.rs:73-124)
__global__ void __launch_bounds__(S_THREADS)
k_s1_all_survive(const __grid_constant__ SegTable tab, DetectCtl* __restrict__ ctl, Survivors sv) {
    const int n = tab.n_queries;
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < n; q += gridDim.x * blockDim.x) sv.query[q] = q;
    if (blockIdx.x == 0 && threadIdx.x < tab.n_seg) {
        const int k = threadIdx.x;
        ctl->n_surv[k] = (k + 1 < tab.n_seg ? tab.seg[k + 1].begin : n) - tab.seg[k].begin;
    }
}

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(S_THREADS)
k_s1b_bbox(const DevLayout L, const __grid_constant__ SegTable tab, const float4* __restrict__ xforms,
           uint8_t* __restrict__ out, const DetectCtl* __restrict__ ctl, Survivors sv) {
    const int k = blockIdx.y;
    const SegDesc& sd = tab.seg[k];
    const int ns = ctl->n_surv[k];
    const int V = sd.n_verts;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < ns; i += gridDim.x * blockDim.x) {
        const int slot = sd.begin + i;
        const int q = sv.query[slot];
        const float4 t4 = __ldg(&xforms[q]);
        const Xf xf{t4.x, t4.y, t4.z, t4.w};
        float x0 = 3.402823466e+38f, y0 = 3.402823466e+38f, x1 = -3.402823466e+38f, y1 = -3.402823466e+38f;
        for (int j = 0; j < V; j++) {
            const float2 v = __ldg(&sd.verts[j]);
            float x, y;
            xf_point(xf, v.x, v.y, x, y);
            x0 = fminf(x0, x), y0 = fminf(y0, y), x1 = fmaxf(x1, x), y1 = fmaxf(y1, y);
        }
        const Rect bb{x0, y0, x1, y1};
        int vroot = -1;
        uint8_t verdict = JG_VERDICT_POLY;
        if (rect_relation(L.cde_bbox, bb) == REL_SURROUNDING) {
            vroot = virtual_root(L, bb);
            verdict = 0;
        }
        out[q] = verdict;
        sv.vroot[slot] = vroot;
        sv.bbox[slot] = make_float4(x0, y0, x1, y1);
    }
}

// ---------------------------------------------------------------------------------------------
template <bool FILT>
__global__ void __launch_bounds__(S_THREADS)
k_s2_edges(const DevLayout L, const __grid_constant__ SegTable gtab, const float4* __restrict__ xforms,
           const uint32_t* __restrict__ filt, uint8_t* out, DetectCtl* __restrict__ ctl, Survivors sv) {
    __shared__ SegTable tab;
    __shared__ int cum[JG_MAX_SEG + 1];  // cum[k] = first flat (survivor, edge) index of segment k
    for (int i = threadIdx.x; i < (int)(sizeof(SegTable) / 4); i += blockDim.x) ((int*)&tab)[i] = ((const int*)&gtab)[i];
    __syncthreads();
    if (threadIdx.x == 0) {
        int acc = 0;
        for (int j = 0; j < tab.n_seg; j++) {
            cum[j] = acc;
            acc += ctl->n_surv[j] * tab.seg[j].n_verts;
        }
        for (int j = tab.n_seg; j <= JG_MAX_SEG; j++) cum[j] = acc;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int total = cum[JG_MAX_SEG];
    volatile uint8_t* vout = out;

    WarpFeed feed;
    feed.init();
    bool active = false;
    int q = 0, seg_hint = 0;
    Seg sg{0, 0, 0, 0};
    Walk w;
    w.start(0);
    for (;;) {
        const unsigned idle = __ballot_sync(0xFFFFFFFFu, !active);
        if (idle == 0xFFFFFFFFu && feed.dry && feed.next >= feed.end) break;
        if (__popc(idle) >= REFILL_MIN || idle == 0xFFFFFFFFu) {
            const int idx = feed.take<S2_CHUNK>(!active, lane, &ctl->s2_next, total);
            int k = 0x7FFFFFFF;
            if (idx >= 0) {
                // indices handed to a warp only grow, so the segment found last time is a lower bound
                k = seg_hint;
                while (k + 1 < tab.n_seg && idx >= cum[k + 1]) k++;
                const SegDesc& sd = tab.seg[k];
                const int r = idx - cum[k];
                const int si = r / sd.n_verts, e = r - si * sd.n_verts;
                const int slot = sd.begin + si;
                q = sv.query[slot];
                if (vout[q] == 0) {  // undecided (not outside the CDE bbox, no edge hit so far)
                    const float4 t4 = __ldg(&xforms[q]);
                    const Xf xf{t4.x, t4.y, t4.z, t4.w};
                    const float2 a = __ldg(&sd.verts[e]);
                    const float2 b = __ldg(&sd.verts[e == sd.n_verts - 1 ? 0 : e + 1]);  // simple_polygon.rs:86-93
                    xf_point(xf, a.x, a.y, sg.sx, sg.sy);
                    xf_point(xf, b.x, b.y, sg.ex, sg.ey);
                    w.start(sv.vroot[slot]);
                    active = true;
                }
            }
            const int kmin = __reduce_min_sync(0xFFFFFFFFu, k);
            if (kmin != 0x7FFFFFFF) seg_hint = kmin;
        }
        if (active) {
            const int r = walk_step<Seg, FILT>(L, w, sg, filt);
            if (r == 1) vout[q] = JG_VERDICT_POLY;
            if (r != 0) active = false;
        }
    }
}

// ---------------------------------------------------------------------------------------------
template <bool FILT>
__global__ void __launch_bounds__(S_THREADS)
k_s3_contain(const DevLayout L, const __grid_constant__ SegTable tab, const float4* __restrict__ xforms,
             const uint32_t* __restrict__ filt, uint8_t* __restrict__ out, const DetectCtl* __restrict__ ctl, Survivors sv) {
    const int k = blockIdx.y;
    const SegDesc& sd = tab.seg[k];
    const int ns = ctl->n_surv[k];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < ns; i += gridDim.x * blockDim.x) {
        const int slot = sd.begin + i;
        const int q = sv.query[slot];
        if (out[q] != 0) continue;
        const int vroot = sv.vroot[slot];
        const int4 m = __ldg(&L.node_meta[vroot]);
        const int n_entire = m.z & 0xFFFF, n_partial = (m.z >> 16) & 0xFFFF;
        if (n_entire > 0) {  // cd_engine.rs:179 `unreachable!` (only reachable under a filter)
            out[q] = JG_VERDICT_REF_PANIC;
            continue;
        }
        const float4 t4 = __ldg(&xforms[q]);
        const Xf xf{t4.x, t4.y, t4.z, t4.w};
        const float4 b4 = sv.bbox[slot];
        const Rect sb{b4.x, b4.y, b4.z, b4.w};
        bool c = false;
        for (int j = 0; j < n_partial && !c; j++) {
            const int hs = __ldg(&L.entry_meta[m.y + j]).x;
            if (FILT && filt_bit(filt, hs)) continue;
            c = containment_collision(L, sd.verts, sd.n_verts, sd.poi_x, sd.poi_y, xf, sb, hs);
        }
        if (c) out[q] = JG_VERDICT_POLY;
    }
}

// ---------------------------------------------------------------------------------------------
static int g_occ[8];
template <class K>
static int persistent_grid(K kernel, int slot, int n_sms) {
    if (g_occ[slot] == 0) {
        int b = 1;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, kernel, S_THREADS, 0);
        g_occ[slot] = b < 1 ? 1 : b;
    }
    return n_sms * g_occ[slot];
}

cudaError_t launch_detect(const DevLayout& L, const SegTable& tab, const float4* d_xf, const uint32_t* d_filt, int mode,
                          uint8_t* d_out, DetectCtl* ctl, const Survivors& sv, int n_sms, cudaStream_t st, int* n_launches) {
    cudaMemsetAsync(ctl, 0, sizeof(DetectCtl), st);
    const int n = tab.n_queries;
    const int want1 = (n + S_THREADS - 1) / S_THREADS;
    if (mode & JG_MODE_POLY_ONLY) {
        const int g = n_sms * 4;
        k_s1_all_survive<<<want1 < g ? want1 : g, S_THREADS, 0, st>>>(tab, ctl, sv);
    } else if (d_filt) {
        const int g = persistent_grid(k_s1_poles<true>, 0, n_sms);
        k_s1_poles<true><<<want1 < g ? want1 : g, S_THREADS, 0, st>>>(L, tab, d_xf, d_filt, d_out, ctl, sv);
    } else {
        const int g = persistent_grid(k_s1_poles<false>, 1, n_sms);
        k_s1_poles<false><<<want1 < g ? want1 : g, S_THREADS, 0, st>>>(L, tab, d_xf, d_filt, d_out, ctl, sv);
    }
    int max_seg = 0;
    for (int k = 0; k < tab.n_seg; k++) {
        const int len = (k + 1 < tab.n_seg ? tab.seg[k + 1].begin : n) - tab.seg[k].begin;
        max_seg = len > max_seg ? len : max_seg;
    }
    int gx = (max_seg + S_THREADS - 1) / S_THREADS;
    const int gx_cap = (n_sms * 8 + tab.n_seg - 1) / tab.n_seg;
    gx = gx < 1 ? 1 : (gx > gx_cap ? gx_cap : gx);
    k_s1b_bbox<<<dim3(gx, tab.n_seg), S_THREADS, 0, st>>>(L, tab, d_xf, d_out, ctl, sv);
    if (d_filt) {
        const int g = persistent_grid(k_s2_edges<true>, 2, n_sms);
        k_s2_edges<true><<<g, S_THREADS, 0, st>>>(L, tab, d_xf, d_filt, d_out, ctl, sv);
        k_s3_contain<true><<<dim3(gx, tab.n_seg), S_THREADS, 0, st>>>(L, tab, d_xf, d_filt, d_out, ctl, sv);
    } else {
        const int g = persistent_grid(k_s2_edges<false>, 3, n_sms);
        k_s2_edges<false><<<g, S_THREADS, 0, st>>>(L, tab, d_xf, d_filt, d_out, ctl, sv);
        k_s3_contain<false><<<dim3(gx, tab.n_seg), S_THREADS, 0, st>>>(L, tab, d_xf, d_filt, d_out, ctl, sv);
    }
    if (n_launches) *n_launches += 4;
    return cudaGetLastError();
}

}  // namespace jg
