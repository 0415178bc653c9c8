// detect pipeline (sm_100a): detect_surrogate_collision + detect_poly_collision for a batch made of up
// to JG_MAX_SEG segments (one segment = one item's queries). One launch sequence per batch:
//
//   k_bin_count / k_bin_scan / k_bin_scatter
//                counting sort of the queries by (segment, rotation octant, Morton cell of the translation
//                on a 32x32 grid over the CDE bbox). Purely a scheduling device: queries that sit next to
//                each other in this order are the same item in almost the same pose, so the lanes of a
//                warp walk the same quadtree nodes and take the same branches. Results never depend on it.
//   k_s1_poles   persistent warps, one lane = one query (in sorted order); lanes REFILL from a shared work
//                counter when their query is decided. Surrogate stage (cd_engine.rs:204-223): poles then
//                piers, each walk from the ROOT. Undecided queries are appended to the survivor list in
//                warp-sized bursts, which keeps the list spatially sorted.
//   k_s2_edges   persistent warps, one lane = one survivor: transformed bbox (simple_polygon.rs:226-253),
//                `bbox.relation_to(shape.bbox) != Surrounding -> collision` (cd_engine.rs:161-163), virtual
//                root (cd_engine.rs:314-327), then the edges IN ORDER with QTNode::collides
//                (qt_node.rs:99-147) and the reference's early exit at the first colliding edge.
//   k_s3_contain one lane per survivor without an edge hit: the `unreachable!` of cd_engine.rs:179 (bit 7) or
//                detect_containment_collision over the virtual root's relevant Partial entries.
//
// Every predicate is the reference's arithmetic (predicates.cuh); the order in which lanes pick up queries
// is never observable because each query's verdict is computed by one lane exactly as the CPU would.
#include "kernels.cuh"
#include "launch.hpp"

namespace jg {

#define S_THREADS 256
#define S_WARPS (S_THREADS / 32)
#define S1_CHUNK 128  // sorted queries fetched per atomic by a warp
#define S2_CHUNK 64   // survivors fetched per atomic by a warp
#define REFILL_MIN 8  // refill once this many lanes of a warp are idle (or none is active)
#define BIN_CELLS 1024
#define BIN_ROT 8

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
    __device__ __forceinline__ bool finished() const { return dry && next >= end; }
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

// ---- spatial binning ------------------------------------------------------------------------------
__device__ __forceinline__ unsigned morton5(unsigned x, unsigned y) {
    unsigned r = 0;
#pragma unroll
    for (int b = 0; b < 5; b++) r |= (((x >> b) & 1u) << (2 * b)) | (((y >> b) & 1u) << (2 * b + 1));
    return r;
}

__global__ void __launch_bounds__(S_THREADS)
k_bin_count(const __grid_constant__ SegTable tab, const DevLayout L, const float4* __restrict__ xforms, int* __restrict__ keys,
            int* __restrict__ hist) {
    const int n = tab.n_queries;
    const float sx = 32.0f / (L.cde_bbox.x_max - L.cde_bbox.x_min), sy = 32.0f / (L.cde_bbox.y_max - L.cde_bbox.y_min);
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < n; q += gridDim.x * blockDim.x) {
        const float4 t = __ldg(&xforms[q]);
        int k = 0;
        for (int j = 1; j < tab.n_seg; j++) k += (q >= tab.seg[j].begin) ? 1 : 0;
        const int cx = min(31, max(0, (int)((t.z - L.cde_bbox.x_min) * sx)));
        const int cy = min(31, max(0, (int)((t.w - L.cde_bbox.y_min) * sy)));
        const int rot = (t.x < 0.0f ? 1 : 0) | (t.y < 0.0f ? 2 : 0) | (fabsf(t.x) < fabsf(t.y) ? 4 : 0);
        const int key = ((k * BIN_ROT + rot) * BIN_CELLS) + (int)morton5((unsigned)cx, (unsigned)cy);
        keys[q] = key;
        atomicAdd(&hist[key], 1);
    }
}

// exclusive scan of hist[0 .. n_bins) in place, one CTA
__global__ void __launch_bounds__(1024) k_bin_scan(int* __restrict__ hist, int n_bins) {
    __shared__ int part[1024];
    const int per = (n_bins + 1023) / 1024;
    const int b0 = min(n_bins, (int)threadIdx.x * per), b1 = min(n_bins, b0 + per);
    int s = 0;
    for (int i = b0; i < b1; i++) s += hist[i];
    part[threadIdx.x] = s;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {  // Hillis-Steele inclusive scan
        const int v = (int)threadIdx.x >= off ? part[threadIdx.x - off] : 0;
        __syncthreads();
        part[threadIdx.x] += v;
        __syncthreads();
    }
    int run = part[threadIdx.x] - s;
    for (int i = b0; i < b1; i++) {
        const int c = hist[i];
        hist[i] = run;
        run += c;
    }
}

__global__ void __launch_bounds__(S_THREADS)
k_bin_scatter(int n, const int* __restrict__ keys, int* __restrict__ cursor, int* __restrict__ order) {
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < n; q += gridDim.x * blockDim.x)
        order[atomicAdd(&cursor[keys[q]], 1)] = q;
}

// ---------------------------------------------------------------------------------------------
// survivor entries: query index in the low 26 bits, segment in the high 6
__device__ __forceinline__ int sv_pack(int q, int k) { return q | (k << 26); }

template <bool FILT>
__global__ void __launch_bounds__(S_THREADS)
k_s1_poles(const DevLayout L, const __grid_constant__ SegTable gtab, const float4* __restrict__ xforms,
           const int* __restrict__ order, const uint32_t* __restrict__ filt, uint8_t* __restrict__ out,
           DetectCtl* __restrict__ ctl, Survivors sv) {
    __shared__ SegTable tab;
    __shared__ int burst[S_WARPS][64];
    for (int i = threadIdx.x; i < (int)(sizeof(SegTable) / 4); i += blockDim.x) ((int*)&tab)[i] = ((const int*)&gtab)[i];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int n = tab.n_queries;
    int* wb = burst[threadIdx.x >> 5];
    int n_burst = 0;  // warp-uniform

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
        if (idle == 0xFFFFFFFFu && feed.finished()) break;
        if (__popc(idle) >= REFILL_MIN || idle == 0xFFFFFFFFu) {
            const int i = feed.take<S1_CHUNK>(!active, lane, &ctl->s1_next, n);
            if (i >= 0) {
                q = __ldg(&order[i]);
                k = seg_of_query(tab, q);
                const float4 t4 = __ldg(&xforms[q]);
                xf = Xf{t4.x, t4.y, t4.z, t4.w};
                ent_i = -1;
                n_ent = tab.seg[k].n_poles + tab.seg[k].n_piers;
                active = true;
                w.depth = -1;  // "entity finished": pick the first one below
            }
        }
        bool survived = false;
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
                    active = false;  // no surrogate collision
                    survived = true;
                }
            }
        }
        // survivors leave in bursts of >= 32 consecutive list slots per warp
        const unsigned sm = __ballot_sync(0xFFFFFFFFu, survived);
        if (sm) {
            if (survived) wb[n_burst + __popc(sm & ((1u << lane) - 1u))] = sv_pack(q, k);
            n_burst += __popc(sm);
            __syncwarp();
            if (n_burst >= 32) {
                int base = 0;
                if (lane == 0) base = atomicAdd(&ctl->n_surv, n_burst);
                base = __shfl_sync(0xFFFFFFFFu, base, 0);
                for (int j = lane; j < n_burst; j += 32) sv.entry[base + j] = wb[j];
                n_burst = 0;
                __syncwarp();
            }
        }
    }
    if (n_burst > 0) {
        int base = 0;
        if (lane == 0) base = atomicAdd(&ctl->n_surv, n_burst);
        base = __shfl_sync(0xFFFFFFFFu, base, 0);
        for (int j = lane; j < n_burst; j += 32) sv.entry[base + j] = wb[j];
    }
}

// JG_MODE_POLY_ONLY: no surrogate stage, every query is a survivor (lbf/benches/quadtree_bench.rs:73-124)
__global__ void __launch_bounds__(S_THREADS)
k_s1_all_survive(const __grid_constant__ SegTable tab, const int* __restrict__ order, DetectCtl* __restrict__ ctl, Survivors sv) {
    const int n = tab.n_queries;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int q = order[i];
        int k = 0;
        for (int j = 1; j < tab.n_seg; j++) k += (q >= tab.seg[j].begin) ? 1 : 0;
        sv.entry[i] = sv_pack(q, k);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) ctl->n_surv = n;
}

// ---------------------------------------------------------------------------------------------
template <bool FILT>
__global__ void __launch_bounds__(S_THREADS)
k_s2_edges(const DevLayout L, const __grid_constant__ SegTable gtab, const float4* __restrict__ xforms,
           const uint32_t* __restrict__ filt, uint8_t* __restrict__ out, DetectCtl* __restrict__ ctl, Survivors sv) {
    __shared__ SegTable tab;
    for (int i = threadIdx.x; i < (int)(sizeof(SegTable) / 4); i += blockDim.x) ((int*)&tab)[i] = ((const int*)&gtab)[i];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int total = ctl->n_surv;

    WarpFeed feed;
    feed.init();
    bool active = false;
    int q = 0, k = 0, e = 0, vroot = 0;
    Xf xf{1, 0, 0, 0};
    Seg sg{0, 0, 0, 0};
    Walk w;
    w.start(0);
    for (;;) {
        const unsigned idle = __ballot_sync(0xFFFFFFFFu, !active);
        if (idle == 0xFFFFFFFFu && feed.finished()) break;
        if (__popc(idle) >= REFILL_MIN || idle == 0xFFFFFFFFu) {
            const int slot = feed.take<S2_CHUNK>(!active, lane, &ctl->s2_next, total);
            if (slot >= 0) {
                const int ent = sv.entry[slot];
                q = ent & 0x3FFFFFF;
                k = (unsigned)ent >> 26;
                const SegDesc& sd = tab.seg[k];
                const float4 t4 = __ldg(&xforms[q]);
                xf = Xf{t4.x, t4.y, t4.z, t4.w};
                // SPolygon::transform_from + generate_bounding_box
                float x0 = 3.402823466e+38f, y0 = 3.402823466e+38f, x1 = -3.402823466e+38f, y1 = -3.402823466e+38f;
                for (int j = 0; j < sd.n_verts; j++) {
                    const float2 v = __ldg(&sd.verts[j]);
                    float x, y;
                    xf_point(xf, v.x, v.y, x, y);
                    x0 = fminf(x0, x), y0 = fminf(y0, y), x1 = fmaxf(x1, x), y1 = fmaxf(y1, y);
                }
                const Rect bb{x0, y0, x1, y1};
                if (rect_relation(L.cde_bbox, bb) != REL_SURROUNDING) {
                    out[q] = JG_VERDICT_POLY;  // cd_engine.rs:161-163
                    sv.vroot[slot] = -1;
                } else {
                    vroot = virtual_root(L, bb);
                    sv.vroot[slot] = vroot;
                    sv.bbox[slot] = make_float4(x0, y0, x1, y1);
                    e = -1;
                    w.depth = -1;  // "edge finished": take edge 0 below
                    active = true;
                }
            }
        }
        if (active) {
            int r = 2;
            if (w.depth >= 0) r = walk_step<Seg, FILT>(L, w, sg, filt);
            if (r == 1) {  // cd_engine.rs:169-172: first colliding edge decides
                out[q] = JG_VERDICT_POLY;
                active = false;
            } else if (r == 2) {
                const SegDesc& sd = tab.seg[k];
                e++;
                if (e < sd.n_verts) {
                    const float2 a = __ldg(&sd.verts[e]);
                    const float2 b = __ldg(&sd.verts[e == sd.n_verts - 1 ? 0 : e + 1]);  // simple_polygon.rs:86-93
                    xf_point(xf, a.x, a.y, sg.sx, sg.sy);
                    xf_point(xf, b.x, b.y, sg.ex, sg.ey);
                    w.start(vroot);
                } else {
                    out[q] = 0;  // no edge hit: k_s3_contain decides
                    active = false;
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
template <bool FILT>
__global__ void __launch_bounds__(S_THREADS)
k_s3_contain(const DevLayout L, const __grid_constant__ SegTable tab, const float4* __restrict__ xforms,
             const uint32_t* __restrict__ filt, uint8_t* __restrict__ out, const DetectCtl* __restrict__ ctl, Survivors sv) {
    const int ns = ctl->n_surv;
    for (int slot = blockIdx.x * blockDim.x + threadIdx.x; slot < ns; slot += gridDim.x * blockDim.x) {
        const int vroot = sv.vroot[slot];
        if (vroot < 0) continue;
        const int ent = sv.entry[slot];
        const int q = ent & 0x3FFFFFF;
        if (out[q] != 0) continue;
        const SegDesc& sd = tab.seg[(unsigned)ent >> 26];
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

size_t detect_bin_words(int n_seg) { return (size_t)n_seg * BIN_ROT * BIN_CELLS; }

cudaError_t launch_detect(const DevLayout& L, const SegTable& tab, const float4* d_xf, const uint32_t* d_filt, int mode,
                          uint8_t* d_out, const DetectScratch& sc, int n_sms, cudaStream_t st, int* n_launches) {
    const int n = tab.n_queries;
    const int n_bins = (int)detect_bin_words(tab.n_seg);
    cudaMemsetAsync(sc.ctl, 0, sizeof(DetectCtl), st);
    cudaMemsetAsync(sc.hist, 0, (size_t)n_bins * sizeof(int), st);
    const int want = (n + S_THREADS - 1) / S_THREADS;
    const int gs = want < n_sms * 8 ? want : n_sms * 8;
    k_bin_count<<<gs, S_THREADS, 0, st>>>(tab, L, d_xf, sc.keys, sc.hist);
    k_bin_scan<<<1, 1024, 0, st>>>(sc.hist, n_bins);
    k_bin_scatter<<<gs, S_THREADS, 0, st>>>(n, sc.keys, sc.hist, sc.order);
    if (mode & JG_MODE_POLY_ONLY) {
        k_s1_all_survive<<<gs, S_THREADS, 0, st>>>(tab, sc.order, sc.ctl, sc.sv);
    } else if (d_filt) {
        const int g = persistent_grid(k_s1_poles<true>, 0, n_sms);
        k_s1_poles<true><<<want < g ? want : g, S_THREADS, 0, st>>>(L, tab, d_xf, sc.order, d_filt, d_out, sc.ctl, sc.sv);
    } else {
        const int g = persistent_grid(k_s1_poles<false>, 1, n_sms);
        k_s1_poles<false><<<want < g ? want : g, S_THREADS, 0, st>>>(L, tab, d_xf, sc.order, d_filt, d_out, sc.ctl, sc.sv);
    }
    if (d_filt) {
        const int g = persistent_grid(k_s2_edges<true>, 2, n_sms);
        k_s2_edges<true><<<want < g ? want : g, S_THREADS, 0, st>>>(L, tab, d_xf, d_filt, d_out, sc.ctl, sc.sv);
        k_s3_contain<true><<<gs, S_THREADS, 0, st>>>(L, tab, d_xf, d_filt, d_out, sc.ctl, sc.sv);
    } else {
        const int g = persistent_grid(k_s2_edges<false>, 3, n_sms);
        k_s2_edges<false><<<want < g ? want : g, S_THREADS, 0, st>>>(L, tab, d_xf, d_filt, d_out, sc.ctl, sc.sv);
        k_s3_contain<false><<<gs, S_THREADS, 0, st>>>(L, tab, d_xf, d_filt, d_out, sc.ctl, sc.sv);
    }
    if (n_launches) *n_launches += 6;
    return cudaGetLastError();
}

}  // namespace jg
