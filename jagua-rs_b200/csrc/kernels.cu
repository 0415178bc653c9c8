// Query kernels (sm_100a): two-stage detect, one-pass collect, plus the FP32 non-FMA peak probe.
#include "kernels.cuh"
#include "launch.hpp"

namespace jg {

// ---------------------------------------------------------------------------------------------
// detect, stage 1 — one thread per query.
//   detect_surrogate_collision (cd_engine.rs:204-223): poles then piers, each from the ROOT.
//   Survivors get the transformed bbox (simple_polygon.rs:226-253 + :122-133), the
//   `bbox.relation_to(shape.bbox) != Surrounding -> true` test of detect_poly_collision
//   (cd_engine.rs:161-163) and their virtual root (cd_engine.rs:314-327), then are appended to the
//   survivor list that stage 2 consumes.
// ---------------------------------------------------------------------------------------------
template <bool FILT>
__global__ void __launch_bounds__(256)
k_detect_stage1(DevLayout L, DevItem it, const float4* __restrict__ xforms, long long n, const uint32_t* __restrict__ filt,
                int mode, uint8_t* __restrict__ out, SurvivorList sv) {
    extern __shared__ float2 s_verts[];
    for (int i = threadIdx.x; i < it.n_verts; i += blockDim.x) s_verts[i] = it.verts[i];
    __syncthreads();

    const int lane = threadIdx.x & 31;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long base = (long long)blockIdx.x * blockDim.x + (threadIdx.x - lane); base < n; base += stride) {
        const long long q = base + lane;
        const bool valid = q < n;
        bool survivor = false;
        int vroot = 0;
        Rect bb{0, 0, 0, 0};
        if (valid) {
            const float4 t4 = __ldg(&xforms[q]);
            const Xf xf{t4.x, t4.y, t4.z, t4.w};
            bool hit = false;
            if (!(mode & JG_MODE_POLY_ONLY)) {
                for (int p = 0; p < it.n_poles && !hit; p++) {
                    const float4 pl = __ldg(&it.poles[p]);
                    Circ c;
                    xf_point(xf, pl.x, pl.y, c.cx, c.cy);  // circle.rs:79-85: radius unchanged
                    c.r = pl.z;
                    hit = qt_collides<Circ, FILT>(L, 0, c, filt);
                }
                for (int p = 0; p < it.n_piers && !hit; p++) {
                    const float4 pr = __ldg(&it.piers[p]);
                    Seg s;
                    xf_point(xf, pr.x, pr.y, s.sx, s.sy);  // edge.rs:126-134
                    xf_point(xf, pr.z, pr.w, s.ex, s.ey);
                    hit = qt_collides<Seg, FILT>(L, 0, s, filt);
                }
            }
            if (hit) {
                out[q] = JG_VERDICT_SURROGATE;
            } else {
                float x0 = 3.402823466e+38f, y0 = 3.402823466e+38f, x1 = -3.402823466e+38f, y1 = -3.402823466e+38f;
                for (int i = 0; i < it.n_verts; i++) {
                    const float2 v = s_verts[i];
                    float x, y;
                    xf_point(xf, v.x, v.y, x, y);
                    x0 = fminf(x0, x), y0 = fminf(y0, y), x1 = fmaxf(x1, x), y1 = fmaxf(y1, y);
                }
                bb = Rect{x0, y0, x1, y1};
                if (rect_relation(L.cde_bbox, bb) != REL_SURROUNDING) {
                    out[q] = JG_VERDICT_POLY;
                } else {
                    survivor = true;
                    vroot = virtual_root(L, bb);
                }
            }
        }
        // warp-aggregated append
        const unsigned m = __ballot_sync(0xFFFFFFFFu, survivor);
        if (m) {
            int pos = 0;
            if (lane == 0) pos = atomicAdd(sv.count, __popc(m));
            pos = __shfl_sync(0xFFFFFFFFu, pos, 0) + __popc(m & ((1u << lane) - 1u));
            if (survivor) {
                sv.query[pos] = (int)q;
                sv.vroot[pos] = vroot;
                sv.bbox[pos] = make_float4(bb.x_min, bb.y_min, bb.x_max, bb.y_max);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// detect, stage 2 — G lanes per surviving query (G = 4, 8, 16, 32).
//   detect_poly_collision (cd_engine.rs:164-196): every lane owns edges lane, lane+G, ... of the
//   transformed polygon and walks them from the virtual root (QTNode::collides); a ballot over the
//   group is the early exit. With no edge hit: an Entire entry in the virtual root is the
//   reference's `unreachable!` (bit7), otherwise the lanes split the virtual root's relevant Partial
//   entries for detect_containment_collision.
// ---------------------------------------------------------------------------------------------
template <int G, bool FILT>
__global__ void __launch_bounds__(256)
k_detect_stage2(DevLayout L, DevItem it, const float4* __restrict__ xforms, const uint32_t* __restrict__ filt,
                uint8_t* __restrict__ out, SurvivorList sv) {
    extern __shared__ float2 s_verts[];
    for (int i = threadIdx.x; i < it.n_verts; i += blockDim.x) s_verts[i] = it.verts[i];
    __syncthreads();

    const int n_surv = *sv.count;
    const int lane = threadIdx.x & 31;
    const int gl = lane & (G - 1);  // lane within the group
    const unsigned gmask = (G == 32) ? 0xFFFFFFFFu : (((1u << G) - 1u) << (lane & ~(G - 1)));
    const int groups_per_block = blockDim.x / G;
    const int n_groups = gridDim.x * groups_per_block;
    const int V = it.n_verts;
    for (int s = blockIdx.x * groups_per_block + threadIdx.x / G; s < n_surv; s += n_groups) {
        const int q = sv.query[s];
        const int vroot = sv.vroot[s];
        const float4 t4 = __ldg(&xforms[q]);
        const Xf xf{t4.x, t4.y, t4.z, t4.w};
        bool hit = false;
        for (int e0 = 0; e0 < V; e0 += G) {
            const int e = e0 + gl;
            bool h = false;
            if (e < V) {
                const float2 a = s_verts[e];
                const float2 b = s_verts[e == V - 1 ? 0 : e + 1];  // simple_polygon.rs:86-93
                Seg sg;
                xf_point(xf, a.x, a.y, sg.sx, sg.sy);
                xf_point(xf, b.x, b.y, sg.ex, sg.ey);
                h = qt_collides<Seg, FILT>(L, vroot, sg, filt);
            }
            if (__ballot_sync(gmask, h) & gmask) { hit = true; break; }
        }
        uint8_t verdict = hit ? JG_VERDICT_POLY : 0;
        if (!hit) {
            const int4 m = __ldg(&L.node_meta[vroot]);
            const int n_entire = m.z & 0xFFFF, n_partial = (m.z >> 16) & 0xFFFF;
            if (n_entire > 0) {
                verdict = JG_VERDICT_REF_PANIC;  // cd_engine.rs:179 (only reachable under a filter)
            } else {
                const float4 b4 = sv.bbox[s];
                const Rect sb{b4.x, b4.y, b4.z, b4.w};
                bool c = false;
                for (int i = gl; i < n_partial && !c; i += G) {
                    const int slot = __ldg(&L.entry_meta[m.y + i]).x;
                    if (FILT && filt_bit(filt, slot)) continue;
                    c = containment_collision(L, s_verts, V, it.poi_x, it.poi_y, xf, sb, slot);
                }
                if (__ballot_sync(gmask, c) & gmask) verdict = JG_VERDICT_POLY;
            }
        }
        if (gl == 0) out[q] = verdict;
    }
}

// ---------------------------------------------------------------------------------------------
// collect — G lanes per query, hazard bitset in shared memory.
//   collect_surrogate_collisions (cd_engine.rs:296-310) then collect_poly_collisions
//   (cd_engine.rs:261-289) into ONE collector (lbf/benches/ci_bench.rs:56-70). The set is a union
//   over independent walks, so lanes may run them in any order.
// ---------------------------------------------------------------------------------------------
template <int G>
__global__ void __launch_bounds__(256)
k_collect(DevLayout L, DevItem it, const float4* __restrict__ xforms, long long n, int mode, uint32_t* __restrict__ out_bits,
          int words) {
    extern __shared__ float2 s_dyn[];
    float2* s_verts = s_dyn;
    uint32_t* s_bits = (uint32_t*)(s_dyn + it.n_verts);
    for (int i = threadIdx.x; i < it.n_verts; i += blockDim.x) s_verts[i] = it.verts[i];
    __syncthreads();

    const int lane = threadIdx.x & 31;
    const int gl = lane & (G - 1);
    const unsigned gmask = (G == 32) ? 0xFFFFFFFFu : (((1u << G) - 1u) << (lane & ~(G - 1)));
    const int groups_per_block = blockDim.x / G;
    const long long n_groups = (long long)gridDim.x * groups_per_block;
    uint32_t* bits = s_bits + (threadIdx.x / G) * words;
    const int V = it.n_verts;
    for (long long q = (long long)blockIdx.x * groups_per_block + threadIdx.x / G; q < n; q += n_groups) {
        for (int i = gl; i < words; i += G) bits[i] = 0;
        __syncwarp(gmask);
        const float4 t4 = __ldg(&xforms[q]);
        const Xf xf{t4.x, t4.y, t4.z, t4.w};
        if (!(mode & JG_MODE_POLY_ONLY)) {
            for (int p = gl; p < it.n_poles + it.n_piers; p += G) {
                if (p < it.n_poles) {
                    const float4 pl = __ldg(&it.poles[p]);
                    Circ c;
                    xf_point(xf, pl.x, pl.y, c.cx, c.cy);
                    c.r = pl.z;
                    qt_collect<Circ>(L, 0, c, bits);
                } else {
                    const float4 pr = __ldg(&it.piers[p - it.n_poles]);
                    Seg sg;
                    xf_point(xf, pr.x, pr.y, sg.sx, sg.sy);
                    xf_point(xf, pr.z, pr.w, sg.ex, sg.ey);
                    qt_collect<Seg>(L, 0, sg, bits);
                }
            }
        }
        // transformed bbox: lanes fold their vertices, then a butterfly over the group (min/max are
        // exact and order-free)
        float x0 = 3.402823466e+38f, y0 = 3.402823466e+38f, x1 = -3.402823466e+38f, y1 = -3.402823466e+38f;
        for (int i = gl; i < V; i += G) {
            const float2 v = s_verts[i];
            float x, y;
            xf_point(xf, v.x, v.y, x, y);
            x0 = fminf(x0, x), y0 = fminf(y0, y), x1 = fmaxf(x1, x), y1 = fmaxf(y1, y);
        }
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) {
            x0 = fminf(x0, __shfl_xor_sync(gmask, x0, o));
            y0 = fminf(y0, __shfl_xor_sync(gmask, y0, o));
            x1 = fmaxf(x1, __shfl_xor_sync(gmask, x1, o));
            y1 = fmaxf(y1, __shfl_xor_sync(gmask, y1, o));
        }
        const Rect sb{x0, y0, x1, y1};
        if (gl == 0 && rect_relation(L.cde_bbox, sb) != REL_SURROUNDING)
            atomicOr(bits + (L.exterior_slot >> 5), 1u << (L.exterior_slot & 31));
        const int vroot = virtual_root(L, sb);
        for (int e = gl; e < V; e += G) {
            const float2 a = s_verts[e];
            const float2 b = s_verts[e == V - 1 ? 0 : e + 1];
            Seg sg;
            xf_point(xf, a.x, a.y, sg.sx, sg.sy);
            xf_point(xf, b.x, b.y, sg.ex, sg.ey);
            qt_collect<Seg>(L, vroot, sg, bits);
        }
        __syncwarp(gmask);
        {  // containment for the virtual root's Partial entries not collected yet
            const int4 m = __ldg(&L.node_meta[vroot]);
            const int n_entire = m.z & 0xFFFF, n_partial = (m.z >> 16) & 0xFFFF;
            for (int i = gl; i < n_partial; i += G) {
                const int slot = __ldg(&L.entry_meta[m.y + n_entire + i]).x;
                if ((((volatile uint32_t*)bits)[slot >> 5] >> (slot & 31)) & 1u) continue;
                if (containment_collision(L, s_verts, V, it.poi_x, it.poi_y, xf, sb, slot))
                    atomicOr(bits + (slot >> 5), 1u << (slot & 31));
            }
        }
        __syncwarp(gmask);
        for (int i = gl; i < words; i += G) out_bits[q * words + i] = ((volatile uint32_t*)bits)[i];
        __syncwarp(gmask);
    }
}

// ---------------------------------------------------------------------------------------------
// FP32 non-FMA peak probe: independent FADD / FMUL chains (what -fmad=false code can issue at best).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_fp32_nofma(float* out, int iters, float a, float b) {
    float r0 = threadIdx.x, r1 = r0 + 1, r2 = r0 + 2, r3 = r0 + 3, r4 = r0 + 4, r5 = r0 + 5, r6 = r0 + 6, r7 = r0 + 7;
    for (int i = 0; i < iters; i++) {
        r0 = __fadd_rn(r0, a); r1 = __fmul_rn(r1, b); r2 = __fadd_rn(r2, a); r3 = __fmul_rn(r3, b);
        r4 = __fadd_rn(r4, a); r5 = __fmul_rn(r5, b); r6 = __fadd_rn(r6, a); r7 = __fmul_rn(r7, b);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = ((r0 + r1) + (r2 + r3)) + ((r4 + r5) + (r6 + r7));
}

__global__ void k_fill(uint32_t* p, size_t n, uint32_t v) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}

// ---------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------
static int g_blocks_per_sm_cache[64];

template <class K>
static int grid_for(K kernel, int slot, int n_sms, size_t smem) {
    if (g_blocks_per_sm_cache[slot] == 0) {
        int b = 1;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, kernel, 256, smem);
        g_blocks_per_sm_cache[slot] = b < 1 ? 1 : b;
    }
    return n_sms * g_blocks_per_sm_cache[slot];
}

cudaError_t launch_detect(const DevLayout& L, const DevItem& it, const float4* d_xf, long long n, const uint32_t* d_filt,
                          int mode, uint8_t* d_out, const SurvivorList& sv, int group, int n_sms, cudaStream_t st,
                          int* n_launches) {
    const size_t smem = (size_t)it.n_verts * sizeof(float2);
    cudaMemsetAsync(sv.count, 0, sizeof(int), st);
    long long want = (n + 255) / 256;
    if (d_filt) {
        int g = grid_for(k_detect_stage1<true>, 0, n_sms, smem);
        k_detect_stage1<true><<<(int)(want < g ? want : g), 256, smem, st>>>(L, it, d_xf, n, d_filt, mode, d_out, sv);
    } else {
        int g = grid_for(k_detect_stage1<false>, 1, n_sms, smem);
        k_detect_stage1<false><<<(int)(want < g ? want : g), 256, smem, st>>>(L, it, d_xf, n, d_filt, mode, d_out, sv);
    }
#define JG_STAGE2(G, F, SLOT)                                                                                  \
    {                                                                                                          \
        int g = grid_for(k_detect_stage2<G, F>, SLOT, n_sms, smem);                                            \
        long long w2 = (n * G + 255) / 256;                                                                    \
        k_detect_stage2<G, F><<<(int)(w2 < g ? w2 : g), 256, smem, st>>>(L, it, d_xf, d_filt, d_out, sv);      \
    }
    if (d_filt) {
        switch (group) {
            case 4: JG_STAGE2(4, true, 2) break;
            case 8: JG_STAGE2(8, true, 3) break;
            case 16: JG_STAGE2(16, true, 4) break;
            default: JG_STAGE2(32, true, 5) break;
        }
    } else {
        switch (group) {
            case 4: JG_STAGE2(4, false, 6) break;
            case 8: JG_STAGE2(8, false, 7) break;
            case 16: JG_STAGE2(16, false, 8) break;
            default: JG_STAGE2(32, false, 9) break;
        }
    }
#undef JG_STAGE2
    if (n_launches) *n_launches += 2;
    return cudaGetLastError();
}

cudaError_t launch_collect(const DevLayout& L, const DevItem& it, const float4* d_xf, long long n, int mode,
                           uint32_t* d_out_bits, int words, int group, int n_sms, cudaStream_t st, int* n_launches) {
    const size_t smem = (size_t)it.n_verts * sizeof(float2) + (size_t)(256 / group) * words * sizeof(uint32_t);
#define JG_COLLECT(G, SLOT)                                                                          \
    {                                                                                                \
        int g = grid_for(k_collect<G>, SLOT, n_sms, smem);                                           \
        long long w2 = (n * G + 255) / 256;                                                          \
        k_collect<G><<<(int)(w2 < g ? w2 : g), 256, smem, st>>>(L, it, d_xf, n, mode, d_out_bits, words); \
    }
    switch (group) {
        case 4: JG_COLLECT(4, 10) break;
        case 8: JG_COLLECT(8, 11) break;
        case 16: JG_COLLECT(16, 12) break;
        default: JG_COLLECT(32, 13) break;
    }
#undef JG_COLLECT
    if (n_launches) *n_launches += 1;
    return cudaGetLastError();
}

cudaError_t launch_fp32_nofma(float* d_out, int blocks, int iters, cudaStream_t st) {
    k_fp32_nofma<<<blocks, 256, 0, st>>>(d_out, iters, 1.0000001f, 0.9999999f);
    return cudaGetLastError();
}

cudaError_t launch_fill(uint32_t* p, size_t n_words, uint32_t v, int n_sms, cudaStream_t st) {
    k_fill<<<n_sms * 8, 256, 0, st>>>(p, n_words, v);
    return cudaGetLastError();
}

}  // namespace jg
