// Query kernels (sm_100a): one-pass collect, the FP32 non-FMA peak probe and the L2 flush fill.
// The detect pipeline lives in detect.cu.
#include "kernels.cuh"
#include "launch.hpp"

namespace jg {

// ---------------------------------------------------------------------------------------------
// collect — G lanes per query, hazard bitset in shared memory.
//   collect_surrogate_collisions (cd_engine.rs:296-310) then collect_poly_collisions
//   (cd_engine.rs:261-289) into ONE collector (lbf/benches/ci_bench.rs:56-70). The set is a union
//   over independent walks, so lanes may run them in any order.
// ---------------------------------------------------------------------------------------------
template <int G>
__global__ void __launch_bounds__(256)
k_collect(DevLayout L, DevItem it, const float4* __restrict__ xforms, long long n, int mode, const uint32_t* __restrict__ seed_bits,
          uint32_t* __restrict__ out_bits, int words) {
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
        for (int i = gl; i < words; i += G) bits[i] = seed_bits ? __ldg(&seed_bits[q * words + i]) : 0u;  // pre-seeded collector
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
template <class K>
static int grid_for(K kernel, int slot, int n_sms, size_t smem) {
    // (not cached: the shared-memory size varies with the item, and the >48 KB opt-in is a per-device attribute)
    (void)slot;
    if (smem > 48 * 1024) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int b = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, kernel, 256, smem);
    return n_sms * (b < 1 ? 1 : b);
}

cudaError_t launch_collect(const DevLayout& L, const DevItem& it, const float4* d_xf, long long n, int mode, const uint32_t* d_seed,
                           uint32_t* d_out_bits, int words, int group, int n_sms, cudaStream_t st, int* n_launches) {
    const size_t smem = (size_t)it.n_verts * sizeof(float2) + (size_t)(256 / group) * words * sizeof(uint32_t);
#define JG_COLLECT(G, SLOT)                                                                          \
    {                                                                                                \
        int g = grid_for(k_collect<G>, SLOT, n_sms, smem);                                           \
        long long w2 = (n * G + 255) / 256;                                                          \
        k_collect<G><<<(int)(w2 < g ? w2 : g), 256, smem, st>>>(L, it, d_xf, n, mode, d_seed, d_out_bits, words); \
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
