// Host-visible launch interface of kernels.cu / detect.cu (kept free of kernel bodies so engine.cu stays small).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "../../include/jagua_cuda.h"
#include "kernels.cuh"

namespace jg {

#define JG_MAX_SEG 32

// One segment of a detect batch: the queries [begin, next segment's begin) all place the same item.
struct SegDesc {
    int begin;
    int n_verts, n_poles, n_piers;
    const float2* verts;
    const float4* poles;  // (x, y, r, 0); only the first n_ff poles are counted in n_poles
    const float4* piers;
    float poi_x, poi_y;
    const float2* rots;  // (cos, sin) of the item's allowed rotations (jg_item_set_rotations), or null
    int n_rots;
    int gbeg;  // collect launches: first group of this segment (launch_collect_grid fills it); unused by detect
};
struct SegTable {
    int n_seg;
    int n_queries;
    SegDesc seg[JG_MAX_SEG];
};

// Where the queries of a launch come from: composed (cos, sin, tx, ty) records, or the compact form
// (index into the item's allowed rotations + translation; transformation.rs:131-139 composed on the host once
// per allowed rotation).
struct XfSrc {
    const float4* xf;
    const uint8_t* ridx;
    const float2* txy;
    const ushort2* ij = nullptr;                   // lattice form (JgLattice): translation = (x0 + i * dx, y0 + j * dy)
    float4 lat = {0.0f, 0.0f, 0.0f, 0.0f};         // x0, y0, dx, dy
};
// jg_lattice_points / the lattice query form: one rounding per product and per sum (no FMA on either side)
__host__ __device__ __forceinline__ float lattice_coord(float o, float d, unsigned k) {
#ifdef __CUDA_ARCH__
    return __fadd_rn(o, __fmul_rn((float)k, d));
#else
    volatile float prod = (float)k * d;  // (the host side is compiled -ffp-contract=off as well)
    return o + prod;
#endif
}
__device__ __forceinline__ float4 load_xf(const XfSrc& s, const SegDesc& sd, int q) {
    // L2 loads: in a streamed batch the queries land in device memory while the kernel is running
    if (s.xf) return __ldcg(&s.xf[q]);
    const int r = (int)__ldcg(&s.ridx[q]);
    const float2 cs = __ldg(&sd.rots[r < sd.n_rots ? r : sd.n_rots - 1]);  // out-of-range indices are clamped
    if (s.ij) {
        const ushort2 k = __ldcg(&s.ij[q]);
        return make_float4(cs.x, cs.y, lattice_coord(s.lat.x, s.lat.z, k.x), lattice_coord(s.lat.y, s.lat.w, k.y));
    }
    const float2 t = __ldcg(&s.txy[q]);
    return make_float4(cs.x, cs.y, t.x, t.y);
}

// Work counters of one detect launch sequence (device memory, zeroed by launch_detect).
struct DetectCtl {
    int s1_next;    // next unclaimed query (k_s1_surr takes 32 at a time)
    int s2_next;    // next unclaimed survivor (k_s2_poly takes a group at a time)
    int n_surv;     // survivors of the surrogate stage
    int fallbacks;  // work items resolved by the register-only walk because a warp stack was full
    // fused kernel (k_detect_fused)
    int n_res;       // survivor slots reserved (n_surv = slots published, in order)
    int done_warps;  // warps that have finished all their stage-1 work
    int avail;       // queries whose transforms have arrived in device memory (streamed batches)
    int error;       // 1: a warp gave up waiting (lost copy); the results are incomplete
    // containment pass (k_s2_contain): survivors without an edge hit whose virtual root holds Partial hazards
    int n_ct;        // entries in the containment list
    int ct_next;     // next unclaimed entry
    int pad[2];
};
#define JG_MAX_LAUNCH_QUERIES (1 << 26)  // survivor entries keep the query index in 26 bits
#define JG_MAX_NODES (1 << 20)           // stage-2 pair words keep the node index in >= 20 bits
#define JG_MAX_ITEM_VERTS 4096           // ... and the edge index in 12

// Device scratch of one detect launch sequence (sized for the largest batch piece seen so far).
struct DetectScratch {
    DetectCtl* ctl;
    int* sv_entry;  // survivors of stage 1: query index within the batch (low 26 bits) | segment (high 6 bits)
    // containment list written by k_s2_grid: survivor entry + virtual root node, and the transformed shape's bbox
    int4* ct_meta;
    float4* ct_bb;
    float4* ct_xf;
};
// Per-layout launch parameters (make_detect_plan).
struct DetectPlan {
    int nodes_in_smem;  // node meta / bbox / kind staged into shared memory by every CTA
    int cap_in;         // capacity of a warp's inner-pair stack
    int gbits;          // log2 of the survivors a stage-2 warp resolves together
    int use_grid;       // depth <= 5 and the node boxes match the nested-midpoint grid: k_s1_grid / k_s2_grid
    long long split_min;  // launches of at least this many queries run the containment tests in k_s2_contain
};
DetectPlan make_detect_plan(int n_nodes, int depth);
// stage 2 of the grid path as a programmatic dependent launch behind stage 1 (process-wide; default from JG_PDL)
void set_dependent_launch(int on);
int dependent_launch();
size_t detect_kinds_bytes(int n_nodes);

// per-filter node summary read by the detect kernels; d_filt == nullptr is NoFilter
cudaError_t launch_node_kinds(const DevLayout& L, int depth, const uint32_t* d_filt, uint8_t* d_kinds, cudaStream_t st);
cudaError_t detect_grid_usable(const uint8_t* d_kinds, int n_nodes, cudaStream_t st, int* ok);
const int* detect_grid_ok_ptr(const uint8_t* d_kinds, int n_nodes);  // device address of k_grid_build's verdict
cudaError_t launch_detect(const DevLayout& L, const DetectPlan& plan, const SegTable& tab, const XfSrc& d_xf,
                          const uint32_t* d_filt, const uint8_t* d_kinds, int mode, uint8_t* d_out, const DetectScratch& sc,
                          int n_sms, cudaStream_t st, int* n_launches);
// one fused kernel (k_detect_fused) for a piece served by the grid; avail_all = 0: the queries are still being
// copied and sc.ctl->avail is bumped by the caller's copy stream
bool detect_fused_applicable(const DetectPlan& plan, const SegTable& tab, int mode);
cudaError_t launch_detect_fused(const DevLayout& L, const SegTable& tab, const XfSrc& d_xf, const uint32_t* d_filt,
                                const uint8_t* d_kinds, uint8_t* d_out, const DetectScratch& sc, int n_sms, int avail_all,
                                cudaStream_t st, int* n_launches);
// collect on the grid: ONE launch for all segments of `tab` (fills tab.seg[k].gbeg); d_seed = pre-seeded collectors or null
bool collect_grid_applicable(const DetectPlan& plan, const SegTable& tab, int words);
cudaError_t launch_collect_grid(const DevLayout& L, SegTable& tab, const uint8_t* d_kinds, const float4* d_xf, int mode,
                                const uint32_t* d_seed, uint32_t* d_out_bits, int words, int* d_next, int n_sms, cudaStream_t st,
                                int* n_launches);
cudaError_t launch_collect(const DevLayout& L, const DevItem& it, const float4* d_xf, long long n, int mode, const uint32_t* d_seed,
                           uint32_t* d_out_bits, int words, int group, int n_sms, cudaStream_t st, int* n_launches);
// one round of a batched placement search (search.cu)
struct SearchRound {
    int mode;      // 0 uniform over box, 1 normal around mean
    int rot_idx;   // fixed rotation index, or -1 (uniform mode): uniform over the item's n_rots rotations
    int n_rots;
    float sd;
    float box[4];
    float mean[2];
    unsigned long long seed, counter;
};
cudaError_t launch_search_sample(const SearchRound& r, long long n, uint8_t* d_ridx, float2* d_txy, int n_sms, cudaStream_t st);
cudaError_t launch_search_reduce(long long n, const uint8_t* d_verdicts, const uint8_t* d_ridx, const float2* d_txy, const float2* d_rot_ext,
                                 int n_rots, unsigned long long* d_best, int n_sms, cudaStream_t st);
cudaError_t launch_fp32_nofma(float* d_out, int blocks, int iters, cudaStream_t st);
cudaError_t launch_fill(uint32_t* p, size_t n_words, uint32_t v, int n_sms, cudaStream_t st);

}  // namespace jg
