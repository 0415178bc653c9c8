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
};
struct SegTable {
    int n_seg;
    int n_queries;
    SegDesc seg[JG_MAX_SEG];
};

// Work counters of one detect launch sequence (device memory, zeroed by launch_detect).
struct DetectCtl {
    int s1_next;  // next unclaimed position of the sorted query order (k_s1_poles)
    int s2_next;  // next unclaimed survivor (k_s2_edges)
    int n_surv;   // survivors of the surrogate stage
    int pad;
};

// Queries that passed the surrogate stage. Appended in warp-sized bursts in arbitrary order (results are
// written by query index, so the order is never observable).
struct Survivors {
    int* entry;    // query index within the batch (low 26 bits) | segment (high 6 bits)
    int* vroot;    // virtual-root node of the transformed bbox, -1 = bbox not inside the CDE bbox
    float4* bbox;  // transformed bbox
};
#define JG_MAX_LAUNCH_QUERIES (1 << 26)

// Device scratch of one detect launch sequence (sized for the largest batch piece seen so far).
struct DetectScratch {
    DetectCtl* ctl;
    int* keys;   // bin key per query
    int* order;  // queries sorted by bin
    int* hist;   // bin counters / cursors, detect_bin_words(n_seg) ints
    Survivors sv;
};
size_t detect_bin_words(int n_seg);

cudaError_t launch_detect(const DevLayout& L, const SegTable& tab, const float4* d_xf, const uint32_t* d_filt, int mode,
                          uint8_t* d_out, const DetectScratch& sc, int n_sms, cudaStream_t st, int* n_launches);
cudaError_t launch_collect(const DevLayout& L, const DevItem& it, const float4* d_xf, long long n, int mode,
                           uint32_t* d_out_bits, int words, int group, int n_sms, cudaStream_t st, int* n_launches);
cudaError_t launch_fp32_nofma(float* d_out, int blocks, int iters, cudaStream_t st);
cudaError_t launch_fill(uint32_t* p, size_t n_words, uint32_t v, int n_sms, cudaStream_t st);

}  // namespace jg
