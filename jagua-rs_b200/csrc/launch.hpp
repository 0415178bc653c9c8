// Host-visible launch interface of kernels.cu (kept free of kernel bodies so engine.cu stays small).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "../../include/jagua_cuda.h"
#include "kernels.cuh"

namespace jg {

// Queries that passed the surrogate stage and the bbox test, appended by stage 1 in arbitrary order
// (results are written by query index, so the order is never observable).
struct SurvivorList {
    int* count;
    int* query;     // query index within the batch
    int* vroot;     // virtual-root node of the transformed bbox
    float4* bbox;   // transformed bbox
};

cudaError_t launch_detect(const DevLayout& L, const DevItem& it, const float4* d_xf, long long n, const uint32_t* d_filt,
                          int mode, uint8_t* d_out, const SurvivorList& sv, int group, int n_sms, cudaStream_t st,
                          int* n_launches);
cudaError_t launch_collect(const DevLayout& L, const DevItem& it, const float4* d_xf, long long n, int mode,
                           uint32_t* d_out_bits, int words, int group, int n_sms, cudaStream_t st, int* n_launches);
cudaError_t launch_fp32_nofma(float* d_out, int blocks, int iters, cudaStream_t st);
cudaError_t launch_fill(uint32_t* p, size_t n_words, uint32_t v, int n_sms, cudaStream_t st);

}  // namespace jg
