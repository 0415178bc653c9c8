// Device-side sample generation and loss reduction for a batched placement search (sm_100a): the two ends of the
// caller loop of the hot path (lbf/src/opt/search.rs:17-106) moved next to the detect kernels, so that one search round -
// draw n samples, resolve them against the layout, keep the feasible one of minimal LBF loss - costs one launch
// sequence and an 8-byte read-back instead of n records up and n verdicts down.
//
//   k_search_sample   UniformRectSampler::sample (lbf/src/samplers/uniform_rect_sampler.rs:18-45) or LSSampler::sample
//                     (ls_sampler.rs:34-120, normal around the reference transformation) for items with a discrete set of
//                     allowed rotations, written in the compact query form of jg_detect_multi_indexed (rotation index
//                     u8 + translation 2 x f32). Counter-based Philox4x32-10 keyed by (seed, round counter): sample q
//                     of a round is a pure function of (seed, counter, q). The reference's SmallRng stream is not
//                     reproduced (sample order is a third-party detail).
//   k_search_reduce   LBFLoss::from_shape + cost (lbf/src/opt/loss.rs:12-35) of every sample the detect kernels left
//                     free, and the minimum: x_max * 10 + y_max of the transformed shape's bbox. For a fixed rotation
//                     max_i fl(fl(c x_i - s y_i) + tx) = fl(max_i fl(c x_i - s y_i) + tx) (rounding is monotone), so
//                     the per-rotation extents are computed once on the host with the kernels' own arithmetic.
#include <cstdint>

#include "launch.hpp"

namespace jg {

__device__ __forceinline__ void philox_round(uint32_t& c0, uint32_t& c1, uint32_t& c2, uint32_t& c3, uint32_t k0, uint32_t k1) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    c0 = hi1 ^ c1 ^ k0;
    c1 = lo1;
    c2 = hi0 ^ c3 ^ k1;
    c3 = lo0;
}
__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        philox_round(c0, c1, c2, c3, k0, k1);
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}
__device__ __forceinline__ float u01(uint32_t u) { return (float)(u >> 8) * (1.0f / 16777216.0f); }  // [0, 1)

__global__ void __launch_bounds__(256)
k_search_sample(const SearchRound r, long long n, uint8_t* __restrict__ ridx, float2* __restrict__ txy) {
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (long long)gridDim.x * blockDim.x) {
        const uint4 u = philox4x32_10((uint32_t)q, (uint32_t)(q >> 32), (uint32_t)r.counter, (uint32_t)(r.counter >> 32),
                                      (uint32_t)r.seed, (uint32_t)(r.seed >> 32));
        float x, y;
        int ri = r.rot_idx;
        if (r.mode == 0) {  // uniform over the sampling box, rotation uniform over the allowed set
            x = r.box[0] + (r.box[2] - r.box[0]) * u01(u.x);
            y = r.box[1] + (r.box[3] - r.box[1]) * u01(u.y);
            if (ri < 0) ri = (int)(((unsigned long long)u.z * (unsigned)r.n_rots) >> 32);
        } else {  // normal around the reference translation (Box-Muller), rotation kept
            const float rad = sqrtf(-2.0f * __logf(1.0f - u01(u.x)));
            float sn, cs;
            __sincosf(6.28318530718f * u01(u.y), &sn, &cs);
            x = r.mean[0] + r.sd * (rad * cs);
            y = r.mean[1] + r.sd * (rad * sn);
        }
        ridx[q] = (uint8_t)ri;
        txy[q] = make_float2(x, y);
    }
}

__device__ __forceinline__ uint32_t order_key(float f) {  // monotone map of finite floats to unsigned
    const uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

// best[0] = min over free samples of (order_key(loss) << 32 | q), best[1] = number of free samples
__global__ void __launch_bounds__(256)
k_search_reduce(long long n, const uint8_t* __restrict__ verdicts, const uint8_t* __restrict__ ridx, const float2* __restrict__ txy,
                const float2* __restrict__ rot_ext, int n_rots, unsigned long long* __restrict__ best) {
    unsigned long long mine = ~0ull;
    unsigned n_free = 0;
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (long long)gridDim.x * blockDim.x) {
        if (verdicts[q] != 0) continue;
        n_free++;
        const int ri = ridx[q] < n_rots ? ridx[q] : n_rots - 1;
        const float2 e = __ldg(&rot_ext[ri]), t = txy[q];
        const float x_max = e.x + t.x, y_max = e.y + t.y;
        const float loss = x_max * 10.0f + y_max;  // loss.rs:31-33 (no FMA: -fmad=false)
        const unsigned long long key = ((unsigned long long)order_key(loss) << 32) | (unsigned long long)(uint32_t)q;
        mine = key < mine ? key : mine;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xFFFFFFFFu, mine, o);
        mine = other < mine ? other : mine;
        n_free += __shfl_xor_sync(0xFFFFFFFFu, n_free, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (mine != ~0ull) atomicMin(&best[0], mine);
        if (n_free) atomicAdd(&best[1], (unsigned long long)n_free);
    }
}

cudaError_t launch_search_sample(const SearchRound& r, long long n, uint8_t* d_ridx, float2* d_txy, int n_sms, cudaStream_t st) {
    const long long want = (n + 255) / 256;
    const int grid = (int)(want < (long long)n_sms * 8 ? want : (long long)n_sms * 8);
    k_search_sample<<<grid < 1 ? 1 : grid, 256, 0, st>>>(r, n, d_ridx, d_txy);
    return cudaGetLastError();
}
cudaError_t launch_search_reduce(long long n, const uint8_t* d_verdicts, const uint8_t* d_ridx, const float2* d_txy, const float2* d_rot_ext,
                                 int n_rots, unsigned long long* d_best, int n_sms, cudaStream_t st) {
    const long long want = (n + 255) / 256;
    const int grid = (int)(want < (long long)n_sms * 8 ? want : (long long)n_sms * 8);
    k_search_reduce<<<grid < 1 ? 1 : grid, 256, 0, st>>>(n, d_verdicts, d_ridx, d_txy, d_rot_ext, n_rots, d_best);
    return cudaGetLastError();
}

}  // namespace jg
