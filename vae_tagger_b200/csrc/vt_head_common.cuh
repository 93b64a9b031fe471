// Device helpers shared by the tag-head translation units (vt_head.cu, vt_head_train.cu).
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>

namespace vt {

__device__ __forceinline__ float h_warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    return v;
}
__device__ __forceinline__ float h_warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xFFFFFFFFu, v, o));
    return v;
}
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

// block-wide sum / max of one value per thread (256 threads); result valid in all threads
__device__ __forceinline__ float block_sum256(float v, float* red) {
    v = h_warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i];
    return t;
}
__device__ __forceinline__ float block_max256(float v, float* red) {
    v = h_warp_max(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = red[0];
#pragma unroll
    for (int i = 1; i < 8; ++i) t = fmaxf(t, red[i]);
    return t;
}

// counter-based generator of the dropout masks (training step): one 64-bit mix per element
__device__ __forceinline__ unsigned long long h_splitmix64(unsigned long long z) {
    z += 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
// keep-mask of nn.Dropout(p): element `idx` of mask stream `stream` is kept when u >= p, u uniform [0,1)
__device__ __forceinline__ bool h_dropout_keep(unsigned long long seed, unsigned stream, unsigned long long idx,
                                               float p) {
    const unsigned long long r = h_splitmix64(seed ^ h_splitmix64((static_cast<unsigned long long>(stream) << 48) ^ idx));
    return static_cast<float>(r >> 40) * (1.0f / 16777216.0f) >= p;
}

}  // namespace vt
