// Hardware probe (development aid): MN-major SWIZZLE_128B tcgen05 operands straight from NHWC tiles -- the layout a
// weight-gradient GEMM wants (K = pixels, M / N = channels, both operands channel-contiguous).
//   A_g [K = 64 pixels][M = 128 channels] bf16, loaded as two TMA boxes of (64 channels x 64 pixels) = 64 rows x 128 B each
//   B_g [K = 64 pixels][N = 64 channels]
//   D[m][n] = sum_k A_g[k][m] * B_g[k][n]
// Hypothesis (CUTLASS canonical MN-major B128 layout ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units): LBO = byte
// distance between 64-element MN blocks, SBO = byte distance between 8-row K groups (1024), K step of 16 = +2048 bytes,
// instruction-descriptor bits 15 / 16 = A / B MN-major.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_mn_probe tools/umma_mn_probe.cu -lcuda
#include <cuda.h>
#include <cuda_bf16.h>
#include <cudaTypedefs.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "../vae_tagger_b200/csrc/vt_ptx.cuh"
using namespace vt;

__device__ __forceinline__ uint64_t desc_mn(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((addr >> 4) & 0x3FFF);
    d |= static_cast<uint64_t>((lbo >> 4) & 0x3FFF) << 16;
    d |= static_cast<uint64_t>((sbo >> 4) & 0x3FFF) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}

__global__ void __launch_bounds__(128, 1)
probe(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, float* out, int lbo, int sbo, int kstep,
      int amaj, int bmaj, int bfmt_f16) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;                 // two boxes: 2 x (64 rows x 128 B)
    uint8_t* sB = smem + 16384;         // 64 rows x 128 B
    uint64_t* bar = reinterpret_cast<uint64_t*>(sB + 8192);
    uint32_t* tptr = reinterpret_cast<uint32_t*>(bar + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); fence_mbar_init(); }
    if (warp == 0) { tmem_alloc(tptr, 64); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tptr;
    if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(&bar[0], 3 * 8192);
        tma_load_3d(sA, &tmA, &bar[0], 0, 0, 0);
        tma_load_3d(sA + 8192, &tmA, &bar[0], 64, 0, 0);
        tma_load_3d(sB, &tmB, &bar[0], 0, 0, 0);
        mbar_wait(&bar[0], 0);
        tc_fence_after();
        uint32_t idesc = umma_idesc_bf16(128, 64);
        if (amaj) idesc |= 1u << 15;
        if (bmaj) idesc |= 1u << 16;
        if (bfmt_f16) idesc &= ~(7u << 10);   // B format = F16 while A stays BF16: is a mixed pair legal?
        for (int k = 0; k < 4; ++k) {
            const uint64_t da = desc_mn(smem_u32(sA) + k * kstep, lbo, sbo);
            const uint64_t db = desc_mn(smem_u32(sB) + k * kstep, lbo, sbo);
            umma_bf16_ss(tmem, da, db, idesc, k != 0);
        }
        umma_commit(&bar[1]);
    }
    mbar_wait(&bar[1], 0);
    tc_fence_after();
    uint32_t r[32];
    for (int j = 0; j < 2; ++j) {
        tmem_ld_32x32(tmem + j * 32 + (static_cast<uint32_t>(warp * 32) << 16), r);
        tmem_ld_wait();
        for (int i = 0; i < 32; ++i) out[(warp * 32 + lane) * 64 + j * 32 + i] = __uint_as_float(r[i]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 64);
}

static PFN_cuTensorMapEncodeTiled_v12000 enc() {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    return reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
}

int main() {
    const int K = 64, M = 128, N = 64;
    std::vector<__nv_bfloat16> hA(K * M), hB(K * N);
    std::vector<float> fA(K * M), fB(K * N);
    unsigned s = 12345;
    auto rnd = [&]() { s = s * 1664525u + 1013904223u; return static_cast<float>(static_cast<int>((s >> 20) % 7) - 3); };
    for (int i = 0; i < K * M; ++i) { fA[i] = rnd(); hA[i] = __float2bfloat16(fA[i]); }
    for (int i = 0; i < K * N; ++i) { fB[i] = rnd(); hB[i] = __float2bfloat16(fB[i]); }
    std::vector<float> ref(M * N, 0.f);
    for (int k = 0; k < K; ++k) for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) ref[m * N + n] += fA[k * M + m] * fB[k * N + n];
    __nv_bfloat16 *dA, *dB; float* dO;
    cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dO, M * N * 4);
    cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
    CUtensorMap tA, tB;
    uint32_t es[3] = {1, 1, 1};
    {
        uint64_t dims[3] = {M, K, 1}; uint64_t str[2] = {2ull * M, 2ull * M * K}; uint32_t box[3] = {64, 64, 1};
        CUresult r = enc()(&tA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, dA, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r) { printf("encode A failed %d\n", r); return 1; }
    }
    {
        uint64_t dims[3] = {N, K, 1}; uint64_t str[2] = {2ull * N, 2ull * N * K}; uint32_t box[3] = {64, 64, 1};
        CUresult r = enc()(&tB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, dB, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r) { printf("encode B failed %d\n", r); return 1; }
    }
    const int smem = 16384 + 8192 + 1024 + 256;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    std::vector<float> hO(M * N);
    struct Cfg { int lbo, sbo, kstep, amaj, bmaj, bf16b; };
    std::vector<Cfg> cfgs = {{8192, 1024, 2048, 1, 1, 0}, {8192, 1024, 2048, 1, 1, 1}};
    for (auto c : cfgs) {
        cudaMemset(dO, 0, M * N * 4);
        probe<<<1, 128, smem>>>(tA, tB, dO, c.lbo, c.sbo, c.kstep, c.amaj, c.bmaj, c.bf16b);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("lbo=%d sbo=%d kstep=%d b_is_f16=%d: CUDA error %s\n", c.lbo, c.sbo, c.kstep, c.bf16b, cudaGetErrorString(e)); return 2; }
        cudaMemcpy(hO.data(), dO, hO.size() * 4, cudaMemcpyDeviceToHost);
        int ok = 0, rows_lo = 0, rows_hi = 0;
        for (int m = 0; m < M; ++m) {
            bool all = true;
            for (int n = 0; n < N; ++n) { const bool eq = hO[m * N + n] == ref[m * N + n]; ok += eq; all &= eq; }
            (m < 64 ? rows_lo : rows_hi) += all;
        }
        printf("lbo=%5d sbo=%5d kstep=%5d amaj=%d bmaj=%d b_is_f16=%d : %5d/%d elements, rows m<64 ok %d/64, m>=64 ok %d/64\n", c.lbo, c.sbo,
               c.kstep, c.amaj, c.bmaj, c.bf16b, ok, M * N, rows_lo, rows_hi);
    }
    return 0;
}
