// Hardware probe (development aid): does a K-major SWIZZLE_128B tcgen05 operand descriptor accept
//   (a) a start address that is only 128-byte aligned (row-shifted view of a TMA-written tile),
//   (b) a stride-byte-offset (8-row group pitch) other than 1024 bytes,
// and what must the descriptor's base_offset field be?  B is an identity matrix, so D[m][n] shows
// which shared-memory row / 16-byte chunk each accumulator element was read from.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_probe tools/umma_probe.cu -lcuda
#include <cuda.h>
#include <cuda_bf16.h>
#include <cudaTypedefs.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "../vae_tagger_b200/csrc/vt_ptx.cuh"
using namespace vt;

constexpr int ROWS = 384;

__global__ void __launch_bounds__(128, 1)
probe(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, float* out, int r0,
      int sbo_bytes, int base_off_mode, int use_lbo) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;                 // ROWS x 128 B
    uint8_t* sB = smem + 512 * 128;     // 64 x 128 B (A region holds two 256-row boxes)
    uint64_t* bar = reinterpret_cast<uint64_t*>(sB + 64 * 128);
    uint32_t* tptr = reinterpret_cast<uint32_t*>(bar + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        fence_mbar_init();
    }
    if (warp == 0) { tmem_alloc(tptr, 64); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tptr;
    if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(&bar[0], 2 * 256 * 128 + 64 * 128);
        tma_load_3d(sA, &tmA, &bar[0], 0, 0, 0);
        tma_load_3d(sA + 256 * 128, &tmA, &bar[0], 0, 256, 0);   // rows 256.. (box 128 rows; OOB rows zero)
        tma_load_3d(sB, &tmB, &bar[0], 0, 0, 0);
        mbar_wait(&bar[0], 0);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(sA) + r0 * 128;
        const uint32_t bo = base_off_mode ? ((a_addr >> 7) & 7) : 0;
        uint64_t da = umma_desc_k_sw128(a_addr, sbo_bytes, bo);
        if (use_lbo) da = (da & ~(0x3FFFull << 16)) | (static_cast<uint64_t>(use_lbo) << 16);
        const uint64_t db = umma_desc_k_sw128(smem_u32(sB));
        const uint32_t idesc = umma_idesc_bf16(128, 64);
        for (int k = 0; k < 4; ++k) umma_bf16_ss(tmem, da + 2 * k, db + 2 * k, idesc, k != 0);
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
    std::vector<__nv_bfloat16> hA(ROWS * 64), hB(64 * 64);
    auto val = [](int r, int c) { return static_cast<float>(((r * 64 + c) % 251) - 125); };
    for (int r = 0; r < ROWS; ++r) for (int c = 0; c < 64; ++c) hA[r * 64 + c] = __float2bfloat16(val(r, c));
    for (int n = 0; n < 64; ++n) for (int k = 0; k < 64; ++k) hB[n * 64 + k] = __float2bfloat16(n == k ? 1.f : 0.f);
    __nv_bfloat16 *dA, *dB; float* dO;
    cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dO, 128 * 64 * 4);
    cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
    CUtensorMap tA, tB;
    uint32_t es[3] = {1, 1, 1};
    {
        uint64_t dims[3] = {64, ROWS, 1}; uint64_t str[2] = {128, 128ull * ROWS}; uint32_t box[3] = {64, 256, 1};
        // second call loads rows 256..511 with the same box: rows >= ROWS are zero filled
        CUresult r = enc()(&tA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, dA, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r) { printf("encode A failed %d\n", r); return 1; }
    }
    {
        uint64_t dims[3] = {64, 64, 1}; uint64_t str[2] = {128, 128 * 64}; uint32_t box[3] = {64, 64, 1};
        CUresult r = enc()(&tB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, dB, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r) { printf("encode B failed %d\n", r); return 1; }
    }
    // the kernel's second A load writes 256 rows at row 256: smem must hold 512 rows for safety
    const int smem = 512 * 128 + 64 * 128 + 1024 + 256;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    std::vector<float> hO(128 * 64);
    struct Cfg { int r0, sbo, bom, lbo; };
    std::vector<Cfg> cfgs;
    for (int bom = 0; bom < 2; ++bom) {
        cfgs.push_back({0, 1024, bom, 0});
        for (int r0 : {1, 3, 8, 11}) cfgs.push_back({r0, 1024, bom, 0});
        cfgs.push_back({0, 1280, bom, 0});
        cfgs.push_back({0, 2304, bom, 0});
        for (int r0 : {1, 10, 19, 37}) { cfgs.push_back({r0, 1280, bom, 0}); cfgs.push_back({r0, 2304, bom, 0}); }
    }
    for (auto c : cfgs) {
        cudaMemset(dO, 0, 128 * 64 * 4);
        probe<<<1, 128, smem>>>(tA, tB, dO, c.r0, c.sbo, c.bom, c.lbo);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("r0=%d sbo=%d bom=%d: CUDA error %s\n", c.r0, c.sbo, c.bom, cudaGetErrorString(e)); return 2; }
        cudaMemcpy(hO.data(), dO, hO.size() * 4, cudaMemcpyDeviceToHost);
        int ok = 0, row_ok = 0;
        int first_bad_m = -1;
        for (int m = 0; m < 128; ++m) {
            const int src = c.r0 + (m / 8) * (c.sbo / 128) + (m % 8);
            bool all = true;
            for (int n = 0; n < 64; ++n) {
                const bool eq = hO[m * 64 + n] == val(src, n);
                ok += eq; all &= eq;
            }
            row_ok += all;
            if (!all && first_bad_m < 0) first_bad_m = m;
        }
        printf("r0=%2d sbo=%4d base_offset_mode=%d : %4d/8192 elements, %3d/128 rows match", c.r0, c.sbo, c.bom, ok, row_ok);
        if (first_bad_m >= 0) {
            // describe what row first_bad_m actually contains: find (row, chunk permutation)
            const int m = first_bad_m;
            printf("  | first bad row m=%d got chunks from:", m);
            for (int ch = 0; ch < 8; ++ch) {
                int found_r = -1, found_c = -1;
                for (int r = 0; r < ROWS && found_r < 0; ++r)
                    for (int c2 = 0; c2 < 8; ++c2) {
                        bool eq = true;
                        for (int e2 = 0; e2 < 8; ++e2) eq &= hO[m * 64 + ch * 8 + e2] == val(r, c2 * 8 + e2);
                        if (eq) { found_r = r; found_c = c2; break; }
                    }
                printf(" (%d,%d)", found_r, found_c);
            }
        }
        printf("\n");
    }
    return 0;
}
