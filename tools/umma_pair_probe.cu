// Hardware probe (development aid): tcgen05.mma.cta_group::2 on a cluster of two CTAs.
//   M = 256 (128 accumulator rows per CTA), N = 64 with each CTA supplying 32 rows of B, K = 64.
//   pass 0: A from shared memory (SS); pass 1: A from tensor memory (TS), the follower CTA signalling
//   "my A is written" with a remote mbarrier arrive.  B is a 64x64 identity split across the pair, so
//   in both CTAs D[m][n] must equal that CTA's A[m][n].
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/umma_pair_probe tools/umma_pair_probe.cu
#include <cuda_fp16.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "../vae_tagger_b200/csrc/vt_ptx.cuh"
using namespace vt;

__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};\n" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
        "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
        "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}

__device__ __host__ inline float aval(int rank, int pass, int m, int k) { return float(rank * 512 + pass * 1024 * 0 + (m % 8) * 64 + k) + (pass ? 0.f : 0.f); }

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) probe(float* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sA = smem;                                  // 128 rows x 128 B
    uint8_t* sB = smem + 16384;                          // 32 rows x 128 B
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 16384 + 4096);   // [0] SS done, [1] A-in-TMEM ready (leader), [2] TS done
    uint32_t* tptr = reinterpret_cast<uint32_t*>(bar + 4);
    const int warp = threadIdx.x >> 5;
    const uint32_t rank = cluster_ctarank();
    if (threadIdx.x == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 2);
        mbar_init(&bar[2], 1);
        fence_mbar_init();
    }
    if (warp == 0) { tmem_alloc_pair(tptr, 256); tmem_relinquish_pair(); }
    // A (SS pass): row m, element k
    for (int i = threadIdx.x; i < 128 * 8; i += 128) {
        const int m = i >> 3, c = i & 7;
        __half v[8];
        for (int e = 0; e < 8; ++e) v[e] = __float2half(aval(rank, 0, m, 8 * c + e));
        *reinterpret_cast<uint4*>(sA + m * 128 + ((c ^ (m & 7)) << 4)) = *reinterpret_cast<uint4*>(v);
    }
    // B half: local row nl = global n - 32*rank; identity
    for (int i = threadIdx.x; i < 32 * 8; i += 128) {
        const int nl = i >> 3, c = i & 7, n = nl + 32 * rank;
        __half v[8];
        for (int e = 0; e < 8; ++e) v[e] = __float2half((8 * c + e) == n ? 1.f : 0.f);
        *reinterpret_cast<uint4*>(sB + nl * 128 + ((c ^ (nl & 7)) << 4)) = *reinterpret_cast<uint4*>(v);
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem = *tptr;
    const uint32_t tmem_d = tmem, tmem_d2 = tmem + 64, tmem_a = tmem + 128;
    const uint32_t idesc = umma_idesc_16(256, 64, true);
    // ---------------- pass 0: SS
    if (rank == 0 && threadIdx.x == 0) {
        const uint64_t da = umma_desc_k_sw128(smem_u32(sA)), db = umma_desc_k_sw128(smem_u32(sB));
        for (int k = 0; k < 4; ++k) umma_f16_ss_pair(tmem_d, da + 2 * k, db + 2 * k, idesc, k != 0);
        umma_commit_pair(&bar[0]);
    }
    mbar_wait(&bar[0], 0);
    tc_fence_after();
    uint32_t r[32];
    for (int j = 0; j < 2; ++j) {
        tmem_ld_32x32(tmem_d + j * 32 + (static_cast<uint32_t>(warp * 32) << 16), r);
        tmem_ld_wait();
        for (int i = 0; i < 32; ++i) out[(rank * 2 + 0) * 8192 + threadIdx.x * 64 + j * 32 + i] = __uint_as_float(r[i]);
    }
    // ---------------- pass 1: TS (A = 2 * value so a stale read of pass 0 data would show)
    {
        const int m = threadIdx.x;
        uint32_t w[32];
        for (int i = 0; i < 32; ++i) {
            const __half2 h = __floats2half2_rn(aval(rank, 1, m, 2 * i) + 1.f, aval(rank, 1, m, 2 * i + 1) + 1.f);
            w[i] = *reinterpret_cast<const uint32_t*>(&h);
        }
        tmem_st_32x32(tmem_a + (static_cast<uint32_t>(warp * 32) << 16), w);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) mbar_arrive_cluster(&bar[1], 0);    // both CTAs arrive on the LEADER's barrier
    if (rank == 0 && threadIdx.x == 0) {
        mbar_wait(&bar[1], 0);
        tc_fence_after();
        const uint64_t db = umma_desc_k_sw128(smem_u32(sB));
        for (int k = 0; k < 4; ++k) umma_f16_ts_pair(tmem_d2, tmem_a + 8 * k, db + 2 * k, idesc, k != 0);
        umma_commit_pair(&bar[2]);
    }
    mbar_wait(&bar[2], 0);
    tc_fence_after();
    for (int j = 0; j < 2; ++j) {
        tmem_ld_32x32(tmem_d2 + j * 32 + (static_cast<uint32_t>(warp * 32) << 16), r);
        tmem_ld_wait();
        for (int i = 0; i < 32; ++i) out[(rank * 2 + 1) * 8192 + threadIdx.x * 64 + j * 32 + i] = __uint_as_float(r[i]);
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 0) tmem_dealloc_pair(tmem, 256);
}

int main() {
    float* d;
    cudaMalloc(&d, 4 * 8192 * 4);
    cudaMemset(d, 0xFF, 4 * 8192 * 4);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
    probe<<<2, 128, 32768>>>(d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
    std::vector<float> h(4 * 8192);
    cudaMemcpy(h.data(), d, h.size() * 4, cudaMemcpyDeviceToHost);
    for (int rank = 0; rank < 2; ++rank)
        for (int pass = 0; pass < 2; ++pass) {
            int bad = 0;
            for (int m = 0; m < 128; ++m)
                for (int n = 0; n < 64; ++n) {
                    const float want = aval(rank, pass, m, n) + (pass ? 1.f : 0.f);
                    const float got = h[(rank * 2 + pass) * 8192 + m * 64 + n];
                    if (got != want) {
                        if (bad < 8) printf("  rank %d pass %d: D[%d][%d] = %g, expected %g\n", rank, pass, m, n, got, want);
                        ++bad;
                    }
                }
            printf("rank %d %s: %d mismatches of 8192\n", rank, pass ? "TS" : "SS", bad);
        }
    return 0;
}
