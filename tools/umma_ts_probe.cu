// Hardware probe (development aid): tcgen05.mma with the A operand in TENSOR MEMORY (fp16, two
// elements per 32-bit column, written with tcgen05.st.32x32b by the thread that owns the row) and B
// in shared memory (K-major SWIZZLE_128B).  B is a 64x64 identity, so D[m][n] must equal A[m][n].
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/umma_ts_probe tools/umma_ts_probe.cu
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
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}

// mode 0: A[m][k] = (m % 8) * 64 + k ; mode 1: A[m][k] = m
__device__ __host__ inline float aval(int mode, int m, int k) { return mode ? float(m) : float((m % 8) * 64 + k); }

__global__ void __launch_bounds__(128, 1) probe(float* out, int mode, int a_col0) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sB = smem;                                  // 64 rows x 128 B
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 64 * 128);
    uint32_t* tptr = reinterpret_cast<uint32_t*>(bar + 2);
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { mbar_init(&bar[0], 1); fence_mbar_init(); }
    if (warp == 0) { tmem_alloc(tptr, 256); tmem_relinquish(); }
    // identity B: row n, element k = (n == k); 16-byte chunk c of row n sits at physical chunk c ^ (n & 7)
    for (int i = threadIdx.x; i < 64 * 8; i += 128) {
        const int n = i >> 3, c = i & 7;
        __half v[8];
        for (int e = 0; e < 8; ++e) v[e] = __float2half((8 * c + e) == n ? 1.f : 0.f);
        *reinterpret_cast<uint4*>(sB + n * 128 + ((c ^ (n & 7)) << 4)) = *reinterpret_cast<uint4*>(v);
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tptr;
    const uint32_t tmem_d = tmem;               // columns 0..63: D (fp32)
    const uint32_t tmem_a = tmem + a_col0;      // 32 columns: A 128 x 64 fp16
    {
        const int m = threadIdx.x;
        uint32_t w[32];
        for (int i = 0; i < 32; ++i) {
            const __half2 h = __floats2half2_rn(aval(mode, m, 2 * i), aval(mode, m, 2 * i + 1));   // .x = low half
            w[i] = *reinterpret_cast<const uint32_t*>(&h);
        }
        tmem_st_32x32(tmem_a + (static_cast<uint32_t>(warp * 32) << 16), w);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) {
        tc_fence_after();
        const uint64_t db = umma_desc_k_sw128(smem_u32(sB));
        const uint32_t idesc = umma_idesc_16(128, 64, true);
        for (int k = 0; k < 4; ++k) umma_f16_ts(tmem_d, tmem_a + 8 * k, db + 2 * k, idesc, k != 0);
        umma_commit(&bar[0]);
    }
    mbar_wait(&bar[0], 0);
    tc_fence_after();
    uint32_t r[32];
    for (int j = 0; j < 2; ++j) {
        tmem_ld_32x32(tmem_d + j * 32 + (static_cast<uint32_t>(warp * 32) << 16), r);
        tmem_ld_wait();
        for (int i = 0; i < 32; ++i) out[threadIdx.x * 64 + j * 32 + i] = __uint_as_float(r[i]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 256);
}

int main() {
    float* d;
    cudaMalloc(&d, 128 * 64 * 4);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384);
    std::vector<float> h(128 * 64);
    for (int mode = 0; mode < 2; ++mode)
        for (int col0 : {64, 128}) {
            cudaMemset(d, 0xFF, 128 * 64 * 4);
            probe<<<1, 128, 16384>>>(d, mode, col0);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("mode %d col0 %d: CUDA error %s\n", mode, col0, cudaGetErrorString(e)); return 1; }
            cudaMemcpy(h.data(), d, h.size() * 4, cudaMemcpyDeviceToHost);
            int bad = 0;
            for (int m = 0; m < 128; ++m)
                for (int n = 0; n < 64; ++n)
                    if (h[m * 64 + n] != aval(mode, m, n)) {
                        if (bad < 12) printf("  mode %d col0 %d: D[%d][%d] = %g, expected %g\n", mode, col0, m, n, h[m * 64 + n], aval(mode, m, n));
                        ++bad;
                    }
            printf("mode %d a_col0 %d: %d mismatches of %d\n", mode, col0, bad, 128 * 64);
        }
    return 0;
}
