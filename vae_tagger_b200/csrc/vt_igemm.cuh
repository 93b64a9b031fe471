// Implicit-GEMM contraction kernel for sm_100a: every dense contraction on the encoder
// path (3x3 convs stride 1 / stride 2, 1x1 shortcut folded in as extra K-slabs, the
// conv_in / conv_out degenerate shapes, attention projections, QK^T, PV) is one launch of
// this kernel with a different "slab table".
//
//   D[m, n] = alpha * sum_{slab s} sum_{k} A_s[m, k] * B[n, kb_s + k]  (+ bias[n]) (+ residual[m, n])
//
//   A: activations, NHWC bf16, described by up to two 5-D TMA tensor maps
//      (c, x, p, y, img); the M-tile is a th x tw patch of output pixels (tw*th = 128) and a
//      slab shifts the patch by (dx, p, dy) -- that is the whole of "im2col": nine shifted
//      TMA boxes with out-of-bounds zero fill, no gather, no pad copy.
//   B: weights (or a second activation), [N][K] K-contiguous bf16, 3-D TMA map (k, n, img).
//   D: fp32 accumulator in TMEM (tcgen05.mma cta_group::1, M=128, N=BLOCK_N, K=16),
//      double buffered so the epilogue of tile i overlaps the MMAs of tile i+1.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer,
// warps 2..5 = epilogue (TMEM -> registers -> bias/residual/GroupNorm partial sums ->
// bf16|fp32 global stores).  Persistent CTAs, static round-robin tile order with the
// n-blocks of one pixel tile adjacent (they share the A tile through L2).
#pragma once
#include "vt_ptx.cuh"

namespace vt {

constexpr int IGEMM_BLOCK_M = 128;
constexpr int IGEMM_BLOCK_K = 64;  // 64 bf16 = one 128-byte swizzle row
constexpr int IGEMM_A_BYTES = IGEMM_BLOCK_M * IGEMM_BLOCK_K * 2;
constexpr int IGEMM_MAX_SLABS = 10;
constexpr int IGEMM_THREADS = 192;

struct IgemmSlab {
    int map;      // 0 / 1: which A tensor map
    int c_base;   // A coordinate 0 of the slab's first 64-channel chunk
    int dx;       // added to the tile's x origin (A coordinate 1)
    int p;        // A coordinate 2 (row/column parity plane for stride-2 maps, else 0)
    int dy;       // added to the tile's y origin (A coordinate 3)
    int kb_base;  // B coordinate 0 of the slab's first chunk
    int nchunks;  // number of 64-wide K chunks in this slab
    int pad_;
};

struct IgemmParams {
    int W, H, NB;          // output pixels per row / rows / images (plain GEMM: W=M, H=1, NB=batch)
    int tw, th;            // M-tile patch, tw*th == 128
    int tiles_x, tiles_y;  // patches per image
    int n_total;           // valid output channels (multiple of 32)
    int n_blocks;          // ceil(n_total / BLOCK_N)
    int num_slabs;
    int a_batched;         // A coordinate 4 = image index (0: shared operand)
    int b_batched;         // B coordinate 2 = image index (0: shared operand)
    int out_fp32;          // output element type: 0 bf16, 1 fp32
    int res_fp32;          // residual element type: 0 bf16, 1 fp32
    int group_size;        // channels per GroupNorm group for the fused statistics; 0 = off
    float alpha;
    const float* bias;              // [n_total] or nullptr
    const void* residual;  // same geometry as out (bf16 or fp32), or nullptr
    void* out;
    long long ld_out;      // elements between consecutive pixels of out / residual
    long long out_bstride;  // elements between consecutive images of out / residual
    double* stats;     // [NB][n_total/group_size][2] running (sum, sum of squares): fp32 per-tile partials,
                       // fp64 atomics across tiles (keeps E[x^2]-E[x]^2 well conditioned)
    IgemmSlab slabs[IGEMM_MAX_SLABS];
};

// Reduce V (power of two <= 32) per-thread values across the warp with V-1 + (5-log2 V)
// shuffles instead of 5*V: after the call lane l holds the total of value (l >> (5 - log2 V)).
template <int V>
__device__ __forceinline__ float warp_multi_reduce(float (&a)[V], int lane) {
    int n = V;
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        if (n > 1) {
            const int half = n / 2;
            const bool upper = (lane & o) != 0;
#pragma unroll
            for (int i = 0; i < V / 2; ++i) {
                if (i < half) {
                    const float send = upper ? a[i] : a[i + half];
                    const float keep = upper ? a[i + half] : a[i];
                    a[i] = keep + __shfl_xor_sync(0xFFFFFFFFu, send, o);
                }
            }
            n = half;
        } else {
            a[0] += __shfl_xor_sync(0xFFFFFFFFu, a[0], o);
        }
    }
    return a[0];
}

template <int G>  // G = channels per group (4, 8, 16): accumulate one 32-column chunk
__device__ __forceinline__ void stats_chunk(const float (&v)[32], bool valid, int lane, float* s_acc) {
    constexpr int NG = 32 / G;
    constexpr int V = 2 * NG;
    float a[V];
#pragma unroll
    for (int k = 0; k < NG; ++k) {
        float s = 0.f, ss = 0.f;
#pragma unroll
        for (int i = 0; i < G; ++i) {
            const float x = v[k * G + i];
            s += x;
            ss = fmaf(x, x, ss);
        }
        a[2 * k] = valid ? s : 0.f;
        a[2 * k + 1] = valid ? ss : 0.f;
    }
    const float tot = warp_multi_reduce<V>(a, lane);
    constexpr int SH = (V == 16) ? 1 : (V == 8) ? 2 : 3;
    if ((lane & ((1 << SH) - 1)) == 0) atomicAdd(&s_acc[lane >> SH], tot);
}

template <int BLOCK_N>
struct IgemmCfg {
    static constexpr int B_BYTES = BLOCK_N * IGEMM_BLOCK_K * 2;
    static constexpr int STAGE_BYTES = IGEMM_A_BYTES + B_BYTES;
    static constexpr int STAGES = (BLOCK_N >= 256) ? 4 : (BLOCK_N >= 128 ? 6 : 8);
    static constexpr int TMEM_COLS = (2 * BLOCK_N <= 32) ? 32 : (2 * BLOCK_N <= 64) ? 64 : (2 * BLOCK_N <= 128) ? 128
                                     : (2 * BLOCK_N <= 256) ? 256 : 512;
    static constexpr int BAR_BYTES = 1024;  // barriers, tmem pointer, stats scratch
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + BAR_BYTES + 1024 /*alignment slack*/;
};

template <int BLOCK_N>
__global__ void __launch_bounds__(IGEMM_THREADS, 1)
igemm_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
             const __grid_constant__ CUtensorMap tmB, const __grid_constant__ IgemmParams P) {
    using Cfg = IgemmCfg<BLOCK_N>;
    constexpr int STAGES = Cfg::STAGES;
    static_assert(2 * BLOCK_N <= 512, "two accumulators must fit TMEM");

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* ctrl = smem + STAGES * Cfg::STAGE_BYTES;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(ctrl);      // [STAGES]
    uint64_t* empty_bar = full_bar + STAGES;                     // [STAGES]
    uint64_t* tfull_bar = empty_bar + STAGES;                    // [2]
    uint64_t* tempty_bar = tfull_bar + 2;                        // [2]
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty_bar + 2);
    float* s_stats = reinterpret_cast<float*>(ctrl + 256);       // [2][128]

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    const int tiles_per_img = P.tiles_x * P.tiles_y;
    const long long total_tiles = 1LL * P.NB * tiles_per_img * P.n_blocks;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA0);
        tma_prefetch_desc(&tmA1);
        tma_prefetch_desc(&tmB);
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull_bar[i], 1);
            mbar_init(&tempty_bar[i], 4);
        }
        fence_mbar_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_ptr, Cfg::TMEM_COLS);
        tmem_relinquish();
    }
    if (warp >= 2) {
        for (int i = threadIdx.x - 64; i < 256; i += 128) s_stats[i] = 0.f;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const int nb = static_cast<int>(tile % P.n_blocks);
                long long m = tile / P.n_blocks;
                const int tx = static_cast<int>(m % P.tiles_x);
                m /= P.tiles_x;
                const int ty = static_cast<int>(m % P.tiles_y);
                const int img = static_cast<int>(m / P.tiles_y);
                const int x0 = tx * P.tw, y0 = ty * P.th, n0 = nb * BLOCK_N;
                for (int s = 0; s < P.num_slabs; ++s) {
                    const IgemmSlab sl = P.slabs[s];
                    const CUtensorMap* mapA = sl.map ? &tmA1 : &tmA0;
                    for (int cc = 0; cc < sl.nchunks; ++cc) {
                        mbar_wait(&empty_bar[stage], phase ^ 1);
                        uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
                        uint8_t* sb = sa + IGEMM_A_BYTES;
                        mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
                        tma_load_5d(sa, mapA, &full_bar[stage], sl.c_base + cc * IGEMM_BLOCK_K, x0 + sl.dx, sl.p,
                                    y0 + sl.dy, P.a_batched ? img : 0);
                        tma_load_3d(sb, &tmB, &full_bar[stage], sl.kb_base + cc * IGEMM_BLOCK_K, n0,
                                    P.b_batched ? img : 0);
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_bf16(IGEMM_BLOCK_M, BLOCK_N);
            int stage = 0;
            uint32_t phase = 0;
            int kblocks = 0;
            for (int s = 0; s < P.num_slabs; ++s) kblocks += P.slabs[s].nchunks;
            uint32_t it = 0;
            for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
                const uint32_t acc = it & 1;
                const uint32_t acc_phase = (it >> 1) & 1;
                mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + acc * BLOCK_N;
                for (int kb = 0; kb < kblocks; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
                    const uint64_t da = umma_desc_k_sw128(sa);
                    const uint64_t db = umma_desc_k_sw128(sa + IGEMM_A_BYTES);
#pragma unroll
                    for (int k = 0; k < IGEMM_BLOCK_K / 16; ++k) {
                        // +32 bytes per K=16 step inside the 128-byte swizzle row (encoded >> 4)
                        umma_bf16_ss(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
                    }
                    umma_commit(&empty_bar[stage]);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit(&tfull_bar[acc]);
            }
        }
        __syncwarp();
    } else {
        // ------------------------------------------------------------ epilogue (warps 2..5)
        const int q = warp & 3;           // TMEM lane quadrant this warp may access
        const int row = q * 32 + lane;    // accumulator row = pixel within the patch
        const int et = threadIdx.x - 64;  // 0..127
        uint32_t it = 0;
        for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
            const int nb = static_cast<int>(tile % P.n_blocks);
            long long m = tile / P.n_blocks;
            const int tx = static_cast<int>(m % P.tiles_x);
            m /= P.tiles_x;
            const int ty = static_cast<int>(m % P.tiles_y);
            const int img = static_cast<int>(m / P.tiles_y);
            const int n0 = nb * BLOCK_N;
            const int x = tx * P.tw + row % P.tw;
            const int y = ty * P.th + row / P.tw;
            const bool valid = (x < P.W) && (y < P.H);
            const long long off = static_cast<long long>(img) * P.out_bstride +
                                  (static_cast<long long>(y) * P.W + x) * P.ld_out + n0;

            const uint32_t acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            float* s_acc = s_stats + acc * 128;
            mbar_wait(&tfull_bar[acc], acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + acc * BLOCK_N + (static_cast<uint32_t>(q * 32) << 16);

#pragma unroll 1
            for (int j = 0; j < BLOCK_N / 32; ++j) {
                uint32_t r[32];
                tmem_ld_32x32(taddr + j * 32, r);
                tmem_ld_wait();
                if (j == BLOCK_N / 32 - 1) {
                    // accumulator fully read: hand the TMEM buffer back to the MMA warp
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tempty_bar[acc]);
                }
                const int nc = n0 + j * 32;
                if (nc >= P.n_total) continue;  // ragged N: whole chunk out of range
                float v[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]) * P.alpha;
                if (P.bias != nullptr) {
                    const float4* bp = reinterpret_cast<const float4*>(P.bias + nc);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float4 b = __ldg(bp + i);
                        v[4 * i] += b.x; v[4 * i + 1] += b.y; v[4 * i + 2] += b.z; v[4 * i + 3] += b.w;
                    }
                }
                if (P.residual != nullptr && valid) {
                    if (P.res_fp32) {
                        const float4* rp = reinterpret_cast<const float4*>(static_cast<const float*>(P.residual) + off + j * 32);
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const float4 u = __ldg(rp + i);
                            v[4 * i] += u.x; v[4 * i + 1] += u.y; v[4 * i + 2] += u.z; v[4 * i + 3] += u.w;
                        }
                    } else {
                        const uint4* rp = reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(P.residual) + off + j * 32);
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const uint4 u = __ldg(rp + i);
                            v[8 * i] += bf16_lo(u.x); v[8 * i + 1] += bf16_hi(u.x);
                            v[8 * i + 2] += bf16_lo(u.y); v[8 * i + 3] += bf16_hi(u.y);
                            v[8 * i + 4] += bf16_lo(u.z); v[8 * i + 5] += bf16_hi(u.z);
                            v[8 * i + 6] += bf16_lo(u.w); v[8 * i + 7] += bf16_hi(u.w);
                        }
                    }
                }
                if (P.group_size != 0) {
                    // channel groups never straddle a 32-column chunk (group_size | 32)
                    if (P.group_size == 4) stats_chunk<4>(v, valid, lane, s_acc + j * 16);
                    else if (P.group_size == 8) stats_chunk<8>(v, valid, lane, s_acc + j * 8);
                    else stats_chunk<16>(v, valid, lane, s_acc + j * 4);
                }
                if (valid) {
                    if (P.out_fp32) {
                        float4* op = reinterpret_cast<float4*>(static_cast<float*>(P.out) + off + j * 32);
#pragma unroll
                        for (int i = 0; i < 8; ++i)
                            op[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
                    } else {
                        uint4* op = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(P.out) + off + j * 32);
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            op[i] = make_uint4(pack_bf16x2(v[8 * i], v[8 * i + 1]), pack_bf16x2(v[8 * i + 2], v[8 * i + 3]),
                                               pack_bf16x2(v[8 * i + 4], v[8 * i + 5]),
                                               pack_bf16x2(v[8 * i + 6], v[8 * i + 7]));
                    }
                }
            }
            if (P.group_size != 0) {
                // all four epilogue warps finished adding into s_acc -> flush to global
                asm volatile("bar.sync 1, 128;" ::: "memory");
                const int nvals = 2 * BLOCK_N / P.group_size;  // (sum, sumsq) per group of this n-block
                if (et < nvals) {
                    const int g_total = P.n_total / P.group_size;
                    const float val = s_acc[et];
                    s_acc[et] = 0.f;
                    const int grp = n0 / P.group_size + (et >> 1);
                    if (grp < g_total)
                        atomicAdd(P.stats + (static_cast<long long>(img) * g_total + grp) * 2 + (et & 1), static_cast<double>(val));
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

}  // namespace vt
