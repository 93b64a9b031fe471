// conv_in of the encoder (3 -> 128 channels, 3x3, pad 1) straight from the image batch.
//
// The contraction is K = 27 (padded to 32): far too thin for an operand pipeline fed by TMA, and the
// image is NCHW fp32 (or NHWC u8) -- not a layout a tensor map can turn into K-major operand rows.
// So the A operand is BUILT in shared memory: four gather warps stage the three image rows a tile
// touches as fp16, then every thread writes the 32-value patch of two pixels straight into the
// 128-byte-swizzled operand rows the tensor core reads.  The weights ([128][64] fp16, k >= 27 zero)
// stay resident in shared memory; two K=16 MMAs per 128-pixel sub-tile; the epilogue (bias, GroupNorm
// partial sums of the output, bf16 pack, stores) is the implicit-GEMM one.
//
// What bounds it (ncu source page, round 2): instruction issue in the epilogue -- ~480 warp instructions per 32x32
// output chunk (alpha/bias FFMA, GroupNorm sum / sum-of-squares, 16-bit pack / unpack, predicates and address
// arithmetic) against 27 MACs per output element on the tensor core; neither loading the image one tile ahead nor
// sixteen epilogue warps instead of eight moved it (both tried and reverted).
//
// Tile = 256 consecutive pixels of one image row (two 128-row sub-tiles) x 128 output channels.
// Algorithmic HBM traffic per image: 3*H*W*4 B read + H*W*128*2 B written (the 64-wide patch matrix the
// first version materialised, 128 B per pixel written and read again, is gone).
#pragma once

#include "vt_igemm.cuh"

namespace vt {

struct ConvInCfg {
    using Epi = IgemmCfg<128, 2>;                     // epilogue geometry: 2 x 128 pixels x 128 channels
    static constexpr int GATHER_WARPS = 4;
    static constexpr int THREADS = 64 + 32 * Epi::EPI_WARPS + 32 * GATHER_WARPS;
    static constexpr int A_STAGE = 2 * IGEMM_A_BYTES;  // two sub-tiles of 128 rows x 128 B
    static constexpr int STAGES = 3;
    static constexpr int B_BYTES = 128 * 128;          // 128 output channels x 64 k (fp16)
    static constexpr int STRIP_W = 264;                // 258 columns used: x0-1 .. x0+256
    static constexpr int STRIP_BYTES = 9 * STRIP_W * 2;  // [kh][c][x] fp16
    static constexpr int STRIP_ALLOC = 2 * ((STRIP_BYTES + 127) / 128 * 128);
    static constexpr int SMEM_BYTES = STAGES * A_STAGE + B_BYTES + Epi::EPI_STAGING_BYTES + STRIP_ALLOC + Epi::BAR_BYTES + 1024;
    static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
};

struct ConvInParams {
    const void* img;   // [N][3][H][W] fp32 in [-1,1]  or  [N][H][W][3] u8
    int in_fmt;        // 0 fp32 NCHW, 1 u8 NHWC (normalised (u/255 - 0.5)/0.5 on the fly)
};

__global__ void __launch_bounds__(ConvInCfg::THREADS, 1)
conv_in_kernel(const __grid_constant__ CUtensorMap tmB, const __grid_constant__ IgemmParams P, const ConvInParams Q) {
    using Cfg = ConvInCfg;
    using Epi = ConvInCfg::Epi;
    constexpr int STAGES = Cfg::STAGES;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* s_b = smem + STAGES * Cfg::A_STAGE;
    float* staging_all = reinterpret_cast<float*>(s_b + Cfg::B_BYTES);
    __half* strips = reinterpret_cast<__half*>(s_b + Cfg::B_BYTES + Epi::EPI_STAGING_BYTES);
    uint8_t* ctrl = s_b + Cfg::B_BYTES + Epi::EPI_STAGING_BYTES + Cfg::STRIP_ALLOC;
    uint64_t* a_full = reinterpret_cast<uint64_t*>(ctrl);   // [STAGES]  128 gather threads arrive
    uint64_t* a_empty = a_full + STAGES;                    // [STAGES]  tcgen05.commit
    uint64_t* tfull_bar = a_empty + STAGES;                 // [2]
    uint64_t* tempty_bar = tfull_bar + 2;                   // [2]
    uint64_t* b_full = tempty_bar + 2;                      // [1]
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(b_full + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t total_tiles = static_cast<uint32_t>(P.NB) * static_cast<uint32_t>(P.tiles_x * P.tiles_y);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmB);
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(&a_full[i], 32 * Cfg::GATHER_WARPS);
            mbar_init(&a_empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull_bar[i], 1);
            mbar_init(&tempty_bar[i], Epi::EPI_WARPS);
        }
        mbar_init(b_full, 1);
        fence_mbar_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_ptr, Epi::TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        // ------------------------------------------------------------ weights: one TMA load, resident
        if (lane == 0) {
            mbar_arrive_expect_tx(b_full, Cfg::B_BYTES);
            tma_load_3d(s_b, &tmB, b_full, 0, 0, 0);
        }
        __syncwarp();
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer (warp-uniform loop)
        constexpr uint32_t idesc = umma_idesc_16(IGEMM_BLOCK_M, 128, true);
        const uint64_t da_base = umma_desc_k_sw128(smem_u32(smem));
        const uint64_t db = umma_desc_k_sw128(smem_u32(s_b));
        mbar_wait(b_full, 0);
        int stage = 0;
        uint32_t phase = 0, it = 0;
        for (uint32_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
            const uint32_t acc = it & 1;
            mbar_wait(&tempty_bar[acc], ((it >> 1) & 1) ^ 1);
            mbar_wait(&a_full[stage], phase);
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + acc * Epi::ACC_COLS;
            const uint64_t so = static_cast<uint64_t>(stage * (Cfg::A_STAGE >> 4));
            if (elect_one()) {
#pragma unroll
                for (int k = 0; k < 2; ++k)          // K = 32: two K=16 steps, +32 bytes inside the swizzle row
#pragma unroll
                    for (int t = 0; t < 2; ++t)
                        umma_bf16_ss(tmem_d + t * 128, da_base + so + (t * (IGEMM_A_BYTES >> 4) + 2 * k), db + 2 * k, idesc, k);
                umma_commit(&a_empty[stage]);
                umma_commit(&tfull_bar[acc]);
            }
            __syncwarp();
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
    } else if (warp < 2 + Epi::EPI_WARPS) {
        // ------------------------------------------------------------ epilogue warps
        // output = raw activation: bf16, or fp16 (P.out_fmt)
        if (P.out_fmt == FMT_F16) {
            if (P.group_size != 0)
                igemm_epilogue<Epi, FMT_F16, 0, true>(P, staging_all, ctrl, tfull_bar, tempty_bar, tmem_base, total_tiles, warp, lane);
            else
                igemm_epilogue<Epi, FMT_F16, 0, false>(P, staging_all, ctrl, tfull_bar, tempty_bar, tmem_base, total_tiles, warp, lane);
        } else if (P.group_size != 0)
            igemm_epilogue<Epi, FMT_BF16, 0, true>(P, staging_all, ctrl, tfull_bar, tempty_bar, tmem_base, total_tiles, warp, lane);
        else
            igemm_epilogue<Epi, FMT_BF16, 0, false>(P, staging_all, ctrl, tfull_bar, tempty_bar, tmem_base, total_tiles, warp, lane);
    } else {
        // ------------------------------------------------------------ gather warps: build the operand rows
        const int gt = threadIdx.x - (64 + 32 * Epi::EPI_WARPS);   // 0..127
        constexpr int NG = 32 * Cfg::GATHER_WARPS;
        constexpr int SW = Cfg::STRIP_W;
        const long long plane = static_cast<long long>(P.H) * P.W;
        int stage = 0;
        uint32_t phase = 0, it = 0;
        for (uint32_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
            const int tx = static_cast<int>(tile % static_cast<uint32_t>(P.tiles_x));
            const uint32_t m = tile / static_cast<uint32_t>(P.tiles_x);
            const int y = static_cast<int>(m % static_cast<uint32_t>(P.tiles_y));
            const int img = static_cast<int>(m / static_cast<uint32_t>(P.tiles_y));
            const int x0 = tx * 256;
            __half* strip = strips + (it & 1) * (Cfg::STRIP_ALLOC / 4);
            // ---- image rows y-1 .. y+1, columns x0-1 .. x0+256, three channels -> fp16 strip [kh*3+c][x]
            // (all loads of a tile are issued before the first is used: one L2 round trip per tile, not one per element)
            if (Q.in_fmt == 0) {
                const float* src = static_cast<const float*>(Q.img) + 3LL * img * plane;
                float v[27];
#pragma unroll
                for (int rowid = 0; rowid < 9; ++rowid) {
                    const int kh = rowid / 3, c = rowid - kh * 3;
                    const int gy = y + kh - 1;
                    const bool y_ok = gy >= 0 && gy < P.H;
                    const float* rowp = src + c * plane + static_cast<long long>(gy) * P.W + (x0 - 1);
#pragma unroll
                    for (int part = 0; part < 3; ++part) {
                        const int xx = gt + NG * part, gx = x0 + xx - 1;
                        float t = 0.f;
                        if (y_ok && xx < 258 && gx >= 0 && gx < P.W) t = __ldg(rowp + xx);
                        v[rowid * 3 + part] = t;
                    }
                }
#pragma unroll
                for (int rowid = 0; rowid < 9; ++rowid)
#pragma unroll
                    for (int part = 0; part < 3; ++part) {
                        const int xx = gt + NG * part;
                        if (xx < 258) strip[rowid * SW + xx] = __float2half_rn(v[rowid * 3 + part]);
                    }
            } else {
                const unsigned char* src = static_cast<const unsigned char*>(Q.img) + 3LL * img * plane;
                unsigned char u[21];
#pragma unroll
                for (int kh = 0; kh < 3; ++kh) {
                    const int gy = y + kh - 1;
                    const bool y_ok = gy >= 0 && gy < P.H;
                    const unsigned char* rowp = src + (static_cast<long long>(gy) * P.W + (x0 - 1)) * 3;
#pragma unroll
                    for (int part = 0; part < 7; ++part) {
                        const int b = gt + NG * part, xx = b / 3, gx = x0 + xx - 1;
                        unsigned char t = 0;
                        if (y_ok && b < 774 && gx >= 0 && gx < P.W) t = __ldg(rowp + b);
                        u[kh * 7 + part] = t;
                    }
                }
#pragma unroll
                for (int kh = 0; kh < 3; ++kh) {
                    const int gy = y + kh - 1;
                    const bool y_ok = gy >= 0 && gy < P.H;
#pragma unroll
                    for (int part = 0; part < 7; ++part) {
                        const int b = gt + NG * part, xx = b / 3, c = b - xx * 3, gx = x0 + xx - 1;
                        const bool ok = y_ok && gx >= 0 && gx < P.W;
                        const float val = ok ? (static_cast<float>(u[kh * 7 + part]) / 255.0f - 0.5f) / 0.5f : 0.f;
                        if (b < 774) strip[(kh * 3 + c) * SW + xx] = __float2half_rn(val);
                    }
                }
            }
            asm volatile("bar.sync 3, %0;" ::"n"(NG) : "memory");
            mbar_wait(&a_empty[stage], phase ^ 1);
            // ---- two pixels per thread: k = (kh*3 + kw)*3 + c  ->  strip[kh*3 + c][x + kw]
            const uint32_t sa = smem_u32(smem + stage * Cfg::A_STAGE);
            const unsigned short* sp = reinterpret_cast<const unsigned short*>(strip);
#pragma unroll
            for (int t = 0; t < 2; ++t) {
                const int px = gt + 128 * t;
                uint32_t w[16];
#pragma unroll
                for (int kp = 0; kp < 16; ++kp) {
                    uint32_t lo = 0, hi = 0;
                    const int k = 2 * kp;
                    if (k < 27) lo = sp[((k / 9) * 3 + (k % 3)) * SW + px + (k / 3) % 3];
                    if (k + 1 < 27) hi = sp[(((k + 1) / 9) * 3 + ((k + 1) % 3)) * SW + px + ((k + 1) / 3) % 3];
                    w[kp] = lo | (hi << 16);
                }
                const uint32_t row = sa + t * IGEMM_A_BYTES + gt * 128;
#pragma unroll
                for (int cidx = 0; cidx < 4; ++cidx)
                    sts128(row + ((cidx ^ (gt & 7)) << 4), __uint_as_float(w[4 * cidx]), __uint_as_float(w[4 * cidx + 1]),
                           __uint_as_float(w[4 * cidx + 2]), __uint_as_float(w[4 * cidx + 3]));
            }
            fence_proxy_async_smem();
            mbar_arrive(&a_full[stage]);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, Epi::TMEM_COLS);
    }
}

}  // namespace vt
