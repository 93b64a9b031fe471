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
// Warp roles: warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer, warps 2.. = epilogue
// (TMEM -> smem staging -> bias/residual/GroupNorm partial sums -> coalesced bf16|fp32 stores).  Persistent CTAs, static round-robin tile order with the
// n-blocks of one pixel tile adjacent (they share the A tile through L2).
#pragma once
#include "vt_ptx.cuh"

namespace vt {

constexpr int IGEMM_BLOCK_M = 128;
constexpr int IGEMM_BLOCK_K = 64;  // 64 bf16 = one 128-byte swizzle row
constexpr int IGEMM_A_BYTES = IGEMM_BLOCK_M * IGEMM_BLOCK_K * 2;
constexpr int IGEMM_MAX_SLABS = 10;

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
    int tw, th;            // M-tile patch, tw*th == 128, tw a power of two
    int tw_log2;
    int sub_dx, sub_dy;    // MT > 1: sub-tile t of a CTA tile sits at patch offset (t*sub_dx, t*sub_dy)
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

template <int BLOCK_N, int MT>  // MT = 128-row sub-tiles per CTA tile (2: two pixel tiles share every weight chunk)
struct IgemmCfg {
    static constexpr int A_BYTES = MT * IGEMM_A_BYTES;
    static constexpr int B_BYTES = BLOCK_N * IGEMM_BLOCK_K * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    // epilogue warps: 4 (one per TMEM lane quadrant) or 8 (two per quadrant, splitting the columns)
    static constexpr int EPI_WARPS = (BLOCK_N == 128 && MT == 1) ? 8 : 4;
    static constexpr int THREADS = 64 + 32 * EPI_WARPS;
    static constexpr int COLS_PER_WARP = BLOCK_N / (EPI_WARPS / 4);
    static constexpr int PASSES_PER_SUB = COLS_PER_WARP / 32;
    static constexpr int PASSES = MT * PASSES_PER_SUB;
    static constexpr int STAGE_ROW_FLOATS = 36;  // 32 columns + 4 pad: conflict-free 16-byte rows
    static constexpr int EPI_STAGING_BYTES = EPI_WARPS * 32 * STAGE_ROW_FLOATS * 4;
    static constexpr int STAGES = (STAGE_BYTES >= 48 * 1024) ? 4 : (STAGE_BYTES >= 32 * 1024 ? 5 : 8);
    static constexpr int ACC_COLS = MT * BLOCK_N;  // TMEM columns of one accumulator set
    static constexpr int TMEM_COLS = (2 * ACC_COLS <= 32) ? 32 : (2 * ACC_COLS <= 64) ? 64 : (2 * ACC_COLS <= 128) ? 128
                                     : (2 * ACC_COLS <= 256) ? 256 : 512;
    static constexpr int BAR_BYTES = 2560;  // barriers, tmem pointer, fp32 + fp64 stats scratch
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_STAGING_BYTES + BAR_BYTES + 1024 /*alignment slack*/;
    static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
    static_assert(2 * ACC_COLS <= 512, "two accumulator sets must fit TMEM");
};

// Epilogue (v2).  Each epilogue warp owns one TMEM lane quadrant (32 accumulator rows = 32 output
// pixels) and a column range, processed in passes of 32 columns:
//   phase A  tcgen05.ld: thread t = row t, 32 fp32 columns -> per-warp smem staging tile [32][32(+4)]
//   phase B  lanes re-map to (row = 4*it + lane/8, 4 columns = lane%8): staging -> registers, * alpha,
//            + bias, + residual (coalesced global loads issued before phase A), GroupNorm partial sums
//            accumulated per lane over its 8 rows, bf16/fp32 pack, coalesced global stores (8 lanes
//            cover 32 consecutive channels of one pixel, a warp instruction covers 4 pixels).
// Statistics: the 4 columns of a lane lie in one group (group sizes are multiples of 4); after the 8
// rows, two shuffles fold the 4 row-slots and 8 lanes add into the CTA's smem accumulators.
template <int BLOCK_N, int MT>
__global__ void __launch_bounds__(IgemmCfg<BLOCK_N, MT>::THREADS, 1)
igemm_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
             const __grid_constant__ CUtensorMap tmB, const __grid_constant__ IgemmParams P) {
    using Cfg = IgemmCfg<BLOCK_N, MT>;
    constexpr int STAGES = Cfg::STAGES;
    constexpr int EPI_WARPS = Cfg::EPI_WARPS;
    constexpr int RF = Cfg::STAGE_ROW_FLOATS;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    float* staging_all = reinterpret_cast<float*>(smem + STAGES * Cfg::STAGE_BYTES);
    uint8_t* ctrl = smem + STAGES * Cfg::STAGE_BYTES + Cfg::EPI_STAGING_BYTES;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(ctrl);      // [STAGES]
    uint64_t* empty_bar = full_bar + STAGES;                     // [STAGES]
    uint64_t* tfull_bar = empty_bar + STAGES;                    // [2]
    uint64_t* tempty_bar = tfull_bar + 2;                        // [2]
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty_bar + 2);
    float* s_stats = reinterpret_cast<float*>(ctrl + 256);       // [2][128]: (sum, sumsq) per group of the n-block

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
            mbar_init(&tempty_bar[i], EPI_WARPS);
        }
        fence_mbar_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_ptr, Cfg::TMEM_COLS);
        tmem_relinquish();
    }
    if (warp >= 2) {
        for (int i = threadIdx.x - 64; i < 256; i += 32 * EPI_WARPS) s_stats[i] = 0.f;
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
                const int x0 = tx * P.tw * (P.sub_dx ? MT : 1), y0 = ty * P.th * (P.sub_dy ? MT : 1), n0 = nb * BLOCK_N;
                for (int s = 0; s < P.num_slabs; ++s) {
                    const IgemmSlab sl = P.slabs[s];
                    const CUtensorMap* mapA = sl.map ? &tmA1 : &tmA0;
                    for (int cc = 0; cc < sl.nchunks; ++cc) {
                        mbar_wait(&empty_bar[stage], phase ^ 1);
                        uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
                        uint8_t* sb = sa + Cfg::A_BYTES;
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
                const uint32_t tmem_d = tmem_base + acc * Cfg::ACC_COLS;
                for (int kb = 0; kb < kblocks; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
                    const uint64_t db = umma_desc_k_sw128(sa + Cfg::A_BYTES);
#pragma unroll
                    for (int k = 0; k < IGEMM_BLOCK_K / 16; ++k) {
#pragma unroll
                        for (int t = 0; t < MT; ++t) {
                            // +32 bytes per K=16 step inside the 128-byte swizzle row (encoded >> 4)
                            const uint64_t da = umma_desc_k_sw128(sa + t * IGEMM_A_BYTES);
                            umma_bf16_ss(tmem_d + t * BLOCK_N, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
                        }
                    }
                    umma_commit(&empty_bar[stage]);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit(&tfull_bar[acc]);
            }
        }
        __syncwarp();
    } else {
        // ------------------------------------------------------------ epilogue warps
        const int ew = warp - 2;            // 0 .. EPI_WARPS-1
        const int q = warp & 3;             // TMEM lane quadrant this warp may access
        const int cg = ew >> 2;             // column range of this warp
        const int col_base = cg * Cfg::COLS_PER_WARP;
        float* stg = staging_all + ew * 32 * RF;
        const int et = threadIdx.x - 64;
        const int rsub = lane >> 3;         // phase B: row slot 0..3
        const int c4 = (lane & 7) * 4;      // phase B: 4-column offset inside the 32-column pass
        const bool has_res = P.residual != nullptr;

        // residual fetch of one pass (coalesced: 8 lanes x 4 channels = 32 consecutive channels of a pixel)
        auto load_res = [&](float4 (&res)[8], const long long* poff, int nc) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                res[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (poff[i] >= 0 && nc < P.n_total) {
                    if (P.res_fp32) {
                        res[i] = __ldg(reinterpret_cast<const float4*>(static_cast<const float*>(P.residual) + poff[i] +
                                                                       nc + c4));
                    } else {
                        const uint2 u = __ldg(reinterpret_cast<const uint2*>(
                            static_cast<const __nv_bfloat16*>(P.residual) + poff[i] + nc + c4));
                        res[i] = make_float4(bf16_lo(u.x), bf16_hi(u.x), bf16_lo(u.y), bf16_hi(u.y));
                    }
                }
            }
        };
        auto tile_geometry = [&](long long tile, int& nb, int& img, long long (&poff)[8 * MT]) {
            nb = static_cast<int>(tile % P.n_blocks);
            long long m = tile / P.n_blocks;
            const int tx = static_cast<int>(m % P.tiles_x);
            m /= P.tiles_x;
            const int ty = static_cast<int>(m % P.tiles_y);
            img = static_cast<int>(m / P.tiles_y);
            const long long img_off = static_cast<long long>(img) * P.out_bstride;
#pragma unroll
            for (int t = 0; t < MT; ++t) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int row = q * 32 + i * 4 + rsub;
                    const int x = (tx * (P.sub_dx ? MT : 1) + t * P.sub_dx) * P.tw + (row & (P.tw - 1));
                    const int y = (ty * (P.sub_dy ? MT : 1) + t * P.sub_dy) * P.th + (row >> P.tw_log2);
                    poff[t * 8 + i] =
                        (x < P.W && y < P.H) ? img_off + (static_cast<long long>(y) * P.W + x) * P.ld_out : -1;
                }
            }
        };

        // running (sum, sumsq) of this CTA for the current (image, n-block): fp64 in shared memory,
        // flushed to global with fp64 atomics only when the (image, n-block) changes -- a persistent CTA
        // walks ~55 consecutive tiles of one image, so this removes ~98 % of the same-address atomics
        double* s_run = reinterpret_cast<double*>(ctrl + 256 + 1024);  // [128], slot et owned by thread et
        if (et < 128) s_run[et] = 0.0;
        int run_img = -1, run_nb = -1;
        auto flush_stats = [&]() {
            const int nvals = 2 * BLOCK_N / P.group_size;
            if (run_img >= 0 && et < nvals) {
                const int g_total = P.n_total / P.group_size;
                const int grp = run_nb * (BLOCK_N / P.group_size) + (et >> 1);
                if (grp < g_total)
                    atomicAdd(P.stats + (static_cast<long long>(run_img) * g_total + grp) * 2 + (et & 1), s_run[et]);
                s_run[et] = 0.0;
            }
        };

        uint32_t it = 0;
        long long tile = blockIdx.x;
        int nb = 0, img = 0;
        long long poff[8 * MT];
        float4 res0[8];
        if (tile < total_tiles) {
            tile_geometry(tile, nb, img, poff);
            if (has_res) load_res(res0, poff, nb * BLOCK_N + col_base);   // in flight while the MMAs run
        }
        for (; tile < total_tiles; ++it) {
            const int n0 = nb * BLOCK_N;
            const uint32_t acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            float* s_acc = s_stats + acc * 128;
            if (P.group_size != 0 && (img != run_img || nb != run_nb)) {
                flush_stats();
                run_img = img; run_nb = nb;
            }
            mbar_wait(&tfull_bar[acc], acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + acc * Cfg::ACC_COLS + col_base + (static_cast<uint32_t>(q * 32) << 16);

            // geometry of the next tile (needed to prefetch its first residual pass at the end of this one)
            const long long next_tile = tile + gridDim.x;
            int nnb = 0, nimg = 0;
            long long npoff[8 * MT];
            if (next_tile < total_tiles) tile_geometry(next_tile, nnb, nimg, npoff);

#pragma unroll
            for (int sub = 0; sub < MT; ++sub) {                 // 128-row sub-tile (unrolled: static register indexing)
            const long long* pf = poff + sub * 8;
#pragma unroll 1
            for (int pc = 0; pc < Cfg::PASSES_PER_SUB; ++pc) {   // 32-column chunk inside this warp's range
                const bool last_pass = (sub == MT - 1) && (pc == Cfg::PASSES_PER_SUB - 1);
                const int nc = n0 + col_base + pc * 32;          // first global column of this pass
                const bool pass_valid = nc < P.n_total;          // ragged N: whole pass out of range
                // ---- phase A: TMEM -> staging
                {
                    uint32_t r[32];
                    tmem_ld_32x32(taddr + sub * BLOCK_N + pc * 32, r);
                    tmem_ld_wait();
                    if (last_pass) {
                        // this warp's part of the accumulator is read: hand the TMEM buffer back
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&tempty_bar[acc]);
                    }
                    float4* dst = reinterpret_cast<float4*>(stg + lane * RF);
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        dst[i] = make_float4(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1]),
                                             __uint_as_float(r[4 * i + 2]), __uint_as_float(r[4 * i + 3]));
                }
                __syncwarp();
                // ---- phase B: staging -> global
                float4 res[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) res[i] = res0[i];
                // prefetch the residual of the next pass (or of the next tile's first pass)
                if (has_res) {
                    if (pc + 1 < Cfg::PASSES_PER_SUB) load_res(res0, pf, nc + 32);
                    else if (sub + 1 < MT) load_res(res0, poff + (sub + 1 < MT ? sub + 1 : 0) * 8, n0 + col_base);
                    else if (next_tile < total_tiles) load_res(res0, npoff, nnb * BLOCK_N + col_base);
                }
                if (pass_valid) {
                    float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (P.bias != nullptr) b4 = __ldg(reinterpret_cast<const float4*>(P.bias + nc + c4));
                    float ssum = 0.f, ssq = 0.f;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        float4 v = *reinterpret_cast<const float4*>(stg + (i * 4 + rsub) * RF + c4);
                        v.x = fmaf(v.x, P.alpha, b4.x); v.y = fmaf(v.y, P.alpha, b4.y);
                        v.z = fmaf(v.z, P.alpha, b4.z); v.w = fmaf(v.w, P.alpha, b4.w);
                        if (has_res) { v.x += res[i].x; v.y += res[i].y; v.z += res[i].z; v.w += res[i].w; }
                        if (pf[i] >= 0) {
                            ssum += (v.x + v.y) + (v.z + v.w);
                            ssq = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, ssq))));
                            if (P.out_fp32) {
                                *reinterpret_cast<float4*>(static_cast<float*>(P.out) + pf[i] + nc + c4) = v;
                            } else {
                                *reinterpret_cast<uint2*>(static_cast<__nv_bfloat16*>(P.out) + pf[i] + nc + c4) =
                                    make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
                            }
                        }
                    }
                    if (P.group_size != 0) {
                        ssum += __shfl_xor_sync(0xFFFFFFFFu, ssum, 8);
                        ssq += __shfl_xor_sync(0xFFFFFFFFu, ssq, 8);
                        ssum += __shfl_xor_sync(0xFFFFFFFFu, ssum, 16);
                        ssq += __shfl_xor_sync(0xFFFFFFFFu, ssq, 16);
                        if (lane < 8) {
                            const int g = (col_base + pc * 32 + c4) / P.group_size;  // group within the n-block
                            atomicAdd(&s_acc[2 * g], ssum);
                            atomicAdd(&s_acc[2 * g + 1], ssq);
                        }
                    }
                }
                __syncwarp();  // staging is reused by the next pass
            }
            }
            if (P.group_size != 0) {
                // all epilogue warps finished adding into s_acc -> fold into the CTA's running fp64 sums
                asm volatile("bar.sync 1, %0;" ::"n"(32 * EPI_WARPS) : "memory");
                const int nvals = 2 * BLOCK_N / P.group_size;  // (sum, sumsq) per group of this n-block
                if (et < nvals) {
                    s_run[et] += static_cast<double>(s_acc[et]);
                    s_acc[et] = 0.f;
                }
            }
            // advance to the next tile
            tile = next_tile;
            nb = nnb; img = nimg;
#pragma unroll
            for (int i = 0; i < 8 * MT; ++i) poff[i] = npoff[i];
        }
        if (P.group_size != 0) flush_stats();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

}  // namespace vt
