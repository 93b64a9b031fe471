// 3x3 stride-1 convolution with GroupNorm(32)+SiLU fused into the operand path ("halo" kernel).
//
//   out = conv3x3( silu( gn(x) ) ) + bias (+ residual),   GroupNorm statistics of `out` in the epilogue
//
// What differs from igemm_kernel:
//   * the A operand of a CTA tile is loaded ONCE per 64-channel chunk as a halo tile
//     ((8*MT+2) x 18 pixels x 64 channels, one 5-D TMA box, out-of-bounds = zero) and the nine taps
//     are nine tcgen05 descriptors that point into that tile at a row offset (dy*HWID + dx): with the
//     128-byte swizzle being a function of the shared-memory address only (checked on hardware with
//     tools/umma_probe.cu) a K-major operand may start at any 128-byte row and use any 8-row-group
//     pitch (here HWID*128 bytes).  L2->SM traffic for A drops 9x -> 1.27x.
//   * four "transform" warps rewrite each halo tile in place before the MMAs read it:
//     raw bf16 x -> fp32 -> *scale_c + shift_c -> SiLU -> fp16 (scale/shift from the producer's
//     (sum, sumsq) statistics, per image), and pixels outside the image are forced to zero (the
//     reference pads AFTER GroupNorm+SiLU).  The separate GroupNorm pass over HBM disappears.
// Accumulators, epilogue, tile scheduler and weight (B) pipeline are those of igemm_kernel.
#pragma once
#include "vt_igemm.cuh"

namespace vt {

// TR = false: accumulator rows = pixels (MT sub-tiles of 8x16), columns = BLOCK_N output channels.
// TR = true ("transposed", for 128-channel layers): accumulator rows = 128 output channels (weights are
//   the A operand), columns = 256 pixels (an 8x32 patch, the halo view is the B operand).  A 128x128
//   MMA reads 8 KB of operands per 64 cycles -- the full shared-memory read bandwidth -- and measured
//   ~45 % tensor-pipe utilisation next to the TMA writes; 128x256 reads 12 KB per 128 cycles.
template <int BLOCK_N, int MT, bool TR, bool PAIR = false>
struct Conv3Cfg {
    static constexpr int kBlockN = BLOCK_N, kMT = MT;
    static constexpr bool kTR = TR;
    static constexpr int PX_W = 8 * MT;               // CTA tile in pixels
    static constexpr int PX_H = TR ? 32 : 16;
    static constexpr int HWID = PX_W + 2;             // halo width in pixels
    static constexpr int HHGT = PX_H + 2;             // halo height
    static constexpr int HROWS = HWID * HHGT;         // 128-byte rows per halo chunk
    static constexpr int HALO_BYTES = (HROWS * 128 + 1023) / 1024 * 1024;
    static constexpr int NHALO = 3;
    // PAIR: two CTAs (neighbouring pixel tiles, same output channels) run every MMA as one
    // tcgen05.mma.cta_group::2 (M = 256); the weight tile -- the B operand -- is split across the pair, so each
    // CTA streams and reads only half of it (shared-memory data pipe: tensor-core operand reads 12 -> 8 KB per MMA)
    static constexpr bool kPair = PAIR;
    static constexpr int B_ROWS = PAIR ? BLOCK_N / 2 : BLOCK_N;   // weight rows this CTA holds
    static constexpr int B_BYTES = B_ROWS * IGEMM_BLOCK_K * 2;    // weight tile of one (tap, 64-channel chunk)
    // transposed (level 0, 2 channel chunks per tile): the epilogue is on the critical path -> two warps per
    // TMEM lane quadrant, paid for with one weight stage; deep-K layers keep four stages and four warps
    // (measured alternatives, 128->128 @1024^2 x2, no residual / residual: 3 halo + 3 stages + 8 epilogue + 4
    // transform warps 535/605 us; 2 halo + 6 stages 554/607; 4 epilogue warps + 4 stages 507/720; this one 518/608)
    static constexpr int BSTAGES = TR ? 3 : (PAIR ? 8 : 4);
    static constexpr int EPI_WARPS = TR ? 8 : 4;
    static constexpr int XF_WARPS = TR ? 8 : 4;   // level 0 has 2 channel chunks per tile: the transform is on the critical path
    static constexpr int THREADS = 64 + 32 * EPI_WARPS + 32 * XF_WARPS;
    static constexpr int COLS_PER_WARP = (TR ? 256 : BLOCK_N) / (EPI_WARPS / 4);  // TMEM columns each epilogue warp walks
    static constexpr int PASSES_PER_SUB = COLS_PER_WARP / 32;
    static constexpr int PASSES = MT * PASSES_PER_SUB;
    static constexpr int STAGE_ROW_FLOATS = 36;
    static constexpr int EPI_STAGING_BYTES = EPI_WARPS * 32 * STAGE_ROW_FLOATS * 4;
    static constexpr int ACC_COLS = TR ? 256 : MT * BLOCK_N;
    static constexpr int TMEM_COLS = 512;
    static constexpr int PART_FLOATS = TR ? (2 * EPI_WARPS * 16) : (2 * EPI_WARPS * (COLS_PER_WARP / 4) * 2);
    static constexpr int SCSH_BYTES = 2 * 512 * 4;     // scale / shift tables, up to 512 input channels
    static constexpr int BAR_BYTES = 256 + 1024 + PART_FLOATS * 4;
    static constexpr int SMEM_BYTES =
        NHALO * HALO_BYTES + BSTAGES * B_BYTES + EPI_STAGING_BYTES + SCSH_BYTES + BAR_BYTES + 1024;
    static_assert(2 * ACC_COLS <= 512, "two accumulator sets must fit TMEM");
    static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
    static_assert(!TR || (BLOCK_N == 128 && MT == 1), "transposed variant: 128 channels x (8x32) pixels");
    static_assert(!PAIR || (!TR && MT == 1), "pair variant: non-transposed, one sub-tile per CTA");
};

// Epilogue of the transposed variant: TMEM lane = output channel, column = pixel.  Each epilogue warp
// owns 32 channels; per pass of 32 pixels (4 image rows x 8) the chunk goes through a column-swizzled
// staging tile so that 4 lanes hold 32 consecutive channels of one pixel (64 contiguous bytes) and a
// warp store covers 8 pixels.  The lane's 8 channels are fixed for the whole kernel: bias lives in
// registers and the GroupNorm partial sums are folded once per tile.
template <typename Cfg, int OUT, int RES, bool STATS, int RAW = FMT_BF16>
__device__ __forceinline__ void conv3t_epilogue(const IgemmParams& P, float* staging_all, uint8_t* ctrl,
                                                uint64_t* tfull_bar, uint64_t* tempty_bar, uint32_t tmem_base,
                                                uint32_t total_tiles, int warp, int lane) {
    constexpr int EPI_WARPS = Cfg::EPI_WARPS;
    constexpr int RF = Cfg::STAGE_ROW_FLOATS;
    constexpr bool OUT_F32 = (OUT == FMT_F32);
    // 16-bit outputs go through a 16-bit staging tile ([channel][32 pixels] = 64-byte rows, bias already added): half
    // the shared-memory wavefronts of the fp32 tile (this kernel is bound by the shared-memory data pipe).  The value
    // is rounded to the output format once more after the residual add; fp32 outputs keep the fp32 tile.
    constexpr bool STG16 = !OUT_F32;
    typedef typename std::conditional<OUT_F32, float, __nv_bfloat16>::type OutT;
    const int ew = warp - 2;
    const int q = warp & 3;                   // TMEM lane quadrant = channels 32q .. 32q+31 of the n-block
    const uint32_t stg = smem_u32(staging_all + ew * 32 * RF);
    const int et = threadIdx.x - 64;
    const int cgrp = lane & 3;                // phase B: 8-channel group inside the warp's 32 channels
    const uint32_t hsel = (cgrp & 1) ? 0x1032u : 0x3210u;   // STG16: PRMT selector that swaps the 16-bit halves on odd groups
    const int swz_a = 8 * ((lane >> 3) & 3);  // phase A: column swizzle of this thread's staging row
    constexpr int NPASS = Cfg::PASSES_PER_SUB;             // 32-pixel passes per warp
    const int pc0 = (ew >> 2) * NPASS;                     // first pass of this warp (two warps per quadrant)

    float* s_part = reinterpret_cast<float*>(ctrl + 256 + 1024);  // [2][EPI_WARPS][8 slots][2]
    if (STATS) {
        for (int i = et; i < Cfg::PART_FLOATS; i += 32 * EPI_WARPS) s_part[i] = 0.f;
        asm volatile("bar.sync 1, %0;" ::"n"(32 * EPI_WARPS) : "memory");
    }
    int ptile = 0;   // pixel tile index inside the image (row of the statistics partial buffer)
    auto decode = [&](uint32_t tile, int& nb, int& x0, int& y0, int& img) {
        nb = static_cast<int>(tile % static_cast<uint32_t>(P.n_blocks));
        uint32_t m = tile / static_cast<uint32_t>(P.n_blocks);
        const int tx = static_cast<int>(m % static_cast<uint32_t>(P.tiles_x));
        m /= static_cast<uint32_t>(P.tiles_x);
        const int ty = static_cast<int>(m % static_cast<uint32_t>(P.tiles_y));
        img = static_cast<int>(m / static_cast<uint32_t>(P.tiles_y));
        ptile = ty * P.tiles_x + tx;
        x0 = tx * Cfg::PX_W;
        y0 = ty * Cfg::PX_H;
    };
    const int ld = static_cast<int>(P.ld_out);
    uint32_t it = 0;
    for (uint32_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        int nb, x0, y0, img;
        decode(tile, nb, x0, y0, img);
        const int ch0 = nb * 128 + q * 32 + cgrp * 8;    // this lane's first output channel
        const bool ch_ok = ch0 < P.n_total;
        const uint32_t acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        const int tile_row = ptile;   // decode() of the NEXT tile (residual prefetch below) overwrites ptile
        float4 b0 = make_float4(0.f, 0.f, 0.f, 0.f), b1 = b0;
        float bias_a = 0.f;   // STG16: bias of the channel this thread holds in phase A (TMEM lane)
        if (P.bias != nullptr && ch_ok) {
            if constexpr (STG16) {
                bias_a = __ldg(P.bias + nb * 128 + q * 32 + lane);
            } else {
                b0 = __ldg(reinterpret_cast<const float4*>(P.bias + ch0));
                b1 = __ldg(reinterpret_cast<const float4*>(P.bias + ch0 + 4));
            }
        }
        const long long img_off = static_cast<long long>(img) * P.out_bstride + ch0;
        OutT* out_img = static_cast<OutT*>(P.out) + img_off;
        // phase B lane mapping: four consecutive pixel columns (one 16-byte staging read per channel) of one
        // image row: column 4*pq + j = 8*row + x  ->  row = pq >> 1, x = 4 * (pq & 1) + j
        const int pq = lane >> 2;
        const int pxb = x0 + 4 * (pq & 1);
        float s_lo = 0.f, q_lo = 0.f, s_hi = 0.f, q_hi = 0.f;

        if (RES == 1 && tile + gridDim.x < total_tiles) {
            // pull the residual rows of this CTA's NEXT tile into L2 now (one pixel = 256 B per thread): the
            // register loads below then hit L2 instead of paying the HBM round trip inside a pass
            int nb2, x2, y2, img2;
            decode(tile + gridDim.x, nb2, x2, y2, img2);
            const int ppx = x2 + (et & 7), ppy = y2 + (et >> 3);
            if (ppx < P.W && ppy < P.H) {
                const __nv_bfloat16* rp = static_cast<const __nv_bfloat16*>(P.residual) +
                                          static_cast<long long>(img2) * P.out_bstride + nb2 * 128 +
                                          (ppy * P.W + ppx) * ld;
                asm volatile("prefetch.global.L2 [%0];" ::"l"(rp));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(rp + 64));
            }
        }
        mbar_wait(&tfull_bar[acc], acc_phase);
        tc_fence_after();
        const uint32_t taddr = tmem_base + acc * Cfg::ACC_COLS + (static_cast<uint32_t>(q * 32) << 16);
#pragma unroll 1
        for (int pc = pc0; pc < pc0 + NPASS; ++pc) {   // 32 pixels = image rows y0+4pc .. +3
            const int py = y0 + 4 * pc + (pq >> 1);
            const bool y_ok = py < P.H && ch_ok;
            const int row_off = (py * P.W + pxb) * ld;
            // residual prefetch for the 4 pixels of this lane in this pass
            uint4 rlo[4], rhi[4];
            if (RES != 0) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    rlo[j] = make_uint4(0u, 0u, 0u, 0u);
                    rhi[j] = rlo[j];
                    if (y_ok && pxb + j < P.W) {
                        const int off = row_off + j * ld;
                        if (RES == 1) {
                            rlo[j] = __ldg(reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(P.residual) + img_off + off));
                        } else {
                            const float* rp = static_cast<const float*>(P.residual) + img_off + off;
                            rlo[j] = __ldg(reinterpret_cast<const uint4*>(rp));
                            rhi[j] = __ldg(reinterpret_cast<const uint4*>(rp + 4));
                        }
                    }
                }
            }
            // ---- phase A: TMEM (row = channel, 32 pixel columns) -> column-swizzled staging
            {
                uint32_t r[32];
                tmem_ld_32x32(taddr + pc * 32, r);
                tmem_ld_wait();
                if (pc == pc0 + NPASS - 1) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tempty_bar[acc]);
                }
                if constexpr (STG16) {
                    // row = channel (64 bytes), 16-byte chunk k at (k ^ sw): conflict-free for the 8-lane store phases
                    // here and for the 16-lane 8-byte reads of phase B
                    const uint32_t dst = stg + lane * 64;
                    const int sw = ((lane >> 1) & 3) ^ (((lane >> 4) & 1) << 1);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        uint32_t w[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            w[j] = pack16x2<OUT>(__uint_as_float(r[8 * k + 2 * j]) + bias_a,
                                                 __uint_as_float(r[8 * k + 2 * j + 1]) + bias_a);
                        sts128(dst + ((k ^ sw) << 4), __uint_as_float(w[0]), __uint_as_float(w[1]), __uint_as_float(w[2]),
                               __uint_as_float(w[3]));
                    }
                } else {
                    const uint32_t dst = stg + lane * RF * 4;
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        sts128(dst + (((4 * k) ^ swz_a) << 2), __uint_as_float(r[4 * k]), __uint_as_float(r[4 * k + 1]),
                               __uint_as_float(r[4 * k + 2]), __uint_as_float(r[4 * k + 3]));
                }
            }
            __syncwarp();
            // ---- phase B: 4 pixels x 8 channels per lane; one 16-byte staging read per channel (the column
            // swizzle of phase A makes the eight lanes of a quarter-warp hit eight different bank groups)
            float vv[4][8];
            if constexpr (STG16) {
                // 4 pixels x 8 channels per lane: one 8-byte read per channel; lanes with an odd channel group walk
                // their rows in the order e ^ 1 so that a 16-lane phase covers all 32 banks
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const int ee = e ^ (cgrp & 1);
                    const int R = cgrp * 8 + ee;
                    const int sw = ((R >> 1) & 3) ^ (((R >> 4) & 1) << 1);
                    uint32_t w0, w1;
                    asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];"
                                 : "=r"(w0), "=r"(w1)
                                 : "r"(stg + R * 64 + (((pq >> 1) ^ sw) << 4) + ((pq & 1) << 3))
                                 : "memory");
                    // register slot e of an odd-group lane therefore holds channel e ^ 1: the residual and the packed
                    // output words get their 16-bit halves swapped to match (one PRMT each), nothing is re-ordered
                    vv[0][e] = raw16_lo<OUT>(w0); vv[1][e] = raw16_hi<OUT>(w0);
                    vv[2][e] = raw16_lo<OUT>(w1); vv[3][e] = raw16_hi<OUT>(w1);
                }
            } else {
                const uint32_t src = stg + ((cgrp * 8) * RF + ((4 * pq) ^ (8 * cgrp))) * 4;
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const float4 t = lds128(src + e * RF * 4);
                    vv[0][e] = t.x; vv[1][e] = t.y; vv[2][e] = t.z; vv[3][e] = t.w;
                }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float v[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) v[e] = vv[i][e];
                if constexpr (!STG16) {
                    v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
                    v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
                }
                if (RES == 1) {
                    uint4 u = rlo[i];
                    if constexpr (STG16) {
                        u.x = __byte_perm(u.x, 0u, hsel); u.y = __byte_perm(u.y, 0u, hsel);
                        u.z = __byte_perm(u.z, 0u, hsel); u.w = __byte_perm(u.w, 0u, hsel);
                    }
                    v[0] += raw16_lo<RAW>(u.x); v[1] += raw16_hi<RAW>(u.x); v[2] += raw16_lo<RAW>(u.y); v[3] += raw16_hi<RAW>(u.y);
                    v[4] += raw16_lo<RAW>(u.z); v[5] += raw16_hi<RAW>(u.z); v[6] += raw16_lo<RAW>(u.w); v[7] += raw16_hi<RAW>(u.w);
                } else if (RES == 2) {
                    const uint4 a = rlo[i], b = rhi[i];
                    v[0] += __uint_as_float(a.x); v[1] += __uint_as_float(a.y); v[2] += __uint_as_float(a.z); v[3] += __uint_as_float(a.w);
                    v[4] += __uint_as_float(b.x); v[5] += __uint_as_float(b.y); v[6] += __uint_as_float(b.z); v[7] += __uint_as_float(b.w);
                }
                const bool ok = y_ok && pxb + i < P.W;
                if (STATS && ok) {
                    s_lo += (v[0] + v[1]) + (v[2] + v[3]);
                    q_lo = fmaf(v[0], v[0], fmaf(v[1], v[1], fmaf(v[2], v[2], fmaf(v[3], v[3], q_lo))));
                    s_hi += (v[4] + v[5]) + (v[6] + v[7]);
                    q_hi = fmaf(v[4], v[4], fmaf(v[5], v[5], fmaf(v[6], v[6], fmaf(v[7], v[7], q_hi))));
                }
                OutT* o = out_img + row_off + i * ld;
                if (OUT_F32) {
                    if (ok) {
                        reinterpret_cast<float4*>(o)[0] = make_float4(v[0], v[1], v[2], v[3]);
                        reinterpret_cast<float4*>(o)[1] = make_float4(v[4], v[5], v[6], v[7]);
                    }
                } else {
                    uint4 pk = make_uint4(pack16x2<OUT>(v[0], v[1]), pack16x2<OUT>(v[2], v[3]),
                                          pack16x2<OUT>(v[4], v[5]), pack16x2<OUT>(v[6], v[7]));
                    if constexpr (STG16) {
                        pk.x = __byte_perm(pk.x, 0u, hsel); pk.y = __byte_perm(pk.y, 0u, hsel);
                        pk.z = __byte_perm(pk.z, 0u, hsel); pk.w = __byte_perm(pk.w, 0u, hsel);
                    }
                    if (ok) *reinterpret_cast<uint4*>(o) = pk;
                }
            }
            __syncwarp();  // staging is reused by the next pass
        }
        if (STATS) {
            // fold the 8 pixel lanes that share a channel group, then lanes 0..3 own two 4-channel slots each
#pragma unroll
            for (int o = 4; o <= 16; o <<= 1) {
                s_lo += __shfl_xor_sync(0xFFFFFFFFu, s_lo, o);
                q_lo += __shfl_xor_sync(0xFFFFFFFFu, q_lo, o);
                s_hi += __shfl_xor_sync(0xFFFFFFFFu, s_hi, o);
                q_hi += __shfl_xor_sync(0xFFFFFFFFu, q_hi, o);
            }
            float* part = s_part + (acc * EPI_WARPS + ew) * 16;
            if (lane < 4) *reinterpret_cast<float4*>(part + lane * 4) = make_float4(s_lo, q_lo, s_hi, q_hi);
            asm volatile("bar.sync 1, %0;" ::"n"(32 * EPI_WARPS) : "memory");
            const int nvals = 2 * 128 / P.group_size;
            const int g_total = P.n_total / P.group_size;
            if (et < nvals && nb * (128 / P.group_size) + (et >> 1) < g_total) {
                // value et = (group g, sum|sumsq): 4-channel slots g*gs/4 .. ; slot s lives in warp q' = s/8
                const int g = et >> 1, which = et & 1;
                const int spg = P.group_size / 4;
                float tot = 0.f;
                for (int sidx = g * spg; sidx < (g + 1) * spg; ++sidx) {
                    // epilogue warps of quadrant qq: ew = (qq + 2) & 3 and that + 4 (warps 2..5 / 6..9 own quadrants 2,3,0,1)
                    const int qq = sidx >> 3, wsel = (qq + 2) & 3;
#pragma unroll
                    for (int h = 0; h < EPI_WARPS / 4; ++h)
                        tot += s_part[(acc * EPI_WARPS + wsel + 4 * h) * 16 + (sidx & 7) * 2 + which];
                }
                // this tile's own row of the partial buffer: no atomics, gn_finalize_kernel adds the rows in order
                const long long row = static_cast<long long>(img) * P.stats_rows + P.stats_row0 + tile_row;
                P.stats_part[(row * g_total + nb * (128 / P.group_size) + g) * 2 + which] = tot;
            }
        }
    }
}

template <int BLOCK_N, int MT, bool TR, bool PAIR, int RAW = FMT_BF16>   // RAW: storage format of raw activations
__global__ void __launch_bounds__(Conv3Cfg<BLOCK_N, MT, TR, PAIR>::THREADS, 1)
conv3_fused_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                   const __grid_constant__ CUtensorMap tmS, const __grid_constant__ IgemmParams P) {
    using Cfg = Conv3Cfg<BLOCK_N, MT, TR, PAIR>;
    constexpr int NHALO = Cfg::NHALO, BST = Cfg::BSTAGES, HWID = Cfg::HWID, HROWS = Cfg::HROWS;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* s_halo = smem;                                   // [NHALO][HALO_BYTES]
    uint8_t* s_b = smem + NHALO * Cfg::HALO_BYTES;            // [BST][B_BYTES]
    float* staging_all = reinterpret_cast<float*>(s_b + BST * Cfg::B_BYTES);
    float* s_sc = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(staging_all) + Cfg::EPI_STAGING_BYTES);
    float* s_sh = s_sc + 512;
    uint8_t* ctrl = reinterpret_cast<uint8_t*>(s_sc) + Cfg::SCSH_BYTES;
    uint64_t* halo_full = reinterpret_cast<uint64_t*>(ctrl);  // [NHALO] TMA landed (raw)
    uint64_t* halo_ready = halo_full + NHALO;                 // [NHALO] transformed, MMA may read
    uint64_t* halo_free = halo_ready + NHALO;                 // [NHALO] MMAs done, TMA may overwrite
    uint64_t* b_full = halo_free + NHALO;                     // [BST]
    uint64_t* b_empty = b_full + BST;                         // [BST]
    uint64_t* tfull_bar = b_empty + BST;                      // [2]
    uint64_t* tempty_bar = tfull_bar + 2;                     // [2]
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty_bar + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = PAIR ? cluster_ctarank() : 0u;      // pair kernels: rank 0 (leader) issues every MMA
    const uint32_t tiles_per_img = static_cast<uint32_t>(P.tiles_x * P.tiles_y);
    const uint32_t total_tiles = static_cast<uint32_t>(P.NB) * tiles_per_img * static_cast<uint32_t>(P.n_blocks);
    const int nchunks = P.cin_chunks;
    // 1x1 shortcut of a channel-changing block: extra items whose A tile is the raw (bf16, untransformed)
    // block input and whose weight columns follow the nine taps (non-transposed variant only)
    const int sc_chunks = TR ? 0 : P.sc_chunks;
    const int items = nchunks + sc_chunks;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        tma_prefetch_desc(&tmS);
        for (int i = 0; i < NHALO; ++i) {
            mbar_init(&halo_full[i], 1);
            mbar_init(&halo_ready[i], (PAIR ? 2 : 1) * Cfg::XF_WARPS);   // pair: both CTAs' transform warps, on the leader
            mbar_init(&halo_free[i], 1);
        }
        for (int i = 0; i < BST; ++i) {
            mbar_init(&b_full[i], 1);
            mbar_init(&b_empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull_bar[i], 1);
            mbar_init(&tempty_bar[i], (PAIR ? 2 : 1) * Cfg::EPI_WARPS);
        }
        fence_mbar_init();
    }
    if (warp == 1) {
        if constexpr (PAIR) { tmem_alloc_pair(tmem_ptr, Cfg::TMEM_COLS); tmem_relinquish_pair(); }
        else { tmem_alloc(tmem_ptr, Cfg::TMEM_COLS); tmem_relinquish(); }
    }
    tc_fence_before();
    __syncthreads();
    if constexpr (PAIR) cluster_sync_all();   // the peer's barriers exist before any remote arrive / TMA signal
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    // pair kernels number the tiles ((pixel-pair * n_blocks + n-block) * 2 + rank): blockIdx.x + k * gridDim.x
    // (gridDim.x even) keeps the rank, and the two CTAs of a cluster walk the same (pixel-pair, n-block) sequence
    auto decode = [&](uint32_t tile, int& nb, int& x0, int& y0, int& img) {
        const uint32_t tq = PAIR ? (tile >> 1) : tile;
        nb = static_cast<int>(tq % static_cast<uint32_t>(P.n_blocks));
        uint32_t m = tq / static_cast<uint32_t>(P.n_blocks);
        if (PAIR) m = 2u * m + (tile & 1u);
        const int tx = static_cast<int>(m % static_cast<uint32_t>(P.tiles_x));
        m /= static_cast<uint32_t>(P.tiles_x);
        const int ty = static_cast<int>(m % static_cast<uint32_t>(P.tiles_y));
        img = static_cast<int>(m / static_cast<uint32_t>(P.tiles_y));
        x0 = tx * Cfg::PX_W;
        y0 = ty * Cfg::PX_H;
    };

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer
        if (lane == 0) {
            int hb = 0, bs = 0;
            uint32_t hphase = 0, bphase = 0;
            // the halo of item i+1 is requested before the nine weight tiles of item i
            uint32_t h_tile = blockIdx.x;
            int h_chunk = 0;
            auto issue_halo = [&]() {
                if (h_tile >= total_tiles) return;
                int nb, x0, y0, img;
                decode(h_tile, nb, x0, y0, img);
                mbar_wait(&halo_free[hb], hphase ^ 1);
                if (h_chunk < nchunks) {
                    mbar_arrive_expect_tx(&halo_full[hb], HROWS * 128);
                    tma_load_5d(s_halo + hb * Cfg::HALO_BYTES, &tmA, &halo_full[hb], h_chunk * IGEMM_BLOCK_K, x0 - 1, 0,
                                y0 - 1, img);
                } else {   // shortcut operand: the plain 8x16 pixel tile, 128 rows of 128 bytes
                    mbar_arrive_expect_tx(&halo_full[hb], 128 * 128);
                    tma_load_5d(s_halo + hb * Cfg::HALO_BYTES, &tmS, &halo_full[hb], (h_chunk - nchunks) * IGEMM_BLOCK_K,
                                x0, 0, y0, img);
                }
                if (++hb == NHALO) { hb = 0; hphase ^= 1; }
                if (++h_chunk == items) { h_chunk = 0; h_tile += gridDim.x; }
            };
            issue_halo();
            for (uint32_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                int nb, x0, y0, img;
                decode(tile, nb, x0, y0, img);
                const int n0 = nb * BLOCK_N;
                for (int c = 0; c < items; ++c) {
                    issue_halo();
                    const int ntaps = c < nchunks ? 9 : 1;
                    for (int tap = 0; tap < ntaps; ++tap) {
                        mbar_wait(&b_empty[bs], bphase ^ 1);
                        const int kb = c < nchunks ? tap * P.gn_C + c * IGEMM_BLOCK_K
                                                   : 9 * P.gn_C + (c - nchunks) * IGEMM_BLOCK_K;
                        if constexpr (PAIR) {
                            // this CTA's half of the weight rows; both halves are counted on the leader's barrier
                            if (rank == 0) mbar_arrive_expect_tx(&b_full[bs], 2 * Cfg::B_BYTES);
                            tma_load_3d_pair(s_b + bs * Cfg::B_BYTES, &tmB, &b_full[bs], kb,
                                             n0 + static_cast<int>(rank) * Cfg::B_ROWS, 0);
                        } else {
                            mbar_arrive_expect_tx(&b_full[bs], Cfg::B_BYTES);
                            tma_load_3d(s_b + bs * Cfg::B_BYTES, &tmB, &b_full[bs], kb, n0, 0);
                        }
                        if (++bs == BST) { bs = 0; bphase ^= 1; }
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer
        // The whole warp walks the loops (warp-uniform control flow keeps the descriptors in uniform
        // registers); one elected lane issues the MMAs and commits.
        if (rank == 0) {
            // fp16 x fp16; transposed: M = 128 channels, N = 256 pixels; pair: M = 2 x 128 pixels
            constexpr uint32_t idesc = umma_idesc_16(PAIR ? 256 : IGEMM_BLOCK_M, TR ? 256 : BLOCK_N, true);
            auto mma = [&](uint32_t d, uint64_t a, uint64_t b, uint32_t id, uint32_t accum) {
                if constexpr (PAIR) umma_f16_ss_pair(d, a, b, id, accum);
                else umma_bf16_ss(d, a, b, id, accum);
            };
            auto commit = [&](uint64_t* bar) {
                if constexpr (PAIR) umma_commit_pair(bar);
                else umma_commit(bar);
            };
            const uint64_t dw_base = umma_desc_k_sw128(smem_u32(s_b));                       // weight tile, stage 0
            const uint64_t dh_base = umma_desc_k_sw128(smem_u32(s_halo), HWID * 128);        // halo tile, buffer 0
            int hb = 0, bs = 0;
            uint32_t hphase = 0, bphase = 0, it = 0;
            for (uint32_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
                const uint32_t acc = it & 1;
                const uint32_t acc_phase = (it >> 1) & 1;
                mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + acc * Cfg::ACC_COLS;
                uint32_t first = 0;
                for (int c = 0; c < nchunks; ++c) {
                    mbar_wait(&halo_ready[hb], hphase);
                    tc_fence_after();
                    const uint64_t dh = dh_base + static_cast<uint64_t>(hb * (Cfg::HALO_BYTES >> 4));
#pragma unroll
                    for (int tap = 0; tap < 9; ++tap) {
                        constexpr int kDummy = 0; (void)kDummy;
                        const int dy = tap / 3, dx = tap % 3;
                        mbar_wait(&b_full[bs], bphase);
                        tc_fence_after();
                        const uint64_t dw = dw_base + static_cast<uint64_t>(bs * (Cfg::B_BYTES >> 4));
                        if (elect_one()) {
#pragma unroll
                            for (int k = 0; k < IGEMM_BLOCK_K / 16; ++k) {
                                if constexpr (TR) {
                                    // A = weights (128 channels), B = halo view: 32 image rows of 8 pixels, one
                                    // 8-row group per image row, HWID halo pixels apart
                                    mma(tmem_d, dw + 2 * k, dh + ((dy * HWID + dx) * 8 + 2 * k), idesc, first | k);
                                } else {
#pragma unroll
                                    for (int t = 0; t < MT; ++t)
                                        mma(tmem_d + t * BLOCK_N, dh + ((dy * HWID + dx + 8 * t) * 8 + 2 * k), dw + 2 * k, idesc,
                                            first | k);
                                }
                            }
                            commit(&b_empty[bs]);
                            if (tap == 8) commit(&halo_free[hb]);
                        }
                        __syncwarp();
                        first = 1;
                        if (++bs == BST) { bs = 0; bphase ^= 1; }
                    }
                    if (++hb == NHALO) { hb = 0; hphase ^= 1; }
                }
                if constexpr (!TR) {
                    // the shortcut operand is the raw block input: raw format x weight columns of the same format
                    constexpr uint32_t idesc_sc = umma_idesc_16(PAIR ? 256 : IGEMM_BLOCK_M, BLOCK_N, RAW == FMT_F16);
                    const uint64_t dx_base = umma_desc_k_sw128(smem_u32(s_halo));                 // plain tile: 1024 B groups
                    for (int c = 0; c < sc_chunks; ++c) {
                        mbar_wait(&halo_ready[hb], hphase);
                        mbar_wait(&b_full[bs], bphase);
                        tc_fence_after();
                        const uint64_t dx = dx_base + static_cast<uint64_t>(hb * (Cfg::HALO_BYTES >> 4));
                        const uint64_t dw = dw_base + static_cast<uint64_t>(bs * (Cfg::B_BYTES >> 4));
                        if (elect_one()) {
#pragma unroll
                            for (int k = 0; k < IGEMM_BLOCK_K / 16; ++k)
                                mma(tmem_d, dx + 2 * k, dw + 2 * k, idesc_sc, 1u);
                            commit(&b_empty[bs]);
                            commit(&halo_free[hb]);
                        }
                        __syncwarp();
                        if (++bs == BST) { bs = 0; bphase ^= 1; }
                        if (++hb == NHALO) { hb = 0; hphase ^= 1; }
                    }
                }
                if (elect_one()) commit(&tfull_bar[acc]);
                __syncwarp();
            }
        }
    } else if (warp < 2 + Cfg::EPI_WARPS) {
        // ------------------------------------------------------------ epilogue warps (2..5)
        // output: fp32, or the raw 16-bit format of this instantiation; residual: none or raw 16-bit
        const int res = P.residual == nullptr ? 0 : 1;
        const int mode = (P.out_fmt == FMT_F32 ? 1 : 0) | (res << 2) | (P.group_size != 0 ? 16 : 0);
#define VT_EPI_CASE(O, R, S)                                                                                  \
    case ((O) | ((R) << 2) | ((S) << 4)):                                                                     \
        if constexpr (TR)                                                                                     \
            conv3t_epilogue<Cfg, (O) ? FMT_F32 : RAW, (R), (S) != 0, RAW>(P, staging_all, ctrl, tfull_bar,      \
                                                      tempty_bar, tmem_base, total_tiles, warp, lane);        \
        else                                                                                                  \
            igemm_epilogue<Cfg, (O) ? FMT_F32 : RAW, (R), (S) != 0, RAW>(P, staging_all, ctrl, tfull_bar,       \
                                                     tempty_bar, tmem_base, total_tiles, warp, lane);         \
        break;
        switch (mode) {
            VT_EPI_CASE(0, 0, 0) VT_EPI_CASE(1, 0, 0) VT_EPI_CASE(0, 1, 0) VT_EPI_CASE(1, 1, 0)
            VT_EPI_CASE(0, 0, 1) VT_EPI_CASE(1, 0, 1) VT_EPI_CASE(0, 1, 1) VT_EPI_CASE(1, 1, 1)
            default: __trap();  // the host launcher rejects every other combination
        }
#undef VT_EPI_CASE
    } else {
        // ------------------------------------------------------------ transform warps (6..9)
        const int xt = threadIdx.x - (64 + 32 * Cfg::EPI_WARPS);  // 0..127
        const int lc = xt & 7;        // logical 8-channel group of this thread (fixed)
        constexpr int RSTEP = 4 * Cfg::XF_WARPS;  // halo rows covered by the transform warps per step
        const int rbase = xt >> 3;    // first halo row of this thread; then +RSTEP per step
        int hb = 0;
        uint32_t hphase = 0;
        int cur_img = -1;
        const double cnt = static_cast<double>(P.H) * P.W * P.gn_gs;
        for (uint32_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            int nb, x0, y0, img;
            decode(tile, nb, x0, y0, img);
            if (img != cur_img) {
                // per-channel scale / shift of this image from the producer's (sum, sumsq)
                asm volatile("bar.sync 2, %0;" ::"n"(32 * Cfg::XF_WARPS) : "memory");
                for (int ch = xt; ch < P.gn_C; ch += 32 * Cfg::XF_WARPS) {
                    const int g = ch / P.gn_gs;
                    const double su = P.gn_stats[(static_cast<long long>(img) * 32 + g) * 2];
                    const double sq = P.gn_stats[(static_cast<long long>(img) * 32 + g) * 2 + 1];
                    const double mean = su / cnt;
                    double var = sq / cnt - mean * mean;
                    var = var > 0.0 ? var : 0.0;
                    const float rstd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(P.gn_eps)));
                    const float ga = P.gn_gamma[ch], be = P.gn_beta[ch];
                    s_sc[ch] = rstd * ga;
                    s_sh[ch] = be - static_cast<float>(mean) * rstd * ga;
                }
                asm volatile("bar.sync 2, %0;" ::"n"(32 * Cfg::XF_WARPS) : "memory");
                cur_img = img;
            }
            for (int c = 0; c < nchunks; ++c) {
                // SiLU(t) = t*sigmoid(t) = h + h*tanh(h) with h = t/2: the halving is folded into the per-channel
                // scale/shift, h is rounded to fp16 and tanh / fma run on half2 (one MUFU per two elements);
                // tests/emulate_bf16.py: no measurable change of the latent error vs the exp/rcp form
                float sc[8], sh[8];
                {
                    const float half_if_silu = P.gn_silu ? 0.5f : 1.0f;
                    const float4 a0 = *reinterpret_cast<const float4*>(s_sc + c * 64 + lc * 8);
                    const float4 a1 = *reinterpret_cast<const float4*>(s_sc + c * 64 + lc * 8 + 4);
                    const float4 b0 = *reinterpret_cast<const float4*>(s_sh + c * 64 + lc * 8);
                    const float4 b1 = *reinterpret_cast<const float4*>(s_sh + c * 64 + lc * 8 + 4);
                    sc[0] = a0.x; sc[1] = a0.y; sc[2] = a0.z; sc[3] = a0.w; sc[4] = a1.x; sc[5] = a1.y; sc[6] = a1.z; sc[7] = a1.w;
                    sh[0] = b0.x; sh[1] = b0.y; sh[2] = b0.z; sh[3] = b0.w; sh[4] = b1.x; sh[5] = b1.y; sh[6] = b1.z; sh[7] = b1.w;
#pragma unroll
                    for (int e = 0; e < 8; ++e) { sc[e] *= half_if_silu; sh[e] *= half_if_silu; }
                }
                mbar_wait(&halo_full[hb], hphase);
                const uint32_t base = smem_u32(s_halo + hb * Cfg::HALO_BYTES);
                const bool silu = P.gn_silu != 0;
                // four halo rows per step: all loads first, then the math, then the stores (the shared-memory
                // accesses are volatile asm and keep their program order, so interleaving them per row would
                // serialise four dependent chains)
#pragma unroll 1
                for (int row0 = rbase; row0 < HROWS; row0 += 4 * RSTEP) {
                    uint32_t u[4][4];
                    uint32_t addr[4];
                    bool inside[4];
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        const int row = row0 + RSTEP * r;
                        const int hy = row / HWID, hx = row - hy * HWID;
                        const int gy = y0 - 1 + hy, gx = x0 - 1 + hx;
                        inside[r] = gy >= 0 && gy < P.H && gx >= 0 && gx < P.W;
                        // physical 16-byte chunk of this thread's channel group in this row (128B swizzle)
                        addr[r] = base + row * 128 + ((lc ^ (row & 7)) << 4);
                        if (row < HROWS)
                            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                                         : "=r"(u[r][0]), "=r"(u[r][1]), "=r"(u[r][2]), "=r"(u[r][3]) : "r"(addr[r]) : "memory");
                    }
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float lo = fmaf(raw16_lo<RAW>(u[r][j]), sc[2 * j], sh[2 * j]);
                            const float hi = fmaf(raw16_hi<RAW>(u[r][j]), sc[2 * j + 1], sh[2 * j + 1]);
                            uint32_t h2 = pack_f16x2(lo, hi);
                            if (silu) {
                                uint32_t th;
                                asm("tanh.approx.f16x2 %0, %1;" : "=r"(th) : "r"(h2));
                                asm("fma.rn.f16x2 %0, %1, %2, %1;" : "=r"(h2) : "r"(h2), "r"(th));
                            }
                            u[r][j] = inside[r] ? h2 : 0u;   // the reference zero-pads AFTER GroupNorm+SiLU
                        }
                    }
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        if (row0 + RSTEP * r < HROWS)
                            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr[r]), "r"(u[r][0]), "r"(u[r][1]),
                                         "r"(u[r][2]), "r"(u[r][3]) : "memory");
                    }
                }
                // generic-proxy writes -> visible to the tensor core's async proxy, then signal the MMA warp
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                    if constexpr (PAIR) mbar_arrive_cluster_cta(&halo_ready[hb], 0);
                    else mbar_arrive(&halo_ready[hb]);
                }
                if (++hb == NHALO) { hb = 0; hphase ^= 1; }
            }
            for (int c = 0; c < sc_chunks; ++c) {
                // shortcut operand: raw bf16 tile, nothing to rewrite -- hand it straight to the MMA warp
                mbar_wait(&halo_full[hb], hphase);
                __syncwarp();
                if (lane == 0) {
                    if constexpr (PAIR) mbar_arrive_cluster_cta(&halo_ready[hb], 0);
                    else mbar_arrive(&halo_ready[hb]);
                }
                if (++hb == NHALO) { hb = 0; hphase ^= 1; }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if constexpr (PAIR) cluster_sync_all();   // neither CTA leaves while the pair still signals / computes
    if (warp == 1) {
        tc_fence_after();
        if constexpr (PAIR) tmem_dealloc_pair(tmem_base, Cfg::TMEM_COLS);
        else tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

}  // namespace vt
