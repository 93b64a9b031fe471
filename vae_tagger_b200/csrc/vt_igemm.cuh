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
#include <type_traits>

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
    int f16;      // operand format of this slab (A and B): 1 = fp16, 0 = bf16
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
    int out_fmt;           // output element type: FMT_BF16 / FMT_F32 / FMT_F16
    int res_fp32;          // residual element type: 0 = 16-bit raw format (bf16, or fp16 when raw_f16), 1 = fp32
    int raw_f16;           // raw activations (residual stream, conv outputs) are stored fp16 instead of bf16
    int group_size;        // channels per GroupNorm group for the fused statistics; 0 = off
    int ax1, ay1, ax2, ay2;  // pixel offset of accumulator row r+8 / r+16 relative to row r (patch shape dependent)
    float alpha;
    const float* bias;              // [n_total] or nullptr
    const void* residual;  // same geometry as out (bf16 or fp32), or nullptr
    void* out;
    long long ld_out;      // elements between consecutive pixels of out / residual
    long long out_bstride;  // elements between consecutive images of out / residual
    int out_row_pitch;      // elements between output rows (0: W * ld_out) -- a strided output view, e.g. one
    int out_px_stride;      // parity of a 2x-upsampled image; elements between output pixels (0: ld_out)
    // GroupNorm statistics of the output, two stages, no atomics (results do not depend on the batch composition
    // or on which CTA processed which tile): every CTA tile stores its fp32 (sum, sum of squares) per group to
    // stats_part[img][pixel tile][group][2]; gn_finalize_kernel adds the tiles of an image in index order in fp64
    // (keeps E[x^2]-E[x]^2 well conditioned) into stats[img][group][2]
    float* stats_part;
    int stats_rows, stats_row0;   // rows per image in stats_part (>= pixel tiles per image) / first row of this launch
    double* stats;
    IgemmSlab slabs[IGEMM_MAX_SLABS];
    // ---- fused GroupNorm+SiLU 3x3 convolution only (vt_conv3.cuh)
    const double* gn_stats;  // [NB][32][2] (sum, sumsq) of the INPUT tensor
    const float* gn_gamma;   // [Cin]
    const float* gn_beta;    // [Cin]
    int gn_C, gn_gs;         // input channels, channels per group
    float gn_eps;
    int gn_silu;
    int cin_chunks;          // Cin / 64
    int sc_chunks;           // Cs / 64: 1x1 shortcut slab (0 = none)
    int pair;                // CTA-pair kernel: tile index = ((pixel-pair * n_blocks + n-block) * 2 + cluster rank)
};

// MT = 128-row sub-tiles per CTA tile (2: two pixel tiles share every weight chunk).
// PAIR: two CTAs with neighbouring pixel tiles form a cluster and run every MMA as one tcgen05.mma.cta_group::2
// (M = 256: 128 rows per CTA); the B tile is split across the pair, so each CTA streams and reads only half of it
// (shared-memory data pipe: 12 -> 8 KB per 128x256x16 MMA, 8 -> 6 KB per 128x128x16).
template <int BLOCK_N, int MT, bool PAIR = false>
struct IgemmCfg {
    static constexpr int kBlockN = BLOCK_N, kMT = MT;
    static constexpr bool kPair = PAIR;
    static constexpr int A_BYTES = MT * IGEMM_A_BYTES;
    static constexpr int B_ROWS = PAIR ? BLOCK_N / 2 : BLOCK_N;
    static constexpr int B_BYTES = B_ROWS * IGEMM_BLOCK_K * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    // epilogue warps: 4 (one per TMEM lane quadrant) or 8 (two per quadrant, splitting the columns)
    // 128-column tiles have few K chunks per tile (level 0, conv_in): their epilogue is on the critical
    // path and latency bound, so two warps per scheduler; 256-column tiles keep a fourth operand stage
    static constexpr int EPI_WARPS = (BLOCK_N == 128) ? 8 : 4;
    static constexpr int THREADS = 64 + 32 * EPI_WARPS;
    static constexpr int COLS_PER_WARP = BLOCK_N / (EPI_WARPS / 4);
    static constexpr int PASSES_PER_SUB = COLS_PER_WARP / 32;
    static constexpr int PASSES = MT * PASSES_PER_SUB;
    static constexpr int STAGE_ROW_FLOATS = 36;  // 32 columns + 4 pad: conflict-free 16-byte rows
    static constexpr int EPI_STAGING_BYTES = EPI_WARPS * 32 * STAGE_ROW_FLOATS * 4;
    // the epilogue is latency bound (one dependent chain per warp): two warps per scheduler are worth more
    // than a fourth operand stage, which is what their staging tiles cost
    static constexpr int STAGES_SOLO = (STAGE_BYTES >= 48 * 1024) ? (EPI_WARPS == 8 ? 3 : 4) : (STAGE_BYTES >= 32 * 1024 ? 5 : 8);
    // pair kernels have smaller stages (half a B tile): as many as fit next to the epilogue staging, at most six
    static constexpr int STAGES_FIT = (227 * 1024 - EPI_STAGING_BYTES - 4096 - 1024) / STAGE_BYTES;
    static constexpr int STAGES = PAIR ? (STAGES_FIT < 6 ? STAGES_FIT : 6) : STAGES_SOLO;
    static constexpr int ACC_COLS = MT * BLOCK_N;  // TMEM columns of one accumulator set
    static constexpr int TMEM_COLS = (2 * ACC_COLS <= 32) ? 32 : (2 * ACC_COLS <= 64) ? 64 : (2 * ACC_COLS <= 128) ? 128
                                     : (2 * ACC_COLS <= 256) ? 256 : 512;
    static constexpr int PART_FLOATS = 2 * EPI_WARPS * (COLS_PER_WARP / 4) * 2;  // [acc][warp][4-col slot][sum,sq]
    static constexpr int BAR_BYTES = 256 + 1024 + PART_FLOATS * 4;  // barriers + tmem pointer | fp64 running sums | partials
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_STAGING_BYTES + BAR_BYTES + 1024 /*alignment slack*/;
    static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
    static_assert(2 * ACC_COLS <= 512, "two accumulator sets must fit TMEM");
};

// ---------------------------------------------------------------------------------------------
// Epilogue.  Each epilogue warp owns one TMEM lane quadrant (32 accumulator rows = 32 output pixels)
// and a column range, processed in passes of 32 columns:
//   phase A  tcgen05.ld: thread t = row t, 32 fp32 columns -> per-warp smem staging tile [32][32(+4)]
//   phase B  lanes re-map to (row = 8*i + lane/4, 8 columns = lane%4): staging -> registers, * alpha,
//            + bias, + residual (prefetched one pass ahead; the first pass of a tile is in flight
//            while the tile's MMAs still run), GroupNorm partial sums, bf16/fp32 pack, stores in which
//            4 lanes cover 32 consecutive channels of a pixel (64 / 128 contiguous bytes).
// Everything that varies per launch but not per element (output type, residual type, statistics)
// is a template parameter, addresses are one 64-bit base per tile plus 32-bit offsets, and the shared
// GroupNorm accumulators are per-warp slots (no shared-memory atomics): the first two versions of this
// epilogue spent ~1000 issue slots per 32x32 chunk on branches, 64-bit address arithmetic and
// compare-and-swap loops and were the limiter of every layer with few K chunks per tile.
template <typename Cfg, int OUT, int RES, bool STATS, int RAW = FMT_BF16>   // RAW: format of a 16-bit residual (RES == 1)
__device__ __forceinline__ void igemm_epilogue(const IgemmParams& P, float* staging_all, uint8_t* ctrl,
                                               uint64_t* tfull_bar, uint64_t* tempty_bar, uint32_t tmem_base,
                                               uint32_t total_tiles, int warp, int lane) {
    constexpr int BLOCK_N = Cfg::kBlockN, MT = Cfg::kMT;
    constexpr int EPI_WARPS = Cfg::EPI_WARPS;
    constexpr int RF = Cfg::STAGE_ROW_FLOATS;
    constexpr int SLOTS = Cfg::COLS_PER_WARP / 4;  // 4-column statistic slots per warp
    constexpr bool OUT_F32 = (OUT == FMT_F32);
    // 16-bit outputs go through a 16-bit staging tile (64-byte rows, 16-byte chunks XOR-swizzled by the row pair):
    // half the shared-memory wavefronts of the fp32 tile; the accumulator is rounded to the output format before
    // alpha / bias / residual are applied and once more afterwards.  fp32 outputs keep the fp32 tile.
    constexpr bool STG16 = !OUT_F32;
    typedef typename std::conditional<OUT_F32, float, __nv_bfloat16>::type OutT;  // 2-byte outputs share the pointer type

    const int ew = warp - 2;            // 0 .. EPI_WARPS-1
    const int q = warp & 3;             // TMEM lane quadrant this warp may access
    const int cg = ew >> 2;             // column range of this warp
    const int col_base = cg * Cfg::COLS_PER_WARP;
    const uint32_t stg = smem_u32(staging_all + ew * 32 * RF);  // shared-space byte address
    const int et = threadIdx.x - 64;
    const int rlane = lane >> 2;        // phase B: row slot 0..7
    const int j8 = (lane & 3) * 8;      // phase B: 8-column offset inside the 32-column pass

    float* s_part = reinterpret_cast<float*>(ctrl + 256 + 1024);          // [2][EPI_WARPS][SLOTS][2]
    if (STATS) {
        for (int i = et; i < Cfg::PART_FLOATS; i += 32 * EPI_WARPS) s_part[i] = 0.f;
        asm volatile("bar.sync 1, %0;" ::"n"(32 * EPI_WARPS) : "memory");
    }

    // per-tile geometry: image, n-block, and for each sub-tile the element offset of this lane's first
    // row (i = 0) inside the image plus a 4-bit validity mask of its rows i = 0..3
    const int out_pxs = P.out_px_stride ? P.out_px_stride : static_cast<int>(P.ld_out);
    const int out_rowp = P.out_row_pitch ? P.out_row_pitch : P.W * static_cast<int>(P.ld_out);
    struct Geo {
        int nb, img, ptile;
        int off0[MT];
        unsigned vmask[MT];
    };
    auto geometry = [&](uint32_t tile, Geo& g) {
        const uint32_t rk = P.pair ? (tile & 1u) : 0u;
        const uint32_t tq = P.pair ? (tile >> 1) : tile;
        g.nb = static_cast<int>(tq % static_cast<uint32_t>(P.n_blocks));
        uint32_t m = tq / static_cast<uint32_t>(P.n_blocks);
        if (P.pair) m = 2u * m + rk;
        const int tx = static_cast<int>(m % static_cast<uint32_t>(P.tiles_x));
        m /= static_cast<uint32_t>(P.tiles_x);
        const int ty = static_cast<int>(m % static_cast<uint32_t>(P.tiles_y));
        g.img = static_cast<int>(m / static_cast<uint32_t>(P.tiles_y));
        g.ptile = ty * P.tiles_x + tx;
        const int r0 = q * 32 + rlane;
#pragma unroll
        for (int t = 0; t < MT; ++t) {
            const int xs = (tx * (P.sub_dx ? MT : 1) + t * P.sub_dx) * P.tw;
            const int ys = (ty * (P.sub_dy ? MT : 1) + t * P.sub_dy) * P.th;
            const int px0 = xs + (r0 & (P.tw - 1)), py0 = ys + (r0 >> P.tw_log2);
            g.off0[t] = py0 * out_rowp + px0 * out_pxs;
            unsigned vm = 0;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int px = px0 + (i & 1) * P.ax1 + (i >> 1) * P.ax2;
                const int py = py0 + (i & 1) * P.ay1 + (i >> 1) * P.ay2;
                if (px < P.W && py < P.H) vm |= 1u << i;
            }
            g.vmask[t] = vm;
        }
    };
    const int step1 = P.ay1 * out_rowp + P.ax1 * out_pxs;  // accumulator row r -> r + 8
    const int step2 = P.ay2 * out_rowp + P.ax2 * out_pxs;  // accumulator row r -> r + 16

    // residual registers of one pass: 8 channels x 4 rows per lane
    struct ResRegs {
        uint4 lo[4];
        uint4 hi[4];  // only used for fp32 residuals
    };
    auto load_res = [&](ResRegs& rr, const Geo& g, int t, int nc) {
        if (RES == 0) return;
        const long long base = static_cast<long long>(g.img) * P.out_bstride + nc + j8;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int off = g.off0[t] + (i & 1) * step1 + (i >> 1) * step2;
            const bool ok = ((g.vmask[t] >> i) & 1u) && nc < P.n_total;
            rr.lo[i] = make_uint4(0u, 0u, 0u, 0u);
            rr.hi[i] = make_uint4(0u, 0u, 0u, 0u);
            if (RES == 1) {
                if (ok) rr.lo[i] = __ldg(reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(P.residual) + base + off));
            } else {
                const float* rp = static_cast<const float*>(P.residual) + base + off;
                if (ok) {
                    rr.lo[i] = __ldg(reinterpret_cast<const uint4*>(rp));
                    rr.hi[i] = __ldg(reinterpret_cast<const uint4*>(rp + 4));
                }
            }
        }
    };

    uint32_t it = 0;
    uint32_t tile = blockIdx.x;
    Geo geo, ngeo;
    ResRegs rnext;
    if (tile < total_tiles) {
        geometry(tile, geo);
        load_res(rnext, geo, 0, geo.nb * BLOCK_N + col_base);   // in flight while the MMAs run
    }
    for (; tile < total_tiles; ++it) {
        const int n0 = geo.nb * BLOCK_N;
        const uint32_t acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        float* part = s_part + (acc * EPI_WARPS + ew) * SLOTS * 2;
        const uint32_t next_tile = tile + gridDim.x;
        if (next_tile < total_tiles) {
            geometry(next_tile, ngeo);
            if (RES == 1) {
                // pull the residual rows of the next tile into L2 (the four lanes that share an accumulator row
                // split that row's 128-byte lines): the per-pass register loads then never wait on HBM
                constexpr int LINES = Cfg::COLS_PER_WARP / 64;
                const int nc0 = ngeo.nb * BLOCK_N + col_base + (lane & 3) * 64;
                if ((lane & 3) < LINES && nc0 < P.n_total) {
                    const __nv_bfloat16* rb = static_cast<const __nv_bfloat16*>(P.residual) +
                                              static_cast<long long>(ngeo.img) * P.out_bstride + nc0;
#pragma unroll
                    for (int t = 0; t < MT; ++t)
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            if ((ngeo.vmask[t] >> i) & 1u)
                                asm volatile("prefetch.global.L2 [%0];" ::"l"(rb + ngeo.off0[t] + (i & 1) * step1 + (i >> 1) * step2));
                }
            }
        }
        OutT* out_img = static_cast<OutT*>(P.out) + static_cast<long long>(geo.img) * P.out_bstride;

        mbar_wait(&tfull_bar[acc], acc_phase);
        tc_fence_after();
        const uint32_t taddr = tmem_base + acc * Cfg::ACC_COLS + col_base + (static_cast<uint32_t>(q * 32) << 16);

#pragma unroll
        for (int sub = 0; sub < MT; ++sub) {                  // 128-row sub-tile (unrolled: static register indexing)
#pragma unroll 1
            for (int pc = 0; pc < Cfg::PASSES_PER_SUB; ++pc) {  // 32-column chunk inside this warp's range
                const bool last_pass = (sub == MT - 1) && (pc == Cfg::PASSES_PER_SUB - 1);
                const int nc = n0 + col_base + pc * 32;         // first global column of this pass
                const bool pass_valid = nc < P.n_total;         // ragged N: whole pass out of range
                // ---- phase A: TMEM -> staging
                {
                    uint32_t r[32];
                    tmem_ld_32x32(taddr + sub * BLOCK_N + pc * 32, r);
                    tmem_ld_wait();
                    if (last_pass) {
                        // this warp's part of the accumulator set is read: hand the TMEM buffer back
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) {
                            // pair kernels: the leader CTA issues the MMAs into both CTAs' accumulators
                            if (P.pair) mbar_arrive_cluster_relaxed(&tempty_bar[acc], 0);
                            else mbar_arrive(&tempty_bar[acc]);
                        }
                    }
                    if constexpr (STG16) {
                        const uint32_t dst = stg + lane * 64;
                        const int sw = (lane >> 1) & 3;
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            uint32_t w[4];
#pragma unroll
                            for (int j = 0; j < 4; ++j)
                                w[j] = pack16x2<OUT>(__uint_as_float(r[8 * k + 2 * j]), __uint_as_float(r[8 * k + 2 * j + 1]));
                            sts128(dst + ((k ^ sw) << 4), __uint_as_float(w[0]), __uint_as_float(w[1]), __uint_as_float(w[2]),
                                   __uint_as_float(w[3]));
                        }
                    } else {
                        const uint32_t dst = stg + lane * RF * 4;
#pragma unroll
                        for (int i = 0; i < 8; ++i)
                            sts128(dst + i * 16, __uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1]),
                                   __uint_as_float(r[4 * i + 2]), __uint_as_float(r[4 * i + 3]));
                    }
                }
                __syncwarp();
                // ---- phase B: staging -> global
                ResRegs rcur = rnext;
                if (RES != 0) {   // prefetch the residual of the next pass / next sub-tile / next tile
                    if (pc + 1 < Cfg::PASSES_PER_SUB) load_res(rnext, geo, sub, nc + 32);
                    else if (sub + 1 < MT) load_res(rnext, geo, sub + 1 < MT ? sub + 1 : 0, n0 + col_base);
                    else if (next_tile < total_tiles) load_res(rnext, ngeo, 0, ngeo.nb * BLOCK_N + col_base);
                }
                if (pass_valid) {
                    float4 b0 = make_float4(0.f, 0.f, 0.f, 0.f), b1 = b0;
                    if (P.bias != nullptr) {
                        b0 = __ldg(reinterpret_cast<const float4*>(P.bias + nc + j8));
                        b1 = __ldg(reinterpret_cast<const float4*>(P.bias + nc + j8 + 4));
                    }
                    float s_lo = 0.f, q_lo = 0.f, s_hi = 0.f, q_hi = 0.f;
                    OutT* optr = out_img + nc + j8;
                    // all staging reads of the pass first (volatile accesses keep program order)
                    float4 sv0[4], sv1[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        if constexpr (STG16) {
                            const int R = 8 * i + rlane;
                            const float4 t = lds128(stg + R * 64 + (((lane & 3) ^ ((R >> 1) & 3)) << 4));
                            const uint32_t w0 = __float_as_uint(t.x), w1 = __float_as_uint(t.y), w2 = __float_as_uint(t.z),
                                           w3 = __float_as_uint(t.w);
                            sv0[i] = make_float4(raw16_lo<OUT>(w0), raw16_hi<OUT>(w0), raw16_lo<OUT>(w1), raw16_hi<OUT>(w1));
                            sv1[i] = make_float4(raw16_lo<OUT>(w2), raw16_hi<OUT>(w2), raw16_lo<OUT>(w3), raw16_hi<OUT>(w3));
                        } else {
                            const uint32_t sp = stg + ((8 * i + rlane) * RF + j8) * 4;
                            sv0[i] = lds128(sp);
                            sv1[i] = lds128(sp + 16);
                        }
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        float4 v0 = sv0[i];
                        float4 v1 = sv1[i];
                        v0.x = fmaf(v0.x, P.alpha, b0.x); v0.y = fmaf(v0.y, P.alpha, b0.y);
                        v0.z = fmaf(v0.z, P.alpha, b0.z); v0.w = fmaf(v0.w, P.alpha, b0.w);
                        v1.x = fmaf(v1.x, P.alpha, b1.x); v1.y = fmaf(v1.y, P.alpha, b1.y);
                        v1.z = fmaf(v1.z, P.alpha, b1.z); v1.w = fmaf(v1.w, P.alpha, b1.w);
                        if (RES == 1) {
                            const uint4 u = rcur.lo[i];
                            v0.x += raw16_lo<RAW>(u.x); v0.y += raw16_hi<RAW>(u.x); v0.z += raw16_lo<RAW>(u.y); v0.w += raw16_hi<RAW>(u.y);
                            v1.x += raw16_lo<RAW>(u.z); v1.y += raw16_hi<RAW>(u.z); v1.z += raw16_lo<RAW>(u.w); v1.w += raw16_hi<RAW>(u.w);
                        } else if (RES == 2) {
                            const uint4 a = rcur.lo[i], b = rcur.hi[i];
                            v0.x += __uint_as_float(a.x); v0.y += __uint_as_float(a.y);
                            v0.z += __uint_as_float(a.z); v0.w += __uint_as_float(a.w);
                            v1.x += __uint_as_float(b.x); v1.y += __uint_as_float(b.y);
                            v1.z += __uint_as_float(b.z); v1.w += __uint_as_float(b.w);
                        }
                        const bool ok = (geo.vmask[sub] >> i) & 1u;
                        if (STATS && ok) {
                            s_lo += (v0.x + v0.y) + (v0.z + v0.w);
                            q_lo = fmaf(v0.x, v0.x, fmaf(v0.y, v0.y, fmaf(v0.z, v0.z, fmaf(v0.w, v0.w, q_lo))));
                            s_hi += (v1.x + v1.y) + (v1.z + v1.w);
                            q_hi = fmaf(v1.x, v1.x, fmaf(v1.y, v1.y, fmaf(v1.z, v1.z, fmaf(v1.w, v1.w, q_hi))));
                        }
                        OutT* o = optr + (geo.off0[sub] + (i & 1) * step1 + (i >> 1) * step2);
                        if (OUT_F32) {
                            if (ok) {
                                reinterpret_cast<float4*>(o)[0] = v0;
                                reinterpret_cast<float4*>(o)[1] = v1;
                            }
                        } else {
                            const uint4 pk = make_uint4(pack16x2<OUT>(v0.x, v0.y), pack16x2<OUT>(v0.z, v0.w),
                                                        pack16x2<OUT>(v1.x, v1.y), pack16x2<OUT>(v1.z, v1.w));
                            if (ok) *reinterpret_cast<uint4*>(o) = pk;
                        }
                    }
                    if (STATS) {
                        // fold the 8 row slots (lanes with equal lane%4), then lanes 0..3 own two 4-column slots each
#pragma unroll
                        for (int o = 4; o <= 16; o <<= 1) {
                            s_lo += __shfl_xor_sync(0xFFFFFFFFu, s_lo, o);
                            q_lo += __shfl_xor_sync(0xFFFFFFFFu, q_lo, o);
                            s_hi += __shfl_xor_sync(0xFFFFFFFFu, s_hi, o);
                            q_hi += __shfl_xor_sync(0xFFFFFFFFu, q_hi, o);
                        }
                        if (lane < 4) {
                            float4* pp = reinterpret_cast<float4*>(part + (pc * 8 + lane * 2) * 2);
                            float4 cur = *pp;
                            cur.x += s_lo; cur.y += q_lo; cur.z += s_hi; cur.w += q_hi;
                            *pp = cur;
                        }
                    }
                }
                __syncwarp();  // staging is reused by the next pass
            }
        }
        if (STATS) {
            // all epilogue warps have written their slots -> this tile's (sum, sumsq) per group, added in a fixed
            // order, stored to the tile's own row of the partial buffer
            asm volatile("bar.sync 1, %0;" ::"n"(32 * EPI_WARPS) : "memory");
            const int nvals = 2 * BLOCK_N / P.group_size;  // (sum, sumsq) per group of this n-block
            const int g_total = P.n_total / P.group_size;
            if (et < nvals && geo.nb * (BLOCK_N / P.group_size) + (et >> 1) < g_total) {
                const int g = et >> 1, which = et & 1;
                const int slots_per_group = P.group_size / 4;
                float tot = 0.f;
                for (int sidx = g * slots_per_group; sidx < (g + 1) * slots_per_group; ++sidx) {
                    const int cgi = sidx / SLOTS, ls = sidx - cgi * SLOTS;
#pragma unroll
                    for (int qq = 0; qq < 4; ++qq) {
                        float* pv = s_part + ((acc * EPI_WARPS + cgi * 4 + qq) * SLOTS + ls) * 2 + which;
                        tot += *pv;
                        *pv = 0.f;
                    }
                }
                const long long row = static_cast<long long>(geo.img) * P.stats_rows + P.stats_row0 + geo.ptile;
                P.stats_part[(row * g_total + geo.nb * (BLOCK_N / P.group_size) + g) * 2 + which] = tot;
            }
        }
        tile = next_tile;
        geo = ngeo;
    }
}

template <int BLOCK_N, int MT, bool PAIR = false>
__global__ void __launch_bounds__(IgemmCfg<BLOCK_N, MT, PAIR>::THREADS, 1)
igemm_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
             const __grid_constant__ CUtensorMap tmB, const __grid_constant__ IgemmParams P) {
    using Cfg = IgemmCfg<BLOCK_N, MT, PAIR>;
    constexpr int STAGES = Cfg::STAGES;
    constexpr int EPI_WARPS = Cfg::EPI_WARPS;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    float* staging_all = reinterpret_cast<float*>(smem + STAGES * Cfg::STAGE_BYTES);
    uint8_t* ctrl = smem + STAGES * Cfg::STAGE_BYTES + Cfg::EPI_STAGING_BYTES;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(ctrl);      // [STAGES]
    uint64_t* empty_bar = full_bar + STAGES;                     // [STAGES]
    uint64_t* tfull_bar = empty_bar + STAGES;                    // [2]
    uint64_t* tempty_bar = tfull_bar + 2;                        // [2]
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty_bar + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = PAIR ? cluster_ctarank() : 0u;     // pair kernels: rank 0 (leader) issues every MMA

    const uint32_t tiles_per_img = static_cast<uint32_t>(P.tiles_x * P.tiles_y);
    const uint32_t total_tiles = static_cast<uint32_t>(P.NB) * tiles_per_img * static_cast<uint32_t>(P.n_blocks);

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
            mbar_init(&tempty_bar[i], (PAIR ? 2 : 1) * EPI_WARPS);   // pair: both CTAs' epilogue warps, on the leader
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

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (uint32_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                // pair kernels number the tiles ((pixel-pair * n_blocks + n-block) * 2 + rank): blockIdx.x + k * gridDim.x
                // (gridDim.x even) keeps the rank, and the two CTAs of a cluster walk the same (pixel-pair, n-block) sequence
                const uint32_t tq = PAIR ? (tile >> 1) : tile;
                const int nb = static_cast<int>(tq % static_cast<uint32_t>(P.n_blocks));
                uint32_t m = tq / static_cast<uint32_t>(P.n_blocks);
                if (PAIR) m = 2u * m + (tile & 1u);
                const int tx = static_cast<int>(m % static_cast<uint32_t>(P.tiles_x));
                m /= static_cast<uint32_t>(P.tiles_x);
                const int ty = static_cast<int>(m % static_cast<uint32_t>(P.tiles_y));
                const int img = static_cast<int>(m / static_cast<uint32_t>(P.tiles_y));
                const int x0 = tx * P.tw * (P.sub_dx ? MT : 1), y0 = ty * P.th * (P.sub_dy ? MT : 1), n0 = nb * BLOCK_N;
                for (int s = 0; s < P.num_slabs; ++s) {
                    const IgemmSlab sl = P.slabs[s];
                    const CUtensorMap* mapA = sl.map ? &tmA1 : &tmA0;
                    for (int cc = 0; cc < sl.nchunks; ++cc) {
                        mbar_wait(&empty_bar[stage], phase ^ 1);
                        uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
                        uint8_t* sb = sa + Cfg::A_BYTES;
                        if constexpr (PAIR) {
                            // both CTAs' A tiles and B halves are counted on the leader's barrier
                            if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * Cfg::STAGE_BYTES);
                            tma_load_5d_pair(sa, mapA, &full_bar[stage], sl.c_base + cc * IGEMM_BLOCK_K, x0 + sl.dx, sl.p,
                                             y0 + sl.dy, P.a_batched ? img : 0);
                            tma_load_3d_pair(sb, &tmB, &full_bar[stage], sl.kb_base + cc * IGEMM_BLOCK_K,
                                             n0 + static_cast<int>(rank) * Cfg::B_ROWS, P.b_batched ? img : 0);
                        } else {
                            mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
                            tma_load_5d(sa, mapA, &full_bar[stage], sl.c_base + cc * IGEMM_BLOCK_K, x0 + sl.dx, sl.p,
                                        y0 + sl.dy, P.a_batched ? img : 0);
                            tma_load_3d(sb, &tmB, &full_bar[stage], sl.kb_base + cc * IGEMM_BLOCK_K, n0,
                                        P.b_batched ? img : 0);
                        }
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
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
            constexpr uint32_t idesc_bf16 = umma_idesc_16(PAIR ? 256 : IGEMM_BLOCK_M, BLOCK_N, false);
            constexpr uint32_t idesc_f16 = umma_idesc_16(PAIR ? 256 : IGEMM_BLOCK_M, BLOCK_N, true);
            auto mma = [&](uint32_t d, uint64_t a, uint64_t b, uint32_t id, uint32_t accum) {
                if constexpr (PAIR) umma_f16_ss_pair(d, a, b, id, accum);
                else umma_bf16_ss(d, a, b, id, accum);
            };
            auto commit = [&](uint64_t* bar) {
                if constexpr (PAIR) umma_commit_pair(bar);
                else umma_commit(bar);
            };
            const uint64_t da_base = umma_desc_k_sw128(smem_u32(smem));                  // stage 0, sub-tile 0
            const uint64_t db_base = umma_desc_k_sw128(smem_u32(smem) + Cfg::A_BYTES);   // stage 0
            int stage = 0;
            uint32_t phase = 0;
            uint32_t it = 0;
            for (uint32_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
                const uint32_t acc = it & 1;
                const uint32_t acc_phase = (it >> 1) & 1;
                mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + acc * Cfg::ACC_COLS;
                uint32_t first = 0;  // 0 until the first MMA of the tile has been issued
                for (int sl = 0; sl < P.num_slabs; ++sl) {
                    const uint32_t idesc = P.slabs[sl].f16 ? idesc_f16 : idesc_bf16;
                    const int nch = P.slabs[sl].nchunks;
                    for (int kb = 0; kb < nch; ++kb) {
                        mbar_wait(&full_bar[stage], phase);
                        tc_fence_after();
                        const uint64_t so = static_cast<uint64_t>(stage * (Cfg::STAGE_BYTES >> 4));
                        if (elect_one()) {
#pragma unroll
                            for (int k = 0; k < IGEMM_BLOCK_K / 16; ++k) {
#pragma unroll
                                for (int t = 0; t < MT; ++t) {
                                    // +32 bytes per K=16 step inside the 128-byte swizzle row (encoded >> 4)
                                    mma(tmem_d + t * BLOCK_N, da_base + so + (t * (IGEMM_A_BYTES >> 4) + 2 * k),
                                        db_base + so + 2 * k, idesc, first | k);
                                }
                            }
                            commit(&empty_bar[stage]);
                        }
                        __syncwarp();
                        first = 1;
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    }
                }
                if (elect_one()) commit(&tfull_bar[acc]);
                __syncwarp();
            }
        }
    } else {
        // ------------------------------------------------------------ epilogue warps
        const int res = P.residual == nullptr ? 0 : (P.res_fp32 ? 2 : 1);
        const int mode = P.out_fmt | (res << 2) | (P.group_size != 0 ? 16 : 0) | (P.raw_f16 && res == 1 ? 32 : 0);
#define VT_EPI_CASE(O, R, S, RAW)                                                                             \
    case ((O) | ((R) << 2) | ((S) << 4) | ((RAW) == FMT_F16 && (R) == 1 ? 32 : 0)):                           \
        igemm_epilogue<Cfg, (O), (R), (S) != 0, (RAW)>(P, staging_all, ctrl, tfull_bar, tempty_bar, tmem_base,  \
                                                         total_tiles, warp, lane);                            \
        break;
        switch (mode) {
            // no residual (RAW irrelevant): bf16 / fp32 / fp16 outputs, with and without statistics
            VT_EPI_CASE(0, 0, 0, FMT_BF16) VT_EPI_CASE(1, 0, 0, FMT_BF16) VT_EPI_CASE(2, 0, 0, FMT_BF16)
            VT_EPI_CASE(0, 0, 1, FMT_BF16) VT_EPI_CASE(1, 0, 1, FMT_BF16) VT_EPI_CASE(2, 0, 1, FMT_BF16)
            // fp32 residual
            VT_EPI_CASE(0, 2, 0, FMT_BF16) VT_EPI_CASE(1, 2, 0, FMT_BF16) VT_EPI_CASE(0, 2, 1, FMT_BF16) VT_EPI_CASE(1, 2, 1, FMT_BF16)
            // bf16 raw residual
            VT_EPI_CASE(0, 1, 0, FMT_BF16) VT_EPI_CASE(1, 1, 0, FMT_BF16) VT_EPI_CASE(0, 1, 1, FMT_BF16) VT_EPI_CASE(1, 1, 1, FMT_BF16)
            // fp16 raw residual
            VT_EPI_CASE(2, 1, 0, FMT_F16) VT_EPI_CASE(1, 1, 0, FMT_F16) VT_EPI_CASE(2, 1, 1, FMT_F16) VT_EPI_CASE(1, 1, 1, FMT_F16)
            default: __trap();  // the host launcher rejects every other combination
        }
#undef VT_EPI_CASE
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
