// Weight gradient of a 3x3 / 1x1 conv (stride 1, or the stride-2 Downsample2D) as a tcgen05 GEMM whose K dimension is
// the PIXELS, with both operands read straight from the NHWC tensors:
//
//   dW[co][tap][ci] = sum_{n,y,x} dY[n,y,x,co] * A[n, s*y + ky - pad, s*x + kx - pad, ci]
//
// NHWC is channel-contiguous, i.e. MN-major for this GEMM (M = co, N = ci, K = pixel).  A TMA box of 64 channels x a
// 16x4 pixel patch lands in shared memory as 64 rows (pixels) of 128 bytes (channels), 128-byte swizzled -- exactly
// the canonical MN-major SWIZZLE_128B operand tile of tcgen05 (checked on hardware with tools/umma_mn_probe.cu:
// leading-dimension byte offset = distance between 64-channel blocks, stride byte offset = 1024 = eight pixel rows,
// a K step of 16 pixels = +2048 bytes, instruction-descriptor bits 15 / 16 = A / B MN-major).  A tap is a shifted box
// of the SAME activation tensor (TMA out-of-bounds zero fill = the conv's padding; the channel coordinate stays a
// multiple of 64, so nothing is misaligned), a stride-2 tap is a box of the (2C, W/2, 2, H/2, N) parity view the
// forward conv uses.  No re-laid-out copies of the operands exist.
//
// One CTA owns 128 output channels x eight 64-wide column boxes (a column box = (tap, 64 input channels); 8 x 64 =
// all 512 TMEM columns, one fp32 accumulator) and a contiguous range of pixel patches (split-K); per 64-pixel stage
// it loads the dY tile once and eight shifted activation boxes, and issues 4 K steps x 2 MMAs of 128 x 256 x 16.
// Partial tiles go to part[split][co][tap*Cin + ci]; wgrad_reduce_kernel (vt_backward.cu) adds them in index order.
//
// 3x3 stride-1 convs (all but 5 layers of the encoder) use wgrad_halo_kernel instead: the eight shifted boxes above
// re-read the same activations eight times from L2 (80 KB of TMA fill per 1024 tensor-pipe cycles per SM -- the fill,
// not the MMA, bounds that kernel).  There a stage is an 8 x 8 pixel patch of dY and ONE 10 x 10 halo tile of 64 input
// channels; the taps are descriptor views of the halo tile: an 8-pixel image row is one group of eight K rows, so the
// stride byte offset is the halo row pitch (10 x 128 B) instead of 1024, a tap moves the start address by
// (ky*10 + kx) x 128 B, and the three kx taps of one kernel row are the three 64-column blocks of ONE N = 192 MMA
// whose leading byte offset is 128 B (block j = the same rows one pixel further).  Shared-memory addresses are swizzled
// by their absolute address bits, so the shifted views read exactly what TMA wrote (the forward fused conv relies on
// the same property).  A CTA owns two (input-channel chunk, kernel row) items = 2 x 192 accumulator columns.
#include "vt_backward.h"
#include "vt_ptx.cuh"

namespace vt {

namespace {

constexpr int WG_BOX = 64 * 128;                    // 64 pixel rows x 128 bytes (64 channels)
constexpr int WG_YB = 2 * WG_BOX;                   // dY tile: two 64-channel blocks
constexpr int WG_AB = 8 * WG_BOX;                   // activation tile: eight column boxes
constexpr int WG_STAGE = WG_YB + WG_AB;             // 80 KB
constexpr int WG_STAGES = 2;
constexpr int WG_THREADS = 192;                     // warp 0 TMA, warp 1 MMA, warps 2..5 epilogue
constexpr int WG_SMEM = WG_STAGES * WG_STAGE + 256 + 1024;

struct WgradMnParams {
    int N, H, W;             // geometry of dY (the conv output)
    int Cout, Cin, ksize, stride;
    int tiles_x, tiles_y;    // 16 x 4 pixel patches per image
    int col_groups, splits, per_split;
    int y_f16, a_f16;        // operand formats: 1 = fp16, 0 = bf16
    float* part;             // [splits][Cout][taps * Cin]
};

__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);
    d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;   // between 64-element blocks of the M / N dimension
    d |= static_cast<uint64_t>(1024 >> 4) << 32;                     // between groups of eight K rows
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;                             // SWIZZLE_128B
    return d;
}

__global__ void __launch_bounds__(WG_THREADS, 1)
wgrad_mn_kernel(const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmA,
                const __grid_constant__ WgradMnParams P) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + WG_STAGES * WG_STAGE);
    uint64_t* empty = full + WG_STAGES;
    uint64_t* done = empty + WG_STAGES;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(done + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    const int s = static_cast<int>(blockIdx.x % static_cast<unsigned>(P.splits));
    const int r = static_cast<int>(blockIdx.x / static_cast<unsigned>(P.splits));
    const int cg = r % P.col_groups, mb = r / P.col_groups;
    const int total = P.N * P.tiles_y * P.tiles_x;
    const int p_begin = s * P.per_split, p_end = min(total, p_begin + P.per_split);
    const int taps = P.ksize * P.ksize, cib = P.Cin / 64, nbox = taps * cib;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmY);
        tma_prefetch_desc(&tmA);
        for (int i = 0; i < WG_STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        mbar_init(done, 1);
        fence_mbar_init();
    }
    if (warp == 1) { tmem_alloc(tmem_ptr, 512); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_ptr;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int p = p_begin; p < p_end; ++p) {
                const int tx = p % P.tiles_x;
                const int q = p / P.tiles_x;
                const int ty = q % P.tiles_y, img = q / P.tiles_y;
                const int x0 = tx * 16, y0 = ty * 4;
                mbar_wait(&empty[stage], phase ^ 1);
                uint8_t* st = smem + stage * WG_STAGE;
                mbar_arrive_expect_tx(&full[stage], WG_STAGE);
                for (int i = 0; i < 2; ++i) tma_load_5d(st + i * WG_BOX, &tmY, &full[stage], mb * 128 + i * 64, x0, 0, y0, img);
                for (int j = 0; j < 8; ++j) {
                    const int b = cg * 8 + j;
                    int c = 2 * P.Cin, xx = x0, pp = 0, yy = y0;   // box beyond the last tap: channel coordinate out of range -> zeros
                    if (b < nbox) {
                        const int tap = b / cib, cb = b - tap * cib;
                        const int ky = P.ksize == 3 ? tap / 3 : 0, kx = P.ksize == 3 ? tap % 3 : 0;
                        if (P.stride == 1) {
                            const int pad = P.ksize == 3 ? 1 : 0;
                            c = cb * 64; xx = x0 + kx - pad; yy = y0 + ky - pad;
                        } else {   // (2C, W/2, 2, H/2, N) view of the conv input: tap (ky, kx) of output (y, x) = input (2y + ky, 2x + kx)
                            c = (kx & 1) * P.Cin + cb * 64; xx = x0 + (kx >> 1); pp = ky & 1; yy = y0 + (ky >> 1);
                        }
                    }
                    tma_load_5d(st + WG_YB + j * WG_BOX, &tmA, &full[stage], c, xx, pp, yy, img);
                }
                if (++stage == WG_STAGES) { stage = 0; phase ^= 1; }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // instruction descriptor: D fp32, M = 128, N = 256, both operands MN-major, formats per operand
        const uint32_t idesc = (1u << 4) | ((P.y_f16 ? 0u : 1u) << 7) | ((P.a_f16 ? 0u : 1u) << 10) | (1u << 15) | (1u << 16) |
                               (static_cast<uint32_t>(256 >> 3) << 17) | (static_cast<uint32_t>(128 >> 4) << 24);
        int stage = 0;
        uint32_t phase = 0;
        uint32_t acc = 0;
        for (int p = p_begin; p < p_end; ++p) {
            mbar_wait(&full[stage], phase);
            tc_fence_after();
            const uint32_t st = smem_u32(smem + stage * WG_STAGE);
            if (elect_one()) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {          // 64 pixels = four K steps of 16 (two 8-row groups = 2048 bytes each)
                    const uint64_t dy = umma_desc_mn_sw128(st + k * 2048, WG_BOX);
#pragma unroll
                    for (int hf = 0; hf < 2; ++hf) {
                        const uint64_t da = umma_desc_mn_sw128(st + WG_YB + hf * 4 * WG_BOX + k * 2048, WG_BOX);
                        umma_bf16_ss(tmem + hf * 256, dy, da, idesc, acc | k);
                    }
                }
                umma_commit(&empty[stage]);
            }
            __syncwarp();
            acc = 1;
            if (++stage == WG_STAGES) { stage = 0; phase ^= 1; }
        }
        if (elect_one()) umma_commit(done);
        __syncwarp();
    } else {
        // ------------------------------------------------------------ epilogue: TMEM -> partial tile in global memory
        const int q = warp & 3;
        const int co = mb * 128 + q * 32 + lane;
        mbar_wait(done, 0);
        tc_fence_after();
        float* rowp = P.part + (static_cast<long long>(s) * P.Cout + co) * taps * P.Cin;
#pragma unroll 1
        for (int c = 0; c < 16; ++c) {
            uint32_t v[32];
            tmem_ld_32x32(tmem + c * 32 + (static_cast<uint32_t>(q * 32) << 16), v);
            tmem_ld_wait();
            const int b = cg * 8 + (c >> 1);
            if (co < P.Cout && b < nbox) {
                const int tap = b / cib, cb = b - tap * cib;
                float4* dst = reinterpret_cast<float4*>(rowp + tap * P.Cin + cb * 64 + (c & 1) * 32);
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    dst[i] = make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]), __uint_as_float(v[4 * i + 2]),
                                         __uint_as_float(v[4 * i + 3]));
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem, 512);
    }
}

// ------------------------------------------------------------------------------------------------ halo variant
constexpr int WH_YB = 2 * WG_BOX;                   // dY tile: 64 pixels (8 x 8) x two 64-channel blocks
constexpr int WH_HALO = 13 * 1024;                  // 10 x 10 halo pixels x 128 bytes = 12800, padded to the swizzle period
constexpr int WH_STAGE = WH_YB + 2 * WH_HALO;       // 43008
constexpr int WH_STAGES = 4;
constexpr int WH_SMEM = WH_STAGES * WH_STAGE + 256 + 1024;
constexpr int WH_HALO_TX = 100 * 128;               // bytes one halo box delivers

struct WgradHaloParams {
    int N, H, W, Cout, Cin;
    int tiles_x, tiles_y;    // 8 x 8 pixel patches per image
    int groups, splits, per_split;
    int y_f16, a_f16;
    float* part;             // [splits][Cout][9 * Cin]
};

__device__ __forceinline__ uint64_t umma_desc_mn_view(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);
    d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}

__global__ void __launch_bounds__(WG_THREADS, 1)
wgrad_halo_kernel(const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmA,
                  const __grid_constant__ WgradHaloParams P) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + WH_STAGES * WH_STAGE);
    uint64_t* empty = full + WH_STAGES;
    uint64_t* done = empty + WH_STAGES;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(done + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    const int s = static_cast<int>(blockIdx.x % static_cast<unsigned>(P.splits));
    const int r = static_cast<int>(blockIdx.x / static_cast<unsigned>(P.splits));
    const int g = r % P.groups, mb = r / P.groups;
    const int total = P.N * P.tiles_y * P.tiles_x;
    const int p_begin = s * P.per_split, p_end = min(total, p_begin + P.per_split);
    // items of this CTA: item i = (input-channel chunk i / 3, kernel row i % 3)
    const int nitems = 3 * (P.Cin / 64);
    const int i0 = 2 * g, i1 = 2 * g + 1;
    const bool has1 = i1 < nitems;
    const int ch0 = i0 / 3, ky0 = i0 - 3 * ch0;
    const int ch1 = has1 ? i1 / 3 : ch0, ky1 = has1 ? i1 - 3 * ch1 : 0;
    const bool two_halos = has1 && ch1 != ch0;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmY);
        tma_prefetch_desc(&tmA);
        for (int i = 0; i < WH_STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        mbar_init(done, 1);
        fence_mbar_init();
    }
    if (warp == 1) { tmem_alloc(tmem_ptr, 512); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_ptr;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            const uint32_t bytes = WH_YB + (two_halos ? 2 : 1) * WH_HALO_TX;
            for (int p = p_begin; p < p_end; ++p) {
                const int tx = p % P.tiles_x;
                const int q = p / P.tiles_x;
                const int ty = q % P.tiles_y, img = q / P.tiles_y;
                const int x0 = tx * 8, y0 = ty * 8;
                mbar_wait(&empty[stage], phase ^ 1);
                uint8_t* st = smem + stage * WH_STAGE;
                mbar_arrive_expect_tx(&full[stage], bytes);
                for (int i = 0; i < 2; ++i) tma_load_5d(st + i * WG_BOX, &tmY, &full[stage], mb * 128 + i * 64, x0, 0, y0, img);
                tma_load_5d(st + WH_YB, &tmA, &full[stage], ch0 * 64, x0 - 1, 0, y0 - 1, img);
                if (two_halos) tma_load_5d(st + WH_YB + WH_HALO, &tmA, &full[stage], ch1 * 64, x0 - 1, 0, y0 - 1, img);
                if (++stage == WH_STAGES) { stage = 0; phase ^= 1; }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // instruction descriptor: D fp32, M = 128, N = 192 (three kx taps x 64 channels), both operands MN-major
        const uint32_t idesc = (1u << 4) | ((P.y_f16 ? 0u : 1u) << 7) | ((P.a_f16 ? 0u : 1u) << 10) | (1u << 15) | (1u << 16) |
                               (static_cast<uint32_t>(192 >> 3) << 17) | (static_cast<uint32_t>(128 >> 4) << 24);
        int stage = 0;
        uint32_t phase = 0;
        uint32_t acc = 0;
        const uint32_t h1_off = two_halos ? WH_HALO : 0;
        for (int p = p_begin; p < p_end; ++p) {
            mbar_wait(&full[stage], phase);
            tc_fence_after();
            const uint32_t st = smem_u32(smem + stage * WH_STAGE);
            if (elect_one()) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {          // K step = patch rows 2k, 2k+1 (two groups of eight pixels)
                    const uint64_t dy = umma_desc_mn_sw128(st + k * 2048, WG_BOX);
                    const uint64_t d0 = umma_desc_mn_view(st + WH_YB + ((2 * k + ky0) * 10) * 128, 128, 1280);
                    umma_bf16_ss(tmem, dy, d0, idesc, acc | k);
                    if (has1) {
                        const uint64_t d1 = umma_desc_mn_view(st + WH_YB + h1_off + ((2 * k + ky1) * 10) * 128, 128, 1280);
                        umma_bf16_ss(tmem + 192, dy, d1, idesc, acc | k);
                    }
                }
                umma_commit(&empty[stage]);
            }
            __syncwarp();
            acc = 1;
            if (++stage == WH_STAGES) { stage = 0; phase ^= 1; }
        }
        if (elect_one()) umma_commit(done);
        __syncwarp();
    } else {
        // ------------------------------------------------------------ epilogue: TMEM -> partial tile in global memory
        const int q = warp & 3;
        const int co = mb * 128 + q * 32 + lane;
        mbar_wait(done, 0);
        tc_fence_after();
        float* rowp = P.part + (static_cast<long long>(s) * P.Cout + co) * 9 * P.Cin;
#pragma unroll 1
        for (int c = 0; c < 12; ++c) {             // 32-column chunks: item c / 6, kx = (c % 6) / 2, half c & 1
            uint32_t v[32];
            tmem_ld_32x32(tmem + c * 32 + (static_cast<uint32_t>(q * 32) << 16), v);
            tmem_ld_wait();
            const int it = c / 6, kx = (c - 6 * it) >> 1;
            if (co < P.Cout && (it == 0 || has1)) {
                const int tap = (it ? ky1 : ky0) * 3 + kx, ch = it ? ch1 : ch0;
                float4* dst = reinterpret_cast<float4*>(rowp + tap * P.Cin + ch * 64 + (c & 1) * 32);
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    dst[i] = make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]), __uint_as_float(v[4 * i + 2]),
                                         __uint_as_float(v[4 * i + 3]));
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem, 512);
    }
}

int make_nhwc_map(CUtensorMap* tm, const void* base, int N, int H, int W, int C, int stride, int box_x = 16, int box_y = 4) {
    uint64_t dims[5], str[4];
    uint32_t box[5] = {64, static_cast<uint32_t>(box_x), 1, static_cast<uint32_t>(box_y), 1};
    if (stride == 1) {
        dims[0] = C; dims[1] = W; dims[2] = 1; dims[3] = H; dims[4] = N;
        str[0] = 2ull * C; str[1] = 2ull * C * W; str[2] = 2ull * C * W; str[3] = 2ull * C * W * H;
    } else {
        dims[0] = 2ull * C; dims[1] = W / 2; dims[2] = 2; dims[3] = H / 2; dims[4] = N;
        str[0] = 2ull * 2 * C; str[1] = 2ull * C * W; str[2] = 2ull * C * W * 2; str[3] = 2ull * C * W * H;
    }
    return make_tmap(tm, base, 5, dims, str, box);
}

}  // namespace

WgradMnPlan bwd_wgrad_mn_plan(int N, int H, int W, int Cout, int Cin, int ks) {
    WgradMnPlan p{};
    const int taps = ks * ks;
    p.tiles_x = (W + 15) / 16;
    p.tiles_y = (H + 3) / 4;
    p.col_groups = (taps * (Cin / 64) + 7) / 8;
    p.m_blocks = (Cout + 127) / 128;
    const int total = N * p.tiles_y * p.tiles_x;
    const int base = p.m_blocks * p.col_groups;
    int splits = std::max(1, (2 * 148 + base - 1) / base);       // ~two waves of CTAs
    splits = std::min(splits, std::max(1, total / 8));            // at least eight 64-pixel stages per CTA
    p.per_split = (total + splits - 1) / splits;
    p.splits = (total + p.per_split - 1) / p.per_split;           // every split owns at least one patch
    int max_splits = p.splits;
    if (ks == 3) {   // halo variant (stride 1): 8 x 8 patches, two (chunk, kernel row) items per CTA
        p.h_tiles_x = (W + 7) / 8;
        p.h_tiles_y = (H + 7) / 8;
        p.h_groups = (3 * (Cin / 64) + 1) / 2;
        const int htotal = N * p.h_tiles_y * p.h_tiles_x;
        const int hbase = p.m_blocks * p.h_groups;
        int hs = std::max(1, (2 * 148 + hbase - 1) / hbase);
        hs = std::min(hs, std::max(1, htotal / 16));              // at least sixteen 64-pixel stages per CTA
        p.h_per_split = (htotal + hs - 1) / hs;
        p.h_splits = (htotal + p.h_per_split - 1) / p.h_per_split;
        max_splits = std::max(max_splits, p.h_splits);
    }
    p.part_bytes = (static_cast<size_t>(max_splits) * Cout * taps * Cin * sizeof(float) + 255) / 256 * 256;
    return p;
}

static bool wgrad_halo_enabled() {
    static const bool on = !(getenv("VT_B200_NO_WGRAD_HALO") && getenv("VT_B200_NO_WGRAD_HALO")[0] == '1');
    return on;
}

// dy: [N][H][W][Cout] 16-bit (y_fmt), a: the conv input [N][stride*H][stride*W][Cin] 16-bit (a_fmt), both NHWC
int bwd_conv_wgrad_mn(const BwdEnv& e, const WgradMnPlan& p, const void* dy, int y_fmt, const void* a, int a_fmt, float* part,
                      int N, int H, int W, int Cout, int Cin, int ks, int stride, int* splits_used) {
    VT_CHECK(!e.fp32, "the tcgen05 weight-gradient kernel serves the 16-bit mode");
    VT_CHECK(Cin % 64 == 0 && Cout % 64 == 0, "weight gradient: channels must be multiples of 64");
    VT_CHECK((ks == 1 || ks == 3) && (stride == 1 || (stride == 2 && ks == 3)), "weight gradient: 3x3 / 1x1 stride 1, or 3x3 stride 2");
    VT_CHECK(y_fmt != FMT_F32 && a_fmt != FMT_F32, "weight gradient operands must be 16-bit");
    CUtensorMap ty, ta;
    if (ks == 3 && stride == 1 && wgrad_halo_enabled()) {
        VT_TRY(make_nhwc_map(&ty, dy, N, H, W, Cout, 1, 8, 8));
        VT_TRY(make_nhwc_map(&ta, a, N, H, W, Cin, 1, 10, 10));
        WgradHaloParams Q{};
        Q.N = N; Q.H = H; Q.W = W; Q.Cout = Cout; Q.Cin = Cin;
        Q.tiles_x = p.h_tiles_x; Q.tiles_y = p.h_tiles_y; Q.groups = p.h_groups; Q.splits = p.h_splits; Q.per_split = p.h_per_split;
        Q.y_f16 = y_fmt == FMT_F16; Q.a_f16 = a_fmt == FMT_F16; Q.part = part;
        static SmemAttrOnce once_h;
        VT_TRY(ensure_dyn_smem(once_h, wgrad_halo_kernel, WH_SMEM));
        const unsigned grid = static_cast<unsigned>(p.m_blocks) * p.h_groups * p.h_splits;
        profiler_begin(e.prof, KC_BWD, e.s, 2.0 * N * H * W * static_cast<double>(Cout) * Cin * 9, 0);
        wgrad_halo_kernel<<<grid, WG_THREADS, WH_SMEM, e.s>>>(ty, ta, Q);
        profiler_end(e.prof, KC_BWD, e.s);
        VT_CUDA(cudaGetLastError());
        *splits_used = p.h_splits;
        return 0;
    }
    *splits_used = p.splits;
    VT_TRY(make_nhwc_map(&ty, dy, N, H, W, Cout, 1));
    VT_TRY(make_nhwc_map(&ta, a, N, stride * H, stride * W, Cin, stride));
    WgradMnParams P{};
    P.N = N; P.H = H; P.W = W; P.Cout = Cout; P.Cin = Cin; P.ksize = ks; P.stride = stride;
    P.tiles_x = p.tiles_x; P.tiles_y = p.tiles_y; P.col_groups = p.col_groups; P.splits = p.splits; P.per_split = p.per_split;
    P.y_f16 = y_fmt == FMT_F16; P.a_f16 = a_fmt == FMT_F16; P.part = part;
    static SmemAttrOnce once;
    VT_TRY(ensure_dyn_smem(once, wgrad_mn_kernel, WG_SMEM));
    const unsigned grid = static_cast<unsigned>(p.m_blocks) * p.col_groups * p.splits;
    profiler_begin(e.prof, KC_BWD, e.s, 2.0 * N * H * W * static_cast<double>(Cout) * Cin * ks * ks, 0);
    wgrad_mn_kernel<<<grid, WG_THREADS, WG_SMEM, e.s>>>(ty, ta, P);
    profiler_end(e.prof, KC_BWD, e.s);
    VT_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace vt
