// Host side of the tcgen05 implicit-GEMM kernel: builds the TMA tensor maps and the slab
// table for each conv / GEMM and launches igemm_kernel<BLOCK_N>.
#include <cudaTypedefs.h>

#include <cstdlib>
#include <mutex>

#include "vt_conv3.cuh"
#include "vt_convin.cuh"
#include "vt_igemm.cuh"
#include "vt_internal.h"

namespace vt {

// cuTensorMapEncodeTiled is a driver entry point; fetch it through the runtime so the
// library has no link-time dependency on libcuda.
static PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
    });
    return fn;
}

// bf16 tensor map, 128-byte swizzle, zero fill out of bounds.
int make_tmap(CUtensorMap* tm, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                     const uint32_t* box) {
    auto fn = get_encode_fn();
    VT_CHECK(fn != nullptr, "cuTensorMapEncodeTiled entry point not available (no CUDA driver?)");
    uint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(base), dims, strides_bytes, box,
                    estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        std::string m = "cuTensorMapEncodeTiled failed with CUresult " + std::to_string(static_cast<int>(r)) + " rank " +
                        std::to_string(rank) + " dims";
        for (int i = 0; i < rank; ++i) m += " " + std::to_string(dims[i]);
        m += " box";
        for (int i = 0; i < rank; ++i) m += " " + std::to_string(box[i]);
        set_error(m);
        return -3;
    }
    return 0;
}

// two-stage GroupNorm statistics: point the kernel at the partial buffer (before the launch) ...
static int bind_stats(IgemmParams& P, const StatsScratch& ws) {
    if (P.stats == nullptr) return 0;
    VT_CHECK(ws.parts >= 1 && ws.part_index >= 0 && ws.part_index < ws.parts, "bad statistics part index");
    const int tiles = P.tiles_x * P.tiles_y;
    const size_t need = static_cast<size_t>(P.NB) * ws.parts * tiles * (P.n_total / P.group_size) * 2 * sizeof(float);
    VT_CHECK(ws.part != nullptr && ws.bytes >= need, "fused GroupNorm statistics need a stats_ws scratch buffer");
    P.stats_part = ws.part;
    P.stats_rows = ws.parts * tiles;
    P.stats_row0 = ws.part_index * tiles;
    return 0;
}
// ... and reduce the per-tile rows in a fixed order right behind it on the same stream
static int finish_stats(const IgemmParams& P, cudaStream_t stream, Profiler* prof) {
    if (P.stats == nullptr || P.stats_row0 + P.tiles_x * P.tiles_y != P.stats_rows) return 0;   // not the last part
    return launch_gn_finalize(P.stats_part, P.stats, P.NB, P.stats_rows, P.n_total / P.group_size, stream, prof);
}

static int num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

template <int BLOCK_N, int MT, bool PAIR>
static int launch_variant(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b, const IgemmParams& P,
                          cudaStream_t stream) {
    using Cfg = IgemmCfg<BLOCK_N, MT, PAIR>;
    static SmemAttrOnce once;
    VT_TRY(ensure_dyn_smem(once, igemm_kernel<BLOCK_N, MT, PAIR>, Cfg::SMEM_BYTES));
    const long long tiles = 1LL * P.NB * P.tiles_x * P.tiles_y * P.n_blocks;
    int grid = static_cast<int>(tiles < num_sms() ? tiles : num_sms());
    if (!PAIR) {
        igemm_kernel<BLOCK_N, MT, PAIR><<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, stream>>>(a0, a1, b, P);
    } else {
        grid &= ~1;   // whole clusters of two CTAs (the tile count is even: dispatch() checked)
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(grid);
        cfg.blockDim = dim3(Cfg::THREADS);
        cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
        cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        VT_CUDA(cudaLaunchKernelEx(&cfg, igemm_kernel<BLOCK_N, MT, PAIR>, a0, a1, b, P));
    }
    VT_CUDA(cudaGetLastError());
    return 0;
}

static int pick_block_n(int n_total) {
    if (n_total >= 256) return 256;
    if (n_total >= 128) return 128;
    return 32;
}

// CTA tile configurations: 128 rows x 256 columns, 2x128 rows x 128 columns (two pixel tiles share
// every weight chunk: same bytes per flop as the 128x256 tile), 128 x 32 for conv_out.
static int pick_mt(int block_n) { return block_n == 128 ? 2 : 1; }

static bool pairing_enabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("VT_B200_NO_PAIR");
        v = !(e && e[0] == '1');
    }
    return v != 0;
}

// CTA pairs (cta_group::2) whenever the pixel tiles of an image pair up: two neighbouring tiles of the same image and
// n-block share every weight chunk, each CTA streaming half of it
static bool igemm_pair_ok(int block_n, const IgemmParams& P) {
    return pairing_enabled() && block_n != 32 && ((P.tiles_x * P.tiles_y) % 2 == 0);
}
// `pair`: the caller built the B tensor map with a box of block_n / 2 rows (each CTA loads its half)
static int dispatch(int block_n, const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b,
                    const IgemmParams& P0, cudaStream_t stream, bool pair) {
    IgemmParams P = P0;
    P.pair = pair ? 1 : 0;
    switch (block_n) {
        case 256: return pair ? launch_variant<256, 1, true>(a0, a1, b, P, stream) : launch_variant<256, 1, false>(a0, a1, b, P, stream);
        case 128: return pair ? launch_variant<128, 2, true>(a0, a1, b, P, stream) : launch_variant<128, 2, false>(a0, a1, b, P, stream);
        case 32: return launch_variant<32, 1, false>(a0, a1, b, P, stream);
    }
    set_error("unsupported BLOCK_N");
    return -2;
}

// 5-D activation map (c, x, p, y, img).  stride 1: dims {C, W, 1, H, N}.
// stride 2: the input [N][Hin][Win][C] is viewed as {2C, Win/2, 2, Hin/2, N}: channel
// coordinate pw*C + c, x = column pair, p = row parity, y = row pair -- so a stride-2 tap
// (kh, kw) is a plain box at (c0 + (kw&1)*C, ox + (kw>>1), kh&1, oy + (kh>>1)) and the
// bottom/right zero padding of Downsample2D is TMA out-of-bounds fill.
static int make_act_map(CUtensorMap* tm, const void* base, int N, int H, int W, int C, int stride, int tw, int th) {
    uint64_t dims[5], str[4];
    uint32_t box[5] = {64, static_cast<uint32_t>(tw), 1, static_cast<uint32_t>(th), 1};
    if (stride == 1) {
        dims[0] = C; dims[1] = W; dims[2] = 1; dims[3] = H; dims[4] = N;
        str[0] = 2ull * C;
        str[1] = 2ull * C * W;  // size-1 dimension: any legal stride
        str[2] = 2ull * C * W;
        str[3] = 2ull * C * W * H;
    } else {
        // Odd sizes: the last pixel / row pair is partial.  Its missing half would be the next row's (image's) first
        // pixel, but only taps of output pixels beyond floor(W/2) x floor(H/2) reach it, and those are never stored
        // (diffusers' pad-right/bottom + stride-2 conv yields floor(size/2) outputs and never reads the pad then).
        dims[0] = 2ull * C; dims[1] = (W + 1) / 2; dims[2] = 2; dims[3] = (H + 1) / 2; dims[4] = N;
        str[0] = 2ull * 2 * C;
        str[1] = 2ull * C * W;
        str[2] = 2ull * C * W * 2;
        str[3] = 2ull * C * W * H;
    }
    return make_tmap(tm, base, 5, dims, str, box);
}

int launch_conv(const ConvOp& op, cudaStream_t stream, Profiler* prof) {
    VT_CHECK(op.ksize == 1 || op.ksize == 3, "conv kernel size must be 1 or 3");
    VT_CHECK(op.stride == 1 || (op.stride == 2 && op.ksize == 3), "stride must be 1, or 2 with a 3x3 kernel");
    VT_CHECK(op.Cin % 64 == 0 && op.Cin > 0, "Cin must be a multiple of 64");
    VT_CHECK(op.Cout % 32 == 0 && op.Cout > 0, "Cout must be a multiple of 32");
    VT_CHECK(op.Cs % 64 == 0, "shortcut channels must be a multiple of 64");
    VT_CHECK(op.sc_in == nullptr || op.stride == 1, "shortcut slab needs a stride-1 conv");
    const int Hout = op.stride == 1 ? op.Hin : op.Hin / 2;
    const int Wout = op.stride == 1 ? op.Win : op.Win / 2;
    VT_CHECK(!op.up2 || (op.ksize == 3 && op.stride == 1 && op.sc_in == nullptr && op.residual == nullptr),
             "sub-pixel upsample conv: 3x3, stride 1, no shortcut / residual");
    const int taps = op.up2 ? 4 : op.ksize * op.ksize;
    const int Ktot = taps * op.Cin + (op.sc_in ? op.Cs : 0);
    const int block_n = pick_block_n(op.Cout);

    IgemmParams P{};
    P.W = Wout; P.H = Hout; P.NB = op.N;
    const int mt = pick_mt(block_n);
    P.tw = 16; P.th = 8; P.tw_log2 = 4;
    P.sub_dx = 0; P.sub_dy = 1;  // sub-tiles stacked vertically: one TMA box of th*mt rows
    P.ax1 = 8; P.ay1 = 0; P.ax2 = 0; P.ay2 = 1;  // 16x8 patch: row r+8 = 8 pixels right, r+16 = next image row
    VT_CHECK(1LL * Hout * Wout * op.Cout < (1LL << 31), "conv output of one image exceeds 2^31 elements");
    P.tiles_x = (Wout + P.tw - 1) / P.tw;
    P.tiles_y = (Hout + P.th * mt - 1) / (P.th * mt);
    P.n_total = op.Cout;
    P.n_blocks = (op.Cout + block_n - 1) / block_n;
    P.a_batched = 1; P.b_batched = 0;
    P.out_fmt = op.out_fmt;
    P.group_size = op.stats ? op.Cout / 32 : 0;
    VT_CHECK(op.out_fmt != 2 || op.residual == nullptr || (op.raw_f16 && !op.residual_fp32),
             "fp16 output takes no residual other than an fp16 one");
    VT_CHECK(op.out_fmt != 0 || op.residual == nullptr || op.residual_fp32 || !op.raw_f16,
             "bf16 output takes no fp16 residual");
    VT_CHECK(op.stats == nullptr || (op.Cout % 32 == 0 && (P.group_size == 4 || P.group_size == 8 || P.group_size == 16)),
             "fused GroupNorm statistics need 4, 8 or 16 channels per group");
    P.alpha = op.alpha; P.raw_f16 = op.raw_f16;
    P.bias = op.bias; P.residual = op.residual; P.res_fp32 = op.residual_fp32; P.out = op.out; P.ld_out = op.Cout;
    P.out_bstride = 1LL * Hout * Wout * op.Cout; P.stats = op.stats;
    if (op.up2) {  // strided view of the 2x-upsampled output: this launch writes one pixel parity
        VT_CHECK(4LL * Hout * Wout * op.Cout < (1LL << 31), "upsampled conv output of one image exceeds 2^31 elements");
        const size_t esz = op.out_fmt == FMT_F32 ? 4 : 2;
        P.out = static_cast<char*>(op.out) + (static_cast<size_t>(op.up_py) * 2 * Wout + op.up_px) * op.Cout * esz;
        P.out_row_pitch = 2 * (2 * Wout) * op.Cout;
        P.out_px_stride = 2 * op.Cout;
        P.out_bstride = 4LL * Hout * Wout * op.Cout;
    }

    int ns = 0;
    if (op.up2) {
        for (int ty = 0; ty < 2; ++ty)
            for (int tx = 0; tx < 2; ++tx) {
                IgemmSlab& s = P.slabs[ns];
                s.map = 0; s.c_base = 0; s.p = 0;
                s.dy = op.up_py == 0 ? ty - 1 : ty;
                s.dx = op.up_px == 0 ? tx - 1 : tx;
                s.kb_base = ns * op.Cin;
                s.nchunks = op.Cin / 64;
                s.f16 = op.in_f16;
                ++ns;
            }
    } else
    for (int kh = 0; kh < op.ksize; ++kh)
        for (int kw = 0; kw < op.ksize; ++kw) {
            IgemmSlab& s = P.slabs[ns];
            s.map = 0;
            if (op.ksize == 1) { s.c_base = 0; s.dx = 0; s.p = 0; s.dy = 0; }
            else if (op.stride == 1) { s.c_base = 0; s.dx = kw - 1; s.p = 0; s.dy = kh - 1; }
            else { s.c_base = (kw & 1) * op.Cin; s.dx = kw >> 1; s.p = kh & 1; s.dy = kh >> 1; }
            s.kb_base = ns * op.Cin;
            s.nchunks = op.Cin / 64;
            s.f16 = op.in_f16;
            ++ns;
        }
    if (op.sc_in) {
        IgemmSlab& s = P.slabs[ns];
        s.map = 1; s.c_base = 0; s.dx = 0; s.p = 0; s.dy = 0;
        s.kb_base = taps * op.Cin; s.nchunks = op.Cs / 64;
        s.f16 = op.raw_f16;  // the shortcut operand is a raw activation: bf16 (fp16 when raw_f16), and so are its weight columns
        ++ns;
    }
    P.num_slabs = ns;

    const bool pair = igemm_pair_ok(block_n, P);
    CUtensorMap a0, a1, b;
    VT_TRY(make_act_map(&a0, op.in, op.N, op.Hin, op.Win, op.Cin, op.stride, P.tw, P.th * mt));
    if (op.sc_in) VT_TRY(make_act_map(&a1, op.sc_in, op.N, Hout, Wout, op.Cs, 1, P.tw, P.th * mt));
    else a1 = a0;
    {
        uint64_t dims[3] = {static_cast<uint64_t>(Ktot), static_cast<uint64_t>(op.Cout), 1};
        uint64_t str[2] = {2ull * Ktot, 2ull * Ktot * op.Cout};
        uint32_t box[3] = {64, static_cast<uint32_t>(pair ? block_n / 2 : block_n), 1};
        VT_TRY(make_tmap(&b, op.w, 3, dims, str, box));
    }
    const double flops = 2.0 * op.N * Hout * Wout * static_cast<double>(op.Cout) * Ktot;
    const double bytes = 2.0 * op.N * (1.0 * op.Hin * op.Win * op.Cin + 1.0 * Hout * Wout * op.Cout);
    VT_TRY(bind_stats(P, op.stats_ws));
    const KernelClass kc = op.kclass >= 0 ? static_cast<KernelClass>(op.kclass) : KC_IGEMM;
    profiler_begin(prof, kc, stream, flops, bytes);
    int rc = dispatch(block_n, a0, a1, b, P, stream, pair);
    profiler_end(prof, kc, stream);
    VT_TRY(rc);
    return finish_stats(P, stream, prof);
}

int launch_conv_in(const ConvInOp& op, cudaStream_t stream, Profiler* prof) {
    VT_CHECK(op.img && op.w && op.out, "conv_in: null operand");
    VT_CHECK(1LL * op.H * op.W * 128 < (1LL << 31), "conv_in output of one image exceeds 2^31 elements");
    IgemmParams P{};
    P.W = op.W; P.H = op.H; P.NB = op.N;
    P.tw = 128; P.th = 1; P.tw_log2 = 7;          // tile = 2 x 128 consecutive pixels of one image row
    P.sub_dx = 1; P.sub_dy = 0;
    P.ax1 = 8; P.ay1 = 0; P.ax2 = 16; P.ay2 = 0;
    P.tiles_x = (op.W + 255) / 256;
    P.tiles_y = op.H;
    P.n_total = 128; P.n_blocks = 1;
    P.out_fmt = op.out_f16 ? FMT_F16 : FMT_BF16;
    P.group_size = op.stats ? 4 : 0;
    P.alpha = 1.f;
    P.bias = op.bias; P.out = op.out; P.ld_out = 128;
    P.out_bstride = 1LL * op.H * op.W * 128; P.stats = op.stats;
    CUtensorMap b;
    {
        uint64_t dims[3] = {64, 128, 1};
        uint64_t str[2] = {128, 128 * 128};
        uint32_t box[3] = {64, 128, 1};
        VT_TRY(make_tmap(&b, op.w, 3, dims, str, box));
    }
    CUtensorMap o;   // output viewed as (channel, x, image row): the store clips pixels beyond the row end
    {
        uint64_t dims[3] = {128, static_cast<uint64_t>(op.W), static_cast<uint64_t>(op.N) * op.H};
        uint64_t str[2] = {256, 256ull * op.W};
        uint32_t box[3] = {64, 32, 1};
        VT_TRY(make_tmap(&o, op.out, 3, dims, str, box));
    }
    VT_CHECK(3LL * op.H * op.W < (1LL << 31), "conv_in image exceeds 2^31 elements");
    ConvInParams Q{op.img, op.in_fmt,
                   (op.W % (op.in_fmt == 0 ? 4 : 16) == 0 && reinterpret_cast<uintptr_t>(op.img) % 16 == 0) ? 1 : 0};
    static SmemAttrOnce once;
    VT_TRY(ensure_dyn_smem(once, conv_in_kernel, ConvInCfg::SMEM_BYTES));
    const long long tiles = 1LL * P.NB * P.tiles_x * P.tiles_y;
    const int grid = static_cast<int>(tiles < num_sms() ? tiles : num_sms());
    const double flops = 2.0 * op.N * op.H * op.W * 128.0 * 27.0;
    const double bytes = 1.0 * op.N * op.H * op.W * ((op.in_fmt ? 3.0 : 12.0) + 256.0);
    VT_TRY(bind_stats(P, op.stats_ws));
    profiler_begin(prof, KC_CONVIN, stream, flops, bytes);
    conv_in_kernel<<<grid, ConvInCfg::THREADS, ConvInCfg::SMEM_BYTES, stream>>>(b, o, P, Q);
    profiler_end(prof, KC_CONVIN, stream);
    VT_CUDA(cudaGetLastError());
    return finish_stats(P, stream, prof);
}

template <int BLOCK_N, int MT, bool TR, bool PAIR, int RAW>
static int launch_conv3_variant(const CUtensorMap& a, const CUtensorMap& b, const CUtensorMap& sc, const IgemmParams& P,
                                cudaStream_t stream) {
    using Cfg = Conv3Cfg<BLOCK_N, MT, TR, PAIR>;
    static SmemAttrOnce once;
    VT_TRY(ensure_dyn_smem(once, conv3_fused_kernel<BLOCK_N, MT, TR, PAIR, RAW>, Cfg::SMEM_BYTES));
    const long long tiles = 1LL * P.NB * P.tiles_x * P.tiles_y * P.n_blocks;
    int grid = static_cast<int>(tiles < num_sms() ? tiles : num_sms());
    if (!PAIR) {
        conv3_fused_kernel<BLOCK_N, MT, TR, PAIR, RAW><<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, stream>>>(a, b, sc, P);
    } else {
        grid &= ~1;   // whole clusters of two CTAs (tiles is even: the launcher checked)
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(grid);
        cfg.blockDim = dim3(Cfg::THREADS);
        cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
        cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        VT_CUDA(cudaLaunchKernelEx(&cfg, conv3_fused_kernel<BLOCK_N, MT, TR, PAIR, RAW>, a, b, sc, P));
    }
    VT_CUDA(cudaGetLastError());
    return 0;
}

// out = conv3x3(silu(gn(in))) + bias (+ residual); in: raw bf16 NHWC; weights fp16 [Cout][9*Cin].
int launch_conv3_fused(const Conv3FusedOp& op, cudaStream_t stream, Profiler* prof) {
    VT_CHECK(op.Cin % 64 == 0 && op.Cin > 0 && op.Cin <= 512, "fused conv: Cin must be a multiple of 64, at most 512");
    VT_CHECK(op.Cout == 128 || (op.Cout >= 256 && op.Cout % 32 == 0), "fused conv: Cout must be 128 or >= 256");
    VT_CHECK(op.Cin % 32 == 0 && (op.Cin / 32) % 4 == 0, "fused conv: GroupNorm(32) groups must hold a multiple of 4 channels");
    VT_CHECK(op.gn_stats && op.gamma && op.beta, "fused conv needs the input statistics and affine parameters");
    VT_CHECK(!op.residual_fp32 && (op.out_fmt == FMT_F32 || op.out_fmt == (op.raw_f16 ? FMT_F16 : FMT_BF16)),
             "fused conv: output must be fp32 or the raw 16-bit format, the residual the raw 16-bit format");
    // 128-channel layers: transposed variant (channels on the accumulator rows, an 8x32 pixel patch on the
    // columns); wider layers: 8x16 pixel patch x 256 channels
    const bool tr = op.Cout == 128;
    const int block_n = tr ? 128 : 256;
    const int pxw = 8, pxh = tr ? 32 : 16;
    const int H = op.H, W = op.W;
    IgemmParams P{};
    P.W = W; P.H = H; P.NB = op.N;
    P.tw = 8; P.th = 16; P.tw_log2 = 3;
    P.sub_dx = 1; P.sub_dy = 0;
    P.ax1 = 0; P.ay1 = 1; P.ax2 = 0; P.ay2 = 2;  // 8x16 patch: accumulator row r+8 is the next image row
    P.tiles_x = (W + pxw - 1) / pxw;
    P.tiles_y = (H + pxh - 1) / pxh;
    P.n_total = op.Cout;
    P.n_blocks = (op.Cout + block_n - 1) / block_n;
    P.num_slabs = 0;
    P.a_batched = 1; P.b_batched = 0;
    P.out_fmt = op.out_fmt;
    P.res_fp32 = op.residual_fp32;
    P.raw_f16 = op.raw_f16;
    P.group_size = op.stats ? op.Cout / 32 : 0;
    VT_CHECK(op.stats == nullptr || P.group_size == 4 || P.group_size == 8 || P.group_size == 16,
             "fused GroupNorm statistics need 4, 8 or 16 channels per group");
    P.alpha = 1.f;
    P.bias = op.bias; P.residual = op.residual; P.out = op.out; P.ld_out = op.Cout;
    P.out_bstride = 1LL * H * W * op.Cout; P.stats = op.stats;
    VT_CHECK(1LL * H * W * op.Cout < (1LL << 31), "conv output of one image exceeds 2^31 elements");
    P.gn_stats = op.gn_stats; P.gn_gamma = op.gamma; P.gn_beta = op.beta;
    P.gn_C = op.Cin; P.gn_gs = op.Cin / 32; P.gn_eps = op.eps; P.gn_silu = op.silu; P.cin_chunks = op.Cin / 64;
    VT_CHECK(op.sc_in == nullptr || (!tr && op.Cs % 64 == 0 && op.Cs > 0), "fused conv: shortcut slab needs Cout >= 256 and Cs % 64 == 0");
    P.sc_chunks = op.sc_in ? op.Cs / 64 : 0;

    // CTA pairs for the 256-wide variant when the pixel tiles pair up (VT_B200_NO_PAIR=1: single-CTA kernel)
    static const bool no_pair = [] { const char* e = getenv("VT_B200_NO_PAIR"); return e && e[0] == '1'; }();
    const bool pair = !tr && !no_pair && ((1LL * op.N * P.tiles_x * P.tiles_y) % 2 == 0) && num_sms() >= 2;
    P.pair = pair ? 1 : 0;

    CUtensorMap a, b, sc;
    VT_TRY(make_act_map(&a, op.in, op.N, H, W, op.Cin, 1, pxw + 2, pxh + 2));
    if (op.sc_in) VT_TRY(make_act_map(&sc, op.sc_in, op.N, H, W, op.Cs, 1, pxw, pxh));
    else sc = a;
    {
        const int Ktot = 9 * op.Cin + (op.sc_in ? op.Cs : 0);
        uint64_t dims[3] = {static_cast<uint64_t>(Ktot), static_cast<uint64_t>(op.Cout), 1};
        uint64_t str[2] = {2ull * Ktot, 2ull * Ktot * op.Cout};
        uint32_t box[3] = {64, static_cast<uint32_t>(pair ? block_n / 2 : block_n), 1};
        VT_TRY(make_tmap(&b, op.w, 3, dims, str, box));
    }
    const double flops = 2.0 * op.N * H * W * static_cast<double>(op.Cout) * (9 * op.Cin + (op.sc_in ? op.Cs : 0));
    const double bytes = 2.0 * op.N * H * W * (1.0 * op.Cin + op.Cout);
    VT_TRY(bind_stats(P, op.stats_ws));
    const KernelClass kc = tr ? KC_CONV3_T : KC_CONV3;
    profiler_begin(prof, kc, stream, flops, bytes);
    int rc;
    if (op.raw_f16)
        rc = tr ? launch_conv3_variant<128, 1, true, false, FMT_F16>(a, b, sc, P, stream)
                : pair ? launch_conv3_variant<256, 1, false, true, FMT_F16>(a, b, sc, P, stream)
                       : launch_conv3_variant<256, 1, false, false, FMT_F16>(a, b, sc, P, stream);
    else
        rc = tr ? launch_conv3_variant<128, 1, true, false, FMT_BF16>(a, b, sc, P, stream)
                : pair ? launch_conv3_variant<256, 1, false, true, FMT_BF16>(a, b, sc, P, stream)
                       : launch_conv3_variant<256, 1, false, false, FMT_BF16>(a, b, sc, P, stream);
    profiler_end(prof, kc, stream);
    VT_TRY(rc);
    return finish_stats(P, stream, prof);
}

int launch_gemm(const GemmOp& op, cudaStream_t stream, Profiler* prof) {
    VT_CHECK(op.K % 64 == 0 && op.K > 0, "GEMM K must be a multiple of 64");
    VT_CHECK(op.N % 32 == 0 && op.N > 0, "GEMM N must be a multiple of 32");
    const long long lda = op.lda ? op.lda : op.K, ldb = op.ldb ? op.ldb : op.K;
    const long long ldo = op.ld_out ? op.ld_out : op.N;
    const int block_n = pick_block_n(op.N);

    IgemmParams P{};
    P.W = op.M; P.H = 1; P.NB = op.batch;
    const int mt = pick_mt(block_n);
    P.tw = 128; P.th = 1; P.tw_log2 = 7;
    P.sub_dx = 1; P.sub_dy = 0;  // sub-tiles are consecutive 128-row blocks
    P.ax1 = 8; P.ay1 = 0; P.ax2 = 16; P.ay2 = 0;  // 128x1 patch: rows are consecutive output rows
    VT_CHECK(1LL * (op.M + 256) * ldo < (1LL << 31), "GEMM output of one batch exceeds 2^31 elements");
    P.tiles_x = (op.M + 128 * mt - 1) / (128 * mt);
    P.tiles_y = 1;
    P.n_total = op.N;
    P.n_blocks = (op.N + block_n - 1) / block_n;
    P.a_batched = op.a_batched; P.b_batched = op.b_batched;
    P.out_fmt = op.out_fmt;
    P.group_size = op.stats ? op.N / 32 : 0;
    VT_CHECK(op.out_fmt != 2 || op.residual == nullptr || (op.raw_f16 && !op.residual_fp32),
             "fp16 output takes no residual other than an fp16 one");
    VT_CHECK(op.out_fmt != 0 || op.residual == nullptr || op.residual_fp32 || !op.raw_f16,
             "bf16 output takes no fp16 residual");
    VT_CHECK(op.stats == nullptr || (P.group_size == 4 || P.group_size == 8 || P.group_size == 16),
             "fused GroupNorm statistics need 4, 8 or 16 channels per group");
    P.alpha = op.alpha; P.raw_f16 = op.raw_f16;
    P.bias = op.bias; P.residual = op.residual; P.res_fp32 = op.residual_fp32; P.out = op.out; P.ld_out = ldo;
    P.out_bstride = op.out_bstride ? op.out_bstride : 1LL * op.M * ldo; P.stats = op.stats;
    P.num_slabs = 1;
    P.slabs[0] = IgemmSlab{0, 0, 0, 0, 0, 0, op.K / 64, op.ab_f16};

    const bool pair = igemm_pair_ok(block_n, P);
    CUtensorMap a, b;
    {
        uint64_t dims[5] = {static_cast<uint64_t>(op.K), static_cast<uint64_t>(op.M), 1, 1,
                            static_cast<uint64_t>(op.a_batched ? op.batch : 1)};
        const uint64_t abs_ = 2ull * (op.a_bstride ? op.a_bstride : lda * op.M);
        uint64_t str[4] = {2ull * lda, abs_, abs_, abs_};
        uint32_t box[5] = {64, static_cast<uint32_t>(128 * mt), 1, 1, 1};
        VT_TRY(make_tmap(&a, op.A, 5, dims, str, box));
    }
    {
        const int brows = op.b_rows > 0 ? op.b_rows : op.N;  // rows beyond brows: TMA out-of-bounds zero fill
        uint64_t dims[3] = {static_cast<uint64_t>(op.K), static_cast<uint64_t>(brows),
                            static_cast<uint64_t>(op.b_batched ? op.batch : 1)};
        uint64_t str[2] = {2ull * ldb, 2ull * (op.b_bstride ? op.b_bstride : ldb * brows)};
        uint32_t box[3] = {64, static_cast<uint32_t>(pair ? block_n / 2 : block_n), 1};
        VT_TRY(make_tmap(&b, op.B, 3, dims, str, box));
    }
    // the epilogue addresses out as ((img*H + y)*W + x)*ld_out: batch stride is M*ld_out
    const double flops = 2.0 * op.batch * static_cast<double>(op.M) * op.N * op.K;
    const double bytes = 2.0 * op.batch * (1.0 * op.M * op.K + 1.0 * op.N * op.K) +
                         (op.out_fmt == 1 ? 4.0 : 2.0) * op.batch * op.M * op.N;
    VT_TRY(bind_stats(P, op.stats_ws));
    const KernelClass kc = op.kclass >= 0 ? static_cast<KernelClass>(op.kclass) : KC_IGEMM;
    profiler_begin(prof, kc, stream, flops, bytes);
    int rc = dispatch(block_n, a, a, b, P, stream, pair);
    profiler_end(prof, kc, stream);
    VT_TRY(rc);
    return finish_stats(P, stream, prof);
}

}  // namespace vt
