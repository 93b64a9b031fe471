// HBM-bound kernels of the encoder path: conv_in patch gather, GroupNorm statistics /
// apply (+SiLU), attention row softmax, moments -> latent, layout conversions.
// All are vectorised (16-byte accesses), channel-innermost (NHWC) and warp-shuffle based.
#include <type_traits>

#include "vt_internal.h"
#include "vt_ptx.cuh"

namespace vt {

static inline int sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xFFFFFFFFu, v, o));
    return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    return v;
}

// ------------------------------------------------------------------------------------------
// conv_in patch gather.  in: fp32 NCHW [N,3,H,W] in [-1,1]  (fmt 0)  or uint8 NHWC [N,H,W,3]
// (fmt 1; normalised (u/255 - 0.5)/0.5 exactly like ToTensor+Normalize, modules.py:136-140).
// out: [N,H,W,64] (bf16 or fp32) with k = (kh*3+kw)*3 + c for k < 27 and zeros above: the K=27
// contraction of conv_in becomes one 64-wide K chunk of the implicit-GEMM kernel.
// 8 threads per pixel, each writes 8 consecutive k.
template <int OFMT>
__global__ void __launch_bounds__(256) im2col3x3_kernel(const void* __restrict__ in, int fmt, void* __restrict__ out,
                                                        int N, int H, int W) {
    const long long total = 1LL * N * H * W * 8;
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += 256LL * gridDim.x) {
        const int j = static_cast<int>(i & 7);
        long long p = i >> 3;
        const int x = static_cast<int>(p % W);
        p /= W;
        const int y = static_cast<int>(p % H);
        const int n = static_cast<int>(p / H);
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int k = j * 8 + e;
            float val = 0.f;
            if (k < 27) {
                const int tap = k / 3, c = k - tap * 3;
                const int yy = y + tap / 3 - 1, xx = x + tap % 3 - 1;
                if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
                    if (fmt == 0) {
                        val = __ldg(static_cast<const float*>(in) + ((1LL * n * 3 + c) * H + yy) * W + xx);
                    } else {
                        const float u = static_cast<float>(
                            __ldg(static_cast<const unsigned char*>(in) + ((1LL * n * H + yy) * W + xx) * 3 + c));
                        val = (u / 255.0f - 0.5f) / 0.5f;
                    }
                }
            }
            v[e] = val;
        }
        if constexpr (OFMT != FMT_F32) {
            uint4 o = make_uint4(pack16x2<OFMT>(v[0], v[1]), pack16x2<OFMT>(v[2], v[3]), pack16x2<OFMT>(v[4], v[5]),
                                 pack16x2<OFMT>(v[6], v[7]));
            reinterpret_cast<uint4*>(out)[i] = o;
        } else {
            float4* o = reinterpret_cast<float4*>(out) + 2 * i;
            o[0] = make_float4(v[0], v[1], v[2], v[3]);
            o[1] = make_float4(v[4], v[5], v[6], v[7]);
        }
    }
}

int launch_im2col3x3(const void* in, int fmt, void* out, int out_fmt, int N, int H, int W, cudaStream_t s,
                     Profiler* prof) {
    const long long total = 1LL * N * H * W * 8;
    const int grid = static_cast<int>(std::min<long long>((total + 255) / 256, 1LL * sm_count() * 16));
    const double bytes = 1.0 * N * H * W * ((fmt ? 3.0 : 12.0) + 64.0 * (out_fmt == FMT_F32 ? 4 : 2));
    profiler_begin(prof, KC_IM2COL, s, 0, bytes);
    if (out_fmt == FMT_F32) im2col3x3_kernel<FMT_F32><<<grid, 256, 0, s>>>(in, fmt, out, N, H, W);
    else if (out_fmt == FMT_F16) im2col3x3_kernel<FMT_F16><<<grid, 256, 0, s>>>(in, fmt, out, N, H, W);
    else im2col3x3_kernel<FMT_BF16><<<grid, 256, 0, s>>>(in, fmt, out, N, H, W);
    profiler_end(prof, KC_IM2COL, s);
    VT_CUDA(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------
// GroupNorm statistics (used by the fp32 verification path, the unfused fallbacks and tests; the 16-bit
// path gets its statistics from the producing contraction's epilogue + gn_finalize_kernel).
// x: [N][HW][C] ; stats: [N][G][2] double (sum, sum of squares), WRITTEN.  One block per (image, group),
// fixed pixel -> thread assignment and a fixed-order tree: the result depends on the image only, not on the
// batch it is part of (no atomics).
template <int IFMT>   // element format of x: FMT_BF16 / FMT_F32 / FMT_F16
__global__ void __launch_bounds__(256) gn_stats_kernel(const void* __restrict__ xv, double* __restrict__ stats,
                                                       long long HW, int C, int G) {
    typedef typename std::conditional<IFMT == FMT_F32, float, bf16>::type T;   // 16-bit formats share the pointer type
    const T* x = static_cast<const T*>(xv);
    __shared__ double sh[2][256];
    const int g = blockIdx.x, n = blockIdx.y;
    const int cpg = C / G;              // a multiple of 4
    const int quads = cpg / 4;          // 16-byte (fp32) / 8-byte (bf16) pieces of one pixel's group slice
    const int qd = threadIdx.x % quads;
    const int lanes = 256 / quads;      // pixel lanes
    const int pl = threadIdx.x / quads;
    double s = 0.0, ss = 0.0;
    if (pl < lanes) {
        float fs = 0.f, fss = 0.f;
        int cnt = 0;
        for (long long p = pl; p < HW; p += lanes) {
            const T* px = x + (1LL * n * HW + p) * C + g * cpg + qd * 4;
            float a, b, c, d;
            if constexpr (IFMT != FMT_F32) {
                const uint2 u = *reinterpret_cast<const uint2*>(px);
                a = raw16_lo<IFMT>(u.x); b = raw16_hi<IFMT>(u.x); c = raw16_lo<IFMT>(u.y); d = raw16_hi<IFMT>(u.y);
            } else {
                const float4 f = *reinterpret_cast<const float4*>(px);
                a = f.x; b = f.y; c = f.z; d = f.w;
            }
            fs += (a + b) + (c + d);
            fss += (a * a + b * b) + (c * c + d * d);
            if (++cnt == 64) { s += fs; ss += fss; fs = fss = 0.f; cnt = 0; }
        }
        s += fs; ss += fss;
    }
    sh[0][threadIdx.x] = s;
    sh[1][threadIdx.x] = ss;
    __syncthreads();
    for (int w = 128; w >= 1; w >>= 1) {
        if (threadIdx.x < w) {
            sh[0][threadIdx.x] += sh[0][threadIdx.x + w];
            sh[1][threadIdx.x] += sh[1][threadIdx.x + w];
        }
        __syncthreads();
    }
    if (threadIdx.x < 2) stats[(1LL * n * G + g) * 2 + threadIdx.x] = sh[threadIdx.x][0];
}

int launch_gn_stats(const void* x, int x_fmt, double* stats, int N, long long HW, int C, int G, cudaStream_t s,
                    Profiler* prof) {
    VT_CHECK(G == 32 && C % (4 * G) == 0 && C / (4 * G) <= 256, "GroupNorm statistics need 32 groups of a multiple of 4 channels");
    dim3 grid(G, N);
    profiler_begin(prof, KC_GN_APPLY, s, 0, 1.0 * N * HW * C * (x_fmt == FMT_F32 ? 4 : 2));
    if (x_fmt == FMT_F32) gn_stats_kernel<FMT_F32><<<grid, 256, 0, s>>>(x, stats, HW, C, G);
    else if (x_fmt == FMT_F16) gn_stats_kernel<FMT_F16><<<grid, 256, 0, s>>>(x, stats, HW, C, G);
    else gn_stats_kernel<FMT_BF16><<<grid, 256, 0, s>>>(x, stats, HW, C, G);
    profiler_end(prof, KC_GN_APPLY, s);
    VT_CUDA(cudaGetLastError());
    return 0;
}

// Second stage of the epilogue statistics (vt_igemm.cuh / vt_conv3.cuh): part[img][tile][G][2] fp32 per-tile
// (sum, sumsq) -> stats[img][G][2] fp64.  A tile row is 2G consecutive floats; a block of 1024 threads owns four
// 16-byte column quads of one image x 256 row lanes: thread (quad, lane) adds rows lane, lane+256, ... in fp64 (all of
// an 8-row batch of loads in flight at once; a warp reads 8 rows x 64 contiguous bytes), then a fixed-order tree over
// the 256 lanes.  The order depends on the tile count only, so an image's statistics are the same bits in any batch.
// (Round 2 first had one block per (image, group) with 4-byte loads 2G floats apart and four loads in flight: 32 us per
// launch at 4096 tiles, 2-3 % of the whole encoder step.)
constexpr int GNF_QUADS = 4, GNF_LANES = 256;
__global__ void __launch_bounds__(GNF_QUADS * GNF_LANES) gn_finalize_kernel(const float* __restrict__ part,
                                                                            double* __restrict__ stats, int tiles, int G) {
    __shared__ double sh[GNF_LANES][GNF_QUADS][4];
    const int n = blockIdx.y;
    const int ql = threadIdx.x & (GNF_QUADS - 1), lane = threadIdx.x >> 2;
    const int quad = blockIdx.x * GNF_QUADS + ql;          // 16-byte column group of the 2G-float row
    const int row_f4 = G / 2;                              // float4 per row
    double a[4] = {0.0, 0.0, 0.0, 0.0};
    if (quad < row_f4) {
        const float4* p = reinterpret_cast<const float4*>(part) + (1LL * n * tiles) * row_f4 + quad;
        int t = lane;
        for (; t + 7 * GNF_LANES < tiles; t += 8 * GNF_LANES) {
            float4 v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = __ldg(p + 1LL * (t + i * GNF_LANES) * row_f4);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                a[0] += static_cast<double>(v[i].x); a[1] += static_cast<double>(v[i].y);
                a[2] += static_cast<double>(v[i].z); a[3] += static_cast<double>(v[i].w);
            }
        }
        for (; t < tiles; t += GNF_LANES) {
            const float4 v = __ldg(p + 1LL * t * row_f4);
            a[0] += static_cast<double>(v.x); a[1] += static_cast<double>(v.y);
            a[2] += static_cast<double>(v.z); a[3] += static_cast<double>(v.w);
        }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) sh[lane][ql][j] = a[j];
    __syncthreads();
    for (int w = GNF_LANES / 2; w >= 1; w >>= 1) {
        if (lane < w) {
#pragma unroll
            for (int j = 0; j < 4; ++j) sh[lane][ql][j] += sh[lane + w][ql][j];
        }
        __syncthreads();
    }
    if (lane == 0 && quad < row_f4) {
#pragma unroll
        for (int j = 0; j < 4; ++j) stats[1LL * n * G * 2 + quad * 4 + j] = sh[0][ql][j];
    }
}

int launch_gn_finalize(const float* part, double* stats, int N, int tiles, int G, cudaStream_t s, Profiler* prof) {
    VT_CHECK(part && stats && N > 0 && tiles > 0 && G > 0, "GroupNorm finalize: bad arguments");
    profiler_begin(prof, KC_GN_APPLY, s, 0, 8.0 * N * tiles * G);
    VT_CHECK(G % 2 == 0, "GroupNorm finalize: odd group count");
    gn_finalize_kernel<<<dim3((G / 2 + GNF_QUADS - 1) / GNF_QUADS, N), GNF_QUADS * GNF_LANES, 0, s>>>(part, stats, tiles, G);
    profiler_end(prof, KC_GN_APPLY, s);
    VT_CUDA(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------
// GroupNorm apply (+ optional SiLU): y = act((x - mean) * rstd * gamma + beta), statistics from
// (sum, sumsq) doubles.  One thread owns 8 consecutive channels of a fixed channel block and
// walks pixels, so scale/shift live in registers; every access is a 16-byte (bf16) or 2x16-byte
// (fp32) vector and a warp touches 512 contiguous bytes.
template <int IFMT, int OFMT, bool FAST>   // IFMT: element format of x (FMT_BF16 / FMT_F32 / FMT_F16)
__global__ void __launch_bounds__(256) gn_apply_kernel(const void* __restrict__ xv, void* __restrict__ yv,
                                                       const double* __restrict__ stats,
                                                       const float* __restrict__ gamma,
                                                       const float* __restrict__ beta, long long HW, int C, int G,
                                                       float eps, int silu) {
    typedef typename std::conditional<IFMT == FMT_F32, float, bf16>::type TI;
    const TI* x = static_cast<const TI*>(xv);
    const int n = blockIdx.y;
    const int cb = C / 8;              // 8-channel blocks per pixel
    const int tpb = min(cb, 256);      // threads spanning the channel dimension
    const int ppb = 256 / tpb;         // pixels per block step
    const int psub = threadIdx.x / tpb;
    const int cpg = C / G;
    const double cnt = static_cast<double>(HW) * cpg;
    const long long chunk = (HW + gridDim.x - 1) / gridDim.x;
    const long long p0 = blockIdx.x * chunk, p1 = min(HW, p0 + chunk);
    for (int c8 = threadIdx.x % tpb; c8 < cb; c8 += tpb) {
        float sc[8], sh[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int c = c8 * 8 + e;
            const int g = c / cpg;
            const double su = stats[(1LL * n * G + g) * 2], sq = stats[(1LL * n * G + g) * 2 + 1];
            const double mean = su / cnt;
            double var = sq / cnt - mean * mean;
            var = var > 0.0 ? var : 0.0;
            const float rstd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
            const float ga = gamma[c], be = beta[c];
            sc[e] = rstd * ga;
            sh[e] = be - static_cast<float>(mean) * rstd * ga;
        }
        for (long long p = p0 + psub; p < p1; p += ppb) {
            const long long off = (1LL * n * HW + p) * C + c8 * 8;
            float v[8];
            if constexpr (IFMT != FMT_F32) {
                const uint4 u = *reinterpret_cast<const uint4*>(x + off);
                v[0] = raw16_lo<IFMT>(u.x); v[1] = raw16_hi<IFMT>(u.x); v[2] = raw16_lo<IFMT>(u.y); v[3] = raw16_hi<IFMT>(u.y);
                v[4] = raw16_lo<IFMT>(u.z); v[5] = raw16_hi<IFMT>(u.z); v[6] = raw16_lo<IFMT>(u.w); v[7] = raw16_hi<IFMT>(u.w);
            } else {
                const float4 a = *reinterpret_cast<const float4*>(x + off);
                const float4 b = *reinterpret_cast<const float4*>(x + off + 4);
                v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                float t = fmaf(v[e], sc[e], sh[e]);
                if (silu) {
                    if constexpr (FAST) t = __fdividef(t, 1.0f + __expf(-t));
                    else t = t / (1.0f + expf(-t));
                }
                v[e] = t;
            }
            if constexpr (OFMT != FMT_F32) {
                *reinterpret_cast<uint4*>(static_cast<bf16*>(yv) + off) =
                    make_uint4(pack16x2<OFMT>(v[0], v[1]), pack16x2<OFMT>(v[2], v[3]), pack16x2<OFMT>(v[4], v[5]),
                               pack16x2<OFMT>(v[6], v[7]));
            } else {
                float* y = static_cast<float*>(yv);
                *reinterpret_cast<float4*>(y + off) = make_float4(v[0], v[1], v[2], v[3]);
                *reinterpret_cast<float4*>(y + off + 4) = make_float4(v[4], v[5], v[6], v[7]);
            }
        }
    }
}

int launch_gn_apply(const void* x, int x_fmt, void* y, int y_fmt, const double* stats, const float* gamma,
                    const float* beta, int N, long long HW, int C, int G, float eps, int silu, cudaStream_t s,
                    Profiler* prof) {
    VT_CHECK(C % 8 == 0 && C % G == 0, "GroupNorm apply needs C % 8 == 0 and C % groups == 0");
    const int tpb = std::min(C / 8, 256), ppb = 256 / tpb;
    // ~16 pixel steps per block at least, and enough blocks to fill the machine a few times
    long long want = (HW + 1LL * ppb * 16 - 1) / (1LL * ppb * 16);
    const long long cap = std::max(1, sm_count() * 16 / N);
    const int chunks = static_cast<int>(std::max<long long>(1, std::min(want, cap)));
    dim3 grid(chunks, N);
    profiler_begin(prof, KC_GN_APPLY, s, 0, 1.0 * N * HW * C * ((x_fmt == FMT_F32 ? 4 : 2) + (y_fmt == FMT_F32 ? 4 : 2)));
#define VT_GN(IF, OF, FAST) \
    gn_apply_kernel<IF, OF, FAST><<<grid, 256, 0, s>>>(x, y, stats, gamma, beta, HW, C, G, eps, silu)
    const bool x_fp32 = x_fmt == FMT_F32;
    if (x_fp32 && y_fmt == FMT_F32) VT_GN(FMT_F32, FMT_F32, false);
    else if (x_fp32 && y_fmt == FMT_F16) VT_GN(FMT_F32, FMT_F16, true);
    else if (x_fp32 && y_fmt == FMT_BF16) VT_GN(FMT_F32, FMT_BF16, true);
    else if (x_fmt == FMT_BF16 && y_fmt == FMT_F16) VT_GN(FMT_BF16, FMT_F16, true);
    else if (x_fmt == FMT_BF16 && y_fmt == FMT_BF16) VT_GN(FMT_BF16, FMT_BF16, true);
    else if (x_fmt == FMT_F16 && y_fmt == FMT_F16) VT_GN(FMT_F16, FMT_F16, true);
    else if (x_fmt == FMT_F16 && y_fmt == FMT_BF16) VT_GN(FMT_F16, FMT_BF16, true);   // backward operand (vt_wgrad.cu)
    else {
        set_error("GroupNorm apply: this input / output format pair is not instantiated");
        return -2;
    }
#undef VT_GN
    profiler_end(prof, KC_GN_APPLY, s);
    VT_CUDA(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------
// Attention row softmax: p = softmax(s) along the last dimension (scores already scaled by
// the GEMM epilogue).  One 256-thread CTA per row; the row (<= 16384 values) lives in
// registers, so it is read once and written once.  Wider rows take the three-pass loop.
struct half_out { __half v; };  // distinct 2-byte element types for the probability output
template <typename TO>
__device__ __forceinline__ void store_prob(TO* p, float v) {
    if constexpr (std::is_same<TO, half_out>::value) p->v = __float2half_rn(v);
    else if constexpr (sizeof(TO) == 2) *p = __float2bfloat16(v);
    else *p = v;
}

template <typename TO, int PER>  // PER values per thread, cols <= 256*PER
__global__ void __launch_bounds__(256) softmax_rows_kernel(const float* __restrict__ s, TO* __restrict__ p,
                                                           int cols, long long ld_s, long long ld_p, int cols_pad) {
    __shared__ float red[8];
    __shared__ float bcast;
    const long long row = blockIdx.x;
    const float* sr = s + row * ld_s;
    TO* pr = p + row * ld_p;
    float v[PER];
    float m = -INFINITY;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
        const int c = i * 256 + threadIdx.x;
        v[i] = c < cols ? sr[c] : -INFINITY;
        m = fmaxf(m, v[i]);
    }
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x < 32) {
        float t = threadIdx.x < 8 ? red[threadIdx.x] : -INFINITY;
        t = warp_max(t);
        if (threadIdx.x == 0) bcast = t;
    }
    __syncthreads();
    m = bcast;
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
        v[i] = expf(v[i] - m);  // exp(-inf) = 0 for the padding
        sum += v[i];
    }
    sum = warp_sum(sum);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sum;
    __syncthreads();
    if (threadIdx.x < 32) {
        float t = threadIdx.x < 8 ? red[threadIdx.x] : 0.f;
        t = warp_sum(t);
        if (threadIdx.x == 0) bcast = t;
    }
    __syncthreads();
    const float inv = 1.0f / bcast;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
        const int c = i * 256 + threadIdx.x;
        if (c < cols) store_prob(pr + c, v[i] * inv);
        else if (c < cols_pad) store_prob(pr + c, 0.f);  // K padding of the following P.V contraction
    }
}

template <typename TO>
__global__ void __launch_bounds__(256) softmax_rows_wide_kernel(const float* __restrict__ s, TO* __restrict__ p,
                                                                int cols, long long ld_s, long long ld_p,
                                                                int cols_pad) {
    __shared__ float red[8];
    __shared__ float bcast;
    const long long row = blockIdx.x;
    const float* sr = s + row * ld_s;
    TO* pr = p + row * ld_p;
    float m = -INFINITY;
    for (int c = threadIdx.x; c < cols; c += 256) m = fmaxf(m, sr[c]);
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = red[0];
        for (int i = 1; i < 8; ++i) t = fmaxf(t, red[i]);
        bcast = t;
    }
    __syncthreads();
    m = bcast;
    float sum = 0.f;
    for (int c = threadIdx.x; c < cols; c += 256) sum += expf(sr[c] - m);
    sum = warp_sum(sum);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int i = 0; i < 8; ++i) t += red[i];
        bcast = t;
    }
    __syncthreads();
    const float inv = 1.0f / bcast;
    for (int c = threadIdx.x; c < cols; c += 256) store_prob(pr + c, expf(sr[c] - m) * inv);
    for (int c = cols + threadIdx.x; c < cols_pad; c += 256) store_prob(pr + c, 0.f);
}

template <typename TO>
static void softmax_dispatch(const float* s, TO* p, long long rows, int cols, long long ld_s, long long ld_p,
                             cudaStream_t st) {
    const unsigned grid = static_cast<unsigned>(rows);
    // probabilities beyond `cols` up to the next multiple of 64 (within the row pitch) are written as zeros
    const int cp = static_cast<int>(std::min<long long>((cols + 63) / 64 * 64, ld_p));
    if (cols <= 256 * 4) softmax_rows_kernel<TO, 4><<<grid, 256, 0, st>>>(s, p, cols, ld_s, ld_p, cp);
    else if (cols <= 256 * 16) softmax_rows_kernel<TO, 16><<<grid, 256, 0, st>>>(s, p, cols, ld_s, ld_p, cp);
    else if (cols <= 256 * 64) softmax_rows_kernel<TO, 64><<<grid, 256, 0, st>>>(s, p, cols, ld_s, ld_p, cp);
    else softmax_rows_wide_kernel<TO><<<grid, 256, 0, st>>>(s, p, cols, ld_s, ld_p, cp);
}

int launch_softmax_rows(const float* s, void* p, int p_fmt, long long rows, int cols, long long ld_s,
                        long long ld_p, cudaStream_t st, Profiler* prof) {
    VT_CHECK(rows > 0 && rows < (1LL << 31) && cols > 0, "softmax shape");
    profiler_begin(prof, KC_SOFTMAX, st, 0, 1.0 * rows * cols * (4 + (p_fmt == FMT_F32 ? 4 : 2)));
    if (p_fmt == FMT_F32) softmax_dispatch<float>(s, static_cast<float*>(p), rows, cols, ld_s, ld_p, st);
    else if (p_fmt == FMT_F16) softmax_dispatch<half_out>(s, static_cast<half_out*>(p), rows, cols, ld_s, ld_p, st);
    else softmax_dispatch<bf16>(s, static_cast<bf16*>(p), rows, cols, ld_s, ld_p, st);
    profiler_end(prof, KC_SOFTMAX, st);
    VT_CUDA(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------
// moments (conv_out, fp32 NHWC [N,h,w,32]) -> DiagonalGaussian outputs, NCHW fp32:
//   mean, logvar = clamp(., -30, 20), latent = (mode|sample) * scale + shift
// sample = mean + exp(0.5 logvar) * eps with eps from a caller tensor (exact parity with a
// host-supplied noise) or from a counter-based generator (splitmix64 + Box-Muller).
__device__ __forceinline__ unsigned long long splitmix64(unsigned long long z) {
    z += 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
__device__ __forceinline__ float counter_normal(unsigned long long seed, unsigned long long idx) {
    const unsigned long long r = splitmix64(seed ^ splitmix64(idx));
    const float u1 = (static_cast<float>(r >> 40) + 1.0f) * (1.0f / 16777216.0f);        // (0,1]
    const float u2 = static_cast<float>((r >> 16) & 0xFFFFFF) * (1.0f / 16777216.0f);    // [0,1)
    return sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
}

__global__ void __launch_bounds__(256) moments_to_latent_kernel(const float* __restrict__ mom,
                                                                float* __restrict__ latent,
                                                                float* __restrict__ mean_out,
                                                                float* __restrict__ logvar_out,
                                                                const float* __restrict__ noise, int N, int HW,
                                                                int LC, int sample, unsigned long long seed,
                                                                float scale, float shift, int apply_scale,
                                                                int apply_shift) {
    // one thread per (n, c, pixel) of the NCHW output: writes coalesced along pixels; the NHWC
    // reads of a warp hit 32 different pixels of one channel (128-byte stride) -- the tensor is
    // tiny (128 KB .. 2 MB per image) and L2 resident.
    const long long total = 1LL * N * LC * HW;
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += 256LL * gridDim.x) {
        const int p = static_cast<int>(i % HW);
        const int c = static_cast<int>((i / HW) % LC);
        const int n = static_cast<int>(i / (1LL * HW * LC));
        const float* m = mom + (1LL * n * HW + p) * (2 * LC);
        const float mean = m[c];
        const float lv = fminf(fmaxf(m[LC + c], -30.0f), 20.0f);
        float z = mean;
        if (sample) {
            const float e = noise ? noise[i] : counter_normal(seed, static_cast<unsigned long long>(i));
            z = mean + expf(0.5f * lv) * e;
        }
        if (apply_scale) z = z * scale;
        if (apply_shift) z = z + shift;
        if (latent) latent[i] = z;
        if (mean_out) mean_out[i] = mean;
        if (logvar_out) logvar_out[i] = lv;
    }
}

int launch_moments_to_latent(const float* moments_nhwc, float* latent, float* mean_out, float* logvar_out,
                             const float* noise, int N, int H, int W, int LC, int sample, unsigned long long seed,
                             float scale, float shift, int apply_scale, int apply_shift, cudaStream_t s,
                             Profiler* prof) {
    const long long total = 1LL * N * LC * H * W;
    const int grid = static_cast<int>(std::min<long long>((total + 255) / 256, 1LL * sm_count() * 8));
    profiler_begin(prof, KC_LATENT, s, 0, total * 12.0);
    moments_to_latent_kernel<<<grid, 256, 0, s>>>(moments_nhwc, latent, mean_out, logvar_out, noise, N, H * W, LC,
                                                  sample, seed, scale, shift, apply_scale, apply_shift);
    profiler_end(prof, KC_LATENT, s);
    VT_CUDA(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------
// Layout conversions (tests, op-level entry points, fp32 path I/O): tiled transpose through
// shared memory so both sides are coalesced.
template <int OFMT>
__global__ void __launch_bounds__(256) nchw_to_nhwc_kernel(const float* __restrict__ in, void* __restrict__ outv,
                                                           int C, long long HW) {
    __shared__ float tile[32][33];
    const int n = blockIdx.z;
    const long long p0 = blockIdx.x * 32LL;
    const int c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int r = ty; r < 32; r += 8) {
        const int c = c0 + r;
        const long long p = p0 + tx;
        tile[r][tx] = (c < C && p < HW) ? in[(1LL * n * C + c) * HW + p] : 0.f;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const long long p = p0 + r;
        const int c = c0 + tx;
        if (c < C && p < HW) {
            const long long o = (1LL * n * HW + p) * C + c;
            if constexpr (OFMT == FMT_BF16) static_cast<bf16*>(outv)[o] = __float2bfloat16(tile[tx][r]);
            else if constexpr (OFMT == FMT_F16) static_cast<__half*>(outv)[o] = __float2half_rn(tile[tx][r]);
            else static_cast<float*>(outv)[o] = tile[tx][r];
        }
    }
}
template <int IFMT>
__global__ void __launch_bounds__(256) nhwc_to_nchw_kernel(const void* __restrict__ inv, float* __restrict__ out,
                                                           int C, long long HW) {
    __shared__ float tile[32][33];
    const int n = blockIdx.z;
    const long long p0 = blockIdx.x * 32LL;
    const int c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int r = ty; r < 32; r += 8) {
        const long long p = p0 + r;
        const int c = c0 + tx;
        float v = 0.f;
        if (c < C && p < HW) {
            const long long o = (1LL * n * HW + p) * C + c;
            if constexpr (IFMT == FMT_BF16) v = __bfloat162float(static_cast<const bf16*>(inv)[o]);
            else if constexpr (IFMT == FMT_F16) v = __half2float(static_cast<const __half*>(inv)[o]);
            else v = static_cast<const float*>(inv)[o];
        }
        tile[r][tx] = v;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int c = c0 + r;
        const long long p = p0 + tx;
        if (c < C && p < HW) out[(1LL * n * C + c) * HW + p] = tile[tx][r];
    }
}

int launch_nchw_to_nhwc(const float* in, void* out, int out_fmt, int N, int C, long long HW, cudaStream_t s) {
    dim3 grid(static_cast<unsigned>((HW + 31) / 32), (C + 31) / 32, N);
    if (out_fmt == FMT_F32) nchw_to_nhwc_kernel<FMT_F32><<<grid, 256, 0, s>>>(in, out, C, HW);
    else if (out_fmt == FMT_F16) nchw_to_nhwc_kernel<FMT_F16><<<grid, 256, 0, s>>>(in, out, C, HW);
    else nchw_to_nhwc_kernel<FMT_BF16><<<grid, 256, 0, s>>>(in, out, C, HW);
    VT_CUDA(cudaGetLastError());
    return 0;
}
int launch_nhwc_to_nchw(const void* in, int in_fmt, float* out, int N, int C, long long HW, cudaStream_t s) {
    dim3 grid(static_cast<unsigned>((HW + 31) / 32), (C + 31) / 32, N);
    if (in_fmt == FMT_F32) nhwc_to_nchw_kernel<FMT_F32><<<grid, 256, 0, s>>>(in, out, C, HW);
    else if (in_fmt == FMT_F16) nhwc_to_nchw_kernel<FMT_F16><<<grid, 256, 0, s>>>(in, out, C, HW);
    else nhwc_to_nchw_kernel<FMT_BF16><<<grid, 256, 0, s>>>(in, out, C, HW);
    VT_CUDA(cudaGetLastError());
    return 0;
}

// ---- VAE decoder I/O
// latent NCHW fp32 [N][LC][HW] -> NHWC [N][HW][CP] (CP = LC padded to 64: one K chunk of the implicit GEMM),
// with DiffusersVAEWrapper.decode's un-shift / un-scale (diffusers_vae_loader.py:88-93) fused in.
template <int OFMT>
__global__ void __launch_bounds__(256) latent_to_nhwc_kernel(const float* __restrict__ z, void* __restrict__ outv,
                                                             int LC, int CP, long long HW, float shift,
                                                             float inv_scale, long long total) {
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += 256LL * gridDim.x) {
        const int c = static_cast<int>(i % CP);
        const long long np = i / CP;
        const long long n = np / HW, p = np - n * HW;
        const float v = c < LC ? (z[(n * LC + c) * HW + p] - shift) * inv_scale : 0.f;
        if constexpr (OFMT == FMT_BF16) static_cast<bf16*>(outv)[i] = __float2bfloat16(v);
        else if constexpr (OFMT == FMT_F16) static_cast<__half*>(outv)[i] = __float2half_rn(fminf(fmaxf(v, -65504.f), 65504.f));
        else static_cast<float*>(outv)[i] = v;
    }
}
int launch_latent_to_nhwc(const float* z, void* out, int out_fmt, int N, int LC, int CP, long long HW, float shift,
                          float inv_scale, cudaStream_t s, Profiler* prof) {
    const long long total = 1LL * N * HW * CP;
    const int grid = static_cast<int>(std::min<long long>((total + 255) / 256, 148 * 16));
    profiler_begin(prof, KC_LATENT, s, 0, 4.0 * N * LC * HW + (out_fmt == FMT_F32 ? 4.0 : 2.0) * total);
    if (out_fmt == FMT_F32) latent_to_nhwc_kernel<FMT_F32><<<grid, 256, 0, s>>>(z, out, LC, CP, HW, shift, inv_scale, total);
    else if (out_fmt == FMT_F16) latent_to_nhwc_kernel<FMT_F16><<<grid, 256, 0, s>>>(z, out, LC, CP, HW, shift, inv_scale, total);
    else latent_to_nhwc_kernel<FMT_BF16><<<grid, 256, 0, s>>>(z, out, LC, CP, HW, shift, inv_scale, total);
    profiler_end(prof, KC_LATENT, s);
    VT_CUDA(cudaGetLastError());
    return 0;
}

// Upsample2D's F.interpolate(scale_factor=2, mode="nearest") on NHWC: out[n][2y+a][2x+b][:] = in[n][y][x][:].
// One thread moves 16 bytes of a source pixel to its four destinations: reads once, writes coalesced.
__global__ void __launch_bounds__(256) upsample2x_nhwc_kernel(const uint4* __restrict__ in, uint4* __restrict__ out,
                                                              int H, int W, int vec_per_px, long long total) {
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += 256LL * gridDim.x) {
        const int v = static_cast<int>(i % vec_per_px);
        const long long px = i / vec_per_px;
        const int x = static_cast<int>(px % W);
        const long long ny = px / W;
        const int y = static_cast<int>(ny % H);
        const long long n = ny / H;
        const uint4 d = in[i];
        const long long o = ((n * 2 * H + 2 * y) * 2 * W + 2 * x) * vec_per_px + v;
        const long long row = 2LL * W * vec_per_px;
        out[o] = d;
        out[o + vec_per_px] = d;
        out[o + row] = d;
        out[o + row + vec_per_px] = d;
    }
}
int launch_upsample2x_nhwc(const void* in, void* out, int elem_bytes, int N, int H, int W, int C, cudaStream_t s,
                           Profiler* prof) {
    VT_CHECK((1LL * C * elem_bytes) % 16 == 0, "upsample: a pixel must be a multiple of 16 bytes");
    const int vpp = C * elem_bytes / 16;
    const long long total = 1LL * N * H * W * vpp;
    const int grid = static_cast<int>(std::min<long long>((total + 255) / 256, 148 * 32));
    profiler_begin(prof, KC_MISC, s, 0, 5.0 * 16.0 * total);
    upsample2x_nhwc_kernel<<<grid, 256, 0, s>>>(static_cast<const uint4*>(in), static_cast<uint4*>(out), H, W, vpp, total);
    profiler_end(prof, KC_MISC, s);
    VT_CUDA(cudaGetLastError());
    return 0;
}

// conv_out result fp32 NHWC [N][HW][CP] (CP = out_channels padded to 32) -> image NCHW fp32 [N][OC][HW]
__global__ void __launch_bounds__(256) nhwc_to_image_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                            int OC, int CP, long long HW, long long total) {
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += 256LL * gridDim.x) {
        const long long p = i % HW;
        const long long nc = i / HW;
        const int c = static_cast<int>(nc % OC);
        const long long n = nc / OC;
        out[i] = in[(n * HW + p) * CP + c];
    }
}
int launch_nhwc_to_image(const float* in, float* out, int N, int OC, int CP, long long HW, cudaStream_t s,
                         Profiler* prof) {
    const long long total = 1LL * N * OC * HW;
    const int grid = static_cast<int>(std::min<long long>((total + 255) / 256, 148 * 32));
    profiler_begin(prof, KC_LATENT, s, 0, 4.0 * N * HW * (CP + OC));
    nhwc_to_image_kernel<<<grid, 256, 0, s>>>(in, out, OC, CP, HW, total);
    profiler_end(prof, KC_LATENT, s);
    VT_CUDA(cudaGetLastError());
    return 0;
}

// Split-K combine: out[r][c] = bias[c] + sum_s part[s][r][c]  (part fp32, out 16-bit; 4 columns / thread)
template <int OFMT>
__global__ void __launch_bounds__(256) splitk_reduce_kernel(const float* __restrict__ part, int splits,
                                                            long long split_stride, const float* __restrict__ bias,
                                                            bf16* __restrict__ out, long long rows, int cols,
                                                            long long ld_out) {
    const long long total = rows * (cols / 4);
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += 256LL * gridDim.x) {
        const long long r = i / (cols / 4);
        const int c = static_cast<int>(i % (cols / 4)) * 4;
        float4 a = bias ? __ldg(reinterpret_cast<const float4*>(bias + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
        for (int s = 0; s < splits; ++s) {
            const float4 v = *reinterpret_cast<const float4*>(part + s * split_stride + r * cols + c);
            a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
        }
        *reinterpret_cast<uint2*>(out + r * ld_out + c) = make_uint2(pack16x2<OFMT>(a.x, a.y), pack16x2<OFMT>(a.z, a.w));
    }
}
int launch_splitk_reduce(const float* part, int splits, long long split_stride, const float* bias, void* out,
                         int out_fmt, long long rows, int cols, long long ld_out, cudaStream_t s, Profiler* prof) {
    VT_CHECK(cols % 4 == 0 && out_fmt != FMT_F32, "split-K combine needs a column count divisible by 4 and a 16-bit output");
    const long long total = rows * (cols / 4);
    const int grid = static_cast<int>(std::min<long long>((total + 255) / 256, 1LL * sm_count() * 8));
    profiler_begin(prof, KC_MISC, s, 0, 4.0 * splits * rows * cols + 2.0 * rows * cols);
    if (out_fmt == FMT_F16)
        splitk_reduce_kernel<FMT_F16><<<grid, 256, 0, s>>>(part, splits, split_stride, bias, static_cast<bf16*>(out), rows, cols, ld_out);
    else
        splitk_reduce_kernel<FMT_BF16><<<grid, 256, 0, s>>>(part, splits, split_stride, bias, static_cast<bf16*>(out), rows, cols, ld_out);
    profiler_end(prof, KC_MISC, s);
    VT_CUDA(cudaGetLastError());
    return 0;
}

// fp32 -> 16-bit cast (weights / test operands)
template <int OFMT>
__global__ void __launch_bounds__(256) cast_f32_16_kernel(const float* __restrict__ in, void* __restrict__ out,
                                                          long long n) {
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += 256LL * gridDim.x) {
        if constexpr (OFMT == FMT_F16) static_cast<__half*>(out)[i] = __float2half_rn(fminf(fmaxf(in[i], -65504.f), 65504.f));
        else static_cast<bf16*>(out)[i] = __float2bfloat16(in[i]);
    }
}
int launch_cast_f32_16(const float* in, void* out, int out_fmt, long long n, cudaStream_t s) {
    const int grid = static_cast<int>(std::min<long long>((n + 255) / 256, 1LL * sm_count() * 8));
    if (out_fmt == FMT_F16) cast_f32_16_kernel<FMT_F16><<<grid, 256, 0, s>>>(in, out, n);
    else cast_f32_16_kernel<FMT_BF16><<<grid, 256, 0, s>>>(in, out, n);
    VT_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace vt
