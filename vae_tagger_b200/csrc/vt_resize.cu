// GPU image preprocessing in front of conv_in: crop + resize of uint8 HWC images, bit-exact with what
// the reference does on the host through Pillow.
//
// Replaces SmartResize.__call__ (modules.py:142-178: crop box, then img.resize((W,H), Image.LANCZOS)) and
// transforms.Resize((res,res)) on a PIL image (modules.py:135: PIL BILINEAR).  The arithmetic is not in the
// reference but in its dependency Pillow (requirements.txt: Pillow>=9.0, unpinned; 12.2.0 in this image),
// src/libImaging/Resample.c, restated here from its published algorithm:
//   * per output pixel a window [xmin, xmin+n) of the input, n <= ksize = 2*ceil(support*max(scale,1))+1,
//     filter taps evaluated in double precision at (x + xmin - center + 0.5)/max(scale,1), normalised to
//     sum 1, then rounded to fixed point with 22 fractional bits (precompute_coeffs, normalize_coeffs_8bpc);
//   * horizontal pass over the rows, then vertical pass, each  out = clip8((2^21 + sum k*in) >> 22)  in
//     int32, with a uint8 intermediate image; a pass is skipped when that dimension does not change;
//   * PIL/Image.py Image.resize (12.2): an image more than 100 times taller than wide that shrinks vertically
//     gets its vertical pass first.
// The coefficient tables are computed on the host (they depend only on (in size, out size, filter)) and
// cached on the device; the two passes are byte-streaming kernels (no tensor cores).  The horizontal pass runs
// on dp4a over three signed base-256 digits of the taps (VT_B200_RESIZE_SCALAR=1: the IMAD form).
#include <math.h>

#include <map>
#include <tuple>
#include <vector>

#include "../../include/vae_tagger_b200.h"
#include "vt_internal.h"
#include "vt_resize.h"

namespace vt {

namespace {

constexpr int kPrecisionBits = 32 - 8 - 2;
constexpr int kHBlock = 128;  // output pixels per CTA of the horizontal pass

double sinc(double x) {
    if (x == 0.0) return 1.0;
    x = x * M_PI;
    return sin(x) / x;
}
double filter_value(int filter, double x) {
    if (filter == VT_FILTER_LANCZOS) return (-3.0 <= x && x < 3.0) ? sinc(x) * sinc(x / 3) : 0.0;
    if (x < 0.0) x = -x;
    return x < 1.0 ? 1.0 - x : 0.0;
}

__device__ __forceinline__ unsigned char clip8(int acc) {
    acc >>= kPrecisionBits;
    return static_cast<unsigned char>(min(max(acc, 0), 255));
}

// Horizontal pass.  grid (ceil(out_w/128), ceil(rows/rows_per_cta)); a CTA stages the span of source
// pixels its 128 output pixels need in shared memory (coalesced), R rows at a time: a thread loads each of its
// taps once and applies it to R rows (R independent accumulation chains hide the load latency).
//   src: rows of 3-byte pixels, already offset to the crop origin; dst: [rows][out_w][3], row pitch dst_pitch
//   kk_t: [ksize][out_w] (transposed: coalesced across the threads of a CTA), bounds: [out_w] (xmin, n)
template <int R>
__global__ void __launch_bounds__(kHBlock) resize_h_kernel(const unsigned char* __restrict__ src, long long src_pitch,
                                                           unsigned char* __restrict__ dst, long long dst_pitch,
                                                           const int* __restrict__ kk_t,
                                                           const int2* __restrict__ bounds, int out_w, int rows,
                                                           int rows_per_cta, int span_pitch) {
    extern __shared__ unsigned char span[];
    const int x0 = blockIdx.x * kHBlock;
    const int xx = x0 + threadIdx.x;
    const bool live = xx < out_w;
    const int2 b = live ? bounds[xx] : make_int2(0, 0);
    const int first = bounds[x0].x;
    const int xl = min(x0 + kHBlock, out_w) - 1;
    const int last = bounds[xl].x + bounds[xl].y;  // windows start and end monotonically
    const int nbytes = (last - first) * 3;
    const int r0 = blockIdx.y * rows_per_cta, r1 = min(rows, r0 + rows_per_cta);
    for (int r = r0; r < r1; r += R) {
        const int nr = min(R, r1 - r);
        __syncthreads();
        for (int q = 0; q < nr; ++q) {
            const unsigned char* row = src + (r + q) * src_pitch + 3LL * first;
            for (int i = threadIdx.x; i < nbytes; i += kHBlock) span[q * span_pitch + i] = row[i];
        }
        __syncthreads();
        if (live) {
            int acc[R][3];
#pragma unroll
            for (int q = 0; q < R; ++q) acc[q][0] = acc[q][1] = acc[q][2] = 1 << (kPrecisionBits - 1);
            const unsigned char* p = span + 3 * (b.x - first);
            for (int k = 0; k < b.y; ++k) {
                const int c = kk_t[1LL * k * out_w + xx];
#pragma unroll
                for (int q = 0; q < R; ++q) {
                    if (q < nr) {
                        acc[q][0] += c * p[q * span_pitch + 3 * k];
                        acc[q][1] += c * p[q * span_pitch + 3 * k + 1];
                        acc[q][2] += c * p[q * span_pitch + 3 * k + 2];
                    }
                }
            }
#pragma unroll
            for (int q = 0; q < R; ++q) {
                if (q < nr) {
                    unsigned char* o = dst + (r + q) * dst_pitch + 3LL * xx;
                    o[0] = clip8(acc[q][0]);
                    o[1] = clip8(acc[q][1]);
                    o[2] = clip8(acc[q][2]);
                }
            }
        }
    }
}

__device__ __forceinline__ int dp4a_u8s8(unsigned a, unsigned b, int c) {
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

// Horizontal pass on dp4a.  Same tiling as resize_h_kernel, but the staged span is de-interleaved into three byte
// planes and the taps come as three signed base-256 digit words per group of four taps:
//   sum_k c_k p_k = 65536 * dp4a(p, c2) + 256 * dp4a(p, c1) + dp4a(p, c0)      (exact: int32 wrap-around arithmetic)
// A thread's window starts at an arbitrary pixel, so its four-pixel words are cut out of two aligned shared-memory
// words with a funnel shift.  Per four taps, channel and row: 1 LDS.32 + 1 SHF + 3 DP4A instead of 4 LDS.U8 + 4 IMAD.
template <int R>
__global__ void __launch_bounds__(kHBlock) resize_h_dp4a_kernel(const unsigned char* __restrict__ src, long long src_pitch,
                                                                unsigned char* __restrict__ dst, long long dst_pitch,
                                                                const unsigned* __restrict__ kd,
                                                                const int2* __restrict__ bounds, int out_w, int rows,
                                                                int rows_per_cta, int plane_pitch) {
    extern __shared__ __align__(16) unsigned char planes[];  // [R][3][plane_pitch]
    const int x0 = blockIdx.x * kHBlock;
    const int xx = x0 + threadIdx.x;
    const bool live = xx < out_w;
    const int2 b = live ? bounds[xx] : make_int2(0, 0);
    const int first = bounds[x0].x;
    const int xl = min(x0 + kHBlock, out_w) - 1;
    const int npx = bounds[xl].x + bounds[xl].y - first;
    const int r0 = blockIdx.y * rows_per_cta, r1 = min(rows, r0 + rows_per_cta);
    const int rel = b.x - first;
    const int sh = (rel & 3) * 8;
    const int ng = (b.y + 3) >> 2;
    for (int r = r0; r < r1; r += R) {
        const int nr = min(R, r1 - r);
        __syncthreads();
        for (int q = 0; q < nr; ++q) {
            const unsigned char* row = src + (r + q) * src_pitch + 3LL * first;
            unsigned char* pl = planes + q * 3 * plane_pitch;
            for (int i = threadIdx.x; i < npx; i += kHBlock) {
                pl[i] = row[3 * i];
                pl[plane_pitch + i] = row[3 * i + 1];
                pl[2 * plane_pitch + i] = row[3 * i + 2];
            }
        }
        __syncthreads();
        if (live) {
            int S[R][3][3];
            unsigned w0[R][3];
#pragma unroll
            for (int q = 0; q < R; ++q)
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    S[q][c][0] = S[q][c][1] = S[q][c][2] = 0;
                    w0[q][c] = *reinterpret_cast<const unsigned*>(planes + (q * 3 + c) * plane_pitch + (rel & ~3));
                }
            for (int g = 0; g < ng; ++g) {
                const unsigned k0 = kd[(1LL * g * 3 + 0) * out_w + xx];
                const unsigned k1 = kd[(1LL * g * 3 + 1) * out_w + xx];
                const unsigned k2 = kd[(1LL * g * 3 + 2) * out_w + xx];
#pragma unroll
                for (int q = 0; q < R; ++q) {
                    if (q < nr) {
#pragma unroll
                        for (int c = 0; c < 3; ++c) {
                            const unsigned w1 = *reinterpret_cast<const unsigned*>(planes + (q * 3 + c) * plane_pitch +
                                                                                   (rel & ~3) + 4 * (g + 1));
                            const unsigned px = __funnelshift_r(w0[q][c], w1, sh);
                            w0[q][c] = w1;
                            S[q][c][0] = dp4a_u8s8(px, k0, S[q][c][0]);
                            S[q][c][1] = dp4a_u8s8(px, k1, S[q][c][1]);
                            S[q][c][2] = dp4a_u8s8(px, k2, S[q][c][2]);
                        }
                    }
                }
            }
#pragma unroll
            for (int q = 0; q < R; ++q) {
                if (q < nr) {
                    unsigned char* o = dst + (r + q) * dst_pitch + 3LL * xx;
#pragma unroll
                    for (int c = 0; c < 3; ++c)
                        o[c] = clip8((1 << (kPrecisionBits - 1)) + S[q][c][2] * 65536 + S[q][c][1] * 256 + S[q][c][0]);
                }
            }
        }
    }
}

// Vertical pass.  One thread per VEC bytes of an output row (a row is out_w*3 bytes: channels and pixels
// are interchangeable here), grid (ceil(row_bytes/VEC/256), out_h).  kk: [out_h][ksize], bounds [out_h].
template <int VEC>
__global__ void __launch_bounds__(256) resize_v_kernel(const unsigned char* __restrict__ src, long long src_pitch,
                                                       unsigned char* __restrict__ dst, long long dst_pitch,
                                                       const int* __restrict__ kk, const int2* __restrict__ bounds,
                                                       int ksize, int row_bytes) {
    const int yy = blockIdx.y;
    const int j = (blockIdx.x * 256 + threadIdx.x) * VEC;
    if (j >= row_bytes) return;
    const int2 b = bounds[yy];
    const int* k = kk + 1LL * yy * ksize;
    int acc[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) acc[v] = 1 << (kPrecisionBits - 1);
    const unsigned char* p = src + b.x * src_pitch + j;
    for (int t = 0; t < b.y; ++t) {
        const int c = k[t];
        if (VEC == 4) {
            const uchar4 u = *reinterpret_cast<const uchar4*>(p);
            acc[0] += c * u.x;
            acc[1 % VEC] += c * u.y;
            acc[2 % VEC] += c * u.z;
            acc[3 % VEC] += c * u.w;
        } else {
            acc[0] += c * p[0];
        }
        p += src_pitch;
    }
    unsigned char* o = dst + yy * dst_pitch + j;
    if (VEC == 4) {
        *reinterpret_cast<uchar4*>(o) = make_uchar4(clip8(acc[0]), clip8(acc[1 % VEC]), clip8(acc[2 % VEC]),
                                                    clip8(acc[3 % VEC]));
    } else {
        o[0] = clip8(acc[0]);
    }
}

// plain crop copy when neither dimension changes (Pillow returns a copy of the cropped image)
__global__ void __launch_bounds__(256) copy_rows_kernel(const unsigned char* __restrict__ src, long long src_pitch,
                                                        unsigned char* __restrict__ dst, long long dst_pitch,
                                                        int row_bytes) {
    const int j = blockIdx.x * 256 + threadIdx.x;
    if (j < row_bytes) dst[blockIdx.y * dst_pitch + j] = src[blockIdx.y * src_pitch + j];
}

}  // namespace

struct ResizeTable {
    int ksize = 0, out_size = 0, max_span = 0;
    int* kk = nullptr;      // [out][ksize]
    int* kk_t = nullptr;    // [ksize][out]
    unsigned* kd = nullptr; // [groups][3][out]: taps 4g..4g+3 as signed base-256 digits (digit d of four taps per word)
    int groups = 0;         // ceil(ksize / 4)
    int2* bounds = nullptr; // [out]
    unsigned long long stamp = 0;
};

struct ResizeCache {
    std::map<std::tuple<int, int, int>, ResizeTable> tables;
    unsigned long long clock = 0;
    void* tmp = nullptr;
    size_t tmp_cap = 0;
};

ResizeCache* resize_cache_create() { return new ResizeCache(); }
void resize_cache_destroy(ResizeCache* c) {
    if (!c) return;
    for (auto& kv : c->tables) {
        cudaFree(kv.second.kk);
        cudaFree(kv.second.kk_t);
        cudaFree(kv.second.kd);
        cudaFree(kv.second.bounds);
    }
    if (c->tmp) cudaFree(c->tmp);
    delete c;
}

// Resample.c precompute_coeffs + normalize_coeffs_8bpc for the box (0, in_size)
void resize_coefficients(int in_size, int out_size, int filter, int* ksize_out, std::vector<int>& bounds,
                         std::vector<int>& kk) {
    const double fsupport = filter == VT_FILTER_LANCZOS ? 3.0 : 1.0;
    double scale = static_cast<double>(in_size) / out_size;
    double filterscale = scale;
    if (filterscale < 1.0) filterscale = 1.0;
    const double support = fsupport * filterscale;
    const int ksize = static_cast<int>(ceil(support)) * 2 + 1;
    bounds.assign(static_cast<size_t>(out_size) * 2, 0);
    kk.assign(static_cast<size_t>(out_size) * ksize, 0);
    std::vector<double> w(ksize);
    const double ss = 1.0 / filterscale;
    for (int xx = 0; xx < out_size; ++xx) {
        const double center = (xx + 0.5) * scale;
        int xmin = static_cast<int>(center - support + 0.5);
        if (xmin < 0) xmin = 0;
        int xmax = static_cast<int>(center + support + 0.5);
        if (xmax > in_size) xmax = in_size;
        xmax -= xmin;
        double ww = 0.0;
        for (int x = 0; x < xmax; ++x) {
            w[x] = filter_value(filter, (x + xmin - center + 0.5) * ss);
            ww += w[x];
        }
        for (int x = 0; x < xmax; ++x) {
            const double v = ww != 0.0 ? w[x] / ww : w[x];
            kk[static_cast<size_t>(xx) * ksize + x] = v < 0 ? static_cast<int>(-0.5 + v * (1 << kPrecisionBits))
                                                            : static_cast<int>(0.5 + v * (1 << kPrecisionBits));
        }
        bounds[2 * xx] = xmin;
        bounds[2 * xx + 1] = xmax;
    }
    *ksize_out = ksize;
}

static int get_table(ResizeCache* c, int in_size, int out_size, int filter, const ResizeTable** out) {
    const auto key = std::make_tuple(in_size, out_size, filter);
    auto it = c->tables.find(key);
    if (it == c->tables.end()) {
        if (c->tables.size() >= 256) {  // evict the least recently used table (cudaFree waits for its users)
            auto old = c->tables.begin();
            for (auto j = c->tables.begin(); j != c->tables.end(); ++j)
                if (j->second.stamp < old->second.stamp) old = j;
            cudaFree(old->second.kk);
            cudaFree(old->second.kk_t);
            cudaFree(old->second.kd);
            cudaFree(old->second.bounds);
            c->tables.erase(old);
        }
        ResizeTable t;
        std::vector<int> bounds, kk;
        resize_coefficients(in_size, out_size, filter, &t.ksize, bounds, kk);
        t.out_size = out_size;
        std::vector<int> kk_t(kk.size());
        for (int xx = 0; xx < out_size; ++xx)
            for (int k = 0; k < t.ksize; ++k)
                kk_t[static_cast<size_t>(k) * out_size + xx] = kk[static_cast<size_t>(xx) * t.ksize + k];
        for (int x0 = 0; x0 < out_size; x0 += kHBlock) {
            const int xl = std::min(x0 + kHBlock, out_size) - 1;
            t.max_span = std::max(t.max_span, bounds[2 * xl] + bounds[2 * xl + 1] - bounds[2 * x0]);
        }
        // signed base-256 digits of the 22-bit taps: c = c2*65536 + c1*256 + c0 with c0, c1 in [-128, 127]; the
        // horizontal pass then runs on dp4a (u8 pixels x s8 digits, exact in int32)
        t.groups = (t.ksize + 3) / 4;
        std::vector<unsigned> kd(static_cast<size_t>(t.groups) * 3 * out_size, 0u);
        for (int xx = 0; xx < out_size; ++xx)
            for (int k = 0; k < t.ksize; ++k) {
                int cc = kk[static_cast<size_t>(xx) * t.ksize + k];
                for (int d = 0; d < 3; ++d) {
                    const int dig = d < 2 ? ((cc + 128) & 255) - 128 : cc;
                    cc = (cc - dig) >> 8;
                    kd[(static_cast<size_t>(k / 4) * 3 + d) * out_size + xx] |=
                        static_cast<unsigned>(dig & 255) << (8 * (k & 3));
                }
            }
        VT_CUDA(cudaMalloc(&t.kd, kd.size() * sizeof(unsigned)));
        VT_CUDA(cudaMemcpy(t.kd, kd.data(), kd.size() * sizeof(unsigned), cudaMemcpyHostToDevice));
        VT_CUDA(cudaMalloc(&t.kk, kk.size() * sizeof(int)));
        VT_CUDA(cudaMalloc(&t.kk_t, kk.size() * sizeof(int)));
        VT_CUDA(cudaMalloc(&t.bounds, bounds.size() * sizeof(int)));
        VT_CUDA(cudaMemcpy(t.kk, kk.data(), kk.size() * sizeof(int), cudaMemcpyHostToDevice));
        VT_CUDA(cudaMemcpy(t.kk_t, kk_t.data(), kk.size() * sizeof(int), cudaMemcpyHostToDevice));
        VT_CUDA(cudaMemcpy(t.bounds, bounds.data(), bounds.size() * sizeof(int), cudaMemcpyHostToDevice));
        it = c->tables.emplace(key, t).first;
    }
    it->second.stamp = ++c->clock;
    *out = &it->second;
    return 0;
}

void smart_crop_box(int ow, int oh, int tw, int th, int* box) {
    // SmartResize, crop_mode 'center' (modules.py:149-172); Python float == C double
    const double target = static_cast<double>(tw) / th;
    const double ratio = static_cast<double>(ow) / oh;
    box[0] = 0; box[1] = 0; box[2] = ow; box[3] = oh;
    if (ratio > target) {
        const int nw = static_cast<int>(oh * target);
        const int left = (ow - nw) / 2;
        box[0] = left; box[2] = left + nw;
    } else if (ratio < target) {
        const int nh = static_cast<int>(ow / target);
        const int top = (oh - nh) / 2;
        box[1] = top; box[3] = top + nh;
    }
}

int resize_u8(ResizeCache* c, const vt_resize_args& a, Profiler* pf) {
    cudaStream_t s = static_cast<cudaStream_t>(a.stream);
    const int cw = a.crop_r - a.crop_l, ch = a.crop_b - a.crop_t;
    const unsigned char* src = static_cast<const unsigned char*>(a.src) + a.crop_t * a.src_stride + 3LL * a.crop_l;
    unsigned char* dst = static_cast<unsigned char*>(a.dst);
    const bool need_h = cw != a.dst_w, need_v = ch != a.dst_h;
    // PIL/Image.py Image.resize (Pillow 12.2): an image more than 100 times taller than wide that shrinks
    // vertically is resized in two calls, vertical pass first
    const bool v_first = ch > 100LL * cw && a.dst_h < ch;
    const double bytes = 3.0 * cw * ch + 3.0 * a.dst_w * a.dst_h +
                         (need_h && need_v ? 6.0 * (v_first ? 1.0 * cw * a.dst_h : 1.0 * a.dst_w * ch) : 0.0);
    profiler_begin(pf, KC_MISC, s, 0, bytes);
    auto ensure_tmp = [&](size_t need) -> int {
        if (need > c->tmp_cap) {
            if (c->tmp) cudaFree(c->tmp);
            c->tmp = nullptr;
            c->tmp_cap = 0;
            VT_CUDA(cudaMalloc(&c->tmp, need));
            c->tmp_cap = need;
        }
        return 0;
    };
    auto pass_h = [&](const unsigned char* in, long long in_pitch, int rows, unsigned char* out,
                      long long out_pitch) -> int {
        const ResizeTable* t = nullptr;
        VT_TRY(get_table(c, cw, a.dst_w, a.filter, &t));
        static const bool scalar = [] { const char* e = getenv("VT_B200_RESIZE_SCALAR"); return e && e[0] == '1'; }();
        const int gx = (a.dst_w + kHBlock - 1) / kHBlock;
        if (!scalar) {
            // plane of one channel of one row: the span, the word the last window may run into, alignment slack
            const int plane_pitch = (t->max_span + 8 + 15) / 16 * 16;
            VT_CHECK(3 * plane_pitch <= 200 * 1024, "horizontal scale factor too large for the staged resize kernel");
            const int R = 12 * plane_pitch <= 96 * 1024 && rows >= 4 ? 4 : 1;
            const size_t smem = static_cast<size_t>(3) * plane_pitch * R;
            if (smem > 48 * 1024) {
                VT_CUDA(cudaFuncSetAttribute(resize_h_dp4a_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
                VT_CUDA(cudaFuncSetAttribute(resize_h_dp4a_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            }
            int rows_per_cta = R;
            while (rows_per_cta < 16 && 1LL * gx * ((rows + rows_per_cta - 1) / rows_per_cta) > 148 * 16) rows_per_cta *= 2;
            const dim3 grid(gx, (rows + rows_per_cta - 1) / rows_per_cta);
            if (R == 4)
                resize_h_dp4a_kernel<4><<<grid, kHBlock, smem, s>>>(in, in_pitch, out, out_pitch, t->kd, t->bounds, a.dst_w,
                                                                    rows, rows_per_cta, plane_pitch);
            else
                resize_h_dp4a_kernel<1><<<grid, kHBlock, smem, s>>>(in, in_pitch, out, out_pitch, t->kd, t->bounds, a.dst_w,
                                                                    rows, rows_per_cta, plane_pitch);
            return 0;
        }
        const int span_pitch = (t->max_span * 3 + 15) / 16 * 16;
        VT_CHECK(span_pitch <= 200 * 1024, "horizontal scale factor too large for the staged resize kernel");
        const int R = 4 * span_pitch <= 96 * 1024 && rows >= 4 ? 4 : 1;  // rows staged together
        const size_t smem = static_cast<size_t>(span_pitch) * R;
        if (smem > 48 * 1024) {
            VT_CUDA(cudaFuncSetAttribute(resize_h_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            VT_CUDA(cudaFuncSetAttribute(resize_h_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        }
        int rows_per_cta = R;
        while (rows_per_cta < 16 && 1LL * gx * ((rows + rows_per_cta - 1) / rows_per_cta) > 148 * 16) rows_per_cta *= 2;
        const dim3 grid(gx, (rows + rows_per_cta - 1) / rows_per_cta);
        if (R == 4)
            resize_h_kernel<4><<<grid, kHBlock, smem, s>>>(in, in_pitch, out, out_pitch, t->kk_t, t->bounds, a.dst_w,
                                                           rows, rows_per_cta, span_pitch);
        else
            resize_h_kernel<1><<<grid, kHBlock, smem, s>>>(in, in_pitch, out, out_pitch, t->kk_t, t->bounds, a.dst_w,
                                                           rows, rows_per_cta, span_pitch);
        return 0;
    };
    auto pass_v = [&](const unsigned char* in, long long in_pitch, int width, unsigned char* out,
                      long long out_pitch) -> int {
        const ResizeTable* t = nullptr;
        VT_TRY(get_table(c, ch, a.dst_h, a.filter, &t));
        const int row_bytes = 3 * width;
        const bool vec = row_bytes % 4 == 0 && in_pitch % 4 == 0 && out_pitch % 4 == 0 &&
                         reinterpret_cast<uintptr_t>(in) % 4 == 0 && reinterpret_cast<uintptr_t>(out) % 4 == 0;
        if (vec)
            resize_v_kernel<4><<<dim3((row_bytes / 4 + 255) / 256, a.dst_h), 256, 0, s>>>(
                in, in_pitch, out, out_pitch, t->kk, t->bounds, t->ksize, row_bytes);
        else
            resize_v_kernel<1><<<dim3((row_bytes + 255) / 256, a.dst_h), 256, 0, s>>>(
                in, in_pitch, out, out_pitch, t->kk, t->bounds, t->ksize, row_bytes);
        return 0;
    };
    if (need_h && need_v) {  // uint8 intermediate image, rows padded to 16 bytes
        const int mid_w = v_first ? cw : a.dst_w, mid_h = v_first ? a.dst_h : ch;
        const long long mid_pitch = (3LL * mid_w + 15) / 16 * 16;
        VT_TRY(ensure_tmp(static_cast<size_t>(mid_pitch) * mid_h));
        unsigned char* mid = static_cast<unsigned char*>(c->tmp);
        if (v_first) {
            VT_TRY(pass_v(src, a.src_stride, cw, mid, mid_pitch));
            VT_TRY(pass_h(mid, mid_pitch, a.dst_h, dst, a.dst_stride));
        } else {
            VT_TRY(pass_h(src, a.src_stride, ch, mid, mid_pitch));
            VT_TRY(pass_v(mid, mid_pitch, a.dst_w, dst, a.dst_stride));
        }
    } else if (need_h) {
        VT_TRY(pass_h(src, a.src_stride, ch, dst, a.dst_stride));
    } else if (need_v) {
        VT_TRY(pass_v(src, a.src_stride, cw, dst, a.dst_stride));
    } else {
        copy_rows_kernel<<<dim3((3 * cw + 255) / 256, ch), 256, 0, s>>>(src, a.src_stride, dst, a.dst_stride, 3 * cw);
    }
    profiler_end(pf, KC_MISC, s);
    VT_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace vt
