// Backward pass of the encoder's building blocks (the reference fine-tunes the VAE through
// autograd in train_vae.py:124-186 and train_full.py:201-256; the arithmetic restated here is the autograd of
// diffusers' ResnetBlock2D = GroupNorm(32) -> SiLU -> conv3x3, twice, plus the identity / 1x1 shortcut).
//
//   * data gradient of a 3x3 / 1x1 stride-1 conv  = the forward implicit-GEMM tcgen05 kernel (vt_igemm.cuh) run on
//     the output gradient with the weights flipped and transposed ([Cin][tap'][Cout], tap' = 8 - tap);
//   * weight gradient = a tcgen05 GEMM whose K dimension is the PIXELS: dW[co][tap][ci] = sum_p dY[p][co] A[p+tap][ci],
//     with BOTH operands read straight from the NHWC tensors as MN-major tiles and every tap a shifted TMA box
//     (vt_wgrad.cu); split-K over pixel patches, fp32 partial tiles, wgrad_reduce_kernel adds them in index order --
//     no atomics, bit-reproducible.  The conv input is handed over as a bf16 NHWC operand: through GroupNorm+SiLU when
//     the conv saw the normalised tensor, a plain fp16 -> bf16 copy otherwise (bwd_conv_wgrad16);
//   * GroupNorm + SiLU backward: two HBM passes (per-channel sums of dt and dt*xhat with a fixed-order two-stage
//     reduce, then the apply pass, with the residual-branch gradient added in the same pass);
//   * bias gradients: fixed-order column sums.
//
// 16-bit mode: activations in the context's raw format, gradients and the re-laid-out operands bf16 (range of a
// gradient is unbounded and no loss scaling exists in the reference), fp32 accumulation, fp32 parameter gradients.
// fp32 mode (verification): everything fp32 on the FFMA pipe (f32_wgrad_kernel, launch_conv_fp32).
#include <cstdlib>
#include <type_traits>

#include "vt_backward.h"
#include "vt_ptx.cuh"

namespace vt {

namespace {

inline size_t al(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

template <int F>
__device__ __forceinline__ void load8(const void* base, long long off, float v[8]) {
    if constexpr (F == FMT_F32) {
        const float* p = static_cast<const float*>(base) + off;
        const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
        const uint4 u = *reinterpret_cast<const uint4*>(static_cast<const bf16*>(base) + off);
        v[0] = raw16_lo<F>(u.x); v[1] = raw16_hi<F>(u.x); v[2] = raw16_lo<F>(u.y); v[3] = raw16_hi<F>(u.y);
        v[4] = raw16_lo<F>(u.z); v[5] = raw16_hi<F>(u.z); v[6] = raw16_lo<F>(u.w); v[7] = raw16_hi<F>(u.w);
    }
}
template <int F>
__device__ __forceinline__ void store8(void* base, long long off, const float v[8]) {
    if constexpr (F == FMT_F32) {
        float* p = static_cast<float*>(base) + off;
        *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
    } else {
        *reinterpret_cast<uint4*>(static_cast<bf16*>(base) + off) =
            make_uint4(pack16x2<F>(v[0], v[1]), pack16x2<F>(v[2], v[3]), pack16x2<F>(v[4], v[5]), pack16x2<F>(v[6], v[7]));
    }
}

// per-channel GroupNorm constants of the 8 channels a thread owns
struct Gn8 {
    float mean[8], rstd[8], ga[8], be[8];
};
__device__ __forceinline__ void gn8_load(Gn8& k, const double* stats, const float* gamma, const float* beta, int n,
                                         int c0, int C, long long HW, float eps) {
    const int cpg = C / 32;
    const double cnt = static_cast<double>(HW) * cpg;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const int c = c0 + e, g = c / cpg;
        const double su = stats[(1LL * n * 32 + g) * 2], sq = stats[(1LL * n * 32 + g) * 2 + 1];
        const double mean = su / cnt;
        double var = sq / cnt - mean * mean;
        var = var > 0.0 ? var : 0.0;
        k.mean[e] = static_cast<float>(mean);
        k.rstd[e] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
        k.ga[e] = gamma[c];
        k.be[e] = beta[c];
    }
}
// xhat and dt = dy * act'(t), t = gamma*xhat + beta.  FAST (16-bit mode): tanh.approx (2^-11 relative) -- below the bf16
// rounding (2^-9) of the gradient that is stored afterwards
template <bool FAST>
__device__ __forceinline__ void gn_dt(float x, float dy, float mean, float rstd, float ga, float be, int silu, float& xhat,
                                      float& dt) {
    xhat = (x - mean) * rstd;
    dt = dy;
    if (silu) {
        const float t = fmaf(ga, xhat, be);
        float sg;
        if (FAST) {   // one MUFU: sigmoid(t) = 0.5 + 0.5 tanh(t/2)
            float th;
            asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(0.5f * t));
            sg = fmaf(0.5f, th, 0.5f);
        } else {
            sg = 1.0f / (1.0f + expf(-t));
        }
        dt = dy * sg * (1.0f + t * (1.0f - sg));
    }
}

// ---------------------------------------------------------------------------------------------------------
// GroupNorm(+SiLU) backward, pass 1: part[n][chunk][c] = (sum_p dt, sum_p dt*xhat) over the chunk's pixels.
// Thread = 8 consecutive channels, walks pixels; the pixel sub-lanes of a block are folded in index order.
template <int XF, int GF>
__global__ void __launch_bounds__(256) gn_bwd_reduce_kernel(const void* __restrict__ x, const void* __restrict__ dy,
                                                            const double* __restrict__ stats,
                                                            const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, float* __restrict__ part,
                                                            long long HW, int C, float eps, int silu) {
    __shared__ float sh[256 * 16];
    const int n = blockIdx.y;
    const int tpb = C / 8;                 // threads across the channels (C <= 2048)
    const int ppb = 256 / tpb;
    const int c8 = threadIdx.x % tpb, psub = threadIdx.x / tpb;
    const long long chunk = (HW + gridDim.x - 1) / gridDim.x;
    const long long p0 = blockIdx.x * chunk, p1 = min(HW, p0 + chunk);
    float a[8], b[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) a[e] = b[e] = 0.f;
    if (psub < ppb) {
        Gn8 k;
        gn8_load(k, stats, gamma, beta, n, c8 * 8, C, HW, eps);
        // two pixels per trip: all four loads are issued before the first is used (same summation order as one
        // pixel per trip)
        for (long long p = p0 + psub; p < p1; p += 2 * ppb) {
            const long long off = (1LL * n * HW + p) * C + c8 * 8;
            const bool two = p + ppb < p1;
            const long long off2 = two ? off + 1LL * ppb * C : off;
            float xv[8], gv[8], xw[8], gw[8];
            load8<XF>(x, off, xv);
            load8<GF>(dy, off, gv);
            load8<XF>(x, off2, xw);
            load8<GF>(dy, off2, gw);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                float xh, dt;
                gn_dt<XF != FMT_F32>(xv[e], gv[e], k.mean[e], k.rstd[e], k.ga[e], k.be[e], silu, xh, dt);
                a[e] += dt;
                b[e] = fmaf(dt, xh, b[e]);
            }
            if (two) {
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    float xh, dt;
                    gn_dt<XF != FMT_F32>(xw[e], gw[e], k.mean[e], k.rstd[e], k.ga[e], k.be[e], silu, xh, dt);
                    a[e] += dt;
                    b[e] = fmaf(dt, xh, b[e]);
                }
            }
        }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        sh[threadIdx.x * 16 + e] = a[e];
        sh[threadIdx.x * 16 + 8 + e] = b[e];
    }
    __syncthreads();
    if (threadIdx.x < tpb) {
        for (int e = 0; e < 16; ++e) {
            float t = 0.f;
            for (int q = 0; q < ppb; ++q) t += sh[(q * tpb + threadIdx.x) * 16 + e];
            const int c = threadIdx.x * 8 + (e & 7);
            part[((1LL * n * gridDim.x + blockIdx.x) * C + c) * 2 + (e >> 3)] = t;
        }
    }
}

// pass 1b: one warp per channel -- lanes stride over the chunks of an image, then a fixed shuffle tree (fp64): the order
// of the additions depends on (chunks) only.  Writes the per-(image, channel) sums times gamma for the group stage
// and d gamma / d beta summed over the images in index order.
__global__ void __launch_bounds__(256) gn_bwd_finalize_kernel(const float* __restrict__ part, const float* __restrict__ gamma,
                                                              float* __restrict__ ab /*[N][C][2]*/, float* __restrict__ dgamma,
                                                              float* __restrict__ dbeta, int N, int chunks, int C, int accumulate) {
    const int c = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (c >= C) return;
    double dg = 0.0, db = 0.0;
    const float ga = gamma[c];
    for (int n = 0; n < N; ++n) {
        double A = 0.0, B = 0.0;
        for (int k = lane; k < chunks; k += 32) {
            const float2 v = *reinterpret_cast<const float2*>(part + ((1LL * n * chunks + k) * C + c) * 2);
            A += v.x;
            B += v.y;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            A += __shfl_xor_sync(0xFFFFFFFFu, A, o);
            B += __shfl_xor_sync(0xFFFFFFFFu, B, o);
        }
        db += A;
        dg += B;
        if (lane == 0) {
            ab[(1LL * n * C + c) * 2] = static_cast<float>(A * ga);
            ab[(1LL * n * C + c) * 2 + 1] = static_cast<float>(B * ga);
        }
    }
    if (lane == 0) {
        dgamma[c] = (accumulate ? dgamma[c] : 0.f) + static_cast<float>(dg);
        dbeta[c] = (accumulate ? dbeta[c] : 0.f) + static_cast<float>(db);
    }
}
// pass 1c: per (image, group) S1 = sum_c gamma_c A_c, S2 = sum_c gamma_c B_c over the group's channels, in channel order
__global__ void gn_bwd_groups_kernel(const float* __restrict__ ab, float* __restrict__ gsum /*[N][32][2]*/, int C) {
    const int n = blockIdx.x, g = threadIdx.x;
    const int cpg = C / 32;
    double s1 = 0.0, s2 = 0.0;
    for (int e = 0; e < cpg; ++e) {
        s1 += ab[(1LL * n * C + g * cpg + e) * 2];
        s2 += ab[(1LL * n * C + g * cpg + e) * 2 + 1];
    }
    gsum[(n * 32 + g) * 2] = static_cast<float>(s1);
    gsum[(n * 32 + g) * 2 + 1] = static_cast<float>(s2);
}

// pass 2: dx = rstd * (gamma*dt - (S1 + xhat*S2)/m) (+ add)
template <int XF, int GF>
__global__ void __launch_bounds__(256) gn_bwd_apply_kernel(const void* __restrict__ x, const void* __restrict__ dy,
                                                           const double* __restrict__ stats,
                                                           const float* __restrict__ gamma,
                                                           const float* __restrict__ beta,
                                                           const float* __restrict__ gsum,
                                                           const void* __restrict__ add, void* __restrict__ dx,
                                                           long long HW, int C, float eps, int silu) {
    const int n = blockIdx.y;
    const int tpb = C / 8;
    const int ppb = 256 / tpb;
    const int c8 = threadIdx.x % tpb, psub = threadIdx.x / tpb;
    if (psub >= ppb) return;
    const int cpg = C / 32;
    const float inv_m = 1.0f / (static_cast<float>(HW) * cpg);
    const long long chunk = (HW + gridDim.x - 1) / gridDim.x;
    const long long p0 = blockIdx.x * chunk, p1 = min(HW, p0 + chunk);
    Gn8 k;
    gn8_load(k, stats, gamma, beta, n, c8 * 8, C, HW, eps);
    float s1[8], s2[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const int g = (c8 * 8 + e) / cpg;
        s1[e] = gsum[(n * 32 + g) * 2] * inv_m;
        s2[e] = gsum[(n * 32 + g) * 2 + 1] * inv_m;
    }
    for (long long p = p0 + psub; p < p1; p += 2 * ppb) {   // two pixels per trip, loads first
        const long long off = (1LL * n * HW + p) * C + c8 * 8;
        const bool two = p + ppb < p1;
        const long long off2 = two ? off + 1LL * ppb * C : off;
        float xv[8], gv[8], av[8], xw[8], gw[8], aw[8], o[8];
        load8<XF>(x, off, xv);
        load8<GF>(dy, off, gv);
        if (add) load8<GF>(add, off, av);
        load8<XF>(x, off2, xw);
        load8<GF>(dy, off2, gw);
        if (add) load8<GF>(add, off2, aw);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            float xh, dt;
            gn_dt<XF != FMT_F32>(xv[e], gv[e], k.mean[e], k.rstd[e], k.ga[e], k.be[e], silu, xh, dt);
            o[e] = k.rstd[e] * (k.ga[e] * dt - (s1[e] + xh * s2[e]));
            if (add) o[e] += av[e];
        }
        store8<GF>(dx, off, o);
        if (two) {
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                float xh, dt;
                gn_dt<XF != FMT_F32>(xw[e], gw[e], k.mean[e], k.rstd[e], k.ga[e], k.be[e], silu, xh, dt);
                o[e] = k.rstd[e] * (k.ga[e] * dt - (s1[e] + xh * s2[e]));
                if (add) o[e] += aw[e];
            }
            store8<GF>(dx, off2, o);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// bias gradient: column sums of g[rows][C], two fixed-order stages
template <int GF>
__global__ void __launch_bounds__(256) colsum_part_kernel(const void* __restrict__ g, float* __restrict__ part,
                                                          long long rows, int C) {
    __shared__ float sh[256 * 8];
    const int tpb = C / 8, ppb = 256 / tpb;
    const int c8 = threadIdx.x % tpb, psub = threadIdx.x / tpb;
    const long long chunk = (rows + gridDim.x - 1) / gridDim.x;
    const long long p0 = blockIdx.x * chunk, p1 = min(rows, p0 + chunk);
    float a[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) a[e] = 0.f;
    if (psub < ppb)
        for (long long p = p0 + psub; p < p1; p += ppb) {
            float v[8];
            load8<GF>(g, p * C + c8 * 8, v);
#pragma unroll
            for (int e = 0; e < 8; ++e) a[e] += v[e];
        }
#pragma unroll
    for (int e = 0; e < 8; ++e) sh[threadIdx.x * 8 + e] = a[e];
    __syncthreads();
    if (threadIdx.x < tpb)
        for (int e = 0; e < 8; ++e) {
            float t = 0.f;
            for (int q = 0; q < ppb; ++q) t += sh[(q * tpb + threadIdx.x) * 8 + e];
            part[1LL * blockIdx.x * C + threadIdx.x * 8 + e] = t;
        }
}
__global__ void colsum_final_kernel(const float* __restrict__ part, float* __restrict__ out, int chunks, int C,
                                    int accumulate) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    double t = 0.0;
    for (int k = 0; k < chunks; ++k) t += part[1LL * k * C + c];
    out[c] = (accumulate ? out[c] : 0.f) + static_cast<float>(t);
}

// ---------------------------------------------------------------------------------------------------------
// flipped / transposed weights of the data-gradient conv: dst[ci][(ks*ks-1-tap)*Cout + co] = w[co][ci][tap]
template <int OFMT>
__global__ void pack_dgrad_weight_kernel(const float* __restrict__ w /*[Cout][Cin][ks][ks]*/, void* __restrict__ dst,
                                         int Cout, int Cin, int ks) {
    const long long total = 1LL * Cout * Cin * ks * ks;
    for (long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < total; i += 1LL * gridDim.x * blockDim.x) {
        const int co = static_cast<int>(i % Cout);
        long long r = i / Cout;
        const int tp = static_cast<int>(r % (ks * ks));
        const int ci = static_cast<int>(r / (ks * ks));
        const float v = w[(1LL * co * Cin + ci) * ks * ks + (ks * ks - 1 - tp)];
        if constexpr (OFMT == FMT_F32) static_cast<float*>(dst)[i] = v;
        else static_cast<bf16*>(dst)[i] = __float2bfloat16(v);
    }
}

// partial tiles part[b][co][taps*Cin] -> dW [Cout][Cin][ks][ks] (OIHW, what torch holds), batches in index order
__global__ void wgrad_reduce_kernel(const float* __restrict__ part, float* __restrict__ dw, int batches, int Cout,
                                    int Cin, int taps, int accumulate) {
    const long long total = 1LL * Cout * Cin * taps;
    for (long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < total; i += 1LL * gridDim.x * blockDim.x) {
        const int ci = static_cast<int>(i % Cin);
        long long r = i / Cin;
        const int tp = static_cast<int>(r % taps);
        const int co = static_cast<int>(r / taps);
        // batches b = 4j + r go to partial sum r (four loads in flight), the four are combined in index order: the
        // association is fixed by `batches` alone
        const float* src = part + 1LL * co * taps * Cin + 1LL * tp * Cin + ci;
        const long long bs = 1LL * Cout * taps * Cin;
        double t0 = 0.0, t1 = 0.0, t2 = 0.0, t3 = 0.0;
        int b = 0;
        for (; b + 4 <= batches; b += 4) {
            t0 += src[(b + 0) * bs]; t1 += src[(b + 1) * bs]; t2 += src[(b + 2) * bs]; t3 += src[(b + 3) * bs];
        }
        for (; b < batches; ++b) t0 += src[b * bs];
        const double t = (t0 + t1) + (t2 + t3);
        float* o = dw + (1LL * co * Cin + ci) * taps + tp;
        *o = (accumulate ? *o : 0.f) + static_cast<float>(t);
    }
}

// ---------------------------------------------------------------------------------------------------------
// fp32 verification mode: weight gradient straight from the NHWC tensors on the FFMA pipe ("TN" GEMM, K = pixels).
// grid (Cout/64, taps * Cin/64, splits); part[split][co][taps*Cin]
// H x W: dimensions of dy (the conv output); the conv input a is (stride*H [+1]) x (stride*W [+1]) = Hi x Wi and the
// tap (ky, kx) of output pixel (y, x) reads input pixel (stride*y + ky - pad, stride*x + kx - pad).
__global__ void __launch_bounds__(256) f32_wgrad_kernel(const float* __restrict__ dy, const float* __restrict__ a,
                                                        float* __restrict__ part, int N, int H, int W, int Cout, int Cin,
                                                        int ks, int stride, int Hi, int Wi) {
    __shared__ float As[16][68], Bs[16][68];
    const int taps = ks * ks;
    const int cit = Cin / 64;
    const int tap = blockIdx.y / cit, ci0 = (blockIdx.y % cit) * 64, co0 = blockIdx.x * 64;
    const int pad = (ks == 3 && stride == 1) ? 1 : 0;
    const int dyo = ks == 3 ? tap / 3 - pad : 0, dxo = ks == 3 ? tap % 3 - pad : 0;
    const long long P = 1LL * N * H * W;
    const long long chunk = (P + gridDim.z - 1) / gridDim.z;
    const long long p0 = blockIdx.z * chunk, p1 = min(P, p0 + chunk);
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int lk = threadIdx.x >> 4, lc = (threadIdx.x & 15) * 4;
    float acc[4][4] = {};
    for (long long pb = p0; pb < p1; pb += 16) {
        const long long p = pb + lk;
        float4 av = make_float4(0.f, 0.f, 0.f, 0.f), bv = av;
        if (p < p1) {
            av = *reinterpret_cast<const float4*>(dy + p * Cout + co0 + lc);
            const int x = static_cast<int>(p % W);
            const long long r = p / W;
            const int y = static_cast<int>(r % H);
            const long long img = r / H;
            const int ys = y * stride + dyo, xs = x * stride + dxo;
            if (ys >= 0 && ys < Hi && xs >= 0 && xs < Wi)
                bv = *reinterpret_cast<const float4*>(a + ((img * Hi + ys) * Wi + xs) * Cin + ci0 + lc);
        }
        *reinterpret_cast<float4*>(&As[lk][lc]) = av;
        *reinterpret_cast<float4*>(&Bs[lk][lc]) = bv;
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
            const float4 a4 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
            const float4 b4 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
            const float aa[4] = {a4.x, a4.y, a4.z, a4.w}, bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(aa[i], bb[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
            part[(1LL * blockIdx.z * Cout + co0 + ty * 4 + i) * taps * Cin + 1LL * tap * Cin + ci0 + tx * 4 + j] = acc[i][j];
}

int grid_for(long long n) { return static_cast<int>(std::min<long long>((n + 255) / 256, 148 * 8)); }

}  // namespace

// =========================================================================================================
int bwd_gn_chunks(int N, long long HW) {
    const long long want = (HW + 255) / 256;
    return static_cast<int>(std::max<long long>(1, std::min<long long>(want, std::max(1, 148 * 4 / std::max(N, 1)))));
}
size_t bwd_gn_scratch_bytes(int N, long long HW, int C) {
    return al(static_cast<size_t>(N) * bwd_gn_chunks(N, HW) * C * 2 * sizeof(float)) +
           al(static_cast<size_t>(N) * 64 * sizeof(float) + static_cast<size_t>(N) * C * 2 * sizeof(float));
}

int bwd_group_norm(const BwdEnv& e, const void* x, const void* dy, const double* stats, const float* gamma,
                   const float* beta, const void* add, void* dx, float* dgamma, float* dbeta, int N, long long HW, int C,
                   float eps, int silu, int accumulate, void* scratch) {
    VT_CHECK(C % 256 == 0 || C == 128, "GroupNorm backward: 128 channels or a multiple of 256");
    VT_CHECK(C <= 1024, "GroupNorm backward: at most 1024 channels");
    const int chunks = bwd_gn_chunks(N, HW);
    float* part = static_cast<float*>(scratch);
    float* gsum = reinterpret_cast<float*>(static_cast<char*>(scratch) + al(static_cast<size_t>(N) * chunks * C * 2 * sizeof(float)));
    dim3 grid(chunks, N);
    profiler_begin(e.prof, KC_GN_APPLY, e.s, 0, 2.0 * N * HW * C * (e.fp32 ? 8.0 : 4.0) + 1.0 * N * HW * C * (e.fp32 ? 4 : 2));
    if (e.fp32) {
        gn_bwd_reduce_kernel<FMT_F32, FMT_F32><<<grid, 256, 0, e.s>>>(x, dy, stats, gamma, beta, part, HW, C, eps, silu);
    } else if (e.raw_fmt == FMT_F16) {
        gn_bwd_reduce_kernel<FMT_F16, FMT_BF16><<<grid, 256, 0, e.s>>>(x, dy, stats, gamma, beta, part, HW, C, eps, silu);
    } else {
        gn_bwd_reduce_kernel<FMT_BF16, FMT_BF16><<<grid, 256, 0, e.s>>>(x, dy, stats, gamma, beta, part, HW, C, eps, silu);
    }
    float* ab = gsum + static_cast<size_t>(N) * 64;
    gn_bwd_finalize_kernel<<<(C + 7) / 8, 256, 0, e.s>>>(part, gamma, ab, dgamma, dbeta, N, chunks, C, accumulate);
    gn_bwd_groups_kernel<<<N, 32, 0, e.s>>>(ab, gsum, C);
    if (e.fp32) {
        gn_bwd_apply_kernel<FMT_F32, FMT_F32><<<grid, 256, 0, e.s>>>(x, dy, stats, gamma, beta, gsum, add, dx, HW, C, eps, silu);
    } else if (e.raw_fmt == FMT_F16) {
        gn_bwd_apply_kernel<FMT_F16, FMT_BF16><<<grid, 256, 0, e.s>>>(x, dy, stats, gamma, beta, gsum, add, dx, HW, C, eps, silu);
    } else {
        gn_bwd_apply_kernel<FMT_BF16, FMT_BF16><<<grid, 256, 0, e.s>>>(x, dy, stats, gamma, beta, gsum, add, dx, HW, C, eps, silu);
    }
    profiler_end(e.prof, KC_GN_APPLY, e.s);
    VT_CUDA(cudaGetLastError());
    return 0;
}

size_t bwd_colsum_scratch_bytes(int C) { return al(static_cast<size_t>(148 * 4) * C * sizeof(float)); }
int bwd_bias_grad(const BwdEnv& e, const void* g, long long rows, int C, float* db, int accumulate, void* scratch) {
    VT_CHECK(C % 8 == 0 && C <= 2048 && 256 % (C / 8) == 0, "bias gradient: channel count must divide 2048");
    const int chunks = static_cast<int>(std::max<long long>(1, std::min<long long>((rows + 255) / 256, 148 * 4)));
    float* part = static_cast<float*>(scratch);
    if (e.fp32) colsum_part_kernel<FMT_F32><<<chunks, 256, 0, e.s>>>(g, part, rows, C);
    else colsum_part_kernel<FMT_BF16><<<chunks, 256, 0, e.s>>>(g, part, rows, C);
    colsum_final_kernel<<<(C + 127) / 128, 128, 0, e.s>>>(part, db, chunks, C, accumulate);
    VT_CUDA(cudaGetLastError());
    return 0;
}

// ---- data gradient
size_t bwd_dgrad_weight_bytes(const BwdEnv& e, int Cout, int Cin, int ks) {
    return al(static_cast<size_t>(Cout) * Cin * ks * ks * (e.fp32 ? 4 : 2));
}
int bwd_pack_dgrad_weight(const BwdEnv& e, const float* w, void* dst, int Cout, int Cin, int ks) {
    const long long total = 1LL * Cout * Cin * ks * ks;
    if (e.fp32) pack_dgrad_weight_kernel<FMT_F32><<<grid_for(total), 256, 0, e.s>>>(w, dst, Cout, Cin, ks);
    else pack_dgrad_weight_kernel<FMT_BF16><<<grid_for(total), 256, 0, e.s>>>(w, dst, Cout, Cin, ks);
    VT_CUDA(cudaGetLastError());
    return 0;
}
int bwd_conv_dgrad(const BwdEnv& e, const void* dy, const void* wd, void* dx, const void* add, int N, int H, int W,
                   int Cout, int Cin, int ks) {
    // a stride-1 conv of the output gradient ([N][H][W][Cout]) with the flipped / transposed weights [Cin][ks*ks*Cout]
    ConvOp op;
    op.in = dy; op.in_f16 = 0; op.raw_f16 = 0; op.N = N; op.Hin = H; op.Win = W; op.Cin = Cout; op.ksize = ks; op.stride = 1;
    op.w = wd; op.Cout = Cin; op.out = dx; op.out_fmt = e.fp32 ? FMT_F32 : FMT_BF16;
    op.residual = add; op.residual_fp32 = e.fp32;
    op.kclass = KC_BWD;
    return e.fp32 ? launch_conv_fp32(op, e.s, e.prof) : launch_conv(op, e.s, e.prof);
}

// ---- weight gradient
// fp32 verification mode: split-K plan of the FFMA weight-gradient kernel
WgradPlan bwd_wgrad_plan(const BwdEnv& e, int N, int H, int W, int Cout, int Cin, int ks) {
    (void)e;
    WgradPlan p{};
    p.taps = ks * ks;
    const long long P = 1LL * N * H * W;
    const int tiles = std::max(1, (Cout / 64) * (p.taps * Cin / 64));
    p.splits = static_cast<int>(std::max<long long>(1, std::min<long long>((P + 1023) / 1024, std::max(1, 148 * 8 / tiles))));
    p.batches = p.splits;
    p.part_bytes = al(static_cast<size_t>(p.batches) * Cout * p.taps * Cin * sizeof(float));
    return p;
}

int bwd_conv_wgrad(const BwdEnv& e, const WgradPlan& p, const void* dy, const void* a, float* part, float* dw, int N, int H,
                   int W, int Cout, int Cin, int ks, int accumulate) {
    VT_CHECK(e.fp32, "bwd_conv_wgrad is the fp32 verification path (16-bit mode: bwd_conv_wgrad16)");
    VT_CHECK(Cout % 64 == 0 && Cin % 64 == 0, "fp32 weight gradient: channels must be multiples of 64");
    dim3 grid(Cout / 64, p.taps * (Cin / 64), p.splits);
    profiler_begin(e.prof, KC_FP32, e.s, 2.0 * N * H * W * Cout * Cin * p.taps, 0);
    f32_wgrad_kernel<<<grid, 256, 0, e.s>>>(static_cast<const float*>(dy), static_cast<const float*>(a), part, N, H, W, Cout, Cin, ks, 1, H, W);
    profiler_end(e.prof, KC_FP32, e.s);
    const long long total = 1LL * Cout * Cin * p.taps;
    wgrad_reduce_kernel<<<grid_for(total), 256, 0, e.s>>>(part, dw, p.batches, Cout, Cin, p.taps, accumulate);
    VT_CUDA(cudaGetLastError());
    return 0;
}


// =========================================================================================================
// Pieces of the whole-encoder backward (vt_train_encoder.cuh): stride-2 downsample convs, conv_in, attention.
namespace {

template <int F>
__device__ __forceinline__ float ld1(const void* p, long long i) {
    if constexpr (F == FMT_F32) return static_cast<const float*>(p)[i];
    else if constexpr (F == FMT_F16) return __half2float(static_cast<const __half*>(p)[i]);
    else return __bfloat162float(static_cast<const bf16*>(p)[i]);
}
template <int F>
__device__ __forceinline__ void st1(void* p, long long i, float v) {
    if constexpr (F == FMT_F32) static_cast<float*>(p)[i] = v;
    else if constexpr (F == FMT_F16) static_cast<__half*>(p)[i] = __float2half_rn(fminf(fmaxf(v, -65504.f), 65504.f));
    else static_cast<bf16*>(p)[i] = __float2bfloat16(v);
}

// out[b][c][r] = in[b][r][c] (r < rows, c < cols), 32x32 tiles; out columns rows..ld_out-1 are left untouched
template <int FI, int FO>
__global__ void __launch_bounds__(256) transpose_kernel(const void* __restrict__ in, void* __restrict__ out, int rows, int cols,
                                                        long long ld_in, long long ld_out, long long in_bs, long long out_bs) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z;
    const int r0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int i = ty; i < 32; i += 8) {
        const int r = r0 + i, c = c0 + tx;
        tile[i][tx] = (r < rows && c < cols) ? ld1<FI>(in, b * in_bs + 1LL * r * ld_in + c) : 0.f;
    }
    __syncthreads();
    for (int i = ty; i < 32; i += 8) {
        const int c = c0 + i, r = r0 + tx;
        if (c < cols && r < rows) st1<FO>(out, b * out_bs + 1LL * c * ld_out + r, tile[tx][i]);
    }
}

// the same for 16-bit -> 16-bit (optionally fp16 -> bf16): 64x64 tiles moved as 32-bit pairs, so a warp reads and
// writes 128 contiguous bytes per row (the 32x32 element-wise form above ran the T x T transposes of the attention
// backward at 1.9 TB/s).  ld_in, ld_out even, 4-byte aligned bases.
template <int FI, int FO>
__global__ void __launch_bounds__(256) transpose16_kernel(const unsigned short* __restrict__ in, unsigned short* __restrict__ out,
                                                          int rows, int cols, long long ld_in, long long ld_out, long long in_bs,
                                                          long long out_bs) {
    __shared__ uint32_t tile[64][33];
    in += blockIdx.z * in_bs;
    out += blockIdx.z * out_bs;
    const int r0 = blockIdx.x * 64, c0 = blockIdx.y * 64;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int i = ty; i < 64; i += 8) {
        const int r = r0 + i, c = c0 + 2 * tx;
        uint32_t v = 0;
        if (r < rows) {
            if (c + 1 < cols) v = *reinterpret_cast<const uint32_t*>(in + 1LL * r * ld_in + c);
            else if (c < cols) v = in[1LL * r * ld_in + c];
        }
        if (FI == FMT_F16 && FO == FMT_BF16) v = pack_bf16x2(f16_lo(v), f16_hi(v));
        tile[i][tx] = v;
    }
    __syncthreads();
    for (int i = ty; i < 64; i += 8) {
        const int c = c0 + i, r = r0 + 2 * tx;
        if (c >= cols || r >= rows) continue;
        const uint32_t w0 = tile[2 * tx][i >> 1], w1 = tile[2 * tx + 1][i >> 1];
        const uint32_t lo = (i & 1) ? (w0 >> 16) : (w0 & 0xFFFFu), hi = (i & 1) ? (w1 >> 16) : (w1 & 0xFFFFu);
        if (r + 1 < rows) *reinterpret_cast<uint32_t*>(out + 1LL * c * ld_out + r) = lo | (hi << 16);
        else out[1LL * c * ld_out + r] = static_cast<unsigned short>(lo);
    }
}

// out[row] = sum_c a[row][c] * b[row][c]   (one warp per row)
template <int FA, int FB>
__global__ void __launch_bounds__(256) rowdot_kernel(const void* __restrict__ a, const void* __restrict__ b,
                                                     float* __restrict__ out, long long rows, int cols) {
    const long long row = blockIdx.x * 8LL + (threadIdx.x >> 5);
    if (row >= rows) return;
    float t = 0.f;
    for (int c = threadIdx.x & 31; c < cols; c += 32) t = fmaf(ld1<FA>(a, row * cols + c), ld1<FB>(b, row * cols + c), t);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xFFFFFFFFu, t, o);
    if ((threadIdx.x & 31) == 0) out[row] = t;
}

// dS[q][k] = P[q][k] * (dP[q][k] - D[q]) * scale for k < T, zero in the row padding
template <int FP, int FO>
__global__ void __launch_bounds__(256) attn_ds_kernel(const void* __restrict__ P, const float* __restrict__ dP,
                                                      const float* __restrict__ D, void* __restrict__ out, long long rows,
                                                      int T, long long tp, float scale) {
    const long long total = rows * tp;
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += 256LL * gridDim.x) {
        const long long q = i / tp;
        const int k = static_cast<int>(i - q * tp);
        st1<FO>(out, i, k < T ? ld1<FP>(P, i) * (dP[i] - D[q]) * scale : 0.f);
    }
}

// 16-bit mode: the same in 64x64 tiles, sixteen elements per thread, writing dS AND its transpose (the operand of dK) in
// one pass over P and dP.  tp % 64 == 0; rows / columns >= T of both outputs are zero (dS rows >= T are not written).
__global__ void __launch_bounds__(256) attn_ds_t_kernel(const __half* __restrict__ P, const float* __restrict__ dP,
                                                        const float* __restrict__ D, bf16* __restrict__ dS,
                                                        bf16* __restrict__ dST, int T, long long tp, float scale) {
    __shared__ uint32_t tile[64][33];
    const int q0 = blockIdx.y * 64, k0 = blockIdx.x * 64;
    const int row = threadIdx.x >> 2, seg = threadIdx.x & 3;
    const int q = q0 + row, kb = k0 + seg * 16;
    uint32_t w[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) w[j] = 0u;
    if (q < T) {
        const long long off = 1LL * q * tp + kb;
        const uint4 p0 = *reinterpret_cast<const uint4*>(P + off), p1 = *reinterpret_cast<const uint4*>(P + off + 8);
        float4 g[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) g[j] = *reinterpret_cast<const float4*>(dP + off + 4 * j);
        const float d = D[q];
        const uint32_t pw[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
        const float gv[16] = {g[0].x, g[0].y, g[0].z, g[0].w, g[1].x, g[1].y, g[1].z, g[1].w,
                              g[2].x, g[2].y, g[2].z, g[2].w, g[3].x, g[3].y, g[3].z, g[3].w};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float a = kb + 2 * j < T ? f16_lo(pw[j]) * (gv[2 * j] - d) * scale : 0.f;
            const float b = kb + 2 * j + 1 < T ? f16_hi(pw[j]) * (gv[2 * j + 1] - d) * scale : 0.f;
            w[j] = pack_bf16x2(a, b);
        }
        *reinterpret_cast<uint4*>(dS + off) = make_uint4(w[0], w[1], w[2], w[3]);
        *reinterpret_cast<uint4*>(dS + off + 8) = make_uint4(w[4], w[5], w[6], w[7]);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) tile[row][seg * 8 + j] = w[j];
    __syncthreads();
    // thread (kk, seg): output row k0 + kk, columns q0 + 16 seg .. + 15 = tile rows 16 seg .. + 15, column kk
    const int kk = threadIdx.x >> 2;
    if (k0 + kk < T) {
        uint32_t o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const uint32_t a = tile[seg * 16 + 2 * j][kk >> 1], b = tile[seg * 16 + 2 * j + 1][kk >> 1];
            o[j] = (kk & 1) ? ((a >> 16) | (b & 0xFFFF0000u)) : ((a & 0xFFFFu) | (b << 16));
        }
        bf16* dst = dST + 1LL * (k0 + kk) * tp + q0 + seg * 16;
        *reinterpret_cast<uint4*>(dst) = make_uint4(o[0], o[1], o[2], o[3]);
        *reinterpret_cast<uint4*>(dst + 8) = make_uint4(o[4], o[5], o[6], o[7]);
    }
}

// ---- stride-2 conv (pad right / bottom): data-gradient weights in the sub-pixel form of launch_conv's up2 mode.
// dX pixel (2y+py, 2x+px) = sum over the 2x2 source taps (ty, tx) of dY; row taps of parity 0: {y-1 <-> ky 2, y <-> ky 0},
// parity 1: {y <-> ky 1, y+1 <-> none}; columns alike.  dst[par][ci][(ty*2+tx)*Cout + co]
template <int OFMT>
__global__ void pack_dgrad_s2_kernel(const float* __restrict__ w /*[Cout][Cin][3][3]*/, void* __restrict__ dst, int Cout,
                                     int Cin) {
    const long long total = 4LL * Cin * 4 * Cout;
    for (long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < total; i += 1LL * gridDim.x * blockDim.x) {
        const int co = static_cast<int>(i % Cout);
        long long r = i / Cout;
        const int t = static_cast<int>(r % 4); r /= 4;
        const int ci = static_cast<int>(r % Cin);
        const int par = static_cast<int>(r / Cin);
        const int py = par >> 1, px = par & 1, ty = t >> 1, tx = t & 1;
        const int ky = py == 0 ? (ty == 0 ? 2 : 0) : (ty == 0 ? 1 : -1);
        const int kx = px == 0 ? (tx == 0 ? 2 : 0) : (tx == 0 ? 1 : -1);
        const float v = (ky >= 0 && kx >= 0) ? w[((1LL * co * Cin + ci) * 3 + ky) * 3 + kx] : 0.f;
        st1<OFMT>(dst, i, v);
    }
}
// fp32 verification mode: direct transposed conv.  dx[n][iy][ix][ci] = sum_{ky,kx,co} dy[n][(iy-ky)/2][(ix-kx)/2][co] w[co][ci][ky][kx]
__global__ void __launch_bounds__(128) f32_dgrad_s2_kernel(const float* __restrict__ dy, const float* __restrict__ w,
                                                           float* __restrict__ dx, int N, int Hi, int Wi, int Ho, int Wo,
                                                           int Cout, int Cin) {
    const long long pix = blockIdx.x;            // n*Hi*Wi + iy*Wi + ix
    const int ix = static_cast<int>(pix % Wi);
    const long long r = pix / Wi;
    const int iy = static_cast<int>(r % Hi);
    const long long n = r / Hi;
    for (int ci = threadIdx.x; ci < Cin; ci += blockDim.x) {
        float acc = 0.f;
        for (int ky = 0; ky < 3; ++ky) {
            const int ty = iy - ky;
            if (ty < 0 || (ty & 1) || ty / 2 >= Ho) continue;
            for (int kx = 0; kx < 3; ++kx) {
                const int tx = ix - kx;
                if (tx < 0 || (tx & 1) || tx / 2 >= Wo) continue;
                const float* g = dy + ((n * Ho + ty / 2) * Wo + tx / 2) * Cout;
                for (int co = 0; co < Cout; ++co) acc = fmaf(g[co], w[((1LL * co * Cin + ci) * 3 + ky) * 3 + kx], acc);
            }
        }
        dx[pix * Cin + ci] = acc;
    }
}

// ---- conv_in (3 -> Cout, 3x3 pad 1): weight gradient straight from the image.  One thread per output channel keeps the
// 27 sums of its row; a block walks a range of pixels with the 27 patch values of each pixel broadcast from shared memory.
template <int GF>
__global__ void __launch_bounds__(128) convin_wgrad_kernel(const void* __restrict__ dy, const void* __restrict__ img, int in_u8,
                                                           float* __restrict__ part, int N, int H, int W, int Cout) {
    __shared__ float patch[32][28];
    const long long P = 1LL * N * H * W;
    const long long chunk = (P + gridDim.x - 1) / gridDim.x;
    const long long p0 = blockIdx.x * chunk, p1 = min(P, p0 + chunk);
    float acc[27];
#pragma unroll
    for (int k = 0; k < 27; ++k) acc[k] = 0.f;
    for (long long pb = p0; pb < p1; pb += 32) {
        __syncthreads();
        for (int i = threadIdx.x; i < 32 * 27; i += blockDim.x) {
            const int j = i / 27, k = i - j * 27;
            const long long p = pb + j;
            float v = 0.f;
            if (p < p1) {
                const int x = static_cast<int>(p % W);
                const long long r = p / W;
                const int y = static_cast<int>(r % H);
                const long long n = r / H;
                const int tap = k / 3, c = k - tap * 3;
                const int ys = y + tap / 3 - 1, xs = x + tap % 3 - 1;
                if (ys >= 0 && ys < H && xs >= 0 && xs < W) {
                    if (in_u8) v = static_cast<const unsigned char*>(img)[((n * H + ys) * W + xs) * 3 + c] * (2.0f / 255.0f) - 1.0f;
                    else v = static_cast<const float*>(img)[((n * 3 + c) * H + ys) * W + xs];
                }
            }
            patch[j][k] = v;
        }
        __syncthreads();
        for (int co = threadIdx.x; co < Cout; co += blockDim.x) {   // Cout == blockDim.x in practice
            const int lim = static_cast<int>(min(32LL, p1 - pb));
            for (int j = 0; j < lim; ++j) {
                const float g = ld1<GF>(dy, (pb + j) * Cout + co);
#pragma unroll
                for (int k = 0; k < 27; ++k) acc[k] = fmaf(g, patch[j][k], acc[k]);
            }
        }
    }
    if (threadIdx.x < Cout)
#pragma unroll
        for (int k = 0; k < 27; ++k) part[(1LL * blockIdx.x * Cout + threadIdx.x) * 27 + k] = acc[k];
}
// part[chunk][co][tap*3 + c] -> dw[co][c][tap]
__global__ void convin_wgrad_reduce_kernel(const float* __restrict__ part, float* __restrict__ dw, int chunks, int Cout,
                                           int accumulate) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Cout * 27) return;
    const int co = i / 27, k = i - co * 27, tap = k / 3, c = k - tap * 3;
    double t = 0.0;
    for (int b = 0; b < chunks; ++b) t += part[(1LL * b * Cout + co) * 27 + k];
    float* o = dw + (co * 3 + c) * 9 + tap;
    *o = (accumulate ? *o : 0.f) + static_cast<float>(t);
}

// channel padding: out[row][Cp] = (in[row][C] | 0), fp32 NHWC source -> gradient format
template <int FO>
__global__ void pad_channels_kernel(const float* __restrict__ in, void* __restrict__ out, long long rows, int C, int Cp) {
    const long long total = rows * Cp;
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += 256LL * gridDim.x) {
        const long long r = i / Cp;
        const int c = static_cast<int>(i - r * Cp);
        st1<FO>(out, i, c < C ? in[r * C + c] : 0.f);
    }
}

}  // namespace

int bwd_transpose(const BwdEnv& e, const void* in, int in_fmt, void* out, int out_fmt, int rows, int cols, long long ld_in,
                  long long ld_out, int batch, long long in_bs, long long out_bs) {
    if (in_fmt != FMT_F32 && out_fmt != FMT_F32 && ld_in % 2 == 0 && ld_out % 2 == 0 && in_bs % 2 == 0 && out_bs % 2 == 0 &&
        reinterpret_cast<uintptr_t>(in) % 4 == 0 && reinterpret_cast<uintptr_t>(out) % 4 == 0 &&
        (in_fmt == out_fmt || (in_fmt == FMT_F16 && out_fmt == FMT_BF16))) {
        dim3 g16((rows + 63) / 64, (cols + 63) / 64, batch);
        const unsigned short* i16 = static_cast<const unsigned short*>(in);
        unsigned short* o16 = static_cast<unsigned short*>(out);
        if (in_fmt == out_fmt) transpose16_kernel<FMT_BF16, FMT_BF16><<<g16, 256, 0, e.s>>>(i16, o16, rows, cols, ld_in, ld_out, in_bs, out_bs);
        else transpose16_kernel<FMT_F16, FMT_BF16><<<g16, 256, 0, e.s>>>(i16, o16, rows, cols, ld_in, ld_out, in_bs, out_bs);
        VT_CUDA(cudaGetLastError());
        return 0;
    }
    dim3 grid((rows + 31) / 32, (cols + 31) / 32, batch);
#define VT_TR(FI, FO) transpose_kernel<FI, FO><<<grid, 256, 0, e.s>>>(in, out, rows, cols, ld_in, ld_out, in_bs, out_bs)
    if (in_fmt == FMT_F32 && out_fmt == FMT_F32) VT_TR(FMT_F32, FMT_F32);
    else if (in_fmt == FMT_F16 && out_fmt == FMT_BF16) VT_TR(FMT_F16, FMT_BF16);
    else if (in_fmt == FMT_BF16 && out_fmt == FMT_BF16) VT_TR(FMT_BF16, FMT_BF16);
    else if (in_fmt == FMT_F16 && out_fmt == FMT_F16) VT_TR(FMT_F16, FMT_F16);
    else { set_error("transpose: format pair not instantiated"); return -2; }
#undef VT_TR
    VT_CUDA(cudaGetLastError());
    return 0;
}

int bwd_rowdot(const BwdEnv& e, const void* a, int a_fmt, const void* b, int b_fmt, float* out, long long rows, int cols) {
    const unsigned grid = static_cast<unsigned>((rows + 7) / 8);
    if (a_fmt == FMT_F32 && b_fmt == FMT_F32) rowdot_kernel<FMT_F32, FMT_F32><<<grid, 256, 0, e.s>>>(a, b, out, rows, cols);
    else if (a_fmt == FMT_BF16 && b_fmt == FMT_F16) rowdot_kernel<FMT_BF16, FMT_F16><<<grid, 256, 0, e.s>>>(a, b, out, rows, cols);
    else { set_error("rowdot: format pair not instantiated"); return -2; }
    VT_CUDA(cudaGetLastError());
    return 0;
}

int bwd_attn_ds(const BwdEnv& e, const void* P, int p_fmt, const float* dP, const float* D, void* out, long long rows, int T,
                long long tp, float scale) {
    const int grid = grid_for(rows * tp);
    if (e.fp32) attn_ds_kernel<FMT_F32, FMT_F32><<<grid, 256, 0, e.s>>>(P, dP, D, out, rows, T, tp, scale);
    else if (p_fmt == FMT_F16) attn_ds_kernel<FMT_F16, FMT_BF16><<<grid, 256, 0, e.s>>>(P, dP, D, out, rows, T, tp, scale);
    else { set_error("attention dS: format not instantiated"); return -2; }
    VT_CUDA(cudaGetLastError());
    return 0;
}

// 16-bit mode: dS and dS^T ([tp][tp] each, bf16) in one pass; P fp16, dP fp32
int bwd_attn_ds_t(const BwdEnv& e, const void* P, int p_fmt, const float* dP, const float* D, void* dS, void* dST, int T, long long tp,
                  float scale) {
    VT_CHECK(!e.fp32 && p_fmt == FMT_F16 && tp % 64 == 0, "fused dS / dS^T pass: 16-bit mode, fp16 probabilities, padded to 64");
    dim3 grid(static_cast<unsigned>(tp / 64), static_cast<unsigned>((T + 63) / 64));
    attn_ds_t_kernel<<<grid, 256, 0, e.s>>>(static_cast<const __half*>(P), dP, D, static_cast<bf16*>(dS), static_cast<bf16*>(dST), T, tp,
                                            scale);
    VT_CUDA(cudaGetLastError());
    return 0;
}

int bwd_pad_channels(const BwdEnv& e, const float* in, void* out, long long rows, int C, int Cp) {
    if (e.fp32) pad_channels_kernel<FMT_F32><<<grid_for(rows * Cp), 256, 0, e.s>>>(in, out, rows, C, Cp);
    else pad_channels_kernel<FMT_BF16><<<grid_for(rows * Cp), 256, 0, e.s>>>(in, out, rows, C, Cp);
    VT_CUDA(cudaGetLastError());
    return 0;
}

// ---- stride-2 downsample conv (Hi x Wi -> Ho x Wo = Hi/2 x Wi/2, even sizes)
size_t bwd_dgrad_s2_weight_bytes(const BwdEnv& e, int Cout, int Cin) {
    return e.fp32 ? 0 : al(static_cast<size_t>(4) * Cin * 4 * Cout * 2);
}
int bwd_conv_s2_dgrad(const BwdEnv& e, const void* dy, const float* w, void* wd_scratch, void* dx, int N, int Hi, int Wi,
                      int Cout, int Cin) {
    VT_CHECK(Hi % 2 == 0 && Wi % 2 == 0, "stride-2 conv backward needs even input sizes");
    const int Ho = Hi / 2, Wo = Wi / 2;
    if (e.fp32) {
        profiler_begin(e.prof, KC_FP32, e.s, 2.0 * N * Ho * Wo * 9.0 * Cout * Cin, 0);
        f32_dgrad_s2_kernel<<<static_cast<unsigned>(1LL * N * Hi * Wi), 128, 0, e.s>>>(
            static_cast<const float*>(dy), w, static_cast<float*>(dx), N, Hi, Wi, Ho, Wo, Cout, Cin);
        profiler_end(e.prof, KC_FP32, e.s);
        VT_CUDA(cudaGetLastError());
        return 0;
    }
    pack_dgrad_s2_kernel<FMT_BF16><<<grid_for(16LL * Cin * Cout), 256, 0, e.s>>>(w, wd_scratch, Cout, Cin);
    VT_CUDA(cudaGetLastError());
    for (int par = 0; par < 4; ++par) {
        ConvOp op;
        op.in = dy; op.in_f16 = 0; op.raw_f16 = 0; op.N = N; op.Hin = Ho; op.Win = Wo; op.Cin = Cout; op.ksize = 3; op.stride = 1;
        op.w = static_cast<const bf16*>(wd_scratch) + static_cast<size_t>(par) * Cin * 4 * Cout; op.Cout = Cin;
        op.out = dx; op.out_fmt = FMT_BF16; op.up2 = 1; op.up_py = par >> 1; op.up_px = par & 1; op.kclass = KC_BWD;
        VT_TRY(launch_conv(op, e.s, e.prof));
    }
    return 0;
}
// fp32 verification mode (the 16-bit mode reads the stride-2 parity view straight from the input: bwd_conv_wgrad16)
int bwd_conv_s2_wgrad(const BwdEnv& e, const WgradPlan& p, const void* dy, const void* x, float* part, float* dw, int N, int Hi,
                      int Wi, int Cout, int Cin, int accumulate) {
    const int Ho = Hi / 2, Wo = Wi / 2;
    VT_CHECK(e.fp32, "bwd_conv_s2_wgrad is the fp32 verification path (16-bit mode: bwd_conv_wgrad16 with stride 2)");
    dim3 grid(Cout / 64, 9 * (Cin / 64), p.splits);
    profiler_begin(e.prof, KC_FP32, e.s, 2.0 * N * Ho * Wo * Cout * Cin * 9.0, 0);
    f32_wgrad_kernel<<<grid, 256, 0, e.s>>>(static_cast<const float*>(dy), static_cast<const float*>(x), part, N, Ho, Wo, Cout,
                                           Cin, 3, 2, Hi, Wi);
    profiler_end(e.prof, KC_FP32, e.s);
    wgrad_reduce_kernel<<<grid_for(9LL * Cout * Cin), 256, 0, e.s>>>(part, dw, p.batches, Cout, Cin, 9, accumulate);
    VT_CUDA(cudaGetLastError());
    return 0;
}

int bwd_convin_wgrad(const BwdEnv& e, const void* dy, const void* img, int in_u8, float* part, float* dw, int N, int H, int W,
                     int Cout, int accumulate) {
    VT_CHECK(Cout <= 128, "conv_in weight gradient: at most 128 output channels");
    const int chunks = bwd_convin_chunks(N, H, W);
    profiler_begin(e.prof, KC_MISC, e.s, 2.0 * N * H * W * 27.0 * Cout, 0);
    if (e.fp32) convin_wgrad_kernel<FMT_F32><<<chunks, 128, 0, e.s>>>(dy, img, in_u8, part, N, H, W, Cout);
    else convin_wgrad_kernel<FMT_BF16><<<chunks, 128, 0, e.s>>>(dy, img, in_u8, part, N, H, W, Cout);
    convin_wgrad_reduce_kernel<<<(Cout * 27 + 127) / 128, 128, 0, e.s>>>(part, dw, chunks, Cout, accumulate);
    profiler_end(e.prof, KC_MISC, e.s);
    VT_CUDA(cudaGetLastError());
    return 0;
}
int bwd_convin_chunks(int N, int H, int W) {
    return static_cast<int>(std::max<long long>(1, std::min<long long>((1LL * N * H * W + 255) / 256, 148 * 8)));
}


// fp16 -> bf16 copy (8 elements per thread): tcgen05 kind::f16 wants both operands of an MMA in ONE 16-bit format (a
// bf16 x fp16 pair traps -- tools/umma_mn_probe.cu), and the gradient side must be bf16 for its range
__global__ void __launch_bounds__(256) f16_to_bf16_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, long long n8) {
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n8; i += 256LL * gridDim.x) {
        const uint4 u = in[i];
        out[i] = make_uint4(pack_bf16x2(f16_lo(u.x), f16_hi(u.x)), pack_bf16x2(f16_lo(u.y), f16_hi(u.y)),
                            pack_bf16x2(f16_lo(u.z), f16_hi(u.z)), pack_bf16x2(f16_lo(u.w), f16_hi(u.w)));
    }
}

// ---- 16-bit weight gradient through the MN-major kernel (vt_wgrad.cu)
size_t bwd_wgrad16_a16_bytes(int N, int H, int W, int Cin, int stride) {
    return al(static_cast<size_t>(N) * stride * H * stride * W * Cin * 2);
}
int bwd_conv_wgrad16(const BwdEnv& e, const void* dy, const void* a_src, int a_fmt, const double* stats, const float* gamma,
                     const float* beta, float eps, int silu, void* a16, float* part, float* dw, int N, int H, int W, int Cout,
                     int Cin, int ks, int stride, int accumulate) {
    const void* a = a_src;
    const long long elems = 1LL * N * stride * H * stride * W * Cin;
    if (stats) {
        VT_CHECK(a16 != nullptr, "the normalised operand needs a scratch buffer");
        VT_TRY(launch_gn_apply(a_src, a_fmt, a16, FMT_BF16, stats, gamma, beta, N, 1LL * stride * H * stride * W, Cin, 32, eps, silu,
                               e.s, e.prof));
        a = a16;
    } else if (a_fmt == FMT_F16) {
        VT_CHECK(a16 != nullptr, "an fp16 operand needs a scratch buffer for its bf16 copy");
        f16_to_bf16_kernel<<<grid_for(elems / 8), 256, 0, e.s>>>(static_cast<const uint4*>(a_src), static_cast<uint4*>(a16), elems / 8);
        VT_CUDA(cudaGetLastError());
        a = a16;
    }
    a_fmt = FMT_BF16;
    const WgradMnPlan p = bwd_wgrad_mn_plan(N, H, W, Cout, Cin, ks);
    int splits = 0;
    VT_TRY(bwd_conv_wgrad_mn(e, p, dy, FMT_BF16, a, a_fmt, part, N, H, W, Cout, Cin, ks, stride, &splits));
    const int taps = ks * ks;
    wgrad_reduce_kernel<<<grid_for(1LL * Cout * Cin * taps), 256, 0, e.s>>>(part, dw, splits, Cout, Cin, taps, accumulate);
    VT_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace vt
