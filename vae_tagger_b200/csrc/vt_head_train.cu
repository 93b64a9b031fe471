// Training step of the tag-decoder head: train-mode forward + loss + analytic backward, fp32.
//
// Replaces, for one batch, the autograd graph of the reference step (train_decoder.py:186-195):
//     decoder.train(); logits = decoder(latent); loss = loss_fn(logits, labels); loss.backward()
// for AttentionClassificationDecoder (modules.py:358-468, cross-attention off) and
// ClassificationDecoder (modules.py:303-349).  The latent comes from the frozen encoder and needs no
// gradient, so the backward stops at the SpatialAttention weights.
//
// Train-mode semantics that differ from the inference kernels in vt_head.cu:
//   * BatchNorm2d uses the batch statistics over (N,H,W) (biased variance) and updates
//     running_mean / running_var (momentum, unbiased variance) and num_batches_tracked;
//   * Dropout: attention weights (p = attention_dropout, modules.py:81) and classifier (.3/.2/.1,
//     modules.py:405-415) masks come from a counter-based generator (seed, mask stream, element), so
//     the backward regenerates them instead of storing them.
// Gradients are ACCUMULATED into one flat fp32 buffer laid out like decoder.parameters()
// (head_param_layout) -- the buffer the data-parallel step all-reduces with NCCL.  Every reduction
// over the batch is two-stage (per-CTA partials, then one sum in a fixed order): results are
// bit-reproducible run to run.
//
// The work is ~1 MB of latent per image and a 1.4 M-parameter MLP: everything is L2 resident and
// latency/HBM bound, so these are plain CUDA-core kernels.
#include <vector>

#include "../../include/vae_tagger_b200.h"
#include "vt_head_common.cuh"
#include "vt_head_train.h"
#include "vt_internal.h"

namespace vt {

namespace {

__device__ __forceinline__ double h_warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    return v;
}
__device__ __forceinline__ double block_sum256_d(double v, double* red) {
    v = h_warp_sum_d(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i];
    return t;
}

// Sums NV per-thread values over the 256 threads of a CTA with ONE shared-memory exchange (warp shuffles
// first): thread k < NV returns the total of value k, other threads return 0.  sm: 8*NV floats.
template <int NV>
__device__ __forceinline__ float block_sum_many(float (&v)[NV], float* sm) {
    static_assert(NV <= 256, "one result per thread");
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < NV; ++k) v[k] = h_warp_sum(v[k]);
    __syncthreads();
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) sm[warp * NV + k] = v[k];
    }
    __syncthreads();
    float t = 0.f;
    if (threadIdx.x < NV) {
#pragma unroll
        for (int w = 0; w < 8; ++w) t += sm[w * NV + threadIdx.x];
    }
    return t;
}

// ------------------------------------------------------------------------------ feature_compress
// conv3x3 C->CO (+bias) -> z[N][CO][HW], and per-CTA partial (sum, sumsq) of z per channel for the
// batch statistics.  grid (ceil(HW/256), N); part[(n*gridDim.x+blockIdx.x)][CO][2] doubles.
template <int CO>
__global__ void __launch_bounds__(256) head_conv_train_kernel(const float* __restrict__ x,   // [N][C][H][W]
                                                              const float* __restrict__ cw,  // [CO][C][3][3]
                                                              const float* __restrict__ cb, float* __restrict__ z,
                                                              double* __restrict__ part, int C, int H, int W) {
    extern __shared__ float sw[];  // CO*C*9 weights, then 8*2*CO floats scratch
    float* red = sw + CO * C * 9;
    const int n = blockIdx.y, HW = H * W;
    for (int i = threadIdx.x; i < CO * C * 9; i += 256) sw[i] = cw[i];
    __syncthreads();
    const int p = blockIdx.x * 256 + threadIdx.x;
    const bool live = p < HW;
    float v[CO];
#pragma unroll
    for (int o = 0; o < CO; ++o) v[o] = cb[o];
    if (live) {
        const int py = p / W, px = p - py * W;
        for (int c = 0; c < C; ++c) {
            const float* xp = x + (1LL * n * C + c) * HW;
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
                const int yy = py + ky - 1;
                if (yy < 0 || yy >= H) continue;
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    const int xx = px + kx - 1;
                    if (xx < 0 || xx >= W) continue;
                    const float xv = xp[yy * W + xx];
#pragma unroll
                    for (int o = 0; o < CO; ++o) v[o] = fmaf(sw[((o * C + c) * 3 + ky) * 3 + kx], xv, v[o]);
                }
            }
        }
    }
    float st[2 * CO];
#pragma unroll
    for (int o = 0; o < CO; ++o) {
        const float t = live ? v[o] : 0.f;
        if (live) z[(1LL * n * CO + o) * HW + p] = t;
        st[2 * o] = t;
        st[2 * o + 1] = t * t;
    }
    const float tot = block_sum_many<2 * CO>(st, red);
    if (threadIdx.x < 2 * CO)
        part[(1LL * n * gridDim.x + blockIdx.x) * CO * 2 + threadIdx.x] = static_cast<double>(tot);
}

// batch statistics from the partials (one CTA): stat[0..Co) = mean, stat[Co..2Co) = 1/sqrt(var+eps);
// running buffers updated like nn.BatchNorm2d in train mode (momentum, unbiased variance).
__global__ void __launch_bounds__(256) head_bn_finalize_kernel(const double* __restrict__ part, int nparts, int Co,
                                                               double count, float eps, float momentum,
                                                               float* __restrict__ stat, float* __restrict__ rmean,
                                                               float* __restrict__ rvar,
                                                               long long* __restrict__ tracked) {
    __shared__ double red[8];
    for (int o = 0; o < Co; ++o) {
        double s1 = 0.0, s2 = 0.0;
        for (int i = threadIdx.x; i < nparts; i += 256) {
            s1 += part[(1LL * i * Co + o) * 2];
            s2 += part[(1LL * i * Co + o) * 2 + 1];
        }
        s1 = block_sum256_d(s1, red);
        s2 = block_sum256_d(s2, red);
        if (threadIdx.x == 0) {
            const double mean = s1 / count;
            double var = s2 / count - mean * mean;
            if (var < 0.0) var = 0.0;
            stat[o] = static_cast<float>(mean);
            stat[Co + o] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
            if (rmean) rmean[o] = (1.0f - momentum) * rmean[o] + momentum * static_cast<float>(mean);
            if (rvar) {
                const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
                rvar[o] = (1.0f - momentum) * rvar[o] + momentum * static_cast<float>(unbiased);
            }
        }
    }
    if (threadIdx.x == 0 && tracked) *tracked += 1;
}

// BatchNorm (batch statistics) -> ReLU -> AdaptiveAvgPool(8,8).  grid (64 cells, N).
__global__ void __launch_bounds__(256) head_bn_relu_pool_kernel(const float* __restrict__ z,
                                                                const float* __restrict__ stat,
                                                                const float* __restrict__ bn_w,
                                                                const float* __restrict__ bn_b,
                                                                float* __restrict__ pooled, int Co, int H, int W) {
    __shared__ float red[8];
    const int n = blockIdx.y, cell = blockIdx.x, HW = H * W;
    const int cy = cell / 8, cx = cell % 8;
    const int y0 = (cy * H) / 8, y1 = ((cy + 1) * H + 7) / 8;
    const int x0 = (cx * W) / 8, x1 = ((cx + 1) * W + 7) / 8;
    const int wh = y1 - y0, ww = x1 - x0;
    for (int o = 0; o < Co; ++o) {
        const float mean = stat[o], a = stat[Co + o] * bn_w[o], b = bn_b[o];
        const float* zp = z + (1LL * n * Co + o) * HW;
        float acc = 0.f;
        for (int i = threadIdx.x; i < wh * ww; i += 256) {
            const float t = (zp[(y0 + i / ww) * W + x0 + i % ww] - mean) * a + b;
            acc += fmaxf(t, 0.f);
        }
        const float t = block_sum256(acc, red);
        if (threadIdx.x == 0) pooled[(1LL * n * Co + o) * 64 + cell] = t / static_cast<float>(wh * ww);
    }
}

// backward of AdaptiveAvgPool + ReLU into dt (written to dz), and the two per-channel sums the
// BatchNorm backward needs: part[blk][Co][2] = (sum dt, sum dt*xhat).  grid (ceil(HW/256), N).
__global__ void __launch_bounds__(256) head_bn_bwd_stats_kernel(const float* __restrict__ z,
                                                                const float* __restrict__ stat,
                                                                const float* __restrict__ bn_w,
                                                                const float* __restrict__ bn_b,
                                                                const float* __restrict__ dpooled,  // [N][Co][64]
                                                                float* __restrict__ dz, double* __restrict__ part,
                                                                int Co, int H, int W) {
    __shared__ float red[8 * 32];
    const int n = blockIdx.y, HW = H * W;
    const int p = blockIdx.x * 256 + threadIdx.x;
    const bool live = p < HW;
    // the (up to four) pooling windows that contain this pixel
    int cells[4];
    float inv_area[4];
    int ncell = 0;
    if (live) {
        const int py = p / W, px = p - py * W;
        for (int cy = 0; cy < 8; ++cy) {
            const int y0 = (cy * H) / 8, y1 = ((cy + 1) * H + 7) / 8;
            if (py < y0 || py >= y1) continue;
            for (int cx = 0; cx < 8; ++cx) {
                const int x0 = (cx * W) / 8, x1 = ((cx + 1) * W + 7) / 8;
                if (px < x0 || px >= x1) continue;
                if (ncell < 4) {
                    cells[ncell] = cy * 8 + cx;
                    inv_area[ncell] = 1.0f / static_cast<float>((y1 - y0) * (x1 - x0));
                    ++ncell;
                }
            }
        }
    }
    float st[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) st[i] = 0.f;
#pragma unroll
    for (int o = 0; o < 16; ++o) {
        if (o < Co && live) {
            const long long idx = (1LL * n * Co + o) * HW + p;
            const float xhat = (z[idx] - stat[o]) * stat[Co + o];
            const float t = xhat * bn_w[o] + bn_b[o];
            float dr = 0.f;
            for (int i = 0; i < ncell; ++i) dr += dpooled[(1LL * n * Co + o) * 64 + cells[i]] * inv_area[i];
            const float dt = t > 0.f ? dr : 0.f;
            dz[idx] = dt;
            st[2 * o] = dt;
            st[2 * o + 1] = dt * xhat;
        }
    }
    const float tot = block_sum_many<32>(st, red);
    if (threadIdx.x < 2 * Co)
        part[(1LL * n * gridDim.x + blockIdx.x) * Co * 2 + threadIdx.x] = static_cast<double>(tot);
}

// sums the partials (one CTA): sums[0..Co) = sum dt, sums[Co..2Co) = sum dt*xhat; BatchNorm affine grads
__global__ void __launch_bounds__(256) head_bn_bwd_finalize_kernel(const double* __restrict__ part, int nparts,
                                                                   int Co, float* __restrict__ sums,
                                                                   float* __restrict__ g_bn_w,
                                                                   float* __restrict__ g_bn_b) {
    __shared__ double red[8];
    for (int o = 0; o < Co; ++o) {
        double s1 = 0.0, s2 = 0.0;
        for (int i = threadIdx.x; i < nparts; i += 256) {
            s1 += part[(1LL * i * Co + o) * 2];
            s2 += part[(1LL * i * Co + o) * 2 + 1];
        }
        s1 = block_sum256_d(s1, red);
        s2 = block_sum256_d(s2, red);
        if (threadIdx.x == 0) {
            sums[o] = static_cast<float>(s1);
            sums[Co + o] = static_cast<float>(s2);
            g_bn_b[o] += static_cast<float>(s1);
            g_bn_w[o] += static_cast<float>(s2);
        }
    }
}

// dz = w*invstd*(dt - mean(dt) - xhat*mean(dt*xhat)), in place over dz (which holds dt)
__global__ void __launch_bounds__(256) head_bn_bwd_apply_kernel(const float* __restrict__ z,
                                                                const float* __restrict__ stat,
                                                                const float* __restrict__ sums,
                                                                const float* __restrict__ bn_w,
                                                                float* __restrict__ dz, int Co, int HW,
                                                                long long total, float inv_count) {
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += 256LL * gridDim.x) {
        const int o = static_cast<int>((i / HW) % Co);
        const float xhat = (z[i] - stat[o]) * stat[Co + o];
        dz[i] = bn_w[o] * stat[Co + o] * (dz[i] - sums[o] * inv_count - xhat * sums[Co + o] * inv_count);
    }
}

// conv3x3 weight gradient: dW[o][c][k] = sum_{n,p} dz[n][o][p] * x[n][c][p+k].  grid (C, bands, N):
// one input channel and one band of rows per CTA, CO*9 accumulators per thread, block reduction,
// partial[(n*bands+band)][o][c][k]; the c == 0 CTAs also reduce the bias gradient sum dz.
template <int CO>
__global__ void __launch_bounds__(256) head_conv_bwd_w_kernel(const float* __restrict__ x,   // [N][C][H][W]
                                                              const float* __restrict__ dz,  // [N][CO][H][W]
                                                              float* __restrict__ part_w, float* __restrict__ part_b,
                                                              int C, int H, int W) {
    __shared__ float red[8 * CO * 9];
    const int c = blockIdx.x, band = blockIdx.y, bands = gridDim.y, n = blockIdx.z;
    const int HW = H * W;
    const int r0 = (band * H) / bands, r1 = ((band + 1) * H) / bands;
    const float* xp = x + (1LL * n * C + c) * HW;
    const float* dzp = dz + 1LL * n * CO * HW;
    float acc[CO * 9];
    float bacc[CO];
#pragma unroll
    for (int i = 0; i < CO * 9; ++i) acc[i] = 0.f;
#pragma unroll
    for (int o = 0; o < CO; ++o) bacc[o] = 0.f;
    for (int i = threadIdx.x; i < (r1 - r0) * W; i += 256) {
        const int py = r0 + i / W, px = i % W;
        float d[CO];
#pragma unroll
        for (int o = 0; o < CO; ++o) {
            d[o] = dzp[1LL * o * HW + py * W + px];
            bacc[o] += d[o];
        }
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            const int yy = py + ky - 1;
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const int xx = px + kx - 1;
                const float xv = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? xp[yy * W + xx] : 0.f;
#pragma unroll
                for (int o = 0; o < CO; ++o) acc[o * 9 + ky * 3 + kx] = fmaf(d[o], xv, acc[o * 9 + ky * 3 + kx]);
            }
        }
    }
    const long long pidx = 1LL * n * bands + band;
    const float tw = block_sum_many<CO * 9>(acc, red);
    if (threadIdx.x < CO * 9) {  // value index = o*9 + k
        const int o = threadIdx.x / 9, k = threadIdx.x % 9;
        part_w[(pidx * CO + o) * C * 9 + c * 9 + k] = tw;
    }
    if (c == 0) {
        const float tb = block_sum_many<CO>(bacc, red);
        if (threadIdx.x < CO) part_b[pidx * CO + threadIdx.x] = tb;
    }
}

// conv3x3 input gradient: dx[n][c][p] = sum_{o,k} dz[n][o][p-k] * W[o][c][k].  grid (ceil(HW/256), N)
template <int CO>
__global__ void __launch_bounds__(256) head_conv_bwd_x_kernel(const float* __restrict__ dz,
                                                              const float* __restrict__ cw,  // [CO][2CO][3][3]
                                                              float* __restrict__ dx, int H, int W) {
    constexpr int C = 2 * CO;
    __shared__ float sw[CO * C * 9];
    for (int i = threadIdx.x; i < CO * C * 9; i += 256) sw[i] = cw[i];
    __syncthreads();
    const int n = blockIdx.y, HW = H * W;
    const int p = blockIdx.x * 256 + threadIdx.x;
    if (p >= HW) return;
    const int py = p / W, px = p - py * W;
    float acc[C];
#pragma unroll
    for (int c = 0; c < C; ++c) acc[c] = 0.f;
    for (int o = 0; o < CO; ++o) {
        const float* dp = dz + (1LL * n * CO + o) * HW;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            const int yy = py - (ky - 1);
            if (yy < 0 || yy >= H) continue;
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const int xx = px - (kx - 1);
                if (xx < 0 || xx >= W) continue;
                const float d = dp[yy * W + xx];
#pragma unroll
                for (int c = 0; c < C; ++c) acc[c] = fmaf(d, sw[((o * C + c) * 3 + ky) * 3 + kx], acc[c]);
            }
        }
    }
#pragma unroll
    for (int c = 0; c < C; ++c) dx[(1LL * n * C + c) * HW + p] = acc[c];
}

// dst[i] += sum over parts of part[p][i]  (fixed order: deterministic)
__global__ void __launch_bounds__(256) head_partial_reduce_kernel(const float* __restrict__ part, int nparts, int P,
                                                                  float* __restrict__ dst) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= P) return;
    float s = 0.f;
    for (int p = 0; p < nparts; ++p) s += part[1LL * p * P + i];
    dst[i] += s;
}

// ------------------------------------------------------------------------------ SpatialAttention bwd
// y = x*g[c]*s[p] (modules.py:36-47).  Step A: dpre[n][p] = s(1-s) * sum_c dy[c][p]*x[c][p]*g[c]
__global__ void __launch_bounds__(256) head_sa_bwd_pre_kernel(const float* __restrict__ x,
                                                              const float* __restrict__ cgate,
                                                              const float* __restrict__ sgate,
                                                              const float* __restrict__ dy,
                                                              float* __restrict__ dpre, int C, int HW) {
    __shared__ float g[64];
    const int n = blockIdx.y;
    if (threadIdx.x < C) g[threadIdx.x] = cgate[1LL * n * C + threadIdx.x];
    __syncthreads();
    for (int p = blockIdx.x * 256 + threadIdx.x; p < HW; p += 256 * gridDim.x) {
        float ds = 0.f;
        for (int c = 0; c < C; ++c) {
            const long long o = (1LL * n * C + c) * HW + p;
            ds = fmaf(dy[o] * x[o], g[c], ds);
        }
        const float s = sgate[1LL * n * HW + p];
        dpre[1LL * n * HW + p] = ds * s * (1.0f - s);
    }
}

// Step B: gradient of the channel gate.  dm = convT7x7(dpre); dxg[c] = dy[c]*s + dm0/C + dm1*[c==argmax];
// dg[n][c] = sum_p dxg[c]*x[c].  grid (bands, N); part_g[(n*bands+band)][C]
__global__ void __launch_bounds__(256) head_sa_bwd_gate_sum_kernel(const float* __restrict__ x,
                                                                   const float* __restrict__ cgate,
                                                                   const float* __restrict__ sgate,
                                                                   const float* __restrict__ dy,
                                                                   const float* __restrict__ dpre,
                                                                   const float* __restrict__ w7,
                                                                   float* __restrict__ part_g, int C, int H, int W) {
    __shared__ float ws[98];
    __shared__ float g[64];
    __shared__ float red[8 * 32];
    const int n = blockIdx.y, HW = H * W;
    if (threadIdx.x < 98) ws[threadIdx.x] = w7[threadIdx.x];
    if (threadIdx.x < C) g[threadIdx.x] = cgate[1LL * n * C + threadIdx.x];
    __syncthreads();
    float acc[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) acc[c] = 0.f;
    const float* dp = dpre + 1LL * n * HW;
    const float inv_c = 1.0f / static_cast<float>(C);
    for (int p = blockIdx.x * 256 + threadIdx.x; p < HW; p += 256 * gridDim.x) {
        const int py = p / W, px = p - py * W;
        float dm0 = 0.f, dm1 = 0.f;
        for (int ky = 0; ky < 7; ++ky) {
            const int yy = py - (ky - 3);
            if (yy < 0 || yy >= H) continue;
            for (int kx = 0; kx < 7; ++kx) {
                const int xx = px - (kx - 3);
                if (xx < 0 || xx >= W) continue;
                const float d = dp[yy * W + xx];
                dm0 = fmaf(d, ws[ky * 7 + kx], dm0);
                dm1 = fmaf(d, ws[49 + ky * 7 + kx], dm1);
            }
        }
        // channel of the maximum of x*g (first one wins, like torch.max)
        int amax = 0;
        float best = -CUDART_INF_F;
        for (int c = 0; c < C; ++c) {
            const float v = x[(1LL * n * C + c) * HW + p] * g[c];
            if (v > best) {
                best = v;
                amax = c;
            }
        }
        const float s = sgate[1LL * n * HW + p];
#pragma unroll
        for (int c = 0; c < 32; ++c) {
            if (c < C) {
                const long long o = (1LL * n * C + c) * HW + p;
                const float dxg = dy[o] * s + dm0 * inv_c + (c == amax ? dm1 : 0.f);
                acc[c] = fmaf(dxg, x[o], acc[c]);
            }
        }
    }
    const float tot = block_sum_many<32>(acc, red);
    if (threadIdx.x < C) part_g[(1LL * n * gridDim.x + blockIdx.x) * C + threadIdx.x] = tot;
}

// Step C: 7x7 weight gradient dW7[ci][k] = sum_{n,p} dpre[p]*map2[ci][p+k].  grid (bands, N, 2):
// one map channel per CTA, 49 accumulators per thread; part_w7[(n*bands+band)][98]
__global__ void __launch_bounds__(256) head_sa_bwd_w7_kernel(const float* __restrict__ map2,
                                                             const float* __restrict__ dpre,
                                                             float* __restrict__ part_w7, int H, int W) {
    __shared__ float red[8 * 49];
    const int n = blockIdx.y, ci = blockIdx.z, HW = H * W;
    const float* mp = map2 + (1LL * n * 2 + ci) * HW;
    const float* dp = dpre + 1LL * n * HW;
    float acc[49];
#pragma unroll
    for (int i = 0; i < 49; ++i) acc[i] = 0.f;
    for (int p = blockIdx.x * 256 + threadIdx.x; p < HW; p += 256 * gridDim.x) {
        const int py = p / W, px = p - py * W;
        const float d = dp[p];
#pragma unroll
        for (int ky = 0; ky < 7; ++ky) {
            const int yy = py + ky - 3;
#pragma unroll
            for (int kx = 0; kx < 7; ++kx) {
                const int xx = px + kx - 3;
                const float m = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? mp[yy * W + xx] : 0.f;
                acc[ky * 7 + kx] = fmaf(d, m, acc[ky * 7 + kx]);
            }
        }
    }
    const float tot = block_sum_many<49>(acc, red);
    if (threadIdx.x < 49) part_w7[(1LL * n * gridDim.x + blockIdx.x) * 98 + ci * 49 + threadIdx.x] = tot;
}

// Step D (one CTA): channel-gate MLP backward, summed over the batch in image order.
//   gate = sigmoid(W2 relu(W1 avg) + W2 relu(W1 max)); only the weights get a gradient.
__global__ void __launch_bounds__(256) head_sa_bwd_mlp_kernel(const float* __restrict__ pool,   // [N][C][2]
                                                              const float* __restrict__ cgate,  // [N][C]
                                                              const float* __restrict__ part_g, int bands,
                                                              const float* __restrict__ w1,  // [Ch][C]
                                                              const float* __restrict__ w2,  // [C][Ch]
                                                              float* __restrict__ g_w1, float* __restrict__ g_w2,
                                                              int N, int C, int Ch) {
    __shared__ float dpg[64];
    __shared__ float hid[2][16];
    __shared__ float dh[2][16];
    const int t = threadIdx.x;
    float acc1 = 0.f, acc2 = 0.f;  // this thread's element of dW1 ([j][c]) and dW2 ([c][j])
    for (int n = 0; n < N; ++n) {
        if (t < C) {
            float dg = 0.f;
            for (int b = 0; b < bands; ++b) dg += part_g[(1LL * n * bands + b) * C + t];
            const float g = cgate[1LL * n * C + t];
            dpg[t] = dg * g * (1.0f - g);
        }
        if (t >= 64 && t < 64 + 2 * Ch) {
            const int which = (t - 64) / Ch, j = (t - 64) % Ch;
            float a = 0.f;
            for (int c = 0; c < C; ++c) a = fmaf(w1[j * C + c], pool[(1LL * n * C + c) * 2 + which], a);
            hid[which][j] = a;
        }
        __syncthreads();
        if (t < C * Ch) {  // dW2[c][j]
            const int c = t / Ch, j = t % Ch;
            acc2 += dpg[c] * (fmaxf(hid[0][j], 0.f) + fmaxf(hid[1][j], 0.f));
        }
        if (t >= 64 && t < 64 + 2 * Ch) {
            const int which = (t - 64) / Ch, j = (t - 64) % Ch;
            float a = 0.f;
            for (int c = 0; c < C; ++c) a = fmaf(w2[c * Ch + j], dpg[c], a);
            dh[which][j] = hid[which][j] > 0.f ? a : 0.f;
        }
        __syncthreads();
        if (t < Ch * C) {  // dW1[j][c]
            const int j = t / C, c = t % C;
            acc1 += dh[0][j] * pool[(1LL * n * C + c) * 2] + dh[1][j] * pool[(1LL * n * C + c) * 2 + 1];
        }
        __syncthreads();
    }
    if (t < C * Ch) {
        g_w1[t] += acc1;
        g_w2[t] += acc2;
    }
}

// ------------------------------------------------------------------------------ self-attention
// MultiHeadSelfAttention in train mode (modules.py:66-91), one thread per token.
// prm = q.w q.b k.w k.b v.w v.b out.w out.b norm.w norm.b, contiguous (the flat parameter order).
// The per-head part (scores, softmax, dropout, P.V and their backward) runs as one CTA per (head, image):
// 8x the parallelism of a CTA per image; the cheap per-image parts (out_proj, projections / LayerNorm
// backward, parameter-gradient sums) are separate one-CTA-per-image kernels.
template <class F>
__device__ __forceinline__ void token_reduce(float (*sm)[65], int t, int count, float* dst, F contrib) {
    for (int base = 0; base < count; base += 64) {
        __syncthreads();
        for (int i = 0; i < 64 && base + i < count; ++i) sm[t][i] = contrib(base + i);
        __syncthreads();
        if (base + t < count) {
            float s = 0.f;
            for (int r = 0; r < 64; ++r) s += sm[r][t];
            dst[base + t] = s;
        }
    }
}

// LayerNorm of token t of image n (eps 1e-5): raw input, normalised value, affine output
__device__ __forceinline__ float mhsa_token_ln(const float* __restrict__ pooled, const float* __restrict__ ln_w,
                                               const float* __restrict__ ln_b, int n, int E, int t, float* xin,
                                               float* xhat, float* xn) {
    for (int e = 0; e < E; ++e) xin[e] = pooled[(1LL * n * E + e) * 64 + t];
    float mean = 0.f;
    for (int e = 0; e < E; ++e) mean += xin[e];
    mean /= static_cast<float>(E);
    float var = 0.f;
    for (int e = 0; e < E; ++e) var += (xin[e] - mean) * (xin[e] - mean);
    var /= static_cast<float>(E);
    const float rstd = 1.0f / sqrtf(var + 1e-5f);
    for (int e = 0; e < E; ++e) {
        xhat[e] = (xin[e] - mean) * rstd;
        xn[e] = xhat[e] * ln_w[e] + ln_b[e];
    }
    return rstd;
}

// One head of one image: grid (heads, N), one thread per query token.
//   BWD = false: ao[n][t][h*hd+d] = dropout(softmax(q k^T / sqrt(hd))) v           (attention output)
//   BWD = true : recomputes the forward of this head, then dqkv[n][{q,k,v}][t][h*hd+d]
template <bool BWD>
__global__ void __launch_bounds__(64) head_mhsa_head_kernel(const float* __restrict__ pooled,  // [N][E][64]
                                                            const float* __restrict__ prm,
                                                            float* __restrict__ ao,            // [N][64][E]
                                                            const float* __restrict__ dfeat,   // [N][E*64]
                                                            float* __restrict__ dqkv,          // [N][3][64][E]
                                                            int E, int heads, float drop_p,
                                                            unsigned long long seed) {
    __shared__ float sk[64][17];
    __shared__ float sv[64][17];
    __shared__ float sq[64][17];
    __shared__ float sd[64][17];
    __shared__ float sm[64][65];
    const int h = blockIdx.x, n = blockIdx.y, t = threadIdx.x;
    const int EE = E * E, hd = E / heads, c0 = h * hd;
    const float *wq = prm, *bq = wq + EE, *wk = bq + E, *bk = wk + EE, *wv = bk + E, *bv = wv + EE, *wo = bv + E,
                *ln_w = wo + EE + E, *ln_b = ln_w + E;
    float xin[16], xhat[16], xn[16], q[16];
    mhsa_token_ln(pooled, ln_w, ln_b, n, E, t, xin, xhat, xn);
    for (int d = 0; d < hd; ++d) {
        const int o = c0 + d;
        float a = bq[o], b = bk[o], c = bv[o];
        for (int e = 0; e < E; ++e) {
            a = fmaf(wq[o * E + e], xn[e], a);
            b = fmaf(wk[o * E + e], xn[e], b);
            c = fmaf(wv[o * E + e], xn[e], c);
        }
        q[d] = a;
        sk[t][d] = b;
        sv[t][d] = c;
        sq[t][d] = a;
    }
    float dao[16];
    if (BWD) {
        for (int d = 0; d < hd; ++d) {
            float a = 0.f;
            for (int o = 0; o < E; ++o) a = fmaf(dfeat[1LL * n * E * 64 + o * 64 + t], wo[o * E + c0 + d], a);
            dao[d] = a;
            sd[t][d] = a;
        }
    }
    __syncthreads();
    const float scale = 1.0f / sqrtf(static_cast<float>(hd));
    const float inv_keep = drop_p > 0.f ? 1.0f / (1.0f - drop_p) : 1.0f;
    float P[64];
    float m = -CUDART_INF_F;
#pragma unroll
    for (int j = 0; j < 64; ++j) {
        float sc = 0.f;
        for (int d = 0; d < hd; ++d) sc = fmaf(q[d], sk[j][d], sc);
        sc *= scale;
        P[j] = sc;
        m = fmaxf(m, sc);
    }
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < 64; ++j) {
        P[j] = expf(P[j] - m);
        sum += P[j];
    }
    const float inv = 1.0f / sum;
    unsigned long long keep = ~0ULL;
    if (drop_p > 0.f) {
        keep = 0ULL;
        const unsigned long long base = ((1ULL * n * heads + h) * 64 + t) * 64;
#pragma unroll
        for (int j = 0; j < 64; ++j)
            if (h_dropout_keep(seed, 0u, base + j, drop_p)) keep |= 1ULL << j;
    }
#pragma unroll
    for (int j = 0; j < 64; ++j) P[j] *= inv;
    if (!BWD) {
        for (int d = 0; d < hd; ++d) {
            float o = 0.f;
#pragma unroll
            for (int j = 0; j < 64; ++j)
                if ((keep >> j) & 1ULL) o = fmaf(P[j] * inv_keep, sv[j][d], o);
            ao[(1LL * n * 64 + t) * E + c0 + d] = o;
        }
        return;
    }
    // dP_j = keep_j * inv_keep * sum_d dao[d]*v[j][d];  dS_j = P_j (dP_j - sum_j' dP_j' P_j') * scale
    float rowdot = 0.f;
#pragma unroll
    for (int j = 0; j < 64; ++j) {
        if ((keep >> j) & 1ULL) {
            float dp = 0.f;
            for (int d = 0; d < hd; ++d) dp = fmaf(dao[d], sv[j][d], dp);
            rowdot = fmaf(dp * inv_keep, P[j], rowdot);
        }
    }
#pragma unroll
    for (int j = 0; j < 64; ++j) sm[t][j] = ((keep >> j) & 1ULL) ? P[j] * inv_keep : 0.f;
    __syncthreads();
    float* dq_out = dqkv + ((1LL * n * 3 + 0) * 64 + t) * E + c0;
    float* dk_out = dqkv + ((1LL * n * 3 + 1) * 64 + t) * E + c0;
    float* dv_out = dqkv + ((1LL * n * 3 + 2) * 64 + t) * E + c0;
    for (int d = 0; d < hd; ++d) {  // dv of token t as a key: column t of the dropped weights
        float a = 0.f;
        for (int r = 0; r < 64; ++r) a = fmaf(sm[r][t], sd[r][d], a);
        dv_out[d] = a;
    }
    __syncthreads();
    float dq[16];
    for (int d = 0; d < hd; ++d) dq[d] = 0.f;
#pragma unroll
    for (int j = 0; j < 64; ++j) {
        float dp = 0.f;
        if ((keep >> j) & 1ULL) {
            for (int d = 0; d < hd; ++d) dp = fmaf(dao[d], sv[j][d], dp);
            dp *= inv_keep;
        }
        const float ds = P[j] * (dp - rowdot) * scale;
        sm[t][j] = ds;
        for (int d = 0; d < hd; ++d) dq[d] = fmaf(ds, sk[j][d], dq[d]);
    }
    __syncthreads();
    for (int d = 0; d < hd; ++d) {  // dk of token t as a key
        float a = 0.f;
        for (int r = 0; r < 64; ++r) a = fmaf(sm[r][t], sq[r][d], a);
        dk_out[d] = a;
        dq_out[d] = dq[d];
    }
}

// out_proj + residual (modules.py:86): feat[n][o*64+t] = bo[o] + sum_e Wo[o][e] ao[n][t][e] + x[n][o][t]
__global__ void __launch_bounds__(64) head_mhsa_out_kernel(const float* __restrict__ pooled,
                                                           const float* __restrict__ prm,
                                                           const float* __restrict__ ao, float* __restrict__ feat,
                                                           int E) {
    const int n = blockIdx.x, t = threadIdx.x;
    const int EE = E * E;
    const float *wo = prm + 3 * (EE + E), *bo = wo + EE;
    float a[16];
    for (int e = 0; e < E; ++e) a[e] = ao[(1LL * n * 64 + t) * E + e];
    for (int o = 0; o < E; ++o) {
        float r = bo[o];
        for (int e = 0; e < E; ++e) r = fmaf(wo[o * E + e], a[e], r);
        feat[1LL * n * E * 64 + o * 64 + t] = r + pooled[(1LL * n * E + o) * 64 + t];
    }
}

// The per-image tail of the backward: projections, LayerNorm and residual -> dpooled[n][e][t], and the
// parameter gradients of this image (sums over its 64 tokens) -> partial[n][4(E*E+E)+2E], prm order.
__global__ void __launch_bounds__(64) head_mhsa_bwd_tail_kernel(const float* __restrict__ pooled,
                                                                const float* __restrict__ prm,
                                                                const float* __restrict__ ao,
                                                                const float* __restrict__ dqkv,
                                                                const float* __restrict__ dfeat,
                                                                float* __restrict__ dpooled,
                                                                float* __restrict__ partial, int E) {
    __shared__ float sm[64][65];
    const int n = blockIdx.x, t = threadIdx.x;
    const int EE = E * E;
    const float *wq = prm, *wk = wq + EE + E, *wv = wk + EE + E, *wo = wv + EE + E, *ln_w = wo + EE + E,
                *ln_b = ln_w + E;
    float xin[16], xhat[16], xn[16], dq[16], dkk[16], dvv[16], dout[16], a_o[16], dxn[16];
    const float rstd = mhsa_token_ln(pooled, ln_w, ln_b, n, E, t, xin, xhat, xn);
    for (int e = 0; e < E; ++e) {
        dq[e] = dqkv[((1LL * n * 3 + 0) * 64 + t) * E + e];
        dkk[e] = dqkv[((1LL * n * 3 + 1) * 64 + t) * E + e];
        dvv[e] = dqkv[((1LL * n * 3 + 2) * 64 + t) * E + e];
        dout[e] = dfeat[1LL * n * E * 64 + e * 64 + t];
        a_o[e] = ao[(1LL * n * 64 + t) * E + e];
    }
    for (int e = 0; e < E; ++e) {  // gradient of the LayerNorm output: Wq^T dq + Wk^T dk + Wv^T dv
        float a = 0.f;
        for (int o = 0; o < E; ++o) {
            a = fmaf(dq[o], wq[o * E + e], a);
            a = fmaf(dkk[o], wk[o * E + e], a);
            a = fmaf(dvv[o], wv[o * E + e], a);
        }
        dxn[e] = a;
    }
    float m1 = 0.f, m2 = 0.f;
    for (int e = 0; e < E; ++e) {
        const float g = dxn[e] * ln_w[e];
        m1 += g;
        m2 += g * xhat[e];
    }
    m1 /= static_cast<float>(E);
    m2 /= static_cast<float>(E);
    for (int e = 0; e < E; ++e)
        dpooled[(1LL * n * E + e) * 64 + t] = rstd * (dxn[e] * ln_w[e] - m1 - xhat[e] * m2) + dout[e];
    float* dst = partial + 1LL * n * (4 * (EE + E) + 2 * E);
    token_reduce(sm, t, EE, dst, [&](int i) { return dq[i / E] * xn[i % E]; });
    token_reduce(sm, t, E, dst + EE, [&](int i) { return dq[i]; });
    token_reduce(sm, t, EE, dst + EE + E, [&](int i) { return dkk[i / E] * xn[i % E]; });
    token_reduce(sm, t, E, dst + 2 * EE + E, [&](int i) { return dkk[i]; });
    token_reduce(sm, t, EE, dst + 2 * EE + 2 * E, [&](int i) { return dvv[i / E] * xn[i % E]; });
    token_reduce(sm, t, E, dst + 3 * EE + 2 * E, [&](int i) { return dvv[i]; });
    token_reduce(sm, t, EE, dst + 3 * EE + 3 * E, [&](int i) { return dout[i / E] * a_o[i % E]; });
    token_reduce(sm, t, E, dst + 4 * EE + 3 * E, [&](int i) { return dout[i]; });
    token_reduce(sm, t, E, dst + 4 * EE + 4 * E, [&](int i) { return dxn[i] * xhat[i]; });
    token_reduce(sm, t, E, dst + 4 * EE + 5 * E, [&](int i) { return dxn[i]; });
}

// ------------------------------------------------------------------------------ cross-attention (optional branch)
// modules.py:450-459: flat += mean_i(CrossAttention(query, tokens)[i]); the mean is a scalar per image, so the gradient
// w.r.t. every element of (out_proj output + residual query) is sum_i dflat[i] / Q.
__global__ void __launch_bounds__(256) head_cross_mean_bwd_kernel(const float* __restrict__ dflat,
                                                                  float* __restrict__ dq, int F, int Q) {
    __shared__ float red[8];
    const int n = blockIdx.x;
    float s = 0.f;
    for (int i = threadIdx.x; i < F; i += 256) s += dflat[1LL * n * F + i];
    const float g = block_sum256(s, red) / static_cast<float>(Q);
    for (int i = threadIdx.x; i < Q; i += 256) dq[1LL * n * Q + i] = g;
}

// Backward of head_cross_attn_kernel (vt_head.cu).  One CTA per image, thread d = one of the 256 embedding dims
// (warp = head, head_dim 32).  Recomputes k, v, the scores and the softmax; given datt[n][256] it writes
//   dq[n][256], dtok[n][E*64] (gradient into the token features) and the per-image parameter gradients
//   part[n] = { dWk[256][E], dbk[256], dWv[256][E], dbv[256] }   (the flat parameter order).
__global__ void __launch_bounds__(256) head_cross_attn_bwd_kernel(const float* __restrict__ feat,
                                                                  const float* __restrict__ q,
                                                                  const float* __restrict__ wk, const float* __restrict__ bk,
                                                                  const float* __restrict__ wv, const float* __restrict__ bv,
                                                                  const float* __restrict__ datt, float* __restrict__ dq,
                                                                  float* __restrict__ dtok, float* __restrict__ part,
                                                                  int E, int heads) {
    __shared__ float tok[16][64];
    __shared__ float sc[8][64];
    __shared__ float dp[8][64];
    extern __shared__ float wpart[];  // [8][E*64]
    const int n = blockIdx.x, d = threadIdx.x;
    const int hd = 256 / heads, h = d / hd, lane = d & 31, warp = d >> 5;
    for (int i = d; i < E * 64; i += 256) tok[i / 64][i % 64] = feat[1LL * n * E * 64 + i];
    __syncthreads();
    float kw[16], vw[16], gwk[16], gwv[16];
    for (int e = 0; e < E; ++e) {
        kw[e] = wk[d * E + e];
        vw[e] = wv[d * E + e];
        gwk[e] = 0.f;
        gwv[e] = 0.f;
    }
    const float scale = 1.0f / sqrtf(static_cast<float>(hd));
    const float qd = q[1LL * n * 256 + d], da = datt[1LL * n * 256 + d];
    for (int j = 0; j < 64; ++j) {
        float kj = bk[d], vj = bv[d];
        for (int e = 0; e < E; ++e) {
            kj = fmaf(kw[e], tok[e][j], kj);
            vj = fmaf(vw[e], tok[e][j], vj);
        }
        float ps = qd * scale * kj, pd = da * vj;
        for (int o = hd / 2; o > 0; o >>= 1) {
            ps += __shfl_xor_sync(0xFFFFFFFFu, ps, o, 32);
            pd += __shfl_xor_sync(0xFFFFFFFFu, pd, o, 32);
        }
        if (lane % hd == 0) {
            sc[h][j] = ps;
            dp[h][j] = pd;
        }
    }
    __syncthreads();
    float m = -CUDART_INF_F;
    for (int j = 0; j < 64; ++j) m = fmaxf(m, sc[h][j]);
    float sum = 0.f;
    for (int j = 0; j < 64; ++j) sum += expf(sc[h][j] - m);
    const float inv = 1.0f / sum;
    float rowdot = 0.f;
    for (int j = 0; j < 64; ++j) rowdot = fmaf(expf(sc[h][j] - m) * inv, dp[h][j], rowdot);
    float gq = 0.f, gbk = 0.f, gbv = 0.f;
    for (int j = 0; j < 64; ++j) {
        const float p = expf(sc[h][j] - m) * inv;
        const float ds = p * (dp[h][j] - rowdot);
        float kj = bk[d];
        for (int e = 0; e < E; ++e) kj = fmaf(kw[e], tok[e][j], kj);
        gq = fmaf(ds * scale, kj, gq);
        const float dk = ds * scale * qd;  // d loss / d k[j][d]
        const float dv = p * da;           // d loss / d v[j][d]
        gbk += dk;
        gbv += dv;
        for (int e = 0; e < E; ++e) {
            gwk[e] = fmaf(dk, tok[e][j], gwk[e]);
            gwv[e] = fmaf(dv, tok[e][j], gwv[e]);
            // gradient into token j, feature e: sum over the 256 dims -- inside the warp here, across warps below
            float t = fmaf(dk, kw[e], dv * vw[e]);
            for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xFFFFFFFFu, t, o, 32);
            if (lane == 0) wpart[warp * E * 64 + e * 64 + j] = t;
        }
    }
    dq[1LL * n * 256 + d] = gq;
    float* pn = part + 1LL * n * (2 * (256 * E + 256));
    for (int e = 0; e < E; ++e) {
        pn[d * E + e] = gwk[e];
        pn[256 * E + 256 + d * E + e] = gwv[e];
    }
    pn[256 * E + d] = gbk;
    pn[2 * 256 * E + 256 + d] = gbv;
    __syncthreads();
    for (int i = d; i < E * 64; i += 256) {
        float t = 0.f;
        for (int w = 0; w < 8; ++w) t += wpart[w * E * 64 + i];
        dtok[1LL * n * E * 64 + i] = t;
    }
}

// dst[i] = a[i] + b[i] (+ c[i])
__global__ void __launch_bounds__(256) head_add_kernel(const float* a, const float* b, const float* c, float* dst,
                                                       long long n) {  // dst may alias a
    const long long i = blockIdx.x * 256LL + threadIdx.x;
    if (i < n) dst[i] = a[i] + b[i] + (c ? c[i] : 0.f);
}

// ------------------------------------------------------------------------------ classifier
// LayerNorm (eps 1e-5) -> ReLU / LeakyReLU(0.2) -> Dropout(p), out of place; stat[b] = (mean, rstd).
__global__ void __launch_bounds__(256) head_ln_act_drop_kernel(const float* __restrict__ a, float* __restrict__ h,
                                                               float* __restrict__ stat,
                                                               const float* __restrict__ w,
                                                               const float* __restrict__ b, int D, int act, float p,
                                                               unsigned long long seed, unsigned stream_id) {
    __shared__ float red[8];
    const float* r = a + 1LL * blockIdx.x * D;
    float s = 0.f;
    for (int i = threadIdx.x; i < D; i += 256) s += r[i];
    const float mean = block_sum256(s, red) / static_cast<float>(D);
    float v = 0.f;
    for (int i = threadIdx.x; i < D; i += 256) v += (r[i] - mean) * (r[i] - mean);
    const float var = block_sum256(v, red) / static_cast<float>(D);
    const float rstd = 1.0f / sqrtf(var + 1e-5f);
    if (threadIdx.x == 0) {
        stat[2 * blockIdx.x] = mean;
        stat[2 * blockIdx.x + 1] = rstd;
    }
    const float inv_keep = p > 0.f ? 1.0f / (1.0f - p) : 1.0f;
    for (int i = threadIdx.x; i < D; i += 256) {
        float t = (r[i] - mean) * rstd * w[i] + b[i];
        if (act == 1) t = fmaxf(t, 0.f);
        else if (act == 2) t = t > 0.f ? t : 0.2f * t;
        if (p > 0.f) t = h_dropout_keep(seed, stream_id, 1ULL * blockIdx.x * D + i, p) ? t * inv_keep : 0.f;
        h[1LL * blockIdx.x * D + i] = t;
    }
}

// backward of the above.  dh (gradient w.r.t. the dropout output) is replaced by the gradient w.r.t.
// the LayerNorm input; dt (gradient w.r.t. the LayerNorm output) is kept for the affine gradients.
__global__ void __launch_bounds__(256) head_ln_act_drop_bwd_kernel(float* __restrict__ dh,
                                                                   const float* __restrict__ a,
                                                                   const float* __restrict__ stat,
                                                                   const float* __restrict__ w,
                                                                   const float* __restrict__ b,
                                                                   float* __restrict__ dt_out, int D, int act,
                                                                   float p, unsigned long long seed,
                                                                   unsigned stream_id) {
    __shared__ float red[8];
    const long long row = 1LL * blockIdx.x * D;
    const float mean = stat[2 * blockIdx.x], rstd = stat[2 * blockIdx.x + 1];
    const float inv_keep = p > 0.f ? 1.0f / (1.0f - p) : 1.0f;
    float s1 = 0.f, s2 = 0.f;
    for (int i = threadIdx.x; i < D; i += 256) {
        const float xhat = (a[row + i] - mean) * rstd;
        const float t = xhat * w[i] + b[i];
        float g = dh[row + i];
        if (p > 0.f) g = h_dropout_keep(seed, stream_id, row + i, p) ? g * inv_keep : 0.f;
        const float slope = t > 0.f ? 1.0f : (act == 2 ? 0.2f : 0.f);
        const float dt = g * slope;
        dt_out[row + i] = dt;
        const float gw = dt * w[i];
        s1 += gw;
        s2 += gw * xhat;
    }
    const float m1 = block_sum256(s1, red) / static_cast<float>(D);
    const float m2 = block_sum256(s2, red) / static_cast<float>(D);
    for (int i = threadIdx.x; i < D; i += 256) {
        const float xhat = (a[row + i] - mean) * rstd;
        dh[row + i] = rstd * (dt_out[row + i] * w[i] - m1 - xhat * m2);
    }
}

// LayerNorm affine gradients: dgamma[d] += sum_b dt[b][d]*xhat[b][d]; dbeta[d] += sum_b dt[b][d]
__global__ void __launch_bounds__(256) head_ln_param_bwd_kernel(const float* __restrict__ dt,
                                                                const float* __restrict__ a,
                                                                const float* __restrict__ stat,
                                                                float* __restrict__ g_w, float* __restrict__ g_b,
                                                                int B, int D) {
    const int d = blockIdx.x * 256 + threadIdx.x;
    if (d >= D) return;
    float sw = 0.f, sb = 0.f;
    for (int b = 0; b < B; ++b) {
        const float v = dt[1LL * b * D + d];
        sw = fmaf(v, (a[1LL * b * D + d] - stat[2 * b]) * stat[2 * b + 1], sw);
        sb += v;
    }
    g_w[d] += sw;
    g_b[d] += sb;
}

// Linear backward, weights: dW[o][i] += sum_b dy[b][o]*x[b][i]; db[o] += sum_b dy[b][o]
__global__ void __launch_bounds__(256) head_linear_bwd_w_kernel(const float* __restrict__ dy,
                                                                const float* __restrict__ x,
                                                                float* __restrict__ g_w, float* __restrict__ g_b,
                                                                int B, int I, int O) {
    const long long idx = blockIdx.x * 256LL + threadIdx.x;
    if (idx >= 1LL * O * I) return;
    const int o = static_cast<int>(idx / I), i = static_cast<int>(idx - 1LL * o * I);
    float acc = 0.f, bacc = 0.f;
    for (int b = 0; b < B; ++b) {
        const float d = dy[1LL * b * O + o];
        acc = fmaf(d, x[1LL * b * I + i], acc);
        bacc += d;
    }
    g_w[idx] += acc;
    if (i == 0) g_b[o] += bacc;
}

// Linear backward, input: dx[b][i] = sum_o dy[b][o]*W[o][i], split over o: grid (ceil(I/128), ceil(O/64)).
// A CTA owns 64 rows of W (each read once, coalesced) and 128 columns; a thread keeps one column's sums
// for 32 batch rows in registers.  part[ks][b][i]; the slices are added in a fixed order afterwards.
__global__ void __launch_bounds__(128) head_linear_bwd_x_kernel(const float* __restrict__ dy,
                                                                const float* __restrict__ w,
                                                                float* __restrict__ part, int B, int I, int O) {
    __shared__ float sdy[32][65];
    const int i = blockIdx.x * 128 + threadIdx.x;
    const int ks = blockIdx.y, o0 = ks * 64;
    const int cnt = min(64, O - o0);
    for (int b0 = 0; b0 < B; b0 += 32) {
        __syncthreads();
        for (int e = threadIdx.x; e < 32 * 64; e += 128) {
            const int bb = e >> 6, oo = e & 63;
            sdy[bb][oo] = (b0 + bb < B && oo < cnt) ? dy[1LL * (b0 + bb) * O + o0 + oo] : 0.f;
        }
        __syncthreads();
        float acc[32];
#pragma unroll
        for (int bb = 0; bb < 32; ++bb) acc[bb] = 0.f;
        if (i < I) {
            for (int oo = 0; oo < cnt; ++oo) {
                const float wv = w[1LL * (o0 + oo) * I + i];
#pragma unroll
                for (int bb = 0; bb < 32; ++bb) acc[bb] = fmaf(sdy[bb][oo], wv, acc[bb]);
            }
#pragma unroll
            for (int bb = 0; bb < 32; ++bb)
                if (b0 + bb < B) part[(1LL * ks * B + b0 + bb) * I + i] = acc[bb];
        }
    }
}

// dst[i] = sum over parts of part[p][i]
__global__ void __launch_bounds__(256) head_partial_assign_kernel(const float* __restrict__ part, int nparts,
                                                                  long long P, float* __restrict__ dst) {
    const long long i = blockIdx.x * 256LL + threadIdx.x;
    if (i >= P) return;
    float s = 0.f;
    for (int p = 0; p < nparts; ++p) s += part[1LL * p * P + i];
    dst[i] = s;
}

__global__ void head_scale_add_kernel(const float* __restrict__ src, float scale, float* __restrict__ dst) {
    if (threadIdx.x == 0 && blockIdx.x == 0) *dst += *src * scale;
}

// dropout keep-masks as 0/1 floats (parity tests only: lets the oracle apply the same masks)
__global__ void __launch_bounds__(256) head_dropout_mask_kernel(float* __restrict__ out, long long n, float p,
                                                                unsigned long long seed, unsigned stream_id) {
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += 256LL * gridDim.x)
        out[i] = (p > 0.f && !h_dropout_keep(seed, stream_id, static_cast<unsigned long long>(i), p)) ? 0.f : 1.f;
}

// ------------------------------------------------------------------------------ optimizer
// sum of squares of a flat buffer, two-stage: part[blockIdx.x] then one CTA
__global__ void __launch_bounds__(256) sumsq_partial_kernel(const float* __restrict__ g, long long n,
                                                            double* __restrict__ part) {
    __shared__ double red[8];
    double s = 0.0;
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += 256LL * gridDim.x) {
        const double v = static_cast<double>(g[i]);
        s += v * v;
    }
    s = block_sum256_d(s, red);
    if (threadIdx.x == 0) part[blockIdx.x] = s;
}
__global__ void __launch_bounds__(256) sumsq_final_kernel(const double* __restrict__ part, int nparts,
                                                          float grad_scale, float* __restrict__ norm_out) {
    __shared__ double red[8];
    double s = 0.0;
    for (int i = threadIdx.x; i < nparts; i += 256) s += part[i];
    s = block_sum256_d(s, red);
    if (threadIdx.x == 0) *norm_out = static_cast<float>(sqrt(s)) * fabsf(grad_scale);
}
// clip_grad_norm_(max_norm) + torch.optim.AdamW (decoupled weight decay) on flat buffers; grads zeroed
__global__ void __launch_bounds__(256) adamw_kernel(float* __restrict__ p, float* __restrict__ g,
                                                    float* __restrict__ m, float* __restrict__ v, long long n,
                                                    float lr, float beta1, float beta2, float eps, float wd,
                                                    float bias1, float bias2_sqrt, float grad_scale, float max_norm,
                                                    const float* __restrict__ norm, int zero_grad) {
    float clip = 1.0f;
    if (max_norm > 0.f) clip = fminf(1.0f, max_norm / (*norm + 1e-6f));
    const float gs = grad_scale * clip;
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += 256LL * gridDim.x) {
        const float gi = g[i] * gs;
        float pi = p[i] * (1.0f - lr * wd);
        const float mi = beta1 * m[i] + (1.0f - beta1) * gi;
        const float vi = beta2 * v[i] + (1.0f - beta2) * gi * gi;
        m[i] = mi;
        v[i] = vi;
        const float denom = sqrtf(vi) / bias2_sqrt + eps;
        pi -= (lr / bias1) * (mi / denom);
        p[i] = pi;
        if (zero_grad) g[i] = 0.f;
    }
}

struct Mlp {
    int n;           // number of Linear layers
    int dims[6];     // n+1 widths
    int act;         // 1 ReLU, 2 LeakyReLU(0.2)
    float drop[5];   // dropout after each hidden layer
};

Mlp mlp_of(const vt_head_config& h) {
    Mlp m{};
    if (h.kind == VT_HEAD_ATTENTION) {
        m.n = 4;
        const int d[5] = {(h.latent_channels / 2) * 64, 1024, 512, 256, h.num_classes};
        for (int i = 0; i < 5; ++i) m.dims[i] = d[i];
        m.act = 1;
        m.drop[0] = 0.3f; m.drop[1] = 0.2f; m.drop[2] = 0.1f;
    } else {
        m.n = 3;
        const int d[4] = {h.plain_flat_dim > 0 ? h.plain_flat_dim : h.latent_channels * 16, 512, 256, h.num_classes};
        for (int i = 0; i < 4; ++i) m.dims[i] = d[i];
        m.act = 2;
        m.drop[0] = 0.3f; m.drop[1] = 0.2f;
    }
    return m;
}

}  // namespace

// ------------------------------------------------------------------------------ parameter layout
std::vector<HeadParamEntry> head_param_layout(const vt_head_config& h) {
    std::vector<HeadParamEntry> L;
    int64_t off = 0;
    auto add = [&](const std::string& name, std::vector<int64_t> shape) {
        int64_t n = 1;
        for (auto d : shape) n *= d;
        L.push_back({name, off, n, shape});
        off += n;
    };
    const int C = h.latent_channels, E = C / 2, T = h.num_classes;
    if (h.kind == VT_HEAD_ATTENTION) {
        if (h.use_spatial_attention) {
            add("spatial_attention.channel_att.0.weight", {C / 8, C, 1, 1});
            add("spatial_attention.channel_att.2.weight", {C, C / 8, 1, 1});
            add("spatial_attention.spatial_att.0.weight", {1, 2, 7, 7});
        }
        add("feature_compress.0.weight", {E, C, 3, 3});
        add("feature_compress.0.bias", {E});
        add("feature_compress.1.weight", {E});
        add("feature_compress.1.bias", {E});
        if (h.use_self_attention) {
            for (const char* k : {"q_proj", "k_proj", "v_proj", "out_proj"}) {
                add(std::string("self_attention_post.") + k + ".weight", {E, E});
                add(std::string("self_attention_post.") + k + ".bias", {E});
            }
            add("self_attention_post.norm.weight", {E});
            add("self_attention_post.norm.bias", {E});
        }
        if (h.use_cross_attention) {  // CrossAttention(query_dim 512, key_dim E, embed_dim 256), modules.py:388-395
            add("cross_attention.q_proj.weight", {256, 512});
            add("cross_attention.q_proj.bias", {256});
            add("cross_attention.k_proj.weight", {256, E});
            add("cross_attention.k_proj.bias", {256});
            add("cross_attention.v_proj.weight", {256, E});
            add("cross_attention.v_proj.bias", {256});
            add("cross_attention.out_proj.weight", {512, 256});
            add("cross_attention.out_proj.bias", {512});
        }
    }
    const Mlp m = mlp_of(h);
    for (int i = 0; i < m.n; ++i) {
        const std::string lin = "classifier." + std::to_string(4 * i);
        add(lin + ".weight", {m.dims[i + 1], m.dims[i]});
        add(lin + ".bias", {m.dims[i + 1]});
        if (i + 1 < m.n) {
            const std::string ln = "classifier." + std::to_string(4 * i + 1);
            add(ln + ".weight", {m.dims[i + 1]});
            add(ln + ".bias", {m.dims[i + 1]});
        }
    }
    if (h.kind == VT_HEAD_ATTENTION && h.use_cross_attention) {  // registered last in the reference (modules.py:421-422)
        add("query_generator.weight", {512, E * 64});
        add("query_generator.bias", {512});
    }
    (void)T;
    return L;
}

namespace {
struct Arena {
    size_t off = 0;
    size_t take(size_t n) {
        const size_t o = off;
        off += (n + 63) / 64 * 64;
        return o;
    }
};
struct TrainWs {
    size_t pool, cgate, map2, x2, sgate, z, dz, dy, dpre, bnstat, bnsums, pooled, dpooled, feat, dfeat;
    size_t a[4], hbuf[4], stat[4], dtbuf, g0, g1, logits, dlogits, loss;
    size_t xq, xqp, xatt, xout, feat2, dxout, dxatt, dxqp, dxq, dfadd, dtok, part_x, part_dx, ao, dqkv, part_bn, part_mhsa, part_cw, part_cb, part_g, part_w7, part_norm;
    size_t total;
};
constexpr int kBands = 8;      // row bands / pixel chunks per image of the spatial-attention reductions
constexpr int kConvBands = 2;  // row bands per image of the conv weight gradient (72 sums per CTA: keep CTAs fat)

TrainWs plan(const vt_head_config& h, int B, int H, int W) {
    TrainWs w{};
    Arena A;
    const size_t C = h.latent_channels, E = C / 2, HW = static_cast<size_t>(H) * W, T = h.num_classes;
    const Mlp m = mlp_of(h);
    const size_t nblk = (HW + 255) / 256;
    if (h.kind == VT_HEAD_ATTENTION) {
        w.pool = A.take(B * C * 2); w.cgate = A.take(B * C); w.map2 = A.take(B * 2 * HW);
        w.x2 = A.take(B * C * HW); w.sgate = A.take(B * HW); w.z = A.take(B * E * HW); w.dz = A.take(B * E * HW);
        w.dy = A.take(B * C * HW); w.dpre = A.take(B * HW); w.bnstat = A.take(2 * E); w.bnsums = A.take(2 * E);
        w.pooled = A.take(B * E * 64); w.dpooled = A.take(B * E * 64);
        w.ao = A.take(B * 64 * E); w.dqkv = A.take(B * 3 * 64 * E);
        w.part_bn = A.take(2 * (B * nblk * E * 2));  // doubles
        w.part_mhsa = A.take(B * (4 * (E * E + E) + 2 * E));
        w.part_cw = A.take(static_cast<size_t>(B) * kConvBands * E * C * 9);
        w.part_cb = A.take(static_cast<size_t>(B) * kConvBands * E);
        w.part_g = A.take(static_cast<size_t>(B) * kBands * C);
        w.part_w7 = A.take(static_cast<size_t>(B) * kBands * 98);
    }
    size_t maxd = 0;
    for (int i = 0; i <= m.n; ++i) maxd = std::max<size_t>(maxd, m.dims[i]);
    w.feat = A.take(B * static_cast<size_t>(m.dims[0])); w.dfeat = A.take(B * static_cast<size_t>(m.dims[0]));
    for (int i = 0; i + 1 < m.n; ++i) {
        w.a[i] = A.take(B * static_cast<size_t>(m.dims[i + 1]));
        w.hbuf[i] = A.take(B * static_cast<size_t>(m.dims[i + 1]));
        w.stat[i] = A.take(2 * static_cast<size_t>(B));
    }
    if (h.kind == VT_HEAD_ATTENTION && h.use_cross_attention) {
        const size_t F = static_cast<size_t>(m.dims[0]);
        w.xq = A.take(B * 512); w.xqp = A.take(B * 256); w.xatt = A.take(B * 256); w.xout = A.take(B * 512);
        w.feat2 = A.take(B * F); w.dxout = A.take(B * 512); w.dxatt = A.take(B * 256); w.dxqp = A.take(B * 256);
        w.dxq = A.take(B * 512); w.dfadd = A.take(B * F); w.dtok = A.take(B * F);
        w.part_x = A.take(static_cast<size_t>(B) * 2 * (256 * E + 256));
    }
    w.dtbuf = A.take(B * maxd); w.g0 = A.take(B * maxd); w.g1 = A.take(B * maxd);
    {
        size_t need = 0;
        for (int i = 0; i < m.n; ++i)
            need = std::max(need, static_cast<size_t>((m.dims[i + 1] + 63) / 64) * B * m.dims[i]);
        if (h.kind == VT_HEAD_ATTENTION && h.use_cross_attention)
            need = std::max(need, static_cast<size_t>(8) * B * std::max(512, m.dims[0]));
        w.part_dx = A.take(need);
    }
    w.logits = A.take(B * T); w.dlogits = A.take(B * T); w.loss = A.take(1);
    w.total = A.off;
    return w;
}
}  // namespace

size_t head_train_workspace_floats(const vt_head_config& h, int B, int H, int W) { return plan(h, B, H, W).total; }

template <int CO>
static void launch_conv_bwd(const float* x, const float* dz, const float* cw, float* part_w, float* part_b,
                            float* dy, int N, int H, int W, cudaStream_t s) {
    head_conv_bwd_w_kernel<CO><<<dim3(2 * CO, kConvBands, N), 256, 0, s>>>(x, dz, part_w, part_b, 2 * CO, H, W);
    if (dy) head_conv_bwd_x_kernel<CO><<<dim3((H * W + 255) / 256, N), 256, 0, s>>>(dz, cw, dy, H, W);
}

#define VT_KC(...)                            \
    do {                                      \
        profiler_begin(pf, KC_HEAD, s, 0, 0); \
        __VA_ARGS__;                          \
        profiler_end(pf, KC_HEAD, s);         \
    } while (0)

int head_train_step(const vt_head_config& h, const vt_head_train_args& a, float* ws, Profiler* pf) {
    cudaStream_t s = static_cast<cudaStream_t>(a.stream);
    const int B = a.batch, C = h.latent_channels, E = C / 2, T = h.num_classes, H = a.lat_h, W = a.lat_w;
    const int HW = H * W;
    const TrainWs w = plan(h, B, H, W);
    const Mlp m = mlp_of(h);
    const auto layout = head_param_layout(h);
    auto find = [&](const std::string& name) -> int64_t {
        for (const auto& e : layout)
            if (e.name == name) return e.offset;
        return -1;
    };
    auto P = [&](const std::string& name) { return a.params + find(name); };
    auto G = [&](const std::string& name) { return a.grads + find(name); };
    const bool drop = a.dropout != 0;
    const bool att = h.kind == VT_HEAD_ATTENTION;
    const int nblk = (HW + 255) / 256;
    const int chunks = std::max(1, std::min(nblk, kBands));
    double* part_bn = reinterpret_cast<double*>(ws + w.part_bn);

    // ------------------------------------------------------------------ forward (train mode)
    const float* y = a.latent;  // input of feature_compress
    if (att) {
        VT_CHECK(E == 4 || E == 8 || E == 12 || E == 16, "head training needs latent_channels 8, 16, 24 or 32");
        if (h.use_spatial_attention) {
            VT_TRY(launch_head_spatial_attention(a.latent, P("spatial_attention.channel_att.0.weight"),
                                                 P("spatial_attention.channel_att.2.weight"),
                                                 P("spatial_attention.spatial_att.0.weight"), ws + w.pool,
                                                 ws + w.cgate, ws + w.map2, ws + w.x2, ws + w.sgate, B, C, H, W, s,
                                                 pf));
            y = ws + w.x2;
        }
        const size_t smem = (static_cast<size_t>(E) * C * 9 + 8 * 2 * E) * sizeof(float);
        const float* cw = P("feature_compress.0.weight");
        const float* cb = P("feature_compress.0.bias");
        profiler_begin(pf, KC_HEAD, s, 2.0 * B * HW * E * C * 9, 4.0 * B * (C + E) * HW);
        switch (E) {
            case 4: head_conv_train_kernel<4><<<dim3(nblk, B), 256, smem, s>>>(y, cw, cb, ws + w.z, part_bn, C, H, W); break;
            case 8: head_conv_train_kernel<8><<<dim3(nblk, B), 256, smem, s>>>(y, cw, cb, ws + w.z, part_bn, C, H, W); break;
            case 12: head_conv_train_kernel<12><<<dim3(nblk, B), 256, smem, s>>>(y, cw, cb, ws + w.z, part_bn, C, H, W); break;
            default: head_conv_train_kernel<16><<<dim3(nblk, B), 256, smem, s>>>(y, cw, cb, ws + w.z, part_bn, C, H, W); break;
        }
        profiler_end(pf, KC_HEAD, s);
        VT_KC(head_bn_finalize_kernel<<<1, 256, 0, s>>>(part_bn, B * nblk, E, static_cast<double>(B) * HW, 1e-5f,
                                                         a.bn_momentum, ws + w.bnstat, a.bn_running_mean,
                                                         a.bn_running_var,
                                                         reinterpret_cast<long long*>(a.bn_num_batches_tracked)));
        VT_KC(head_bn_relu_pool_kernel<<<dim3(64, B), 256, 0, s>>>(ws + w.z, ws + w.bnstat,
                                                                     P("feature_compress.1.weight"),
                                                                     P("feature_compress.1.bias"), ws + w.pooled, E,
                                                                     H, W));
        if (h.use_self_attention) {
            VT_CHECK(h.attention_heads >= 1 && E % h.attention_heads == 0, "embed_dim not divisible by heads");
            VT_KC(head_mhsa_head_kernel<false><<<dim3(h.attention_heads, B), 64, 0, s>>>(
                ws + w.pooled, P("self_attention_post.q_proj.weight"), ws + w.ao, nullptr, nullptr, E,
                h.attention_heads, drop ? a.attention_dropout : 0.f, a.seed));
            VT_KC(head_mhsa_out_kernel<<<B, 64, 0, s>>>(ws + w.pooled, P("self_attention_post.q_proj.weight"),
                                                         ws + w.ao, ws + w.feat, E));
        }
    } else if (h.plain_flat_dim > 0) {
        VT_CHECK(h.plain_flat_dim == C * H * W, "latent size differs from the one the plain head was built for");
    } else {
        VT_TRY(launch_head_adaptive_pool(a.latent, ws + w.feat, B, C, H, W, 4, 4, s, pf));
    }
    const float* feat = (att && !h.use_self_attention) ? ws + w.pooled
                        : (!att && h.plain_flat_dim > 0)  ? a.latent
                                                          : ws + w.feat;
    const bool cross = att && h.use_cross_attention;
    const float* cls_in = feat;  // input of the classifier
    if (cross) {
        VT_CHECK(h.attention_heads == 8, "the cross-attention branch needs attention_heads = 8");
        VT_TRY(launch_head_linear(feat, P("query_generator.weight"), P("query_generator.bias"), ws + w.xq, B, m.dims[0],
                                  512, s, pf));
        VT_TRY(launch_head_linear(ws + w.xq, P("cross_attention.q_proj.weight"), P("cross_attention.q_proj.bias"),
                                  ws + w.xqp, B, 512, 256, s, pf));
        VT_TRY(launch_head_cross_attention(feat, ws + w.xqp, P("cross_attention.k_proj.weight"),
                                           P("cross_attention.k_proj.bias"), P("cross_attention.v_proj.weight"),
                                           P("cross_attention.v_proj.bias"), ws + w.xatt, B, E, h.attention_heads, s, pf));
        VT_TRY(launch_head_linear(ws + w.xatt, P("cross_attention.out_proj.weight"), P("cross_attention.out_proj.bias"),
                                  ws + w.xout, B, 256, 512, s, pf));
        VT_TRY(launch_head_cross_add(ws + w.xout, ws + w.xq, feat, ws + w.feat2, B, 512, m.dims[0], s, pf));
        cls_in = ws + w.feat2;
    }
    float* logits = a.logits ? a.logits : ws + w.logits;
    {
        const float* cur = cls_in;
        for (int i = 0; i < m.n; ++i) {
            const std::string lin = "classifier." + std::to_string(4 * i);
            float* out = (i + 1 == m.n) ? logits : ws + w.a[i];
            VT_TRY(launch_head_linear(cur, P(lin + ".weight"), P(lin + ".bias"), out, B, m.dims[i], m.dims[i + 1], s,
                                      pf));
            if (i + 1 < m.n) {
                const std::string ln = "classifier." + std::to_string(4 * i + 1);
                VT_KC(head_ln_act_drop_kernel<<<B, 256, 0, s>>>(out, ws + w.hbuf[i], ws + w.stat[i],
                                                                 P(ln + ".weight"), P(ln + ".bias"), m.dims[i + 1],
                                                                 m.act, drop ? m.drop[i] : 0.f, a.seed, 1u + i));
                cur = ws + w.hbuf[i];
            }
        }
    }
    // ------------------------------------------------------------------ loss (mean over B*T) + dlogits
    const long long nl = 1LL * B * T;
    VT_CUDA(cudaMemsetAsync(ws + w.loss, 0, sizeof(float), s));
    VT_TRY(launch_focal_loss(logits, a.targets, ws + w.loss, ws + w.dlogits, nl, a.focal_alpha, a.focal_gamma,
                             a.loss_scale / static_cast<float>(nl), s, pf, a.class_weights, T));
    if (a.loss)
        head_scale_add_kernel<<<1, 32, 0, s>>>(ws + w.loss, a.loss_scale / static_cast<float>(nl), a.loss);
    if (!a.grads) {
        VT_CUDA(cudaGetLastError());
        return 0;
    }
    // ------------------------------------------------------------------ backward: classifier
    float* gcur = ws + w.dlogits;  // gradient w.r.t. the current layer's output
    float* gbuf[2] = {ws + w.g0, ws + w.g1};
    for (int i = m.n - 1; i >= 0; --i) {
        const std::string lin = "classifier." + std::to_string(4 * i);
        const float* xin = i == 0 ? cls_in : ws + w.hbuf[i - 1];
        const int I = m.dims[i], O = m.dims[i + 1];
        VT_KC(head_linear_bwd_w_kernel<<<static_cast<int>((1LL * O * I + 255) / 256), 256, 0, s>>>(
            gcur, xin, G(lin + ".weight"), G(lin + ".bias"), B, I, O));
        if (i == 0 && !att) break;  // the pooled latent needs no gradient
        float* gx = i == 0 ? ws + w.dfeat : gbuf[i & 1];
        VT_KC(head_linear_bwd_x_kernel<<<dim3((I + 127) / 128, (O + 63) / 64), 128, 0, s>>>(
            gcur, P(lin + ".weight"), ws + w.part_dx, B, I, O));
        VT_KC(head_partial_assign_kernel<<<static_cast<int>((1LL * B * I + 255) / 256), 256, 0, s>>>(
            ws + w.part_dx, (O + 63) / 64, 1LL * B * I, gx));
        gcur = gx;
        if (i > 0) {
            const std::string ln = "classifier." + std::to_string(4 * (i - 1) + 1);
            VT_KC(head_ln_act_drop_bwd_kernel<<<B, 256, 0, s>>>(gcur, ws + w.a[i - 1], ws + w.stat[i - 1],
                                                                 P(ln + ".weight"), P(ln + ".bias"), ws + w.dtbuf, I,
                                                                 m.act, drop ? m.drop[i - 1] : 0.f, a.seed,
                                                                 1u + (i - 1)));
            VT_KC(head_ln_param_bwd_kernel<<<(I + 255) / 256, 256, 0, s>>>(ws + w.dtbuf, ws + w.a[i - 1],
                                                                            ws + w.stat[i - 1], G(ln + ".weight"),
                                                                            G(ln + ".bias"), B, I));
        }
    }
    auto lin_bwd = [&](const float* dy, const float* x, const std::string& name, float* dx, int I, int O) -> int {
        VT_KC(head_linear_bwd_w_kernel<<<static_cast<int>((1LL * O * I + 255) / 256), 256, 0, s>>>(
            dy, x, G(name + ".weight"), G(name + ".bias"), B, I, O));
        VT_KC(head_linear_bwd_x_kernel<<<dim3((I + 127) / 128, (O + 63) / 64), 128, 0, s>>>(dy, P(name + ".weight"),
                                                                                           ws + w.part_dx, B, I, O));
        VT_KC(head_partial_assign_kernel<<<static_cast<int>((1LL * B * I + 255) / 256), 256, 0, s>>>(
            ws + w.part_dx, (O + 63) / 64, 1LL * B * I, dx));
        return 0;
    };
    if (cross) {
        // -------------------------------------------------------------- cross-attention branch (modules.py:450-459)
        const int F = m.dims[0];
        VT_KC(head_cross_mean_bwd_kernel<<<B, 256, 0, s>>>(ws + w.dfeat, ws + w.dxout, F, 512));
        VT_TRY(lin_bwd(ws + w.dxout, ws + w.xatt, "cross_attention.out_proj", ws + w.dxatt, 256, 512));
        const size_t smem_x = static_cast<size_t>(8) * E * 64 * sizeof(float);
        VT_KC(head_cross_attn_bwd_kernel<<<B, 256, smem_x, s>>>(
            feat, ws + w.xqp, P("cross_attention.k_proj.weight"), P("cross_attention.k_proj.bias"),
            P("cross_attention.v_proj.weight"), P("cross_attention.v_proj.bias"), ws + w.dxatt, ws + w.dxqp, ws + w.dtok,
            ws + w.part_x, E, h.attention_heads));
        const int PX = 2 * (256 * E + 256);
        VT_KC(head_partial_reduce_kernel<<<(PX + 255) / 256, 256, 0, s>>>(ws + w.part_x, B, PX,
                                                                         G("cross_attention.k_proj.weight")));
        VT_TRY(lin_bwd(ws + w.dxqp, ws + w.xq, "cross_attention.q_proj", ws + w.dxq, 512, 256));
        // the query also feeds the residual of CrossAttention: d query = q_proj path + d(out + query)
        VT_KC(head_add_kernel<<<static_cast<int>((1LL * B * 512 + 255) / 256), 256, 0, s>>>(ws + w.dxq, ws + w.dxout, nullptr,
                                                                                           ws + w.dxq, 1LL * B * 512));
        VT_TRY(lin_bwd(ws + w.dxq, feat, "query_generator", ws + w.dfadd, F, 512));
        // d features = identity path of "flat + mean" + query_generator path + token path of k / v
        VT_KC(head_add_kernel<<<static_cast<int>((1LL * B * F + 255) / 256), 256, 0, s>>>(ws + w.dfeat, ws + w.dfadd,
                                                                                         ws + w.dtok, ws + w.dfeat,
                                                                                         1LL * B * F));
    }
    if (att) {
        // -------------------------------------------------------------- self-attention
        const float* dpooled = ws + w.dfeat;
        if (h.use_self_attention) {
            const int PM = 4 * (E * E + E) + 2 * E;
            VT_KC(head_mhsa_head_kernel<true><<<dim3(h.attention_heads, B), 64, 0, s>>>(
                ws + w.pooled, P("self_attention_post.q_proj.weight"), ws + w.ao, ws + w.dfeat, ws + w.dqkv, E,
                h.attention_heads, drop ? a.attention_dropout : 0.f, a.seed));
            VT_KC(head_mhsa_bwd_tail_kernel<<<B, 64, 0, s>>>(ws + w.pooled, P("self_attention_post.q_proj.weight"),
                                                              ws + w.ao, ws + w.dqkv, ws + w.dfeat, ws + w.dpooled,
                                                              ws + w.part_mhsa, E));
            VT_KC(head_partial_reduce_kernel<<<(PM + 255) / 256, 256, 0, s>>>(
                ws + w.part_mhsa, B, PM, G("self_attention_post.q_proj.weight")));
            dpooled = ws + w.dpooled;
        }
        // -------------------------------------------------------------- pool / ReLU / BatchNorm / conv
        VT_KC(head_bn_bwd_stats_kernel<<<dim3(nblk, B), 256, 0, s>>>(ws + w.z, ws + w.bnstat,
                                                                      P("feature_compress.1.weight"),
                                                                      P("feature_compress.1.bias"), dpooled,
                                                                      ws + w.dz, part_bn, E, H, W));
        VT_KC(head_bn_bwd_finalize_kernel<<<1, 256, 0, s>>>(part_bn, B * nblk, E, ws + w.bnsums,
                                                             G("feature_compress.1.weight"),
                                                             G("feature_compress.1.bias")));
        const long long total = 1LL * B * E * HW;
        VT_KC(head_bn_bwd_apply_kernel<<<static_cast<int>(std::min<long long>((total + 255) / 256, 148 * 16)), 256,
                                          0, s>>>(ws + w.z, ws + w.bnstat, ws + w.bnsums,
                                                  P("feature_compress.1.weight"), ws + w.dz, E, HW, total,
                                                  1.0f / (static_cast<float>(B) * HW)));
        float* dy = h.use_spatial_attention ? ws + w.dy : nullptr;
        profiler_begin(pf, KC_HEAD, s, 0, 0);
        switch (E) {
            case 4: launch_conv_bwd<4>(y, ws + w.dz, P("feature_compress.0.weight"), ws + w.part_cw, ws + w.part_cb, dy, B, H, W, s); break;
            case 8: launch_conv_bwd<8>(y, ws + w.dz, P("feature_compress.0.weight"), ws + w.part_cw, ws + w.part_cb, dy, B, H, W, s); break;
            case 12: launch_conv_bwd<12>(y, ws + w.dz, P("feature_compress.0.weight"), ws + w.part_cw, ws + w.part_cb, dy, B, H, W, s); break;
            default: launch_conv_bwd<16>(y, ws + w.dz, P("feature_compress.0.weight"), ws + w.part_cw, ws + w.part_cb, dy, B, H, W, s); break;
        }
        profiler_end(pf, KC_HEAD, s);
        VT_KC(head_partial_reduce_kernel<<<(E * C * 9 + 255) / 256, 256, 0, s>>>(
            ws + w.part_cw, B * kConvBands, E * C * 9, G("feature_compress.0.weight")));
        VT_KC(head_partial_reduce_kernel<<<1, 256, 0, s>>>(ws + w.part_cb, B * kConvBands, E,
                                                            G("feature_compress.0.bias")));
        // -------------------------------------------------------------- spatial attention
        if (h.use_spatial_attention) {
            VT_KC(head_sa_bwd_pre_kernel<<<dim3(chunks, B), 256, 0, s>>>(a.latent, ws + w.cgate, ws + w.sgate,
                                                                          ws + w.dy, ws + w.dpre, C, HW));
            VT_KC(head_sa_bwd_gate_sum_kernel<<<dim3(chunks, B), 256, 0, s>>>(
                a.latent, ws + w.cgate, ws + w.sgate, ws + w.dy, ws + w.dpre,
                P("spatial_attention.spatial_att.0.weight"), ws + w.part_g, C, H, W));
            VT_KC(head_sa_bwd_w7_kernel<<<dim3(chunks, B, 2), 256, 0, s>>>(ws + w.map2, ws + w.dpre,
                                                                           ws + w.part_w7, H, W));
            VT_KC(head_partial_reduce_kernel<<<1, 256, 0, s>>>(ws + w.part_w7, B * chunks, 98,
                                                                G("spatial_attention.spatial_att.0.weight")));
            VT_KC(head_sa_bwd_mlp_kernel<<<1, 256, 0, s>>>(ws + w.pool, ws + w.cgate, ws + w.part_g, chunks,
                                                            P("spatial_attention.channel_att.0.weight"),
                                                            P("spatial_attention.channel_att.2.weight"),
                                                            G("spatial_attention.channel_att.0.weight"),
                                                            G("spatial_attention.channel_att.2.weight"), B, C,
                                                            C / 8));
        }
    }
    VT_CUDA(cudaGetLastError());
    return 0;
}

int head_dropout_masks(const vt_head_config& h, int B, float attention_dropout, unsigned long long seed,
                       float* attn, float* const* cls, cudaStream_t s) {
    const Mlp m = mlp_of(h);
    if (attn && h.kind == VT_HEAD_ATTENTION && h.use_self_attention) {
        const long long n = 1LL * B * h.attention_heads * 64 * 64;
        head_dropout_mask_kernel<<<static_cast<int>(std::min<long long>((n + 255) / 256, 1184)), 256, 0, s>>>(
            attn, n, attention_dropout, seed, 0u);
    }
    for (int i = 0; i + 1 < m.n; ++i) {
        if (!cls || !cls[i]) continue;
        const long long n = 1LL * B * m.dims[i + 1];
        head_dropout_mask_kernel<<<static_cast<int>(std::min<long long>((n + 255) / 256, 1184)), 256, 0, s>>>(
            cls[i], n, m.drop[i], seed, 1u + i);
    }
    VT_CUDA(cudaGetLastError());
    return 0;
}

int launch_adamw(float* p, float* g, float* m, float* v, long long n, float lr, float beta1, float beta2, float eps,
                 float wd, long long step, float grad_scale, float max_norm, double* scratch /*>=1184 doubles*/,
                 float* norm_out, int zero_grad, cudaStream_t s, Profiler* pf) {
    const int grid = static_cast<int>(std::max<long long>(1, std::min<long long>((n + 255) / 256, 1184)));
    VT_KC(sumsq_partial_kernel<<<grid, 256, 0, s>>>(g, n, scratch));
    VT_KC(sumsq_final_kernel<<<1, 256, 0, s>>>(scratch, grid, grad_scale, norm_out));
    const float bias1 = 1.0f - static_cast<float>(pow(static_cast<double>(beta1), static_cast<double>(step)));
    const float bias2 = 1.0f - static_cast<float>(pow(static_cast<double>(beta2), static_cast<double>(step)));
    VT_KC(adamw_kernel<<<grid, 256, 0, s>>>(p, g, m, v, n, lr, beta1, beta2, eps, wd, bias1, sqrtf(bias2),
                                              grad_scale, max_norm, norm_out, zero_grad));
    VT_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace vt
