// Tag-decoder head kernels (reference modules.py:15-91, :303-475), eval-mode forward, fp32.
// The head reads 16 x h x w floats per image (1 MB at 1024^2) and does ~44 MFLOP: it is
// latency/HBM bound, so these are warp-shuffle CUDA-core kernels sized to keep the whole
// batch on chip (everything is L2 resident), not tensor-core kernels.
#include <math_constants.h>

#include "vt_head_common.cuh"
#include "vt_internal.h"

namespace vt {

// ---- SpatialAttention step 1 (modules.py:38-39): global average and max per (image, channel)
// latent NCHW fp32; grid = N*C blocks; pool[n][c][0]=avg, [1]=max
__global__ void __launch_bounds__(256) head_pool_kernel(const float* __restrict__ x, float* __restrict__ pool,
                                                        int HW) {
    __shared__ float red[8];
    const float* p = x + 1LL * blockIdx.x * HW;
    float s = 0.f, m = -CUDART_INF_F;
    for (int i = threadIdx.x; i < HW; i += 256) {
        const float v = p[i];
        s += v;
        m = fmaxf(m, v);
    }
    s = block_sum256(s, red);
    m = block_max256(m, red);
    if (threadIdx.x == 0) {
        pool[2 * blockIdx.x] = s / static_cast<float>(HW);
        pool[2 * blockIdx.x + 1] = m;
    }
}

// ---- SpatialAttention step 2 (modules.py:38-45): channel gate = sigmoid(MLP(avg)+MLP(max)),
// then per pixel the channel mean / max of the gated tensor -> 2-channel map m[n][2][HW].
// cgate[n][c] is also written for the next kernel.
__global__ void __launch_bounds__(256) head_channel_gate_kernel(const float* __restrict__ x,
                                                                const float* __restrict__ pool,
                                                                const float* __restrict__ w1,  // [C/r][C]
                                                                const float* __restrict__ w2,  // [C][C/r]
                                                                float* __restrict__ cgate, float* __restrict__ map2,
                                                                int C, int Ch, int HW) {
    __shared__ float g[64];
    __shared__ float hid[2][16];
    const int n = blockIdx.y;
    if (threadIdx.x < 2 * Ch) {
        const int which = threadIdx.x / Ch, j = threadIdx.x % Ch;
        float a = 0.f;
        for (int c = 0; c < C; ++c) a = fmaf(w1[j * C + c], pool[(1LL * n * C + c) * 2 + which], a);
        hid[which][j] = fmaxf(a, 0.f);
    }
    __syncthreads();
    if (threadIdx.x < C) {
        float a = 0.f, b = 0.f;
        for (int j = 0; j < Ch; ++j) {
            a = fmaf(w2[threadIdx.x * Ch + j], hid[0][j], a);
            b = fmaf(w2[threadIdx.x * Ch + j], hid[1][j], b);
        }
        const float gate = sigmoidf_(a + b);
        g[threadIdx.x] = gate;
        if (blockIdx.x == 0) cgate[1LL * n * C + threadIdx.x] = gate;
    }
    __syncthreads();
    for (int p = blockIdx.x * 256 + threadIdx.x; p < HW; p += 256 * gridDim.x) {
        float s = 0.f, m = -CUDART_INF_F;
        for (int c = 0; c < C; ++c) {
            const float v = x[(1LL * n * C + c) * HW + p] * g[c];
            s += v;
            m = fmaxf(m, v);
        }
        map2[(1LL * n * 2) * HW + p] = s / static_cast<float>(C);
        map2[(1LL * n * 2 + 1) * HW + p] = m;
    }
}

// ---- SpatialAttention step 3 (modules.py:46-47): spatial gate = sigmoid(conv7x7(map2)) (pad 3, no
// bias); y = x * cgate * sgate, NCHW fp32.
__global__ void __launch_bounds__(256) head_spatial_gate_kernel(const float* __restrict__ x,
                                                                const float* __restrict__ cgate,
                                                                const float* __restrict__ map2,
                                                                const float* __restrict__ w7,  // [1][2][7][7]
                                                                float* __restrict__ y,
                                                                float* __restrict__ sgate,  // [N][HW] or nullptr
                                                                int C, int H, int W) {
    __shared__ float ws[98];
    __shared__ float g[64];
    const int n = blockIdx.y;
    const int HW = H * W;
    if (threadIdx.x < 98) ws[threadIdx.x] = w7[threadIdx.x];
    if (threadIdx.x < C) g[threadIdx.x] = cgate[1LL * n * C + threadIdx.x];
    __syncthreads();
    for (int p = blockIdx.x * 256 + threadIdx.x; p < HW; p += 256 * gridDim.x) {
        const int py = p / W, px = p - py * W;
        float a = 0.f;
        for (int ch = 0; ch < 2; ++ch) {
            const float* mp = map2 + (1LL * n * 2 + ch) * HW;
            for (int ky = 0; ky < 7; ++ky) {
                const int yy = py + ky - 3;
                if (yy < 0 || yy >= H) continue;
                for (int kx = 0; kx < 7; ++kx) {
                    const int xx = px + kx - 3;
                    if (xx < 0 || xx >= W) continue;
                    a = fmaf(ws[(ch * 7 + ky) * 7 + kx], mp[yy * W + xx], a);
                }
            }
        }
        const float sg = sigmoidf_(a);
        if (sgate) sgate[1LL * n * HW + p] = sg;
        for (int c = 0; c < C; ++c) {
            const long long o = (1LL * n * C + c) * HW + p;
            y[o] = x[o] * g[c] * sg;
        }
    }
}

// ---- feature_compress (modules.py:377-382): conv3x3 C->C/2 (+bias) -> BatchNorm (running stats)
// -> ReLU -> AdaptiveAvgPool(8,8).  One CTA per (pooled cell, image); each thread owns pixels of
// the cell's window and all C/2 output channels; block reduction per channel.
// out: pooled[n][oc][64]
__global__ void __launch_bounds__(256) head_compress_kernel(const float* __restrict__ x,   // [N][C][H][W]
                                                            const float* __restrict__ cw,  // [Co][C][3][3]
                                                            const float* __restrict__ cb,
                                                            const float* __restrict__ bn_w,
                                                            const float* __restrict__ bn_b,
                                                            const float* __restrict__ bn_rm,
                                                            const float* __restrict__ bn_rv, float bn_eps,
                                                            float* __restrict__ pooled, int C, int Co, int H,
                                                            int W) {
    extern __shared__ float sw[];  // Co*C*9 weights, then 8 floats scratch
    float* red = sw + Co * C * 9;
    const int n = blockIdx.y, cell = blockIdx.x;
    const int cy = cell / 8, cx = cell % 8;
    // adaptive pooling window: [floor(i*H/8), ceil((i+1)*H/8))
    const int y0 = (cy * H) / 8, y1 = ((cy + 1) * H + 7) / 8;
    const int x0 = (cx * W) / 8, x1 = ((cx + 1) * W + 7) / 8;
    const int wh = y1 - y0, ww = x1 - x0;
    for (int i = threadIdx.x; i < Co * C * 9; i += 256) sw[i] = cw[i];
    __syncthreads();
    float acc[16];
#pragma unroll
    for (int o = 0; o < 16; ++o) acc[o] = 0.f;
    for (int i = threadIdx.x; i < wh * ww; i += 256) {
        const int py = y0 + i / ww, px = x0 + i % ww;
        float v[16];
#pragma unroll
        for (int o = 0; o < 16; ++o) v[o] = (o < Co) ? cb[o] : 0.f;
        for (int c = 0; c < C; ++c) {
            const float* xp = x + (1LL * n * C + c) * H * W;
            for (int ky = 0; ky < 3; ++ky) {
                const int yy = py + ky - 1;
                if (yy < 0 || yy >= H) continue;
                for (int kx = 0; kx < 3; ++kx) {
                    const int xx = px + kx - 1;
                    if (xx < 0 || xx >= W) continue;
                    const float xv = xp[yy * W + xx];
#pragma unroll
                    for (int o = 0; o < 16; ++o)
                        if (o < Co) v[o] = fmaf(sw[((o * C + c) * 3 + ky) * 3 + kx], xv, v[o]);
                }
            }
        }
#pragma unroll
        for (int o = 0; o < 16; ++o) {
            if (o < Co) {
                const float t = (v[o] - bn_rm[o]) / sqrtf(bn_rv[o] + bn_eps) * bn_w[o] + bn_b[o];
                acc[o] += fmaxf(t, 0.f);
            }
        }
    }
    for (int o = 0; o < Co; ++o) {
        const float t = block_sum256(acc[o], red);
        if (threadIdx.x == 0) pooled[(1LL * n * Co + o) * 64 + cell] = t / static_cast<float>(wh * ww);
    }
}

// ---- MultiHeadSelfAttention on the 8x8 grid (modules.py:66-91): tokens [64][E], LN(E), q/k/v
// E->E, heads x head_dim, softmax over 64 keys, out_proj + residual; E <= 16.
// Two kernels: (1) one CTA per (head, image), one thread per query token: LayerNorm, this head's q/k/v,
// softmax(q k^T / sqrt(hd)) v -> ao[n][t][e]  (8x the parallelism of a CTA per image -- the per-thread work is a
// serial chain of 64 exponentials per head); (2) one CTA per image: out_proj + residual, written as the
// flattened NCHW feature f[n][e*64 + token] (modules.py:448).
__device__ __forceinline__ void mhsa_ln_token(const float* __restrict__ pooled, const float* __restrict__ ln_w,
                                              const float* __restrict__ ln_b, int n, int E, int t, float* xn) {
    float xin[16];
    for (int e = 0; e < E; ++e) xin[e] = pooled[(1LL * n * E + e) * 64 + t];
    float mean = 0.f;
    for (int e = 0; e < E; ++e) mean += xin[e];
    mean /= static_cast<float>(E);
    float var = 0.f;
    for (int e = 0; e < E; ++e) var += (xin[e] - mean) * (xin[e] - mean);
    var /= static_cast<float>(E);
    const float rstd = 1.0f / sqrtf(var + 1e-5f);
    for (int e = 0; e < E; ++e) xn[e] = (xin[e] - mean) * rstd * ln_w[e] + ln_b[e];
}

__global__ void __launch_bounds__(64) head_mhsa_heads_kernel(const float* __restrict__ pooled,  // [N][E][64]
                                                             const float* __restrict__ ln_w, const float* __restrict__ ln_b,
                                                             const float* __restrict__ wq, const float* __restrict__ bq,
                                                             const float* __restrict__ wk, const float* __restrict__ bk,
                                                             const float* __restrict__ wv, const float* __restrict__ bv,
                                                             float* __restrict__ ao,  // [N][64][E]
                                                             int E, int heads) {
    __shared__ float sk[64][17];
    __shared__ float sv[64][17];
    const int h = blockIdx.x, n = blockIdx.y, t = threadIdx.x;
    const int hd = E / heads, c0 = h * hd;
    float xn[16], q[16];
    mhsa_ln_token(pooled, ln_w, ln_b, n, E, t, xn);
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
    }
    __syncthreads();
    const float scale = 1.0f / sqrtf(static_cast<float>(hd));
    float sc[64];
    float m = -CUDART_INF_F;
#pragma unroll
    for (int j = 0; j < 64; ++j) {
        float s = 0.f;
        for (int d = 0; d < hd; ++d) s = fmaf(q[d], sk[j][d], s);
        s *= scale;
        sc[j] = s;
        m = fmaxf(m, s);
    }
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < 64; ++j) {
        sc[j] = expf(sc[j] - m);
        sum += sc[j];
    }
    const float inv = 1.0f / sum;
    for (int d = 0; d < hd; ++d) {
        float o = 0.f;
#pragma unroll
        for (int j = 0; j < 64; ++j) o = fmaf(sc[j] * inv, sv[j][d], o);
        ao[(1LL * n * 64 + t) * E + c0 + d] = o;
    }
}

__global__ void __launch_bounds__(64) head_mhsa_out_kernel_eval(const float* __restrict__ pooled,
                                                                const float* __restrict__ ao,
                                                                const float* __restrict__ wo, const float* __restrict__ bo,
                                                                float* __restrict__ feat, int E, int enabled) {
    const int n = blockIdx.x, t = threadIdx.x;
    if (!enabled) {
        for (int e = 0; e < E; ++e) feat[1LL * n * E * 64 + e * 64 + t] = pooled[(1LL * n * E + e) * 64 + t];
        return;
    }
    float a[16];
    for (int e = 0; e < E; ++e) a[e] = ao[(1LL * n * 64 + t) * E + e];
    for (int o = 0; o < E; ++o) {
        float r = bo[o];
        for (int e = 0; e < E; ++e) r = fmaf(wo[o * E + e], a[e], r);
        feat[1LL * n * E * 64 + o * 64 + t] = r + pooled[(1LL * n * E + o) * 64 + t];
    }
}

// ---- CrossAttention (modules.py:93-122), the optional --use_cross_attention branch (modules.py:450-459):
// one query vector per image (q = q_proj(query_generator(flat)), computed by the Linear kernel) attends over
// the 64 tokens of the 8x8 feature grid.  One CTA per image, thread d = one of the 256 embedding dims:
// k/v projections of the tokens (E -> 256), per-head scores (head_dim 32, warp = head), softmax over the
// 64 keys, weighted sum of the values.  feat: [N][E*64] (channel-major tokens), out: [N][256].
__global__ void __launch_bounds__(256) head_cross_attn_kernel(const float* __restrict__ feat,
                                                              const float* __restrict__ q,   // [N][256]
                                                              const float* __restrict__ wk, const float* __restrict__ bk,
                                                              const float* __restrict__ wv, const float* __restrict__ bv,
                                                              float* __restrict__ out, int E, int heads) {
    __shared__ float tok[16][64];
    __shared__ float sc[8][64];
    const int n = blockIdx.x, d = threadIdx.x;
    const int hd = 256 / heads, h = d / hd;
    for (int i = d; i < E * 64; i += 256) tok[i / 64][i % 64] = feat[1LL * n * E * 64 + i];
    __syncthreads();
    float kw[16], vw[16];
    for (int e = 0; e < E; ++e) {
        kw[e] = wk[d * E + e];
        vw[e] = wv[d * E + e];
    }
    const float qd = q[1LL * n * 256 + d] / sqrtf(static_cast<float>(hd));
    // scores[h][j] = sum over the head's dims of q[d] * k[j][d]: each thread adds its dim (shared-memory atomics
    // would be order dependent: reduce inside the warp instead; hd = 32 -> one warp per head)
    const int lane = d & 31;
    for (int j = 0; j < 64; ++j) {
        float kj = bk[d];
        for (int e = 0; e < E; ++e) kj = fmaf(kw[e], tok[e][j], kj);
        float part = qd * kj;
        for (int o = hd / 2; o > 0; o >>= 1) part += __shfl_xor_sync(0xFFFFFFFFu, part, o, 32);
        if (lane % hd == 0) sc[h][j] = part;
    }
    __syncthreads();
    float m = -CUDART_INF_F;
    for (int j = 0; j < 64; ++j) m = fmaxf(m, sc[h][j]);
    float sum = 0.f, acc = 0.f;
    for (int j = 0; j < 64; ++j) {
        const float p = expf(sc[h][j] - m);
        float vj = bv[d];
        for (int e = 0; e < E; ++e) vj = fmaf(vw[e], tok[e][j], vj);
        sum += p;
        acc = fmaf(p, vj, acc);
    }
    out[1LL * n * 256 + d] = acc / sum;
}

// flat[n][:] += mean_i(attended[n][i] + query[n][i])   (modules.py:118,459: out_proj output + residual query,
// its mean over the 512 dims broadcast onto every feature).  One CTA per image.
__global__ void __launch_bounds__(256) head_cross_add_kernel(const float* __restrict__ attended,
                                                             const float* __restrict__ query, const float* __restrict__ in,
                                                             float* __restrict__ out, int Q, int F) {
    __shared__ float red[8];
    const int n = blockIdx.x;
    float s = 0.f;
    for (int i = threadIdx.x; i < Q; i += 256) s += attended[1LL * n * Q + i] + query[1LL * n * Q + i];
    const float mean = block_sum256(s, red) / static_cast<float>(Q);
    for (int i = threadIdx.x; i < F; i += 256) out[1LL * n * F + i] = in[1LL * n * F + i] + mean;
}

// ---- AdaptiveAvgPool2d((OH,OW)) for the plain ClassificationDecoder (modules.py:313, :339-340):
// out[n][c*OH*OW + cell]
__global__ void __launch_bounds__(64) head_adaptive_pool_kernel(const float* __restrict__ x, float* __restrict__ out,
                                                                int C, int H, int W, int OH, int OW) {
    const int n = blockIdx.y;
    const int idx = blockIdx.x;  // c*OH*OW + cell
    const int c = idx / (OH * OW), cell = idx % (OH * OW);
    const int cy = cell / OW, cx = cell % OW;
    const int y0 = (cy * H) / OH, y1 = ((cy + 1) * H + OH - 1) / OH;
    const int x0 = (cx * W) / OW, x1 = ((cx + 1) * W + OW - 1) / OW;
    const int wh = y1 - y0, ww = x1 - x0;
    float s = 0.f;
    for (int i = threadIdx.x; i < wh * ww; i += 64)
        s += x[((1LL * n * C + c) * H + y0 + i / ww) * W + x0 + i % ww];
    s = h_warp_sum(s);
    __shared__ float r[2];
    if ((threadIdx.x & 31) == 0) r[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) out[1LL * n * C * OH * OW + idx] = (r[0] + r[1]) / static_cast<float>(wh * ww);
}

// ---- Linear: y[b][o] = bias[o] + sum_i W[o][i] x[b][i].  One warp per output neuron; the neuron's
// weight row is read ONCE (coalesced, two loads in flight per lane) and applied to BT batch rows held in
// registers; the activations (B x I floats) are L1/L2 resident and shared by every warp.
template <int BT>
__global__ void __launch_bounds__(256) head_linear_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                          const float* __restrict__ bias, float* __restrict__ y,
                                                          int B, int I, int O) {
    const int o = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (o >= O) return;
    const float* wr = w + 1LL * o * I;
    for (int b0 = 0; b0 < B; b0 += BT) {
        float acc[BT];
#pragma unroll
        for (int bb = 0; bb < BT; ++bb) acc[bb] = 0.f;
        const int nb = min(BT, B - b0);
        if (nb == BT) {
            int i = lane;
            for (; i + 32 < I; i += 64) {
                const float w0 = wr[i], w1 = wr[i + 32];
#pragma unroll
                for (int bb = 0; bb < BT; ++bb) {
                    const float* xr = x + 1LL * (b0 + bb) * I + i;
                    acc[bb] = fmaf(w1, xr[32], fmaf(w0, xr[0], acc[bb]));
                }
            }
            for (; i < I; i += 32) {
                const float w0 = wr[i];
#pragma unroll
                for (int bb = 0; bb < BT; ++bb) acc[bb] = fmaf(w0, x[1LL * (b0 + bb) * I + i], acc[bb]);
            }
        } else {
            for (int i = lane; i < I; i += 32) {
                const float w0 = wr[i];
#pragma unroll
                for (int bb = 0; bb < BT; ++bb)
                    if (bb < nb) acc[bb] = fmaf(w0, x[1LL * (b0 + bb) * I + i], acc[bb]);
            }
        }
#pragma unroll
        for (int bb = 0; bb < BT; ++bb) {
            const float t = h_warp_sum(acc[bb]);
            if (lane == 0 && bb < nb) y[1LL * (b0 + bb) * O + o] = t + (bias ? bias[o] : 0.f);
        }
    }
}

// ---- LayerNorm (eps 1e-5) + activation in place: act 1 = ReLU, 2 = LeakyReLU(0.2); CTA per row
__global__ void __launch_bounds__(256) head_ln_act_kernel(float* __restrict__ x, const float* __restrict__ w,
                                                          const float* __restrict__ b, int D, int act) {
    __shared__ float red[8];
    float* r = x + 1LL * blockIdx.x * D;
    float s = 0.f;
    for (int i = threadIdx.x; i < D; i += 256) s += r[i];
    const float mean = block_sum256(s, red) / static_cast<float>(D);
    float v = 0.f;
    for (int i = threadIdx.x; i < D; i += 256) v += (r[i] - mean) * (r[i] - mean);
    const float var = block_sum256(v, red) / static_cast<float>(D);
    const float rstd = 1.0f / sqrtf(var + 1e-5f);
    for (int i = threadIdx.x; i < D; i += 256) {
        float t = (r[i] - mean) * rstd * w[i] + b[i];
        if (act == 1) t = fmaxf(t, 0.f);
        else if (act == 2) t = t > 0.f ? t : 0.2f * t;
        r[i] = t;
    }
}

// ---- get_confidence (modules.py:470-475) + threshold count (infer_full.py:114):
// conf = sigmoid(logits); descending sort of (conf, index) per image in shared memory (bitonic,
// ties broken by ascending index so the order is deterministic); count of conf >= thr.
__global__ void __launch_bounds__(1024) head_confidence_kernel(const float* __restrict__ logits,
                                                               float* __restrict__ conf_sorted,
                                                               long long* __restrict__ idx_sorted,
                                                               int* __restrict__ count, float* __restrict__ probs,
                                                               int T, int P2, float thr) {
    extern __shared__ unsigned char sm_raw[];
    float* key = reinterpret_cast<float*>(sm_raw);
    int* val = reinterpret_cast<int*>(key + P2);
    __shared__ int cnt;
    const int n = blockIdx.x;
    if (threadIdx.x == 0) cnt = 0;
    __syncthreads();
    int local = 0;
    for (int i = threadIdx.x; i < P2; i += blockDim.x) {
        float c = -1.0f;  // padding sorts last (sigmoid >= 0)
        if (i < T) {
            c = sigmoidf_(logits[1LL * n * T + i]);
            if (probs) probs[1LL * n * T + i] = c;
            if (c >= thr) ++local;
        }
        key[i] = c;
        val[i] = i;
    }
    if (local) atomicAdd(&cnt, local);
    __syncthreads();
    for (int k = 2; k <= P2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < P2; i += blockDim.x) {
                const int l = i ^ j;
                if (l > i) {
                    const float a = key[i], b = key[l];
                    const int ia = val[i], ib = val[l];
                    // "a before b" in the final descending order
                    const bool a_first = (a > b) || (a == b && ia < ib);
                    const bool desc_block = ((i & k) == 0);
                    if (desc_block ? !a_first : a_first) {
                        key[i] = b; key[l] = a;
                        val[i] = ib; val[l] = ia;
                    }
                }
            }
            __syncthreads();
        }
    }
    for (int i = threadIdx.x; i < T; i += blockDim.x) {
        if (conf_sorted) conf_sorted[1LL * n * T + i] = key[i];
        if (idx_sorted) idx_sorted[1LL * n * T + i] = val[i];
    }
    if (threadIdx.x == 0 && count) count[n] = cnt;
}

// ---- FocalLoss forward + analytic backward (improved_losses.py:47-56):
//   bce = max(x,0) - x*y + log1p(exp(-|x|)); pt = exp(-bce); fl = alpha*(1-pt)^gamma*bce
//   dfl/dx = alpha*(sigmoid(x)-y)*[(1-pt)^gamma + gamma*(1-pt)^(gamma-1)*pt*bce]
// loss_sum accumulates sum(fl) (one atomic per block); grad = grad_scale * dfl/dx.
__global__ void __launch_bounds__(256) focal_loss_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                         float* __restrict__ loss_sum, float* __restrict__ grad,
                                                         long long n, float alpha, float gamma, float grad_scale,
                                                         const float* __restrict__ class_w, int T) {
    __shared__ float red[8];
    float local = 0.f;
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += 256LL * gridDim.x) {
        const float xv = x[i], yv = y[i];
        const float cw = class_w ? class_w[i % T] : 1.0f;  // ClassBalancedLoss: per-class weight (improved_losses.py:66-72)
        const float bce = fmaxf(xv, 0.f) - xv * yv + log1pf(expf(-fabsf(xv)));
        const float pt = expf(-bce);
        const float om = 1.0f - pt;
        const float mod = (gamma == 0.f) ? 1.0f : powf(om, gamma);
        local += cw * alpha * mod * bce;
        if (grad) {
            const float dbce = sigmoidf_(xv) - yv;
            float dmod = 0.f;
            if (gamma != 0.f && om > 0.f) dmod = gamma * powf(om, gamma - 1.0f) * pt;
            grad[i] = grad_scale * cw * alpha * dbce * (mod + dmod * bce);
        }
    }
    const float t = block_sum256(local, red);
    if (threadIdx.x == 0 && loss_sum) atomicAdd(loss_sum, t);
}

// =========================================================================================
int launch_head_spatial_attention(const float* latent, const float* w1, const float* w2, const float* w7,
                                  float* pool, float* cgate, float* map2, float* out, float* sgate, int N, int C,
                                  int H, int W, cudaStream_t s, Profiler* prof) {
    VT_CHECK(C <= 64 && C % 8 == 0 && C / 8 <= 16, "SpatialAttention supports up to 64 channels");
    const int HW = H * W;
    profiler_begin(prof, KC_HEAD, s, 0, 4.0 * N * C * HW);
    head_pool_kernel<<<N * C, 256, 0, s>>>(latent, pool, HW);
    profiler_end(prof, KC_HEAD, s);
    const int chunks = std::max(1, std::min((HW + 255) / 256, 8));
    profiler_begin(prof, KC_HEAD, s, 0, 4.0 * N * (C + 2) * HW);
    head_channel_gate_kernel<<<dim3(chunks, N), 256, 0, s>>>(latent, pool, w1, w2, cgate, map2, C, C / 8, HW);
    profiler_end(prof, KC_HEAD, s);
    profiler_begin(prof, KC_HEAD, s, 0, 4.0 * N * (2 * C + 2) * HW);
    head_spatial_gate_kernel<<<dim3(chunks, N), 256, 0, s>>>(latent, cgate, map2, w7, out, sgate, C, H, W);
    profiler_end(prof, KC_HEAD, s);
    VT_CUDA(cudaGetLastError());
    return 0;
}

int launch_head_compress(const float* x, const float* cw, const float* cb, const float* bn_w, const float* bn_b,
                         const float* bn_rm, const float* bn_rv, float bn_eps, float* pooled, int N, int C, int H,
                         int W, cudaStream_t s, Profiler* prof) {
    const int Co = C / 2;
    VT_CHECK(Co <= 16 && Co >= 1, "feature_compress supports up to 32 latent channels");
    const size_t smem = (static_cast<size_t>(Co) * C * 9 + 8) * sizeof(float);
    profiler_begin(prof, KC_HEAD, s, 2.0 * N * H * W * Co * C * 9, 4.0 * N * C * H * W);
    head_compress_kernel<<<dim3(64, N), 256, smem, s>>>(x, cw, cb, bn_w, bn_b, bn_rm, bn_rv, bn_eps, pooled, C, Co, H,
                                                        W);
    profiler_end(prof, KC_HEAD, s);
    VT_CUDA(cudaGetLastError());
    return 0;
}

int launch_head_mhsa(const float* pooled, const float* const* p /*10 pointers*/, float* feat, float* ao /*[N][64][E]*/,
                     int N, int E, int heads, int enabled, cudaStream_t s, Profiler* prof) {
    VT_CHECK(E <= 16 && heads >= 1 && E % heads == 0, "self-attention embed dim must be <= 16 and divisible by heads");
    profiler_begin(prof, KC_HEAD, s, 0, 8.0 * N * E * 64);
    if (enabled)
        head_mhsa_heads_kernel<<<dim3(heads, N), 64, 0, s>>>(pooled, p[0], p[1], p[2], p[3], p[4], p[5], p[6], p[7], ao, E,
                                                             heads);
    head_mhsa_out_kernel_eval<<<N, 64, 0, s>>>(pooled, ao, p[8], p[9], feat, E, enabled);
    profiler_end(prof, KC_HEAD, s);
    VT_CUDA(cudaGetLastError());
    return 0;
}

int launch_head_cross_attention(const float* feat, const float* q, const float* wk, const float* bk, const float* wv,
                                const float* bv, float* out, int N, int E, int heads, cudaStream_t s,
                                Profiler* prof) {
    VT_CHECK(E <= 16 && heads >= 1 && 256 % heads == 0 && (256 / heads == 32),
             "cross-attention kernel: embed 256 with 8 heads of 32 (modules.py:396-398)");
    profiler_begin(prof, KC_HEAD, s, 0, 4.0 * N * (E * 64 + 512));
    head_cross_attn_kernel<<<N, 256, 0, s>>>(feat, q, wk, bk, wv, bv, out, E, heads);
    profiler_end(prof, KC_HEAD, s);
    VT_CUDA(cudaGetLastError());
    return 0;
}

int launch_head_cross_add(const float* attended, const float* query, const float* in, float* out, int N, int Q, int F,
                          cudaStream_t s, Profiler* prof) {
    profiler_begin(prof, KC_HEAD, s, 0, 4.0 * N * (2 * Q + 2 * F));
    head_cross_add_kernel<<<N, 256, 0, s>>>(attended, query, in, out, Q, F);
    profiler_end(prof, KC_HEAD, s);
    VT_CUDA(cudaGetLastError());
    return 0;
}

int launch_head_adaptive_pool(const float* x, float* out, int N, int C, int H, int W, int OH, int OW,
                              cudaStream_t s, Profiler* prof) {
    profiler_begin(prof, KC_HEAD, s, 0, 4.0 * N * C * H * W);
    head_adaptive_pool_kernel<<<dim3(C * OH * OW, N), 64, 0, s>>>(x, out, C, H, W, OH, OW);
    profiler_end(prof, KC_HEAD, s);
    VT_CUDA(cudaGetLastError());
    return 0;
}

int launch_head_linear(const float* x, const float* w, const float* b, float* y, int B, int I, int O,
                       cudaStream_t s, Profiler* prof) {
    profiler_begin(prof, KC_HEAD, s, 2.0 * B * I * O, 4.0 * I * O);
    if (B >= 16) head_linear_kernel<32><<<(O + 7) / 8, 256, 0, s>>>(x, w, b, y, B, I, O);
    else head_linear_kernel<8><<<(O + 7) / 8, 256, 0, s>>>(x, w, b, y, B, I, O);
    profiler_end(prof, KC_HEAD, s);
    VT_CUDA(cudaGetLastError());
    return 0;
}

int launch_head_ln_act(float* x, const float* w, const float* b, int B, int D, int act, cudaStream_t s,
                       Profiler* prof) {
    profiler_begin(prof, KC_HEAD, s, 0, 8.0 * B * D);
    head_ln_act_kernel<<<B, 256, 0, s>>>(x, w, b, D, act);
    profiler_end(prof, KC_HEAD, s);
    VT_CUDA(cudaGetLastError());
    return 0;
}

int launch_head_confidence(const float* logits, float* conf_sorted, long long* idx_sorted, int* count,
                           float* probs, int B, int T, float thr, cudaStream_t s, Profiler* prof) {
    int p2 = 1;
    while (p2 < T) p2 <<= 1;
    VT_CHECK(p2 <= 16384, "confidence sort supports up to 16384 tags");
    const size_t smem = static_cast<size_t>(p2) * 8;
    static SmemAttrOnce once;
    VT_TRY(ensure_dyn_smem(once, head_confidence_kernel, 16384 * 8));
    const int threads = std::max(32, std::min(1024, p2 / 2));
    profiler_begin(prof, KC_HEAD, s, 0, 16.0 * B * T);
    head_confidence_kernel<<<B, threads, smem, s>>>(logits, conf_sorted, idx_sorted, count, probs, T, p2, thr);
    profiler_end(prof, KC_HEAD, s);
    VT_CUDA(cudaGetLastError());
    return 0;
}

int launch_focal_loss(const float* logits, const float* targets, float* loss_sum, float* grad, long long n,
                      float alpha, float gamma, float grad_scale, cudaStream_t s, Profiler* prof, const float* class_w,
                      int T) {
    const int grid = static_cast<int>(std::max<long long>(1, std::min<long long>((n + 255) / 256, 148 * 8)));
    profiler_begin(prof, KC_HEAD, s, 0, 12.0 * n);
    focal_loss_kernel<<<grid, 256, 0, s>>>(logits, targets, loss_sum, grad, n, alpha, gamma, grad_scale, class_w, T);
    profiler_end(prof, KC_HEAD, s);
    VT_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace vt
