// VAE fine-tuning losses of the reference (improved_losses.py), forward + analytic backward in one call
// (SURVEY.md 8f-4):
//   * ImprovedTripletLoss (improved_losses.py:74-109) and ContrastiveLoss (:6-37) over flattened latents
//     [B][D] (D = 16*h*w: 262 144 at 1024^2), cosine (F.normalize eps 1e-12) or euclidean (F.pairwise_distance eps
//     1e-6) distance, label-overlap weights;
//   * F.mse_loss reconstruction term of CombinedLoss / train_vae.py (:278, train_vae.py:137);
//   * AdaptiveLossWeights (:111-125): softmax(log_w / T) weighted sum of four scalar losses.
// Three stages, all fixed order (no atomics): per-(row, chunk) partial dot products -> one block per row folds them,
// reads the labels, forms the row loss and the coefficients of the gradient (every gradient is a linear
// combination of the three input rows plus a constant) -> one elementwise pass writes the gradients.
#include "vt_internal.h"
#include "vt_losses.h"

namespace vt {

namespace {

constexpr int NDOT = 8;   // aa, pp, nn, ap, an, sum a, sum p, sum n

__global__ void __launch_bounds__(256) rowdots_kernel(const float* __restrict__ a, const float* __restrict__ p,
                                                      const float* __restrict__ n, float* __restrict__ part,
                                                      long long D) {
    __shared__ float sh[NDOT][256];
    const int b = blockIdx.y;
    const long long chunk = ((D + gridDim.x - 1) / gridDim.x + 3) / 4 * 4;
    const long long d0 = blockIdx.x * chunk, d1 = min(D, d0 + chunk);
    const float* ar = a + b * D;
    const float* pr = p + b * D;
    const float* nr = n ? n + b * D : nullptr;
    float v[NDOT] = {};
    for (long long d = d0 + threadIdx.x; d < d1; d += 256) {
        const float x = ar[d], y = pr[d], z = nr ? nr[d] : 0.f;
        v[0] = fmaf(x, x, v[0]); v[1] = fmaf(y, y, v[1]); v[2] = fmaf(z, z, v[2]);
        v[3] = fmaf(x, y, v[3]); v[4] = fmaf(x, z, v[4]);
        v[5] += x; v[6] += y; v[7] += z;
    }
#pragma unroll
    for (int k = 0; k < NDOT; ++k) sh[k][threadIdx.x] = v[k];
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o)
#pragma unroll
            for (int k = 0; k < NDOT; ++k) sh[k][threadIdx.x] += sh[k][threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x < NDOT) part[(1LL * b * gridDim.x + blockIdx.x) * NDOT + threadIdx.x] = sh[threadIdx.x][0];
}

// coef[b][12]: grad_a = c0 a + c1 p + c2 n + c9 ; grad_p = c3 a + c4 p + c10 ; grad_n = c5 a + c6 n + c11 (c7, c8 unused)
__global__ void __launch_bounds__(128) embed_rows_kernel(const float* __restrict__ part, int chunks,
                                                         const float* __restrict__ la, const float* __restrict__ lp,
                                                         int T, long long D, int B, int kind, int euclid, float margin,
                                                         float* __restrict__ rowloss, float* __restrict__ coef) {
    __shared__ double dots[NDOT];
    __shared__ float lsum[3][128];
    const int b = blockIdx.x;
    if (threadIdx.x < NDOT) {
        double t = 0.0;
        for (int k = 0; k < chunks; ++k) t += part[(1LL * b * chunks + k) * NDOT + threadIdx.x];
        dots[threadIdx.x] = t;
    }
    // label sums: overlap = sum la*lp, sa = sum la, un = sum (la + lp - la*lp)
    float ov = 0.f, sa = 0.f, un = 0.f;
    if (la && lp)
        for (int t = threadIdx.x; t < T; t += 128) {
            const float x = la[1LL * b * T + t], y = lp[1LL * b * T + t];
            ov = fmaf(x, y, ov); sa += x; un += x + y - x * y;
        }
    lsum[0][threadIdx.x] = ov; lsum[1][threadIdx.x] = sa; lsum[2][threadIdx.x] = un;
    __syncthreads();
    for (int o = 64; o > 0; o >>= 1) {
        if (threadIdx.x < o)
            for (int k = 0; k < 3; ++k) lsum[k][threadIdx.x] += lsum[k][threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x != 0) return;
    const double aa = dots[0], pp = dots[1], nn = dots[2], ap = dots[3], an = dots[4], s_a = dots[5], s_p = dots[6], s_n = dots[7];
    const bool has_labels = la && lp;
    float c[12] = {};
    double loss = 0.0;
    const double invB = 1.0 / B;
    if (kind == 0) {   // ImprovedTripletLoss
        const double w = has_labels ? 1.0 + 0.5 * (lsum[0][0] / (static_cast<double>(lsum[1][0]) + 1e-8)) : 1.0;
        if (!euclid) {
            const double eps = 1e-12;
            const double ra = sqrt(aa), rp = sqrt(pp), rn = sqrt(nn);
            const double na = fmax(ra, eps), np_ = fmax(rp, eps), nn_ = fmax(rn, eps);
            const double sp = ap / (na * np_), sn = an / (na * nn_);
            const double h = sn - sp + margin;    // pos_dist - neg_dist + margin
            if (h > 0.0) {
                loss = h * w;
                const double s = w * invB;
                c[0] = static_cast<float>(ra > eps ? s * (sp - sn) / (na * na) : 0.0);
                c[1] = static_cast<float>(-s / (na * np_));
                c[2] = static_cast<float>(s / (na * nn_));
                c[3] = static_cast<float>(-s / (na * np_));
                c[4] = static_cast<float>(rp > eps ? s * sp / (np_ * np_) : 0.0);
                c[5] = static_cast<float>(s / (na * nn_));
                c[6] = static_cast<float>(rn > eps ? -s * sn / (nn_ * nn_) : 0.0);
            }
        } else {
            const double eps = 1e-6;
            const double dp2 = aa - 2 * ap + pp + 2 * eps * (s_a - s_p) + D * eps * eps;
            const double dn2 = aa - 2 * an + nn + 2 * eps * (s_a - s_n) + D * eps * eps;
            const double dp = sqrt(fmax(dp2, 0.0)), dn = sqrt(fmax(dn2, 0.0));
            const double h = dp - dn + margin;
            if (h > 0.0) {
                loss = h * w;
                const double s = w * invB;
                const double ip = dp > 0.0 ? s / dp : 0.0, in_ = dn > 0.0 ? s / dn : 0.0;
                c[0] = static_cast<float>(ip - in_); c[1] = static_cast<float>(-ip); c[2] = static_cast<float>(in_);
                c[9] = static_cast<float>(eps * (ip - in_));
                c[3] = static_cast<float>(-ip); c[4] = static_cast<float>(ip); c[10] = static_cast<float>(-eps * ip);
                c[5] = static_cast<float>(in_); c[6] = static_cast<float>(-in_); c[11] = static_cast<float>(eps * in_);
            }
        }
    } else {           // ContrastiveLoss
        const double sim = lsum[0][0] / (static_cast<double>(lsum[2][0]) + 1e-8);
        const bool similar = static_cast<float>(sim) > 0.3f;
        const double w = similar ? sim : 1.0 - sim;
        double d, t;    // distance, d loss / d distance
        if (!euclid) {
            const double eps = 1e-12;
            const double ra = sqrt(aa), rp = sqrt(pp);
            const double na = fmax(ra, eps), np_ = fmax(rp, eps);
            const double s12 = ap / (na * np_);
            d = 1.0 - s12;
            if (similar) { loss = w * d * d; t = 2.0 * w * d * invB; }
            else { const double g = fmax(margin - d, 0.0); loss = w * g * g; t = -2.0 * w * g * invB; }
            // d d / d e1 = -(e2/(n1 n2) - s12 e1/n1^2)
            c[0] = static_cast<float>(ra > eps ? t * s12 / (na * na) : 0.0);
            c[1] = static_cast<float>(-t / (na * np_));
            c[3] = static_cast<float>(-t / (na * np_));
            c[4] = static_cast<float>(rp > eps ? t * s12 / (np_ * np_) : 0.0);
        } else {
            const double eps = 1e-6;
            const double d2 = aa - 2 * ap + pp + 2 * eps * (s_a - s_p) + D * eps * eps;
            d = sqrt(fmax(d2, 0.0));
            if (similar) { loss = w * d * d; t = 2.0 * w * d * invB; }
            else { const double g = fmax(margin - d, 0.0); loss = w * g * g; t = -2.0 * w * g * invB; }
            const double k = d > 0.0 ? t / d : 0.0;
            c[0] = static_cast<float>(k); c[1] = static_cast<float>(-k); c[9] = static_cast<float>(eps * k);
            c[3] = static_cast<float>(-k); c[4] = static_cast<float>(k); c[10] = static_cast<float>(-eps * k);
        }
    }
    rowloss[b] = static_cast<float>(loss);
    for (int k = 0; k < 12; ++k) coef[b * 12 + k] = c[k];
}

__global__ void mean_rows_kernel(const float* __restrict__ rowloss, int B, float* __restrict__ out) {
    double t = 0.0;
    for (int b = 0; b < B; ++b) t += rowloss[b];
    *out = static_cast<float>(t / B);
}

__global__ void __launch_bounds__(256) embed_grad_kernel(const float* __restrict__ a, const float* __restrict__ p,
                                                         const float* __restrict__ n, const float* __restrict__ coef,
                                                         float* __restrict__ ga, float* __restrict__ gp,
                                                         float* __restrict__ gn, long long D) {
    const int b = blockIdx.y;
    const float* c = coef + b * 12;
    const float c0 = c[0], c1 = c[1], c2 = c[2], c3 = c[3], c4 = c[4], c5 = c[5], c6 = c[6], k0 = c[9], k1 = c[10], k2 = c[11];
    for (long long d = blockIdx.x * 256LL + threadIdx.x; d < D; d += 256LL * gridDim.x) {
        const long long i = b * D + d;
        const float x = a[i], y = p[i], z = n ? n[i] : 0.f;
        if (ga) ga[i] = fmaf(c0, x, fmaf(c1, y, fmaf(c2, z, k0)));
        if (gp) gp[i] = fmaf(c3, x, fmaf(c4, y, k1));
        if (gn) gn[i] = fmaf(c5, x, fmaf(c6, z, k2));
    }
}

// ---- mean squared error, two fixed-order stages + gradient
__global__ void __launch_bounds__(256) mse_part_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                       float* __restrict__ part, float* __restrict__ gx, long long n,
                                                       float gscale) {
    __shared__ float sh[256];
    const long long chunk = (n + gridDim.x - 1) / gridDim.x;
    const long long i0 = blockIdx.x * chunk, i1 = min(n, i0 + chunk);
    float t = 0.f;
    for (long long i = i0 + threadIdx.x; i < i1; i += 256) {
        const float d = x[i] - y[i];
        t = fmaf(d, d, t);
        if (gx) gx[i] = gscale * d;
    }
    sh[threadIdx.x] = t;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) part[blockIdx.x] = sh[0];
}
__global__ void mse_final_kernel(const float* __restrict__ part, int chunks, double inv_n, float* __restrict__ out) {
    double t = 0.0;
    for (int k = 0; k < chunks; ++k) t += part[k];
    *out = static_cast<float>(t * inv_n);
}

// ---- AdaptiveLossWeights: w = softmax(log_w / T); total = sum w_i L_i; d total / d L_i = w_i;
// d total / d log_w_j = w_j (L_j - total) / T
__global__ void adaptive_weights_kernel(const float* __restrict__ log_w, const float* __restrict__ losses, int n, float temp,
                                        float* __restrict__ total, float* __restrict__ weights,
                                        float* __restrict__ grad_log_w) {
    double m = -1e300;
    for (int i = 0; i < n; ++i) m = fmax(m, static_cast<double>(log_w[i]) / temp);
    double z = 0.0;
    for (int i = 0; i < n; ++i) z += exp(static_cast<double>(log_w[i]) / temp - m);
    double tot = 0.0;
    for (int i = 0; i < n; ++i) {
        const double w = exp(static_cast<double>(log_w[i]) / temp - m) / z;
        weights[i] = static_cast<float>(w);
        tot += w * losses[i];
    }
    *total = static_cast<float>(tot);
    if (grad_log_w)
        for (int i = 0; i < n; ++i) grad_log_w[i] = static_cast<float>(weights[i] * (losses[i] - tot) / temp);
}

}  // namespace

size_t embed_loss_scratch_bytes(int B, long long D) {
    const int chunks = embed_loss_chunks(B, D);
    return (static_cast<size_t>(B) * chunks * NDOT + static_cast<size_t>(B) * 13) * sizeof(float) + 256;
}
int embed_loss_chunks(int B, long long D) {
    const long long want = (D + 4095) / 4096;
    return static_cast<int>(std::max<long long>(1, std::min<long long>(want, std::max(1, 148 * 4 / std::max(B, 1)))));
}

int launch_embed_loss(const vt_embed_loss_args& a, void* scratch, cudaStream_t s, Profiler* prof) {
    VT_CHECK(a.a && a.p && a.loss, "null pointers");
    VT_CHECK(a.kind == 0 || a.kind == 1, "kind: 0 triplet, 1 contrastive");
    VT_CHECK(a.kind == 1 || a.n != nullptr, "the triplet loss needs the negative embeddings");
    VT_CHECK(a.kind == 0 || (a.labels_a && a.labels_p), "the contrastive loss needs both label matrices");
    VT_CHECK((a.labels_a == nullptr) == (a.labels_p == nullptr), "label matrices go together");
    VT_CHECK(a.B > 0 && a.D > 0 && a.B < 65536, "bad batch / embedding size");
    const int chunks = embed_loss_chunks(a.B, a.D);
    float* part = static_cast<float*>(scratch);
    float* rowloss = part + static_cast<size_t>(a.B) * chunks * NDOT;
    float* coef = rowloss + a.B;
    profiler_begin(prof, KC_MISC, s, 0, 4.0 * a.B * a.D * (a.n ? 6 : 4));
    rowdots_kernel<<<dim3(chunks, a.B), 256, 0, s>>>(a.a, a.p, a.kind == 0 ? a.n : nullptr, part, a.D);
    embed_rows_kernel<<<a.B, 128, 0, s>>>(part, chunks, a.labels_a, a.labels_p, a.T, a.D, a.B, a.kind, a.similarity != 0,
                                         a.margin, rowloss, coef);
    mean_rows_kernel<<<1, 1, 0, s>>>(rowloss, a.B, a.loss);
    if (a.grad_a || a.grad_p || a.grad_n) {
        const int gx = static_cast<int>(std::max<long long>(1, std::min<long long>((a.D + 1023) / 1024, 148 * 8 / a.B + 1)));
        embed_grad_kernel<<<dim3(gx, a.B), 256, 0, s>>>(a.a, a.p, a.kind == 0 ? a.n : nullptr, coef, a.grad_a, a.grad_p,
                                                       a.kind == 0 ? a.grad_n : nullptr, a.D);
    }
    profiler_end(prof, KC_MISC, s);
    VT_CUDA(cudaGetLastError());
    return 0;
}

int launch_mse_loss(const float* x, const float* y, long long n, float* loss, float* grad_x, void* scratch, cudaStream_t s,
                    Profiler* prof) {
    VT_CHECK(x && y && loss && n > 0, "bad arguments");
    const int chunks = static_cast<int>(std::max<long long>(1, std::min<long long>((n + 4095) / 4096, 148 * 8)));
    profiler_begin(prof, KC_MISC, s, 0, 4.0 * n * (grad_x ? 3 : 2));
    mse_part_kernel<<<chunks, 256, 0, s>>>(x, y, static_cast<float*>(scratch), grad_x, n, static_cast<float>(2.0 / n));
    mse_final_kernel<<<1, 1, 0, s>>>(static_cast<float*>(scratch), chunks, 1.0 / n, loss);
    profiler_end(prof, KC_MISC, s);
    VT_CUDA(cudaGetLastError());
    return 0;
}

int launch_adaptive_weights(const float* log_w, const float* losses, int n, float temp, float* total, float* weights,
                            float* grad_log_w, cudaStream_t s) {
    VT_CHECK(log_w && losses && total && weights && n > 0 && n <= 64 && temp > 0.f, "bad arguments");
    adaptive_weights_kernel<<<1, 1, 0, s>>>(log_w, losses, n, temp, total, weights, grad_log_w);
    VT_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace vt
