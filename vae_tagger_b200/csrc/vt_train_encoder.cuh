// Training pass of the encoder: a forward that keeps every activation the backward needs (the
// "tape"), and the backward over the whole schedule -- what autograd does for the reference when it fine-tunes the
// VAE (train_full.py:201-256: triplet / contrastive loss on posterior samples; train_vae.py:124-186).  Included by
// vt_api.cu behind the encoder schedule (EncRun, run_attention, AttnPlan).  Same kernels and formats as inference for
// the forward; the backward is built from vt_backward.cu:
//   conv 3x3 / 1x1 stride 1 ... bwd_conv_dgrad (forward tcgen05 kernel on flipped weights) + bwd_conv_wgrad16 (vt_wgrad.cu:
//                               MN-major tcgen05 GEMM over the pixels, split-K with a fixed-order reduce)
//   Downsample2D (stride 2) ... bwd_conv_s2_dgrad (sub-pixel form: four 2x2-tap convs of the output gradient) +
//                               bwd_conv_wgrad16 on the (2C, W/2, 2, H/2, N) parity view of the input
//   GroupNorm(+SiLU) .......... bwd_group_norm (the residual / shortcut gradient is added in its apply pass)
//   attention ................. scores and probabilities are rebuilt per image (S = q k^T, P = softmax), then
//                               dV = P^T dO, dP = dO V^T, dS = P o (dP - rowsum(dO o O)) / sqrt(C), dQ = dS K, dK = dS^T Q
//                               as tcgen05 GEMMs on bf16 operands; the four projections are 1x1 convs
//   conv_in ................... weight gradient straight from the image (K = 27)
//   conv_out .................. 32 moment channels padded to one 64-wide chunk
// Gradients are written (or accumulated) into caller buffers bound by parameter name (vt_encoder_grad_bind).
#pragma once

namespace {

struct TapeOp {
    int kind = 0;                 // 0 ResnetBlock2D, 1 Downsample2D, 2 attention, 3 Upsample2D (nearest 2x + conv3x3)
    std::string prefix;           // parameter name prefix
    const ResnetW* res = nullptr;
    Act x;                        // input of the op (raw format)
    const double* st_x = nullptr; // GroupNorm statistics of x
    Act h;                        // ResnetBlock2D: conv1 output
    const double* st_h = nullptr;
    int H = 0, W = 0;             // spatial size of x
    int cin = 0, cout = 0;
    // attention: normalised tokens, [q|k], V^T, O (operand format)
    void *Tn = nullptr, *QK = nullptr, *Vt = nullptr, *O = nullptr;
};

}  // namespace

struct EncTape {
    bool valid = false;
    int n = 0, H = 0, W = 0, fp32 = 0, in_fmt = 0;
    const void* images = nullptr;
    DevBuf arena, stats, statpart, mom;
    Act x0;                      // conv_in output
    const double* st_x0 = nullptr;
    std::vector<TapeOp> ops;
    Act xf;                      // input of conv_norm_out
    const double* st_xf = nullptr;
    int lh = 0, lw = 0;
    // decoder tapes: the latent in NHWC padded to one K chunk (conv_in's operand) and the un-scale factor applied to it
    Act z;
    float inv_scale = 1.f;
    void release() {
        arena.release(); stats.release(); statpart.release(); mom.release();
    }
};

namespace {

// [N][LC][hw] mean / logvar gradients (NCHW fp32, either may be null) -> [N][hw][CP] moment gradient (gradient format)
template <int FO>
__global__ void moments_grad_kernel(const float* __restrict__ gm, const float* __restrict__ gl, void* __restrict__ out, int LC,
                                    int CP, long long HW, long long total) {
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += 256LL * gridDim.x) {
        const int c = static_cast<int>(i % CP);
        const long long np = i / CP;
        const long long n = np / HW, p = np - n * HW;
        float v = 0.f;
        if (c < LC) v = gm ? gm[(n * LC + c) * HW + p] : 0.f;
        else if (c < 2 * LC) v = gl ? gl[(n * LC + (c - LC)) * HW + p] : 0.f;
        if constexpr (FO == FMT_F32) static_cast<float*>(out)[i] = v;
        else static_cast<bf16*>(out)[i] = __float2bfloat16(v);
    }
}

// dx[n][y][x][c] = sum of the 2x2 block of du[n][2y..2y+1][2x..2x+1][c]  (backward of the nearest-neighbour 2x upsample)
template <int F>
__global__ void __launch_bounds__(256) sumpool2x2_kernel(const void* __restrict__ du, void* __restrict__ dx, int N, int H, int W,
                                                         int C) {
    const int c8n = C / 8;
    const long long total = 1LL * N * H * W * c8n;
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += 256LL * gridDim.x) {
        const int c8 = static_cast<int>(i % c8n);
        long long r = i / c8n;
        const int x = static_cast<int>(r % W); r /= W;
        const int y = static_cast<int>(r % H);
        const long long nn = r / H;
        float acc[8] = {};
#pragma unroll
        for (int dy = 0; dy < 2; ++dy)
#pragma unroll
            for (int dxx = 0; dxx < 2; ++dxx) {
                const long long off = ((nn * 2 * H + 2 * y + dy) * 2 * W + 2 * x + dxx) * C + c8 * 8;
                float v[8];
                if constexpr (F == FMT_F32) {
                    const float4 a = *reinterpret_cast<const float4*>(static_cast<const float*>(du) + off);
                    const float4 b = *reinterpret_cast<const float4*>(static_cast<const float*>(du) + off + 4);
                    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
                } else {
                    const uint4 u = *reinterpret_cast<const uint4*>(static_cast<const bf16*>(du) + off);
                    v[0] = bf16_lo(u.x); v[1] = bf16_hi(u.x); v[2] = bf16_lo(u.y); v[3] = bf16_hi(u.y);
                    v[4] = bf16_lo(u.z); v[5] = bf16_hi(u.z); v[6] = bf16_lo(u.w); v[7] = bf16_hi(u.w);
                }
#pragma unroll
                for (int e = 0; e < 8; ++e) acc[e] += v[e];
            }
        const long long o = ((nn * H + y) * W + x) * C + c8 * 8;
        if constexpr (F == FMT_F32) {
            *reinterpret_cast<float4*>(static_cast<float*>(dx) + o) = make_float4(acc[0], acc[1], acc[2], acc[3]);
            *reinterpret_cast<float4*>(static_cast<float*>(dx) + o + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
        } else {
            *reinterpret_cast<uint4*>(static_cast<bf16*>(dx) + o) =
                make_uint4(pack_bf16x2(acc[0], acc[1]), pack_bf16x2(acc[2], acc[3]), pack_bf16x2(acc[4], acc[5]), pack_bf16x2(acc[6], acc[7]));
        }
    }
}
// image gradient NCHW fp32 [N][OC][HW] -> NHWC [N][HW][CP] gradient format (channels >= OC zero)
template <int FO>
__global__ void image_grad_kernel(const float* __restrict__ g, void* __restrict__ out, int OC, int CP, long long HW, long long total) {
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += 256LL * gridDim.x) {
        const int c = static_cast<int>(i % CP);
        const long long np = i / CP;
        const long long n = np / HW, p = np - n * HW;
        const float v = c < OC ? g[(n * OC + c) * HW + p] : 0.f;
        if constexpr (FO == FMT_F32) static_cast<float*>(out)[i] = v;
        else static_cast<bf16*>(out)[i] = __float2bfloat16(v);
    }
}
// latent gradient: NHWC [N][HW][CP] (gradient format) -> NCHW fp32 [N][LC][HW], times the un-scale factor of the forward
template <int FI>
__global__ void latent_grad_kernel(const void* __restrict__ g, float* __restrict__ out, int LC, int CP, long long HW, float mul,
                                   long long total, int accumulate) {
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += 256LL * gridDim.x) {
        const long long p = i % HW;
        long long r = i / HW;
        const int c = static_cast<int>(r % LC);
        const long long n = r / LC;
        float v;
        if constexpr (FI == FMT_F32) v = static_cast<const float*>(g)[(n * HW + p) * CP + c];
        else v = __bfloat162float(static_cast<const bf16*>(g)[(n * HW + p) * CP + c]);
        out[i] = (accumulate ? out[i] : 0.f) + v * mul;
    }
}
// dst[co][ci][k] (ci < Cd) (+)= src[co][ci][k] (ci < Cs): the leading input channels of a padded weight gradient
__global__ void weight_subset_kernel(float* __restrict__ dst, const float* __restrict__ src, int Co, int Cd, int Cs, int K, int accumulate) {
    const long long total = 1LL * Co * Cd * K;
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += 256LL * gridDim.x) {
        const int k = static_cast<int>(i % K);
        long long r = i / K;
        const int ci = static_cast<int>(r % Cd);
        const long long co = r / Cd;
        dst[i] = (accumulate ? dst[i] : 0.f) + src[(co * Cs + ci) * K + k];
    }
}
// padded copy of a conv weight: dst[co][ci][k] = src[co][ci][k] for co < Co, ci < Ci, zero elsewhere ([CoP][CiP][K])
__global__ void weight_pad_kernel(float* __restrict__ dst, const float* __restrict__ src, int Co, int Ci, int CoP, int CiP, int K) {
    const long long total = 1LL * CoP * CiP * K;
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += 256LL * gridDim.x) {
        const int k = static_cast<int>(i % K);
        long long r = i / K;
        const int ci = static_cast<int>(r % CiP);
        const long long co = r / CiP;
        dst[i] = (co < Co && ci < Ci) ? src[(co * Ci + ci) * K + k] : 0.f;
    }
}
__global__ void copy_or_add_kernel(float* __restrict__ dst, const float* __restrict__ src, long long n, int accumulate) {
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += 256LL * gridDim.x)
        dst[i] = (accumulate ? dst[i] : 0.f) + src[i];
}
// rows [r0, r0 + rows) of a [R][cols] matrix copied / added into a [rows][cols] destination
int copy_or_add(float* dst, const float* src, long long n, int accumulate, cudaStream_t s) {
    copy_or_add_kernel<<<static_cast<int>(std::min<long long>((n + 255) / 256, 1184)), 256, 0, s>>>(dst, src, n, accumulate);
    VT_CUDA(cudaGetLastError());
    return 0;
}

struct EncBwd {
    vt_ctx* c;
    EncTape* t;
    BwdEnv e;
    int acc;          // accumulate into the bound gradient buffers
    int n;
    size_t es;        // bytes per activation / gradient element
    int gf, xf, of;   // gradient / raw / operand formats
    // which half of the VAE: parameter tensors, bound gradient buffers, mid-block attention weights
    std::map<std::string, Param>* params = nullptr;
    std::map<std::string, float*>* grads = nullptr;
    const AttnW* attnw = nullptr;
    const char* what = "encoder";

    float* G(const std::string& name, int* err) {
        auto it = grads->find(name);
        if (it == grads->end() || it->second == nullptr) {
            set_error(std::string("no gradient buffer bound for ") + what + " parameter " + name);
            *err = -4;
            return nullptr;
        }
        return it->second;
    }
    const float* Wt(const std::string& name) { return (*params)[name].dev; }

    // scratch: bump allocations, grown on demand, shared by every tape of the context.  bind() carves the scratch of
    // ONE helper call out of c->tbws (every helper starts again at its beginning); bind_layer() carves what a whole
    // layer keeps across helper calls (c->tlws)
    Carver cv;
    int bind() { int r = cv.bind(c->tbws); cv.items.clear(); return r; }
    int bind_layer() { int r = cv.bind(c->tlws); cv.items.clear(); return r; }

    // weight + bias gradient of a stride-1 conv (ks 1 or 3) whose input is `a_src`:
    //   16-bit mode: a_src is a raw NHWC tensor, optionally seen through GroupNorm(+SiLU) (st/gamma/beta non-null)
    //   fp32 mode:   a_src is the conv's actual fp32 input (the caller applies the normalisation)
    int conv_wgrad(const void* dy, const void* a_src, int a_fmt, const double* st, const float* gamma, const float* beta,
                   int silu, int H, int W, int Cout, int Cin, int ks, float* dw, float* db) {
        void *a16 = nullptr, *cs = nullptr;
        float* part = nullptr;
        if (e.fp32) {
            const WgradPlan p = bwd_wgrad_plan(e, n, H, W, Cout, Cin, ks);
            cv.want(&part, p.part_bytes); cv.want(&cs, bwd_colsum_scratch_bytes(Cout));
            VT_TRY(bind());
            VT_TRY(bwd_conv_wgrad(e, p, dy, a_src, part, dw, n, H, W, Cout, Cin, ks, acc));
        } else {
            cv.want(&a16, bwd_wgrad16_a16_bytes(n, H, W, Cin, 1));
            cv.want(&part, bwd_wgrad_mn_plan(n, H, W, Cout, Cin, ks).part_bytes);
            cv.want(&cs, bwd_colsum_scratch_bytes(Cout));
            VT_TRY(bind());
            VT_TRY(bwd_conv_wgrad16(e, dy, a_src, a_fmt, st, gamma, beta, 1e-6f, silu, a16, part, dw, n, H, W, Cout, Cin, ks, 1, acc));
        }
        if (db) VT_TRY(bwd_bias_grad(e, dy, 1LL * n * H * W, Cout, db, acc, cs));
        return 0;
    }
    int conv_dgrad(const void* dy, const float* w_oihw, void* dx, const void* add, int H, int W, int Cout, int Cin, int ks) {
        void* wd = nullptr;
        cv.want(&wd, bwd_dgrad_weight_bytes(e, Cout, Cin, ks));
        VT_TRY(bind());
        VT_TRY(bwd_pack_dgrad_weight(e, w_oihw, wd, Cout, Cin, ks));
        return bwd_conv_dgrad(e, dy, wd, dx, add, n, H, W, Cout, Cin, ks);
    }
    int gn_bwd(const void* x, const void* dy, const double* st, const std::string& norm, const void* add, void* dx, long long HW,
               int C, int silu) {
        int err = 0;
        float* dg = G(norm + ".weight", &err);
        float* db = G(norm + ".bias", &err);
        if (err) return err;
        void* sc = nullptr;
        cv.want(&sc, bwd_gn_scratch_bytes(n, HW, C));
        VT_TRY(bind());
        return bwd_group_norm(e, x, dy, st, Wt(norm + ".weight"), Wt(norm + ".bias"), add, dx, dg, db, n, HW, C, 1e-6f, silu, acc, sc);
    }
    // fp32 mode: the normalised operand a conv saw, rebuilt into `T`
    int normalised(const void* x, const double* st, const std::string& norm, void* T, long long HW, int C, int silu) {
        return launch_gn_apply(x, FMT_F32, T, FMT_F32, st, Wt(norm + ".weight"), Wt(norm + ".bias"), n, HW, C, 32, 1e-6f, silu, e.s,
                               c->prof);
    }

    // ---- ResnetBlock2D.  dOut: gradient of the block output; dX receives the gradient of the block input.
    // tmpA / tmpB: two more gradient-sized buffers.
    int resnet(const TapeOp& op, const void* dOut, void* dX, void* tmpA, void* tmpB) {
        const int H = op.H, W = op.W, Cin = op.cin, Cout = op.cout;
        const long long HW = 1LL * H * W;
        const std::string& p = op.prefix;
        int err = 0;
        float *gw1 = G(p + ".conv1.weight", &err), *gb1 = G(p + ".conv1.bias", &err);
        float *gw2 = G(p + ".conv2.weight", &err), *gb2 = G(p + ".conv2.bias", &err);
        if (err) return err;
        const bool sc = Cin != Cout;
        void* T = nullptr;   // fp32 mode: the normalised operand of a conv, rebuilt
        cv.want(&T, e.fp32 ? static_cast<size_t>(n) * HW * std::max(Cin, Cout) * 4 : 0);
        VT_TRY(bind_layer());
        // conv2: weight / bias gradients (operand = silu(norm2(h))), data gradient -> tmpA
        if (e.fp32) {
            VT_TRY(normalised(op.h.p, op.st_h, p + ".norm2", T, HW, Cout, 1));
            VT_TRY(conv_wgrad(dOut, T, FMT_F32, nullptr, nullptr, nullptr, 0, H, W, Cout, Cout, 3, gw2, gb2));
        } else {
            VT_TRY(conv_wgrad(dOut, op.h.p, xf, op.st_h, Wt(p + ".norm2.weight"), Wt(p + ".norm2.bias"), 1, H, W, Cout, Cout, 3, gw2, gb2));
        }
        if (sc) {
            float *gws = G(p + ".conv_shortcut.weight", &err), *gbs = G(p + ".conv_shortcut.bias", &err);
            if (err) return err;
            VT_TRY(conv_wgrad(dOut, op.x.p, xf, nullptr, nullptr, nullptr, 0, H, W, Cout, Cin, 1, gws, gbs));
        }
        VT_TRY(conv_dgrad(dOut, Wt(p + ".conv2.weight"), tmpA, nullptr, H, W, Cout, Cout, 3));
        // norm2 + SiLU -> dH in tmpB
        VT_TRY(gn_bwd(op.h.p, tmpA, op.st_h, p + ".norm2", nullptr, tmpB, HW, Cout, 1));
        // conv1
        if (e.fp32) {
            VT_TRY(normalised(op.x.p, op.st_x, p + ".norm1", T, HW, Cin, 1));
            VT_TRY(conv_wgrad(tmpB, T, FMT_F32, nullptr, nullptr, nullptr, 0, H, W, Cout, Cin, 3, gw1, gb1));
        } else {
            VT_TRY(conv_wgrad(tmpB, op.x.p, xf, op.st_x, Wt(p + ".norm1.weight"), Wt(p + ".norm1.bias"), 1, H, W, Cout, Cin, 3, gw1, gb1));
        }
        VT_TRY(conv_dgrad(tmpB, Wt(p + ".conv1.weight"), tmpA, nullptr, H, W, Cout, Cin, 3));
        // shortcut branch gradient, added inside norm1's apply pass
        const void* add = dOut;
        if (sc) {
            VT_TRY(conv_dgrad(dOut, Wt(p + ".conv_shortcut.weight"), tmpB, nullptr, H, W, Cout, Cin, 1));
            add = tmpB;
        }
        return gn_bwd(op.x.p, tmpA, op.st_x, p + ".norm1", add, dX, HW, Cin, 1);
    }

    // ---- Downsample2D: x [H][W] -> [H/2][W/2]
    int down(const TapeOp& op, const void* dOut, void* dX) {
        const int Hi = op.H, Wi = op.W, C = op.cin, Ho = Hi / 2, Wo = Wi / 2;
        const std::string& p = op.prefix;
        int err = 0;
        float *gw = G(p + ".weight", &err), *gb = G(p + ".bias", &err);
        if (err) return err;
        void *cs = nullptr, *wd = nullptr, *a16 = nullptr;
        float* part = nullptr;
        const WgradPlan pl = bwd_wgrad_plan(e, n, Ho, Wo, C, C, 3);
        cv.want(&a16, e.fp32 ? 0 : bwd_wgrad16_a16_bytes(n, Ho, Wo, C, 2));
        cv.want(&part, e.fp32 ? pl.part_bytes : bwd_wgrad_mn_plan(n, Ho, Wo, C, C, 3).part_bytes);
        cv.want(&cs, bwd_colsum_scratch_bytes(C)); cv.want(&wd, bwd_dgrad_s2_weight_bytes(e, C, C));
        VT_TRY(bind());
        if (e.fp32) {
            VT_TRY(bwd_conv_s2_wgrad(e, pl, dOut, op.x.p, part, gw, n, Hi, Wi, C, C, acc));
        } else {
            // the conv input as stored (a raw activation), read through the stride-2 parity view
            VT_TRY(bwd_conv_wgrad16(e, dOut, op.x.p, xf, nullptr, nullptr, nullptr, 0.f, 0, a16, part, gw, n, Ho, Wo, C, C, 3, 2, acc));
        }
        VT_TRY(bwd_bias_grad(e, dOut, 1LL * n * Ho * Wo, C, gb, acc, cs));
        return bwd_conv_s2_dgrad(e, dOut, Wt(p + ".weight"), wd, dX, n, Hi, Wi, C, C);
    }


    // ---- Upsample2D: nearest 2x of x [H][W] then conv3x3 -> [2H][2W].  The upsampled operand is rebuilt into `U`
    // (raw format), the data gradient of the conv is summed over each 2x2 block.
    int upsample(const TapeOp& op, const void* dOut, void* dX, void* tmpA) {
        const int H = op.H, W = op.W, C = op.cin;
        const std::string& p = op.prefix;
        int err = 0;
        float *gw = G(p + ".weight", &err), *gb = G(p + ".bias", &err);
        if (err) return err;
        void* U = nullptr;
        cv.want(&U, static_cast<size_t>(n) * 4 * H * W * C * es);
        VT_TRY(bind_layer());
        VT_TRY(launch_upsample2x_nhwc(op.x.p, U, static_cast<int>(es), n, H, W, C, e.s, c->prof));
        VT_TRY(conv_wgrad(dOut, U, xf, nullptr, nullptr, nullptr, 0, 2 * H, 2 * W, C, C, 3, gw, gb));
        VT_TRY(conv_dgrad(dOut, Wt(p + ".weight"), tmpA, nullptr, 2 * H, 2 * W, C, C, 3));
        const long long total = 1LL * n * H * W * (C / 8);
        const int grid = static_cast<int>(std::min<long long>((total + 255) / 256, 148 * 16));
        if (e.fp32) sumpool2x2_kernel<FMT_F32><<<grid, 256, 0, e.s>>>(tmpA, dX, n, H, W, C);
        else sumpool2x2_kernel<FMT_BF16><<<grid, 256, 0, e.s>>>(tmpA, dX, n, H, W, C);
        VT_CUDA(cudaGetLastError());
        return 0;
    }

    // ---- attention.  tmpA / tmpB: gradient-sized buffers ([n][tokens][C])
    int attention(const TapeOp& op, const void* dOut, void* dX, void* tmpA, void* tmpB) {
        const int h = op.H, w_ = op.W, C = op.cin;
        const long long T = 1LL * h * w_, tp = (T + 63) / 64 * 64;
        const std::string& p = op.prefix;
        const float scale = 1.0f / sqrtf(static_cast<float>(C));
        int err = 0;
        float *gwq = G(p + ".to_q.weight", &err), *gbq = G(p + ".to_q.bias", &err), *gwk = G(p + ".to_k.weight", &err),
              *gbk = G(p + ".to_k.bias", &err), *gwv = G(p + ".to_v.weight", &err), *gbv = G(p + ".to_v.bias", &err),
              *gwo = G(p + ".to_out.0.weight", &err), *gbo = G(p + ".to_out.0.bias", &err);
        if (err) return err;
        const size_t ge = es;                    // gradient element size (= operand element size in both modes)
        const int bfmt = e.fp32 ? FMT_F32 : FMT_BF16;   // operands of the backward GEMMs
        // ---- per-image softmax backward on rebuilt scores: everything the layer keeps across helper calls
        void *S, *P, *PT, *dS, *dST, *Vf, *QT, *KT, *dOT, *dQK, *dV;
        float* D;
        float *gqk = nullptr, *gbqk = nullptr;
        void* wqk = nullptr;
        cv.want(&gqk, static_cast<size_t>(2) * C * C * 4); cv.want(&gbqk, static_cast<size_t>(2) * C * 4);
        cv.want(&wqk, static_cast<size_t>(2) * C * C * 4);
        cv.want(&S, static_cast<size_t>(tp) * tp * 4); cv.want(&P, static_cast<size_t>(tp) * tp * ge);
        cv.want(&PT, static_cast<size_t>(tp) * tp * ge); cv.want(&dS, static_cast<size_t>(tp) * tp * ge);
        cv.want(&dST, static_cast<size_t>(tp) * tp * ge);
        cv.want(&Vf, static_cast<size_t>(n) * T * C * ge); cv.want(&QT, static_cast<size_t>(C) * tp * ge);
        cv.want(&KT, static_cast<size_t>(C) * tp * ge); cv.want(&dOT, static_cast<size_t>(C) * tp * ge);
        cv.want(&dQK, static_cast<size_t>(n) * T * 2 * C * ge); cv.want(&dV, static_cast<size_t>(n) * T * C * ge);
        cv.want(&D, static_cast<size_t>(n) * T * 4);
        VT_TRY(bind_layer());
        // ---- out projection: dWo = dOut^T O, dbo, dO = dOut Wo  (1x1 convs over the token grid)
        VT_TRY(conv_wgrad(dOut, op.O, of, nullptr, nullptr, nullptr, 0, h, w_, C, C, 1, gwo, gbo));
        void* dO = tmpA;
        VT_TRY(conv_dgrad(dOut, Wt(p + ".to_out.0.weight"), dO, nullptr, h, w_, C, C, 1));
        const size_t sq = static_cast<size_t>(tp) * tp * ge;
        VT_CUDA(cudaMemsetAsync(PT, 0, sq, e.s));
        VT_CUDA(cudaMemsetAsync(dST, 0, sq, e.s));
        VT_CUDA(cudaMemsetAsync(P, 0, sq, e.s));
        VT_CUDA(cudaMemsetAsync(QT, 0, static_cast<size_t>(C) * tp * ge, e.s));
        VT_CUDA(cudaMemsetAsync(KT, 0, static_cast<size_t>(C) * tp * ge, e.s));
        VT_CUDA(cudaMemsetAsync(dOT, 0, static_cast<size_t>(C) * tp * ge, e.s));
        auto gemm = [&](GemmOp& g) -> int {
            g.kclass = KC_BWD;
            if (e.fp32) { g.out_fmt = FMT_F32; return launch_gemm_fp32(g, e.s, c->prof); }
            return launch_gemm(g, e.s, c->prof);
        };
        {   // V (with its bias: O = P (v + b_v)) in token-major layout, backward-operand format: [n][T][C]
            GemmOp g;
            g.A = op.Tn; g.B = e.fp32 ? static_cast<const void*>(attnw->v.w32) : static_cast<const void*>(attnw->v.w16);
            g.batch = n; g.M = static_cast<int>(T); g.N = C; g.K = C; g.a_batched = 1; g.b_batched = 0;
            g.bias = attnw->v.bias; g.out = Vf; g.out_fmt = bfmt; g.ab_f16 = 1;
            VT_TRY(gemm(g));
        }
        // D[q] = sum_c dO[q][c] O[q][c]
        VT_TRY(bwd_rowdot(e, dO, bfmt, op.O, of, D, 1LL * n * T, C));
        const char* QK = static_cast<const char*>(op.QK);
        for (int i = 0; i < n; ++i) {
            const char* q_i = QK + static_cast<size_t>(i) * T * 2 * C * es;
            const char* dO_i = static_cast<const char*>(dO) + static_cast<size_t>(i) * T * C * ge;
            const char* V_i = static_cast<const char*>(Vf) + static_cast<size_t>(i) * T * C * ge;
            {   // S = scale q k^T
                GemmOp g;
                g.A = q_i; g.lda = 2 * C; g.B = q_i + static_cast<size_t>(C) * es; g.ldb = 2 * C;
                g.batch = 1; g.M = static_cast<int>(T); g.N = static_cast<int>(tp); g.b_rows = static_cast<int>(T); g.K = C;
                g.alpha = scale; g.out = S; g.out_fmt = FMT_F32; g.ab_f16 = 1;
                VT_TRY(gemm(g));
            }
            VT_TRY(launch_softmax_rows(static_cast<const float*>(S), P, of, T, static_cast<int>(T), tp, tp, e.s, c->prof));
            {   // dP = dO V^T  (into the score buffer)
                GemmOp g;
                g.A = dO_i; g.B = V_i; g.batch = 1; g.M = static_cast<int>(T); g.N = static_cast<int>(tp); g.b_rows = static_cast<int>(T);
                g.K = C; g.out = S; g.out_fmt = FMT_F32; g.ab_f16 = 0;
                VT_TRY(gemm(g));
            }
            // dS, and the transposes (square, zero padded) and [C][tokens] views of q, k, dO
            if (!e.fp32 && of == FMT_F16) {
                VT_TRY(bwd_attn_ds_t(e, P, of, static_cast<const float*>(S), D + 1LL * i * T, dS, dST, static_cast<int>(T), tp, scale));
            } else {
                VT_TRY(bwd_attn_ds(e, P, of, static_cast<const float*>(S), D + 1LL * i * T, dS, T, static_cast<int>(T), tp, scale));
                VT_TRY(bwd_transpose(e, dS, bfmt, dST, bfmt, static_cast<int>(T), static_cast<int>(T), tp, tp, 1, 0, 0));
            }
            VT_TRY(bwd_transpose(e, P, of, PT, bfmt, static_cast<int>(T), static_cast<int>(T), tp, tp, 1, 0, 0));
            VT_TRY(bwd_transpose(e, q_i, of, QT, bfmt, static_cast<int>(T), C, 2 * C, tp, 1, 0, 0));
            VT_TRY(bwd_transpose(e, q_i + static_cast<size_t>(C) * es, of, KT, bfmt, static_cast<int>(T), C, 2 * C, tp, 1, 0, 0));
            VT_TRY(bwd_transpose(e, dO_i, bfmt, dOT, bfmt, static_cast<int>(T), C, C, tp, 1, 0, 0));
            {   // dV[k][c] = sum_q P[q][k] dO[q][c]
                GemmOp g;
                g.A = PT; g.lda = tp; g.B = dOT; g.ldb = tp; g.batch = 1; g.M = static_cast<int>(T); g.N = C; g.K = static_cast<int>(tp);
                g.out = static_cast<char*>(dV) + static_cast<size_t>(i) * T * C * ge; g.out_fmt = bfmt; g.ab_f16 = 0;
                VT_TRY(gemm(g));
            }
            {   // dQ[q][c] = sum_k dS[q][k] K[k][c]  -> left half of [dQ | dK]
                GemmOp g;
                g.A = dS; g.lda = tp; g.B = KT; g.ldb = tp; g.batch = 1; g.M = static_cast<int>(T); g.N = C; g.K = static_cast<int>(tp);
                g.out = static_cast<char*>(dQK) + static_cast<size_t>(i) * T * 2 * C * ge; g.ld_out = 2 * C; g.out_fmt = bfmt; g.ab_f16 = 0;
                VT_TRY(gemm(g));
            }
            {   // dK[k][c] = sum_q dS[q][k] Q[q][c]  -> right half
                GemmOp g;
                g.A = dST; g.lda = tp; g.B = QT; g.ldb = tp; g.batch = 1; g.M = static_cast<int>(T); g.N = C; g.K = static_cast<int>(tp);
                g.out = static_cast<char*>(dQK) + (static_cast<size_t>(i) * T * 2 * C + C) * ge; g.ld_out = 2 * C; g.out_fmt = bfmt;
                g.ab_f16 = 0;
                VT_TRY(gemm(g));
            }
        }
        // ---- projections: weight / bias gradients (operand = normalised tokens), data gradient into the tokens
        // [q|k] is one stacked 1x1 conv with 2C outputs: gradient buffers of to_q / to_k are separate tensors, so the
        // stacked weight gradient goes through a scratch buffer
        const int acc_saved = acc;
        acc = 0;
        VT_TRY(conv_wgrad(dQK, op.Tn, of, nullptr, nullptr, nullptr, 0, h, w_, 2 * C, C, 1, gqk, gbqk));
        acc = acc_saved;
        VT_TRY(copy_or_add(gwq, gqk, 1LL * C * C, acc, e.s));
        VT_TRY(copy_or_add(gwk, gqk + 1LL * C * C, 1LL * C * C, acc, e.s));
        VT_TRY(copy_or_add(gbq, gbqk, C, acc, e.s));
        VT_TRY(copy_or_add(gbk, gbqk + C, C, acc, e.s));
        VT_TRY(conv_wgrad(dV, op.Tn, of, nullptr, nullptr, nullptr, 0, h, w_, C, C, 1, gwv, gbv));
        // dTn = dQ Wq + dK Wk + dV Wv
        VT_CUDA(cudaMemcpyAsync(wqk, Wt(p + ".to_q.weight"), static_cast<size_t>(C) * C * 4, cudaMemcpyDeviceToDevice, e.s));
        VT_CUDA(cudaMemcpyAsync(static_cast<float*>(wqk) + 1LL * C * C, Wt(p + ".to_k.weight"), static_cast<size_t>(C) * C * 4,
                                cudaMemcpyDeviceToDevice, e.s));
        VT_TRY(conv_dgrad(dQK, static_cast<const float*>(wqk), tmpB, nullptr, h, w_, 2 * C, C, 1));
        VT_TRY(conv_dgrad(dV, Wt(p + ".to_v.weight"), tmpA, tmpB, h, w_, C, C, 1));
        // ---- group_norm (no activation), residual added in the apply pass
        return gn_bwd(op.x.p, tmpA, op.st_x, p + ".group_norm", dOut, dX, T, C, 0);
    }
};

int run_encoder_train_forward(vt_ctx* c, const vt_encode_args* a, int slot) {
    const vt_encoder_config& cfg = c->ecfg;
    if (c->tapes[slot] == nullptr) c->tapes[slot] = new EncTape();
    EncTape& tp = *c->tapes[slot];
    tp.valid = false;
    tp.ops.clear();
    const int H = a->height, Wd = a->width, n = a->batch;
    const int fp32 = a->precision == VT_PREC_FP32;
    const size_t es = fp32 ? 4 : 2;
    const int nb = cfg.num_blocks;
    const int LC = cfg.latent_channels;
    const int C0 = cfg.block_out_channels[0];
    const int Cm = cfg.block_out_channels[nb - 1];
    VT_CHECK(H % (1 << (nb - 1)) == 0 && Wd % (1 << (nb - 1)) == 0 && H >= 8 && Wd >= 8,
             "the training pass needs image sizes divisible by 8 (every level halves exactly)");
    const int lh = H >> (nb - 1), lw = Wd >> (nb - 1);
    const long long tokens = 1LL * lh * lw;
    cudaStream_t s = static_cast<cudaStream_t>(a->stream);

    // ---- arena: every activation of the pass + one scratch buffer for normalised operands
    const AttnPlan pl = plan_attention(cfg.mid_block_add_attention != 0, n, tokens, Cm, es, fp32);
    size_t total = 0;
    auto sz = [&](long long hw, int C) { return align_up(static_cast<size_t>(n) * hw * C * es, 1024); };
    const size_t scratch_b = sz(1LL * H * Wd, C0);
    total += scratch_b + sz(1LL * H * Wd, C0);   // T scratch, conv_in output
    {
        int hh = H, ww = Wd;
        for (int b = 0; b < nb; ++b) {
            total += 2 * cfg.layers_per_block * sz(1LL * hh * ww, cfg.block_out_channels[b]);
            if (b < nb - 1) { hh /= 2; ww /= 2; total += sz(1LL * hh * ww, cfg.block_out_channels[b]); }
        }
        total += 4 * sz(tokens, Cm) + 2 * sz(tokens, Cm) + pl.attn_bytes;   // two mid resnets, attention (Tn, out, workspace)
    }
    VT_TRY(tp.arena.ensure(total));
    char* base = static_cast<char*>(tp.arena.p);
    size_t off = 0;
    auto take = [&](size_t bytes) { void* p = base + off; off += align_up(bytes, 1024); return p; };
    void* T = take(scratch_b);

    const int groups = cfg.norm_num_groups;
    const int max_slots = 64;
    VT_TRY(tp.stats.ensure(static_cast<size_t>(max_slots) * n * groups * 2 * sizeof(double)));
    VT_TRY(tp.statpart.ensure(stats_scratch_bytes(n, H, Wd)));
    VT_TRY(tp.mom.ensure(static_cast<size_t>(n) * tokens * 2 * LC * sizeof(float)));
    EncRun R{c, s, fp32, n, static_cast<double*>(tp.stats.p), 0, groups};
    R.ws.part = static_cast<float*>(tp.statpart.p); R.ws.bytes = tp.statpart.cap;
    {
        const char* e = getenv("VT_B200_NO_FUSED_GN");
        R.use_fused = !(e && e[0] == '1');
        const char* f = getenv("VT_B200_NO_FLASH");
        R.use_flash = !(f && f[0] == '1');
    }
    auto peek_stats = [&]() { return R.stats_base + static_cast<size_t>(R.stats_used) * n * groups * 2; };

    // ---- conv_in
    const char* img = static_cast<const char*>(a->images);
    double* st_x = R.new_stats();
    Act X{take(sz(1LL * H * Wd, C0)), R.raw_fmt()};
    const bool convin_direct = !fp32 && C0 == 128 && c->conv_in.f16;
    if (convin_direct) {
        ConvInOp op;
        op.img = img; op.in_fmt = a->in_fmt; op.N = n; op.H = H; op.W = Wd;
        op.w = c->conv_in.w16; op.bias = c->conv_in.bias; op.out = X.p; op.stats = st_x; op.stats_ws = R.ws;
        op.out_f16 = R.raw16();
        VT_TRY(launch_conv_in(op, s, c->prof));
    } else {
        VT_TRY(launch_im2col3x3(img, a->in_fmt, T, R.opd_fmt(), n, H, Wd, s, c->prof));
        ConvW w = c->conv_in;
        w.Cin = 64; w.ksize = 1; w.Cs = 0;
        VT_TRY(R.conv(T, H, Wd, w, 1, nullptr, nullptr, X, st_x));
    }
    tp.x0 = X; tp.st_x0 = st_x;

    auto resnet = [&](const ResnetW& rw, const std::string& prefix, int hh, int ww) -> int {
        TapeOp op;
        op.kind = 0; op.prefix = prefix; op.res = &rw; op.x = X; op.st_x = st_x; op.H = hh; op.W = ww; op.cin = rw.cin; op.cout = rw.cout;
        op.h = Act{take(sz(1LL * hh * ww, rw.cout)), R.raw_fmt()};
        Act out{take(sz(1LL * hh * ww, rw.cout)), R.raw_fmt()};
        double* st_o = R.new_stats();
        op.st_h = peek_stats();   // EncRun::resnet draws the next slot for conv1's output
        VT_TRY(R.resnet(rw, X, st_x, hh, ww, 0, T, op.h.p, out, st_o));
        tp.ops.push_back(op);
        X = out; st_x = st_o;
        return 0;
    };

    int h = H, w_ = Wd;
    for (int b = 0; b < nb; ++b) {
        for (int l = 0; l < cfg.layers_per_block; ++l)
            VT_TRY(resnet(c->down[b][l], "down_blocks." + std::to_string(b) + ".resnets." + std::to_string(l), h, w_));
        if (c->downsample[b].Cout != 0) {
            TapeOp op;
            op.kind = 1; op.prefix = "down_blocks." + std::to_string(b) + ".downsamplers.0.conv"; op.x = X; op.st_x = st_x;
            op.H = h; op.W = w_; op.cin = op.cout = c->downsample[b].Cout;
            double* st_o = R.new_stats();
            Act out{take(sz(1LL * (h / 2) * (w_ / 2), op.cout)), R.raw_fmt()};
            VT_TRY(R.conv(X.p, h, w_, c->downsample[b], 2, nullptr, nullptr, out, st_o));
            tp.ops.push_back(op);
            X = out; st_x = st_o;
            h /= 2; w_ /= 2;
        }
    }
    VT_TRY(resnet(c->mid0, "mid_block.resnets.0", h, w_));
    if (cfg.mid_block_add_attention) {
        TapeOp op;
        op.kind = 2; op.prefix = "mid_block.attentions.0"; op.x = X; op.st_x = st_x; op.H = h; op.W = w_; op.cin = op.cout = Cm;
        op.Tn = take(sz(tokens, Cm));
        char* ab = static_cast<char*>(take(pl.attn_bytes));
        op.QK = ab; op.Vt = ab + pl.qk_b; op.O = ab + pl.qk_b + pl.vt_b + pl.s_b + pl.p_b;
        double* st_o = R.new_stats();
        Act out{take(sz(tokens, Cm)), R.raw_fmt()};
        VT_TRY(run_attention(R, c->attn, pl, ab, X, st_x, op.Tn, out, st_o, n, h, w_, fp32, es));
        tp.ops.push_back(op);
        X = out; st_x = st_o;
    }
    VT_TRY(resnet(c->mid1, "mid_block.resnets.1", h, w_));
    VT_CHECK(off <= total, "training arena overflow");
    tp.xf = X; tp.st_xf = st_x;
    // ---- conv_norm_out + SiLU + conv_out -> moments -> DiagonalGaussian outputs
    VT_TRY(R.gn(X, T, st_x, c->norm_out, tokens, Cm, 1));
    VT_TRY(R.conv(T, h, w_, c->conv_out, 1, nullptr, nullptr, Act{tp.mom.p, FMT_F32}, nullptr));
    VT_TRY(launch_moments_to_latent(static_cast<const float*>(tp.mom.p), a->latent, a->mean, a->logvar, a->noise, n, h, w_, LC,
                                    a->sample, a->seed, cfg.scaling_factor, cfg.shift_factor,
                                    a->apply_scale_shift && cfg.has_scaling_factor, a->apply_scale_shift && cfg.has_shift_factor,
                                    s, c->prof));
    VT_CHECK(R.stats_used <= max_slots, "GroupNorm statistics slots exhausted");
    tp.n = n; tp.H = H; tp.W = Wd; tp.fp32 = fp32; tp.in_fmt = a->in_fmt; tp.images = a->images; tp.lh = h; tp.lw = w_;
    tp.valid = true;
    return 0;
}

int run_encoder_backward(vt_ctx* c, const vt_encoder_backward_args* a) {
    VT_CHECK(a->slot >= 0 && a->slot < VT_MAX_TAPES && c->tapes[a->slot] != nullptr && c->tapes[a->slot]->valid,
             "vt_encoder_backward needs a preceding vt_encoder_train_forward on the same slot");
    EncTape& tp = *c->tapes[a->slot];
    const vt_encoder_config& cfg = c->ecfg;
    const int n = tp.n, nb = cfg.num_blocks, LC = cfg.latent_channels;
    const int C0 = cfg.block_out_channels[0], Cm = cfg.block_out_channels[nb - 1];
    const long long tokens = 1LL * tp.lh * tp.lw;
    EncBwd B{c, &tp, BwdEnv{}, a->accumulate != 0, n, static_cast<size_t>(tp.fp32 ? 4 : 2), 0, 0, 0};
    B.params = &c->eparams; B.grads = &c->egrads; B.attnw = &c->attn; B.what = "encoder";
    B.e.s = static_cast<cudaStream_t>(a->stream); B.e.prof = c->prof; B.e.fp32 = tp.fp32;
    B.e.raw_fmt = c->raw_f16 ? FMT_F16 : FMT_BF16;
    B.gf = tp.fp32 ? FMT_F32 : FMT_BF16; B.xf = tp.fp32 ? FMT_F32 : B.e.raw_fmt; B.of = tp.fp32 ? FMT_F32 : FMT_F16;
    const size_t gmax = align_up(static_cast<size_t>(n) * tp.H * tp.W * C0 * B.es, 1024);
    // three rotating gradient buffers (current gradient, two temporaries of the layer) + the next gradient
    VT_TRY(c->tg0.ensure(4 * gmax));
    char* gb = static_cast<char*>(c->tg0.p);
    void* cur = gb;
    void* nxt = gb + gmax;
    void* tmpA = gb + 2 * gmax;
    void* tmpB = gb + 3 * gmax;
    cudaStream_t s = B.e.s;
    int err = 0;

    // ---- conv_out: moment gradient [n][tokens][64] (2*LC used)
    const int CP = 64;
    {
        void* dM = tmpA;
        const long long total = 1LL * n * tokens * CP;
        const int grid = static_cast<int>(std::min<long long>((total + 255) / 256, 1184));
        if (tp.fp32) moments_grad_kernel<FMT_F32><<<grid, 256, 0, s>>>(a->grad_mean, a->grad_logvar, dM, LC, CP, tokens, total);
        else moments_grad_kernel<FMT_BF16><<<grid, 256, 0, s>>>(a->grad_mean, a->grad_logvar, dM, LC, CP, tokens, total);
        VT_CUDA(cudaGetLastError());
        float *gw = B.G("conv_out.weight", &err), *gbias = B.G("conv_out.bias", &err);
        if (err) return err;
        // weights / gradients padded to 64 output channels in a scratch buffer that survives the nested binds
        float *w64 = nullptr, *gw64 = nullptr, *gb64 = nullptr;
        void* Tn = nullptr;
        B.cv.want(&w64, static_cast<size_t>(CP) * Cm * 9 * 4); B.cv.want(&gw64, static_cast<size_t>(CP) * Cm * 9 * 4);
        B.cv.want(&gb64, CP * 4); B.cv.want(&Tn, tp.fp32 ? static_cast<size_t>(n) * tokens * Cm * 4 : 0);
        VT_TRY(B.bind_layer());
        VT_CUDA(cudaMemsetAsync(w64, 0, static_cast<size_t>(CP) * Cm * 9 * 4, s));
        VT_CUDA(cudaMemcpyAsync(w64, B.Wt("conv_out.weight"), static_cast<size_t>(2 * LC) * Cm * 9 * 4, cudaMemcpyDeviceToDevice, s));
        const int acc_saved = B.acc;
        B.acc = 0;
        if (tp.fp32) {
            VT_TRY(B.normalised(tp.xf.p, tp.st_xf, "conv_norm_out", Tn, tokens, Cm, 1));
            VT_TRY(B.conv_wgrad(dM, Tn, FMT_F32, nullptr, nullptr, nullptr, 0, tp.lh, tp.lw, CP, Cm, 3, gw64, gb64));
        } else {
            VT_TRY(B.conv_wgrad(dM, tp.xf.p, B.xf, tp.st_xf, B.Wt("conv_norm_out.weight"), B.Wt("conv_norm_out.bias"), 1, tp.lh, tp.lw,
                                CP, Cm, 3, gw64, gb64));
        }
        B.acc = acc_saved;
        VT_TRY(copy_or_add(gw, gw64, 1LL * 2 * LC * Cm * 9, B.acc, s));
        VT_TRY(copy_or_add(gbias, gb64, 2 * LC, B.acc, s));
        VT_TRY(B.conv_dgrad(dM, w64, tmpB, nullptr, tp.lh, tp.lw, CP, Cm, 3));
        // conv_norm_out + SiLU
        VT_TRY(B.gn_bwd(tp.xf.p, tmpB, tp.st_xf, "conv_norm_out", nullptr, cur, tokens, Cm, 1));
    }
    // ---- the tape in reverse
    for (int i = static_cast<int>(tp.ops.size()) - 1; i >= 0; --i) {
        const TapeOp& op = tp.ops[i];
        if (op.kind == 0) VT_TRY(B.resnet(op, cur, nxt, tmpA, tmpB));
        else if (op.kind == 1) VT_TRY(B.down(op, cur, nxt));
        else VT_TRY(B.attention(op, cur, nxt, tmpA, tmpB));
        std::swap(cur, nxt);
    }
    // ---- conv_in: weight / bias gradients (no data gradient: the image is the leaf)
    {
        float *gw = B.G("conv_in.weight", &err), *gbias = B.G("conv_in.bias", &err);
        if (err) return err;
        float* part = nullptr;
        void* cs = nullptr;
        B.cv.want(&part, static_cast<size_t>(bwd_convin_chunks(n, tp.H, tp.W)) * C0 * 27 * 4); B.cv.want(&cs, bwd_colsum_scratch_bytes(C0));
        VT_TRY(B.bind());
        VT_TRY(bwd_convin_wgrad(B.e, cur, tp.images, tp.in_fmt == VT_IN_U8_NHWC, part, gw, n, tp.H, tp.W, C0, B.acc));
        VT_TRY(bwd_bias_grad(B.e, cur, 1LL * n * tp.H * tp.W, C0, gbias, B.acc, cs));
    }
    return 0;
}

// =========================================================================================================
// Decoder half (diffusers Decoder; the reference back-propagates the reconstruction MSE through vae.decode in
// train_vae.py:124-186 and CombinedLoss, improved_losses.py:278)
int run_decoder_train_forward(vt_ctx* c, const vt_decode_args* a, int slot) {
    const vt_encoder_config& cfg = c->ecfg;
    if (c->dtapes[slot] == nullptr) c->dtapes[slot] = new EncTape();
    EncTape& tp = *c->dtapes[slot];
    tp.valid = false;
    tp.ops.clear();
    const int n = a->batch, lh = a->lat_h, lw = a->lat_w;
    const int fp32 = a->precision == VT_PREC_FP32;
    const size_t es = fp32 ? 4 : 2;
    const int nb = cfg.num_blocks;
    const int LC = cfg.latent_channels;
    const int H = lh << (nb - 1), Wd = lw << (nb - 1);
    const int Cm = cfg.block_out_channels[nb - 1], C0 = cfg.block_out_channels[0];
    const long long tokens = 1LL * lh * lw;
    cudaStream_t s = static_cast<cudaStream_t>(a->stream);

    const AttnPlan pl = plan_attention(cfg.mid_block_add_attention != 0, n, tokens, Cm, es, fp32);
    auto sz = [&](long long hw, int C) { return align_up(static_cast<size_t>(n) * hw * C * es, 1024); };
    size_t total = 0, scratch_b = sz(tokens, std::max(Cm, 64));
    {
        total += sz(tokens, 64) + sz(tokens, Cm);                          // padded latent, conv_in output
        total += 4 * sz(tokens, Cm) + 2 * sz(tokens, Cm) + pl.attn_bytes;   // mid resnets, attention
        int hh = lh, ww = lw;
        for (int b = 0; b < nb; ++b) {
            const int cout = cfg.block_out_channels[nb - 1 - b];
            total += 2 * static_cast<size_t>(c->up[b].size()) * sz(1LL * hh * ww, cout);
            scratch_b = std::max(scratch_b, sz(1LL * hh * ww, std::max(cout, c->up[b][0].cin)));
            if (c->upsample[b].Cout != 0) {
                hh *= 2; ww *= 2;
                total += sz(1LL * hh * ww, cout);
                scratch_b = std::max(scratch_b, sz(1LL * hh * ww, cout));
            }
        }
        scratch_b = std::max(scratch_b, align_up(static_cast<size_t>(n) * H * Wd * 32 * 4, 1024));   // padded fp32 conv_out tile
        total += 2 * scratch_b;
    }
    VT_TRY(tp.arena.ensure(total));
    char* base = static_cast<char*>(tp.arena.p);
    size_t off = 0;
    auto take = [&](size_t bytes) { void* p = base + off; off += align_up(bytes, 1024); return p; };
    void* T = take(scratch_b);
    void* T2 = take(scratch_b);
    const int groups = cfg.norm_num_groups;
    const int max_slots = 64;
    VT_TRY(tp.stats.ensure(static_cast<size_t>(max_slots) * n * groups * 2 * sizeof(double)));
    VT_TRY(tp.statpart.ensure(stats_scratch_bytes(n, H, Wd)));
    EncRun R{c, s, fp32, n, static_cast<double*>(tp.stats.p), 0, groups};
    R.ws.part = static_cast<float*>(tp.statpart.p); R.ws.bytes = tp.statpart.cap;
    {
        const char* e = getenv("VT_B200_NO_FUSED_GN");
        R.use_fused = !(e && e[0] == '1');
        const char* f = getenv("VT_B200_NO_FLASH");
        R.use_flash = !(f && f[0] == '1');
    }
    auto peek_stats = [&]() { return R.stats_base + static_cast<size_t>(R.stats_used) * n * groups * 2; };

    const float shift = (a->apply_scale_shift && cfg.has_shift_factor) ? cfg.shift_factor : 0.f;
    const float inv_scale = (a->apply_scale_shift && cfg.has_scaling_factor) ? 1.0f / cfg.scaling_factor : 1.0f;
    tp.z = Act{take(sz(tokens, 64)), R.raw_fmt()};
    tp.inv_scale = inv_scale;
    VT_TRY(launch_latent_to_nhwc(a->latent, tp.z.p, R.raw_fmt(), n, LC, c->dconv_in.Cin, tokens, shift, inv_scale, s, c->prof));
    double* st_x = R.new_stats();
    Act X{take(sz(tokens, Cm)), R.raw_fmt()};
    VT_TRY(R.conv(tp.z.p, lh, lw, c->dconv_in, 1, nullptr, nullptr, X, st_x));
    tp.x0 = X; tp.st_x0 = st_x;

    auto resnet = [&](const ResnetW& rw, const std::string& prefix, int hh, int ww) -> int {
        TapeOp op;
        op.kind = 0; op.prefix = prefix; op.res = &rw; op.x = X; op.st_x = st_x; op.H = hh; op.W = ww; op.cin = rw.cin; op.cout = rw.cout;
        op.h = Act{take(sz(1LL * hh * ww, rw.cout)), R.raw_fmt()};
        Act out{take(sz(1LL * hh * ww, rw.cout)), R.raw_fmt()};
        double* st_o = R.new_stats();
        op.st_h = peek_stats();
        VT_TRY(R.resnet(rw, X, st_x, hh, ww, 0, T, op.h.p, out, st_o));
        tp.ops.push_back(op);
        X = out; st_x = st_o;
        return 0;
    };
    int h = lh, w_ = lw;
    VT_TRY(resnet(c->dmid0, "mid_block.resnets.0", h, w_));
    if (cfg.mid_block_add_attention) {
        TapeOp op;
        op.kind = 2; op.prefix = "mid_block.attentions.0"; op.x = X; op.st_x = st_x; op.H = h; op.W = w_; op.cin = op.cout = Cm;
        op.Tn = take(sz(tokens, Cm));
        char* ab = static_cast<char*>(take(pl.attn_bytes));
        op.QK = ab; op.Vt = ab + pl.qk_b; op.O = ab + pl.qk_b + pl.vt_b + pl.s_b + pl.p_b;
        double* st_o = R.new_stats();
        Act out{take(sz(tokens, Cm)), R.raw_fmt()};
        VT_TRY(run_attention(R, c->dattn, pl, ab, X, st_x, op.Tn, out, st_o, n, h, w_, fp32, es));
        tp.ops.push_back(op);
        X = out; st_x = st_o;
    }
    VT_TRY(resnet(c->dmid1, "mid_block.resnets.1", h, w_));
    for (int b = 0; b < nb; ++b) {
        for (size_t l = 0; l < c->up[b].size(); ++l)
            VT_TRY(resnet(c->up[b][l], "up_blocks." + std::to_string(b) + ".resnets." + std::to_string(l), h, w_));
        if (c->upsample[b].Cout != 0) {
            const ConvW& uw = c->upsample[b];
            TapeOp op;
            op.kind = 3; op.prefix = "up_blocks." + std::to_string(b) + ".upsamplers.0.conv"; op.x = X; op.st_x = st_x;
            op.H = h; op.W = w_; op.cin = op.cout = uw.Cin;
            double* st_o = R.new_stats();
            Act out{take(sz(4LL * h * w_, uw.Cout)), R.raw_fmt()};
            // explicit form (the sub-pixel form of inference computes the same function with pre-summed taps)
            VT_TRY(launch_upsample2x_nhwc(X.p, T2, static_cast<int>(es), n, h, w_, uw.Cin, s, c->prof));
            h *= 2; w_ *= 2;
            VT_TRY(R.conv(T2, h, w_, uw, 1, nullptr, nullptr, out, st_o));
            tp.ops.push_back(op);
            X = out; st_x = st_o;
        }
    }
    VT_CHECK(off <= total, "training arena overflow");
    tp.xf = X; tp.st_xf = st_x;
    VT_TRY(R.gn(X, T, st_x, c->dnorm_out, 1LL * h * w_, C0, 1));
    VT_TRY(R.conv(T, h, w_, c->dconv_out, 1, nullptr, nullptr, Act{T2, FMT_F32}, nullptr));
    VT_TRY(launch_nhwc_to_image(static_cast<const float*>(T2), a->image, n, 3, c->dconv_out.Cout, 1LL * H * Wd, s, c->prof));
    VT_CHECK(R.stats_used <= max_slots, "GroupNorm statistics slots exhausted");
    tp.n = n; tp.H = H; tp.W = Wd; tp.fp32 = fp32; tp.lh = lh; tp.lw = lw;
    tp.valid = true;
    return 0;
}

int run_decoder_backward(vt_ctx* c, const vt_decoder_backward_args* a) {
    VT_CHECK(a->slot >= 0 && a->slot < VT_MAX_TAPES && c->dtapes[a->slot] != nullptr && c->dtapes[a->slot]->valid,
             "vt_decoder_backward needs a preceding vt_decoder_train_forward on the same slot");
    EncTape& tp = *c->dtapes[a->slot];
    const vt_encoder_config& cfg = c->ecfg;
    const int n = tp.n, nb = cfg.num_blocks, LC = cfg.latent_channels;
    const int C0 = cfg.block_out_channels[0], Cm = cfg.block_out_channels[nb - 1];
    const long long HW = 1LL * tp.H * tp.W, tokens = 1LL * tp.lh * tp.lw;
    EncBwd B{c, &tp, BwdEnv{}, a->accumulate != 0, n, static_cast<size_t>(tp.fp32 ? 4 : 2), 0, 0, 0};
    B.params = &c->dparams; B.grads = &c->dgrads; B.attnw = &c->dattn; B.what = "decoder";
    B.e.s = static_cast<cudaStream_t>(a->stream); B.e.prof = c->prof; B.e.fp32 = tp.fp32;
    B.e.raw_fmt = c->raw_f16 ? FMT_F16 : FMT_BF16;
    B.gf = tp.fp32 ? FMT_F32 : FMT_BF16; B.xf = tp.fp32 ? FMT_F32 : B.e.raw_fmt; B.of = tp.fp32 ? FMT_F32 : FMT_F16;
    cudaStream_t s = B.e.s;
    // the largest gradient tensor of the schedule (FLUX: the 256-channel upsample conv at full resolution)
    size_t gel = static_cast<size_t>(n) * HW * std::max(C0, 64);
    for (const TapeOp& op : tp.ops) {
        const size_t px = static_cast<size_t>(n) * op.H * op.W;
        gel = std::max(gel, px * std::max(op.cin, op.cout) * (op.kind == 3 ? 4 : 1));
    }
    const size_t gmax = align_up(gel * B.es, 1024);
    VT_TRY(c->tg0.ensure(4 * gmax));
    char* gb = static_cast<char*>(c->tg0.p);
    void* cur = gb;
    void* nxt = gb + gmax;
    void* tmpA = gb + 2 * gmax;
    void* tmpB = gb + 3 * gmax;
    int err = 0;
    const int CP = 64;
    {   // ---- conv_out (C0 -> 3, padded to 64 outputs) + conv_norm_out
        void* dI = tmpA;
        const long long total = 1LL * n * HW * CP;
        const int grid = static_cast<int>(std::min<long long>((total + 255) / 256, 148 * 16));
        if (tp.fp32) image_grad_kernel<FMT_F32><<<grid, 256, 0, s>>>(a->grad_image, dI, 3, CP, HW, total);
        else image_grad_kernel<FMT_BF16><<<grid, 256, 0, s>>>(a->grad_image, dI, 3, CP, HW, total);
        VT_CUDA(cudaGetLastError());
        float *gw = B.G("conv_out.weight", &err), *gbias = B.G("conv_out.bias", &err);
        if (err) return err;
        float *w64 = nullptr, *gw64 = nullptr, *gb64 = nullptr;
        void* Tn = nullptr;
        B.cv.want(&w64, static_cast<size_t>(CP) * C0 * 9 * 4); B.cv.want(&gw64, static_cast<size_t>(CP) * C0 * 9 * 4);
        B.cv.want(&gb64, CP * 4); B.cv.want(&Tn, tp.fp32 ? static_cast<size_t>(n) * HW * C0 * 4 : 0);
        VT_TRY(B.bind_layer());
        VT_CUDA(cudaMemsetAsync(w64, 0, static_cast<size_t>(CP) * C0 * 9 * 4, s));
        VT_CUDA(cudaMemcpyAsync(w64, B.Wt("conv_out.weight"), static_cast<size_t>(3) * C0 * 9 * 4, cudaMemcpyDeviceToDevice, s));
        const int acc_saved = B.acc;
        B.acc = 0;
        if (tp.fp32) {
            VT_TRY(B.normalised(tp.xf.p, tp.st_xf, "conv_norm_out", Tn, HW, C0, 1));
            VT_TRY(B.conv_wgrad(dI, Tn, FMT_F32, nullptr, nullptr, nullptr, 0, tp.H, tp.W, CP, C0, 3, gw64, gb64));
        } else {
            VT_TRY(B.conv_wgrad(dI, tp.xf.p, B.xf, tp.st_xf, B.Wt("conv_norm_out.weight"), B.Wt("conv_norm_out.bias"), 1, tp.H, tp.W,
                                CP, C0, 3, gw64, gb64));
        }
        B.acc = acc_saved;
        VT_TRY(copy_or_add(gw, gw64, 1LL * 3 * C0 * 9, B.acc, s));
        VT_TRY(copy_or_add(gbias, gb64, 3, B.acc, s));
        VT_TRY(B.conv_dgrad(dI, w64, tmpB, nullptr, tp.H, tp.W, CP, C0, 3));
        VT_TRY(B.gn_bwd(tp.xf.p, tmpB, tp.st_xf, "conv_norm_out", nullptr, cur, HW, C0, 1));
    }
    for (int i = static_cast<int>(tp.ops.size()) - 1; i >= 0; --i) {
        const TapeOp& op = tp.ops[i];
        if (op.kind == 0) VT_TRY(B.resnet(op, cur, nxt, tmpA, tmpB));
        else if (op.kind == 2) VT_TRY(B.attention(op, cur, nxt, tmpA, tmpB));
        else VT_TRY(B.upsample(op, cur, nxt, tmpA));
        std::swap(cur, nxt);
    }
    {   // ---- conv_in (latent padded to 64 channels -> Cm): weight / bias gradients, data gradient -> latent gradient
        float *gw = B.G("conv_in.weight", &err), *gbias = B.G("conv_in.bias", &err);
        if (err) return err;
        float *w64 = nullptr, *gw64 = nullptr;
        void* cs = nullptr;
        B.cv.want(&w64, static_cast<size_t>(Cm) * CP * 9 * 4); B.cv.want(&gw64, static_cast<size_t>(Cm) * CP * 9 * 4);
        B.cv.want(&cs, bwd_colsum_scratch_bytes(Cm));
        VT_TRY(B.bind_layer());
        {   // weight gradient against the padded operand (written, not accumulated), then its LC leading input channels
            const int keep = B.acc;
            B.acc = 0;
            VT_TRY(B.conv_wgrad(cur, tp.z.p, B.xf, nullptr, nullptr, nullptr, 0, tp.lh, tp.lw, Cm, CP, 3, gw64, nullptr));
            B.acc = keep;
        }
        VT_TRY(bwd_bias_grad(B.e, cur, 1LL * n * tokens, Cm, gbias, B.acc, cs));
        const long long wt = 1LL * Cm * LC * 9;
        weight_subset_kernel<<<static_cast<int>(std::min<long long>((wt + 255) / 256, 1184)), 256, 0, s>>>(gw, gw64, Cm, LC, CP, 9, B.acc);
        VT_CUDA(cudaGetLastError());
        if (a->grad_latent) {
            const long long wp = 1LL * Cm * CP * 9;
            weight_pad_kernel<<<static_cast<int>(std::min<long long>((wp + 255) / 256, 1184)), 256, 0, s>>>(w64, B.Wt("conv_in.weight"), Cm, LC, Cm, CP, 9);
            VT_CUDA(cudaGetLastError());
            VT_TRY(B.conv_dgrad(cur, w64, nxt, nullptr, tp.lh, tp.lw, Cm, CP, 3));
            const long long total = 1LL * n * LC * tokens;
            const int grid = static_cast<int>(std::min<long long>((total + 255) / 256, 1184));
            if (tp.fp32) latent_grad_kernel<FMT_F32><<<grid, 256, 0, s>>>(nxt, a->grad_latent, LC, CP, tokens, tp.inv_scale, total, 0);
            else latent_grad_kernel<FMT_BF16><<<grid, 256, 0, s>>>(nxt, a->grad_latent, LC, CP, tokens, tp.inv_scale, total, 0);
            VT_CUDA(cudaGetLastError());
        }
    }
    return 0;
}

}  // namespace
