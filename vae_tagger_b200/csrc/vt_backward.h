// Internal interface of the encoder backward building blocks (vt_backward.cu) used by the C-ABI layer.
#pragma once
#include "vt_internal.h"

namespace vt {

struct BwdEnv {
    cudaStream_t s = nullptr;
    Profiler* prof = nullptr;
    int fp32 = 0;        // verification mode: every tensor fp32, FFMA kernels
    int raw_fmt = 0;     // 16-bit mode: storage format of the forward activations (FMT_F16 / FMT_BF16); gradients are bf16
};

// GroupNorm(32)(+SiLU) backward over NHWC tensors.  x: forward input (raw format / fp32), dy: gradient of the output,
// stats: (sum, sumsq) of x per (image, group), add: optional tensor added to dx (gradient format).
int bwd_gn_chunks(int N, long long HW);
size_t bwd_gn_scratch_bytes(int N, long long HW, int C);
int bwd_group_norm(const BwdEnv& e, const void* x, const void* dy, const double* stats, const float* gamma,
                   const float* beta, const void* add, void* dx, float* dgamma, float* dbeta, int N, long long HW, int C,
                   float eps, int silu, int accumulate, void* scratch);

// db[c] (+)= sum over rows of g[rows][C] (gradient format)
size_t bwd_colsum_scratch_bytes(int C);
int bwd_bias_grad(const BwdEnv& e, const void* g, long long rows, int C, float* db, int accumulate, void* scratch);

// data gradient of a stride-1 conv: dx[N][H][W][Cin] = conv(dy[N][H][W][Cout], flipped / transposed weights) (+ add)
size_t bwd_dgrad_weight_bytes(const BwdEnv& e, int Cout, int Cin, int ks);
int bwd_pack_dgrad_weight(const BwdEnv& e, const float* w /*[Cout][Cin][ks][ks] fp32*/, void* dst, int Cout, int Cin, int ks);
int bwd_conv_dgrad(const BwdEnv& e, const void* dy, const void* wd, void* dx, const void* add, int N, int H, int W,
                   int Cout, int Cin, int ks);

// weight gradient, fp32 verification mode (FFMA, straight from the NHWC fp32 tensors; a = the conv's actual input)
struct WgradPlan {
    int taps, splits, batches;
    size_t part_bytes;               // fp32 split-K partial tiles
};
WgradPlan bwd_wgrad_plan(const BwdEnv& e, int N, int H, int W, int Cout, int Cin, int ks);
int bwd_conv_wgrad(const BwdEnv& e, const WgradPlan& p, const void* dy, const void* a, float* part, float* dw, int N, int H,
                   int W, int Cout, int Cin, int ks, int accumulate);

// ---- weight gradient straight from the NHWC tensors (vt_wgrad.cu): MN-major tcgen05 operands, taps = shifted TMA boxes
struct WgradMnPlan {
    int tiles_x, tiles_y, col_groups, m_blocks, splits, per_split;
    int h_tiles_x, h_tiles_y, h_groups, h_splits, h_per_split;   // halo variant (3x3 stride 1)
    size_t part_bytes;      // fp32 split-K partial tiles [splits][Cout][taps*Cin] (the larger of the two variants)
};
WgradMnPlan bwd_wgrad_mn_plan(int N, int H, int W, int Cout, int Cin, int ks);
int bwd_conv_wgrad_mn(const BwdEnv& e, const WgradMnPlan& p, const void* dy, int y_fmt, const void* a, int a_fmt, float* part,
                      int N, int H, int W, int Cout, int Cin, int ks, int stride, int* splits_used);
// 16-bit mode, whole weight gradient of one conv: the conv input as a bf16 NHWC operand in `a16` -- through
// GroupNorm(+SiLU) when stats != null, a plain fp16 -> bf16 copy otherwise (both MMA operands must share one format and
// the gradient side is bf16); a bf16 input is used as stored -- then the MN-major GEMM and the fixed-order reduce into
// dw[Cout][Cin][ks][ks].  H, W: size of dy (the conv OUTPUT); the input is (stride*H) x (stride*W).
size_t bwd_wgrad16_a16_bytes(int N, int H, int W, int Cin, int stride);
int bwd_conv_wgrad16(const BwdEnv& e, const void* dy, const void* a_src, int a_fmt, const double* stats, const float* gamma,
                     const float* beta, float eps, int silu, void* a16, float* part, float* dw, int N, int H, int W, int Cout,
                     int Cin, int ks, int stride, int accumulate);

// ---- pieces of the whole-encoder backward
int bwd_transpose(const BwdEnv& e, const void* in, int in_fmt, void* out, int out_fmt, int rows, int cols, long long ld_in,
                  long long ld_out, int batch, long long in_bs, long long out_bs);
int bwd_rowdot(const BwdEnv& e, const void* a, int a_fmt, const void* b, int b_fmt, float* out, long long rows, int cols);
int bwd_attn_ds(const BwdEnv& e, const void* P, int p_fmt, const float* dP, const float* D, void* out, long long rows, int T,
                long long tp, float scale);
int bwd_attn_ds_t(const BwdEnv& e, const void* P, int p_fmt, const float* dP, const float* D, void* dS, void* dST, int T, long long tp,
                  float scale);
int bwd_pad_channels(const BwdEnv& e, const float* in, void* out, long long rows, int C, int Cp);
// stride-2 downsample conv (diffusers Downsample2D: pad right / bottom, 3x3, stride 2), even input sizes
size_t bwd_dgrad_s2_weight_bytes(const BwdEnv& e, int Cout, int Cin);
int bwd_conv_s2_dgrad(const BwdEnv& e, const void* dy, const float* w /*OIHW fp32*/, void* wd_scratch, void* dx, int N, int Hi,
                      int Wi, int Cout, int Cin);
int bwd_conv_s2_wgrad(const BwdEnv& e, const WgradPlan& p, const void* dy, const void* x, float* part, float* dw, int N, int Hi,
                      int Wi, int Cout, int Cin, int accumulate);
// conv_in: dw[Cout][3][3][3] from the image itself (fp32 NCHW in [-1,1] or uint8 NHWC normalised to it)
int bwd_convin_chunks(int N, int H, int W);
int bwd_convin_wgrad(const BwdEnv& e, const void* dy, const void* img, int in_u8, float* part /*[chunks][Cout][27]*/, float* dw,
                     int N, int H, int W, int Cout, int accumulate);

}  // namespace vt
