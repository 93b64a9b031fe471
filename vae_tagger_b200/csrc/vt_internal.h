// Internal (non-ABI) declarations shared by the vae-tagger B200 translation units.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <algorithm>
#include <atomic>
#include <string>

namespace vt {

typedef __nv_bfloat16 bf16;

// ---- error plumbing: int return codes + thread-local message, no exceptions across the ABI
void set_error(const std::string& msg);
const char* last_error();
#define VT_CUDA(expr)                                                                                   \
    do {                                                                                                \
        cudaError_t _e = (expr);                                                                        \
        if (_e != cudaSuccess) {                                                                        \
            ::vt::set_error(std::string(#expr) + " failed: " + cudaGetErrorString(_e) + " (" __FILE__ ":" + \
                            std::to_string(__LINE__) + ")");                                            \
            return -1;                                                                                  \
        }                                                                                               \
    } while (0)
#define VT_CHECK(cond, msg)                                                           \
    do {                                                                              \
        if (!(cond)) {                                                                \
            ::vt::set_error(std::string(msg) + " [" #cond "] (" __FILE__ ":" +        \
                            std::to_string(__LINE__) + ")");                          \
            return -2;                                                                \
        }                                                                             \
    } while (0)
#define VT_TRY(expr)              \
    do {                          \
        int _r = (expr);          \
        if (_r != 0) return _r;   \
    } while (0)

// ---- cudaFuncAttributeMaxDynamicSharedMemorySize is a per-DEVICE attribute of a kernel: set it once per
// (kernel, device), not once per process (one context per device may live in the same process)
struct SmemAttrOnce {
    std::atomic<unsigned long long> done{0};   // bit d: set on device d (devices >= 64 set it on every call)
};
template <typename F>
int ensure_dyn_smem(SmemAttrOnce& once, F func, int bytes) {
    int dev = 0;
    VT_CUDA(cudaGetDevice(&dev));
    const unsigned long long bit = dev < 64 ? (1ull << dev) : 0ull;
    if (bit && (once.done.load(std::memory_order_acquire) & bit)) return 0;
    VT_CUDA(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    once.done.fetch_or(bit, std::memory_order_release);
    return 0;
}

// ---- kernel-class accounting (launch counts + optional CUDA-event timing per class)
enum KernelClass {
    KC_IGEMM = 0,      // tcgen05 implicit-GEMM (convs, projections, QK^T, PV)
    KC_GN_APPLY = 1,   // GroupNorm apply (+SiLU)
    KC_IM2COL = 2,     // conv_in patch gather
    KC_SOFTMAX = 3,    // attention row softmax
    KC_LATENT = 4,     // moments -> latent (mode/sample, scale/shift)
    KC_HEAD = 5,       // tag head kernels
    KC_FP32 = 6,       // fp32 verification-mode kernels
    KC_MISC = 7,
    // the tensor kernels that carry the step, each under its own class so that bench.py can put every one of
    // them against the roofline (KC_IGEMM keeps the remaining implicit-GEMM launches: stride-2 downsamples,
    // projections, conv_out, the decoder's sub-pixel convs, the fallbacks)
    KC_CONV3_T = 8,    // conv3_fused_kernel<128,1,true>: 128-channel layers, transposed accumulator
    KC_CONV3 = 9,      // conv3_fused_kernel<256,...>: 256/512-channel layers (CTA pairs)
    KC_FLASH = 10,     // flash_d512_kernel
    KC_CONVIN = 11,    // conv_in_kernel
    KC_BWD = 12,       // backward-pass contractions (dgrad / wgrad)
    KC_COUNT = 13
};
struct Profiler;
Profiler* profiler_create();
void profiler_destroy(Profiler*);
void profiler_enable(Profiler*, bool timing);
void profiler_begin(Profiler*, KernelClass, cudaStream_t, double flops, double bytes);
void profiler_end(Profiler*, KernelClass, cudaStream_t);
// out: [KC_COUNT][4] = launches, milliseconds, flops, bytes ; resets when reset != 0
int profiler_read(Profiler*, double* out, int reset);

// Scratch for the two-stage GroupNorm statistics of a contraction's output: the kernel's epilogue stores one
// fp32 (sum, sumsq) row of 32 groups per CTA tile, gn_finalize_kernel reduces the rows of an image in a fixed
// order.  One buffer per stream (the finalize runs right behind the producer on the same stream).
struct StatsScratch {
    float* part = nullptr;
    size_t bytes = 0;
    // several launches that together produce ONE tensor (the four parities of a sub-pixel upsample conv): launch k
    // of `parts` writes rows [k * tiles, (k+1) * tiles) of every image and only the last one reduces
    int parts = 1, part_index = 0;
};
// enough rows for every tiling the launchers use on an H x W output (>= 64-pixel 2-D patches; conv_in: row tiles
// of 256 pixels; GEMMs: 128 rows of M = H*W)
inline size_t stats_scratch_bytes(int n, int H, int W) {
    const size_t patches = ((static_cast<size_t>(H) + 7) / 8) * ((static_cast<size_t>(W) + 7) / 8);
    const size_t row_tiles = static_cast<size_t>(H) * ((static_cast<size_t>(W) + 255) / 256);
    return static_cast<size_t>(n) * std::max(patches, row_tiles) * 64 * sizeof(float) + 4096;
}

// ---- tcgen05 implicit GEMM (vt_igemm.cu)
struct ConvOp {
    // input activation, NHWC bf16
    const void* in = nullptr;  // 16-bit (tcgen05 path) or fp32 (launch_conv_fp32)
    int in_f16 = 0;            // tcgen05 path: main operand + its weight columns are fp16 (else bf16)
    int raw_f16 = 0;           // tcgen05 path: 16-bit residual / shortcut operand (+ its weight columns) are fp16 (else bf16)
    int N = 0, Hin = 0, Win = 0, Cin = 0;
    int ksize = 3;   // 1 or 3
    int stride = 1;  // 1 (pad 1 for 3x3) or 2 (pad right/bottom by 1, diffusers Downsample2D)
    // packed weights [Cout][ksize*ksize*Cin (+ Cs)] bf16, K contiguous, tap-major then channel
    const void* w = nullptr;
    int Cout = 0;
    // optional 1x1 shortcut operand folded in as an extra K-slab: [N][Hout][Wout][Cs]
    const void* sc_in = nullptr;
    int Cs = 0;
    const float* bias = nullptr;     // [Cout]
    const void* residual = nullptr;  // [N][Hout][Wout][Cout]
    int residual_fp32 = 0;           // residual element type (tcgen05 path; the fp32 path is all fp32)
    void* out = nullptr;             // [N][Hout][Wout][Cout]
    int out_fmt = 0;                 // 0 bf16, 1 fp32, 2 fp16
    double* stats = nullptr;  // [N][32][2] GroupNorm (sum, sumsq) of the output (group = Cout/32 channels)
    StatsScratch stats_ws;    // per-tile partials of the statistics (needed when stats != nullptr)
    float alpha = 1.f;
    // Sub-pixel form of "nearest 2x upsample, then conv3x3" (tcgen05 path only): the output pixels of parity
    // (up_py, up_px) of the 2Hin x 2Win result are a 2x2-tap conv of the SOURCE image whose taps are sums of the
    // 3x3 taps; w = [Cout][4*Cin] (tap (ty,tx) at K offset (ty*2+tx)*Cin), out = the full upsampled tensor.
    int up2 = 0, up_py = 0, up_px = 0;
    int kclass = -1;   // profiler class of the launch (default KC_IGEMM)
};
int launch_conv(const ConvOp& op, cudaStream_t stream, Profiler* prof);

// conv_in (3 -> 128, 3x3, pad 1) straight from the image batch: the operand rows are built in shared
// memory by gather warps (vt_convin.cuh), no patch matrix in HBM
struct ConvInOp {
    const void* img = nullptr;  // [N][3][H][W] fp32 or [N][H][W][3] u8
    int in_fmt = 0;             // VT_IN_F32_NCHW / VT_IN_U8_NHWC
    int N = 0, H = 0, W = 0;
    const void* w = nullptr;    // [128][64] fp16, k = (kh*3+kw)*3+c, k >= 27 zero
    const float* bias = nullptr;
    void* out = nullptr;        // [N][H][W][128] bf16 (out_f16: fp16)
    int out_f16 = 0;
    double* stats = nullptr;    // [N][32][2] or null
    StatsScratch stats_ws;
};
int launch_conv_in(const ConvInOp& op, cudaStream_t stream, Profiler* prof);

struct GemmOp {
    // D[b][m][n] = alpha * sum_k A[b?][m][k] * B[b?][n][k] (+ bias[n]) (+ residual[b][m][n])
    // element type of A / B / residual: bf16 (tcgen05 path) or fp32 (launch_gemm_fp32)
    const void* A = nullptr;
    const void* B = nullptr;
    int batch = 1, M = 0, N = 0, K = 0;
    int b_rows = 0;  // rows of B that exist (default N); rows b_rows..N-1 read as zero (N padded to the tile size)
    int a_batched = 1, b_batched = 1;
    long long lda = 0, ldb = 0;              // row strides in elements (default K)
    long long a_bstride = 0, b_bstride = 0;  // batch strides in elements (default rows * ld)
    int kclass = -1;   // profiler class of the launch (default KC_IGEMM)
    const float* bias = nullptr;
    const void* residual = nullptr;
    int residual_fp32 = 0;
    void* out = nullptr;
    int out_fmt = 0;            // 0 bf16, 1 fp32, 2 fp16
    int ab_f16 = 0;             // tcgen05 path: A and B are fp16 (else bf16)
    int raw_f16 = 0;            // tcgen05 path: a 16-bit residual is fp16 (else bf16)
    long long ld_out = 0;       // default N
    long long out_bstride = 0;  // default M * ld_out (also used for residual)
    double* stats = nullptr;    // [batch][32][2] over (m, group of N/32 columns)
    StatsScratch stats_ws;
    float alpha = 1.f;
};
int launch_gemm(const GemmOp& op, cudaStream_t stream, Profiler* prof);

// 16-bit tensor map, 128-byte swizzle, zero fill out of bounds (vt_igemm.cu)
int make_tmap(CUtensorMap* tm, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
              const uint32_t* box);

// Fused attention for head_dim 512 (vt_flash.cu): O = softmax(scale * Q K^T) V + bias_v, fp16 operands.
//   qk: [n][tokens][2C] (q | k), vt: [n][C][tokens] (V transposed), out: [n][tokens][C] fp16
struct FlashOp {
    const void* qk = nullptr;
    const void* vt = nullptr;
    const float* bias_v = nullptr;
    void* out = nullptr;
    int n = 0, tokens = 0, C = 512;
    long long ld_vt = 0;  // elements between rows of vt (default tokens; a multiple of 8)
    float scale = 1.f;
};
int launch_flash_attention(const FlashOp& op, cudaStream_t stream, Profiler* prof);

// 3x3 stride-1 conv with GroupNorm(32)+SiLU of the INPUT fused into the operand path (vt_conv3.cuh)
struct Conv3FusedOp {
    const void* in = nullptr;   // raw activation, 16-bit NHWC [N][H][W][Cin]: bf16, or fp16 when raw_f16
    int raw_f16 = 0;            // raw activations (in, sc_in, residual, 16-bit out) and the shortcut weight columns are fp16
    int N = 0, H = 0, W = 0, Cin = 0, Cout = 0;
    const double* gn_stats = nullptr;  // [N][32][2] (sum, sumsq) of `in`
    const float* gamma = nullptr;      // [Cin]
    const float* beta = nullptr;       // [Cin]
    float eps = 1e-6f;
    int silu = 1;
    const void* w = nullptr;           // fp16 [Cout][9*Cin (+Cs bf16 shortcut columns)], tap-major then channel
    const void* sc_in = nullptr;       // optional 1x1 shortcut operand, raw bf16 [N][H][W][Cs] (Cout >= 256 only)
    int Cs = 0;
    const float* bias = nullptr;
    const void* residual = nullptr;    // [N][H][W][Cout]
    int residual_fp32 = 0;
    void* out = nullptr;
    int out_fmt = 0;
    double* stats = nullptr;           // (sum, sumsq) of the output
    StatsScratch stats_ws;
};
int launch_conv3_fused(const Conv3FusedOp& op, cudaStream_t stream, Profiler* prof);

// ---- HBM-bound kernels (vt_elementwise.cu)
int launch_im2col3x3(const void* in, int fmt, void* out, int out_fmt, int N, int H, int W, cudaStream_t,
                     Profiler*);
int launch_gn_stats(const void* x, int x_fmt /*FMT_BF16 | FMT_F32 | FMT_F16*/, double* stats, int N, long long HW, int C, int G, cudaStream_t,
                    Profiler*);
// second stage of the epilogue statistics: part[N][tiles][G][2] fp32 -> stats[N][G][2] fp64, fixed order
int launch_gn_finalize(const float* part, double* stats, int N, int tiles, int G, cudaStream_t, Profiler*);
int launch_gn_apply(const void* x, int x_fmt /*FMT_BF16 | FMT_F32 | FMT_F16*/, void* y, int y_fmt, const double* stats, const float* gamma,
                    const float* beta, int N, long long HW, int C, int G, float eps, int silu, cudaStream_t,
                    Profiler*);
int launch_softmax_rows(const float* s, void* p, int p_fmt, long long rows, int cols, long long ld_s,
                        long long ld_p, cudaStream_t, Profiler*);
int launch_moments_to_latent(const float* moments_nhwc, float* latent, float* mean_out, float* logvar_out,
                             const float* noise, int N, int H, int W, int LC, int sample, unsigned long long seed,
                             float scale, float shift, int apply_scale, int apply_shift, cudaStream_t, Profiler*);
int launch_nchw_to_nhwc(const float* in, void* out, int out_fmt, int N, int C, long long HW, cudaStream_t);
int launch_nhwc_to_nchw(const void* in, int in_fmt, float* out, int N, int C, long long HW, cudaStream_t);
// VAE decoder I/O
int launch_latent_to_nhwc(const float* z, void* out, int out_fmt, int N, int LC, int CP, long long HW, float shift,
                          float inv_scale, cudaStream_t, Profiler*);
int launch_upsample2x_nhwc(const void* in, void* out, int elem_bytes, int N, int H, int W, int C, cudaStream_t,
                           Profiler*);
int launch_nhwc_to_image(const float* in, float* out, int N, int OC, int CP, long long HW, cudaStream_t, Profiler*);
int launch_cast_f32_16(const float* in, void* out, int out_fmt, long long n, cudaStream_t);
int launch_splitk_reduce(const float* part, int splits, long long split_stride, const float* bias, void* out,
                         int out_fmt, long long rows, int cols, long long ld_out, cudaStream_t, Profiler*);

// ---- fp32 verification mode (vt_fp32.cu): FFMA implicit GEMM, NHWC fp32, same operand packing
int launch_conv_fp32(const ConvOp& op, cudaStream_t, Profiler*);
int launch_gemm_fp32(const GemmOp& op, cudaStream_t, Profiler*);

// ---- tag head (vt_head.cu), fp32, NCHW latent
int launch_head_spatial_attention(const float* latent, const float* w1, const float* w2, const float* w7,
                                  float* pool, float* cgate, float* map2, float* out, float* sgate /*[N][HW] or null*/,
                                  int N, int C, int H, int W, cudaStream_t, Profiler*);
int launch_head_compress(const float* x, const float* cw, const float* cb, const float* bn_w, const float* bn_b,
                         const float* bn_rm, const float* bn_rv, float bn_eps, float* pooled, int N, int C, int H,
                         int W, cudaStream_t, Profiler*);
int launch_head_mhsa(const float* pooled, const float* const* params10, float* feat, float* ao /*[N][64][E] scratch*/,
                     int N, int E, int heads, int enabled, cudaStream_t, Profiler*);
int launch_head_cross_attention(const float* feat, const float* q, const float* wk, const float* bk, const float* wv,
                                const float* bv, float* out, int N, int E, int heads, cudaStream_t, Profiler*);
int launch_head_cross_add(const float* attended, const float* query, const float* in, float* out, int N, int Q, int F,
                          cudaStream_t, Profiler*);
int launch_head_adaptive_pool(const float* x, float* out, int N, int C, int H, int W, int OH, int OW, cudaStream_t,
                              Profiler*);
int launch_head_linear(const float* x, const float* w, const float* b, float* y, int B, int I, int O, cudaStream_t,
                       Profiler*);
int launch_head_ln_act(float* x, const float* w, const float* b, int B, int D, int act, cudaStream_t, Profiler*);
int launch_head_confidence(const float* logits, float* conf_sorted, long long* idx_sorted, int* count, float* probs,
                           int B, int T, float thr, cudaStream_t, Profiler*);
int launch_focal_loss(const float* logits, const float* targets, float* loss_sum, float* grad, long long n,
                      float alpha, float gamma, float grad_scale, cudaStream_t, Profiler*,
                      const float* class_w = nullptr /*[T] per-class weights*/, int T = 1);

}  // namespace vt
