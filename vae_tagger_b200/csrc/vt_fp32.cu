// fp32 verification mode: the same contractions as vt_igemm.cu (implicit-GEMM conv over NHWC
// activations with the shortcut K-slab, batched NT GEMM) on the FFMA pipe, fp32 in / fp32
// accumulate / fp32 out.  It exists for the "fp32 mode: latent relative L2 <= 1e-4" bar of the
// north star (single-pass tensor-core math cannot meet it) and as an independent on-device
// cross-check of the tcgen05 path.  64x64 output tile, K step 16, 256 threads x (4x4) outputs.
#include "vt_internal.h"

namespace vt {

struct F32Problem {
    // A operand: conv gather (ksize 1/3, stride 1/2) over NHWC input, optional shortcut slab
    const float* in;
    int Hin, Win, Cin, ksize, stride, Hout, Wout;
    const float* sc_in;  // [.. Hout, Wout, Cs]
    int Cs;
    // plain GEMM A (when in == nullptr): A[b][m][k], row stride lda, batch stride a_bs (0 = shared)
    const float* A;
    long long lda, a_bs;
    // B operand: [n][k] row stride ldb, batch stride b_bs (0 = shared)
    const float* B;
    long long ldb, b_bs;
    int M;  // rows per batch (Hout*Wout for convs)
    int N, K, batch;
    int b_rows;  // rows of B that exist (rows b_rows..N-1 read as zero)
    const float* bias;
    const float* residual;
    float* out;
    long long ld_out;  // out/residual row stride; batch stride = M * ld_out
    float alpha;
};

constexpr int F_BM = 64, F_BN = 64, F_BK = 16;

__global__ void __launch_bounds__(256) f32_contract_kernel(const F32Problem P) {
    __shared__ float As[F_BK][F_BM + 4];
    __shared__ float Bs[F_BK][F_BN + 4];
    const int b = blockIdx.z;
    const int m0 = blockIdx.x * F_BM, n0 = blockIdx.y * F_BN;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;  // 16 x 16 threads, 4x4 outputs each
    float acc[4][4] = {};

    // loader mapping: each thread loads 4 consecutive k of one row (A) and one row (B)
    const int lrow = threadIdx.x >> 2;        // 0..63
    const int lk = (threadIdx.x & 3) * 4;     // 0,4,8,12
    const int am = m0 + lrow;
    int oy = 0, ox = 0;
    if (P.in != nullptr && am < P.M) { oy = am / P.Wout; ox = am - oy * P.Wout; }
    const int pad = (P.ksize == 3 && P.stride == 1) ? 1 : 0;
    const int Kconv = P.in ? P.ksize * P.ksize * P.Cin : 0;

    for (int k0 = 0; k0 < P.K; k0 += F_BK) {
        // ---- A tile
        float4 av = make_float4(0.f, 0.f, 0.f, 0.f);
        const int k = k0 + lk;
        if (am < P.M && k < P.K) {
            if (P.in != nullptr) {
                if (k < Kconv) {
                    const int tap = k / P.Cin, ci = k - tap * P.Cin;  // Cin % 4 == 0: the 4 k share a tap
                    const int kh = tap / P.ksize, kw = tap - kh * P.ksize;
                    const int iy = oy * P.stride + kh - pad, ix = ox * P.stride + kw - pad;
                    if (iy >= 0 && iy < P.Hin && ix >= 0 && ix < P.Win)
                        av = *reinterpret_cast<const float4*>(P.in + ((1LL * b * P.Hin + iy) * P.Win + ix) * P.Cin + ci);
                } else {
                    av = *reinterpret_cast<const float4*>(P.sc_in + (1LL * b * P.M + am) * P.Cs + (k - Kconv));
                }
            } else {
                av = *reinterpret_cast<const float4*>(P.A + b * P.a_bs + 1LL * am * P.lda + k);
            }
        }
        As[lk + 0][lrow] = av.x; As[lk + 1][lrow] = av.y; As[lk + 2][lrow] = av.z; As[lk + 3][lrow] = av.w;
        // ---- B tile
        float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
        const int bn = n0 + lrow;
        if (bn < P.b_rows && k < P.K) bv = *reinterpret_cast<const float4*>(P.B + b * P.b_bs + 1LL * bn * P.ldb + k);
        Bs[lk + 0][lrow] = bv.x; Bs[lk + 1][lrow] = bv.y; Bs[lk + 2][lrow] = bv.z; Bs[lk + 3][lrow] = bv.w;
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < F_BK; ++kk) {
            const float4 a4 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
            const float4 b4 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
            const float a[4] = {a4.x, a4.y, a4.z, a4.w};
            const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= P.M) continue;
        const long long off = (1LL * b * P.M + m) * P.ld_out;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n >= P.N) continue;
            float v = acc[i][j] * P.alpha;
            if (P.bias) v += P.bias[n];
            if (P.residual) v += P.residual[off + n];
            P.out[off + n] = v;
        }
    }
}

static int launch_f32(const F32Problem& P, cudaStream_t s, Profiler* prof) {
    VT_CHECK(P.K % 4 == 0, "fp32 contraction needs K % 4 == 0");
    dim3 grid((P.M + F_BM - 1) / F_BM, (P.N + F_BN - 1) / F_BN, P.batch);
    VT_CHECK(grid.y < 65536 && grid.z < 65536, "fp32 contraction grid too large");
    profiler_begin(prof, KC_FP32, s, 2.0 * P.batch * P.M * static_cast<double>(P.N) * P.K, 0);
    f32_contract_kernel<<<grid, 256, 0, s>>>(P);
    profiler_end(prof, KC_FP32, s);
    VT_CUDA(cudaGetLastError());
    return 0;
}

// weights: [Cout][ksize*ksize*Cin (+Cs)] fp32, same packing as the bf16 path
int launch_conv_fp32(const ConvOp& op, cudaStream_t s, Profiler* prof) {
    VT_CHECK(op.Cin % 4 == 0 && op.Cs % 4 == 0, "fp32 conv needs channel counts divisible by 4");
    F32Problem P{};
    P.in = reinterpret_cast<const float*>(op.in);
    P.Hin = op.Hin; P.Win = op.Win; P.Cin = op.Cin; P.ksize = op.ksize; P.stride = op.stride;
    P.Hout = op.stride == 1 ? op.Hin : op.Hin / 2;
    P.Wout = op.stride == 1 ? op.Win : op.Win / 2;
    P.sc_in = reinterpret_cast<const float*>(op.sc_in);
    P.Cs = op.sc_in ? op.Cs : 0;
    P.K = op.ksize * op.ksize * op.Cin + P.Cs;
    P.B = reinterpret_cast<const float*>(op.w); P.ldb = P.K; P.b_bs = 0;
    P.M = P.Hout * P.Wout; P.N = op.Cout; P.b_rows = op.Cout; P.batch = op.N;
    P.bias = op.bias;
    P.residual = reinterpret_cast<const float*>(op.residual);
    P.out = static_cast<float*>(op.out); P.ld_out = op.Cout;
    P.alpha = op.alpha;
    return launch_f32(P, s, prof);
}

int launch_gemm_fp32(const GemmOp& op, cudaStream_t s, Profiler* prof) {
    F32Problem P{};
    P.in = nullptr;
    P.A = reinterpret_cast<const float*>(op.A);
    P.lda = op.lda ? op.lda : op.K;
    P.a_bs = op.a_batched ? (op.a_bstride ? op.a_bstride : P.lda * op.M) : 0;
    P.B = reinterpret_cast<const float*>(op.B);
    P.ldb = op.ldb ? op.ldb : op.K;
    P.b_rows = op.b_rows > 0 ? op.b_rows : op.N;
    P.b_bs = op.b_batched ? (op.b_bstride ? op.b_bstride : P.ldb * P.b_rows) : 0;
    P.M = op.M; P.N = op.N; P.K = op.K; P.batch = op.batch;
    P.bias = op.bias;
    P.residual = reinterpret_cast<const float*>(op.residual);
    P.out = static_cast<float*>(op.out);
    P.ld_out = op.ld_out ? op.ld_out : op.N;
    P.alpha = op.alpha;
    VT_CHECK(P.lda % 4 == 0 && P.ldb % 4 == 0, "fp32 GEMM needs row strides divisible by 4");
    return launch_f32(P, s, prof);
}

}  // namespace vt
