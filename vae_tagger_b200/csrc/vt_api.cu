// C-ABI of the encode+tag path (include/vae_tagger_b200.h): context, parameter store and
// repacking, the encoder schedule (which kernel runs on which buffer, in what order), the
// tag-head schedule, the host end-to-end call and the single-op test entry points.
#include <map>
#include <memory>
#include <vector>

#include "../../include/vae_tagger_b200.h"
#include "vt_head_train.h"
#include "vt_internal.h"
#include "vt_backward.h"
#include "vt_losses.h"
#include "vt_resize.h"
#include "vt_ptx.cuh"

using namespace vt;

namespace {

struct Param {
    std::vector<int64_t> shape;
    float* dev = nullptr;
    size_t numel = 0;
};

struct ConvW {      // packed conv / linear operand: [Cout][Ktot]
    bf16* w16 = nullptr;   // 16-bit copy (fp16 or bf16 bits, see f16)
    float* w32 = nullptr;
    float* bias = nullptr;  // [Cout] (conv2 + shortcut biases folded)
    int Cin = 0, Cout = 0, ksize = 1, Cs = 0, Ktot = 0;
    bool f16 = true;  // 16-bit copy: main columns fp16 (else bf16); shortcut columns are always bf16
};
struct NormW {
    const float* gamma = nullptr;
    const float* beta = nullptr;
};
struct ResnetW {
    NormW norm1, norm2;
    ConvW conv1, conv2;  // conv2 carries the 1x1 shortcut as an extra K slab when cin != cout
    int cin = 0, cout = 0;
};
struct AttnW {
    NormW gn;
    ConvW qk;   // [2C][C]  rows 0..C-1 = to_q, C..2C-1 = to_k
    ConvW v;    // [C][C]   used as the A operand of the V^T GEMM; bias applied after P.V
    ConvW out;  // [C][C]
    int C = 0;
};

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return 0;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        VT_CUDA(cudaMalloc(&p, bytes));
        cap = bytes;
        return 0;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

// carve buffers out of the single-op workspace: register (pointer, bytes) pairs, then bind them all at once
struct Carver {
    struct Item { void** p; size_t bytes; };
    std::vector<Item> items;
    template <typename T> void want(T** p, size_t bytes) { items.push_back({reinterpret_cast<void**>(p), (bytes + 1023) / 1024 * 1024}); }
    int bind(DevBuf& ws) {
        size_t total = 0;
        for (auto& it : items) total += it.bytes;
        VT_TRY(ws.ensure(total));
        char* b = static_cast<char*>(ws.p);
        for (auto& it : items) { *it.p = b; b += it.bytes; }
        return 0;
    }
};
}  // namespace

struct EncTape;   // activations kept by the encoder's training forward (vt_train_encoder.cuh)

struct vt_ctx {
    int device = 0;
    Profiler* prof = nullptr;
    // 16-bit mode: storage format of RAW activations (residual stream, conv outputs) and of the weight columns
    // that multiply them.  fp16 (default): the format the reference's own fp16 autocast stores them in
    // (infer_full.py:100) -- 11-bit mantissa, latent error ~5e-3; bf16 (VT_B200_RAW_BF16=1 when the context is
    // created): 8-bit mantissa, unbounded range, latent error 0.8-1.2e-2.  Fixed per context: the packed weights
    // depend on it.
    bool raw_f16 = true;

    // ---- encoder
    vt_encoder_config ecfg{};
    bool ecfg_set = false, enc_ready = false;
    std::map<std::string, Param> eparams;
    std::vector<void*> epacked;  // allocations owned by the packed representation
    ConvW conv_in, conv_out;
    std::vector<std::vector<ResnetW>> down;  // [block][layer]
    std::vector<ConvW> downsample;           // per block (Cout == 0: none)
    ResnetW mid0, mid1;
    AttnW attn;
    NormW norm_out;
    // Two execution lanes: consecutive micro-batches of one vt_encode call alternate between two
    // internal streams with private workspaces, so the HBM-bound passes (GroupNorm apply, softmax) of
    // one micro-batch run under the tensor-bound contractions of the other and kernel tails overlap.
    struct Lane {
        cudaStream_t stream = nullptr;
        cudaEvent_t done = nullptr;
        DevBuf arena;  // activation workspace
        DevBuf stats;  // GroupNorm (sum, sumsq) slots
        DevBuf statpart;  // per-tile partials of the layer in flight (two-stage statistics, vt_internal.h)
        DevBuf mom;    // conv_out moments fp32 NHWC
        DevBuf hws;    // tag-head workspace when the head runs per micro-batch on this lane (vt_infer_host)
    } lanes[2];
    cudaEvent_t ev_start = nullptr;

    // ---- encoder training: the tape of the last training forward, gradient buffers by parameter name
    EncTape* tapes[VT_MAX_TAPES] = {};
    EncTape* dtapes[VT_MAX_TAPES] = {};    // decoder training forwards
    std::map<std::string, float*> egrads, dgrads;
    DevBuf tbws, tlws, tg0;   // backward scratch: one helper call / one layer / the rotating gradient buffers

    // ---- VAE decoder; shares ecfg with the encoder
    bool dec_ready = false;
    std::map<std::string, Param> dparams;
    std::vector<void*> dpacked;
    ConvW dconv_in, dconv_out;
    std::vector<std::vector<ResnetW>> up;  // [block][layer], layers_per_block + 1 each
    std::vector<ConvW> upsample;           // per block (Cout == 0: none)
    std::vector<bf16*> upsample_sp;        // per block: sub-pixel weights [4][C][4C] (nullptr: none)
    ResnetW dmid0, dmid1;
    AttnW dattn;
    NormW dnorm_out;

    // ---- head
    vt_head_config hcfg{};
    bool hcfg_set = false, head_ready = false;
    std::map<std::string, Param> hparams;
    DevBuf hws;  // head workspace
    DevBuf e2e;  // device staging of vt_infer_host

    DevBuf opws;  // single-op entry points
    DevBuf optws; // optimizer scratch
    ResizeCache* resize = nullptr;  // preprocessing coefficient tables
};

namespace {

int set_device(vt_ctx* c) {
    VT_CHECK(c != nullptr, "null context");
    VT_CUDA(cudaSetDevice(c->device));
    return 0;
}

int store_param(std::map<std::string, Param>& m, const char* name, const float* data, const int64_t* shape,
                int ndim) {
    VT_CHECK(name && data && (shape || ndim == 0) && ndim >= 0 && ndim <= 8, "bad parameter arguments");
    Param& p = m[name];
    size_t n = 1;
    std::vector<int64_t> sh(shape, shape + ndim);
    for (auto d : sh) {
        VT_CHECK(d > 0, "parameter dimensions must be positive");
        n *= static_cast<size_t>(d);
    }
    if (p.dev == nullptr || p.numel != n) {
        if (p.dev) cudaFree(p.dev);
        p.dev = nullptr;
        VT_CUDA(cudaMalloc(&p.dev, n * sizeof(float)));
    }
    p.shape = sh;
    p.numel = n;
    VT_CUDA(cudaMemcpy(p.dev, data, n * sizeof(float), cudaMemcpyDefault));
    return 0;
}

void free_params(std::map<std::string, Param>& m) {
    for (auto& kv : m)
        if (kv.second.dev) cudaFree(kv.second.dev);
    m.clear();
}

// ---- weight packing kernels: OIHW fp32 -> [Cout][Ktot] with k = (kh*ks+kw)*Cin + ci, at column
// offset k_off (the shortcut slab lands behind the taps).
template <int OFMT>  // FMT_BF16 / FMT_F32 / FMT_F16
__global__ void pack_weight_kernel(const float* __restrict__ src, void* __restrict__ dstv, int Cout, int Cin, int ks,
                                   int Ktot, int k_off, int cin_pad = 0) {
    const int CinP = cin_pad > 0 ? cin_pad : Cin;  // K index stride per tap (input channels zero-padded)
    const long long total = 1LL * Cout * Cin * ks * ks;
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += 256LL * gridDim.x) {
        // i indexes src = ((co*Cin + ci)*ks + kh)*ks + kw
        long long r = i;
        const int kw = static_cast<int>(r % ks); r /= ks;
        const int kh = static_cast<int>(r % ks); r /= ks;
        const int ci = static_cast<int>(r % Cin);
        const int co = static_cast<int>(r / Cin);
        const long long d = 1LL * co * Ktot + k_off + (kh * ks + kw) * CinP + ci;
        if constexpr (OFMT == FMT_BF16) static_cast<bf16*>(dstv)[d] = __float2bfloat16(src[i]);
        else if constexpr (OFMT == FMT_F16)
            static_cast<__half*>(dstv)[d] = __float2half_rn(fminf(fmaxf(src[i], -65504.f), 65504.f));
        else static_cast<float*>(dstv)[d] = src[i];
    }
}
// Sub-pixel weights of "nearest 2x upsample + conv3x3": for output parity (py,px) the 3x3 taps collapse onto
// 2x2 source pixels.  Rows: py = 0 -> source rows {y-1: ky 0; y: ky 1+2}, py = 1 -> {y: ky 0+1; y+1: ky 2};
// columns alike.  dst[par][co][(ty*2+tx)*C + ci], summed in fp32 then rounded to bf16 (raw operand).
template <int OFMT>   // raw 16-bit format of the operand these weights multiply
__global__ void pack_subpixel_kernel(const float* __restrict__ src /*[Co][C][3][3]*/, bf16* __restrict__ dst, int Co,
                                     int C) {
    const long long total = 4LL * Co * 4 * C;
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += 256LL * gridDim.x) {
        long long r = i;
        const int ci = static_cast<int>(r % C); r /= C;
        const int tap = static_cast<int>(r % 4); r /= 4;
        const int co = static_cast<int>(r % Co);
        const int par = static_cast<int>(r / Co);
        const int py = par >> 1, px = par & 1, ty = tap >> 1, tx = tap & 1;
        // kernel rows / columns that land on this source tap
        const int ky0 = py == 0 ? (ty == 0 ? 0 : 1) : (ty == 0 ? 0 : 2), ky1 = py == 0 ? (ty == 0 ? 0 : 2) : (ty == 0 ? 1 : 2);
        const int kx0 = px == 0 ? (tx == 0 ? 0 : 1) : (tx == 0 ? 0 : 2), kx1 = px == 0 ? (tx == 0 ? 0 : 2) : (tx == 0 ? 1 : 2);
        float a = 0.f;
        for (int ky = ky0; ky <= ky1; ++ky)
            for (int kx = kx0; kx <= kx1; ++kx) a += src[((1LL * co * C + ci) * 3 + ky) * 3 + kx];
        if constexpr (OFMT == FMT_F16) reinterpret_cast<__half*>(dst)[i] = __float2half_rn(fminf(fmaxf(a, -65504.f), 65504.f));
        else dst[i] = __float2bfloat16(a);
    }
}
__global__ void add_vec_kernel(float* __restrict__ a, const float* __restrict__ b, int n) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i < n) a[i] += b[i];
}

struct Packer {
    vt_ctx* c;
    std::map<std::string, Param>* params;  // eparams (encoder) or dparams (decoder)
    std::vector<void*>* packed;            // allocations owned by the packed representation
    const char* what;
    int get(const std::string& name, std::initializer_list<int64_t> shape, const Param** out) {
        auto it = params->find(name);
        if (it == params->end()) {
            set_error(std::string(what) + " parameter missing: " + name);
            return -4;
        }
        const Param& p = it->second;
        std::vector<int64_t> want(shape);
        if (p.shape != want) {
            std::string m = std::string(what) + " parameter " + name + " has shape [";
            for (auto d : p.shape) m += std::to_string(d) + ",";
            m += "] expected [";
            for (auto d : want) m += std::to_string(d) + ",";
            set_error(m + "]");
            return -4;
        }
        *out = &p;
        return 0;
    }
    template <typename T>
    int alloc(T** p, size_t n) {
        void* q = nullptr;
        VT_CUDA(cudaMalloc(&q, n * sizeof(T)));
        VT_CUDA(cudaMemset(q, 0, n * sizeof(T)));
        packed->push_back(q);
        *p = static_cast<T*>(q);
        return 0;
    }
    // conv weight `prefix`.weight [Cout][Cin][ks][ks] (+ optional shortcut [Cout][Cs][1][1]); Kpad pads K.
    // f16: the main operand of this conv is a normalised (bounded) tensor -> fp16 weight columns; otherwise it is a
    // raw activation and the columns take the context's raw format, like the shortcut columns always do.
    int conv(const std::string& prefix, int Cin, int Cout, int ks, const std::string& sc_prefix, int Cs, int Kpad,
             ConvW* w, bool f16 = true) {
        f16 = f16 || c->raw_f16;
        const Param *pw, *pb;
        VT_TRY(get(prefix + ".weight", {Cout, Cin, ks, ks}, &pw));
        VT_TRY(get(prefix + ".bias", {Cout}, &pb));
        const int Ktot = std::max(Kpad, ks * ks * Cin + Cs);
        w->Cin = Cin; w->Cout = Cout; w->ksize = ks; w->Cs = Cs; w->Ktot = Ktot;
        VT_TRY(alloc(&w->w16, static_cast<size_t>(Cout) * Ktot));
        VT_TRY(alloc(&w->w32, static_cast<size_t>(Cout) * Ktot));
        VT_TRY(alloc(&w->bias, static_cast<size_t>(Cout)));
        const int grid = 1024;
        w->f16 = f16;
        if (f16) pack_weight_kernel<FMT_F16><<<grid, 256>>>(pw->dev, w->w16, Cout, Cin, ks, Ktot, 0);
        else pack_weight_kernel<FMT_BF16><<<grid, 256>>>(pw->dev, w->w16, Cout, Cin, ks, Ktot, 0);
        pack_weight_kernel<FMT_F32><<<grid, 256>>>(pw->dev, w->w32, Cout, Cin, ks, Ktot, 0);
        VT_CUDA(cudaMemcpy(w->bias, pb->dev, Cout * sizeof(float), cudaMemcpyDeviceToDevice));
        if (Cs > 0) {
            const Param *sw, *sb;
            VT_TRY(get(sc_prefix + ".weight", {Cout, Cs, 1, 1}, &sw));
            VT_TRY(get(sc_prefix + ".bias", {Cout}, &sb));
            if (c->raw_f16) pack_weight_kernel<FMT_F16><<<grid, 256>>>(sw->dev, w->w16, Cout, Cs, 1, Ktot, ks * ks * Cin);
            else pack_weight_kernel<FMT_BF16><<<grid, 256>>>(sw->dev, w->w16, Cout, Cs, 1, Ktot, ks * ks * Cin);
            pack_weight_kernel<FMT_F32><<<grid, 256>>>(sw->dev, w->w32, Cout, Cs, 1, Ktot, ks * ks * Cin);
            add_vec_kernel<<<(Cout + 255) / 256, 256>>>(w->bias, sb->dev, Cout);
        }
        VT_CUDA(cudaGetLastError());
        return 0;
    }
    // conv with the input channels zero-padded to CinP (one 64-wide K chunk) and the output channels to CoutP
    // (a 32-column accumulator tile): the decoder's conv_in (16 -> 512) and conv_out (128 -> 3)
    int conv_padded(const std::string& prefix, int Cin, int Cout, int ks, int CinP, int CoutP, ConvW* w, bool f16) {
        const Param *pw, *pb;
        VT_TRY(get(prefix + ".weight", {Cout, Cin, ks, ks}, &pw));
        VT_TRY(get(prefix + ".bias", {Cout}, &pb));
        const int Ktot = ks * ks * CinP;
        f16 = f16 || c->raw_f16;
        w->Cin = CinP; w->Cout = CoutP; w->ksize = ks; w->Cs = 0; w->Ktot = Ktot; w->f16 = f16;
        VT_TRY(alloc(&w->w16, static_cast<size_t>(CoutP) * Ktot));
        VT_TRY(alloc(&w->w32, static_cast<size_t>(CoutP) * Ktot));
        VT_TRY(alloc(&w->bias, static_cast<size_t>(CoutP)));
        if (f16) pack_weight_kernel<FMT_F16><<<256, 256>>>(pw->dev, w->w16, Cout, Cin, ks, Ktot, 0, CinP);
        else pack_weight_kernel<FMT_BF16><<<256, 256>>>(pw->dev, w->w16, Cout, Cin, ks, Ktot, 0, CinP);
        pack_weight_kernel<FMT_F32><<<256, 256>>>(pw->dev, w->w32, Cout, Cin, ks, Ktot, 0, CinP);
        VT_CUDA(cudaMemcpy(w->bias, pb->dev, Cout * sizeof(float), cudaMemcpyDeviceToDevice));
        VT_CUDA(cudaGetLastError());
        return 0;
    }
    // the four sub-pixel weight sets of an upsample conv (the operand is a raw activation: the context's raw format)
    int conv_subpixel(const std::string& prefix, int C, bf16** w4) {
        const Param* pw;
        VT_TRY(get(prefix + ".weight", {C, C, 3, 3}, &pw));
        VT_TRY(alloc(w4, static_cast<size_t>(4) * C * 4 * C));
        if (c->raw_f16) pack_subpixel_kernel<FMT_F16><<<1024, 256>>>(pw->dev, *w4, C, C);
        else pack_subpixel_kernel<FMT_BF16><<<1024, 256>>>(pw->dev, *w4, C, C);
        VT_CUDA(cudaGetLastError());
        return 0;
    }
    // linear weights [out][in] stacked along the output dimension
    int linear(std::initializer_list<std::string> prefixes, int In, int OutEach, bool with_bias, ConvW* w) {
        const int n = static_cast<int>(prefixes.size());
        w->Cin = In; w->Cout = n * OutEach; w->ksize = 1; w->Cs = 0; w->Ktot = In; w->f16 = true;
        VT_TRY(alloc(&w->w16, static_cast<size_t>(w->Cout) * In));
        VT_TRY(alloc(&w->w32, static_cast<size_t>(w->Cout) * In));
        VT_TRY(alloc(&w->bias, static_cast<size_t>(w->Cout)));
        int i = 0;
        for (const auto& pre : prefixes) {
            const Param *pw, *pb;
            VT_TRY(get(pre + ".weight", {OutEach, In}, &pw));
            VT_TRY(get(pre + ".bias", {OutEach}, &pb));
            const size_t off = static_cast<size_t>(i) * OutEach * In;
            VT_TRY(launch_cast_f32_16(pw->dev, w->w16 + off, FMT_F16, 1LL * OutEach * In, nullptr));
            VT_CUDA(cudaMemcpy(w->w32 + off, pw->dev, sizeof(float) * OutEach * In, cudaMemcpyDeviceToDevice));
            if (with_bias)
                VT_CUDA(cudaMemcpy(w->bias + static_cast<size_t>(i) * OutEach, pb->dev, sizeof(float) * OutEach,
                                   cudaMemcpyDeviceToDevice));
            ++i;
        }
        return 0;
    }
    int norm(const std::string& prefix, int C, NormW* n) {
        const Param *g, *b;
        VT_TRY(get(prefix + ".weight", {C}, &g));
        VT_TRY(get(prefix + ".bias", {C}, &b));
        n->gamma = g->dev;
        n->beta = b->dev;
        return 0;
    }
    int resnet(const std::string& prefix, int cin, int cout, ResnetW* r) {
        r->cin = cin; r->cout = cout;
        VT_TRY(norm(prefix + ".norm1", cin, &r->norm1));
        VT_TRY(conv(prefix + ".conv1", cin, cout, 3, "", 0, 0, &r->conv1));
        VT_TRY(norm(prefix + ".norm2", cout, &r->norm2));
        VT_TRY(conv(prefix + ".conv2", cout, cout, 3, prefix + ".conv_shortcut", cin != cout ? cin : 0, 0, &r->conv2));
        return 0;
    }
};

void free_packed_decoder(vt_ctx* c) {
    for (void* p : c->dpacked) cudaFree(p);
    c->dpacked.clear();
    c->up.clear();
    c->upsample.clear();
    c->upsample_sp.clear();
    c->dec_ready = false;
}

void free_packed(vt_ctx* c) {
    for (void* p : c->epacked) cudaFree(p);
    c->epacked.clear();
    c->down.clear();
    c->downsample.clear();
    c->enc_ready = false;
}

// ---------------------------------------------------------------------------------------------
// The encoder schedule.  Activations are NHWC in four ping-pong buffers of the workspace arena;
// GroupNorm statistics come from the producing contraction's epilogue (bf16 mode) or a separate
// reduction (fp32 mode).
//
// Formats of bf16 mode ("16-bit tensor-core mode"; accumulation is always fp32 in TMEM):
//   * raw activations (residual stream, conv1 outputs) are stored bf16: unbounded range, 8-bit mantissa;
//   * every *bounded* MMA operand -- GroupNorm+SiLU outputs, the normalised image patches, q/k/v, softmax
//     probabilities, the attention output, and all weights that multiply them -- is fp16 (11-bit
//     mantissa): the operand rounding that dominates the pipeline's error shrinks 8x.  Measured with
//     tests/emulate_bf16.py against the fp32 oracle: all-bf16 operands 1.06e-2 latent rel-L2 (over the
//     1e-2 bar), fp16 bounded operands + bf16 storage 0.80e-2.  The reference itself infers under
//     fp16 autocast (infer_full.py:100);
//   * operands that ARE raw activations (downsample conv input, 1x1 shortcut input) stay bf16, with
//     bf16 weight columns.
struct Act {
    void* p = nullptr;
    int fmt = 0;  // FMT_BF16 / FMT_F32 / FMT_F16
};

struct EncRun {
    vt_ctx* c;
    cudaStream_t s;
    int fp32;     // verification mode: everything fp32
    int n;        // images in this micro-batch
    double* stats_base;
    int stats_used = 0;
    int groups;
    StatsScratch ws{};  // this lane's partial buffer for the epilogue statistics
    int use_fused = 1;  // GroupNorm+SiLU fused into the 3x3 convs (VT_B200_NO_FUSED_GN=1 disables)
    int use_flash = 1;  // fused attention kernel (VT_B200_NO_FLASH=1: score-matrix path)

    double* new_stats() {
        double* p = stats_base + static_cast<size_t>(stats_used) * n * groups * 2;
        ++stats_used;
        return p;
    }
    const void* W(const ConvW& w) const { return fp32 ? static_cast<const void*>(w.w32) : static_cast<const void*>(w.w16); }
    int raw_fmt() const { return fp32 ? FMT_F32 : (c->raw_f16 ? FMT_F16 : FMT_BF16); }   // residual stream / conv1 outputs
    int raw16() const { return !fp32 && c->raw_f16; }
    int opd_fmt() const { return fp32 ? FMT_F32 : FMT_F16; }    // bounded MMA operands

    int conv(const void* in, int H, int Wd, const ConvW& w, int stride, const void* sc_in, const Act* residual,
             Act out, double* st) {
        ConvOp op;
        op.in = in; op.N = n; op.Hin = H; op.Win = Wd; op.Cin = w.Cin; op.ksize = w.ksize; op.stride = stride;
        op.in_f16 = w.f16; op.raw_f16 = raw16();
        op.w = W(w); op.Cout = w.Cout; op.sc_in = sc_in; op.Cs = w.Cs; op.bias = w.bias;
        if (residual) { op.residual = residual->p; op.residual_fp32 = residual->fmt == FMT_F32; }
        op.out = out.p; op.out_fmt = out.fmt;
        if (fp32) {
            VT_TRY(launch_conv_fp32(op, s, c->prof));
            if (st) {
                const int Ho = stride == 1 ? H : H / 2, Wo = stride == 1 ? Wd : Wd / 2;
                VT_TRY(launch_gn_stats(out.p, 1, st, n, 1LL * Ho * Wo, w.Cout, groups, s, c->prof));
            }
            return 0;
        }
        op.stats = st; op.stats_ws = ws;
        return launch_conv(op, s, c->prof);
    }
    int gemm(GemmOp& op, long long rows_for_stats) {
        if (fp32) {
            double* st = op.stats;
            op.stats = nullptr;
            op.out_fmt = FMT_F32;
            VT_TRY(launch_gemm_fp32(op, s, c->prof));
            if (st) VT_TRY(launch_gn_stats(op.out, 1, st, op.batch, rows_for_stats, op.N, groups, s, c->prof));
            return 0;
        }
        op.stats_ws = ws; op.raw_f16 = raw16();
        return launch_gemm(op, s, c->prof);
    }
    // normalised operand: fp16 in 16-bit mode (it feeds TMA), fp32 in verification mode
    int gn(Act x, void* y, const double* st, const NormW& nw, long long HW, int C, int silu) {
        return launch_gn_apply(x.p, x.fmt, y, opd_fmt(), st, nw.gamma, nw.beta, n, HW, C, groups, 1e-6f, silu,
                               s, c->prof);
    }
    int conv_fused(Act x, const double* st_x, const NormW& nw, int H, int Wd, const ConvW& w, const Act* residual,
                   Act out, double* st, const void* sc_in = nullptr) {
        Conv3FusedOp op;
        op.sc_in = sc_in; op.Cs = sc_in ? w.Cs : 0;
        op.in = x.p; op.N = n; op.H = H; op.W = Wd; op.Cin = w.Cin; op.Cout = w.Cout; op.gn_stats = st_x;
        op.gamma = nw.gamma; op.beta = nw.beta; op.w = w.w16; op.bias = w.bias;
        if (residual) { op.residual = residual->p; op.residual_fp32 = residual->fmt == FMT_F32; }
        op.out = out.p; op.out_fmt = out.fmt; op.stats = st; op.stats_ws = ws; op.raw_f16 = raw16();
        return launch_conv3_fused(op, s, c->prof);
    }
    // ResnetBlock2D: out = x (+shortcut) + conv2(silu(norm2(conv1(silu(norm1(x))))))
    // 16-bit mode: both GroupNorm+SiLU steps are fused into the operand path of the following 3x3 conv
    // (vt_conv3.cuh), including the 1x1 shortcut slab of the channel-changing blocks.
    int resnet(const ResnetW& r, Act x, const double* st_x, int H, int Wd, int level, void* T, void* Hb, Act out,
               double* st_out) {
        (void)level;
        const long long HW = 1LL * H * Wd;
        double* st_h = new_stats();
        Act h{Hb, raw_fmt()};
        if (!fp32 && use_fused) {
            VT_TRY(conv_fused(x, st_x, r.norm1, H, Wd, r.conv1, nullptr, h, st_h));
            if (r.cin == r.cout) return conv_fused(h, st_h, r.norm2, H, Wd, r.conv2, &x, out, st_out);
            if (r.cout >= 256) {
                // channel-changing block: the 1x1 shortcut of the block input is an extra K slab of conv2
                VT_CHECK(x.fmt == raw_fmt(), "shortcut operand must be in the raw 16-bit format");
                return conv_fused(h, st_h, r.norm2, H, Wd, r.conv2, nullptr, out, st_out, x.p);
            }
            VT_TRY(gn(h, T, st_h, r.norm2, HW, r.cout, 1));
            return conv(T, H, Wd, r.conv2, 1, x.p, nullptr, out, st_out);
        }
        VT_TRY(gn(x, T, st_x, r.norm1, HW, r.cin, 1));
        VT_TRY(conv(T, H, Wd, r.conv1, 1, nullptr, nullptr, h, st_h));
        VT_TRY(gn(h, T, st_h, r.norm2, HW, r.cout, 1));
        if (r.cin != r.cout) {
            VT_CHECK(fp32 || x.fmt == raw_fmt(), "shortcut operand must be in the raw 16-bit format");
            return conv(T, H, Wd, r.conv2, 1, x.p, nullptr, out, st_out);
        }
        return conv(T, H, Wd, r.conv2, 1, nullptr, &x, out, st_out);
    }
};

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct AttnPlan {  // workspace of the mid-block attention (after the four activation buffers)
    size_t attn_bytes = 0;
    long long rows_per_chunk = 0;
    int ipc = 1;
    size_t qk_b = 0, vt_b = 0, s_b = 0, p_b = 0, o_b = 0, part_b = 0;
    int pv_splits = 1;
};

AttnPlan plan_attention(bool enabled, int n, long long tokens, int Cm, size_t es, int fp32) {
    AttnPlan pl;
    if (!enabled) return pl;
    const size_t s_budget = 64ull << 20;  // score tile kept L2 resident
    const long long max_rows = static_cast<long long>(s_budget / (tokens * 4));
    if (max_rows >= tokens) {
        pl.rows_per_chunk = tokens;
        pl.ipc = static_cast<int>(std::min<long long>(n, std::max<long long>(1, max_rows / tokens)));
    } else {
        pl.rows_per_chunk = std::max<long long>(128, max_rows / 128 * 128);
        pl.ipc = 1;
    }
    // rows of V^T, of the score matrix and of the probabilities are pitched to a multiple of 64 tokens (TMA row
    // alignment, K chunks of the P.V contraction); the padding holds zeros
    const long long tp = (tokens + 63) / 64 * 64;
    pl.qk_b = align_up(static_cast<size_t>(n) * tokens * 2 * Cm * es, 1024);
    pl.vt_b = align_up(static_cast<size_t>(n) * tp * Cm * es, 1024);
    pl.s_b = align_up(static_cast<size_t>(pl.ipc) * pl.rows_per_chunk * tp * 4, 1024);
    pl.p_b = align_up(static_cast<size_t>(pl.ipc) * pl.rows_per_chunk * tp * es, 1024);
    pl.o_b = pl.vt_b;
    // P.V of a row chunk has only ceil(rows/128) * C/256 output tiles: split K (= tokens) across
    // CTAs when that leaves most of the machine idle, combine the fp32 partials afterwards
    if (!fp32 && pl.ipc == 1) {
        const long long tiles = (pl.rows_per_chunk + 127) / 128 * ((Cm + 255) / 256);
        const long long kchunks = tp / 64;
        int want = static_cast<int>(std::min<long long>(16, 148 / std::max<long long>(1, tiles)));
        while (want > 1 && kchunks % want != 0) --want;
        pl.pv_splits = std::max(1, want);
        if (pl.pv_splits > 1) pl.part_b = align_up(static_cast<size_t>(pl.pv_splits) * pl.rows_per_chunk * Cm * 4, 1024);
    }
    pl.attn_bytes = pl.qk_b + pl.vt_b + pl.s_b + pl.p_b + pl.o_b + pl.part_b;
    return pl;
}

struct MidW {
    const ResnetW* mid0;
    const AttnW* attn;
    const ResnetW* mid1;
    bool has_attn;
};

// Single-head attention of the mid block over all tokens (diffusers Attention + AttnProcessor2_0):
// out = to_out(softmax(q k^T / sqrt(C)) v) + x.  T receives the normalised tokens, ab the [q|k], V^T, (S, P,) O buffers.
int run_attention(EncRun& R, const AttnW& A, const AttnPlan& pl, char* ab, Act X, const double* st_x, void* T, Act out,
                  double* st_o, int n, int h, int w_, int fp32, size_t es) {
    vt_ctx* c = R.c;
    cudaStream_t s = R.s;
    const long long tokens = 1LL * h * w_;
    const long long tp = (tokens + 63) / 64 * 64;   // row pitch of V^T / scores / probabilities (plan_attention)
    const long long rows_per_chunk = pl.rows_per_chunk;
    const int ipc = pl.ipc, pv_splits = pl.pv_splits;
    const size_t qk_b = pl.qk_b, vt_b = pl.vt_b, s_b = pl.s_b, p_b = pl.p_b, o_b = pl.o_b;
    {
        const int C = A.C;
        char* QK = ab;
        char* Vt = QK + qk_b;
        char* S = Vt + vt_b;
        char* P = S + s_b;
        char* O = P + p_b;
        char* PART = O + o_b;
        VT_TRY(R.gn(X, T, st_x, A.gn, tokens, C, 0));
        {   // [q | k] = t Wqk^T + b : [n][tokens][2C]
            GemmOp g;
            g.A = T; g.B = R.W(A.qk); g.batch = n; g.M = static_cast<int>(tokens); g.N = 2 * C; g.K = C;
            g.a_batched = 1; g.b_batched = 0; g.bias = A.qk.bias; g.out = QK; g.ab_f16 = 1; g.out_fmt = R.opd_fmt();
            VT_TRY(R.gemm(g, tokens));
        }
        {   // V^T = Wv t^T : [n][C][tokens]  (bias b_v is added after P.V: softmax rows sum to one)
            GemmOp g;
            g.A = R.W(A.v); g.B = T; g.batch = n; g.M = C; g.N = static_cast<int>(tp); g.b_rows = static_cast<int>(tokens);
            g.K = C;
            g.a_batched = 0; g.b_batched = 1; g.out = Vt; g.ab_f16 = 1; g.out_fmt = R.opd_fmt();
            VT_TRY(R.gemm(g, 0));
        }
        const float scale = 1.0f / sqrtf(static_cast<float>(C));
        if (!fp32 && C == 512 && R.use_flash) {
            // fused attention (vt_flash.cu): scores and probabilities never leave the SM
            FlashOp f;
            f.qk = QK; f.vt = Vt; f.bias_v = A.v.bias; f.out = O; f.n = n; f.tokens = static_cast<int>(tokens);
            f.ld_vt = tp; f.C = C; f.scale = scale;
            VT_TRY(launch_flash_attention(f, s, c->prof));
        } else
        for (int i0 = 0; i0 < n; i0 += ipc) {
            const int nb_img = std::min(ipc, n - i0);
            for (long long r0 = 0; r0 < tokens; r0 += rows_per_chunk) {
                const int rows = static_cast<int>(std::min<long long>(rows_per_chunk, tokens - r0));
                {   // S = scale * Q K^T (fp32)
                    GemmOp g;
                    g.A = QK + (static_cast<size_t>(i0) * tokens + r0) * 2 * C * es;
                    g.lda = 2 * C; g.a_bstride = tokens * 2 * C;
                    g.B = QK + (static_cast<size_t>(i0) * tokens * 2 * C + C) * es;
                    g.ldb = 2 * C; g.b_bstride = tokens * 2 * C;
                    g.batch = nb_img; g.M = rows; g.N = static_cast<int>(tp); g.b_rows = static_cast<int>(tokens); g.K = C;
                    g.alpha = scale; g.out = S; g.out_fmt = FMT_F32; g.ab_f16 = 1;
                    VT_TRY(R.gemm(g, 0));
                }
                VT_TRY(launch_softmax_rows(reinterpret_cast<const float*>(S), P, R.opd_fmt(), 1LL * nb_img * rows,
                                           static_cast<int>(tokens), tp, tp, s, c->prof));
                if (pv_splits > 1) {
                    // O = P V + b_v with K split: batch index = K slice, fp32 partials, then combine
                    const long long ks = tp / pv_splits;
                    GemmOp g;
                    g.A = P; g.lda = tp; g.a_bstride = ks;
                    g.B = Vt + static_cast<size_t>(i0) * C * tp * es; g.ldb = tp; g.b_bstride = ks;
                    g.batch = pv_splits; g.M = rows; g.N = C; g.K = static_cast<int>(ks);
                    g.out = PART; g.out_fmt = FMT_F32; g.ab_f16 = 1; g.ld_out = C; g.out_bstride = 1LL * rows * C;
                    VT_TRY(R.gemm(g, 0));
                    VT_TRY(launch_splitk_reduce(reinterpret_cast<const float*>(PART), pv_splits, 1LL * rows * C, A.v.bias,
                                                O + (static_cast<size_t>(i0) * tokens + r0) * C * es, FMT_F16, rows, C, C,
                                                s, c->prof));
                } else {   // O = P V + b_v
                    GemmOp g;
                    g.A = P; g.lda = tp; g.a_bstride = 1LL * rows * tp;
                    g.B = Vt + static_cast<size_t>(i0) * C * tp * es; g.ldb = tp; g.b_bstride = 1LL * C * tp;
                    g.batch = nb_img; g.M = rows; g.N = C; g.K = static_cast<int>(tp);
                    g.bias = A.v.bias; g.ab_f16 = 1; g.out_fmt = R.opd_fmt();
                    g.out = O + (static_cast<size_t>(i0) * tokens + r0) * C * es;
                    g.ld_out = C; g.out_bstride = tokens * C;
                    VT_TRY(R.gemm(g, 0));
                }
            }
        }
        {   // out = O Wo^T + b_o + x
            GemmOp g;
            g.A = O; g.B = R.W(A.out); g.batch = n; g.M = static_cast<int>(tokens); g.N = C; g.K = C;
            g.a_batched = 1; g.b_batched = 0; g.bias = A.out.bias; g.residual = X.p; g.residual_fp32 = X.fmt == FMT_F32;
            g.out = out.p; g.out_fmt = out.fmt; g.ab_f16 = 1; g.stats = st_o;
            VT_TRY(R.gemm(g, tokens));
        }
    }
    return 0;
}

// UNetMidBlock2D (shared by the encoder and the decoder): resnet, single-head attention over all
// tokens, resnet.  X / spare / st_x are the caller's ping-pong state.
int run_mid_block(EncRun& R, const MidW& M, const AttnPlan& pl, char* ab, Act& X, void*& spare, double*& st_x, void* T,
                  void* Hb, int n, int h, int w_, int fp32, size_t es) {
    vt_ctx* c = R.c;
    cudaStream_t s = R.s;
    const long long tokens = 1LL * h * w_;
    const long long tp = (tokens + 63) / 64 * 64;   // row pitch of V^T / scores / probabilities (plan_attention)
    const long long rows_per_chunk = pl.rows_per_chunk;
    const int ipc = pl.ipc, pv_splits = pl.pv_splits;
    const size_t qk_b = pl.qk_b, vt_b = pl.vt_b, s_b = pl.s_b, p_b = pl.p_b, o_b = pl.o_b;
    auto advance = [&](Act out) { spare = X.p; X = out; };
    const int lvl = 0;
    // ---- mid block
    {
        double* st_o = R.new_stats();
        Act out{spare, R.raw_fmt()};
        VT_TRY(R.resnet((*M.mid0), X, st_x, h, w_, lvl, T, Hb, out, st_o));
        advance(out);
        st_x = st_o;
    }
    if (M.has_attn) {
        double* st_o = R.new_stats();
        Act out{spare, R.raw_fmt()};
        VT_TRY(run_attention(R, *M.attn, pl, ab, X, st_x, T, out, st_o, n, h, w_, fp32, es));
        advance(out);
        st_x = st_o;
    }
    {
        double* st_o = R.new_stats();
        Act out{spare, R.raw_fmt()};
        VT_TRY(R.resnet((*M.mid1), X, st_x, h, w_, lvl, T, Hb, out, st_o));
        advance(out);
        st_x = st_o;
    }
    return 0;
}

int run_encoder_microbatch(vt_ctx* c, vt_ctx::Lane& L, const vt_encode_args* a, int img0, int n, cudaStream_t s) {
    const vt_encoder_config& cfg = c->ecfg;
    const int H = a->height, Wd = a->width;
    const int fp32 = a->precision == VT_PREC_FP32;
    const size_t es = fp32 ? 4 : 2;
    const int C0 = cfg.block_out_channels[0];
    const int LC = cfg.latent_channels;
    const int nb = cfg.num_blocks;
    const int lh = H >> (nb - 1), lw = Wd >> (nb - 1);
    const long long tokens = 1LL * lh * lw;
    const int Cm = cfg.block_out_channels[nb - 1];

    // ---- workspace layout.  One activation buffer holds the largest tensor of any level (level 0).
    size_t act = static_cast<size_t>(n) * H * Wd * C0 * es;
    act = align_up(act, 1024);
    const AttnPlan pl = plan_attention(cfg.mid_block_add_attention != 0, n, tokens, Cm, es, fp32);
    const size_t attn_bytes = pl.attn_bytes;
    VT_TRY(L.arena.ensure(4 * act + attn_bytes));
    char* base = static_cast<char*>(L.arena.p);
    void* Xp = base;           // residual stream
    void* T = base + act;      // normalised operand
    void* Hb = base + 2 * act; // conv1 output / conv_in gather
    void* Yp = base + 3 * act; // block output
    char* ab = base + 4 * act;

    const int groups = cfg.norm_num_groups;
    const int max_slots = 64;
    VT_TRY(L.stats.ensure(static_cast<size_t>(max_slots) * n * groups * 2 * sizeof(double)));
    VT_TRY(L.statpart.ensure(stats_scratch_bytes(n, H, Wd)));
    VT_TRY(L.mom.ensure(static_cast<size_t>(n) * tokens * 2 * LC * sizeof(float)));

    EncRun R{c, s, fp32, n, static_cast<double*>(L.stats.p), 0, groups};
    R.ws.part = static_cast<float*>(L.statpart.p); R.ws.bytes = L.statpart.cap;
    {
        const char* e = getenv("VT_B200_NO_FUSED_GN");
        R.use_fused = !(e && e[0] == '1');
        const char* f = getenv("VT_B200_NO_FLASH");
        R.use_flash = !(f && f[0] == '1');
    }

    // ---- conv_in: gather the 3x3x3 patches (K = 27 padded to 64) then one K chunk of contraction
    const char* img = static_cast<const char*>(a->images);
    const size_t img_stride = a->in_fmt == VT_IN_U8_NHWC ? static_cast<size_t>(H) * Wd * 3
                                                         : static_cast<size_t>(H) * Wd * 3 * sizeof(float);
    double* st_x = R.new_stats();
    Act X{Xp, R.raw_fmt()};
    const bool convin_direct = !fp32 && cfg.block_out_channels[0] == 128 && c->conv_in.f16 &&
                               !(getenv("VT_B200_NO_CONVIN") && getenv("VT_B200_NO_CONVIN")[0] == '1');
    if (convin_direct) {
        // operand rows built in shared memory from the image itself (vt_convin.cuh)
        ConvInOp op;
        op.img = img + img_stride * img0; op.in_fmt = a->in_fmt; op.N = n; op.H = H; op.W = Wd;
        op.w = c->conv_in.w16; op.bias = c->conv_in.bias; op.out = X.p; op.stats = st_x; op.stats_ws = R.ws;
        op.out_f16 = R.raw16();
        VT_TRY(launch_conv_in(op, s, c->prof));
    } else {
        VT_TRY(launch_im2col3x3(img + img_stride * img0, a->in_fmt, Hb, R.opd_fmt(), n, H, Wd, s, c->prof));
        ConvW w = c->conv_in;  // viewed as a 1x1 conv over the 64-wide gathered patches
        w.Cin = 64; w.ksize = 1; w.Cs = 0;
        VT_TRY(R.conv(Hb, H, Wd, w, 1, nullptr, nullptr, X, st_x));
    }
    void* spare = Yp;
    auto advance = [&](Act out) { spare = X.p; X = out; };

    int h = H, w_ = Wd;
    for (int b = 0; b < nb; ++b) {
        const int nl = static_cast<int>(c->down[b].size());
        const bool has_down = c->downsample[b].Cout != 0;
        for (int l = 0; l < nl; ++l) {
            double* st_o = R.new_stats();
            Act out{spare, R.raw_fmt()};
            VT_TRY(R.resnet(c->down[b][l], X, st_x, h, w_, b, T, Hb, out, st_o));
            advance(out);
            st_x = st_o;
        }
        if (has_down) {
            double* st_o = R.new_stats();
            Act out{spare, R.raw_fmt()};
            VT_CHECK(fp32 || X.fmt == R.raw_fmt(), "downsample operand must be in the raw 16-bit format");
            VT_TRY(R.conv(X.p, h, w_, c->downsample[b], 2, nullptr, nullptr, out, st_o));
            advance(out);
            st_x = st_o;
            h /= 2; w_ /= 2;
        }
    }
    {   // ---- mid block
        MidW M{&c->mid0, &c->attn, &c->mid1, cfg.mid_block_add_attention != 0};
        VT_TRY(run_mid_block(R, M, pl, ab, X, spare, st_x, T, Hb, n, h, w_, fp32, es));
    }
    // ---- conv_norm_out + SiLU + conv_out -> moments (fp32 NHWC) -> DiagonalGaussian outputs
    VT_TRY(R.gn(X, T, st_x, c->norm_out, tokens, Cm, 1));
    VT_TRY(R.conv(T, h, w_, c->conv_out, 1, nullptr, nullptr, Act{L.mom.p, FMT_F32}, nullptr));
    const size_t lat_stride = static_cast<size_t>(LC) * tokens;
    VT_TRY(launch_moments_to_latent(static_cast<const float*>(L.mom.p),
                                    a->latent ? a->latent + lat_stride * img0 : nullptr,
                                    a->mean ? a->mean + lat_stride * img0 : nullptr,
                                    a->logvar ? a->logvar + lat_stride * img0 : nullptr,
                                    a->noise ? a->noise + lat_stride * img0 : nullptr, n, h, w_, LC, a->sample,
                                    a->seed + 0x9E37ULL * static_cast<unsigned long long>(img0),
                                    cfg.scaling_factor, cfg.shift_factor,
                                    a->apply_scale_shift && cfg.has_scaling_factor,
                                    a->apply_scale_shift && cfg.has_shift_factor, s, c->prof));
    VT_CHECK(R.stats_used <= max_slots, "GroupNorm statistics slots exhausted");
    return 0;
}

}  // namespace
#include "vt_train_encoder.cuh"
namespace {

// ---------------------------------------------------------------------------------------------
// The decoder schedule (diffusers Decoder.forward; reference call sites diffusers_vae_loader.py:72-76,
// :88-94): conv_in, mid block, four UpDecoderBlock2D (layers_per_block + 1 resnets, nearest-2x upsample +
// conv3x3 on all but the last), conv_norm_out + SiLU + conv_out.  Same kernels and formats as the encoder;
// the nearest-neighbour upsample is an explicit HBM pass into the operand buffer of the following conv.
int run_decoder_microbatch(vt_ctx* c, vt_ctx::Lane& L, const vt_decode_args* a, int img0, int n, cudaStream_t s) {
    const vt_encoder_config& cfg = c->ecfg;
    const int fp32 = a->precision == VT_PREC_FP32;
    const size_t es = fp32 ? 4 : 2;
    const int nb = cfg.num_blocks;
    const int LC = cfg.latent_channels;
    const int lh = a->lat_h, lw = a->lat_w;
    const int H = lh << (nb - 1), Wd = lw << (nb - 1);
    const long long tokens = 1LL * lh * lw;
    const int Cm = cfg.block_out_channels[nb - 1];
    const int groups = cfg.norm_num_groups;

    // ---- workspace: four ping-pong buffers sized for the largest tensor of the schedule
    size_t elems = static_cast<size_t>(n) * tokens * std::max(Cm, 64);
    {
        int h = lh, w = lw;
        for (int b = 0; b < nb; ++b) {
            const int cout = cfg.block_out_channels[nb - 1 - b];
            elems = std::max(elems, static_cast<size_t>(n) * h * w * std::max(cout, b ? cfg.block_out_channels[nb - b] : Cm));
            if (b < nb - 1) {
                h *= 2; w *= 2;
                elems = std::max(elems, static_cast<size_t>(n) * h * w * cout);
            }
        }
        elems = std::max(elems, static_cast<size_t>(n) * H * Wd * 32 * 4 / es);  // padded fp32 conv_out tile
    }
    const size_t act = align_up(elems * es, 1024);
    const AttnPlan pl = plan_attention(cfg.mid_block_add_attention != 0, n, tokens, Cm, es, fp32);
    VT_TRY(L.arena.ensure(4 * act + pl.attn_bytes));
    char* base = static_cast<char*>(L.arena.p);
    void* Xp = base;
    void* T = base + act;
    void* Hb = base + 2 * act;
    void* Yp = base + 3 * act;
    char* ab = base + 4 * act;
    const int max_slots = 64;
    VT_TRY(L.stats.ensure(static_cast<size_t>(max_slots) * n * groups * 2 * sizeof(double)));
    VT_TRY(L.statpart.ensure(stats_scratch_bytes(n, H, Wd)));

    EncRun R{c, s, fp32, n, static_cast<double*>(L.stats.p), 0, groups};
    R.ws.part = static_cast<float*>(L.statpart.p); R.ws.bytes = L.statpart.cap;
    {
        const char* e = getenv("VT_B200_NO_FUSED_GN");
        R.use_fused = !(e && e[0] == '1');
        const char* f = getenv("VT_B200_NO_FLASH");
        R.use_flash = !(f && f[0] == '1');
    }
    const bool use_subpixel = !(getenv("VT_B200_NO_SUBPIXEL") && getenv("VT_B200_NO_SUBPIXEL")[0] == '1');
    // ---- (z - shift) / scale (diffusers_vae_loader.py:88-93), NCHW fp32 -> NHWC padded to one K chunk
    const size_t lat_stride = static_cast<size_t>(LC) * tokens;
    const float shift = (a->apply_scale_shift && cfg.has_shift_factor) ? cfg.shift_factor : 0.f;
    const float inv_scale = (a->apply_scale_shift && cfg.has_scaling_factor) ? 1.0f / cfg.scaling_factor : 1.0f;
    VT_TRY(launch_latent_to_nhwc(a->latent + lat_stride * img0, Hb, R.raw_fmt(), n, LC, c->dconv_in.Cin, tokens, shift,
                                 inv_scale, s, c->prof));
    double* st_x = R.new_stats();
    Act X{Xp, R.raw_fmt()};
    VT_TRY(R.conv(Hb, lh, lw, c->dconv_in, 1, nullptr, nullptr, X, st_x));
    void* spare = Yp;
    auto advance = [&](Act out) { spare = X.p; X = out; };
    int h = lh, w_ = lw;
    {
        MidW M{&c->dmid0, &c->dattn, &c->dmid1, cfg.mid_block_add_attention != 0};
        VT_TRY(run_mid_block(R, M, pl, ab, X, spare, st_x, T, Hb, n, h, w_, fp32, es));
    }
    for (int b = 0; b < nb; ++b) {
        for (size_t l = 0; l < c->up[b].size(); ++l) {
            double* st_o = R.new_stats();
            Act out{spare, R.raw_fmt()};
            VT_TRY(R.resnet(c->up[b][l], X, st_x, h, w_, b, T, Hb, out, st_o));
            advance(out);
            st_x = st_o;
        }
        if (c->upsample[b].Cout != 0) {
            const ConvW& uw = c->upsample[b];
            const int C = uw.Cin;
            double* st_o = R.new_stats();
            Act out{spare, R.raw_fmt()};
            if (!fp32 && use_subpixel && c->upsample_sp[b]) {
                // Upsample2D without the upsampled tensor: four 2x2-tap convs of the source, one per output pixel
                // parity, with pre-summed taps -- 16/36 of the FLOPs and no extra HBM pass
                for (int par = 0; par < 4; ++par) {
                    ConvOp op;
                    op.in = X.p; op.N = n; op.Hin = h; op.Win = w_; op.Cin = C; op.ksize = 3; op.stride = 1;
                    op.in_f16 = R.raw16(); op.raw_f16 = R.raw16();
                    op.w = c->upsample_sp[b] + static_cast<size_t>(par) * C * 4 * C; op.Cout = C;
                    op.bias = uw.bias; op.out = out.p; op.out_fmt = out.fmt; op.stats = st_o;
                    op.stats_ws = R.ws; op.stats_ws.parts = 4; op.stats_ws.part_index = par;   // one tensor, four launches
                    op.up2 = 1; op.up_py = par >> 1; op.up_px = par & 1;
                    VT_TRY(launch_conv(op, s, c->prof));
                }
                h *= 2; w_ *= 2;
            } else {
                VT_TRY(launch_upsample2x_nhwc(X.p, T, static_cast<int>(es), n, h, w_, C, s, c->prof));
                h *= 2; w_ *= 2;
                VT_TRY(R.conv(T, h, w_, uw, 1, nullptr, nullptr, out, st_o));
            }
            advance(out);
            st_x = st_o;
        }
    }
    // ---- conv_norm_out + SiLU + conv_out (output channels padded to 32, fp32 NHWC) -> image NCHW fp32
    const int C0 = cfg.block_out_channels[0];
    VT_TRY(R.gn(X, T, st_x, c->dnorm_out, 1LL * h * w_, C0, 1));
    VT_TRY(R.conv(T, h, w_, c->dconv_out, 1, nullptr, nullptr, Act{Hb, FMT_F32}, nullptr));
    VT_TRY(launch_nhwc_to_image(static_cast<const float*>(Hb), a->image + static_cast<size_t>(3) * H * Wd * img0, n, 3,
                                c->dconv_out.Cout, 1LL * H * Wd, s, c->prof));
    VT_CHECK(R.stats_used <= max_slots, "GroupNorm statistics slots exhausted");
    return 0;
}

const float* hp(vt_ctx* c, const std::string& name) {
    auto it = c->hparams.find(name);
    return it == c->hparams.end() ? nullptr : it->second.dev;
}
int hcheck(vt_ctx* c, const std::string& name, std::initializer_list<int64_t> shape) {
    auto it = c->hparams.find(name);
    if (it == c->hparams.end()) {
        set_error("head parameter missing: " + name);
        return -4;
    }
    if (it->second.shape != std::vector<int64_t>(shape)) {
        std::string m = "head parameter " + name + " has shape [";
        for (auto d : it->second.shape) m += std::to_string(d) + ",";
        m += "] expected [";
        for (auto d : shape) m += std::to_string(d) + ",";
        set_error(m + "]");
        return -4;
    }
    return 0;
}

}  // namespace

// =============================================================================================
extern "C" {

const char* vt_last_error(void) { return vt::last_error(); }
int vt_abi_version(void) { return VT_ABI_VERSION; }

int vt_ctx_create(int device, vt_ctx** out) {
    VT_CHECK(out != nullptr, "null output pointer");
    int count = 0;
    VT_CUDA(cudaGetDeviceCount(&count));
    VT_CHECK(device >= 0 && device < count, "CUDA device index out of range");
    VT_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    VT_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        set_error(std::string("vae_tagger_b200 needs an sm_100a (B200) device; found ") + prop.name + " (sm_" +
                  std::to_string(prop.major) + std::to_string(prop.minor) + ")");
        return -5;
    }
    vt_ctx* c = new vt_ctx();
    c->device = device;
    {
        const char* e = getenv("VT_B200_RAW_BF16");
        c->raw_f16 = !(e && e[0] == '1');
    }
    c->prof = profiler_create();
    for (auto& L : c->lanes) {
        VT_CUDA(cudaStreamCreateWithFlags(&L.stream, cudaStreamNonBlocking));
        VT_CUDA(cudaEventCreateWithFlags(&L.done, cudaEventDisableTiming));
    }
    VT_CUDA(cudaEventCreateWithFlags(&c->ev_start, cudaEventDisableTiming));
    *out = c;
    return 0;
}

int vt_ctx_destroy(vt_ctx* c) {
    if (!c) return 0;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    free_packed(c);
    free_packed_decoder(c);
    free_params(c->eparams);
    free_params(c->dparams);
    free_params(c->hparams);
    for (auto& L : c->lanes) {
        L.arena.release(); L.stats.release(); L.statpart.release(); L.mom.release(); L.hws.release();
        if (L.stream) cudaStreamDestroy(L.stream);
        if (L.done) cudaEventDestroy(L.done);
    }
    if (c->ev_start) cudaEventDestroy(c->ev_start);
    c->hws.release(); c->e2e.release(); c->opws.release(); c->optws.release();
    for (auto& t : c->tapes)
        if (t) { t->release(); delete t; t = nullptr; }
    for (auto& t : c->dtapes)
        if (t) { t->release(); delete t; t = nullptr; }
    c->tbws.release(); c->tlws.release(); c->tg0.release();
    resize_cache_destroy(c->resize);
    profiler_destroy(c->prof);
    delete c;
    return 0;
}

// ------------------------------------------------------------------------------------- encoder
int vt_encoder_configure(vt_ctx* c, const vt_encoder_config* cfg) {
    VT_TRY(set_device(c));
    VT_CHECK(cfg != nullptr, "null config");
    VT_CHECK(cfg->in_channels == 3, "encoder in_channels must be 3");
    VT_CHECK(cfg->num_blocks >= 1 && cfg->num_blocks <= 8, "num_blocks must be in 1..8");
    VT_CHECK(cfg->norm_num_groups == 32, "norm_num_groups must be 32");
    VT_CHECK(cfg->layers_per_block >= 1 && cfg->layers_per_block <= 8, "layers_per_block must be in 1..8");
    for (int i = 0; i < cfg->num_blocks; ++i) {
        const int ch = cfg->block_out_channels[i];
        VT_CHECK(ch == 128 || ch == 256 || ch == 512,
                 "block_out_channels entries must be 128, 256 or 512 (tile configs exist for those)");
    }
    VT_CHECK(cfg->latent_channels == 16, "latent_channels must be 16 (conv_out is a 32-column tile)");
    free_packed(c);
    c->ecfg = *cfg;
    c->ecfg_set = true;
    return 0;
}

int vt_encoder_set_param(vt_ctx* c, const char* name, const float* data, const int64_t* shape, int ndim) {
    VT_TRY(set_device(c));
    c->enc_ready = false;
    return store_param(c->eparams, name, data, shape, ndim);
}

int vt_encoder_finalize(vt_ctx* c) {
    VT_TRY(set_device(c));
    VT_CHECK(c->ecfg_set, "vt_encoder_configure has not been called");
    free_packed(c);
    const vt_encoder_config& cfg = c->ecfg;
    Packer P{c, &c->eparams, &c->epacked, "encoder"};
    const int C0 = cfg.block_out_channels[0];
    VT_TRY(P.conv("conv_in", 3, C0, 3, "", 0, 64, &c->conv_in));
    c->down.resize(cfg.num_blocks);
    c->downsample.resize(cfg.num_blocks);
    int cin = C0;
    for (int b = 0; b < cfg.num_blocks; ++b) {
        const int cout = cfg.block_out_channels[b];
        c->down[b].resize(cfg.layers_per_block);
        for (int l = 0; l < cfg.layers_per_block; ++l) {
            VT_TRY(P.resnet("down_blocks." + std::to_string(b) + ".resnets." + std::to_string(l), l == 0 ? cin : cout,
                            cout, &c->down[b][l]));
        }
        if (b < cfg.num_blocks - 1)
            VT_TRY(P.conv("down_blocks." + std::to_string(b) + ".downsamplers.0.conv", cout, cout, 3, "", 0, 0,
                          &c->downsample[b], /*f16=*/false));
        cin = cout;
    }
    const int Cm = cfg.block_out_channels[cfg.num_blocks - 1];
    VT_TRY(P.resnet("mid_block.resnets.0", Cm, Cm, &c->mid0));
    VT_TRY(P.resnet("mid_block.resnets.1", Cm, Cm, &c->mid1));
    if (cfg.mid_block_add_attention) {
        const std::string a = "mid_block.attentions.0";
        c->attn.C = Cm;
        VT_TRY(P.norm(a + ".group_norm", Cm, &c->attn.gn));
        VT_TRY(P.linear({a + ".to_q", a + ".to_k"}, Cm, Cm, true, &c->attn.qk));
        VT_TRY(P.linear({a + ".to_v"}, Cm, Cm, true, &c->attn.v));
        VT_TRY(P.linear({a + ".to_out.0"}, Cm, Cm, true, &c->attn.out));
    }
    VT_TRY(P.norm("conv_norm_out", Cm, &c->norm_out));
    VT_TRY(P.conv("conv_out", Cm, 2 * cfg.latent_channels, 3, "", 0, 0, &c->conv_out));
    VT_CUDA(cudaDeviceSynchronize());
    c->enc_ready = true;
    return 0;
}

// host_src != nullptr: a->images is a device staging buffer that is filled from host_src (pinned) one
// micro-batch at a time, on the stream that runs the micro-batch -- with two lanes the upload of micro-batch
// k+1 overlaps the kernels of micro-batch k (vt_infer_host).
struct TagTail {  // the tag head of a micro-batch, run right behind its encoder on the same stream
    float threshold;
    float* conf;
    int64_t* idx;
    int32_t* cnt;
};
static int tag_impl(vt_ctx* c, const vt_tag_args* a, DevBuf& hws);

static int encode_impl(vt_ctx* c, const vt_encode_args* a, const char* host_src, const TagTail* tail = nullptr) {
    VT_TRY(set_device(c));
    VT_CHECK(a != nullptr, "null arguments");
    VT_CHECK(c->enc_ready, "encoder parameters not finalised (vt_encoder_finalize)");
    VT_CHECK(a->images != nullptr, "null image pointer");
    VT_CHECK(a->batch > 0, "batch must be positive");
    const int down = 1 << (c->ecfg.num_blocks - 1);
    // any size from 2^(num_blocks-1) up: every Downsample2D halves with floor, like the reference's
    VT_CHECK(a->height >= down && a->width >= down, "image height and width must be at least 2^(num_blocks-1)");
    VT_CHECK(a->in_fmt == VT_IN_F32_NCHW || a->in_fmt == VT_IN_U8_NHWC, "unknown image format");
    VT_CHECK(a->precision == VT_PREC_BF16 || a->precision == VT_PREC_FP32, "unknown precision");
    cudaStream_t s = static_cast<cudaStream_t>(a->stream);
    int mb = a->micro_batch;
    if (mb <= 0) {
        // keep the top-level activation of one pass near 1 GB (bf16): 2 lanes x 4 ping-pong buffers
        const double per_img = 1.0 * a->height * a->width * c->ecfg.block_out_channels[0] * 2;
        mb = static_cast<int>(std::max(1.0, std::min(32.0, (1.1e9) / per_img)));
    }
    const size_t img_stride = static_cast<size_t>(a->height) * a->width * 3 * (a->in_fmt == VT_IN_U8_NHWC ? 1 : 4);
    auto upload = [&](int i0, int n, cudaStream_t st) -> int {
        if (host_src)
            VT_CUDA(cudaMemcpyAsync(const_cast<char*>(static_cast<const char*>(a->images)) + img_stride * i0,
                                    host_src + img_stride * i0, img_stride * n, cudaMemcpyHostToDevice, st));
        return 0;
    };
    auto tag_tail = [&](vt_ctx::Lane& L, int i0, int n, cudaStream_t st) -> int {
        if (!tail) return 0;
        const int lh = a->height / down, lw = a->width / down, T = c->hcfg.num_classes;
        vt_tag_args t{};
        t.latent = a->latent + static_cast<size_t>(i0) * c->ecfg.latent_channels * lh * lw;
        t.batch = n; t.lat_h = lh; t.lat_w = lw; t.threshold = tail->threshold;
        t.conf_sorted = tail->conf + static_cast<size_t>(i0) * T;
        t.idx_sorted = tail->idx + static_cast<size_t>(i0) * T;
        t.count = tail->cnt + i0;
        t.stream = st;
        return tag_impl(c, &t, L.hws);
    };
    if (mb >= a->batch || a->single_lane) {
        for (int i0 = 0; i0 < a->batch; i0 += mb) {
            const int n = std::min(mb, a->batch - i0);
            VT_TRY(upload(i0, n, s));
            VT_TRY(run_encoder_microbatch(c, c->lanes[0], a, i0, n, s));
            VT_TRY(tag_tail(c->lanes[0], i0, n, s));
        }
        return 0;
    }
    // several micro-batches: alternate between the two lanes
    VT_CUDA(cudaEventRecord(c->ev_start, s));
    for (auto& L : c->lanes) VT_CUDA(cudaStreamWaitEvent(L.stream, c->ev_start, 0));
    int k = 0;
    for (int i0 = 0; i0 < a->batch; i0 += mb, ++k) {
        vt_ctx::Lane& L = c->lanes[k & 1];
        const int n = std::min(mb, a->batch - i0);
        VT_TRY(upload(i0, n, L.stream));
        VT_TRY(run_encoder_microbatch(c, L, a, i0, n, L.stream));
        VT_TRY(tag_tail(L, i0, n, L.stream));
    }
    for (auto& L : c->lanes) {
        VT_CUDA(cudaEventRecord(L.done, L.stream));
        VT_CUDA(cudaStreamWaitEvent(s, L.done, 0));
    }
    return 0;
}

int vt_encode(vt_ctx* c, const vt_encode_args* a) { return encode_impl(c, a, nullptr); }

// ------------------------------------------------------------------------------------- VAE decoder
int vt_decoder_set_param(vt_ctx* c, const char* name, const float* data, const int64_t* shape, int ndim) {
    VT_TRY(set_device(c));
    c->dec_ready = false;
    return store_param(c->dparams, name, data, shape, ndim);
}

int vt_decoder_finalize(vt_ctx* c) {
    VT_TRY(set_device(c));
    VT_CHECK(c->ecfg_set, "vt_encoder_configure has not been called");
    free_packed_decoder(c);
    const vt_encoder_config& cfg = c->ecfg;
    Packer P{c, &c->dparams, &c->dpacked, "decoder"};
    const int nb = cfg.num_blocks;
    const int Cm = cfg.block_out_channels[nb - 1];
    VT_CHECK(cfg.latent_channels <= 64, "decoder: latent_channels must be at most 64");
    // the latent is a raw (unbounded) activation: bf16 operand, bf16 weight columns
    VT_TRY(P.conv_padded("conv_in", cfg.latent_channels, Cm, 3, 64, Cm, &c->dconv_in, /*f16=*/false));
    VT_TRY(P.resnet("mid_block.resnets.0", Cm, Cm, &c->dmid0));
    VT_TRY(P.resnet("mid_block.resnets.1", Cm, Cm, &c->dmid1));
    if (cfg.mid_block_add_attention) {
        const std::string a = "mid_block.attentions.0";
        c->dattn.C = Cm;
        VT_TRY(P.norm(a + ".group_norm", Cm, &c->dattn.gn));
        VT_TRY(P.linear({a + ".to_q", a + ".to_k"}, Cm, Cm, true, &c->dattn.qk));
        VT_TRY(P.linear({a + ".to_v"}, Cm, Cm, true, &c->dattn.v));
        VT_TRY(P.linear({a + ".to_out.0"}, Cm, Cm, true, &c->dattn.out));
    }
    c->up.resize(nb);
    c->upsample.resize(nb);
    c->upsample_sp.assign(nb, nullptr);
    int cin = Cm;
    for (int b = 0; b < nb; ++b) {
        const int cout = cfg.block_out_channels[nb - 1 - b];
        c->up[b].resize(cfg.layers_per_block + 1);
        for (int l = 0; l <= cfg.layers_per_block; ++l)
            VT_TRY(P.resnet("up_blocks." + std::to_string(b) + ".resnets." + std::to_string(l), l == 0 ? cin : cout,
                            cout, &c->up[b][l]));
        if (b < nb - 1) {
            const std::string up = "up_blocks." + std::to_string(b) + ".upsamplers.0.conv";
            VT_TRY(P.conv(up, cout, cout, 3, "", 0, 0, &c->upsample[b], /*f16=*/false));
            VT_TRY(P.conv_subpixel(up, cout, &c->upsample_sp[b]));
        }
        cin = cout;
    }
    const int C0 = cfg.block_out_channels[0];
    VT_TRY(P.norm("conv_norm_out", C0, &c->dnorm_out));
    VT_TRY(P.conv_padded("conv_out", C0, 3, 3, C0, 32, &c->dconv_out, /*f16=*/true));
    VT_CUDA(cudaDeviceSynchronize());
    c->dec_ready = true;
    return 0;
}

int vt_decode(vt_ctx* c, const vt_decode_args* a) {
    VT_TRY(set_device(c));
    VT_CHECK(a != nullptr, "null arguments");
    VT_CHECK(c->dec_ready, "decoder parameters not finalised (vt_decoder_finalize)");
    VT_CHECK(a->latent != nullptr && a->image != nullptr, "null latent / image pointer");
    VT_CHECK(a->batch > 0 && a->lat_h > 0 && a->lat_w > 0, "batch and latent size must be positive");
    VT_CHECK(a->precision == VT_PREC_BF16 || a->precision == VT_PREC_FP32, "unknown precision");
    cudaStream_t s = static_cast<cudaStream_t>(a->stream);
    int mb = a->micro_batch;
    if (mb <= 0) {
        const int up = 1 << (c->ecfg.num_blocks - 1);
        const double per_img = 4.0 * a->lat_h * up * a->lat_w * up * c->ecfg.block_out_channels[0];
        mb = static_cast<int>(std::max(1.0, std::min(32.0, (2.2e9) / per_img)));
    }
    for (int i0 = 0; i0 < a->batch; i0 += mb)
        VT_TRY(run_decoder_microbatch(c, c->lanes[0], a, i0, std::min(mb, a->batch - i0), s));
    return 0;
}

// ------------------------------------------------------------------------------------- head
int vt_head_configure(vt_ctx* c, const vt_head_config* cfg) {
    VT_TRY(set_device(c));
    VT_CHECK(cfg != nullptr, "null config");
    VT_CHECK(cfg->kind == VT_HEAD_ATTENTION || cfg->kind == VT_HEAD_PLAIN, "unknown head kind");
    VT_CHECK(cfg->latent_channels >= 8 && cfg->latent_channels <= 32 && cfg->latent_channels % 8 == 0,
             "latent_channels must be 8, 16, 24 or 32");
    VT_CHECK(cfg->num_classes >= 1 && cfg->num_classes <= 16384, "num_classes must be in 1..16384");
    if (cfg->kind == VT_HEAD_ATTENTION && cfg->use_cross_attention)
        VT_CHECK(cfg->attention_heads == 8, "the cross-attention branch needs attention_heads = 8 (embed 256, head_dim 32)");
    if (cfg->kind == VT_HEAD_ATTENTION && cfg->use_self_attention)
        VT_CHECK(cfg->attention_heads >= 1 && (cfg->latent_channels / 2) % cfg->attention_heads == 0,
                 "embed_dim must be divisible by num_heads (modules.py:56)");
    c->hcfg = *cfg;
    c->hcfg_set = true;
    c->head_ready = false;
    return 0;
}

int vt_head_set_param(vt_ctx* c, const char* name, const float* data, const int64_t* shape, int ndim) {
    VT_TRY(set_device(c));
    c->head_ready = false;
    return store_param(c->hparams, name, data, shape, ndim);
}

int vt_head_finalize(vt_ctx* c) {
    VT_TRY(set_device(c));
    VT_CHECK(c->hcfg_set, "vt_head_configure has not been called");
    const vt_head_config& h = c->hcfg;
    const int C = h.latent_channels, E = C / 2, T = h.num_classes;
    if (h.kind == VT_HEAD_ATTENTION) {
        if (h.use_spatial_attention) {
            VT_TRY(hcheck(c, "spatial_attention.channel_att.0.weight", {C / 8, C, 1, 1}));
            VT_TRY(hcheck(c, "spatial_attention.channel_att.2.weight", {C, C / 8, 1, 1}));
            VT_TRY(hcheck(c, "spatial_attention.spatial_att.0.weight", {1, 2, 7, 7}));
        }
        VT_TRY(hcheck(c, "feature_compress.0.weight", {E, C, 3, 3}));
        VT_TRY(hcheck(c, "feature_compress.0.bias", {E}));
        for (const char* k : {"weight", "bias", "running_mean", "running_var"})
            VT_TRY(hcheck(c, std::string("feature_compress.1.") + k, {E}));
        if (h.use_self_attention) {
            for (const char* k : {"q_proj", "k_proj", "v_proj", "out_proj"}) {
                VT_TRY(hcheck(c, std::string("self_attention_post.") + k + ".weight", {E, E}));
                VT_TRY(hcheck(c, std::string("self_attention_post.") + k + ".bias", {E}));
            }
            VT_TRY(hcheck(c, "self_attention_post.norm.weight", {E}));
            VT_TRY(hcheck(c, "self_attention_post.norm.bias", {E}));
        }
        if (h.use_cross_attention) {
            VT_TRY(hcheck(c, "query_generator.weight", {512, E * 64}));
            VT_TRY(hcheck(c, "query_generator.bias", {512}));
            VT_TRY(hcheck(c, "cross_attention.q_proj.weight", {256, 512}));
            VT_TRY(hcheck(c, "cross_attention.q_proj.bias", {256}));
            for (const char* k : {"k_proj", "v_proj"}) {
                VT_TRY(hcheck(c, std::string("cross_attention.") + k + ".weight", {256, E}));
                VT_TRY(hcheck(c, std::string("cross_attention.") + k + ".bias", {256}));
            }
            VT_TRY(hcheck(c, "cross_attention.out_proj.weight", {512, 256}));
            VT_TRY(hcheck(c, "cross_attention.out_proj.bias", {512}));
        }
        const int dims[5] = {E * 64, 1024, 512, 256, T};
        const int lin[4] = {0, 4, 8, 12}, ln[3] = {1, 5, 9};
        for (int i = 0; i < 4; ++i) {
            VT_TRY(hcheck(c, "classifier." + std::to_string(lin[i]) + ".weight", {dims[i + 1], dims[i]}));
            VT_TRY(hcheck(c, "classifier." + std::to_string(lin[i]) + ".bias", {dims[i + 1]}));
            if (i < 3) {
                VT_TRY(hcheck(c, "classifier." + std::to_string(ln[i]) + ".weight", {dims[i + 1]}));
                VT_TRY(hcheck(c, "classifier." + std::to_string(ln[i]) + ".bias", {dims[i + 1]}));
            }
        }
    } else {
        const int dims[4] = {h.plain_flat_dim > 0 ? h.plain_flat_dim : C * 16, 512, 256, T};
        const int lin[3] = {0, 4, 8}, ln[2] = {1, 5};
        for (int i = 0; i < 3; ++i) {
            VT_TRY(hcheck(c, "classifier." + std::to_string(lin[i]) + ".weight", {dims[i + 1], dims[i]}));
            VT_TRY(hcheck(c, "classifier." + std::to_string(lin[i]) + ".bias", {dims[i + 1]}));
            if (i < 2) {
                VT_TRY(hcheck(c, "classifier." + std::to_string(ln[i]) + ".weight", {dims[i + 1]}));
                VT_TRY(hcheck(c, "classifier." + std::to_string(ln[i]) + ".bias", {dims[i + 1]}));
            }
        }
    }
    c->head_ready = true;
    return 0;
}

static int tag_impl(vt_ctx* c, const vt_tag_args* a, DevBuf& hws);
int vt_tag(vt_ctx* c, const vt_tag_args* a) {
    VT_CHECK(c != nullptr, "null context");
    return tag_impl(c, a, c->hws);
}

static int tag_impl(vt_ctx* c, const vt_tag_args* a, DevBuf& hws) {
    VT_TRY(set_device(c));
    VT_CHECK(a != nullptr, "null arguments");
    VT_CHECK(c->head_ready, "head parameters not finalised (vt_head_finalize)");
    VT_CHECK(a->latent != nullptr && a->batch > 0 && a->lat_h > 0 && a->lat_w > 0, "bad latent arguments");
    const vt_head_config& h = c->hcfg;
    const int B = a->batch, C = h.latent_channels, E = C / 2, T = h.num_classes;
    const int H = a->lat_h, W = a->lat_w, HW = H * W;
    cudaStream_t s = static_cast<cudaStream_t>(a->stream);
    Profiler* pf = c->prof;
    // workspace (floats)
    size_t off = 0;
    auto take = [&](size_t n) { size_t o = off; off += (n + 63) / 64 * 64; return o; };
    const size_t o_pool = take(static_cast<size_t>(B) * C * 2), o_cg = take(static_cast<size_t>(B) * C);
    const size_t o_map = take(static_cast<size_t>(B) * 2 * HW), o_x2 = take(static_cast<size_t>(B) * C * HW);
    const size_t o_pooled = take(static_cast<size_t>(B) * E * 64), o_feat = take(static_cast<size_t>(B) * 1024);
    const size_t o_a = take(static_cast<size_t>(B) * 1024), o_b = take(static_cast<size_t>(B) * 1024);
    const size_t o_logits = take(static_cast<size_t>(B) * T);
    const size_t o_x1 = take(static_cast<size_t>(B) * (512 + 256 + 256 + 512)), o_feat2 = take(static_cast<size_t>(B) * 1024);
    const size_t o_ao = take(static_cast<size_t>(B) * 64 * E);
    VT_TRY(hws.ensure(off * sizeof(float)));
    float* ws = static_cast<float*>(hws.p);
    float* logits = a->logits ? a->logits : ws + o_logits;

    if (h.kind == VT_HEAD_ATTENTION) {
        const float* x = a->latent;
        if (h.use_spatial_attention) {
            VT_TRY(launch_head_spatial_attention(a->latent, hp(c, "spatial_attention.channel_att.0.weight"),
                                                 hp(c, "spatial_attention.channel_att.2.weight"),
                                                 hp(c, "spatial_attention.spatial_att.0.weight"), ws + o_pool,
                                                 ws + o_cg, ws + o_map, ws + o_x2, nullptr, B, C, H, W, s, pf));
            x = ws + o_x2;
        }
        VT_TRY(launch_head_compress(x, hp(c, "feature_compress.0.weight"), hp(c, "feature_compress.0.bias"),
                                    hp(c, "feature_compress.1.weight"), hp(c, "feature_compress.1.bias"),
                                    hp(c, "feature_compress.1.running_mean"), hp(c, "feature_compress.1.running_var"),
                                    1e-5f, ws + o_pooled, B, C, H, W, s, pf));
        const std::string sp = "self_attention_post.";
        const float* mp[10] = {hp(c, sp + "norm.weight"), hp(c, sp + "norm.bias"),
                               hp(c, sp + "q_proj.weight"), hp(c, sp + "q_proj.bias"),
                               hp(c, sp + "k_proj.weight"), hp(c, sp + "k_proj.bias"),
                               hp(c, sp + "v_proj.weight"), hp(c, sp + "v_proj.bias"),
                               hp(c, sp + "out_proj.weight"), hp(c, sp + "out_proj.bias")};
        VT_TRY(launch_head_mhsa(ws + o_pooled, mp, ws + o_feat, ws + o_ao, B, E, h.use_self_attention ? h.attention_heads : 1,
                                h.use_self_attention, s, pf));
        const int dims[5] = {E * 64, 1024, 512, 256, T};
        const int lin[4] = {0, 4, 8, 12}, ln[3] = {1, 5, 9};
        const float* cur = ws + o_feat;
        float* bufs[2] = {ws + o_a, ws + o_b};
        if (h.use_cross_attention) {
            // modules.py:450-459: query = query_generator(flat); attended = CrossAttention(query, tokens);
            // flat += mean(attended) -- the Linear layers are the batched warp-per-neuron kernel
            float* query = ws + o_x1;
            float* qp = query + static_cast<size_t>(B) * 512;
            float* att = qp + static_cast<size_t>(B) * 256;
            float* outp = att + static_cast<size_t>(B) * 256;
            VT_TRY(launch_head_linear(cur, hp(c, "query_generator.weight"), hp(c, "query_generator.bias"), query, B,
                                      E * 64, 512, s, pf));
            VT_TRY(launch_head_linear(query, hp(c, "cross_attention.q_proj.weight"), hp(c, "cross_attention.q_proj.bias"),
                                      qp, B, 512, 256, s, pf));
            VT_TRY(launch_head_cross_attention(cur, qp, hp(c, "cross_attention.k_proj.weight"),
                                               hp(c, "cross_attention.k_proj.bias"), hp(c, "cross_attention.v_proj.weight"),
                                               hp(c, "cross_attention.v_proj.bias"), att, B, E, h.attention_heads, s, pf));
            VT_TRY(launch_head_linear(att, hp(c, "cross_attention.out_proj.weight"),
                                      hp(c, "cross_attention.out_proj.bias"), outp, B, 256, 512, s, pf));
            VT_TRY(launch_head_cross_add(outp, query, cur, ws + o_feat2, B, 512, E * 64, s, pf));
            cur = ws + o_feat2;
        }
        for (int i = 0; i < 4; ++i) {
            float* out = i == 3 ? logits : bufs[i & 1];
            VT_TRY(launch_head_linear(cur, hp(c, "classifier." + std::to_string(lin[i]) + ".weight"),
                                      hp(c, "classifier." + std::to_string(lin[i]) + ".bias"), out, B, dims[i],
                                      dims[i + 1], s, pf));
            if (i < 3)
                VT_TRY(launch_head_ln_act(out, hp(c, "classifier." + std::to_string(ln[i]) + ".weight"),
                                          hp(c, "classifier." + std::to_string(ln[i]) + ".bias"), B, dims[i + 1], 1,
                                          s, pf));
            cur = out;
        }
    } else {
        const float* cur = ws + o_feat;
        if (h.plain_flat_dim > 0) {  // use_adaptive_pooling=False: latent.reshape(B, -1) is the NCHW buffer itself
            VT_CHECK(h.plain_flat_dim == C * H * W, "latent size differs from the one the plain head was built for");
            cur = a->latent;
        } else {
            VT_TRY(launch_head_adaptive_pool(a->latent, ws + o_feat, B, C, H, W, 4, 4, s, pf));
        }
        const int dims[4] = {h.plain_flat_dim > 0 ? h.plain_flat_dim : C * 16, 512, 256, T};
        const int lin[3] = {0, 4, 8}, ln[2] = {1, 5};
        float* bufs[2] = {ws + o_a, ws + o_b};
        for (int i = 0; i < 3; ++i) {
            float* out = i == 2 ? logits : bufs[i & 1];
            VT_TRY(launch_head_linear(cur, hp(c, "classifier." + std::to_string(lin[i]) + ".weight"),
                                      hp(c, "classifier." + std::to_string(lin[i]) + ".bias"), out, B, dims[i],
                                      dims[i + 1], s, pf));
            if (i < 2)
                VT_TRY(launch_head_ln_act(out, hp(c, "classifier." + std::to_string(ln[i]) + ".weight"),
                                          hp(c, "classifier." + std::to_string(ln[i]) + ".bias"), B, dims[i + 1], 2,
                                          s, pf));
            cur = out;
        }
    }
    if (a->conf_sorted || a->idx_sorted || a->count || a->probs)
        VT_TRY(launch_head_confidence(logits, a->conf_sorted, reinterpret_cast<long long*>(a->idx_sorted), a->count,
                                      a->probs, B, T, a->threshold, s, pf));
    return 0;
}

// ------------------------------------------------------------------------------------- head training
int vt_head_param_count(vt_ctx* c, int32_t* n_tensors, int64_t* n_floats) {
    VT_CHECK(c != nullptr && c->hcfg_set, "vt_head_configure has not been called");
    const auto L = head_param_layout(c->hcfg);
    if (n_tensors) *n_tensors = static_cast<int32_t>(L.size());
    if (n_floats) *n_floats = L.empty() ? 0 : L.back().offset + L.back().numel;
    return 0;
}

int vt_head_param_layout(vt_ctx* c, int32_t index, char* name, int32_t name_cap, int64_t* offset, int64_t* numel) {
    VT_CHECK(c != nullptr && c->hcfg_set, "vt_head_configure has not been called");
    const auto L = head_param_layout(c->hcfg);
    VT_CHECK(index >= 0 && index < static_cast<int32_t>(L.size()), "parameter index out of range");
    const auto& e = L[index];
    if (name) {
        VT_CHECK(name_cap > static_cast<int32_t>(e.name.size()), "name buffer too small");
        std::copy(e.name.begin(), e.name.end(), name);
        name[e.name.size()] = 0;
    }
    if (offset) *offset = e.offset;
    if (numel) *numel = e.numel;
    return 0;
}

int vt_head_train_step(vt_ctx* c, const vt_head_train_args* a) {
    VT_TRY(set_device(c));
    VT_CHECK(a != nullptr, "null arguments");
    VT_CHECK(c->hcfg_set, "vt_head_configure has not been called");
    VT_CHECK(a->latent && a->targets && a->params, "latent, targets and params are required");
    VT_CHECK(a->batch > 0 && a->lat_h > 0 && a->lat_w > 0, "bad latent arguments");
    VT_CHECK(a->attention_dropout >= 0.f && a->attention_dropout < 1.f, "attention_dropout must be in [0,1)");
    if (c->hcfg.kind == VT_HEAD_ATTENTION)
        VT_CHECK(1LL * a->batch * a->lat_h * a->lat_w > 1, "BatchNorm in train mode needs more than one value per channel");
    VT_TRY(c->hws.ensure(head_train_workspace_floats(c->hcfg, a->batch, a->lat_h, a->lat_w) * sizeof(float)));
    return head_train_step(c->hcfg, *a, static_cast<float*>(c->hws.p), c->prof);
}

int vt_head_dropout_masks(vt_ctx* c, int batch, float attention_dropout, uint64_t seed, float* attn, float* cls0,
                          float* cls1, float* cls2, void* stream) {
    VT_TRY(set_device(c));
    VT_CHECK(c->hcfg_set && batch > 0, "vt_head_configure has not been called");
    float* cls[3] = {cls0, cls1, cls2};
    return head_dropout_masks(c->hcfg, batch, attention_dropout, seed, attn, cls, static_cast<cudaStream_t>(stream));
}

int vt_adamw_step(vt_ctx* c, float* params, float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                  float beta1, float beta2, float eps, float weight_decay, int64_t step, float grad_scale,
                  float max_norm, int zero_grad, float* norm_out, void* stream) {
    VT_TRY(set_device(c));
    VT_CHECK(params && grads && exp_avg && exp_avg_sq && n > 0 && step >= 1, "bad optimizer arguments");
    VT_TRY(c->optws.ensure(1184 * sizeof(double) + 64));
    double* scratch = static_cast<double*>(c->optws.p);
    float* norm = norm_out ? norm_out : reinterpret_cast<float*>(scratch + 1184);
    return launch_adamw(params, grads, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, weight_decay, step, grad_scale,
                        max_norm, scratch, norm, zero_grad, static_cast<cudaStream_t>(stream), c->prof);
}

// ------------------------------------------------------------------------------------- preprocessing
int vt_resize_u8(vt_ctx* c, const vt_resize_args* a) {
    VT_TRY(set_device(c));
    VT_CHECK(a != nullptr && a->src != nullptr && a->dst != nullptr, "null arguments");
    VT_CHECK(a->filter == VT_FILTER_LANCZOS || a->filter == VT_FILTER_BILINEAR, "unknown resize filter");
    VT_CHECK(a->src_w > 0 && a->src_h > 0 && a->dst_w > 0 && a->dst_h > 0, "image sizes must be positive");
    VT_CHECK(a->crop_l >= 0 && a->crop_t >= 0 && a->crop_r <= a->src_w && a->crop_b <= a->src_h &&
                 a->crop_l < a->crop_r && a->crop_t < a->crop_b,
             "crop box must be a non-empty box inside the image");
    VT_CHECK(a->src_stride >= 3LL * a->src_w && a->dst_stride >= 3LL * a->dst_w, "row stride smaller than a row");
    if (!c->resize) c->resize = resize_cache_create();
    return resize_u8(c->resize, *a, c->prof);
}

int vt_resize_u8_batch(vt_ctx* c, const vt_resize_args* items, int n, void* stream) {
    VT_CHECK(items != nullptr && n >= 0, "null arguments");
    for (int i = 0; i < n; ++i) {
        vt_resize_args a = items[i];
        a.stream = stream;
        VT_TRY(vt_resize_u8(c, &a));
    }
    return 0;
}

int vt_smart_crop_box(int src_w, int src_h, int dst_w, int dst_h, int32_t* box4) {
    VT_CHECK(box4 != nullptr && src_w > 0 && src_h > 0 && dst_w > 0 && dst_h > 0, "bad crop box arguments");
    int b[4];
    smart_crop_box(src_w, src_h, dst_w, dst_h, b);
    for (int i = 0; i < 4; ++i) box4[i] = b[i];
    return 0;
}

int vt_resize_coefficients(int in_size, int out_size, int filter, int32_t* ksize, int32_t* bounds, int32_t* kk) {
    VT_CHECK(in_size > 0 && out_size > 0 && ksize != nullptr, "bad coefficient arguments");
    VT_CHECK(filter == VT_FILTER_LANCZOS || filter == VT_FILTER_BILINEAR, "unknown resize filter");
    std::vector<int> b, k;
    int ks = 0;
    resize_coefficients(in_size, out_size, filter, &ks, b, k);
    *ksize = ks;
    if (bounds) std::copy(b.begin(), b.end(), bounds);
    if (kk) std::copy(k.begin(), k.end(), kk);
    return 0;
}

// ------------------------------------------------------------------------------------- e2e
int vt_infer_host(vt_ctx* c, const vt_infer_host_args* a) {
    VT_TRY(set_device(c));
    VT_CHECK(a != nullptr && a->images_host != nullptr, "null arguments");
    VT_CHECK(c->enc_ready && c->head_ready, "encoder and head must be finalised");
    cudaStream_t s = static_cast<cudaStream_t>(a->stream);
    const int B = a->batch, H = a->height, W = a->width;
    const int LC = c->ecfg.latent_channels, T = c->hcfg.num_classes;
    const int down = 1 << (c->ecfg.num_blocks - 1);
    VT_CHECK(B > 0 && H >= down && W >= down, "bad image shape");
    const int lh = H / down, lw = W / down;
    const size_t img_b = align_up(static_cast<size_t>(B) * H * W * 3 * (a->in_fmt == VT_IN_U8_NHWC ? 1 : 4), 256);
    const size_t lat_b = align_up(static_cast<size_t>(B) * LC * lh * lw * 4, 256);
    const size_t conf_b = align_up(static_cast<size_t>(B) * T * 4, 256), idx_b = align_up(static_cast<size_t>(B) * T * 8, 256);
    const size_t cnt_b = align_up(static_cast<size_t>(B) * 4, 256);
    VT_TRY(c->e2e.ensure(img_b + lat_b + conf_b + idx_b + cnt_b));
    char* d = static_cast<char*>(c->e2e.p);
    char* d_img = d; float* d_lat = reinterpret_cast<float*>(d + img_b);
    float* d_conf = reinterpret_cast<float*>(d + img_b + lat_b);
    int64_t* d_idx = reinterpret_cast<int64_t*>(d + img_b + lat_b + conf_b);
    int32_t* d_cnt = reinterpret_cast<int32_t*>(d + img_b + lat_b + conf_b + idx_b);
    // the upload is issued per micro-batch inside encode_impl, on the stream that consumes it
    vt_encode_args e{};
    e.images = d_img; e.in_fmt = a->in_fmt; e.batch = B; e.height = H; e.width = W; e.precision = a->precision;
    e.sample = 0; e.apply_scale_shift = 1; e.latent = d_lat; e.micro_batch = a->micro_batch; e.stream = a->stream;
    // the head of each micro-batch runs right behind its encoder on the same lane: it hides under the other
    // lane's contractions instead of trailing the whole batch
    const TagTail tail{a->threshold, d_conf, d_idx, d_cnt};
    VT_TRY(encode_impl(c, &e, static_cast<const char*>(a->images_host), &tail));
    if (a->conf_sorted_host)
        VT_CUDA(cudaMemcpyAsync(a->conf_sorted_host, d_conf, static_cast<size_t>(B) * T * 4, cudaMemcpyDeviceToHost, s));
    if (a->idx_sorted_host)
        VT_CUDA(cudaMemcpyAsync(a->idx_sorted_host, d_idx, static_cast<size_t>(B) * T * 8, cudaMemcpyDeviceToHost, s));
    if (a->count_host)
        VT_CUDA(cudaMemcpyAsync(a->count_host, d_cnt, static_cast<size_t>(B) * 4, cudaMemcpyDeviceToHost, s));
    if (a->latent_host)
        VT_CUDA(cudaMemcpyAsync(a->latent_host, d_lat, static_cast<size_t>(B) * LC * lh * lw * 4,
                                cudaMemcpyDeviceToHost, s));
    VT_CUDA(cudaStreamSynchronize(s));
    return 0;
}

int vt_infer(vt_ctx* c, const vt_infer_args* a) {
    VT_TRY(set_device(c));
    VT_CHECK(a != nullptr && a->images != nullptr, "null arguments");
    VT_CHECK(c->enc_ready && c->head_ready, "encoder and head must be finalised");
    VT_CHECK(a->latent && a->conf_sorted && a->idx_sorted && a->count, "latent, conf_sorted, idx_sorted and count are required");
    vt_encode_args e{};
    e.images = a->images; e.in_fmt = a->in_fmt; e.batch = a->batch; e.height = a->height; e.width = a->width;
    e.precision = a->precision; e.sample = 0; e.apply_scale_shift = 1; e.latent = a->latent;
    e.micro_batch = a->micro_batch; e.single_lane = a->single_lane; e.stream = a->stream;
    const TagTail tail{a->threshold, a->conf_sorted, a->idx_sorted, a->count};
    return encode_impl(c, &e, nullptr, &tail);
}

int vt_focal_loss(vt_ctx* c, const float* logits, const float* targets, int64_t n, float alpha, float gamma,
                  float grad_scale, float* loss_sum, float* grad, void* stream) {
    VT_TRY(set_device(c));
    VT_CHECK(logits && targets && n > 0, "bad focal loss arguments");
    return launch_focal_loss(logits, targets, loss_sum, grad, n, alpha, gamma, grad_scale,
                             static_cast<cudaStream_t>(stream), c->prof);
}

// ------------------------------------------------------------------------------------- accounting
int vt_profile_enable(vt_ctx* c, int timing) {
    VT_TRY(set_device(c));
    profiler_enable(c->prof, timing != 0);
    return 0;
}
int vt_profile_read(vt_ctx* c, double* out, int reset) {
    VT_TRY(set_device(c));
    VT_CHECK(out != nullptr, "null output");
    VT_CUDA(cudaDeviceSynchronize());
    return profiler_read(c->prof, out, reset);
}

// ------------------------------------------------------------------------------------- single ops
// precision: VT_PREC_BF16 (bf16 operands), VT_PREC_FP32 (FFMA path), VT_PREC_F16 (fp16 operands, the
// format the encoder schedule uses for bounded operands; raw shortcut operands stay bf16)
static int op_fmt(int precision) { return precision == VT_PREC_FP32 ? FMT_F32 : (precision == VT_PREC_F16 ? FMT_F16 : FMT_BF16); }

int vt_op_conv2d(vt_ctx* c, const float* x, const float* w, const float* bias, const float* residual,
                 const float* sc_x, const float* sc_w, int N, int Cin, int H, int W, int Cout, int ksize, int stride,
                 int Cs, int precision, float* out, double* stats, void* stream) {
    VT_TRY(set_device(c));
    VT_CHECK(x && w && out, "null pointers");
    VT_CHECK((sc_x == nullptr) == (sc_w == nullptr), "shortcut operand and weight go together");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int fp32 = precision == VT_PREC_FP32;
    const int fmt = op_fmt(precision);
    const int raw = fp32 ? FMT_F32 : (c->raw_f16 ? FMT_F16 : FMT_BF16);   // residual / shortcut operand storage
    const size_t es = fp32 ? 4 : 2;
    const int Ho = stride == 1 ? H : H / 2, Wo = stride == 1 ? W : W / 2;
    if (!sc_x) Cs = 0;
    const int Ktot = ksize * ksize * Cin + Cs;
    const size_t b_x = align_up(static_cast<size_t>(N) * H * W * Cin * es, 256);
    const size_t b_o = align_up(static_cast<size_t>(N) * Ho * Wo * Cout * 4, 256);
    const size_t b_r = align_up(static_cast<size_t>(N) * Ho * Wo * Cout * es, 256);
    const size_t b_s = align_up(static_cast<size_t>(N) * Ho * Wo * std::max(Cs, 1) * es, 256);
    const size_t b_w = align_up(static_cast<size_t>(Cout) * Ktot * es, 256);
    const size_t b_st = align_up(stats_scratch_bytes(N, Ho, Wo), 256);
    VT_TRY(c->opws.ensure(b_x + b_o + b_r + b_s + b_w + b_st));
    char* p = static_cast<char*>(c->opws.p);
    void* dx = p; void* dout = p + b_x; void* dres = p + b_x + b_o; void* dsc = p + b_x + b_o + b_r;
    void* dw = p + b_x + b_o + b_r + b_s;
    StatsScratch ws;
    ws.part = reinterpret_cast<float*>(p + b_x + b_o + b_r + b_s + b_w); ws.bytes = b_st;
    VT_TRY(launch_nchw_to_nhwc(x, dx, fmt, N, Cin, 1LL * H * W, s));
    if (residual) VT_TRY(launch_nchw_to_nhwc(residual, dres, raw, N, Cout, 1LL * Ho * Wo, s));
    if (sc_x) VT_TRY(launch_nchw_to_nhwc(sc_x, dsc, raw, N, Cs, 1LL * Ho * Wo, s));
    VT_CUDA(cudaMemsetAsync(dw, 0, b_w, s));
    if (fmt == FMT_F32) {
        pack_weight_kernel<FMT_F32><<<256, 256, 0, s>>>(w, dw, Cout, Cin, ksize, Ktot, 0);
        if (sc_w) pack_weight_kernel<FMT_F32><<<256, 256, 0, s>>>(sc_w, dw, Cout, Cs, 1, Ktot, ksize * ksize * Cin);
    } else {
        if (fmt == FMT_F16) pack_weight_kernel<FMT_F16><<<256, 256, 0, s>>>(w, dw, Cout, Cin, ksize, Ktot, 0);
        else pack_weight_kernel<FMT_BF16><<<256, 256, 0, s>>>(w, dw, Cout, Cin, ksize, Ktot, 0);
        if (sc_w) {
            if (c->raw_f16) pack_weight_kernel<FMT_F16><<<256, 256, 0, s>>>(sc_w, dw, Cout, Cs, 1, Ktot, ksize * ksize * Cin);
            else pack_weight_kernel<FMT_BF16><<<256, 256, 0, s>>>(sc_w, dw, Cout, Cs, 1, Ktot, ksize * ksize * Cin);
        }
    }
    VT_CUDA(cudaGetLastError());
    ConvOp op;
    op.raw_f16 = !fp32 && c->raw_f16;
    op.in = dx; op.in_f16 = fmt == FMT_F16; op.N = N; op.Hin = H; op.Win = W; op.Cin = Cin; op.ksize = ksize;
    op.stride = stride; op.w = dw;
    op.Cout = Cout; op.sc_in = sc_x ? dsc : nullptr; op.Cs = Cs; op.bias = bias; op.residual = residual ? dres : nullptr;
    op.out = dout; op.out_fmt = FMT_F32;
    if (fp32) {
        VT_TRY(launch_conv_fp32(op, s, c->prof));
        if (stats) VT_TRY(launch_gn_stats(dout, 1, stats, N, 1LL * Ho * Wo, Cout, 32, s, c->prof));
    } else {
        op.stats = stats; op.stats_ws = ws;
        VT_TRY(launch_conv(op, s, c->prof));
    }
    return launch_nhwc_to_nchw(dout, FMT_F32, out, N, Cout, 1LL * Ho * Wo, s);
}

int vt_op_conv3_fused(vt_ctx* c, const float* x, const float* gamma, const float* beta, const float* w,
                      const float* bias, const float* residual, const float* sc_x, const float* sc_w, int N, int Cin,
                      int H, int W, int Cout, int Cs, float eps, int silu, float* out, double* stats, void* stream) {
    VT_TRY(set_device(c));
    VT_CHECK(x && gamma && beta && w && out, "null pointers");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const long long HW = 1LL * H * W;
    const size_t b_x = align_up(static_cast<size_t>(N) * HW * Cin * 2, 1024);
    const size_t b_o = align_up(static_cast<size_t>(N) * HW * Cout * 4, 1024);
    const size_t b_r = align_up(static_cast<size_t>(N) * HW * Cout * 2, 1024);
    VT_CHECK((sc_x == nullptr) == (sc_w == nullptr), "shortcut operand and weight go together");
    if (!sc_x) Cs = 0;
    const int Ktot = 9 * Cin + Cs;
    const size_t b_w = align_up(static_cast<size_t>(Cout) * Ktot * 2, 1024);
    const size_t b_s = 1024 + static_cast<size_t>(N) * 64 * sizeof(double);
    const size_t b_c = align_up(static_cast<size_t>(N) * HW * std::max(Cs, 1) * 2, 1024);
    const size_t b_st = align_up(stats_scratch_bytes(N, H, W), 1024);
    VT_TRY(c->opws.ensure(b_x + b_o + b_r + b_w + b_s + b_c + b_st));
    char* p = static_cast<char*>(c->opws.p);
    void* dx = p; void* dout = p + b_x; void* dres = p + b_x + b_o; void* dw = p + b_x + b_o + b_r;
    double* st_in = reinterpret_cast<double*>(p + b_x + b_o + b_r + b_w);
    void* dsc = p + b_x + b_o + b_r + b_w + b_s;
    const int raw = c->raw_f16 ? FMT_F16 : FMT_BF16;    // storage of the raw input, the residual and the shortcut operand
    VT_TRY(launch_nchw_to_nhwc(x, dx, raw, N, Cin, HW, s));
    if (residual) VT_TRY(launch_nchw_to_nhwc(residual, dres, raw, N, Cout, HW, s));
    if (sc_x) VT_TRY(launch_nchw_to_nhwc(sc_x, dsc, raw, N, Cs, HW, s));
    pack_weight_kernel<FMT_F16><<<256, 256, 0, s>>>(w, dw, Cout, Cin, 3, Ktot, 0);
    if (sc_w) {
        if (c->raw_f16) pack_weight_kernel<FMT_F16><<<256, 256, 0, s>>>(sc_w, dw, Cout, Cs, 1, Ktot, 9 * Cin);
        else pack_weight_kernel<FMT_BF16><<<256, 256, 0, s>>>(sc_w, dw, Cout, Cs, 1, Ktot, 9 * Cin);
    }
    VT_CUDA(cudaGetLastError());
    VT_TRY(launch_gn_stats(dx, raw, st_in, N, HW, Cin, 32, s, c->prof));
    Conv3FusedOp op;
    op.in = dx; op.N = N; op.H = H; op.W = W; op.Cin = Cin; op.Cout = Cout; op.gn_stats = st_in; op.gamma = gamma;
    op.raw_f16 = c->raw_f16;
    op.beta = beta; op.eps = eps; op.silu = silu; op.w = dw; op.bias = bias; op.residual = residual ? dres : nullptr;
    op.out = dout; op.out_fmt = FMT_F32; op.stats = stats; op.sc_in = sc_x ? dsc : nullptr; op.Cs = Cs;
    op.stats_ws.part = reinterpret_cast<float*>(p + b_x + b_o + b_r + b_w + b_s + b_c); op.stats_ws.bytes = b_st;
    VT_TRY(launch_conv3_fused(op, s, c->prof));
    return launch_nhwc_to_nchw(dout, FMT_F32, out, N, Cout, HW, s);
}

int vt_op_flash_attention(vt_ctx* c, const float* qk, const float* vt, const float* bias_v, int n, int tokens,
                          float scale, float* out, void* stream) {
    VT_TRY(set_device(c));
    VT_CHECK(qk && vt && out, "null pointers");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const size_t nqk = static_cast<size_t>(n) * tokens * 1024, nv = static_cast<size_t>(n) * 512 * tokens;
    const size_t b_qk = align_up(nqk * 2, 1024), b_v = align_up(nv * 2, 1024), b_o = align_up(nv * 2, 1024);
    VT_TRY(c->opws.ensure(b_qk + b_v + b_o));
    char* p = static_cast<char*>(c->opws.p);
    VT_TRY(launch_cast_f32_16(qk, p, FMT_F16, static_cast<long long>(nqk), s));
    VT_TRY(launch_cast_f32_16(vt, p + b_qk, FMT_F16, static_cast<long long>(nv), s));
    FlashOp f;
    f.qk = p; f.vt = p + b_qk; f.bias_v = bias_v; f.out = p + b_qk + b_v; f.n = n; f.tokens = tokens; f.C = 512; f.scale = scale;
    VT_TRY(launch_flash_attention(f, s, c->prof));
    // widen fp16 [n*tokens][512] -> fp32 (layout kernel with one "pixel" per row)
    VT_CHECK(static_cast<long long>(n) * tokens < 65536, "op entry point: n * tokens must be below 65536");
    return launch_nhwc_to_nchw(p + b_qk + b_v, FMT_F16, out, n * tokens, 512, 1, s);
}

int vt_op_gemm_nt(vt_ctx* c, const float* A, const float* B, const float* bias, int batch, int M, int N, int K,
                  int b_batched, float alpha, int precision, float* out, void* stream) {
    VT_TRY(set_device(c));
    VT_CHECK(A && B && out, "null pointers");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    GemmOp g;
    g.batch = batch; g.M = M; g.N = N; g.K = K; g.a_batched = 1; g.b_batched = b_batched; g.bias = bias;
    g.alpha = alpha; g.out = out; g.out_fmt = FMT_F32;
    if (precision == VT_PREC_FP32) {
        g.A = A; g.B = B;
        return launch_gemm_fp32(g, s, c->prof);
    }
    const int fmt = op_fmt(precision);
    const size_t na = static_cast<size_t>(batch) * M * K, nbb = static_cast<size_t>(b_batched ? batch : 1) * N * K;
    VT_TRY(c->opws.ensure(align_up(na * 2, 256) + nbb * 2));
    void* da = c->opws.p;
    void* db = static_cast<char*>(c->opws.p) + align_up(na * 2, 256);
    VT_TRY(launch_cast_f32_16(A, da, fmt, static_cast<long long>(na), s));
    VT_TRY(launch_cast_f32_16(B, db, fmt, static_cast<long long>(nbb), s));
    g.A = da; g.B = db; g.ab_f16 = fmt == FMT_F16;
    return launch_gemm(g, s, c->prof);
}

int vt_op_group_norm(vt_ctx* c, const float* x, const float* gamma, const float* beta, int N, int C, int H, int W,
                     int groups, float eps, int silu, int precision, float* out, void* stream) {
    VT_TRY(set_device(c));
    VT_CHECK(x && gamma && beta && out, "null pointers");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int fp32 = precision == VT_PREC_FP32;
    const int ofmt = op_fmt(precision);          // output: bf16 / fp32 / fp16
    const int ifmt = fp32 ? FMT_F32 : ((ofmt == FMT_F16 && c->raw_f16) ? FMT_F16 : FMT_BF16);  // raw input storage
    const size_t es = fp32 ? 4 : 2;
    const long long HW = 1LL * H * W;
    const size_t b_x = align_up(static_cast<size_t>(N) * HW * C * es, 256);
    const size_t b_s = align_up(static_cast<size_t>(N) * groups * 2 * sizeof(double), 256);
    VT_TRY(c->opws.ensure(2 * b_x + b_s));
    char* p = static_cast<char*>(c->opws.p);
    void* dx = p; void* dy = p + b_x; double* st = reinterpret_cast<double*>(p + 2 * b_x);
    VT_TRY(launch_nchw_to_nhwc(x, dx, ifmt, N, C, HW, s));
    VT_CUDA(cudaMemsetAsync(st, 0, b_s, s));
    VT_TRY(launch_gn_stats(dx, ifmt, st, N, HW, C, groups, s, c->prof));
    VT_TRY(launch_gn_apply(dx, ifmt, dy, ofmt, st, gamma, beta, N, HW, C, groups, eps, silu, s, c->prof));
    return launch_nhwc_to_nchw(dy, ofmt, out, N, C, HW, s);
}

int vt_op_softmax_rows(vt_ctx* c, const float* sc, int64_t rows, int cols, int precision, float* out, void* stream) {
    VT_TRY(set_device(c));
    VT_CHECK(sc && out, "null pointers");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (precision == VT_PREC_FP32) return launch_softmax_rows(sc, out, FMT_F32, rows, cols, cols, cols, s, c->prof);
    const int fmt = op_fmt(precision);
    VT_TRY(c->opws.ensure(static_cast<size_t>(rows) * cols * 2));
    VT_TRY(launch_softmax_rows(sc, c->opws.p, fmt, rows, cols, cols, cols, s, c->prof));
    // widen to fp32 through the layout kernel with C = cols, HW = 1 per row
    return launch_nhwc_to_nchw(c->opws.p, fmt, out, static_cast<int>(rows), cols, 1, s);
}

// ------------------------------------------------------------------------------------- backward ops (8f-4)
}  // extern "C"
namespace {
BwdEnv bwd_env(vt_ctx* c, int precision, void* stream) {
    BwdEnv e;
    e.s = static_cast<cudaStream_t>(stream); e.prof = c->prof; e.fp32 = precision == VT_PREC_FP32;
    e.raw_fmt = c->raw_f16 ? FMT_F16 : FMT_BF16;
    return e;
}
}  // namespace
extern "C" {

int vt_op_conv2d_backward(vt_ctx* c, const float* x, const float* w, const float* grad_out, int N, int Cin, int H, int W,
                          int Cout, int ksize, int precision, float* grad_x, float* grad_w, float* grad_b, void* stream) {
    VT_TRY(set_device(c));
    VT_CHECK(x && w && grad_out, "null pointers");
    VT_CHECK(ksize == 1 || ksize == 3, "kernel size 1 or 3");
    VT_CHECK(Cin % 64 == 0 && Cout % 64 == 0, "channels must be multiples of 64");
    const BwdEnv e = bwd_env(c, precision, stream);
    const int gf = e.fp32 ? FMT_F32 : FMT_BF16, xf = e.fp32 ? FMT_F32 : e.raw_fmt;
    const size_t es = e.fp32 ? 4 : 2;
    const long long HW = 1LL * H * W;
    const WgradPlan plan = bwd_wgrad_plan(e, N, H, W, Cout, Cin, ksize);
    void *dX, *dG, *dDx, *wd, *cs, *a16;
    float* part;
    Carver cv;
    cv.want(&dX, N * HW * Cin * es); cv.want(&dG, N * HW * Cout * es); cv.want(&dDx, N * HW * Cin * es);
    cv.want(&a16, e.fp32 ? 0 : bwd_wgrad16_a16_bytes(N, H, W, Cin, 1));
    cv.want(&wd, bwd_dgrad_weight_bytes(e, Cout, Cin, ksize));
    cv.want(&part, e.fp32 ? plan.part_bytes : bwd_wgrad_mn_plan(N, H, W, Cout, Cin, ksize).part_bytes);
    cv.want(&cs, bwd_colsum_scratch_bytes(Cout));
    VT_TRY(cv.bind(c->opws));
    VT_TRY(launch_nchw_to_nhwc(x, dX, xf, N, Cin, HW, e.s));
    VT_TRY(launch_nchw_to_nhwc(grad_out, dG, gf, N, Cout, HW, e.s));
    if (grad_x) {
        VT_TRY(bwd_pack_dgrad_weight(e, w, wd, Cout, Cin, ksize));
        VT_TRY(bwd_conv_dgrad(e, dG, wd, dDx, nullptr, N, H, W, Cout, Cin, ksize));
        VT_TRY(launch_nhwc_to_nchw(dDx, gf, grad_x, N, Cin, HW, e.s));
    }
    if (grad_w) {
        if (e.fp32) VT_TRY(bwd_conv_wgrad(e, plan, dG, dX, part, grad_w, N, H, W, Cout, Cin, ksize, 0));
        else VT_TRY(bwd_conv_wgrad16(e, dG, dX, xf, nullptr, nullptr, nullptr, 0.f, 0, a16, part, grad_w, N, H, W, Cout, Cin, ksize, 1, 0));
    }
    if (grad_b) VT_TRY(bwd_bias_grad(e, dG, N * HW, Cout, grad_b, 0, cs));
    return 0;
}

int vt_op_group_norm_backward(vt_ctx* c, const float* x, const float* gamma, const float* beta, const float* grad_y, int N,
                              int C, int H, int W, float eps, int silu, int precision, float* grad_x, float* grad_gamma,
                              float* grad_beta, void* stream) {
    VT_TRY(set_device(c));
    VT_CHECK(x && gamma && beta && grad_y && grad_x && grad_gamma && grad_beta, "null pointers");
    const BwdEnv e = bwd_env(c, precision, stream);
    const int gf = e.fp32 ? FMT_F32 : FMT_BF16, xf = e.fp32 ? FMT_F32 : e.raw_fmt;
    const size_t es = e.fp32 ? 4 : 2;
    const long long HW = 1LL * H * W;
    void *dX, *dG, *dDx, *sc;
    double* st;
    Carver cv;
    cv.want(&dX, N * HW * C * es); cv.want(&dG, N * HW * C * es); cv.want(&dDx, N * HW * C * es);
    cv.want(&st, static_cast<size_t>(N) * 64 * sizeof(double)); cv.want(&sc, bwd_gn_scratch_bytes(N, HW, C));
    VT_TRY(cv.bind(c->opws));
    VT_TRY(launch_nchw_to_nhwc(x, dX, xf, N, C, HW, e.s));
    VT_TRY(launch_nchw_to_nhwc(grad_y, dG, gf, N, C, HW, e.s));
    VT_TRY(launch_gn_stats(dX, xf, st, N, HW, C, 32, e.s, c->prof));
    VT_TRY(bwd_group_norm(e, dX, dG, st, gamma, beta, nullptr, dDx, grad_gamma, grad_beta, N, HW, C, eps, silu, 0, sc));
    return launch_nhwc_to_nchw(dDx, gf, grad_x, N, C, HW, e.s);
}

int vt_op_resnet_block_backward(vt_ctx* c, const float* x, const vt_resnet_block_params* pr, const float* grad_out, int N,
                                int Cin, int Cout, int H, int W, int precision, const vt_resnet_block_grads* g, void* stream) {
    VT_TRY(set_device(c));
    VT_CHECK(x && pr && grad_out && g, "null pointers");
    VT_CHECK(pr->norm1_w && pr->norm1_b && pr->conv1_w && pr->conv1_b && pr->norm2_w && pr->norm2_b && pr->conv2_w && pr->conv2_b,
             "missing block parameters");
    VT_CHECK(g->x && g->norm1_w && g->norm1_b && g->conv1_w && g->conv1_b && g->norm2_w && g->norm2_b && g->conv2_w && g->conv2_b,
             "missing gradient outputs");
    const bool sc = Cin != Cout;
    VT_CHECK(!sc || (pr->sc_w && pr->sc_b && g->sc_w && g->sc_b), "a channel-changing block needs its 1x1 shortcut");
    VT_CHECK(Cin % 128 == 0 && Cout % 128 == 0, "channels must be multiples of 128");
    const BwdEnv e = bwd_env(c, precision, stream);
    const int gf = e.fp32 ? FMT_F32 : FMT_BF16, xf = e.fp32 ? FMT_F32 : e.raw_fmt, of = e.fp32 ? FMT_F32 : FMT_F16;
    const size_t es = e.fp32 ? 4 : 2;
    const long long HW = 1LL * H * W;
    const float eps = 1e-6f;
    const int Cmax = std::max(Cin, Cout);
    WgradPlan plan = bwd_wgrad_plan(e, N, H, W, Cout, Cmax, 3);    // fp32 mode: FFMA split-K plan
    plan.part_bytes = align_up(static_cast<size_t>(plan.batches) * Cout * 9 * Cmax * sizeof(float), 256);
    if (!e.fp32) plan.part_bytes = bwd_wgrad_mn_plan(N, H, W, Cout, Cmax, 3).part_bytes;   // >= the 1x1 shortcut's and conv1's
    void *X, *H1, *T, *dOut, *dA, *dH, *dX, *dSc, *w1, *wd1, *wd2, *wds, *a16, *gsc, *cs;
    double *st_x, *st_h;
    float* part;
    Carver cv;
    cv.want(&X, N * HW * Cin * es); cv.want(&H1, N * HW * Cout * es); cv.want(&T, N * HW * Cmax * es);
    cv.want(&dOut, N * HW * Cout * es); cv.want(&dA, N * HW * Cmax * es); cv.want(&dH, N * HW * Cout * es);
    cv.want(&dX, N * HW * Cin * es); cv.want(&dSc, sc ? N * HW * Cin * es : 0);
    cv.want(&w1, static_cast<size_t>(Cout) * 9 * Cin * es);
    cv.want(&wd1, bwd_dgrad_weight_bytes(e, Cout, Cin, 3)); cv.want(&wd2, bwd_dgrad_weight_bytes(e, Cout, Cout, 3));
    cv.want(&wds, sc ? bwd_dgrad_weight_bytes(e, Cout, Cin, 1) : 0);
    cv.want(&a16, e.fp32 ? 0 : bwd_wgrad16_a16_bytes(N, H, W, Cmax, 1));
    cv.want(&part, std::max(plan.part_bytes, bwd_wgrad_mn_plan(N, H, W, Cout, Cin, 3).part_bytes));
    cv.want(&gsc, bwd_gn_scratch_bytes(N, HW, Cmax)); cv.want(&cs, bwd_colsum_scratch_bytes(Cout));
    cv.want(&st_x, static_cast<size_t>(N) * 64 * sizeof(double)); cv.want(&st_h, static_cast<size_t>(N) * 64 * sizeof(double));
    VT_TRY(cv.bind(c->opws));

    // ---- forward, first half: h = conv1(silu(norm1(x))) + b1 (stored in the raw format), statistics of x and h
    VT_TRY(launch_nchw_to_nhwc(x, X, xf, N, Cin, HW, e.s));
    VT_TRY(launch_nchw_to_nhwc(grad_out, dOut, gf, N, Cout, HW, e.s));
    VT_TRY(launch_gn_stats(X, xf, st_x, N, HW, Cin, 32, e.s, c->prof));
    VT_TRY(launch_gn_apply(X, xf, T, of, st_x, pr->norm1_w, pr->norm1_b, N, HW, Cin, 32, eps, 1, e.s, c->prof));
    if (e.fp32) pack_weight_kernel<FMT_F32><<<256, 256, 0, e.s>>>(pr->conv1_w, w1, Cout, Cin, 3, 9 * Cin, 0);
    else pack_weight_kernel<FMT_F16><<<256, 256, 0, e.s>>>(pr->conv1_w, w1, Cout, Cin, 3, 9 * Cin, 0);
    VT_CUDA(cudaGetLastError());
    {
        ConvOp op;
        op.in = T; op.in_f16 = 1; op.raw_f16 = c->raw_f16; op.N = N; op.Hin = H; op.Win = W; op.Cin = Cin; op.ksize = 3;
        op.stride = 1; op.w = w1; op.Cout = Cout; op.bias = pr->conv1_b; op.out = H1; op.out_fmt = xf;
        VT_TRY(e.fp32 ? launch_conv_fp32(op, e.s, c->prof) : launch_conv(op, e.s, c->prof));
    }
    VT_TRY(launch_gn_stats(H1, xf, st_h, N, HW, Cout, 32, e.s, c->prof));

    // ---- conv2 (+ shortcut): weight / bias gradients, data gradient
    VT_TRY(bwd_bias_grad(e, dOut, N * HW, Cout, g->conv2_b, 0, cs));
    if (sc) VT_TRY(bwd_bias_grad(e, dOut, N * HW, Cout, g->sc_b, 0, cs));
    WgradPlan p1 = plan; p1.taps = 1;
    if (e.fp32) {
        VT_TRY(launch_gn_apply(H1, xf, T, of, st_h, pr->norm2_w, pr->norm2_b, N, HW, Cout, 32, eps, 1, e.s, c->prof));
        VT_TRY(bwd_conv_wgrad(e, plan, dOut, T, part, g->conv2_w, N, H, W, Cout, Cout, 3, 0));
        if (sc) VT_TRY(bwd_conv_wgrad(e, p1, dOut, X, part, g->sc_w, N, H, W, Cout, Cin, 1, 0));
    } else {
        VT_TRY(bwd_conv_wgrad16(e, dOut, H1, xf, st_h, pr->norm2_w, pr->norm2_b, eps, 1, a16, part, g->conv2_w, N, H, W, Cout, Cout, 3, 1, 0));
        if (sc) VT_TRY(bwd_conv_wgrad16(e, dOut, X, xf, nullptr, nullptr, nullptr, 0.f, 0, a16, part, g->sc_w, N, H, W, Cout, Cin, 1, 1, 0));
    }
    VT_TRY(bwd_pack_dgrad_weight(e, pr->conv2_w, wd2, Cout, Cout, 3));
    VT_TRY(bwd_conv_dgrad(e, dOut, wd2, dA, nullptr, N, H, W, Cout, Cout, 3));
    // ---- norm2 + SiLU
    VT_TRY(bwd_group_norm(e, H1, dA, st_h, pr->norm2_w, pr->norm2_b, nullptr, dH, g->norm2_w, g->norm2_b, N, HW, Cout, eps, 1, 0, gsc));
    // ---- conv1
    VT_TRY(bwd_bias_grad(e, dH, N * HW, Cout, g->conv1_b, 0, cs));
    if (e.fp32) {
        VT_TRY(launch_gn_apply(X, xf, T, of, st_x, pr->norm1_w, pr->norm1_b, N, HW, Cin, 32, eps, 1, e.s, c->prof));
        VT_TRY(bwd_conv_wgrad(e, plan, dH, T, part, g->conv1_w, N, H, W, Cout, Cin, 3, 0));
    } else {
        VT_TRY(bwd_conv_wgrad16(e, dH, X, xf, st_x, pr->norm1_w, pr->norm1_b, eps, 1, a16, part, g->conv1_w, N, H, W, Cout, Cin, 3, 1, 0));
    }
    VT_TRY(bwd_pack_dgrad_weight(e, pr->conv1_w, wd1, Cout, Cin, 3));
    VT_TRY(bwd_conv_dgrad(e, dH, wd1, dA, nullptr, N, H, W, Cout, Cin, 3));
    // ---- shortcut branch gradient, added inside norm1's apply pass
    const void* add = dOut;
    if (sc) {
        VT_TRY(bwd_pack_dgrad_weight(e, pr->sc_w, wds, Cout, Cin, 1));
        VT_TRY(bwd_conv_dgrad(e, dOut, wds, dSc, nullptr, N, H, W, Cout, Cin, 1));
        add = dSc;
    }
    VT_TRY(bwd_group_norm(e, X, dA, st_x, pr->norm1_w, pr->norm1_b, add, dX, g->norm1_w, g->norm1_b, N, HW, Cin, eps, 1, 0, gsc));
    return launch_nhwc_to_nchw(dX, gf, g->x, N, Cin, HW, e.s);
}

// ------------------------------------------------------------------------------------- fine-tuning losses (8f-4)
int vt_embed_loss(vt_ctx* c, const vt_embed_loss_args* a) {
    VT_TRY(set_device(c));
    VT_CHECK(a != nullptr, "null arguments");
    VT_TRY(c->optws.ensure(embed_loss_scratch_bytes(a->B, a->D)));
    return launch_embed_loss(*a, c->optws.p, static_cast<cudaStream_t>(a->stream), c->prof);
}
int vt_mse_loss(vt_ctx* c, const float* x, const float* y, int64_t n, float* loss, float* grad_x, void* stream) {
    VT_TRY(set_device(c));
    VT_TRY(c->optws.ensure(148 * 8 * sizeof(float)));
    return launch_mse_loss(x, y, n, loss, grad_x, c->optws.p, static_cast<cudaStream_t>(stream), c->prof);
}
int vt_adaptive_loss_weights(vt_ctx* c, const float* log_w, const float* losses, int n, float temperature, float* total,
                             float* weights, float* grad_log_w, void* stream) {
    VT_TRY(set_device(c));
    return launch_adaptive_weights(log_w, losses, n, temperature, total, weights, grad_log_w, static_cast<cudaStream_t>(stream));
}

// ------------------------------------------------------------------------------------- encoder training (8f-4)
int vt_encoder_train_forward(vt_ctx* c, const vt_encode_args* a, int slot) {
    VT_TRY(set_device(c));
    VT_CHECK(a != nullptr, "null arguments");
    VT_CHECK(slot >= 0 && slot < VT_MAX_TAPES, "tape slot out of range");
    VT_CHECK(c->enc_ready, "encoder parameters not finalised (vt_encoder_finalize)");
    VT_CHECK(a->images != nullptr, "null image pointer");
    VT_CHECK(a->batch > 0 && a->height > 0 && a->width > 0, "batch and image size must be positive");
    VT_CHECK(a->in_fmt == VT_IN_F32_NCHW || a->in_fmt == VT_IN_U8_NHWC, "unknown input format");
    return run_encoder_train_forward(c, a, slot);
}
int vt_encoder_grad_bind(vt_ctx* c, const char* name, float* grad) {
    VT_CHECK(c != nullptr && name != nullptr, "null arguments");
    VT_CHECK(c->eparams.count(name) != 0, std::string("unknown encoder parameter ") + name);
    if (grad) c->egrads[name] = grad;
    else c->egrads.erase(name);
    return 0;
}
int vt_encoder_tape_release(vt_ctx* c, int slot) {
    VT_TRY(set_device(c));
    VT_CHECK(slot >= 0 && slot < VT_MAX_TAPES, "tape slot out of range");
    if (c->tapes[slot]) {
        VT_CUDA(cudaDeviceSynchronize());
        c->tapes[slot]->release();
        delete c->tapes[slot];
        c->tapes[slot] = nullptr;
    }
    return 0;
}
int vt_encoder_backward(vt_ctx* c, const vt_encoder_backward_args* a) {
    VT_TRY(set_device(c));
    VT_CHECK(a != nullptr, "null arguments");
    VT_CHECK(a->grad_mean != nullptr || a->grad_logvar != nullptr, "no output gradient");
    return run_encoder_backward(c, a);
}

int vt_decoder_train_forward(vt_ctx* c, const vt_decode_args* a, int slot) {
    VT_TRY(set_device(c));
    VT_CHECK(a != nullptr, "null arguments");
    VT_CHECK(slot >= 0 && slot < VT_MAX_TAPES, "tape slot out of range");
    VT_CHECK(c->dec_ready, "decoder parameters not finalised (vt_decoder_finalize)");
    VT_CHECK(a->latent != nullptr && a->image != nullptr, "null latent / image pointer");
    VT_CHECK(a->batch > 0 && a->lat_h > 0 && a->lat_w > 0, "batch and latent size must be positive");
    return run_decoder_train_forward(c, a, slot);
}
int vt_decoder_grad_bind(vt_ctx* c, const char* name, float* grad) {
    VT_CHECK(c != nullptr && name != nullptr, "null arguments");
    VT_CHECK(c->dparams.count(name) != 0, std::string("unknown decoder parameter ") + name);
    if (grad) c->dgrads[name] = grad;
    else c->dgrads.erase(name);
    return 0;
}
int vt_decoder_backward(vt_ctx* c, const vt_decoder_backward_args* a) {
    VT_TRY(set_device(c));
    VT_CHECK(a != nullptr && a->grad_image != nullptr, "null arguments");
    return run_decoder_backward(c, a);
}
int vt_decoder_tape_release(vt_ctx* c, int slot) {
    VT_TRY(set_device(c));
    VT_CHECK(slot >= 0 && slot < VT_MAX_TAPES, "tape slot out of range");
    if (c->dtapes[slot]) {
        VT_CUDA(cudaDeviceSynchronize());
        c->dtapes[slot]->release();
        delete c->dtapes[slot];
        c->dtapes[slot] = nullptr;
    }
    return 0;
}

}  // extern "C"
