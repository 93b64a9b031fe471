// Flash-style fused attention for the FLUX VAE mid block: one head, head_dim = 512, N = H*W/64 tokens.
//
//   O = softmax(scale * Q K^T) V + b_v          (fp16 operands, fp32 accumulation in TMEM)
//
// One CTA = 128 queries x one half (256 columns) of d_v, looping over 128-key tiles:
//   S(j)   = Q K_j^T            tcgen05.mma M=128 N=128, K = 512 streamed as 8 chunks of (Q_c, K_c)
//   P(j)   = exp2(c*S - m)      4 softmax warps, one query row per thread (no shuffles), fp16 -> smem
//   O     += P(j) V_j           tcgen05.mma M=128 N=256, K = 128
// TMEM: two S buffers (2 x 128 columns) + O (256 columns) = 512 columns: that is why d_v is split over
// two CTAs (an O tile of 128 x 512 fp32 alone would fill TMEM) -- QK^T is computed twice, PV once.
// The running maximum is only raised when a row's new maximum exceeds it by more than 2^8 (lazy
// rescale: P stays within fp16 range, O is rescaled in TMEM only then).  Scores never touch HBM.
#include "vt_internal.h"
#include "vt_ptx.cuh"

namespace vt {

namespace {

constexpr int FQ = 128;                 // queries per CTA
constexpr int FK = 128;                 // keys per tile
constexpr int FD = 512;                 // head dim
constexpr int FDV = 256;                // d_v columns per CTA
constexpr int QK_STAGE = 2 * 16384;     // Q chunk + K chunk (128 rows x 128 B each)
constexpr int QK_STAGES = 3;
constexpr int V_STAGE = FDV * 128;      // 256 d_v rows x 64 keys x 2 B
constexpr int V_STAGES = 2;
constexpr int P_BYTES = 2 * 16384;      // two 64-key chunks of 128 rows x 128 B
constexpr int FLASH_SMEM = QK_STAGES * QK_STAGE + V_STAGES * V_STAGE + P_BYTES + 1024 + 1024;
constexpr int FLASH_THREADS = 192;      // warp 0 TMA, warp 1 MMA, warps 2..5 softmax / epilogue

__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};\n" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
        "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
        "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}
__device__ __forceinline__ float fast_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__global__ void __launch_bounds__(FLASH_THREADS, 1)
flash_d512_kernel(const __grid_constant__ CUtensorMap tmQK, const __grid_constant__ CUtensorMap tmV,
                  __half* __restrict__ out, const float* __restrict__ bias_v, int tokens, int q_tiles,
                  float scale_log2) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* s_qk = smem;                                   // [QK_STAGES][Q 16 KB | K 16 KB]
    uint8_t* s_v = smem + QK_STAGES * QK_STAGE;             // [V_STAGES][32 KB]
    uint8_t* s_p = s_v + V_STAGES * V_STAGE;                // [2 chunks][16 KB]
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_p + P_BYTES);
    uint64_t* qk_full = bars;            // [3]
    uint64_t* qk_empty = bars + 3;       // [3]
    uint64_t* v_full = bars + 6;         // [2]
    uint64_t* v_empty = bars + 8;        // [2]
    uint64_t* s_full = bars + 10;        // [2]  S(j) accumulated
    uint64_t* s_empty = bars + 12;       // [2]  S buffer read by the softmax warps
    uint64_t* p_full = bars + 14;        // [1]  P(j) written to shared memory
    uint64_t* p_empty = bars + 15;       // [1]  P(j) V_j done: P may be overwritten, O may be rescaled
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 16);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // blockIdx.x = (img * q_tiles + q_tile) * 2 + half
    const int half = blockIdx.x & 1;
    const int qt = (blockIdx.x >> 1) % q_tiles;
    const int img = (blockIdx.x >> 1) / q_tiles;
    const int q0 = qt * FQ;
    const int key_tiles = (tokens + FK - 1) / FK;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmQK);
        tma_prefetch_desc(&tmV);
        for (int i = 0; i < QK_STAGES; ++i) { mbar_init(&qk_full[i], 1); mbar_init(&qk_empty[i], 1); }
        for (int i = 0; i < V_STAGES; ++i) { mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], 4); }
        mbar_init(p_full, 4);
        mbar_init(p_empty, 1);
        fence_mbar_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_ptr, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    const uint32_t tmem_s = tmem_base;          // columns 0..255: S double buffer
    const uint32_t tmem_o = tmem_base + 256;    // columns 256..511: O

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer
        if (lane == 0) {
            int qs = 0, vs = 0;
            uint32_t qph = 0, vph = 0;
            auto load_s_tile = [&](int j) {      // 8 (Q_c, K_c) chunk pairs of key tile j
                for (int c = 0; c < FD / 64; ++c) {
                    mbar_wait(&qk_empty[qs], qph ^ 1);
                    uint8_t* st = s_qk + qs * QK_STAGE;
                    mbar_arrive_expect_tx(&qk_full[qs], QK_STAGE);
                    tma_load_3d(st, &tmQK, &qk_full[qs], c * 64, q0, img);                    // Q chunk
                    tma_load_3d(st + 16384, &tmQK, &qk_full[qs], FD + c * 64, j * FK, img);  // K chunk
                    if (++qs == QK_STAGES) { qs = 0; qph ^= 1; }
                }
            };
            load_s_tile(0);
            for (int j = 0; j < key_tiles; ++j) {
                if (j + 1 < key_tiles) load_s_tile(j + 1);
                for (int kc = 0; kc < FK / 64; ++kc) {   // V^T chunks of key tile j: [256 d_v rows][64 keys]
                    mbar_wait(&v_empty[vs], vph ^ 1);
                    mbar_arrive_expect_tx(&v_full[vs], V_STAGE);
                    tma_load_3d(s_v + vs * V_STAGE, &tmV, &v_full[vs], j * FK + kc * 64, half * FDV, img);
                    if (++vs == V_STAGES) { vs = 0; vph ^= 1; }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer (warp-uniform loops)
        constexpr uint32_t idesc_s = umma_idesc_16(128, FK, true);     // S: M=128 queries, N=128 keys
        constexpr uint32_t idesc_o = umma_idesc_16(128, FDV, true);    // O: M=128 queries, N=256 d_v
        const uint64_t dq_base = umma_desc_k_sw128(smem_u32(s_qk));
        const uint64_t dk_base = umma_desc_k_sw128(smem_u32(s_qk) + 16384);
        const uint64_t dv_base = umma_desc_k_sw128(smem_u32(s_v));
        const uint64_t dp_base = umma_desc_k_sw128(smem_u32(s_p));
        int qs = 0, vs = 0;
        uint32_t qph = 0, vph = 0;
        auto issue_s = [&](int j) {
            const uint32_t sb = j & 1;
            mbar_wait(&s_empty[sb], ((j >> 1) & 1) ^ 1);
            tc_fence_after();
            for (int c = 0; c < FD / 64; ++c) {
                mbar_wait(&qk_full[qs], qph);
                tc_fence_after();
                const uint64_t so = static_cast<uint64_t>(qs * (QK_STAGE >> 4));
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_bf16_ss(tmem_s + sb * FK, dq_base + so + 2 * k, dk_base + so + 2 * k, idesc_s, (c | k) != 0);
                    umma_commit(&qk_empty[qs]);
                    if (c == FD / 64 - 1) umma_commit(&s_full[sb]);
                }
                __syncwarp();
                if (++qs == QK_STAGES) { qs = 0; qph ^= 1; }
            }
        };
        issue_s(0);
        for (int j = 0; j < key_tiles; ++j) {
            if (j + 1 < key_tiles) issue_s(j + 1);            // overlaps the softmax of tile j
            mbar_wait(p_full, j & 1);
            tc_fence_after();
            for (int kc = 0; kc < FK / 64; ++kc) {
                mbar_wait(&v_full[vs], vph);
                tc_fence_after();
                const uint64_t vo = static_cast<uint64_t>(vs * (V_STAGE >> 4));
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_bf16_ss(tmem_o, dp_base + (kc * (16384 >> 4) + 2 * k), dv_base + vo + 2 * k, idesc_o,
                                     (j | kc | k) != 0);
                    umma_commit(&v_empty[vs]);
                    if (kc == FK / 64 - 1) umma_commit(p_empty);
                }
                __syncwarp();
                if (++vs == V_STAGES) { vs = 0; vph ^= 1; }
            }
        }
    } else {
        // ------------------------------------------------------------ softmax warps (one query row per thread)
        const int q = warp & 3;                           // TMEM lane quadrant
        const int row = q * 32 + lane;                    // query row inside the tile
        const uint32_t lane_sel = static_cast<uint32_t>(q * 32) << 16;
        const uint32_t p_row = smem_u32(s_p) + row * 128;
        float m_used = -INFINITY;                         // maximum the exponentials are taken against (log2 units)
        float l = 0.f;
        for (int j = 0; j < key_tiles; ++j) {
            const uint32_t sb = j & 1;
            mbar_wait(&s_full[sb], (j >> 1) & 1);
            tc_fence_after();
            uint32_t r[4][32];
#pragma unroll
            for (int cchunk = 0; cchunk < 4; ++cchunk) tmem_ld_32x32(tmem_s + sb * FK + cchunk * 32 + lane_sel, r[cchunk]);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_empty[sb]);     // S(j+2) may overwrite this buffer
            // scale to log2 units, mask keys beyond the sequence, row maximum
            const int key0 = j * FK;
            float mx = -INFINITY;
#pragma unroll
            for (int cchunk = 0; cchunk < 4; ++cchunk)
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    float v = __uint_as_float(r[cchunk][i]) * scale_log2;
                    if (key0 + cchunk * 32 + i >= tokens) v = -INFINITY;
                    r[cchunk][i] = __float_as_uint(v);
                    mx = fmaxf(mx, v);
                }
            // lazy rescale: raise the reference maximum only when it is exceeded by more than 2^8
            const bool raise = mx > m_used + 8.0f;
            const float m_new = raise ? mx : m_used;
            const float alpha = (raise && j > 0) ? exp2f(m_used - m_new) : 1.0f;
            const bool any_raise = __any_sync(0xFFFFFFFFu, raise && j > 0);
            if (j > 0) {
                mbar_wait(p_empty, (j - 1) & 1);          // P(j-1) V done: O is stable, P may be overwritten
                tc_fence_after();
            }
            if (any_raise) {
                // rescale this warp's 32 rows of O in TMEM (rows that keep their maximum use alpha = 1)
#pragma unroll 1
                for (int cchunk = 0; cchunk < FDV / 32; ++cchunk) {
                    uint32_t o[32];
                    tmem_ld_32x32(tmem_o + cchunk * 32 + lane_sel, o);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                    tmem_st_32x32(tmem_o + cchunk * 32 + lane_sel, o);
                }
                tmem_st_wait();
                l *= alpha;
            }
            m_used = m_new;
            // P = exp2(s - m), fp16, into the K-major 128B-swizzled A-operand layout
            float lsum = 0.f;
#pragma unroll
            for (int cchunk = 0; cchunk < 4; ++cchunk) {
                uint32_t pk[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const float a = fast_exp2(__uint_as_float(r[cchunk][2 * i]) - m_used);
                    const float b = fast_exp2(__uint_as_float(r[cchunk][2 * i + 1]) - m_used);
                    lsum += a + b;
                    pk[i] = pack_f16x2(a, b);
                }
                // 32 keys = 64 bytes = four 16-byte chunks of this row: keys [cchunk*32, +32)
#pragma unroll
                for (int w4 = 0; w4 < 4; ++w4) {
                    const int chunk16 = cchunk * 4 + w4;                 // 0..15 across the 128 keys
                    const int kc = chunk16 >> 3, pos = (chunk16 & 7) ^ (row & 7);
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(p_row + kc * 16384 + pos * 16),
                                 "r"(pk[4 * w4]), "r"(pk[4 * w4 + 1]), "r"(pk[4 * w4 + 2]), "r"(pk[4 * w4 + 3]) : "memory");
                }
            }
            l += lsum;
            fence_proxy_async_smem();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(p_full);
        }
        // ---- epilogue: O / l + b_v -> fp16 global
        mbar_wait(p_empty, (key_tiles - 1) & 1);
        tc_fence_after();
        const float inv = 1.0f / l;
        const int qrow = q0 + row;
        __half* orow = out + (static_cast<long long>(img) * tokens + qrow) * FD + half * FDV;
#pragma unroll 1
        for (int cchunk = 0; cchunk < FDV / 32; ++cchunk) {
            uint32_t o[32];
            tmem_ld_32x32(tmem_o + cchunk * 32 + lane_sel, o);
            tmem_ld_wait();
            if (qrow < tokens) {
                const float* bv = bias_v + half * FDV + cchunk * 32;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    uint32_t w[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int cidx = i * 8 + e * 2;
                        const float a = __uint_as_float(o[cidx]) * inv + (bias_v ? bv[cidx] : 0.f);
                        const float b = __uint_as_float(o[cidx + 1]) * inv + (bias_v ? bv[cidx + 1] : 0.f);
                        w[e] = pack_f16x2(a, b);
                    }
                    *reinterpret_cast<uint4*>(orow + cchunk * 32 + i * 8) = make_uint4(w[0], w[1], w[2], w[3]);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace

int launch_flash_attention(const FlashOp& op, cudaStream_t stream, Profiler* prof) {
    VT_CHECK(op.C == FD, "fused attention is specialised for head_dim 512");
    VT_CHECK(op.tokens > 0 && op.tokens % 8 == 0 && op.n > 0, "fused attention: token count must be a positive multiple of 8");
    CUtensorMap tq, tv;
    {
        uint64_t dims[3] = {static_cast<uint64_t>(2 * FD), static_cast<uint64_t>(op.tokens), static_cast<uint64_t>(op.n)};
        uint64_t str[2] = {2ull * 2 * FD, 2ull * 2 * FD * op.tokens};
        uint32_t box[3] = {64, 128, 1};
        VT_TRY(make_tmap(&tq, op.qk, 3, dims, str, box));
    }
    {
        uint64_t dims[3] = {static_cast<uint64_t>(op.tokens), static_cast<uint64_t>(FD), static_cast<uint64_t>(op.n)};
        uint64_t str[2] = {2ull * op.tokens, 2ull * op.tokens * FD};
        uint32_t box[3] = {64, static_cast<uint32_t>(FDV), 1};
        VT_TRY(make_tmap(&tv, op.vt, 3, dims, str, box));
    }
    static bool attr_set = false;
    if (!attr_set) {
        VT_CUDA(cudaFuncSetAttribute(flash_d512_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FLASH_SMEM));
        attr_set = true;
    }
    const int q_tiles = (op.tokens + FQ - 1) / FQ;
    const int grid = op.n * q_tiles * 2;
    const double key_t = (op.tokens + FK - 1) / FK * FK;
    const double flops = 2.0 * op.n * q_tiles * FQ * key_t * FD * 3.0;  // QK^T twice (two d_v halves) + PV
    profiler_begin(prof, KC_IGEMM, stream, flops, 2.0 * op.n * op.tokens * FD * 4.0);
    flash_d512_kernel<<<grid, FLASH_THREADS, FLASH_SMEM, stream>>>(tq, tv, static_cast<__half*>(op.out), op.bias_v,
                                                                   op.tokens, q_tiles,
                                                                   op.scale * 1.4426950408889634f);
    profiler_end(prof, KC_IGEMM, stream);
    VT_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace vt
