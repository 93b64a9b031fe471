// Flash-style fused attention for the FLUX VAE mid block: one head, head_dim = 512, N = H*W/64 tokens.
//
//   O = softmax(scale * Q K^T) V + b_v          (fp16 operands, fp32 accumulation in TMEM)
//
// One CTA = 128 queries x one half (256 columns) of d_v, looping over 128-key tiles:
//   S(j)   = Q K_j^T            tcgen05.mma M=128 N=128, K = 512: Q resident in shared memory (8 chunks),
//                               K_j streamed as 8 chunks of 128 keys x 64 dims
//   P(j)   = exp2(c*S - m)      8 softmax warps: a query row is owned by two threads (64 keys each, the row
//                               maximum is exchanged through shared memory); fp16 P is written back into
//                               TENSOR MEMORY over its own S buffer (two values per column)
//   O     += P(j) V_j           tcgen05.mma with the A operand in tensor memory, M=128 N=128 per d_v quarter,
//                               V_j streamed as 4 chunks of 128 d_v rows x 64 keys
// TMEM: two S/P buffers (2 x 128 columns) + O (256 columns) = 512 columns: that is why d_v is split over
// two CTAs (an O tile of 128 x 512 fp32 alone would fill TMEM) -- QK^T is computed twice, PV once.
// CTA PAIRS: two CTAs with neighbouring query tiles (same d_v half) form a cluster and run every MMA as
// one tcgen05.mma.cta_group::2 (M = 256: 128 query rows per CTA): the B operand -- K_j for S, V_j for O --
// is split across the pair, so each CTA streams and reads only HALF of every K / V chunk.
// Shared memory per CTA: Q 128 KB + one ring of twelve 8 KB half-chunks carrying K and V in consumption
// order.  The first version re-streamed Q with every key tile, staged P through shared memory and ran
// single-CTA MMAs: 320 KB of L2->SM traffic and ~700 KB of shared-memory traffic per key tile, which (not
// the tensor pipe) set its speed; now 96 KB and ~320 KB.
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
constexpr int CHUNK = 16384;            // 128 rows x 128 B (64 fp16, 128B-swizzled K-major)
constexpr int Q_BYTES = (FD / 64) * CHUNK;
constexpr int HCHUNK = CHUNK / 2;       // this CTA's half of a K / V chunk: 64 rows x 128 B
constexpr int RING = 12;                // K half-chunks (8 per tile) and V half-chunks (4 per tile) in consumption order
constexpr int XCHG_BYTES = 2 * 2 * FQ * 4;   // [tile parity][column half][row] row maxima / row sums
constexpr int FLASH_SMEM = Q_BYTES + RING * HCHUNK + 256 + XCHG_BYTES;   // no alignment slack: see the kernel
static_assert(FLASH_SMEM <= 227 * 1024, "shared memory budget");
constexpr int SM_WARPS = 8;              // softmax warps: two per TMEM lane quadrant, each owns 64 of a tile's 128 keys
constexpr int FLASH_THREADS = 64 + 32 * SM_WARPS;  // warp 0 TMA, warp 1 MMA, warps 2..9 softmax / epilogue

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

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(FLASH_THREADS, 1)
flash_d512_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                  const __grid_constant__ CUtensorMap tmV, __half* __restrict__ out,
                  const float* __restrict__ bias_v, int tokens, int q_pairs, float scale_log2) {
    // Q + ring fill the 227 KB to within 1 KB, so there is no room for an alignment pad: the dynamic window is
    // declared 1024-byte aligned (what the 128B-swizzled tiles need) and the kernel refuses to run otherwise
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw;
    if ((smem_u32(smem) & 1023u) != 0u) __trap();
    uint8_t* s_q = smem;                                    // [8 chunks][128 queries x 64 dims]
    uint8_t* s_ring = smem + Q_BYTES;                       // [RING][8 KB]
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_ring + RING * HCHUNK);
    uint64_t* q_full = bars;                 // [1]     (leader's is waited on)
    uint64_t* ring_full = bars + 1;          // [RING]  (leader's is waited on: both CTAs' TMA bytes land there)
    uint64_t* ring_empty = ring_full + RING; // [RING]  tcgen05.commit multicast to both CTAs
    uint64_t* s_full = ring_empty + RING;    // [2]  S(j) accumulated (multicast)
    uint64_t* p_full = s_full + 2;           // [2]  P(j) written to tensor memory in BOTH CTAs (leader's)
    uint64_t* o_done = p_full + 2;           // [1]  P(j) V_j accumulated: O may be rescaled / read (multicast)
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(o_done + 1);
    float* xchg = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);   // [2][2][FQ]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // blockIdx.x = (((img * q_pairs + q_pair) * 2 + half) * 2 + rank): the cluster is the two query tiles of a pair
    const uint32_t rank = cluster_ctarank();
    const int half = (blockIdx.x >> 1) & 1;
    const int qp = (blockIdx.x >> 2) % q_pairs;
    const int img = (blockIdx.x >> 2) / q_pairs;
    const int q0 = (qp * 2 + static_cast<int>(rank)) * FQ;   // may lie beyond the sequence (odd tile count): TMA zero-fills, stores are masked
    const int key_tiles = (tokens + FK - 1) / FK;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmQ);
        tma_prefetch_desc(&tmK);
        tma_prefetch_desc(&tmV);
        mbar_init(q_full, 1);
        for (int i = 0; i < RING; ++i) { mbar_init(&ring_full[i], 1); mbar_init(&ring_empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&s_full[i], 1); mbar_init(&p_full[i], 2 * SM_WARPS); }
        mbar_init(o_done, 1);
        fence_mbar_init();
    }
    if (warp == 1) {
        tmem_alloc_pair(tmem_ptr, 512);
        tmem_relinquish_pair();
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();      // the peer's barriers are initialised before any remote arrive / TMA signal
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    const uint32_t tmem_s = tmem_base;          // columns 0..255: S double buffer (P aliases the first 64 columns of each)
    const uint32_t tmem_o = tmem_base + 256;    // columns 256..511: O

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer (both CTAs, own halves)
        if (lane == 0) {
            if (rank == 0) mbar_arrive_expect_tx(q_full, 2 * Q_BYTES);
            for (int c = 0; c < FD / 64; ++c) tma_load_3d_pair(s_q + c * CHUNK, &tmQ, q_full, c * 64, q0, img);
            int rs = 0;
            uint32_t rph = 0;
            auto put = [&](const CUtensorMap* map, int c0, int c1) {
                mbar_wait(&ring_empty[rs], rph ^ 1);
                if (rank == 0) mbar_arrive_expect_tx(&ring_full[rs], 2 * HCHUNK);
                tma_load_3d_pair(s_ring + rs * HCHUNK, map, &ring_full[rs], c0, c1 + static_cast<int>(rank) * 64, img);
                if (++rs == RING) { rs = 0; rph ^= 1; }
            };
            auto load_k = [&](int j) {           // this CTA's 64 keys of key tile j, 8 chunks of 64 dims
                for (int c = 0; c < FD / 64; ++c) put(&tmK, FD + c * 64, j * FK);
            };
            load_k(0);
            for (int j = 0; j < key_tiles; ++j) {
                if (j + 1 < key_tiles) load_k(j + 1);
                for (int kc = 0; kc < FK / 64; ++kc)     // V^T: this CTA's 64 of the 128 d_v rows x 64 keys
                    for (int h = 0; h < FDV / 128; ++h) put(&tmV, j * FK + kc * 64, half * FDV + h * 128);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer: leader CTA only
        if (rank == 0) {
            constexpr uint32_t idesc = umma_idesc_16(256, 128, true);   // pair MMA: 2 x 128 queries; 128 keys / 128 d_v
            const uint64_t dq_base = umma_desc_k_sw128(smem_u32(s_q));
            const uint64_t dr_base = umma_desc_k_sw128(smem_u32(s_ring));
            int rs = 0;
            uint32_t rph = 0;
            mbar_wait(q_full, 0);
            auto issue_s = [&](int j) {
                // S(j) overwrites the buffer that held P(j-2): P(j-2) V was issued earlier and the tensor pipe
                // executes in issue order, so no barrier is needed for that hazard
                const uint32_t sb = j & 1;
                for (int c = 0; c < FD / 64; ++c) {
                    mbar_wait(&ring_full[rs], rph);
                    tc_fence_after();
                    const uint64_t ro = static_cast<uint64_t>(rs * (HCHUNK >> 4));
                    const uint64_t qo = static_cast<uint64_t>(c * (CHUNK >> 4));
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma_f16_ss_pair(tmem_s + sb * FK, dq_base + qo + 2 * k, dr_base + ro + 2 * k, idesc, (c | k) != 0);
                        umma_commit_pair(&ring_empty[rs]);
                        if (c == FD / 64 - 1) umma_commit_pair(&s_full[sb]);
                    }
                    __syncwarp();
                    if (++rs == RING) { rs = 0; rph ^= 1; }
                }
            };
            issue_s(0);
            for (int j = 0; j < key_tiles; ++j) {
                if (j + 1 < key_tiles) issue_s(j + 1);            // overlaps the softmax of tile j
                const uint32_t sb = j & 1;
                mbar_wait(&p_full[sb], (j >> 1) & 1);
                tc_fence_after();
                for (int kc = 0; kc < FK / 64; ++kc)
                    for (int h = 0; h < FDV / 128; ++h) {
                        mbar_wait(&ring_full[rs], rph);
                        tc_fence_after();
                        const uint64_t ro = static_cast<uint64_t>(rs * (HCHUNK >> 4));
                        if (elect_one()) {
#pragma unroll
                            for (int k = 0; k < 4; ++k)   // A = P(j) keys [kc*64 + 16k, +16): 8 tensor-memory columns
                                umma_f16_ts_pair(tmem_o + h * 128, tmem_s + sb * FK + kc * 32 + k * 8, dr_base + ro + 2 * k,
                                                 idesc, (j | kc | k) != 0);
                            umma_commit_pair(&ring_empty[rs]);
                            if (kc == FK / 64 - 1 && h == FDV / 128 - 1) umma_commit_pair(o_done);
                        }
                        __syncwarp();
                        if (++rs == RING) { rs = 0; rph ^= 1; }
                    }
            }
        }
    } else {
        // ------------------------------------------------------------ softmax warps
        // thread = (query row, column half hs): 64 keys of every tile, 128 of the 256 O columns
        const int q = warp & 3;                           // TMEM lane quadrant (hardware: warp id % 4)
        const int hs = (warp - 2) >> 2;                   // 0: warps 2..5, 1: warps 6..9
        const int row = q * 32 + lane;                    // query row inside the tile
        const uint32_t lane_sel = static_cast<uint32_t>(q * 32) << 16;
        const int pair_bar = 1 + q;                       // named barrier of the two warps that share a quadrant
        float m_used = -INFINITY;                         // maximum the exponentials are taken against (log2 units)
        float l = 0.f;                                    // this thread's part of the row sum
        for (int j = 0; j < key_tiles; ++j) {
            const uint32_t sb = j & 1;
            mbar_wait(&s_full[sb], (j >> 1) & 1);
            tc_fence_after();
            uint32_t r[2][32];
#pragma unroll
            for (int cchunk = 0; cchunk < 2; ++cchunk)
                tmem_ld_32x32(tmem_s + sb * FK + hs * 64 + cchunk * 32 + lane_sel, r[cchunk]);
            tmem_ld_wait();
            // row maximum of the raw scores (four independent chains); the softmax scale is positive, so it is
            // applied to the maximum here and folded into one FFMA per element below.  Keys beyond the sequence
            // exist only in a ragged last tile: masked to -inf there, no per-element test anywhere else.
            const int key0 = j * FK + hs * 64;
            if (key0 + 64 > tokens) {
#pragma unroll
                for (int cchunk = 0; cchunk < 2; ++cchunk)
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        if (key0 + cchunk * 32 + i >= tokens) r[cchunk][i] = __float_as_uint(-INFINITY);
            }
            float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
            for (int cchunk = 0; cchunk < 2; ++cchunk)
#pragma unroll
                for (int i = 0; i < 32; ++i) mx4[i & 3] = fmaxf(mx4[i & 3], __uint_as_float(r[cchunk][i]));
            float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3])) * scale_log2;
            // the other half of the row lives in the partner warp: exchange the maxima
            float* xs = xchg + (j & 1) * (2 * FQ);
            xs[hs * FQ + row] = mx;
            asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
            mx = fmaxf(mx, xs[(hs ^ 1) * FQ + row]);
            // lazy rescale: raise the reference maximum only when it is exceeded by more than 2^8
            const bool raise = mx > m_used + 8.0f;
            const float m_new = raise ? mx : m_used;
            const float alpha = (raise && j > 0) ? exp2f(m_used - m_new) : 1.0f;
            const bool any_raise = __any_sync(0xFFFFFFFFu, raise && j > 0);
            if (any_raise) {
                // O must be stable: P(j-1) V accumulated (P(j-2) V certainly is -- S(j) was issued after it)
                mbar_wait(o_done, (j - 1) & 1);
                tc_fence_after();
                // rescale this thread's 128 O columns of its row (rows that keep their maximum use alpha = 1)
#pragma unroll 1
                for (int cchunk = 0; cchunk < 4; ++cchunk) {
                    uint32_t o[32];
                    tmem_ld_32x32(tmem_o + hs * 128 + cchunk * 32 + lane_sel, o);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                    tmem_st_32x32(tmem_o + hs * 128 + cchunk * 32 + lane_sel, o);
                }
                l *= alpha;
            }
            m_used = m_new;
            // P = exp2(s - m) as fp16 pairs, written over this row's S values: keys 2w, 2w+1 -> column w
            float ls4[4] = {0.f, 0.f, 0.f, 0.f};
            uint32_t pk[32];
#pragma unroll
            for (int cc = 0; cc < 2; ++cc)
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const float a = fast_exp2(fmaf(__uint_as_float(r[cc][2 * i]), scale_log2, -m_used));
                    const float b = fast_exp2(fmaf(__uint_as_float(r[cc][2 * i + 1]), scale_log2, -m_used));
                    ls4[i & 3] += a + b;
                    pk[cc * 16 + i] = pack_f16x2(a, b);
                }
            tmem_st_32x32(tmem_s + sb * FK + hs * 32 + lane_sel, pk);
            l += (ls4[0] + ls4[1]) + (ls4[2] + ls4[3]);
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster_relaxed(&p_full[sb], 0);   // the leader issues P V for both CTAs
        }
        // ---- epilogue: O / l + b_v -> fp16 global; the row sum is the two threads' parts added up
        {
            float* xs = xchg + (key_tiles & 1) * (2 * FQ);
            xs[hs * FQ + row] = l;
            asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
            l += xs[(hs ^ 1) * FQ + row];
        }
        mbar_wait(o_done, (key_tiles - 1) & 1);
        tc_fence_after();
        const float inv = 1.0f / l;
        const int qrow = q0 + row;
        __half* orow = out + (static_cast<long long>(img) * tokens + qrow) * FD + half * FDV + hs * 128;
#pragma unroll 1
        for (int cchunk = 0; cchunk < 4; ++cchunk) {
            uint32_t o[32];
            tmem_ld_32x32(tmem_o + hs * 128 + cchunk * 32 + lane_sel, o);
            tmem_ld_wait();
            if (qrow < tokens) {
                const float* bv = bias_v + half * FDV + hs * 128 + cchunk * 32;
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
    cluster_sync_all();      // neither CTA leaves (or frees tensor memory) while the pair still works
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc_pair(tmem_base, 512);
    }
}

}  // namespace

int launch_flash_attention(const FlashOp& op, cudaStream_t stream, Profiler* prof) {
    VT_CHECK(op.C == FD, "fused attention is specialised for head_dim 512");
    const long long ldv = op.ld_vt ? op.ld_vt : op.tokens;
    VT_CHECK(op.tokens > 0 && op.n > 0 && ldv >= op.tokens && ldv % 8 == 0,
             "fused attention: the row pitch of V^T must be a multiple of 8 elements and at least the token count");
    CUtensorMap tq, tk, tv;
    {
        uint64_t dims[3] = {static_cast<uint64_t>(2 * FD), static_cast<uint64_t>(op.tokens), static_cast<uint64_t>(op.n)};
        uint64_t str[2] = {2ull * 2 * FD, 2ull * 2 * FD * op.tokens};
        uint32_t box[3] = {64, 128, 1};
        VT_TRY(make_tmap(&tq, op.qk, 3, dims, str, box));
        uint32_t boxk[3] = {64, 64, 1};     // one CTA's half of a 128-key chunk
        VT_TRY(make_tmap(&tk, op.qk, 3, dims, str, boxk));
    }
    {
        uint64_t dims[3] = {static_cast<uint64_t>(op.tokens), static_cast<uint64_t>(FD), static_cast<uint64_t>(op.n)};
        uint64_t str[2] = {2ull * ldv, 2ull * ldv * FD};
        uint32_t box[3] = {64, 64, 1};      // 64 keys x one CTA's 64 of the 128 d_v rows
        VT_TRY(make_tmap(&tv, op.vt, 3, dims, str, box));
    }
    static SmemAttrOnce once;
    VT_TRY(ensure_dyn_smem(once, flash_d512_kernel, FLASH_SMEM));
    const int q_tiles = (op.tokens + FQ - 1) / FQ;
    const int q_pairs = (q_tiles + 1) / 2;
    const int grid = op.n * q_pairs * 4;   // pair x d_v half
    const double key_t = (op.tokens + FK - 1) / FK * FK;
    // algorithmic work: QK^T and PV once each (the second QK^T of the d_v split is overhead, not counted)
    const double flops = 2.0 * op.n * q_tiles * FQ * key_t * FD * 2.0;
    profiler_begin(prof, KC_FLASH, stream, flops, 2.0 * op.n * op.tokens * FD * 4.0);
    flash_d512_kernel<<<grid, FLASH_THREADS, FLASH_SMEM, stream>>>(tq, tk, tv, static_cast<__half*>(op.out), op.bias_v,
                                                                   op.tokens, q_pairs,
                                                                   op.scale * 1.4426950408889634f);
    profiler_end(prof, KC_FLASH, stream);
    VT_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace vt
