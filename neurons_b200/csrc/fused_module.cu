// The whole VanillaTemporalModule.forward at C = 320 (d_h = 40, 8 heads) in ONE persistent kernel
// (motion_module.py:134-158 -> :210-222 -> :270-329; motion_module_new.py:258-287, :441-471, :497-518).
//
// Why one kernel: at C = 320 every stage of the multi-kernel pipeline is bound by HBM, not by the tensor cores -- one call moves
// ~1.76 GB (fp32 residual stream read + written by each of the 4 residual GEMMs and 3 LayerNorms, q|k|v, the [N,4C] GEGLU
// activations) for 295 GFLOP: 271 us of pure HBM time against 180 us of tensor time at the cuBLAS peak.  But the module is LOCAL to
// a spatial position once the GroupNorm statistics are known: a tile of 128 tokens = (128 / F) positions x F frames can run
// GroupNorm-apply -> proj_in -> [LN + PE -> QKV -> attention -> to_out (+h)] x A -> LN -> GEGLU -> ff_out (+h) -> proj_out (+x)
// without ever leaving the SM.  HBM traffic per call drops to x (twice: statistics + tile) and y; only the 4.5 MB of weights
// stream through shared memory (from L2), once per tile.
//
// Data placement per CTA (one 128-token tile at a time, persistent over tiles):
//   TMEM   H  = columns [0, 320)    the fp32 RESIDUAL STREAM: proj_in writes it, to_out / ff_out accumulate onto it with the MMA's
//                                   own accumulate input (the residual add costs nothing), proj_out finally overwrites it.  Biases
//                                   are not added into TMEM: readers add the cumulative bias vector cb_k (packed at pack time).
//          S  = columns [320, 512)  scratch accumulators: two 80-column buffers for the q / k / v units of a head pair,
//                                   or one 128-column buffer for a GEGLU chunk.
//   SMEM   A1  80 KB  the A operand of the current stage (GroupNorm tokens / LayerNorm output / bf16 h), K-major WITHOUT
//                     swizzle: [40 chunks of 8 channels][128 rows][16 B] -- the thread that owns a row (= its TMEM lane)
//                     writes 16-byte pieces at consecutive addresses across the warp: conflict-free.
//          T   60 KB  q | k | v of one head pair as bf16 (30 chunks, rows position-major and XOR-swizzled for ldmatrix),
//                     later the GEGLU activation chunks (3 x 16 KB, the A operand of the chained ff_out MMAs);
//          CTX 20 KB  attention output of the head pair in A-operand layout;   T + CTX double as the x staging tile.
//          RING 3 x 20 KB weight stages filled by TMA (128-byte swizzle; to_out's 16-channel tail by a 1-D bulk copy).
// Row order inside the tile: m = f * ppt + pl (frame-major; ppt = 128 / F positions): a warp's 32 TMEM lanes are consecutive
// positions of one or two frames, so the x loads and y stores of the NCHW tensors are full 32-byte sectors per channel.
//
// Roles (640 threads): warp 0 = TMA producer of the weight stream, warp 1 = tcgen05.mma issuer (one thread), warp 2 = TMEM
// allocator, warps 4-19 = 16 "epilogue" warps (4 per TMEM lane quadrant) that do everything else: x load + GroupNorm apply,
// LayerNorm (+PE) straight out of TMEM, q|k|v dumps, the mma.sync temporal attention, GEGLU, the y store.
// The three programs (producer / MMA / epilogue) are the same static sequence per tile; they meet only at mbarriers.
//
// GEGLU (+) ff_out chaining: chunk j of the 4C axis (128 packed W1 rows = 64 value/gate pairs) is computed into S, GEGLU'd by the
// epilogue warps into a 128 x 64 bf16 A tile, and immediately consumed by ff_out's K = 64 slice -- the [N, 4C] intermediate
// never exists.  The MMA order G_0, G_1, F_0, G_2, F_1, ... keeps the tensor pipe busy while the epilogue works.
#include <cuda.h>

#include <cstdlib>
#include <cstring>
#include <mutex>
#include <type_traits>
#include <vector>

#include "attention_core.cuh"
#include "common.cuh"
#include "ptx.cuh"

namespace nmm {

constexpr int FM_C = 320, FM_DH = 40;
constexpr int FM_EW = 16;                                  // epilogue warps
constexpr int FM_THREADS = 128 + 32 * FM_EW;               // 640
constexpr int FM_ETHREADS = 32 * FM_EW;
// Warp roles.  The epilogue warps are warps 0-15 and the three single-thread roles sit ABOVE them: the SM's issue arbiter prefers the
// highest warp id among the eligible warps, and a TMA producer / MMA issuer that waits behind four busy epilogue warps of its scheduler
// starves the tensor pipe (measured: 110 instead of 52 cycles per N = 80 MMA with the roles at warps 0 / 1).
constexpr int FM_W_PRODUCER = FM_EW, FM_W_MMA = FM_EW + 1, FM_W_ALLOC = FM_EW + 2, FM_W_PREFETCH = FM_EW + 3;
constexpr uint32_t FM_CHUNK = 2048;                        // bytes of one 8-channel chunk of a 128-row operand tile
constexpr uint32_t FM_A1 = 0;                              // 40 chunks
constexpr uint32_t FM_T = 40 * FM_CHUNK;                   // 30 chunks (q: 0-9, k: 10-19, v: 20-29)
constexpr uint32_t FM_CTX = FM_T + 30 * FM_CHUNK;          // 10 chunks
constexpr uint32_t FM_RING_BYTES = 61440;                  // weight ring: 3 x 20 KB (single CTA) or 6 x 10 KB (CTA pair: each CTA stages half)
constexpr int FM_MAX_STAGES = 6;
constexpr uint32_t FM_RING = FM_CTX + 10 * FM_CHUNK;
constexpr uint32_t FM_SCR = FM_RING + FM_RING_BYTES;       // 4 KB: GroupNorm mean/rstd table, LayerNorm partial sums
constexpr uint32_t FM_BAR = FM_SCR + 4096;
constexpr uint32_t FM_SMEM_BYTES = FM_BAR + 1024 + 1024;   // + alignment slack
constexpr uint32_t FM_ACT_BYTES = 8 * FM_CHUNK;            // one GEGLU activation tile: 128 rows x 64 channels
constexpr int FM_ACT_BUFS = 3;
constexpr int FM_FF_CHUNKS = 4 * FM_C / 64;                // 20
constexpr uint32_t FM_TAIL_BYTES = 2 * 160 * 16;           // to_out tail of one N half: [2 chunks][160 rows][16 B]
// TMEM columns
constexpr uint32_t FM_TM_H = 0, FM_TM_S = 320, FM_TM_COLS = 512;
static_assert(FM_SMEM_BYTES <= 227 * 1024, "shared memory budget");

struct FmMaps {                 // weight tensor maps (bf16, 64-element = 128-byte swizzled rows)
    CUtensorMap win, wout;      // [320, 320],   box 160 rows
    CUtensorMap w1;             // [2560, 320],  box 128 rows (GEGLU-interleaved rows, api.cu)
    CUtensorMap w2;             // [320, 1280],  box 160 rows
    CUtensorMap wqkv[NMM_MAX_ATTN];   // [960, 320] tile order (q|k|v rows of a head pair adjacent), box 80 rows
    CUtensorMap wo[NMM_MAX_ATTN];     // [320, 320], box 160 rows
    CUtensorMap tail[NMM_MAX_ATTN];   // CTA pairs: wo_tail as [4 * 2 * 2 * 160 rows][8] (16-byte rows, no swizzle), box 80 rows
};

struct FmParams {
    const bf16 *x; bf16 *y;
    int64_t xsb, xsc, xsf, ysb, ysc, ysf;
    int B, F, P, A, ppt, tiles_per_b;
    int64_t ntiles;
    // GroupNorm
    const double *gn_partial; int gn_splits; double gn_count; float gn_eps;
    // fp32 vector blocks (packed by nmm_pack_params, staged through shared memory phase by phase); cb_k = cumulative bias of the
    // residual stream after k bias-carrying GEMMs (the GEMMs accumulate onto TMEM without their bias)
    const float *vec_attn[NMM_MAX_ATTN];          // cb_i[320] | ln gamma_i[320] | (ln beta_i + pe_i[f])[max_len or 1][320]
    const float *vec_ff;                          // cb_A[320] | ff gamma[320] | ff beta[320]
    const float *vec_fin;                         // cb_{A+1}[320] | b_out[320]
    const float *b1;                              // GEGLU bias, packed (interleaved) order [2560]
    const bf16 *wo_tail[NMM_MAX_ATTN];            // [4 pairs][2 halves][2 chunks][160 rows][8]
    int pos_enc, y_vec16;
    float ln_eps, scale_log2e;
    float *stage_dump; int stage_id;              // tests: fp32 [N, 320] snapshot of the residual stream after stage `stage_id`
    float2 *y_part;                               // or null: [tile][F][32 groups] (sum, sum of squares) of the tile's y values as stored
    unsigned long long *trace;                    // -DNMM_TRACE builds: clock64 stamps of CTA 0's first tile (development)
};

#ifdef NMM_TRACE
constexpr int FM_TRACE_SLOTS = 256;
#define FM_TRACE(slot)                                                                                     \
    do {                                                                                                   \
        if (p.trace != nullptr && blockIdx.x == 0 && t == tile0 && lane == 0) p.trace[slot] = clock64();   \
    } while (0)
#else
#define FM_TRACE(slot) do { } while (0)
#endif

// ---- small device helpers -----------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void fm_sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint4 fm_lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t fm_lds_u16(uint32_t addr) {
    uint16_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr) : "memory");
    return (uint32_t)v;
}
__device__ __forceinline__ void fm_sts32(uint32_t addr, uint32_t v) { asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ void fm_bar_epi() { asm volatile("bar.sync 1, %0;" ::"n"(FM_ETHREADS) : "memory"); }      // all 16 epilogue warps
__device__ __forceinline__ void fm_bar_quad(int q) { asm volatile("bar.sync %0, 128;" ::"r"(2 + q) : "memory"); }    // the 4 warps of a lane quadrant

// row of (local position pl, frame f) inside the q|k|v tile: position-major (the F frames of a position are F consecutive
// 16-byte pieces of a chunk -> one ldmatrix reads 128 contiguous bytes) with the low 3 frame bits XOR-ed by the position, so that
// the dump -- whose lanes are consecutive POSITIONS of one frame -- is conflict-free too.
template <int F>
__device__ __forceinline__ uint32_t fm_trow(int pl, int f) { return (uint32_t)(pl * F + (f & 8) + ((f & 7) ^ (pl & 7))); }

// ---- temporal attention on the q|k|v tile (same arithmetic as attention_core.cuh; other addressing) ----------------------------------
// One warp, NP problems = (position pl[u], head hd[u] of the pair).  O (bf16) overwrites the problem's q slot.
template <int F, int NP>
__device__ __forceinline__ void fm_attention(uint32_t T, const int (&pl)[NP], const int (&hd)[NP], int lane, float scale_log2e) {
    constexpr int NT = F / 8;
    const int lrow = lane & 7, lmat = lane >> 3;
    const int crow = lane >> 2, ccol = (lane & 3) * 2;
    uint32_t qb[NP], kb[NP], vb[NP];          // chunk bases of the problem's head
    uint32_t r_lo[NP], r_hi[NP];              // tile-row byte offsets of frame lrow / lrow + 8
#pragma unroll
    for (int u = 0; u < NP; u++) {
        qb[u] = T + (uint32_t)(hd[u] * 5) * FM_CHUNK;
        kb[u] = qb[u] + 10 * FM_CHUNK;
        vb[u] = qb[u] + 20 * FM_CHUNK;
        r_lo[u] = fm_trow<F>(pl[u], lrow) * 16;
        r_hi[u] = F == 16 ? fm_trow<F>(pl[u], lrow + 8) * 16 : 0u;
    }
    float s[NP][NT][4];
#pragma unroll
    for (int u = 0; u < NP; u++)
#pragma unroll
        for (int j = 0; j < NT; j++) { s[u][j][0] = s[u][j][1] = s[u][j][2] = s[u][j][3] = 0.f; }
    // ---- S = Q K^T ----
    if constexpr (F == 8) {
#pragma unroll
        for (int u = 0; u < NP; u++) {
            uint32_t a0, a1, a2, a3, b0, b1, b2, b3;
            ldsm_x4(qb[u] + (uint32_t)lmat * FM_CHUNK + r_lo[u], a0, a1, a2, a3);        // chunks 0-3 of q, frames 0-7
            ldsm_x4(kb[u] + (uint32_t)lmat * FM_CHUNK + r_lo[u], b0, b1, b2, b3);
            mma_k16(s[u][0], a0, 0u, a1, 0u, b0, b1);
            mma_k16(s[u][0], a2, 0u, a3, 0u, b2, b3);
        }
#pragma unroll
        for (int u = 0; u < NP; u++) {
            uint32_t a0, b0;
            ldsm_x1(qb[u] + 4 * FM_CHUNK + r_lo[u], a0);                                   // channels 32-39
            ldsm_x1(kb[u] + 4 * FM_CHUNK + r_lo[u], b0);
            mma_k8(s[u][0], a0, 0u, b0);
        }
    } else {
#pragma unroll
        for (int c0 = 0; c0 < 4; c0 += 2) {
#pragma unroll
            for (int u = 0; u < NP; u++) {
                uint32_t a0, a1, a2, a3;
                // matrices: (frames 0-7, chunk c0), (frames 8-15, c0), (frames 0-7, c0 + 1), (frames 8-15, c0 + 1)
                ldsm_x4(qb[u] + (uint32_t)(c0 + (lmat >> 1)) * FM_CHUNK + ((lmat & 1) ? r_hi[u] : r_lo[u]), a0, a1, a2, a3);
#pragma unroll
                for (int j = 0; j < NT; j++) {
                    uint32_t b0, b1;
                    ldsm_x2(kb[u] + (uint32_t)(c0 + (lmat & 1)) * FM_CHUNK + (j ? r_hi[u] : r_lo[u]), b0, b1);
                    mma_k16(s[u][j], a0, a1, a2, a3, b0, b1);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < NP; u++) {
            uint32_t a0, a1;
            ldsm_x2(qb[u] + 4 * FM_CHUNK + ((lmat & 1) ? r_hi[u] : r_lo[u]), a0, a1);
#pragma unroll
            for (int j = 0; j < NT; j++) {
                uint32_t b0;
                ldsm_x1(kb[u] + 4 * FM_CHUNK + (j ? r_hi[u] : r_lo[u]), b0);
                mma_k8(s[u][j], a0, a1, b0);
            }
        }
    }
    // ---- softmax over the keys (fp32, base-2), rows crow (regs 0,1) and crow + 8 (regs 2,3; F == 16) ----
    uint32_t ph[NP][4], pw[NP][4];
#pragma unroll
    for (int u = 0; u < NP; u++) {
        float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
        for (int j = 0; j < NT; j++) { mx0 = fmaxf(mx0, fmaxf(s[u][j][0], s[u][j][1])); mx1 = fmaxf(mx1, fmaxf(s[u][j][2], s[u][j][3])); }
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
        if constexpr (F == 16) { mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2)); }
        float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
        for (int j = 0; j < NT; j++) {
            s[u][j][0] = exp2f((s[u][j][0] - mx0) * scale_log2e); s[u][j][1] = exp2f((s[u][j][1] - mx0) * scale_log2e);
            sum0 += s[u][j][0] + s[u][j][1];
            if constexpr (F == 16) {
                s[u][j][2] = exp2f((s[u][j][2] - mx1) * scale_log2e); s[u][j][3] = exp2f((s[u][j][3] - mx1) * scale_log2e);
                sum1 += s[u][j][2] + s[u][j][3];
            }
        }
        sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1); sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
        if constexpr (F == 16) { sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1); sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2); }
        const float inv0 = 1.0f / sum0, inv1 = (F == 16) ? 1.0f / sum1 : 0.f;
        ph[u][0] = ph[u][1] = ph[u][2] = ph[u][3] = 0u; pw[u][0] = pw[u][1] = pw[u][2] = pw[u][3] = 0u;
        split_bf16x2(s[u][0][0] * inv0, s[u][0][1] * inv0, ph[u][0], pw[u][0]);
        if constexpr (F == 16) {
            split_bf16x2(s[u][0][2] * inv1, s[u][0][3] * inv1, ph[u][1], pw[u][1]);
            split_bf16x2(s[u][1][0] * inv0, s[u][1][1] * inv0, ph[u][2], pw[u][2]);
            split_bf16x2(s[u][1][2] * inv1, s[u][1][3] * inv1, ph[u][3], pw[u][3]);
        }
    }
    __syncwarp();                                    // every lane has read its q rows: they may be overwritten with O
    // O (frame `row`, head-dim columns col, col + 1 packed in v) -> the problem's q slot
    auto store = [&](int u, int row, int col, uint32_t v) {
        fm_sts32(qb[u] + (uint32_t)(col >> 3) * FM_CHUNK + fm_trow<F>(pl[u], row) * 16 + (uint32_t)((col & 7) * 2), v);
    };
    if constexpr (F == 8) {
        static_assert(F != 8 || NP == 2, "8 frames: two problems per warp share every m16n8k16 through a block-diagonal P");
        // ldmatrix matrices 0, 2 <- problem 0's V rows (keys 0-7), 1, 3 <- problem 1's
        const uint32_t vsel = (lmat & 1) ? vb[NP - 1] + r_lo[NP - 1] : vb[0] + r_lo[0];
#pragma unroll
        for (int n0 = 0; n0 < FM_DH; n0 += 16) {
            uint32_t bv[4] = {0u, 0u, 0u, 0u};
            if (n0 + 16 <= FM_DH) ldsm_x4_t(vsel + (uint32_t)(n0 / 8 + (lmat >> 1)) * FM_CHUNK, bv[0], bv[1], bv[2], bv[3]);
            else ldsm_x2_t(vsel + (uint32_t)(n0 / 8) * FM_CHUNK, bv[0], bv[1]);
#pragma unroll
            for (int q = 0; q < 2; q++) {
                if (n0 + 8 * q < FM_DH) {
                    float o[4] = {0.f, 0.f, 0.f, 0.f};
                    mma_k16(o, ph[0][0], 0u, 0u, ph[NP - 1][0], bv[2 * q], bv[2 * q + 1]);
                    mma_k16(o, pw[0][0], 0u, 0u, pw[NP - 1][0], bv[2 * q], bv[2 * q + 1]);
                    store(0, crow, ccol + n0 + 8 * q, pack_bf16x2(o[0], o[1]));
                    store(NP - 1, crow, ccol + n0 + 8 * q, pack_bf16x2(o[2], o[3]));
                }
            }
        }
    } else {
#pragma unroll
        for (int n0 = 0; n0 + 16 <= FM_DH; n0 += 16) {
#pragma unroll
            for (int u = 0; u < NP; u++) {
                uint32_t bv[4];
                // matrices: (keys 0-7, n0), (keys 8-15, n0), (keys 0-7, n0 + 8), (keys 8-15, n0 + 8)
                ldsm_x4_t(vb[u] + (uint32_t)(n0 / 8 + (lmat >> 1)) * FM_CHUNK + ((lmat & 1) ? r_hi[u] : r_lo[u]), bv[0], bv[1], bv[2], bv[3]);
#pragma unroll
                for (int q = 0; q < 2; q++) {
                    float o[4] = {0.f, 0.f, 0.f, 0.f};
                    mma_k16(o, ph[u][0], ph[u][1], ph[u][2], ph[u][3], bv[2 * q], bv[2 * q + 1]);
                    mma_k16(o, pw[u][0], pw[u][1], pw[u][2], pw[u][3], bv[2 * q], bv[2 * q + 1]);
                    store(u, crow, ccol + n0 + 8 * q, pack_bf16x2(o[0], o[1]));
                    store(u, crow + 8, ccol + n0 + 8 * q, pack_bf16x2(o[2], o[3]));
                }
            }
        }
#pragma unroll
        for (int u = 0; u < NP; u++) {
            uint32_t b0, b1;
            ldsm_x2_t(vb[u] + 4 * FM_CHUNK + ((lmat & 1) ? r_hi[u] : r_lo[u]), b0, b1);
            float o[4] = {0.f, 0.f, 0.f, 0.f};
            mma_k16(o, ph[u][0], ph[u][1], ph[u][2], ph[u][3], b0, b1);
            mma_k16(o, pw[u][0], pw[u][1], pw[u][2], pw[u][3], b0, b1);
            store(u, crow, ccol + 32, pack_bf16x2(o[0], o[1]));
            store(u, crow + 8, ccol + 32, pack_bf16x2(o[2], o[3]));
        }
    }
}

// ---- the kernel -----------------------------------------------------------------------------------------------------------------------
// CG = CTAs per cluster.  CG == 2: a CTA pair runs two tiles in lock-step with ONE tcgen05.mma.cta_group::2 stream issued by the even
// CTA (M = 256: rows 0-127 -> this CTA's TMEM, 128-255 -> the peer's); each CTA stages only its HALF of every weight tile.  Why it
// matters here: shared memory moves 128 B/clk/SM, and at cta_group::1 the weight stream alone (TMA fill writes + B operand reads,
// 64 B/clk each at full MMA rate) takes all of it before the A operand is read once (measured: ~115 cycles per MMA whatever its
// N).  The pair halves both.  The epilogue warps of both CTAs arrive on the leader's barriers; tcgen05.commit multicasts back.
template <int F, int CG>
__global__ void __launch_bounds__(FM_THREADS, 1) fused_module_kernel(const __grid_constant__ FmMaps maps, const FmParams p) {
    constexpr int PPT = 128 / F;                       // positions per tile
    // The rank inside the (2, 1, 1) cluster IS blockIdx.x & 1 -- and it is taken from blockIdx on purpose: the compiler must KNOW the
    // rank is uniform.  With %cluster_ctarank (asm or intrinsic) the leader-only MMA role was compiled as divergent code: every
    // tcgen05.mma operand built in vector registers and moved to the uniform file by a waterfall loop, ~100 cycles per MMA.
    const uint32_t rank = CG > 1 ? (blockIdx.x & 1u) : 0u;
    const bool leader = rank == 0;
    extern __shared__ unsigned char smem_raw[];
    // (broadcast through a shuffle: the compiler then knows the base -- and every descriptor / barrier address derived from it -- is
    //  warp-uniform and keeps them in uniform registers for the tcgen05 / TMA instructions)
    const uint32_t sb = __shfl_sync(0xffffffffu, (ptx::smem_u32(smem_raw) + 1023u) & ~1023u, 0);
    unsigned char *sm = smem_raw + (sb - ptx::smem_u32(smem_raw));
    const uint32_t A1 = sb + FM_A1, T = sb + FM_T, CTX = sb + FM_CTX, RING = sb + FM_RING, BAR = sb + FM_BAR;
    float *scr = reinterpret_cast<float *>(sm + FM_SCR);
    // Weight ring: FM_STAGES stages of FM_STAGE_BYTES.  One fill = what feeds 4 (N = 160 / 128) or 8 (N = 80) MMAs:
    //   N = 320 GEMMs: one 64-wide k-block of ONE 160-column half (160 / CG rows);  q|k|v units: two k-blocks of 80 / CG rows;
    //   GEGLU chunk: one k-block of 128 / CG rows;  to_out tail: [2 halves][2 chunks][160 / CG rows][16 B].
    constexpr int FM_STAGES = 3 * CG;
    constexpr uint32_t FM_STAGE_BYTES = FM_RING_BYTES / FM_STAGES;
    auto wfull = [&](int s) { return BAR + 8u * s; };
    auto wempty = [&](int s) { return BAR + 8u * (FM_MAX_STAGES + s); };
    const uint32_t a1_ready = BAR + 8u * 12, h_done = BAR + 8u * 13;
    auto s_full = [&](int b) { return BAR + 8u * (14 + b); };
    auto s_free = [&](int b) { return BAR + 8u * (16 + b); };
    const uint32_t ctx_ready = BAR + 8u * 18, ctx_free = BAR + 8u * 19, g_full = BAR + 8u * 20, g_free = BAR + 8u * 21;
    auto act_ready = [&](int b) { return BAR + 8u * (22 + b); };
    auto act_free = [&](int b) { return BAR + 8u * (25 + b); };
    const uint32_t x_taken = BAR + 8u * 28;
    const uint32_t tmem_slot = BAR + 8u * 30;
    uint32_t n_pref = 0;

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);      // provably warp-uniform: role branches stay uniform
    const int lane = threadIdx.x & 31;
    if (warp == FM_W_PRODUCER && lane == 0) {
        ptx::prefetch_tensormap(&maps.win); ptx::prefetch_tensormap(&maps.wout); ptx::prefetch_tensormap(&maps.w1); ptx::prefetch_tensormap(&maps.w2);
        for (int i = 0; i < p.A; i++) { ptx::prefetch_tensormap(&maps.wqkv[i]); ptx::prefetch_tensormap(&maps.wo[i]); }
    }
    if (warp == FM_W_MMA && lane == 0) {
        for (int s = 0; s < FM_STAGES; s++) { ptx::mbar_init(wfull(s), 1); ptx::mbar_init(wempty(s), 1); }
        // barriers the MMA thread waits on collect one arrive per epilogue warp of EVERY CTA of the cluster (they live in the leader)
        ptx::mbar_init(a1_ready, FM_EW * CG); ptx::mbar_init(h_done, 1);
        for (int b = 0; b < 2; b++) { ptx::mbar_init(s_full(b), 1); ptx::mbar_init(s_free(b), FM_EW * CG); }
        ptx::mbar_init(ctx_ready, FM_EW * CG); ptx::mbar_init(ctx_free, 1);
        ptx::mbar_init(g_full, 1); ptx::mbar_init(g_free, FM_EW * CG);
        ptx::mbar_init(x_taken, 1);
        for (int b = 0; b < FM_ACT_BUFS; b++) { ptx::mbar_init(act_ready(b), FM_EW * CG); ptx::mbar_init(act_free(b), 1); }
        ptx::fence_mbar_init();
    }
    if (warp == FM_W_ALLOC) ptx::tmem_alloc<CG>(tmem_slot, FM_TM_COLS);
    pdl_wait();                       // everything above overlapped the previous kernel (gn_stats) -- now its sums are visible
    pdl_launch_dependents();
    ptx::tc_fence_before();
    if (CG > 1) ptx::cluster_sync();  // the peer's barriers must exist before any remote arrive / complete_tx
    else __syncthreads();
    ptx::tc_fence_after();
    uint32_t tmem_base_ld;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base_ld) : "r"(tmem_slot));
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, tmem_base_ld, 0);

    const int64_t tile0 = blockIdx.x, tstep = gridDim.x;
    const int A = p.A;

    if (warp == FM_W_PRODUCER) {
        // =========================== weight producer (every CTA; CG == 2: its half of every W tile) ===========================
        // (the whole warp runs the schedule; one elected lane issues)
        {
            int stage = 0; uint32_t phase = 0;
            // `bytes` = what THIS CTA loads; with a CTA pair both halves are counted on the leader's barrier (the one the MMA thread waits on)
            auto begin = [&](uint32_t bytes) {
                ptx::mbar_wait(wempty(stage), phase ^ 1u);
                if (CG == 1 || leader) { if (ptx::elect_one()) ptx::mbar_expect_tx(wfull(stage), bytes * CG); }
            };
            auto next = [&]() { __syncwarp(); if (++stage == FM_STAGES) { stage = 0; phase ^= 1u; } };
            auto box = [&](const CUtensorMap *m, uint32_t off, int k_elem, int row) {
                const uint32_t dst = RING + (uint32_t)stage * FM_STAGE_BYTES + off;
                if (ptx::elect_one()) {
                    if constexpr (CG == 1) ptx::tma_load_2d(m, wfull(stage), dst, k_elem, row);
                    else ptx::tma_load_2d_2sm(m, wfull(stage) & ptx::PEER_MASK, dst, k_elem, row);
                }
            };
            constexpr int HR = 160 / CG, UR = 80 / CG, GR = 128 / CG;     // rows this CTA stages of an N = 160 block / a q|k|v unit / a GEGLU chunk
            // one k-block of N = 320: two fills, one per 160-column half (N <= 256 per MMA; each CTA of a pair stages 80 rows of each half)
            auto kblock320 = [&](const CUtensorMap *m, int k_elem) {
                for (int half = 0; half < 2; half++) { begin(HR * 128); box(m, 0, k_elem, half * 160 + (int)rank * HR); next(); }
            };
            for (int64_t t = tile0; t < p.ntiles; t += tstep) {
                for (int kb = 0; kb < 5; kb++) kblock320(&maps.win, kb * 64);
                for (int i = 0; i < A; i++) {
                    auto tproj = [&](int hp) {                   // to_out slice of head pair hp: K = 64 (swizzled box) + 16 (un-swizzled tail)
                        kblock320(&maps.wo[i], hp * 80);
                        begin(2 * 2 * HR * 16);                  // [2 halves][2 chunks][HR rows][16 B]
                        if constexpr (CG == 1) {
                            if (ptx::elect_one())
                                ptx::bulk_load_1d(RING + (uint32_t)stage * FM_STAGE_BYTES, p.wo_tail[i] + (size_t)hp * FM_TAIL_BYTES, 2 * FM_TAIL_BYTES, wfull(stage));
                        } else {
                            for (int h = 0; h < 2; h++)
                                for (int ch = 0; ch < 2; ch++)
                                    box(&maps.tail[i], (uint32_t)(h * 2 + ch) * HR * 16, 0, ((hp * 2 + h) * 2 + ch) * 160 + (int)rank * HR);
                        }
                        next();
                    };
                    for (int hp = 0; hp < 4; hp++) {
                        for (int s = 0; s < 3; s++) {            // q / k / v unit of the pair: K = 320 as fills of 2 + 2 + 1 k-blocks
                            const int row = hp * 240 + s * 80 + (int)rank * UR;
                            for (int kb0 = 0; kb0 < 5; kb0 += 2) {
                                const int nkb = 5 - kb0 < 2 ? 1 : 2;
                                begin((uint32_t)nkb * UR * 128);
                                for (int j = 0; j < nkb; j++) box(&maps.wqkv[i], (uint32_t)j * UR * 128, (kb0 + j) * 64, row);
                                next();
                            }
                            if (i == 0) FM_TRACE(220 + 3 * hp + s);
                            if (s == 2 && hp > 0) tproj(hp - 1);       // after the pair's last unit: the attention of pair hp waits for that unit, not for this
                        }
                    }
                    tproj(3);
                }
                for (int j = 0; j <= FM_FF_CHUNKS; j++) {
                    if (j < FM_FF_CHUNKS) {
                        for (int kb = 0; kb < 5; kb++) { begin(GR * 128); box(&maps.w1, 0, kb * 64, j * 128 + (int)rank * GR); next(); }
                        FM_TRACE(232 + j);
                    }
                    if (j > 0) kblock320(&maps.w2, (j - 1) * 64);
                }
                for (int kb = 0; kb < 5; kb++) kblock320(&maps.wout, kb * 64);
            }
        }
    } else if (warp == FM_W_MMA) {
        // =========================== MMA issuer (one thread; of the even CTA when CG == 2) ===========================
        // (the whole warp runs the schedule and the barrier waits; one elected lane issues the tcgen05 instructions)
        if (leader) {
            constexpr uint32_t MM = 128 * CG;
            const uint32_t id160 = ptx::umma_idesc_bf16(MM, 160), id80 = ptx::umma_idesc_bf16(MM, 80), id128 = ptx::umma_idesc_bf16(MM, 128);
            int stage = 0; uint32_t phase = 0;
            uint32_t n_a1 = 0, n_ctx = 0, n_unit = 0, n_g = 0;   // running use counters of the barriers (phase = count parity)
            auto wait_fill = [&]() { ptx::mbar_wait(wfull(stage), phase); ptx::tc_fence_after(); };
            auto commit = [&](uint32_t bar) { if (ptx::elect_one()) ptx::umma_commit<CG>(bar); __syncwarp(); };
            auto mma = [&](uint32_t d, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t acc) { if (ptx::elect_one()) ptx::umma_bf16<CG>(d, ad, bd, idesc, acc); };
            auto release = [&]() { commit(wempty(stage)); if (++stage == FM_STAGES) { stage = 0; phase ^= 1u; } };
            auto stage_addr = [&]() { return RING + (uint32_t)stage * FM_STAGE_BYTES; };
            auto adesc = [&](uint32_t base, int chunk) { return ptx::umma_smem_desc_interleave(base + (uint32_t)chunk * FM_CHUNK, FM_CHUNK, 128); };
            // barriers the epilogue warps (of both CTAs) arrive on: acquire at cluster scope
            auto wait_epi = [&](uint32_t bar, uint32_t parity) {
                ptx::mbar_wait(bar, parity);
                ptx::tc_fence_after();
            };
            auto wait_a1 = [&]() { wait_epi(a1_ready, n_a1 & 1u); n_a1++; };
            // H (+)= A . W^T for one 64-wide k-block whose A chunks start at (abase, chunk0): N = 320 as two fills / two N = 160 blocks
            auto mma320_kblock = [&](uint32_t abase, int chunk0, bool acc_first) {
                for (int half = 0; half < 2; half++) {
                    wait_fill();
                    const uint64_t bd = ptx::umma_smem_desc_sw128(stage_addr());
#pragma unroll
                    for (int k = 0; k < 4; k++)
                        mma(tmem_base + FM_TM_H + half * 160, adesc(abase, chunk0 + 2 * k), bd + (uint64_t)(k * 2), id160, (acc_first || k != 0) ? 1u : 0u);
                    release();
                }
            };
            for (int64_t t = tile0; t < p.ntiles; t += tstep) {
                wait_a1();                                       // GroupNorm tokens are in A1 (and the previous tile's y epilogue is done with H)
                FM_TRACE(100);
                for (int kb = 0; kb < 5; kb++) mma320_kblock(A1, kb * 8, kb != 0);
                commit(h_done);
                FM_TRACE(101);
                for (int i = 0; i < A; i++) {
                    wait_a1();                                   // LayerNorm_i(h) + pe
                    auto tproj = [&](bool last) {                // H += ctx . Wo[:, pair]^T
                        wait_epi(ctx_ready, n_ctx & 1u); n_ctx++;
                        mma320_kblock(CTX, 0, true);
                        wait_fill();                             // channels 64-79 of the pair: un-swizzled 16-channel weight tail
                        constexpr uint32_t HR = 160 / CG;
                        for (int half = 0; half < 2; half++)
                            mma(tmem_base + FM_TM_H + half * 160, adesc(CTX, 8),
                                               ptx::umma_smem_desc_interleave(stage_addr() + (uint32_t)half * 2 * HR * 16, HR * 16, 128), id160, 1u);
                        release();
                        commit(ctx_free);
                        if (last) commit(h_done);
                    };
                    for (int hp = 0; hp < 4; hp++) {
                        for (int s = 0; s < 3; s++) {
                            const int b = n_unit & 1;
                            wait_epi(s_free(b), ((n_unit >> 1) & 1u) ^ 1u);      // the epilogue has drained this buffer's previous unit
                            constexpr int UR = 80 / CG;
                            for (int kb0 = 0; kb0 < 5; kb0 += 2) {
                                const int nkb = 5 - kb0 < 2 ? 1 : 2;
                                wait_fill();
                                for (int j = 0; j < nkb; j++) {
                                    const int kb = kb0 + j;
                                    const uint64_t bd = ptx::umma_smem_desc_sw128(stage_addr() + (uint32_t)j * UR * 128);
#pragma unroll
                                    for (int k = 0; k < 4; k++)
                                        mma(tmem_base + FM_TM_S + b * 80, adesc(A1, kb * 8 + 2 * k), bd + (uint64_t)(k * 2), id80, (kb | k) != 0 ? 1u : 0u);
                                }
                                release();
                            }
                            commit(s_full(b));
                            n_unit++;
                            FM_TRACE(102 + 20 * i + 3 * hp + s);
                            if (s == 2 && hp > 0) { tproj(false); FM_TRACE(102 + 20 * i + 12 + hp - 1); }
                        }
                    }
                    tproj(true);
                    FM_TRACE(102 + 20 * i + 15);
                }
                wait_a1();                                       // LayerNorm_ff(h)
                for (int j = 0; j <= FM_FF_CHUNKS; j++) {
                    if (j < FM_FF_CHUNKS) {                      // G_j: S = A1 . W1[chunk j]^T
                        wait_epi(g_free, (n_g & 1u) ^ 1u);
                        for (int kb = 0; kb < 5; kb++) {
                            wait_fill();
                            const uint64_t bd = ptx::umma_smem_desc_sw128(stage_addr());
#pragma unroll
                            for (int k = 0; k < 4; k++)
                                mma(tmem_base + FM_TM_S, adesc(A1, kb * 8 + 2 * k), bd + (uint64_t)(k * 2), id128, (kb | k) != 0 ? 1u : 0u);
                            release();
                        }
                        commit(g_full);
                        n_g++;
                        FM_TRACE(150 + j);
                    }
                    if (j > 0) {                                 // F_{j-1}: H += act . W2[:, chunk j-1]^T
                        const uint32_t g = n_g - (j < FM_FF_CHUNKS ? 2u : 1u);       // global index of chunk j - 1
                        const int b = (int)(g % FM_ACT_BUFS);
                        wait_epi(act_ready(b), (g / FM_ACT_BUFS) & 1u);
                        mma320_kblock(T + (uint32_t)b * FM_ACT_BYTES, 0, true);
                        commit(act_free(b));
                        FM_TRACE(175 + j - 1);
                    }
                }
                commit(h_done);
                wait_a1();                                       // bf16(h)
                FM_TRACE(198);
                for (int kb = 0; kb < 5; kb++) mma320_kblock(A1, kb * 8, kb != 0);      // proj_out overwrites H
                commit(h_done);
                FM_TRACE(199);
            }
        }
    } else if (warp == FM_W_PREFETCH) {
        // =========================== x prefetcher: the NEXT tile's 80 KB of x into L2 while this tile computes ===========================
        for (int64_t t = tile0; t < p.ntiles; t += tstep) {
            const int64_t tn = t + tstep;
            if (tn >= p.ntiles) break;
            const int b = (int)(tn / p.tiles_per_b);
            const int p0 = (int)(tn - (int64_t)b * p.tiles_per_b) * PPT;
            const bf16 *xb = p.x + (int64_t)b * p.xsb + p0;
            for (int it = lane; it < FM_C * F; it += 32) {            // one (channel, frame) run of PPT positions (32 or 16 bytes) each
                const int c = it / F, f = it - c * F;
                asm volatile("prefetch.global.L2 [%0];" ::"l"(xb + (int64_t)c * p.xsc + (int64_t)f * p.xsf));
            }
            // pace: one tile ahead is enough (wait until the epilogue warps have published this tile's tokens)
            ptx::mbar_wait(x_taken, n_pref & 1u); n_pref++;
        }
    } else if (warp < FM_EW) {
        // =========================== epilogue warps ===========================
        const int ew = warp, q = ew & 3, sub = ew >> 2;
        const int et = (int)threadIdx.x;
        const int m = q * 32 + lane;                             // tile row = TMEM lane
        const int f_m = m / PPT, pl_m = m % PPT;
        const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
        const uint32_t VA = T + 10 * FM_CHUNK;                   // vector staging of the attention LayerNorms: the k | v part of the tile
        const uint32_t VF = T + 3 * FM_ACT_BYTES;                // ... of the feed-forward / output phases: behind the activation buffers
        uint32_t n_hd = 0, n_unit = 0, n_ctx = 0, n_g = 0;
        auto wait_h = [&]() { ptx::mbar_wait(h_done, n_hd & 1u); n_hd++; ptx::tc_fence_after(); };
        // arrive on a barrier the MMA thread waits on (it lives in the even CTA of a pair)
        auto arrive_mma = [&](uint32_t bar) {
            if constexpr (CG == 1) ptx::mbar_arrive(bar); else ptx::mbar_arrive_remote(bar & ptx::PEER_MASK);
        };
        auto publish_a1 = [&]() {                                // generic-proxy writes of A1 -> visible to the tensor core, then arrive
            ptx::fence_proxy_async();
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) arrive_mma(a1_ready);
        };
        // global fp32 vectors -> shared memory (all epilogue threads, 16 bytes each): with 226 KB of shared memory there is no L1
        // left, every __ldg of a bias / affine vector would be an L2 round trip on the critical path.
        auto stage_vec = [&](uint32_t dst, const float *src, int n4) {
            for (int i = et; i < n4; i += FM_ETHREADS) {
                const float4 v = __ldg(reinterpret_cast<const float4 *>(src) + i);
                fm_sts128(dst + (uint32_t)i * 16, __float_as_uint(v.x), __float_as_uint(v.y), __float_as_uint(v.z), __float_as_uint(v.w));
            }
        };
        auto lds4 = [&](uint32_t addr) { const uint4 v = fm_lds128(addr); return make_float4(__uint_as_float(v.x), __uint_as_float(v.y), __uint_as_float(v.z), __uint_as_float(v.w)); };
        // LayerNorm (+ beta + pe) of the residual stream (TMEM H + cumulative bias) -> A1.  Each thread: its row, 80 of the 320 columns.
        // Vector block at `vb` (shared memory): cb[320] | gamma[320] | (beta + pe[f])[rows][320].
        auto layer_norm = [&](uint32_t vb, int bpe_row) {
            const int cbeg = 80 * sub;
            const uint32_t cbv = vb + (uint32_t)cbeg * 4, gmv = vb + 1280 + (uint32_t)cbeg * 4, btv = vb + 2560 + (uint32_t)bpe_row * 1280 + (uint32_t)cbeg * 4;
            const uint32_t h0 = ptx::tmem_ld1(t_lane + FM_TM_H);
            ptx::tmem_ld_wait();
            float x0;
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x0) : "r"(vb));
            x0 += __uint_as_float(h0);                               // shift by the row's first element: keeps the one-pass variance well conditioned
            // (packed fp32x2 arithmetic: the phase is instruction-bound -- 16 warps x 80 columns -- and FADD2 / FFMA2 halve the FP count;
            //  same operation order per element as the scalar form, hence the same roundings)
            const uint64_t nx0 = f32x2_pack(-x0, -x0);
            uint64_t s1p = f32x2_pack(0.f, 0.f), s2p = s1p;
#pragma unroll
            for (int k = 0; k < 5; k++) {
                uint32_t r[16];
                ptx::tmem_ld16(t_lane + FM_TM_H + cbeg + 16 * k, r);
                float4 c4[4];
#pragma unroll
                for (int j4 = 0; j4 < 4; j4++) c4[j4] = lds4(cbv + (uint32_t)(16 * k + 4 * j4) * 4);
                ptx::tmem_ld_wait();
#pragma unroll
                for (int j4 = 0; j4 < 4; j4++) {
                    const uint64_t d01 = f32x2_add(f32x2_add(f32x2_pack(__uint_as_float(r[4 * j4]), __uint_as_float(r[4 * j4 + 1])), f32x2_pack(c4[j4].x, c4[j4].y)), nx0);
                    const uint64_t d23 = f32x2_add(f32x2_add(f32x2_pack(__uint_as_float(r[4 * j4 + 2]), __uint_as_float(r[4 * j4 + 3])), f32x2_pack(c4[j4].z, c4[j4].w)), nx0);
                    s1p = f32x2_add(s1p, f32x2_add(d01, d23));
                    s2p = f32x2_fma(d01, d01, s2p);
                    s2p = f32x2_fma(d23, d23, s2p);
                }
            }
            float s1a, s1b, s2a, s2b;
            f32x2_unpack(s1p, s1a, s1b);
            f32x2_unpack(s2p, s2a, s2b);
            reinterpret_cast<float2 *>(scr)[m * 4 + sub] = make_float2(s1a + s1b, s2a + s2b);
            fm_bar_quad(q);
            float S1 = 0.f, S2 = 0.f;
#pragma unroll
            for (int i = 0; i < 4; i++) { const float2 t2 = reinterpret_cast<const float2 *>(scr)[m * 4 + i]; S1 += t2.x; S2 += t2.y; }
            const float md = S1 * (1.0f / FM_C);
            const float mean = x0 + md;
            const float rstd = rsqrtf(fmaxf(S2 * (1.0f / FM_C) - md * md, 0.f) + p.ln_eps);
            const uint64_t nmean = f32x2_pack(-mean, -mean), rs2 = f32x2_pack(rstd, rstd);
#pragma unroll
            for (int k = 0; k < 5; k++) {
                const int c0 = cbeg + 16 * k;
                uint32_t r[16];
                ptx::tmem_ld16(t_lane + FM_TM_H + c0, r);
                float4 c4[4];
#pragma unroll
                for (int j4 = 0; j4 < 4; j4++) c4[j4] = lds4(cbv + (uint32_t)(16 * k + 4 * j4) * 4);
                ptx::tmem_ld_wait();
                uint32_t o[8];
#pragma unroll
                for (int j4 = 0; j4 < 4; j4++) {
                    const float4 gm = lds4(gmv + (uint32_t)(16 * k + 4 * j4) * 4);
                    const float4 bt = lds4(btv + (uint32_t)(16 * k + 4 * j4) * 4);
                    // ((h + cb) - mean) * rstd * gamma + (beta + pe)
                    uint64_t t01 = f32x2_add(f32x2_add(f32x2_pack(__uint_as_float(r[4 * j4]), __uint_as_float(r[4 * j4 + 1])), f32x2_pack(c4[j4].x, c4[j4].y)), nmean);
                    uint64_t t23 = f32x2_add(f32x2_add(f32x2_pack(__uint_as_float(r[4 * j4 + 2]), __uint_as_float(r[4 * j4 + 3])), f32x2_pack(c4[j4].z, c4[j4].w)), nmean);
                    t01 = f32x2_fma(f32x2_mul(t01, rs2), f32x2_pack(gm.x, gm.y), f32x2_pack(bt.x, bt.y));
                    t23 = f32x2_fma(f32x2_mul(t23, rs2), f32x2_pack(gm.z, gm.w), f32x2_pack(bt.z, bt.w));
                    float a0, a1, a2, a3;
                    f32x2_unpack(t01, a0, a1);
                    f32x2_unpack(t23, a2, a3);
                    o[2 * j4] = pack_bf16x2(a0, a1);
                    o[2 * j4 + 1] = pack_bf16x2(a2, a3);
                }
                const uint32_t dst = A1 + (uint32_t)(c0 >> 3) * FM_CHUNK + (uint32_t)m * 16;
                fm_sts128(dst, o[0], o[1], o[2], o[3]);
                fm_sts128(dst + FM_CHUNK, o[4], o[5], o[6], o[7]);
            }
            publish_a1();
        };
        // tests: snapshot of the residual stream (H + cumulative bias, from the staged vector block) of this tile's rows
        auto dump_stage = [&](int id, uint32_t cbv, int64_t token) {
            if (p.stage_dump == nullptr || p.stage_id != id) return;
#pragma unroll
            for (int k = 0; k < 5; k++) {
                const int c0 = 80 * sub + 16 * k;
                uint32_t r[16];
                ptx::tmem_ld16(t_lane + FM_TM_H + c0, r);
                ptx::tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; j++) {
                    float cbj;
                    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(cbj) : "r"(cbv + (uint32_t)(c0 + j) * 4));
                    p.stage_dump[token * FM_C + c0 + j] = __uint_as_float(r[j]) + cbj;
                }
            }
        };
        // x tile -> staging (T + CTX = 80 KB): one 32-bit word per (channel PAIR, row): word (c / 2) * 128 + m = {x[c][m], x[c + 1][m]}, so that the
        // row's owner reads / writes two channels per shared-memory access.  Global side: 16-byte pieces of 8 positions of one channel.
        auto prmt = [](uint32_t a, uint32_t b, uint32_t sel) { uint32_t d; asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel)); return d; };
        auto load_x_tile = [&](const bf16 *xb, uint4 (&v)[10]) {
#pragma unroll
            for (int i = 0; i < 5; i++) {
                const int it = et + FM_ETHREADS * i;                   // 2560 items: (channel pair, piece of 8 positions)
                const int cp = it >> 4, piece = it & 15;
                const int m0 = piece * 8, f = m0 / PPT, pl0 = m0 % PPT;
                const bf16 *src = xb + (int64_t)(2 * cp) * p.xsc + (int64_t)f * p.xsf + pl0;
                v[2 * i] = __ldg(reinterpret_cast<const uint4 *>(src));
                v[2 * i + 1] = __ldg(reinterpret_cast<const uint4 *>(src + p.xsc));
            }
        };
        auto store_x_tile = [&](const uint4 (&v)[10]) {
#pragma unroll
            for (int i = 0; i < 5; i++) {
                const int it = et + FM_ETHREADS * i;
                const uint4 a = v[2 * i], c = v[2 * i + 1];
                const uint32_t dst = T + (uint32_t)((it >> 4) * 128 + (it & 15) * 8) * 4;
                fm_sts128(dst, prmt(a.x, c.x, 0x5410), prmt(a.x, c.x, 0x7632), prmt(a.y, c.y, 0x5410), prmt(a.y, c.y, 0x7632));
                fm_sts128(dst + 16, prmt(a.z, c.z, 0x5410), prmt(a.z, c.z, 0x7632), prmt(a.w, c.w, 0x5410), prmt(a.w, c.w, 0x7632));
            }
        };
        auto lds32 = [](uint32_t addr) { uint32_t v; asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory"); return v; };

        for (int64_t t = tile0; t < p.ntiles; t += tstep) {
            const int b = (int)(t / p.tiles_per_b);
            const int p0 = (int)(t - (int64_t)b * p.tiles_per_b) * PPT;
            const int64_t token = ((int64_t)b * F + f_m) * p.P + p0 + pl_m;        // this thread's row in (b, f, p) token order
            const bf16 *xb = p.x + (int64_t)b * p.xsb + p0;
#ifdef NMM_TRACE
#define FM_ETRACE(slot) do { if (ew == 0) FM_TRACE(slot); } while (0)
#else
#define FM_ETRACE(slot) do { } while (0)
#endif
            FM_ETRACE(0);
            // ---- phase 0: x tile -> staging, GroupNorm statistics of the F images -> scr, normalised tokens -> A1 ----
            // (gamma / beta of the GroupNorm are folded into proj_in's weight and bias at pack time: tokens = (x - mean) * rstd)
            {
                uint4 v[10];
                load_x_tile(xb, v);
                for (int i = et; i < F * NMM_GN_GROUPS; i += FM_ETHREADS) {
                    float mean, rstd;
                    gn_finalize_one(p.gn_partial, (b * F + i / NMM_GN_GROUPS) * NMM_GN_GROUPS + i % NMM_GN_GROUPS, p.gn_splits, p.gn_count, p.gn_eps, mean, rstd);
                    scr[2 * i] = rstd; scr[2 * i + 1] = -mean * rstd;
                }
                store_x_tile(v);
                fm_bar_epi();
                if (warp == 0 && lane == 0) ptx::mbar_arrive(x_taken);       // the prefetcher may run one more tile ahead
                FM_ETRACE(1);
                for (int ck = 10 * sub; ck < 10 * sub + 10; ck++) {         // this row, 8 channels (4 staged pairs) at a time
                    uint32_t w[4];
#pragma unroll
                    for (int i2 = 0; i2 < 4; i2++) {
                        const int c = ck * 8 + 2 * i2;                      // both channels of a pair lie in one GroupNorm group (10 channels)
                        const float2 ab = reinterpret_cast<const float2 *>(scr)[f_m * NMM_GN_GROUPS + c / (FM_C / NMM_GN_GROUPS)];
                        const uint32_t xw = lds32(T + (uint32_t)((ck * 4 + i2) * 128 + m) * 4);
                        w[i2] = pack_bf16x2(fmaf(bf16_lo(xw), ab.x, ab.y), fmaf(bf16_hi(xw), ab.x, ab.y));
                    }
                    fm_sts128(A1 + (uint32_t)ck * FM_CHUNK + (uint32_t)m * 16, w[0], w[1], w[2], w[3]);
                }
                publish_a1();
            }
            FM_ETRACE(2);
            fm_bar_epi();                                            // every warp is done with the x staging: T is free for the vectors
            stage_vec(VA, p.vec_attn[0], 160);                                       // cb_0 | gamma_0
            stage_vec(VA + 2560, p.vec_attn[0] + 2 * FM_C, (p.pos_enc ? F : 1) * 80);   // (beta_0 + pe[f]), f < F
            fm_bar_epi();
            wait_h();                                                // proj_in done: H = tokens . W_in'^T
            FM_ETRACE(3);
            dump_stage(0, VA, token);
            // ---- attention blocks ----
            for (int i = 0; i < A; i++) {
                layer_norm(VA, p.pos_enc ? f_m : 0);
                // The k | v dumps below overwrite the vector block at VA.  They are already ordered after every warp's last read of it
                // (a1_ready collects all 16 x CG warps -> MMA -> tcgen05.commit -> s_full), but only through mbarriers, which
                // compute-sanitizer's racecheck does not follow: this barrier costs nothing (the warps would wait for the MMA anyway)
                // and makes the ordering visible to the tool.
                fm_bar_epi();
                FM_ETRACE(4 + 20 * i);
                for (int hp = 0; hp < 4; hp++) {
                    for (int s = 0; s < 3; s++) {
                        // dump this warp's share of the unit's 80 accumulator columns as bf16 into the q|k|v tile
                        const int bsel = n_unit & 1;
                        ptx::mbar_wait(s_full(bsel), (n_unit >> 1) & 1u);
                        n_unit++;
                        ptx::tc_fence_after();
                        const int cb0 = sub < 2 ? 24 * sub : 48 + 16 * (sub - 2);           // 3 + 3 + 2 + 2 chunks
                        const uint32_t tcol = t_lane + FM_TM_S + bsel * 80 + cb0;
                        uint32_t ra[16], rb[8];
                        ptx::tmem_ld16(tcol, ra);
                        if (sub < 2) ptx::tmem_ld8(tcol + 16, rb);
                        ptx::tmem_ld_wait();
                        ptx::tc_fence_before();
                        __syncwarp();
                        if (lane == 0) arrive_mma(s_free(bsel));                       // the accumulator buffer may be overwritten
                        const uint32_t dst = T + (uint32_t)(s * 10 + (cb0 >> 3)) * FM_CHUNK + fm_trow<F>(pl_m, f_m) * 16;
#define FM_PK(a, i) pack_bf16x2(__uint_as_float(a[i]), __uint_as_float(a[i + 1]))
                        fm_sts128(dst, FM_PK(ra, 0), FM_PK(ra, 2), FM_PK(ra, 4), FM_PK(ra, 6));
                        fm_sts128(dst + FM_CHUNK, FM_PK(ra, 8), FM_PK(ra, 10), FM_PK(ra, 12), FM_PK(ra, 14));
                        if (sub < 2) fm_sts128(dst + 2 * FM_CHUNK, FM_PK(rb, 0), FM_PK(rb, 2), FM_PK(rb, 4), FM_PK(rb, 6));
#undef FM_PK
                    }
                    fm_bar_epi();                                    // the pair's q | k | v tile is complete
                    FM_ETRACE(4 + 20 * i + 1 + 4 * hp);
                    if constexpr (F == 8) {                          // 32 problems: (position, head) = ew and ew + 16
                        const int pls[2] = {ew >> 1, (ew >> 1) + 8}, hds[2] = {ew & 1, ew & 1};
                        fm_attention<8, 2>(T, pls, hds, lane, p.scale_log2e);
                    } else {                                         // 16 problems
                        const int pls[1] = {ew >> 1}, hds[1] = {ew & 1};
                        fm_attention<16, 1>(T, pls, hds, lane, p.scale_log2e);
                    }
                    fm_bar_epi();                                    // every problem's O sits in its q slot
                    FM_ETRACE(4 + 20 * i + 2 + 4 * hp);
                    ptx::mbar_wait(ctx_free, (n_ctx & 1u) ^ 1u);     // the previous pair's to_out MMAs have read CTX
                    n_ctx++;
                    for (int it = et; it < 10 * 128; it += FM_ETHREADS) {          // q slots (tile order) -> CTX (A-operand order)
                        const int j = it >> 7, mm = it & 127;
                        const uint4 v = fm_lds128(T + (uint32_t)j * FM_CHUNK + fm_trow<F>(mm % PPT, mm / PPT) * 16);
                        fm_sts128(CTX + (uint32_t)j * FM_CHUNK + (uint32_t)mm * 16, v.x, v.y, v.z, v.w);
                    }
                    ptx::fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) arrive_mma(ctx_ready);
                    fm_bar_epi();                                    // the tile may be overwritten by the next pair's dumps
                    FM_ETRACE(4 + 20 * i + 3 + 4 * hp);
                }
                // the next LayerNorm's vectors while the last to_out MMAs run (k | v part of the tile is free; CTX is not)
                if (i + 1 < A) {
                    stage_vec(VA, p.vec_attn[i + 1], 160);
                    stage_vec(VA + 2560, p.vec_attn[i + 1] + 2 * FM_C, (p.pos_enc ? F : 1) * 80);
                } else {
                    stage_vec(VA, p.vec_ff, 240);                    // cb_A | gamma_ff | beta_ff
                }
                fm_bar_epi();
                wait_h();                                            // to_out of all four pairs accumulated onto H
                FM_ETRACE(4 + 20 * i + 17);
                dump_stage(1 + i, VA, token);
            }
            // ---- feed-forward: LayerNorm -> [GEGLU chunk -> ff_out slice] x 20 ----
            layer_norm(VA, 0);
            FM_ETRACE(44);
            // GEGLU biases (packed order) and the output-phase vectors behind the activation buffers (the LayerNorm block above sits inside them:
            // every warp has left layer_norm once the first accumulator chunk arrives -- a1_ready needs all 16 warps)
            stage_vec(VF, p.b1, 8 * FM_C / 4);
            stage_vec(VF + 8 * FM_C * 4, p.vec_fin, 160);            // cb_{A+1} | b_out
            fm_bar_epi();
            for (int j = 0; j < FM_FF_CHUNKS; j++) {
                if (j == 10) FM_ETRACE(240);
                ptx::mbar_wait(g_full, n_g & 1u);
                ptx::tc_fence_after();
                if (j == 10) FM_ETRACE(241);
                uint32_t r[32];
                ptx::tmem_ld32(t_lane + FM_TM_S + 32 * sub, r);
                ptx::tmem_ld_wait();
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) arrive_mma(g_free);             // S may take the next chunk while the GELUs run
                if (j == 10) FM_ETRACE(242);
                const uint32_t b1v = VF + (uint32_t)(j * 128 + 32 * sub) * 4;
                uint32_t o[8];
#pragma unroll
                for (int i4 = 0; i4 < 8; i4++) {                     // accumulator columns 4i .. 4i+3 = value 2q, value 2q+1, gate 2q, gate 2q+1
                    const float4 bb = lds4(b1v + (uint32_t)i4 * 16);
                    float v0, v1, g0, g1, y0, y1;
                    f32x2_unpack(f32x2_add(f32x2_pack(__uint_as_float(r[4 * i4]), __uint_as_float(r[4 * i4 + 1])), f32x2_pack(bb.x, bb.y)), v0, v1);
                    f32x2_unpack(f32x2_add(f32x2_pack(__uint_as_float(r[4 * i4 + 2]), __uint_as_float(r[4 * i4 + 3])), f32x2_pack(bb.z, bb.w)), g0, g1);
                    geglu_pair(v0, g0, v1, g1, y0, y1);
                    o[i4] = pack_bf16x2(y0, y1);
                }
                const int ab = (int)(n_g % FM_ACT_BUFS);
                if (j == 10) FM_ETRACE(243);
                ptx::mbar_wait(act_free(ab), ((n_g / FM_ACT_BUFS) & 1u) ^ 1u);       // ff_out has consumed this buffer's previous chunk
                if (j == 10) FM_ETRACE(244);
                n_g++;
                const uint32_t dst = T + (uint32_t)ab * FM_ACT_BYTES + (uint32_t)(2 * sub) * FM_CHUNK + (uint32_t)m * 16;
                fm_sts128(dst, o[0], o[1], o[2], o[3]);
                fm_sts128(dst + FM_CHUNK, o[4], o[5], o[6], o[7]);
                ptx::fence_proxy_async();
                __syncwarp();
                if (lane == 0) arrive_mma(act_ready(ab));
                FM_ETRACE(45 + j);
            }
            wait_h();                                                // h = h + ff(...)
            FM_ETRACE(65);
            const uint32_t cbf = VF + 8 * FM_C * 4, bov = cbf + 1280;
            dump_stage(1 + A, cbf, token);
            // ---- bf16(h) -> A1, the A operand of proj_out ----
#pragma unroll
            for (int k = 0; k < 5; k++) {
                const int c0 = 80 * sub + 16 * k;
                uint32_t r[16];
                ptx::tmem_ld16(t_lane + FM_TM_H + c0, r);
                float4 c4[4];
#pragma unroll
                for (int j4 = 0; j4 < 4; j4++) c4[j4] = lds4(cbf + (uint32_t)(c0 + 4 * j4) * 4);
                ptx::tmem_ld_wait();
                float o[16];
#pragma unroll
                for (int j4 = 0; j4 < 4; j4++) {
                    o[4 * j4] = __uint_as_float(r[4 * j4]) + c4[j4].x; o[4 * j4 + 1] = __uint_as_float(r[4 * j4 + 1]) + c4[j4].y;
                    o[4 * j4 + 2] = __uint_as_float(r[4 * j4 + 2]) + c4[j4].z; o[4 * j4 + 3] = __uint_as_float(r[4 * j4 + 3]) + c4[j4].w;
                }
                const uint32_t dst = A1 + (uint32_t)(c0 >> 3) * FM_CHUNK + (uint32_t)m * 16;
                fm_sts128(dst, pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]), pack_bf16x2(o[4], o[5]), pack_bf16x2(o[6], o[7]));
                fm_sts128(dst + FM_CHUNK, pack_bf16x2(o[8], o[9]), pack_bf16x2(o[10], o[11]), pack_bf16x2(o[12], o[13]), pack_bf16x2(o[14], o[15]));
            }
            publish_a1();
            FM_ETRACE(66);
            // ---- y[b, c, f, p] = acc + b_out[c] + x[b, c, f, p] through the staging tile: x in with 16-byte loads (issued before the
            // wait for proj_out), the sum formed in place by the row's owner, y out with 16-byte stores ----
            {
                uint4 v[10];
                load_x_tile(xb, v);
                // the x tile is about to overwrite the vector staging (it lies inside T + CTX): park b_out in the scratch area
                for (int i = et; i < 80; i += FM_ETHREADS) {
                    const uint4 w = fm_lds128(bov + (uint32_t)i * 16);
                    fm_sts128(sb + FM_SCR + (uint32_t)i * 16, w.x, w.y, w.z, w.w);
                }
                fm_bar_epi();                                        // all reads of VF done (cb_f above, b_out just now)
                store_x_tile(v);
                fm_bar_epi();
                wait_h();                                            // H = h . W_out^T
                FM_ETRACE(67);
                // (EMIT: also accumulate this row's per-GroupNorm-group sums of the ROUNDED outputs -- the next InflatedGroupNorm's statistics)
                auto y_rows = [&](auto EMIT) {
                    constexpr bool kEmit = decltype(EMIT)::value;
                    float gv[kEmit ? 16 : 1];                             // this thread's 80 channels = 8 groups of 10: (sum, sum of squares) pairs
                    if constexpr (kEmit) {
#pragma unroll
                        for (int i = 0; i < 16; i++) gv[i] = 0.f;
                    }
#pragma unroll
                    for (int k = 0; k < 5; k++) {
                        const int c0 = 80 * sub + 16 * k;
                        uint32_t r[16];
                        ptx::tmem_ld16(t_lane + FM_TM_H + c0, r);
                        float4 b4[4];
#pragma unroll
                        for (int j4 = 0; j4 < 4; j4++) b4[j4] = lds4(sb + FM_SCR + (uint32_t)(c0 + 4 * j4) * 4);
                        uint32_t xw[8];
#pragma unroll
                        for (int j = 0; j < 8; j++) xw[j] = lds32(T + (uint32_t)((c0 / 2 + j) * 128 + m) * 4);
                        ptx::tmem_ld_wait();
                        const float bb[16] = {b4[0].x, b4[0].y, b4[0].z, b4[0].w, b4[1].x, b4[1].y, b4[1].z, b4[1].w,
                                              b4[2].x, b4[2].y, b4[2].z, b4[2].w, b4[3].x, b4[3].y, b4[3].z, b4[3].w};
#pragma unroll
                        for (int j = 0; j < 8; j++) {
                            const uint32_t yw = pack_bf16x2(__uint_as_float(r[2 * j]) + bb[2 * j] + bf16_lo(xw[j]),
                                                            __uint_as_float(r[2 * j + 1]) + bb[2 * j + 1] + bf16_hi(xw[j]));
                            fm_sts32(T + (uint32_t)((c0 / 2 + j) * 128 + m) * 4, yw);
                            if constexpr (kEmit) {
                                const int gl = (16 * k + 2 * j) / 10;      // compile-time: both channels of the pair lie in one group
                                const float y0 = bf16_lo(yw), y1 = bf16_hi(yw);
                                gv[2 * gl] += y0 + y1;
                                gv[2 * gl + 1] = fmaf(y0, y0, fmaf(y1, y1, gv[2 * gl + 1]));
                            }
                        }
                    }
                    if constexpr (kEmit) {
                        // the PPT rows of one frame are PPT consecutive lanes: reduce-scatter over them, then lane pl_m of the frame holds
                        // floats [pl_m * 16 / PPT, ...) of the frame's 16 (= 8 groups x (sum, sum of squares)) totals
                        lane_group_reduce_scatter<16, PPT>(gv, lane);
                        float *dst = reinterpret_cast<float *>(p.y_part + ((int64_t)t * F + f_m) * NMM_GN_GROUPS + 8 * sub) + pl_m * (16 / PPT);
#pragma unroll
                        for (int i = 0; i < 16 / PPT; i++) dst[i] = gv[i];
                    }
                };
                if (p.y_part != nullptr) y_rows(std::true_type{}); else y_rows(std::false_type{});
                ptx::tc_fence_before();      // this tile's TMEM reads are ordered before the next tile's a1_ready arrive -> proj_in may overwrite H
                fm_bar_epi();
                bf16 *yb = p.y + (int64_t)b * p.ysb + p0;
                if (p.y_vec16) {
#pragma unroll
                    for (int i = 0; i < 5; i++) {
                        const int it = et + FM_ETHREADS * i;
                        const int cp = it >> 4, piece = it & 15;
                        const int m0 = piece * 8, f = m0 / PPT, pl0 = m0 % PPT;
                        const uint32_t src = T + (uint32_t)(cp * 128 + piece * 8) * 4;
                        const uint4 lo = fm_lds128(src), hi = fm_lds128(src + 16);
                        bf16 *dst = yb + (int64_t)(2 * cp) * p.ysc + (int64_t)f * p.ysf + pl0;
                        *reinterpret_cast<uint4 *>(dst) = make_uint4(prmt(lo.x, lo.y, 0x5410), prmt(lo.z, lo.w, 0x5410), prmt(hi.x, hi.y, 0x5410), prmt(hi.z, hi.w, 0x5410));
                        *reinterpret_cast<uint4 *>(dst + p.ysc) = make_uint4(prmt(lo.x, lo.y, 0x7632), prmt(lo.z, lo.w, 0x7632), prmt(hi.x, hi.y, 0x7632), prmt(hi.z, hi.w, 0x7632));
                    }
                } else {                                             // y rows not 16-byte aligned: element stores (lanes = consecutive positions)
                    for (int cp = sub; cp < FM_C / 2; cp += 4) {
                        const uint32_t w = lds32(T + (uint32_t)(cp * 128 + m) * 4);
                        uint16_t *d0 = reinterpret_cast<uint16_t *>(yb + (int64_t)(2 * cp) * p.ysc + (int64_t)f_m * p.ysf);
                        d0[pl_m] = (uint16_t)(w & 0xffffu);
                        reinterpret_cast<uint16_t *>(yb + (int64_t)(2 * cp + 1) * p.ysc + (int64_t)f_m * p.ysf)[pl_m] = (uint16_t)(w >> 16);
                    }
                }
                fm_bar_epi();                                        // the staging tile is free for the next tile's x
            }
            FM_ETRACE(68);
        }
    }
    __syncwarp();
    ptx::tc_fence_before();
    if (CG > 1) ptx::cluster_sync();  // the peer may still read this CTA's shared memory / arrive on its barriers
    else __syncthreads();
    if (warp == FM_W_ALLOC) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc<CG>(tmem_base, FM_TM_COLS);
    }
}

// ---- host side ------------------------------------------------------------------------------------------------------------------------
typedef CUresult (*FmEncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                    const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static FmEncodeTiledFn fm_encode_fn() {
    static FmEncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void *ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<FmEncodeTiledFn>(ptr);
    });
    return fn;
}
// bf16 weight [rows, cols] row-major; box = box_rows x 64 columns, 128-byte swizzle
static int fm_weight_map(CUtensorMap *tm, const void *ptr, int rows, int cols, int box_rows) {
    FmEncodeTiledFn fn = fm_encode_fn();
    if (!fn) return fail(NMM_ERR_DEVICE, "cuTensorMapEncodeTiled not available (no CUDA driver?)");
    if (!aligned(ptr, 16)) return fail(NMM_ERR_BAD_ARG, "packed weights must be 16-byte aligned");
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(NMM_ERR_CUDA, "cuTensorMapEncodeTiled (fused module weights) failed with CUresult %d", (int)r);
    return NMM_OK;
}

// Which modules keep the extra packed tensors of the fused kernel (shape-independent part; api.cu packed_layout)
bool fused_module_weights(const Geo &g) {
    return g.dtype == NMM_BF16 && !g.ln_fold && g.C == FM_C && g.heads == 8 && g.layers == 1 && g.A >= 1 && g.A <= NMM_MAX_ATTN;
}
// ... and which calls can run on it
bool fused_module_eligible(const Geo &g, const nmm_shape *s, const void *x) {
    if (!opt(NMM_OPT_FUSED_MODULE) || !fused_module_weights(g)) return false;
    if (g.F != 8 && g.F != 16) return false;
    if (g.P % (128 / g.F) != 0) return false;
    return aligned(x, 16) && s->x_stride_b % 8 == 0 && s->x_stride_c % 8 == 0 && s->x_stride_f % 8 == 0;
}

#ifdef NMM_TRACE
static unsigned long long *g_fm_trace = nullptr;
extern "C" __attribute__((visibility("default"))) int nmm_debug_fm_trace_dump(const char *path) {
    if (!g_fm_trace) return -1;
    static unsigned long long host[FM_TRACE_SLOTS];
    cudaDeviceSynchronize();
    cudaMemcpy(host, g_fm_trace, sizeof(host), cudaMemcpyDeviceToHost);
    FILE *f = fopen(path, "w");
    if (!f) return -2;
    for (int i = 0; i < FM_TRACE_SLOTS; i++) fprintf(f, "%d %llu\n", i, host[i]);
    fclose(f);
    cudaMemset(g_fm_trace, 0, sizeof(host));
    return 0;
}
#endif

// wo_tail as a 2-D tensor of 16-byte rows ([4 pairs][2 halves][2 chunks][160 rows] x 8 bf16), box = 160 rows, no swizzle
static int fm_tail_map(CUtensorMap *tm, const void *ptr) {
    FmEncodeTiledFn fn = fm_encode_fn();
    if (!fn) return fail(NMM_ERR_DEVICE, "cuTensorMapEncodeTiled not available (no CUDA driver?)");
    cuuint64_t dims[2] = {8, 4 * 2 * 2 * 160};
    cuuint64_t strides[1] = {16};
    cuuint32_t box[2] = {8, 80};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(NMM_ERR_CUDA, "cuTensorMapEncodeTiled (to_out tail) failed with CUresult %d", (int)r);
    return NMM_OK;
}

// Tensor maps depend only on the packed buffer's addresses: encoded once per (module, cluster size), not per launch.
namespace {
struct MapKey { const void *w; int A, cg; bool operator==(const MapKey &o) const { return w == o.w && A == o.A && cg == o.cg; } };
struct MapEntry { MapKey key; FmMaps maps; };
std::mutex g_map_mu;
std::vector<MapEntry> g_map_cache;        // small (one entry per live C = 320 module x cluster size): linear search
}  // namespace

static int fm_maps_for(const FusedArgs &a, int cg, FmMaps *out) {
    const MapKey key{a.w_in_g, a.A, cg};
    {
        std::lock_guard<std::mutex> lk(g_map_mu);
        for (const MapEntry &e : g_map_cache)
            if (e.key == key) { *out = e.maps; return NMM_OK; }
    }
    FmMaps maps;
    memset(&maps, 0, sizeof(maps));
    int rc;
    if ((rc = fm_weight_map(&maps.win, a.w_in_g, FM_C, FM_C, 160 / cg)) != NMM_OK) return rc;
    if ((rc = fm_weight_map(&maps.wout, a.w_out, FM_C, FM_C, 160 / cg)) != NMM_OK) return rc;
    if ((rc = fm_weight_map(&maps.w1, a.w1, 8 * FM_C, FM_C, 128 / cg)) != NMM_OK) return rc;
    if ((rc = fm_weight_map(&maps.w2, a.w2, FM_C, 4 * FM_C, 160 / cg)) != NMM_OK) return rc;
    for (int i = 0; i < a.A; i++) {
        if ((rc = fm_weight_map(&maps.wqkv[i], a.wqkv_t[i], 3 * FM_C, FM_C, 80 / cg)) != NMM_OK) return rc;
        if ((rc = fm_weight_map(&maps.wo[i], a.wo[i], FM_C, FM_C, 160 / cg)) != NMM_OK) return rc;
        if (cg == 2 && (rc = fm_tail_map(&maps.tail[i], a.wo_tail[i])) != NMM_OK) return rc;
    }
    {
        std::lock_guard<std::mutex> lk(g_map_mu);
        if (g_map_cache.size() >= 512) g_map_cache.clear();
        g_map_cache.push_back(MapEntry{key, maps});
    }
    *out = maps;
    return NMM_OK;
}

int launch_fused_module(const FusedArgs &a, cudaStream_t st) {
    const int64_t ntiles_ = (int64_t)a.B * (a.P / (128 / a.F));
    // CTA pairs whenever the tiles pair up (every UNet level at CFG batch 2); a lone / odd tile count runs the single-CTA variant
    const int64_t force = opt(NMM_OPT_FUSED_CLUSTER);
    const int cg = (force == 1 || ntiles_ % 2 != 0) ? 1 : 2;
    FmMaps maps;
    int rc = fm_maps_for(a, cg, &maps);
    if (rc != NMM_OK) return rc;
    FmParams p;
    memset(&p, 0, sizeof(p));
    p.x = (const bf16 *)a.x; p.y = (bf16 *)a.y;
    p.xsb = a.xsb; p.xsc = a.xsc; p.xsf = a.xsf; p.ysb = a.ysb; p.ysc = a.ysc; p.ysf = a.ysf;
    p.B = a.B; p.F = a.F; p.P = a.P; p.A = a.A; p.ppt = 128 / a.F; p.tiles_per_b = a.P / p.ppt;
    p.ntiles = (int64_t)a.B * p.tiles_per_b;
    p.gn_partial = a.gn_partial; p.gn_splits = a.gn_splits; p.gn_count = a.gn_count; p.gn_eps = a.gn_eps;
    for (int i = 0; i < a.A; i++) { p.vec_attn[i] = a.vec_attn[i]; p.wo_tail[i] = (const bf16 *)a.wo_tail[i]; }
    p.vec_ff = a.vec_ff; p.vec_fin = a.vec_fin; p.b1 = a.b1; p.pos_enc = a.pos_enc;
    p.y_vec16 = aligned(a.y, 16) && a.ysb % 8 == 0 && a.ysc % 8 == 0 && a.ysf % 8 == 0;
    p.ln_eps = a.ln_eps; p.scale_log2e = (1.0f / sqrtf((float)FM_DH)) * 1.4426950408889634f;
    p.stage_dump = a.stage_dump; p.stage_id = a.stage_id;
    p.y_part = a.y_part;
#ifdef NMM_TRACE
    if (!g_fm_trace) { cudaMalloc(&g_fm_trace, FM_TRACE_SLOTS * 8); cudaMemset(g_fm_trace, 0, FM_TRACE_SLOTS * 8); }
    p.trace = g_fm_trace;
#endif
    if (p.ntiles <= 0) return NMM_OK;

    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int max_grid = sms / cg * cg;
    const int grid = (int)(p.ntiles < max_grid ? p.ntiles : max_grid);
    auto kern = a.F == 8 ? (cg == 2 ? fused_module_kernel<8, 2> : fused_module_kernel<8, 1>) : (cg == 2 ? fused_module_kernel<16, 2> : fused_module_kernel<16, 1>);
    static DeviceOnce once[4];
    NMM_CUDA_OK(once[(a.F == 16 ? 2 : 0) + (cg - 1)].max_smem(kern, (int)FM_SMEM_BYTES));
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(FM_THREADS);
    cfg.dynamicSmemBytes = FM_SMEM_BYTES;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    attr[1].id = cudaLaunchAttributeClusterDimension;
    attr[1].val.clusterDim.x = (unsigned)cg;
    attr[1].val.clusterDim.y = 1;
    attr[1].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = cg > 1 ? 2 : 1;
    {
        const double N = (double)a.B * a.F * a.P;
        // algorithmic work: every Linear of the module + the attention; bytes: x, y and the weights once
        const double flops = 2.0 * N * FM_C * FM_C * (14.0 + 4.0 * a.A) + 4.0 * a.A * N * a.F * FM_C;
        const double bytes = 2.0 * N * FM_C * 2 + (double)FM_C * FM_C * (14.0 + 4.0 * a.A) * 2;
        ProfScope prof(K_FUSED_MODULE, st, flops, bytes);
        cudaError_t le = cudaLaunchKernelEx(&cfg, kern, maps, p);
        if (le != cudaSuccess) return fail(NMM_ERR_CUDA, "cudaLaunchKernelEx(fused_module_kernel) failed: %s", cudaGetErrorString(le));
    }
    NMM_LAUNCHED("fused_module_kernel");
    return NMM_OK;
}

}  // namespace nmm
