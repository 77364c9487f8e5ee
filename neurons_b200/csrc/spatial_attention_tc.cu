// Spatial self-attention on the 5th-generation tensor cores (tcgen05 + TMEM + TMA) -- the long-sequence levels of the UNet's
// Transformer3DModel blocks (arithmetic: CrossAttention._attention, motion_module_new.py:258-287):
//     O = softmax(Q K^T * d_h^-1/2) V       per (image, head),  Q [Lq, d_h], K / V [Lkv, d_h],  d_h in {40, 80},  Lkv >= 256
// (the mma.sync kernel of spatial_attention.cu keeps d_h = 160, the 77-token cross-attention and the small levels: its legacy HMMA pipe
// saturates at ~260 TFLOP/s, profiles/r2_ncu_spatial_attention.txt).
//
// Two kernels share the design below.  spatial_attention_tc2_kernel (the DEFAULT): one CTA = TWO 128-query tiles of one (image, head), one
// CTA per SM, 2 x 4 softmax warps + TMA warp + MMA warp; the two tiles share every K / V tile (the TMA pays per 16-byte segment of the
// core-matrix order, which bound the one-tile kernel once the SFU was relieved).  spatial_attention_tc_kernel: one 128-query tile per CTA, two
// CTAs per SM -- the predecessor, kept with its A/B switches (NMM_OPT_SPATIAL_ATTN: 23 one-tile, 2 two softmax threads per row, 10..14 share
// of polynomial exponentials, 15 no K / V traffic (timing only), 16 cp.async loader, 20..22 K / V stages of the two-tile kernel, 24 text
// cross-attention on this kernel, 26 two-tile kernel without K / V traffic (timing only), 27 P through tensor memory); measurements in profiles/r2_spatial_attention_probe.txt, story in DESIGN.md section 5.  Per 128-key tile t:
//   TMA warp   K(t), V(t) -> shared memory as [16-byte channel chunk][key][8 channels]: a 4-D tensor map (8, rows, chunks, images) whose box
//              (8, 128, d_h/8, 1) lands exactly in the un-swizzled core-matrix order tcgen05 reads (K-major for Q / K, N-major for V);
//              rows past the end of the image are zero-filled by the hardware; K and V are separate streams (lanes 0 / 1)
//   MMA warp   S = Q K(t)^T  (M128 N128 K16 x d_h/16, d_h = 40 zero-padded to 48) -> TMEM; issued as soon as the softmax warps have READ
//              S(t-1) into registers, i.e. under their exponentials;   O += P(t) V(t)  (M128 N48|96 K16 x 8) -> TMEM: O accumulates in
//              tensor memory over ALL tiles, and V's extra ONES column makes column d_h of O the row sum of the bf16 weights
//   softmax    tcgen05.ld of the thread's 128 scores, row max, p = 2^((s - m_ref) * scale*log2e): 5 of 8 on the SFU (ex2.approx), 3 of 8 by
//              a polynomial on the FMA / ALU pipes, packed fp32x2 arithmetic; half of them before the wait for P(t-1) V(t-1), the rest
//              interleaved with the 16-byte stores of the bf16 P tile in the K-major operand order ([key chunk][row][16 B]).  m_ref is the
//              row's REFERENCE max, raised only when the running max exceeds it by more than 2^8 (then O's row in TMEM is rescaled
//              once: rare after the first tiles); p <= 256 is exact enough in bf16 and the final O / l is independent of m_ref.
// Nothing but Q, K, V is read from and O written to HBM; the score matrix (537 MB per frame in the reference at the 64 x 64 level) never exists.
#include <mutex>

#include "common.cuh"
#include "ptx.cuh"

namespace nmm {

constexpr int FT_BM = 128;          // queries per CTA
constexpr int FT_BN = 128;          // keys per tile
constexpr int FT_TMEM_COLS = 256;
constexpr int FT_O_COL = 128;       // O accumulator columns [128, 128 + DHP)
constexpr uint32_t FT_CHUNK = FT_BN * 16;      // bytes of one 16-byte chunk column of a 128-row operand tile

template <int DH>
struct FtCfg {
    static constexpr int DHP = (DH + 15) / 16 * 16;        // K extent of S = Q K^T: 48 (zero-padded), 80
    static constexpr int NV = (DH + 1 + 15) / 16 * 16;     // N extent of O = P V: d_h channels + the ONES column (row sums), 48 / 96
    static constexpr int CH = DH / 8, CHP = DHP / 8, CHV = NV / 8;
    static constexpr int NS = DH <= 40 ? 2 : 1;            // K / V ring depth (2 CTAs per SM must fit)
    static constexpr uint32_t TILE_BYTES = CHP * FT_CHUNK; // one Q / K tile incl. the zero pad chunk
    static constexpr uint32_t VTILE_BYTES = CHV * FT_CHUNK;// one V tile incl. the pad chunks (the first holds the ones column)
    static constexpr uint32_t TX_BYTES = CH * FT_CHUNK;    // bytes a TMA box delivers
    static constexpr uint32_t P_BYTES = (FT_BN / 8) * FT_CHUNK;
    static constexpr uint32_t OFF_Q = 0, OFF_K = TILE_BYTES, OFF_V = OFF_K + NS * TILE_BYTES, OFF_P = OFF_V + NS * VTILE_BYTES;
    static constexpr uint32_t OFF_BAR = OFF_P + P_BYTES;
    static constexpr uint32_t OFF_MAX = OFF_BAR + 256;      // SP == 2: row-max exchange between the two threads of a row, [2 parities][2 halves][128] floats
    static constexpr uint32_t SMEM = OFF_MAX + 2048 + 128;  // + alignment slack
};

struct FtParams {
    void *o;
    int64_t o_rs, o_bs;
    const void *k, *v;          // cp.async loader (CPA): K / V base pointers (column slice of the projection output), row / image strides
    int64_t kv_rs, kv_bs;
    int Lq, Lkv;
    int kv_div;                 // K / V image of query image i = i / kv_div (text cross-attention: the frames of a clip share K / V)
    float scale_log2e;
#ifdef NMM_TRACE                // development build only (python -m neurons_b200.build --trace): timing experiments + per-tile timeline
    unsigned long long *trace;  // per-tile clock64 timeline of CTA (0,0,0), [role 0 softmax / 1 mma][tile < 16][event < 8], or null
    int debug;                  // results invalid: 1 = no ex2, 2 = no P stores, 4 = P V reduced to one MMA; 8 = pack by truncation
#endif
};

__device__ __forceinline__ float ft_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// SP = softmax threads per query row (1: a thread owns all 128 scores of its row; 2: warps w and w + 4 own 64 columns each of the same
// TMEM lanes -- 4 softmax warps per scheduler at 2 CTAs / SM instead of 2, which is what keeps the SFU fed while other warps are in their
// TMEM-load / max / P-store stretches).
// 2^x for x <= ~8 WITHOUT the SFU: round x to the nearest integer n with the 1.5 * 2^23 trick, 2^(x - n) by a degree-3 minimax polynomial
// on [-0.5, 0.5] (max relative error 7.5e-5, against the 2^-9 of the bf16 rounding that follows), n added to the exponent field.  9 FMA- /
// ALU-pipe instructions; FT_POLY (template parameter) of every 8 weights go this way, because the kernel is bounded by the 16 / clk / SM SFU (ex2 and the
// F2FP packing share it): the timeline of scripts/spatial_attn_trace.py shows the SFU ~95 % busy while the softmax warps are in their
// exponentials and those stretches making up 76 % of a tile's time.
__device__ __forceinline__ float ft_exp2_poly(float x) {
    x = fmaxf(x, -125.0f);
    const float fi = x + 12582912.0f;
    const float f = x - (fi - 12582912.0f);
    float p = fmaf(0.0551716536f, f, 0.2426111251f);
    p = fmaf(p, f, 0.6932609677f);
    p = fmaf(p, f, 0.9999280572f);
    return __uint_as_float(__float_as_uint(p) + (__float_as_uint(fi) << 23));
}

// Two of them at once on the packed fp32x2 pipe (FADD2 / FFMA2: one instruction, two lanes): 6 packed + 6 scalar instructions per pair
// instead of 18 -- the kernel's issue slots are its busiest resource once the SFU is relieved.
__device__ __forceinline__ void ft_exp2_poly2(uint64_t x, float &y0, float &y1) {
    float x0, x1;
    f32x2_unpack(x, x0, x1);
    const uint64_t xc = f32x2_pack(fmaxf(x0, -125.0f), fmaxf(x1, -125.0f));
    const uint64_t magic = f32x2_pack(12582912.0f, 12582912.0f);
    const uint64_t fi = f32x2_add(xc, magic);
    const uint64_t t = f32x2_add(fi, f32x2_pack(-12582912.0f, -12582912.0f));
    const uint64_t f = f32x2_fma(t, f32x2_pack(-1.0f, -1.0f), xc);
    uint64_t p = f32x2_fma(f32x2_pack(0.0551716536f, 0.0551716536f), f, f32x2_pack(0.2426111251f, 0.2426111251f));
    p = f32x2_fma(p, f, f32x2_pack(0.6932609677f, 0.6932609677f));
    p = f32x2_fma(p, f, f32x2_pack(0.9999280572f, 0.9999280572f));
    float p0, p1, i0, i1;
    f32x2_unpack(p, p0, p1);
    f32x2_unpack(fi, i0, i1);
    y0 = __uint_as_float(__float_as_uint(p0) + (__float_as_uint(i0) << 23));
    y1 = __uint_as_float(__float_as_uint(p1) + (__float_as_uint(i1) << 23));
}

#ifndef NMM_TRACE
#define FT_TRACE(role, t, ev) do { } while (0)
#define FT_DEBUG(bit) false
#else
#define FT_DEBUG(bit) ((p.debug & (bit)) != 0)
#define FT_TRACE(role, t, ev)                                                                                                   \
    do {                                                                                                                        \
        if (p.trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && lane == 0 && (t) < 16 &&           \
            ((role) == 1 || warp == 0))                                                                                         \
            p.trace[((role) * 16 + (t)) * 8 + (ev)] = (unsigned long long)clock64();                                            \
    } while (0)
#endif

template <int DH, int SP, int FT_POLY, bool NOKV = false, bool CPA = false>
__global__ void __launch_bounds__(64 + 128 * SP, 2)
spatial_attention_tc_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k, const __grid_constant__ CUtensorMap tm_v,
                            const FtParams p) {
    using Cfg = FtCfg<DH>;
    constexpr int DHP = Cfg::DHP, NV = Cfg::NV, CH = Cfg::CH, CHP = Cfg::CHP, CHV = Cfg::CHV, NS = Cfg::NS;
    extern __shared__ uint8_t ft_smem_raw[];
    const uint32_t base = (ptx::smem_u32(ft_smem_raw) + 127u) & ~127u;
    uint8_t *gbase = ft_smem_raw + (base - ptx::smem_u32(ft_smem_raw));
    const uint32_t s_q = base + Cfg::OFF_Q, s_k = base + Cfg::OFF_K, s_v = base + Cfg::OFF_V, s_p = base + Cfg::OFF_P, s_bar = base + Cfg::OFF_BAR;
    // barriers (8 bytes each)
    const uint32_t b_q = s_bar, b_sfull = s_bar + 8, b_sfree = s_bar + 16, b_pfull = s_bar + 24, b_pv = s_bar + 32;
    auto b_kfull = [&](int s) { return s_bar + 40 + 8 * s; };
    auto b_kempty = [&](int s) { return s_bar + 56 + 8 * s; };
    auto b_vfull = [&](int s) { return s_bar + 72 + 8 * s; };
    auto b_vempty = [&](int s) { return s_bar + 88 + 8 * s; };
    volatile uint32_t *tmem_slot = reinterpret_cast<volatile uint32_t *>(gbase + Cfg::OFF_BAR + 128);

    constexpr int FT_THREADS = 64 + 128 * SP, W_TMA = 4 * SP, W_MMA = 4 * SP + 1, COLS = FT_BN / SP;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q0 = blockIdx.x * FT_BM, head = blockIdx.y, img = blockIdx.z;
    const int nt = (p.Lkv + FT_BN - 1) / FT_BN;

    if (tid == 0) {
        ptx::mbar_init(b_q, 1); ptx::mbar_init(b_sfull, 1); ptx::mbar_init(b_sfree, 4 * SP); ptx::mbar_init(b_pfull, 4 * SP); ptx::mbar_init(b_pv, 1);
        for (int s = 0; s < NS; s++) { ptx::mbar_init(b_kfull(s), CPA ? 16 : 1); ptx::mbar_init(b_kempty(s), 1); ptx::mbar_init(b_vfull(s), CPA ? 16 : 1); ptx::mbar_init(b_vempty(s), 1); }
        ptx::fence_mbar_init();
    }
    // Pad chunks (TMA never writes them).  Q / K: zeros (0 x 0 instead of 0 x garbage in the padded k-step).  V: channel d_h of every key
    // is 1.0, the rest 0 -- column d_h of O = P V is then the row sum of the bf16 weights the tensor core used, accumulated in fp32 in
    // TMEM beside O: the softmax threads do no summation at all, and O / l is an exact weighted mean of V rows.
    if constexpr (CHP > CH) {
        for (int i = tid; i < (1 + NS) * (int)(FT_CHUNK / 16); i += FT_THREADS) {
            const int tile = i / (int)(FT_CHUNK / 16), r = i - tile * (int)(FT_CHUNK / 16);
            *reinterpret_cast<uint4 *>(gbase + tile * Cfg::TILE_BYTES + CH * FT_CHUNK + r * 16) = make_uint4(0u, 0u, 0u, 0u);
        }
    }
    for (int i = tid; i < NS * (CHV - CH) * (int)(FT_CHUNK / 16); i += FT_THREADS) {
        const int per = (CHV - CH) * (int)(FT_CHUNK / 16);
        const int st = i / per, r = i - st * per;            // r < FT_BN: the first pad chunk (ones column)
        *reinterpret_cast<uint4 *>(gbase + Cfg::OFF_V + st * Cfg::VTILE_BYTES + CH * FT_CHUNK + r * 16) = make_uint4(r < FT_BN ? 0x3F80u : 0u, 0u, 0u, 0u);
    }
    ptx::fence_proxy_async();
    if (warp == W_MMA) ptx::tmem_alloc<1>(ptx::smem_u32(const_cast<uint32_t *>(tmem_slot)), FT_TMEM_COLS);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    pdl_wait();
    pdl_launch_dependents();

    if (warp == W_TMA) {
        // ===================== TMA producer: lane 0 streams K, lane 1 streams V =====================
        // (two independent streams: K(t+1) is wanted early -- S(t+1) is issued under the softmax of tile t -- and must not queue behind the
        // wait for V's buffer, which P(t) V(t) releases late; with one stream the d_h = 80 kernel, whose rings are one stage deep, lost 35 %)
        if constexpr (CPA) {
            // K / V through the LSU instead of the TMA unit: a K or V tile in the un-swizzled core-matrix order is 128 x d_h/8 separate 16-byte
            // segments for the TMA (1280 per key tile at d_h = 40, shared by the SM's two CTAs: ~90 % of a tile's time), but one cp.async
            // per lane here -- lanes 0-15 stream K, 16-31 stream V; 8 consecutive lanes take 8 consecutive keys of one 16-byte channel chunk
            // (128 contiguous bytes of shared memory), the two lane groups adjacent chunks (32 / 64 contiguous bytes per key in global
            // memory).  Completion: cp.async.mbarrier.arrive.noinc on the full barrier (16 arrivals); the MMA warp adds a proxy fence.
            const int hl = lane & 15, is_v = lane >> 4;
            if (lane == 0) {
                ptx::prefetch_tensormap(&tm_q);
                ptx::mbar_expect_tx(b_q, Cfg::TX_BYTES);
                ptx::tma_load_4d(&tm_q, b_q, s_q, 0, q0, head * CH, img);
            }
            const bf16 *src0 = (const bf16 *)(is_v ? p.v : p.k) + (int64_t)img * p.kv_bs + head * DH;
            for (int t = 0; t < nt; t++) {
                const int s = t % NS, fill = t / NS;
                const uint32_t full = is_v ? b_vfull(s) : b_kfull(s), empty = is_v ? b_vempty(s) : b_kempty(s);
                if (fill > 0) ptx::mbar_wait(empty, (uint32_t)(fill - 1) & 1u);
                const uint32_t dst0 = (is_v ? s_v + s * Cfg::VTILE_BYTES : s_k + s * Cfg::TILE_BYTES) + (uint32_t)(hl & 7) * 16;
                if (!(NOKV && fill > 0)) {
#pragma unroll 2
                    for (int rg = 0; rg < FT_BN / 8; rg++) {
                        const int key = t * FT_BN + rg * 8 + (hl & 7);
                        const bool ok = key < p.Lkv;
                        const bf16 *src = src0 + (int64_t)(ok ? key : p.Lkv - 1) * p.kv_rs;
#pragma unroll
                        for (int cp = 0; cp < (CH + 1) / 2; cp++) {
                            const int c = 2 * cp + (hl >> 3);
                            if (c < CH) {
                                const int sz = ok ? 16 : 0;
                                asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst0 + (uint32_t)c * FT_CHUNK + (uint32_t)rg * 128), "l"(src + c * 8),
                                             "r"(sz)
                                             : "memory");
                            }
                        }
                    }
                }
                asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(full) : "memory");
            }
        } else
        if (lane == 0) {
            ptx::prefetch_tensormap(&tm_q); ptx::prefetch_tensormap(&tm_k);
            ptx::mbar_expect_tx(b_q, Cfg::TX_BYTES);
            ptx::tma_load_4d(&tm_q, b_q, s_q, 0, q0, head * CH, img);
            for (int t = 0; t < nt; t++) {
                const int s = t % NS, fill = t / NS;
                if (fill > 0) ptx::mbar_wait(b_kempty(s), (uint32_t)(fill - 1) & 1u);
                if (NOKV && fill > 0) { ptx::mbar_arrive(b_kfull(s)); continue; }      // timing experiment (results invalid): no K traffic after the first fill
                ptx::mbar_expect_tx(b_kfull(s), Cfg::TX_BYTES);
                ptx::tma_load_4d(&tm_k, b_kfull(s), s_k + s * Cfg::TILE_BYTES, 0, t * FT_BN, head * CH, img);
            }
        } else if (lane == 1) {
            ptx::prefetch_tensormap(&tm_v);
            for (int t = 0; t < nt; t++) {
                const int s = t % NS, fill = t / NS;
                if (fill > 0) ptx::mbar_wait(b_vempty(s), (uint32_t)(fill - 1) & 1u);
                if (NOKV && fill > 0) { ptx::mbar_arrive(b_vfull(s)); continue; }
                ptx::mbar_expect_tx(b_vfull(s), Cfg::TX_BYTES);
                ptx::tma_load_4d(&tm_v, b_vfull(s), s_v + s * Cfg::VTILE_BYTES, 0, t * FT_BN, head * CH, img);
            }
        }
        __syncwarp();
    } else if (warp == W_MMA) {
        // ===================== MMA issuer =====================
        const uint32_t idesc_s = ptx::umma_idesc_bf16(FT_BM, FT_BN);
        const uint32_t idesc_o = ptx::umma_idesc_bf16(FT_BM, NV) | ptx::UMMA_IDESC_B_MN_MAJOR;
        // V is an N-major operand in the un-swizzled layout: 16-byte channel chunks FT_CHUNK apart (SBO), 8-key groups 128 B apart (LBO)
        // (the other assignment of the two fields was tried on B200 and is wrong: profiles/r2_spatial_attention_probe.txt history)
        const uint32_t v_lbo = 128u, v_sbo = FT_CHUNK;
        auto issue_s = [&](int t) {           // S(t) = Q K(t)^T
            const int s = t % NS;
            FT_TRACE(1, t, 0);
            ptx::mbar_wait(b_kfull(s), (uint32_t)(t / NS) & 1u);
            if constexpr (CPA) ptx::fence_proxy_async();                    // cp.async wrote K through the generic proxy; the MMA reads it through the async proxy
            FT_TRACE(1, t, 1);
            if (t > 0) ptx::mbar_wait(b_sfree, (uint32_t)(t - 1) & 1u);      // the softmax warps hold S(t-1) in registers
            ptx::tc_fence_after();
            FT_TRACE(1, t, 2);
            if (ptx::elect_one()) {
                const uint64_t a = ptx::umma_smem_desc_interleave(s_q, FT_CHUNK, 128);
                const uint64_t b = ptx::umma_smem_desc_interleave(s_k + s * Cfg::TILE_BYTES, FT_CHUNK, 128);
#pragma unroll
                for (int k = 0; k < DHP / 16; k++)
                    ptx::umma_bf16<1>(tmem, a + (uint64_t)((k * 2 * FT_CHUNK) >> 4), b + (uint64_t)((k * 2 * FT_CHUNK) >> 4), idesc_s, k != 0 ? 1u : 0u);
                ptx::umma_commit<1>(b_sfull);
                ptx::umma_commit<1>(b_kempty(s));
            }
            __syncwarp();
        };
        ptx::mbar_wait(b_q, 0);
        issue_s(0);
        for (int t = 0; t < nt; t++) {
            if (t + 1 < nt) issue_s(t + 1);
            const int s = t % NS;
            FT_TRACE(1, t, 3);
            ptx::mbar_wait(b_vfull(s), (uint32_t)(t / NS) & 1u);
            if constexpr (CPA) ptx::fence_proxy_async();
            FT_TRACE(1, t, 4);
            ptx::mbar_wait(b_pfull, (uint32_t)t & 1u);
            ptx::tc_fence_after();
            FT_TRACE(1, t, 5);
            if (ptx::elect_one()) {          // O += P(t) V(t)
                const uint64_t a = ptx::umma_smem_desc_interleave(s_p, FT_CHUNK, 128);
                const uint64_t b = ptx::umma_smem_desc_interleave(s_v + s * Cfg::VTILE_BYTES, v_lbo, v_sbo);
#pragma unroll
                for (int k = 0; k < (FT_DEBUG(4) ? 1 : FT_BN / 16); k++)
                    ptx::umma_bf16<1>(tmem + FT_O_COL, a + (uint64_t)((k * 2 * FT_CHUNK) >> 4), b + (uint64_t)((k * 16 * 16) >> 4), idesc_o, (t | k) != 0 ? 1u : 0u);
                ptx::umma_commit<1>(b_pv);
                ptx::umma_commit<1>(b_vempty(s));
            }
            __syncwarp();
            FT_TRACE(1, t, 6);
        }
    } else {
        // ===================== softmax: query row = TMEM lane (warp % 4) * 32 + lane; columns [half * COLS, (half + 1) * COLS) =====
        const int half = warp >> 2, qw = warp & 3, lrow = qw * 32 + lane;
        const uint32_t t_row = tmem + ((uint32_t)(qw * 32) << 16);
        const float sl = p.scale_log2e;
        float *maxbuf = reinterpret_cast<float *>(gbase + Cfg::OFF_MAX);
        float m_ref = -INFINITY;
        for (int t = 0; t < nt; t++) {
            FT_TRACE(0, t, 0);
            ptx::mbar_wait(b_sfull, (uint32_t)t & 1u);
            ptx::tc_fence_after();
            FT_TRACE(0, t, 1);
            uint32_t sr[COLS];
#pragma unroll
            for (int c = 0; c < COLS / 32; c++) ptx::tmem_ld32(t_row + half * COLS + c * 32, *reinterpret_cast<uint32_t (*)[32]>(&sr[c * 32]));
            ptx::tmem_ld_wait();
            ptx::tc_fence_before();
            __syncwarp();
            if (ptx::elect_one()) ptx::mbar_arrive(b_sfree);          // S(t) is in registers: the tensor core may overwrite it with S(t+1)
            __syncwarp();
            FT_TRACE(0, t, 2);
            float *sf = reinterpret_cast<float *>(sr);
            if ((t + 1) * FT_BN > p.Lkv) {                            // keys past the end of the image (zero-filled K rows): no weight
                const int valid = p.Lkv - t * FT_BN - half * COLS;
#pragma unroll
                for (int c = 0; c < COLS; c++)
                    if (c >= valid) sf[c] = -INFINITY;
            }
            // (few warps per scheduler: four independent chains for the max, or their latency is the critical path)
            float mx4[4] = {sf[0], sf[1], sf[2], sf[3]};
#pragma unroll
            for (int c = 4; c < COLS; c += 4) {
                mx4[0] = fmaxf(mx4[0], sf[c]); mx4[1] = fmaxf(mx4[1], sf[c + 1]); mx4[2] = fmaxf(mx4[2], sf[c + 2]); mx4[3] = fmaxf(mx4[3], sf[c + 3]);
            }
            float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
            if constexpr (SP == 2) {                                  // the row's other 64 columns belong to warp (warp ^ 4): exchange the maxima
                float *mb = maxbuf + (t & 1) * 256;                   // double-buffered by tile parity: one named barrier per tile is enough
                mb[half * 128 + lrow] = mx;
                asm volatile("bar.sync %0, 64;" ::"r"(1 + qw) : "memory");
                mx = fmaxf(mx, mb[(half ^ 1) * 128 + lrow]);
            }
            // reference max: raise it only when the running max is more than 2^8 above it (p <= 256 otherwise)
            const bool raise = (mx - m_ref) * sl > 8.0f;              // true at t = 0 (m_ref = -inf)
            const float m_new = raise ? mx : m_ref;
            // exponentials first (they need only m_new), packed to bf16 in place of the scores; the wait for P(t-1) V(t-1) comes after them,
            // when it has long completed.  (The row sum is column d_h of O: see the ones column of V above.  Packing by truncation -- a PRMT
            // instead of F2FP, which shares the SFU pipe with ex2 -- was measured: same time, slightly larger error; round-to-nearest kept.)
            const float ms = m_new * sl;
            const uint32_t prow = s_p + (uint32_t)lrow * 16 + (uint32_t)(half * (COLS / 8)) * FT_CHUNK;
            const uint64_t sl2 = f32x2_pack(sl, sl), nms2 = f32x2_pack(-ms, -ms);
            auto exp_chunk = [&](int kc, uint32_t (&q)[4]) {          // 8 scores -> 8 bf16 weights: 8 - FT_POLY on the SFU, FT_POLY on the FMA / ALU pipes
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const uint64_t x = f32x2_fma(f32x2_pack(sf[kc * 8 + 2 * i], sf[kc * 8 + 2 * i + 1]), sl2, nms2);      // (s - m_ref) * scale * log2 e, two at once
                    float x0, x1, e0, e1;
                    f32x2_unpack(x, x0, x1);
#ifdef NMM_TRACE
                    if (FT_DEBUG(1)) { q[i] = pack_bf16x2(x0, x1); continue; }
                    if (FT_DEBUG(8)) { q[i] = __byte_perm(__float_as_uint(ft_exp2(x0)), __float_as_uint(ft_exp2(x1)), 0x7632); continue; }
#endif
                    constexpr int FIRST = 8 - FT_POLY;                // elements FIRST .. 7 of every 8 take the polynomial
                    if (2 * i >= FIRST) ft_exp2_poly2(x, e0, e1);
                    else { e0 = ft_exp2(x0); e1 = (2 * i + 1 >= FIRST) ? ft_exp2_poly(x1) : ft_exp2(x1); }
                    q[i] = pack_bf16x2(e0, e1);
                }
            };
            auto store_chunk = [&](int kc, const uint32_t (&q)[4]) {
#ifdef NMM_TRACE
                if (FT_DEBUG(2)) { if (q[0] == 0x12345678u && q[1] == q[2] + q[3]) asm volatile("st.shared.b32 [%0], %1;" ::"r"(prow), "r"(q[3]) : "memory"); return; }
#endif
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(prow + kc * FT_CHUNK), "r"(q[0]), "r"(q[1]), "r"(q[2]), "r"(q[3]) : "memory");
            };
            // first half of the exponentials before the wait for P(t-1) V(t-1) (P's buffer is busy until then), the rest interleaved with
            // the stores: no burst of 16-byte stores + proxy fence at the end of the tile with the SFU idle behind it
            constexpr int NCH = COLS / 8, PRE = NCH / 2;
            uint32_t pk[PRE][4];
#pragma unroll
            for (int kc = 0; kc < PRE; kc++) exp_chunk(kc, pk[kc]);
            FT_TRACE(0, t, 3);
            if (t > 0) ptx::mbar_wait(b_pv, (uint32_t)(t - 1) & 1u);
            FT_TRACE(0, t, 4);  // P(t-1) V(t-1) done: P's buffer is free and O's rows are at rest
            if (t > 0 && __any_sync(0xffffffffu, raise) && half == 0) {   // (both threads of a row see the same maxima; the first one rescales)
                const float f = ft_exp2((m_ref - m_new) * sl);        // 1 for the rows that keep their reference
                ptx::tc_fence_after();
#pragma unroll
                for (int c = 0; c < NV; c += 16) {                    // O's row and its sum column
                    uint32_t orow[16];
                    ptx::tmem_ld16(t_row + FT_O_COL + c, orow);
                    ptx::tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 16; i++) orow[i] = __float_as_uint(__uint_as_float(orow[i]) * f);
                    ptx::tmem_st16(t_row + FT_O_COL + c, orow);
                }
                ptx::tmem_st_wait();
            }
            m_ref = m_new;
#pragma unroll
            for (int kc = 0; kc < PRE; kc++) {
                store_chunk(kc, pk[kc]);
                uint32_t q[4];
                exp_chunk(PRE + kc, q);
                store_chunk(PRE + kc, q);
            }
            FT_TRACE(0, t, 5);
            ptx::fence_proxy_async();          // generic-proxy stores -> visible to the tensor core's async-proxy reads
            ptx::tc_fence_before();
            __syncwarp();
            if (ptx::elect_one()) ptx::mbar_arrive(b_pfull);
            __syncwarp();
            FT_TRACE(0, t, 6);
        }
        // ---- O / l -> global (d_h contiguous bf16 per query row; the row's threads take alternate 16-channel pieces) ----
        ptx::mbar_wait(b_pv, (uint32_t)(nt - 1) & 1u);
        ptx::tc_fence_after();
        const uint32_t lraw = ptx::tmem_ld1(t_row + FT_O_COL + DH);      // the row sum: column d_h of O
        ptx::tmem_ld_wait();
        const float inv = 1.0f / __uint_as_float(lraw);
        const int row = q0 + lrow;
        bf16 *og = (bf16 *)p.o + (int64_t)img * p.o_bs + (int64_t)row * p.o_rs + head * DH;
#pragma unroll
        for (int c = 2 * half; c < CH; c += 2 * SP) {
            uint32_t orow[16];
            ptx::tmem_ld16(t_row + FT_O_COL + c * 8, orow);
            ptx::tmem_ld_wait();
            if (row < p.Lq) {
                *reinterpret_cast<uint4 *>(og + c * 8) = make_uint4(pack_bf16x2(__uint_as_float(orow[0]) * inv, __uint_as_float(orow[1]) * inv),
                                                                     pack_bf16x2(__uint_as_float(orow[2]) * inv, __uint_as_float(orow[3]) * inv),
                                                                     pack_bf16x2(__uint_as_float(orow[4]) * inv, __uint_as_float(orow[5]) * inv),
                                                                     pack_bf16x2(__uint_as_float(orow[6]) * inv, __uint_as_float(orow[7]) * inv));
                if (c + 1 < CH)
                    *reinterpret_cast<uint4 *>(og + c * 8 + 8) = make_uint4(pack_bf16x2(__uint_as_float(orow[8]) * inv, __uint_as_float(orow[9]) * inv),
                                                                             pack_bf16x2(__uint_as_float(orow[10]) * inv, __uint_as_float(orow[11]) * inv),
                                                                             pack_bf16x2(__uint_as_float(orow[12]) * inv, __uint_as_float(orow[13]) * inv),
                                                                             pack_bf16x2(__uint_as_float(orow[14]) * inv, __uint_as_float(orow[15]) * inv));
            }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == W_MMA) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc<1>(tmem, FT_TMEM_COLS);
    }
}

// =====================================================================================================================================
// Two query tiles per CTA (256 queries of one (image, head)), ONE CTA per SM: the same two softmax pipelines per SM as two CTAs of the
// kernel above, but they share every K / V tile -- half the TMA work per query (the K / V segment rate is what binds that kernel once the
// SFU is relieved) and room for two-stage K / V rings at d_h = 80 too (192 KB of shared memory, 448 of the 512 TMEM columns).
// Warps 0-3: softmax of query tile 0, warps 4-7: query tile 1 (warp w and w + 4 own the same TMEM lanes, different columns), warp 8: TMA,
// warp 9: MMA.  TMEM: S0 [0,128) S1 [128,256) O0 [256, 256+NV) O1 [384, 384+NV).  One softmax thread per row, 3 of 8 exponentials by polynomial.
template <int DH, int NS_>
struct Ft2Cfg {
    using B = FtCfg<DH>;
    static constexpr int NS = NS_;
    static constexpr uint32_t OFF_Q = 0, OFF_K = 2 * B::TILE_BYTES, OFF_V = OFF_K + NS * B::TILE_BYTES, OFF_P = OFF_V + NS * B::VTILE_BYTES;
    static constexpr uint32_t OFF_BAR = OFF_P + 2 * B::P_BYTES;
    static constexpr uint32_t SMEM = OFF_BAR + 256 + 128;
};

template <int DH, int NS_, bool NOKV = false, bool PT = false>
__global__ void __launch_bounds__(320, 1)
spatial_attention_tc2_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k, const __grid_constant__ CUtensorMap tm_v,
                             const FtParams p) {
    using Cfg = FtCfg<DH>;
    using C2 = Ft2Cfg<DH, NS_>;
    constexpr int DHP = Cfg::DHP, NV = Cfg::NV, CH = Cfg::CH, CHP = Cfg::CHP, CHV = Cfg::CHV, NS = C2::NS, FT_POLY = 3;
    constexpr int THREADS = 320, W_TMA = 8, W_MMA = 9;
    extern __shared__ uint8_t ft_smem_raw[];
    const uint32_t base = (ptx::smem_u32(ft_smem_raw) + 127u) & ~127u;
    uint8_t *gbase = ft_smem_raw + (base - ptx::smem_u32(ft_smem_raw));
    const uint32_t s_q = base + C2::OFF_Q, s_k = base + C2::OFF_K, s_v = base + C2::OFF_V, s_p = base + C2::OFF_P, s_bar = base + C2::OFF_BAR;
    // barriers (8 bytes each): per query tile sfull / sfree / pfull / pv; shared q, K / V rings
    const uint32_t b_q = s_bar;
    auto b_sfull = [&](int qt) { return s_bar + 8 + 8 * qt; };
    auto b_sfree = [&](int qt) { return s_bar + 24 + 8 * qt; };
    auto b_pfull = [&](int qt) { return s_bar + 40 + 8 * qt; };
    auto b_pv = [&](int qt) { return s_bar + 56 + 8 * qt; };
    auto b_kfull = [&](int s) { return s_bar + 72 + 8 * s; };        // up to 4 stages each
    auto b_kempty = [&](int s) { return s_bar + 104 + 8 * s; };
    auto b_vfull = [&](int s) { return s_bar + 136 + 8 * s; };
    auto b_vempty = [&](int s) { return s_bar + 168 + 8 * s; };
    volatile uint32_t *tmem_slot = reinterpret_cast<volatile uint32_t *>(gbase + C2::OFF_BAR + 240);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q0 = blockIdx.x * 2 * FT_BM, head = blockIdx.y, img = blockIdx.z;
    const int nt = (p.Lkv + FT_BN - 1) / FT_BN;
    const int kv_img = img / p.kv_div;

    if (tid == 0) {
        ptx::mbar_init(b_q, 1);
        for (int qt = 0; qt < 2; qt++) { ptx::mbar_init(b_sfull(qt), 1); ptx::mbar_init(b_sfree(qt), 4); ptx::mbar_init(b_pfull(qt), 4); ptx::mbar_init(b_pv(qt), 1); }
        for (int s = 0; s < NS; s++) { ptx::mbar_init(b_kfull(s), 1); ptx::mbar_init(b_kempty(s), 1); ptx::mbar_init(b_vfull(s), 1); ptx::mbar_init(b_vempty(s), 1); }
        ptx::fence_mbar_init();
    }
    // pad chunks: zeros for Q (2 tiles) and K (NS stages), the ones column for V (see the one-tile kernel)
    if constexpr (CHP > CH) {
        for (int i = tid; i < (2 + NS) * (int)(FT_CHUNK / 16); i += THREADS) {
            const int tile = i / (int)(FT_CHUNK / 16), r = i - tile * (int)(FT_CHUNK / 16);
            *reinterpret_cast<uint4 *>(gbase + tile * Cfg::TILE_BYTES + CH * FT_CHUNK + r * 16) = make_uint4(0u, 0u, 0u, 0u);
        }
    }
    for (int i = tid; i < NS * (CHV - CH) * (int)(FT_CHUNK / 16); i += THREADS) {
        const int per = (CHV - CH) * (int)(FT_CHUNK / 16);
        const int st = i / per, r = i - st * per;
        *reinterpret_cast<uint4 *>(gbase + C2::OFF_V + st * Cfg::VTILE_BYTES + CH * FT_CHUNK + r * 16) = make_uint4(r < FT_BN ? 0x3F80u : 0u, 0u, 0u, 0u);
    }
    ptx::fence_proxy_async();
    if (warp == W_MMA) ptx::tmem_alloc<1>(ptx::smem_u32(const_cast<uint32_t *>(tmem_slot)), 512);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    pdl_wait();
    pdl_launch_dependents();

    if (warp == W_TMA) {
        if (lane == 0) {
            ptx::prefetch_tensormap(&tm_q); ptx::prefetch_tensormap(&tm_k);
            ptx::mbar_expect_tx(b_q, 2 * Cfg::TX_BYTES);
            ptx::tma_load_4d(&tm_q, b_q, s_q, 0, q0, head * CH, img);
            ptx::tma_load_4d(&tm_q, b_q, s_q + Cfg::TILE_BYTES, 0, q0 + FT_BM, head * CH, img);       // rows past the image: zero-filled
            for (int t = 0; t < nt; t++) {
                const int s = t % NS, fill = t / NS;
                if (fill > 0) ptx::mbar_wait(b_kempty(s), (uint32_t)(fill - 1) & 1u);
                if (NOKV && fill > 0) { ptx::mbar_arrive(b_kfull(s)); continue; }          // timing experiment (results invalid)
                ptx::mbar_expect_tx(b_kfull(s), Cfg::TX_BYTES);
                ptx::tma_load_4d(&tm_k, b_kfull(s), s_k + s * Cfg::TILE_BYTES, 0, t * FT_BN, head * CH, kv_img);
            }
        } else if (lane == 1) {
            ptx::prefetch_tensormap(&tm_v);
            for (int t = 0; t < nt; t++) {
                const int s = t % NS, fill = t / NS;
                if (fill > 0) ptx::mbar_wait(b_vempty(s), (uint32_t)(fill - 1) & 1u);
                if (NOKV && fill > 0) { ptx::mbar_arrive(b_vfull(s)); continue; }
                ptx::mbar_expect_tx(b_vfull(s), Cfg::TX_BYTES);
                ptx::tma_load_4d(&tm_v, b_vfull(s), s_v + s * Cfg::VTILE_BYTES, 0, t * FT_BN, head * CH, kv_img);
            }
        }
        __syncwarp();
    } else if (warp == W_MMA) {
        const uint32_t idesc_s = ptx::umma_idesc_bf16(FT_BM, FT_BN);
        const uint32_t idesc_o = ptx::umma_idesc_bf16(FT_BM, NV) | ptx::UMMA_IDESC_B_MN_MAJOR;
        auto issue_s = [&](int t) {           // S_qt(t) = Q_qt K(t)^T for both query tiles; K(t)'s buffer is released after the second
            const int s = t % NS;
            ptx::mbar_wait(b_kfull(s), (uint32_t)(t / NS) & 1u);
            for (int qt = 0; qt < 2; qt++) {
                if (t > 0) ptx::mbar_wait(b_sfree(qt), (uint32_t)(t - 1) & 1u);
                ptx::tc_fence_after();
                if (ptx::elect_one()) {
                    const uint64_t a = ptx::umma_smem_desc_interleave(s_q + qt * Cfg::TILE_BYTES, FT_CHUNK, 128);
                    const uint64_t b = ptx::umma_smem_desc_interleave(s_k + s * Cfg::TILE_BYTES, FT_CHUNK, 128);
#pragma unroll
                    for (int k = 0; k < DHP / 16; k++)
                        ptx::umma_bf16<1>(tmem + qt * 128, a + (uint64_t)((k * 2 * FT_CHUNK) >> 4), b + (uint64_t)((k * 2 * FT_CHUNK) >> 4), idesc_s, k != 0 ? 1u : 0u);
                    ptx::umma_commit<1>(b_sfull(qt));
                    if (qt == 1) ptx::umma_commit<1>(b_kempty(s));
                }
                __syncwarp();
            }
        };
        ptx::mbar_wait(b_q, 0);
        issue_s(0);
        for (int t = 0; t < nt; t++) {
            if (t + 1 < nt) issue_s(t + 1);
            const int s = t % NS;
            ptx::mbar_wait(b_vfull(s), (uint32_t)(t / NS) & 1u);
            for (int qt = 0; qt < 2; qt++) {
                ptx::mbar_wait(b_pfull(qt), (uint32_t)t & 1u);
                ptx::tc_fence_after();
                if (ptx::elect_one()) {          // O_qt += P_qt(t) V(t)
                    const uint64_t a = ptx::umma_smem_desc_interleave(s_p + qt * Cfg::P_BYTES, FT_CHUNK, 128);
                    const uint64_t b = ptx::umma_smem_desc_interleave(s_v + s * Cfg::VTILE_BYTES, 128u, FT_CHUNK);
#pragma unroll
                    for (int k = 0; k < FT_BN / 16; k++) {
                        if constexpr (PT)          // P(t) sits in tensor memory (64 columns behind O_qt: two bf16 keys per column)
                            ptx::umma_bf16_ts(tmem + 256 + qt * 128, tmem + 256 + qt * 128 + 64 + k * 8, b + (uint64_t)((k * 16 * 16) >> 4), idesc_o, (t | k) != 0 ? 1u : 0u);
                        else
                            ptx::umma_bf16<1>(tmem + 256 + qt * 128, a + (uint64_t)((k * 2 * FT_CHUNK) >> 4), b + (uint64_t)((k * 16 * 16) >> 4), idesc_o, (t | k) != 0 ? 1u : 0u);
                    }
                    ptx::umma_commit<1>(b_pv(qt));
                    if (qt == 1) ptx::umma_commit<1>(b_vempty(s));
                }
                __syncwarp();
            }
        }
    } else {
        // ===================== softmax of query tile qt: thread = query row = TMEM lane =====================
        const int qt = warp >> 2, qw = warp & 3, lrow = qw * 32 + lane;
        const uint32_t t_s = tmem + ((uint32_t)(qw * 32) << 16) + qt * 128, t_o = tmem + ((uint32_t)(qw * 32) << 16) + 256 + qt * 128;
        const uint32_t prow = s_p + qt * Cfg::P_BYTES + (uint32_t)lrow * 16;
        const float sl = p.scale_log2e;
        float m_ref = -INFINITY;
        for (int t = 0; t < nt; t++) {
            ptx::mbar_wait(b_sfull(qt), (uint32_t)t & 1u);
            ptx::tc_fence_after();
            uint32_t sr[FT_BN];
#pragma unroll
            for (int c = 0; c < FT_BN / 32; c++) ptx::tmem_ld32(t_s + c * 32, *reinterpret_cast<uint32_t (*)[32]>(&sr[c * 32]));
            ptx::tmem_ld_wait();
            ptx::tc_fence_before();
            __syncwarp();
            if (ptx::elect_one()) ptx::mbar_arrive(b_sfree(qt));
            __syncwarp();
            float *sf = reinterpret_cast<float *>(sr);
            if ((t + 1) * FT_BN > p.Lkv) {
                const int valid = p.Lkv - t * FT_BN;
#pragma unroll
                for (int c = 0; c < FT_BN; c++)
                    if (c >= valid) sf[c] = -INFINITY;
            }
            float mx4[4] = {sf[0], sf[1], sf[2], sf[3]};
#pragma unroll
            for (int c = 4; c < FT_BN; c += 4) {
                mx4[0] = fmaxf(mx4[0], sf[c]); mx4[1] = fmaxf(mx4[1], sf[c + 1]); mx4[2] = fmaxf(mx4[2], sf[c + 2]); mx4[3] = fmaxf(mx4[3], sf[c + 3]);
            }
            const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
            const bool raise = (mx - m_ref) * sl > 8.0f;
            const float m_new = raise ? mx : m_ref;
            const float ms = m_new * sl;
            const uint64_t sl2 = f32x2_pack(sl, sl), nms2 = f32x2_pack(-ms, -ms);
            auto exp_chunk = [&](int kc, uint32_t (&q)[4]) {
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const uint64_t x = f32x2_fma(f32x2_pack(sf[kc * 8 + 2 * i], sf[kc * 8 + 2 * i + 1]), sl2, nms2);
                    float x0, x1, e0, e1;
                    f32x2_unpack(x, x0, x1);
                    constexpr int FIRST = 8 - FT_POLY;
                    if (2 * i >= FIRST) ft_exp2_poly2(x, e0, e1);
                    else { e0 = ft_exp2(x0); e1 = (2 * i + 1 >= FIRST) ? ft_exp2_poly(x1) : ft_exp2(x1); }
                    q[i] = pack_bf16x2(e0, e1);
                }
            };
            auto store_chunk = [&](int kc, const uint32_t (&q)[4]) {
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(prow + kc * FT_CHUNK), "r"(q[0]), "r"(q[1]), "r"(q[2]), "r"(q[3]) : "memory");
            };
            constexpr int NCH = FT_BN / 8, PRE = NCH / 2;
            if constexpr (PT) {
                // P through tensor memory: all 64 bf16x2 columns in registers, then two tcgen05.st of 32 columns into [O_qt + 64, O_qt + 128)
                uint32_t pk2[2][32];
#pragma unroll
                for (int kc = 0; kc < NCH; kc++) {
                    uint32_t q[4];
                    exp_chunk(kc, q);
#pragma unroll
                    for (int i = 0; i < 4; i++) pk2[kc / 8][(kc % 8) * 4 + i] = q[i];
                }
                if (t > 0) ptx::mbar_wait(b_pv(qt), (uint32_t)(t - 1) & 1u);
                ptx::tc_fence_after();
                if (t > 0 && __any_sync(0xffffffffu, raise)) {
                    const float f = ft_exp2((m_ref - m_new) * sl);
#pragma unroll
                    for (int c = 0; c < NV; c += 16) {
                        uint32_t orow[16];
                        ptx::tmem_ld16(t_o + c, orow);
                        ptx::tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 16; i++) orow[i] = __float_as_uint(__uint_as_float(orow[i]) * f);
                        ptx::tmem_st16(t_o + c, orow);
                    }
                }
                m_ref = m_new;
                ptx::tmem_st32(t_o + 64, pk2[0]);
                ptx::tmem_st32(t_o + 96, pk2[1]);
                ptx::tmem_st_wait();
            } else {
            uint32_t pk[PRE][4];
#pragma unroll
            for (int kc = 0; kc < PRE; kc++) exp_chunk(kc, pk[kc]);
            if (t > 0) ptx::mbar_wait(b_pv(qt), (uint32_t)(t - 1) & 1u);
            if (t > 0 && __any_sync(0xffffffffu, raise)) {
                const float f = ft_exp2((m_ref - m_new) * sl);
                ptx::tc_fence_after();
#pragma unroll
                for (int c = 0; c < NV; c += 16) {
                    uint32_t orow[16];
                    ptx::tmem_ld16(t_o + c, orow);
                    ptx::tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 16; i++) orow[i] = __float_as_uint(__uint_as_float(orow[i]) * f);
                    ptx::tmem_st16(t_o + c, orow);
                }
                ptx::tmem_st_wait();
            }
            m_ref = m_new;
#pragma unroll
            for (int kc = 0; kc < PRE; kc++) {
                store_chunk(kc, pk[kc]);
                uint32_t q[4];
                exp_chunk(PRE + kc, q);
                store_chunk(PRE + kc, q);
            }
            }
            ptx::fence_proxy_async();
            ptx::tc_fence_before();
            __syncwarp();
            if (ptx::elect_one()) ptx::mbar_arrive(b_pfull(qt));
            __syncwarp();
        }
        ptx::mbar_wait(b_pv(qt), (uint32_t)(nt - 1) & 1u);
        ptx::tc_fence_after();
        const uint32_t lraw = ptx::tmem_ld1(t_o + DH);
        ptx::tmem_ld_wait();
        const float inv = 1.0f / __uint_as_float(lraw);
        const int row = q0 + qt * FT_BM + lrow;
        bf16 *og = (bf16 *)p.o + (int64_t)img * p.o_bs + (int64_t)row * p.o_rs + head * DH;
#pragma unroll
        for (int c = 0; c < CH; c += 2) {
            uint32_t orow[16];
            ptx::tmem_ld16(t_o + c * 8, orow);
            ptx::tmem_ld_wait();
            if (row < p.Lq) {
                *reinterpret_cast<uint4 *>(og + c * 8) = make_uint4(pack_bf16x2(__uint_as_float(orow[0]) * inv, __uint_as_float(orow[1]) * inv),
                                                                     pack_bf16x2(__uint_as_float(orow[2]) * inv, __uint_as_float(orow[3]) * inv),
                                                                     pack_bf16x2(__uint_as_float(orow[4]) * inv, __uint_as_float(orow[5]) * inv),
                                                                     pack_bf16x2(__uint_as_float(orow[6]) * inv, __uint_as_float(orow[7]) * inv));
                if (c + 1 < CH)
                    *reinterpret_cast<uint4 *>(og + c * 8 + 8) = make_uint4(pack_bf16x2(__uint_as_float(orow[8]) * inv, __uint_as_float(orow[9]) * inv),
                                                                             pack_bf16x2(__uint_as_float(orow[10]) * inv, __uint_as_float(orow[11]) * inv),
                                                                             pack_bf16x2(__uint_as_float(orow[12]) * inv, __uint_as_float(orow[13]) * inv),
                                                                             pack_bf16x2(__uint_as_float(orow[14]) * inv, __uint_as_float(orow[15]) * inv));
            }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == W_MMA) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc<1>(tmem, 512);
    }
}

// ---- host side ----------------------------------------------------------------------------------------------------------------------
typedef CUresult (*FtEncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                    const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static FtEncodeTiledFn ft_encode_fn() {
    static FtEncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void *ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<FtEncodeTiledFn>(ptr);
    });
    return fn;
}
// rows of `row_elems` bf16 viewed as (8 channels, rows of one image, 16-byte chunks of a row, images): box (8, 128, d_h / 8, 1)
static int ft_map(CUtensorMap *tm, const void *ptr, int64_t rows, int64_t row_stride, int64_t image_stride, int64_t images, int dh) {
    FtEncodeTiledFn fn = ft_encode_fn();
    if (!fn) return fail(NMM_ERR_DEVICE, "cuTensorMapEncodeTiled not available (no CUDA driver?)");
    cuuint64_t dims[4] = {8, (cuuint64_t)rows, (cuuint64_t)(row_stride / 8), (cuuint64_t)images};
    cuuint64_t strides[3] = {(cuuint64_t)row_stride * 2, 16, (cuuint64_t)image_stride * 2};
    cuuint32_t box[4] = {8, FT_BN, (cuuint32_t)(dh / 8), 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void *>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(NMM_ERR_CUDA, "cuTensorMapEncodeTiled (spatial attention operand) failed with CUresult %d", (int)r);
    return NMM_OK;
}

#ifdef NMM_TRACE
static unsigned long long *g_ft_trace = nullptr;
extern "C" __attribute__((visibility("default"))) int nmm_debug_ft_trace(unsigned long long *host_out) {      // 2 x 16 x 8 clock64 values
    if (!g_ft_trace) { cudaMalloc(&g_ft_trace, 2 * 16 * 8 * 8); cudaMemset(g_ft_trace, 0, 2 * 16 * 8 * 8); return 1; }
    cudaDeviceSynchronize();
    cudaMemcpy(host_out, g_ft_trace, 2 * 16 * 8 * 8, cudaMemcpyDeviceToHost);
    return 0;
}
#endif

bool spatial_attention_tc_eligible(const FlashArgs &a) {
    if (a.dtype != NMM_BF16 || (a.dh != 40 && a.dh != 80) || a.images % a.kv_div != 0) return false;
    // self-attention from 256 keys up; the 77-key text cross-attention (kv_div > 1) only behind NMM_OPT_SPATIAL_ATTN = 24 (A/B)
    if (a.kv_div != 1 || a.Lkv < 256) { if (opt(NMM_OPT_SPATIAL_ATTN) != 24) return false; }
    // the tensor maps address whole rows: a row stride that covers the row's channels and image strides that are whole rows
    return a.q_rs % 8 == 0 && a.kv_rs % 8 == 0 && a.q_bs % 8 == 0 && a.kv_bs % 8 == 0 && a.o_rs % 8 == 0 && a.o_bs % 8 == 0 && aligned(a.q, 16) &&
           aligned(a.k, 16) && aligned(a.v, 16) && aligned(a.o, 16);
}

template <int DH, int SP, int PL, bool NOKV = false, bool CPA = false>
static int launch_ft(const FlashArgs &a, int debug, cudaStream_t st) {
    using Cfg = FtCfg<DH>;
    static DeviceOnce once;
    NMM_CUDA_OK(once.max_smem(spatial_attention_tc_kernel<DH, SP, PL, NOKV, CPA>, (int)Cfg::SMEM));
    CUtensorMap tq, tk, tv;
    int rc;
    // q / k / v are column slices of a wider row (q | k | v of one projection): the map starts at the slice, the chunk dimension spans the
    // rest of the row (the kernel only ever asks for the head's d_h / 8 chunks)
    if ((rc = ft_map(&tq, a.q, a.Lq, a.q_rs, a.q_bs, a.images, DH)) != NMM_OK) return rc;
    if ((rc = ft_map(&tk, a.k, a.Lkv, a.kv_rs, a.kv_bs, a.images, DH)) != NMM_OK) return rc;
    if ((rc = ft_map(&tv, a.v, a.Lkv, a.kv_rs, a.kv_bs, a.images, DH)) != NMM_OK) return rc;
    FtParams p;
    p.o = a.o; p.o_rs = a.o_rs; p.o_bs = a.o_bs; p.k = a.k; p.v = a.v; p.kv_rs = a.kv_rs; p.kv_bs = a.kv_bs; p.Lq = a.Lq; p.Lkv = a.Lkv; p.kv_div = 1; p.scale_log2e = a.scale_log2e;
#ifdef NMM_TRACE
    p.debug = debug & 15; p.trace = g_ft_trace;
#else
    (void)debug;
#endif
    const dim3 grid((unsigned)ceil_div(a.Lq, FT_BM), (unsigned)a.heads, (unsigned)a.images);
    const double per = (double)a.images * a.heads;
    ProfScope prof(K_SPATIAL_ATTN, st, 4.0 * per * a.Lq * (double)a.Lkv * DH, 2.0 * 4.0 * per * a.Lq * DH);
    NMM_CUDA_OK(launch_pdl(spatial_attention_tc_kernel<DH, SP, PL, NOKV, CPA>, grid, dim3(64 + 128 * SP), (size_t)Cfg::SMEM, st, tq, tk, tv, p));
    NMM_LAUNCHED("spatial_attention_tc_kernel");
    return NMM_OK;
}

template <int DH, int NS, bool NOKV = false, bool PT = false>
static int launch_ft2(const FlashArgs &a, cudaStream_t st) {
    using C2 = Ft2Cfg<DH, NS>;
    static_assert(!PT || DH == 40, "P through TMEM: 2 x (128 S + 48 O + 64 P) columns only fit at d_h = 40");
    static DeviceOnce once;
    NMM_CUDA_OK(once.max_smem(spatial_attention_tc2_kernel<DH, NS, NOKV, PT>, (int)C2::SMEM));
    CUtensorMap tq, tk, tv;
    int rc;
    if ((rc = ft_map(&tq, a.q, a.Lq, a.q_rs, a.q_bs, a.images, DH)) != NMM_OK) return rc;
    if ((rc = ft_map(&tk, a.k, a.Lkv, a.kv_rs, a.kv_bs, a.images / a.kv_div, DH)) != NMM_OK) return rc;
    if ((rc = ft_map(&tv, a.v, a.Lkv, a.kv_rs, a.kv_bs, a.images / a.kv_div, DH)) != NMM_OK) return rc;
    FtParams p;
    memset(&p, 0, sizeof(p));
    p.o = a.o; p.o_rs = a.o_rs; p.o_bs = a.o_bs; p.Lq = a.Lq; p.Lkv = a.Lkv; p.kv_div = a.kv_div; p.scale_log2e = a.scale_log2e;
    const dim3 grid((unsigned)ceil_div(a.Lq, 2 * FT_BM), (unsigned)a.heads, (unsigned)a.images);
    const double per = (double)a.images * a.heads;
    ProfScope prof(K_SPATIAL_ATTN, st, 4.0 * per * a.Lq * (double)a.Lkv * DH, 2.0 * 4.0 * per * a.Lq * DH);
    NMM_CUDA_OK(launch_pdl(spatial_attention_tc2_kernel<DH, NS, NOKV, PT>, grid, dim3(320), (size_t)C2::SMEM, st, tq, tk, tv, p));
    NMM_LAUNCHED("spatial_attention_tc2_kernel");
    return NMM_OK;
}

int launch_spatial_attention_tc(const FlashArgs &a, int variant, cudaStream_t st) {
    // default: two query tiles per CTA (one CTA per SM; K / V tiles shared by the two softmax pipelines; 3 K / V stages at d_h = 40, 2 at 80).  20 / 21 / 22: force 2 / 3 / 4 K / V stages
    // (d_h = 80: 2 only); 23: the one-tile kernel (two CTAs per SM) that the variants below select explicitly
    if (variant == 0 || variant == 21 || variant == 24) return a.dh == 40 ? launch_ft2<40, 3>(a, st) : launch_ft2<80, 2>(a, st);
    if (variant == 20) return a.dh == 40 ? launch_ft2<40, 2>(a, st) : launch_ft2<80, 2>(a, st);
    if (variant == 27) return a.dh == 40 ? launch_ft2<40, 3, false, true>(a, st) : launch_ft2<80, 2>(a, st);      // A/B: P through tensor memory (d_h = 40)
    if (variant == 26) return a.dh == 40 ? launch_ft2<40, 3, true>(a, st) : launch_ft2<80, 2, true>(a, st);      // timing only: no K / V traffic
    if (variant == 22) return a.dh == 40 ? launch_ft2<40, 4>(a, st) : launch_ft2<80, 2>(a, st);
    // NMM_OPT_SPATIAL_ATTN: 0 / 3 = one softmax thread per query row (measured faster once the polynomial share relieved the SFU), 2 = two;
    // 10..14 = polynomial share 0 / 2 / 3 / 4 / 6 of 8 (A/B; 3 is the default);
    // 100 + bits = timing experiments of the -DNMM_TRACE development build (results invalid; ignored by the production build)
    const int debug = variant >= 100 ? variant - 100 : 0;
    const int sp = variant == 2 ? 2 : 1;      // (variant 3 / 23: one thread per row, one query tile per CTA)
    if (variant == 16) return a.dh == 40 ? launch_ft<40, 1, 3, false, true>(a, 0, st) : launch_ft<80, 1, 3, false, true>(a, 0, st);     // A/B: cp.async K / V loader
    if (variant == 17) return a.dh == 40 ? launch_ft<40, 1, 3, true, true>(a, 0, st) : launch_ft<80, 1, 3, true, true>(a, 0, st);
    if (variant == 15) return a.dh == 40 ? launch_ft<40, 1, 3, true>(a, 0, st) : launch_ft<80, 1, 3, true>(a, 0, st);      // timing experiment: no K / V traffic
    if (variant >= 10 && variant <= 14) {          // A/B: polynomial share 0 / 2 / 3 / 4 / 6 of 8 (one softmax thread per row)
        const int pl = variant - 10;
        if (a.dh == 40) return pl == 0 ? launch_ft<40, 1, 0>(a, 0, st) : pl == 1 ? launch_ft<40, 1, 2>(a, 0, st) : pl == 2 ? launch_ft<40, 1, 3>(a, 0, st) : pl == 3 ? launch_ft<40, 1, 4>(a, 0, st) : launch_ft<40, 1, 6>(a, 0, st);
        if (a.dh == 80) return pl == 0 ? launch_ft<80, 1, 0>(a, 0, st) : pl == 1 ? launch_ft<80, 1, 2>(a, 0, st) : pl == 2 ? launch_ft<80, 1, 3>(a, 0, st) : pl == 3 ? launch_ft<80, 1, 4>(a, 0, st) : launch_ft<80, 1, 6>(a, 0, st);
    }
    if (a.dh == 40) return sp == 1 ? launch_ft<40, 1, 3>(a, debug, st) : launch_ft<40, 2, 3>(a, debug, st);
    if (a.dh == 80) return sp == 1 ? launch_ft<80, 1, 3>(a, debug, st) : launch_ft<80, 2, 3>(a, debug, st);
    return fail(NMM_ERR_UNSUPPORTED, "tcgen05 spatial attention: d_h %d", a.dh);
}

}  // namespace nmm
