// Spatial self-attention on the 5th-generation tensor cores (tcgen05 + TMEM + TMA) -- the long-sequence levels of the UNet's
// Transformer3DModel blocks (arithmetic: CrossAttention._attention, motion_module_new.py:258-287):
//     O = softmax(Q K^T * d_h^-1/2) V       per (image, head),  Q [Lq, d_h], K / V [Lkv, d_h],  d_h in {40, 80},  Lkv >= 256
// (the mma.sync kernel of spatial_attention.cu keeps d_h = 160, the 77-token cross-attention and the small levels: its legacy HMMA pipe
// saturates at ~260 TFLOP/s, profiles/r2_ncu_spatial_attention.txt).
//
// One CTA = 128 queries of one (image, head); 192 threads = 4 softmax warps (one query row per thread = one TMEM lane), 1 TMA warp,
// 1 MMA warp; 2 CTAs per SM (256 TMEM columns, <= 92 KB shared memory each), so one CTA's tensor-core phases run under the other's
// exponentials.  Per 128-key tile t:
//   TMA warp   K(t), V(t) -> shared memory as [16-byte channel chunk][key][8 channels]: a 4-D tensor map (8, rows, chunks, images) whose box
//              (8, 128, d_h/8, 1) lands exactly in the un-swizzled core-matrix order tcgen05 reads (K-major for Q / K, N-major for V);
//              rows past the end of the image are zero-filled by the hardware
//   MMA warp   S = Q K(t)^T  (M128 N128 K16 x d_h/16, d_h = 40 zero-padded to 48) -> TMEM columns [0, 128); issued as soon as the softmax
//              warps have READ S(t-1) into registers, i.e. under their exponentials;   O += P(t) V(t)  (M128 N48|80 K16 x 8) -> TMEM columns
//              [128, 128 + d_h): O accumulates in tensor memory over ALL tiles
//   softmax    tcgen05.ld of the thread's 128 scores, row max, p = 2^((s - m_ref) * scale*log2e) in fp32 (ex2.approx), bf16 P written to
//              shared memory in the K-major operand order ([key chunk][row][16 B]: conflict-free 16-byte stores).  m_ref is the row's
//              REFERENCE max, raised only when the running max exceeds it by more than 2^8 (then O's row in TMEM and the running sum are
//              rescaled once: rare after the first tiles); p <= 256 is exact enough in bf16 and the final O / l is independent of m_ref.
// Nothing but Q, K, V is read from and O written to HBM; the score matrix (537 MB per frame in the reference at the 64 x 64 level) never exists.
#include <mutex>

#include "common.cuh"
#include "ptx.cuh"

namespace nmm {

constexpr int FT_BM = 128;          // queries per CTA
constexpr int FT_BN = 128;          // keys per tile
constexpr int FT_THREADS = 192;
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
    static constexpr uint32_t SMEM = OFF_BAR + 256 + 128;   // + alignment slack
};

struct FtParams {
    void *o;
    int64_t o_rs, o_bs;
    int Lq, Lkv;
    float scale_log2e;
    int swap_v_desc;            // development: LBO / SBO of the N-major V descriptor exchanged
};

__device__ __forceinline__ float ft_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <int DH>
__global__ void __launch_bounds__(FT_THREADS, 2)
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

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q0 = blockIdx.x * FT_BM, head = blockIdx.y, img = blockIdx.z;
    const int nt = (p.Lkv + FT_BN - 1) / FT_BN;

    if (tid == 0) {
        ptx::mbar_init(b_q, 1); ptx::mbar_init(b_sfull, 1); ptx::mbar_init(b_sfree, 4); ptx::mbar_init(b_pfull, 4); ptx::mbar_init(b_pv, 1);
        for (int s = 0; s < NS; s++) { ptx::mbar_init(b_kfull(s), 1); ptx::mbar_init(b_kempty(s), 1); ptx::mbar_init(b_vfull(s), 1); ptx::mbar_init(b_vempty(s), 1); }
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
    if (warp == 5) ptx::tmem_alloc<1>(ptx::smem_u32(const_cast<uint32_t *>(tmem_slot)), FT_TMEM_COLS);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    pdl_wait();
    pdl_launch_dependents();

    if (warp == 4) {
        // ===================== TMA producer =====================
        if (ptx::elect_one()) {
            ptx::prefetch_tensormap(&tm_q); ptx::prefetch_tensormap(&tm_k); ptx::prefetch_tensormap(&tm_v);
            ptx::mbar_expect_tx(b_q, Cfg::TX_BYTES);
            ptx::tma_load_4d(&tm_q, b_q, s_q, 0, q0, head * CH, img);
            for (int t = 0; t < nt; t++) {
                const int s = t % NS, fill = t / NS;
                if (fill > 0) ptx::mbar_wait(b_kempty(s), (uint32_t)(fill - 1) & 1u);
                ptx::mbar_expect_tx(b_kfull(s), Cfg::TX_BYTES);
                ptx::tma_load_4d(&tm_k, b_kfull(s), s_k + s * Cfg::TILE_BYTES, 0, t * FT_BN, head * CH, img);
                if (fill > 0) ptx::mbar_wait(b_vempty(s), (uint32_t)(fill - 1) & 1u);
                ptx::mbar_expect_tx(b_vfull(s), Cfg::TX_BYTES);
                ptx::tma_load_4d(&tm_v, b_vfull(s), s_v + s * Cfg::VTILE_BYTES, 0, t * FT_BN, head * CH, img);
            }
        }
        __syncwarp();
    } else if (warp == 5) {
        // ===================== MMA issuer =====================
        const uint32_t idesc_s = ptx::umma_idesc_bf16(FT_BM, FT_BN);
        const uint32_t idesc_o = ptx::umma_idesc_bf16(FT_BM, NV) | ptx::UMMA_IDESC_B_MN_MAJOR;
        const uint32_t v_lbo = p.swap_v_desc ? FT_CHUNK : 128u, v_sbo = p.swap_v_desc ? 128u : FT_CHUNK;
        auto issue_s = [&](int t) {           // S(t) = Q K(t)^T
            const int s = t % NS;
            ptx::mbar_wait(b_kfull(s), (uint32_t)(t / NS) & 1u);
            if (t > 0) ptx::mbar_wait(b_sfree, (uint32_t)(t - 1) & 1u);      // the softmax warps hold S(t-1) in registers
            ptx::tc_fence_after();
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
            ptx::mbar_wait(b_vfull(s), (uint32_t)(t / NS) & 1u);
            ptx::mbar_wait(b_pfull, (uint32_t)t & 1u);
            ptx::tc_fence_after();
            if (ptx::elect_one()) {          // O += P(t) V(t)
                const uint64_t a = ptx::umma_smem_desc_interleave(s_p, FT_CHUNK, 128);
                const uint64_t b = ptx::umma_smem_desc_interleave(s_v + s * Cfg::VTILE_BYTES, v_lbo, v_sbo);
#pragma unroll
                for (int k = 0; k < FT_BN / 16; k++)
                    ptx::umma_bf16<1>(tmem + FT_O_COL, a + (uint64_t)((k * 2 * FT_CHUNK) >> 4), b + (uint64_t)((k * 16 * 16) >> 4), idesc_o, (t | k) != 0 ? 1u : 0u);
                ptx::umma_commit<1>(b_pv);
                ptx::umma_commit<1>(b_vempty(s));
            }
            __syncwarp();
        }
    } else {
        // ===================== softmax: thread = query row q0 + tid = TMEM lane tid =====================
        const uint32_t t_row = tmem + ((uint32_t)(warp * 32) << 16);
        const float sl = p.scale_log2e;
        float m_ref = -INFINITY;
        for (int t = 0; t < nt; t++) {
            ptx::mbar_wait(b_sfull, (uint32_t)t & 1u);
            ptx::tc_fence_after();
            uint32_t sr[128];
#pragma unroll
            for (int c = 0; c < 4; c++) ptx::tmem_ld32(t_row + c * 32, *reinterpret_cast<uint32_t (*)[32]>(&sr[c * 32]));
            ptx::tmem_ld_wait();
            ptx::tc_fence_before();
            __syncwarp();
            if (ptx::elect_one()) ptx::mbar_arrive(b_sfree);          // S(t) is in registers: the tensor core may overwrite it with S(t+1)
            __syncwarp();
            float *sf = reinterpret_cast<float *>(sr);
            if ((t + 1) * FT_BN > p.Lkv) {                            // keys past the end of the image (zero-filled K rows): no weight
                const int valid = p.Lkv - t * FT_BN;
#pragma unroll
                for (int c = 0; c < 128; c++)
                    if (c >= valid) sf[c] = -INFINITY;
            }
            // (one warp per scheduler per CTA: four independent chains for the max and for the sum, or their latency is the critical path)
            float mx4[4] = {sf[0], sf[1], sf[2], sf[3]};
#pragma unroll
            for (int c = 4; c < 128; c += 4) {
                mx4[0] = fmaxf(mx4[0], sf[c]); mx4[1] = fmaxf(mx4[1], sf[c + 1]); mx4[2] = fmaxf(mx4[2], sf[c + 2]); mx4[3] = fmaxf(mx4[3], sf[c + 3]);
            }
            const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
            // reference max: raise it only when the running max is more than 2^8 above it (p <= 256 otherwise)
            const bool raise = (mx - m_ref) * sl > 8.0f;              // true at t = 0 (m_ref = -inf)
            const float m_new = raise ? mx : m_ref;
            // exponentials first (they need only m_new), packed to bf16 in place of the scores; the wait for P(t-1) V(t-1) comes after them,
            // when it has long completed.  (The row sum is column d_h of O: see the ones column of V above.  Packing by truncation -- a PRMT
            // instead of F2FP, which shares the SFU pipe with ex2 -- was measured: same time, slightly larger error; round-to-nearest kept.)
            const float ms = m_new * sl;
            uint32_t pk[64];
#pragma unroll
            for (int c = 0; c < 64; c++) pk[c] = pack_bf16x2(ft_exp2(fmaf(sf[2 * c], sl, -ms)), ft_exp2(fmaf(sf[2 * c + 1], sl, -ms)));
            if (t > 0) ptx::mbar_wait(b_pv, (uint32_t)(t - 1) & 1u);  // P(t-1) V(t-1) done: P's buffer is free and O's rows are at rest
            if (t > 0 && __any_sync(0xffffffffu, raise)) {
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
            const uint32_t prow = s_p + (uint32_t)tid * 16;
#pragma unroll
            for (int kc = 0; kc < FT_BN / 8; kc++)
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(prow + kc * FT_CHUNK), "r"(pk[4 * kc]), "r"(pk[4 * kc + 1]), "r"(pk[4 * kc + 2]),
                             "r"(pk[4 * kc + 3])
                             : "memory");
            ptx::fence_proxy_async();          // generic-proxy stores -> visible to the tensor core's async-proxy reads
            ptx::tc_fence_before();
            __syncwarp();
            if (ptx::elect_one()) ptx::mbar_arrive(b_pfull);
            __syncwarp();
        }
        // ---- O / l -> global (one query row per thread: d_h contiguous bf16) ----
        ptx::mbar_wait(b_pv, (uint32_t)(nt - 1) & 1u);
        ptx::tc_fence_after();
        const uint32_t lraw = ptx::tmem_ld1(t_row + FT_O_COL + DH);      // the row sum: column d_h of O
        ptx::tmem_ld_wait();
        const float inv = 1.0f / __uint_as_float(lraw);
        const int row = q0 + tid;
        bf16 *og = (bf16 *)p.o + (int64_t)img * p.o_bs + (int64_t)row * p.o_rs + head * DH;
#pragma unroll
        for (int c = 0; c < CH; c += 2) {
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
    if (warp == 5) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc<1>(tmem, FT_TMEM_COLS);
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

bool spatial_attention_tc_eligible(const FlashArgs &a) {
    if (a.dtype != NMM_BF16 || (a.dh != 40 && a.dh != 80) || a.kv_div != 1 || a.Lkv < 256) return false;
    // the tensor maps address whole rows: a row stride that covers the row's channels and image strides that are whole rows
    return a.q_rs % 8 == 0 && a.kv_rs % 8 == 0 && a.q_bs % 8 == 0 && a.kv_bs % 8 == 0 && a.o_rs % 8 == 0 && a.o_bs % 8 == 0 && aligned(a.q, 16) &&
           aligned(a.k, 16) && aligned(a.v, 16) && aligned(a.o, 16);
}

template <int DH>
static int launch_ft(const FlashArgs &a, int swap_v, cudaStream_t st) {
    using Cfg = FtCfg<DH>;
    static DeviceOnce once;
    NMM_CUDA_OK(once.max_smem(spatial_attention_tc_kernel<DH>, (int)Cfg::SMEM));
    CUtensorMap tq, tk, tv;
    int rc;
    // q / k / v are column slices of a wider row (q | k | v of one projection): the map starts at the slice, the chunk dimension spans the
    // rest of the row (the kernel only ever asks for the head's d_h / 8 chunks)
    if ((rc = ft_map(&tq, a.q, a.Lq, a.q_rs, a.q_bs, a.images, DH)) != NMM_OK) return rc;
    if ((rc = ft_map(&tk, a.k, a.Lkv, a.kv_rs, a.kv_bs, a.images, DH)) != NMM_OK) return rc;
    if ((rc = ft_map(&tv, a.v, a.Lkv, a.kv_rs, a.kv_bs, a.images, DH)) != NMM_OK) return rc;
    FtParams p;
    p.o = a.o; p.o_rs = a.o_rs; p.o_bs = a.o_bs; p.Lq = a.Lq; p.Lkv = a.Lkv; p.scale_log2e = a.scale_log2e; p.swap_v_desc = swap_v;
    const dim3 grid((unsigned)ceil_div(a.Lq, FT_BM), (unsigned)a.heads, (unsigned)a.images);
    const double per = (double)a.images * a.heads;
    ProfScope prof(K_SPATIAL_ATTN, st, 4.0 * per * a.Lq * (double)a.Lkv * DH, 2.0 * 4.0 * per * a.Lq * DH);
    NMM_CUDA_OK(launch_pdl(spatial_attention_tc_kernel<DH>, grid, dim3(FT_THREADS), (size_t)Cfg::SMEM, st, tq, tk, tv, p));
    NMM_LAUNCHED("spatial_attention_tc_kernel");
    return NMM_OK;
}

int launch_spatial_attention_tc(const FlashArgs &a, int swap_v, cudaStream_t st) {
    if (a.dh == 40) return launch_ft<40>(a, swap_v, st);
    if (a.dh == 80) return launch_ft<80>(a, swap_v, st);
    return fail(NMM_ERR_UNSUPPORTED, "tcgen05 spatial attention: d_h %d", a.dh);
}

}  // namespace nmm
