// fp32-GRADE spatial attention on the tensor cores (NMM_F32X3 mode of the spatial transformer): the flash-style kernel of
// spatial_attention.cu with every operand as a two-term bf16 split, x = hi + lo + O(2^-17 |x|), and three MMAs per product
//     S  = Qh Kh^T + Qh Kl^T + Ql Kh^T          O += Ph Vh + Ph Vl + Pl Vh            (fp32 accumulation; the lo.lo terms are below 2^-17)
// like the 3 x bf16 Linear layers (gemm_tcgen05.cu, X3).  q, k, v arrive as bf16 planes -- row = hi plane | lo plane, the format
// launch_convert_rows(..., NMM_F32X3) writes -- and the softmax weights are split in registers; the row sums add the fp32 weights; O leaves as
// plain fp32.  One CTA = 64 queries of one (image, head) (4 warps), 64-key stages (hi + lo tiles of K and V: 4 tiles per stage, double-buffered
// cp.async).  Reference arithmetic: CrossAttention._attention, motion_module_new.py:258-287 in fp32.
// This is what makes fp32 activations (the reference as shipped) usable on the spatial transformer beyond test sizes: ~3x the MMA work of the
// bf16 kernel instead of a CUDA-core kernel that streams K / V from L2 per query.
#include "attention_core.cuh"
#include "common.cuh"

namespace nmm {

constexpr int FX_BM = 64, FX_BN = 64, FX_THREADS = 128;

template <int DH>
struct FxCfg {
    static constexpr int PITCH = ((DH / 8) % 2 == 1) ? DH : DH + 8;
    static constexpr int CH = DH / 8;
    static constexpr int TILE = FX_BN * PITCH * 2;                      // one 64-row bf16 tile (Q, K or V; hi or lo)
    static constexpr int SMEM = (2 + 2 * 4) * TILE;                     // Qh Ql | stage 0: Kh Kl Vh Vl | stage 1: ...
    static constexpr int KS16 = DH / 16;
    static constexpr bool TAIL8 = (DH % 16) == 8;
    static constexpr int NT = DH / 8;
    static constexpr bool QRES = DH <= 80;                              // Q fragments (hi + lo) resident in registers; d_h = 160 reloads them
};

__device__ __forceinline__ void fx_cp16(uint32_t saddr, const void *g, bool valid) {
    const int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(saddr), "l"(g), "r"(sz) : "memory");
}
__device__ __forceinline__ float fx_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <int DH>
__global__ void __launch_bounds__(FX_THREADS, DH <= 40 ? 4 : DH <= 80 ? 2 : 1) spatial_attention_x3_kernel(const FlashArgs a) {
    using Cfg = FxCfg<DH>;
    constexpr int PITCH = Cfg::PITCH, CH = Cfg::CH, NT = Cfg::NT, KS16 = Cfg::KS16, TILE = Cfg::TILE;
    extern __shared__ __align__(128) uint8_t fx_smem[];
    pdl_wait();
    pdl_launch_dependents();
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q0 = blockIdx.x * FX_BM, head = blockIdx.y, img = blockIdx.z;
    const bf16 *qg = (const bf16 *)a.q + (int64_t)img * a.q_bs + head * DH;
    const int kv_img = img / a.kv_div;
    const bf16 *kg = (const bf16 *)a.k + (int64_t)kv_img * a.kv_bs + head * DH;
    const bf16 *vg = (const bf16 *)a.v + (int64_t)kv_img * a.kv_bs + head * DH;
    const uint32_t sq = (uint32_t)__cvta_generic_to_shared(fx_smem);          // Qh at sq, Ql at sq + TILE
    const uint32_t st0 = sq + 2 * TILE;                                        // stage s: Kh, Kl, Vh, Vl at st0 + (4 s + {0,1,2,3}) TILE
    const int Lq = a.Lq, Lkv = a.Lkv;

    for (int i = tid; i < FX_BM * CH; i += FX_THREADS) {
        const int r = i / CH, c = i - r * CH;
        const int row = q0 + r;
        const bool ok = row < Lq;
        const bf16 *src = qg + (int64_t)(ok ? row : Lq - 1) * a.q_rs + c * 8;
        const uint32_t so = (uint32_t)(r * PITCH + c * 8) * 2;
        fx_cp16(sq + so, src, ok);
        fx_cp16(sq + TILE + so, src + a.q_lo_off, ok);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    auto load_kv = [&](int t, int stage) {
        const uint32_t sb = st0 + stage * 4 * TILE;
        for (int i = tid; i < FX_BN * CH; i += FX_THREADS) {
            const int r = i / CH, c = i - r * CH;
            const int key = t * FX_BN + r;
            const bool ok = key < Lkv;
            const int64_t off = (int64_t)(ok ? key : Lkv - 1) * a.kv_rs + c * 8;
            const uint32_t so = (uint32_t)(r * PITCH + c * 8) * 2;
            fx_cp16(sb + so, kg + off, ok);
            fx_cp16(sb + TILE + so, kg + off + a.kv_lo_off, ok);
            fx_cp16(sb + 2 * TILE + so, vg + off, ok);
            fx_cp16(sb + 3 * TILE + so, vg + off + a.kv_lo_off, ok);
        }
    };
    load_kv(0, 0);
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 1;" ::: "memory");
    __syncthreads();

    const uint32_t qa = sq + (uint32_t)((warp * 16 + (lane & 7) + 8 * ((lane >> 3) & 1)) * PITCH + 8 * (lane >> 4)) * 2;
    const uint32_t qa8 = sq + (uint32_t)((warp * 16 + (lane & 7) + 8 * ((lane >> 3) & 1)) * PITCH) * 2 + KS16 * 32;
    uint32_t qh[Cfg::QRES ? KS16 : 1][4], ql[Cfg::QRES ? KS16 : 1][4];
    uint32_t qth[2] = {0u, 0u}, qtl[2] = {0u, 0u};
    if constexpr (Cfg::QRES) {
#pragma unroll
        for (int ks = 0; ks < KS16; ks++) {
            ldsm_x4(qa + ks * 32, qh[ks][0], qh[ks][1], qh[ks][2], qh[ks][3]);
            ldsm_x4(qa + TILE + ks * 32, ql[ks][0], ql[ks][1], ql[ks][2], ql[ks][3]);
        }
    }
    if constexpr (Cfg::TAIL8) {
        ldsm_x2(qa8, qth[0], qth[1]);
        ldsm_x2(qa8 + TILE, qtl[0], qtl[1]);
    }

    float o[NT][4];
#pragma unroll
    for (int n = 0; n < NT; n++) { o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f; }
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
    const float sl = a.scale_log2e;
    const int nt = (Lkv + FX_BN - 1) / FX_BN;
    const uint32_t k_lane = (uint32_t)((lane & 7) * PITCH + 8 * ((lane >> 3) & 1)) * 2;                       // x2: keys 8j.., channels 0-7 | 8-15 of a k16 step
    const uint32_t v_lane = (uint32_t)(((lane & 7) + 8 * ((lane >> 3) & 1)) * PITCH + 8 * (lane >> 4)) * 2;   // x4.trans: keys 0-7 | 8-15, channel tiles n | n + 1

    for (int t = 0; t < nt; t++) {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        if (t + 1 < nt) load_kv(t + 1, (t + 1) & 1);
        asm volatile("cp.async.commit_group;" ::: "memory");
        const uint32_t sb = st0 + (t & 1) * 4 * TILE;
        const uint32_t skh = sb + k_lane, skl = sb + TILE + k_lane, svh = sb + 2 * TILE + v_lane, svl = sb + 3 * TILE + v_lane;

        // ---- S = Qh Kh^T + Qh Kl^T + Ql Kh^T (16 x 64 per warp) ----
        float s[8][4];
#pragma unroll
        for (int j = 0; j < 8; j++) {
            s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
            const uint32_t ko = (uint32_t)(8 * j * PITCH) * 2;
#pragma unroll
            for (int ks = 0; ks < KS16; ks++) {
                uint32_t h0, h1, g0, g1;
                ldsm_x2(skh + ko + ks * 32, h0, h1);
                ldsm_x2(skl + ko + ks * 32, g0, g1);
                uint32_t ah[4], al[4];
                if constexpr (Cfg::QRES) {
#pragma unroll
                    for (int i = 0; i < 4; i++) { ah[i] = qh[ks][i]; al[i] = ql[ks][i]; }
                } else {
                    ldsm_x4(qa + ks * 32, ah[0], ah[1], ah[2], ah[3]);
                    ldsm_x4(qa + TILE + ks * 32, al[0], al[1], al[2], al[3]);
                }
                mma_k16(s[j], ah[0], ah[1], ah[2], ah[3], h0, h1);
                mma_k16(s[j], ah[0], ah[1], ah[2], ah[3], g0, g1);
                mma_k16(s[j], al[0], al[1], al[2], al[3], h0, h1);
            }
            if constexpr (Cfg::TAIL8) {
                uint32_t h0, g0;
                ldsm_x1(skh + ko + KS16 * 32, h0);
                ldsm_x1(skl + ko + KS16 * 32, g0);
                mma_k8(s[j], qth[0], qth[1], h0);
                mma_k8(s[j], qth[0], qth[1], g0);
                mma_k8(s[j], qtl[0], qtl[1], h0);
            }
        }
        if ((t + 1) * FX_BN > Lkv) {
            const int kbase = t * FX_BN + 2 * (lane & 3);
#pragma unroll
            for (int j = 0; j < 8; j++) {
                if (kbase + 8 * j >= Lkv) { s[j][0] = -INFINITY; s[j][2] = -INFINITY; }
                if (kbase + 8 * j + 1 >= Lkv) { s[j][1] = -INFINITY; s[j][3] = -INFINITY; }
            }
        }
        // ---- online softmax in fp32; weights split into bf16 hi + lo ----
        float mx0 = m0, mx1 = m1;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            mx0 = fmaxf(mx0, fmaxf(s[j][0], s[j][1]));
            mx1 = fmaxf(mx1, fmaxf(s[j][2], s[j][3]));
        }
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
        const float al0 = fx_exp2((m0 - mx0) * sl), al1 = fx_exp2((m1 - mx1) * sl);
        m0 = mx0; m1 = mx1;
        const float ms0 = mx0 * sl, ms1 = mx1 * sl;
        float r0 = 0.f, r1 = 0.f;
        uint32_t ph[8][2], pl[8][2];
#pragma unroll
        for (int j = 0; j < 8; j++) {
            // exp2f, not ex2.approx: the approximation's 2^-22 relative error is fine, but keep the checker-grade path simple and exact-ish
            const float e0 = exp2f(fmaf(s[j][0], sl, -ms0)), e1 = exp2f(fmaf(s[j][1], sl, -ms0));
            const float e2 = exp2f(fmaf(s[j][2], sl, -ms1)), e3 = exp2f(fmaf(s[j][3], sl, -ms1));
            r0 += e0 + e1; r1 += e2 + e3;
            split_bf16x2(e0, e1, ph[j][0], pl[j][0]);
            split_bf16x2(e2, e3, ph[j][1], pl[j][1]);
        }
        l0 = fmaf(l0, al0, r0); l1 = fmaf(l1, al1, r1);
#pragma unroll
        for (int n = 0; n < NT; n++) { o[n][0] *= al0; o[n][1] *= al0; o[n][2] *= al1; o[n][3] *= al1; }
        // ---- O += Ph Vh + Ph Vl + Pl Vh ----
#pragma unroll
        for (int kk = 0; kk < 4; kk++) {
            const uint32_t vo = (uint32_t)(16 * kk * PITCH) * 2;
            const uint32_t a0 = ph[2 * kk][0], a1 = ph[2 * kk][1], a2 = ph[2 * kk + 1][0], a3 = ph[2 * kk + 1][1];
            const uint32_t c0 = pl[2 * kk][0], c1 = pl[2 * kk][1], c2 = pl[2 * kk + 1][0], c3 = pl[2 * kk + 1][1];
#pragma unroll
            for (int n = 0; n + 2 <= NT; n += 2) {
                uint32_t b0, b1, b2, b3, d0, d1, d2, d3;
                ldsm_x4_t(svh + vo + n * 16, b0, b1, b2, b3);
                ldsm_x4_t(svl + vo + n * 16, d0, d1, d2, d3);
                mma_k16(o[n], a0, a1, a2, a3, b0, b1); mma_k16(o[n], a0, a1, a2, a3, d0, d1); mma_k16(o[n], c0, c1, c2, c3, b0, b1);
                mma_k16(o[n + 1], a0, a1, a2, a3, b2, b3); mma_k16(o[n + 1], a0, a1, a2, a3, d2, d3); mma_k16(o[n + 1], c0, c1, c2, c3, b2, b3);
            }
            if constexpr (NT % 2 == 1) {
                uint32_t b0, b1, d0, d1;
                ldsm_x2_t(svh + vo + (NT - 1) * 16, b0, b1);
                ldsm_x2_t(svl + vo + (NT - 1) * 16, d0, d1);
                mma_k16(o[NT - 1], a0, a1, a2, a3, b0, b1); mma_k16(o[NT - 1], a0, a1, a2, a3, d0, d1); mma_k16(o[NT - 1], c0, c1, c2, c3, b0, b1);
            }
        }
    }
    // ---- O / l -> fp32 rows (each quad writes 32 contiguous bytes) ----
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float i0 = 1.f / l0, i1 = 1.f / l1;
    const int crow = lane >> 2, ccol = (lane & 3) * 2;
    float *og = (float *)a.o + (int64_t)img * a.o_bs + head * DH;
    const int row0 = q0 + warp * 16 + crow, row1 = row0 + 8;
#pragma unroll
    for (int n = 0; n < NT; n++) {
        if (row0 < Lq) *reinterpret_cast<float2 *>(og + (int64_t)row0 * a.o_rs + 8 * n + ccol) = make_float2(o[n][0] * i0, o[n][1] * i0);
        if (row1 < Lq) *reinterpret_cast<float2 *>(og + (int64_t)row1 * a.o_rs + 8 * n + ccol) = make_float2(o[n][2] * i1, o[n][3] * i1);
    }
}

template <int DH>
static int launch_fx(const FlashArgs &a, cudaStream_t st) {
    using Cfg = FxCfg<DH>;
    static DeviceOnce once;
    NMM_CUDA_OK(once.max_smem(spatial_attention_x3_kernel<DH>, Cfg::SMEM));
    const dim3 grid((unsigned)ceil_div(a.Lq, FX_BM), (unsigned)a.heads, (unsigned)a.images);
    const double per = (double)a.images * a.heads;
    ProfScope prof(K_SPATIAL_ATTN, st, 4.0 * per * a.Lq * (double)a.Lkv * DH, 4.0 * (2.0 * per * a.Lq * DH + 2.0 * per / a.kv_div * a.Lkv * DH));
    NMM_CUDA_OK(launch_pdl(spatial_attention_x3_kernel<DH>, grid, dim3(FX_THREADS), (size_t)Cfg::SMEM, st, a));
    NMM_LAUNCHED("spatial_attention_x3_kernel");
    return NMM_OK;
}

// q, k, v: bf16 hi | lo planes (hi at the pointer, lo `*_lo_off` elements further in the same row); o: fp32
int launch_spatial_attention_x3(const FlashArgs &a, cudaStream_t st) {
    if (a.Lq <= 0 || a.Lkv <= 0 || a.images <= 0 || a.heads <= 0 || a.kv_div <= 0) return fail(NMM_ERR_BAD_ARG, "spatial attention: non-positive size");
    if (!aligned(a.q, 16) || !aligned(a.k, 16) || !aligned(a.v, 16) || !aligned(a.o, 8) || a.q_rs % 8 || a.kv_rs % 8 || a.q_bs % 8 || a.kv_bs % 8 ||
        a.q_lo_off % 8 || a.kv_lo_off % 8 || a.o_rs % 2 || a.o_bs % 2)
        return fail(NMM_ERR_BAD_ARG, "fp32-grade spatial attention: misaligned operand planes");
    switch (a.dh) {
        case 40: return launch_fx<40>(a, st);
        case 80: return launch_fx<80>(a, st);
        case 160: return launch_fx<160>(a, st);
        default: break;
    }
    return fail(NMM_ERR_UNSUPPORTED, "fp32-grade spatial attention: head dim %d (supported: 40, 80, 160)", a.dh);
}

}  // namespace nmm
