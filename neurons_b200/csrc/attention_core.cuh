// Tensor-core temporal attention on a shared-memory tile: the compute shared by the stand-alone attention kernel
// (attention_kernel.cu) and the QKV GEMM's fused epilogue (gemm_tcgen05.cu).
// Tile layout: row (pl * F + f) holds q | k | v of position pl, frame f: 3 segments of W = HB * DH bf16 each, pitch 3 W + 8 elements
// (pitch / 2 words = 4 mod 32 for W in {80, 320}: the 8 rows of an ldmatrix hit distinct banks).  One warp per (position, head)
// problem: S = Q K^T and O = P V on mma.sync (m16n8k16 / m16n8k8 bf16 -> fp32), fp32 base-2 softmax in the accumulator
// fragments, P fed back as a bf16 hi + lo pair; O (bf16) overwrites the problem's q slot.
// Reference arithmetic: CrossAttention._attention, motion_module_new.py:258-287 (scale d_h^-1/2, softmax over frames, no mask).
#pragma once
#include "common.cuh"

namespace nmm {

__device__ __forceinline__ void ldsm_x1(uint32_t addr, uint32_t &r0) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x1.shared.b16 {%0}, [%1];" : "=r"(r0) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x2(uint32_t addr, uint32_t &r0, uint32_t &r1) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0, %1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t &r0, uint32_t &r1, uint32_t &r2, uint32_t &r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x1_t(uint32_t addr, uint32_t &r0) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x1.trans.shared.b16 {%0}, [%1];" : "=r"(r0) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x2_t(uint32_t addr, uint32_t &r0, uint32_t &r1) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0, %1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t &r0, uint32_t &r1, uint32_t &r2, uint32_t &r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma_k16(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma_k8(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t b0) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5}, {%6}, {%0, %1, %2, %3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a0), "r"(a1), "r"(b0));
}
// (split_bf16x2: two fp32 -> bf16 hi and bf16 lo (residual) pairs, common.cuh)


// NP independent (position, head) problems of one warp, their instruction streams interleaved phase by phase: the chain
// ldmatrix -> mma -> shuffle-softmax -> mma of a single problem is latency-bound, two in flight roughly halve the time per problem
// when few warps share the tile (the fused QKV epilogue has 8).  qb[u] = shared-memory address of problem u's q rows (row 0, its head's
// first column); the result overwrites those q columns.
// `store(u, row, col, v)`: two adjacent bf16 context values (packed in v) of problem u, frame `row`, head-dim columns col, col + 1.
// The stand-alone kernel puts them into the problem's q slot (SmemQSlotStore); the fused QKV epilogue writes them straight to global.
template <int W>
struct SmemQSlotStore {
    const uint32_t *qb;
    __device__ __forceinline__ void operator()(int u, int row, int col, uint32_t v) const {
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(qb[u] + (uint32_t)row * ((3 * W + 8) * 2) + (uint32_t)(col * 2)), "r"(v) : "memory");
    }
};

template <int F, int DH, int W, int NP, typename Store>
__device__ __forceinline__ void attention_problems(const uint32_t (&qb)[NP], int lane, float scale_log2e, const Store &store) {
    constexpr int NT = F / 8;
    constexpr uint32_t RS = (3 * W + 8) * 2;    // row stride in bytes
    constexpr uint32_t KOFF = W * 2, VOFF = 2 * W * 2;
    const int lrow = lane & 7, lmat = lane >> 3;          // ldmatrix: lanes 8m..8m+7 address the rows of matrix m
    const int crow = lane >> 2, ccol = (lane & 3) * 2;     // accumulator fragment: row crow (and crow + 8), columns ccol, ccol + 1
    // per-lane ldmatrix offsets relative to qb (constant offsets are added as immediates below)
    const uint32_t q_off = F == 16 ? (uint32_t)(lrow + 8 * (lmat & 1)) * RS + (uint32_t)(16 * (lmat >> 1)) : (uint32_t)lrow * RS + (uint32_t)(16 * lmat);
    const uint32_t k_off = KOFF + (uint32_t)lrow * RS + (uint32_t)(16 * lmat);          // 4 matrices = 32 columns (F == 8) ...
    const uint32_t k_off2 = KOFF + (uint32_t)lrow * RS + (uint32_t)(16 * (lmat & 1));   // ... or 2 matrices = 16 columns
    float s[NP][NT][4];
#pragma unroll
    for (int u = 0; u < NP; u++)
#pragma unroll
        for (int j = 0; j < NT; j++) { s[u][j][0] = s[u][j][1] = s[u][j][2] = s[u][j][3] = 0.f; }
    // ---- S = Q K^T ------------------------------------------------------------------------------------------------------------
    if constexpr (F == 8) {
        // 32 head-dim columns per step: one ldmatrix.x4 for Q (rows 0-7), one for K (keys 0-7), two k16 MMAs
#pragma unroll
        for (int k0 = 0; k0 + 32 <= DH; k0 += 32) {
#pragma unroll
            for (int u = 0; u < NP; u++) {
                uint32_t a0, a1, a2, a3, b0, b1, b2, b3;
                ldsm_x4(qb[u] + q_off + k0 * 2, a0, a1, a2, a3);
                ldsm_x4(qb[u] + k_off + k0 * 2, b0, b1, b2, b3);
                mma_k16(s[u][0], a0, 0u, a1, 0u, b0, b1);
                mma_k16(s[u][0], a2, 0u, a3, 0u, b2, b3);
            }
        }
        constexpr int K1 = DH / 32 * 32;
        if constexpr (DH - K1 >= 16) {
#pragma unroll
            for (int u = 0; u < NP; u++) {
                uint32_t a0, a2, b0, b1;
                ldsm_x2(qb[u] + q_off + K1 * 2, a0, a2);
                ldsm_x2(qb[u] + k_off + K1 * 2, b0, b1);
                mma_k16(s[u][0], a0, 0u, a2, 0u, b0, b1);
            }
        }
        constexpr int K2 = DH / 16 * 16;
        if constexpr (DH - K2 == 8) {
#pragma unroll
            for (int u = 0; u < NP; u++) {
                uint32_t a0, b0;
                ldsm_x1(qb[u] + q_off + K2 * 2, a0);
                ldsm_x1(qb[u] + k_off + K2 * 2, b0);
                mma_k8(s[u][0], a0, 0u, b0);
            }
        }
    } else {
#pragma unroll
        for (int k0 = 0; k0 + 16 <= DH; k0 += 16) {
#pragma unroll
            for (int u = 0; u < NP; u++) {
                uint32_t a0, a1, a2, a3;
                ldsm_x4(qb[u] + q_off + k0 * 2, a0, a1, a2, a3);
#pragma unroll
                for (int j = 0; j < NT; j++) {
                    uint32_t b0, b1;
                    ldsm_x2(qb[u] + k_off2 + (uint32_t)(8 * j) * RS + k0 * 2, b0, b1);
                    mma_k16(s[u][j], a0, a1, a2, a3, b0, b1);
                }
            }
        }
        constexpr int K2 = DH / 16 * 16;
        if constexpr (DH - K2 == 8) {
#pragma unroll
            for (int u = 0; u < NP; u++) {
                uint32_t a0, a1;
                ldsm_x2(qb[u] + (uint32_t)(lrow + 8 * (lmat & 1)) * RS + K2 * 2, a0, a1);
#pragma unroll
                for (int j = 0; j < NT; j++) {
                    uint32_t b0;
                    ldsm_x1(qb[u] + KOFF + (uint32_t)(8 * j + lrow) * RS + K2 * 2, b0);
                    mma_k8(s[u][j], a0, a1, b0);
                }
            }
        }
    }
    // ---- softmax (fp32, base-2 exponentials) over the keys of row crow (regs 0,1) and row crow + 8 (regs 2,3; F == 16 only) -------
    uint32_t ph[NP][4], pl_[NP][4];
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
        // P as the A operand of the second MMA (accumulator layout == A layout for these shapes), bf16 hi + lo
        ph[u][0] = ph[u][1] = ph[u][2] = ph[u][3] = 0u; pl_[u][0] = pl_[u][1] = pl_[u][2] = pl_[u][3] = 0u;
        split_bf16x2(s[u][0][0] * inv0, s[u][0][1] * inv0, ph[u][0], pl_[u][0]);
        if constexpr (F == 16) {
            split_bf16x2(s[u][0][2] * inv1, s[u][0][3] * inv1, ph[u][1], pl_[u][1]);
            split_bf16x2(s[u][1][0] * inv0, s[u][1][1] * inv0, ph[u][2], pl_[u][2]);
            split_bf16x2(s[u][1][2] * inv1, s[u][1][3] * inv1, ph[u][3], pl_[u][3]);
        }
    }
    __syncwarp();                                   // all of this warp's reads of the q rows are done: reuse them for O
    // ---- O = P V: one ldmatrix.x4.trans feeds 32 (F == 8) or 16 (F == 16) output columns ---------------------------------------------
    if constexpr (F == 8 && NP == 2) {
        // Two 8-frame problems share every m16n8k16: A = blockdiag(P_0, P_1) (rows 0-7 / keys 0-7 = problem 0, rows 8-15 / keys 8-15 =
        // problem 1), B = [V_0; V_1] -> D rows 0-7 = P_0 V_0, rows 8-15 = P_1 V_1.  Half the tensor-core instructions of two separate
        // m16n8k8 streams -- legacy mma.sync shares the tensor pipe with the tcgen05 mainloop, so the count is what matters here.
        const uint32_t vsel = qb[(lmat & 1)] + VOFF + (uint32_t)lrow * RS;     // ldmatrix matrices 0,2 <- problem 0's V rows; 1,3 <- problem 1's
#pragma unroll
        for (int n0 = 0; n0 < DH; n0 += 16) {
            uint32_t bv[4] = {0u, 0u, 0u, 0u};
            if (n0 + 16 <= DH) ldsm_x4_t(vsel + (uint32_t)((n0 + 8 * (lmat >> 1)) * 2), bv[0], bv[1], bv[2], bv[3]);
            else ldsm_x2_t(vsel + (uint32_t)(n0 * 2), bv[0], bv[1]);                   // last 8 columns (lanes 0-15 address)
#pragma unroll
            for (int q = 0; q < 2; q++) {
                if (n0 + 8 * q < DH) {
                    float o[4] = {0.f, 0.f, 0.f, 0.f};
                    mma_k16(o, ph[0][0], 0u, 0u, ph[1][0], bv[2 * q], bv[2 * q + 1]);
                    mma_k16(o, pl_[0][0], 0u, 0u, pl_[1][0], bv[2 * q], bv[2 * q + 1]);
                    store(0, crow, ccol + (n0 + 8 * q), pack_bf16x2(o[0], o[1]));
                    store(1, crow, ccol + (n0 + 8 * q), pack_bf16x2(o[2], o[3]));
                }
            }
        }
    } else if constexpr (F == 8) {
        const uint32_t v_off = VOFF + (uint32_t)lrow * RS + (uint32_t)(16 * lmat);
#pragma unroll
        for (int n0 = 0; n0 + 32 <= DH; n0 += 32) {
#pragma unroll
            for (int u = 0; u < NP; u++) {
                uint32_t bv[4];
                ldsm_x4_t(qb[u] + v_off + n0 * 2, bv[0], bv[1], bv[2], bv[3]);
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    float o[4] = {0.f, 0.f, 0.f, 0.f};
                    mma_k8(o, ph[u][0], 0u, bv[q]);
                    mma_k8(o, pl_[u][0], 0u, bv[q]);
                    store(u, crow, ccol + (n0 + 8 * q), pack_bf16x2(o[0], o[1]));
                }
            }
        }
        constexpr int N1 = DH / 32 * 32;
#pragma unroll
        for (int n0 = N1; n0 < DH; n0 += 8) {
#pragma unroll
            for (int u = 0; u < NP; u++) {
                uint32_t b0;
                ldsm_x1_t(qb[u] + VOFF + (uint32_t)lrow * RS + n0 * 2, b0);
                float o[4] = {0.f, 0.f, 0.f, 0.f};
                mma_k8(o, ph[u][0], 0u, b0);
                mma_k8(o, pl_[u][0], 0u, b0);
                store(u, crow, ccol + n0, pack_bf16x2(o[0], o[1]));
            }
        }
    } else {
        // matrices: (keys 0-7, n0), (keys 8-15, n0), (keys 0-7, n0 + 8), (keys 8-15, n0 + 8)
        const uint32_t v_off = VOFF + (uint32_t)(lrow + 8 * (lmat & 1)) * RS + (uint32_t)(16 * (lmat >> 1));
#pragma unroll
        for (int n0 = 0; n0 + 16 <= DH; n0 += 16) {
#pragma unroll
            for (int u = 0; u < NP; u++) {
                uint32_t bv[4];
                ldsm_x4_t(qb[u] + v_off + n0 * 2, bv[0], bv[1], bv[2], bv[3]);
#pragma unroll
                for (int q = 0; q < 2; q++) {
                    float o[4] = {0.f, 0.f, 0.f, 0.f};
                    mma_k16(o, ph[u][0], ph[u][1], ph[u][2], ph[u][3], bv[2 * q], bv[2 * q + 1]);
                    mma_k16(o, pl_[u][0], pl_[u][1], pl_[u][2], pl_[u][3], bv[2 * q], bv[2 * q + 1]);
                    store(u, crow, ccol + (n0 + 8 * q), pack_bf16x2(o[0], o[1]));
                    store(u, crow + 8, ccol + (n0 + 8 * q), pack_bf16x2(o[2], o[3]));
                }
            }
        }
        constexpr int N1 = DH / 16 * 16;
        if constexpr (DH - N1 == 8) {
#pragma unroll
            for (int u = 0; u < NP; u++) {
                uint32_t b0, b1;
                ldsm_x2_t(qb[u] + VOFF + (uint32_t)(lrow + 8 * (lmat & 1)) * RS + N1 * 2, b0, b1);
                float o[4] = {0.f, 0.f, 0.f, 0.f};
                mma_k16(o, ph[u][0], ph[u][1], ph[u][2], ph[u][3], b0, b1);
                mma_k16(o, pl_[u][0], pl_[u][1], pl_[u][2], pl_[u][3], b0, b1);
                store(u, crow, ccol + N1, pack_bf16x2(o[0], o[1]));
                store(u, crow + 8, ccol + N1, pack_bf16x2(o[2], o[3]));
            }
        }
    }
}

// All problems of a tile, one per warp per pass (the stand-alone attention kernel: many CTAs per SM hide the latency).
template <int F, int DH, int W>      // F in {8, 16}; d_h % 8 == 0; W = heads per tile * d_h
__device__ __forceinline__ void attention_tile_mma(uint32_t sm_u32, int npos, int warp, int nwarps, int lane, float scale_log2e) {
    constexpr int HB = W / DH;
    constexpr uint32_t RS = (3 * W + 8) * 2;
    for (int prob = warp; prob < npos * HB; prob += nwarps) {
        const int pl = prob / HB, hd = prob - pl * HB;
        const uint32_t qb[1] = {sm_u32 + (uint32_t)(pl * F) * RS + (uint32_t)(hd * DH * 2)};
        attention_problems<F, DH, W, 1>(qb, lane, scale_log2e, SmemQSlotStore<W>{qb});
    }
}

}  // namespace nmm
