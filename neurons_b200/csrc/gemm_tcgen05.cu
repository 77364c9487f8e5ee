// bf16 GEMM on the 5th-generation tensor cores:  D[M,N] = A[M,K] . W[N,K]^T, fp32 accumulation in TMEM,
// fused epilogues (epilogue.cuh).  This is the production kernel for every nn.Linear on the motion-module
// path (motion_module.py:145,152,289,297,298,321; motion_module_new.py:466,516).
//
// Structure (persistent, warp-specialised, one CTA per SM):
//   warp 0      TMA producer: cp.async.bulk.tensor 2-D tiles of A (128 x 64) and W (block_n x 64), 128-byte swizzle,
//               into a `stages`-deep shared-memory ring; completion via mbarrier complete_tx.
//   warp 1      MMA issuer: one elected thread issues tcgen05.mma.cta_group::1.kind::f16 (M=128, N=block_n, K=16)
//               4x per stage, tcgen05.commit releases the stage / publishes the accumulator.
//   warp 2      TMEM allocator (2 accumulator buffers of block_n fp32 columns -> epilogue overlaps the next tile).
//   warps 4-11  epilogue: tcgen05.ld (lane == output row) -> warp-private shared-memory transpose -> bias / residual / GEGLU
//               with fully coalesced 16-byte global accesses; the NCHW store of proj_out is coalesced along p as is.
// Both operands are K-major in global memory ([rows, K] row-major), which is exactly how activations
// (token-major) and nn.Linear weights ([out, in]) are laid out, so no transposes are ever materialised.
#include <cuda.h>

#include <cstdlib>
#include <cstring>
#include <mutex>

#include "common.cuh"
#include "epilogue.cuh"
#include "ptx.cuh"

namespace nmm {

constexpr int TC_BM = 128, TC_BK = 64;
constexpr int TC_EPI_WARPS = 8;                           // two warps per TMEM lane quadrant, alternating column chunks
constexpr int TC_THREADS = 128 + 32 * TC_EPI_WARPS;       // warps 0-3: TMA / MMA / TMEM alloc / spare; warps 4-11: epilogue
constexpr int TC_A_BYTES = TC_BM * TC_BK * 2;            // 16 KB per stage
constexpr int TC_SMEM_BUDGET = 220 * 1024;
constexpr int TC_STAGE_PITCH = 36;                        // floats per staged row: 32 columns + 4 pad (conflict-free 16-byte accesses)
constexpr int TC_STAGE_BYTES = TC_EPI_WARPS * 32 * TC_STAGE_PITCH * 4;   // 36 KB of warp-private transpose buffers

struct TcParams {
    int64_t M;
    int N, K;
    int block_n, stages, n_tiles, tmem_cols;
    int64_t m_tiles;
    int cluster;             // CTAs per tile along M: 1 (cta_group::1) or 2 (cta_group::2 pair, each CTA stages half of W)
    int64_t cluster_tiles;   // ceil(m_tiles / cluster) * n_tiles
    unsigned long long *trace;   // NMM_TRACE builds only: per-tile timestamps of CTA 0 (development instrumentation)
    int debug;               // NMM_GEMM_DEBUG (timing experiments only, results invalid): 1 = epilogue does nothing, 2 = no TMA loads
};


#ifdef NMM_TRACE
// event slots per tile (CTA 0 only, first TRACE_TILES tiles): 0 mma:tile start, 1 mma:accumulator free, 2 mma:first stage full,
// 3 mma:all issued, 4 prod:first load issued, 5 prod:last load issued, 6 epi(w4):before tfull wait, 7 epi:accumulator ready,
// 8..11 epi: chunk k done, 12 epi: released, 13 epi(w8): ready, 14 epi(w8): released
constexpr int TRACE_TILES = 48, TRACE_SLOTS = 16;
#define TRACE(tile_no, slot)                                                                          \
    do {                                                                                              \
        if (p.trace != nullptr && blockIdx.x == 0 && (tile_no) < TRACE_TILES) p.trace[(tile_no) * TRACE_SLOTS + (slot)] = clock64(); \
    } while (0)
#else
#define TRACE(tile_no, slot) do { } while (0)
#endif

// ---- coalesced epilogue ------------------------------------------------------------------------------------------
// tcgen05.ld hands each lane ONE ROW of the accumulator (lane == TMEM lane == output row), so a direct store makes every
// lane of a warp hit a different 128-byte line.  Each epilogue warp therefore transposes its 32 x 32 fp32 chunk through a
// private shared-memory buffer: after the transpose 8 consecutive lanes own 32 consecutive columns of one row
// (128 contiguous bytes of fp32), so the residual read-modify-write and all stores are fully coalesced.
template <int EPI>
__device__ __forceinline__ void epi_chunk(const EpiParams &e, float *stage, int lane, int64_t row0, int col0, int width,
                                          const uint32_t (&acc)[32], const float4 (&res)[8], int debug) {
    (void)debug;
    // registers (row per lane) -> staging buffer
    float4 *mine = reinterpret_cast<float4 *>(stage + lane * TC_STAGE_PITCH);
#pragma unroll
    for (int j = 0; j < 8; j++)
        if (j * 4 < width)
            mine[j] = make_float4(__uint_as_float(acc[4 * j]), __uint_as_float(acc[4 * j + 1]), __uint_as_float(acc[4 * j + 2]),
                                  __uint_as_float(acc[4 * j + 3]));
    __syncwarp();
    const int cl = (lane & 7) * 4, rl = lane >> 3;
    const int col = col0 + cl;
    if (cl < width) {                                    // warp-uniform per 8-lane group; width is 16 or 32
        float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (e.bias != nullptr) b4 = __ldg(reinterpret_cast<const float4 *>(e.bias + col));
        // straight-line code (no branches inside the unrolled loops) so the 8 rows overlap: all loads, then math, then
        // predicated stores
        float4 v[8];
#pragma unroll
        for (int it = 0; it < 8; it++) v[it] = *reinterpret_cast<const float4 *>(stage + (it * 4 + rl) * TC_STAGE_PITCH + cl);
#pragma unroll
        for (int it = 0; it < 8; it++) {
            v[it].x += b4.x; v[it].y += b4.y; v[it].z += b4.z; v[it].w += b4.w;
            if constexpr (EPI == NMM_EPI_RESIDUAL) { v[it].x += res[it].x; v[it].y += res[it].y; v[it].z += res[it].z; v[it].w += res[it].w; }
        }
        if constexpr (EPI == NMM_EPI_GEGLU) {
            uint32_t o[8];
#pragma unroll
            for (int it = 0; it < 8; it++) o[it] = pack_bf16x2(v[it].x * gelu_erf_fast(v[it].y), v[it].z * gelu_erf_fast(v[it].w));
            bf16 *dst = reinterpret_cast<bf16 *>(e.out) + (row0 + rl) * (e.N / 2) + col / 2;
#pragma unroll
            for (int it = 0; it < 8; it++)
                if (row0 + it * 4 + rl < e.M) *reinterpret_cast<uint32_t *>(dst + (int64_t)(it * 4) * (e.N / 2)) = o[it];
        } else {
            if (e.h != nullptr) {
                float *dst = e.h + (row0 + rl) * e.N + col;
#pragma unroll
                for (int it = 0; it < 8; it++)
                    if (row0 + it * 4 + rl < e.M) *reinterpret_cast<float4 *>(dst + (int64_t)(it * 4) * e.N) = v[it];
            }
            if (e.out != nullptr) {
                bf16 *dst = reinterpret_cast<bf16 *>(e.out) + (row0 + rl) * e.N + col;
#pragma unroll
                for (int it = 0; it < 8; it++)
                    if (row0 + it * 4 + rl < e.M)
                        *reinterpret_cast<uint2 *>(dst + (int64_t)(it * 4) * e.N) =
                            make_uint2(pack_bf16x2(v[it].x, v[it].y), pack_bf16x2(v[it].z, v[it].w));
            }
        }
    }
    __syncwarp();            // the next chunk overwrites the staging buffer
}

template <int EPI, int CG>
__global__ void __launch_bounds__(TC_THREADS, 1)
linear_tc_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_w, TcParams p, EpiParams e) {
    // CG == 1: one CTA per 128 x block_n tile (tcgen05.mma.cta_group::1).
    // CG == 2: a CTA pair (cluster of 2 along M) computes a 256 x block_n tile with ONE tcgen05.mma.cta_group::2 stream issued
    //          by the even CTA: each CTA stages its own 128 rows of A and HALF of the W tile (block_n/2 rows); the tensor core
    //          reads the two W halves from both CTAs' shared memory and each CTA's TMEM receives its own 128 x block_n
    //          accumulator.  Per SM and k-block the shared-memory fill drops from 16 KB + block_n*128 B to 16 KB + block_n*64 B
    //          -- the L2 -> SM ingest rate, not HBM or the tensor pipe, is what bounds the single-CTA kernel.
    extern __shared__ unsigned char smem_raw[];
    const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;     // SWIZZLE_128B tiles need 1024-byte alignment
    const uint32_t b_rows = (uint32_t)p.block_n / CG;                           // W rows staged by this CTA
    const uint32_t stage_bytes = TC_A_BYTES + b_rows * TC_BK * 2;
    const uint32_t bar_base = smem_base + (uint32_t)p.stages * stage_bytes;
    // barriers: full[stages], empty[stages], tmem_full[2], tmem_empty[2]; then the TMEM base address
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (p.stages + s); };
    auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * p.stages + s); };
    auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * p.stages + 2 + s); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * p.stages + 4);
    // warp-private transpose buffers of the epilogue warps (generic pointer: plain ld/st.shared)
    float *stage_base = reinterpret_cast<float *>(smem_raw + (smem_base - ptx::smem_u32(smem_raw)) + (size_t)p.stages * stage_bytes + 256);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_kb = (p.K + TC_BK - 1) / TC_BK;
    // Tile schedule: a cluster walks (m_group, n_blk) pairs; CTA `rank` of the cluster owns m-block m_group*CG + rank.
    // Both CTAs of a pair run the same number of iterations (a phantom m-block past the end is all zero-fill + masked).
    const uint32_t rank = CG > 1 ? ptx::cluster_ctarank() : 0u;
    const bool leader = rank == 0;
    const int64_t cluster_id = blockIdx.x / CG, num_clusters = gridDim.x / CG;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&tm_a);
        ptx::prefetch_tensormap(&tm_w);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < p.stages; s++) { ptx::mbar_init(full_bar(s), 1); ptx::mbar_init(empty_bar(s), 1); }
        // tmem_empty collects one arrive per epilogue warp of every CTA of the pair (it lives in the MMA-issuing CTA)
        for (int s = 0; s < 2; s++) { ptx::mbar_init(tfull_bar(s), 1); ptx::mbar_init(tempty_bar(s), TC_EPI_WARPS * CG); }
        ptx::fence_mbar_init();
    }
    if (warp == 2) ptx::tmem_alloc<CG>(tmem_slot, (uint32_t)p.tmem_cols);
    ptx::tc_fence_before();
    if (CG > 1) ptx::cluster_sync();                 // the peer's barriers must exist before any remote arrive / complete_tx
    else __syncthreads();
    ptx::tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    if (warp == 0) {
        // ===================== TMA producer (every CTA: its A rows, its slice of W) =====================
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            int tile_no = 0;
            for (int64_t ct = cluster_id; ct < p.cluster_tiles; ct += num_clusters, tile_no++) {
                const int64_t m_grp = ct / p.n_tiles;
                const int n_blk = (int)(ct - m_grp * p.n_tiles);
                const int64_t m_blk = m_grp * CG + rank;
                for (int kb = 0; kb < num_kb; kb++) {
                    ptx::mbar_wait(empty_bar(stage), phase ^ 1u);       // the MMAs that read this stage have retired
                    if (kb == 0) TRACE(tile_no, 4);
                    if (kb == num_kb - 1) TRACE(tile_no, 5);
                    const uint32_t sa = smem_base + (uint32_t)stage * stage_bytes;
                    if (p.debug & 2) {                                  // timing experiment: MMA on whatever is in shared memory
                        if (leader) ptx::mbar_arrive(full_bar(stage));
                    } else if (CG == 1) {
                        ptx::mbar_expect_tx(full_bar(stage), stage_bytes);
                        ptx::tma_load_2d(&tm_a, full_bar(stage), sa, kb * TC_BK, (int32_t)(m_blk * TC_BM));
                        ptx::tma_load_2d(&tm_w, full_bar(stage), sa + TC_A_BYTES, kb * TC_BK, n_blk * p.block_n);
                    } else {
                        // both CTAs' bytes are counted on the even CTA's barrier (the only one the MMA thread waits on)
                        if (leader) ptx::mbar_expect_tx(full_bar(stage), 2u * stage_bytes);
                        const uint32_t bar0 = full_bar(stage) & ptx::PEER_MASK;
                        ptx::tma_load_2d_2sm(&tm_a, bar0, sa, kb * TC_BK, (int32_t)(m_blk * TC_BM));
                        ptx::tma_load_2d_2sm(&tm_w, bar0, sa + TC_A_BYTES, kb * TC_BK, n_blk * p.block_n + (int)(rank * b_rows));
                    }
                    if (++stage == p.stages) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (one thread of the even CTA) =====================
        if (lane == 0 && leader) {
            const uint32_t idesc = ptx::umma_idesc_bf16(TC_BM * CG, (uint32_t)p.block_n);
            int stage = 0; uint32_t phase = 0;
            int as = 0; uint32_t aphase = 0;
            int tile_no = 0;
            for (int64_t ct = cluster_id; ct < p.cluster_tiles; ct += num_clusters, tile_no++) {
                TRACE(tile_no, 0);
                ptx::mbar_wait(tempty_bar(as), aphase ^ 1u);           // every epilogue warp (of both CTAs) drained this accumulator
                TRACE(tile_no, 1);
                ptx::tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(as * p.block_n);
                for (int kb = 0; kb < num_kb; kb++) {
                    ptx::mbar_wait(full_bar(stage), phase);              // TMA bytes (of both CTAs) have landed
                    if (kb == 0) TRACE(tile_no, 2);
                    ptx::tc_fence_after();
                    const uint32_t sa = smem_base + (uint32_t)stage * stage_bytes;
                    const uint64_t a_desc = ptx::umma_smem_desc_sw128(sa);
                    const uint64_t b_desc = ptx::umma_smem_desc_sw128(sa + TC_A_BYTES);
#pragma unroll
                    for (int k = 0; k < TC_BK / 16; k++)                 // +32 bytes per K=16 step inside the swizzle atom
                        ptx::umma_bf16<CG>(d_tmem, a_desc + (uint64_t)(k * 2), b_desc + (uint64_t)(k * 2), idesc, (kb | k) != 0 ? 1u : 0u);
                    ptx::umma_commit<CG>(empty_bar(stage));               // stage reusable (in both CTAs) once these MMAs retire
                    if (++stage == p.stages) { stage = 0; phase ^= 1u; }
                }
                ptx::umma_commit<CG>(tfull_bar(as));                      // accumulator complete (signalled in both CTAs)
                TRACE(tile_no, 3);
                if (++as == 2) { as = 0; aphase ^= 1u; }
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue (every CTA: its own 128 TMEM lanes) =====================
        const int q = warp & 3;                                           // TMEM lane quadrant this warp may access
        const int half = (warp - 4) >> 2;                                 // which of the quadrant's two warps
        float *stage = stage_base + (warp - 4) * 32 * TC_STAGE_PITCH;
        int as = 0; uint32_t aphase = 0;
        int tile_no = 0;
        for (int64_t ct = cluster_id; ct < p.cluster_tiles; ct += num_clusters, tile_no++) {
            const int64_t m_grp = ct / p.n_tiles;
            const int n_blk = (int)(ct - m_grp * p.n_tiles);
            const int64_t m_blk = m_grp * CG + rank;
            const int64_t row0 = m_blk * TC_BM + q * 32;
            if (warp == 4 && lane == 0) TRACE(tile_no, 6);
            const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * p.block_n);
            bool waited = false;
            auto wait_acc = [&]() {
                if (!waited) {
                    ptx::mbar_wait(tfull_bar(as), aphase);
                    ptx::tc_fence_after();
                    waited = true;
                    if (warp == 4 && lane == 0) TRACE(tile_no, 7);
                    if (warp == 8 && lane == 0) TRACE(tile_no, 13);
                }
            };
            if (p.debug & 1) {
                // timing experiment: drain nothing
            } else if constexpr (EPI == NMM_EPI_OUTPUT) {
                // y[b,c,f,p] = acc + bias[c] + x[b,c,f,p]: the output is channel-major, rows (p) are the contiguous axis
                if (e.nchw_vec) {
                    // 32 x 32 chunk through the transpose buffer (pitch 33: conflict-free both ways); afterwards 4 lanes own
                    // 32 consecutive positions of one channel -> 16-byte loads of x and stores of y, 64 contiguous bytes per channel
                    const int cq = lane >> 2, rg = lane & 3;
                    const int64_t rowg = row0 + 8 * rg;                      // first of this lane's 8 rows
                    const bool rows_ok = row0 < p.M;                          // M % 32 == 0 on this path: all-or-nothing per warp
                    const int64_t bfi = rowg / e.P;
                    const int pp = (int)(rowg - bfi * e.P);
                    const int64_t bb = bfi / e.F, ff = bfi - bb * e.F;
                    const bf16 *xrow = reinterpret_cast<const bf16 *>(e.x) + bb * e.xsb + ff * e.xsf + pp;
                    bf16 *yrow = reinterpret_cast<bf16 *>(e.y) + bb * e.ysb + ff * e.ysf + pp;
                    for (int c0 = half * 32; c0 < p.block_n; c0 += 64) {
                        const int width = min(32, p.block_n - c0);
                        const int col0 = n_blk * p.block_n + c0;
                        uint4 xin[4];
#pragma unroll
                        for (int j = 0; j < 4; j++) {
                            const int c = cq + 8 * j;
                            xin[j] = (rows_ok && c < width) ? __ldg(reinterpret_cast<const uint4 *>(xrow + (int64_t)(col0 + c) * e.xsc))
                                                            : make_uint4(0u, 0u, 0u, 0u);
                        }
                        wait_acc();
                        uint32_t lo[16], hi[16];
                        ptx::tmem_ld16(t_row + (uint32_t)c0, lo);
                        if (width > 16) ptx::tmem_ld16(t_row + (uint32_t)c0 + 16u, hi);
                        ptx::tmem_ld_wait();
                        uint32_t *srow = reinterpret_cast<uint32_t *>(stage) + lane * 33;
#pragma unroll
                        for (int j = 0; j < 16; j++) srow[j] = lo[j];
                        if (width > 16) {
#pragma unroll
                            for (int j = 0; j < 16; j++) srow[16 + j] = hi[j];
                        }
                        __syncwarp();
                        if (rows_ok) {
#pragma unroll
                            for (int j = 0; j < 4; j++) {
                                const int c = cq + 8 * j;
                                if (c < width) {
                                    const float bias = e.bias ? __ldg(e.bias + col0 + c) : 0.f;
                                    const float *sc = stage + (8 * rg) * 33 + c;
                                    const uint32_t xw[4] = {xin[j].x, xin[j].y, xin[j].z, xin[j].w};
                                    uint32_t o[4];
#pragma unroll
                                    for (int i = 0; i < 4; i++)
                                        o[i] = pack_bf16x2(sc[(2 * i) * 33] + bias + bf16_lo(xw[i]), sc[(2 * i + 1) * 33] + bias + bf16_hi(xw[i]));
                                    *reinterpret_cast<uint4 *>(yrow + (int64_t)(col0 + c) * e.ysc) = make_uint4(o[0], o[1], o[2], o[3]);
                                }
                            }
                        }
                        __syncwarp();
                    }
                } else {
                    // generic (ragged P / unaligned) path: one position per lane, 2-byte accesses coalesced along p
                    wait_acc();
                    const int64_t row = row0 + lane;
                    for (int c0 = half * 16; c0 < p.block_n; c0 += 32) {
                        uint32_t r[16];
                        ptx::tmem_ld16(t_row + (uint32_t)c0, r);
                        ptx::tmem_ld_wait();
                        const int col0 = n_blk * p.block_n + c0;
                        if (row < p.M && col0 < p.N) {
                            float acc[16];
#pragma unroll
                            for (int j = 0; j < 16; j++) acc[j] = __uint_as_float(r[j]);
                            epilogue_apply<EPI, bf16, 16>(e, row, col0, acc);
                        }
                    }
                }
            } else {
                const int cl = (lane & 7) * 4, rl = lane >> 3;
                for (int c0 = half * 32; c0 < p.block_n; c0 += 64) {
                    const int width = min(32, p.block_n - c0);
                    const int col0 = n_blk * p.block_n + c0;
                    float4 res[8];
                    if constexpr (EPI == NMM_EPI_RESIDUAL) {
                        // prefetch the residual rows in the coalesced layout before waiting on the accumulator
#pragma unroll
                        for (int it = 0; it < 8; it++) {
                            const int64_t row = row0 + it * 4 + rl;
                            res[it] = (cl < width && row < e.M) ? *reinterpret_cast<const float4 *>(e.h + row * e.N + col0 + cl)
                                                                : make_float4(0.f, 0.f, 0.f, 0.f);
                        }
                    }
                    wait_acc();
                    uint32_t r[32];
                    {
                        uint32_t lo[16], hi[16];
                        ptx::tmem_ld16(t_row + (uint32_t)c0, lo);
                        if (width > 16) ptx::tmem_ld16(t_row + (uint32_t)c0 + 16u, hi);
                        ptx::tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 16; j++) { r[j] = lo[j]; r[16 + j] = (width > 16) ? hi[j] : 0u; }
                    }
                    epi_chunk<EPI>(e, stage, lane, row0, col0, width, r, res, p.debug);
                    if (warp == 4 && lane == 0 && c0 / 64 < 4) TRACE(tile_no, 8 + c0 / 64);
                }
            }
            wait_acc();                                                   // a warp without a chunk in this tile still follows the phases
            ptx::tc_fence_before();
            __syncwarp();
            if (warp == 4 && lane == 0) TRACE(tile_no, 12);
            if (warp == 8 && lane == 0) TRACE(tile_no, 14);
            if (lane == 0) {
                if (CG == 1) ptx::mbar_arrive(tempty_bar(as));
                else ptx::mbar_arrive_cluster(tempty_bar(as) & ptx::PEER_MASK);      // the even CTA's barrier
            }
            if (++as == 2) { as = 0; aphase ^= 1u; }
        }
    }
    __syncwarp();                                    // lanes 1-31 of the single-thread roles rejoin lane 0 before the aligned barrier
    ptx::tc_fence_before();
    if (CG > 1) ptx::cluster_sync();                 // the peer may still read this CTA's shared memory / arrive on its barriers
    else __syncthreads();
    if (warp == 2) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc<CG>(tmem_base, (uint32_t)p.tmem_cols);
    }
}

// ---- host side ----------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

// [rows, cols] row-major bf16 with leading dimension ld (elements); box = box_rows x 64 columns, 128-byte swizzle.
static int make_tmap(CUtensorMap *tm, const void *ptr, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return fail(NMM_ERR_DEVICE, "cuTensorMapEncodeTiled not available (no CUDA driver?)");
    if (!aligned(ptr, 16) || (ld * 2) % 16 != 0) return fail(NMM_ERR_BAD_ARG, "TMA operand must be 16-byte aligned with a 16-byte multiple row pitch");
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(ptr), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(NMM_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return NMM_OK;
}

#ifdef NMM_TRACE
static unsigned long long *g_trace_dev = nullptr;
extern "C" __attribute__((visibility("default"))) int nmm_debug_trace_dump(const char *path) {
    if (!g_trace_dev) return -1;
    static unsigned long long host[TRACE_TILES * TRACE_SLOTS];
    cudaDeviceSynchronize();
    cudaMemcpy(host, g_trace_dev, sizeof(host), cudaMemcpyDeviceToHost);
    FILE *f = fopen(path, "w");
    if (!f) return -2;
    for (int t = 0; t < TRACE_TILES; t++) {
        for (int s = 0; s < TRACE_SLOTS; s++) fprintf(f, "%llu ", host[t * TRACE_SLOTS + s]);
        fprintf(f, "\n");
    }
    fclose(f);
    cudaMemset(g_trace_dev, 0, sizeof(host));
    return 0;
}
#endif

static int num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    }
    return n;
}

// N tile: a multiple of 16 in [16, 256] that divides N; the largest one that still yields >= 2 waves of tiles,
// otherwise the largest that yields >= 1 wave, otherwise the smallest divisor >= 64 (small-M levels of the UNet).
static int choose_block_n(int64_t m_tiles, int N, int sms) {
    int best2 = 0, best1 = 0, smallest = 0;
    for (int bn = 256; bn >= 16; bn -= 16) {
        if (N % bn) continue;
        const int64_t tiles = m_tiles * (N / bn);
        if (!best2 && tiles >= 2 * sms) best2 = bn;
        if (!best1 && tiles >= sms) best1 = bn;
        if (bn >= 64 || smallest == 0) smallest = bn;
    }
    if (best2) return best2;
    if (best1) return best1;
    return smallest;
}

template <int EPI, int CG>
static int launch_tc_t(const CUtensorMap &ta, const CUtensorMap &tw, const TcParams &p, const EpiParams &e, size_t smem, int grid,
                       cudaStream_t st, double flops, double bytes) {
    auto kern = linear_tc_kernel<EPI, CG>;
    static bool attr_set = false;     // per template instantiation
    if (!attr_set) {
        NMM_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BUDGET + 2048));
        attr_set = true;
    }
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(TC_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)CG;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = CG > 1 ? 1 : 0;
    {
        ProfScope prof(K_LINEAR_TC, st, flops, bytes);
        cudaError_t le = cudaLaunchKernelEx(&cfg, kern, ta, tw, p, e);
        if (le != cudaSuccess) return fail(NMM_ERR_CUDA, "cudaLaunchKernelEx(linear_tc_kernel) failed: %s", cudaGetErrorString(le));
    }
    NMM_LAUNCHED("linear_tc_kernel");
    return NMM_OK;
}

int launch_linear_tc(const LinearArgs &a, cudaStream_t st) {
    if (a.N % 16 != 0) return fail(NMM_ERR_UNSUPPORTED, "tcgen05 GEMM needs N %% 16 == 0 (N=%d)", a.N);
    if (a.K % 8 != 0) return fail(NMM_ERR_UNSUPPORTED, "tcgen05 GEMM needs K %% 8 == 0 (K=%d)", a.K);
    if (a.M <= 0) return NMM_OK;
    TcParams p;
    p.M = a.M; p.N = a.N; p.K = a.K;
    p.m_tiles = ceil_div(a.M, TC_BM);
    const int sms = num_sms();
    // CTA pairs along M (tcgen05 cta_group::2): each CTA stages half of every W tile, which cuts the L2 -> SM fill per FLOP
    // by a third -- the quantity that bounds the single-CTA kernel.  148 SMs = 74 pairs.  NMM_GEMM_CLUSTER=1|2 forces a mode.
    static const int force_cluster = getenv("NMM_GEMM_CLUSTER") ? atoi(getenv("NMM_GEMM_CLUSTER")) : 0;
    static const int debug_flags = getenv("NMM_GEMM_DEBUG") ? atoi(getenv("NMM_GEMM_DEBUG")) : 0;
    static const int force_bn = getenv("NMM_GEMM_BLOCK_N") ? atoi(getenv("NMM_GEMM_BLOCK_N")) : 0;
    p.debug = debug_flags;
    p.trace = nullptr;
#ifdef NMM_TRACE
    if (!g_trace_dev) { cudaMalloc(&g_trace_dev, TRACE_TILES * TRACE_SLOTS * 8); cudaMemset(g_trace_dev, 0, TRACE_TILES * TRACE_SLOTS * 8); }
    p.trace = g_trace_dev;
#endif
    p.cluster = (p.m_tiles >= 2 && sms % 2 == 0) ? 2 : 1;
    if (force_cluster == 1 || force_cluster == 2) p.cluster = force_cluster;
    const int64_t m_groups = ceil_div(p.m_tiles, p.cluster);
    p.block_n = choose_block_n(m_groups * p.cluster, a.N, sms);
    if (force_bn >= 16 && force_bn <= 256 && force_bn % 16 == 0 && a.N % force_bn == 0) p.block_n = force_bn;
    p.n_tiles = a.N / p.block_n;
    p.cluster_tiles = m_groups * p.n_tiles;
    const size_t stage_bytes = (size_t)TC_A_BYTES + (size_t)(p.block_n / p.cluster) * TC_BK * 2;     // per CTA
    int stages = (int)((TC_SMEM_BUDGET - 1024 - 512 - TC_STAGE_BYTES) / stage_bytes);
    if (stages > 8) stages = 8;
    if (stages < 2) return fail(NMM_ERR_UNSUPPORTED, "tcgen05 GEMM: tile does not fit shared memory");
    p.stages = stages;
    int cols = 32;
    while (cols < 2 * p.block_n) cols <<= 1;
    p.tmem_cols = cols;
    // alignment slack + operand ring + 256 B of barriers + transpose buffers
    const size_t smem = 1024 + (size_t)stages * stage_bytes + 256 + TC_STAGE_BYTES;
    CUtensorMap ta, tw;
    int rc = make_tmap(&ta, a.A, a.M, a.K, a.K, TC_BM);
    if (rc != NMM_OK) return rc;
    rc = make_tmap(&tw, a.W, a.N, a.K, a.K, p.block_n / p.cluster);      // each CTA of a cluster fetches its slice of the W tile
    if (rc != NMM_OK) return rc;
    const int64_t ctas = p.cluster_tiles * p.cluster;
    const int max_grid = sms / p.cluster * p.cluster;
    const int grid = (int)(ctas < max_grid ? ctas : max_grid);
    EpiParams e = epi_params_of(a);
    const double fl = linear_flops(a), by = linear_bytes(a, 2);
#define TC_DISPATCH(EPI)                                                                                   \
    return p.cluster == 2 ? launch_tc_t<EPI, 2>(ta, tw, p, e, smem, grid, st, fl, by)                      \
                          : launch_tc_t<EPI, 1>(ta, tw, p, e, smem, grid, st, fl, by)
    switch (a.epilogue) {
        case NMM_EPI_STORE: TC_DISPATCH(NMM_EPI_STORE);
        case NMM_EPI_RESIDUAL: TC_DISPATCH(NMM_EPI_RESIDUAL);
        case NMM_EPI_GEGLU: TC_DISPATCH(NMM_EPI_GEGLU);
        case NMM_EPI_OUTPUT: TC_DISPATCH(NMM_EPI_OUTPUT);
        default: return fail(NMM_ERR_BAD_ARG, "unknown epilogue %d", a.epilogue);
    }
#undef TC_DISPATCH
}

}  // namespace nmm
