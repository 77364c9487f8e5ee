// Micro-probes behind the fused C = 320 module kernel (fused_module.cu), run once on a B200:
//   1. correctness of a NO-SWIZZLE ("interleave", 8-row x 16-byte core matrices) K-major A descriptor next to a SW128 B operand
//      (which of LBO / SBO is the K-chunk stride and which the 8-row-group stride);
//   2. tcgen05.mma issue-to-retire cycles per instruction against N (cta_group::1, M = 128, K = 16) for both A layouts:
//      is a small-N MMA bound by the 4 KB A re-read from shared memory?
//   3. per-SM TMA ingest from an L2-resident 4.7 MB weight buffer with all SMs pulling at once (the W stream of the fused kernel).
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o scripts/micro/bin/umma_probe scripts/micro/umma_probe.cu
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../neurons_b200/csrc/ptx.cuh"

using namespace nmm;
typedef __nv_bfloat16 bf16;

#define CK(x)                                                                                  \
    do {                                                                                       \
        cudaError_t e_ = (x);                                                                  \
        if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } \
    } while (0)

__device__ __forceinline__ uint64_t desc_interleave(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;     // layout type 0 = no swizzle
}

// ---- probe 1 ------------------------------------------------------------------------------------------------------------------
// mode 0: A sw128 (control); 1: A interleave, LBO = K-chunk stride, SBO = 8-row stride; 2: swapped
__global__ void __launch_bounds__(128, 1) probe_desc(const bf16 *A, const bf16 *B, float *D, int N, int mode) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char *sm = smem_raw + (base - ptx::smem_u32(smem_raw));
    const uint32_t a_off = 0, b_off = 16384, bar_off = 16384 + 32768, slot_off = bar_off + 8;
    const int t = threadIdx.x, warp = t >> 5;
    // A [128 x 64]
    for (int i = t; i < 128 * 64; i += 128) {
        const int r = i / 64, k = i % 64;
        uint32_t off;
        if (mode == 0) off = r * 128 + (((k / 8) ^ (r & 7)) << 4) + (k % 8) * 2;
        else off = (k / 8) * 2048 + r * 16 + (k % 8) * 2;
        *reinterpret_cast<bf16 *>(sm + a_off + off) = A[i];
    }
    // B [N x 64], SW128 K-major
    for (int i = t; i < N * 64; i += 128) {
        const int r = i / 64, k = i % 64;
        *reinterpret_cast<bf16 *>(sm + b_off + r * 128 + (((k / 8) ^ (r & 7)) << 4) + (k % 8) * 2) = B[i];
    }
    if (t == 0) { ptx::mbar_init(base + bar_off, 1); ptx::fence_mbar_init(); }
    if (warp == 0) ptx::tmem_alloc<1>(base + slot_off, 256);
    ptx::fence_proxy_async();
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<uint32_t *>(sm + slot_off);
    if (t == 0) {
        const uint32_t idesc = ptx::umma_idesc_bf16(128, N);
        for (int k = 0; k < 4; k++) {
            uint64_t ad;
            if (mode == 0) ad = ptx::umma_smem_desc_sw128(base + a_off) + (uint64_t)(k * 2);
            else if (mode == 1) ad = desc_interleave(base + a_off + k * 2 * 2048, 2048, 128);
            else ad = desc_interleave(base + a_off + k * 2 * 2048, 128, 2048);
            const uint64_t bd = ptx::umma_smem_desc_sw128(base + b_off) + (uint64_t)(k * 2);
            ptx::umma_bf16<1>(tmem, ad, bd, idesc, k > 0 ? 1u : 0u);
        }
        ptx::umma_commit<1>(base + bar_off);
    }
    ptx::mbar_wait(base + bar_off, 0);
    ptx::tc_fence_after();
    const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
    for (int c0 = 0; c0 < N; c0 += 16) {
        uint32_t r[16];
        ptx::tmem_ld16(trow + c0, r);
        ptx::tmem_ld_wait();
        for (int j = 0; j < 16; j++) D[(size_t)t * N + c0 + j] = __uint_as_float(r[j]);
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 0) ptx::tmem_dealloc<1>(tmem, 256);
}

// ---- probe 2 ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1) probe_mma_rate(int N, int a_mode, int iters, long long *cycles) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char *sm = smem_raw + (base - ptx::smem_u32(smem_raw));
    const uint32_t a_off = 0, b_off = 16384, bar_off = 16384 + 32768, slot_off = bar_off + 8;
    const int t = threadIdx.x, warp = t >> 5;
    for (int i = t; i < (16384 + 32768) / 4; i += 128) reinterpret_cast<uint32_t *>(sm)[i] = 0;
    if (t == 0) { ptx::mbar_init(base + bar_off, 1); ptx::fence_mbar_init(); }
    if (warp == 0) ptx::tmem_alloc<1>(base + slot_off, 256);
    ptx::fence_proxy_async();
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<uint32_t *>(sm + slot_off);
    if (t == 0) {
        const uint32_t idesc = ptx::umma_idesc_bf16(128, N);
        const long long t0 = clock64();
        for (int i = 0; i < iters; i++) {
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const uint64_t ad = a_mode == 0 ? ptx::umma_smem_desc_sw128(base + a_off) + (uint64_t)(k * 2) : desc_interleave(base + a_off + k * 2 * 2048, 2048, 128);
                const uint64_t bd = ptx::umma_smem_desc_sw128(base + b_off) + (uint64_t)(k * 2);
                ptx::umma_bf16<1>(tmem, ad, bd, idesc, 1u);
            }
        }
        ptx::umma_commit<1>(base + bar_off);
        ptx::mbar_wait(base + bar_off, 0);
        const long long t1 = clock64();
        cycles[blockIdx.x] = t1 - t0;
    }
    __syncthreads();
    ptx::tc_fence_after();
    if (warp == 0) ptx::tmem_dealloc<1>(tmem, 256);
}

// ---- probe 3 ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1) probe_tma_bw(const __grid_constant__ CUtensorMap tm, int nbox_mask, int stages_log2, int box_bytes, int iters, long long *cycles) {
    // stages and the number of boxes are powers of two: no integer division in the issue loop (the first version of this probe
    // measured its own loop overhead, ~400 cycles per iteration)
    extern __shared__ unsigned char smem_raw[];
    const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
    const int stages = 1 << stages_log2;
    const uint32_t bar0 = base + (uint32_t)stages * (uint32_t)box_bytes;
    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; s++) ptx::mbar_init(bar0 + 8 * s, 1);
        ptx::fence_mbar_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const int rows = box_bytes / 128;
        int box = (blockIdx.x * 37) & nbox_mask;
        const long long t0 = clock64();
        for (int i = 0; i < iters + stages; i++) {
            const int s = i & (stages - 1);
            if (i >= stages) ptx::mbar_wait(bar0 + 8 * s, ((i >> stages_log2) - 1) & 1);
            if (i < iters) {
                ptx::mbar_expect_tx(bar0 + 8 * s, (uint32_t)box_bytes);
                ptx::tma_load_2d(&tm, bar0 + 8 * s, base + (uint32_t)s * (uint32_t)box_bytes, 0, box * rows);
                box = (box + 1) & nbox_mask;
            }
        }
        cycles[blockIdx.x] = clock64() - t0;
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    CK(cudaSetDevice(0));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    printf("device %s, %d SMs\n", prop.name, sms);

    // ---- probe 1
    {
        const int N = 80;
        std::vector<bf16> hA(128 * 64), hB(N * 64);
        srand(1);
        for (auto &v : hA) v = __float2bfloat16((float)(rand() % 9 - 4));
        for (auto &v : hB) v = __float2bfloat16((float)(rand() % 9 - 4));
        std::vector<float> ref(128 * N);
        for (int m = 0; m < 128; m++)
            for (int n = 0; n < N; n++) {
                float s = 0;
                for (int k = 0; k < 64; k++) s += __bfloat162float(hA[m * 64 + k]) * __bfloat162float(hB[n * 64 + k]);
                ref[m * N + n] = s;
            }
        bf16 *dA, *dB;
        float *dD;
        CK(cudaMalloc(&dA, hA.size() * 2)); CK(cudaMalloc(&dB, hB.size() * 2)); CK(cudaMalloc(&dD, ref.size() * 4));
        CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
        CK(cudaFuncSetAttribute(probe_desc, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        for (int mode = 0; mode < 3; mode++) {
            CK(cudaMemset(dD, 0, ref.size() * 4));
            probe_desc<<<1, 128, 64 * 1024>>>(dA, dB, dD, N, mode);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("probe1 mode %d: CUDA error %s\n", mode, cudaGetErrorString(e)); return 1; }
            std::vector<float> out(ref.size());
            CK(cudaMemcpy(out.data(), dD, out.size() * 4, cudaMemcpyDeviceToHost));
            double md = 0;
            for (size_t i = 0; i < out.size(); i++) md = fmax(md, fabs((double)out[i] - ref[i]));
            printf("probe1 A-layout mode %d (0 sw128, 1 interleave LBO=Kchunk SBO=8rows, 2 swapped): max|diff| = %g %s\n", mode, md, md == 0 ? "EXACT" : "WRONG");
        }
    }
    // ---- probe 2
    {
        long long *dc;
        CK(cudaMalloc(&dc, sms * 8));
        CK(cudaFuncSetAttribute(probe_mma_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        const int Ns[] = {32, 48, 64, 80, 96, 128, 160, 192, 240, 256};
        const int iters = 2000;
        for (int am = 0; am < 2; am++)
            for (int N : Ns) {
                for (int rep = 0; rep < 2; rep++) {
                    probe_mma_rate<<<sms, 128, 64 * 1024>>>(N, am, iters, dc);
                    CK(cudaDeviceSynchronize());
                }
                std::vector<long long> hc(sms);
                CK(cudaMemcpy(hc.data(), dc, sms * 8, cudaMemcpyDeviceToHost));
                double mean = 0;
                for (auto v : hc) mean += (double)v;
                mean /= sms;
                printf("probe2 A=%s N=%3d: %.1f cycles per K16 MMA (ideal N/2 = %d)\n", am ? "interleave" : "sw128     ", N, mean / (iters * 4.0), N / 2);
            }
    }
    // ---- probe 3
    {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres));
        EncodeTiledFn fn = (EncodeTiledFn)p;
        const int64_t rows = 36864;      // x 128 B = 4.7 MB, stays in L2
        bf16 *dW;
        CK(cudaMalloc(&dW, rows * 128));
        CK(cudaMemset(dW, 0, rows * 128));
        long long *dc;
        CK(cudaMalloc(&dc, sms * 8));
        CK(cudaFuncSetAttribute(probe_tma_bw, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        const int boxes_rows[] = {40, 80, 128, 160};
        for (int br : boxes_rows) {
            CUtensorMap tm;
            cuuint64_t dims[2] = {64, (cuuint64_t)rows};
            cuuint64_t strides[1] = {128};
            cuuint32_t box[2] = {64, (cuuint32_t)br};
            cuuint32_t estr[2] = {1, 1};
            CUresult r = fn(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dW, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
            const int box_bytes = br * 128;
            int nbox = 1;
            while (nbox * 2 * br <= rows) nbox *= 2;          // power of two
            const int grids[] = {sms, sms / 2, 16};
            for (int grid : grids)
                for (int sl = 0; sl <= 3; sl++) {
                    const int stages = 1 << sl;
                    if ((size_t)stages * box_bytes > 180 * 1024) continue;
                    const int iters = 4000;
                    const size_t smem = (size_t)stages * box_bytes + 2048;
                    for (int rep = 0; rep < 2; rep++) {
                        probe_tma_bw<<<grid, 128, smem>>>(tm, nbox - 1, sl, box_bytes, iters, dc);
                        CK(cudaDeviceSynchronize());
                    }
                    std::vector<long long> hc(grid);
                    CK(cudaMemcpy(hc.data(), dc, grid * 8, cudaMemcpyDeviceToHost));
                    double mean = 0, mx = 0;
                    for (auto v : hc) { mean += (double)v; mx = fmax(mx, (double)v); }
                    mean /= grid;
                    printf("probe3 box %3d rows (%5d B) grid %3d stages %2d (%3d KB in flight): %.1f B/clk/SM mean, %.1f at the slowest; chip %.0f B/clk; %.0f cycles per box\n", br,
                           box_bytes, grid, stages, stages * box_bytes / 1024, (double)iters * box_bytes / mean, (double)iters * box_bytes / mx,
                           (double)iters * box_bytes / mean * grid, mean / iters);
                }
        }
    }
    return 0;
}
