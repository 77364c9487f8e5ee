// Microbenchmark: tcgen05.ld (TMEM -> registers) throughput per SM, alone and concurrent with tcgen05.mma.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_bw tmem_bw.cu && ./tmem_bw
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int X>
__device__ __forceinline__ void ld(uint32_t taddr, uint32_t& sink) {
    if constexpr (X == 16) {
        uint32_t r[16];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                       "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(taddr) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int i = 0; i < 16; i++) sink ^= r[i];
    } else {
        uint32_t r[32];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                       "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
                       "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]),
                       "=r"(r[30]), "=r"(r[31]) : "r"(taddr) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int i = 0; i < 32; i++) sink ^= r[i];
    }
}

// mode bit0: epilogue warps do LDTM; bit1: one thread issues MMAs (M=128,N=256,K=16 bf16 on zeroed smem)
template <int X>
__global__ void __launch_bounds__(384, 1) k(int iters, int nwarps_ld, int mode, long long* cycles, uint32_t* out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint32_t tmem_slot;
    __shared__ uint64_t bar;
    __shared__ long long t_end[12];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (threadIdx.x == 32) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tbase = tmem_slot;
    long long t0 = clock64();
    uint32_t sink = 0;
    if (warp == 1 && (mode & 2)) {
        if (lane == 0) {
            const uint32_t sa = (smem_u32(smem) + 1023u) & ~1023u;
            uint64_t adesc = ((uint64_t)((sa & 0x3FFFF) >> 4)) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
            uint64_t bdesc = ((uint64_t)(((sa + 16384) & 0x3FFFF) >> 4)) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((256u >> 3) << 17) | ((128u >> 4) << 24);
            for (int i = 0; i < iters; i++) {
                // 20 MMAs (one K=320 tile) into columns 256..511
                for (int j = 0; j < 20; j++) {
                    uint32_t acc = j > 0;
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                                 ::"r"(tbase + 256u), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
                }
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
            // wait for completion
            uint32_t ok = 0;
            while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
        }
    } else if (warp >= 4 && warp < 4 + nwarps_ld && (mode & 1)) {
        const int q = warp & 3;
        const uint32_t trow = tbase + ((uint32_t)(q * 32) << 16);
        for (int i = 0; i < iters; i++) {
            // read a 256-column accumulator (this warp's share: columns split between the warps of the quadrant)
            const int share = nwarps_ld / 4, me = (warp - 4) / 4;
            for (int c = me * X; c < 256; c += share * X) ld<X>(trow + c, sink);
        }
    }
    __syncwarp();
    if (lane == 0) t_end[warp] = clock64();          // per-warp end time (BAR.SYNC does not block a following clock read)
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
        long long t1 = 0;
        for (int w = 0; w < 12; w++) t1 = t_end[w] > t1 ? t_end[w] : t1;
        cycles[blockIdx.x] = t1 - t0;
    }
    if (sink == 0x12345678) out[0] = sink;
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(512u) : "memory");
}

template <int X>
void run(const char* name, int nwarps_ld, int mode) {
    long long* cyc; uint32_t* out;
    cudaMalloc(&cyc, 148 * 8); cudaMalloc(&out, 4);
    cudaFuncSetAttribute(k<X>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    const int iters = 200;
    k<X><<<148, 384, 64 * 1024>>>(iters, nwarps_ld, mode, cyc, out);
    cudaError_t le = cudaGetLastError();
    cudaError_t e = cudaDeviceSynchronize();
    if (le != cudaSuccess) printf("launch error: %s\n", cudaGetErrorString(le));
    long long h[148];
    cudaMemcpy(h, cyc, 148 * 8, cudaMemcpyDeviceToHost);
    long long mx = 0; for (int i = 0; i < 148; i++) mx = h[i] > mx ? h[i] : mx;
    double per_tile = (double)mx / iters;
    printf("  raw cycles: max %lld, sm0 %lld sm1 %lld\n", mx, h[0], h[1]);
    printf("%-44s x%-3d warps=%d mode=%d : %8.0f cycles/tile  (LDTM %6.1f B/clk/SM, MMA %5.1f%% of 2560-cycle floor) %s\n", name, X, nwarps_ld, mode,
           per_tile, (mode & 1) ? 128.0 * 256 * 4 / per_tile : 0.0, (mode & 2) ? 100.0 * 2560 / per_tile : 0.0, e == cudaSuccess ? "" : cudaGetErrorString(e));
    cudaFree(cyc); cudaFree(out);
}

int main() {
    run<16>("LDTM only", 4, 1);
    run<32>("LDTM only", 4, 1);
    run<16>("LDTM only", 8, 1);
    run<32>("LDTM only", 8, 1);
    run<32>("MMA only (20 x M128 N256 K16 per tile)", 4, 2);
    run<32>("LDTM + MMA concurrently", 4, 3);
    run<32>("LDTM + MMA concurrently", 8, 3);
    run<16>("LDTM + MMA concurrently", 8, 3);
    return 0;
}
