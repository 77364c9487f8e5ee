// How fast is legacy mma.sync (HMMA) on B200?  One warp per SM sub-partition issues chains of m16n8k16 bf16 MMAs:
// dependent (same accumulator) -> latency; independent (4 accumulators) -> issue interval.  Also ldmatrix latency.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_lat mma_sync_latency.cu && ./mma_lat
#include <cstdio>
#include <cstdint>
__device__ __forceinline__ void mma16(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__global__ void k(long long *out, int warps_active) {
    __shared__ __align__(16) uint32_t sm[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = 0x3f803f80u;
    __syncthreads();
    if ((threadIdx.x >> 5) >= warps_active) return;
    float d[4][4] = {};
    uint32_t a = 0x3f803f80u, b = 0x3c003c00u;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 64; i++) { mma16(d[0], a, a, a, a, b, b); }           // dependent chain
    long long t1 = clock64();
#pragma unroll 1
    for (int i = 0; i < 16; i++) { mma16(d[0], a, a, a, a, b, b); mma16(d[1], a, a, a, a, b, b); mma16(d[2], a, a, a, a, b, b); mma16(d[3], a, a, a, a, b, b); }
    long long t2 = clock64();
    uint32_t r0, r1, r2, r3, addr = (uint32_t)__cvta_generic_to_shared(sm) + (threadIdx.x & 31) * 16;
#pragma unroll 1
    for (int i = 0; i < 64; i++) {
        asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
        addr ^= (r0 & 16);                                                    // dependent address
    }
    long long t3 = clock64();
    if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t1; out[2] = t3 - t2; }
    if (d[0][0] + d[1][1] + d[2][2] + d[3][3] + r1 == 12345.f) out[3] = 1;
}
int main() {
    long long *o; cudaMallocManaged(&o, 64);
    for (int w : {1, 4, 8, 16}) {
        k<<<1, 512>>>(o, w); cudaDeviceSynchronize();
        printf("warps %2d: dependent mma %5.1f cyc each | 4 independent %5.1f cyc each | dependent ldmatrix.x4 %5.1f cyc\n", w, o[0] / 64.0, o[1] / 64.0, o[2] / 64.0);
    }
    return 0;
}
