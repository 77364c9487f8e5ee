// fp32 parity-mode GEMM:  D[M,N] = A[M,K] . W[N,K]^T  on the fp32 FMA pipe, fused epilogues.
//
// Used only when nmm_shape.dtype == NMM_F32 (bar: max-abs 1e-4 against the fp32 reference, which single-pass
// bf16/TF32 tensor-core products cannot meet -- SURVEY 7 "fp32 mode").  The production path is the tcgen05
// kernel in gemm_tcgen05.cu.  Reference arithmetic: every nn.Linear on the path (motion_module.py:145,152,
// 289,297,298,321; motion_module_new.py:466,516).
#include "common.cuh"
#include "epilogue.cuh"

namespace nmm {

constexpr int SG_BM = 128, SG_BN = 64, SG_BK = 16, SG_THREADS = 256;

template <int EPI>
__global__ void __launch_bounds__(SG_THREADS) linear_simt_kernel(const float *__restrict__ A, const float *__restrict__ W,
                                                                 int K, EpiParams e) {
    pdl_wait();                    // PDL: the previous kernel has completed (no-op without the launch attribute)
    pdl_launch_dependents();       // let the next kernel's launch + prologue overlap this kernel
    __shared__ __align__(16) float As[SG_BK][SG_BM + 4];
    __shared__ __align__(16) float Bs[SG_BK][SG_BN + 4];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int64_t m0 = (int64_t)blockIdx.y * SG_BM;
    const int n0 = blockIdx.x * SG_BN;
    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < K; k0 += SG_BK) {
#pragma unroll
        for (int r = 0; r < 2; r++) {
            const int i = tid + r * SG_THREADS;
            const int row = i >> 2, kq = i & 3;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (m0 + row < e.M) v = __ldg(reinterpret_cast<const float4 *>(A + (m0 + row) * K + k0 + kq * 4));
            As[kq * 4 + 0][row] = v.x; As[kq * 4 + 1][row] = v.y; As[kq * 4 + 2][row] = v.z; As[kq * 4 + 3][row] = v.w;
        }
        {
            const int row = tid >> 2, kq = tid & 3;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (n0 + row < e.N) v = __ldg(reinterpret_cast<const float4 *>(W + (int64_t)(n0 + row) * K + k0 + kq * 4));
            Bs[kq * 4 + 0][row] = v.x; Bs[kq * 4 + 1][row] = v.y; Bs[kq * 4 + 2][row] = v.z; Bs[kq * 4 + 3][row] = v.w;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < SG_BK; k++) {
            float4 a0 = *reinterpret_cast<const float4 *>(&As[k][ty * 8]);
            float4 a1 = *reinterpret_cast<const float4 *>(&As[k][ty * 8 + 4]);
            float4 b = *reinterpret_cast<const float4 *>(&Bs[k][tx * 4]);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 8; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
        }
        __syncthreads();
    }
    const int col0 = n0 + tx * 4;
    if (col0 >= e.N) return;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const int64_t row = m0 + ty * 8 + i;
        if (row < e.M) epilogue_apply<EPI, float, 4>(e, row, col0, acc[i]);
    }
}

int launch_linear_simt(const LinearArgs &a, cudaStream_t st) {
    if (a.K % SG_BK != 0 || a.N % 4 != 0) return fail(NMM_ERR_UNSUPPORTED, "fp32 GEMM needs K %% 16 == 0 and N %% 4 == 0 (K=%d N=%d)", a.K, a.N);
    if (!aligned(a.A, 16) || !aligned(a.W, 16)) return fail(NMM_ERR_BAD_ARG, "fp32 GEMM operands must be 16-byte aligned");
    if (a.M <= 0) return NMM_OK;
    const int64_t mt = ceil_div(a.M, SG_BM);
    if (mt > 65535) return fail(NMM_ERR_UNSUPPORTED, "fp32 GEMM: too many rows");
    dim3 grid((a.N + SG_BN - 1) / SG_BN, (unsigned)mt), block(SG_THREADS);
    EpiParams e = epi_params_of(a);
    const float *A = (const float *)a.A, *W = (const float *)a.W;
    ProfScope prof(K_LINEAR_SIMT, st, linear_flops(a), linear_bytes(a, 4));
    switch (a.epilogue) {
        case NMM_EPI_STORE: launch_pdl(linear_simt_kernel<NMM_EPI_STORE>, grid, block, 0, st, A, W, a.K, e); break;
        case NMM_EPI_RESIDUAL: launch_pdl(linear_simt_kernel<NMM_EPI_RESIDUAL>, grid, block, 0, st, A, W, a.K, e); break;
        case NMM_EPI_GEGLU: launch_pdl(linear_simt_kernel<NMM_EPI_GEGLU>, grid, block, 0, st, A, W, a.K, e); break;
        case NMM_EPI_OUTPUT: launch_pdl(linear_simt_kernel<NMM_EPI_OUTPUT>, grid, block, 0, st, A, W, a.K, e); break;
        default: return fail(NMM_ERR_BAD_ARG, "unknown epilogue %d", a.epilogue);
    }
    NMM_LAUNCHED("linear_simt_kernel");
    return NMM_OK;
}

}  // namespace nmm
