// Classifier-free guidance + deterministic DDIM update, fused into one elementwise pass.
// Reference loop: pipeline_neuroclips.py:478-483
//     noise_pred = noise_pred_uncond + guidance_scale * (noise_pred_text - noise_pred_uncond)
//     latents = scheduler.step(noise_pred, t, latents).prev_sample
// with diffusers' DDIMScheduler.step (eta = 0, epsilon prediction, no clipping; un-vendored diffusers 0.11.1, restated in
// neurons_b200/sampler.py):  x0 = (x - sqrt(1 - a_t) eps) / sqrt(a_t);  x_prev = sqrt(a_prev) x0 + sqrt(1 - a_prev) eps.
// HBM-bound: reads latents + 2 eps, writes latents = 4 passes of n elements instead of the ~14 of the unfused torch expression.
#include "common.cuh"

namespace nmm {

template <typename T> struct Pack16;
template <> struct Pack16<float> { static constexpr int N = 4; };
template <> struct Pack16<bf16> { static constexpr int N = 8; };

template <typename T>
__device__ __forceinline__ void unpack16(const uint4 &v, float (&f)[Pack16<T>::N]) {
    if constexpr (sizeof(T) == 4) {
        f[0] = __uint_as_float(v.x); f[1] = __uint_as_float(v.y); f[2] = __uint_as_float(v.z); f[3] = __uint_as_float(v.w);
    } else {
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int i = 0; i < 4; i++) { f[2 * i] = bf16_lo(w[i]); f[2 * i + 1] = bf16_hi(w[i]); }
    }
}
template <typename T>
__device__ __forceinline__ uint4 pack16(const float (&f)[Pack16<T>::N]) {
    if constexpr (sizeof(T) == 4) return make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]), __float_as_uint(f[3]));
    else return make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
}

// x_prev = cx * x + ce * eps,   eps = eu + g * (ec - eu)
//   cx = sqrt(a_prev / a_t),  ce = sqrt(1 - a_prev) - sqrt(a_prev) * sqrt(1 - a_t) / sqrt(a_t)      (computed in double on the host)
template <typename T, bool VEC>
__global__ void __launch_bounds__(256) cfg_ddim_kernel(T *__restrict__ x, const T *__restrict__ eu, const T *__restrict__ ec, int64_t n,
                                                       float g, float cx, float ce) {
    pdl_wait();
    pdl_launch_dependents();
    if constexpr (VEC) {
        constexpr int V = Pack16<T>::N;
        const int64_t nv = n / V;
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += (int64_t)gridDim.x * blockDim.x) {
            float xf[V], uf[V], cf[V];
            unpack16<T>(reinterpret_cast<const uint4 *>(x)[i], xf);
            unpack16<T>(__ldg(reinterpret_cast<const uint4 *>(eu) + i), uf);
            if (ec) unpack16<T>(__ldg(reinterpret_cast<const uint4 *>(ec) + i), cf);
#pragma unroll
            for (int k = 0; k < V; k++) {
                const float eps = ec ? fmaf(g, cf[k] - uf[k], uf[k]) : uf[k];
                xf[k] = fmaf(ce, eps, cx * xf[k]);
            }
            reinterpret_cast<uint4 *>(x)[i] = pack16<T>(xf);
        }
    } else {
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
            const float u = to_f32(eu[i]);
            const float eps = ec ? fmaf(g, to_f32(ec[i]) - u, u) : u;
            x[i] = from_f32<T>(fmaf(ce, eps, cx * to_f32(x[i])));
        }
    }
}

int launch_cfg_ddim(int dtype, int64_t n, void *x, const void *eu, const void *ec, float g, double a_t, double a_prev, cudaStream_t st) {
    if (n <= 0) return NMM_OK;
    if (!(a_t > 0.0 && a_t <= 1.0 && a_prev > 0.0 && a_prev <= 1.0)) return fail(NMM_ERR_BAD_ARG, "alphas_cumprod must be in (0, 1]");
    const float cx = (float)sqrt(a_prev / a_t);
    const float ce = (float)(sqrt(1.0 - a_prev) - sqrt(a_prev) * sqrt(1.0 - a_t) / sqrt(a_t));
    const int es = dtype == NMM_BF16 ? 2 : 4, V = 16 / es;
    const bool vec = n % V == 0 && aligned(x, 16) && aligned(eu, 16) && (ec == nullptr || aligned(ec, 16));
    const int64_t work = vec ? n / V : n;
    int64_t blocks = ceil_div(work, 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
#define CFG_CASE(T, VV) launch_pdl(cfg_ddim_kernel<T, VV>, (unsigned)blocks, 256, 0, st, (T *)x, (const T *)eu, (const T *)ec, n, g, cx, ce)
    if (dtype == NMM_BF16) { if (vec) CFG_CASE(bf16, true); else CFG_CASE(bf16, false); }
    else { if (vec) CFG_CASE(float, true); else CFG_CASE(float, false); }
#undef CFG_CASE
    NMM_LAUNCHED("cfg_ddim_kernel");
    return NMM_OK;
}

}  // namespace nmm
