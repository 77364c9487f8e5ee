// Short-sequence temporal self-attention: softmax(q k^T * d_h^-1/2) v over the frame axis (F <= 32),
// independently per (batch, spatial position, head).
//
// Reference arithmetic: CrossAttention._attention, motion_module_new.py:258-287 (baddbmm(beta=0, alpha=scale)
// -> softmax(dim=-1) -> bmm), with reshape_heads_to_batch_dim / reshape_batch_dim_to_heads (:181-193) and the
// "(b f) d c -> (b d) f c" / inverse rearranges of VersatileAttention.forward (motion_module.py:275,327) folded
// into the addressing: a token row is n = (b*F + f)*P + p, so the F rows of one position are P rows apart.
//
// Mapping: one thread per (position, head, query frame).  A CTA stages the q|k|v rows of PB consecutive
// positions x all F frames in shared memory with 16-byte coalesced loads (each frame contributes one
// contiguous PB*3C-element run), keeps the F scores in registers, does the softmax in fp32 and writes its
// context row back through shared memory so the global stores are 16-byte coalesced as well.
// HBM-bound: bytes = 4*N*C*s (read qkv, write ctx), flops = 4*N*F*C.
#include "common.cuh"

namespace nmm {

template <typename T, int VEC> struct RowVec;
template <> struct RowVec<bf16, 8> {
    static __device__ __forceinline__ void load(const bf16 *p, float (&f)[8]) {
        uint4 v = *reinterpret_cast<const uint4 *>(p);
        f[0] = bf16_lo(v.x); f[1] = bf16_hi(v.x); f[2] = bf16_lo(v.y); f[3] = bf16_hi(v.y);
        f[4] = bf16_lo(v.z); f[5] = bf16_hi(v.z); f[6] = bf16_lo(v.w); f[7] = bf16_hi(v.w);
    }
    static __device__ __forceinline__ void store(bf16 *p, const float (&f)[8]) {
        *reinterpret_cast<uint4 *>(p) = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]),
                                                   pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
    }
};
template <> struct RowVec<float, 4> {
    static __device__ __forceinline__ void load(const float *p, float (&f)[4]) {
        float4 v = *reinterpret_cast<const float4 *>(p);
        f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
    }
    static __device__ __forceinline__ void store(float *p, const float (&f)[4]) {
        *reinterpret_cast<float4 *>(p) = make_float4(f[0], f[1], f[2], f[3]);
    }
};
template <typename T> struct RowVec<T, 1> {
    static __device__ __forceinline__ void load(const T *p, float (&f)[1]) { f[0] = to_f32(*p); }
    static __device__ __forceinline__ void store(T *p, const float (&f)[1]) { *p = from_f32<T>(f[0]); }
};

// A CTA handles PB positions x HB heads (blockIdx.y selects the head group; HB == heads for the UNet's 320/640-channel
// levels, fewer for 1280 channels so the tile fits shared memory).
// smem layout: [PB positions][F frames][q(HB*dh) | k(HB*dh) | v(HB*dh) | PAD] elements of T; PAD keeps consecutive frame
// rows on different banks (row pitch in bytes = 3*HB*dh*s + 16).
template <typename T, int VEC, int FMAX>
__global__ void __launch_bounds__(256) temporal_attention_kernel(const T *__restrict__ qkv, T *__restrict__ ctx, int B, int F,
                                                                 int P, int C, int heads, int PB, int HB, float scale) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *sm = reinterpret_cast<T *>(smem_raw);
    constexpr int PAD = 16 / (int)sizeof(T);
    const int dh = C / heads;
    const int W = HB * dh;                      // channels of this head group
    const int c_off = blockIdx.y * W;           // first channel of the head group
    const int pitch = 3 * W + PAD;
    const int tiles_per_img = (P + PB - 1) / PB;
    const int b = blockIdx.x / tiles_per_img;
    const int p0 = (blockIdx.x % tiles_per_img) * PB;
    const int npos = min(PB, P - p0);
    const int tid = threadIdx.x, nthr = blockDim.x;

    // ---- stage q|k|v with cp.async (16-byte LDGSTS, no register staging): one warp per (frame, position) row at a time,
    //      lanes striding over the row's vectors; the three segments (q, k, v of this head group) are W elements each --------
    {
        constexpr int LV = 16 / (int)sizeof(T);                     // elements per 16-byte copy
        const int warp = tid >> 5, lane = tid & 31, nwarps = nthr >> 5;
        const int rows = F * npos;
        if (VEC > 1) {                                              // host guarantees 16-byte alignment when VEC > 1
            const int vec_per_seg = W / LV;
            for (int r = warp; r < rows; r += nwarps) {
                const int f = r / npos, pl = r - f * npos;
                const T *src_row = qkv + ((int64_t)(b * F + f) * P + p0 + pl) * (3 * C) + c_off;
                const uint32_t dst_row = (uint32_t)__cvta_generic_to_shared(sm + (pl * F + f) * pitch);
#pragma unroll
                for (int seg = 0; seg < 3; seg++)
                    for (int v = lane; v < vec_per_seg; v += 32)
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_row + (uint32_t)((seg * W + v * LV) * sizeof(T))),
                                     "l"(src_row + seg * C + v * LV)
                                     : "memory");
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        } else {
            for (int r = warp; r < rows; r += nwarps) {
                const int f = r / npos, pl = r - f * npos;
                const T *src_row = qkv + ((int64_t)(b * F + f) * P + p0 + pl) * (3 * C) + c_off;
                T *dst_row = sm + (pl * F + f) * pitch;
                for (int seg = 0; seg < 3; seg++)
                    for (int e = lane; e < W; e += 32) dst_row[seg * W + e] = src_row[seg * C + e];
            }
        }
    }
    __syncthreads();

    // ---- one thread per (pl, head, f) ---------------------------------------------------------------
    const int per_pos = HB * F;
    const int pl = tid / per_pos;
    const int hf = tid % per_pos;
    const int head = hf / F, f = hf % F;
    const bool active = pl < npos;             // tid < PB*heads*F by launch configuration
    if (active) {
        T *qrow = sm + (pl * F + f) * pitch + head * dh;             // this thread's query row (reused for output)
        const T *kbase = sm + (pl * F) * pitch + W + head * dh;
        const T *vbase = sm + (pl * F) * pitch + 2 * W + head * dh;
        float sc[FMAX];
#pragma unroll
        for (int j = 0; j < FMAX; j++) sc[j] = 0.f;
        for (int d = 0; d < dh; d += VEC) {
            float q[VEC];
            RowVec<T, VEC>::load(qrow + d, q);
#pragma unroll
            for (int j = 0; j < FMAX; j++) {
                if (j < F) {
                    float k[VEC];
                    RowVec<T, VEC>::load(kbase + j * pitch + d, k);
#pragma unroll
                    for (int e = 0; e < VEC; e++) sc[j] = fmaf(q[e], k[e], sc[j]);
                }
            }
        }
        // softmax over the F keys (fp32; the reference computes it in the input dtype -- fp32 as shipped)
        float mx = -INFINITY;
#pragma unroll
        for (int j = 0; j < FMAX; j++) if (j < F) { sc[j] *= scale; mx = fmaxf(mx, sc[j]); }
        float den = 0.f;
#pragma unroll
        for (int j = 0; j < FMAX; j++) if (j < F) { sc[j] = __expf(sc[j] - mx); den += sc[j]; }
        const float inv = 1.0f / den;
#pragma unroll
        for (int j = 0; j < FMAX; j++) if (j < F) sc[j] *= inv;
        for (int d = 0; d < dh; d += VEC) {
            float o[VEC];
#pragma unroll
            for (int e = 0; e < VEC; e++) o[e] = 0.f;
#pragma unroll
            for (int j = 0; j < FMAX; j++) {
                if (j < F) {
                    float v[VEC];
                    RowVec<T, VEC>::load(vbase + j * pitch + d, v);
#pragma unroll
                    for (int e = 0; e < VEC; e++) o[e] = fmaf(sc[j], v[e], o[e]);
                }
            }
            RowVec<T, VEC>::store(qrow + d, o);                       // q slot is private to this thread
        }
    }
    __syncthreads();

    // ---- write ctx rows (the first W elements of each staged row): one warp per row, 16-byte coalesced stores ----------
    {
        constexpr int LV = 16 / (int)sizeof(T);
        const int warp = tid >> 5, lane = tid & 31, nwarps = nthr >> 5;
        const int rows = F * npos;
        for (int r = warp; r < rows; r += nwarps) {
            const int f2 = r / npos, pl2 = r - f2 * npos;
            const T *src_row = sm + (pl2 * F + f2) * pitch;
            T *dst_row = ctx + ((int64_t)(b * F + f2) * P + p0 + pl2) * C + c_off;
            if (VEC > 1) {
                for (int v = lane; v < W / LV; v += 32)
                    *reinterpret_cast<uint4 *>(dst_row + v * LV) = *reinterpret_cast<const uint4 *>(src_row + v * LV);
            } else {
                for (int e = lane; e < W; e += 32) dst_row[e] = src_row[e];
            }
        }
    }
}

template <typename T, int VEC>
static int launch_attn_t(const Geo &g, const T *qkv, T *ctx, cudaStream_t st) {
    if (g.heads * g.F > 256 && g.F > 256) return fail(NMM_ERR_UNSUPPORTED, "frames too large");
    const size_t smem_cap = 200 * 1024, smem_pref = 32 * 1024;       // <= 32 KB: ~7 CTAs per SM, so load / compute / store phases of different CTAs overlap
    auto row_bytes = [&](int hb) { return (size_t)(3 * hb * g.dh) * sizeof(T) + 16; };
    // heads per CTA: the largest divisor of `heads` whose tile fits the preferred budget, else the largest that fits at all
    int HB = 0, PB = 1;
    for (int pass = 0; pass < 2 && HB == 0; pass++) {
        const size_t budget = pass == 0 ? smem_pref : smem_cap;
        for (int hb = g.heads; hb >= 1; hb--) {
            if (g.heads % hb || hb * g.F > 256) continue;
            int pb = 256 / (hb * g.F);
            if (pb > g.P) pb = g.P;
            while (pb > 1 && (size_t)pb * g.F * row_bytes(hb) > budget) pb--;
            if ((size_t)pb * g.F * row_bytes(hb) <= budget) { HB = hb; PB = pb; break; }
        }
    }
    if (HB == 0) return fail(NMM_ERR_UNSUPPORTED, "attention tile does not fit shared memory (C=%d, F=%d)", g.C, g.F);
    const int per_pos = HB * g.F;
    const size_t smem = (size_t)PB * g.F * row_bytes(HB);
    const int threads = ((PB * per_pos + 31) / 32) * 32;
    const int64_t blocks = (int64_t)g.B * ceil_div(g.P, PB);
    if (blocks > 0x7fffffff) return fail(NMM_ERR_UNSUPPORTED, "attention grid too large");
    const dim3 grid((unsigned)blocks, (unsigned)(g.heads / HB));
    const float scale = 1.0f / sqrtf((float)g.dh);
#define ATTN_CASE(FM)                                                                                                    \
    do {                                                                                                                 \
        auto kern = temporal_attention_kernel<T, VEC, FM>;                                                               \
        static bool attr_set = false; /* once per instantiation; never inside a stream capture after warm-up */         \
        if (!attr_set) {                                                                                                 \
            NMM_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cap));          \
            attr_set = true;                                                                                             \
        }                                                                                                                \
        ProfScope prof(K_ATTENTION, st, 4.0 * g.N * g.F * g.C, 4.0 * g.N * g.C * sizeof(T));                             \
        kern<<<grid, threads, smem, st>>>(qkv, ctx, g.B, g.F, g.P, g.C, g.heads, PB, HB, scale);                          \
    } while (0)
    if (g.F <= 8) ATTN_CASE(8);
    else if (g.F <= 16) ATTN_CASE(16);
    else ATTN_CASE(32);
#undef ATTN_CASE
    NMM_LAUNCHED("temporal_attention_kernel");
    return NMM_OK;
}

int launch_temporal_attention(const Geo &g, const void *qkv, void *ctx, cudaStream_t st) {
    if (g.F > NMM_MAX_FRAMES) return fail(NMM_ERR_UNSUPPORTED, "frames %d > %d", g.F, NMM_MAX_FRAMES);
    const bool al = aligned(qkv, 16) && aligned(ctx, 16);
    if (g.dtype == NMM_BF16) {
        if (g.dh % 8 == 0 && al) return launch_attn_t<bf16, 8>(g, (const bf16 *)qkv, (bf16 *)ctx, st);
        return launch_attn_t<bf16, 1>(g, (const bf16 *)qkv, (bf16 *)ctx, st);
    }
    if (g.dh % 4 == 0 && al) return launch_attn_t<float, 4>(g, (const float *)qkv, (float *)ctx, st);
    return launch_attn_t<float, 1>(g, (const float *)qkv, (float *)ctx, st);
}

}  // namespace nmm
