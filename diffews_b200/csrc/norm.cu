// K4 / K7 — GroupNorm(+SiLU), LayerNorm and the row softmax of the VAE attention: bandwidth-bound kernels,
// 128-bit vectorised loads/stores, fp32 statistics, warp-shuffle / fixed-order reductions (deterministic).
// ref: nn.GroupNorm + SiLU inside diffusers ResnetBlock2D / Transformer2DModel.norm / conv_norm_out
//      (diffews/models/unet_2d_condition.py:1246-1248), BasicTransformerBlock.norm1/2/3 (upstream).
#include <atomic>
#include <cstdlib>

#include "common.cuh"
#include "ptx.cuh"

namespace dfw {
extern std::atomic<long long> g_launches;
namespace {

// silu(x) = x * sigmoid(x) = 0.5 x (1 + tanh(x/2)): one MUFU op (tanh.approx.f32, max rel. error 2^-11 — the same
// order as the fp16 / bf16 rounding of the stored result) instead of ex2 + rcp.
__device__ __forceinline__ float silu_f(float x) {
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * x));
    const float hx = 0.5f * x;
    return fmaf(hx, t, hx);
}

// load 8 consecutive channels as fp32.  XD: 0 = bf16, 1 = fp32, 2 = fp16 storage
template <int XD>
__device__ __forceinline__ void load8(const void* base, long long elem_off, float (&v)[8]) {
    if constexpr (XD == 1) {
        const float4* p = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + elem_off);
        float4 a = __ldg(p), b = __ldg(p + 1);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
        uint4 u = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(base) + elem_off));
        const float2 a = unpack_h2(u.x, XD == 2), b = unpack_h2(u.y, XD == 2), c = unpack_h2(u.z, XD == 2),
                     d = unpack_h2(u.w, XD == 2);
        v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
    }
}
// store 8 consecutive channels as bf16 (f16 == 0) or fp16 (f16 != 0); both are 2-byte tensor-core operand formats
__device__ __forceinline__ void store8_16(void* base, long long elem_off, const float (&v)[8], int f16) {
    uint4 o;
    if (f16) {
        o.x = pack_f16x2(v[0], v[1]); o.y = pack_f16x2(v[2], v[3]);
        o.z = pack_f16x2(v[4], v[5]); o.w = pack_f16x2(v[6], v[7]);
    } else {
        o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]);
        o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
    }
    *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(base) + elem_off) = o;
}

// ---------------------------------------------------------------------------------------------------------
// GroupNorm stage 1: per (image, row-chunk) partial sums per group.  blockDim = V * RPI (V = C/8).
// partial layout: [N, nchunks, groups, 2]
// ---------------------------------------------------------------------------------------------------------
template <int XD>
__global__ void gn_stats_kernel(const void* __restrict__ x, float* __restrict__ partial, int HW, int C, int groups,
                                int rows_per_chunk, int RPI) {
    extern __shared__ float sm[];  // [RPI][C][2]
    pdl_wait();
    const int V = C / 8;
    const int n = blockIdx.y, chunk = blockIdx.x, nchunks = gridDim.x;
    const int vc = threadIdx.x % V, r = threadIdx.x / V;
    const int row0 = chunk * rows_per_chunk;
    const int row1 = min(HW, row0 + rows_per_chunk);
    float s[8], ss[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { s[j] = 0.f; ss[j] = 0.f; }
    const long long img_off = static_cast<long long>(n) * HW * C + vc * 8;
    int row = row0 + r;
    for (; row + 3 * RPI < row1; row += 4 * RPI) {          // 4 independent 16-byte loads in flight per thread
        float v0[8], v1[8], v2[8], v3[8];
        load8<XD>(x, img_off + static_cast<long long>(row) * C, v0);
        load8<XD>(x, img_off + static_cast<long long>(row + RPI) * C, v1);
        load8<XD>(x, img_off + static_cast<long long>(row + 2 * RPI) * C, v2);
        load8<XD>(x, img_off + static_cast<long long>(row + 3 * RPI) * C, v3);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            s[j] += (v0[j] + v1[j]) + (v2[j] + v3[j]);
            ss[j] = fmaf(v0[j], v0[j], fmaf(v1[j], v1[j], fmaf(v2[j], v2[j], fmaf(v3[j], v3[j], ss[j]))));
        }
    }
    for (; row < row1; row += RPI) {
        float v[8];
        load8<XD>(x, img_off + static_cast<long long>(row) * C, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) { s[j] += v[j]; ss[j] = fmaf(v[j], v[j], ss[j]); }
    }
    float* my = sm + (static_cast<size_t>(r) * C + vc * 8) * 2;
#pragma unroll
    for (int j = 0; j < 8; ++j) { my[2 * j] = s[j]; my[2 * j + 1] = ss[j]; }
    __syncthreads();
    const int cpg = C / groups;
    for (int g = threadIdx.x; g < groups; g += blockDim.x) {
        float gs = 0.f, gss = 0.f;
        for (int rr = 0; rr < RPI; ++rr) {
            const float* base = sm + (static_cast<size_t>(rr) * C + g * cpg) * 2;
            for (int c = 0; c < cpg; ++c) { gs += base[2 * c]; gss += base[2 * c + 1]; }
        }
        float* out = partial + ((static_cast<size_t>(n) * nchunks + chunk) * groups + g) * 2;
        out[0] = gs; out[1] = gss;
    }
}

// stage 2: y = [silu](x * scale + shift).  Same (row-chunk, image) decomposition as stage 1: a thread owns 8 fixed
// channels and walks rows, so there is no per-vector index arithmetic.  Every CTA first folds the stage-1 partials of
// its image (fixed order -> deterministic; a few KB from L2) into mean / rstd per group — no separate finalize launch.
template <int XD>
__global__ void gn_apply_kernel(const void* __restrict__ x, const float* __restrict__ partial,
                                const float* __restrict__ gamma, const float* __restrict__ beta,
                                void* __restrict__ y, int HW, int C, int groups, int nchunks, float eps,
                                int rows_per_chunk, int RPI, int apply_silu, int y_f16) {
    pdl_wait();
    __shared__ float s_red[8][64][2];
    __shared__ float s_mean[64], s_rstd[64];
    const int V = C / 8;
    const int n = blockIdx.y, chunk = blockIdx.x;
    const int cpg = C / groups;
    {
        // 8 slices x groups threads (first 8*groups threads of the CTA; blockDim >= 128 whenever C >= 128*... else loop)
        const int nthr = blockDim.x;
        for (int idx = threadIdx.x; idx < 8 * groups; idx += nthr) {
            const int g = idx % groups, sl = idx / groups;
            float s = 0.f, ss = 0.f;
            // independent 8-byte loads, 8 in flight per thread: the fold is a chain of L2 latencies otherwise
            // (nchunks is 2-8 x #SMs when the statistics come from a conv epilogue)
            const float2* pp = reinterpret_cast<const float2*>(partial) + static_cast<size_t>(n) * nchunks * groups + g;
#pragma unroll 8
            for (int c = sl; c < nchunks; c += 8) {
                const float2 v = __ldg(pp + static_cast<size_t>(c) * groups);
                s += v.x; ss += v.y;
            }
            s_red[sl][g][0] = s; s_red[sl][g][1] = ss;
        }
        __syncthreads();
        for (int g = threadIdx.x; g < groups; g += nthr) {
            double s = 0.0, ss = 0.0;
#pragma unroll
            for (int sl = 0; sl < 8; ++sl) { s += s_red[sl][g][0]; ss += s_red[sl][g][1]; }
            const double cnt = static_cast<double>(HW) * cpg;
            const double mean = s / cnt;
            double var = ss / cnt - mean * mean;
            if (var < 0.0) var = 0.0;
            s_mean[g] = static_cast<float>(mean);
            s_rstd[g] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
        }
        __syncthreads();
    }
    const int vc = threadIdx.x % V, r = threadIdx.x / V;
    const int row0 = chunk * rows_per_chunk;
    const int row1 = min(HW, row0 + rows_per_chunk);
    float sc[8], sh[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int c = vc * 8 + j;
        const int g = c / cpg;
        sc[j] = s_rstd[g] * __ldg(gamma + c);
        sh[j] = __ldg(beta + c) - s_mean[g] * sc[j];
    }
    const long long img_off = static_cast<long long>(n) * HW * C + vc * 8;
    auto finish = [&](float (&v)[8], long long off) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float a = fmaf(v[j], sc[j], sh[j]);
            if (apply_silu) a = silu_f(a);
            v[j] = a;
        }
        store8_16(y, off, v, y_f16);
    };
    int row = row0 + r;
    for (; row + 3 * RPI < row1; row += 4 * RPI) {
        float v0[8], v1[8], v2[8], v3[8];
        const long long o0 = img_off + static_cast<long long>(row) * C, o1 = o0 + static_cast<long long>(RPI) * C,
                        o2 = o1 + static_cast<long long>(RPI) * C, o3 = o2 + static_cast<long long>(RPI) * C;
        load8<XD>(x, o0, v0); load8<XD>(x, o1, v1); load8<XD>(x, o2, v2); load8<XD>(x, o3, v3);
        finish(v0, o0); finish(v1, o1); finish(v2, o2); finish(v3, o3);
    }
    for (; row < row1; row += RPI) {
        float v[8];
        const long long o = img_off + static_cast<long long>(row) * C;
        load8<XD>(x, o, v);
        finish(v, o);
    }
}

// ---------------------------------------------------------------------------------------------------------
// GroupNorm (+SiLU) backward (BASELINE config 4: training-shape forward + backward of the GroupNorm kernels;
// ref: the nn.GroupNorm + SiLU autograd of every diffusers ResnetBlock2D under train_tools/...v3.py:1320-1396).
//   z = xh * gamma + beta, xh = (x - mean) * rstd, y = silu(z) or z;   dz = dy * act'(z)
//   dgamma_c = sum dz xh, dbeta_c = sum dz;  dx = rstd (dz gamma - S1/M - xh S2/M),
//   S1 = sum_{group} dz gamma, S2 = sum_{group} dz gamma xh, M = HW * cpg.
// Same (row-chunk, image) decomposition and fixed-order (deterministic) folds as the forward.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float act_grad(float z, int silu) {
    if (!silu) return 1.0f;
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * z));
    const float sg = fmaf(0.5f, t, 0.5f);                     // sigmoid(z)
    return sg * fmaf(z, 1.0f - sg, 1.0f);
}
template <int XD>
__device__ __forceinline__ void store8(void* base, long long off, const float (&v)[8]) {
    if constexpr (XD == 1) {
        float4* p = reinterpret_cast<float4*>(reinterpret_cast<float*>(base) + off);
        p[0] = make_float4(v[0], v[1], v[2], v[3]);
        p[1] = make_float4(v[4], v[5], v[6], v[7]);
    } else {
        store8_16(base, off, v, XD == 2);
    }
}
// fold the forward statistics partials of image n into s_mean / s_rstd (identical to gn_apply_kernel's prologue)
__device__ __forceinline__ void gn_fold_stats(const float* __restrict__ partial, int n, int nchunks, int groups, int HW,
                                              int cpg, float eps, float (*s_red)[64][2], float* s_mean, float* s_rstd) {
    for (int idx = threadIdx.x; idx < 8 * groups; idx += blockDim.x) {
        const int g = idx % groups, sl = idx / groups;
        float s = 0.f, ss = 0.f;
        const float2* pp = reinterpret_cast<const float2*>(partial) + static_cast<size_t>(n) * nchunks * groups + g;
#pragma unroll 8
        for (int c = sl; c < nchunks; c += 8) {
            const float2 v = __ldg(pp + static_cast<size_t>(c) * groups);
            s += v.x; ss += v.y;
        }
        s_red[sl][g][0] = s; s_red[sl][g][1] = ss;
    }
    __syncthreads();
    for (int g = threadIdx.x; g < groups; g += blockDim.x) {
        double s = 0.0, ss = 0.0;
#pragma unroll
        for (int sl = 0; sl < 8; ++sl) { s += s_red[sl][g][0]; ss += s_red[sl][g][1]; }
        const double cnt = static_cast<double>(HW) * cpg;
        const double mean = s / cnt;
        double var = ss / cnt - mean * mean;
        if (var < 0.0) var = 0.0;
        s_mean[g] = static_cast<float>(mean);
        s_rstd[g] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
    }
    __syncthreads();
}
// pass 1: per (image, row chunk, channel) a = sum dz, b = sum dz * xh   ->  ab_partial [N][nchunks][C][2]
template <int XD>
__global__ void gn_bwd_reduce_kernel(const void* __restrict__ x, const void* __restrict__ dy,
                                     const float* __restrict__ partial, const float* __restrict__ gamma,
                                     const float* __restrict__ beta, float* __restrict__ ab_partial, int HW, int C,
                                     int groups, float eps, int rows_per_chunk, int RPI, int silu) {
    extern __shared__ float sm[];                                 // [RPI][C][2]
    __shared__ float s_red[8][64][2];
    __shared__ float s_mean[64], s_rstd[64];
    const int V = C / 8, cpg = C / groups;
    const int n = blockIdx.y, chunk = blockIdx.x, nchunks = gridDim.x;
    gn_fold_stats(partial, n, nchunks, groups, HW, cpg, eps, s_red, s_mean, s_rstd);
    const int vc = threadIdx.x % V, r = threadIdx.x / V;
    const int row0 = chunk * rows_per_chunk, row1 = min(HW, row0 + rows_per_chunk);
    float mu[8], rs[8], ga[8], be[8], a[8], b[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int c = vc * 8 + j, g = c / cpg;
        mu[j] = s_mean[g]; rs[j] = s_rstd[g]; ga[j] = __ldg(gamma + c); be[j] = __ldg(beta + c);
        a[j] = 0.f; b[j] = 0.f;
    }
    const long long img_off = static_cast<long long>(n) * HW * C + vc * 8;
    for (int row = row0 + r; row < row1; row += RPI) {
        float xv[8], gv[8];
        const long long o = img_off + static_cast<long long>(row) * C;
        load8<XD>(x, o, xv); load8<XD>(dy, o, gv);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float xh = (xv[j] - mu[j]) * rs[j];
            const float dz = gv[j] * act_grad(fmaf(xh, ga[j], be[j]), silu);
            a[j] += dz; b[j] = fmaf(dz, xh, b[j]);
        }
    }
    float* my = sm + (static_cast<size_t>(r) * C + vc * 8) * 2;
#pragma unroll
    for (int j = 0; j < 8; ++j) { my[2 * j] = a[j]; my[2 * j + 1] = b[j]; }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float sa = 0.f, sb = 0.f;
        for (int rr = 0; rr < RPI; ++rr) { sa += sm[(static_cast<size_t>(rr) * C + c) * 2]; sb += sm[(static_cast<size_t>(rr) * C + c) * 2 + 1]; }
        float* out = ab_partial + ((static_cast<size_t>(n) * nchunks + chunk) * C + c) * 2;
        out[0] = sa; out[1] = sb;
    }
}
// pass 2 (one CTA per image): ab[n][c] = sum over chunks; gs[n][g] = (S1, S2) / M; ms[n][g] = (mean, rstd)
__global__ void gn_bwd_fold_kernel(const float* __restrict__ partial, const float* __restrict__ ab_partial,
                                   const float* __restrict__ gamma, float* __restrict__ ab, float* __restrict__ gs,
                                   float* __restrict__ ms, int nchunks, int HW, int C, int groups, float eps) {
    __shared__ float s_red[8][64][2];
    __shared__ float s_mean[64], s_rstd[64];
    extern __shared__ float s_ab[];                               // [C][2]
    const int n = blockIdx.x, cpg = C / groups;
    gn_fold_stats(partial, n, nchunks, groups, HW, cpg, eps, s_red, s_mean, s_rstd);
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float sa = 0.f, sb = 0.f;
        for (int k = 0; k < nchunks; ++k) {
            const float2 v = __ldg(reinterpret_cast<const float2*>(ab_partial) + (static_cast<size_t>(n) * nchunks + k) * C + c);
            sa += v.x; sb += v.y;
        }
        s_ab[2 * c] = sa; s_ab[2 * c + 1] = sb;
        ab[(static_cast<size_t>(n) * C + c) * 2] = sa; ab[(static_cast<size_t>(n) * C + c) * 2 + 1] = sb;
    }
    __syncthreads();
    for (int g = threadIdx.x; g < groups; g += blockDim.x) {
        float s1 = 0.f, s2 = 0.f;
        for (int c = g * cpg; c < (g + 1) * cpg; ++c) {
            const float ga = __ldg(gamma + c);
            s1 = fmaf(ga, s_ab[2 * c], s1); s2 = fmaf(ga, s_ab[2 * c + 1], s2);
        }
        const float inv = 1.0f / (static_cast<float>(HW) * cpg);
        gs[(static_cast<size_t>(n) * groups + g) * 2] = s1 * inv; gs[(static_cast<size_t>(n) * groups + g) * 2 + 1] = s2 * inv;
        ms[(static_cast<size_t>(n) * groups + g) * 2] = s_mean[g]; ms[(static_cast<size_t>(n) * groups + g) * 2 + 1] = s_rstd[g];
    }
}
// pass 3: dgamma_c = sum_n b, dbeta_c = sum_n a (fixed order over images)
__global__ void gn_bwd_param_kernel(const float* __restrict__ ab, float* __restrict__ dgamma, float* __restrict__ dbeta,
                                    int N, int C) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float sa = 0.f, sb = 0.f;
    for (int n = 0; n < N; ++n) { sa += ab[(static_cast<size_t>(n) * C + c) * 2]; sb += ab[(static_cast<size_t>(n) * C + c) * 2 + 1]; }
    dbeta[c] = sa; dgamma[c] = sb;
}
// pass 4: dx
template <int XD>
__global__ void gn_bwd_apply_kernel(const void* __restrict__ x, const void* __restrict__ dy, const float* __restrict__ gs,
                                    const float* __restrict__ ms, const float* __restrict__ gamma,
                                    const float* __restrict__ beta, void* __restrict__ dx, int HW, int C, int groups,
                                    int rows_per_chunk, int RPI, int silu) {
    const int V = C / 8, cpg = C / groups;
    const int n = blockIdx.y, chunk = blockIdx.x;
    const int vc = threadIdx.x % V, r = threadIdx.x / V;
    const int row0 = chunk * rows_per_chunk, row1 = min(HW, row0 + rows_per_chunk);
    float mu[8], rs[8], ga[8], be[8], m1[8], m2[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int c = vc * 8 + j, g = c / cpg;
        const size_t o = (static_cast<size_t>(n) * groups + g) * 2;
        mu[j] = __ldg(ms + o); rs[j] = __ldg(ms + o + 1); m1[j] = __ldg(gs + o); m2[j] = __ldg(gs + o + 1);
        ga[j] = __ldg(gamma + c); be[j] = __ldg(beta + c);
    }
    const long long img_off = static_cast<long long>(n) * HW * C + vc * 8;
    for (int row = row0 + r; row < row1; row += RPI) {
        float xv[8], gv[8], o8[8];
        const long long o = img_off + static_cast<long long>(row) * C;
        load8<XD>(x, o, xv); load8<XD>(dy, o, gv);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float xh = (xv[j] - mu[j]) * rs[j];
            const float dz = gv[j] * act_grad(fmaf(xh, ga[j], be[j]), silu);
            o8[j] = rs[j] * (fmaf(dz, ga[j], -m1[j]) - xh * m2[j]);
        }
        store8<XD>(dx, o, o8);
    }
}

// Per-(image, channel) scale / shift of a GroupNorm whose statistics came from a conv epilogue: out[n][0][c] = rstd *
// gamma, out[n][1][c] = beta - mean * rstd * gamma.  Same fixed-order fold (8 slices, then fp64) as gn_apply_kernel, so
// a convolution that applies them on the fly (dfw_conv2d_igemm_gnin) sees exactly the numbers gn_apply would use.
__global__ void gn_scale_shift_kernel(const float* __restrict__ partial, int nchunks, const float* __restrict__ gamma,
                                      const float* __restrict__ beta, float* __restrict__ out, double HW, int C,
                                      int groups, float eps) {
    __shared__ float s_red[8][64][2];
    __shared__ float s_mean[64], s_rstd[64];
    const int n = blockIdx.x, cpg = C / groups;
    for (int idx = threadIdx.x; idx < 8 * groups; idx += blockDim.x) {
        const int g = idx % groups, sl = idx / groups;
        float s = 0.f, ss = 0.f;
        const float2* pp = reinterpret_cast<const float2*>(partial) + static_cast<size_t>(n) * nchunks * groups + g;
#pragma unroll 8
        for (int c = sl; c < nchunks; c += 8) {
            const float2 v = __ldg(pp + static_cast<size_t>(c) * groups);
            s += v.x; ss += v.y;
        }
        s_red[sl][g][0] = s; s_red[sl][g][1] = ss;
    }
    __syncthreads();
    for (int g = threadIdx.x; g < groups; g += blockDim.x) {
        double s = 0.0, ss = 0.0;
#pragma unroll
        for (int sl = 0; sl < 8; ++sl) { s += s_red[sl][g][0]; ss += s_red[sl][g][1]; }
        const double cnt = HW * cpg;
        const double mean = s / cnt;
        double var = ss / cnt - mean * mean;
        if (var < 0.0) var = 0.0;
        s_mean[g] = static_cast<float>(mean);
        s_rstd[g] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const int g = c / cpg;
        const float sc = s_rstd[g] * __ldg(gamma + c);
        out[(static_cast<size_t>(n) * 2) * C + c] = sc;
        out[(static_cast<size_t>(n) * 2 + 1) * C + c] = __ldg(beta + c) - s_mean[g] * sc;
    }
}

// ---------------------------------------------------------------------------------------------------------
// LayerNorm: one warp per row, row held in registers (C <= 2048), exact two-pass statistics.
// ---------------------------------------------------------------------------------------------------------
template <int XD>
__global__ void layernorm_kernel(const void* __restrict__ x, const float* __restrict__ gamma,
                                 const float* __restrict__ beta, void* __restrict__ y, int M, int C, float eps,
                                 int y_f16) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= M) return;
    const int V = C / 8;
    const long long row_off = static_cast<long long>(warp) * C;
    float v[8][8];
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int vc = lane + 32 * k;
        if (vc < V) {
            load8<XD>(x, row_off + vc * 8, v[k]);
#pragma unroll
            for (int j = 0; j < 8; ++j) s += v[k][j];
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / C;
    float ss = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int vc = lane + 32 * k;
        if (vc < V) {
#pragma unroll
            for (int j = 0; j < 8; ++j) { const float d = v[k][j] - mean; ss = fmaf(d, d, ss); }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const float rstd = rsqrtf(ss / C + eps);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int vc = lane + 32 * k;
        if (vc < V) {
            const float4* gp = reinterpret_cast<const float4*>(gamma + vc * 8);
            const float4* bp = reinterpret_cast<const float4*>(beta + vc * 8);
            float4 g0 = __ldg(gp), g1 = __ldg(gp + 1), b0 = __ldg(bp), b1 = __ldg(bp + 1);
            float o[8];
            o[0] = (v[k][0] - mean) * rstd * g0.x + b0.x; o[1] = (v[k][1] - mean) * rstd * g0.y + b0.y;
            o[2] = (v[k][2] - mean) * rstd * g0.z + b0.z; o[3] = (v[k][3] - mean) * rstd * g0.w + b0.w;
            o[4] = (v[k][4] - mean) * rstd * g1.x + b1.x; o[5] = (v[k][5] - mean) * rstd * g1.y + b1.y;
            o[6] = (v[k][6] - mean) * rstd * g1.z + b1.z; o[7] = (v[k][7] - mean) * rstd * g1.w + b1.w;
            store8_16(y, row_off + vc * 8, o, y_f16);
        }
    }
}

// Same arithmetic (exact two-pass statistics on the register-resident row), laid out for bandwidth: 4-element vectors
// (a warp instruction covers 512 contiguous bytes of an fp32 row) and NV = ceil(C / 128) compile-time steps, so the
// 320 / 640 / 1280-channel rows of the UNet keep all lanes busy and every load of a row is in flight at once.
template <int XD>
__device__ __forceinline__ float4 ln_load4(const void* x, long long off) {
    if constexpr (XD == 1) {
        return __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(x) + off));
    } else {
        const uint2 r = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const uint16_t*>(x) + off));
        const float2 a = unpack_h2(r.x, XD == 2), b = unpack_h2(r.y, XD == 2);
        return make_float4(a.x, a.y, b.x, b.y);
    }
}
template <int XD> struct LnRaw { using T = uint2; };
template <> struct LnRaw<1> { using T = float4; };
template <int XD>
__device__ __forceinline__ typename LnRaw<XD>::T ln_load_raw(const void* x, long long off) {
    if constexpr (XD == 1) return __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(x) + off));
    else return __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const uint16_t*>(x) + off));
}
template <int XD>
__device__ __forceinline__ float4 ln_cvt(typename LnRaw<XD>::T r) {
    if constexpr (XD == 1) {
        return r;
    } else {
        const float2 a = unpack_h2(r.x, XD == 2), b = unpack_h2(r.y, XD == 2);
        return make_float4(a.x, a.y, b.x, b.y);
    }
}
template <int XD, int NV>
__global__ void __launch_bounds__(256) layernorm_v4_kernel(const void* __restrict__ x, const float* __restrict__ gamma,
                                                           const float* __restrict__ beta, void* __restrict__ y, int M,
                                                           int C, float eps, int y_f16) {
    // Grid-stride over rows with the NEXT row's loads issued before the current row is reduced, normalised and stored: a
    // warp that loads, reduces (10 shuffles twice) and stores one row and then exits leaves the memory pipe idle for a
    // third of its life (26.6 us for 84 MB at M = 65536, C = 320 = 3.2 TB/s; CTA turnover on top).
    pdl_wait();
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= M) return;
    const int V = C / 4;
    const float invC = 1.0f / C;
    typename LnRaw<XD>::T nx[NV];                          // next row, as loaded (16-bit rows stay packed: half the registers)
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        const int vc = lane + 32 * k;
        nx[k] = ln_load_raw<XD>(x, static_cast<long long>(row) * C + (vc < V ? vc : 0) * 4);
    }
    for (; row < M; row += nwarps) {
        const long long row_off = static_cast<long long>(row) * C;
        float4 v[NV];
#pragma unroll
        for (int k = 0; k < NV; ++k) v[k] = (lane + 32 * k < V) ? ln_cvt<XD>(nx[k]) : make_float4(0.f, 0.f, 0.f, 0.f);
        const int nrow = row + nwarps;
        if (nrow < M) {
#pragma unroll
            for (int k = 0; k < NV; ++k) {
                const int vc = lane + 32 * k;
                if (vc < V) nx[k] = ln_load_raw<XD>(x, static_cast<long long>(nrow) * C + vc * 4);
            }
        }
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < NV; ++k) s += (v[k].x + v[k].y) + (v[k].z + v[k].w);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        const float mean = s / C;
        float ss = 0.f;
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            if (lane + 32 * k < V) {
                const float a = v[k].x - mean, b = v[k].y - mean, c = v[k].z - mean, d = v[k].w - mean;
                ss = fmaf(a, a, ss); ss = fmaf(b, b, ss); ss = fmaf(c, c, ss); ss = fmaf(d, d, ss);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
        const float rstd = rsqrtf(ss / C + eps);
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            const int vc = lane + 32 * k;
            if (vc < V) {
                const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + vc);
                const float4 b = __ldg(reinterpret_cast<const float4*>(beta) + vc);
                uint2 o;
                o.x = pack_h2((v[k].x - mean) * rstd * g.x + b.x, (v[k].y - mean) * rstd * g.y + b.y, y_f16);
                o.y = pack_h2((v[k].z - mean) * rstd * g.z + b.z, (v[k].w - mean) * rstd * g.w + b.w, y_f16);
                *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(y) + row_off + vc * 4) = o;
            }
        }
    }
    (void)invC;
}
template <int XD>
static bool launch_layernorm_v4(const void* x, const float* gamma, const float* beta, void* y, int M, int C, float eps,
                                int y_f16, cudaStream_t stream) {
    const int nv = (C / 4 + 31) / 32;
    const int full = (M + 7) / 8;                       // one warp per row
    const int cap = sm_count() * (nv <= 5 ? 4 : 2);     // resident CTAs of 256 threads at 50-60 / 80-120 registers per thread
    const int blocks = full < cap ? full : cap;
    if (nv <= 3) launch_k(layernorm_v4_kernel<XD, 3>, blocks, 256, 0, stream, x, gamma, beta, y, M, C, eps, y_f16);
    else if (nv <= 5) launch_k(layernorm_v4_kernel<XD, 5>, blocks, 256, 0, stream, x, gamma, beta, y, M, C, eps, y_f16);
    else if (nv <= 10) launch_k(layernorm_v4_kernel<XD, 10>, blocks, 256, 0, stream, x, gamma, beta, y, M, C, eps, y_f16);
    else if (nv <= 16) launch_k(layernorm_v4_kernel<XD, 16>, blocks, 256, 0, stream, x, gamma, beta, y, M, C, eps, y_f16);
    else return false;
    return true;
}

// ---------------------------------------------------------------------------------------------------------
// LayerNorm backward (training shapes).  Statistics are recomputed from x (nothing is saved by the
// forward).  Warp per row, the row's x and dy held in registers:
//     xh = (x - mean) * rstd ;  a = mean_c(dy * gamma) ;  b = mean_c(dy * gamma * xh)
//     dx = rstd * (dy * gamma - a - xh * b)
// dgamma = sum_rows dy * xh and dbeta = sum_rows dy: every warp accumulates its rows into its own [2, C] slice of
// shared memory (owner-only read-modify-write, no atomics), the CTA folds its 8 slices in warp order into
// partial[cta][2][C], and a second kernel folds the CTAs in index order -- deterministic.
// ---------------------------------------------------------------------------------------------------------
constexpr int LNB_WARPS = 8;
template <int XD>
__device__ __forceinline__ void ln_store4(void* y, long long off, float4 v) {
    if constexpr (XD == 1) {
        *reinterpret_cast<float4*>(reinterpret_cast<float*>(y) + off) = v;
    } else {
        uint2 o;
        o.x = pack_h2(v.x, v.y, XD == 2);
        o.y = pack_h2(v.z, v.w, XD == 2);
        *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(y) + off) = o;
    }
}
template <int XD, int NV>
__global__ void __launch_bounds__(LNB_WARPS * 32) layernorm_bwd_kernel(const void* __restrict__ x, const void* __restrict__ dy,
                                                                       const float* __restrict__ gamma, void* __restrict__ dx,
                                                                       float* __restrict__ partial, int M, int C, float eps) {
    extern __shared__ float lnb_sm[];                       // [LNB_WARPS][2][C]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int V = C / 4;
    float* my = lnb_sm + static_cast<size_t>(warp) * 2 * C;
    for (int c = lane; c < 2 * C; c += 32) my[c] = 0.f;
    __syncwarp();
    float4 g[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        const int vc = lane + 32 * k;
        g[k] = vc < V ? __ldg(reinterpret_cast<const float4*>(gamma) + vc) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const float invC = 1.0f / C;
    for (int row = blockIdx.x * LNB_WARPS + warp; row < M; row += gridDim.x * LNB_WARPS) {
        const long long row_off = static_cast<long long>(row) * C;
        float4 v[NV], d[NV];
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            const int vc = lane + 32 * k;
            const bool ok = vc < V;
            v[k] = ok ? ln_load4<XD>(x, row_off + vc * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
            d[k] = ok ? ln_load4<XD>(dy, row_off + vc * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
            s += (v[k].x + v[k].y) + (v[k].z + v[k].w);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        const float mean = s * invC;
        float ss = 0.f;
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            if (lane + 32 * k < V) {
                v[k].x -= mean; v[k].y -= mean; v[k].z -= mean; v[k].w -= mean;
                ss = fmaf(v[k].x, v[k].x, ss); ss = fmaf(v[k].y, v[k].y, ss);
                ss = fmaf(v[k].z, v[k].z, ss); ss = fmaf(v[k].w, v[k].w, ss);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
        const float rstd = rsqrtf(ss * invC + eps);
        float a = 0.f, b = 0.f;
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            const int vc = lane + 32 * k;
            if (vc < V) {
                v[k].x *= rstd; v[k].y *= rstd; v[k].z *= rstd; v[k].w *= rstd;          // xh
                float* dg = my + vc * 4;
                float* db = my + C + vc * 4;
                dg[0] += d[k].x * v[k].x; dg[1] += d[k].y * v[k].y; dg[2] += d[k].z * v[k].z; dg[3] += d[k].w * v[k].w;
                db[0] += d[k].x; db[1] += d[k].y; db[2] += d[k].z; db[3] += d[k].w;
                d[k].x *= g[k].x; d[k].y *= g[k].y; d[k].z *= g[k].z; d[k].w *= g[k].w;  // dy * gamma
                a += (d[k].x + d[k].y) + (d[k].z + d[k].w);
                b += (d[k].x * v[k].x + d[k].y * v[k].y) + (d[k].z * v[k].z + d[k].w * v[k].w);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            a += __shfl_xor_sync(0xffffffffu, a, o);
            b += __shfl_xor_sync(0xffffffffu, b, o);
        }
        a *= invC; b *= invC;
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            const int vc = lane + 32 * k;
            if (vc < V) {
                float4 o4;
                o4.x = rstd * (d[k].x - a - v[k].x * b); o4.y = rstd * (d[k].y - a - v[k].y * b);
                o4.z = rstd * (d[k].z - a - v[k].z * b); o4.w = rstd * (d[k].w - a - v[k].w * b);
                ln_store4<XD>(dx, row_off + vc * 4, o4);
            }
        }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < 2 * C; c += LNB_WARPS * 32) {
        float acc = 0.f;
#pragma unroll
        for (int w = 0; w < LNB_WARPS; ++w) acc += lnb_sm[static_cast<size_t>(w) * 2 * C + c];
        partial[static_cast<size_t>(blockIdx.x) * 2 * C + c] = acc;
    }
}
__global__ void layernorm_bwd_fold_kernel(const float* __restrict__ partial, int nblocks, int C, float* __restrict__ dgamma,
                                          float* __restrict__ dbeta) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= 2 * C) return;
    float acc = 0.f;
    for (int b = 0; b < nblocks; ++b) acc += partial[static_cast<size_t>(b) * 2 * C + c];
    if (c < C) dgamma[c] = acc;
    else dbeta[c - C] = acc;
}
inline int lnb_blocks(int M) {
    const int want = (M + LNB_WARPS - 1) / LNB_WARPS, cap = 2 * sm_count();
    return want < cap ? want : cap;
}

// ---------------------------------------------------------------------------------------------------------
// Row softmax (fp32 logits -> bf16 probabilities), one CTA per row.
// ---------------------------------------------------------------------------------------------------------
__global__ void softmax_rows_kernel(const float* __restrict__ s, uint16_t* __restrict__ p, int L, float scale,
                                    int y_f16) {
    pdl_wait();
    __shared__ float red[32];
    const float* row = s + static_cast<long long>(blockIdx.x) * L;
    uint16_t* out = p + static_cast<long long>(blockIdx.x) * L;
    const int nvec = L / 4;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    float m = -INFINITY;
    for (int i = threadIdx.x; i < nvec; i += blockDim.x) {
        float4 v = __ldg(reinterpret_cast<const float4*>(row) + i);
        m = fmaxf(m, fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) red[warp] = m;
    __syncthreads();
    m = red[0];
    for (int w = 1; w < nwarps; ++w) m = fmaxf(m, red[w]);
    __syncthreads();
    const float c = scale * 1.4426950408889634f;
    const float mc = m * c;
    float sum = 0.f;
    for (int i = threadIdx.x; i < nvec; i += blockDim.x) {
        float4 v = __ldg(reinterpret_cast<const float4*>(row) + i);
        sum += exp2f(fmaf(v.x, c, -mc)) + exp2f(fmaf(v.y, c, -mc)) + exp2f(fmaf(v.z, c, -mc)) +
               exp2f(fmaf(v.w, c, -mc));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (lane == 0) red[warp] = sum;
    __syncthreads();
    sum = 0.f;
    for (int w = 0; w < nwarps; ++w) sum += red[w];
    const float inv = 1.0f / sum;
    for (int i = threadIdx.x; i < nvec; i += blockDim.x) {
        float4 v = __ldg(reinterpret_cast<const float4*>(row) + i);
        uint2 o;
        o.x = pack_h2(exp2f(fmaf(v.x, c, -mc)) * inv, exp2f(fmaf(v.y, c, -mc)) * inv, y_f16);
        o.y = pack_h2(exp2f(fmaf(v.z, c, -mc)) * inv, exp2f(fmaf(v.w, c, -mc)) * inv, y_f16);
        *reinterpret_cast<uint2*>(out + 4 * i) = o;
    }
}

struct GnPlan {
    int V, RPI, threads, nchunks, rows_per_chunk;
    size_t smem;
};
GnPlan gn_plan(int N, int HW, int C) {
    GnPlan pl;
    pl.V = C / 8;
    pl.RPI = pl.V >= 256 ? 1 : 256 / pl.V;
    if (pl.RPI > HW) pl.RPI = HW;
    pl.threads = pl.V * pl.RPI;
    // exactly one wave: the kernels run 3 CTAs of <= 320 threads per SM (register-limited, ncu), and a second partial
    // wave plus the per-CTA statistics fold cost 30 % on the L2-sized tensors
    const int per_sm = get_option(DFW_OPT_GN_CTAS_PER_SM) > 0 ? get_option(DFW_OPT_GN_CTAS_PER_SM) : 3;
    int want = (per_sm * sm_count() + (per_sm > 3 ? N - 1 : 0)) / N;
    if (want < 1) want = 1;
    int max_chunks = (HW + pl.RPI * 16 - 1) / (pl.RPI * 16);  // at least 16 rows per thread
    if (max_chunks < 1) max_chunks = 1;
    pl.nchunks = want < max_chunks ? want : max_chunks;
    if (pl.nchunks < 1) pl.nchunks = 1;
    pl.rows_per_chunk = (HW + pl.nchunks - 1) / pl.nchunks;
    pl.nchunks = (HW + pl.rows_per_chunk - 1) / pl.rows_per_chunk;
    pl.smem = static_cast<size_t>(pl.RPI) * C * 2 * sizeof(float);
    return pl;
}

}  // namespace
}  // namespace dfw

extern "C" {

long long dfw_groupnorm_workspace_bytes(int N, int HW, int C, int groups) {
    if (N <= 0 || HW <= 0 || C <= 0 || groups <= 0 || C % 8 != 0) return -1;
    dfw::GnPlan pl = dfw::gn_plan(N, HW, C);
    long long partial = static_cast<long long>(N) * pl.nchunks * groups * 2;
    long long ss = static_cast<long long>(N) * C * 2;
    return (partial + ss) * static_cast<long long>(sizeof(float)) + 256;
}

int dfw_groupnorm_silu(const void* x, int x_dtype, const float* gamma, const float* beta, void* y, int y_f16, int N,
                       int HW, int C, int groups, float eps, int apply_silu, void* workspace, void* stream_) {
    using namespace dfw;
    int rc = require_sm100();
    if (rc != DFW_OK) return rc;
    DFW_REQUIRE(x && gamma && beta && y && workspace);
    DFW_REQUIRE(x_dtype >= 0 && x_dtype <= 2);
    DFW_REQUIRE(N > 0 && HW > 0 && C > 0 && C % 8 == 0 && groups > 0 && groups <= 64 && C % groups == 0);
    DFW_REQUIRE(C / 8 <= 1024);
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    GnPlan pl = gn_plan(N, HW, C);
    float* partial = reinterpret_cast<float*>(workspace);
    size_t partial_elems = static_cast<size_t>(N) * pl.nchunks * groups * 2;
    partial_elems = (partial_elems + 63) / 64 * 64;
    (void)partial_elems;
    dim3 grid(pl.nchunks, N);
    DFW_REQUIRE(pl.smem <= 48 * 1024);
#define DFW_GN_LAUNCH(XD)                                                                                          \
    launch_k(gn_stats_kernel<XD>, grid, pl.threads, pl.smem, stream, x, partial, HW, C, groups, pl.rows_per_chunk, pl.RPI); \
    launch_k(gn_apply_kernel<XD>, grid, pl.threads, 0, stream, x, partial, gamma, beta, y, HW, C, groups, pl.nchunks, eps,  \
             pl.rows_per_chunk, pl.RPI, apply_silu, y_f16);
    if (x_dtype == 1) { DFW_GN_LAUNCH(1) } else if (x_dtype == 2) { DFW_GN_LAUNCH(2) } else { DFW_GN_LAUNCH(0) }
#undef DFW_GN_LAUNCH
    g_launches.fetch_add(2);
    DFW_CHECK_CUDA(cudaGetLastError());
    return DFW_OK;
}

int dfw_groupnorm_from_partial(const void* x, int x_dtype, const float* partial, int nchunks, const float* gamma,
                               const float* beta, void* y, int y_f16, int N, int HW, int C, int groups, float eps,
                               int apply_silu, void* stream_) {
    using namespace dfw;
    int rc = require_sm100();
    if (rc != DFW_OK) return rc;
    DFW_REQUIRE(x && partial && gamma && beta && y && nchunks > 0);
    DFW_REQUIRE(x_dtype >= 0 && x_dtype <= 2);
    DFW_REQUIRE(N > 0 && HW > 0 && C > 0 && C % 8 == 0 && groups > 0 && groups <= 64 && C % groups == 0 && C / 8 <= 1024);
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    GnPlan pl = gn_plan(N, HW, C);
    dim3 grid(pl.nchunks, N);
#define DFW_GN_APPLY(XD)                                                                                            \
    launch_k(gn_apply_kernel<XD>, grid, pl.threads, 0, stream, x, partial, gamma, beta, y, HW, C, groups, nchunks, eps,      \
             pl.rows_per_chunk, pl.RPI, apply_silu, y_f16);
    if (x_dtype == 1) { DFW_GN_APPLY(1) } else if (x_dtype == 2) { DFW_GN_APPLY(2) } else { DFW_GN_APPLY(0) }
#undef DFW_GN_APPLY
    g_launches.fetch_add(1);
    DFW_CHECK_CUDA(cudaGetLastError());
    return DFW_OK;
}

long long dfw_groupnorm_bwd_workspace_bytes(int N, int HW, int C, int groups) {
    if (N <= 0 || HW <= 0 || C <= 0 || groups <= 0 || C % 8 != 0) return -1;
    const dfw::GnPlan pl = dfw::gn_plan(N, HW, C);
    const size_t fl = static_cast<size_t>(N) * pl.nchunks * groups * 2 + static_cast<size_t>(N) * pl.nchunks * C * 2 +
                      static_cast<size_t>(N) * C * 2 + static_cast<size_t>(N) * groups * 4;
    return static_cast<long long>(fl * sizeof(float) + 256);
}

int dfw_groupnorm_silu_bwd(const void* x, const void* dy, int dtype, const float* gamma, const float* beta, void* dx,
                           float* dgamma, float* dbeta, int N, int HW, int C, int groups, float eps, int apply_silu,
                           void* workspace, void* stream_) {
    using namespace dfw;
    int rc = require_sm100();
    if (rc != DFW_OK) return rc;
    DFW_REQUIRE(x && dy && gamma && beta && dx && dgamma && dbeta && workspace);
    DFW_REQUIRE(dtype >= 0 && dtype <= 2);
    DFW_REQUIRE(N > 0 && HW > 0 && C > 0 && C % 8 == 0 && groups > 0 && groups <= 64 && C % groups == 0 && C / 8 <= 1024);
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    GnPlan pl = gn_plan(N, HW, C);
    DFW_REQUIRE(pl.smem <= 48 * 1024 && static_cast<size_t>(C) * 2 * sizeof(float) <= 40 * 1024);
    float* partial = reinterpret_cast<float*>(workspace);
    float* ab_partial = partial + static_cast<size_t>(N) * pl.nchunks * groups * 2;
    float* ab = ab_partial + static_cast<size_t>(N) * pl.nchunks * C * 2;
    float* gs = ab + static_cast<size_t>(N) * C * 2;
    float* ms = gs + static_cast<size_t>(N) * groups * 2;
    dim3 grid(pl.nchunks, N);
#define DFW_GN_BWD(XD)                                                                                                  \
    gn_stats_kernel<XD><<<grid, pl.threads, pl.smem, stream>>>(x, partial, HW, C, groups, pl.rows_per_chunk, pl.RPI);   \
    gn_bwd_reduce_kernel<XD><<<grid, pl.threads, pl.smem, stream>>>(x, dy, partial, gamma, beta, ab_partial, HW, C,     \
                                                                    groups, eps, pl.rows_per_chunk, pl.RPI, apply_silu); \
    gn_bwd_fold_kernel<<<N, 256, static_cast<size_t>(C) * 2 * sizeof(float), stream>>>(partial, ab_partial, gamma, ab, gs, \
                                                                                       ms, pl.nchunks, HW, C, groups, eps); \
    gn_bwd_param_kernel<<<(C + 255) / 256, 256, 0, stream>>>(ab, dgamma, dbeta, N, C);                                  \
    gn_bwd_apply_kernel<XD><<<grid, pl.threads, 0, stream>>>(x, dy, gs, ms, gamma, beta, dx, HW, C, groups,             \
                                                             pl.rows_per_chunk, pl.RPI, apply_silu);
    if (dtype == 1) { DFW_GN_BWD(1) } else if (dtype == 2) { DFW_GN_BWD(2) } else { DFW_GN_BWD(0) }
#undef DFW_GN_BWD
    g_launches.fetch_add(5);
    DFW_CHECK_CUDA(cudaGetLastError());
    return DFW_OK;
}

int dfw_gn_scale_shift(const float* partial, int nchunks, const float* gamma, const float* beta, float* scale_shift,
                       int N, long long HW, int C, int groups, float eps, void* stream_) {
    using namespace dfw;
    int rc = require_sm100();
    if (rc != DFW_OK) return rc;
    DFW_REQUIRE(partial && gamma && beta && scale_shift && nchunks > 0);
    DFW_REQUIRE(N > 0 && HW > 0 && C > 0 && groups > 0 && groups <= 64 && C % groups == 0);
    gn_scale_shift_kernel<<<N, 256, 0, static_cast<cudaStream_t>(stream_)>>>(partial, nchunks, gamma, beta, scale_shift,
                                                                            static_cast<double>(HW), C, groups, eps);
    g_launches.fetch_add(1);
    DFW_CHECK_CUDA(cudaGetLastError());
    return DFW_OK;
}

int dfw_layernorm(const void* x, int x_dtype, const float* gamma, const float* beta, void* y, int y_f16, int M, int C,
                  float eps, void* stream_) {
    using namespace dfw;
    int rc = require_sm100();
    if (rc != DFW_OK) return rc;
    DFW_REQUIRE(x && gamma && beta && y && M > 0 && C > 0 && C % 8 == 0 && C <= 2048);
    DFW_REQUIRE(x_dtype >= 0 && x_dtype <= 2);
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const int warps_per_block = 8;
    const int blocks = (M + warps_per_block - 1) / warps_per_block;
    if (C % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 &&
        (x_dtype == 1 ? launch_layernorm_v4<1>(x, gamma, beta, y, M, C, eps, y_f16, stream)
         : x_dtype == 2 ? launch_layernorm_v4<2>(x, gamma, beta, y, M, C, eps, y_f16, stream)
                        : launch_layernorm_v4<0>(x, gamma, beta, y, M, C, eps, y_f16, stream))) {
    } else if (x_dtype == 1)
        layernorm_kernel<1><<<blocks, warps_per_block * 32, 0, stream>>>(x, gamma, beta, y, M, C, eps, y_f16);
    else if (x_dtype == 2)
        layernorm_kernel<2><<<blocks, warps_per_block * 32, 0, stream>>>(x, gamma, beta, y, M, C, eps, y_f16);
    else
        layernorm_kernel<0><<<blocks, warps_per_block * 32, 0, stream>>>(x, gamma, beta, y, M, C, eps, y_f16);
    g_launches.fetch_add(1);
    DFW_CHECK_CUDA(cudaGetLastError());
    return DFW_OK;
}

long long dfw_layernorm_bwd_workspace_bytes(int M, int C) {
    if (M <= 0 || C <= 0) return -1;
    return static_cast<long long>(dfw::lnb_blocks(M)) * 2 * C * sizeof(float);
}

int dfw_layernorm_bwd(const void* x, const void* dy, int dtype, const float* gamma, void* dx, float* dgamma, float* dbeta,
                      int M, int C, float eps, void* workspace, void* stream_) {
    using namespace dfw;
    int rc = require_sm100();
    if (rc != DFW_OK) return rc;
    DFW_REQUIRE(x && dy && gamma && dx && dgamma && dbeta && workspace && M > 0 && C > 0 && C % 8 == 0 && C <= 1280);
    DFW_REQUIRE(dtype >= 0 && dtype <= 2);
    DFW_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(dx)) & 15) == 0);
    cudaStream_t st = static_cast<cudaStream_t>(stream_);
    const int blocks = lnb_blocks(M);
    const size_t smem = static_cast<size_t>(LNB_WARPS) * 2 * C * sizeof(float);
    float* partial = reinterpret_cast<float*>(workspace);
    const int nv = (C / 4 + 31) / 32;
#define DFW_LNB_LAUNCH(XD, NV)                                                                                      \
    do {                                                                                                            \
        DFW_CHECK_CUDA(cudaFuncSetAttribute(layernorm_bwd_kernel<XD, NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                            static_cast<int>(smem)));                                              \
        layernorm_bwd_kernel<XD, NV><<<blocks, LNB_WARPS * 32, smem, st>>>(x, dy, gamma, dx, partial, M, C, eps);      \
    } while (0)
#define DFW_LNB_DTYPE(NV)                                  \
    do {                                                   \
        if (dtype == 1) DFW_LNB_LAUNCH(1, NV);             \
        else if (dtype == 2) DFW_LNB_LAUNCH(2, NV);        \
        else DFW_LNB_LAUNCH(0, NV);                        \
    } while (0)
    if (nv <= 3) DFW_LNB_DTYPE(3);
    else if (nv <= 5) DFW_LNB_DTYPE(5);
    else DFW_LNB_DTYPE(10);
#undef DFW_LNB_DTYPE
#undef DFW_LNB_LAUNCH
    layernorm_bwd_fold_kernel<<<(2 * C + 127) / 128, 128, 0, st>>>(partial, blocks, C, dgamma, dbeta);
    g_launches.fetch_add(2);
    DFW_CHECK_CUDA(cudaGetLastError());
    return DFW_OK;
}

int dfw_softmax_rows(const float* s, void* p, int y_f16, int M, int L, float scale, void* stream_) {
    using namespace dfw;
    int rc = require_sm100();
    if (rc != DFW_OK) return rc;
    DFW_REQUIRE(s && p && M > 0 && L > 0 && L % 4 == 0);
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    launch_k(softmax_rows_kernel, M, 256, 0, stream, s, reinterpret_cast<uint16_t*>(p), L, scale, y_f16);
    g_launches.fetch_add(1);
    DFW_CHECK_CUDA(cudaGetLastError());
    return DFW_OK;
}

}  // extern "C"
