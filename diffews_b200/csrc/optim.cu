// K10 — training-step tail (first pieces): MSE loss forward + backward, global gradient-norm clipping
// and a fused multi-tensor AdamW step.  All HBM-bound; one launch covers every parameter tensor of the model through a
// descriptor table (the SD-2.1 UNet has ~690 tensors / 866 M parameters).
//
// ref: train_tools/train_icl_multitask_nocrop_nearest_nshot_v3.py:1384   loss = F.mse_loss(pred.float(), target.float(), "mean")
//      :1393  accelerator.clip_grad_norm_(unet.parameters(), max_grad_norm)   (torch.nn.utils.clip_grad_norm_, L2)
//      :1186-1194, :1394  torch.optim.AdamW(lr, betas, weight_decay, eps).step()   (torch/optim/adamw.py single-tensor path:
//             p *= 1 - lr*wd ; m.lerp_(g, 1-b1) ; v = v*b2 + (1-b2) g g ; p += -(lr / (1-b1^t)) * m / (sqrt(v)/sqrt(1-b2^t) + eps))
// fp32 parameters, gradients and moments (the reference trains fp32 weights under fp16 autocast); optionally the step also
// writes the 16-bit tensor-core operand copy of every parameter, which saves the separate cast pass of the next forward.
//
// Determinism: every reduction is a fixed-shape tree (per-chunk partials, then one block folds the partials in index
// order) — no floating-point atomics.
#include <atomic>
#include <cmath>

#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.cuh"

namespace dfw {
extern std::atomic<long long> g_launches;
namespace {

struct AdamTensor {                 // mirrors DfwAdamTensor (include/diffews_b200.h)
    float* p;
    const float* g;
    float* m;
    float* v;
    uint16_t* p16;
    long long n;
};

constexpr int OPT_THREADS = 256;

__device__ __forceinline__ float block_sum(float x, float* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) red[w] = x;
    __syncthreads();
    float s = 0.f;
    if (w == 0) {
        s = l < OPT_THREADS / 32 ? red[l] : 0.f;
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    }
    __syncthreads();
    return s;                        // valid in warp 0
}

// per-chunk sum of squares of the gradients (chunk c of tensor chunk_tensor[c] starts at element chunk_offset[c])
__global__ void __launch_bounds__(OPT_THREADS) sumsq_chunks_kernel(const AdamTensor* __restrict__ ts,
                                                                   const int* __restrict__ chunk_tensor,
                                                                   const long long* __restrict__ chunk_offset,
                                                                   int chunk_elems, float* __restrict__ partial) {
    __shared__ float red[OPT_THREADS / 32];
    const AdamTensor t = ts[chunk_tensor[blockIdx.x]];
    const long long off = chunk_offset[blockIdx.x];
    const long long n = min(static_cast<long long>(chunk_elems), t.n - off);
    const float* g = t.g + off;
    float acc = 0.f;
    if ((reinterpret_cast<uintptr_t>(g) & 15) == 0) {
        const long long n4 = n >> 2;
        for (long long i = threadIdx.x; i < n4; i += OPT_THREADS) {
            const float4 x = __ldg(reinterpret_cast<const float4*>(g) + i);
            acc += (x.x * x.x + x.y * x.y) + (x.z * x.z + x.w * x.w);
        }
        for (long long i = (n4 << 2) + threadIdx.x; i < n; i += OPT_THREADS) acc += g[i] * g[i];
    } else {
        for (long long i = threadIdx.x; i < n; i += OPT_THREADS) acc += g[i] * g[i];
    }
    const float s = block_sum(acc, red);
    if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

// one block: total = sum of partials in index order (fixed tree), norm = sqrt(total),
// coef = min(1, max_norm / (norm + 1e-6))   (torch.nn.utils.clip_grad_norm_)
__global__ void __launch_bounds__(OPT_THREADS) clip_coef_kernel(const float* __restrict__ partial, int n, float max_norm,
                                                                float* __restrict__ norm_out, float* __restrict__ coef_out) {
    __shared__ float red[OPT_THREADS / 32];
    float acc = 0.f;
    for (int i = threadIdx.x; i < n; i += OPT_THREADS) acc += partial[i];
    const float s = block_sum(acc, red);
    if (threadIdx.x == 0) {
        const float norm = sqrtf(s);
        norm_out[0] = norm;
        const float c = max_norm / (norm + 1e-6f);
        // a non-finite norm (an inf / NaN gradient, e.g. an fp16 overflow under the reference's GradScaler) poisons the
        // coefficient on purpose: dfw_adamw_step skips the whole update when it reads a non-finite grad_scale
        coef_out[0] = isfinite(norm) ? (c < 1.0f ? c : 1.0f) : __int_as_float(0x7fc00000);
    }
}

struct AdamScalars {
    float decay;          // 1 - lr * weight_decay
    float w1;             // 1 - beta1
    float beta2, w2;      // beta2, 1 - beta2
    float bc2_sqrt;       // sqrt(1 - beta2^t)
    float neg_step;       // -(lr / (1 - beta1^t))
    float eps;
};

template <int P16>   // 0: no 16-bit copy, 1: bf16, 2: fp16
__global__ void __launch_bounds__(OPT_THREADS) adamw_chunks_kernel(const AdamTensor* __restrict__ ts,
                                                                   const int* __restrict__ chunk_tensor,
                                                                   const long long* __restrict__ chunk_offset,
                                                                   int chunk_elems, const AdamScalars sc,
                                                                   const float* __restrict__ grad_scale) {
    const AdamTensor t = ts[chunk_tensor[blockIdx.x]];
    const long long off = chunk_offset[blockIdx.x];
    const long long n = min(static_cast<long long>(chunk_elems), t.n - off);
    const float gs = grad_scale ? __ldg(grad_scale) : 1.0f;
    if (!isfinite(gs)) return;        // found-inf guard: parameters and moments stay untouched (GradScaler semantics)
    auto upd = [&](float& p, float g, float& m, float& v) {
        g *= gs;
        p *= sc.decay;
        m = m + sc.w1 * (g - m);
        v = v * sc.beta2 + sc.w2 * (g * g);
        const float denom = sqrtf(v) / sc.bc2_sqrt + sc.eps;
        p = p + sc.neg_step * (m / denom);
    };
    auto to16 = [&](float x) -> uint16_t {
        if (P16 == 2) return __half_as_ushort(__float2half_rn(x));
        return __bfloat16_as_ushort(__float2bfloat16_rn(x));
    };
    float* p = t.p + off;
    const float* g = t.g + off;
    float* m = t.m + off;
    float* v = t.v + off;
    uint16_t* p16 = P16 ? t.p16 + off : nullptr;
    const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                       reinterpret_cast<uintptr_t>(v)) & 15) == 0 && (!P16 || (reinterpret_cast<uintptr_t>(p16) & 7) == 0);
    long long done = 0;
    if (vec) {
        const long long n4 = n >> 2;
        for (long long i = threadIdx.x; i < n4; i += OPT_THREADS) {
            float4 pp = reinterpret_cast<float4*>(p)[i];
            const float4 gg = __ldg(reinterpret_cast<const float4*>(g) + i);
            float4 mm = reinterpret_cast<float4*>(m)[i];
            float4 vv = reinterpret_cast<float4*>(v)[i];
            upd(pp.x, gg.x, mm.x, vv.x); upd(pp.y, gg.y, mm.y, vv.y); upd(pp.z, gg.z, mm.z, vv.z); upd(pp.w, gg.w, mm.w, vv.w);
            reinterpret_cast<float4*>(p)[i] = pp;
            reinterpret_cast<float4*>(m)[i] = mm;
            reinterpret_cast<float4*>(v)[i] = vv;
            if (P16) {
                uint2 h;
                h.x = to16(pp.x) | (static_cast<uint32_t>(to16(pp.y)) << 16);
                h.y = to16(pp.z) | (static_cast<uint32_t>(to16(pp.w)) << 16);
                reinterpret_cast<uint2*>(p16)[i] = h;
            }
        }
        done = n4 << 2;
    }
    for (long long i = done + threadIdx.x; i < n; i += OPT_THREADS) {
        float pp = p[i], mm = m[i], vv = v[i];
        upd(pp, g[i], mm, vv);
        p[i] = pp; m[i] = mm; v[i] = vv;
        if (P16) p16[i] = to16(pp);
    }
}

// MSE: per-block partial of (pred - target)^2 and, in the same pass, dpred = upstream * 2 (pred - target) / n
__global__ void __launch_bounds__(OPT_THREADS) mse_partial_kernel(const float* __restrict__ pred,
                                                                  const float* __restrict__ target, long long n,
                                                                  float grad_mul, float* __restrict__ dpred,
                                                                  float* __restrict__ partial) {
    __shared__ float red[OPT_THREADS / 32];
    float acc = 0.f;
    for (long long i = static_cast<long long>(blockIdx.x) * OPT_THREADS + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * OPT_THREADS) {
        const float d = pred[i] - target[i];
        acc += d * d;
        if (dpred) dpred[i] = grad_mul * d;
    }
    const float s = block_sum(acc, red);
    if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

__global__ void __launch_bounds__(OPT_THREADS) mse_final_kernel(const float* __restrict__ partial, int n, float inv_n,
                                                                float* __restrict__ loss) {
    __shared__ float red[OPT_THREADS / 32];
    float acc = 0.f;
    for (int i = threadIdx.x; i < n; i += OPT_THREADS) acc += partial[i];
    const float s = block_sum(acc, red);
    if (threadIdx.x == 0) loss[0] = s * inv_n;
}

constexpr int MSE_BLOCKS = 592;     // 4 x 148 SMs

// GEGLU backward (diffusers GEGLU: v, g = proj(x).chunk(2, -1); y = v * gelu_erf(g)):
//   dv = dy * gelu(g) ;  dg = dy * v * (Phi(g) + g * phi(g)),  Phi = 0.5 (1 + erf(g / sqrt 2)), phi = exp(-g^2 / 2) / sqrt(2 pi)
// h [M, 2F] (value | gate), dy [M, F], dh [M, 2F], one dtype (0 bf16 / 1 fp32 / 2 fp16).  Thread = 4 consecutive columns.
template <int XD>
__device__ __forceinline__ float4 load4(const void* x, long long off) {
    if constexpr (XD == 1) {
        return __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(x) + off));
    } else {
        const uint2 r = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const uint16_t*>(x) + off));
        if constexpr (XD == 2) {
            const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&r.x)), b = __half22float2(*reinterpret_cast<const __half2*>(&r.y));
            return make_float4(a.x, a.y, b.x, b.y);
        } else {
            const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&r.x)), b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&r.y));
            return make_float4(a.x, a.y, b.x, b.y);
        }
    }
}
template <int XD>
__device__ __forceinline__ void store4(void* y, long long off, float4 v) {
    if constexpr (XD == 1) {
        *reinterpret_cast<float4*>(reinterpret_cast<float*>(y) + off) = v;
    } else if constexpr (XD == 2) {
        uint2 o;
        *reinterpret_cast<__half2*>(&o.x) = __floats2half2_rn(v.x, v.y);
        *reinterpret_cast<__half2*>(&o.y) = __floats2half2_rn(v.z, v.w);
        *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(y) + off) = o;
    } else {
        uint2 o;
        *reinterpret_cast<__nv_bfloat162*>(&o.x) = __floats2bfloat162_rn(v.x, v.y);
        *reinterpret_cast<__nv_bfloat162*>(&o.y) = __floats2bfloat162_rn(v.z, v.w);
        *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(y) + off) = o;
    }
}
template <int XD>
__global__ void __launch_bounds__(OPT_THREADS) geglu_bwd_kernel(const void* __restrict__ h, const void* __restrict__ dy,
                                                                void* __restrict__ dh, long long M, int F) {
    const int F4 = F / 4;
    const long long total = M * F4;
    for (long long i = static_cast<long long>(blockIdx.x) * OPT_THREADS + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * OPT_THREADS) {
        const long long row = i / F4;
        const int c = static_cast<int>(i - row * F4) * 4;
        const float4 v = load4<XD>(h, row * 2 * F + c), g = load4<XD>(h, row * 2 * F + F + c), d = load4<XD>(dy, row * F + c);
        auto one = [](float vv, float gg, float dd, float& dv, float& dg) {
            const float cdf = 0.5f * (1.0f + erff(gg * 0.70710678118654752f));
            const float pdf = 0.39894228040143268f * expf(-0.5f * gg * gg);
            dv = dd * (gg * cdf);
            dg = dd * vv * (cdf + gg * pdf);
        };
        float4 dv, dg;
        one(v.x, g.x, d.x, dv.x, dg.x); one(v.y, g.y, d.y, dv.y, dg.y); one(v.z, g.z, d.z, dv.z, dg.z); one(v.w, g.w, d.w, dv.w, dg.w);
        store4<XD>(dh, row * 2 * F + c, dv);
        store4<XD>(dh, row * 2 * F + F + c, dg);
    }
}

}  // namespace
}  // namespace dfw

extern "C" {

int dfw_grad_norm_clip_coef(const void* tensors, const int* chunk_tensor, const long long* chunk_offset, int n_chunks,
                            int chunk_elems, float max_norm, float* partial, float* norm_out, float* coef_out,
                            void* stream_) {
    using namespace dfw;
    int rc = require_sm100();
    if (rc != DFW_OK) return rc;
    DFW_REQUIRE(tensors && chunk_tensor && chunk_offset && partial && norm_out && coef_out);
    DFW_REQUIRE(n_chunks > 0 && chunk_elems > 0 && chunk_elems % 4 == 0 && max_norm > 0.0f);
    cudaStream_t st = static_cast<cudaStream_t>(stream_);
    sumsq_chunks_kernel<<<n_chunks, OPT_THREADS, 0, st>>>(reinterpret_cast<const AdamTensor*>(tensors), chunk_tensor,
                                                           chunk_offset, chunk_elems, partial);
    clip_coef_kernel<<<1, OPT_THREADS, 0, st>>>(partial, n_chunks, max_norm, norm_out, coef_out);
    g_launches.fetch_add(2);
    DFW_CHECK_CUDA(cudaGetLastError());
    return DFW_OK;
}

int dfw_adamw_step(const void* tensors, const int* chunk_tensor, const long long* chunk_offset, int n_chunks,
                   int chunk_elems, double lr, double beta1, double beta2, double eps, double weight_decay, int step,
                   const float* grad_scale, int p16_format, void* stream_) {
    using namespace dfw;
    int rc = require_sm100();
    if (rc != DFW_OK) return rc;
    DFW_REQUIRE(tensors && chunk_tensor && chunk_offset && n_chunks > 0 && chunk_elems > 0 && chunk_elems % 4 == 0);
    DFW_REQUIRE(step >= 1 && beta1 >= 0.0 && beta1 < 1.0 && beta2 >= 0.0 && beta2 < 1.0 && eps >= 0.0);
    DFW_REQUIRE(p16_format >= 0 && p16_format <= 2);
    // the scalars torch computes in Python doubles (torch/optim/adamw.py _single_tensor_adamw), then rounds to fp32
    AdamScalars sc;
    // (hyper-parameters arrive as doubles: torch keeps them as Python floats, and 1 - float(0.999) is 1.3e-5 off 0.001)
    sc.decay = static_cast<float>(1.0 - lr * weight_decay);
    sc.w1 = static_cast<float>(1.0 - beta1);
    sc.beta2 = static_cast<float>(beta2);
    sc.w2 = static_cast<float>(1.0 - beta2);
    const double bc1 = 1.0 - std::pow(beta1, step);
    const double bc2 = 1.0 - std::pow(beta2, step);
    sc.bc2_sqrt = static_cast<float>(std::sqrt(bc2));
    sc.neg_step = static_cast<float>(-(lr / bc1));
    sc.eps = static_cast<float>(eps);
    cudaStream_t st = static_cast<cudaStream_t>(stream_);
    const AdamTensor* ts = reinterpret_cast<const AdamTensor*>(tensors);
    if (p16_format == 0) adamw_chunks_kernel<0><<<n_chunks, OPT_THREADS, 0, st>>>(ts, chunk_tensor, chunk_offset, chunk_elems, sc, grad_scale);
    else if (p16_format == 1) adamw_chunks_kernel<1><<<n_chunks, OPT_THREADS, 0, st>>>(ts, chunk_tensor, chunk_offset, chunk_elems, sc, grad_scale);
    else adamw_chunks_kernel<2><<<n_chunks, OPT_THREADS, 0, st>>>(ts, chunk_tensor, chunk_offset, chunk_elems, sc, grad_scale);
    g_launches.fetch_add(1);
    DFW_CHECK_CUDA(cudaGetLastError());
    return DFW_OK;
}

long long dfw_mse_workspace_floats(void) { return dfw::MSE_BLOCKS; }

int dfw_mse_loss(const float* pred, const float* target, long long n, float upstream, float* loss_out, float* dpred,
                 float* workspace, void* stream_) {
    using namespace dfw;
    int rc = require_sm100();
    if (rc != DFW_OK) return rc;
    DFW_REQUIRE(pred && target && loss_out && workspace && n > 0);
    cudaStream_t st = static_cast<cudaStream_t>(stream_);
    long long want = (n + OPT_THREADS - 1) / OPT_THREADS;
    const int blocks = static_cast<int>(want < MSE_BLOCKS ? want : MSE_BLOCKS);
    const float inv_n = static_cast<float>(1.0 / static_cast<double>(n));
    mse_partial_kernel<<<blocks, OPT_THREADS, 0, st>>>(pred, target, n, upstream * 2.0f * inv_n, dpred, workspace);
    mse_final_kernel<<<1, OPT_THREADS, 0, st>>>(workspace, blocks, inv_n, loss_out);
    g_launches.fetch_add(2);
    DFW_CHECK_CUDA(cudaGetLastError());
    return DFW_OK;
}

int dfw_geglu_bwd(const void* h, const void* dy, void* dh, int dtype, long long M, int F, void* stream_) {
    using namespace dfw;
    int rc = require_sm100();
    if (rc != DFW_OK) return rc;
    DFW_REQUIRE(h && dy && dh && M > 0 && F > 0 && F % 8 == 0 && dtype >= 0 && dtype <= 2);
    DFW_REQUIRE(((reinterpret_cast<uintptr_t>(h) | reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(dh)) & 15) == 0);
    cudaStream_t st = static_cast<cudaStream_t>(stream_);
    const long long total = M * (F / 4);
    long long want = (total + OPT_THREADS - 1) / OPT_THREADS;
    const long long cap = 16LL * sm_count();
    const int blocks = static_cast<int>(want < cap ? want : cap);
    if (dtype == 1) geglu_bwd_kernel<1><<<blocks, OPT_THREADS, 0, st>>>(h, dy, dh, M, F);
    else if (dtype == 2) geglu_bwd_kernel<2><<<blocks, OPT_THREADS, 0, st>>>(h, dy, dh, M, F);
    else geglu_bwd_kernel<0><<<blocks, OPT_THREADS, 0, st>>>(h, dy, dh, M, F);
    g_launches.fetch_add(1);
    DFW_CHECK_CUDA(cudaGetLastError());
    return DFW_OK;
}

}  // extern "C"
