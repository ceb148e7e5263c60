// K1b — backward of the KV-fused attention (BASELINE config 4: training-shape forward + backward, query tokens against
// their own plus the 7 supports' keys / values; ref: the autograd of xformers.ops.memory_efficient_attention on
// cat([key, folded bank]) in diffews/models/attention_processor.py:251-271 under the training step
// train_tools/train_icl_multitask_nocrop_nearest_nshot_v3.py:1320-1396).
//
// First correct CUDA path, not a flash-style fused kernel: every contraction runs on the tcgen05 batched GEMM
// (dfw_bmm_nt), with the logits materialised per (episode, head) in the caller's workspace:
//     S   = scale Q K^T            S^T  = scale K Q^T                    (two orientations instead of transposing L x L)
//     lse = row logsumexp(S),  delta = rowsum(dO . O)
//     dP  = dO V^T                 dP^T = V dO^T
//     dS  = exp(S - lse) (dP - delta)          [rows = queries]          dS^T, P^T likewise with column statistics
//     dQ  = scale dS K,   dK = scale dS^T Q,   dV = P^T dO               (K / Q / dO transposed head-major copies as the
//                                                                         "weight" operand: contraction over L)
// K = [K_self ; K_bank] is gathered head-major once; dK / dV are scattered back to the self and bank tensors.
// Deterministic (no atomics).  Memory: 4 fp32 + 3 16-bit L_q x L_k matrices per (episode, head).
#include <atomic>

#include "common.cuh"
#include "ptx.cuh"

namespace dfw {
extern std::atomic<long long> g_launches;
namespace {

constexpr int D = 64;

// src 16-bit [B, L, row_stride] (+ head * 64)  ->  dst [B*h, Ldst, 64] rows [row0, row0 + L)  and, if dstT != nullptr,
// dstT [B*h, 64, Ldst] columns [row0, row0 + L).  One thread = 8 channels of one (b, head, l).
__global__ void gather_heads_kernel(const uint16_t* __restrict__ src, long long batch_stride, int row_stride,
                                    uint16_t* __restrict__ dst, uint16_t* __restrict__ dstT, int B, int heads, int L,
                                    int Ldst, int row0) {
    const long long total = static_cast<long long>(B) * heads * L * 8;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int u = static_cast<int>(i & 7);
        long long t = i >> 3;
        const int l = static_cast<int>(t % L); t /= L;
        const int h = static_cast<int>(t % heads);
        const int b = static_cast<int>(t / heads);
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(src + b * batch_stride + static_cast<long long>(l) * row_stride +
                                                            h * D + u * 8));
        const long long bh = static_cast<long long>(b) * heads + h;
        *reinterpret_cast<uint4*>(dst + (bh * Ldst + row0 + l) * D + u * 8) = v;
        if (dstT != nullptr) {
            const uint16_t* e = reinterpret_cast<const uint16_t*>(&v);
#pragma unroll
            for (int j = 0; j < 8; ++j) dstT[(bh * D + u * 8 + j) * Ldst + row0 + l] = e[j];
        }
    }
}

// delta[bh, l] = sum_d dO * O  (both in the caller's [B, L, row_stride] layout); 8 lanes per (b, head, l)
__global__ void delta_kernel(const uint16_t* __restrict__ o, long long o_bs, int o_rs, const uint16_t* __restrict__ dout,
                             long long do_bs, int do_rs, float* __restrict__ delta, int B, int heads, int L, int f16) {
    const long long total = static_cast<long long>(B) * heads * L * 8;
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    const bool active = i < total;
    const long long ii = active ? i : 0;
    const int u = static_cast<int>(ii & 7);
    long long t = ii >> 3;
    const int l = static_cast<int>(t % L); t /= L;
    const int h = static_cast<int>(t % heads);
    const int b = static_cast<int>(t / heads);
    const uint4 a = __ldg(reinterpret_cast<const uint4*>(o + b * o_bs + static_cast<long long>(l) * o_rs + h * D + u * 8));
    const uint4 g = __ldg(reinterpret_cast<const uint4*>(dout + b * do_bs + static_cast<long long>(l) * do_rs + h * D + u * 8));
    const uint32_t* aw = reinterpret_cast<const uint32_t*>(&a);
    const uint32_t* gw = reinterpret_cast<const uint32_t*>(&g);
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float2 x = unpack_h2(aw[j], f16), y = unpack_h2(gw[j], f16);
        s = fmaf(x.x, y.x, fmaf(x.y, y.y, s));
    }
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    if (active && u == 0) delta[(static_cast<long long>(b) * heads + h) * L + l] = s;
}

// lse2[row] = log2(sum_j exp2(S[row, j] * log2e))  for fp32 S [rows, Lk]; one CTA per row
__global__ void row_lse_kernel(const float* __restrict__ S, float* __restrict__ lse2, int Lk) {
    __shared__ float red[32];
    const float* row = S + static_cast<long long>(blockIdx.x) * Lk;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    constexpr float L2E = 1.4426950408889634f;
    float m = -INFINITY;
    for (int j = threadIdx.x; j < Lk; j += blockDim.x) m = fmaxf(m, __ldg(row + j));
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) red[warp] = m;
    __syncthreads();
    m = red[0];
    for (int w = 1; w < nw; ++w) m = fmaxf(m, red[w]);
    __syncthreads();
    float s = 0.f;
    for (int j = threadIdx.x; j < Lk; j += blockDim.x) s += exp2f((__ldg(row + j) - m) * L2E);
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) red[warp] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < nw; ++w) t += red[w];
        lse2[blockIdx.x] = m * L2E + log2f(t);
    }
}

// p = exp2(S * log2e - lse2[q]);  dS = p * (dP - delta[q])  -> 16-bit.  S, dP fp32 [BH, R, Ccols]; q = the QUERY index of
// an element: its row (rows_are_queries) or its column.  lse2 / delta [BH, Lq].  4 elements per thread.
__global__ void ds_kernel(const float* __restrict__ S, const float* __restrict__ dP, const float* __restrict__ lse2,
                          const float* __restrict__ delta, uint16_t* __restrict__ outP, uint16_t* __restrict__ outdS,
                          long long total4, int R, int Ccols, int Lq, int rows_are_queries, int f16) {
    constexpr float L2E = 1.4426950408889634f;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total4;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long e0 = i * 4;
        const int c0 = static_cast<int>(e0 % Ccols);
        const long long rr = e0 / Ccols;
        const int r = static_cast<int>(rr % R);
        const long long bh = rr / R;
        const float4 s = __ldg(reinterpret_cast<const float4*>(S + e0));
        const float4 g = __ldg(reinterpret_cast<const float4*>(dP + e0));
        float l[4], d[4];
        if (rows_are_queries) {
            const float lv = __ldg(lse2 + bh * Lq + r), dv = __ldg(delta + bh * Lq + r);
            l[0] = l[1] = l[2] = l[3] = lv; d[0] = d[1] = d[2] = d[3] = dv;
        } else {
            const float4 lv = __ldg(reinterpret_cast<const float4*>(lse2 + bh * Lq + c0));
            const float4 dv = __ldg(reinterpret_cast<const float4*>(delta + bh * Lq + c0));
            l[0] = lv.x; l[1] = lv.y; l[2] = lv.z; l[3] = lv.w; d[0] = dv.x; d[1] = dv.y; d[2] = dv.z; d[3] = dv.w;
        }
        const float p0 = exp2f(fmaf(s.x, L2E, -l[0])), p1 = exp2f(fmaf(s.y, L2E, -l[1])),
                    p2 = exp2f(fmaf(s.z, L2E, -l[2])), p3 = exp2f(fmaf(s.w, L2E, -l[3]));
        if (outP != nullptr) {
            uint2 o; o.x = pack_h2(p0, p1, f16); o.y = pack_h2(p2, p3, f16);
            *reinterpret_cast<uint2*>(outP + e0) = o;
        }
        uint2 o; o.x = pack_h2(p0 * (g.x - d[0]), p1 * (g.y - d[1]), f16); o.y = pack_h2(p2 * (g.z - d[2]), p3 * (g.w - d[3]), f16);
        *reinterpret_cast<uint2*>(outdS + e0) = o;
    }
}

// src fp32 [B*h, Lsrc, 64] rows [row0, row0 + L)  ->  dst 16-bit [B, L, row_stride] (+ head * 64)
__global__ void scatter_heads_kernel(const float* __restrict__ src, uint16_t* __restrict__ dst, long long batch_stride,
                                     int row_stride, int B, int heads, int L, int Lsrc, int row0, int f16) {
    const long long total = static_cast<long long>(B) * heads * L * 8;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int u = static_cast<int>(i & 7);
        long long t = i >> 3;
        const int l = static_cast<int>(t % L); t /= L;
        const int h = static_cast<int>(t % heads);
        const int b = static_cast<int>(t / heads);
        const float4* p = reinterpret_cast<const float4*>(src + ((static_cast<long long>(b) * heads + h) * Lsrc + row0 + l) * D + u * 8);
        const float4 a = __ldg(p), c = __ldg(p + 1);
        uint4 o;
        o.x = pack_h2(a.x, a.y, f16); o.y = pack_h2(a.z, a.w, f16); o.z = pack_h2(c.x, c.y, f16); o.w = pack_h2(c.z, c.w, f16);
        *reinterpret_cast<uint4*>(dst + b * batch_stride + static_cast<long long>(l) * row_stride + h * D + u * 8) = o;
    }
}

inline int grid1d(long long total, int threads) {
    long long b = (total + threads - 1) / threads;
    const long long cap = static_cast<long long>(sm_count()) * 32;
    return static_cast<int>(b < 1 ? 1 : (b > cap ? cap : b));
}
inline size_t al(size_t x) { return (x + 255) / 256 * 256; }

struct BwdWs {
    size_t qh, qhT, kh, khT, vh, doh, dohT, S, ST, dP, dPT, dS, dST, PT, lse, delta, dqh, dkh, dvh, total;
};
BwdWs plan_ws(long long BH, long long Lq, long long Lk) {
    BwdWs w{};
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += al(bytes); return o; };
    w.qh = take(BH * Lq * D * 2); w.qhT = take(BH * Lq * D * 2);
    w.kh = take(BH * Lk * D * 2); w.khT = take(BH * Lk * D * 2);
    w.vh = take(BH * Lk * D * 2);
    w.doh = take(BH * Lq * D * 2); w.dohT = take(BH * Lq * D * 2);
    w.S = take(BH * Lq * Lk * 4); w.ST = take(BH * Lq * Lk * 4);
    w.dP = take(BH * Lq * Lk * 4); w.dPT = take(BH * Lq * Lk * 4);
    w.dS = take(BH * Lq * Lk * 2); w.dST = take(BH * Lq * Lk * 2); w.PT = take(BH * Lq * Lk * 2);
    w.lse = take(BH * Lq * 4); w.delta = take(BH * Lq * 4);
    w.dqh = take(BH * Lq * D * 4); w.dkh = take(BH * Lk * D * 4); w.dvh = take(BH * Lk * D * 4);
    w.total = off;
    return w;
}

}  // namespace
}  // namespace dfw

namespace dfw {

// the round-1 path (logits materialised), kept behind DFW_OPT_ATTN_BWD_UNFUSED for A/B measurements (attn_bwd_fused.cu)
long long attn_bwd_unfused_workspace_bytes(int B, int heads, int Lq, int Ls, int Lb) {
    if (B <= 0 || heads <= 0 || Lq <= 0 || Ls <= 0 || Lb < 0) return -1;
    return static_cast<long long>(dfw::plan_ws(static_cast<long long>(B) * heads, Lq, static_cast<long long>(Ls) + Lb).total);
}

int attn_bwd_unfused(const void* q, long long q_batch_stride, int q_row_stride, const void* k_self,
                         const void* v_self, long long kv_self_batch_stride, int kv_self_row_stride,
                         const void* k_bank, const void* v_bank, long long kv_bank_batch_stride, int kv_bank_row_stride,
                         const void* o, const void* d_o, long long o_batch_stride, int o_row_stride, void* dq, void* dk_self,
                         void* dv_self, void* dk_bank, void* dv_bank, int B, int heads, int Lq, int Ls, int Lb, float scale,
                         int f16, void* workspace, void* stream_) {
    int rc = require_sm100();
    if (rc != DFW_OK) return rc;
    DFW_REQUIRE(q && k_self && v_self && o && d_o && dq && dk_self && dv_self && workspace);
    DFW_REQUIRE(B > 0 && heads > 0 && Lq > 0 && Ls > 0 && Lb >= 0);
    DFW_REQUIRE(Lb == 0 || (k_bank && v_bank && dk_bank && dv_bank));
    DFW_REQUIRE(Lq % 64 == 0 && Ls % 64 == 0 && Lb % 64 == 0);          // every L is a GEMM contraction length here
    DFW_REQUIRE(q_row_stride % 8 == 0 && kv_self_row_stride % 8 == 0 && o_row_stride % 8 == 0 && kv_bank_row_stride % 8 == 0);
    DFW_REQUIRE(q_batch_stride % 8 == 0 && kv_self_batch_stride % 8 == 0 && o_batch_stride % 8 == 0 && kv_bank_batch_stride % 8 == 0);
    cudaStream_t st = static_cast<cudaStream_t>(stream_);
    const int BH = B * heads, Lk = Ls + Lb;
    DFW_REQUIRE(static_cast<long long>(BH) <= 65535);
    const BwdWs w = plan_ws(BH, Lq, Lk);
    uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
    DFW_REQUIRE((reinterpret_cast<uintptr_t>(ws) & 255) == 0);
    auto h16 = [&](size_t off) { return reinterpret_cast<uint16_t*>(ws + off); };
    auto f32 = [&](size_t off) { return reinterpret_cast<float*>(ws + off); };
    const uint16_t* q16 = reinterpret_cast<const uint16_t*>(q);
    const int T = 256;
    // ---- head-major operands (and their transposes where they serve as the contraction-over-L weight operand) ----
    gather_heads_kernel<<<grid1d(static_cast<long long>(BH) * Lq * 8, T), T, 0, st>>>(q16, q_batch_stride, q_row_stride,
                                                                                   h16(w.qh), h16(w.qhT), B, heads, Lq, Lq, 0);
    gather_heads_kernel<<<grid1d(static_cast<long long>(BH) * Ls * 8, T), T, 0, st>>>(
        reinterpret_cast<const uint16_t*>(k_self), kv_self_batch_stride, kv_self_row_stride, h16(w.kh), h16(w.khT), B, heads, Ls, Lk, 0);
    gather_heads_kernel<<<grid1d(static_cast<long long>(BH) * Ls * 8, T), T, 0, st>>>(
        reinterpret_cast<const uint16_t*>(v_self), kv_self_batch_stride, kv_self_row_stride, h16(w.vh), nullptr, B, heads, Ls, Lk, 0);
    if (Lb > 0) {
        gather_heads_kernel<<<grid1d(static_cast<long long>(BH) * Lb * 8, T), T, 0, st>>>(
            reinterpret_cast<const uint16_t*>(k_bank), kv_bank_batch_stride, kv_bank_row_stride, h16(w.kh), h16(w.khT), B, heads, Lb, Lk, Ls);
        gather_heads_kernel<<<grid1d(static_cast<long long>(BH) * Lb * 8, T), T, 0, st>>>(
            reinterpret_cast<const uint16_t*>(v_bank), kv_bank_batch_stride, kv_bank_row_stride, h16(w.vh), nullptr, B, heads, Lb, Lk, Ls);
    }
    gather_heads_kernel<<<grid1d(static_cast<long long>(BH) * Lq * 8, T), T, 0, st>>>(
        reinterpret_cast<const uint16_t*>(d_o), o_batch_stride, o_row_stride, h16(w.doh), h16(w.dohT), B, heads, Lq, Lq, 0);
    {
        const long long total = static_cast<long long>(BH) * Lq * 8;
        delta_kernel<<<static_cast<unsigned>((total + T - 1) / T), T, 0, st>>>(
            reinterpret_cast<const uint16_t*>(o), o_batch_stride, o_row_stride, reinterpret_cast<const uint16_t*>(d_o),
            o_batch_stride, o_row_stride, f32(w.delta), B, heads, Lq, f16);
    }
    g_launches.fetch_add(Lb > 0 ? 7 : 5);
    DFW_CHECK_CUDA(cudaGetLastError());
    const int F = (f16 ? DFW_EPI_F16 : 0), F32 = F | DFW_EPI_OUT_F32;
    // ---- logits in both orientations, dP in both orientations (contraction over d = 64) ----
    rc = dfw_bmm_nt(h16(w.qh), h16(w.kh), D, static_cast<long long>(Lk) * D, nullptr, f32(w.S), BH, Lq, D, Lk, F32, scale, st);
    if (rc != DFW_OK) return rc;
    rc = dfw_bmm_nt(h16(w.kh), h16(w.qh), D, static_cast<long long>(Lq) * D, nullptr, f32(w.ST), BH, Lk, D, Lq, F32, scale, st);
    if (rc != DFW_OK) return rc;
    rc = dfw_bmm_nt(h16(w.doh), h16(w.vh), D, static_cast<long long>(Lk) * D, nullptr, f32(w.dP), BH, Lq, D, Lk, F32, 1.0f, st);
    if (rc != DFW_OK) return rc;
    rc = dfw_bmm_nt(h16(w.vh), h16(w.doh), D, static_cast<long long>(Lq) * D, nullptr, f32(w.dPT), BH, Lk, D, Lq, F32, 1.0f, st);
    if (rc != DFW_OK) return rc;
    row_lse_kernel<<<static_cast<unsigned>(static_cast<long long>(BH) * Lq), 256, 0, st>>>(f32(w.S), f32(w.lse), Lk);
    {
        const long long total4 = static_cast<long long>(BH) * Lq * Lk / 4;
        ds_kernel<<<grid1d(total4, T), T, 0, st>>>(f32(w.S), f32(w.dP), f32(w.lse), f32(w.delta), nullptr, h16(w.dS), total4,
                                                   Lq, Lk, Lq, 1, f16);
        ds_kernel<<<grid1d(total4, T), T, 0, st>>>(f32(w.ST), f32(w.dPT), f32(w.lse), f32(w.delta), h16(w.PT), h16(w.dST),
                                                   total4, Lk, Lq, Lq, 0, f16);
    }
    g_launches.fetch_add(3);
    DFW_CHECK_CUDA(cudaGetLastError());
    // ---- gradients (contraction over L): dQ = scale dS K, dK = scale dS^T Q, dV = P^T dO ----
    rc = dfw_bmm_nt(h16(w.dS), h16(w.khT), Lk, static_cast<long long>(D) * Lk, nullptr, f32(w.dqh), BH, Lq, Lk, D, F32, scale, st);
    if (rc != DFW_OK) return rc;
    rc = dfw_bmm_nt(h16(w.dST), h16(w.qhT), Lq, static_cast<long long>(D) * Lq, nullptr, f32(w.dkh), BH, Lk, Lq, D, F32, scale, st);
    if (rc != DFW_OK) return rc;
    rc = dfw_bmm_nt(h16(w.PT), h16(w.dohT), Lq, static_cast<long long>(D) * Lq, nullptr, f32(w.dvh), BH, Lk, Lq, D, F32, 1.0f, st);
    if (rc != DFW_OK) return rc;
    // ---- back to the callers' layouts ----
    scatter_heads_kernel<<<grid1d(static_cast<long long>(BH) * Lq * 8, T), T, 0, st>>>(
        f32(w.dqh), reinterpret_cast<uint16_t*>(dq), q_batch_stride, q_row_stride, B, heads, Lq, Lq, 0, f16);
    scatter_heads_kernel<<<grid1d(static_cast<long long>(BH) * Ls * 8, T), T, 0, st>>>(
        f32(w.dkh), reinterpret_cast<uint16_t*>(dk_self), kv_self_batch_stride, kv_self_row_stride, B, heads, Ls, Lk, 0, f16);
    scatter_heads_kernel<<<grid1d(static_cast<long long>(BH) * Ls * 8, T), T, 0, st>>>(
        f32(w.dvh), reinterpret_cast<uint16_t*>(dv_self), kv_self_batch_stride, kv_self_row_stride, B, heads, Ls, Lk, 0, f16);
    if (Lb > 0) {
        scatter_heads_kernel<<<grid1d(static_cast<long long>(BH) * Lb * 8, T), T, 0, st>>>(
            f32(w.dkh), reinterpret_cast<uint16_t*>(dk_bank), kv_bank_batch_stride, kv_bank_row_stride, B, heads, Lb, Lk, Ls, f16);
        scatter_heads_kernel<<<grid1d(static_cast<long long>(BH) * Lb * 8, T), T, 0, st>>>(
            f32(w.dvh), reinterpret_cast<uint16_t*>(dv_bank), kv_bank_batch_stride, kv_bank_row_stride, B, heads, Lb, Lk, Ls, f16);
    }
    g_launches.fetch_add(Lb > 0 ? 5 : 3);
    DFW_CHECK_CUDA(cudaGetLastError());
    return DFW_OK;
}

}  // namespace dfw
