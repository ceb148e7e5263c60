// K1 — KV-fused flash attention forward for sm_100a (head_dim 64), and K2 — short-context cross attention.
//
// ref: diffews/models/attention_processor.py:251-271 (MyXFormersAttnProcessor): on the query pass
//        key   = cat([key_self,   fold(k_bank)], dim=1)      fold = shot-major concatenation of the k supports
//        value = cat([value_self, fold(v_bank)], dim=1)
//        out   = xformers.ops.memory_efficient_attention(q, key, value, scale=attn.scale)
//      (:351-365 is the SDPA variant, :153-164 the unfused bmm variant — same maths).
// Here the concatenation is never materialised: the K/V ring is fed from two TMA tensor maps (self, then bank).
//
// CTA = 128 query rows of one (episode, head).  Warp roles:
//   warp 0     TMA producer: Q once, then (K_j, V_j) 128-key tiles through a 3-stage ring
//   warp 1     MMA issuer:   S_j = Q K_j^T (tcgen05, 128x128x64, fp32 in TMEM, double-buffered)
//                            O_j = P_j V_j  (128x64x128, P from smem K-major, V MN-major straight from TMA)
//   warp 2     TMEM allocator
//   warps 4-7  softmax: one thread per query row (= TMEM lane): running max / sum in the log2 domain, P_j -> bf16 ->
//              swizzled smem, O accumulated in registers with the per-tile rescale (no TMEM read-modify-write).
#include <atomic>

#include "common.cuh"
#include "ptx.cuh"

namespace dfw {
extern std::atomic<long long> g_launches;
namespace {

constexpr int ATT_M = 128;       // query rows per CTA
constexpr int ATT_N = 128;       // keys per tile
constexpr int ATT_D = 64;        // head dim
constexpr int KV_STAGES = 3;
constexpr int TILE_BYTES = 128 * 128;             // 128 rows x 128 B
constexpr int ATT_THREADS = 256;
constexpr int ATT_SMEM = TILE_BYTES /*Q*/ + KV_STAGES * 2 * TILE_BYTES /*K,V*/ + 2 * 2 * TILE_BYTES /*P x2*/ +
                         1024 + 256;
constexpr int ATT_TMEM_COLS = 512;
// All 16-bit tensors of a call (Q, K, V, P, O) share one format, bf16 or fp16: tcgen05 kind::f16 requires the A and B
// operand of an MMA to have the same format (a bf16 x fp16 mix raises an illegal-instruction trap on sm_100).

struct AttnMaps {
    CUtensorMap q, k_self, v_self, k_bank, v_bank;
};
struct AttnParams {
    int Lq, Ls, Lb;
    int n_self, n_bank;   // number of 128-key tiles per source
    float scale_log2;     // scale * log2(e)
    uint16_t* o;
    long long o_batch_stride;
    int o_row_stride;
    int f16;              // 16-bit tensors are fp16 (else bf16)
};

__global__ void __launch_bounds__(ATT_THREADS, 1)
attn_kvfused_kernel(const __grid_constant__ AttnMaps maps, const __grid_constant__ AttnParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_u32 = smem_u32(smem_raw);
    const uint32_t base = (raw_u32 + 1023u) & ~1023u;
    const uint32_t sQ = base;
    auto sK = [&](int s) { return base + TILE_BYTES + s * 2 * TILE_BYTES; };
    auto sV = [&](int s) { return base + TILE_BYTES + s * 2 * TILE_BYTES + TILE_BYTES; };
    auto sP = [&](int b) { return base + TILE_BYTES + KV_STAGES * 2 * TILE_BYTES + b * 2 * TILE_BYTES; };
    const uint32_t bar_base = base + TILE_BYTES + KV_STAGES * 2 * TILE_BYTES + 4 * TILE_BYTES;
    const uint32_t q_full = bar_base;
    auto kv_full = [&](int s) { return bar_base + 8u * (1 + s); };
    auto kv_empty = [&](int s) { return bar_base + 8u * (1 + KV_STAGES + s); };
    auto s_full = [&](int b) { return bar_base + 8u * (1 + 2 * KV_STAGES + b); };
    auto p_full = [&](int b) { return bar_base + 8u * (3 + 2 * KV_STAGES + b); };
    auto pv_done = [&](int b) { return bar_base + 8u * (5 + 2 * KV_STAGES + b); };
    const uint32_t tmem_slot = bar_base + 8u * (7 + 2 * KV_STAGES);
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw_u32));
    uint8_t* sP_generic = smem_raw + (sP(0) - raw_u32);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * ATT_M;
    const int head = blockIdx.y;
    const int b = blockIdx.z;
    const int ntiles = p.n_self + p.n_bank;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&maps.q);
        tma_prefetch_desc(&maps.k_self);
        tma_prefetch_desc(&maps.v_self);
        if (p.n_bank) { tma_prefetch_desc(&maps.k_bank); tma_prefetch_desc(&maps.v_bank); }
    }
    if (warp == 1 && lane == 0) {
        mbar_init(q_full, 1);
        for (int s = 0; s < KV_STAGES; ++s) { mbar_init(kv_full(s), 1); mbar_init(kv_empty(s), 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(s_full(i), 1); mbar_init(p_full(i), 128); mbar_init(pv_done(i), 1); }
        fence_mbar_init();
    }
    if (warp == 2) {
        tmem_alloc(tmem_slot, ATT_TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    auto tS = [&](int i) { return tmem_base + i * 128; };
    auto tO = [&](int i) { return tmem_base + 256 + i * 64; };

    if (warp == 0) {
        if (lane == 0) {
            mbar_arrive_expect_tx(q_full, TILE_BYTES);
            tma_load_3d(sQ, &maps.q, q_full, head * ATT_D, q0, b);
            for (int j = 0; j < ntiles; ++j) {
                const int s = j % KV_STAGES;
                const uint32_t ph = (j / KV_STAGES) & 1;
                mbar_wait(kv_empty(s), ph ^ 1u, 10);
                mbar_arrive_expect_tx(kv_full(s), 2 * TILE_BYTES);
                if (j < p.n_self) {
                    tma_load_3d(sK(s), &maps.k_self, kv_full(s), head * ATT_D, j * ATT_N, b);
                    tma_load_3d(sV(s), &maps.v_self, kv_full(s), head * ATT_D, j * ATT_N, b);
                } else {
                    const int jb = j - p.n_self;
                    tma_load_3d(sK(s), &maps.k_bank, kv_full(s), head * ATT_D, jb * ATT_N, b);
                    tma_load_3d(sV(s), &maps.v_bank, kv_full(s), head * ATT_D, jb * ATT_N, b);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t fmt = p.f16 ? 0u : 1u;
            const uint32_t idesc_s = umma_idesc(ATT_M, ATT_N, fmt, fmt, 0);   // S = Q K^T : B (=K) is K-major
            const uint32_t idesc_o = umma_idesc(ATT_M, ATT_D, fmt, fmt, 1);   // O = P V   : B (=V) is MN-major
            auto issue_s = [&](int j) {
                const int s = j % KV_STAGES;
                mbar_wait(kv_full(s), (j / KV_STAGES) & 1, 11);
                tc_fence_after();
                const uint64_t adesc = umma_desc_sw128(sQ);
                const uint64_t bdesc = umma_desc_sw128(sK(s));
#pragma unroll
                for (int k = 0; k < ATT_D / 16; ++k)
                    umma_ss(tS(j & 1), adesc + 2u * k, bdesc + 2u * k, idesc_s, k > 0 ? 1u : 0u);
                tc_commit(s_full(j & 1));
            };
            mbar_wait(q_full, 0, 12);
            issue_s(0);
            for (int j = 0; j < ntiles; ++j) {
                if (j + 1 < ntiles) issue_s(j + 1);
                mbar_wait(p_full(j & 1), (j >> 1) & 1, 13);
                tc_fence_after();
                const int s = j % KV_STAGES;
#pragma unroll
                for (int ks = 0; ks < ATT_N / 16; ++ks) {
                    const uint64_t adesc = umma_desc_sw128(sP(j & 1) + (ks >> 2) * TILE_BYTES) + 2u * (ks & 3);
                    const uint64_t bdesc = umma_desc_sw128(sV(s) + ks * 16 * 128);
                    umma_ss(tO(j & 1), adesc, bdesc, idesc_o, ks > 0 ? 1u : 0u);
                }
                tc_commit(kv_empty(s));
                tc_commit(pv_done(j & 1));
            }
        }
    } else if (warp >= 4) {
        const int qd = warp - 4;
        const int row = qd * 32 + lane;
        const uint32_t lane_off = static_cast<uint32_t>(qd * 32) << 16;
        float m_run = -INFINITY, l_run = 0.f, alpha_pending = 0.f;
        float o_acc[ATT_D];
#pragma unroll
        for (int i = 0; i < ATT_D; ++i) o_acc[i] = 0.f;

        auto accumulate_o = [&](int j, float alpha) {
            mbar_wait(pv_done(j & 1), (j >> 1) & 1, 14);
            tc_fence_after();
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                uint32_t v[32];
                tmem_ld_32x32(tO(j & 1) + lane_off + c * 32, v);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) o_acc[c * 32 + i] = fmaf(o_acc[c * 32 + i], alpha, __uint_as_float(v[i]));
            }
        };

        for (int j = 0; j < ntiles; ++j) {
            int valid;
            if (j < p.n_self) valid = min(ATT_N, p.Ls - j * ATT_N);
            else valid = min(ATT_N, p.Lb - (j - p.n_self) * ATT_N);
            mbar_wait(s_full(j & 1), (j >> 1) & 1, 15);
            tc_fence_after();
            const uint32_t ts = tS(j & 1) + lane_off;
            // pass 1: row max
            float mx = -INFINITY;
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
                uint32_t v[32];
                tmem_ld_32x32(ts + c * 32, v);
                tmem_ld_wait();
                if (valid == ATT_N) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(v[i]));
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        if (c * 32 + i < valid) mx = fmaxf(mx, __uint_as_float(v[i]));
                }
            }
            const float m_new = fmaxf(m_run, mx * p.scale_log2);
            const float alpha = exp2f(m_run - m_new);   // first tile: exp2(-inf) = 0
            float psum = 0.f;
            // pass 2: p = exp2(s*c - m), bf16, swizzled store into the K-major P tile
            uint8_t* pbuf = sP_generic + (j & 1) * 2 * TILE_BYTES;
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
                uint32_t v[32];
                tmem_ld_32x32(ts + c * 32, v);
                tmem_ld_wait();
                float pf[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    float e = exp2f(fmaf(__uint_as_float(v[i]), p.scale_log2, -m_new));
                    if (valid != ATT_N && c * 32 + i >= valid) e = 0.f;
                    pf[i] = e;
                    psum += e;
                }
                uint8_t* chunk = pbuf + (c >> 1) * TILE_BYTES + row * 128;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    uint4 w;
                    w.x = pack_h2(pf[u * 8 + 0], pf[u * 8 + 1], p.f16);
                    w.y = pack_h2(pf[u * 8 + 2], pf[u * 8 + 3], p.f16);
                    w.z = pack_h2(pf[u * 8 + 4], pf[u * 8 + 5], p.f16);
                    w.w = pack_h2(pf[u * 8 + 6], pf[u * 8 + 7], p.f16);
                    const int unit = ((c & 1) * 4 + u) ^ (row & 7);
                    *reinterpret_cast<uint4*>(chunk + unit * 16) = w;
                }
            }
            l_run = fmaf(l_run, alpha, psum);
            m_run = m_new;
            fence_proxy_async_smem();
            tc_fence_before();
            mbar_arrive(p_full(j & 1));
            if (j >= 1) accumulate_o(j - 1, alpha_pending);
            alpha_pending = alpha;
        }
        accumulate_o(ntiles - 1, alpha_pending);
        if (q0 + row < p.Lq) {
            const float inv = 1.0f / l_run;
            uint16_t* op = p.o + static_cast<long long>(b) * p.o_batch_stride +
                                static_cast<long long>(q0 + row) * p.o_row_stride + head * ATT_D;
#pragma unroll
            for (int i = 0; i < ATT_D; i += 8) {
                uint4 w;
                w.x = pack_h2(o_acc[i] * inv, o_acc[i + 1] * inv, p.f16);
                w.y = pack_h2(o_acc[i + 2] * inv, o_acc[i + 3] * inv, p.f16);
                w.z = pack_h2(o_acc[i + 4] * inv, o_acc[i + 5] * inv, p.f16);
                w.w = pack_h2(o_acc[i + 6] * inv, o_acc[i + 7] * inv, p.f16);
                *reinterpret_cast<uint4*>(op + i) = w;
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, ATT_TMEM_COLS);
    }
}

int make_seq_map(CUtensorMap* m, const void* base, int C, int L, int B, int row_stride, long long batch_stride) {
    const uint64_t dims[3] = {static_cast<uint64_t>(C), static_cast<uint64_t>(L), static_cast<uint64_t>(B)};
    const uint64_t strides[2] = {static_cast<uint64_t>(row_stride) * 2, static_cast<uint64_t>(batch_stride) * 2};
    const uint32_t box[3] = {ATT_D, ATT_N, 1};
    return encode_tmap_bf16_sw128(m, base, 3, dims, strides, box);
}

// ---------------------------------------------------------------------------------------------------------
// K2: cross attention to Lctx <= 128 prompt tokens. 8 lanes per (token, head); online softmax over the keys.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void unpack8(const uint4& u, int f16, float (&f)[8]) {
    const float2 a = unpack_h2(u.x, f16), b = unpack_h2(u.y, f16), c = unpack_h2(u.z, f16), d = unpack_h2(u.w, f16);
    f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}

__global__ void cross_attn_kernel(const uint16_t* __restrict__ q, const uint16_t* __restrict__ k,
                                  const uint16_t* __restrict__ v, long long kv_batch_stride,
                                  uint16_t* __restrict__ o, long long total_groups, int L, int heads, int Lctx,
                                  float scale_log2, int f16) {
    const long long gid = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 3;
    const int sub = threadIdx.x & 7;
    const bool active = gid < total_groups;
    const long long g = active ? gid : 0;
    const int h = static_cast<int>(g % heads);
    const long long tok = g / heads;            // b*L + l
    const long long bidx = tok / L;
    const int C = heads * 64;
    const long long qoff = tok * C + h * 64 + sub * 8;
    const uint4 qv = __ldg(reinterpret_cast<const uint4*>(q + qoff));
    float qf[8];
    unpack8(qv, f16, qf);
    float m = -INFINITY, l = 0.f;
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const uint16_t* kb = k + bidx * kv_batch_stride + h * 64 + sub * 8;
    const uint16_t* vb = v + bidx * kv_batch_stride + h * 64 + sub * 8;
    for (int j = 0; j < Lctx; ++j) {
        const uint4 kv = __ldg(reinterpret_cast<const uint4*>(kb + static_cast<long long>(j) * C));
        float kf[8];
        unpack8(kv, f16, kf);
        float d = qf[0] * kf[0] + qf[1] * kf[1] + qf[2] * kf[2] + qf[3] * kf[3] + qf[4] * kf[4] + qf[5] * kf[5] +
                  qf[6] * kf[6] + qf[7] * kf[7];
        d += __shfl_xor_sync(0xffffffffu, d, 1);
        d += __shfl_xor_sync(0xffffffffu, d, 2);
        d += __shfl_xor_sync(0xffffffffu, d, 4);
        const float s = d * scale_log2;
        const float m_new = fmaxf(m, s);
        const float alpha = exp2f(m - m_new);
        const float pj = exp2f(s - m_new);
        l = fmaf(l, alpha, pj);
        const uint4 vv = __ldg(reinterpret_cast<const uint4*>(vb + static_cast<long long>(j) * C));
        float vf[8];
        unpack8(vv, f16, vf);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = fmaf(acc[i], alpha, pj * vf[i]);
        m = m_new;
    }
    if (active) {
        const float inv = 1.0f / l;
        uint4 w;
        w.x = pack_h2(acc[0] * inv, acc[1] * inv, f16);
        w.y = pack_h2(acc[2] * inv, acc[3] * inv, f16);
        w.z = pack_h2(acc[4] * inv, acc[5] * inv, f16);
        w.w = pack_h2(acc[6] * inv, acc[7] * inv, f16);
        *reinterpret_cast<uint4*>(o + qoff) = w;
    }
}

}  // namespace
}  // namespace dfw

extern "C" {

int dfw_attn_kvfused_fwd(const void* q, long long q_batch_stride, int q_row_stride, const void* k_self,
                         const void* v_self, long long kv_self_batch_stride, int kv_self_row_stride,
                         const void* k_bank, const void* v_bank, long long kv_bank_batch_stride,
                         int kv_bank_row_stride, void* o, long long o_batch_stride, int o_row_stride, int B,
                         int heads, int Lq, int Ls, int Lb, float scale, int f16, void* stream_) {
    using namespace dfw;
    int rc = require_sm100();
    if (rc != DFW_OK) return rc;
    DFW_REQUIRE(q && k_self && v_self && o);
    DFW_REQUIRE(B > 0 && heads > 0 && Lq > 0 && Ls > 0 && Lb >= 0);
    DFW_REQUIRE(Lb == 0 || (k_bank && v_bank));
    DFW_REQUIRE(q_row_stride % 8 == 0 && kv_self_row_stride % 8 == 0 && o_row_stride % 8 == 0);
    DFW_REQUIRE(q_batch_stride % 8 == 0 && kv_self_batch_stride % 8 == 0 && o_batch_stride % 8 == 0);
    DFW_REQUIRE(B <= 65535 && heads <= 65535);
    const int C = heads * ATT_D;
    AttnMaps maps;
    rc = make_seq_map(&maps.q, q, C, Lq, B, q_row_stride, q_batch_stride);
    if (rc != DFW_OK) return rc;
    rc = make_seq_map(&maps.k_self, k_self, C, Ls, B, kv_self_row_stride, kv_self_batch_stride);
    if (rc != DFW_OK) return rc;
    rc = make_seq_map(&maps.v_self, v_self, C, Ls, B, kv_self_row_stride, kv_self_batch_stride);
    if (rc != DFW_OK) return rc;
    if (Lb > 0) {
        DFW_REQUIRE(kv_bank_row_stride % 8 == 0 && kv_bank_batch_stride % 8 == 0);
        rc = make_seq_map(&maps.k_bank, k_bank, C, Lb, B, kv_bank_row_stride, kv_bank_batch_stride);
        if (rc != DFW_OK) return rc;
        rc = make_seq_map(&maps.v_bank, v_bank, C, Lb, B, kv_bank_row_stride, kv_bank_batch_stride);
        if (rc != DFW_OK) return rc;
    } else {
        maps.k_bank = maps.k_self;
        maps.v_bank = maps.v_self;
    }
    AttnParams p{};
    p.Lq = Lq; p.Ls = Ls; p.Lb = Lb;
    p.n_self = (Ls + ATT_N - 1) / ATT_N;
    p.n_bank = (Lb + ATT_N - 1) / ATT_N;
    p.scale_log2 = scale * 1.4426950408889634f;
    p.o = reinterpret_cast<uint16_t*>(o);
    p.f16 = f16;
    p.o_batch_stride = o_batch_stride;
    p.o_row_stride = o_row_stride;
    static bool attr_set = false;
    if (!attr_set) {
        DFW_CHECK_CUDA(cudaFuncSetAttribute(attn_kvfused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM));
        attr_set = true;
    }
    dim3 grid((Lq + ATT_M - 1) / ATT_M, heads, B);
    attn_kvfused_kernel<<<grid, ATT_THREADS, ATT_SMEM, static_cast<cudaStream_t>(stream_)>>>(maps, p);
    g_launches.fetch_add(1);
    DFW_CHECK_CUDA(cudaGetLastError());
    return DFW_OK;
}

int dfw_cross_attn_fwd(const void* q, const void* k, const void* v, long long kv_batch_stride, void* o, int B,
                       int L, int heads, int Lctx, float scale, int f16, void* stream_) {
    using namespace dfw;
    int rc = require_sm100();
    if (rc != DFW_OK) return rc;
    DFW_REQUIRE(q && k && v && o && B > 0 && L > 0 && heads > 0 && Lctx > 0 && Lctx <= 128);
    DFW_REQUIRE(kv_batch_stride % 8 == 0);
    const long long groups = static_cast<long long>(B) * L * heads;
    const long long threads = groups * 8;
    const long long blocks = (threads + 255) / 256;
    DFW_REQUIRE(blocks < (1LL << 31));
    cross_attn_kernel<<<static_cast<unsigned int>(blocks), 256, 0, static_cast<cudaStream_t>(stream_)>>>(
        reinterpret_cast<const uint16_t*>(q), reinterpret_cast<const uint16_t*>(k),
        reinterpret_cast<const uint16_t*>(v), kv_batch_stride, reinterpret_cast<uint16_t*>(o), groups, L,
        heads, Lctx, scale * 1.4426950408889634f, f16);
    g_launches.fetch_add(1);
    DFW_CHECK_CUDA(cudaGetLastError());
    return DFW_OK;
}

}  // extern "C"
