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
#include <cstdlib>

#include <type_traits>

#include "common.cuh"
#include "ptx.cuh"
#include "attn_softmax.cuh"

namespace dfw {
extern std::atomic<long long> g_launches;
#ifdef DFW_ATTN_TRACE
extern long long* g_attn_trace;
#endif
namespace {

constexpr int ATT_M = 128;       // query rows per tile (one tcgen05 M)
constexpr int ATT_QT = 2;        // query tiles per CTA (ping-pong: softmax of one overlaps the MMAs of the other)
constexpr int ATT_N = 128;       // keys per tile
constexpr int ATT_D = 64;        // head dim
constexpr int KV_STAGES = 3;
constexpr int TILE_BYTES = 128 * 128;             // 128 rows x 128 B
constexpr int ATT_THREADS = 384;                  // 4 control warps + 2 x 4 softmax warps
constexpr int ATT_SMEM = ATT_QT * TILE_BYTES /*Q*/ + KV_STAGES * 2 * TILE_BYTES /*K,V*/ +
                         ATT_QT * 2 * TILE_BYTES /*P*/ + 1024 + 256;
constexpr int ATT_SBUF = 3;                       // S accumulators rotate over 3 TMEM buffers: the MMA warp computes
                                                  // S_X(j+1) while the softmax group of X still reads S_X(j)
constexpr int ATT_TMEM_COLS = 512;                // S buffers [0,128) [128,256) [256,384); O_A [384,448) O_B [448,512)
constexpr float ATT_RESCALE_TAU = 8.0f;           // lazy rescale threshold in the log2 domain (P <= 2^8)
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
    float* lse;           // optional [B, heads, Lq]: log2-domain logsumexp of the scaled logits (for the backward); v3 only
#ifdef DFW_ATTN_TRACE
    long long* trace;     // debug build only (scripts/attn_trace.py): clock64 stamps of CTA (0,0,0), [warp][tile][8]
#endif
};
#ifdef DFW_ATTN_TRACE
#define ATT_STAMP(slot) do { if (p.trace && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && lane == 0 && j < 64) \
        p.trace[(warp * 64 + j) * 8 + (slot)] = clock64(); } while (0)
#else
#define ATT_STAMP(slot) do { } while (0)
#endif

// CTA = 256 query rows (two 128-row tiles A, B) of one (episode, head).
//   warp 0      TMA producer: Q_A, Q_B once; (K_j, V_j) 128-key tiles through a 3-stage ring, from two tensor maps
//               (self keys first, then the support bank) — the concatenation is never materialised
//   warp 1      MMA issuer: S_X(j) = Q_X K_j^T into TMEM (128x128x64), O_X += P_X(j) V_j accumulated IN TMEM
//               (128x64x128, V consumed MN-major straight from the TMA tile); schedule A,B interleaved
//   warp 2      TMEM allocator
//   warps 4-7   softmax of tile A, warps 8-11 softmax of tile B: one thread per query row (= TMEM lane); running max
//               and sum in the log2 domain; O is only rescaled (tcgen05.ld -> scale -> tcgen05.st) when some row of
//               the warp raises its max by more than 2^TAU, so the common case never touches O; P -> 16-bit ->
//               128B-swizzled smem (the K-major A operand of the PV MMA)
template <bool F16>
__global__ void __launch_bounds__(ATT_THREADS, 1)
attn_kvfused_v2_kernel(const __grid_constant__ AttnMaps maps, const __grid_constant__ AttnParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_u32 = smem_u32(smem_raw);
    const uint32_t base = (raw_u32 + 1023u) & ~1023u;
    constexpr bool MI = false;      // (the three-issuer schedule of round 1 was slower and is gone)
    constexpr int KVS = KV_STAGES;
    constexpr int NPB = 1;
    auto sQ = [&](int x) { return base + x * TILE_BYTES; };
    const uint32_t kv_base = base + ATT_QT * TILE_BYTES;
    auto sK = [&](int s) { return kv_base + s * 2 * TILE_BYTES; };
    auto sV = [&](int s) { return kv_base + s * 2 * TILE_BYTES + TILE_BYTES; };
    const uint32_t p_base = kv_base + KVS * 2 * TILE_BYTES;
    auto sP = [&](int x, int pb) { return p_base + (x * NPB + pb) * 2 * TILE_BYTES; };
    const uint32_t bar_base = p_base + ATT_QT * NPB * 2 * TILE_BYTES;
    const uint32_t q_full = bar_base;
    auto kv_full = [&](int s) { return bar_base + 8u * (1 + s); };
    auto kv_empty = [&](int s) { return bar_base + 8u * (1 + KVS + s); };
    auto s_full = [&](int buf) { return bar_base + 8u * (1 + 2 * KVS + buf); };      // one per S buffer
    auto p_full = [&](int x) { return bar_base + 8u * (4 + 2 * KVS + x); };
    auto pv_done = [&](int x, int pb) { return bar_base + 8u * (6 + 2 * KVS + x * 2 + pb); };   // one per P buffer
    const uint32_t tmem_slot = bar_base + 8u * (10 + 2 * KVS);
    auto s_free = [&](int buf) { return bar_base + 8u * (11 + 2 * KVS + buf); };     // MI only: S buffer read out
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw_u32));
    uint8_t* sP_generic = smem_raw + (p_base - raw_u32);

    // warp index via shuffle = provably warp-uniform; role loops run on all lanes, TMA / tcgen05 issue sits in
    // elect_one() regions (see igemm.cu: `if (lane == 0)` around the loop costs an ELECT/BRA.U.ANY loop per instruction)
    const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * ATT_M * ATT_QT;
    const int head = blockIdx.y;
    const int b = blockIdx.z;
    const int ntiles = p.n_self + p.n_bank;
    const bool has_b = (q0 + ATT_M) < p.Lq;          // second query tile holds at least one valid row

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&maps.q);
        tma_prefetch_desc(&maps.k_self);
        tma_prefetch_desc(&maps.v_self);
        if (p.n_bank) { tma_prefetch_desc(&maps.k_bank); tma_prefetch_desc(&maps.v_bank); }
    }
    if (warp == 1 && lane == 0) {
        mbar_init(q_full, 1);
        for (int s = 0; s < KVS; ++s) { mbar_init(kv_full(s), 1); mbar_init(kv_empty(s), (MI && has_b) ? 2 : 1); }
        for (int x = 0; x < ATT_SBUF; ++x) { mbar_init(s_full(x), 1); mbar_init(s_free(x), 128); }
        for (int x = 0; x < ATT_QT; ++x) {
            mbar_init(p_full(x), 128);
            for (int pb = 0; pb < NPB; ++pb) mbar_init(pv_done(x, pb), 1);
        }
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
    pdl_wait();
    auto tS = [&](int buf) { return tmem_base + buf * 128; };
    auto tO = [&](int x) { return tmem_base + ATT_SBUF * 128 + x * 64; };

    if (warp == 0) {
        {
            if (elect_one()) {
                mbar_arrive_expect_tx(q_full, (has_b ? 2 : 1) * TILE_BYTES);
                tma_load_3d(sQ(0), &maps.q, q_full, head * ATT_D, q0, b);
                if (has_b) tma_load_3d(sQ(1), &maps.q, q_full, head * ATT_D, q0 + ATT_M, b);
            }
            __syncwarp();
            for (int j = 0; j < ntiles; ++j) {
                const int s = j % KVS;
                const uint32_t ph = (j / KVS) & 1;
                mbar_wait(kv_empty(s), ph ^ 1u, 10);
                if (elect_one()) {
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
                __syncwarp();
            }
        }
    } else if (warp == 1) {
        {
            const uint32_t fmt = F16 ? 0u : 1u;
            const uint32_t idesc_s = umma_idesc(ATT_M, ATT_N, fmt, fmt, 0);   // S = Q K^T : B (=K) is K-major
            const uint32_t idesc_o = umma_idesc(ATT_M, ATT_D, fmt, fmt, 1);   // O = P V   : B (=V) is MN-major
            const int nq = has_b ? 2 : 1;
            // S tiles are numbered in issue order, seq = nq * j + x, and live in TMEM buffer seq % 3.  S(seq) may be
            // issued once K_j has landed and the softmax of S(seq - 3) has finished (its p_full was waited two PVs ago).
            auto issue_s = [&](int x, int j) {          // called by the elected lane
                const int s = j % KVS;
                const int buf = (nq * j + x) % ATT_SBUF;
                const uint64_t adesc = umma_desc_sw128(sQ(x));
                const uint64_t bdesc = umma_desc_sw128(sK(s));
#pragma unroll
                for (int k = 0; k < ATT_D / 16; ++k)
                    umma_ss(tS(buf), adesc + 2u * k, bdesc + 2u * k, idesc_s, k > 0 ? 1u : 0u);
                tc_commit(s_full(buf));
            };
            mbar_wait(q_full, 0, 12);
            mbar_wait(kv_full(0), 0, 11);
            tc_fence_after();
            if (elect_one()) {
                for (int x = 0; x < nq; ++x) issue_s(x, 0);
            }
            __syncwarp();
            for (int j = 0; j < ntiles; ++j) {
                const int s = j % KVS;
                if (j + 1 < ntiles) {
                    mbar_wait(kv_full((j + 1) % KVS), ((j + 1) / KVS) & 1, 11);
                    tc_fence_after();
                }
                for (int x = 0; x < nq; ++x) {
                    // run ahead: S_X(j+1) goes to the tensor pipe before P_X(j) is waited for (its buffer, last used by
                    // S(seq - 3), was released by a p_full wait of an earlier step)
                    if (j + 1 < ntiles) {
                        if (elect_one()) issue_s(x, j + 1);
                        __syncwarp();
                    }
                    ATT_STAMP(x * 2);
                    mbar_wait(p_full(x), j & 1, 13);                          // P_X(j) in smem, S_X(j) consumed, O_X rescaled
                    tc_fence_after();
                    ATT_STAMP(x * 2 + 1);
                    if (elect_one()) {
#pragma unroll
                        for (int ks = 0; ks < ATT_N / 16; ++ks) {
                            const uint64_t adesc = umma_desc_sw128(sP(x, 0) + (ks >> 2) * TILE_BYTES) + 2u * (ks & 3);
                            const uint64_t bdesc = umma_desc_sw128(sV(s) + ks * 16 * 128);
                            umma_ss(tO(x), adesc, bdesc, idesc_o, (j > 0 || ks > 0) ? 1u : 0u);
                        }
                        tc_commit(pv_done(x, 0));
                        if (x == nq - 1) tc_commit(kv_empty(s));              // K_j / V_j fully consumed by both tiles
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp >= 4) {
        const int x = (warp - 4) >> 2;                // query tile of this softmax group
        const int qd = (warp - 4) & 3;                // TMEM lane quadrant
        const int row = qd * 32 + lane;
        const int qrow0 = q0 + x * ATT_M;
        if (x == 0 || has_b) {
            const uint32_t lane_off = static_cast<uint32_t>(qd * 32) << 16;
            const int nq = has_b ? 2 : 1;
            const uint32_t to = tO(x) + lane_off;
            float m_used = -INFINITY, l_run = 0.f;
            for (int j = 0; j < ntiles; ++j) {
                uint8_t* pbuf = sP_generic + (x * NPB + j % NPB) * 2 * TILE_BYTES;
                int valid;
                if (j < p.n_self) valid = min(ATT_N, p.Ls - j * ATT_N);
                else valid = min(ATT_N, p.Lb - (j - p.n_self) * ATT_N);
                const int seq = nq * j + x, sbuf = seq % ATT_SBUF;
                const uint32_t ts = tS(sbuf) + lane_off;
                ATT_STAMP(0);
                mbar_wait(s_full(sbuf), (seq / ATT_SBUF) & 1, 15);
                tc_fence_after();
                ATT_STAMP(1);
                // The TMEM loads are software-pipelined: the load of chunk c+1 is in flight while chunk c is
                // processed (tcgen05.wait::ld waits for everything outstanding, so it is placed after the compute).
                // The whole tile body is instantiated twice (full tile / ragged last tile) so the common full-tile path
                // carries no per-element masking; exp2 is one MUFU (ex2.approx.ftz), the row max uses 3-input max.
                float psum = 0.f;
                auto run_tile = [&](auto full_tag) {
                    constexpr bool FULL = decltype(full_tag)::value;
                    uint32_t va[32], vb[32];
                    auto chunk_max = [&](const uint32_t (&v)[32], int c, float& mx) {
                        if constexpr (FULL) {
#pragma unroll
                            for (int i = 0; i < 32; i += 2)
                                mx = fmax3(mx, __uint_as_float(v[i]), __uint_as_float(v[i + 1]));
                        } else {
#pragma unroll
                            for (int i = 0; i < 32; ++i)
                                if (c * 32 + i < valid) mx = fmaxf(mx, __uint_as_float(v[i]));
                        }
                    };
                    // pass 1: row max of the scaled logits
                    float mx = -INFINITY;
                    tmem_ld_32x32(ts, va);
                    tmem_ld_wait(); tmem_regs_ready(va);
                    tmem_ld_32x32(ts + 32, vb);
                    chunk_max(va, 0, mx);
                    tmem_ld_wait(); tmem_regs_ready(vb);
                    tmem_ld_32x32(ts + 64, va);
                    chunk_max(vb, 1, mx);
                    tmem_ld_wait(); tmem_regs_ready(va);
                    tmem_ld_32x32(ts + 96, vb);
                    chunk_max(va, 2, mx);
                    tmem_ld_wait(); tmem_regs_ready(vb);
                    tmem_ld_32x32(ts, va);                               // chunk 0 again, for pass 2
                    chunk_max(vb, 3, mx);
                    tmem_ld_wait(); tmem_regs_ready(va);
                    const float m_new = fmaxf(m_used, mx * p.scale_log2);
                    ATT_STAMP(2);
                    // lazy rescale: only when some row of the warp moved its max by more than 2^TAU (warp-uniform
                    // branch, tcgen05.ld/st are warp-collective)
                    if (__any_sync(0xffffffffu, m_new > m_used + ATT_RESCALE_TAU)) {
                        const float alpha = ex2_approx(m_used - m_new);   // 0 on the first tile
                        if (j > 0) {
                            mbar_wait(pv_done(x, (j - 1) % NPB), ((j - 1) / NPB) & 1, 14);       // O_X holds tiles < j
                            tc_fence_after();
#pragma unroll
                            for (int c = 0; c < 2; ++c) {
                                uint32_t v[32];
                                tmem_ld_32x32(to + c * 32, v);
                                tmem_ld_wait(); tmem_regs_ready(v);
#pragma unroll
                                for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
                                tmem_st_32x32(to + c * 32, v);
                            }
                            tmem_st_wait();
                        }
                        l_run *= alpha;
                        m_used = m_new;
                    }
                    // pass 2: p = exp2(s*c - m_used) -> 16-bit -> swizzled K-major P tile.  The P buffer is still being
                    // read by the PV MMA of tile j-1 until pv_done (S no longer orders this: it is computed ahead).
                    ATT_STAMP(3);
                    if (j >= NPB) mbar_wait(pv_done(x, j % NPB), (j / NPB - 1) & 1, 17);   // PV(j - NPB) has read this P buffer
                    ATT_STAMP(4);
                    const float sc = p.scale_log2, mu = m_used;
                    auto chunk_p = [&](const uint32_t (&v)[32], int c) {
                        float pf[32];
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            float e = ex2_approx(fmaf(__uint_as_float(v[i]), sc, -mu));
                            if constexpr (!FULL) { if (c * 32 + i >= valid) e = 0.f; }
                            pf[i] = e;
                        }
#pragma unroll
                        for (int i = 0; i < 32; i += 4) psum += (pf[i] + pf[i + 1]) + (pf[i + 2] + pf[i + 3]);
                        uint8_t* chunk = pbuf + (c >> 1) * TILE_BYTES + row * 128;
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            uint4 w;
                            w.x = pack_h2(pf[u * 8 + 0], pf[u * 8 + 1], F16);
                            w.y = pack_h2(pf[u * 8 + 2], pf[u * 8 + 3], F16);
                            w.z = pack_h2(pf[u * 8 + 4], pf[u * 8 + 5], F16);
                            w.w = pack_h2(pf[u * 8 + 6], pf[u * 8 + 7], F16);
                            const int unit = ((c & 1) * 4 + u) ^ (row & 7);
                            *reinterpret_cast<uint4*>(chunk + unit * 16) = w;
                        }
                    };
                    tmem_ld_32x32(ts + 32, vb);
                    chunk_p(va, 0);
                    tmem_ld_wait(); tmem_regs_ready(vb);
                    tmem_ld_32x32(ts + 64, va);
                    chunk_p(vb, 1);
                    tmem_ld_wait(); tmem_regs_ready(va);
                    tmem_ld_32x32(ts + 96, vb);
                    chunk_p(va, 2);
                    tmem_ld_wait(); tmem_regs_ready(vb);
                    if constexpr (MI) {                  // S_X(j) is now entirely in registers: hand the buffer back
                        tc_fence_before();
                        mbar_arrive(s_free(sbuf));
                    }
                    chunk_p(vb, 3);
                };
                if (valid == ATT_N) run_tile(std::true_type{});
                else run_tile(std::false_type{});
                l_run += psum;
                ATT_STAMP(5);
                fence_proxy_async_smem();
                tc_fence_before();
                mbar_arrive(p_full(x));
            }
            // epilogue: O_X / l
            mbar_wait(pv_done(x, (ntiles - 1) % NPB), ((ntiles - 1) / NPB) & 1, 16);
            tc_fence_after();
            const float inv = 1.0f / l_run;
            const bool row_ok = (qrow0 + row) < p.Lq;
            uint16_t* op = p.o + static_cast<long long>(b) * p.o_batch_stride +
                           static_cast<long long>(qrow0 + row) * p.o_row_stride + head * ATT_D;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                uint32_t v[32];
                tmem_ld_32x32(to + c * 32, v);
                tmem_ld_wait(); tmem_regs_ready(v);
                if (row_ok) {
#pragma unroll
                    for (int i = 0; i < 32; i += 8) {
                        uint4 w;
                        w.x = pack_h2(__uint_as_float(v[i]) * inv, __uint_as_float(v[i + 1]) * inv, F16);
                        w.y = pack_h2(__uint_as_float(v[i + 2]) * inv, __uint_as_float(v[i + 3]) * inv, F16);
                        w.z = pack_h2(__uint_as_float(v[i + 4]) * inv, __uint_as_float(v[i + 5]) * inv, F16);
                        w.w = pack_h2(__uint_as_float(v[i + 6]) * inv, __uint_as_float(v[i + 7]) * inv, F16);
                        *reinterpret_cast<uint4*>(op + c * 32 + i) = w;
                    }
                }
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

// ---------------------------------------------------------------------------------------------------------
// v3 (round 2, default).  Same CTA tile (2 x 128 query rows of one (episode, head)), same TMA producer and K/V ring fed
// from two tensor maps, same 3 rotating S accumulators + 2 O accumulators in TMEM.  What changed, each on evidence from
// scripts/microbench (profiles/r02_microbench_*.log):
//   * The "one tcgen05.mma per ~82 cycles per issuing thread" floor of round 1 was the issue loop itself (descriptor
//     assembly, a runtime modulo): with the descriptors hoisted one thread issues M128 N128 K16 at 64.2 and M128 N64 K16
//     at 32.1 cycles (99.6 % / 99.7 % of the tensor pipe, umma_ts.cu).  The MMA warp keeps per-stage descriptor bases in
//     registers and adds immediates.
//   * P never goes through shared memory.  The softmax threads write the 16-bit probabilities back into TMEM, over the
//     first 64 columns of the very S accumulator they came from (each thread holds its whole row in registers by then),
//     and O += P V runs with the A operand in TMEM (tcgen05.mma ".ts" form).  With P in smem the PV instruction read 6 KiB
//     of operands per 32 cycles of math and ran at 48 cycles (the 128 B/clk smem port); from TMEM it runs at 32.  It also
//     removes the 16 STS.128 + proxy fence per row, 64 KiB of smem (now two more K/V stages) and the wait for "PV(j-1) has
//     read the P buffer": the buffer of S(seq) is next written by S(seq + 3), which the same thread issues after PV(seq),
//     and MMAs of one thread execute in order.
//   * One pass over the logits.  The row maximum is not recomputed per tile (attn_softmax.cuh): P = 2^(s c - m_used) is
//     evaluated against the maximum in use and only when a row sum of the warp exceeds 2^10 (or on the first / a ragged
//     tile) the warp re-reads S, takes the exact maximum, rescales O and l and redoes the tile.
//   * NPOLY/16 of the exponentials run on the FMA pipe (degree-3 polynomial, packed fp32x2 arithmetic) next to the
//     16-lane XU; scale-and-subtract, row sum and the fp16 pack are packed instructions too.
// ---------------------------------------------------------------------------------------------------------
constexpr int KV_STAGES3 = 5;
constexpr int ATT_SMEM3 = ATT_QT * TILE_BYTES + KV_STAGES3 * 2 * TILE_BYTES + 1024 + 256;
static_assert(ATT_SMEM3 <= 227 * 1024, "dynamic smem limit of sm_100");
#ifndef ATT_PACE
#define ATT_PACE 0
#endif
#ifndef ATT_MMA_WARP
#define ATT_MMA_WARP 1
#endif
#ifndef DFW_ATT_NPOLY
#define DFW_ATT_NPOLY 5
#endif
constexpr int ATT_NPOLY = DFW_ATT_NPOLY;                      // of every 16 logit pairs on the polynomial path (softmax_rate.cu)

__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "}\n"
        ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(acc)
        : "memory");
}

template <bool F16>
__global__ void __launch_bounds__(ATT_THREADS, 1)
attn_kvfused_kernel(const __grid_constant__ AttnMaps maps, const __grid_constant__ AttnParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_u32 = smem_u32(smem_raw);
    const uint32_t base = (raw_u32 + 1023u) & ~1023u;
    constexpr int KVS = KV_STAGES3;
    auto sQ = [&](int x) { return base + x * TILE_BYTES; };
    const uint32_t kv_base = base + ATT_QT * TILE_BYTES;
    auto sK = [&](int s) { return kv_base + s * 2 * TILE_BYTES; };
    auto sV = [&](int s) { return kv_base + s * 2 * TILE_BYTES + TILE_BYTES; };
    const uint32_t bar_base = kv_base + KVS * 2 * TILE_BYTES;
    const uint32_t q_full = bar_base;
    auto kv_full = [&](int s) { return bar_base + 8u * (1 + s); };
    auto kv_empty = [&](int s) { return bar_base + 8u * (1 + KVS + s); };
    auto s_full = [&](int buf) { return bar_base + 8u * (1 + 2 * KVS + buf); };      // one per S buffer
    auto p_full = [&](int x) { return bar_base + 8u * (4 + 2 * KVS + x); };
    auto pv_done = [&](int x) { return bar_base + 8u * (6 + 2 * KVS + x); };
    const uint32_t tmem_slot = bar_base + 8u * (8 + 2 * KVS);
    auto pace = [&](int i) { return bar_base + 8u * (9 + 2 * KVS + i); };            // MMA issue pacing (ATT_PACE groups in flight)
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw_u32));

    const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * ATT_M * ATT_QT;
    const int head = blockIdx.y;
    const int b = blockIdx.z;
    const int ntiles = p.n_self + p.n_bank;
    const bool has_b = (q0 + ATT_M) < p.Lq;          // second query tile holds at least one valid row
    const int nq = has_b ? 2 : 1;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&maps.q);
        tma_prefetch_desc(&maps.k_self);
        tma_prefetch_desc(&maps.v_self);
        if (p.n_bank) { tma_prefetch_desc(&maps.k_bank); tma_prefetch_desc(&maps.v_bank); }
    }
    if (warp == ATT_MMA_WARP && lane == 0) {
        mbar_init(q_full, 1);
        for (int s = 0; s < KVS; ++s) { mbar_init(kv_full(s), 1); mbar_init(kv_empty(s), 1); }
        for (int x = 0; x < ATT_SBUF; ++x) mbar_init(s_full(x), 1);
        for (int x = 0; x < ATT_QT; ++x) { mbar_init(p_full(x), 128); mbar_init(pv_done(x), 1); }
        for (int i = 0; i < (ATT_PACE > 0 ? ATT_PACE : 1); ++i) mbar_init(pace(i), 1);
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
    pdl_wait();
    auto tS = [&](int buf) { return tmem_base + buf * 128; };
    auto tO = [&](int x) { return tmem_base + ATT_SBUF * 128 + x * 64; };
    // register budget (setmaxnreg inside each role branch, so ptxas allocates per role): the softmax threads hold a
    // whole row of 128 logits in the exact path; the control warps need few
    if (warp == 0) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 48;");
        if (elect_one()) {
            mbar_arrive_expect_tx(q_full, nq * TILE_BYTES);
            tma_load_3d(sQ(0), &maps.q, q_full, head * ATT_D, q0, b);
            if (has_b) tma_load_3d(sQ(1), &maps.q, q_full, head * ATT_D, q0 + ATT_M, b);
        }
        __syncwarp();
        int s = 0;
        uint32_t ph = 1;
        for (int j = 0; j < ntiles; ++j) {
            mbar_wait(kv_empty(s), ph, 10);
            if (elect_one()) {
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
            __syncwarp();
            if (++s == KVS) { s = 0; ph ^= 1u; }
        }
    } else if (warp == ATT_MMA_WARP) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 48;");
        const uint32_t fmt = F16 ? 0u : 1u;
        const uint32_t idesc_s = umma_idesc(ATT_M, ATT_N, fmt, fmt, 0);   // S = Q K^T : B (= K) is K-major
        const uint32_t idesc_o = umma_idesc(ATT_M, ATT_D, fmt, fmt, 1);   // O += P V  : A (= P) in TMEM, B (= V) MN-major
        // descriptor bases, hoisted: a stage / k-step only moves the 14-bit start-address field (address >> 4)
        const uint64_t dq0 = umma_desc_sw128(sQ(0)), dq1 = umma_desc_sw128(sQ(1));
        const uint64_t dk0 = umma_desc_sw128(sK(0)), dv0 = umma_desc_sw128(sV(0));
        constexpr uint32_t STAGE_STEP = (2 * TILE_BYTES) >> 4;
        // S tiles are numbered in issue order, seq = nq * j + x, and live in TMEM buffer seq % 3; P(seq) overwrites the
        // first 64 columns of that buffer and is consumed by PV(seq), issued before S(seq + 3) by this same thread.
        // Pacing: a tcgen05.mma issued into a full tensor-pipe queue blocks its warp AT the instruction, and while it sits
        // there the other warps of the sub-partition issue their MUFU / tcgen05.ld instructions more slowly (the softmax warps
        // that share the MMA warp's sub-partition ran their exponentials 43 % slower, profiles/r02_attn_timeline_v3_a.log).
        // So MMAs go out in groups of four (128 - 256 tensor-pipe cycles) with at most ATT_PACE groups in flight: before
        // group g the warp waits -- parked on an mbarrier, not blocked at issue -- for the commit of group g - ATT_PACE.
        uint32_t grp = 0;
        auto pace_wait = [&]() {                                 // all lanes
            if (ATT_PACE > 0 && grp >= ATT_PACE) mbar_wait(pace(grp % ATT_PACE), ((grp / ATT_PACE) - 1) & 1, 20);
        };
        auto pace_commit = [&]() {                               // elected lane
            if (ATT_PACE > 0) tc_commit(pace(grp % ATT_PACE));
        };
        auto issue_s = [&](int x, int stage, int buf) {          // all lanes; S(x, .) = one group
            pace_wait();
            if (elect_one()) {
                const uint64_t adesc = x ? dq1 : dq0;
                const uint64_t bdesc = dk0 + static_cast<uint64_t>(stage * STAGE_STEP);
                const uint32_t d = tS(buf);
#pragma unroll
                for (int k = 0; k < ATT_D / 16; ++k) umma_ss(d, adesc + 2u * k, bdesc + 2u * k, idesc_s, k > 0 ? 1u : 0u);
                tc_commit(s_full(buf));
                pace_commit();
            }
            __syncwarp();
            ++grp;
        };
        mbar_wait(q_full, 0, 12);
        mbar_wait(kv_full(0), 0, 11);
        tc_fence_after();
        for (int x = 0; x < nq; ++x) issue_s(x, 0, x);
        int s = 0, s1 = (KVS > 1) ? 1 : 0;              // stage of tile j, of tile j + 1
        uint32_t ph1 = (KVS > 1) ? 0u : 1u;             // kv_full parity of tile j + 1
        int buf = 0;                                    // S buffer of (x = 0, j)
        for (int j = 0; j < ntiles; ++j) {
            ATT_STAMP(4);
            if (j + 1 < ntiles) {
                mbar_wait(kv_full(s1), ph1, 11);
                tc_fence_after();
            }
            ATT_STAMP(5);
            int bx = buf;                               // buffer of (x, j)
            int bn = buf + nq; if (bn >= ATT_SBUF) bn -= ATT_SBUF;      // buffer of (x, j + 1) = seq + nq
            for (int x = 0; x < nq; ++x) {
                if (j + 1 < ntiles) issue_s(x, s1, bn);     // run ahead: S_X(j+1) goes to the tensor pipe before P_X(j) is waited for
                ATT_STAMP(x * 2);
                mbar_wait(p_full(x), j & 1, 13);        // P_X(j) in TMEM, O_X rescaled if it had to be
                tc_fence_after();
                ATT_STAMP(x * 2 + 1);
                const uint64_t bdesc = dv0 + static_cast<uint64_t>(s * STAGE_STEP);
                const uint32_t d = tO(x), a = tS(bx);
#pragma unroll
                for (int half = 0; half < 2; ++half) {      // PV_X(j) = two groups of four k-steps
                    pace_wait();
                    if (elect_one()) {
#pragma unroll
                        for (int ks = half * 4; ks < half * 4 + 4; ++ks)
                            umma_ts(d, a + ks * 8, bdesc + static_cast<uint64_t>(ks * ((16 * 128) >> 4)), idesc_o, (j > 0 || ks > 0) ? 1u : 0u);
                        if (half == 1) {
#ifdef ATT_DUMMY_COMMITS
#pragma unroll
                            for (int dc = 0; dc < ATT_DUMMY_COMMITS; ++dc) tc_commit(pace(0));     // experiment: marginal cost of a commit
#endif
                            tc_commit(pv_done(x));
                            if (x == nq - 1) tc_commit(kv_empty(s));        // K_j / V_j fully consumed by both tiles
                        }
                        pace_commit();
                    }
                    __syncwarp();
                    ++grp;
                }
                ATT_STAMP(7 - x);
                if (++bx == ATT_SBUF) bx = 0;
                if (++bn == ATT_SBUF) bn = 0;
            }
            buf += nq; if (buf >= ATT_SBUF) buf -= ATT_SBUF;
            s = s1;
            if (++s1 == KVS) { s1 = 0; ph1 ^= 1u; }
        }
    } else if (warp < 4) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 48;");
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
        const int x = (warp - 4) >> 2;                // query tile of this softmax group
        const int qd = (warp - 4) & 3;                // TMEM lane quadrant
        const int row = qd * 32 + lane;
        const int qrow0 = q0 + x * ATT_M;
        if (x == 0 || has_b) {
            const uint32_t lane_off = static_cast<uint32_t>(qd * 32) << 16;
            const uint32_t to = tO(x) + lane_off;
            const float sc = p.scale_log2;
            float m_used = -INFINITY, l_run = 0.f;
            int sbuf = x;                             // seq = nq * j + x -> buffer seq % 3, parity (seq / 3) & 1
            uint32_t sph = 0;
            for (int j = 0; j < ntiles; ++j) {
                int valid;
                if (j < p.n_self) valid = min(ATT_N, p.Ls - j * ATT_N);
                else valid = min(ATT_N, p.Lb - (j - p.n_self) * ATT_N);
                const uint32_t ts = tS(sbuf) + lane_off;
                ATT_STAMP(0);
                mbar_wait(s_full(sbuf), sph, 15);
                tc_fence_after();
                ATT_STAMP(1);
                uint32_t pk[64];
                float psum = 0.f;
                bool slow = (j == 0) || (valid != ATT_N);         // warp-uniform
                if (!slow) {
                    psum = exp_row128_tmem<F16, ATT_NPOLY>(ts, sc, m_used, pk);
                    slow = __any_sync(0xffffffffu, !(psum <= SOFTMAX_TRIGGER));
                }
                ATT_STAMP(2);
                if (slow) {
                    // exact path: whole row from TMEM (still intact: P is stored below), true maximum, O / l rescale
                    uint32_t sv[128];
                    tmem_ld_row128(ts, sv);
                    if (valid != ATT_N) {
#pragma unroll
                        for (int i = 0; i < 128; ++i)
                            if (i >= valid) sv[i] = 0xff800000u;                     // -inf: masked key
                    }
                    const float m_new = fmaxf(m_used, row_max128(sv) * sc);
                    const float alpha = ex2_approx(m_used - m_new);                   // 0 on the first tile
                    if (j > 0 && __any_sync(0xffffffffu, m_new != m_used)) {
                        mbar_wait(pv_done(x), (j - 1) & 1, 14);                       // O_X holds tiles < j
                        tc_fence_after();
#pragma unroll
                        for (int c = 0; c < 2; ++c) {
                            uint32_t v[32];
                            tmem_ld_32x32(to + c * 32, v);
                            tmem_ld_wait(); tmem_regs_ready(v);
#pragma unroll
                            for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
                            tmem_st_32x32(to + c * 32, v);
                        }
                    }
                    l_run *= alpha;
                    m_used = m_new;
                    psum = exp_row128_staged<F16, ATT_NPOLY, 2>(sv, sc, m_used, pk);
                }
                l_run += psum;
                ATT_STAMP(3);
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    uint32_t (&v)[16] = *reinterpret_cast<uint32_t (*)[16]>(&pk[c * 16]);
                    tmem_st_32x16(ts + c * 16, v);
                }
                tmem_st_wait();
                ATT_STAMP(4);
                // S_X(j+1) is computed ahead, so a warp can start (and finish) tile j+1 while a sibling is still in the
                // exact path of tile j.  Its arrival must not be counted in phase j of p_full (the MMA warp would issue
                // PV(j) before the sibling's rows of P exist), and parity waits on pv_done are only unambiguous while a
                // waiter is at most one phase behind.  Both follow from observing EVERY phase of pv_done in order: before
                // arriving for tile j, wait for PV(j-1) -- issued a whole tile ago, so normally one successful try_wait.
                if (j > 0) mbar_wait(pv_done(x), (j - 1) & 1, 18);
                tc_fence_before();
                mbar_arrive(p_full(x));
                ATT_STAMP(5);
                sbuf += nq;
                if (sbuf >= ATT_SBUF) { sbuf -= ATT_SBUF; sph ^= 1u; }
            }
            // epilogue: O_X / l  (PV(ntiles - 2) was observed before the last arrive, so this parity wait is unambiguous)
            mbar_wait(pv_done(x), (ntiles - 1) & 1, 16);
            tc_fence_after();
            const float inv = 1.0f / l_run;
            const bool row_ok = (qrow0 + row) < p.Lq;
            if (p.lse != nullptr && row_ok)
                p.lse[(static_cast<long long>(b) * gridDim.y + head) * p.Lq + qrow0 + row] = m_used + log2f(l_run);
            uint16_t* op = p.o + static_cast<long long>(b) * p.o_batch_stride +
                           static_cast<long long>(qrow0 + row) * p.o_row_stride + head * ATT_D;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                uint32_t v[32];
                tmem_ld_32x32(to + c * 32, v);
                tmem_ld_wait(); tmem_regs_ready(v);
                if (row_ok) {
#pragma unroll
                    for (int i = 0; i < 32; i += 8) {
                        uint4 w;
                        w.x = pack_h2(__uint_as_float(v[i]) * inv, __uint_as_float(v[i + 1]) * inv, F16);
                        w.y = pack_h2(__uint_as_float(v[i + 2]) * inv, __uint_as_float(v[i + 3]) * inv, F16);
                        w.z = pack_h2(__uint_as_float(v[i + 4]) * inv, __uint_as_float(v[i + 5]) * inv, F16);
                        w.w = pack_h2(__uint_as_float(v[i + 6]) * inv, __uint_as_float(v[i + 7]) * inv, F16);
                        *reinterpret_cast<uint4*>(op + c * 32 + i) = w;
                    }
                }
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

// ---------------------------------------------------------------------------------------------------------
// v4 (round 2, second half).  The v3 timeline (profiles/r02_attn_timeline_v3_a.log) shows the two query tiles of a CTA
// running in strict alternation: a key-tile pair takes 2435 cycles although the two softmax warps of a sub-partition need
// only ~1400 XU cycles / ~1360 issue slots between them -- each warp spends 570-620 cycles waiting for its next S, and the
// single MMA warp needs 1035 cycles to get PV_B(j) and S_A(j+2) issued (it serves A and B in a fixed order and sits behind
// whichever softmax group is late).  With three rotating S buffers that order cannot be relaxed: S_A(j+1) lands in the
// buffer of P_B(j-1).  Here every query tile owns TWO S buffers -- 96-key tiles make four of them fit beside the two O
// accumulators (4 x 96 + 2 x 64 = 512 TMEM columns) -- and its own MMA-issuing warp (warps 1 and 3, different
// sub-partitions): loop { S_X(j+1) -> other buffer; wait P_X(j); O_X += P_X(j) V_j }.  A tile never waits for its sibling;
// the only shared resource is the K/V ring (a stage is released by both issuers' commits).  Softmax code, P-in-TMEM,
// lazy maximum and the polynomial exp2 share are v3's.
// MEASURED (profiles/r02_bench_attn_v4.json, back-to-back launches): correct on every shape of the suite and 3-5 % SLOWER
// than v3 (B16 h5 4096x8192: 820 vs 860 TFLOP/s; 9216x18432: 854 vs 884): decoupling the tiles buys about what the fixed
// per-tile cost (barrier round trips, P store, S wait) loses when it is amortised over 96 instead of 128 keys.  So the
// A/B chain was not the limiter: with two softmax warps per sub-partition the loop is bound by per-warp instruction
// latency (ncu: issue slots 56 %, XU 61 %, neither saturated).  Kept selectable (DFW_OPT_ATTN_V4), off.
// ---------------------------------------------------------------------------------------------------------
constexpr int ATT_N4 = 96;                                   // keys per tile
constexpr int KV4_TILE = ATT_N4 * 128;                       // 12 KiB: 96 rows x 128 B (12 SWIZZLE_128B atoms)
constexpr int KV_STAGES4 = 6;
constexpr int ATT_SMEM4 = ATT_QT * TILE_BYTES + KV_STAGES4 * 2 * KV4_TILE + 1024 + 256;
static_assert(ATT_SMEM4 <= 227 * 1024, "dynamic smem limit of sm_100");
static_assert(4 * ATT_N4 + 2 * ATT_D <= 512, "TMEM budget");

template <bool F16>
__global__ void __launch_bounds__(ATT_THREADS, 1)
attn_kvfused_v4_kernel(const __grid_constant__ AttnMaps maps, const __grid_constant__ AttnParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_u32 = smem_u32(smem_raw);
    const uint32_t base = (raw_u32 + 1023u) & ~1023u;
    constexpr int KVS = KV_STAGES4;
    constexpr int NCH = ATT_N4 / 32;
    auto sQ = [&](int x) { return base + x * TILE_BYTES; };
    const uint32_t kv_base = base + ATT_QT * TILE_BYTES;
    auto sK = [&](int s) { return kv_base + s * 2 * KV4_TILE; };
    auto sV = [&](int s) { return kv_base + s * 2 * KV4_TILE + KV4_TILE; };
    const uint32_t bar_base = kv_base + KVS * 2 * KV4_TILE;
    const uint32_t q_full = bar_base;
    auto kv_full = [&](int s) { return bar_base + 8u * (1 + s); };
    auto kv_empty = [&](int s) { return bar_base + 8u * (1 + KVS + s); };
    auto s_full = [&](int x, int par) { return bar_base + 8u * (1 + 2 * KVS + 2 * x + par); };
    auto p_full = [&](int x) { return bar_base + 8u * (5 + 2 * KVS + x); };
    auto pv_done = [&](int x) { return bar_base + 8u * (7 + 2 * KVS + x); };
    const uint32_t tmem_slot = bar_base + 8u * (9 + 2 * KVS);
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw_u32));

    const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * ATT_M * ATT_QT;
    const int head = blockIdx.y;
    const int b = blockIdx.z;
    const int ntiles = p.n_self + p.n_bank;
    const bool has_b = (q0 + ATT_M) < p.Lq;          // second query tile holds at least one valid row
    const int nq = has_b ? 2 : 1;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&maps.q);
        tma_prefetch_desc(&maps.k_self);
        tma_prefetch_desc(&maps.v_self);
        if (p.n_bank) { tma_prefetch_desc(&maps.k_bank); tma_prefetch_desc(&maps.v_bank); }
    }
    if (warp == 1 && lane == 0) {
        mbar_init(q_full, 1);
        for (int s = 0; s < KVS; ++s) { mbar_init(kv_full(s), 1); mbar_init(kv_empty(s), nq); }     // released by every issuer
        for (int x = 0; x < ATT_QT; ++x) {
            mbar_init(s_full(x, 0), 1); mbar_init(s_full(x, 1), 1);
            mbar_init(p_full(x), 128); mbar_init(pv_done(x), 1);
        }
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
    pdl_wait();
    auto tS = [&](int x, int par) { return tmem_base + (2 * x + par) * ATT_N4; };
    auto tO = [&](int x) { return tmem_base + 4 * ATT_N4 + x * ATT_D; };
    if (warp == 0) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
        if (elect_one()) {
            mbar_arrive_expect_tx(q_full, nq * TILE_BYTES);
            tma_load_3d(sQ(0), &maps.q, q_full, head * ATT_D, q0, b);
            if (has_b) tma_load_3d(sQ(1), &maps.q, q_full, head * ATT_D, q0 + ATT_M, b);
        }
        __syncwarp();
        int s = 0;
        uint32_t ph = 1;
        for (int j = 0; j < ntiles; ++j) {
            mbar_wait(kv_empty(s), ph, 10);
            if (elect_one()) {
                mbar_arrive_expect_tx(kv_full(s), 2 * KV4_TILE);
                if (j < p.n_self) {
                    tma_load_3d(sK(s), &maps.k_self, kv_full(s), head * ATT_D, j * ATT_N4, b);
                    tma_load_3d(sV(s), &maps.v_self, kv_full(s), head * ATT_D, j * ATT_N4, b);
                } else {
                    const int jb = j - p.n_self;
                    tma_load_3d(sK(s), &maps.k_bank, kv_full(s), head * ATT_D, jb * ATT_N4, b);
                    tma_load_3d(sV(s), &maps.v_bank, kv_full(s), head * ATT_D, jb * ATT_N4, b);
                }
            }
            __syncwarp();
            if (++s == KVS) { s = 0; ph ^= 1u; }
        }
    } else if (warp == 1 || warp == 3) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
        const int x = warp >> 1;                          // warp 1 issues for tile A, warp 3 for tile B
        if (x == 0 || has_b) {
            const uint32_t fmt = F16 ? 0u : 1u;
            const uint32_t idesc_s = umma_idesc(ATT_M, ATT_N4, fmt, fmt, 0);  // S = Q K^T : B (= K) is K-major
            const uint32_t idesc_o = umma_idesc(ATT_M, ATT_D, fmt, fmt, 1);   // O += P V  : A (= P) in TMEM, B (= V) MN-major
            const uint64_t dq = umma_desc_sw128(sQ(x));
            const uint64_t dk0 = umma_desc_sw128(sK(0)), dv0 = umma_desc_sw128(sV(0));
            constexpr uint32_t STAGE_STEP = (2 * KV4_TILE) >> 4;
            auto issue_s = [&](int stage, int par) {          // all lanes
                if (elect_one()) {
                    const uint64_t bdesc = dk0 + static_cast<uint64_t>(stage * STAGE_STEP);
                    const uint32_t d = tS(x, par);
#pragma unroll
                    for (int k = 0; k < ATT_D / 16; ++k) umma_ss(d, dq + 2u * k, bdesc + 2u * k, idesc_s, k > 0 ? 1u : 0u);
                    tc_commit(s_full(x, par));
                }
                __syncwarp();
            };
            mbar_wait(q_full, 0, 12);
            mbar_wait(kv_full(0), 0, 11);
            tc_fence_after();
            issue_s(0, 0);
            int s = 0, s1 = 1;
            uint32_t ph1 = 0u;
            for (int j = 0; j < ntiles; ++j) {
                if (j + 1 < ntiles) {
                    mbar_wait(kv_full(s1), ph1, 11);
                    tc_fence_after();
                    // S_X(j+1) -> the buffer P_X(j-1) lived in: PV_X(j-1) was issued by this thread, MMAs of one thread run in order
                    issue_s(s1, (j + 1) & 1);
                }
                mbar_wait(p_full(x), j & 1, 13);            // P_X(j) in TMEM, O_X rescaled if it had to be
                tc_fence_after();
                if (elect_one()) {
                    const uint64_t bdesc = dv0 + static_cast<uint64_t>(s * STAGE_STEP);
                    const uint32_t d = tO(x), a = tS(x, j & 1);
#pragma unroll
                    for (int ks = 0; ks < ATT_N4 / 16; ++ks)
                        umma_ts(d, a + ks * 8, bdesc + static_cast<uint64_t>(ks * ((16 * 128) >> 4)), idesc_o, (j > 0 || ks > 0) ? 1u : 0u);
                    tc_commit(pv_done(x));
                    tc_commit(kv_empty(s));                 // this tile is done with K_j / V_j
                }
                __syncwarp();
                s = s1;
                if (++s1 == KVS) { s1 = 0; ph1 ^= 1u; }
            }
        }
    } else if (warp < 4) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
        const int x = (warp - 4) >> 2;                // query tile of this softmax group
        const int qd = (warp - 4) & 3;                // TMEM lane quadrant
        const int row = qd * 32 + lane;
        const int qrow0 = q0 + x * ATT_M;
        if (x == 0 || has_b) {
            const uint32_t lane_off = static_cast<uint32_t>(qd * 32) << 16;
            const uint32_t to = tO(x) + lane_off;
            const float sc = p.scale_log2;
            float m_used = -INFINITY, l_run = 0.f;
            for (int j = 0; j < ntiles; ++j) {
                int valid;
                if (j < p.n_self) valid = min(ATT_N4, p.Ls - j * ATT_N4);
                else valid = min(ATT_N4, p.Lb - (j - p.n_self) * ATT_N4);
                const uint32_t ts = tS(x, j & 1) + lane_off;
                mbar_wait(s_full(x, j & 1), (j >> 1) & 1, 15);
                tc_fence_after();
                uint32_t pk[ATT_N4 / 2];
                float psum = 0.f;
                bool slow = (j == 0) || (valid != ATT_N4);         // warp-uniform
                if (!slow) {
                    psum = exp_row_tmem_n<F16, ATT_NPOLY, NCH>(ts, sc, m_used, pk);
                    slow = __any_sync(0xffffffffu, !(psum <= SOFTMAX_TRIGGER));
                }
                if (slow) {
                    // exact path: whole row from TMEM (still intact: P is stored below), true maximum, O / l rescale
                    uint32_t sv[ATT_N4];
                    tmem_ld_row_n<NCH>(ts, sv);
                    if (valid != ATT_N4) {
#pragma unroll
                        for (int i = 0; i < ATT_N4; ++i)
                            if (i >= valid) sv[i] = 0xff800000u;                     // -inf: masked key
                    }
                    const float m_new = fmaxf(m_used, row_max_n<ATT_N4>(sv) * sc);
                    const float alpha = ex2_approx(m_used - m_new);                   // 0 on the first tile
                    if (j > 0 && __any_sync(0xffffffffu, m_new != m_used)) {
                        mbar_wait(pv_done(x), (j - 1) & 1, 14);                       // O_X holds tiles < j
                        tc_fence_after();
#pragma unroll
                        for (int c = 0; c < 2; ++c) {
                            uint32_t v[32];
                            tmem_ld_32x32(to + c * 32, v);
                            tmem_ld_wait(); tmem_regs_ready(v);
#pragma unroll
                            for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
                            tmem_st_32x32(to + c * 32, v);
                        }
                    }
                    l_run *= alpha;
                    m_used = m_new;
                    psum = exp_row_staged_n<F16, ATT_NPOLY, 2, NCH>(sv, sc, m_used, pk);
                }
                l_run += psum;
#pragma unroll
                for (int c = 0; c < ATT_N4 / 32; ++c) {
                    uint32_t (&v)[16] = *reinterpret_cast<uint32_t (*)[16]>(&pk[c * 16]);
                    tmem_st_32x16(ts + c * 16, v);
                }
                tmem_st_wait();
                // observe EVERY phase of pv_done in order (see v3): before arriving for tile j, wait for PV(j-1)
                if (j > 0) mbar_wait(pv_done(x), (j - 1) & 1, 18);
                tc_fence_before();
                mbar_arrive(p_full(x));
            }
            mbar_wait(pv_done(x), (ntiles - 1) & 1, 16);
            tc_fence_after();
            const float inv = 1.0f / l_run;
            const bool row_ok = (qrow0 + row) < p.Lq;
            if (p.lse != nullptr && row_ok)
                p.lse[(static_cast<long long>(b) * gridDim.y + head) * p.Lq + qrow0 + row] = m_used + log2f(l_run);
            uint16_t* op = p.o + static_cast<long long>(b) * p.o_batch_stride +
                           static_cast<long long>(qrow0 + row) * p.o_row_stride + head * ATT_D;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                uint32_t v[32];
                tmem_ld_32x32(to + c * 32, v);
                tmem_ld_wait(); tmem_regs_ready(v);
                if (row_ok) {
#pragma unroll
                    for (int i = 0; i < 32; i += 8) {
                        uint4 w;
                        w.x = pack_h2(__uint_as_float(v[i]) * inv, __uint_as_float(v[i + 1]) * inv, F16);
                        w.y = pack_h2(__uint_as_float(v[i + 2]) * inv, __uint_as_float(v[i + 3]) * inv, F16);
                        w.z = pack_h2(__uint_as_float(v[i + 4]) * inv, __uint_as_float(v[i + 5]) * inv, F16);
                        w.w = pack_h2(__uint_as_float(v[i + 6]) * inv, __uint_as_float(v[i + 7]) * inv, F16);
                        *reinterpret_cast<uint4*>(op + c * 32 + i) = w;
                    }
                }
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

int make_seq_map(CUtensorMap* m, const void* base, int C, int L, int B, int row_stride, long long batch_stride,
                 int box_rows = ATT_N) {
    const uint64_t dims[3] = {static_cast<uint64_t>(C), static_cast<uint64_t>(L), static_cast<uint64_t>(B)};
    const uint64_t strides[2] = {static_cast<uint64_t>(row_stride) * 2, static_cast<uint64_t>(batch_stride) * 2};
    const uint32_t box[3] = {ATT_D, static_cast<uint32_t>(box_rows), 1};
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
    pdl_wait();
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

// ---------------------------------------------------------------------------------------------------------
// K2b: cross attention to a FIXED short prompt, collapsed algebraically.  With the prompt's K / V constant (same
// empty-prompt embedding for every sample, Lctx = 2 at eval: pipeline:591-601, :690-692) the whole attn2 block
//     to_out(softmax(to_q(x) K^T * scale) V) + bias
// is  logits = x @ Wlog^T  with  Wlog[(h, j)] = scale * K[j, h] @ Wq[h]   (a [heads * Lctx, C] matrix, one small GEMM)
// and out    = bias + sum_{h, j} softmax_j(logits[h, :])[j] * U[(h, j)]   with  U[(h, j)] = Wo[:, h] @ V[j, h]  ([.., C]),
// i.e. two C x C GEMMs and the attention kernel become one skinny GEMM plus this bandwidth kernel (+ residual).
// CTA = 16 tokens: phase 1 one thread per (token, logit) -> probabilities in smem; phase 2 one thread per
// (token, 8 channels): acc = bias + sum p * U (U read through L1), + residual, 16-byte store.
// ---------------------------------------------------------------------------------------------------------
constexpr int XC_TOK = 16;      // tokens per CTA
constexpr int XC_OCT = 16;      // 8-channel octets per CTA: the CTA's slice of U (heads*Lctx x 128 floats) stays in L1
template <int RD>    // residual / output dtype: 0 bf16, 1 fp32, 2 fp16
__global__ void __launch_bounds__(256) cross_attn_collapsed_kernel(const float* __restrict__ logits, int ld_logits,
                                                                   const float* __restrict__ U, const float* __restrict__ bias,
                                                                   const void* __restrict__ residual, void* __restrict__ out,
                                                                   long long M, int C, int heads, int Lctx) {
    // thread = (token, 8 channels); no shared memory, no barrier: the per-head softmax over the Lctx (2 at eval) logits
    // is recomputed by the 16 threads of a token (a handful of MUFU ops) and every load is independent of the others
    const int ol = threadIdx.x % XC_OCT, tl = threadIdx.x / XC_OCT;
    const long long m = static_cast<long long>(blockIdx.y) * XC_TOK + tl;
    const int oc = blockIdx.x * XC_OCT + ol;
    if (m >= M || oc * 8 >= C) return;
    const long long off = m * C + oc * 8;
    float acc[8];
    if (residual != nullptr) {
        if constexpr (RD == 1) {
            const float4* rp = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(residual) + off);
            const float4 r0 = __ldg(rp), r1 = __ldg(rp + 1);
            acc[0] = r0.x; acc[1] = r0.y; acc[2] = r0.z; acc[3] = r0.w; acc[4] = r1.x; acc[5] = r1.y; acc[6] = r1.z; acc[7] = r1.w;
        } else {
            const uint4 r = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(residual) + off));
            unpack8(r, RD == 2, acc);
        }
    } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    }
    {
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + oc * 8)), b1 = __ldg(reinterpret_cast<const float4*>(bias + oc * 8) + 1);
        acc[0] += b0.x; acc[1] += b0.y; acc[2] += b0.z; acc[3] += b0.w; acc[4] += b1.x; acc[5] += b1.y; acc[6] += b1.z; acc[7] += b1.w;
    }
    const float* lg = logits + m * ld_logits;
    const float* up = U + oc * 8;
    for (int h = 0; h < heads; ++h) {
        float l[8];
        float mx = -INFINITY;
        for (int j = 0; j < Lctx; ++j) { l[j] = __ldg(lg + h * Lctx + j); mx = fmaxf(mx, l[j]); }
        float sum = 0.f;
        for (int j = 0; j < Lctx; ++j) { l[j] = __expf(l[j] - mx); sum += l[j]; }
        const float inv = __fdividef(1.0f, sum);
        for (int j = 0; j < Lctx; ++j) {
            const float pj = l[j] * inv;
            const float4* u4 = reinterpret_cast<const float4*>(up + static_cast<size_t>(h * Lctx + j) * C);
            const float4 u0 = __ldg(u4), u1 = __ldg(u4 + 1);
            acc[0] = fmaf(pj, u0.x, acc[0]); acc[1] = fmaf(pj, u0.y, acc[1]); acc[2] = fmaf(pj, u0.z, acc[2]); acc[3] = fmaf(pj, u0.w, acc[3]);
            acc[4] = fmaf(pj, u1.x, acc[4]); acc[5] = fmaf(pj, u1.y, acc[5]); acc[6] = fmaf(pj, u1.z, acc[6]); acc[7] = fmaf(pj, u1.w, acc[7]);
        }
    }
    if constexpr (RD == 1) {
        float4* op = reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + off);
        op[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
        op[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
    } else {
        constexpr int f16 = (RD == 2);
        uint4 o;
        o.x = pack_h2(acc[0], acc[1], f16); o.y = pack_h2(acc[2], acc[3], f16);
        o.z = pack_h2(acc[4], acc[5], f16); o.w = pack_h2(acc[6], acc[7], f16);
        *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(out) + off) = o;
    }
}

}  // namespace
}  // namespace dfw

#ifdef DFW_ATTN_TRACE
namespace dfw { long long* g_attn_trace = nullptr; }
extern "C" void dfw_attn_trace_buffer(void* p) { dfw::g_attn_trace = reinterpret_cast<long long*>(p); }
#endif

extern "C" {

static int attn_fwd_impl(const void* q, long long q_batch_stride, int q_row_stride, const void* k_self,
                         const void* v_self, long long kv_self_batch_stride, int kv_self_row_stride,
                         const void* k_bank, const void* v_bank, long long kv_bank_batch_stride,
                         int kv_bank_row_stride, void* o, long long o_batch_stride, int o_row_stride, int B,
                         int heads, int Lq, int Ls, int Lb, float scale, int f16, float* lse, void* stream_) {
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
    const bool v2 = get_option(DFW_OPT_ATTN_V2) && lse == nullptr;
    const bool v4 = !v2 && get_option(DFW_OPT_ATTN_V4);
    const int kt = v4 ? ATT_N4 : ATT_N;                  // keys per tile
    AttnMaps maps;
    rc = make_seq_map(&maps.q, q, C, Lq, B, q_row_stride, q_batch_stride, ATT_M);
    if (rc != DFW_OK) return rc;
    rc = make_seq_map(&maps.k_self, k_self, C, Ls, B, kv_self_row_stride, kv_self_batch_stride, kt);
    if (rc != DFW_OK) return rc;
    rc = make_seq_map(&maps.v_self, v_self, C, Ls, B, kv_self_row_stride, kv_self_batch_stride, kt);
    if (rc != DFW_OK) return rc;
    if (Lb > 0) {
        DFW_REQUIRE(kv_bank_row_stride % 8 == 0 && kv_bank_batch_stride % 8 == 0);
        rc = make_seq_map(&maps.k_bank, k_bank, C, Lb, B, kv_bank_row_stride, kv_bank_batch_stride, kt);
        if (rc != DFW_OK) return rc;
        rc = make_seq_map(&maps.v_bank, v_bank, C, Lb, B, kv_bank_row_stride, kv_bank_batch_stride, kt);
        if (rc != DFW_OK) return rc;
    } else {
        maps.k_bank = maps.k_self;
        maps.v_bank = maps.v_self;
    }
    AttnParams p{};
    p.Lq = Lq; p.Ls = Ls; p.Lb = Lb;
    p.n_self = (Ls + kt - 1) / kt;
    p.n_bank = (Lb + kt - 1) / kt;
    p.scale_log2 = scale * 1.4426950408889634f;
    p.o = reinterpret_cast<uint16_t*>(o);
    p.f16 = f16;
    p.lse = lse;
    p.o_batch_stride = o_batch_stride;
    p.o_row_stride = o_row_stride;
#ifdef DFW_ATTN_TRACE
    p.trace = g_attn_trace;
#endif
    static bool attr_set = false;
    if (!attr_set) {
        DFW_CHECK_CUDA(cudaFuncSetAttribute(attn_kvfused_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM3));
        DFW_CHECK_CUDA(cudaFuncSetAttribute(attn_kvfused_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM3));
        DFW_CHECK_CUDA(cudaFuncSetAttribute(attn_kvfused_v4_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM4));
        DFW_CHECK_CUDA(cudaFuncSetAttribute(attn_kvfused_v4_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM4));
        DFW_CHECK_CUDA(cudaFuncSetAttribute(attn_kvfused_v2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM));
        DFW_CHECK_CUDA(cudaFuncSetAttribute(attn_kvfused_v2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM));
        attr_set = true;
    }
    dim3 grid((Lq + ATT_M * ATT_QT - 1) / (ATT_M * ATT_QT), heads, B);
    cudaStream_t st = static_cast<cudaStream_t>(stream_);
    if (v2) {        // round-1 kernel (P through smem, two passes over S), kept for A/B measurements
        if (f16) DFW_CHECK_CUDA(launch_k(attn_kvfused_v2_kernel<true>, grid, ATT_THREADS, ATT_SMEM, st, maps, p));
        else DFW_CHECK_CUDA(launch_k(attn_kvfused_v2_kernel<false>, grid, ATT_THREADS, ATT_SMEM, st, maps, p));
    } else if (v4) {
        if (f16) DFW_CHECK_CUDA(launch_k(attn_kvfused_v4_kernel<true>, grid, ATT_THREADS, ATT_SMEM4, st, maps, p));
        else DFW_CHECK_CUDA(launch_k(attn_kvfused_v4_kernel<false>, grid, ATT_THREADS, ATT_SMEM4, st, maps, p));
    } else {
        if (f16) DFW_CHECK_CUDA(launch_k(attn_kvfused_kernel<true>, grid, ATT_THREADS, ATT_SMEM3, st, maps, p));
        else DFW_CHECK_CUDA(launch_k(attn_kvfused_kernel<false>, grid, ATT_THREADS, ATT_SMEM3, st, maps, p));
    }
    g_launches.fetch_add(1);
    DFW_CHECK_CUDA(cudaGetLastError());
    return DFW_OK;
}

int dfw_attn_kvfused_fwd(const void* q, long long q_batch_stride, int q_row_stride, const void* k_self,
                         const void* v_self, long long kv_self_batch_stride, int kv_self_row_stride,
                         const void* k_bank, const void* v_bank, long long kv_bank_batch_stride,
                         int kv_bank_row_stride, void* o, long long o_batch_stride, int o_row_stride, int B,
                         int heads, int Lq, int Ls, int Lb, float scale, int f16, void* stream_) {
    return attn_fwd_impl(q, q_batch_stride, q_row_stride, k_self, v_self, kv_self_batch_stride, kv_self_row_stride, k_bank,
                         v_bank, kv_bank_batch_stride, kv_bank_row_stride, o, o_batch_stride, o_row_stride, B, heads, Lq, Ls,
                         Lb, scale, f16, nullptr, stream_);
}

int dfw_attn_kvfused_fwd_lse(const void* q, long long q_batch_stride, int q_row_stride, const void* k_self,
                             const void* v_self, long long kv_self_batch_stride, int kv_self_row_stride,
                             const void* k_bank, const void* v_bank, long long kv_bank_batch_stride,
                             int kv_bank_row_stride, void* o, long long o_batch_stride, int o_row_stride, int B,
                             int heads, int Lq, int Ls, int Lb, float scale, int f16, float* lse, void* stream_) {
    if (lse == nullptr) return DFW_ERR_INVALID;
    return attn_fwd_impl(q, q_batch_stride, q_row_stride, k_self, v_self, kv_self_batch_stride, kv_self_row_stride, k_bank,
                         v_bank, kv_bank_batch_stride, kv_bank_row_stride, o, o_batch_stride, o_row_stride, B, heads, Lq, Ls,
                         Lb, scale, f16, lse, stream_);
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
    DFW_CHECK_CUDA(launch_k(cross_attn_kernel, static_cast<unsigned int>(blocks), 256, 0, static_cast<cudaStream_t>(stream_),
                            reinterpret_cast<const uint16_t*>(q), reinterpret_cast<const uint16_t*>(k),
                            reinterpret_cast<const uint16_t*>(v), kv_batch_stride, reinterpret_cast<uint16_t*>(o), groups, L,
                            heads, Lctx, scale * 1.4426950408889634f, f16));
    g_launches.fetch_add(1);
    DFW_CHECK_CUDA(cudaGetLastError());
    return DFW_OK;
}

int dfw_cross_attn_collapsed(const float* logits, int ld_logits, const float* U, const float* bias, const void* residual,
                             void* out, int dtype, long long M, int C, int heads, int Lctx, void* stream_) {
    using namespace dfw;
    int rc = require_sm100();
    if (rc != DFW_OK) return rc;
    DFW_REQUIRE(logits && U && bias && out && M > 0 && C > 0 && C % 8 == 0 && heads > 0 && Lctx > 0);
    DFW_REQUIRE(heads * Lctx <= 96 && Lctx <= 8 && ld_logits >= heads * Lctx && dtype >= 0 && dtype <= 2);
    DFW_REQUIRE((M + XC_TOK - 1) / XC_TOK <= 65535);
    cudaStream_t st = static_cast<cudaStream_t>(stream_);
    const dim3 grid((C / 8 + XC_OCT - 1) / XC_OCT, static_cast<unsigned>((M + XC_TOK - 1) / XC_TOK));
    if (dtype == 1) cross_attn_collapsed_kernel<1><<<grid, 256, 0, st>>>(logits, ld_logits, U, bias, residual, out, M, C, heads, Lctx);
    else if (dtype == 2) cross_attn_collapsed_kernel<2><<<grid, 256, 0, st>>>(logits, ld_logits, U, bias, residual, out, M, C, heads, Lctx);
    else cross_attn_collapsed_kernel<0><<<grid, 256, 0, st>>>(logits, ld_logits, U, bias, residual, out, M, C, heads, Lctx);
    g_launches.fetch_add(1);
    DFW_CHECK_CUDA(cudaGetLastError());
    return DFW_OK;
}

}  // extern "C"
