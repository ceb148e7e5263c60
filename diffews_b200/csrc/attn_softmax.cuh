// Softmax inner loop of the KV-fused attention kernel (attn.cu, v3), shared with scripts/microbench/softmax_rate.cu.
//
// One thread owns one query row (= one TMEM lane) and the 128 logits of a key tile, held in registers for both passes
// (so the S accumulator in TMEM is free again right after the load).  Head dim 64 makes the exponentials the bound:
// at full tensor rate a 128 x 128 tile takes 512 cycles, the 16 384 exponentials take 1024 on the SM's 16-lane XU.
// So a compile-time fraction of the exponentials (NPOLY of every 8 packed pairs) is evaluated on the FMA pipe instead:
//   2^x = 2^n * p(f),  n = round(x), f = x - n in [-0.5, 0.5],  p = degree-3 minimax (max relative error 7.5e-5, below
//   the 4.9e-4 half-ulp of the fp16 P it is rounded to), n added into the exponent field with one integer add-shift.
// Packed fp32x2 instructions (FFMA2 / FADD2, sm_100) carry the scale-and-subtract, the range reduction, the Horner steps
// and the row sum: two logits per issue slot.
#pragma once
#include "ptx.cuh"

namespace dfw {

__device__ __forceinline__ uint32_t cvt_f16x2(float lo, float hi) {      // one F2FP: {hi, lo} -> packed fp16
    uint32_t r;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
__device__ __forceinline__ uint32_t cvt_bf16x2(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}

constexpr float EXP2_C0 = 0.9999280571937561f, EXP2_C1 = 0.6932609677314758f, EXP2_C2 = 0.2426111251115799f,
                EXP2_C3 = 0.055171653628349304f;
constexpr float EXP2_MAGIC = 12582912.0f;        // 1.5 * 2^23: x + MAGIC rounds x to the nearest integer (|x| < 2^22)

// 2^x for a packed pair on the FMA / ALU pipes (no MUFU).  x is clamped to [-126, 126]: below, the result (~1e-38) is 0 in
// fp16 anyway; above, 2^126 still drives the row sum over SOFTMAX_TRIGGER (an unclamped x >= 129 would wrap the exponent
// field into the sign bit and yield a tiny NEGATIVE value that the trigger cannot see).
__device__ __forceinline__ void exp2_poly2(uint64_t x2, float& e0, float& e1) {
    float x0, x1;
    f2_unpack(x2, x0, x1);
    x0 = fminf(fmaxf(x0, -126.0f), 126.0f);
    x1 = fminf(fmaxf(x1, -126.0f), 126.0f);
    const uint64_t xc = f2_pack(x0, x1);
    const uint64_t t2 = f2_add(xc, f2_pack(EXP2_MAGIC, EXP2_MAGIC));
    const uint64_t n2 = f2_add(t2, f2_pack(-EXP2_MAGIC, -EXP2_MAGIC));
    const uint64_t f2 = f2_fma(n2, f2_pack(-1.0f, -1.0f), xc);
    uint64_t p2 = f2_fma(f2_pack(EXP2_C3, EXP2_C3), f2, f2_pack(EXP2_C2, EXP2_C2));
    p2 = f2_fma(p2, f2, f2_pack(EXP2_C1, EXP2_C1));
    p2 = f2_fma(p2, f2, f2_pack(EXP2_C0, EXP2_C0));
    float t0, t1, p0, p1;
    f2_unpack(t2, t0, t1);
    f2_unpack(p2, p0, p1);
    e0 = __int_as_float(__float_as_int(p0) + (__float_as_int(t0) << 23));
    e1 = __int_as_float(__float_as_int(p1) + (__float_as_int(t1) << 23));
}

// Row max of 128 raw logits (3-input max, four independent chains).
__device__ __forceinline__ float row_max128(const uint32_t (&s)[128]) {
    float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
    for (int i = 0; i < 128; i += 8) {
        m0 = fmax3(m0, __uint_as_float(s[i]), __uint_as_float(s[i + 1]));
        m1 = fmax3(m1, __uint_as_float(s[i + 2]), __uint_as_float(s[i + 3]));
        m2 = fmax3(m2, __uint_as_float(s[i + 4]), __uint_as_float(s[i + 5]));
        m3 = fmax3(m3, __uint_as_float(s[i + 6]), __uint_as_float(s[i + 7]));
    }
    return fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
}

// p = 2^(s * sc - mu) for 32 logits -> 16 packed 16-bit pairs in pk[], row sum accumulated in sum2 (two packed
// accumulators).  NPOLY of every 16 pairs take the polynomial path, spread evenly over the 16.
template <bool F16, int NPOLY>
__device__ __forceinline__ void exp_chunk32(const uint32_t* s, uint64_t sc2, uint64_t nmu2, uint64_t (&sum2)[2],
                                            uint32_t* pk) {
#pragma unroll
    for (int pi = 0; pi < 16; ++pi) {
        const uint64_t x2 = f2_fma(f2_pack(__uint_as_float(s[2 * pi]), __uint_as_float(s[2 * pi + 1])), sc2, nmu2);
        float e0, e1;
        // pair pi takes the polynomial when floor((pi + 1) * NPOLY / 16) > floor(pi * NPOLY / 16)
        const bool poly = ((pi + 1) * NPOLY) / 16 > (pi * NPOLY) / 16;
        if (poly) {
            exp2_poly2(x2, e0, e1);
        } else {
            float x0, x1;
            f2_unpack(x2, x0, x1);
            e0 = ex2_approx(x0);
            e1 = ex2_approx(x1);
        }
        sum2[pi & 1] = f2_add(sum2[pi & 1], f2_pack(e0, e1));
        pk[pi] = F16 ? cvt_f16x2(e0, e1) : cvt_bf16x2(e0, e1);
    }
}

// Staged variant (same results): all scale-and-subtracts, then every MUFU of the chunk back to back, then the polynomial
// pairs (FMA-pipe work that runs while the XU drains), then the sums and conversions -- the consumers of a MUFU result
// sit as far from its issue as the chunk allows.  NACC packed sum accumulators.
template <bool F16, int NPOLY, int NACC>
__device__ __forceinline__ void exp_chunk32_staged(const uint32_t* s, uint64_t sc2, uint64_t nmu2, uint64_t (&sum2)[NACC],
                                                   uint32_t* pk) {
    uint64_t x2[16];
    float e[32];
#pragma unroll
    for (int pi = 0; pi < 16; ++pi)
        x2[pi] = f2_fma(f2_pack(__uint_as_float(s[2 * pi]), __uint_as_float(s[2 * pi + 1])), sc2, nmu2);
#pragma unroll
    for (int pi = 0; pi < 16; ++pi) {
        const bool poly = ((pi + 1) * NPOLY) / 16 > (pi * NPOLY) / 16;
        if (!poly) {
            float x0, x1;
            f2_unpack(x2[pi], x0, x1);
            e[2 * pi] = ex2_approx(x0);
            e[2 * pi + 1] = ex2_approx(x1);
        }
    }
#pragma unroll
    for (int pi = 0; pi < 16; ++pi) {
        const bool poly = ((pi + 1) * NPOLY) / 16 > (pi * NPOLY) / 16;
        if (poly) exp2_poly2(x2[pi], e[2 * pi], e[2 * pi + 1]);
    }
#pragma unroll
    for (int pi = 0; pi < 16; ++pi) {
        sum2[pi % NACC] = f2_add(sum2[pi % NACC], f2_pack(e[2 * pi], e[2 * pi + 1]));
        pk[pi] = F16 ? cvt_f16x2(e[2 * pi], e[2 * pi + 1]) : cvt_bf16x2(e[2 * pi], e[2 * pi + 1]);
    }
}

template <bool F16, int NPOLY, int NACC>
__device__ __forceinline__ float exp_row128_staged(const uint32_t (&s)[128], float sc, float mu, uint32_t (&pk)[64]) {
    const uint64_t sc2 = f2_pack(sc, sc), nmu2 = f2_pack(-mu, -mu);
    uint64_t sum2[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) sum2[i] = 0ull;
#pragma unroll
    for (int c = 0; c < 4; ++c) exp_chunk32_staged<F16, NPOLY, NACC>(&s[c * 32], sc2, nmu2, sum2, &pk[c * 16]);
    float tot = 0.f;
#pragma unroll
    for (int i = 0; i < NACC; ++i) { float a, b; f2_unpack(sum2[i], a, b); tot += a + b; }
    return tot;
}

// All 128 logits of a row: pk[64] packed probabilities, returns the row sum of this tile.
template <bool F16, int NPOLY>
__device__ __forceinline__ float exp_row128(const uint32_t (&s)[128], float sc, float mu, uint32_t (&pk)[64]) {
    const uint64_t sc2 = f2_pack(sc, sc), nmu2 = f2_pack(-mu, -mu);
    uint64_t sum2[2] = {0ull, 0ull};
#pragma unroll
    for (int c = 0; c < 4; ++c) exp_chunk32<F16, NPOLY>(&s[c * 32], sc2, nmu2, sum2, &pk[c * 16]);
    float a, b, c, d;
    f2_unpack(sum2[0], a, b);
    f2_unpack(sum2[1], c, d);
    return (a + b) + (c + d);
}

// The running row maximum is NOT recomputed per tile.  P = 2^(s*c - m_used) is evaluated against the maximum in use;
// as long as the tile's row sum stays <= SOFTMAX_TRIGGER every p is <= SOFTMAX_TRIGGER (finite in fp16, exact scaling in
// the fp32 accumulators), so nothing needs to change.  Only when some row of the warp exceeds it (or on the first tile,
// m_used = -inf -> sum = +inf) the warp takes the slow path: exact row max, new m_used, O / l rescale, exponentials redone.
#ifndef DFW_SOFTMAX_TRIGGER
#define DFW_SOFTMAX_TRIGGER 1024.0f
#endif
constexpr float SOFTMAX_TRIGGER = DFW_SOFTMAX_TRIGGER;

// 16 consecutive 32-bit TMEM columns <- registers (thread i of the warp writes lane base + i)
__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
          "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
}

// Fast path over one key tile with the TMEM loads software-pipelined in 32-column chunks (the load of chunk c+1 is in
// flight while chunk c is exponentiated; tcgen05.wait::ld waits for everything outstanding, hence the placement).
// Returns the tile's row sum; pk[64] = packed probabilities.
template <bool F16, int NPOLY>
__device__ __forceinline__ float exp_row128_tmem(uint32_t ts, float sc, float mu, uint32_t (&pk)[64]) {
    const uint64_t sc2 = f2_pack(sc, sc), nmu2 = f2_pack(-mu, -mu);
    uint64_t sum2[2] = {0ull, 0ull};
    uint32_t va[32], vb[32];
    tmem_ld_32x32(ts, va);
    tmem_ld_wait(); tmem_regs_ready(va);
    tmem_ld_32x32(ts + 32, vb);
    exp_chunk32<F16, NPOLY>(va, sc2, nmu2, sum2, &pk[0]);
    tmem_ld_wait(); tmem_regs_ready(vb);
    tmem_ld_32x32(ts + 64, va);
    exp_chunk32<F16, NPOLY>(vb, sc2, nmu2, sum2, &pk[16]);
    tmem_ld_wait(); tmem_regs_ready(va);
    tmem_ld_32x32(ts + 96, vb);
    exp_chunk32<F16, NPOLY>(va, sc2, nmu2, sum2, &pk[32]);
    tmem_ld_wait(); tmem_regs_ready(vb);
    exp_chunk32<F16, NPOLY>(vb, sc2, nmu2, sum2, &pk[48]);
    float a, b, c, d;
    f2_unpack(sum2[0], a, b);
    f2_unpack(sum2[1], c, d);
    return (a + b) + (c + d);
}

// 128 consecutive fp32 TMEM columns -> registers: four x32 loads, tcgen05.wait::ld, and the compiler-level fence that
// keeps arithmetic on the destination registers below the wait (the loads are asynchronous: without the fence nvcc is
// free to schedule a consumer between the load and the wait -- a timing-dependent read of stale registers, seen as
// sporadic NaN rows when the exact path ran after a fast attempt).
__device__ __forceinline__ void tmem_ld_row128(uint32_t taddr, uint32_t (&s)[128]) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        uint32_t (&v)[32] = *reinterpret_cast<uint32_t (*)[32]>(&s[c * 32]);
        tmem_ld_32x32(taddr + c * 32, v);
    }
    tmem_ld_wait();
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        uint32_t (&v)[32] = *reinterpret_cast<uint32_t (*)[32]>(&s[c * 32]);
        tmem_regs_ready(v);
    }
}

// ---- generic tile widths (v4 kernel: 96-key tiles = three 32-column chunks) ------------------------------------------
template <int N>
__device__ __forceinline__ float row_max_n(const uint32_t (&s)[N]) {
    static_assert(N % 8 == 0, "row_max_n");
    float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
    for (int i = 0; i < N; i += 8) {
        m0 = fmax3(m0, __uint_as_float(s[i]), __uint_as_float(s[i + 1]));
        m1 = fmax3(m1, __uint_as_float(s[i + 2]), __uint_as_float(s[i + 3]));
        m2 = fmax3(m2, __uint_as_float(s[i + 4]), __uint_as_float(s[i + 5]));
        m3 = fmax3(m3, __uint_as_float(s[i + 6]), __uint_as_float(s[i + 7]));
    }
    return fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
}

template <bool F16, int NPOLY, int NACC, int NCH>
__device__ __forceinline__ float exp_row_staged_n(const uint32_t (&s)[NCH * 32], float sc, float mu, uint32_t (&pk)[NCH * 16]) {
    const uint64_t sc2 = f2_pack(sc, sc), nmu2 = f2_pack(-mu, -mu);
    uint64_t sum2[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) sum2[i] = 0ull;
#pragma unroll
    for (int c = 0; c < NCH; ++c) exp_chunk32_staged<F16, NPOLY, NACC>(&s[c * 32], sc2, nmu2, sum2, &pk[c * 16]);
    float tot = 0.f;
#pragma unroll
    for (int i = 0; i < NACC; ++i) { float a, b; f2_unpack(sum2[i], a, b); tot += a + b; }
    return tot;
}

// Fast path, NCH chunks of 32 columns, TMEM loads software-pipelined one chunk ahead (see exp_row128_tmem).
template <bool F16, int NPOLY, int NCH>
__device__ __forceinline__ float exp_row_tmem_n(uint32_t ts, float sc, float mu, uint32_t (&pk)[NCH * 16]) {
    const uint64_t sc2 = f2_pack(sc, sc), nmu2 = f2_pack(-mu, -mu);
    uint64_t sum2[2] = {0ull, 0ull};
    uint32_t va[32], vb[32];
    tmem_ld_32x32(ts, va);
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        uint32_t (&cur)[32] = (c & 1) ? vb : va;
        uint32_t (&nxt)[32] = (c & 1) ? va : vb;
        tmem_ld_wait(); tmem_regs_ready(cur);
        if (c + 1 < NCH) tmem_ld_32x32(ts + (c + 1) * 32, nxt);
        exp_chunk32<F16, NPOLY>(cur, sc2, nmu2, sum2, &pk[c * 16]);
    }
    float a, b, c2, d;
    f2_unpack(sum2[0], a, b);
    f2_unpack(sum2[1], c2, d);
    return (a + b) + (c2 + d);
}

template <int NCH>
__device__ __forceinline__ void tmem_ld_row_n(uint32_t taddr, uint32_t (&s)[NCH * 32]) {
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        uint32_t (&v)[32] = *reinterpret_cast<uint32_t (*)[32]>(&s[c * 32]);
        tmem_ld_32x32(taddr + c * 32, v);
    }
    tmem_ld_wait();
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        uint32_t (&v)[32] = *reinterpret_cast<uint32_t (*)[32]>(&s[c * 32]);
        tmem_regs_ready(v);
    }
}

}  // namespace dfw
