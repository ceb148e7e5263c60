// K3/K6 — tcgen05 implicit-GEMM convolution + Linear for sm_100a.
//
// One persistent CTA per SM, warp-specialised:
//   warp 0      TMA producer: per (filter tap, 64-channel block) one 4-D box of the NHWC activation tensor
//               (128 output pixels x 64 channels; halo / zero padding comes from TMA out-of-bounds fill) and one
//               2-D box of the [Cout, taps*Cin] weight matrix, both SWIZZLE_128B, into a STAGES-deep smem ring.
//   warp 1      MMA issuer: tcgen05.mma (cta_group::1, kind::f16, M=128, N=BLOCK_N, K=16) x4 per ring slot,
//               fp32 accumulators in TMEM, double-buffered so the epilogue of tile i overlaps the mainloop of i+1.
//   warp 2      TMEM allocator.
//   warps 4-7   epilogue: tcgen05.ld (thread = accumulator row), + bias (+ per-image time-embedding bias),
//               * scale, SiLU / GEGLU, + residual, convert, 16-byte stores to NHWC.
//
// Replaces the cuDNN / cuBLAS calls behind every nn.Conv2d / nn.Linear on the reference path
// (ref: diffews/models/unet_2d_condition.py:1118-1121,1161,1191,1226,1249 and the diffusers-0.25 blocks they
//  reach; diffews/marigold_pipeline_rgb_latent_noise.py:852-853,901-902 for the VAE).
#include <atomic>
#include <cstdlib>
#include <type_traits>

#include "common.cuh"
#include "ptx.cuh"

#ifndef DFW_GNIN_DBG
#define DFW_GNIN_DBG 0          // ablation switches of the GN_IN transform (scripts/build_variant.sh); 0 in the product
#endif

#ifndef T128_A_SLOTS
// patch ring / weight-tile ring depth of igemm_t128_kernel (36 KiB / 16 KiB per slot).  A patch feeds 1536 cycles of MMAs,
// a weight tile 512: two patches cover the load latency, three weight tiles do not (the MMA warp waited ~1000-1700 cycles
// per 64-channel block on w_full).  2 + 5 against 3 + 3: 512->512 at 128^2 1504 -> 1596 TFLOP/s, 64^2 1581 -> 1615, the
// 128- and 256-channel layers unchanged (profiles/r02_t128_ring_depth.log).
#define T128_A_SLOTS 2
#define T128_W_SLOTS 5
#endif
#ifndef DFW_GNIN_TRACE
#define DFW_GNIN_TRACE 0        // 1: CTA 0 of the GN_IN kernel accumulates phase cycle counts (dfw_debug_t128_trace)
#endif
#if DFW_GNIN_TRACE
__device__ long long g_t128_trace[16];
#define TR_T0() const long long tr_t0_ = clock64()
#define TR_ADD(var) var += clock64() - tr_t0_
#else
#define TR_T0()
#define TR_ADD(var)
#endif

namespace dfw {
extern std::atomic<long long> g_launches;

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;
constexpr int A_TILE_BYTES = BLOCK_M * BLOCK_K * 2;  // 16 KiB
constexpr int IGEMM_THREADS = 384;      // warps: 0 TMA, 1 MMA, 2 TMEM alloc, 3 residual loader, 4-11 two epilogue groups
constexpr int GN_SLOTS_PER_CTA = 2;     // one GroupNorm partial slot per epilogue group
constexpr int MAX_TAPS = 9;

struct IgemmMaps {
    CUtensorMap a[4];
    CUtensorMap b;
    CUtensorMap out;   // TMA-store epilogue: y   as [out_ch, W, H, N], box [64 B, TW, TH, TN], SWIZZLE_64B
    CUtensorMap res;   //                     residual, same geometry
};

struct IgemmParams {
    int N, H, W;      // output images / height / width (GEMM rows = N*H*W)
    int Cout;         // weight rows
    int kb_per_tap;   // Cin / 64
    int ntaps;
    int tap_map[MAX_TAPS];
    int tap_dh[MAX_TAPS];
    int tap_dw[MAX_TAPS];
    int TW, TH, TN;   // output tile = TN images x TH rows x TW cols = 128 pixels
    int tiles_w, tiles_h, tiles_nimg;
    int n_tiles;      // Cout tiles
    int total_tiles;
    const float* bias;
    int bias_sample_stride;
    const void* residual;
    void* out;
    int out_ch;       // channels of y
    float out_scale;
    int flags;
    int w_batched;    // 1: per-image weights (batched GEMM): the B tensor map's 3rd coordinate is the image index
    float* gn_partial; // optional: per-(image, CTA, group) partial (sum, sumsq) of y for the GroupNorm that consumes y
    int gn_cpg;        // channels per group of that GroupNorm (must divide the chunk width; TN == 1)
    long long gn_img_stride;   // floats between consecutive images in gn_partial (slot of this CTA: + blockIdx.x * 64)
    int halo;         // 1: 3x3 stride-1 "halo" mainloop (vertical taps reuse one (TH+2) x TW patch per horizontal offset)
    int tma_epi;      // 1: epilogue stages 64-byte-wide column chunks in smem and uses TMA stores / residual TMA loads
    int has_res;
    // weight-stationary mode (1x1 / linear with a short K): the kb_per_tap weight tiles of the current Cout tile stay in
    // smem while the CTA walks the pixel tiles (tiles ordered Cout-tile-major), the ring carries activation tiles only
    int b_resident;
    int a_slots;      // ring depth in that mode (<= STAGES, <= kb_per_tap)
    int m_tiles;      // pixel tiles (tiles_w * tiles_h * tiles_nimg)
};

// TPU = output tiles per work unit.  TPU = 2 ("paired tiles", only when the layer has a single Cout tile) lets two
// consecutive 128-pixel tiles share every weight tile: 48 KiB of operands per 2 x (128x128x64) MMAs instead of
// 2 x 32 KiB, which is what bounds the Cout = 128 layers (L2 -> smem operand traffic, not the tensor pipe).
template <int BLOCK_N, int TPU = 1>
struct IgemmCfg {
    static constexpr int B_TILE_BYTES = BLOCK_N * 128;
    static constexpr int STAGE_BYTES = TPU * A_TILE_BYTES + B_TILE_BYTES;
    static constexpr int STAGES = (STAGE_BYTES >= 48 * 1024) ? 4 : (STAGE_BYTES > 32 * 1024 ? 5 : (STAGE_BYTES == 32 * 1024 ? 6 : 8));
    // epilogue smem, shared by the two epilogue flavours: TMA path 2 x 8 KiB out + 2 x 8 KiB residual chunk buffers;
    // direct path 4 x 4.5 KiB transpose staging + 2 KiB row table
    static constexpr int EPI_BYTES = 4 * 8192;
    static constexpr int SUB_COLS = BLOCK_N <= 32 ? 32 : (BLOCK_N <= 64 ? 64 : (BLOCK_N <= 128 ? 128 : 256));
    static constexpr int ACC_COLS = TPU * SUB_COLS;          // TMEM columns of one work unit's accumulators
    static constexpr int TMEM_COLS = 2 * ACC_COLS;
    static_assert(TMEM_COLS <= 512, "TMEM budget");
    static constexpr int GN_BYTES = 8 * 64 * 4;      // per-epilogue-warp GroupNorm partial sums (32 groups x 2)
    // "halo" mainloop for 3x3 / stride-1 convolutions (tile = 8 rows x 16 cols): the operand area is split into an
    // A ring of (8+2) x 16-pixel patches (one per horizontal filter offset; the 3 vertical taps read the SAME patch at
    // row offsets 0 / 16 / 32, which stay 1024-byte aligned, so the SWIZZLE_128B phase is preserved) and a B ring.
    static constexpr int PATCH_BYTES = (8 + 2) * 16 * 128;              // 20 KiB
    static constexpr int HALO_A_SLOTS = (TPU == 2) ? 2 : ((BLOCK_N >= 160) ? 3 : 4);
    static constexpr int HALO_B_RAW = (STAGES * STAGE_BYTES - HALO_A_SLOTS * TPU * PATCH_BYTES) / B_TILE_BYTES;
    static constexpr int HALO_B_SLOTS = HALO_B_RAW > 8 ? 8 : HALO_B_RAW;
    static_assert(HALO_B_SLOTS >= 3, "halo ring sizing");
    // the dynamic smem array is declared __align__(1024); 512 B of slack remain and the kernel traps if the base ever
    // needs more than that to reach the 1024-byte alignment SWIZZLE_128B wants
    static constexpr int SMEM_USED = STAGES * STAGE_BYTES + EPI_BYTES + GN_BYTES + 512 /*barriers*/;
    static constexpr int SMEM_BYTES = SMEM_USED + 512 /*align slack*/;
    static_assert(STAGE_BYTES % 1024 == 0, "stage must keep 1024B alignment for SWIZZLE_128B");
    static_assert(SMEM_BYTES <= 232448, "smem budget");
};

// Exact-GELU x * Phi(x) with erfc from Abramowitz & Stegun 7.1.28: erfc(z) = (1 + a1 z + ... + a6 z^6)^-16, |error| <= 3e-7
// (x * Phi(x): <= 8e-7 absolute, far below the 16-bit output rounding).  One MUFU (rcp) + 6 FMA + 4 squarings instead of
// erff()'s ~30 instructions or 7.1.26's rcp + ex2: a MUFU blocks the sub-partition's issue port for 8 cycles, and the
// GEGLU projections (K = 320 / 640) are bound by this epilogue, not by their MMAs.
__device__ __forceinline__ float gelu_erf(float x) {
    const float z = fabsf(x) * 0.70710678118654752f;
    float poly = fmaf(z, 0.0000430638f, 0.0002765672f);
    poly = fmaf(poly, z, 0.0001520143f);
    poly = fmaf(poly, z, 0.0092705272f);
    poly = fmaf(poly, z, 0.0422820123f);
    poly = fmaf(poly, z, 0.0705230784f);
    poly = fmaf(poly, z, 1.0f);
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(poly));
    r *= r; r *= r; r *= r; r *= r;                       // poly^-16 = erfc(z)
    const float hq = 0.5f * r;
    return x * (x >= 0.f ? 1.0f - hq : hq);
}
// Two GEGLU outputs at once, (val + bv) * gelu(gate + bg), on packed fp32x2 instructions: the same IEEE operations per lane
// as gelu_erf (bit-identical results), ~25 issue slots per pair instead of ~41 -- the K = 320 GEGLU projection spent 4000
// epilogue issue slots per sub-partition on a tile whose MMAs take 2560 cycles (ncu: tensor pipe 40 %, issue 57 %).
__device__ __forceinline__ void geglu_pair(float v0, float v1, float g0, float g1, float bv0, float bv1, float bg0, float bg1,
                                           float& o0, float& o1) {
    const uint64_t x2 = f2_add(f2_pack(g0, g1), f2_pack(bg0, bg1));
    float x0, x1;
    f2_unpack(x2, x0, x1);
    const uint64_t z2 = f2_mul(f2_pack(fabsf(x0), fabsf(x1)), f2_pack(0.70710678118654752f, 0.70710678118654752f));
    uint64_t p2 = f2_fma(z2, f2_pack(0.0000430638f, 0.0000430638f), f2_pack(0.0002765672f, 0.0002765672f));
    p2 = f2_fma(p2, z2, f2_pack(0.0001520143f, 0.0001520143f));
    p2 = f2_fma(p2, z2, f2_pack(0.0092705272f, 0.0092705272f));
    p2 = f2_fma(p2, z2, f2_pack(0.0422820123f, 0.0422820123f));
    p2 = f2_fma(p2, z2, f2_pack(0.0705230784f, 0.0705230784f));
    p2 = f2_fma(p2, z2, f2_pack(1.0f, 1.0f));
    float p0, p1, r0, r1;
    f2_unpack(p2, p0, p1);
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(p0));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(p1));
    uint64_t r2 = f2_pack(r0, r1);
    r2 = f2_mul(r2, r2); r2 = f2_mul(r2, r2); r2 = f2_mul(r2, r2); r2 = f2_mul(r2, r2);      // poly^-16 = erfc(z)
    const uint64_t hq2 = f2_mul(r2, f2_pack(0.5f, 0.5f));
    float h0, h1;
    f2_unpack(hq2, h0, h1);
    const uint64_t phi2 = f2_pack(x0 >= 0.f ? 1.0f - h0 : h0, x1 >= 0.f ? 1.0f - h1 : h1);
    const uint64_t ge2 = f2_mul(x2, phi2);
    const uint64_t val2 = f2_add(f2_pack(v0, v1), f2_pack(bv0, bv1));
    f2_unpack(f2_mul(val2, ge2), o0, o1);
}
__device__ __forceinline__ float silu(float x) { return x / (1.0f + __expf(-x)); }

// Butterfly reduce-scatter over the warp: on return v[0] of lane l holds the sum over all 32 lanes of element
// (l >> (5 - log2 NV)) of the input vectors.  NV + ... shuffles instead of 5 * NV for NV independent all-reduces.
template <int NV>
__device__ __forceinline__ float warp_reduce_scatter(float (&v)[NV], int lane) {
    int bit = 16;
#pragma unroll
    for (int half = NV / 2; half >= 1; half /= 2, bit >>= 1) {
        const bool upper = (lane & bit) != 0;
#pragma unroll
        for (int i = 0; i < half; ++i) {
            const float send = upper ? v[i] : v[i + half];
            const float keep = upper ? v[i + half] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
        }
    }
    float r = v[0];
    for (; bit >= 1; bit >>= 1) r += __shfl_xor_sync(0xffffffffu, r, bit);
    return r;
}

// GroupNorm partial statistics of one epilogue chunk: f[0..CW) are the final output values of this thread's row for
// channels [col0, col0 + CW); CPG channels per group.  Each warp owns 64 floats of smem (32 groups x (sum, sumsq)).
template <int CPG, int CW>
__device__ __forceinline__ void gn_chunk_stats(const float (&f)[32], bool row_valid, int col0, int lane, float* wacc) {
    constexpr int G = CW / CPG;
    constexpr int NV = 2 * G;
    float v[NV];
#pragma unroll
    for (int g = 0; g < G; ++g) {
        float s = 0.f, ss = 0.f;
#pragma unroll
        for (int i = 0; i < CPG; ++i) { const float x = f[g * CPG + i]; s += x; ss = fmaf(x, x, ss); }
        v[g] = row_valid ? s : 0.f;
        v[G + g] = row_valid ? ss : 0.f;
    }
    const float r = warp_reduce_scatter<NV>(v, lane);
    constexpr int LPV = 32 / NV;                       // lanes holding the same element
    if ((lane & (LPV - 1)) == 0) {
        const int idx = lane / LPV;                    // 0..NV-1 : [0,G) sums, [G,2G) sums of squares
        const int g = idx % G, k = idx / G;
        float* slot = wacc + ((col0 / CPG + g) * 2 + k);
        *slot += r;
    }
}

struct TileCoord {
    int n_tile, n0, h0, w0;
};
__device__ __forceinline__ TileCoord decode_tile(const IgemmParams& p, int t) {
    TileCoord c;
    int m;
    if (p.b_resident) { c.n_tile = t / p.m_tiles; m = t - c.n_tile * p.m_tiles; }     // Cout-tile-major
    else { c.n_tile = t % p.n_tiles; m = t / p.n_tiles; }
    c.w0 = (m % p.tiles_w) * p.TW;
    m /= p.tiles_w;
    c.h0 = (m % p.tiles_h) * p.TH;
    c.n0 = (m / p.tiles_h) * p.TN;
    return c;
}

// ---------------------------------------------------------------------------------------------------------
// Epilogue.  tcgen05.ld hands every thread one accumulator ROW (32 consecutive columns per load).  Writing rows
// straight to global would make every store instruction touch 32 different cache lines, so each warp transposes
// its 32x32 chunk through a padded smem staging buffer: afterwards 8 consecutive lanes own one row's 32 columns
// (4 columns each) and every global load/store instruction covers 4 complete rows of the chunk (4 x 128 B in fp32).
// ---------------------------------------------------------------------------------------------------------
constexpr int STG_LD = 36;                       // staging row pitch in floats (pad 4: conflict-free float4 access)
constexpr int STG_BYTES_PER_WARP = 32 * STG_LD * 4;

struct EpiRow {                                  // per accumulator row, computed once per tile
    long long pix;                               // output pixel index, -1 if the row is outside the tensor
    int img;
    int pad;
};

__device__ __forceinline__ void stage_rows(float* stg, int lane, const uint32_t (&v)[32]) {
    float4* dst = reinterpret_cast<float4*>(stg + lane * STG_LD);
#pragma unroll
    for (int j = 0; j < 8; ++j)
        dst[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                             __uint_as_float(v[4 * j + 3]));
}

// Raw residual values of one 32x32 chunk in the transposed ("8 lanes per row") layout, fetched BEFORE the chunk's
// accumulators are read so that the global-load latency overlaps the TMEM load / smem transpose of the chunk and the
// stores of the previous one (the loads must not be interleaved with the stores: out and residual may alias as far
// as the compiler knows, and an in-order LDG -> FADD -> STG chain per row group serialises on DRAM latency).
struct ResChunk {
    uint4 raw[8];      // fp32 residual: 4 floats; 16-bit residual: .x/.y hold 4 halves
};

__device__ __forceinline__ void prefetch_residual(const IgemmParams& p, const EpiRow* rows, int lane, int col_in0,
                                                  int col_out0, int NC, ResChunk& rc) {
    if (p.residual == nullptr) return;
    const int cq = (lane & 7) * 4, sub = lane >> 3;
    const int ncols = min(NC, p.Cout - col_in0);
    const bool vec = (p.out_ch % 4 == 0) && (ncols > 0) && (ncols % 4 == 0);
    if (!vec || cq >= ncols) return;     // the scalar tail path loads inline
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const long long pix = rows[4 * i + sub].pix;
        rc.raw[i] = make_uint4(0u, 0u, 0u, 0u);
        if (pix < 0) continue;
        const long long off = pix * p.out_ch + col_out0 + cq;
        if (p.flags & DFW_EPI_RES_F32) {
            rc.raw[i] = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const float*>(p.residual) + off));
        } else {
            const uint2 t = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const uint16_t*>(p.residual) + off));
            rc.raw[i].x = t.x; rc.raw[i].y = t.y;
        }
    }
}

// Processes one staged chunk of NC (<= 32) columns starting at weight row `col_in0`, output column `col_out0`.
__device__ __forceinline__ void epilogue_store_chunk(const IgemmParams& p, const float* stg, const EpiRow* rows,
                                                     int lane, int col_in0, int col_out0, int NC, const ResChunk& rc) {
    const int cq = (lane & 7) * 4;               // this lane's 4 columns inside the chunk
    const int sub = lane >> 3;                   // row within a group of 4
    const int f16 = p.flags & DFW_EPI_F16;       // 16-bit tensors are fp16 (else bf16)
    const int ncols = min(NC, p.Cout - col_in0); // valid columns in this chunk
    if (ncols <= 0) return;
    const bool vec = (p.out_ch % 4 == 0) && (ncols % 4 == 0);
    const bool lane_has_cols = cq < ncols;
    float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
    const bool shared_bias = (p.bias != nullptr) && (p.bias_sample_stride == 0);
    if (shared_bias && lane_has_cols) {
        if (vec) bias4 = __ldg(reinterpret_cast<const float4*>(p.bias + col_in0 + cq));
        else {
            const float* bp = p.bias + col_in0 + cq;
            bias4.x = __ldg(bp);
            if (cq + 1 < ncols) bias4.y = __ldg(bp + 1);
            if (cq + 2 < ncols) bias4.z = __ldg(bp + 2);
            if (cq + 3 < ncols) bias4.w = __ldg(bp + 3);
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int r = 4 * i + sub;
        const EpiRow row = rows[r];
        if (row.pix < 0 || !lane_has_cols) continue;
        float4 f = *reinterpret_cast<const float4*>(stg + r * STG_LD + cq);
        if (p.bias != nullptr) {
            if (shared_bias) { f.x += bias4.x; f.y += bias4.y; f.z += bias4.z; f.w += bias4.w; }
            else {
                const float* bp = p.bias + static_cast<long long>(row.img) * p.bias_sample_stride + col_in0 + cq;
                if (vec) { const float4 b = __ldg(reinterpret_cast<const float4*>(bp)); f.x += b.x; f.y += b.y; f.z += b.z; f.w += b.w; }
                else {
                    f.x += __ldg(bp);
                    if (cq + 1 < ncols) f.y += __ldg(bp + 1);
                    if (cq + 2 < ncols) f.z += __ldg(bp + 2);
                    if (cq + 3 < ncols) f.w += __ldg(bp + 3);
                }
            }
        }
        if (p.out_scale != 1.0f) { f.x *= p.out_scale; f.y *= p.out_scale; f.z *= p.out_scale; f.w *= p.out_scale; }
        if (p.flags & DFW_EPI_SILU) { f.x = silu(f.x); f.y = silu(f.y); f.z = silu(f.z); f.w = silu(f.w); }
        const long long off = row.pix * p.out_ch + col_out0 + cq;
        if (vec) {
            if (p.residual != nullptr) {
                if (p.flags & DFW_EPI_RES_F32) {
                    f.x += __uint_as_float(rc.raw[i].x); f.y += __uint_as_float(rc.raw[i].y);
                    f.z += __uint_as_float(rc.raw[i].z); f.w += __uint_as_float(rc.raw[i].w);
                } else {
                    const float2 r0 = unpack_h2(rc.raw[i].x, f16), r1 = unpack_h2(rc.raw[i].y, f16);
                    f.x += r0.x; f.y += r0.y; f.z += r1.x; f.w += r1.y;
                }
            }
            if (p.flags & DFW_EPI_OUT_F32) {
                __stcs(reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + off), f);
            } else {
                uint2 o;
                o.x = pack_h2(f.x, f.y, f16);
                o.y = pack_h2(f.z, f.w, f16);
                *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(p.out) + off) = o;
            }
        } else {
            const float fv[4] = {f.x, f.y, f.z, f.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (cq + j >= ncols) break;
                float val = fv[j];
                if (p.residual != nullptr) {
                    if (p.flags & DFW_EPI_RES_F32) val += reinterpret_cast<const float*>(p.residual)[off + j];
                    else if (f16) val += __half2float(reinterpret_cast<const __half*>(p.residual)[off + j]);
                    else val += __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p.residual)[off + j]);
                }
                if (p.flags & DFW_EPI_OUT_F32) reinterpret_cast<float*>(p.out)[off + j] = val;
                else if (f16) reinterpret_cast<__half*>(p.out)[off + j] = __float2half_rn(val);
                else reinterpret_cast<__nv_bfloat16*>(p.out)[off + j] = __float2bfloat16_rn(val);
            }
        }
    }
}

template <int BLOCK_N, int TPU>
__global__ void __launch_bounds__(IGEMM_THREADS, 1)
igemm_kernel(const __grid_constant__ IgemmMaps maps, const __grid_constant__ IgemmParams p) {
    using Cfg = IgemmCfg<BLOCK_N, TPU>;
    constexpr int STAGES = Cfg::STAGES;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t smem_raw_u32 = smem_u32(smem_raw);
    const uint32_t smem_base = (smem_raw_u32 + 1023u) & ~1023u;
    if (smem_base - smem_raw_u32 > static_cast<uint32_t>(Cfg::SMEM_BYTES - Cfg::SMEM_USED)) __trap();
    const uint32_t epi_base = smem_base + STAGES * Cfg::STAGE_BYTES;          // staging + row table
    const uint32_t gn_base = epi_base + Cfg::EPI_BYTES;
    const uint32_t bar_base = gn_base + Cfg::GN_BYTES;
    auto sA = [&](int s, int sub) { return smem_base + s * Cfg::STAGE_BYTES + sub * A_TILE_BYTES; };
    auto sB = [&](int s) { return smem_base + s * Cfg::STAGE_BYTES + TPU * A_TILE_BYTES; };
    const int units = (p.total_tiles + TPU - 1) / TPU;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
    auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + a); };
    auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + 2 + a); };
    auto res_full = [&](int b) { return bar_base + 8u * (2 * STAGES + 4 + b); };     // 4 epilogue chunk buffers
    auto buf_free = [&](int b) { return bar_base + 8u * (2 * STAGES + 8 + b); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 12);
    const uint32_t hbar = bar_base + 8u * (2 * STAGES + 13);
    auto ha_full = [&](int i) { return hbar + 8u * i; };
    auto ha_empty = [&](int i) { return hbar + 8u * (4 + i); };
    auto hb_full = [&](int i) { return hbar + 8u * (8 + i); };
    auto hb_empty = [&](int i) { return hbar + 8u * (16 + i); };
    auto hA = [&](int slot, int sub) { return smem_base + (slot * TPU + sub) * Cfg::PATCH_BYTES; };
    auto hB = [&](int slot) { return smem_base + Cfg::HALO_A_SLOTS * TPU * Cfg::PATCH_BYTES + slot * Cfg::B_TILE_BYTES; };
    volatile uint32_t* tmem_slot_ptr =
        reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_raw_u32));
    uint8_t* epi_generic = smem_raw + (epi_base - smem_raw_u32);
    // TMA epilogue: 4 chunk buffers of 8 KiB used IN PLACE — the residual chunk is TMA-loaded into a buffer, every
    // thread adds its accumulator row to its own row of the buffer and writes the result back, then the same buffer
    // is TMA-stored; a buffer is recycled when its store has finished reading it (4-deep ring for loads and stores).
    auto epi_buf = [&](int b) { return epi_base + 8192u * b; };
    // TMA-epilogue chunk geometry: 64 bytes of output per row -> 32 columns (16-bit y) or 16 columns (fp32 y)
    const int CW = (p.flags & DFW_EPI_OUT_F32) ? 16 : 32;
    const int out_cols_per_tile = (p.flags & DFW_EPI_GEGLU) ? BLOCK_N / 2 : BLOCK_N;
    const int chunks_per_tile = out_cols_per_tile / CW;

    // warp index through a shuffle so the compiler knows it is warp-uniform: the role loops below run on all 32 lanes
    // and only the tcgen05 / TMA instructions sit inside elect_one() regions (a plain `if (lane == 0)` around a whole
    // loop makes ptxas wrap every uniform-datapath instruction in an ELECT / BRA.U.ANY serialisation loop, which made
    // the issue of an N = 128 MMA slower than the MMA itself)
    const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        for (int i = 0; i < 4; ++i) tma_prefetch_desc(&maps.a[i]);
        tma_prefetch_desc(&maps.b);
        if (p.tma_epi) {
            tma_prefetch_desc(&maps.out);
            if (p.has_res) tma_prefetch_desc(&maps.res);
        }
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tfull_bar(a), 1);
            mbar_init(tempty_bar(a), p.tma_epi ? 8 : 4);
        }
        for (int a = 0; a < 4; ++a) {
            mbar_init(res_full(a), 1);
            mbar_init(buf_free(a), 1);
        }
        for (int i = 0; i < 4; ++i) { mbar_init(ha_full(i), 1); mbar_init(ha_empty(i), 1); }
        for (int i = 0; i < 8; ++i) { mbar_init(hb_full(i), 1); mbar_init(hb_empty(i), 1); }
        fence_mbar_init();
    }
    if (warp == 2) {
        tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    pdl_launch_dependents();
    pdl_wait();                      // everything above overlapped the previous kernel's tail; global memory from here on

    const int kblocks = p.ntaps * p.kb_per_tap;

    if (warp == 0 && p.halo) {
        {
            int as = 0, bs = 0;
            uint32_t aph = 0, bph = 0;
            for (int u = blockIdx.x; u < units; u += gridDim.x) {
                const int t0 = u * TPU;
                const int nsub = min(TPU, p.total_tiles - t0);
                TileCoord tc[TPU];
#pragma unroll
                for (int sub = 0; sub < TPU; ++sub) tc[sub] = decode_tile(p, min(t0 + sub, p.total_tiles - 1));
                for (int kb = 0; kb < p.kb_per_tap; ++kb) {
                    for (int dw = -1; dw <= 1; ++dw) {
                        mbar_wait(ha_empty(as), aph ^ 1u, 7);
                        if (elect_one()) {
                            mbar_arrive_expect_tx(ha_full(as), nsub * Cfg::PATCH_BYTES);
#pragma unroll
                            for (int sub = 0; sub < TPU; ++sub)
                                if (sub < nsub)
                                    tma_load_4d(hA(as, sub), &maps.a[1], ha_full(as), kb * BLOCK_K, tc[sub].w0 + dw,
                                                tc[sub].h0 - 1, tc[sub].n0);
                        }
                        __syncwarp();
                        if (++as == Cfg::HALO_A_SLOTS) { as = 0; aph ^= 1u; }
                        for (int dh = -1; dh <= 1; ++dh) {
                            const int tap = (dh + 1) * 3 + (dw + 1);
                            mbar_wait(hb_empty(bs), bph ^ 1u, 8);
                            if (elect_one()) {
                                mbar_arrive_expect_tx(hb_full(bs), Cfg::B_TILE_BYTES);
                                tma_load_3d(hB(bs), &maps.b, hb_full(bs), (tap * p.kb_per_tap + kb) * BLOCK_K,
                                            tc[0].n_tile * BLOCK_N, 0);
                            }
                            __syncwarp();
                            if (++bs == Cfg::HALO_B_SLOTS) { bs = 0; bph ^= 1u; }
                        }
                    }
                }
            }
        }
    } else if (warp == 0 && TPU == 1 && p.b_resident) {
        // weight-stationary producer: B tiles reloaded only when the Cout tile changes (after the MMAs of every earlier unit
        // have drained: b_done completes once per unit, and with a_slots <= kb_per_tap this warp is never more than one
        // unit ahead of the MMA warp, so the parity wait is unambiguous), A tiles through the ring
        int slot = 0, cur_n = -1;
        uint32_t phase = 0, nunit = 0;
        const uint32_t b_full = ha_full(0), b_done = ha_empty(0);
        for (int u = blockIdx.x; u < units; u += gridDim.x, ++nunit) {
            const TileCoord tc = decode_tile(p, u);
            if (tc.n_tile != cur_n) {
                if (nunit > 0) mbar_wait(b_done, (nunit - 1) & 1u, 30);
                if (elect_one()) {
                    mbar_arrive_expect_tx(b_full, p.kb_per_tap * Cfg::B_TILE_BYTES);
                    for (int kb = 0; kb < p.kb_per_tap; ++kb)
                        tma_load_3d(smem_base + kb * Cfg::B_TILE_BYTES, &maps.b, b_full, kb * BLOCK_K, tc.n_tile * BLOCK_N, 0);
                }
                __syncwarp();
                cur_n = tc.n_tile;
            }
            const CUtensorMap* am = &maps.a[p.tap_map[0]];
            for (int kb = 0; kb < p.kb_per_tap; ++kb) {
                mbar_wait(empty_bar(slot), phase ^ 1u, 1);
                if (elect_one()) {
                    mbar_arrive_expect_tx(full_bar(slot), A_TILE_BYTES);
                    tma_load_4d(smem_base + p.kb_per_tap * Cfg::B_TILE_BYTES + slot * A_TILE_BYTES, am, full_bar(slot),
                                kb * BLOCK_K, tc.w0 + p.tap_dw[0], tc.h0 + p.tap_dh[0], tc.n0);
                }
                __syncwarp();
                if (++slot == p.a_slots) { slot = 0; phase ^= 1u; }
            }
        }
    } else if (warp == 0) {
        {
            int stage = 0;
            uint32_t phase = 0;
            for (int u = blockIdx.x; u < units; u += gridDim.x) {
                const int t0 = u * TPU;
                const int nsub = min(TPU, p.total_tiles - t0);
                TileCoord tc[TPU];
#pragma unroll
                for (int sub = 0; sub < TPU; ++sub) tc[sub] = decode_tile(p, min(t0 + sub, p.total_tiles - 1));
                for (int tap = 0; tap < p.ntaps; ++tap) {
                    const CUtensorMap* am = &maps.a[p.tap_map[tap]];
                    for (int kb = 0; kb < p.kb_per_tap; ++kb) {
                        mbar_wait(empty_bar(stage), phase ^ 1u, 1);
                        if (elect_one()) {
                            mbar_arrive_expect_tx(full_bar(stage), nsub * A_TILE_BYTES + Cfg::B_TILE_BYTES);
#pragma unroll
                            for (int sub = 0; sub < TPU; ++sub)
                                if (sub < nsub)
                                    tma_load_4d(sA(stage, sub), am, full_bar(stage), kb * BLOCK_K,
                                                tc[sub].w0 + p.tap_dw[tap], tc[sub].h0 + p.tap_dh[tap], tc[sub].n0);
                            tma_load_3d(sB(stage), &maps.b, full_bar(stage), (tap * p.kb_per_tap + kb) * BLOCK_K,
                                        tc[0].n_tile * BLOCK_N, p.w_batched ? tc[0].n0 : 0);
                        }
                        __syncwarp();
                        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                    }
                }
            }
        }
    } else if (warp == 1 && p.halo) {
        {
            const uint32_t fmt = (p.flags & DFW_EPI_F16) ? 0u : 1u;
            const uint32_t idesc = umma_idesc(BLOCK_M, BLOCK_N, fmt, fmt, 0);
            int as = 0, bs = 0;
            uint32_t aph = 0, bph = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int u = blockIdx.x; u < units; u += gridDim.x) {
                const int nsub = min(TPU, p.total_tiles - u * TPU);
                mbar_wait(tempty_bar(acc), acc_phase ^ 1u, 2);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * Cfg::ACC_COLS;
                bool first = true;
                for (int kb = 0; kb < p.kb_per_tap; ++kb) {
                    for (int dw = 0; dw < 3; ++dw) {
                        mbar_wait(ha_full(as), aph, 9);
                        for (int dh = 0; dh < 3; ++dh) {
                            mbar_wait(hb_full(bs), bph, 3);
                            tc_fence_after();
                            if (elect_one()) {
                                const uint64_t bdesc = umma_desc_sw128(hB(bs));
#pragma unroll
                                for (int sub = 0; sub < TPU; ++sub) {
                                    if (sub < nsub) {
                                        // vertical tap dh reads patch rows [16*dh, 16*dh + 128): + dh * 2048 bytes
                                        const uint64_t adesc = umma_desc_sw128(hA(as, sub) + dh * 16 * 128);
#pragma unroll
                                        for (int k = 0; k < BLOCK_K / 16; ++k)
                                            umma_ss(d_tmem + sub * Cfg::SUB_COLS, adesc + 2u * k, bdesc + 2u * k, idesc,
                                                    (!first || k > 0) ? 1u : 0u);
                                    }
                                }
                                tc_commit(hb_empty(bs));
                                if (dh == 2) tc_commit(ha_empty(as));
                                if (dh == 2 && dw == 2 && kb == p.kb_per_tap - 1) tc_commit(tfull_bar(acc));
                            }
                            __syncwarp();
                            first = false;
                            if (++bs == Cfg::HALO_B_SLOTS) { bs = 0; bph ^= 1u; }
                        }
                        if (++as == Cfg::HALO_A_SLOTS) { as = 0; aph ^= 1u; }
                    }
                }
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1u;
            }
        }
    } else if (warp == 1 && TPU == 1 && p.b_resident) {
        const uint32_t fmt = (p.flags & DFW_EPI_F16) ? 0u : 1u;
        const uint32_t idesc = umma_idesc(BLOCK_M, BLOCK_N, fmt, fmt, 0);
        const uint32_t b_full = ha_full(0), b_done = ha_empty(0);
        int slot = 0, acc = 0, cur_n = -1;
        uint32_t phase = 0, acc_phase = 0, bphase = 0;
        for (int u = blockIdx.x; u < units; u += gridDim.x) {
            const int n_tile = p.b_resident ? u / p.m_tiles : 0;
            mbar_wait(tempty_bar(acc), acc_phase ^ 1u, 2);
            if (n_tile != cur_n) {
                mbar_wait(b_full, bphase, 31);
                bphase ^= 1u;
                cur_n = n_tile;
            }
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * Cfg::ACC_COLS;
            for (int kb = 0; kb < p.kb_per_tap; ++kb) {
                mbar_wait(full_bar(slot), phase, 3);
                tc_fence_after();
                if (elect_one()) {
                    const uint64_t adesc = umma_desc_sw128(smem_base + p.kb_per_tap * Cfg::B_TILE_BYTES + slot * A_TILE_BYTES);
                    const uint64_t bdesc = umma_desc_sw128(smem_base + kb * Cfg::B_TILE_BYTES);
#pragma unroll
                    for (int k = 0; k < BLOCK_K / 16; ++k)
                        umma_ss(d_tmem, adesc + 2u * k, bdesc + 2u * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
                    tc_commit(empty_bar(slot));
                    if (kb == p.kb_per_tap - 1) { tc_commit(tfull_bar(acc)); tc_commit(b_done); }
                }
                __syncwarp();
                if (++slot == p.a_slots) { slot = 0; phase ^= 1u; }
            }
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1u;
        }
    } else if (warp == 1) {
        {
            const uint32_t fmt = (p.flags & DFW_EPI_F16) ? 0u : 1u;       // A and B must share one 16-bit format
            const uint32_t idesc = umma_idesc(BLOCK_M, BLOCK_N, fmt, fmt, 0);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int u = blockIdx.x; u < units; u += gridDim.x) {
                const int nsub = min(TPU, p.total_tiles - u * TPU);
                mbar_wait(tempty_bar(acc), acc_phase ^ 1u, 2);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * Cfg::ACC_COLS;
                for (int kbi = 0; kbi < kblocks; ++kbi) {
                    mbar_wait(full_bar(stage), phase, 3);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint64_t bdesc = umma_desc_sw128(sB(stage));
#pragma unroll
                        for (int sub = 0; sub < TPU; ++sub) {
                            if (sub < nsub) {
                                const uint64_t adesc = umma_desc_sw128(sA(stage, sub));
#pragma unroll
                                for (int k = 0; k < BLOCK_K / 16; ++k)
                                    umma_ss(d_tmem + sub * Cfg::SUB_COLS, adesc + 2u * k, bdesc + 2u * k, idesc,
                                            (kbi > 0 || k > 0) ? 1u : 0u);
                            }
                        }
                        tc_commit(empty_bar(stage));
                        if (kbi == kblocks - 1) tc_commit(tfull_bar(acc));
                    }
                    __syncwarp();
                    if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                }
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1u;
            }
        }
    } else if (warp == 3) {
        // residual loader for the TMA epilogue: one 8 KiB chunk (128 rows x 64 B) per epilogue chunk, 4-deep ring
        if (lane == 0 && p.tma_epi && p.has_res) {
            uint32_t g = 0;
            for (int u = blockIdx.x; u < units; u += gridDim.x)
            for (int t = u * TPU; t < min(u * TPU + TPU, p.total_tiles); ++t) {
                const TileCoord tc = decode_tile(p, t);
                for (int c = 0; c < chunks_per_tile; ++c, ++g) {
                    const int b = g & 3;
                    mbar_wait(buf_free(b), ((g >> 2) & 1) ^ 1u, 5);       // the store that last used this buffer is done
                    mbar_arrive_expect_tx(res_full(b), 8192);
                    tma_load_4d(epi_buf(b), &maps.res, res_full(b), tc.n_tile * out_cols_per_tile + c * CW, tc.w0,
                                tc.h0, tc.n0);
                }
            }
        }
    } else if (warp >= 4 && p.tma_epi) {
        // ---------------- TMA epilogue: thread = accumulator row; 64-byte chunks through swizzled smem ----------------
        // Two groups of four warps (one warp of each group per SM sub-partition, so the sub-partition's scheduler has a
        // second warp to issue from while the first waits on TMEM / smem / a barrier); group eg takes the chunks with
        // (global chunk index & 1) == eg and owns chunk buffers {eg, eg + 2}, named barrier 1 + eg and its own stores.
        const int eg = (warp - 4) >> 2;
        const int q = (warp - 4) & 3;                 // TMEM lane quarter = warp % 4
        const int row = q * 32 + lane;
        const int f16 = p.flags & DFW_EPI_F16;
        const bool out_f32 = (p.flags & DFW_EPI_OUT_F32) != 0;
        const bool geglu = (BLOCK_N == 256) && (p.flags & DFW_EPI_GEGLU);
        const bool issuer = (q == 0 && lane == 0);
        const bool has_res = p.has_res != 0;
        const bool has_bias = p.bias != nullptr;
        const bool has_gn = p.gn_partial != nullptr;
        const int epi_flags = p.flags;
        const float out_scale = p.out_scale;
        const int gn_cpg = p.gn_cpg;
        const int bar_id = 1 + eg;
        const uint32_t row_off = static_cast<uint32_t>(row) * 64u;
        const uint32_t sw = static_cast<uint32_t>((row >> 1) & 3);          // SWIZZLE_64B: unit ^= (row/2) % 4
        // fused GroupNorm statistics of the output (consumed by the next layer's GroupNorm): per-warp smem partials,
        // flushed to gn_partial[image][cta][group][2] whenever the CTA moves on to another image (TN == 1)
        float* wacc_all = reinterpret_cast<float*>(smem_raw + (gn_base - smem_raw_u32));
        float* wacc = wacc_all + (eg * 4 + q) * 64;
        const int gn_th = row / p.TW, gn_tw = row % p.TW;
        int gn_img = -1;
        auto gn_flush = [&](int img) {
            named_bar_sync(bar_id, 128);
            const int e = q * 32 + lane;
            if (e < 64) {
                const float* wg = wacc_all + eg * 256;
                const float tot = (wg[e] + wg[64 + e]) + (wg[128 + e] + wg[192 + e]);
                p.gn_partial[static_cast<size_t>(img) * p.gn_img_stride + (blockIdx.x * GN_SLOTS_PER_CTA + eg) * 64 + e] = tot;
            }
            named_bar_sync(bar_id, 128);
            if (lane < 32) { wacc[lane] = 0.f; wacc[32 + lane] = 0.f; }
            __syncwarp();
        };
        if (has_gn) {          // (the host zero-fills gn_partial: CTAs write only the images they touch)
            wacc[lane] = 0.f; wacc[32 + lane] = 0.f;
            __syncwarp();
        }
        int acc = 0;
        uint32_t acc_phase = 0, g = 0;
        for (int u = blockIdx.x; u < units; u += gridDim.x) {
            mbar_wait(tfull_bar(acc), acc_phase, 4);
            tc_fence_after();
          for (int t = u * TPU; t < min(u * TPU + TPU, p.total_tiles); ++t) {
            const TileCoord tc = decode_tile(p, t);
            bool gn_row_valid = false;
            if (has_gn) {
                if (tc.n0 != gn_img) {
                    if (gn_img >= 0) gn_flush(gn_img);
                    gn_img = tc.n0;
                }
                gn_row_valid = (tc.h0 + gn_th < p.H) && (tc.w0 + gn_tw < p.W);
            }
            const uint32_t taddr = tmem_base + acc * Cfg::ACC_COLS + (t - u * TPU) * Cfg::SUB_COLS +
                                   (static_cast<uint32_t>(q * 32) << 16);
#pragma unroll 1
            for (int c = 0; c < chunks_per_tile; ++c, ++g) {
                if ((g & 1u) != static_cast<uint32_t>(eg)) continue;
                const int b = g & 3;
                float f[32];
                const int col_out0 = tc.n_tile * out_cols_per_tile + c * CW;
                if (geglu) {
                    if constexpr (BLOCK_N == 256) {
                        uint32_t v[32], gt[32];
                        tmem_ld_32x32(taddr + c * 32, v);
                        tmem_ld_32x32(taddr + 128 + c * 32, gt);
                        tmem_ld_wait(); tmem_regs_ready(v); tmem_regs_ready(gt);      // compiler fence: consumers of the async load stay below the wait
                        const float* bp = p.bias ? p.bias + tc.n_tile * 256 + c * 32 : nullptr;
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            float4 bv = make_float4(0.f, 0.f, 0.f, 0.f), bg = bv;
                            if (bp) { bv = __ldg(reinterpret_cast<const float4*>(bp + j)); bg = __ldg(reinterpret_cast<const float4*>(bp + 128 + j)); }
                            geglu_pair(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(gt[j]),
                                       __uint_as_float(gt[j + 1]), bv.x, bv.y, bg.x, bg.y, f[j], f[j + 1]);
                            geglu_pair(__uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]), __uint_as_float(gt[j + 2]),
                                       __uint_as_float(gt[j + 3]), bv.z, bv.w, bg.z, bg.w, f[j + 2], f[j + 3]);
                        }
                    }
                } else {
                    const int col_in0 = tc.n_tile * BLOCK_N + c * CW;
                    const bool bias_vec = has_bias && (col_in0 + CW <= p.Cout);
                    if (out_f32) {
                        uint32_t v[16];
                        tmem_ld_32x16(taddr + c * 16, v);
                        float4 bb[4];                             // bias loads overlap the TMEM read
                        if (bias_vec) {
#pragma unroll
                            for (int j = 0; j < 4; ++j) bb[j] = __ldg(reinterpret_cast<const float4*>(p.bias + col_in0) + j);
                        }
                        tmem_ld_wait(); tmem_regs_ready16(v);      // compiler fence: consumers of the async load stay below the wait
#pragma unroll
                        for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]);
                        if (bias_vec) {
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                f[4 * j] += bb[j].x; f[4 * j + 1] += bb[j].y; f[4 * j + 2] += bb[j].z; f[4 * j + 3] += bb[j].w;
                            }
                        }
                    } else {
                        if constexpr (BLOCK_N >= 32) {
                            uint32_t v[32];
                            tmem_ld_32x32(taddr + c * 32, v);
                            float4 bb[8];
                            if (bias_vec) {
#pragma unroll
                                for (int j = 0; j < 8; ++j) bb[j] = __ldg(reinterpret_cast<const float4*>(p.bias + col_in0) + j);
                            }
                            tmem_ld_wait(); tmem_regs_ready(v);      // compiler fence: consumers of the async load stay below the wait
#pragma unroll
                            for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
                            if (bias_vec) {
#pragma unroll
                                for (int j = 0; j < 8; ++j) {
                                    f[4 * j] += bb[j].x; f[4 * j + 1] += bb[j].y; f[4 * j + 2] += bb[j].z; f[4 * j + 3] += bb[j].w;
                                }
                            }
                        }
                    }
                    if (has_bias && !bias_vec) {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (j < CW && col_in0 + j < p.Cout) f[j] += __ldg(p.bias + col_in0 + j);
                    }
                    if (out_scale != 1.0f) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) f[j] *= out_scale;
                    }
                    if (epi_flags & DFW_EPI_SILU) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) f[j] = silu(f[j]);
                    }
                }
                if (has_res) {
                    mbar_wait(res_full(b), (g >> 2) & 1, 6);
                    const uint32_t rb = epi_buf(b) + row_off;
#pragma unroll
                    for (uint32_t u = 0; u < 4; ++u) {
                        uint4 r;
                        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                                     : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(rb + ((u ^ sw) << 4)));
                        if (out_f32) {
                            f[4 * u] += __uint_as_float(r.x); f[4 * u + 1] += __uint_as_float(r.y);
                            f[4 * u + 2] += __uint_as_float(r.z); f[4 * u + 3] += __uint_as_float(r.w);
                        } else {
                            const float2 a0 = unpack_h2(r.x, f16), a1 = unpack_h2(r.y, f16), a2 = unpack_h2(r.z, f16),
                                         a3 = unpack_h2(r.w, f16);
                            f[8 * u] += a0.x; f[8 * u + 1] += a0.y; f[8 * u + 2] += a1.x; f[8 * u + 3] += a1.y;
                            f[8 * u + 4] += a2.x; f[8 * u + 5] += a2.y; f[8 * u + 6] += a3.x; f[8 * u + 7] += a3.y;
                        }
                    }
                }
                if (has_gn && col_out0 < p.out_ch) {
                    if (out_f32) {
                        if (gn_cpg == 4) gn_chunk_stats<4, 16>(f, gn_row_valid, col_out0, lane, wacc);
                        else if (gn_cpg == 8) gn_chunk_stats<8, 16>(f, gn_row_valid, col_out0, lane, wacc);
                        else gn_chunk_stats<16, 16>(f, gn_row_valid, col_out0, lane, wacc);
                    } else {
                        if (gn_cpg == 4) gn_chunk_stats<4, 32>(f, gn_row_valid, col_out0, lane, wacc);
                        else if (gn_cpg == 8) gn_chunk_stats<8, 32>(f, gn_row_valid, col_out0, lane, wacc);
                        else gn_chunk_stats<16, 32>(f, gn_row_valid, col_out0, lane, wacc);
                    }
                }
                if (!has_res) {
                    // no loader in the loop: this group's store two chunks ago (same buffer) must have finished reading
                    if (issuer) tma_store_wait_read<1>();
                    named_bar_sync(bar_id, 128);
                }
                const uint32_t ob = epi_buf(b) + row_off;
#pragma unroll
                for (uint32_t u = 0; u < 4; ++u) {
                    uint4 w;
                    if (out_f32) {
                        w.x = __float_as_uint(f[4 * u]); w.y = __float_as_uint(f[4 * u + 1]);
                        w.z = __float_as_uint(f[4 * u + 2]); w.w = __float_as_uint(f[4 * u + 3]);
                    } else {
                        w.x = pack_h2(f[8 * u], f[8 * u + 1], f16); w.y = pack_h2(f[8 * u + 2], f[8 * u + 3], f16);
                        w.z = pack_h2(f[8 * u + 4], f[8 * u + 5], f16); w.w = pack_h2(f[8 * u + 6], f[8 * u + 7], f16);
                    }
                    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};"
                                 ::"r"(ob + ((u ^ sw) << 4)), "r"(w.x), "r"(w.y), "r"(w.z), "r"(w.w) : "memory");
                }
                fence_proxy_async_smem();
                named_bar_sync(bar_id, 128);
                if (issuer) {
                    tma_store_4d(&maps.out, epi_buf(b), col_out0, tc.w0, tc.h0, tc.n0);
                    tma_store_commit();
                    if (has_res && g >= 2) {
                        tma_store_wait_read<1>();                    // this group's previous store (chunk g-2) has drained
                        mbar_arrive(buf_free((g - 2) & 3));          // its buffer: hand it back to the residual loader
                    }
                }
            }
          }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(acc));
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1u;
        }
        if (has_gn && gn_img >= 0) gn_flush(gn_img);
        if (issuer) tma_store_wait_all<0>();
    } else if (warp >= 4 && warp < 8) {
        const int q = warp - 4;
        const int row = q * 32 + lane;
        const int tn = row / (p.TH * p.TW);
        const int rem = row % (p.TH * p.TW);
        const int th = rem / p.TW;
        const int tw = rem % p.TW;
        float* stg = reinterpret_cast<float*>(epi_generic + q * STG_BYTES_PER_WARP);
        EpiRow* rows = reinterpret_cast<EpiRow*>(epi_generic + 4 * STG_BYTES_PER_WARP) + q * 32;
        const bool geglu = (BLOCK_N == 256) && (p.flags & DFW_EPI_GEGLU);
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int u = blockIdx.x; u < units; u += gridDim.x) {
            mbar_wait(tfull_bar(acc), acc_phase, 4);
            tc_fence_after();
          for (int t = u * TPU; t < min(u * TPU + TPU, p.total_tiles); ++t) {
            const TileCoord tc = decode_tile(p, t);
            {
                const int img = tc.n0 + tn, hh = tc.h0 + th, ww = tc.w0 + tw;
                const bool ok = (img < p.N) && (hh < p.H) && (ww < p.W);
                EpiRow er;
                er.pix = ok ? (static_cast<long long>(img) * p.H + hh) * p.W + ww : -1;
                er.img = img;
                er.pad = 0;
                __syncwarp();
                rows[lane] = er;
                __syncwarp();
            }
            const uint32_t taddr = tmem_base + acc * Cfg::ACC_COLS + (t - u * TPU) * Cfg::SUB_COLS +
                                   (static_cast<uint32_t>(q * 32) << 16);
            if (geglu) {
                if constexpr (BLOCK_N == 256) {
                    const int cq = (lane & 7) * 4, sub = lane >> 3;
#pragma unroll 1
                    for (int c = 0; c < 4; ++c) {
                        uint32_t v[32];
                        float4 val[8];
                        const int cv = tc.n_tile * 256 + c * 32 + cq;             // permuted weight row of this lane's values
                        tmem_ld_32x32(taddr + c * 32, v);
                        tmem_ld_wait(); tmem_regs_ready(v);      // compiler fence: consumers of the async load stay below the wait
                        stage_rows(stg, lane, v);
                        __syncwarp();
                        float4 bv = make_float4(0.f, 0.f, 0.f, 0.f), bg = bv;
                        if (p.bias) {
                            bv = __ldg(reinterpret_cast<const float4*>(p.bias + cv));
                            bg = __ldg(reinterpret_cast<const float4*>(p.bias + cv + 128));
                        }
#pragma unroll
                        for (int i = 0; i < 8; ++i) val[i] = *reinterpret_cast<const float4*>(stg + (4 * i + sub) * STG_LD + cq);
                        __syncwarp();
                        tmem_ld_32x32(taddr + 128 + c * 32, v);
                        tmem_ld_wait(); tmem_regs_ready(v);      // compiler fence: consumers of the async load stay below the wait
                        stage_rows(stg, lane, v);
                        __syncwarp();
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int r = 4 * i + sub;
                            const long long pix = rows[r].pix;
                            if (pix < 0) continue;
                            const float4 g = *reinterpret_cast<const float4*>(stg + r * STG_LD + cq);
                            uint2 o;
                            const int f16 = p.flags & DFW_EPI_F16;
                            o.x = pack_h2((val[i].x + bv.x) * gelu_erf(g.x + bg.x), (val[i].y + bv.y) * gelu_erf(g.y + bg.y), f16);
                            o.y = pack_h2((val[i].z + bv.z) * gelu_erf(g.z + bg.z), (val[i].w + bv.w) * gelu_erf(g.w + bg.w), f16);
                            *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(p.out) + pix * p.out_ch +
                                                      tc.n_tile * 128 + c * 32 + cq) = o;
                        }
                        __syncwarp();
                    }
                }
            } else if constexpr (BLOCK_N >= 32) {
                ResChunk rc_next;
                prefetch_residual(p, rows, lane, tc.n_tile * BLOCK_N, tc.n_tile * BLOCK_N, 32, rc_next);
#pragma unroll 1
                for (int c = 0; c < BLOCK_N / 32; ++c) {
                    const ResChunk rc = rc_next;
                    uint32_t v[32];
                    tmem_ld_32x32(taddr + c * 32, v);
                    tmem_ld_wait(); tmem_regs_ready(v);      // compiler fence: consumers of the async load stay below the wait
                    stage_rows(stg, lane, v);
                    __syncwarp();
                    const int col0 = tc.n_tile * BLOCK_N + c * 32;
                    if (c + 1 < BLOCK_N / 32) prefetch_residual(p, rows, lane, col0 + 32, col0 + 32, 32, rc_next);
                    epilogue_store_chunk(p, stg, rows, lane, col0, col0, 32, rc);
                    __syncwarp();
                }
            } else {
                uint32_t v[32];
                uint32_t v16[16];
                tmem_ld_32x16(taddr, v16);
                tmem_ld_wait(); tmem_regs_ready16(v16);      // compiler fence: consumers of the async load stay below the wait
#pragma unroll
                for (int j = 0; j < 16; ++j) { v[j] = v16[j]; v[16 + j] = 0u; }
                stage_rows(stg, lane, v);
                __syncwarp();
                const int col0 = tc.n_tile * BLOCK_N;
                ResChunk rc;
                prefetch_residual(p, rows, lane, col0, col0, 16, rc);
                epilogue_store_chunk(p, stg, rows, lane, col0, col0, 16, rc);
                __syncwarp();
            }
          }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(acc));
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1u;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

// ---------------------------------------------------------------------------------------------------------
// "T128": channel-major accumulator variant for the Cout = 128, 3x3, stride-1 convolutions (the 512^2 layers of the
// VAE, 1/7 of the whole step).  With pixels as M and 128 channels as N every tcgen05.mma (128x128x16) has to read
// 4 KiB of A and 4 KiB of B from smem in its 64 cycles -- the 128 B/clk of the smem port, so the tensor pipe stalls
// on operands (73 % active in ncu).  Here the roles are swapped: A = the weight tile [128 ch x 64], B = a 16 x 16-pixel
// patch [256 px x 64]: one 128x256x16 MMA reads 4 + 8 KiB in 128 cycles (96 B/clk, like the N = 256 layers, 91 %).
// The accumulator is then D^T: TMEM lane = output channel, column = pixel of the tile.  Epilogue: a thread owns one
// channel, so bias is a scalar, the GroupNorm statistics are two running registers (no shuffles until the image
// changes), and each warp writes its own [64 px x 32 ch] staging block with 2-byte stores (a warp store = one 64 B
// pixel row, conflict-free) and TMA-stores it -- no CTA-level barrier anywhere in the epilogue.
//   warp 0  patch producer: (16+2) x 16 px x 64 ch per (channel block, horizontal tap), the 3 vertical taps are the
//           same patch at +0 / +2048 / +4096 B
//   warp 1  MMA issuer      warp 2  TMEM allocator + weight-tile producer      warp 3  residual loader (TMA)
//   warps 4-11  epilogue: channel quarter q = warp % 4, pixel half eg = (warp - 4) / 4 (tile rows 8 eg .. 8 eg + 7)
// ---------------------------------------------------------------------------------------------------------
template <bool GN_IN>
struct T128CfgT {
    static constexpr int PATCH_BYTES = 18 * 16 * 128;       // 36 KiB
    static constexpr int A_SLOTS = T128_A_SLOTS;
    static constexpr int W_TILE_BYTES = 128 * 128;          // 16 KiB
    static constexpr int W_SLOTS = T128_W_SLOTS;
    static constexpr int STG_BYTES = 64 * 64;               // 64 pixels x 32 channels x 2 B
    // GN_IN: per-CTA ring of transformed (16+2) x (16+2)-pixel x 64-channel blocks in global memory (L2-resident)
    static constexpr int XF_BUFS = 4;
    static constexpr int XF_BUF_BYTES = 18 * 18 * 128;      // 40.5 KiB
    static constexpr int NBAR = 2 * A_SLOTS + 2 * W_SLOTS + 4 + 32 + 2 * XF_BUFS;
    static constexpr int XF_WARPS = 8;                         // GN_IN: operand-transform warps (GroupNorm + SiLU on the fly)
    static constexpr int THREADS = IGEMM_THREADS + (GN_IN ? 32 * XF_WARPS : 0);     // 12 (+ 8) warps
    static constexpr int SS_MAX_CIN = 512;                  // GN_IN: scale / shift of the current image live in smem (2 x Cin floats)
    static constexpr int SS_BYTES = GN_IN ? 2 * SS_MAX_CIN * 4 + 16 : 0;
    static constexpr int SMEM_USED = A_SLOTS * PATCH_BYTES + W_SLOTS * W_TILE_BYTES + 16 * STG_BYTES + 8 * NBAR + 16 + SS_BYTES;
    static constexpr int SMEM_BYTES = SMEM_USED + 512;      // smem is declared __align__(1024); the kernel traps otherwise
    static_assert(SMEM_BYTES <= 232448, "smem budget");
};
struct T128Maps {
    CUtensorMap patch, w, out, res;
};
struct T128Params {
    int N, H, W;
    int ktaps;            // 3: 3x3 / pad 1;  1: 1x1 (the patch is the bare 16 x 16 tile)
    int kb_per_tap;
    int tiles_w, tiles_h, total_units;
    int n_slabs;          // ceil(Cout / 128): a unit = (image, 128-channel slab, 16 x 16 tile), tiles fastest
    int Cout;             // may end inside the last slab (1x1 only): TMA zero-fills / clips the channels beyond it
    int gn_cpg;           // channels per GroupNorm group (4 / 8 / 16)
    const float* bias;
    int f16, has_res;
    float out_scale;
    float* gn_partial;
    long long gn_img_stride;
    const float* gn_in;   // GN_IN: [N][2][Cin] fp32 scale / shift of the GroupNorm (+ SiLU) applied to x on the fly
    int Cin;
    const void* x;        // GN_IN: the raw activation tensor (the transform warps read it with plain loads)
    void* xf_scratch;     // GN_IN: [gridDim.x][XF_BUFS][18][18][64] 16-bit ring of transformed blocks (maps.patch views it)
};

// GN_IN: the convolution consumes silu(groupnorm(x)) without that tensor ever reaching HBM.  Eight extra warps read the
// raw (16+2) x (16+2)-pixel neighbourhood of every (tile, 64-channel block) ONCE with plain 16-byte loads (8 threads = the
// 128 contiguous bytes of a pixel; two register sets, so the loads of block i + 1 are in flight while block i is
// processed), apply x * scale[n, c] + shift[n, c] and SiLU in registers (fp32, one tanh.approx per element; pixels outside
// the image become 0 = the zero padding of the normalised tensor) and store the block into a small per-CTA ring in global
// memory (4 x 40.5 KiB per CTA, 24 MB in all: it lives in L2).  The ordinary TMA patch producer then loads the three
// horizontally shifted patches from the ring exactly as it would from the activation tensor, so the shared-memory side
// of the kernel is the plain one.  Measured dead ends (profiles/r02_gnin_*.json): rewriting each TMA-loaded patch in place
// (3 transforms + an smem round trip per element: MUFU-bound, 0.6x the plain kernel); writing the three shifted patches
// with st.shared from registers (one transform per element, but generic-proxy smem stores get ~30 B/clk next to a running
// tcgen05.mma that reads 96 B/clk: 2800-3100 cycles per block for the stores alone against 4608 of MMA).
template <bool GN_IN>
__global__ void __launch_bounds__(T128CfgT<GN_IN>::THREADS, 1)
igemm_t128_kernel(const __grid_constant__ T128Maps maps, const __grid_constant__ T128Params p) {
    using Cfg = T128CfgT<GN_IN>;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    if (base - raw > static_cast<uint32_t>(Cfg::SMEM_BYTES - Cfg::SMEM_USED)) __trap();
    auto sP = [&](int s) { return base + s * Cfg::PATCH_BYTES; };
    auto sW = [&](int s) { return base + Cfg::A_SLOTS * Cfg::PATCH_BYTES + s * Cfg::W_TILE_BYTES; };
    const uint32_t stg_base = base + Cfg::A_SLOTS * Cfg::PATCH_BYTES + Cfg::W_SLOTS * Cfg::W_TILE_BYTES;
    auto sidx = [&](int w8, int b) { return w8 * 2 + b; };      // staging block of (warp, sub-block)
    auto sS = [&](int w8, int b) { return stg_base + sidx(w8, b) * Cfg::STG_BYTES; };
    const uint32_t bar = stg_base + 16 * Cfg::STG_BYTES;
    auto pa_full = [&](int i) { return bar + 8u * i; };
    auto pa_empty = [&](int i) { return bar + 8u * (Cfg::A_SLOTS + i); };
    auto w_full = [&](int i) { return bar + 8u * (2 * Cfg::A_SLOTS + i); };
    auto w_empty = [&](int i) { return bar + 8u * (2 * Cfg::A_SLOTS + Cfg::W_SLOTS + i); };
    auto xf_ready = [&](int i) { return bar + 8u * (2 * Cfg::A_SLOTS + 2 * Cfg::W_SLOTS + i); };                   // GN_IN ring
    auto xf_free = [&](int i) { return bar + 8u * (2 * Cfg::A_SLOTS + 2 * Cfg::W_SLOTS + Cfg::XF_BUFS + i); };
    const uint32_t b2 = bar + 8u * (2 * Cfg::A_SLOTS + 2 * Cfg::W_SLOTS + 2 * Cfg::XF_BUFS);
    auto tfull = [&](int a) { return b2 + 8u * a; };
    auto tempty = [&](int a) { return b2 + 8u * (2 + a); };
    auto res_full = [&](int i) { return b2 + 8u * (4 + i); };        // i = warp8 * 2 + buffer
    auto buf_free = [&](int i) { return b2 + 8u * (20 + i); };
    const uint32_t tmem_slot = b2 + 8u * 36;
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw));
    const uint32_t ss_base = (tmem_slot + 16 + 15) & ~15u;   // GN_IN: 2 x Cin floats, 16-byte aligned
    (void)ss_base;

    const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&maps.patch); tma_prefetch_desc(&maps.w); tma_prefetch_desc(&maps.out);
        if (p.has_res) tma_prefetch_desc(&maps.res);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < Cfg::A_SLOTS; ++i) { mbar_init(pa_full(i), 1); mbar_init(pa_empty(i), 1); }
        for (int i = 0; i < Cfg::W_SLOTS; ++i) { mbar_init(w_full(i), 1); mbar_init(w_empty(i), 1); }
        for (int i = 0; i < Cfg::XF_BUFS; ++i) { mbar_init(xf_ready(i), Cfg::XF_WARPS); mbar_init(xf_free(i), 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(tfull(a), 1); mbar_init(tempty(a), 8); }
        for (int i = 0; i < 16; ++i) { mbar_init(res_full(i), 1); mbar_init(buf_free(i), 1); }
        fence_mbar_init();
    }
    if (warp == 2) {
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    pdl_launch_dependents();
    pdl_wait();
    const int tiles_per_img = p.tiles_w * p.tiles_h;
    auto decode = [&](int u, int& n, int& slab, int& h0, int& w0) {
        const int ns = u / tiles_per_img;              // image * n_slabs + slab
        const int r = u - ns * tiles_per_img;
        n = ns / p.n_slabs;
        slab = ns - n * p.n_slabs;
        h0 = (r / p.tiles_w) * 16;
        w0 = (r % p.tiles_w) * 16;
    };

    // GN_IN: 640 threads leave 96 registers per thread, but the transform warps want two 44-register load buffers; the
    // four single-thread role warps give theirs up (setmaxnreg is per warpgroup: 0 = roles, 1-2 = epilogue, 3-4 = transform)
    if (warp < 4) {
    if constexpr (GN_IN) asm volatile("setmaxnreg.dec.sync.aligned.u32 48;");
    if (warp == 0) {
        int as = 0;
        uint32_t aph = 0;
        int xb = 0;                      // GN_IN: ring buffer of the current channel block
        uint32_t xph = 0;
        for (int u = blockIdx.x; u < p.total_units; u += gridDim.x) {
            int n, slab, h0, w0;
            decode(u, n, slab, h0, w0);
            const int pad = p.ktaps >> 1;
            const uint32_t patch_bytes = (16 + 2 * pad) * 16 * 128;
            for (int kb = 0; kb < p.kb_per_tap; ++kb) {
                if constexpr (GN_IN) mbar_wait(xf_ready(xb), xph, 20);        // the transformed block is in the ring
                for (int dw = 0; dw < p.ktaps; ++dw) {
                    mbar_wait(pa_empty(as), aph ^ 1u, 21);
                    if (elect_one()) {
                        mbar_arrive_expect_tx(pa_full(as), patch_bytes);
                        if constexpr (GN_IN)
                            tma_load_4d(sP(as), &maps.patch, pa_full(as), 0, dw, 0, blockIdx.x * Cfg::XF_BUFS + xb);
                        else
                            tma_load_4d(sP(as), &maps.patch, pa_full(as), kb * BLOCK_K, w0 + dw - pad, h0 - pad, n);
                    }
                    __syncwarp();
                    if (++as == Cfg::A_SLOTS) { as = 0; aph ^= 1u; }
                }
                if (++xb == Cfg::XF_BUFS) { xb = 0; xph ^= 1u; }
            }
        }
    } else if (warp == 2) {
        int ws = 0;
        uint32_t wph = 0;
        for (int u = blockIdx.x; u < p.total_units; u += gridDim.x) {
            const int slab = (u / tiles_per_img) % p.n_slabs;
            for (int kb = 0; kb < p.kb_per_tap; ++kb)
                for (int dw = 0; dw < p.ktaps; ++dw)
                    for (int dh = 0; dh < p.ktaps; ++dh) {
                        mbar_wait(w_empty(ws), wph ^ 1u, 22);
                        if (elect_one()) {
                            mbar_arrive_expect_tx(w_full(ws), Cfg::W_TILE_BYTES);
                            tma_load_3d(sW(ws), &maps.w, w_full(ws), ((dh * p.ktaps + dw) * p.kb_per_tap + kb) * BLOCK_K,
                                        slab * 128, 0);
                        }
                        __syncwarp();
                        if (++ws == Cfg::W_SLOTS) { ws = 0; wph ^= 1u; }
                    }
        }
    } else if (warp == 1) {
        const uint32_t fmt = p.f16 ? 0u : 1u;
        const uint32_t idesc = umma_idesc(128, 256, fmt, fmt, 0);
        int as = 0, ws = 0, acc = 0, xb = 0;
        uint32_t aph = 0, wph = 0, acc_phase = 0;
        long long trA = 0, trW = 0, trAcc = 0, trTot = clock64();
        for (int u = blockIdx.x; u < p.total_units; u += gridDim.x) {
            { TR_T0(); mbar_wait(tempty(acc), acc_phase ^ 1u, 23); TR_ADD(trAcc); }
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * 256;
            bool first = true;
            const int last = p.ktaps - 1;
            for (int kb = 0; kb < p.kb_per_tap; ++kb)
                for (int dw = 0; dw < p.ktaps; ++dw) {
                    { TR_T0(); mbar_wait(pa_full(as), aph, 24); TR_ADD(trA); }
                    for (int dh = 0; dh < p.ktaps; ++dh) {
                        { TR_T0(); mbar_wait(w_full(ws), wph, 25); TR_ADD(trW); }
                        tc_fence_after();
                        if (elect_one()) {
                            const uint64_t adesc = umma_desc_sw128(sW(ws));
                            const uint64_t bdesc = umma_desc_sw128(sP(as) + dh * 16 * 128);   // tile row r = patch row r + dh
#pragma unroll
                            for (int k = 0; k < BLOCK_K / 16; ++k)
                                umma_ss(d_tmem, adesc + 2u * k, bdesc + 2u * k, idesc, (!first || k > 0) ? 1u : 0u);
                            tc_commit(w_empty(ws));
                            if (dh == last) tc_commit(pa_empty(as));
                            if (GN_IN && dh == last && dw == last) tc_commit(xf_free(xb));   // (its TMA reads ended long before)
                            if (dh == last && dw == last && kb == p.kb_per_tap - 1) tc_commit(tfull(acc));
                        }
                        __syncwarp();
                        first = false;
                        if (++ws == Cfg::W_SLOTS) { ws = 0; wph ^= 1u; }
                    }
                    if (++as == Cfg::A_SLOTS) { as = 0; aph ^= 1u; }
                    if (dw == last && ++xb == Cfg::XF_BUFS) xb = 0;
                }
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1u;
        }
#if DFW_GNIN_TRACE
        if (blockIdx.x == 0 && lane == 0) {
            g_t128_trace[0] += clock64() - trTot; g_t128_trace[1] += trA; g_t128_trace[2] += trW; g_t128_trace[3] += trAcc;
        }
#else
        (void)trA; (void)trW; (void)trAcc; (void)trTot;
#endif
    } else if (warp == 3) {
        if (p.has_res) {
            uint32_t k = 0;
            for (int u = blockIdx.x; u < p.total_units; u += gridDim.x, ++k) {
                int n, slab, h0, w0;
                decode(u, n, slab, h0, w0);
                for (int sb = 0; sb < 2; ++sb)
                    for (int w8 = 0; w8 < 8; ++w8) {
                        const int i = sidx(w8, sb);
                        mbar_wait(buf_free(i), (k & 1u) ^ 1u, 26);        // the store that last used this block has drained
                        if (elect_one()) {
                            mbar_arrive_expect_tx(res_full(i), Cfg::STG_BYTES);
                            tma_load_4d(sS(w8, sb), &maps.res, res_full(i), slab * 128 + 32 * (w8 & 3), w0,
                                        h0 + 8 * (w8 >> 2) + 4 * sb, n);
                        }
                        __syncwarp();
                    }
            }
        }
    }
    } else if (warp >= 12) {
        if constexpr (GN_IN) {
            asm volatile("setmaxnreg.inc.sync.aligned.u32 136;");
            // Transform warps.  Thread = (16-byte unit `un` of a pixel's 128-byte channel block, pixel lane p0 of 32); a warp
            // instruction reads / writes 4 whole pixels.  The raw (16 + 2 pad)^2 neighbourhood is split so that every address
            // is affine in the loop index: MAIN = rows rr + 2 i (i < 8 + pad) x the 16 tile columns (always inside the image
            // horizontally); HALO = the left / right neighbour columns (pad only), 36 pixels = slot 9 of every lane + slot 10
            // of four lanes.  Ring block layout: [18 rows][18 columns][64 channels] (pad = 0 uses its 16 x 16 corner).
            constexpr int MAXU = 11;
            const int t = threadIdx.x - 12 * 32;
            const int un = t & 7, p0 = t >> 3;
            const int x = p0 & 15, rr = p0 >> 4;
            const int pad = p.ktaps >> 1;
            const int NI = 8 + pad;
            const int hc = p0 & 1;                                   // halo: 0 = left neighbour column, 1 = right
            const int hr9 = p0 >> 1, hr10 = 16 + (p0 >> 1);          // halo rows of slots 9 / 10 (slot 10: p0 < 4 only)
            const int C8 = p.Cin >> 3;                               // 16-byte units per pixel
            const long long rs2 = 2LL * p.W * C8;
            uint4* ring = reinterpret_cast<uint4*>(p.xf_scratch) +
                          static_cast<size_t>(blockIdx.x) * Cfg::XF_BUFS * (Cfg::XF_BUF_BYTES / 16) + un;
            const int om = ((rr * 18) + pad + x) * 8;               // main slot i: + i * 36 * 8
            const int o9 = (hr9 * 18 + 17 * hc) * 8, o10 = (hr10 * 18 + 17 * hc) * 8;
            float* ss_smem = reinterpret_cast<float*>(smem_raw + (ss_base - raw));
            int xb = 0;
            uint32_t xph = 0;
            auto transform = [&](auto F16TAG, uint4 (&v)[MAXU], uint32_t inb, const float4 (&ss)[4]) {
                constexpr bool F16 = decltype(F16TAG)::value;
#pragma unroll
                for (int i = 0; i < MAXU; ++i) {
                    if (!(DFW_GNIN_DBG & 2) && (inb & (1u << i))) {      // outside the image: stays 0 (padding of the NORMALISED tensor)
                        uint32_t* vw = reinterpret_cast<uint32_t*>(&v[i]);
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            float x0, x1;
                            if constexpr (F16) {
                                const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&vw[j]));
                                x0 = f.x; x1 = f.y;
                            } else {
                                x0 = bf16_lo(vw[j]); x1 = bf16_hi(vw[j]);
                            }
                            const float4 sc4 = ss[j >> 1], sh4 = ss[2 + (j >> 1)];
                            const float g0 = fmaf(x0, (j & 1) ? sc4.z : sc4.x, (j & 1) ? sh4.z : sh4.x);
                            const float g1 = fmaf(x1, (j & 1) ? sc4.w : sc4.y, (j & 1) ? sh4.w : sh4.y);
                            float t0, t1;
                            asm("tanh.approx.f32 %0, %1;" : "=f"(t0) : "f"(g0));
                            asm("tanh.approx.f32 %0, %1;" : "=f"(t1) : "f"(g1));
                            const float y0 = fmaf(g0, t0, g0), y1 = fmaf(g1, t1, g1);
                            vw[j] = F16 ? pack_f16x2(y0, y1) : pack_bf16x2(y0, y1);
                        }
                    }
                }
            };
            long long trL = 0, trS = 0, trWr = 0, trN = 0, trF = 0, trTot = clock64();
            // raw loads of one (tile, channel block): 11 x 16 bytes per thread, zero where the pixel is outside the image
            auto issue = [&](uint4 (&v)[MAXU], uint32_t& inb, int& n, int u, int kb) {
                int slab, h0, w0;
                decode(u, n, slab, h0, w0);
                const uint4* ximg = reinterpret_cast<const uint4*>(p.x) + static_cast<long long>(n) * p.H * p.W * C8 + un +
                                    kb * (BLOCK_K / 8);
                const bool top = pad && h0 == 0, bot = pad && h0 + 16 == p.H;
                const bool colok = pad && !(hc == 0 ? w0 == 0 : w0 + 16 == p.W);
                inb = (1u << NI) - 1u;
                if (top && rr == 0) inb &= ~1u;
                if (bot && rr == 1) inb &= ~(1u << (NI - 1));
                if (colok && !(top && hr9 == 0)) inb |= 1u << 9;
                if (colok && p0 < 4 && !(bot && hr10 == 17)) inb |= 1u << 10;
                const uint4* gm = ximg + (static_cast<long long>(h0 - pad + rr) * p.W + (w0 + x)) * C8;
                const uint4* g9 = ximg + (static_cast<long long>(h0 - 1 + hr9) * p.W + (w0 - 1 + 17 * hc)) * C8;
                const uint4* g10 = g9 + 16LL * p.W * C8;
#pragma unroll
                for (int i = 0; i < MAXU; ++i) {
                    v[i] = make_uint4(0u, 0u, 0u, 0u);
                    if (inb & (1u << i)) {
#if DFW_GNIN_DBG & 1
                        v[i] = make_uint4(0x3c003800u + h0, 0x3c003800u, 0x3c003800u + w0, 0x3c003800u);
#else
                        v[i] = __ldg(i < 9 ? gm + i * rs2 : (i == 9 ? g9 : g10));
#endif
                    }
                }
            };
            // transform in registers, then into the ring
            // The proxy fence that publishes the ring stores of block i waits for them to reach L2 (MEMBAR.GPU, ~1000
            // cycles right after the stores), so it is issued one block late, after the transform of block i + 1.
            int pend = -1;                             // ring block whose stores are not yet published
            auto publish = [&]() {
                if (pend >= 0) {
                    fence_proxy_async_all();          // generic-proxy global stores -> visible to the TMA (async proxy) reads
                    __syncwarp();
                    if (lane == 0) mbar_arrive(xf_ready(pend));
                    pend = -1;
                }
            };
            int ss_n = -1;                              // image whose (halved) scale / shift are in smem
            auto process = [&](uint4 (&v)[MAXU], uint32_t inb, int n, int kb) {
                TR_T0();
                if (n != ss_n) {                        // a CTA meets the images in order: every transform warp gets here together
                    named_bar_sync(1, 32 * Cfg::XF_WARPS);             // everybody is done with the previous image's values
                    const float* src = p.gn_in + static_cast<size_t>(n) * 2 * p.Cin;
                    for (int i = t; i < 2 * p.Cin; i += 32 * Cfg::XF_WARPS)
                        ss_smem[i] = 0.5f * __ldg(src + i);            // silu(y) = h + h tanh(h), h = y / 2 (exact scaling)
                    named_bar_sync(1, 32 * Cfg::XF_WARPS);
                    ss_n = n;
                }
                float4 ss[4];
                {
                    const float4* s4 = reinterpret_cast<const float4*>(ss_smem + kb * BLOCK_K + un * 8);
                    ss[0] = s4[0]; ss[1] = s4[1];
                    const float4* h4 = reinterpret_cast<const float4*>(ss_smem + p.Cin + kb * BLOCK_K + un * 8);
                    ss[2] = h4[0]; ss[3] = h4[1];
                }
                if (p.f16) transform(std::true_type{}, v, inb, ss);
                else transform(std::false_type{}, v, inb, ss);
#if DFW_GNIN_TRACE
                asm volatile("" :: "r"(v[0].x), "r"(v[5].y), "r"(v[8].z), "r"(v[9].w));
                TR_ADD(trL); ++trN;
#endif
                { TR_T0(); publish(); TR_ADD(trF); }
                { TR_T0(); mbar_wait(xf_free(xb), xph ^ 1u, 29); TR_ADD(trS); }      // the MMAs that read this ring block are done
                {
                    TR_T0();
                    uint4* dst = ring + xb * (Cfg::XF_BUF_BYTES / 16);
#pragma unroll
                    for (int i = 0; i < 9; ++i)
                        if (!(DFW_GNIN_DBG & 4) && i < NI) dst[om + i * (36 * 8)] = v[i];
                    if (!(DFW_GNIN_DBG & 4) && pad) {
                        dst[o9] = v[9];
                        if (p0 < 4) dst[o10] = v[10];
                    }
                    TR_ADD(trWr);
                }
                pend = xb;
                if (++xb == Cfg::XF_BUFS) { xb = 0; xph ^= 1u; }
            };
            // software pipeline over the flattened (tile, channel block) stream: the loads of block i + 1 are in flight
            // (second register set) while block i is transformed and written
            uint4 vA[MAXU], vB[MAXU];
            uint32_t inbA = 0, inbB = 0;
            int nA = 0, nB = 0, kbA = 0, kbB = 0;
            int u = blockIdx.x, kb = 0;
            auto advance = [&](int& uu, int& kk) { if (++kk == p.kb_per_tap) { kk = 0; uu += gridDim.x; } };
            if (u < p.total_units) { issue(vA, inbA, nA, u, kb); kbA = kb; }
            while (u < p.total_units) {
                advance(u, kb);
                if (u < p.total_units) { issue(vB, inbB, nB, u, kb); kbB = kb; }
                process(vA, inbA, nA, kbA);
                if (u >= p.total_units) break;
                advance(u, kb);
                if (u < p.total_units) { issue(vA, inbA, nA, u, kb); kbA = kb; }
                process(vB, inbB, nB, kbB);
            }
            publish();
#if DFW_GNIN_TRACE
            if (blockIdx.x == 0 && t == 0) {
                g_t128_trace[4] += clock64() - trTot; g_t128_trace[5] += trL; g_t128_trace[6] += trS; g_t128_trace[7] += trWr;
                g_t128_trace[8] += trN; g_t128_trace[9] += trF;
            }
#else
            (void)trL; (void)trS; (void)trWr; (void)trN; (void)trTot; (void)trF;
#endif
        }
    } else {
        if constexpr (GN_IN) asm volatile("setmaxnreg.dec.sync.aligned.u32 80;");
        const int w8 = warp - 4;
        const int q = w8 & 3, eg = w8 >> 2;
        const int ch = 32 * q + lane;                    // channel inside the 128-channel slab
        const float out_scale = p.out_scale;
        const bool has_res = p.has_res != 0, has_gn = p.gn_partial != nullptr;
        const int f16 = p.f16;
        uint8_t* stg_generic = smem_raw + (stg_base - raw);
        float gs = 0.f, gss = 0.f, bias_v = 0.f;
        int gn_key = -1;                                  // image * n_slabs + slab the running sums belong to
        const int cpg = p.gn_cpg;
        auto gn_flush = [&](int key) {                    // a CTA meets every (image, slab) in one contiguous run of units
            float s = gs, ss = gss;
            for (int o = 1; o < cpg; o <<= 1) {           // cpg (4 / 8 / 16) consecutive lanes = one group
                s += __shfl_xor_sync(0xffffffffu, s, o);  ss += __shfl_xor_sync(0xffffffffu, ss, o);
            }
            if ((lane & (cpg - 1)) == 0) {
                const int img = key / p.n_slabs, slab = key - img * p.n_slabs;
                float* dst = p.gn_partial + static_cast<size_t>(img) * p.gn_img_stride +
                             (blockIdx.x * GN_SLOTS_PER_CTA + eg) * 64 + ((slab * 128 + ch) / cpg) * 2;
                dst[0] = s; dst[1] = ss;
            }
            gs = 0.f; gss = 0.f;
        };
        int acc = 0, free_i = -1;       // free_i: staging block whose store is still draining (handed back to the loader
        uint32_t acc_phase = 0, k = 0;  // half a sub-block later, so its refill runs a whole sub-block ahead of its reuse)
        for (int u = blockIdx.x; u < p.total_units; u += gridDim.x, ++k) {
            int n, slab, h0, w0;
            decode(u, n, slab, h0, w0);
            const int key = n * p.n_slabs + slab;
            if (key != gn_key) {
                if (has_gn && gn_key >= 0) gn_flush(gn_key);
                gn_key = key;
                bias_v = (p.bias && slab * 128 + ch < p.Cout) ? __ldg(p.bias + slab * 128 + ch) : 0.f;
            }
            mbar_wait(tfull(acc), acc_phase, 27);
            tc_fence_after();
            const uint32_t taddr = tmem_base + acc * 256 + eg * 128 + (static_cast<uint32_t>(q * 32) << 16);
#pragma unroll 1
            for (int sb = 0; sb < 2; ++sb) {
                uint16_t* stage = reinterpret_cast<uint16_t*>(stg_generic + sidx(w8, sb) * Cfg::STG_BYTES);
                if (has_res) {
                    mbar_wait(res_full(sidx(w8, sb)), k & 1u, 28);
                } else {
                    if (lane == 0) tma_store_wait_read<1>();          // this block's previous store (two stores ago) has drained
                    __syncwarp();
                }
#pragma unroll 1
                for (int half = 0; half < 2; ++half) {
                    uint32_t v[32];
                    tmem_ld_32x32(taddr + sb * 64 + half * 32, v);
                    tmem_ld_wait(); tmem_regs_ready(v);      // compiler fence: consumers of the async load stay below the wait
                    uint16_t* sp = stage + (half * 32) * 32 + lane;          // [pixel][32 channels]
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        float f = (__uint_as_float(v[i]) + bias_v) * out_scale;
                        if (has_res) {
                            const uint32_t r = sp[i * 32];
                            f += f16 ? __half2float(__ushort_as_half(static_cast<unsigned short>(r)))
                                     : __uint_as_float(r << 16);
                        }
                        gs += f; gss = fmaf(f, f, gss);
                        sp[i * 32] = static_cast<uint16_t>(pack_h2(f, 0.f, f16) & 0xffffu);
                    }
                    if (has_res && half == 0 && free_i >= 0) {
                        if (lane == 0) {
                            tma_store_wait_read<0>();                   // the previous sub-block's store has drained
                            mbar_arrive(buf_free(free_i));
                        }
                        free_i = -1;
                    }
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                    tma_store_4d(&maps.out, sS(w8, sb), slab * 128 + 32 * q, w0, h0 + 8 * eg + 4 * sb, n);
                    tma_store_commit();
                }
                free_i = sidx(w8, sb);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty(acc));
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1u;
        }
        if (has_gn && gn_key >= 0) gn_flush(gn_key);
        if (lane == 0) tma_store_wait_all<0>();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// Pick the output tile (TW x TH x TN = 128 pixels, powers of two).  Images: TW, then TH, by the padding of their own
// extent (ties to the larger tile), TN = what is left -- a choice that does NOT depend on N, so an episode computed in a
// batch takes the same kernel path (halo / generic mainloop, fused-statistics grouping) as computed alone and the results
// are bit-identical (tests: batched == singletons).  Token matrices (1 x M "images", N = 1, M >= 128): fewest padded rows over all
// shapes, ties to the wider tile (by its own padding a 539-row matrix -- the 7 x 77 prompt tokens of a 7-shot support
// pass -- went to TW = 1: 539 tiles of one valid row each).
void choose_tile(int W, int H, int N, int& TW, int& TH, int& TN) {
    if (H == 1 && N == 1 && W >= 128) {
        long long best = -1;
        TW = 128; TH = 1; TN = 1;
        for (int tw = 128; tw >= 1; tw /= 2) {
            const int tn = 128 / tw;
            const long long cost = (static_cast<long long>((W + tw - 1) / tw) * tw) * tn;
            if (best < 0 || cost < best) { best = cost; TW = tw; TN = tn; }
        }
        return;
    }
    auto pick = [](int extent, int cap) {
        int best = 1;
        long long best_cost = -1;
        for (int t = 1; t <= cap; t *= 2) {
            const long long cost = static_cast<long long>((extent + t - 1) / t) * t;
            if (best_cost < 0 || cost <= best_cost) { best = t; best_cost = cost; }
        }
        return best;
    };
    TW = pick(W, 128);
    TH = pick(H, 128 / TW);
    TN = 128 / (TW * TH);
}

// Host side of the T128 variant (see igemm_t128_kernel).  Returns DFW_OK after launching, or 1 when the layer is not
// eligible (the caller falls through to the generic kernel).
int try_launch_t128(const void* x, const void* w, const float* bias, const void* residual, void* y, int N, int Hin, int Win,
                    int Cin, int Cout, int ksize, int stride, int pad_mode, int flags, float out_scale, int bias_sample_stride,
                    long long w_row_stride, long long w_batch_stride, int up_phase, float* gn_partial, int gn_groups,
                    long long gn_img_stride, cudaStream_t stream, const float* gn_in = nullptr, bool dry_run = false,
                    void* xf_scratch = nullptr) {
    const bool enabled = get_option(DFW_OPT_T128) != 0;
    const int max_cout = get_option(DFW_OPT_T128_MAXC);
    // a token matrix [M, K] (linear layers arrive as a 1 x M image) is the image [M / 16, 16]: a 16 x 16 tile is then 256
    // consecutive rows
    // (measured: a gain only for the projections with a residual; the others stay on the generic kernel)
    if (ksize == 1 && N == 1 && Hin == 1 && Win % 256 == 0 && residual != nullptr) { Hin = Win / 16; Win = 16; }
    const bool ragged = Cout % 128 != 0;                 // 1x1 only: the last slab is partly outside the tensor
    if (!enabled || (ksize != 3 && ksize != 1) || stride != 1 || pad_mode != 0 || up_phase >= 0 || w_batch_stride != 0 || w_row_stride != 0 ||
        Cout % 32 != 0 || Cout < 128 || Cout > max_cout || (ragged && (ksize != 1 || gn_partial != nullptr)) ||
        (gn_partial != nullptr && Cout > 512) || Cin % BLOCK_K != 0 || Win % 16 != 0 || Hin % 16 != 0 || bias_sample_stride != 0 ||
        (flags & (DFW_EPI_OUT_F32 | DFW_EPI_RES_F32 | DFW_EPI_GEGLU | DFW_EPI_SILU)) != 0 ||
        (gn_partial != nullptr && gn_groups != 32) ||
        ((reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(residual)) & 15) != 0)
        return 1;
    const int n_slabs = (Cout + 127) / 128;
    const long long units = static_cast<long long>(N) * n_slabs * (Hin / 16) * (Win / 16);
    if (units < 4LL * sm_count() || units >= (1LL << 31)) return 1;
    if (gn_in != nullptr && Cin > T128CfgT<true>::SS_MAX_CIN) return 1;
    if (dry_run) return DFW_OK;
    T128Maps maps;
    T128Params p{};
    p.N = N; p.H = Hin; p.W = Win;
    p.ktaps = ksize;
    p.kb_per_tap = Cin / BLOCK_K;
    p.tiles_w = Win / 16; p.tiles_h = Hin / 16;
    p.total_units = static_cast<int>(units);
    p.n_slabs = n_slabs;
    p.Cout = Cout;
    p.gn_cpg = Cout / 32;
    p.bias = bias;
    p.f16 = (flags & DFW_EPI_F16) ? 1 : 0;
    p.has_res = residual != nullptr ? 1 : 0;
    p.out_scale = out_scale;
    p.gn_partial = gn_partial;
    p.gn_img_stride = gn_img_stride > 0 ? gn_img_stride : static_cast<long long>(sm_count()) * GN_SLOTS_PER_CTA * 64;
    p.gn_in = gn_in;
    p.Cin = Cin;
    p.x = x;
    p.xf_scratch = xf_scratch;
    const uint64_t esz = 2;
    const int grid = p.total_units < sm_count() ? p.total_units : sm_count();
    int rc;
    if (gn_in != nullptr) {
        if (xf_scratch == nullptr || (reinterpret_cast<uintptr_t>(xf_scratch) & 127) != 0) return DFW_ERR_INVALID;
        // the ring as a tensor [grid * XF_BUFS][18][18][64]; the patch of tap dw is the box at column dw
        const uint64_t dims[4] = {BLOCK_K, 18, 18, static_cast<uint64_t>(grid) * T128CfgT<true>::XF_BUFS};
        const uint64_t strides[3] = {BLOCK_K * esz, 18 * BLOCK_K * esz, 18 * 18 * BLOCK_K * esz};
        const uint32_t box[4] = {BLOCK_K, 16, static_cast<uint32_t>(16 + ksize - 1), 1};
        rc = encode_tmap_bf16_sw128(&maps.patch, xf_scratch, 4, dims, strides, box);
        if (rc != DFW_OK) return rc;
    } else {
        const uint64_t dims[4] = {static_cast<uint64_t>(Cin), static_cast<uint64_t>(Win), static_cast<uint64_t>(Hin),
                                  static_cast<uint64_t>(N)};
        const uint64_t strides[3] = {Cin * esz, static_cast<uint64_t>(Win) * Cin * esz,
                                     static_cast<uint64_t>(Hin) * Win * Cin * esz};
        const uint32_t box[4] = {BLOCK_K, 16, static_cast<uint32_t>(16 + ksize - 1), 1};
        rc = encode_tmap_bf16_sw128(&maps.patch, x, 4, dims, strides, box);
        if (rc != DFW_OK) return rc;
    }
    {
        const uint64_t Kt = static_cast<uint64_t>(ksize) * ksize * Cin;
        const uint64_t dims[3] = {Kt, static_cast<uint64_t>(Cout), 1};
        const uint64_t strides[2] = {Kt * esz, Kt * Cout * esz};
        const uint32_t box[3] = {BLOCK_K, 128, 1};
        rc = encode_tmap_bf16_sw128(&maps.w, w, 3, dims, strides, box);
        if (rc != DFW_OK) return rc;
    }
    {
        const uint64_t dims[4] = {static_cast<uint64_t>(Cout), static_cast<uint64_t>(Win), static_cast<uint64_t>(Hin),
                                  static_cast<uint64_t>(N)};
        const uint64_t strides[3] = {Cout * esz, static_cast<uint64_t>(Win) * Cout * esz,
                                     static_cast<uint64_t>(Hin) * Win * Cout * esz};
        const uint32_t box[4] = {32, 16, 4, 1};
        rc = encode_tmap(&maps.out, y, 2, 0, 4, dims, strides, box);
        if (rc != DFW_OK) return rc;
        if (residual != nullptr) {
            rc = encode_tmap(&maps.res, residual, 2, 0, 4, dims, strides, box);
            if (rc != DFW_OK) return rc;
        } else {
            maps.res = maps.out;
        }
    }
    static bool attr_set = false;
    if (!attr_set) {
        DFW_CHECK_CUDA(cudaFuncSetAttribute(igemm_t128_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, T128CfgT<false>::SMEM_BYTES));
        DFW_CHECK_CUDA(cudaFuncSetAttribute(igemm_t128_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, T128CfgT<true>::SMEM_BYTES));
        attr_set = true;
    }
    if (gn_in != nullptr)
        DFW_CHECK_CUDA(launch_k(igemm_t128_kernel<true>, grid, T128CfgT<true>::THREADS, T128CfgT<true>::SMEM_BYTES, stream, maps, p));
    else
        DFW_CHECK_CUDA(launch_k(igemm_t128_kernel<false>, grid, T128CfgT<false>::THREADS, T128CfgT<false>::SMEM_BYTES, stream, maps, p));
    g_launches.fetch_add(1);
    DFW_CHECK_CUDA(cudaGetLastError());
    return DFW_OK;
}

template <int BLOCK_N, int TPU = 1>
int launch_igemm(const IgemmMaps& maps, IgemmParams& p, cudaStream_t stream) {
    using Cfg = IgemmCfg<BLOCK_N, TPU>;
    static bool attr_set = false;
    if (!attr_set) {
        DFW_CHECK_CUDA(cudaFuncSetAttribute(igemm_kernel<BLOCK_N, TPU>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            Cfg::SMEM_BYTES));
        attr_set = true;
    }
    p.n_tiles = (p.Cout + BLOCK_N - 1) / BLOCK_N;
    p.total_tiles = p.tiles_w * p.tiles_h * p.tiles_nimg * p.n_tiles;
    const int units = (p.total_tiles + TPU - 1) / TPU;
    const int grid = units < sm_count() ? units : sm_count();
    DFW_CHECK_CUDA(launch_k(igemm_kernel<BLOCK_N, TPU>, grid, IGEMM_THREADS, Cfg::SMEM_BYTES, stream, maps, p));
    g_launches.fetch_add(1);
    DFW_CHECK_CUDA(cudaGetLastError());
    return DFW_OK;
}

int igemm_dispatch(const void* x, const void* w, const float* bias, int bias_sample_stride, const void* residual,
                   void* y, int N, int Hin, int Win, int Cin, int Cout, int ksize, int stride, int pad_mode,
                   int flags, float out_scale, cudaStream_t stream, long long w_row_stride = 0,
                   long long w_batch_stride = 0, int up_phase = -1, float* gn_partial = nullptr, int gn_groups = 0,
                   long long gn_img_stride = 0) {
    // up_phase >= 0: this launch computes output phase (ph, pw) = (up_phase / 2, up_phase % 2) of a
    // "nearest-2x upsample -> 3x3 conv": a 2x2-tap convolution over the LOW-resolution input x (weights pre-summed
    // on the host, K order = (a*2 + b)*Cin + c), written to y[:, ph::2, pw::2, :] of the [N, 2H, 2W, Cout] output.
    int rc = require_sm100();
    if (rc != DFW_OK) return rc;
    DFW_REQUIRE(x && w && y);
    DFW_REQUIRE(N > 0 && Hin > 0 && Win > 0 && Cout > 0);
    DFW_REQUIRE(Cin > 0 && Cin % BLOCK_K == 0);
    DFW_REQUIRE(ksize == 1 || ksize == 3);
    DFW_REQUIRE(stride == 1 || stride == 2);
    DFW_REQUIRE(pad_mode == 0 || (pad_mode == 1 && stride == 2 && ksize == 3));
    DFW_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(w) & 15) == 0);
    if (stride == 2) DFW_REQUIRE(ksize == 3 && Hin % 2 == 0 && Win % 2 == 0);
    const bool geglu = (flags & DFW_EPI_GEGLU) != 0;
    if (geglu) DFW_REQUIRE(Cout % 256 == 0 && residual == nullptr && !(flags & DFW_EPI_OUT_F32));

    // Cout = 128 3x3 stride-1 layers on large images: channel-major accumulator kernel
    rc = try_launch_t128(x, w, bias, residual, y, N, Hin, Win, Cin, Cout, ksize, stride, pad_mode, flags, out_scale,
                         bias_sample_stride, w_row_stride, w_batch_stride, up_phase, gn_partial, gn_groups, gn_img_stride,
                         stream);
    if (rc != 1) return rc;

    IgemmMaps maps;
    IgemmParams p{};
    const int Hout = Hin / stride, Wout = Win / stride;
    p.N = N; p.H = Hout; p.W = Wout; p.Cout = Cout;
    p.kb_per_tap = Cin / BLOCK_K;
    p.ntaps = (up_phase >= 0) ? 4 : ksize * ksize;
    if (up_phase >= 0) DFW_REQUIRE(ksize == 3 && stride == 1 && pad_mode == 0 && residual == nullptr);
    choose_tile(Wout, Hout, N, p.TW, p.TH, p.TN);
    // halo mainloop: 3x3 / stride 1 on images that tile exactly into 8 x 16 output patches
    const bool halo_enabled = get_option(DFW_OPT_HALO) != 0;
    p.halo = (halo_enabled && ksize == 3 && stride == 1 && up_phase < 0 && w_batch_stride == 0 && Wout % 16 == 0 &&
              Hout % 8 == 0) ? 1 : 0;
    if (p.halo) { p.TW = 16; p.TH = 8; p.TN = 1; }
    if (w_batch_stride > 0 && p.TN != 1) {       // per-image weights: a tile must not straddle images
        p.TN = 1;
        if (Hout == 1) { p.TW = 128; p.TH = 1; } else { p.TH = 128 / p.TW; }
    }
    p.tiles_w = (Wout + p.TW - 1) / p.TW;
    p.tiles_h = (Hout + p.TH - 1) / p.TH;
    p.tiles_nimg = (N + p.TN - 1) / p.TN;
    p.bias = bias; p.bias_sample_stride = bias_sample_stride;
    p.residual = residual; p.out = y;
    p.out_ch = geglu ? Cout / 2 : Cout;
    p.out_scale = out_scale; p.flags = flags;

    const uint32_t box[4] = {BLOCK_K, static_cast<uint32_t>(p.TW), static_cast<uint32_t>(p.TH),
                             static_cast<uint32_t>(p.TN)};
    const uint64_t esz = 2;
    if (stride == 1) {
        const uint64_t dims[4] = {static_cast<uint64_t>(Cin), static_cast<uint64_t>(Win),
                                  static_cast<uint64_t>(Hin), static_cast<uint64_t>(N)};
        const uint64_t strides[3] = {Cin * esz, static_cast<uint64_t>(Win) * Cin * esz,
                                     static_cast<uint64_t>(Hin) * Win * Cin * esz};
        rc = encode_tmap_bf16_sw128(&maps.a[0], x, 4, dims, strides, box);
        if (rc != DFW_OK) return rc;
        for (int i = 1; i < 4; ++i) maps.a[i] = maps.a[0];
        if (p.halo) {       // a[1]: (TH+2) x TW patch box for the halo mainloop
            const uint32_t pbox[4] = {BLOCK_K, 16, 10, 1};
            rc = encode_tmap_bf16_sw128(&maps.a[1], x, 4, dims, strides, pbox);
            if (rc != DFW_OK) return rc;
        }
        const int pad = (ksize - 1) / 2;
        if (up_phase >= 0) {
            const int ph = up_phase >> 1, pw = up_phase & 1;
            for (int a = 0; a < 2; ++a)
                for (int b = 0; b < 2; ++b) {
                    const int t = a * 2 + b;
                    p.tap_map[t] = 0; p.tap_dh[t] = a + ph - 1; p.tap_dw[t] = b + pw - 1;
                }
        } else {
            for (int kh = 0; kh < ksize; ++kh)
                for (int kw = 0; kw < ksize; ++kw) {
                    const int t = kh * ksize + kw;
                    p.tap_map[t] = 0; p.tap_dh[t] = kh - pad; p.tap_dw[t] = kw - pad;
                }
        }
    } else {
        // stride 2: four phase views x[:, ph::2, pw::2, :] of the input, each a plain stride-1 TMA tensor.
        const uint64_t dims[4] = {static_cast<uint64_t>(Cin), static_cast<uint64_t>(Win / 2),
                                  static_cast<uint64_t>(Hin / 2), static_cast<uint64_t>(N)};
        const uint64_t strides[3] = {2 * Cin * esz, 2 * static_cast<uint64_t>(Win) * Cin * esz,
                                     static_cast<uint64_t>(Hin) * Win * Cin * esz};
        for (int ph = 0; ph < 2; ++ph)
            for (int pw = 0; pw < 2; ++pw) {
                const uint8_t* base =
                    reinterpret_cast<const uint8_t*>(x) + (static_cast<uint64_t>(ph) * Win + pw) * Cin * esz;
                rc = encode_tmap_bf16_sw128(&maps.a[ph * 2 + pw], base, 4, dims, strides, box);
                if (rc != DFW_OK) return rc;
            }
        auto phase_of = [&](int k, int& ph, int& d) {
            if (pad_mode == 0) { ph = (k + 1) & 1; d = (k == 0) ? -1 : 0; }   // in = 2*o + k - 1
            else               { ph = k & 1;       d = (k == 2) ? 1 : 0; }    // in = 2*o + k
        };
        for (int kh = 0; kh < 3; ++kh)
            for (int kw = 0; kw < 3; ++kw) {
                int ph, pw, dh, dw;
                phase_of(kh, ph, dh);
                phase_of(kw, pw, dw);
                const int t = kh * 3 + kw;
                p.tap_map[t] = ph * 2 + pw; p.tap_dh[t] = dh; p.tap_dw[t] = dw;
            }
    }

    // Cout tile width: among the widths that divide Cout, the one with the cheapest schedule on this many SMs
    // (rounds of the persistent loop x per-tile cost) -- e.g. 1280 channels on 32 pixel tiles: 256 -> 160 tiles = 2 rounds
    // of 256 columns, 160 -> 256 tiles = 2 rounds of 160 columns; on 8 pixel tiles 128 -> 80 tiles = 1 round of 128.
    int block_n;
    if (geglu) block_n = 256;
    else if (Cout <= 16) block_n = 16;
    else {
        const long long m_tiles = static_cast<long long>(p.tiles_w) * p.tiles_h * p.tiles_nimg;
        const int cand[3] = {256, 160, 128};
        double best = 0.0;
        block_n = 0;
        for (int i = 0; i < 3; ++i) {
            const int bn = cand[i];
            if (Cout % bn != 0 && !(bn == 128 && block_n == 0)) continue;     // 128 is the ragged-Cout fallback
            const long long tiles = m_tiles * ((Cout + bn - 1) / bn);
            const long long rounds = (tiles + sm_count() - 1) / sm_count();
            // per-tile cost in accumulator columns: + fixed pipeline fill / drain; N = 128 MMAs run ~15 % below the
            // N = 256 rate per column (measured)
            const double cost = static_cast<double>(rounds) * (bn + 24) * (bn == 128 ? 1.15 : 1.0);
            if (block_n == 0 || cost < best - 1e-9) { best = cost; block_n = bn; }
        }
    }
    {
        const uint64_t Kt = static_cast<uint64_t>(p.ntaps) * Cin;
        const uint64_t row = w_row_stride > 0 ? static_cast<uint64_t>(w_row_stride) : Kt;
        p.w_batched = w_batch_stride > 0 ? 1 : 0;
        if (p.w_batched) DFW_REQUIRE(p.TN == 1 && (w_batch_stride * esz) % 16 == 0);   // one image per tile
        DFW_REQUIRE((row * esz) % 16 == 0);
        const uint64_t dims[3] = {Kt, static_cast<uint64_t>(Cout), static_cast<uint64_t>(p.w_batched ? N : 1)};
        const uint64_t strides[2] = {row * esz, (p.w_batched ? static_cast<uint64_t>(w_batch_stride) : row * Cout) * esz};
        const uint32_t bbox[3] = {BLOCK_K, static_cast<uint32_t>(block_n), 1};
        rc = encode_tmap_bf16_sw128(&maps.b, w, 3, dims, strides, bbox);
        if (rc != DFW_OK) return rc;
    }
    // TMA-store epilogue whenever the output (and residual) rows are TMA-addressable and of one element size
    {
        const bool out_f32 = (flags & DFW_EPI_OUT_F32) != 0, res_f32 = (flags & DFW_EPI_RES_F32) != 0;
        const uint64_t oesz = out_f32 ? 4 : 2;
        const bool ok = block_n != 16 && bias_sample_stride == 0 && (p.out_ch * oesz) % 16 == 0 &&
                        (reinterpret_cast<uintptr_t>(y) & 15) == 0 &&
                        (residual == nullptr || (res_f32 == out_f32 && (reinterpret_cast<uintptr_t>(residual) & 15) == 0));
        p.tma_epi = ok ? 1 : 0;
        p.has_res = (ok && residual != nullptr) ? 1 : 0;
        p.gn_partial = nullptr;
        p.gn_cpg = 0;
        if (gn_partial != nullptr) {
            // fused GroupNorm statistics: TMA epilogue, one image per tile, channels-per-group in {4, 8, 16}
            DFW_REQUIRE(ok && p.TN == 1 && gn_groups == 32 && p.out_ch % 32 == 0 && !geglu);
            const int cpg = p.out_ch / 32;
            DFW_REQUIRE(cpg == 4 || cpg == 8 || cpg == 16);
            p.gn_partial = gn_partial;
            p.gn_cpg = cpg;
            p.gn_img_stride = gn_img_stride > 0 ? gn_img_stride : static_cast<long long>(sm_count()) * GN_SLOTS_PER_CTA * 64;
        }
        if (up_phase >= 0) DFW_REQUIRE(ok);          // the strided phase view is only reachable through the TMA store
        if (ok && up_phase >= 0) {
            const int ph = up_phase >> 1, pw = up_phase & 1;
            const uint64_t W2 = 2ull * Wout, H2 = 2ull * Hout;
            const uint64_t dims[4] = {static_cast<uint64_t>(p.out_ch), static_cast<uint64_t>(Wout),
                                      static_cast<uint64_t>(Hout), static_cast<uint64_t>(N)};
            const uint64_t strides[3] = {2 * p.out_ch * oesz, 2 * W2 * p.out_ch * oesz, H2 * W2 * p.out_ch * oesz};
            const uint32_t obox[4] = {static_cast<uint32_t>(64 / oesz), static_cast<uint32_t>(p.TW),
                                      static_cast<uint32_t>(p.TH), static_cast<uint32_t>(p.TN)};
            uint8_t* base = reinterpret_cast<uint8_t*>(y) + (static_cast<uint64_t>(ph) * W2 + pw) * p.out_ch * oesz;
            rc = encode_tmap(&maps.out, base, static_cast<int>(oesz), 64, 4, dims, strides, obox);
            if (rc != DFW_OK) return rc;
            maps.res = maps.out;
        } else if (ok) {
            const uint64_t dims[4] = {static_cast<uint64_t>(p.out_ch), static_cast<uint64_t>(Wout),
                                      static_cast<uint64_t>(Hout), static_cast<uint64_t>(N)};
            const uint64_t strides[3] = {p.out_ch * oesz, static_cast<uint64_t>(Wout) * p.out_ch * oesz,
                                         static_cast<uint64_t>(Hout) * Wout * p.out_ch * oesz};
            const uint32_t obox[4] = {static_cast<uint32_t>(64 / oesz), static_cast<uint32_t>(p.TW),
                                      static_cast<uint32_t>(p.TH), static_cast<uint32_t>(p.TN)};
            rc = encode_tmap(&maps.out, y, static_cast<int>(oesz), 64, 4, dims, strides, obox);
            if (rc != DFW_OK) return rc;
            if (residual != nullptr) {
                rc = encode_tmap(&maps.res, residual, static_cast<int>(oesz), 64, 4, dims, strides, obox);
                if (rc != DFW_OK) return rc;
            } else {
                maps.res = maps.out;
            }
        } else {
            maps.out = maps.b;
            maps.res = maps.b;
        }
    }
    const long long m_tiles = static_cast<long long>(p.tiles_w) * p.tiles_h * p.tiles_nimg;
    p.m_tiles = static_cast<int>(m_tiles);
    // weight-stationary mode: 1x1 / linear layers whose whole K extent of one Cout tile fits beside >= 3 activation slots
    // (K = 320 at N = 160, K <= 512 at N = 128) and that have enough pixel tiles for every CTA to reuse the weights.
    // MEASURED (scripts/bench_linear.py, profiles/r02_bench_linear_stationary.log): bit-identical, and no faster -- 32.7 vs
    // 30.7 us at M 65536 K 320 N 320, 128.8-130.5 vs 129.8-130.2 ms per step: re-reading the 100 KB weight tile per pixel tile
    // from L2 was not what bounds the K = 320 token GEMMs (they are short kernels: ramp + HBM).  Off by default.
    auto try_resident = [&](int stages, int stage_bytes, int b_tile_bytes) {
        p.b_resident = 0; p.a_slots = 0;
        if (!get_option(DFW_OPT_B_RESIDENT) || p.ntaps != 1 || p.w_batched || p.gn_partial != nullptr || up_phase >= 0) return;
        const int ring = stages * stage_bytes, b_bytes = p.kb_per_tap * b_tile_bytes;
        int a = (ring - b_bytes) / A_TILE_BYTES;
        if (a > stages) a = stages;
        if (a > p.kb_per_tap) a = p.kb_per_tap;
        if (b_bytes < ring && a >= 3 && m_tiles >= 2LL * sm_count() && (Cout + block_n - 1) / block_n >= 1) {
            p.b_resident = 1; p.a_slots = a;
        }
    };
    switch (block_n) {
        case 16: return launch_igemm<16>(maps, p, stream);
        case 128: {
            // paired tiles when there is a single Cout tile and enough work to keep every SM busy with pairs
            if (Cout <= 128 && m_tiles >= 4LL * sm_count()) return launch_igemm<128, 2>(maps, p, stream);
            try_resident(IgemmCfg<128>::STAGES, IgemmCfg<128>::STAGE_BYTES, IgemmCfg<128>::B_TILE_BYTES);
            return launch_igemm<128>(maps, p, stream);
        }
        case 160:
            try_resident(IgemmCfg<160>::STAGES, IgemmCfg<160>::STAGE_BYTES, IgemmCfg<160>::B_TILE_BYTES);
            return launch_igemm<160>(maps, p, stream);
        default: return launch_igemm<256>(maps, p, stream);
    }
}

}  // namespace
}  // namespace dfw

extern "C" {

int dfw_conv2d_igemm(const void* x, const void* w, const float* bias, int bias_sample_stride, const void* residual,
                     void* y, int N, int Hin, int Win, int Cin, int Cout, int ksize, int stride, int pad_mode,
                     int flags, float out_scale, void* stream) {
    return dfw::igemm_dispatch(x, w, bias, bias_sample_stride, residual, y, N, Hin, Win, Cin, Cout, ksize, stride,
                               pad_mode, flags, out_scale, static_cast<cudaStream_t>(stream));
}

int dfw_upconv2x_igemm(const void* x, const void* w4, const float* bias, void* y, int N, int Hin, int Win, int Cin,
                       int Cout, int flags, float* gn_partial, void* stream) {
    // four phase convolutions (2x2 taps each) = nearest-2x upsample followed by a 3x3 / pad 1 convolution
    const size_t esz = 2;
    // partial layout [N][4 phases x #SMs][32][2]: each phase launch owns a slice of #SMs slots inside every image
    const long long slots = 4LL * dfw::sm_count() * dfw::GN_SLOTS_PER_CTA;
    if (gn_partial) {
        if (cudaMemsetAsync(gn_partial, 0, static_cast<size_t>(N) * slots * 64 * sizeof(float),
                            static_cast<cudaStream_t>(stream)) != cudaSuccess) return DFW_ERR_CUDA;
    }
    for (int phase = 0; phase < 4; ++phase) {
        const uint8_t* wp = reinterpret_cast<const uint8_t*>(w4) + static_cast<size_t>(phase) * Cout * 4 * Cin * esz;
        int rc = dfw::igemm_dispatch(x, wp, bias, 0, nullptr, y, N, Hin, Win, Cin, Cout, 3, 1, 0, flags, 1.0f,
                                     static_cast<cudaStream_t>(stream), 0, 0, phase,
                                     gn_partial ? gn_partial + static_cast<size_t>(phase) * dfw::sm_count() * dfw::GN_SLOTS_PER_CTA * 64 : nullptr,
                                     gn_partial ? 32 : 0, slots * 64);
        if (rc != DFW_OK) return rc;
    }
    return DFW_OK;
}

int dfw_conv_gnstats_supported(int N, int Hout, int Wout, int Cout) {
    if (N <= 0 || Hout <= 0 || Wout <= 0 || Cout % 32 != 0) return 0;
    const int cpg = Cout / 32;
    if (cpg != 4 && cpg != 8 && cpg != 16) return 0;
    int TW, TH, TN;
    dfw::choose_tile(Wout, Hout, N, TW, TH, TN);
    return TN == 1 ? 1 : 0;
}

long long dfw_gn_partial_floats(int N) {
    return N > 0 ? static_cast<long long>(N) * dfw::sm_count() * dfw::GN_SLOTS_PER_CTA * 64 : -1;
}

int dfw_conv2d_igemm_gnstats(const void* x, const void* w, const float* bias, const void* residual, void* y, int N,
                             int Hin, int Win, int Cin, int Cout, int ksize, int stride, int pad_mode, int flags,
                             float out_scale, float* gn_partial, void* stream) {
    if (gn_partial == nullptr) return DFW_ERR_INVALID;
    if (cudaMemsetAsync(gn_partial, 0, static_cast<size_t>(dfw_gn_partial_floats(N)) * sizeof(float),
                        static_cast<cudaStream_t>(stream)) != cudaSuccess) return DFW_ERR_CUDA;
    return dfw::igemm_dispatch(x, w, bias, 0, residual, y, N, Hin, Win, Cin, Cout, ksize, stride, pad_mode, flags,
                               out_scale, static_cast<cudaStream_t>(stream), 0, 0, -1, gn_partial, 32);
}

int dfw_conv_gnin_supported(int N, int H, int W, int Cin, int Cout, int ksize) {
    if (Cin > dfw::T128CfgT<true>::SS_MAX_CIN) return 0;       // the image's scale / shift vectors live in shared memory
    return dfw_conv_t128_eligible(N, H, W, Cin, Cout, ksize);
}

int dfw_conv_t128_eligible(int N, int H, int W, int Cin, int Cout, int ksize) {
    if (dfw::require_sm100() != DFW_OK) return 0;
    static unsigned char dummy[16] __attribute__((aligned(16)));
    return dfw::try_launch_t128(dummy, dummy, nullptr, nullptr, dummy, N, H, W, Cin, Cout, ksize, 1, 0, DFW_EPI_F16, 1.0f, 0,
                                0, 0, -1, nullptr, 32, 0, nullptr, nullptr, true) == DFW_OK ? 1 : 0;
}

int dfw_conv2d_igemm_gnin(const void* x, const float* gn_scale_shift, const void* w, const float* bias,
                          const void* residual, void* y, int N, int Hin, int Win, int Cin, int Cout, int ksize, int flags,
                          float* gn_partial_out, void* scratch, void* stream) {
    int rc = dfw::require_sm100();
    if (rc != DFW_OK) return rc;
    if (!x || !gn_scale_shift || !w || !y || !scratch) return DFW_ERR_INVALID;
    if (gn_partial_out != nullptr &&
        cudaMemsetAsync(gn_partial_out, 0, static_cast<size_t>(dfw_gn_partial_floats(N)) * sizeof(float),
                        static_cast<cudaStream_t>(stream)) != cudaSuccess) return DFW_ERR_CUDA;
    rc = dfw::try_launch_t128(x, w, bias, residual, y, N, Hin, Win, Cin, Cout, ksize, 1, 0, flags, 1.0f, 0, 0, 0, -1,
                              gn_partial_out, 32, 0, static_cast<cudaStream_t>(stream), gn_scale_shift, false, scratch);
    return rc == 1 ? DFW_ERR_INVALID : rc;        // 1 = shape not eligible (ask dfw_conv_gnin_supported first)
}

long long dfw_conv_gnin_scratch_bytes(void) {
    return static_cast<long long>(dfw::sm_count()) * dfw::T128CfgT<true>::XF_BUFS * dfw::T128CfgT<true>::XF_BUF_BYTES;
}

int dfw_bmm_nt(const void* x, const void* w, long long w_row_stride, long long w_batch_stride, const float* bias,
               void* y, int B, int M, int K, int Nout, int flags, float out_scale, void* stream) {
    // y[b] (M x Nout) = x[b] (M x K) @ w[b]^T, w[b] element (n, k) at w[b*w_batch_stride + n*w_row_stride + k]
    return dfw::igemm_dispatch(x, w, bias, 0, nullptr, y, B, 1, M, K, Nout, 1, 1, 0, flags, out_scale,
                               static_cast<cudaStream_t>(stream), w_row_stride, w_batch_stride);
}

int dfw_linear(const void* x, const void* w, const float* bias, const void* residual, void* y, int M, int K,
               int Nout, int flags, float out_scale, void* stream) {
    // a Linear is a 1x1 "convolution" over an image of height 1 and width M
    return dfw::igemm_dispatch(x, w, bias, 0, residual, y, 1, 1, M, K, Nout, 1, 1, 0, flags, out_scale,
                               static_cast<cudaStream_t>(stream));
}

}  // extern "C"

#if DFW_GNIN_TRACE
extern "C" int dfw_debug_t128_trace(long long* out16, int reset) {
    cudaDeviceSynchronize();
    if (cudaMemcpyFromSymbol(out16, g_t128_trace, sizeof(long long) * 16) != cudaSuccess) return -2;
    if (reset) { long long z[16] = {0}; cudaMemcpyToSymbol(g_t128_trace, z, sizeof(z)); }
    return 0;
}
#endif
