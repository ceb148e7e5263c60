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

#include "common.cuh"
#include "ptx.cuh"

namespace dfw {
extern std::atomic<long long> g_launches;

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;
constexpr int A_TILE_BYTES = BLOCK_M * BLOCK_K * 2;  // 16 KiB
constexpr int IGEMM_THREADS = 256;
constexpr int MAX_TAPS = 9;

struct IgemmMaps {
    CUtensorMap a[4];
    CUtensorMap b;
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
};

template <int BLOCK_N>
struct IgemmCfg {
    static constexpr int B_TILE_BYTES = BLOCK_N * 128;
    static constexpr int STAGE_BYTES = A_TILE_BYTES + B_TILE_BYTES;
    static constexpr int STAGES = (BLOCK_N >= 256) ? 4 : (BLOCK_N >= 128 ? 6 : 8);
    static constexpr int ACC_COLS = BLOCK_N <= 32 ? 32 : (BLOCK_N <= 64 ? 64 : (BLOCK_N <= 128 ? 128 : 256));
    static constexpr int TMEM_COLS = 2 * ACC_COLS;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
    static_assert(STAGE_BYTES % 1024 == 0, "stage must keep 1024B alignment for SWIZZLE_128B");
    static_assert(SMEM_BYTES <= 232448, "smem budget");
};

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float silu(float x) { return x / (1.0f + __expf(-x)); }

struct TileCoord {
    int n_tile, n0, h0, w0;
};
__device__ __forceinline__ TileCoord decode_tile(const IgemmParams& p, int t) {
    TileCoord c;
    c.n_tile = t % p.n_tiles;
    int m = t / p.n_tiles;
    c.w0 = (m % p.tiles_w) * p.TW;
    m /= p.tiles_w;
    c.h0 = (m % p.tiles_h) * p.TH;
    c.n0 = (m / p.tiles_h) * p.TN;
    return c;
}

template <int CH>
__device__ __forceinline__ void epilogue_chunk(const IgemmParams& p, float (&f)[CH], int col_in0, int col_out0,
                                               int img, long long pix, bool row_valid) {
    // f: raw accumulators for weight rows [col_in0, col_in0+CH). Output columns [col_out0, col_out0+CH).
    if (!row_valid || col_in0 >= p.Cout) return;
    const int ncols = min(CH, p.Cout - col_in0);
    if (p.bias != nullptr) {
        const float* bp = p.bias + static_cast<long long>(img) * p.bias_sample_stride + col_in0;
#pragma unroll
        for (int j = 0; j < CH; ++j)
            if (j < ncols) f[j] += __ldg(bp + j);
    }
    if (p.out_scale != 1.0f) {
#pragma unroll
        for (int j = 0; j < CH; ++j) f[j] *= p.out_scale;
    }
    if (p.flags & DFW_EPI_SILU) {
#pragma unroll
        for (int j = 0; j < CH; ++j) f[j] = silu(f[j]);
    }
    const long long off = pix * p.out_ch + col_out0;
    const bool vec_ok = (ncols == CH) && (p.out_ch % 8 == 0);
    if (p.residual != nullptr) {
        if (p.flags & DFW_EPI_RES_F32) {
            const float* rp = reinterpret_cast<const float*>(p.residual) + off;
            if (vec_ok) {
#pragma unroll
                for (int j = 0; j < CH; j += 4) {
                    float4 r = *reinterpret_cast<const float4*>(rp + j);
                    f[j] += r.x; f[j + 1] += r.y; f[j + 2] += r.z; f[j + 3] += r.w;
                }
            } else {
                for (int j = 0; j < ncols; ++j) f[j] += rp[j];
            }
        } else {
            const __nv_bfloat16* rp = reinterpret_cast<const __nv_bfloat16*>(p.residual) + off;
            if (vec_ok) {
#pragma unroll
                for (int j = 0; j < CH; j += 8) {
                    uint4 r = *reinterpret_cast<const uint4*>(rp + j);
                    f[j] += bf16_lo(r.x); f[j + 1] += bf16_hi(r.x);
                    f[j + 2] += bf16_lo(r.y); f[j + 3] += bf16_hi(r.y);
                    f[j + 4] += bf16_lo(r.z); f[j + 5] += bf16_hi(r.z);
                    f[j + 6] += bf16_lo(r.w); f[j + 7] += bf16_hi(r.w);
                }
            } else {
                for (int j = 0; j < ncols; ++j) f[j] += __bfloat162float(rp[j]);
            }
        }
    }
    if (p.flags & DFW_EPI_OUT_F32) {
        float* op = reinterpret_cast<float*>(p.out) + off;
        if (vec_ok) {
#pragma unroll
            for (int j = 0; j < CH; j += 4)
                *reinterpret_cast<float4*>(op + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
        } else {
            for (int j = 0; j < ncols; ++j) op[j] = f[j];
        }
    } else {
        __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(p.out) + off;
        if (vec_ok) {
#pragma unroll
            for (int j = 0; j < CH; j += 8) {
                uint4 o;
                o.x = pack_bf16x2(f[j], f[j + 1]);
                o.y = pack_bf16x2(f[j + 2], f[j + 3]);
                o.z = pack_bf16x2(f[j + 4], f[j + 5]);
                o.w = pack_bf16x2(f[j + 6], f[j + 7]);
                *reinterpret_cast<uint4*>(op + j) = o;
            }
        } else {
            for (int j = 0; j < ncols; ++j) op[j] = __float2bfloat16_rn(f[j]);
        }
    }
}

template <int BLOCK_N>
__global__ void __launch_bounds__(IGEMM_THREADS, 1)
igemm_kernel(const __grid_constant__ IgemmMaps maps, const __grid_constant__ IgemmParams p) {
    using Cfg = IgemmCfg<BLOCK_N>;
    constexpr int STAGES = Cfg::STAGES;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_raw_u32 = smem_u32(smem_raw);
    const uint32_t smem_base = (smem_raw_u32 + 1023u) & ~1023u;
    const uint32_t bar_base = smem_base + STAGES * Cfg::STAGE_BYTES;
    auto sA = [&](int s) { return smem_base + s * Cfg::STAGE_BYTES; };
    auto sB = [&](int s) { return smem_base + s * Cfg::STAGE_BYTES + A_TILE_BYTES; };
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
    auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + a); };
    auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + 2 + a); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 4);
    volatile uint32_t* tmem_slot_ptr =
        reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_raw_u32));

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        for (int i = 0; i < 4; ++i) tma_prefetch_desc(&maps.a[i]);
        tma_prefetch_desc(&maps.b);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tfull_bar(a), 1);
            mbar_init(tempty_bar(a), 4);
        }
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

    const int kblocks = p.ntaps * p.kb_per_tap;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
                const TileCoord tc = decode_tile(p, t);
                for (int tap = 0; tap < p.ntaps; ++tap) {
                    const CUtensorMap* am = &maps.a[p.tap_map[tap]];
                    const int hh = tc.h0 + p.tap_dh[tap];
                    const int ww = tc.w0 + p.tap_dw[tap];
                    for (int kb = 0; kb < p.kb_per_tap; ++kb) {
                        mbar_wait(empty_bar(stage), phase ^ 1u, 1);
                        mbar_arrive_expect_tx(full_bar(stage), Cfg::STAGE_BYTES);
                        tma_load_4d(sA(stage), am, full_bar(stage), kb * BLOCK_K, ww, hh, tc.n0);
                        tma_load_2d(sB(stage), &maps.b, full_bar(stage), (tap * p.kb_per_tap + kb) * BLOCK_K,
                                    tc.n_tile * BLOCK_N);
                        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_bf16(BLOCK_M, BLOCK_N, 0);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
                mbar_wait(tempty_bar(acc), acc_phase ^ 1u, 2);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * Cfg::ACC_COLS;
                for (int kbi = 0; kbi < kblocks; ++kbi) {
                    mbar_wait(full_bar(stage), phase, 3);
                    tc_fence_after();
                    const uint64_t adesc = umma_desc_sw128(sA(stage));
                    const uint64_t bdesc = umma_desc_sw128(sB(stage));
#pragma unroll
                    for (int k = 0; k < BLOCK_K / 16; ++k)
                        umma_ss(d_tmem, adesc + 2u * k, bdesc + 2u * k, idesc, (kbi > 0 || k > 0) ? 1u : 0u);
                    tc_commit(empty_bar(stage));
                    if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                }
                tc_commit(tfull_bar(acc));
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1u;
            }
        }
    } else if (warp >= 4) {
        const int q = warp - 4;
        const int row = q * 32 + lane;
        const int tn = row / (p.TH * p.TW);
        const int rem = row % (p.TH * p.TW);
        const int th = rem / p.TW;
        const int tw = rem % p.TW;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
            const TileCoord tc = decode_tile(p, t);
            const int img = tc.n0 + tn, hh = tc.h0 + th, ww = tc.w0 + tw;
            const bool row_valid = (img < p.N) && (hh < p.H) && (ww < p.W);
            const long long pix = (static_cast<long long>(img) * p.H + hh) * p.W + ww;
            mbar_wait(tfull_bar(acc), acc_phase, 4);
            tc_fence_after();
            const uint32_t taddr = tmem_base + acc * Cfg::ACC_COLS + (static_cast<uint32_t>(q * 32) << 16);
            if constexpr (BLOCK_N == 256) {
                if (p.flags & DFW_EPI_GEGLU) {
#pragma unroll 1
                    for (int c = 0; c < 4; ++c) {
                        uint32_t v[32], g[32];
                        tmem_ld_32x32(taddr + c * 32, v);
                        tmem_ld_32x32(taddr + 128 + c * 32, g);
                        tmem_ld_wait();
                        if (row_valid) {
                            float f[32];
                            const int cv = tc.n_tile * 256 + c * 32;
                            const float* bv = p.bias ? p.bias + cv : nullptr;
#pragma unroll
                            for (int j = 0; j < 32; ++j) {
                                float val = __uint_as_float(v[j]);
                                float gate = __uint_as_float(g[j]);
                                if (bv) { val += __ldg(bv + j); gate += __ldg(bv + 128 + j); }
                                f[j] = val * gelu_erf(gate);
                            }
                            __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(p.out) + pix * p.out_ch +
                                                tc.n_tile * 128 + c * 32;
#pragma unroll
                            for (int j = 0; j < 32; j += 8) {
                                uint4 o;
                                o.x = pack_bf16x2(f[j], f[j + 1]);
                                o.y = pack_bf16x2(f[j + 2], f[j + 3]);
                                o.z = pack_bf16x2(f[j + 4], f[j + 5]);
                                o.w = pack_bf16x2(f[j + 6], f[j + 7]);
                                *reinterpret_cast<uint4*>(op + j) = o;
                            }
                        }
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(tempty_bar(acc));
                    acc ^= 1;
                    if (acc == 0) acc_phase ^= 1u;
                    continue;
                }
            }
            if constexpr (BLOCK_N >= 32) {
#pragma unroll 1
                for (int c = 0; c < BLOCK_N / 32; ++c) {
                    uint32_t v[32];
                    tmem_ld_32x32(taddr + c * 32, v);
                    tmem_ld_wait();
                    float f[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
                    const int col0 = tc.n_tile * BLOCK_N + c * 32;
                    epilogue_chunk<32>(p, f, col0, col0, img, pix, row_valid);
                }
            } else {
                uint32_t v[16];
                tmem_ld_32x16(taddr, v);
                tmem_ld_wait();
                float f[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]);
                const int col0 = tc.n_tile * BLOCK_N;
                epilogue_chunk<16>(p, f, col0, col0, img, pix, row_valid);
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

// Pick the output tile (TW x TH x TN = 128 pixels) that wastes the fewest out-of-range pixels.
void choose_tile(int W, int H, int N, int& TW, int& TH, int& TN) {
    auto pick = [](int extent, int cap) {
        int best = 1;
        long long best_cost = -1;
        for (int t = 1; t <= cap; t *= 2) {
            long long cost = static_cast<long long>((extent + t - 1) / t) * t;
            if (best_cost < 0 || cost <= best_cost) { best = t; best_cost = cost; }
        }
        return best;
    };
    TW = pick(W, 128);
    TH = pick(H, 128 / TW);
    TN = 128 / (TW * TH);
    (void)N;
}

template <int BLOCK_N>
int launch_igemm(const IgemmMaps& maps, IgemmParams& p, cudaStream_t stream) {
    using Cfg = IgemmCfg<BLOCK_N>;
    static bool attr_set = false;
    if (!attr_set) {
        DFW_CHECK_CUDA(cudaFuncSetAttribute(igemm_kernel<BLOCK_N>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            Cfg::SMEM_BYTES));
        attr_set = true;
    }
    p.n_tiles = (p.Cout + BLOCK_N - 1) / BLOCK_N;
    p.total_tiles = p.tiles_w * p.tiles_h * p.tiles_nimg * p.n_tiles;
    const int grid = p.total_tiles < sm_count() ? p.total_tiles : sm_count();
    igemm_kernel<BLOCK_N><<<grid, IGEMM_THREADS, Cfg::SMEM_BYTES, stream>>>(maps, p);
    g_launches.fetch_add(1);
    DFW_CHECK_CUDA(cudaGetLastError());
    return DFW_OK;
}

int igemm_dispatch(const void* x, const void* w, const float* bias, int bias_sample_stride, const void* residual,
                   void* y, int N, int Hin, int Win, int Cin, int Cout, int ksize, int stride, int pad_mode,
                   int flags, float out_scale, cudaStream_t stream) {
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

    IgemmMaps maps;
    IgemmParams p{};
    const int Hout = Hin / stride, Wout = Win / stride;
    p.N = N; p.H = Hout; p.W = Wout; p.Cout = Cout;
    p.kb_per_tap = Cin / BLOCK_K;
    p.ntaps = ksize * ksize;
    choose_tile(Wout, Hout, N, p.TW, p.TH, p.TN);
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
        const int pad = (ksize - 1) / 2;
        for (int kh = 0; kh < ksize; ++kh)
            for (int kw = 0; kw < ksize; ++kw) {
                const int t = kh * ksize + kw;
                p.tap_map[t] = 0; p.tap_dh[t] = kh - pad; p.tap_dw[t] = kw - pad;
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

    int block_n;
    if (geglu) block_n = 256;
    else if (Cout <= 16) block_n = 16;
    else if (Cout % 256 == 0) block_n = 256;
    else if (Cout % 160 == 0) block_n = 160;
    else block_n = 128;
    {
        const uint64_t Kt = static_cast<uint64_t>(p.ntaps) * Cin;
        const uint64_t dims[2] = {Kt, static_cast<uint64_t>(Cout)};
        const uint64_t strides[1] = {Kt * esz};
        const uint32_t bbox[2] = {BLOCK_K, static_cast<uint32_t>(block_n)};
        rc = encode_tmap_bf16_sw128(&maps.b, w, 2, dims, strides, bbox);
        if (rc != DFW_OK) return rc;
    }
    switch (block_n) {
        case 16: return launch_igemm<16>(maps, p, stream);
        case 128: return launch_igemm<128>(maps, p, stream);
        case 160: return launch_igemm<160>(maps, p, stream);
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

int dfw_linear(const void* x, const void* w, const float* bias, const void* residual, void* y, int M, int K,
               int Nout, int flags, float out_scale, void* stream) {
    // a Linear is a 1x1 "convolution" over an image of height 1 and width M
    return dfw::igemm_dispatch(x, w, bias, 0, residual, y, 1, 1, M, K, Nout, 1, 1, 0, flags, out_scale,
                               static_cast<cudaStream_t>(stream));
}

}  // extern "C"
