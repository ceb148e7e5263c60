// Gradient kernels of the training step for sm_100a.
//
//   wgrad_kernel     dW[co, tap, ci] = sum over pixels of dy[p, co] * x[p + tap, ci]   (tcgen05, both operands MN-major)
//   wgrad_reduce     fixed-order sum of the pixel-split partials (+ scale, optional accumulate) -> fp32 gradient
//   wperm_kernel     w[R, T, C] -> w'[C, T', R] 16-bit (tap map with zeros): the operand of the data-gradient convolution
//   colsum_*         per-channel sums of dy over row groups (bias gradients, per-image time-embedding gradients)
//   downsum2x        backward of the nearest-2x upsample;  split_channels: backward of the channel concat
//   geglu_fwd        y = v * gelu(g) on the (value | gate) pre-activations (the training forward keeps them for geglu_bwd)
//   nchw_to_nhwc_pad fp32 NCHW -> 16-bit NHWC with the channel axis zero-padded (d loss / d prediction -> conv_out's dgrad)
//
// ref: train_tools/train_icl_multitask_nocrop_nearest_nshot_v3.py:1386 `accelerator.backward(loss)` — torch autograd
//      through every nn.Conv2d / nn.Linear of the UNet (cuDNN wgrad / dgrad, cuBLAS) is what these replace.
//
// Weight gradient as a GEMM: per filter tap, D[co, ci] += A[co, p] * B[ci, p] with the pixel index p as the contraction.
// Both tensors are stored pixel-major ([N, H, W, C]: channels contiguous), i.e. the contraction index is the SLOW one:
// that is the "MN-major" operand form of tcgen05, so dy and x are TMA-loaded IN PLACE as [64 pixels x 64 channels]
// SWIZZLE_128B blocks (rows = pixels) and never transposed.  The shifted x tile of a tap is the same TMA box at
// (w + dw, h + dh): zero padding = TMA out-of-bounds fill, stride 2 = the four phase views of x (as in igemm.cu).
// One CTA per SM, persistent over units (pixel split, tap, ci tile, co tile); accumulators double-buffered in TMEM.
#include <atomic>

#include "common.cuh"
#include "ptx.cuh"

namespace dfw {
extern std::atomic<long long> g_launches;
namespace {

constexpr int WG_THREADS = 256;          // warps: 0 TMA, 1 MMA, 2 TMEM alloc, 3 idle, 4-7 epilogue
constexpr int WG_BLK = 64 * 64 * 2;      // one [64 pixels x 64 channels] block
constexpr int WG_OPERAND_BYTES = 4 * (2 + 4) * WG_BLK;      // 4 stages of the widest tile (BN = 256)
constexpr int WG_SMEM = WG_OPERAND_BYTES + 1024 /*barriers*/ + 1024 /*align slack*/;
constexpr int WG_MAX_STAGES = 8;

struct WgradMaps {
    CUtensorMap dy;
    CUtensorMap x[4];
};
struct WgradParams {
    int Cout, Cout_store, Cin;
    int ntaps;
    int tap_map[9], tap_dh[9], tap_dw[9];
    int TW, TH, TN;                    // pixel chunk = TW x TH x TN = 64 pixels of dy's grid
    int chunks_w, chunks_h, total_chunks;
    int tiles_m, tiles_n, NB;          // ci tile = 64 * NB channels
    int splits, chunks_per_split;
    int stages, stage_bytes;
    float* partial;                    // [splits][Cout_store][ld]
    long long split_stride;            // floats
    int ld;                            // ntaps * Cin
    int f16;
    int total_units;
};

__host__ __device__ constexpr uint32_t umma_idesc_mnmn(uint32_t M, uint32_t N, uint32_t fmt) {
    return (1u << 4) | (fmt << 7) | (fmt << 10) | (1u << 15) | (1u << 16) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

__global__ void __launch_bounds__(WG_THREADS, 1)
wgrad_kernel(const __grid_constant__ WgradMaps maps, const __grid_constant__ WgradParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    const uint32_t bar = base + WG_OPERAND_BYTES;
    auto full = [&](int s) { return bar + 8u * s; };
    auto empty = [&](int s) { return bar + 8u * (WG_MAX_STAGES + s); };
    auto tfull = [&](int a) { return bar + 8u * (2 * WG_MAX_STAGES + a); };
    auto tempty = [&](int a) { return bar + 8u * (2 * WG_MAX_STAGES + 2 + a); };
    const uint32_t tmem_slot = bar + 8u * (2 * WG_MAX_STAGES + 4);
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw));
    auto sA = [&](int s) { return base + s * p.stage_bytes; };
    auto sB = [&](int s) { return base + s * p.stage_bytes + 2 * WG_BLK; };

    const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&maps.dy);
        for (int i = 0; i < 4; ++i) tma_prefetch_desc(&maps.x[i]);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < WG_MAX_STAGES; ++s) { mbar_init(full(s), 1); mbar_init(empty(s), 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(tfull(a), 1); mbar_init(tempty(a), 4); }
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

    // unit -> (pixel split, tap, ci tile, co tile); co tiles fastest so neighbouring CTAs share the x tile in L2
    auto decode = [&](int u, int& s, int& t, int& tn, int& tm) {
        tm = u % p.tiles_m; u /= p.tiles_m;
        tn = u % p.tiles_n; u /= p.tiles_n;
        t = u % p.ntaps;
        s = u / p.ntaps;
    };
    auto chunk_range = [&](int s, int& c0, int& c1) {
        c0 = s * p.chunks_per_split;
        c1 = min(c0 + p.chunks_per_split, p.total_chunks);
    };

    if (warp == 0) {
        int stage = 0;
        uint32_t phase = 0;
        const uint32_t tx = static_cast<uint32_t>((2 + p.NB) * WG_BLK);
        for (int u = blockIdx.x; u < p.total_units; u += gridDim.x) {
            int s, t, tn, tm, c0, c1;
            decode(u, s, t, tn, tm);
            chunk_range(s, c0, c1);
            const CUtensorMap* xm = &maps.x[p.tap_map[t]];
            const int dh = p.tap_dh[t], dw = p.tap_dw[t];
            for (int c = c0; c < c1; ++c) {
                const int cw = c % p.chunks_w;
                const int r = c / p.chunks_w;
                const int ch = r % p.chunks_h, cn = r / p.chunks_h;
                const int w0 = cw * p.TW, h0 = ch * p.TH, n0 = cn * p.TN;
                mbar_wait(empty(stage), phase ^ 1u, 40);
                if (elect_one()) {
                    mbar_arrive_expect_tx(full(stage), tx);
                    for (int mb = 0; mb < 2; ++mb)
                        tma_load_4d(sA(stage) + mb * WG_BLK, &maps.dy, full(stage), tm * 128 + mb * 64, w0, h0, n0);
                    for (int nb = 0; nb < p.NB; ++nb)
                        tma_load_4d(sB(stage) + nb * WG_BLK, xm, full(stage), (tn * p.NB + nb) * 64, w0 + dw, h0 + dh, n0);
                }
                __syncwarp();
                if (++stage == p.stages) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (warp == 1) {
        const uint32_t idesc = umma_idesc_mnmn(128, static_cast<uint32_t>(64 * p.NB), p.f16 ? 0u : 1u);
        int stage = 0, acc = 0;
        uint32_t phase = 0, acc_phase = 0;
        for (int u = blockIdx.x; u < p.total_units; u += gridDim.x) {
            int s, t, tn, tm, c0, c1;
            decode(u, s, t, tn, tm);
            chunk_range(s, c0, c1);
            mbar_wait(tempty(acc), acc_phase ^ 1u, 41);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * 256;
            for (int c = c0; c < c1; ++c) {
                mbar_wait(full(stage), phase, 42);
                tc_fence_after();
                if (elect_one()) {
                    // MN-major SWIZZLE_128B: 64-channel column blocks WG_BLK bytes apart (LBO), 8-pixel groups 1024 B apart
                    const uint64_t adesc = umma_desc_sw128(sA(stage), WG_BLK);
                    const uint64_t bdesc = umma_desc_sw128(sB(stage), WG_BLK);
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks)          // 16 pixels per MMA = 16 rows x 128 B
                        umma_ss(d_tmem, adesc + static_cast<uint64_t>(ks * ((16 * 128) >> 4)),
                                bdesc + static_cast<uint64_t>(ks * ((16 * 128) >> 4)), idesc, (c > c0 || ks > 0) ? 1u : 0u);
                    tc_commit(empty(stage));
                    if (c == c1 - 1) tc_commit(tfull(acc));
                }
                __syncwarp();
                if (++stage == p.stages) { stage = 0; phase ^= 1u; }
            }
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1u;
        }
    } else if (warp >= 4) {
        const int q = warp - 4;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int u = blockIdx.x; u < p.total_units; u += gridDim.x) {
            int s, t, tn, tm, c0, c1;
            decode(u, s, t, tn, tm);
            mbar_wait(tfull(acc), acc_phase, 43);
            tc_fence_after();
            const int co = tm * 128 + q * 32 + lane;
            const uint32_t taddr = tmem_base + acc * 256 + (static_cast<uint32_t>(q * 32) << 16);
            float* row = p.partial + static_cast<long long>(s) * p.split_stride + static_cast<long long>(co) * p.ld +
                         static_cast<long long>(t) * p.Cin;
            for (int cb = 0; cb < 2 * p.NB; ++cb) {
                uint32_t v[32];
                tmem_ld_32x32(taddr + cb * 32, v);
                tmem_ld_wait(); tmem_regs_ready(v);
                const int ci0 = tn * p.NB * 64 + cb * 32;
                if (co < p.Cout_store && ci0 < p.Cin) {        // Cin % 32 == 0
                    float4* dst = reinterpret_cast<float4*>(row + ci0);
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        dst[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                             __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty(acc));
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1u;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float4* __restrict__ partial, long long split_stride4, int splits,
                                                           float4* __restrict__ out, long long n4, float scale, int accumulate) {
    for (long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x; i < n4; i += static_cast<long long>(gridDim.x) * 256) {
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int s = 0; s < splits; ++s) {              // fixed order: deterministic
            const float4 v = __ldcs(partial + s * split_stride4 + i);
            a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
        }
        a.x *= scale; a.y *= scale; a.z *= scale; a.w *= scale;
        if (accumulate) { const float4 o = out[i]; a.x += o.x; a.y += o.y; a.z += o.z; a.w += o.w; }
        out[i] = a;
    }
}

struct WgradPlan {
    int TW, TH, TN, chunks_w, chunks_h, chunks_n, total_chunks, tiles_m, tiles_n, NB, splits, cps, ntaps;
};

// 64-pixel chunk (TW x TH x TN, powers of two) covering the [N, H, W] grid with the fewest padded pixels; ties -> wider
void choose_chunk(int W, int H, int N, int& TW, int& TH, int& TN) {
    long long best = -1;
    TW = 64; TH = 1; TN = 1;
    for (int tw = 64; tw >= 1; tw /= 2)
        for (int th = 64 / tw; th >= 1; th /= 2) {
            const int tn = 64 / (tw * th);
            const long long cost = (static_cast<long long>((W + tw - 1) / tw) * tw) * (static_cast<long long>((H + th - 1) / th) * th) *
                                   (static_cast<long long>((N + tn - 1) / tn) * tn);
            if (best < 0 || cost < best) { best = cost; TW = tw; TH = th; TN = tn; }
        }
}

WgradPlan wgrad_plan(int N, int Ho, int Wo, int Cin, int Cout, int ksize) {
    WgradPlan pl;
    pl.ntaps = ksize * ksize;
    choose_chunk(Wo, Ho, N, pl.TW, pl.TH, pl.TN);
    pl.chunks_w = (Wo + pl.TW - 1) / pl.TW;
    pl.chunks_h = (Ho + pl.TH - 1) / pl.TH;
    pl.chunks_n = (N + pl.TN - 1) / pl.TN;
    pl.total_chunks = pl.chunks_w * pl.chunks_h * pl.chunks_n;
    pl.tiles_m = (Cout + 127) / 128;
    // ci tile width 64 * NB: least padded work, narrower tiles penalised for their lower MMA rate
    double best = 0.0;
    pl.NB = 4;
    for (int nb = 4; nb >= 1; --nb) {
        const int bn = 64 * nb;
        const double cost = static_cast<double>((Cin + bn - 1) / bn) * bn * (nb == 1 ? 1.5 : (nb == 2 ? 1.1 : 1.0));
        if (nb == 4 || cost < best - 1e-9) { best = cost; pl.NB = nb; }
    }
    pl.tiles_n = (Cin + 64 * pl.NB - 1) / (64 * pl.NB);
    const long long tiles = static_cast<long long>(pl.tiles_m) * pl.tiles_n * pl.ntaps;
    long long want = (2LL * sm_count() + tiles - 1) / tiles;            // >= 2 units per SM
    const long long max_splits = (pl.total_chunks + 7) / 8;             // >= 8 chunks (512 pixels) per unit
    if (want > max_splits) want = max_splits;
    if (want < 1) want = 1;
    if (want > 64) want = 64;
    pl.cps = static_cast<int>((pl.total_chunks + want - 1) / want);
    pl.splits = (pl.total_chunks + pl.cps - 1) / pl.cps;
    return pl;
}

}  // namespace
}  // namespace dfw

extern "C" {

long long dfw_conv_wgrad_workspace_bytes(int N, int Hin, int Win, int Cin, int Cout, int ksize, int stride) {
    if (N <= 0 || Hin <= 0 || Win <= 0 || Cin <= 0 || Cout <= 0 || (ksize != 1 && ksize != 3) || (stride != 1 && stride != 2))
        return -1;
    const dfw::WgradPlan pl = dfw::wgrad_plan(N, Hin / stride, Win / stride, Cin, Cout, ksize);
    return static_cast<long long>(pl.splits) * Cout * ksize * ksize * Cin * 4;
}

int dfw_conv_wgrad(const void* x, const void* dy, float* dw, int N, int Hin, int Win, int Cin, int Cout, int cout_store,
                   int ksize, int stride, int f16, float scale, int accumulate, void* workspace, void* stream_) {
    using namespace dfw;
    int rc = require_sm100();
    if (rc != DFW_OK) return rc;
    DFW_REQUIRE(x && dy && dw && workspace);
    DFW_REQUIRE(N > 0 && Hin > 0 && Win > 0 && Cin > 0 && Cin % 32 == 0 && Cout > 0 && Cout % 8 == 0);
    DFW_REQUIRE(cout_store > 0 && cout_store <= Cout);
    DFW_REQUIRE(ksize == 1 || ksize == 3);
    DFW_REQUIRE(stride == 1 || (stride == 2 && ksize == 3 && Hin % 2 == 0 && Win % 2 == 0));
    DFW_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(dw) |
                  reinterpret_cast<uintptr_t>(workspace)) & 15) == 0);
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const int Ho = Hin / stride, Wo = Win / stride;
    const WgradPlan pl = wgrad_plan(N, Ho, Wo, Cin, Cout, ksize);
    WgradMaps maps;
    WgradParams p{};
    p.Cout = Cout; p.Cout_store = cout_store; p.Cin = Cin;
    p.ntaps = pl.ntaps;
    p.TW = pl.TW; p.TH = pl.TH; p.TN = pl.TN;
    p.chunks_w = pl.chunks_w; p.chunks_h = pl.chunks_h; p.total_chunks = pl.total_chunks;
    p.tiles_m = pl.tiles_m; p.tiles_n = pl.tiles_n; p.NB = pl.NB;
    p.splits = pl.splits; p.chunks_per_split = pl.cps;
    p.stage_bytes = (2 + pl.NB) * WG_BLK;
    p.stages = WG_OPERAND_BYTES / p.stage_bytes;
    if (p.stages > WG_MAX_STAGES) p.stages = WG_MAX_STAGES;
    p.ld = pl.ntaps * Cin;
    p.partial = static_cast<float*>(workspace);
    p.split_stride = static_cast<long long>(cout_store) * p.ld;
    p.f16 = f16 ? 1 : 0;
    const long long units = static_cast<long long>(pl.splits) * pl.ntaps * pl.tiles_n * pl.tiles_m;
    DFW_REQUIRE(units < (1LL << 31));
    p.total_units = static_cast<int>(units);
    const uint64_t esz = 2;
    const uint32_t box[4] = {64, static_cast<uint32_t>(pl.TW), static_cast<uint32_t>(pl.TH), static_cast<uint32_t>(pl.TN)};
    {
        const uint64_t dims[4] = {static_cast<uint64_t>(Cout), static_cast<uint64_t>(Wo), static_cast<uint64_t>(Ho),
                                  static_cast<uint64_t>(N)};
        const uint64_t strides[3] = {Cout * esz, static_cast<uint64_t>(Wo) * Cout * esz,
                                     static_cast<uint64_t>(Ho) * Wo * Cout * esz};
        rc = encode_tmap_bf16_sw128(&maps.dy, dy, 4, dims, strides, box);
        if (rc != DFW_OK) return rc;
    }
    if (stride == 1) {
        const uint64_t dims[4] = {static_cast<uint64_t>(Cin), static_cast<uint64_t>(Win), static_cast<uint64_t>(Hin),
                                  static_cast<uint64_t>(N)};
        const uint64_t strides[3] = {Cin * esz, static_cast<uint64_t>(Win) * Cin * esz,
                                     static_cast<uint64_t>(Hin) * Win * Cin * esz};
        rc = encode_tmap_bf16_sw128(&maps.x[0], x, 4, dims, strides, box);
        if (rc != DFW_OK) return rc;
        for (int i = 1; i < 4; ++i) maps.x[i] = maps.x[0];
        const int pad = (ksize - 1) / 2;
        for (int kh = 0; kh < ksize; ++kh)
            for (int kw = 0; kw < ksize; ++kw) {
                const int t = kh * ksize + kw;
                p.tap_map[t] = 0; p.tap_dh[t] = kh - pad; p.tap_dw[t] = kw - pad;
            }
    } else {
        // stride 2 / pad 1: in = 2*o + k - 1 -> phase view x[:, ph::2, pw::2, :] at offset d (igemm.cu, pad_mode 0)
        const uint64_t dims[4] = {static_cast<uint64_t>(Cin), static_cast<uint64_t>(Win / 2), static_cast<uint64_t>(Hin / 2),
                                  static_cast<uint64_t>(N)};
        const uint64_t strides[3] = {2 * Cin * esz, 2 * static_cast<uint64_t>(Win) * Cin * esz,
                                     static_cast<uint64_t>(Hin) * Win * Cin * esz};
        for (int ph = 0; ph < 2; ++ph)
            for (int pw = 0; pw < 2; ++pw) {
                const uint8_t* b = reinterpret_cast<const uint8_t*>(x) + (static_cast<uint64_t>(ph) * Win + pw) * Cin * esz;
                rc = encode_tmap_bf16_sw128(&maps.x[ph * 2 + pw], b, 4, dims, strides, box);
                if (rc != DFW_OK) return rc;
            }
        for (int kh = 0; kh < 3; ++kh)
            for (int kw = 0; kw < 3; ++kw) {
                const int t = kh * 3 + kw;
                const int ph = (kh + 1) & 1, pw = (kw + 1) & 1;
                p.tap_map[t] = ph * 2 + pw; p.tap_dh[t] = (kh == 0) ? -1 : 0; p.tap_dw[t] = (kw == 0) ? -1 : 0;
            }
    }
    static bool attr_set = false;
    if (!attr_set) {
        DFW_CHECK_CUDA(cudaFuncSetAttribute(wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM));
        attr_set = true;
    }
    const int grid = p.total_units < sm_count() ? p.total_units : sm_count();
    wgrad_kernel<<<grid, WG_THREADS, WG_SMEM, stream>>>(maps, p);
    g_launches.fetch_add(1);
    DFW_CHECK_CUDA(cudaGetLastError());
    const long long n4 = static_cast<long long>(cout_store) * p.ld / 4;
    long long blocks = (n4 + 255) / 256;
    if (blocks > 4LL * sm_count()) blocks = 4LL * sm_count();
    wgrad_reduce_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(reinterpret_cast<const float4*>(workspace),
                                                                     p.split_stride / 4, p.splits, reinterpret_cast<float4*>(dw),
                                                                     n4, scale, accumulate);
    g_launches.fetch_add(1);
    DFW_CHECK_CUDA(cudaGetLastError());
    return DFW_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------------------------
// small layout / reduction kernels
// ---------------------------------------------------------------------------------------------------------------------
namespace dfw {
namespace {

// out[c, t', r] = tapmap[t'] >= 0 ? in[r, tapmap[t'], c] : 0      in [R, T, C], out [C, T_out, R], 16-bit
// 32 x 32 tiles of the (r, c) plane through smem (coalesced on both sides); grid = (C/32, R/32, T_out)
struct TapMap { int t[16]; };
__global__ void __launch_bounds__(256) wperm_kernel(const uint16_t* __restrict__ in, uint16_t* __restrict__ out, int R, int T,
                                                    int C, int T_out, TapMap tm) {
    __shared__ uint16_t tile[32][33];
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32, to = blockIdx.z;
    const int ti = tm.t[to];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;       // 32 x 8
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int r = r0 + ty + 8 * k, c = c0 + tx;
        uint16_t v = 0;
        if (ti >= 0 && r < R && c < C) v = in[(static_cast<long long>(r) * T + ti) * C + c];
        tile[ty + 8 * k][tx] = v;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int c = c0 + ty + 8 * k, r = r0 + tx;
        if (c < C && r < R) out[(static_cast<long long>(c) * T_out + to) * R + r] = tile[tx][ty + 8 * k];
    }
}

template <int XD>   // 0 bf16, 1 fp32, 2 fp16
__device__ __forceinline__ float ldf(const void* x, long long i) {
    if (XD == 1) return reinterpret_cast<const float*>(x)[i];
    if (XD == 2) return __half2float(reinterpret_cast<const __half*>(x)[i]);
    return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(x)[i]);
}

// stage 1: partial[g][chunk][c] = sum of x[g * rows_per_group + r, c] over the chunk's rows.  A thread owns 8 consecutive
// channels (one 16-byte load per row for 16-bit tensors); the 256 threads of a CTA cover VPR = C / 8 vectors x 256 / VPR
// rows at a time, four rows in flight per thread; the row lanes are folded through smem in a fixed order.
template <int XD>
__global__ void __launch_bounds__(256) colsum_partial_kernel(const void* __restrict__ x, float* __restrict__ partial,
                                                             long long rows_per_group, int C, int chunks) {
    __shared__ float red[256 * 8];
    const int g = blockIdx.z, chunk = blockIdx.y;
    const int VPR = C / 8;                                   // vectors per row (host: C % 8 == 0, VPR <= 256 per x-block)
    const int vpb = VPR < 256 ? VPR : 256;                   // vectors handled by this block
    const int RP = 256 / vpb;                                // rows in parallel
    const int tx = threadIdx.x % vpb, ty = threadIdx.x / vpb;
    const int v = blockIdx.x * 256 + tx;
    const bool active = ty < RP && v < VPR;
    const long long per = (rows_per_group + chunks - 1) / chunks;
    const long long r0 = chunk * per, r1 = min(r0 + per, rows_per_group);
    const long long base = static_cast<long long>(g) * rows_per_group;
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    auto ld8 = [&](long long r, float (&f)[8]) {
        const long long o = (base + r) * C + static_cast<long long>(v) * 8;
        if (XD == 1) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(x) + o));
            const float4 b = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(x) + o) + 1);
            f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
        } else {
            const uint4 u = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(x) + o));
            const float2 a = unpack_h2(u.x, XD == 2), b = unpack_h2(u.y, XD == 2), c = unpack_h2(u.z, XD == 2), d = unpack_h2(u.w, XD == 2);
            f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
        }
    };
    if (active) {
        long long r = r0 + ty;
        for (; r + 3LL * RP < r1; r += 4LL * RP) {
            float f0[8], f1[8], f2[8], f3[8];
            ld8(r, f0); ld8(r + RP, f1); ld8(r + 2LL * RP, f2); ld8(r + 3LL * RP, f3);
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] += (f0[i] + f1[i]) + (f2[i] + f3[i]);
        }
        for (; r < r1; r += RP) {
            float f0[8];
            ld8(r, f0);
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] += f0[i];
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) red[threadIdx.x * 8 + i] = active ? acc[i] : 0.f;
    __syncthreads();
    if (ty == 0 && v < VPR) {
        float* dst = partial + (static_cast<long long>(g) * chunks + chunk) * C + static_cast<long long>(v) * 8;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float t = 0.f;
            for (int k = 0; k < RP; ++k) t += red[(k * vpb + tx) * 8 + i];
            dst[i] = t;
        }
    }
}
__global__ void __launch_bounds__(256) colsum_final_kernel(const float* __restrict__ partial, float* __restrict__ out, int C,
                                                           int chunks, int groups, float scale, int accumulate) {
    const long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
    if (i >= static_cast<long long>(groups) * C) return;
    const int g = static_cast<int>(i / C), c = static_cast<int>(i % C);
    float s = 0.f;
    for (int k = 0; k < chunks; ++k) s += partial[(static_cast<long long>(g) * chunks + k) * C + c];
    s *= scale;
    out[i] = accumulate ? out[i] + s : s;
}

// dx[n, h, w, :] = dy[n, 2h, 2w, :] + dy[n, 2h, 2w+1, :] + dy[n, 2h+1, 2w, :] + dy[n, 2h+1, 2w+1, :]   (16-bit, 8 ch / thread)
__global__ void __launch_bounds__(256) downsum2x_kernel(const uint4* __restrict__ dy, uint4* __restrict__ dx, long long total,
                                                        int H, int W, int C8, int f16) {
    for (long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * 256) {
        const int c = static_cast<int>(i % C8);
        long long r = i / C8;
        const int w = static_cast<int>(r % W); r /= W;
        const int h = static_cast<int>(r % H);
        const long long n = r / H;
        const long long W2 = 2LL * W;
        const long long o = ((n * 2 * H + 2 * h) * W2 + 2 * w) * C8 + c;
        const uint4 a = dy[o], b = dy[o + C8], cc = dy[o + W2 * C8], d = dy[o + W2 * C8 + C8];
        const uint32_t* pa = reinterpret_cast<const uint32_t*>(&a); const uint32_t* pb = reinterpret_cast<const uint32_t*>(&b);
        const uint32_t* pc = reinterpret_cast<const uint32_t*>(&cc); const uint32_t* pd = reinterpret_cast<const uint32_t*>(&d);
        uint4 out;
        uint32_t* po = reinterpret_cast<uint32_t*>(&out);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 fa = unpack_h2(pa[j], f16), fb = unpack_h2(pb[j], f16), fc = unpack_h2(pc[j], f16), fd = unpack_h2(pd[j], f16);
            po[j] = pack_h2((fa.x + fb.x) + (fc.x + fd.x), (fa.y + fb.y) + (fc.y + fd.y), f16);
        }
        dx[i] = out;
    }
}

// a[rows, Ca] = y[rows, :Ca], b[rows, Cb] = y[rows, Ca:]   (16-byte vectors; Ca, Cb multiples of 8 halves)
__global__ void __launch_bounds__(256) split_channels_kernel(const uint4* __restrict__ y, uint4* __restrict__ a, uint4* __restrict__ b,
                                                             long long rows, int Ca8, int Cb8) {
    const int C8 = Ca8 + Cb8;
    const long long total = rows * C8;
    for (long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * 256) {
        const long long r = i / C8;
        const int c = static_cast<int>(i - r * C8);
        const uint4 v = y[i];
        if (c < Ca8) a[r * Ca8 + c] = v; else b[r * Cb8 + (c - Ca8)] = v;
    }
}

__device__ __forceinline__ float gelu_exact(float g) { return 0.5f * g * (1.0f + erff(g * 0.70710678118654752f)); }
// y[row, c] = h[row, c] * gelu(h[row, F + c])   16-bit, 2 elements per thread
__global__ void __launch_bounds__(256) geglu_fwd_kernel(const uint32_t* __restrict__ h, uint32_t* __restrict__ y, long long M,
                                                        int F2, int f16) {
    const long long total = M * F2;
    for (long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * 256) {
        const long long row = i / F2;
        const int c = static_cast<int>(i - row * F2);
        const float2 v = unpack_h2(h[row * 2 * F2 + c], f16), g = unpack_h2(h[row * 2 * F2 + F2 + c], f16);
        y[i] = pack_h2(v.x * gelu_exact(g.x), v.y * gelu_exact(g.y), f16);
    }
}

// y[n, h, w, c] = c < C ? x[n, c, h, w] * scale : 0       fp32 NCHW -> 16-bit NHWC with Cpad channels
__global__ void __launch_bounds__(256) nchw_to_nhwc_pad_kernel(const float* __restrict__ x, uint16_t* __restrict__ y, long long total,
                                                               int C, int HW, int Cpad, float scale, int f16) {
    for (long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * 256) {
        const int c = static_cast<int>(i % Cpad);
        const long long pix = i / Cpad;
        const long long n = pix / HW, hw = pix % HW;
        const float v = c < C ? x[(n * C + c) * HW + hw] * scale : 0.f;
        y[i] = static_cast<uint16_t>(pack_h2(v, 0.f, f16) & 0xffffu);
    }
}

int grid_1d(long long total, int cap_mult = 8) {
    long long b = (total + 255) / 256;
    const long long cap = static_cast<long long>(cap_mult) * sm_count();
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return static_cast<int>(b);
}

}  // namespace
}  // namespace dfw

extern "C" {

int dfw_weight_permute(const void* w, void* out, int R, int T, int C, int T_out, const int* tap_map, void* stream) {
    using namespace dfw;
    int rc = require_sm100();
    if (rc != DFW_OK) return rc;
    DFW_REQUIRE(w && out && tap_map && R > 0 && T > 0 && C > 0 && T_out > 0 && T_out <= 16);
    TapMap tm;
    for (int i = 0; i < 16; ++i) tm.t[i] = -1;
    for (int i = 0; i < T_out; ++i) { DFW_REQUIRE(tap_map[i] >= -1 && tap_map[i] < T); tm.t[i] = tap_map[i]; }
    const dim3 grid((C + 31) / 32, (R + 31) / 32, T_out);
    DFW_REQUIRE(grid.y < 65536);
    wperm_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const uint16_t*>(w), static_cast<uint16_t*>(out),
                                                                     R, T, C, T_out, tm);
    g_launches.fetch_add(1);
    DFW_CHECK_CUDA(cudaGetLastError());
    return DFW_OK;
}

int dfw_colsum_chunks(long long rows_per_group, int groups) {
    if (rows_per_group <= 0 || groups <= 0) return -1;
    long long want = (4LL * dfw::sm_count() + groups - 1) / groups;
    const long long cap = (rows_per_group + 63) / 64;
    if (want > cap) want = cap;
    if (want < 1) want = 1;
    return static_cast<int>(want);
}

int dfw_colsum(const void* x, int dtype, float* out, long long rows_per_group, int groups, int C, float scale, int accumulate,
               float* workspace /* groups * dfw_colsum_chunks() * C floats */, void* stream_) {
    using namespace dfw;
    int rc = require_sm100();
    if (rc != DFW_OK) return rc;
    DFW_REQUIRE(x && out && workspace && rows_per_group > 0 && groups > 0 && groups < 65536 && C > 0 && C % 8 == 0);
    DFW_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0);
    DFW_REQUIRE(dtype >= 0 && dtype <= 2);
    cudaStream_t st = static_cast<cudaStream_t>(stream_);
    const int chunks = dfw_colsum_chunks(rows_per_group, groups);
    const dim3 grid((C / 8 + 255) / 256, chunks, groups);
    if (dtype == 1) colsum_partial_kernel<1><<<grid, 256, 0, st>>>(x, workspace, rows_per_group, C, chunks);
    else if (dtype == 2) colsum_partial_kernel<2><<<grid, 256, 0, st>>>(x, workspace, rows_per_group, C, chunks);
    else colsum_partial_kernel<0><<<grid, 256, 0, st>>>(x, workspace, rows_per_group, C, chunks);
    g_launches.fetch_add(1);
    DFW_CHECK_CUDA(cudaGetLastError());
    const long long n = static_cast<long long>(groups) * C;
    colsum_final_kernel<<<static_cast<int>((n + 255) / 256), 256, 0, st>>>(workspace, out, C, chunks, groups, scale, accumulate);
    g_launches.fetch_add(1);
    DFW_CHECK_CUDA(cudaGetLastError());
    return DFW_OK;
}

int dfw_downsum2x_nhwc(const void* dy, void* dx, int f16, int N, int H, int W, int C, void* stream) {
    using namespace dfw;
    int rc = require_sm100();
    if (rc != DFW_OK) return rc;
    DFW_REQUIRE(dy && dx && N > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0);
    const long long total = static_cast<long long>(N) * H * W * (C / 8);
    downsum2x_kernel<<<grid_1d(total), 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const uint4*>(dy),
                                                                                    static_cast<uint4*>(dx), total, H, W, C / 8, f16);
    g_launches.fetch_add(1);
    DFW_CHECK_CUDA(cudaGetLastError());
    return DFW_OK;
}

int dfw_split_channels(const void* y, void* a, void* b, long long rows, int Ca, int Cb, void* stream) {
    using namespace dfw;
    int rc = require_sm100();
    if (rc != DFW_OK) return rc;
    DFW_REQUIRE(y && a && b && rows > 0 && Ca > 0 && Cb > 0 && Ca % 8 == 0 && Cb % 8 == 0);
    const long long total = rows * ((Ca + Cb) / 8);
    split_channels_kernel<<<grid_1d(total), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const uint4*>(y), static_cast<uint4*>(a), static_cast<uint4*>(b), rows, Ca / 8, Cb / 8);
    g_launches.fetch_add(1);
    DFW_CHECK_CUDA(cudaGetLastError());
    return DFW_OK;
}

int dfw_geglu_fwd(const void* h, void* y, int f16, long long M, int F, void* stream) {
    using namespace dfw;
    int rc = require_sm100();
    if (rc != DFW_OK) return rc;
    DFW_REQUIRE(h && y && M > 0 && F > 0 && F % 2 == 0);
    const long long total = M * (F / 2);
    geglu_fwd_kernel<<<grid_1d(total), 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const uint32_t*>(h),
                                                                                    static_cast<uint32_t*>(y), M, F / 2, f16);
    g_launches.fetch_add(1);
    DFW_CHECK_CUDA(cudaGetLastError());
    return DFW_OK;
}

int dfw_nchw_f32_to_nhwc16_pad(const float* x, void* y, int N, int C, int H, int W, int Cpad, float scale, int f16, void* stream) {
    using namespace dfw;
    int rc = require_sm100();
    if (rc != DFW_OK) return rc;
    DFW_REQUIRE(x && y && N > 0 && C > 0 && H > 0 && W > 0 && Cpad >= C);
    const long long total = static_cast<long long>(N) * H * W * Cpad;
    nchw_to_nhwc_pad_kernel<<<grid_1d(total), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, static_cast<uint16_t*>(y), total, C,
                                                                                           H * W, Cpad, scale, f16);
    g_launches.fetch_add(1);
    DFW_CHECK_CUDA(cudaGetLastError());
    return DFW_OK;
}

}  // extern "C"
