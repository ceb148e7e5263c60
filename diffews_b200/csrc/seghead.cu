// K10 — fused VAE decoder head: GroupNorm-apply + SiLU + conv3x3 128 -> 3 + clip + (.)*0.5+0.5 + (.)*255 + uint8, one pass.
//
// ref: diffews/marigold_pipeline_rgb_latent_noise.py:887-905 (decode_seg: vae.decoder(...) whose tail is
//        conv_norm_out (GroupNorm 32, eps 1e-6) -> SiLU -> conv_out 128 -> 3, upstream Decoder.forward; .clip(-1, 1)),
//      :787-795 (clip(-1,1); seg * 0.5 + 0.5; * 255) and :534 (clip(0,255).astype(uint8): truncation).
//
// The input [N,H,W,128] (67 MB per 512^2 image in 16 bits) is read ONCE; nothing but the 3 uint8 planes is written.
// With 3 output channels a tcgen05 tile would be > 80 % padding and a direct CUDA-core convolution needs 3456 FMAs per
// pixel (0.43 ms per 16-image batch at the SM's 128 FMA/clk, 2.6x the HBM time).  Instead the 3x3 convolution is split as
//      T[p, (dy,dx,o)] = sum_c a[p, c] * w[o, c, dy, dx]          one GEMM, M = pixels, N = 27 (-> 32), K = 128
//      out[y, x, o]    = sum_{dy,dx} T[(y+dy-1, x+dx-1), (dy,dx,o)]   9 shifted adds per output
// so every activated pixel vector enters the tensor cores exactly once (32 warp-level mma.sync.m16n8k16 per 16 pixels:
// 0.06 ms per batch at the measured 8 cycles per mma per sub-partition, profiles/r02_microbench_hmma_rate.log) and the
// shifted adds run on shared-memory accumulators: a ring of three output rows per CTA, updated in three barrier-separated
// phases (one per dx) in which all writers hit distinct addresses -- no atomics.  The GroupNorm affine + SiLU is applied
// while the raw tile moves global -> registers -> shared (fp32 affine, one tanh.approx per element).  Bound: HBM, then
// the XU (one MUFU per element = 0.13 ms per batch).
//
// CTA = 256 threads = 8 warps, 2 CTAs per SM (their load / transform / mma phases interleave).  A CTA owns a contiguous
// range of output rows; per input row it walks 128-pixel chunks: 8 x LDG.128 per thread prefetched one chunk ahead,
// transform -> smem tile [128 px][128 ch] (272 B pitch: conflict-free ldmatrix), one 16-pixel m-tile per warp.
#include <atomic>

#include "common.cuh"
#include "ptx.cuh"

namespace dfw {
extern std::atomic<long long> g_launches;
namespace {

constexpr int SH_C = 128;                // channels (the decoder's last block)
constexpr int SH_THREADS = 256;
constexpr int SH_CHUNK = 128;            // pixels per chunk = 8 warps x 16
constexpr int SH_PITCH = 272;            // bytes per pixel row of the activated tile (256 + 16: ldmatrix rows hit distinct banks)
constexpr int SH_WB_U32 = 8 * 4 * 32 * 2;  // B fragments: [k-step][n-tile][lane][2]

struct SegHeadParams {
    const uint16_t* x;        // [N, H, W, 128] 16-bit
    const float* ss;          // [N, 2, 128] GroupNorm scale / shift
    const uint32_t* wb;       // SH_WB_U32 prepared B fragments
    float bias[3];
    uint8_t* out_u8;          // [N, 3, H, W] or null
    float* out_f32;           // [N, 3, H, W] float in [0, 255] or null
    int N, H, W;
    long long rows_total;     // N * H
    int f16;
};

__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t (&r)[4]) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
template <bool F16>
__device__ __forceinline__ void mma_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    if constexpr (F16)
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
    else
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float tanh_approx(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// silu(x * scale + shift) for 8 channels of one pixel (one 16-byte piece), back to 16 bits
template <bool F16>
__device__ __forceinline__ uint4 gn_silu8(const uint4& raw, const float2* __restrict__ ss8) {
    uint32_t in[4] = {raw.x, raw.y, raw.z, raw.w}, out[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 v = unpack_h2(in[i], F16 ? 1 : 0);
        const float2 s0 = ss8[2 * i], s1 = ss8[2 * i + 1];
        const float h0 = 0.5f * fmaf(v.x, s0.x, s0.y), h1 = 0.5f * fmaf(v.y, s1.x, s1.y);
        out[i] = pack_h2(fmaf(h0, tanh_approx(h0), h0), fmaf(h1, tanh_approx(h1), h1), F16 ? 1 : 0);   // y * sigmoid(y)
    }
    return make_uint4(out[0], out[1], out[2], out[3]);
}

template <bool F16>
__global__ void __launch_bounds__(SH_THREADS, 2) seg_head_kernel(const __grid_constant__ SegHeadParams p) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int W = p.W, H = p.H;
    const int ring_row = (W + 2) * 3;                       // floats per output-accumulator row (x = -1 .. W)
    uint8_t* tile = smem;                                    // [SH_CHUNK][SH_PITCH]
    uint32_t* wb = reinterpret_cast<uint32_t*>(smem + SH_CHUNK * SH_PITCH);
    float2* ssm = reinterpret_cast<float2*>(wb + SH_WB_U32);  // [128] (scale, shift) of the current image
    float* ring = reinterpret_cast<float*>(ssm + SH_C);      // [3][W + 2][3]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < SH_WB_U32; i += SH_THREADS) wb[i] = __ldg(p.wb + i);

    // contiguous range of output rows (global row index g = n * H + y) of this CTA
    const long long g_lo = p.rows_total * blockIdx.x / gridDim.x, g_hi = p.rows_total * (blockIdx.x + 1) / gridDim.x;
    const int nchunk = (W + SH_CHUNK - 1) / SH_CHUNK;
    // scatter roles of this lane's two accumulator columns c = (lane & 3) * 2 + {0, 1} (see the scatter below)
    const int dyj[2] = {((lane & 3) * 2) / 3, ((lane & 3) * 2 + 1) / 3}, oj[2] = {((lane & 3) * 2) % 3, ((lane & 3) * 2 + 1) % 3};
    const int ninth_phase[2] = {(lane & 3) == 0 ? 0 : (lane & 3) == 1 ? 2 : -1, (lane & 3) == 0 ? 1 : -1};
    const int piece_c = tid & 15;                            // this thread's 8-channel piece
    const int piece_p = tid >> 4;                            // and first pixel (stride 16) within a chunk

    long long g = g_lo;
    while (g < g_hi) {
        const int n = static_cast<int>(g / H), ya = static_cast<int>(g % H);
        const int yb = static_cast<int>(min(static_cast<long long>(H), ya + (g_hi - g)));      // output rows [ya, yb) of image n
        __syncthreads();                                                                     // previous segment fully drained
        for (int i = tid; i < SH_C; i += SH_THREADS)
            ssm[i] = make_float2(__ldg(p.ss + (static_cast<size_t>(n) * 2) * SH_C + i), __ldg(p.ss + (static_cast<size_t>(n) * 2 + 1) * SH_C + i));
        for (int i = tid; i < 3 * ring_row; i += SH_THREADS) ring[i] = 0.f;
        const int r_lo = max(ya - 1, 0), r_hi = min(yb, H - 1);                              // input rows that contribute
        const uint16_t* ximg = p.x + static_cast<size_t>(n) * H * W * SH_C;
        // prefetch of (row, chunk) items, one ahead
        uint4 pre[8];
        auto prefetch = [&](int r, int ch) {
            const uint4* src = reinterpret_cast<const uint4*>(ximg + (static_cast<size_t>(r) * W + ch * SH_CHUNK) * SH_C);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int px = piece_p + 16 * i;
                pre[i] = (ch * SH_CHUNK + px < W) ? __ldg(src + px * 16 + piece_c) : make_uint4(0, 0, 0, 0);
            }
        };
        prefetch(r_lo, 0);
        for (int r = r_lo; r <= r_hi; ++r) {
            for (int ch = 0; ch < nchunk; ++ch) {
                __syncthreads();                            // tile free (previous chunk's ldmatrix done), ring phases done, ssm / ring init visible
                {
                    const float2* ss8 = ssm + piece_c * 8;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int px = piece_p + 16 * i;
                        *reinterpret_cast<uint4*>(tile + px * SH_PITCH + piece_c * 16) = gn_silu8<F16>(pre[i], ss8);
                    }
                }
                __syncthreads();
                {   // next item's loads are in flight during the mma / scatter below
                    int r2 = r, c2 = ch + 1;
                    if (c2 == nchunk) { c2 = 0; ++r2; }
                    if (r2 <= r_hi) prefetch(r2, c2);
                }
                const int x0 = ch * SH_CHUNK + warp * 16;   // first pixel of this warp's m-tile
                float acc[4][4];
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
                if (x0 < W) {
                    const uint32_t a_base = smem_u32(tile) + (warp * 16 + (lane & 7) + 8 * ((lane >> 3) & 1)) * SH_PITCH + 16 * (lane >> 4);
#pragma unroll
                    for (int kc = 0; kc < 8; ++kc) {
                        uint32_t a[4];
                        ldmatrix_x4(a_base + kc * 32, a);
#pragma unroll
                        for (int nt = 0; nt < 4; ++nt) {
                            const uint2 b = *reinterpret_cast<const uint2*>(wb + ((kc * 4 + nt) * 32 + lane) * 2);
                            mma_16816<F16>(acc[nt], a, b.x, b.y);
                        }
                    }
                }
                // scatter: T[x', (dy,dx,o)] -> ring[(r - dy + 1) % 3][x' - dx + 1][o]; phase dx: all writers distinct.
                // Column order (dfw_seg_head_prepare_weights): n-tile nt < 3 holds dx = nt with (dy, o) = (c / 3, c % 3) for
                // c = 0..7; n-tile 3 holds the ninth combination (dy, o) = (2, 2) of dx = 0, 1, 2 in its columns 0, 1, 2.
                // So the phase of every accumulator is static and (dy, o) depend on the lane only (hoisted: dyj / oj).
                float* const row_base[3] = {ring + ((r + 1) % 3) * ring_row, ring + (r % 3) * ring_row, ring + ((r + 2) % 3) * ring_row};   // dy = 0, 1, 2
                const int xr = x0 + (lane >> 2) + 2;         // ring column of fragment row lane/4 for dx = 0 (+1: column 0 is x = -1)
#pragma unroll
                for (int ph = 0; ph < 3; ++ph) {
                    if (x0 < W) {
#pragma unroll
                        for (int jj = 0; jj < 2; ++jj) {
                            float* dst = (dyj[jj] == 0 ? row_base[0] : dyj[jj] == 1 ? row_base[1] : row_base[2]) + (xr - ph) * 3 + oj[jj];
                            dst[0] += acc[ph][jj];
                            dst[24] += acc[ph][2 + jj];       // fragment row + 8
                        }
                        if (ninth_phase[0] == ph) { float* dst = row_base[2] + (xr - ph) * 3 + 2; dst[0] += acc[3][0]; dst[24] += acc[3][2]; }
                        if (ninth_phase[1] == ph) { float* dst = row_base[2] + (xr - ph) * 3 + 2; dst[0] += acc[3][1]; dst[24] += acc[3][3]; }
                    }
                    if (ph < 2) __syncthreads();
                }
            }
            // input row r done: output row y = r - 1 is final (and y = r too when r is the image's last row)
            __syncthreads();
            const int y_done_hi = (r == H - 1) ? r : r - 1;
            for (int y = r - 1; y <= y_done_hi; ++y) {
                if (y >= ya && y < yb) {
                    const float* src = ring + ((y + 3) % 3) * ring_row + 3;      // x = 0
                    for (int i = tid; i < 3 * W; i += SH_THREADS) {
                        const int o = i / W, x = i - o * W;
                        float v = src[x * 3 + o] + p.bias[o];
                        v = fminf(fmaxf(v, -1.0f), 1.0f);
                        v = __fadd_rn(__fmul_rn(v, 0.5f), 0.5f);                 // (seg * 0.5) + 0.5, torch op order
                        v = __fmul_rn(v, 255.0f);
                        const size_t off = ((static_cast<size_t>(n) * 3 + o) * H + y) * W + x;
                        if (p.out_f32) p.out_f32[off] = v;
                        if (p.out_u8) p.out_u8[off] = static_cast<uint8_t>(fminf(fmaxf(v, 0.0f), 255.0f));   // truncation
                    }
                }
            }
            __syncthreads();
            {   // the slot of y = r - 1 (final, or outside the band / image: then it only holds stray sums) becomes y = r + 2's
                float* z = ring + ((r - 1 + 3) % 3) * ring_row;
                for (int i = tid; i < ring_row; i += SH_THREADS) z[i] = 0.f;
            }
        }
        g += (yb - ya);
    }
}

}  // namespace
}  // namespace dfw

extern "C" {

long long dfw_seg_head_weight_u32(void) { return dfw::SH_WB_U32; }

// Host-side re-layout of conv_out's weight [3, 128, 3, 3] (fp32, OIHW) into the mma.sync B fragments the kernel reads:
// wb[((kc * 4 + nt) * 32 + lane) * 2 + reg] = pack(Wt[nt*8 + lane/4][kc*16 + (lane%4)*2 + 8*reg + {0,1}]) with
// Wt[nt*8 + c] = w[o][.][dy][dx]: nt < 3 -> dx = nt, (dy, o) = (c / 3, c % 3); nt = 3 -> (dy, o) = (2, 2), dx = c < 3; else 0.  `out` holds dfw_seg_head_weight_u32() uint32.
int dfw_seg_head_prepare_weights(const float* w_oihw, int f16, uint32_t* out) {
    if (!w_oihw || !out) return DFW_ERR_INVALID;
    auto to16 = [&](float v) -> uint32_t {
        if (f16) { __half h = __float2half_rn(v); return *reinterpret_cast<unsigned short*>(&h); }
        __nv_bfloat16 h = __float2bfloat16_rn(v);
        return *reinterpret_cast<unsigned short*>(&h);
    };
    auto wt = [&](int row, int c) -> float {          // row = nt * 8 + col: the column order the scatter of the kernel assumes
        const int nt = row / 8, col = row % 8;
        int dy, dx, o;
        if (nt < 3) { dx = nt; dy = col / 3; o = col % 3; }
        else if (col < 3) { dx = col; dy = 2; o = 2; }
        else return 0.f;
        return w_oihw[((o * 128 + c) * 3 + dy) * 3 + dx];
    };
    for (int kc = 0; kc < 8; ++kc)
        for (int nt = 0; nt < 4; ++nt)
            for (int lane = 0; lane < 32; ++lane)
                for (int reg = 0; reg < 2; ++reg) {
                    const int n = nt * 8 + lane / 4, k = kc * 16 + (lane % 4) * 2 + 8 * reg;
                    out[((kc * 4 + nt) * 32 + lane) * 2 + reg] = to16(wt(n, k)) | (to16(wt(n, k + 1)) << 16);
                }
    return DFW_OK;
}

int dfw_seg_head_u8(const void* x, int f16, const float* scale_shift, const uint32_t* wb, const float* bias_host,
                    uint8_t* out_u8, float* out_f32, int N, int H, int W, void* stream_) {
    using namespace dfw;
    int rc = require_sm100();
    if (rc != DFW_OK) return rc;
    DFW_REQUIRE(x && scale_shift && wb && bias_host && (out_u8 || out_f32) && N > 0 && H > 0 && W > 0);
    DFW_REQUIRE(W % 16 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0);
    SegHeadParams p{};
    p.x = reinterpret_cast<const uint16_t*>(x); p.ss = scale_shift; p.wb = wb;
    p.bias[0] = bias_host[0]; p.bias[1] = bias_host[1]; p.bias[2] = bias_host[2];
    p.out_u8 = out_u8; p.out_f32 = out_f32; p.N = N; p.H = H; p.W = W; p.f16 = f16;
    p.rows_total = static_cast<long long>(N) * H;
    const size_t smem = static_cast<size_t>(SH_CHUNK) * SH_PITCH + SH_WB_U32 * 4 + SH_C * 8 + 3ull * (W + 2) * 3 * 4;
    DFW_REQUIRE(smem <= 110 * 1024);
    static bool attr_set = false;
    if (!attr_set) {
        DFW_CHECK_CUDA(cudaFuncSetAttribute(seg_head_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024));
        DFW_CHECK_CUDA(cudaFuncSetAttribute(seg_head_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024));
        attr_set = true;
    }
    long long ctas = 2LL * sm_count();
    if (ctas > p.rows_total) ctas = p.rows_total;           // at least one output row per CTA
    cudaStream_t st = static_cast<cudaStream_t>(stream_);
    if (f16) seg_head_kernel<true><<<static_cast<unsigned>(ctas), SH_THREADS, smem, st>>>(p);
    else seg_head_kernel<false><<<static_cast<unsigned>(ctas), SH_THREADS, smem, st>>>(p);
    g_launches.fetch_add(1);
    DFW_CHECK_CUDA(cudaGetLastError());
    return DFW_OK;
}

}  // extern "C"
