// Layout / small-channel helpers (CUDA cores; bandwidth-bound, 128-bit vectorised where the layout allows).
#include <atomic>

#include "common.cuh"
#include "ptx.cuh"

namespace dfw {
extern std::atomic<long long> g_launches;
namespace {

// ---------------------------------------------------------------------------------------------------------
// nearest 2x upsample, NHWC bf16: one thread per input 16-byte vector, four stores.
// ref: diffusers Upsample2D: F.interpolate(scale_factor=2.0, mode="nearest") (upstream)
// ---------------------------------------------------------------------------------------------------------
__global__ void upsample2x_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, long long total_vecs, int H,
                                  int W, int V) {
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total_vecs; i += stride) {
        const int vc = static_cast<int>(i % V);
        long long pix = i / V;
        const int w = static_cast<int>(pix % W);
        pix /= W;
        const int h = static_cast<int>(pix % H);
        const long long n = pix / H;
        const uint4 v = __ldg(x + i);
        const long long W2 = 2LL * W;
        const long long o = ((n * 2 * H + 2 * h) * W2 + 2 * w) * V + vc;
        y[o] = v;
        y[o + V] = v;
        y[o + W2 * V] = v;
        y[o + W2 * V + V] = v;
    }
}

// same, fp32 input -> bf16 output (fuses the fp32 residual-stream -> bf16 MMA-operand cast)
__global__ void upsample2x_f32_kernel(const float4* __restrict__ x, uint4* __restrict__ y, long long total_vecs, int H,
                                      int W, int V, int y_f16) {
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total_vecs; i += stride) {
        const int vc = static_cast<int>(i % V);
        long long pix = i / V;
        const int w = static_cast<int>(pix % W);
        pix /= W;
        const int h = static_cast<int>(pix % H);
        const long long n = pix / H;
        const float4 a = __ldg(x + 2 * i), b = __ldg(x + 2 * i + 1);
        uint4 v;
        v.x = pack_h2(a.x, a.y, y_f16); v.y = pack_h2(a.z, a.w, y_f16);
        v.z = pack_h2(b.x, b.y, y_f16); v.w = pack_h2(b.z, b.w, y_f16);
        const long long W2 = 2LL * W;
        const long long o = ((n * 2 * H + 2 * h) * W2 + 2 * w) * V + vc;
        y[o] = v;
        y[o + V] = v;
        y[o + W2 * V] = v;
        y[o + W2 * V + V] = v;
    }
}

__global__ void cast_f32_bf16_kernel(const float4* __restrict__ x, uint4* __restrict__ y, long long nvec, int y_f16) {
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < nvec; i += stride) {
        const float4 a = __ldg(x + 2 * i), b = __ldg(x + 2 * i + 1);
        uint4 v;
        v.x = pack_h2(a.x, a.y, y_f16); v.y = pack_h2(a.z, a.w, y_f16);
        v.z = pack_h2(b.x, b.y, y_f16); v.w = pack_h2(b.z, b.w, y_f16);
        y[i] = v;
    }
}

// channel concat (rows x (Ca+Cb)), 16-byte vectors
__global__ void concat_kernel(const uint4* __restrict__ a, const uint4* __restrict__ b, uint4* __restrict__ y,
                              long long rows, int Va, int Vb) {
    const int V = Va + Vb;
    const long long total = rows * V;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int vc = static_cast<int>(i % V);
        const long long r = i / V;
        y[i] = (vc < Va) ? __ldg(a + r * Va + vc) : __ldg(b + r * Vb + (vc - Va));
    }
}

// ---------------------------------------------------------------------------------------------------------
// 3x3 / stride 1 / pad 1 conv with Cin <= 8 from fp32 NCHW input to bf16 NHWC output.
// CTA = 32 pixels of one image row x all Cout; thread = (pixel, 8-channel lane); weights + input patch in smem.
// ---------------------------------------------------------------------------------------------------------
constexpr int SC_PIX = 32;
constexpr int SC_ROWS = 8;
__global__ void __launch_bounds__(256)
conv3x3_small_cin_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                         void* __restrict__ y_, int y_dtype, int H, int W, int Cin, int Cout) {
    extern __shared__ float smf[];
    // layout: wsm [9*Cin][Cout] (k-major so that 8 consecutive couts are contiguous), patch [Cin][3][SC_PIX+2]
    float* wsm = smf;
    float* patch = smf + 9 * Cin * Cout;
    const int tiles_w = (W + SC_PIX - 1) / SC_PIX;
    const int tiles_h = (H + SC_ROWS - 1) / SC_ROWS;
    const int tw = blockIdx.x % tiles_w;
    const int th = (blockIdx.x / tiles_w) % tiles_h;
    const int n = blockIdx.x / (tiles_w * tiles_h);
    const int w0 = tw * SC_PIX;
    // weights: global [Cout][3][3][Cin] -> smem [(tap*Cin + ci)][Cout]   (loaded once per CTA, reused for SC_ROWS rows)
    const int K = 9 * Cin;
    for (int i = threadIdx.x; i < K * Cout; i += blockDim.x) {
        const int co = i / K, k = i % K;
        wsm[k * Cout + co] = __ldg(w + i);
    }
    const int PW = SC_PIX + 2;
    const int px = threadIdx.x % SC_PIX;
    const int cl = threadIdx.x / SC_PIX;  // 0..7
    const int wo = w0 + px;
    const bool valid = wo < W;
    for (int hr = 0; hr < SC_ROWS; ++hr) {
        const int h = th * SC_ROWS + hr;
        if (h >= H) break;
        __syncthreads();
        for (int i = threadIdx.x; i < Cin * 3 * PW; i += blockDim.x) {
            const int ppx = i % PW;
            const int r = (i / PW) % 3;
            const int ci = i / (3 * PW);
            const int hh = h + r - 1, ww = w0 + ppx - 1;
            float v = 0.f;
            if (hh >= 0 && hh < H && ww >= 0 && ww < W)
                v = __ldg(x + ((static_cast<long long>(n) * Cin + ci) * H + hh) * W + ww);
            patch[i] = v;
        }
        __syncthreads();
        const long long out_row = ((static_cast<long long>(n) * H + h) * W + wo) * Cout;
        for (int co = cl * 8; co < Cout; co += 64) {
            float acc[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = bias ? __ldg(bias + co + j) : 0.f;
            for (int ci = 0; ci < Cin; ++ci) {
#pragma unroll
                for (int r = 0; r < 3; ++r) {
#pragma unroll
                    for (int s = 0; s < 3; ++s) {
                        const float xv = patch[(ci * 3 + r) * PW + px + s];
                        const float4* wp =
                            reinterpret_cast<const float4*>(wsm + ((r * 3 + s) * Cin + ci) * Cout + co);
                        const float4 w0v = wp[0], w1v = wp[1];
                        acc[0] = fmaf(xv, w0v.x, acc[0]); acc[1] = fmaf(xv, w0v.y, acc[1]);
                        acc[2] = fmaf(xv, w0v.z, acc[2]); acc[3] = fmaf(xv, w0v.w, acc[3]);
                        acc[4] = fmaf(xv, w1v.x, acc[4]); acc[5] = fmaf(xv, w1v.y, acc[5]);
                        acc[6] = fmaf(xv, w1v.z, acc[6]); acc[7] = fmaf(xv, w1v.w, acc[7]);
                    }
                }
            }
            if (valid) {
                if (y_dtype == 1) {
                    float* y = reinterpret_cast<float*>(y_) + out_row + co;
                    *reinterpret_cast<float4*>(y) = make_float4(acc[0], acc[1], acc[2], acc[3]);
                    *reinterpret_cast<float4*>(y + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
                } else {
                    uint4 o;
                    const int hf = (y_dtype == 2);
                    o.x = pack_h2(acc[0], acc[1], hf); o.y = pack_h2(acc[2], acc[3], hf);
                    o.z = pack_h2(acc[4], acc[5], hf); o.w = pack_h2(acc[6], acc[7], hf);
                    *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(y_) + out_row + co) = o;
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// im2col for 3x3 / stride 1 / pad 1 convolutions with a tiny input channel count: fp32 NCHW -> 16-bit rows
// [N*H*W, Kpad], k = (kh*3 + kw)*Cin + c, zero-padded to Kpad (multiple of 64) so the convolution becomes one
// tcgen05 GEMM with K = Kpad.  8 lanes cover one pixel's row (16 B each), so a warp writes 4 complete rows.
// ---------------------------------------------------------------------------------------------------------
__global__ void im2col3x3_small_kernel(const float* __restrict__ x, uint16_t* __restrict__ y, long long total_vecs,
                                       int H, int W, int Cin, int Kpad, int y_f16) {
    const int VPR = Kpad / 8;                     // 16-byte vectors per output row
    const int K = 9 * Cin;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total_vecs; i += stride) {
        const int v = static_cast<int>(i % VPR);
        long long pix = i / VPR;
        const int w = static_cast<int>(pix % W);
        const long long t = pix / W;
        const int h = static_cast<int>(t % H);
        const long long n = t / H;
        float f[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int k = v * 8 + j;
            float val = 0.f;
            if (k < K) {
                const int tap = k / Cin, c = k - tap * Cin;
                const int hh = h + tap / 3 - 1, ww = w + tap % 3 - 1;
                if (hh >= 0 && hh < H && ww >= 0 && ww < W) val = __ldg(x + ((n * Cin + c) * H + hh) * W + ww);
            }
            f[j] = val;
        }
        uint4 o;
        o.x = pack_h2(f[0], f[1], y_f16); o.y = pack_h2(f[2], f[3], y_f16);
        o.z = pack_h2(f[4], f[5], y_f16); o.w = pack_h2(f[6], f[7], y_f16);
        *reinterpret_cast<uint4*>(y + i * 8) = o;
    }
}

// Compile-time (Cin, Kpad) flavour used by every layer on the path (3->128, 4->320, 4->512: Kpad 64; 8->320: Kpad 128):
// grid = (pixel blocks of a row, H, N), so no 64-bit index divisions, and tap / channel come from constant divisors.
constexpr int IM2COL_ROWS = 16;
template <int CIN, int KPAD>
__global__ void __launch_bounds__(256) im2col3x3_small_t_kernel(const float* __restrict__ x, uint16_t* __restrict__ y,
                                                                int H, int W, int y_f16) {
    constexpr int VPR = KPAD / 8;                 // 16-byte vectors per output row
    constexpr int PIX = 256 / VPR;                // pixels per CTA
    constexpr int K = 9 * CIN;
    const int v = threadIdx.x % VPR;
    const int w = blockIdx.x * PIX + threadIdx.x / VPR;
    const int n = blockIdx.z;
    if (w >= W) return;
    const float* xn = x + static_cast<size_t>(n) * CIN * H * W;
    // Everything that does not depend on the image row is computed once per thread: for each of the 8 elements of this
    // thread's 16-byte vector the offset of its tap relative to the pixel, and a 3-bit mask saying which vertical tap
    // (top / centre / bottom) it reads (0 = padding column of K, or a horizontal tap outside the image).
    int off[8];
    int rmask[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int k = v * 8 + j;
        const int tap = k / CIN, c = k - tap * CIN;
        const int dh = tap / 3, dw = tap - dh * 3;
        const int ww = w + dw - 1;
        off[j] = c * H * W + (dh - 1) * W + (dw - 1);
        rmask[j] = (k < K && ww >= 0 && ww < W) ? (1 << dh) : 0;
    }
    // IM2COL_ROWS image rows per thread: the loads of several rows are in flight together and vertically adjacent
    // taps hit in L1
#pragma unroll 4
    for (int r = 0; r < IM2COL_ROWS; ++r) {
        const int h = blockIdx.y * IM2COL_ROWS + r;
        if (h < H) {
            const int base = h * W + w;
            const int rows_ok = (h > 0 ? 1 : 0) | 2 | (h < H - 1 ? 4 : 0);
            float f[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = (rmask[j] & rows_ok) ? __ldg(xn + base + off[j]) : 0.f;
            uint4 o;
            o.x = pack_h2(f[0], f[1], y_f16); o.y = pack_h2(f[2], f[3], y_f16);
            o.z = pack_h2(f[4], f[5], y_f16); o.w = pack_h2(f[6], f[7], y_f16);
            *reinterpret_cast<uint4*>(y + ((static_cast<size_t>(n) * H + h) * W + w) * KPAD + v * 8) = o;
        }
    }
}

template <int CIN, int KPAD>
static void launch_im2col_t(const float* x, uint16_t* y, int N, int H, int W, int y_f16, cudaStream_t st) {
    constexpr int PIX = 256 / (KPAD / 8);
    const dim3 grid((W + PIX - 1) / PIX, (H + IM2COL_ROWS - 1) / IM2COL_ROWS, N);
    im2col3x3_small_t_kernel<CIN, KPAD><<<grid, 256, 0, st>>>(x, y, H, W, y_f16);
}

// ---------------------------------------------------------------------------------------------------------
// 1x1 conv on <= 8 channels with arbitrary element strides (NCHW <-> NHWC), fp32.
// ---------------------------------------------------------------------------------------------------------
struct PwParams {
    float w[64];
    float b[8];
};
__global__ void pointwise_small_kernel(const float* __restrict__ x, long long x_ns, long long x_ps, long long x_cs,
                                       PwParams pw, float in_scale, float out_scale, float* __restrict__ y,
                                       long long y_ns, long long y_ps, long long y_cs, long long total, int HW,
                                       int Cin, int Cout) {
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
        const long long n = i / HW, p = i % HW;
        float in[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) in[c] = (c < Cin) ? __ldg(x + n * x_ns + p * x_ps + c * x_cs) * in_scale : 0.f;
#pragma unroll
        for (int co = 0; co < 8; ++co) {
            if (co < Cout) {
                float a = pw.b[co];
#pragma unroll
                for (int c = 0; c < 8; ++c)
                    if (c < Cin) a = fmaf(pw.w[co * 8 + c], in[c], a);
                y[n * y_ns + p * y_ps + co * y_cs] = a * out_scale;
            }
        }
    }
}

// fp32 NHWC rows -> fp32 NCHW with affine + clamp
__global__ void nhwc_to_nchw_kernel(const float* __restrict__ x, int x_row_stride, float* __restrict__ y,
                                    long long total, int C, int HW, float scale, float shift, float lo, float hi) {
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
        // i enumerates output elements (n, c, p) so that stores are coalesced
        const long long p = i % HW;
        const long long nc = i / HW;
        const int c = static_cast<int>(nc % C);
        const long long n = nc / C;
        float v = __ldg(x + (n * HW + p) * x_row_stride + c);
        v = fminf(fmaxf(fmaf(v, scale, shift), lo), hi);
        y[i] = v;
    }
}

// seg post-processing (pipeline:787-795 and :534): clip(-1,1) -> *0.5+0.5 -> *255 ; uint8 truncation
__global__ void seg_post_kernel(const float* __restrict__ dec, int row_stride, float* __restrict__ seg_f32,
                                uint8_t* __restrict__ seg_u8, long long total, int HW) {
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
        const long long p = i % HW;
        const long long nc = i / HW;
        const int c = static_cast<int>(nc % 3);
        const long long n = nc / 3;
        float v = __ldg(dec + (n * HW + p) * row_stride + c);
        v = fminf(fmaxf(v, -1.0f), 1.0f);
        v = __fadd_rn(__fmul_rn(v, 0.5f), 0.5f);   // (seg * 0.5) + 0.5, no FMA contraction: match torch op order
        v = __fmul_rn(v, 255.0f);
        if (seg_f32) seg_f32[i] = v;
        if (seg_u8) {
            float c8 = fminf(fmaxf(v, 0.0f), 255.0f);
            seg_u8[i] = static_cast<uint8_t>(c8);  // truncation, like numpy astype(uint8) on [0,255]
        }
    }
}

int grid_for(long long total, int threads) {
    long long blocks = (total + threads - 1) / threads;
    const long long cap = static_cast<long long>(sm_count()) * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return static_cast<int>(blocks);
}

}  // namespace
}  // namespace dfw

extern "C" {

int dfw_upsample2x_nhwc(const void* x, int x_f32, void* y, int y_f16, int N, int H, int W, int C, void* stream_) {
    using namespace dfw;
    int rc = require_sm100();
    if (rc != DFW_OK) return rc;
    DFW_REQUIRE(x && y && N > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0);
    const int V = C / 8;
    const long long total = static_cast<long long>(N) * H * W * V;
    if (x_f32)
        upsample2x_f32_kernel<<<grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream_)>>>(
            reinterpret_cast<const float4*>(x), reinterpret_cast<uint4*>(y), total, H, W, V, y_f16);
    else
        upsample2x_kernel<<<grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream_)>>>(
            reinterpret_cast<const uint4*>(x), reinterpret_cast<uint4*>(y), total, H, W, V);
    g_launches.fetch_add(1);
    DFW_CHECK_CUDA(cudaGetLastError());
    return DFW_OK;
}

int dfw_concat_channels(const void* a, const void* b, void* y, long long rows, int Ca, int Cb, int elem_bytes,
                        void* stream_) {
    using namespace dfw;
    int rc = require_sm100();
    if (rc != DFW_OK) return rc;
    DFW_REQUIRE(a && b && y && rows > 0 && Ca > 0 && Cb > 0 && (elem_bytes == 2 || elem_bytes == 4));
    const int per_vec = 16 / elem_bytes;
    DFW_REQUIRE(Ca % per_vec == 0 && Cb % per_vec == 0);
    const long long total = rows * ((Ca + Cb) / per_vec);
    concat_kernel<<<grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream_)>>>(
        reinterpret_cast<const uint4*>(a), reinterpret_cast<const uint4*>(b), reinterpret_cast<uint4*>(y), rows,
        Ca / per_vec, Cb / per_vec);
    g_launches.fetch_add(1);
    DFW_CHECK_CUDA(cudaGetLastError());
    return DFW_OK;
}

int dfw_cast_f32_to_16(const float* x, void* y, int y_f16, long long n, void* stream_) {
    using namespace dfw;
    int rc = require_sm100();
    if (rc != DFW_OK) return rc;
    DFW_REQUIRE(x && y && n > 0 && n % 8 == 0);
    cast_f32_bf16_kernel<<<grid_for(n / 8, 256), 256, 0, static_cast<cudaStream_t>(stream_)>>>(
        reinterpret_cast<const float4*>(x), reinterpret_cast<uint4*>(y), n / 8, y_f16);
    g_launches.fetch_add(1);
    DFW_CHECK_CUDA(cudaGetLastError());
    return DFW_OK;
}

int dfw_im2col3x3_small(const float* x, void* y, int y_f16, int N, int H, int W, int Cin, int Kpad, void* stream_) {
    using namespace dfw;
    int rc = require_sm100();
    if (rc != DFW_OK) return rc;
    DFW_REQUIRE(x && y && N > 0 && H > 0 && W > 0 && Cin >= 1 && Cin <= 16);
    DFW_REQUIRE(Kpad % 64 == 0 && Kpad >= 9 * Cin);
    const long long total = static_cast<long long>(N) * H * W * (Kpad / 8);
    cudaStream_t st = static_cast<cudaStream_t>(stream_);
    uint16_t* y16 = reinterpret_cast<uint16_t*>(y);
    const bool grid_ok = H <= 65535 && N <= 65535 && static_cast<long long>(Cin) * H * W < (1LL << 31);
    if (grid_ok && Cin == 3 && Kpad == 64) launch_im2col_t<3, 64>(x, y16, N, H, W, y_f16, st);
    else if (grid_ok && Cin == 4 && Kpad == 64) launch_im2col_t<4, 64>(x, y16, N, H, W, y_f16, st);
    else if (grid_ok && Cin == 8 && Kpad == 128) launch_im2col_t<8, 128>(x, y16, N, H, W, y_f16, st);
    else
        im2col3x3_small_kernel<<<grid_for(total, 256), 256, 0, st>>>(x, y16, total, H, W, Cin, Kpad, y_f16);
    g_launches.fetch_add(1);
    DFW_CHECK_CUDA(cudaGetLastError());
    return DFW_OK;
}

int dfw_conv3x3_small_cin(const float* x, const float* w, const float* bias, void* y, int y_dtype, int N, int H,
                          int W, int Cin, int Cout, void* stream_) {
    using namespace dfw;
    int rc = require_sm100();
    if (rc != DFW_OK) return rc;
    DFW_REQUIRE(x && w && y && N > 0 && H > 0 && W > 0);
    DFW_REQUIRE(Cin >= 1 && Cin <= 8 && Cout % 64 == 0 && Cout <= 512);
    const size_t smem = (static_cast<size_t>(9) * Cin * Cout + static_cast<size_t>(Cin) * 3 * (SC_PIX + 2)) * 4;
    DFW_REQUIRE(smem <= 200 * 1024);
    static size_t smem_set = 0;
    if (smem > 48 * 1024 && smem > smem_set) {
        DFW_CHECK_CUDA(cudaFuncSetAttribute(conv3x3_small_cin_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            200 * 1024));
        smem_set = 200 * 1024;
    }
    const int tiles_w = (W + SC_PIX - 1) / SC_PIX;
    const long long blocks = static_cast<long long>(N) * ((H + SC_ROWS - 1) / SC_ROWS) * tiles_w;
    DFW_REQUIRE(blocks < (1LL << 31));
    conv3x3_small_cin_kernel<<<static_cast<int>(blocks), 256, smem, static_cast<cudaStream_t>(stream_)>>>(
        x, w, bias, y, y_dtype, H, W, Cin, Cout);
    g_launches.fetch_add(1);
    DFW_CHECK_CUDA(cudaGetLastError());
    return DFW_OK;
}

int dfw_pointwise_small(const float* x, long long x_ns, long long x_ps, long long x_cs, const float* w_host,
                        const float* b_host, float in_scale, float out_scale, float* y, long long y_ns,
                        long long y_ps, long long y_cs, int N, int HW, int Cin, int Cout, void* stream_) {
    using namespace dfw;
    int rc = require_sm100();
    if (rc != DFW_OK) return rc;
    DFW_REQUIRE(x && w_host && y && N > 0 && HW > 0 && Cin >= 1 && Cin <= 8 && Cout >= 1 && Cout <= 8);
    PwParams pw{};
    for (int co = 0; co < Cout; ++co) {
        for (int c = 0; c < Cin; ++c) pw.w[co * 8 + c] = w_host[co * Cin + c];
        pw.b[co] = b_host ? b_host[co] : 0.f;
    }
    const long long total = static_cast<long long>(N) * HW;
    pointwise_small_kernel<<<grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream_)>>>(
        x, x_ns, x_ps, x_cs, pw, in_scale, out_scale, y, y_ns, y_ps, y_cs, total, HW, Cin, Cout);
    g_launches.fetch_add(1);
    DFW_CHECK_CUDA(cudaGetLastError());
    return DFW_OK;
}

int dfw_nhwc_f32_to_nchw_f32(const float* x, int x_row_stride, float* y, int N, int C, int HW, float scale,
                             float shift, float lo, float hi, void* stream_) {
    using namespace dfw;
    int rc = require_sm100();
    if (rc != DFW_OK) return rc;
    DFW_REQUIRE(x && y && N > 0 && C > 0 && HW > 0 && x_row_stride >= C);
    const long long total = static_cast<long long>(N) * C * HW;
    nhwc_to_nchw_kernel<<<grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream_)>>>(
        x, x_row_stride, y, total, C, HW, scale, shift, lo, hi);
    g_launches.fetch_add(1);
    DFW_CHECK_CUDA(cudaGetLastError());
    return DFW_OK;
}

int dfw_seg_post(const float* dec, int row_stride, float* seg_f32, uint8_t* seg_u8, int N, int HW, void* stream_) {
    using namespace dfw;
    int rc = require_sm100();
    if (rc != DFW_OK) return rc;
    DFW_REQUIRE(dec && (seg_f32 || seg_u8) && N > 0 && HW > 0 && row_stride >= 3);
    const long long total = static_cast<long long>(N) * 3 * HW;
    seg_post_kernel<<<grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream_)>>>(dec, row_stride, seg_f32,
                                                                                           seg_u8, total, HW);
    g_launches.fetch_add(1);
    DFW_CHECK_CUDA(cudaGetLastError());
    return DFW_OK;
}

}  // extern "C"
