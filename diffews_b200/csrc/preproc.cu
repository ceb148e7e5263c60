// K9 — episode preprocessing on the GPU: the data format either side of the hot path.
//
// ref: evaluation_util/data/dataset.py:36-40   transforms.Resize((S,S)) -> ToTensor() -> Normalize([0.5],[0.5])
//      evaluation_util/data/coco.py:38-47      transform(query / supports); F.interpolate(mask, (S,S), mode='nearest')
//      evaluation_util/data/coco.py:92-93      mask = (label == class + 1)
//      evaluation_util/data/pascal.py:78-83    boundary = floor(mask / 255) (ignore index), mask = (label == class + 1)
//      evaluation_util/data/fss.py:80-84       mask = (L >= 128)
//
// `Resize` on a PIL image is Pillow's ImagingResample (third-party, src/libImaging/Resample.c): a separable two-pass
// convolution — horizontal first, uint8 intermediate, then vertical — with per-output-pixel coefficient windows computed
// in double precision and quantised to 22-bit fixed point.  This file restates that algorithm so the result is the same
// BYTES as the reference's CPU path: the coefficient arithmetic uses explicit round-to-nearest double intrinsics in the
// reference's operation order (no FMA contraction), the accumulation is the same int32 sum with the same rounding
// constant, and ToTensor / Normalize are the same three fp32 operations (true division by 255, subtract, divide).
//
// All images of a batch (different sizes) are processed by ONE launch per pass: the caller packs the decoded images
// into one device buffer and passes a descriptor table (offset, h, w, row stride) that lives in the same buffer.
#include <atomic>
#include <cstdlib>

#include "common.cuh"

namespace dfw {
extern std::atomic<long long> g_launches;
namespace {

constexpr int PRECISION_BITS = 32 - 8 - 2;      // Resample.c

struct ImageDesc {                               // mirrors DfwImageDesc (include/diffews_b200.h)
    long long offset;
    int h, w, row_stride, param;
};

struct CoefLayout {                              // workspace: [bounds int2][kk int32][tmp u8]
    int out_max, ks;
    long long kk_off, tmp_off;                   // byte offsets
    long long tmp_per_image;
};

__host__ __device__ inline int ksize_for(int in_size, int out_size) {
    const double scale = static_cast<double>(in_size) / out_size;
    const double fs = scale < 1.0 ? 1.0 : scale;
    return static_cast<int>(ceil(fs)) * 2 + 1;
}

// Resample.c precompute_coeffs + normalize_coeffs_8bpc for ONE output index of one axis, bilinear filter, box = the
// whole axis.  Writes ks fixed-point coefficients (zero padded) to k[] and returns (first source index, count).
__device__ __forceinline__ int2 resample_coeffs(int in_size, int out_size, int xx, int ks, int* k) {
    const double scale = __ddiv_rn(static_cast<double>(in_size), static_cast<double>(out_size));
    const double filterscale = scale < 1.0 ? 1.0 : scale;
    const double support = filterscale;                       // bilinear support 1.0 * filterscale
    const double ss = __ddiv_rn(1.0, filterscale);
    const double center = __dmul_rn(static_cast<double>(xx) + 0.5, scale);   // in0 == 0
    int xmin = static_cast<int>(__dadd_rn(__dsub_rn(center, support), 0.5));
    if (xmin < 0) xmin = 0;
    int xmax = static_cast<int>(__dadd_rn(__dadd_rn(center, support), 0.5));
    if (xmax > in_size) xmax = in_size;
    int n = xmax - xmin;
    if (n > ks) n = ks;                                       // cannot happen (ks is sized from the largest image)
    double ww = 0.0;
    for (int x = 0; x < n; ++x) {
        double t = __dmul_rn(__dadd_rn(__dsub_rn(static_cast<double>(x + xmin), center), 0.5), ss);
        if (t < 0.0) t = -t;
        ww = __dadd_rn(ww, t < 1.0 ? __dsub_rn(1.0, t) : 0.0);
    }
    for (int x = 0; x < ks; ++x) {
        int q = 0;
        if (x < n) {
            double t = __dmul_rn(__dadd_rn(__dsub_rn(static_cast<double>(x + xmin), center), 0.5), ss);
            if (t < 0.0) t = -t;
            double w = t < 1.0 ? __dsub_rn(1.0, t) : 0.0;
            if (ww != 0.0) w = __ddiv_rn(w, ww);
            const double v = __dmul_rn(w, static_cast<double>(1 << PRECISION_BITS));
            q = w < 0.0 ? static_cast<int>(__dadd_rn(-0.5, v)) : static_cast<int>(__dadd_rn(0.5, v));
        }
        k[x] = q;
    }
    return make_int2(xmin, n);
}

// Two-pass path, kernel 1: one thread per (output index, axis, image).
__global__ void preproc_coeffs_kernel(const ImageDesc* __restrict__ descs, int2* __restrict__ bounds,
                                      int* __restrict__ kk, int out_w, int out_h, int out_max, int ks) {
    const int img = blockIdx.z, axis = blockIdx.y;
    const int xx = blockIdx.x * blockDim.x + threadIdx.x;
    const int out_size = axis == 0 ? out_w : out_h;
    if (xx >= out_size) return;
    const ImageDesc d = descs[img];
    const long long slot = (static_cast<long long>(img) * 2 + axis) * out_max + xx;
    bounds[slot] = resample_coeffs(axis == 0 ? d.w : d.h, out_size, xx, ks, kk + slot * ks);
}

__device__ __forceinline__ uint8_t clip8(int v) {
    v >>= PRECISION_BITS;
    return static_cast<uint8_t>(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// Horizontal pass: src [h, w, 3] u8 -> tmp [h, out_w, 3] u8.  Thread = one output pixel (3 channels).
__global__ void preproc_horizontal_kernel(const uint8_t* __restrict__ base, const ImageDesc* __restrict__ descs,
                                          const int2* __restrict__ bounds, const int* __restrict__ kk,
                                          uint8_t* __restrict__ tmp, long long tmp_per_image, int out_w, int out_max,
                                          int ks) {
    const int img = blockIdx.z;
    const ImageDesc d = descs[img];
    const int xx = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (xx >= out_w || y >= d.h) return;
    const long long slot = (static_cast<long long>(img) * 2 + 0) * out_max + xx;
    const int2 b = bounds[slot];
    const int* k = kk + slot * ks;
    const uint8_t* row = base + d.offset + static_cast<long long>(y) * d.row_stride + b.x * 3;
    int s0 = 1 << (PRECISION_BITS - 1), s1 = s0, s2 = s0;
    for (int x = 0; x < b.y; ++x) {
        const int c = __ldg(k + x);
        s0 += static_cast<int>(__ldg(row + 3 * x + 0)) * c;
        s1 += static_cast<int>(__ldg(row + 3 * x + 1)) * c;
        s2 += static_cast<int>(__ldg(row + 3 * x + 2)) * c;
    }
    uint8_t* o = tmp + static_cast<long long>(img) * tmp_per_image + (static_cast<long long>(y) * out_w + xx) * 3;
    o[0] = clip8(s0); o[1] = clip8(s1); o[2] = clip8(s2);
}

// Vertical pass + ToTensor + Normalize: tmp [h, out_w, 3] u8 -> dst [3, out_h, out_w] fp32 (and / or u8 HWC).
__global__ void preproc_vertical_kernel(const ImageDesc* __restrict__ descs, const int2* __restrict__ bounds,
                                        const int* __restrict__ kk, const uint8_t* __restrict__ tmp,
                                        long long tmp_per_image, float* __restrict__ dst_f32,
                                        uint8_t* __restrict__ dst_u8, int out_w, int out_h, int out_max, int ks,
                                        float mean, float stdv) {
    __shared__ float lut[256];
    // ToTensor: fp32(v) / 255 (true division); Normalize: (x - mean) / std — the reference's three fp32 roundings
    for (int v = threadIdx.y * blockDim.x + threadIdx.x; v < 256; v += blockDim.x * blockDim.y)
        lut[v] = __fdiv_rn(__fsub_rn(__fdiv_rn(static_cast<float>(v), 255.0f), mean), stdv);
    __syncthreads();
    const int img = blockIdx.z;
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int yy = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= out_w || yy >= out_h) return;
    const long long slot = (static_cast<long long>(img) * 2 + 1) * out_max + yy;
    const int2 b = bounds[slot];
    const int* k = kk + slot * ks;
    const uint8_t* col = tmp + static_cast<long long>(img) * tmp_per_image +
                         (static_cast<long long>(b.x) * out_w + x) * 3;
    int s0 = 1 << (PRECISION_BITS - 1), s1 = s0, s2 = s0;
    for (int j = 0; j < b.y; ++j) {
        const int c = __ldg(k + j);
        const uint8_t* p = col + static_cast<long long>(j) * out_w * 3;
        s0 += static_cast<int>(p[0]) * c;
        s1 += static_cast<int>(p[1]) * c;
        s2 += static_cast<int>(p[2]) * c;
    }
    const uint8_t r = clip8(s0), g = clip8(s1), bl = clip8(s2);
    const long long plane = static_cast<long long>(out_h) * out_w;
    const long long pix = static_cast<long long>(yy) * out_w + x;
    if (dst_f32) {
        float* o = dst_f32 + static_cast<long long>(img) * 3 * plane + pix;
        o[0] = lut[r]; o[plane] = lut[g]; o[2 * plane] = lut[bl];
    }
    if (dst_u8) {
        uint8_t* o = dst_u8 + (static_cast<long long>(img) * plane + pix) * 3;
        o[0] = r; o[1] = g; o[2] = bl;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Fused path (default): coefficient kernel + ONE resample launch, no intermediate image in HBM.  CTA = one 16 x 64
// output tile of one image:
//   B  the source rectangle the tile's coefficient windows touch is staged in smem with aligned 32-bit loads (rows
//      keep their global byte misalignment, so word i of a row in smem is word i of the row in HBM)
//   C  horizontal pass smem -> smem (uint8, as Pillow's intermediate image); thread = one output column, strided rows
//   D  vertical pass + ToTensor + Normalize; thread = 4 consecutive pixels of one output row -> three 128-bit plane
//      stores (and / or the uint8 HWC image).
// The per-image coefficient tables (a few KB, written by preproc_coeffs_kernel) are read through L1.
// HBM traffic = source bytes (+ the 1-2 row / column overlap between neighbouring tiles, absorbed by L2) + output.
// ---------------------------------------------------------------------------------------------------------------
constexpr int FT_H = 16, FT_W = 64, FT_THREADS = 256;
static_assert(FT_H * FT_W / 4 == FT_THREADS, "pass D: one 4-pixel group per thread");

struct FusedLayout {            // dynamic smem carve-up (bytes), computed on the host from the largest image
    int ks, rmax, row_bytes;    // coefficient slots, max source rows per tile, smem bytes per staged source row
    int off_hz, off_shift, total;
};

__global__ void __launch_bounds__(FT_THREADS)
preproc_fused_kernel(const uint8_t* __restrict__ base, const ImageDesc* __restrict__ descs,
                     const int2* __restrict__ bounds, const int* __restrict__ kk, float* __restrict__ dst_f32,
                     uint8_t* __restrict__ dst_u8, int out_w, int out_h, int out_max, float mean, float stdv,
                     const FusedLayout L) {
    extern __shared__ __align__(16) uint8_t fsm[];
    __shared__ float lut[256];
    uint8_t* src = fsm;
    uint8_t* hz = fsm + L.off_hz;
    int* shift = reinterpret_cast<int*>(fsm + L.off_shift);
    const int t = threadIdx.x;
    const int img = blockIdx.z;
    const ImageDesc d = descs[img];
    const int x0 = blockIdx.x * FT_W, y0 = blockIdx.y * FT_H;
    const int tw = min(FT_W, out_w - x0), th = min(FT_H, out_h - y0);
    lut[t] = __fdiv_rn(__fsub_rn(__fdiv_rn(static_cast<float>(t), 255.0f), mean), stdv);     // FT_THREADS == 256
    const long long slot_x = (static_cast<long long>(img) * 2 + 0) * out_max + x0;
    const long long slot_y = (static_cast<long long>(img) * 2 + 1) * out_max + y0;
    // source rectangle (windows are monotone in the output index)
    const int2 bxl = __ldg(bounds + slot_x + tw - 1), byl = __ldg(bounds + slot_y + th - 1);
    const int c0 = __ldg(bounds + slot_x).x, c1 = bxl.x + bxl.y;
    const int r0 = __ldg(bounds + slot_y).x, r1 = byl.x + byl.y;
    const int nrows = r1 - r0, nbytes = (c1 - c0) * 3;
    if (nrows > L.rmax || nbytes + 7 > L.row_bytes) { asm volatile("trap;"); }      // host sizing bug: never silently wrong
    // B: stage rows r0..r1, bytes [c0*3, c1*3) with aligned words
    const int warp = t >> 5, lane = t & 31;
    for (int r = warp; r < nrows; r += FT_THREADS / 32) {
        const long long g = d.offset + static_cast<long long>(r0 + r) * d.row_stride + static_cast<long long>(c0) * 3;
        const int sh = static_cast<int>(g & 3);
        const uint32_t* gp = reinterpret_cast<const uint32_t*>(base + (g - sh));
        uint32_t* sp = reinterpret_cast<uint32_t*>(src + r * L.row_bytes);
        const int nwords = (sh + nbytes + 3) >> 2;
        for (int w = lane; w < nwords; w += 32) sp[w] = __ldg(gp + w);
        if (lane == 0) shift[r] = sh;
    }
    __syncthreads();
    // C: horizontal pass, smem -> smem.  Thread = output column xx (window and coefficients fixed), rows r, r+4, ...
    {
        const int xx = t & (FT_W - 1);
        if (xx < tw) {
            const int2 bw = __ldg(bounds + slot_x + xx);
            const int* k = kk + (slot_x + xx) * L.ks;
            const int poff = (bw.x - c0) * 3;
            for (int r = t / FT_W; r < nrows; r += FT_THREADS / FT_W) {
                const uint8_t* p = src + r * L.row_bytes + shift[r] + poff;
                int s0 = 1 << (PRECISION_BITS - 1), s1 = s0, s2 = s0;
                for (int x = 0; x < bw.y; ++x) {
                    const int c = __ldg(k + x);
                    s0 += static_cast<int>(p[3 * x + 0]) * c;
                    s1 += static_cast<int>(p[3 * x + 1]) * c;
                    s2 += static_cast<int>(p[3 * x + 2]) * c;
                }
                uint8_t* o = hz + (r * FT_W + xx) * 3;
                o[0] = clip8(s0); o[1] = clip8(s1); o[2] = clip8(s2);
            }
        }
    }
    __syncthreads();
    // D: vertical pass + ToTensor + Normalize.  Thread = pixels x4 .. x4+3 of output row yy (12 bytes per source row).
    const int yy = t / (FT_W / 4), x4 = (t & (FT_W / 4 - 1)) * 4;
    if (yy >= th || x4 >= tw) return;
    const int2 bh = __ldg(bounds + slot_y + yy);
    const int* k = kk + (slot_y + yy) * L.ks;
    int acc[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) acc[i] = 1 << (PRECISION_BITS - 1);
    const uint32_t* p = reinterpret_cast<const uint32_t*>(hz + ((bh.x - r0) * FT_W + x4) * 3);
    for (int j = 0; j < bh.y; ++j) {
        const int c = __ldg(k + j);
        const uint32_t w0 = p[0], w1 = p[1], w2 = p[2];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            acc[i] += static_cast<int>((w0 >> (8 * i)) & 255u) * c;
            acc[4 + i] += static_cast<int>((w1 >> (8 * i)) & 255u) * c;
            acc[8 + i] += static_cast<int>((w2 >> (8 * i)) & 255u) * c;
        }
        p += FT_W * 3 / 4;
    }
    uint8_t v[12];                                   // v[3 * pixel + channel]
#pragma unroll
    for (int i = 0; i < 12; ++i) v[i] = clip8(acc[i]);
    const long long plane = static_cast<long long>(out_h) * out_w;
    const long long pix = static_cast<long long>(y0 + yy) * out_w + x0 + x4;
    const int npx = min(4, tw - x4);
    if (dst_f32) {
        float* o = dst_f32 + static_cast<long long>(img) * 3 * plane + pix;
        if (npx == 4 && (out_w & 3) == 0 && (reinterpret_cast<uintptr_t>(dst_f32) & 15) == 0) {
#pragma unroll
            for (int ch = 0; ch < 3; ++ch)
                *reinterpret_cast<float4*>(o + ch * plane) = make_float4(lut[v[ch]], lut[v[3 + ch]], lut[v[6 + ch]], lut[v[9 + ch]]);
        } else {
            for (int i = 0; i < npx; ++i) {
                o[i] = lut[v[3 * i]]; o[plane + i] = lut[v[3 * i + 1]]; o[2 * plane + i] = lut[v[3 * i + 2]];
            }
        }
    }
    if (dst_u8) {
        uint8_t* o = dst_u8 + (static_cast<long long>(img) * plane + pix) * 3;
        for (int i = 0; i < npx * 3; ++i) o[i] = v[i];
    }
}

// Upper bounds of the source extent one tile touches along an axis: the last window ends at most
// (T-1)*scale + 2*support + 1 source samples after the first one starts; +2 of slack for the roundings.
inline int tile_extent(int in_size, int out_size, int T) {
    const double scale = static_cast<double>(in_size) / out_size;
    const double support = scale < 1.0 ? 1.0 : scale;
    return static_cast<int>((T - 1) * scale + 2.0 * support) + 3;
}

FusedLayout fused_layout(int max_h, int max_w, int out_h, int out_w) {
    FusedLayout L;
    const int kh = ksize_for(max_h, out_h), kw = ksize_for(max_w, out_w);
    L.ks = kh > kw ? kh : kw;
    L.rmax = tile_extent(max_h, out_h, FT_H);
    const int cmax = tile_extent(max_w, out_w, FT_W);
    L.row_bytes = (cmax * 3 + 3 + 4 + 15) & ~15;
    int off = L.rmax * L.row_bytes;
    L.off_hz = off; off += L.rmax * FT_W * 3;
    off = (off + 15) & ~15;
    L.off_shift = off; off += L.rmax * 4;
    L.total = off;
    return L;
}

// F.interpolate(mode='nearest') source index (ATen UpSampleKernel.cpp nearest_idx): fp32 product, floor, clamp.
__device__ __forceinline__ int nearest_src(int dst, int in_size, int out_size) {
    if (out_size == in_size) return dst;
    if (out_size == 2 * in_size) return dst >> 1;
    const float scale = __fdiv_rn(static_cast<float>(in_size), static_cast<float>(out_size));
    const int s = static_cast<int>(floorf(__fmul_rn(static_cast<float>(dst), scale)));
    return s < in_size - 1 ? s : in_size - 1;
}

// Label mask [h, w] u8 -> binary mask [out_h, out_w] fp32 (+ optional ignore boundary).  mode 0: label == param
// (COCO / PASCAL: param = class + 1); mode 1: label >= 128 (FSS-1000).  boundary = floor(label / 255) = (label == 255).
// Thread = 4 consecutive output pixels of one row (one 128-bit store per output tensor).
__global__ void preproc_mask_kernel(const uint8_t* __restrict__ base, const ImageDesc* __restrict__ descs,
                                    float* __restrict__ mask_out, float* __restrict__ boundary_out, int out_h,
                                    int out_w, int mode) {
    const int img = blockIdx.z;
    const ImageDesc d = descs[img];
    const int x4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const int yy = blockIdx.y * blockDim.y + threadIdx.y;
    if (x4 >= out_w || yy >= out_h) return;
    const int sy = nearest_src(yy, d.h, out_h);
    const uint8_t* row = base + d.offset + static_cast<long long>(sy) * d.row_stride;
    const int npx = min(4, out_w - x4);
    float m[4] = {0.f, 0.f, 0.f, 0.f}, bd[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if (i < npx) {
            const int v = __ldg(row + nearest_src(x4 + i, d.w, out_w));
            m[i] = (mode == 0 ? (v == d.param) : (v >= 128)) ? 1.0f : 0.0f;
            bd[i] = v == 255 ? 1.0f : 0.0f;
        }
    }
    const long long o = (static_cast<long long>(img) * out_h + yy) * out_w + x4;
    const bool vec = npx == 4 && (out_w & 3) == 0 && (reinterpret_cast<uintptr_t>(mask_out) & 15) == 0 &&
                     (reinterpret_cast<uintptr_t>(boundary_out) & 15) == 0;
    if (vec) {
        *reinterpret_cast<float4*>(mask_out + o) = make_float4(m[0], m[1], m[2], m[3]);
        if (boundary_out) *reinterpret_cast<float4*>(boundary_out + o) = make_float4(bd[0], bd[1], bd[2], bd[3]);
    } else {
        for (int i = 0; i < npx; ++i) {
            mask_out[o + i] = m[i];
            if (boundary_out) boundary_out[o + i] = bd[i];
        }
    }
}

CoefLayout coef_layout(int n, int max_h, int max_w, int out_h, int out_w) {
    CoefLayout L;
    L.out_max = out_h > out_w ? out_h : out_w;
    const int kh = ksize_for(max_h, out_h), kw = ksize_for(max_w, out_w);
    L.ks = kh > kw ? kh : kw;
    const long long slots = static_cast<long long>(n) * 2 * L.out_max;
    L.kk_off = slots * static_cast<long long>(sizeof(int2));
    L.tmp_off = L.kk_off + slots * L.ks * static_cast<long long>(sizeof(int));
    L.tmp_off = (L.tmp_off + 255) & ~255LL;
    L.tmp_per_image = (static_cast<long long>(max_h) * out_w * 3 + 255) & ~255LL;
    return L;
}

}  // namespace
}  // namespace dfw

extern "C" {

long long dfw_preproc_workspace_bytes(int n, int max_h, int max_w, int out_h, int out_w) {
    if (n <= 0 || max_h <= 0 || max_w <= 0 || out_h <= 0 || out_w <= 0) return -1;
    const dfw::CoefLayout L = dfw::coef_layout(n, max_h, max_w, out_h, out_w);
    return L.tmp_off + static_cast<long long>(n) * L.tmp_per_image;
}

int dfw_resize_normalize_u8(const void* base, const void* descs, int n, int max_h, int max_w, float* dst_f32,
                            uint8_t* dst_u8, int out_h, int out_w, float mean, float stdv, void* workspace,
                            long long workspace_bytes, void* stream_) {
    using namespace dfw;
    int rc = require_sm100();
    if (rc != DFW_OK) return rc;
    DFW_REQUIRE(base && descs && workspace && (dst_f32 || dst_u8));
    DFW_REQUIRE(n > 0 && n <= 65535 && max_h > 0 && max_w > 0 && out_h > 0 && out_w > 0);
    DFW_REQUIRE(max_h < (1 << 15) && max_w < (1 << 15) && out_h < (1 << 15) && out_w < (1 << 15));
    DFW_REQUIRE(stdv != 0.0f);
    DFW_REQUIRE((reinterpret_cast<uintptr_t>(descs) & 7) == 0 && (reinterpret_cast<uintptr_t>(workspace) & 255) == 0);
    DFW_REQUIRE(workspace_bytes >= dfw_preproc_workspace_bytes(n, max_h, max_w, out_h, out_w));
    cudaStream_t st = static_cast<cudaStream_t>(stream_);
    const ImageDesc* d = reinterpret_cast<const ImageDesc*>(descs);
    const uint8_t* b = reinterpret_cast<const uint8_t*>(base);
    DFW_REQUIRE((reinterpret_cast<uintptr_t>(base) & 3) == 0);
    const FusedLayout F = fused_layout(max_h, max_w, out_h, out_w);
    const bool two_pass_forced = get_option(DFW_OPT_PREPROC_TWO_PASS) != 0;
    const CoefLayout L = coef_layout(n, max_h, max_w, out_h, out_w);
    uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
    int2* bounds = reinterpret_cast<int2*>(ws);
    int* kk = reinterpret_cast<int*>(ws + L.kk_off);
    uint8_t* tmp = ws + L.tmp_off;
    preproc_coeffs_kernel<<<dim3((L.out_max + 127) / 128, 2, n), 128, 0, st>>>(d, bounds, kk, out_w, out_h, L.out_max, L.ks);
    if (!two_pass_forced && F.total <= 46 * 1024 && (out_h + FT_H - 1) / FT_H <= 65535) {
        // fused resample (no intermediate image); very large reductions (> 46 KB of staging) take the two-pass path
        preproc_fused_kernel<<<dim3((out_w + FT_W - 1) / FT_W, (out_h + FT_H - 1) / FT_H, n), FT_THREADS, F.total, st>>>(
            b, d, bounds, kk, dst_f32, dst_u8, out_w, out_h, L.out_max, mean, stdv, F);
        g_launches.fetch_add(2);
        DFW_CHECK_CUDA(cudaGetLastError());
        return DFW_OK;
    }
    const dim3 blk(64, 4);
    preproc_horizontal_kernel<<<dim3((out_w + 63) / 64, (max_h + 3) / 4, n), blk, 0, st>>>(b, d, bounds, kk, tmp, L.tmp_per_image,
                                                                                            out_w, L.out_max, L.ks);
    preproc_vertical_kernel<<<dim3((out_w + 63) / 64, (out_h + 3) / 4, n), blk, 0, st>>>(d, bounds, kk, tmp, L.tmp_per_image,
                                                                                          dst_f32, dst_u8, out_w, out_h,
                                                                                          L.out_max, L.ks, mean, stdv);
    g_launches.fetch_add(3);
    DFW_CHECK_CUDA(cudaGetLastError());
    return DFW_OK;
}

int dfw_mask_nearest(const void* base, const void* descs, int n, float* mask_out, float* boundary_out, int out_h,
                     int out_w, int mode, void* stream_) {
    using namespace dfw;
    int rc = require_sm100();
    if (rc != DFW_OK) return rc;
    DFW_REQUIRE(base && descs && mask_out && n > 0 && n <= 65535 && out_h > 0 && out_w > 0 && (mode == 0 || mode == 1));
    DFW_REQUIRE((out_h + 3) / 4 <= 65535);
    DFW_REQUIRE((reinterpret_cast<uintptr_t>(descs) & 7) == 0);
    preproc_mask_kernel<<<dim3((out_w + 255) / 256, (out_h + 3) / 4, n), dim3(64, 4), 0, static_cast<cudaStream_t>(stream_)>>>(
        reinterpret_cast<const uint8_t*>(base), reinterpret_cast<const ImageDesc*>(descs), mask_out, boundary_out, out_h,
        out_w, mode);
    g_launches.fetch_add(1);
    DFW_CHECK_CUDA(cudaGetLastError());
    return DFW_OK;
}

}  // extern "C"
