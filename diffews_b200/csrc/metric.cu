// K8 — reverse-threshold (rthres) binarisation + intersection/union histogram, integer counts.
//
// ref: evaluation_util/main_oss.py:128-134   pred = to_tensor(PIL) (uint8/255, fp32, CPU);
//                                            thr  = pred.max() * r_threshold;  mask = pred.mean(dim=1) > thr
//      evaluation_util/common/evaluation.py:12-39   histc(bins=2) of pred[pred==gt], pred, gt; ignore_index 255
//      evaluation_util/common/logger.py:35-37       index_add_ into [2, nclass]
//
// The fp32 CPU expression is reproduced bit-exactly: ((R/255 + G/255) + B/255) / 3 with IEEE division and no FMA
// contraction (the integer rule 4(R+G+B) > 3*max is wrong on 2292 exact-tie combinations).
// Counting: per-thread -> warp shuffle reduction -> shared-memory atomics -> one global atomic per CTA per counter;
// the last CTA of each episode (ticket) folds the counters into area_inter / area_union.
#include <atomic>

#include "common.cuh"
#include "ptx.cuh"

namespace dfw {
extern std::atomic<long long> g_launches;
namespace {

struct EpisodeWs {            // 64 bytes per episode
    unsigned int max_u8;
    unsigned int ticket;
    unsigned long long cnt[6];  // inter0, inter1, pred0, pred1, gt0, gt1
    unsigned long long pad;
};
static_assert(sizeof(EpisodeWs) == 64, "workspace record");

__global__ void rthres_max_kernel(const uint8_t* __restrict__ pred, EpisodeWs* __restrict__ ws, long long per_ep) {
    const int b = blockIdx.y;
    const uint8_t* p = pred + static_cast<long long>(b) * per_ep;
    unsigned int m = 0;
    // 128-bit loads only when this episode's plane group starts 16-byte aligned (3*H*W % 16 != 0 shifts episodes b > 0)
    const long long nvec = (reinterpret_cast<uintptr_t>(p) & 15) == 0 ? per_ep / 16 : 0;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < nvec; i += stride) {
        uint4 v = __ldg(reinterpret_cast<const uint4*>(p) + i);
        unsigned int a = __vmaxu4(__vmaxu4(v.x, v.y), __vmaxu4(v.z, v.w));
        a = max(max(a & 0xFFu, (a >> 8) & 0xFFu), max((a >> 16) & 0xFFu, a >> 24));
        m = max(m, a);
    }
    for (long long i = nvec * 16 + static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < per_ep; i += stride)
        m = max(m, static_cast<unsigned int>(p[i]));
    m = __reduce_max_sync(0xffffffffu, m);
    if ((threadIdx.x & 31) == 0 && m > 0) atomicMax(&ws[b].max_u8, m);
}

__device__ __forceinline__ bool rthres_pixel(unsigned int r, unsigned int g, unsigned int bl, float thr) {
    const float fr = __fdiv_rn(static_cast<float>(r), 255.0f);
    const float fg = __fdiv_rn(static_cast<float>(g), 255.0f);
    const float fb = __fdiv_rn(static_cast<float>(bl), 255.0f);
    const float mean = __fdiv_rn(__fadd_rn(__fadd_rn(fr, fg), fb), 3.0f);
    return mean > thr;
}

__global__ void __launch_bounds__(256)
rthres_hist_kernel(const uint8_t* __restrict__ pred, int pred_is_mask, const uint8_t* __restrict__ gt,
                   const uint8_t* __restrict__ ignore, float r_threshold, EpisodeWs* __restrict__ ws,
                   long long* __restrict__ area_inter, long long* __restrict__ area_union,
                   uint8_t* __restrict__ mask_out, int HW) {
    __shared__ unsigned int s_cnt[6];
    __shared__ bool s_last;
    const int b = blockIdx.y;
    if (threadIdx.x < 6) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const float thr = __fmul_rn(__fdiv_rn(static_cast<float>(ws[b].max_u8), 255.0f), r_threshold);
    // pred_is_mask: pred is an already-binarised [B,H,W] {0,1} mask (Evaluator.classify_prediction drop-in)
    const uint8_t* pr = pred + static_cast<long long>(b) * (pred_is_mask ? 1 : 3) * HW;
    const uint8_t* pg = pred_is_mask ? pr : pr + HW;
    const uint8_t* pb = pred_is_mask ? pr : pg + HW;
    const uint8_t* gtb = gt + static_cast<long long>(b) * HW;
    const uint8_t* igb = ignore ? ignore + static_cast<long long>(b) * HW : nullptr;
    uint8_t* mo = mask_out ? mask_out + static_cast<long long>(b) * HW : nullptr;
    unsigned int c[6] = {0, 0, 0, 0, 0, 0};
    auto tally = [&](unsigned int r, unsigned int g, unsigned int bl, unsigned int gv, unsigned int ig) -> uint8_t {
        unsigned int pv = pred_is_mask ? (r ? 1u : 0u) : (rthres_pixel(r, g, bl, thr) ? 1u : 0u);
        if (ig) { gv = 255u; pv = 255u; }          // gt += ignore*255 ; pred[gt==255] = 255
        if (pv == gv && pv < 2u) c[pv]++;          // histc(pred[pred==gt]) drops the 255 bin
        if (pv < 2u) c[2 + pv]++;
        if (gv < 2u) c[4 + gv]++;
        return static_cast<uint8_t>(pv);
    };
    // 32-bit vector path only when every plane of this episode is 4-byte aligned (odd H*W or an offset view: scalar)
    const uintptr_t align_or = reinterpret_cast<uintptr_t>(pr) | reinterpret_cast<uintptr_t>(pg) | reinterpret_cast<uintptr_t>(pb) |
                               reinterpret_cast<uintptr_t>(gtb) | reinterpret_cast<uintptr_t>(igb) | reinterpret_cast<uintptr_t>(mo);
    const int nvec = (align_or & 3) == 0 ? HW / 4 : 0;
    const int stride = gridDim.x * blockDim.x;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
        const unsigned int r4 = __ldg(reinterpret_cast<const unsigned int*>(pr) + i);
        const unsigned int g4 = __ldg(reinterpret_cast<const unsigned int*>(pg) + i);
        const unsigned int b4 = __ldg(reinterpret_cast<const unsigned int*>(pb) + i);
        const unsigned int t4 = __ldg(reinterpret_cast<const unsigned int*>(gtb) + i);
        const unsigned int i4 = igb ? __ldg(reinterpret_cast<const unsigned int*>(igb) + i) : 0u;
        unsigned int m4 = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int sh = 8 * k;
            const uint8_t pv = tally((r4 >> sh) & 0xFFu, (g4 >> sh) & 0xFFu, (b4 >> sh) & 0xFFu, (t4 >> sh) & 0xFFu,
                                     (i4 >> sh) & 0xFFu);
            m4 |= static_cast<unsigned int>(pv) << sh;
        }
        if (mo) reinterpret_cast<unsigned int*>(mo)[i] = m4;
    }
    for (int i = nvec * 4 + blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += stride) {
        const uint8_t pv = tally(pr[i], pg[i], pb[i], gtb[i], igb ? igb[i] : 0u);
        if (mo) mo[i] = pv;
    }
    // warp-aggregated counts -> smem atomics -> one global atomic per CTA per counter
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        const unsigned int s = __reduce_add_sync(0xffffffffu, c[k]);
        if ((threadIdx.x & 31) == 0 && s) atomicAdd(&s_cnt[k], s);
    }
    __syncthreads();
    if (threadIdx.x < 6 && s_cnt[threadIdx.x])
        atomicAdd(&ws[b].cnt[threadIdx.x], static_cast<unsigned long long>(s_cnt[threadIdx.x]));
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(&ws[b].ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (s_last && threadIdx.x < 2) {
        __threadfence();
        const int v = threadIdx.x;
        volatile unsigned long long* cnt = ws[b].cnt;
        const long long inter = static_cast<long long>(cnt[v]);
        const long long pa = static_cast<long long>(cnt[2 + v]);
        const long long ga = static_cast<long long>(cnt[4 + v]);
        area_inter[b * 2 + v] = inter;
        area_union[b * 2 + v] = pa + ga - inter;
    }
}

__global__ void iou_accumulate_kernel(const long long* __restrict__ inter, const long long* __restrict__ uni,
                                      const long long* __restrict__ class_id, long long* inter_buf,
                                      long long* union_buf, int B, int nclass) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * 2) return;
    const int b = i / 2, v = i % 2;
    const long long c = class_id[b];
    if (c < 0 || c >= nclass) {        // the reference's index_add_ (logger.py:36-37) raises a device-side assert here
        printf("[dfw] iou_accumulate: class_id[%d] = %lld is outside [0, %d)\n", b, c, nclass);
        __trap();
    }
    atomicAdd(reinterpret_cast<unsigned long long*>(inter_buf + static_cast<long long>(v) * nclass + c),
              static_cast<unsigned long long>(inter[b * 2 + v]));
    atomicAdd(reinterpret_cast<unsigned long long*>(union_buf + static_cast<long long>(v) * nclass + c),
              static_cast<unsigned long long>(uni[b * 2 + v]));
}

}  // namespace
}  // namespace dfw

extern "C" {

long long dfw_rthres_workspace_bytes(int B) { return B > 0 ? static_cast<long long>(B) * 64 : -1; }

int dfw_rthres_iou_hist(const uint8_t* pred_u8, int pred_is_mask, const uint8_t* gt, const uint8_t* ignore,
                        float r_threshold, long long* area_inter, long long* area_union, uint8_t* mask_out, int B,
                        int H, int W, void* workspace, void* stream_) {
    using namespace dfw;
    int rc = require_sm100();
    if (rc != DFW_OK) return rc;
    DFW_REQUIRE(pred_u8 && gt && area_inter && area_union && workspace && B > 0 && H > 0 && W > 0);
    const int HW = H * W;     // any size / alignment: the kernels take the vector paths only where the planes are aligned
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    EpisodeWs* ws = reinterpret_cast<EpisodeWs*>(workspace);
    DFW_CHECK_CUDA(cudaMemsetAsync(ws, 0, static_cast<size_t>(B) * sizeof(EpisodeWs), stream));
    int bx = (HW / 4 + 255) / 256;
    int cap = (4 * sm_count() + B - 1) / B;
    if (cap < 1) cap = 1;
    if (bx > cap) bx = cap;
    if (bx < 1) bx = 1;
    dim3 grid(bx, B);
    if (!pred_is_mask) rthres_max_kernel<<<grid, 256, 0, stream>>>(pred_u8, ws, 3LL * HW);
    rthres_hist_kernel<<<grid, 256, 0, stream>>>(pred_u8, pred_is_mask, gt, ignore, r_threshold, ws, area_inter,
                                                  area_union, mask_out, HW);
    g_launches.fetch_add(pred_is_mask ? 1 : 2);
    DFW_CHECK_CUDA(cudaGetLastError());
    return DFW_OK;
}

int dfw_iou_accumulate(const long long* area_inter, const long long* area_union, const long long* class_id,
                       long long* inter_buf, long long* union_buf, int B, int nclass, void* stream_) {
    using namespace dfw;
    int rc = require_sm100();
    if (rc != DFW_OK) return rc;
    DFW_REQUIRE(area_inter && area_union && class_id && inter_buf && union_buf && B > 0 && nclass > 0);
    iou_accumulate_kernel<<<(2 * B + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream_)>>>(
        area_inter, area_union, class_id, inter_buf, union_buf, B, nclass);
    g_launches.fetch_add(1);
    DFW_CHECK_CUDA(cudaGetLastError());
    return DFW_OK;
}

}  // extern "C"
