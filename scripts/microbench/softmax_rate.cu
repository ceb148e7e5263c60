// Micro-benchmark 3 (round 2): the softmax inner loop of the attention kernel in isolation (attn_softmax.cuh).
// 8 softmax warps (2 per sub-partition, like the kernel's two query tiles) loop over "key tiles": 128 fp32 logits per
// thread from TMEM -> row max -> p = 2^(s*c - m) with NPOLY/8 of the exponentials on the FMA pipe -> fp16 pairs ->
// tcgen05.st into the P columns.  Reported: cycles per tile pair (= both warps of a sub-partition one tile each); the
// tensor pipe needs 1024 cycles for the same pair, so tensor-pipe utilisation <= 1024 / that.
// Also checks the numerics of the polynomial path and the P layout against the host.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../diffews_b200/csrc -o softmax_rate softmax_rate.cu
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "attn_softmax.cuh"
using namespace dfw;

constexpr int THREADS = 384;

template <int NPOLY, bool SETMAXNREG>
__global__ void __launch_bounds__(THREADS, 1) softmax_kernel(const float* S_in, uint32_t* P_out, float* sum_out,
                                                               long long* cyc, int tiles, float sc) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 2) { tmem_alloc(smem_u32(&slot), 512); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = slot;
    if (warp >= 4) {
        if (SETMAXNREG) asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
        const int x = (warp - 4) >> 2, qd = warp & 3, row = qd * 32 + lane;
        const uint32_t lane_off = static_cast<uint32_t>(qd * 32) << 16;
        const uint32_t tS = tmem + lane_off + x * 256, tP = tS + 128;
        // seed S of this row from global memory
        {
            uint32_t v[32];
            for (int c = 0; c < 4; ++c) {
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(S_in[(x * 128 + row) * 128 + c * 32 + i]);
                tmem_st_32x32(tS + c * 32, v);
            }
            tmem_st_wait();
        }
        float m_used = -INFINITY, l_run = 0.f;
        long long c0 = 0;
        for (int j = 0; j < tiles + 1; ++j) {
            if (j == 1) c0 = clock64();            // tile 0 warms up
            uint32_t s[128];
            tmem_ld_row128(tS, s);
            tmem_ld_wait();
            const float mx = row_max128(s);
            const float m_new = fmaxf(m_used, mx * sc);
            if (__any_sync(0xffffffffu, m_new > m_used + 8.0f)) {
                l_run *= ex2_approx(m_used - m_new);
                m_used = m_new;
            }
            const uint64_t sc2 = f2_pack(sc, sc), nmu2 = f2_pack(-m_used, -m_used);
            uint64_t sum2[2] = {0ull, 0ull};
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                uint32_t pk[16];
                exp_chunk32<true, NPOLY>(&s[c * 32], sc2, nmu2, sum2, pk);
                tmem_st_32x16(tP + c * 16, pk);
            }
            float a, b, c, d;
            f2_unpack(sum2[0], a, b);
            f2_unpack(sum2[1], c, d);
            l_run += (a + b) + (c + d);
            tmem_st_wait();
        }
        const long long c1 = clock64();
        if (lane == 0) cyc[blockIdx.x * 8 + (warp - 4)] = c1 - c0;
        if (blockIdx.x == 0) {
            sum_out[x * 128 + row] = l_run / (tiles + 1);
            uint32_t v[32];
            for (int c = 0; c < 2; ++c) {
                tmem_ld_32x32(tP + c * 32, v);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) P_out[(x * 128 + row) * 64 + c * 32 + i] = v[i];
            }
        }
    } else if (SETMAXNREG) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

template <int NPOLY, bool SMR>
void run(const float* dS, const std::vector<float>& hS, float sc) {
    const int tiles = 200, blocks = 148;
    uint32_t* dP; float* dsum; long long* dc;
    cudaMalloc(&dP, 256 * 64 * 4); cudaMalloc(&dsum, 256 * 4); cudaMalloc(&dc, blocks * 8 * 8);
    softmax_kernel<NPOLY, SMR><<<blocks, THREADS>>>(dS, dP, dsum, dc, tiles, sc);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("NPOLY %d: %s\n", NPOLY, cudaGetErrorString(e)); exit(1); }
    std::vector<long long> hc(blocks * 8);
    std::vector<uint32_t> hP(256 * 64);
    std::vector<float> hsum(256);
    cudaMemcpy(hc.data(), dc, hc.size() * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(hP.data(), dP, hP.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(hsum.data(), dsum, hsum.size() * 4, cudaMemcpyDeviceToHost);
    double avg = 0;
    for (auto v : hc) avg += v;
    avg /= hc.size();
    const double per_tile = avg / tiles;
    // numerics: P and row sums vs the host
    double max_rel = 0, max_sum_rel = 0;
    for (int r = 0; r < 256; ++r) {
        float mx = -INFINITY;
        for (int k = 0; k < 128; ++k) mx = fmaxf(mx, hS[r * 128 + k]);
        const float mu = mx * sc;
        double sum = 0;
        for (int k = 0; k < 128; ++k) {
            const double ref = exp2(double(hS[r * 128 + k]) * sc - mu);
            const uint32_t w = hP[r * 64 + k / 2];
            const __half_raw hr{static_cast<unsigned short>((k & 1) ? (w >> 16) : (w & 0xffff))};
            const double got = __half2float(__half(hr));
            sum += ref;
            if (ref > 1e-3) max_rel = fmax(max_rel, fabs(got - ref) / ref);
        }
        max_sum_rel = fmax(max_sum_rel, fabs(hsum[r] - sum) / sum);
    }
    printf("NPOLY %d/8%s: %7.1f cycles per tile per warp (= per tile pair per sub-partition) -> tensor-pipe ceiling %5.1f %% ; "
           "P max rel err %.2e, row-sum rel err %.2e %s\n", NPOLY, SMR ? " setmaxnreg" : "", per_tile, 100.0 * 1024.0 / per_tile,
           max_rel, max_sum_rel, (max_rel < 2e-3 && max_sum_rel < 1e-4) ? "NUMERICS_OK" : "NUMERICS_BAD");
    cudaFree(dP); cudaFree(dsum); cudaFree(dc);
}

int main() {
    std::vector<float> hS(256 * 128);
    srand(3);
    for (auto& v : hS) v = ((rand() % 20001) - 10000) / 10000.0f * 24.0f;      // raw logits in [-24, 24]
    float* dS;
    cudaMalloc(&dS, hS.size() * 4);
    cudaMemcpy(dS, hS.data(), hS.size() * 4, cudaMemcpyHostToDevice);
    const float sc = 0.125f * 1.4426950408889634f;
    run<0, false>(dS, hS, sc);
    run<1, false>(dS, hS, sc);
    run<2, false>(dS, hS, sc);
    run<3, false>(dS, hS, sc);
    run<4, false>(dS, hS, sc);
    run<0, true>(dS, hS, sc);
    run<2, true>(dS, hS, sc);
    run<3, true>(dS, hS, sc);
    run<4, true>(dS, hS, sc);
    printf("status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
