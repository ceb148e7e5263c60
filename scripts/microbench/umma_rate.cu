// Micro-benchmark: sustained rate of tcgen05.mma (kind::f16, cta_group::1, fp32 accumulate) per instruction shape, as the
// KV-fused attention kernel issues them: one elected thread issues R back-to-back MMAs on operands already in shared
// memory (contents irrelevant), commits to an mbarrier and waits.  Reported: SM cycles per MMA instruction and the
// fraction of the dense peak (4096 MAC / clk / SM).
//   S = Q K^T : M128 N128 K16, B K-major          O += P V : M128 N64 K16, B MN-major (V straight from the TMA tile)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../diffews_b200/csrc -o umma_rate umma_rate.cu
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>

#include "ptx.cuh"
using namespace dfw;

template <int M, int N, int BMN>
__global__ void __launch_bounds__(128, 1) kern(long long* out, int reps, int distinct, int issuers) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t sA = base, sB = base + 64 * 1024, bar = base + 160 * 1024, slot = bar + 16;
    volatile uint32_t* slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (slot - smem_u32(smem_raw)));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 40 * 1024; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw)[i] = 0x3c003c00u;
    if (warp == 0 && lane == 0) { mbar_init(bar, 1); mbar_init(bar + 8, 1); fence_mbar_init(); }
    if (warp == 3) { tmem_alloc(slot, 512); tmem_relinquish(); }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *slot_ptr;
    if (warp < issuers) {                                        // warp w issues into TMEM columns [256w, 256w+256), own mbarrier
        const uint32_t idesc = umma_idesc(M, N, 0, 0, BMN);
        long long t0 = 0, t1 = 0;
        for (int round = 0; round < 2; ++round) {               // round 0 warms up
            t0 = clock64();
            if (elect_one()) {
                for (int r = 0; r < reps; ++r) {
                    const int k = r % distinct;                  // walk over `distinct` K-slices like the real kernels do
                    const uint64_t adesc = umma_desc_sw128(sA + (k >> 2) * 16384) + 2u * (k & 3);
                    const uint64_t bdesc = BMN ? umma_desc_sw128(sB + k * 16 * 128) : (umma_desc_sw128(sB + (k >> 2) * 16384) + 2u * (k & 3));
                    umma_ss(tmem + (issuers == 1 ? (r & 1) * 256 : warp * 256), adesc, bdesc, idesc, r >= 2 ? 1u : 0u);
                }
                tc_commit(bar + 8 * warp);
            }
            __syncwarp();
            mbar_wait(bar + 8 * warp, round & 1, 1);
            tc_fence_after();
            t1 = clock64();
        }
        if (lane == 0 && warp == 0) out[blockIdx.x] = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 3) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

template <int M, int N, int BMN>
void run(const char* name, long long* out) {
    const int reps = 512, blocks = 148, smem = 164 * 1024;
    cudaFuncSetAttribute(kern<M, N, BMN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int issuers : {1, 2}) {
        if (issuers == 2 && N > 128) continue;                   // two accumulators of N columns each must fit 256-column halves
        const int distinct = 8;
        kern<M, N, BMN><<<blocks, 128, smem>>>(out, reps, distinct, issuers);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
        long long h[148];
        cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
        double avg = 0;
        for (int i = 0; i < blocks; ++i) avg += h[i];
        avg /= blocks;
        const double per = avg / (double(reps) * issuers), ideal = double(M) * N * 16 / 4096.0;
        printf("%-34s %d issuing warp(s): %7.1f cycles / MMA (ideal %5.1f) = %5.1f %% of dense peak\n", name, issuers, per, ideal,
               100.0 * ideal / per);
    }
}

int main() {
    long long* out;
    cudaMalloc(&out, 148 * sizeof(long long));
    run<128, 128, 0>("M128 N128 K16  B K-major  (S=QK^T)", out);
    run<128, 64, 1>("M128 N64  K16  B MN-major (O+=PV)", out);
    run<128, 64, 0>("M128 N64  K16  B K-major", out);
    run<128, 128, 1>("M128 N128 K16  B MN-major", out);
    run<128, 256, 0>("M128 N256 K16  B K-major", out);
    printf("status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
