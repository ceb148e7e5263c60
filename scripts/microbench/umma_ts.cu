// Micro-benchmark 2 (round 2): tcgen05.mma issue / execution rate with HOISTED descriptors, and with the A operand in
// TMEM (".ts": P of the PV product never goes through shared memory), plus a numerical check of the TMEM A layout.
//
//   part 1  rate, one or two issuing warps, descriptors precomputed outside the issue loop:
//             SS  M128 N128 K16 (S = Q K^T), SS M128 N64 K16 MN-major B (O += P V), TS M128 N64 K16 MN-major B,
//             and the attention mix of one key tile: 4 x SS N128 + 8 x TS N64 per issuer
//   part 2  TS numerics: A[128 x 128] fp16 written to TMEM with tcgen05.st (thread = row, column c = elements 2c, 2c+1),
//           B = V[128 keys x 64] MN-major in 128B-swizzled smem, D = A B read back and compared with the host product
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../diffews_b200/csrc -o umma_ts umma_ts.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_fp16.h>

#include "ptx.cuh"
using namespace dfw;

__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "}\n"
        ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(acc)
        : "memory");
}

// mode 0: SS N128 K-major B   1: SS N64 MN-major B   2: TS N64 MN-major B   3: attention mix (4 SS N128 + 8 TS N64)
template <int MODE>
__global__ void __launch_bounds__(128, 1) rate_kernel(long long* out, int reps, int issuers) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t sA = base, sB = base + 64 * 1024, bar = base + 160 * 1024, slot = bar + 32;
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
    if (warp < issuers) {
        const uint32_t idesc_s = umma_idesc(128, 128, 0, 0, 0);
        const uint32_t idesc_o = umma_idesc(128, 64, 0, 0, 1);
        // TMEM plan of one issuer (256 columns): S [0,128)  P [128,192)  O [192,256)
        const uint32_t t0c = tmem + warp * 256;
        uint64_t a_s[4], b_s[4], a_p[8], b_v[8];
#pragma unroll
        for (int k = 0; k < 4; ++k) { a_s[k] = umma_desc_sw128(sA) + 2u * k; b_s[k] = umma_desc_sw128(sB) + 2u * k; }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            a_p[k] = umma_desc_sw128(sA + 16384 + (k >> 2) * 16384) + 2u * (k & 3);
            b_v[k] = umma_desc_sw128(sB + 16384 + k * 16 * 128);
        }
        long long c0 = 0, c1 = 0;
        for (int round = 0; round < 2; ++round) {
            c0 = clock64();
            if (elect_one()) {
                for (int r = 0; r < reps; ++r) {
                    if (MODE == 0 || MODE == 3) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) umma_ss(t0c, a_s[k], b_s[k], idesc_s, k > 0 ? 1u : 0u);
                    }
                    if (MODE == 1) {
#pragma unroll
                        for (int k = 0; k < 8; ++k) umma_ss(t0c + 192, a_p[k], b_v[k], idesc_o, 1u);
                    }
                    if (MODE == 2 || MODE == 3) {
#pragma unroll
                        for (int k = 0; k < 8; ++k) umma_ts(t0c + 192, t0c + 128 + k * 8, b_v[k], idesc_o, 1u);
                    }
                }
                tc_commit(bar + 8 * warp);
            }
            __syncwarp();
            mbar_wait(bar + 8 * warp, round & 1, 1);
            tc_fence_after();
            c1 = clock64();
        }
        if (lane == 0 && warp == 0) out[blockIdx.x] = c1 - c0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 3) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

template <int MODE>
void run_rate(const char* name, int per_rep, double ideal_per_rep, long long* out) {
    const int reps = 256, blocks = 148, smem = 164 * 1024;
    cudaFuncSetAttribute(rate_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int issuers : {1, 2}) {
        rate_kernel<MODE><<<blocks, 128, smem>>>(out, reps, issuers);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); exit(1); }
        long long h[148];
        cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
        double avg = 0;
        for (int i = 0; i < blocks; ++i) avg += h[i];
        avg /= blocks;
        const double per_rep_cyc = avg / reps;             // wall cycles per repetition (both issuers run concurrently)
        printf("%-44s %d issuer(s): %7.1f cycles / rep / issuer-set = %6.1f cycles / MMA ; tensor-pipe work %6.1f -> %5.1f %%\n",
               name, issuers, per_rep_cyc, per_rep_cyc / (per_rep * issuers), ideal_per_rep * issuers,
               100.0 * ideal_per_rep * issuers / per_rep_cyc);
    }
}

// ---- TS numerics -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1) ts_check_kernel(const __half* P, const __half* V, float* O) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t sV = base, bar = base + 32 * 1024, slot = bar + 16;
    volatile uint32_t* slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (slot - smem_u32(smem_raw)));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // V [128 keys][64 d] -> the TMA SWIZZLE_128B image: row = key (128 B), 16-byte unit u stored at u ^ (key & 7)
    for (int i = threadIdx.x; i < 128 * 8; i += blockDim.x) {
        const int key = i >> 3, u = i & 7;
        *reinterpret_cast<uint4*>(gen + key * 128 + ((u ^ (key & 7)) * 16)) = *reinterpret_cast<const uint4*>(V + key * 64 + u * 8);
    }
    if (warp == 0 && lane == 0) { mbar_init(bar, 1); fence_mbar_init(); }
    if (warp == 3) { tmem_alloc(slot, 256); tmem_relinquish(); }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *slot_ptr;
    const uint32_t lane_off = static_cast<uint32_t>(warp * 32) << 16;
    // thread = row; P row -> 64 packed columns at TMEM columns [128,192)
    const int row = warp * 32 + lane;
    uint32_t v[32];
    for (int c = 0; c < 2; ++c) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = reinterpret_cast<const uint32_t*>(P + row * 128)[c * 32 + i];
        tmem_st_32x32(tmem + lane_off + 128 + c * 32, v);
    }
    tmem_st_wait();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 0) {
        if (elect_one()) {
            const uint32_t idesc_o = umma_idesc(128, 64, 0, 0, 1);
#pragma unroll
            for (int k = 0; k < 8; ++k)
                umma_ts(tmem + 192, tmem + 128 + k * 8, umma_desc_sw128(sV + k * 16 * 128), idesc_o, k > 0 ? 1u : 0u);
            tc_commit(bar);
        }
        __syncwarp();
    }
    mbar_wait(bar, 0, 2);
    tc_fence_after();
    for (int c = 0; c < 2; ++c) {
        tmem_ld_32x32(tmem + lane_off + 192 + c * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) O[row * 64 + c * 32 + i] = __uint_as_float(v[i]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 3) { tc_fence_after(); tmem_dealloc(tmem, 256); }
}

int main() {
    long long* out;
    cudaMalloc(&out, 148 * sizeof(long long));
    run_rate<0>("SS M128 N128 K16 x4 (S = Q K^T), hoisted desc", 4, 4 * 64.0, out);
    run_rate<1>("SS M128 N64 K16 x8 (O += P V, P in smem)", 8, 8 * 32.0, out);
    run_rate<2>("TS M128 N64 K16 x8 (O += P V, P in TMEM)", 8, 8 * 32.0, out);
    run_rate<3>("mix: 4 SS N128 + 8 TS N64 (one key tile)", 12, 512.0, out);

    // numerics of the TMEM A layout
    std::vector<__half> hP(128 * 128), hV(128 * 64);
    srand(1);
    for (auto& x : hP) x = __float2half((rand() % 2001 - 1000) / 1000.0f);
    for (auto& x : hV) x = __float2half((rand() % 2001 - 1000) / 1000.0f);
    __half *dP, *dV; float* dO;
    cudaMalloc(&dP, hP.size() * 2); cudaMalloc(&dV, hV.size() * 2); cudaMalloc(&dO, 128 * 64 * 4);
    cudaMemcpy(dP, hP.data(), hP.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dV, hV.data(), hV.size() * 2, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(ts_check_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 40 * 1024);
    ts_check_kernel<<<1, 128, 40 * 1024>>>(dP, dV, dO);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("ts_check: %s\n", cudaGetErrorString(e)); return 1; }
    std::vector<float> hO(128 * 64);
    cudaMemcpy(hO.data(), dO, hO.size() * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0, maxref = 0;
    for (int r = 0; r < 128; ++r)
        for (int d = 0; d < 64; ++d) {
            double acc = 0;
            for (int k = 0; k < 128; ++k) acc += double(__half2float(hP[r * 128 + k])) * double(__half2float(hV[k * 64 + d]));
            maxerr = fmax(maxerr, fabs(acc - hO[r * 64 + d]));
            maxref = fmax(maxref, fabs(acc));
        }
    printf("TS numerics: max |O - ref| = %.3e (max |ref| %.2f) -> %s\n", maxerr, maxref, maxerr < 1e-3 * maxref ? "TS_LAYOUT_OK" : "TS_LAYOUT_MISMATCH");
    printf("status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
