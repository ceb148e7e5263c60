// Micro-benchmark: does a MUFU (ex2.approx) warp instruction block the sub-partition's issue port for the 8 cycles the
// quarter-rate XU datapath needs, or can FMA-pipe instructions of the same / another warp issue underneath it?
// (attention: the softmax of head-dim-64 attention is bound by this.)
//
// Each thread runs ITER iterations of: 8 independent ex2 chains, each followed by K independent FFMAs.
// Reported: SM cycles per warp-level MUFU instruction per sub-partition, for W warps per sub-partition.
//   blocking model:    8 + K (W = 1) -- FFMAs never hide
//   concurrent model:  max(8, 1 + K)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_issue mufu_issue.cu ; run: ./mufu_issue
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ float ex2(float x) {
    float y;
    asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <int K>
__global__ void __launch_bounds__(512) kern(float* out, long long* cyc, int iters, float a, float b) {
    float x[8], y[16];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = -1.0f - 0.01f * (threadIdx.x + i);
#pragma unroll
    for (int i = 0; i < 16; ++i) y[i] = 0.5f + 0.001f * (threadIdx.x + i);
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            x[i] = ex2(x[i]) - 1.5f;                      // MUFU (+ one FADD keeping the chain in range)
#pragma unroll
            for (int k = 0; k < K; ++k) {
                float& v = y[(i * K + k) & 15];
                asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(v) : "f"(a), "f"(b));
            }
        }
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i];
#pragma unroll
    for (int i = 0; i < 16; ++i) s += y[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int K>
void run(int warps_per_smsp, float* out, long long* cyc) {
    const int iters = 2000, blocks = 148, threads = warps_per_smsp * 4 * 32;
    kern<K><<<blocks, threads>>>(out, cyc, 10, 0.999f, 0.001f);
    cudaDeviceSynchronize();
    kern<K><<<blocks, threads>>>(out, cyc, iters, 0.999f, 0.001f);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < blocks; ++i) avg += h[i];
    avg /= blocks;
    // per sub-partition: warps_per_smsp warps x iters x 8 MUFU warp-instructions
    const double per_mufu = avg / (double(warps_per_smsp) * iters * 8);
    printf("K=%2d FFMA/MUFU  W=%d warps/SMSP : %.2f cycles per MUFU warp-instruction per SMSP  (blocking %d, concurrent %d)\n",
           K, warps_per_smsp, per_mufu, 8 + K + 1, (K + 2) > 8 ? K + 2 : 8);
}

int main() {
    float* out; long long* cyc;
    cudaMalloc(&out, 148 * 512 * sizeof(float));
    cudaMalloc(&cyc, 148 * sizeof(long long));
    for (int w : {1, 2, 4}) {
        run<0>(w, out, cyc); run<1>(w, out, cyc); run<2>(w, out, cyc); run<4>(w, out, cyc);
        run<6>(w, out, cyc); run<8>(w, out, cyc); run<12>(w, out, cyc);
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
