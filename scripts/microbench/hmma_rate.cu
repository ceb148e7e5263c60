// Micro-benchmark 5 (round 2): rate of the legacy warp-level mma.sync.m16n8k16 (fp16 in, fp32 accumulate) on sm_100a --
// the candidate for the N = 3 decoder-head convolution, where a tcgen05 tile would be > 80 % padding.
// Reported: SM cycles per mma.sync per sub-partition and the implied dense TFLOP/s at 1.9 GHz on 148 SMs.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o hmma_rate hmma_rate.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(512) kern(float* out, long long* cyc, int iters) {
    uint32_t a[4] = {0x3c003c00u + threadIdx.x, 0x3c003c00u, 0x38003800u, 0x3c003c00u};
    uint32_t b[2] = {0x3c003c00u, 0x34003400u + threadIdx.x};
    float c[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.f;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                         : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
    float* out; long long* cyc;
    cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 148 * 8);
    for (int w : {1, 2, 4}) {
        const int iters = 2000, threads = w * 128;
        kern<<<148, threads>>>(out, cyc, 10);
        kern<<<148, threads>>>(out, cyc, iters);
        cudaDeviceSynchronize();
        long long hc[148];
        cudaMemcpy(hc, cyc, sizeof(hc), cudaMemcpyDeviceToHost);
        double avg = 0;
        for (int i = 0; i < 148; ++i) avg += hc[i];
        avg /= 148;
        const double per = avg / (double(w) * iters * 8);
        printf("mma.sync m16n8k16 f16->f32, W=%d warps/SMSP: %.2f cycles per mma per SMSP = %.0f dense TFLOP/s at 1.9 GHz x 148 SMs\n", w, per,
               4096.0 / per * 4 * 148 * 1.9e9 / 1e12);
    }
    printf("status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
