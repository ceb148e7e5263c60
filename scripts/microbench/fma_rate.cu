// Micro-benchmark 4 (round 2): issue rate of FFMA / FFMA2 (packed fp32x2) / HFMA2 / FMNMX / F2FP per sub-partition, W warps each.
// Reported: SM cycles per warp-instruction per sub-partition (1.0 = one instruction per clock).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o fma_rate fma_rate.cu
#include <cstdint>
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

template <int OP>
__global__ void __launch_bounds__(512) kern(float* out, long long* cyc, int iters, float a, float b) {
    float r[16];
    unsigned long long q[8];
    uint32_t h[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { r[i] = 0.5f + 0.001f * (threadIdx.x + i); h[i] = 0x3c003c00u + threadIdx.x + i; }
#pragma unroll
    for (int i = 0; i < 8; ++i) q[i] = (static_cast<unsigned long long>(__float_as_uint(r[2 * i + 1])) << 32) | __float_as_uint(r[2 * i]);
    const unsigned long long a2 = (static_cast<unsigned long long>(__float_as_uint(a)) << 32) | __float_as_uint(a);
    const unsigned long long b2 = (static_cast<unsigned long long>(__float_as_uint(b)) << 32) | __float_as_uint(b);
    const uint32_t ha = 0x3bff3bffu, hb = 0x00010001u;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (OP == 0) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(r[i]) : "f"(a), "f"(b));
            if (OP == 1) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(q[i & 7]) : "l"(a2), "l"(b2));
            if (OP == 2) asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(h[i]) : "r"(ha), "r"(hb));
            if (OP == 3) asm volatile("max.f32 %0, %0, %1;" : "+f"(r[i]) : "f"(a));
            if (OP == 4) asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h[i]) : "f"(r[i]), "f"(r[(i + 1) & 15]));
            if (OP == 5) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(r[i]) : "f"(r[(i + 5) & 15]), "f"(r[(i + 9) & 15]));   // 3 distinct registers
            if (OP == 6) {      // FFMA2 and FMNMX alternating (two pipes)
                if (i & 1) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(q[i & 7]) : "l"(a2), "l"(b2));
                else asm volatile("max.f32 %0, %0, %1;" : "+f"(r[i]) : "f"(a));
            }
            if (OP == 7) {      // HFMA2 and FMNMX alternating
                if (i & 1) asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(h[i]) : "r"(ha), "r"(hb));
                else asm volatile("max.f32 %0, %0, %1;" : "+f"(r[i]) : "f"(a));
            }
        }
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += r[i] + __uint_as_float(h[i]);
#pragma unroll
    for (int i = 0; i < 8; ++i) s += __uint_as_float(static_cast<uint32_t>(q[i])) + __uint_as_float(static_cast<uint32_t>(q[i] >> 32));
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char* name, float* out, long long* cyc) {
    for (int w : {1, 2, 4}) {
        const int iters = 2000, blocks = 148, threads = w * 128;
        kern<OP><<<blocks, threads>>>(out, cyc, 10, 0.999f, 0.001f);
        kern<OP><<<blocks, threads>>>(out, cyc, iters, 0.999f, 0.001f);
        cudaDeviceSynchronize();
        long long hc[148];
        cudaMemcpy(hc, cyc, sizeof(hc), cudaMemcpyDeviceToHost);
        double avg = 0;
        for (int i = 0; i < blocks; ++i) avg += hc[i];
        avg /= blocks;
        printf("%-28s W=%d warps/SMSP: %.2f cycles per warp-instruction per SMSP\n", name, w, avg / (double(w) * iters * 16));
    }
}

int main() {
    float* out; long long* cyc;
    cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 148 * 8);
    run<0>("FFMA (2 reg + same)", out, cyc);
    run<5>("FFMA (3 distinct regs)", out, cyc);
    run<1>("FFMA2 (fp32x2)", out, cyc);
    run<2>("HFMA2", out, cyc);
    run<3>("FMNMX", out, cyc);
    run<4>("F2FP (cvt.f16x2.f32)", out, cyc);
    run<6>("FFMA2 / FMNMX alternating", out, cyc);
    run<7>("HFMA2 / FMNMX alternating", out, cyc);
    printf("status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
