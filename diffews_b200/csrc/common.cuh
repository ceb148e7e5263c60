// Shared host-side helpers: status codes, launch checks, TMA tensor-map encoding through the driver entry point
// (the library must dlopen on a machine without libcuda.so, so nothing links against -lcuda).
#pragma once
#include <cstdint>
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>

#include "../../include/diffews_b200.h"

namespace dfw {

#define DFW_CHECK_CUDA(expr)                                                                          \
    do {                                                                                              \
        cudaError_t _e = (expr);                                                                      \
        if (_e != cudaSuccess) {                                                                      \
            fprintf(stderr, "[dfw] %s failed: %s (%s:%d)\n", #expr, cudaGetErrorString(_e), __FILE__, \
                    __LINE__);                                                                        \
            return DFW_ERR_CUDA;                                                                      \
        }                                                                                             \
    } while (0)

#define DFW_REQUIRE(cond)                                                                         \
    do {                                                                                          \
        if (!(cond)) {                                                                            \
            fprintf(stderr, "[dfw] invalid argument: %s (%s:%d)\n", #cond, __FILE__, __LINE__);   \
            return DFW_ERR_INVALID;                                                               \
        }                                                                                         \
    } while (0)

// Returns DFW_OK if the current device is sm_100 (B200); the library has no other code path.
int require_sm100();

// Encode a bf16 tensor map with SWIZZLE_128B (inner box extent must be 64 elements = 128 bytes)
// dims[0] is the contiguous dimension. strides_bytes[i] is the stride of dims[i+1].
int encode_tmap_bf16_sw128(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                           const uint64_t* strides_bytes, const uint32_t* box);
// General form: elem_bytes 2|4, swizzle_bytes 0|32|64|128 (inner box extent * elem_bytes must equal the swizzle span
// when swizzling).  Used for the epilogue's TMA stores / residual loads.
int encode_tmap(CUtensorMap* out, const void* base, int elem_bytes, int swizzle_bytes, int rank, const uint64_t* dims,
                const uint64_t* strides_bytes, const uint32_t* box);

// Process-wide option table (include/diffews_b200.h DFW_OPT_*): explicit, no environment variables.
int get_option(int option);

// Kernel launch with the programmatic-dependent-launch attribute when DFW_OPT_PDL is set (default: plain launch).  Only for kernels that call
// pdl_wait() before their first global access.
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

inline int sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    }
    return n;
}

}  // namespace dfw
