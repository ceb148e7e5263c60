#include "common.cuh"

#include <atomic>
#include <cstdlib>
#include <mutex>

namespace dfw {

std::atomic<long long> g_launches{0};

static std::atomic<int> g_options[DFW_OPT_COUNT] = {
    {0},        // DFW_OPT_PDL
    {1},        // DFW_OPT_T128
    {1 << 20},  // DFW_OPT_T128_MAXC
    {1},        // DFW_OPT_HALO
    {3},        // DFW_OPT_GN_CTAS_PER_SM
    {0},        // DFW_OPT_PREPROC_TWO_PASS
    {0},        // DFW_OPT_ATTN_V2
    {1},        // DFW_OPT_SEG_HEAD
    {0},        // DFW_OPT_ATTN_BWD_UNFUSED
    {0},        // DFW_OPT_ATTN_V4
    {0},        // DFW_OPT_B_RESIDENT
};

int get_option(int option) {
    return (option >= 0 && option < DFW_OPT_COUNT) ? g_options[option].load(std::memory_order_relaxed) : -1;
}

bool pdl_enabled() {
    // off by default: correct (bring-up + parity suites pass under it, CUDA-graph capture keeps the edges) but neutral on
    // B200 -- the step is power-capped, so hiding the ~2 us launch gaps buys no time (121.1 / 119.5 vs 120.7 / 121.4 eps/s)
    return get_option(DFW_OPT_PDL) != 0;
}

int require_sm100() {
    static int cached = 1;  // 1 = unknown
    if (cached != 1) return cached;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return DFW_ERR_CUDA;
    int major = 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return DFW_ERR_CUDA;
    cached = (major == 10) ? DFW_OK : DFW_ERR_ARCH;
    if (cached != DFW_OK) fprintf(stderr, "[dfw] device compute capability %d.x is not sm_100: no fallback path\n", major);
    return cached;
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
    static PFN_encodeTiled fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
        if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = reinterpret_cast<PFN_encodeTiled>(p);
    });
    return fn;
}

int encode_tmap_bf16_sw128(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                           const uint64_t* strides_bytes, const uint32_t* box) {
    return encode_tmap(out, base, 2, 128, rank, dims, strides_bytes, box);
}

int encode_tmap(CUtensorMap* out, const void* base, int elem_bytes, int swizzle_bytes, int rank, const uint64_t* dims,
                const uint64_t* strides_bytes, const uint32_t* box) {
    PFN_encodeTiled enc = get_encode();
    if (!enc) {
        fprintf(stderr, "[dfw] cuTensorMapEncodeTiled entry point unavailable\n");
        return DFW_ERR_CUDA;
    }
    cuuint64_t gdim[5];
    cuuint64_t gstr[4];
    cuuint32_t bdim[5];
    cuuint32_t estr[5];
    for (int i = 0; i < rank; ++i) {
        gdim[i] = dims[i];
        bdim[i] = box[i];
        estr[i] = 1;
        if (i > 0) gstr[i - 1] = strides_bytes[i - 1];
    }
    const CUtensorMapDataType dt = (elem_bytes == 4) ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_UINT16;
    const CUtensorMapSwizzle sw = (swizzle_bytes == 128) ? CU_TENSOR_MAP_SWIZZLE_128B
                                  : (swizzle_bytes == 64) ? CU_TENSOR_MAP_SWIZZLE_64B
                                  : (swizzle_bytes == 32) ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE;
    CUresult r = enc(out, dt, static_cast<cuuint32_t>(rank), const_cast<void*>(base),
                     gdim, gstr, bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        fprintf(stderr, "[dfw] cuTensorMapEncodeTiled failed (%d): rank %d dims", static_cast<int>(r), rank);
        for (int i = 0; i < rank; ++i) fprintf(stderr, " %llu", static_cast<unsigned long long>(dims[i]));
        fprintf(stderr, " strides");
        for (int i = 0; i + 1 < rank; ++i) fprintf(stderr, " %llu", static_cast<unsigned long long>(strides_bytes[i]));
        fprintf(stderr, " box");
        for (int i = 0; i < rank; ++i) fprintf(stderr, " %u", box[i]);
        fprintf(stderr, " base %p\n", base);
        return DFW_ERR_CUDA;
    }
    return DFW_OK;
}

}  // namespace dfw

extern "C" {
int dfw_version(void) { return 2; }
int dfw_set_option(int option, int value) {
    if (option < 0 || option >= DFW_OPT_COUNT) return DFW_ERR_INVALID;
    dfw::g_options[option].store(value, std::memory_order_relaxed);
    return DFW_OK;
}
int dfw_get_option(int option) { return dfw::get_option(option); }
int dfw_device_ok(void) { return dfw::require_sm100(); }
long long dfw_launch_count(void) { return dfw::g_launches.load(); }
}
