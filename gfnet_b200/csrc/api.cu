// C-ABI housekeeping: version, error strings, device query, TMA descriptor encoding.
#include "common.cuh"
#include <cudaTypedefs.h>
#include <mutex>

extern "C" int gfb_abi_version(void) { return GFB_ABI_VERSION; }

extern "C" const char* gfb_strerror(int code) {
    switch (code) {
        case GFB_OK: return "ok";
        case GFB_EINVAL: return "gfnet_b200: invalid argument (null pointer, bad shape or enum)";
        case GFB_EUNSUPPORTED: return "gfnet_b200: configuration not supported by the requested kernel";
        case GFB_EALIGN: return "gfnet_b200: pointer not aligned as required";
        case GFB_EWORKSPACE: return "gfnet_b200: workspace too small";
        case GFB_ENODEVICE: return "gfnet_b200: no sm_100 device or CUDA driver entry point unavailable";
        default: break;
    }
    if (code > 0) return cudaGetErrorString((cudaError_t)code);
    return "gfnet_b200: unknown error";
}

extern "C" int gfb_device_info(int* sm_count, int* cc_major, int* cc_minor) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, dev);
    if (e != cudaSuccess) return (int)e;
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    return GFB_OK;
}

static PFN_cuTensorMapEncodeTiled_v12000 g_encode = nullptr;
static std::once_flag g_encode_once;

int gfb_encode_tmap_f32(CUtensorMap* map, const void* base, int rank, const uint64_t* dims,
                        const uint64_t* strides_bytes, const uint32_t* box, int swizzle) {
    std::call_once(g_encode_once, [] {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPointByVersion("cuTensorMapEncodeTiled", &fn, 12000, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            g_encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
    });
    if (!g_encode) return GFB_ENODEVICE;
    cuuint64_t gd[5], gs[4];
    cuuint32_t bx[5], es[5];
    for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
    for (int i = 0; i + 1 < rank; ++i) gs[i] = strides_bytes[i];
    CUtensorMapSwizzle sw = swizzle == 3 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE;
    CUresult r = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<void*>(base), gd, gs, bx, es,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? GFB_OK : GFB_EINVAL;
}
