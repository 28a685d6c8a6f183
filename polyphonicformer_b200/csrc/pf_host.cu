// Error reporting, device checks, TMA descriptor encoding.
#include <stdarg.h>
#include <string.h>

#include "pf_internal.h"
#include <cstdlib>

namespace pf {

static thread_local char g_err[512] = "";
static thread_local int g_launches = 0;

int set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

void count_launch(int n) { g_launches += n; }
void reset_launch_count() { g_launches = 0; }

int check_device() {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return set_error(PF_ERR_ARCH, "no CUDA device: %s", cudaGetErrorString(e));
    static thread_local int cached_dev = -1, cached_major = 0;
    if (cached_dev != dev) {
        int major = 0;
        e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
        if (e != cudaSuccess) return set_error(PF_ERR_ARCH, "cudaDeviceGetAttribute: %s", cudaGetErrorString(e));
        cached_dev = dev;
        cached_major = major;
    }
    if (cached_major != 10)
        return set_error(PF_ERR_ARCH, "device %d is sm_%dx; libpf_decoder is built for sm_100a only", dev, cached_major);
    return PF_OK;
}

int num_sms() {
    int dev = 0, n = 0;
    cudaGetDevice(&dev);
    static thread_local int cached_dev = -1, cached_n = 0;
    if (cached_dev == dev) return cached_n;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    // experiment knob ("SM-partitioned batch windows"): size the persistent streaming kernels for
    // fewer SMs so that another window's small-N block can run beside them
    if (const char* lim = getenv("PF_SM_LIMIT")) {
        const int v = atoi(lim);
        if (v >= 8 && v < n) n = v;
    }
    cached_dev = dev;
    cached_n = n;
    return n;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn) return fn;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !p) return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(p);
    return fn;
}

int make_tmap_bf16_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t pitch_elems,
                      uint32_t box_rows, uint32_t box_cols) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return set_error(PF_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) return set_error(PF_ERR_ALIGN, "TMA base not 16-byte aligned");
    if ((pitch_elems * 2) % 16 != 0) return set_error(PF_ERR_ALIGN, "TMA row pitch %llu elems not a 16-byte multiple",
                                                       (unsigned long long)pitch_elems);
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {pitch_elems * 2};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(PF_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return PF_OK;
}

int make_tmap_bf16_3d(CUtensorMap* out, const void* base, uint64_t d2, uint64_t d1, uint64_t d0, uint32_t box1,
                      uint32_t box0) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return set_error(PF_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) return set_error(PF_ERR_ALIGN, "TMA base not 16-byte aligned");
    if ((d0 * 2) % 16 != 0) return set_error(PF_ERR_ALIGN, "TMA row pitch not a 16-byte multiple");
    cuuint64_t dims[3] = {d0, d1, d2};
    cuuint64_t strides[2] = {d0 * 2, d0 * d1 * 2};
    cuuint32_t box[3] = {box0, box1, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(PF_ERR_CUDA, "cuTensorMapEncodeTiled(3d) failed with CUresult %d", (int)r);
    return PF_OK;
}

int make_tmap_bf16_blocked(CUtensorMap* out, const void* base, uint64_t blocks) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return set_error(PF_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) return set_error(PF_ERR_ALIGN, "TMA base not 16-byte aligned");
    cuuint64_t dims[3] = {8, 128, blocks};
    cuuint64_t strides[2] = {16, 2048};
    cuuint32_t box[3] = {8, 128, 8};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(PF_ERR_CUDA, "cuTensorMapEncodeTiled(blocked) failed with CUresult %d", (int)r);
    return PF_OK;
}

int make_tmap_f32_3d(CUtensorMap* out, const void* base, uint64_t d2, uint64_t d1, uint64_t d0, uint32_t box1,
                     uint32_t box0) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return set_error(PF_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) return set_error(PF_ERR_ALIGN, "TMA base not 16-byte aligned");
    if ((d0 * 4) % 16 != 0) return set_error(PF_ERR_ALIGN, "TMA row pitch not a 16-byte multiple");
    cuuint64_t dims[3] = {d0, d1, d2};
    cuuint64_t strides[2] = {d0 * 4, d0 * d1 * 4};
    cuuint32_t box[3] = {box0, box1, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(PF_ERR_CUDA, "cuTensorMapEncodeTiled(f32 3d) failed with CUresult %d", (int)r);
    return PF_OK;
}


int make_tmap_bf16_nd(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                      const uint32_t* box) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return set_error(PF_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    if (rank < 2 || rank > 5) return set_error(PF_ERR_ARG, "TMA rank %d", rank);
    if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) return set_error(PF_ERR_ALIGN, "TMA base not 16-byte aligned");
    cuuint64_t d[5], st[4];
    cuuint32_t b[5], estr[5];
    for (int i = 0; i < rank; ++i) d[i] = dims[i], b[i] = box[i], estr[i] = 1;
    for (int i = 0; i + 1 < rank; ++i) {
        if (strides_bytes[i] % 16 != 0) return set_error(PF_ERR_ALIGN, "TMA stride %d not a 16-byte multiple", i);
        st[i] = strides_bytes[i];
    }
    if (box[0] * 2 != 128) return set_error(PF_ERR_ARG, "TMA box inner extent must be 128 bytes for the 128-byte swizzle");
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), d, st, b, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(PF_ERR_CUDA, "cuTensorMapEncodeTiled(nd) failed with CUresult %d", (int)r);
    return PF_OK;
}

}  // namespace pf

extern "C" int pf_version(void) { return 100; }
extern "C" const char* pf_last_error_string(void) { return pf::g_err; }
extern "C" int pf_last_launch_count(void) { return pf::g_launches; }
