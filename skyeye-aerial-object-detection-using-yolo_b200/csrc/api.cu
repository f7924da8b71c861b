// Library plumbing: version, thread-local error string, device check, TMA tensor-map encoding.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>

#include "common.cuh"

namespace skb {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

constexpr int MAX_DEV = 64;
static int g_dev_checked[MAX_DEV];  // per device ordinal: 0 unknown, 1 ok
static int g_num_sms[MAX_DEV];
static std::mutex g_dev_mu;

bool PerDeviceOnce::first() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= MAX_DEV) return true;  // cannot tell: redo the (idempotent) action
    std::lock_guard<std::mutex> lk(g_dev_mu);
    const unsigned long long bit = 1ULL << dev;
    if (seen & bit) return false;
    seen |= bit;
    return true;
}
void PerDeviceOnce::reset_current() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= MAX_DEV) return;
    std::lock_guard<std::mutex> lk(g_dev_mu);
    seen &= ~(1ULL << dev);
}

int check_device() {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) {
        set_error("no CUDA device: %s", cudaGetErrorString(e));
        return SKB_ERR_CUDA;
    }
    if (dev >= 0 && dev < MAX_DEV && g_dev_checked[dev] == 1) return SKB_OK;
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, dev);
    if (e != cudaSuccess) {
        set_error("cudaGetDeviceProperties: %s", cudaGetErrorString(e));
        return SKB_ERR_CUDA;
    }
    if (prop.major != 10) {
        set_error("libskyeye_b200 needs an sm_100 (B200) device, found sm_%d%d (%s); there is no fallback path",
                  prop.major, prop.minor, prop.name);
        return SKB_ERR_ARCH;
    }
    if (dev >= 0 && dev < MAX_DEV) {
        std::lock_guard<std::mutex> lk(g_dev_mu);
        g_num_sms[dev] = prop.multiProcessorCount;
        g_dev_checked[dev] = 1;
    }
    return SKB_OK;
}

int num_sms() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= MAX_DEV) return 148;
    return g_num_sms[dev] > 0 ? g_num_sms[dev] : 148;
}

int pdl_enabled() {  // tuning knob (not part of the ABI): SKB_PDL=0 launches every kernel fully serialised
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("SKB_PDL");
        v = e ? (atoi(e) != 0) : 1;
    }
    return v;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;
static std::mutex g_mu;

static int resolve_encode() {
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_encode) return SKB_OK;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !fn) {
        set_error("cuTensorMapEncodeTiled not available from the driver (%s)", cudaGetErrorString(e));
        return SKB_ERR_CUDA;
    }
    g_encode = (EncodeTiledFn)fn;
    return SKB_OK;
}

int encode_tensor_map(CUtensorMap* out, const void* base, int elem_bytes, int rank, const uint64_t* dims,
                      const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes) {
    int rc = resolve_encode();
    if (rc != SKB_OK) return rc;
    cuuint64_t gdims[5], gstr[4];
    cuuint32_t gbox[5], estr[5];
    for (int i = 0; i < rank; ++i) {
        gdims[i] = dims[i];
        gbox[i] = box[i];
        estr[i] = 1;
    }
    for (int i = 0; i < rank - 1; ++i) gstr[i] = strides_bytes[i];
    CUtensorMapSwizzle sw = swizzle_bytes == 128  ? CU_TENSOR_MAP_SWIZZLE_128B
                            : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                            : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                                                  : CU_TENSOR_MAP_SWIZZLE_NONE;
    CUtensorMapDataType dt = elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    CUresult r = g_encode(out, dt, (cuuint32_t)rank, const_cast<void*>(base), gdims, gstr, gbox, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (CUresult %d): rank %d dims [%llu %llu %llu %llu %llu] box [%u %u %u %u %u] "
                  "stride0 %llu base %p",
                  (int)r, rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
                  (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0),
                  (unsigned long long)(rank > 4 ? dims[4] : 0), box[0], rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0,
                  rank > 3 ? box[3] : 0, rank > 4 ? box[4] : 0, (unsigned long long)(rank > 1 ? strides_bytes[0] : 0), base);
        return SKB_ERR_CUDA;
    }
    return SKB_OK;
}

}  // namespace skb

extern "C" {

int skb_version(void) { return SKB_VERSION; }
const char* skb_last_error(void) { return skb::g_err; }
int skb_device_check(void) { return skb::check_device(); }

}  // extern "C"
