// Anchor decode (DetectionHead.process_detections, skyeye/core/models/detector.py:88-145) fused with
// the [B,na,h,w,no] relayout of the raw logits (detector.py:81-82).  HBM-bound: one pass over the
// head-conv output, fp32 throughout (bf16 decode loses up to 117 px at random init).
#include "common.cuh"

namespace skb {

struct DecodeLevel {
    const float* raw;   // [B, h, w, pitch] fp32, channel = a*no + o
    float* raw_out;     // [B, na, h, w, no] or null
    long pitch;
    int h, w;
    long row0;          // first detection row of this level
    float stride;       // max(H/h, W/w) as a float (detector.py:107-109)
    float anchor[8][2]; // anchors[i] * stride (detector.py:119-121, quirk X16)
};
struct DecodeParams {
    DecodeLevel lv[4];
    int levels, na, no, B;
    long rows_per_image;
    long cells_total;  // sum_l B*na*h*w
    long cum[5];       // prefix of per-level B*na*h*w
};

__device__ __forceinline__ float sigmoid_acc(float x) { return 1.0f / (1.0f + expf(-x)); }

// one thread per (level, b, a, y, x) cell; it handles the `no` channels of its cell
__global__ void decode_kernel(const DecodeParams p, float* __restrict__ det) {
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < p.cells_total; i += (long)gridDim.x * blockDim.x) {
        int l = 0;
        while (l + 1 < p.levels && i >= p.cum[l + 1]) ++l;
        const DecodeLevel& L = p.lv[l];
        long j = i - p.cum[l];
        const int x = (int)(j % L.w); j /= L.w;
        const int y = (int)(j % L.h); j /= L.h;
        const int a = (int)(j % p.na);
        const int b = (int)(j / p.na);
        const float* src = L.raw + (((long)b * L.h + y) * L.w + x) * L.pitch + a * p.no;
        float* dst = det + ((long)b * p.rows_per_image + L.row0 + ((long)a * L.h + y) * L.w + x) * p.no;
        float* rdst = L.raw_out ? L.raw_out + ((((long)b * p.na + a) * L.h + y) * L.w + x) * p.no : nullptr;
        for (int o = 0; o < p.no; ++o) {
            const float r = src[o];
            if (rdst) rdst[o] = r;
            const float s = sigmoid_acc(r);
            float v;
            if (o < 2) {
                // (s*2 - 0.5 + grid) * stride   (detector.py:137); grid order (x, y) (detector.py:115)
                const float g = o == 0 ? (float)x : (float)y;
                v = __fmul_rn(__fadd_rn(__fsub_rn(__fmul_rn(s, 2.0f), 0.5f), g), L.stride);
            } else if (o < 4) {
                // (s*2)^2 * anchor_grid          (detector.py:138)
                const float t = __fmul_rn(s, 2.0f);
                v = __fmul_rn(__fmul_rn(t, t), L.anchor[a][o - 2]);
            } else {
                v = s;
            }
            dst[o] = v;
        }
    }
}

}  // namespace skb

using namespace skb;

extern "C" int skb_decode_f32(const skb_view* raw, int32_t levels, int32_t na, int32_t no, const float* anchors_host, int32_t in_h,
                              int32_t in_w, float* det, float* const* raw_out, void* stream) {
    int rc = check_device();
    if (rc != SKB_OK) return rc;
    SKB_REQUIRE(raw && det && anchors_host, SKB_ERR_ARG, "decode: null argument");
    SKB_REQUIRE(levels >= 1 && levels <= 4 && na >= 1 && na <= 8 && no >= 5, SKB_ERR_UNSUPPORTED, "decode: levels=%d na=%d no=%d", levels, na, no);
    DecodeParams p;
    memset(&p, 0, sizeof(p));
    p.levels = levels; p.na = na; p.no = no; p.B = raw[0].n;
    long row = 0;
    p.cum[0] = 0;
    for (int l = 0; l < levels; ++l) {
        const skb_view& v = raw[l];
        SKB_REQUIRE(v.ptr && v.dtype == SKB_F32 && v.n == p.B && v.c >= na * no && v.pitch >= na * no, SKB_ERR_ARG, "decode: level %d view", l);
        DecodeLevel& L = p.lv[l];
        L.raw = (const float*)v.ptr; L.pitch = v.pitch; L.h = v.h; L.w = v.w; L.row0 = row;
        L.raw_out = raw_out ? raw_out[l] : nullptr;
        const float sh = (float)((double)in_h / (double)v.h), sw = (float)((double)in_w / (double)v.w);
        L.stride = sh > sw ? sh : sw;
        for (int a = 0; a < na; ++a)
            for (int k = 0; k < 2; ++k) L.anchor[a][k] = anchors_host[(l * na + a) * 2 + k] * L.stride;
        row += (long)na * v.h * v.w;
        p.cum[l + 1] = p.cum[l] + (long)p.B * na * v.h * v.w;
    }
    p.rows_per_image = row;
    p.cells_total = p.cum[levels];
    long g = (p.cells_total + 255) / 256;
    long cap = (long)num_sms() * 16;
    decode_kernel<<<(int)(g > cap ? cap : (g < 1 ? 1 : g)), 256, 0, (cudaStream_t)stream>>>(p, det);
    SKB_LAUNCH_CHECK();
    return SKB_OK;
}
