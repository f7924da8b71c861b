// Anchor decode (DetectionHead.process_detections, skyeye/core/models/detector.py:88-145) fused with
// the [B,na,h,w,no] relayout of the raw logits (detector.py:81-82).  HBM-bound: one pass over the
// head-conv output, fp32 throughout (bf16 decode loses up to 117 px at random init).
#include "common.cuh"

namespace skb {

constexpr int DEC_PIX = 64;       // pixels per CTA chunk (32 when a pixel carries more than 128 head channels)
constexpr int DEC_THREADS = 256;
constexpr int DEC_MAX_CH = 256;   // na * no staged per pixel: 3 * 85 = 255 for the reference default num_classes = 80 (detector.py:28)
constexpr int DEC_SMEM_FLOATS = 64 * 128;

struct DecodeLevel {
    const float* raw;   // [B, h, w, pitch] fp32, channel = a*no + o
    float* raw_out;     // [B, na, h, w, no] or null
    long pitch;
    int h, w;
    long row0;          // first detection row of this level
    float stride;       // max(H/h, W/w) as a float (detector.py:107-109)
    float anchor[8][2]; // anchors[i] * stride (detector.py:119-121, quirk X16)
    int chunks_per_image;
    int chunk0;         // first CTA of this level
};
struct DecodeParams {
    DecodeLevel lv[4];
    int levels, na, no, B;
    int pix;            // pixels per CTA: pix * 4 * ceil(na*no / 4) <= DEC_SMEM_FLOATS
    unsigned int no_magic;  // ceil(2^20 / no): i / no == (i * no_magic) >> 20 for every i < DEC_PIX * no (checked on the host)
    long rows_per_image;
};

// 1 / (1 + e^-x): ex2-based exponential (relative error ~1e-7 for |x| < 20) and a correctly rounded reciprocal.  The
// library expf + IEEE division cost ~60 instructions per element and made this kernel issue-bound (82 % of the issue
// slots at 2.1 TB/s); the decode tolerance is 1e-5 relative.
__device__ __forceinline__ float sigmoid_acc(float x) { return __frcp_rn(1.0f + __expf(-x)); }

// One CTA = 64 consecutive pixels of one (level, image).  The head-conv rows (na*no floats per pixel)
// are staged in shared memory with coalesced 16-byte loads; outputs are then produced in (anchor,
// pixel, output) order, which is the memory order of BOTH destinations (det rows and raw_outputs), so
// every global store is coalesced.  (The one-thread-per-cell version read and wrote 60-byte strided
// records: 0.94 TB/s.)
__global__ void __launch_bounds__(DEC_THREADS)
decode_kernel(const DecodeParams p, float* __restrict__ det) {
    __shared__ __align__(16) float sin_[DEC_SMEM_FLOATS];
    __shared__ float sgx[DEC_PIX], sgy[DEC_PIX];  // grid (x, y) of the chunk's pixels
    int l = 0;
    while (l + 1 < p.levels && (int)blockIdx.x >= p.lv[l + 1].chunk0) ++l;
    const DecodeLevel& L = p.lv[l];
    const int cb = (int)blockIdx.x - L.chunk0;
    const int b = cb / L.chunks_per_image;
    const int hw = L.h * L.w;
    const int pix0 = (cb - b * L.chunks_per_image) * p.pix;
    const int npix = min(p.pix, hw - pix0);
    const int nch = p.na * p.no;
    const int nch4 = (nch + 3) >> 2;  // 16-byte vectors per pixel (host guarantees pitch >= 4*nch4)
    const float* src = L.raw + ((long)b * hw + pix0) * L.pitch;
    for (int i = threadIdx.x; i < npix * nch4; i += DEC_THREADS) {
        const int px = i / nch4, v = i - px * nch4;
        const float4 t = *reinterpret_cast<const float4*>(src + (long)px * L.pitch + 4 * v);
        *reinterpret_cast<float4*>(&sin_[px * (4 * nch4) + 4 * v]) = t;
    }
    if ((int)threadIdx.x < npix) {
        const int pix = pix0 + (int)threadIdx.x;
        const int y = pix / L.w;
        sgx[threadIdx.x] = (float)(pix - y * L.w);
        sgy[threadIdx.x] = (float)y;
    }
    __syncthreads();
    const int per_a = npix * p.no;
    for (int a = 0; a < p.na; ++a) {
        // destination runs: npix*no consecutive floats each
        float* dst = det + ((long)b * p.rows_per_image + L.row0 + (long)a * hw + pix0) * p.no;
        float* rdst = L.raw_out ? L.raw_out + (((long)b * p.na + a) * hw + pix0) * p.no : nullptr;
        const float aw = L.anchor[a][0], ah = L.anchor[a][1];
        // one output element: raw logit r and decoded value v of flat index i = pixel * no + output
        auto elem = [&](int i, float& r, float& v) {
            const int px = (int)(((unsigned int)i * p.no_magic) >> 20), o = i - px * p.no;
            r = sin_[px * (4 * nch4) + a * p.no + o];
            const float s = sigmoid_acc(r);
            if (o < 2) {
                // (s*2 - 0.5 + grid) * stride   (detector.py:137); grid order (x, y) (detector.py:115)
                const float g = o == 0 ? sgx[px] : sgy[px];
                v = __fmul_rn(__fadd_rn(__fsub_rn(__fmul_rn(s, 2.0f), 0.5f), g), L.stride);
            } else if (o < 4) {
                // (s*2)^2 * anchor_grid          (detector.py:138)
                const float t = __fmul_rn(s, 2.0f);
                v = __fmul_rn(__fmul_rn(t, t), o == 2 ? aw : ah);
            } else {
                v = s;
            }
        };
        // both destination runs are contiguous: 16-byte stores of four consecutive elements when the runs allow it (every
        // detector shape: the run starts are multiples of 4 floats), scalar stores otherwise
        const bool vec = (per_a & 3) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0 && (!rdst || (reinterpret_cast<uintptr_t>(rdst) & 15) == 0);
        if (vec) {
            for (int i4 = 4 * (int)threadIdx.x; i4 < per_a; i4 += 4 * DEC_THREADS) {
                float4 rr, vv;
                elem(i4, rr.x, vv.x);
                elem(i4 + 1, rr.y, vv.y);
                elem(i4 + 2, rr.z, vv.z);
                elem(i4 + 3, rr.w, vv.w);
                if (rdst) *reinterpret_cast<float4*>(rdst + i4) = rr;
                *reinterpret_cast<float4*>(dst + i4) = vv;
            }
        } else {
            for (int i = threadIdx.x; i < per_a; i += DEC_THREADS) {
                float r, v;
                elem(i, r, v);
                if (rdst) rdst[i] = r;
                dst[i] = v;
            }
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
    SKB_REQUIRE(levels >= 1 && levels <= 4 && na >= 1 && na <= 8 && no >= 5 && na * no <= DEC_MAX_CH, SKB_ERR_UNSUPPORTED,
                "decode: levels=%d na=%d no=%d", levels, na, no);
    DecodeParams p;
    memset(&p, 0, sizeof(p));
    p.levels = levels; p.na = na; p.no = no; p.B = raw[0].n;
    p.pix = (na * no + 3) / 4 * 4 <= 128 ? DEC_PIX : DEC_PIX / 2;
    p.no_magic = ((1u << 20) + (unsigned int)no - 1u) / (unsigned int)no;
    for (unsigned int i = 0; i < (unsigned int)(DEC_PIX * no); ++i)
        SKB_REQUIRE(((i * p.no_magic) >> 20) == i / (unsigned int)no, SKB_ERR_UNSUPPORTED, "decode: no=%d outside the fast-division range", no);
    const int nch4 = (na * no + 3) / 4 * 4;
    long row = 0;
    int chunk = 0;
    for (int l = 0; l < levels; ++l) {
        const skb_view& v = raw[l];
        SKB_REQUIRE(v.ptr && v.dtype == SKB_F32 && v.n == p.B && v.c >= na * no && v.pitch >= nch4 && v.pitch % 4 == 0 &&
                        ((uintptr_t)v.ptr & 15) == 0,
                    SKB_ERR_ARG, "decode: level %d view (needs fp32, pitch %% 4 == 0, pitch >= %d, 16-byte aligned)", l, nch4);
        DecodeLevel& L = p.lv[l];
        L.raw = (const float*)v.ptr; L.pitch = v.pitch; L.h = v.h; L.w = v.w; L.row0 = row;
        L.raw_out = raw_out ? raw_out[l] : nullptr;
        const float sh = (float)((double)in_h / (double)v.h), sw = (float)((double)in_w / (double)v.w);
        L.stride = sh > sw ? sh : sw;
        for (int a = 0; a < na; ++a)
            for (int k = 0; k < 2; ++k) L.anchor[a][k] = anchors_host[(l * na + a) * 2 + k] * L.stride;
        row += (long)na * v.h * v.w;
        L.chunks_per_image = (v.h * v.w + p.pix - 1) / p.pix;
        L.chunk0 = chunk;
        chunk += L.chunks_per_image * p.B;
    }
    p.rows_per_image = row;
    if (chunk == 0) return SKB_OK;
    decode_kernel<<<chunk, DEC_THREADS, 0, (cudaStream_t)stream>>>(p, det);
    SKB_LAUNCH_CHECK();
    return SKB_OK;
}
