// HBM-bound kernels of the forward path: Focus space-to-depth (+NCHW fp32 -> NHWC bf16), 5x5 max
// pool (SPP cascade), CBAM gate, cross-layer-attention core, LayerNorm.  All are vectorised (16-byte
// loads/stores of 8 bf16 channels), coalesced along the NHWC channel axis, and write straight into
// channel slices of concat buffers (views) so torch.cat never materialises.
#include <float.h>

#include "common.cuh"

namespace skb {

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
    f[0] = bf16_lo(u.x); f[1] = bf16_hi(u.x); f[2] = bf16_lo(u.y); f[3] = bf16_hi(u.y);
    f[4] = bf16_lo(u.z); f[5] = bf16_hi(u.z); f[6] = bf16_lo(u.w); f[7] = bf16_hi(u.w);
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
    uint4 u;
    u.x = pack_bf16x2(f[0], f[1]); u.y = pack_bf16x2(f[2], f[3]);
    u.z = pack_bf16x2(f[4], f[5]); u.w = pack_bf16x2(f[6], f[7]);
    return u;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// ---------------------------------------------------------------------------------------------
// Focus (blocks.py:170-181): y[n, oy, ox, p*3 + c] = img[n, c, 2*oy + dy(p), 2*ox + dx(p)]
// patches: p0 TL (0,0), p1 BL (1,0), p2 TR (0,1), p3 BR (1,1); channels >= 12 are zero.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float2 focus_ld2(const float* p) { return *reinterpret_cast<const float2*>(p); }
__device__ __forceinline__ float2 focus_ld2(const uint8_t* p) {  // uint8 image: img.float() / 255 (validate.py:237-238)
    const uchar2 u = *reinterpret_cast<const uchar2*>(p);
    return make_float2(__fdiv_rn((float)u.x, 255.0f), __fdiv_rn((float)u.y, 255.0f));
}
template <typename T>
__global__ void focus_kernel(const T* __restrict__ img, int N, int H, int W, __nv_bfloat16* __restrict__ y, int pitch, int cpad) {
    const int Ho = H / 2, Wo = W / 2;
    const long total = (long)N * Ho * Wo;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int ox = (int)(i % Wo);
        const int oy = (int)((i / Wo) % Ho);
        const int n = (int)(i / ((long)Wo * Ho));
        float v[12];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const T* base = img + (((long)n * 3 + c) * H + 2 * oy) * W + 2 * ox;
            const float2 top = focus_ld2(base);
            const float2 bot = focus_ld2(base + W);
            v[0 * 3 + c] = top.x;  // TL
            v[1 * 3 + c] = bot.x;  // BL
            v[2 * 3 + c] = top.y;  // TR
            v[3 * 3 + c] = bot.y;  // BR
        }
        uint4* o = reinterpret_cast<uint4*>(y + i * pitch);
        uint4 a, b;
        a.x = pack_bf16x2(v[0], v[1]); a.y = pack_bf16x2(v[2], v[3]); a.z = pack_bf16x2(v[4], v[5]); a.w = pack_bf16x2(v[6], v[7]);
        b.x = pack_bf16x2(v[8], v[9]); b.y = pack_bf16x2(v[10], v[11]); b.z = 0u; b.w = 0u;
        o[0] = a;
        o[1] = b;
        const uint4 z = make_uint4(0u, 0u, 0u, 0u);
        for (int g = 2; g < cpad / 8; ++g) o[g] = z;
    }
}

// Same space-to-depth into the padded 16-channel scratch of skb_focus_conv_bf16: row = [0 | Wo pixels | 0 0 0]
// One CTA per output row (n, oy): no per-element index division (the flat 64-bit i % Wp, i / Wp form spent most of its
// issue slots on it: 1.7 TB/s).
// TILED = true: image n is a window of a larger frame (4K drone frames cut into 1280^2 tiles): `tiles` holds
// (frame, y0, x0) per image and the window is read in place -- the tile batch never exists in memory.  Window origins may
// be odd (x0 = 853), so that variant loads single pixels.
__device__ __forceinline__ float focus_ld1(const float* p) { return *p; }
__device__ __forceinline__ float focus_ld1(const uint8_t* p) { return __fdiv_rn((float)*p, 255.0f); }
template <typename T, bool TILED>
__global__ void focus_pad_kernel(const T* __restrict__ img, int N, int H, int W, __nv_bfloat16* __restrict__ y,
                                 const int* __restrict__ tiles, int FH, int FW) {
    const int Ho = H / 2, Wo = W / 2, Wp = Wo + 4;
    const int oy = (int)(blockIdx.x % (unsigned int)Ho);
    const int n = (int)(blockIdx.x / (unsigned int)Ho);
    const T* src;       // pixel (0, 2*oy, 0) of image n
    long plane, pitch;  // elements between channel planes / rows
    if (TILED) {
        const int f = tiles[3 * n], y0 = tiles[3 * n + 1], x0 = tiles[3 * n + 2];
        plane = (long)FH * FW;
        pitch = FW;
        src = img + (long)f * 3 * plane + (long)(y0 + 2 * oy) * FW + x0;
    } else {
        plane = (long)H * W;
        pitch = W;
        src = img + (long)n * 3 * plane + (long)(2 * oy) * W;
    }
    for (int col = threadIdx.x; col < Wp; col += blockDim.x) {
        const long i = ((long)n * Ho + oy) * Wp + col;
        uint4 a = make_uint4(0u, 0u, 0u, 0u), b = a;
        const int ox = col - 1;
        if (ox >= 0 && ox < Wo) {
            float v[12];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const T* base = src + c * plane + 2 * ox;
                float2 top, bot;
                if (TILED) {
                    top = make_float2(focus_ld1(base), focus_ld1(base + 1));
                    bot = make_float2(focus_ld1(base + pitch), focus_ld1(base + pitch + 1));
                } else {
                    top = focus_ld2(base);
                    bot = focus_ld2(base + pitch);
                }
                v[0 * 3 + c] = top.x; v[1 * 3 + c] = bot.x; v[2 * 3 + c] = top.y; v[3 * 3 + c] = bot.y;
            }
            a.x = pack_bf16x2(v[0], v[1]); a.y = pack_bf16x2(v[2], v[3]); a.z = pack_bf16x2(v[4], v[5]); a.w = pack_bf16x2(v[6], v[7]);
            b.x = pack_bf16x2(v[8], v[9]); b.y = pack_bf16x2(v[10], v[11]);
        }
        uint4* o = reinterpret_cast<uint4*>(y + i * 16);
        o[0] = a;
        o[1] = b;
    }
}

// uint8 images whose rows are multiples of 16 pixels (every detector input: sizes are multiples of the stride 32): a thread
// produces EIGHT output pixels from six 16-byte loads (3 channels x 2 rows x 16 input columns) and writes 256 contiguous bytes.
// The one-pixel-per-thread form issued six 2-byte loads per 32-byte output (26 % of the HBM peak, ncu: profiles/
// r2l_ncu_memory_bound_kernels.md).  Same arithmetic per element (__fdiv_rn(u8, 255)): bit-identical output.
__global__ void __launch_bounds__(256)
focus_pad8_u8_kernel(const uint8_t* __restrict__ img, int N, int H, int W, __nv_bfloat16* __restrict__ y) {
    const int Ho = H / 2, Wo = W / 2, Wp = Wo + 4, segs = Wo / 8;
    const long plane = (long)H * W;
    const long items = (long)N * Ho * segs;
    for (long it = blockIdx.x * (long)blockDim.x + threadIdx.x; it < items; it += (long)gridDim.x * blockDim.x) {
        const int seg = (int)(it % segs);
        const long r = it / segs;            // padded output row (n * Ho + oy)
        const int oy = (int)(r % Ho);
        const long n = r / Ho;
        const uint8_t* src = img + n * 3 * plane + (long)(2 * oy) * W + 16 * seg;
        uint4 top[3], bot[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            top[c] = __ldg(reinterpret_cast<const uint4*>(src + c * plane));
            bot[c] = __ldg(reinterpret_cast<const uint4*>(src + c * plane + W));
        }
        uint4* o = reinterpret_cast<uint4*>(y + (r * Wp + 1 + 8 * seg) * 16);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float v[12];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const unsigned int tw = (&top[c].x)[i >> 1] >> ((i & 1) * 16), bw = (&bot[c].x)[i >> 1] >> ((i & 1) * 16);
                v[0 * 3 + c] = __fdiv_rn((float)(tw & 0xffu), 255.0f);
                v[1 * 3 + c] = __fdiv_rn((float)(bw & 0xffu), 255.0f);
                v[2 * 3 + c] = __fdiv_rn((float)((tw >> 8) & 0xffu), 255.0f);
                v[3 * 3 + c] = __fdiv_rn((float)((bw >> 8) & 0xffu), 255.0f);
            }
            uint4 a, b;
            a.x = pack_bf16x2(v[0], v[1]); a.y = pack_bf16x2(v[2], v[3]); a.z = pack_bf16x2(v[4], v[5]); a.w = pack_bf16x2(v[6], v[7]);
            b.x = pack_bf16x2(v[8], v[9]); b.y = pack_bf16x2(v[10], v[11]); b.z = 0u; b.w = 0u;
            o[2 * i] = a;
            o[2 * i + 1] = b;
        }
        const uint4 z = make_uint4(0u, 0u, 0u, 0u);
        if (seg == 0) { uint4* o0 = reinterpret_cast<uint4*>(y + (r * Wp) * 16); o0[0] = z; o0[1] = z; }   // left pad column
        if (seg == segs - 1) {                                                                                // three right pad columns
            uint4* o1 = reinterpret_cast<uint4*>(y + (r * Wp + 1 + Wo) * 16);
#pragma unroll
            for (int q = 0; q < 6; ++q) o1[q] = z;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// letterbox + BGR->RGB + HWC->CHW (augmentation.py:442-496, detect.py:131-132), OpenCV's 8-bit INTER_LINEAR
// ---------------------------------------------------------------------------------------------
struct LinCoef {
    int s0, s1, a0, a1;  // source indices and 11-bit fixed-point weights
};
// cv::resize INTER_LINEAR coefficient of destination index d: f = (float)((d + 0.5) * scale - 0.5), split into
// floor + fraction, clamped at the borders, weights = cvRound((1 - f) * 2048), cvRound(f * 2048)
__device__ __forceinline__ LinCoef lin_coef(int d, double scale, int n_src) {
    float f = (float)(((double)d + 0.5) * scale - 0.5);
    int s = (int)floorf(f);
    f -= (float)s;
    if (s < 0) { f = 0.f; s = 0; }
    if (s >= n_src - 1) { f = 0.f; s = n_src - 1; }
    LinCoef c;
    c.s0 = s;
    c.s1 = min(s + 1, n_src - 1);
    c.a0 = __float2int_rn((1.0f - f) * 2048.0f);
    c.a1 = __float2int_rn(f * 2048.0f);
    return c;
}
__global__ void letterbox_kernel(const uint8_t* __restrict__ src, int h0, int w0, int pitch, uint8_t* __restrict__ dst, int H, int W,
                                 int new_h, int new_w, int top, int left, int pad, double sx, double sy, int resize) {
    const long total = (long)H * W;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int x = (int)(i % W), y = (int)(i / W);
        const int rx = x - left, ry = y - top;
        int b = pad, g = pad, r = pad;
        if (rx >= 0 && rx < new_w && ry >= 0 && ry < new_h) {
            if (!resize) {
                const uint8_t* px = src + (long)ry * pitch + rx * 3;
                b = px[0]; g = px[1]; r = px[2];
            } else {
                const LinCoef cx = lin_coef(rx, sx, w0), cy = lin_coef(ry, sy, h0);
                const uint8_t* r0 = src + (long)cy.s0 * pitch;
                const uint8_t* r1 = src + (long)cy.s1 * pitch;
                int out[3];
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const int h0v = r0[cx.s0 * 3 + c] * cx.a0 + r0[cx.s1 * 3 + c] * cx.a1;  // horizontal pass, x 2048
                    const int h1v = r1[cx.s0 * 3 + c] * cx.a0 + r1[cx.s1 * 3 + c] * cx.a1;
                    out[c] = (((cy.a0 * (h0v >> 4)) >> 16) + ((cy.a1 * (h1v >> 4)) >> 16) + 2) >> 2;
                }
                b = out[0]; g = out[1]; r = out[2];
            }
        }
        dst[i] = (uint8_t)r;                 // RGB planes
        dst[total + i] = (uint8_t)g;
        dst[2 * total + i] = (uint8_t)b;
    }
}

// ---------------------------------------------------------------------------------------------
// MaxPool2d(5, 1, 2) on bf16 NHWC (padding acts as -inf)
// ---------------------------------------------------------------------------------------------
// One thread per (image, column x, group of 8 channels) walks down the column and keeps the horizontal 5-maxima of the
// five most recent rows in registers: 5 loads per output instead of 25 (the 25-load form was issue-bound: 69 % of the
// issue slots at 1.0 TB/s).  bf16 max is exact, so the result is bit-identical.
__global__ void maxpool5_kernel(const __nv_bfloat16* __restrict__ x, long xpitch, __nv_bfloat16* __restrict__ y, long ypitch,
                                int N, int H, int W, int C8, int segs) {
    const long total = (long)N * segs * W * C8;  // a column is cut into `segs` row segments (more threads in flight)
    const int rows_per = (H + segs - 1) / segs;
    const __nv_bfloat162 ninf = __float2bfloat162_rn(-INFINITY);
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int g = (int)(i % C8);
        long t = i / C8;
        const int px = (int)(t % W);
        t /= W;
        const int seg = (int)(t % segs);
        const int n = (int)(t / segs);
        const int y0 = seg * rows_per, y1 = min(H, y0 + rows_per);
        const int xa = max(px - 2, 0), xb = min(px + 2, W - 1);
        const __nv_bfloat16* col = x + ((long)n * H * W) * xpitch + g * 8;
        auto rowmax = [&](int r, __nv_bfloat162 (&m)[4]) {
            m[0] = m[1] = m[2] = m[3] = ninf;
            if (r < 0 || r >= H) return;
            const __nv_bfloat16* rowp = col + ((long)r * W) * xpitch;
            for (int xx = xa; xx <= xb; ++xx) {
                const uint4 u = *reinterpret_cast<const uint4*>(rowp + (long)xx * xpitch);
                const __nv_bfloat162* v = reinterpret_cast<const __nv_bfloat162*>(&u);
                m[0] = __hmax2(m[0], v[0]); m[1] = __hmax2(m[1], v[1]);
                m[2] = __hmax2(m[2], v[2]); m[3] = __hmax2(m[3], v[3]);
            }
        };
        __nv_bfloat162 r0[4], r1[4], r2[4], r3[4], r4[4];  // horizontal maxima of rows py-2 .. py+2
        rowmax(y0 - 2, r0); rowmax(y0 - 1, r1); rowmax(y0, r2); rowmax(y0 + 1, r3);
        for (int py = y0; py < y1; ++py) {
            rowmax(py + 2, r4);
            __nv_bfloat162 m[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) m[q] = __hmax2(__hmax2(__hmax2(r0[q], r1[q]), __hmax2(r2[q], r3[q])), r4[q]);
            *reinterpret_cast<uint4*>(y + (((long)n * H + py) * W + px) * ypitch + g * 8) = *reinterpret_cast<uint4*>(m);
#pragma unroll
            for (int q = 0; q < 4; ++q) { r0[q] = r1[q]; r1[q] = r2[q]; r2[q] = r3[q]; r3[q] = r4[q]; }
        }
    }
}

// SPP's three pools in ONE pass (blocks.py:143-149): a CTA owns the whole H x W map of 16 channels of one image (32 bytes per
// pixel = one DRAM sector), reads it once into shared memory and runs the 5x5 cascade there (5, 9 = 5o5, 13 = 5o5o5: exact for
// max with -inf padding), writing each stage's result to its concat slice.  Three chained launches read and wrote the map three
// times; this reads it once and writes three slices.  Separable: row maxima A -> T, column maxima T -> A (A is dead once its row
// maxima exist, so two buffers do: 2 x H*W*32 B = 102 KB at 40 x 40, two CTAs of 512 threads per SM); larger maps keep the
// cascade of launches.  A thread owns one 16-byte vector column-slot and walks it with a sliding window (no per-item division).
constexpr int SPP_THREADS = 512;
__global__ void __launch_bounds__(SPP_THREADS, 2)
spp_fused_kernel(const __nv_bfloat16* __restrict__ x, long xpitch, __nv_bfloat16* __restrict__ y5, __nv_bfloat16* __restrict__ y9,
                 __nv_bfloat16* __restrict__ y13, long ypitch, int H, int W, int groups) {
    extern __shared__ __align__(16) uint4 spp_smem[];
    const int items = H * W * 2;  // 16-byte vectors: [y][x][2]
    uint4* A = spp_smem;          // stage input / output
    uint4* T = A + items;         // row maxima
    const int n = blockIdx.x / groups, cg = blockIdx.x - n * groups;
    const long img = (long)n * H * W;
    for (int i = threadIdx.x; i < items; i += SPP_THREADS) {
        const int pix = i >> 1, v = i & 1;
        A[i] = __ldg(reinterpret_cast<const uint4*>(x + (img + pix) * xpitch + cg * 16 + v * 8));
    }
    __syncthreads();
    auto vmax = [](uint4 a, const uint4& b) {
        __nv_bfloat162* pa = reinterpret_cast<__nv_bfloat162*>(&a);
        const __nv_bfloat162* pb = reinterpret_cast<const __nv_bfloat162*>(&b);
#pragma unroll
        for (int q = 0; q < 4; ++q) pa[q] = __hmax2(pa[q], pb[q]);
        return a;
    };
    __nv_bfloat16* outs[3] = {y5, y9, y13};
    const int W2 = W * 2;
#pragma unroll 1
    for (int stage = 0; stage < 3; ++stage) {
        for (int i = threadIdx.x; i < items; i += SPP_THREADS) {  // horizontal 5-max (clamped window = -inf padding)
            const int pix = i >> 1, px = pix % W;
            uint4 m = A[i];
            if (px >= 1) m = vmax(m, A[i - 2]);
            if (px >= 2) m = vmax(m, A[i - 4]);
            if (px + 1 < W) m = vmax(m, A[i + 2]);
            if (px + 2 < W) m = vmax(m, A[i + 4]);
            T[i] = m;
        }
        __syncthreads();
        __nv_bfloat16* out = outs[stage] + cg * 16;
        for (int i = threadIdx.x; i < items; i += SPP_THREADS) {  // vertical 5-max, store
            const int pix = i >> 1, v = i & 1, py = pix / W;
            uint4 m = T[i];
            if (py >= 1) m = vmax(m, T[i - W2]);
            if (py >= 2) m = vmax(m, T[i - 2 * W2]);
            if (py + 1 < H) m = vmax(m, T[i + W2]);
            if (py + 2 < H) m = vmax(m, T[i + 2 * W2]);
            A[i] = m;  // the stage's output is the next stage's input
            *reinterpret_cast<uint4*>(out + (img + pix) * ypitch + v * 8) = m;
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// CBAM (attention.py:37-60, 80-98)
// ---------------------------------------------------------------------------------------------
// stage 1: per-image partial channel sums / maxima over a slab of pixels. grid (slabs, N)
__global__ void cbam_pool_kernel(const __nv_bfloat16* __restrict__ x, long pitch, int HW, int C, int slabs,
                                 float* __restrict__ psum, float* __restrict__ pmax) {
    const int n = blockIdx.y, slab = blockIdx.x;
    const int C8 = C / 8;
    const int lanes = blockDim.x / C8;  // pixel lanes per channel group (host guarantees >= 1)
    const int g = threadIdx.x % C8, pl = threadIdx.x / C8;
    const int per = (HW + slabs - 1) / slabs;
    const int p0 = slab * per, p1 = min(HW, p0 + per);
    float s[8], m[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { s[j] = 0.f; m[j] = -FLT_MAX; }
    if (pl < lanes) {
        // eight independent 16-byte loads in flight per thread (one load per iteration was latency-bound: 64 dependent round trips
        // per thread, 39 us for the 105 MB map of skyeye_l's P4 level)
        constexpr int U = 8;
        for (int pb = p0 + pl; pb < p1; pb += lanes * U) {
            uint4 u[U];
#pragma unroll
            for (int k = 0; k < U; ++k) {
                const int p = pb + k * lanes;
                u[k] = p < p1 ? __ldg(reinterpret_cast<const uint4*>(x + ((long)n * HW + p) * pitch + g * 8)) : make_uint4(0u, 0u, 0u, 0u);
            }
#pragma unroll
            for (int k = 0; k < U; ++k) {
                if (pb + k * lanes < p1) {
                    float f[8];
                    unpack8(u[k], f);
#pragma unroll
                    for (int j = 0; j < 8; ++j) { s[j] += f[j]; m[j] = fmaxf(m[j], f[j]); }
                }
            }
        }
    }
    extern __shared__ float sh[];  // [lanes][C] sums then [lanes][C] maxima
    float* ss = sh;
    float* sm = sh + (size_t)lanes * C;
    if (pl < lanes) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { ss[pl * C + g * 8 + j] = s[j]; sm[pl * C + g * 8 + j] = m[j]; }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float a = 0.f, b = -FLT_MAX;
        for (int l = 0; l < lanes; ++l) { a += ss[l * C + c]; b = fmaxf(b, sm[l * C + c]); }
        psum[((long)n * slabs + slab) * C + c] = a;
        pmax[((long)n * slabs + slab) * C + c] = b;
    }
}
// stage 2: att[n][c] = sigmoid(W1 relu(W0 avg) + W1 relu(W0 max)). grid N, smem 2C + 2R floats
__global__ void cbam_mlp_kernel(const float* __restrict__ psum, const float* __restrict__ pmax, int slabs, int HW, int C, int R,
                                const float* __restrict__ w0, const float* __restrict__ w1, float* __restrict__ att) {
    extern __shared__ float sh[];
    float* avg = sh; float* mx = sh + C; float* ha = sh + 2 * C; float* hm = ha + R;
    const int n = blockIdx.x;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float a = 0.f, b = -FLT_MAX;
        for (int s0 = 0; s0 < slabs; s0 += 8) {  // eight independent loads of each array in flight (the sum order stays s = 0, 1, ...)
            float ps[8], pm[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const bool ok = s0 + k < slabs;
                ps[k] = ok ? __ldg(psum + ((long)n * slabs + s0 + k) * C + c) : 0.f;
                pm[k] = ok ? __ldg(pmax + ((long)n * slabs + s0 + k) * C + c) : -FLT_MAX;
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) { a += ps[k]; b = fmaxf(b, pm[k]); }
        }
        avg[c] = a / (float)HW;
        mx[c] = b;
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    for (int r = warp; r < R; r += nw) {
        float a = 0.f, b = 0.f;
#pragma unroll 8
        for (int c = lane; c < C; c += 32) { const float w = __ldg(w0 + r * C + c); a += w * avg[c]; b += w * mx[c]; }
        a = warp_sum(a); b = warp_sum(b);
        if (lane == 0) { ha[r] = fmaxf(a, 0.f); hm[r] = fmaxf(b, 0.f); }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float a = 0.f;
#pragma unroll 8
        for (int r = 0; r < R; ++r) a += __ldg(w1 + c * R + r) * (ha[r] + hm[r]);
        att[(long)n * C + c] = 1.0f / (1.0f + expf(-a));
    }
}
// stage 3: per pixel mean/max over channels of x*att. one warp per pixel
__global__ void cbam_stats_kernel(const __nv_bfloat16* __restrict__ x, long pitch, const float* __restrict__ att, long npix, int HW,
                                  int C, float* __restrict__ stats) {
    const int lane = threadIdx.x & 31;
    const long wid = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5;
    const long nw = ((long)gridDim.x * blockDim.x) >> 5;
    for (long p = wid; p < npix; p += nw) {
        const int n = (int)(p / HW);
        float s = 0.f, m = -FLT_MAX;
        for (int c = lane * 8; c < C; c += 256) {
            const uint4 u = *reinterpret_cast<const uint4*>(x + p * pitch + c);
            float f[8];
            unpack8(u, f);
            const float4 a0 = *reinterpret_cast<const float4*>(att + (long)n * C + c);
            const float4 a1 = *reinterpret_cast<const float4*>(att + (long)n * C + c + 4);
            f[0] *= a0.x; f[1] *= a0.y; f[2] *= a0.z; f[3] *= a0.w; f[4] *= a1.x; f[5] *= a1.y; f[6] *= a1.z; f[7] *= a1.w;
#pragma unroll
            for (int j = 0; j < 8; ++j) { s += f[j]; m = fmaxf(m, f[j]); }
        }
        s = warp_sum(s);
        m = warp_max(m);
        if (lane == 0) { stats[p * 2] = s / (float)C; stats[p * 2 + 1] = m; }
    }
}
// stage 4: sa = sigmoid(conv7x7([mean, max])), y = x * att * sa. one warp per pixel
__global__ void cbam_apply_kernel(const __nv_bfloat16* __restrict__ x, long xpitch, const float* __restrict__ att,
                                  const float* __restrict__ stats, const float* __restrict__ w7, int N, int H, int W, int C,
                                  __nv_bfloat16* __restrict__ y, long ypitch) {
    __shared__ float sw[98];
    for (int i = threadIdx.x; i < 98; i += blockDim.x) sw[i] = w7[i];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const long npix = (long)N * H * W;
    const long wid = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5;
    const long nw = ((long)gridDim.x * blockDim.x) >> 5;
    for (long p = wid; p < npix; p += nw) {
        const int px = (int)(p % W);
        const int py = (int)((p / W) % H);
        const int n = (int)(p / ((long)W * H));
        float acc = 0.f;
        for (int t = lane; t < 49; t += 32) {
            const int dy = t / 7 - 3, dx = t % 7 - 3;
            const int yy = py + dy, xx = px + dx;
            if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
                const float2 st = *reinterpret_cast<const float2*>(stats + (((long)n * H + yy) * W + xx) * 2);
                acc += sw[t] * st.x + sw[49 + t] * st.y;
            }
        }
        acc = warp_sum(acc);
        const float sa = 1.0f / (1.0f + expf(-acc));
        for (int c = lane * 8; c < C; c += 256) {
            const uint4 u = *reinterpret_cast<const uint4*>(x + p * xpitch + c);
            float f[8];
            unpack8(u, f);
            const float4 a0 = *reinterpret_cast<const float4*>(att + (long)n * C + c);
            const float4 a1 = *reinterpret_cast<const float4*>(att + (long)n * C + c + 4);
            f[0] = f[0] * a0.x * sa; f[1] = f[1] * a0.y * sa; f[2] = f[2] * a0.z * sa; f[3] = f[3] * a0.w * sa;
            f[4] = f[4] * a1.x * sa; f[5] = f[5] * a1.y * sa; f[6] = f[6] * a1.z * sa; f[7] = f[7] * a1.w * sa;
            *reinterpret_cast<uint4*>(y + p * ypitch + c) = pack8(f);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Cross-layer attention core (attention.py:196-238 with R4, closed form SURVEY.md §8 A10)
// ---------------------------------------------------------------------------------------------
struct Bilin {
    int y0, y1, x0, x1;
    float ly, lx;
};
// F.interpolate(mode='bilinear', align_corners=False): src = max(0, (dst + .5) * in/out - .5)
__device__ __forceinline__ Bilin bilin_coords(int y, int x, int H, int W, int Hk, int Wk) {
    Bilin b;
    const float sy = fmaxf(((float)y + 0.5f) * ((float)Hk / (float)H) - 0.5f, 0.f);
    const float sx = fmaxf(((float)x + 0.5f) * ((float)Wk / (float)W) - 0.5f, 0.f);
    b.y0 = (int)sy; b.x0 = (int)sx;
    b.y1 = min(b.y0 + 1, Hk - 1); b.x1 = min(b.x0 + 1, Wk - 1);
    b.ly = sy - (float)b.y0; b.lx = sx - (float)b.x0;
    return b;
}
__device__ __forceinline__ void bilin_load8(const __nv_bfloat16* __restrict__ t, long pitch, int n, int Hk, int Wk, const Bilin& b, int c, float (&o)[8]) {
    const long base = (long)n * Hk * Wk;
    float a[8], bb[8], cc[8], d[8];
    unpack8(*reinterpret_cast<const uint4*>(t + (base + (long)b.y0 * Wk + b.x0) * pitch + c), a);
    unpack8(*reinterpret_cast<const uint4*>(t + (base + (long)b.y0 * Wk + b.x1) * pitch + c), bb);
    unpack8(*reinterpret_cast<const uint4*>(t + (base + (long)b.y1 * Wk + b.x0) * pitch + c), cc);
    unpack8(*reinterpret_cast<const uint4*>(t + (base + (long)b.y1 * Wk + b.x1) * pitch + c), d);
    const float w00 = (1.f - b.ly) * (1.f - b.lx), w01 = (1.f - b.ly) * b.lx, w10 = b.ly * (1.f - b.lx), w11 = b.ly * b.lx;
    // packed fp32x2: the same multiply + three fused multiply-adds per channel, two channels per issue slot (the CLA kernels
    // are issue-bound: 63-69 % of the issue slots at 2.1 TB/s)
    const float2 v00 = make_float2(w00, w00), v01 = make_float2(w01, w01), v10 = make_float2(w10, w10), v11 = make_float2(w11, w11);
#pragma unroll
    for (int j = 0; j < 8; j += 2) {
        float2 r = fmul2(v00, make_float2(a[j], a[j + 1]));
        r = ffma2(v01, make_float2(bb[j], bb[j + 1]), r);
        r = ffma2(v10, make_float2(cc[j], cc[j + 1]), r);
        r = ffma2(v11, make_float2(d[j], d[j + 1]), r);
        o[j] = r.x; o[j + 1] = r.y;
    }
}
// scores s[n,g,y,x] = scale * sum_{c in head g} q * bilinear(k).  One CTA per image row (n, y): warps
// stride over x two pixels at a time (10 independent 16-byte loads in flight per lane), a head is a
// lane-aligned group of cph/8 lanes (segmented shuffle reduce), and the row's scores are staged in
// shared memory so the global writes are coalesced.
__device__ __forceinline__ void bilin_y(int y, int H, int Hk, int& y0, int& y1, float& ly) {
    const float sy = fmaxf(((float)y + 0.5f) * ((float)Hk / (float)H) - 0.5f, 0.f);
    y0 = (int)sy; y1 = min(y0 + 1, Hk - 1); ly = sy - (float)y0;
}
__global__ void __launch_bounds__(256)
cla_score_kernel(const __nv_bfloat16* __restrict__ q, long qpitch, const __nv_bfloat16* __restrict__ k, long kpitch,
                 int N, int H, int W, int Hk, int Wk, int Cq, int heads, float scale, float* __restrict__ s) {
    extern __shared__ float srow[];  // [heads][W]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const int n = blockIdx.x / H, py = blockIdx.x - n * H;
    const int cph = Cq / heads, lph = cph >> 3;  // lanes per head (power of two, host-checked)
    Bilin b0, b1;
    bilin_y(py, H, Hk, b0.y0, b0.y1, b0.ly);
    b1.y0 = b0.y0; b1.y1 = b0.y1; b1.ly = b0.ly;
    const float rx = (float)Wk / (float)W;
    for (int x0 = warp; x0 < W; x0 += 2 * nwarp) {
        const int x1 = x0 + nwarp;
        const bool two = x1 < W;
        {
            const float sx = fmaxf(((float)x0 + 0.5f) * rx - 0.5f, 0.f);
            b0.x0 = (int)sx; b0.x1 = min(b0.x0 + 1, Wk - 1); b0.lx = sx - (float)b0.x0;
            const float sx1 = fmaxf(((float)(two ? x1 : x0) + 0.5f) * rx - 0.5f, 0.f);
            b1.x0 = (int)sx1; b1.x1 = min(b1.x0 + 1, Wk - 1); b1.lx = sx1 - (float)b1.x0;
        }
        const long p0 = ((long)n * H + py) * W + x0, p1 = two ? p0 + nwarp : p0;
        for (int cbase = 0; cbase < Cq; cbase += 256) {
            const int c = cbase + lane * 8;
            float part0 = 0.f, part1 = 0.f;
            if (c < Cq) {
                float q0[8], q1[8], k0[8], k1[8];
                unpack8(*reinterpret_cast<const uint4*>(q + p0 * qpitch + c), q0);
                unpack8(*reinterpret_cast<const uint4*>(q + p1 * qpitch + c), q1);
                bilin_load8(k, kpitch, n, Hk, Wk, b0, c, k0);
                bilin_load8(k, kpitch, n, Hk, Wk, b1, c, k1);
#pragma unroll
                for (int j = 0; j < 8; ++j) { part0 += q0[j] * k0[j]; part1 += q1[j] * k1[j]; }
            }
            for (int o = lph >> 1; o > 0; o >>= 1) {
                part0 += __shfl_xor_sync(0xffffffffu, part0, o);
                part1 += __shfl_xor_sync(0xffffffffu, part1, o);
            }
            if (c < Cq && (lane & (lph - 1)) == 0) {
                const int g = c / cph;
                srow[g * W + x0] = part0 * scale;
                if (two) srow[g * W + x1] = part1 * scale;
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < heads * W; i += blockDim.x) {
        const int g = i / W, x = i - g * W;
        s[(((long)n * heads + g) * H + py) * W + x] = srow[i];
    }
}
// column softmax statistics over image rows: 32 columns x 8 row partitions per CTA (online max / sum per partition, merged
// through shared memory); the one-thread-per-column form ran 10 240 threads for the whole P3 level.
__global__ void __launch_bounds__(256) cla_colstat_kernel(const float* __restrict__ s, int NG, int H, int W, float* __restrict__ st) {
    __shared__ float sm_m[8][32], sm_s[8][32];
    const int tx = threadIdx.x & 31, part = threadIdx.x >> 5;
    const int xblocks = (W + 31) / 32;
    const int ng = blockIdx.x / xblocks;
    const int x = (blockIdx.x - ng * xblocks) * 32 + tx;
    const int rows_per = (H + 7) / 8;
    const int y0 = part * rows_per, y1 = min(H, y0 + rows_per);
    float m = -FLT_MAX, sum = 0.f;
    if (x < W) {
        const float* col = s + (long)ng * H * W + x;
        for (int y = y0; y < y1; ++y) m = fmaxf(m, col[(long)y * W]);
        for (int y = y0; y < y1; ++y) sum += expf(col[(long)y * W] - m);
    }
    sm_m[part][tx] = m;
    sm_s[part][tx] = sum;
    __syncthreads();
    if (part == 0 && x < W) {
        float mm = sm_m[0][tx];
#pragma unroll
        for (int q = 1; q < 8; ++q) mm = fmaxf(mm, sm_m[q][tx]);
        float tot = 0.f;
#pragma unroll
        for (int q = 0; q < 8; ++q) tot += sm_s[q][tx] * expf(sm_m[q][tx] - mm);  // empty partitions: sum 0
        const long i = (long)ng * W + x;
        st[i * 2] = mm;
        st[i * 2 + 1] = 1.0f / tot;
    }
}
// o[n,y,x,c] = r2 * softmax_y(s)[head(c)] * bilinear(v)[c].  One CTA per image row: the row's attention
// weights are computed once into shared memory, then warps stride over x two pixels at a time.
__global__ void __launch_bounds__(256)
cla_apply_kernel(const float* __restrict__ s, const float* __restrict__ st, const __nv_bfloat16* __restrict__ v, long vpitch,
                 int N, int H, int W, int Hk, int Wk, int Cv, int heads, float r2, __nv_bfloat16* __restrict__ o, long opitch) {
    extern __shared__ float arow[];  // [heads][W]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const int n = blockIdx.x / H, py = blockIdx.x - n * H;
    for (int i = threadIdx.x; i < heads * W; i += blockDim.x) {
        const int g = i / W, x = i - g * W;
        const long ng = (long)n * heads + g;
        const float sc = s[(ng * H + py) * W + x];
        const float2 ms = *reinterpret_cast<const float2*>(st + (ng * W + x) * 2);
        arow[i] = r2 * expf(sc - ms.x) * ms.y;
    }
    __syncthreads();
    const int cph = Cv / heads;
    Bilin b0, b1;
    bilin_y(py, H, Hk, b0.y0, b0.y1, b0.ly);
    b1.y0 = b0.y0; b1.y1 = b0.y1; b1.ly = b0.ly;
    const float rx = (float)Wk / (float)W;
    for (int x0 = warp; x0 < W; x0 += 2 * nwarp) {
        const int x1 = x0 + nwarp;
        const bool two = x1 < W;
        {
            const float sx = fmaxf(((float)x0 + 0.5f) * rx - 0.5f, 0.f);
            b0.x0 = (int)sx; b0.x1 = min(b0.x0 + 1, Wk - 1); b0.lx = sx - (float)b0.x0;
            const float sx1 = fmaxf(((float)(two ? x1 : x0) + 0.5f) * rx - 0.5f, 0.f);
            b1.x0 = (int)sx1; b1.x1 = min(b1.x0 + 1, Wk - 1); b1.lx = sx1 - (float)b1.x0;
        }
        const long p0 = ((long)n * H + py) * W + x0, p1 = p0 + nwarp;
        for (int c = lane * 8; c < Cv; c += 256) {
            float v0[8], v1[8];
            bilin_load8(v, vpitch, n, Hk, Wk, b0, c, v0);
            bilin_load8(v, vpitch, n, Hk, Wk, b1, c, v1);
            const int g = c / cph;
            const float w0 = arow[g * W + x0], w1 = arow[g * W + (two ? x1 : x0)];
#pragma unroll
            for (int j = 0; j < 8; ++j) { v0[j] *= w0; v1[j] *= w1; }
            *reinterpret_cast<uint4*>(o + p0 * opitch + c) = pack8(v0);
            if (two) *reinterpret_cast<uint4*>(o + p1 * opitch + c) = pack8(v1);
        }
    }
}

// ---- exact 2x ratio (H = 2 Hk, W = 2 Wk: every CLA call of the detector) -------------------------------------------------
// With align_corners = False and scale 2 the source coordinate of output row y is y/2 - 0.25: output rows 2b-1 and 2b both
// interpolate source rows (b-1, b), with weights (0.75, 0.25) and (0.25, 0.75); likewise for columns.  So the 2 x 2 output block
// {2bi-1, 2bi} x {2bj-1, 2bj} needs exactly FOUR source pixels, which one thread loads and unpacks once (the generic kernels
// load and unpack four source vectors per OUTPUT pixel: they are issue-bound at ~2 TB/s).  Border blocks (bi = 0 or Hk, bj = 0
// or Wk) have one valid row / column and a clamped source index, where both source rows coincide and the weights do not matter.
// Interpolation is separable (columns, then rows) in packed fp32x2 arithmetic.
struct Blk2x {
    int r0, r1, c0, c1;   // source rows / columns
    int ya, yb, xa, xb;   // output rows 2bi-1, 2bi and columns 2bj-1, 2bj (ya / xa < 0 and yb >= H / xb >= W are skipped)
};
__device__ __forceinline__ Blk2x blk2x(int bi, int bj, int Hk, int Wk) {
    Blk2x b;
    b.r0 = max(bi - 1, 0); b.r1 = min(bi, Hk - 1); b.c0 = max(bj - 1, 0); b.c1 = min(bj, Wk - 1);
    b.ya = 2 * bi - 1; b.yb = 2 * bi; b.xa = 2 * bj - 1; b.xb = 2 * bj;
    return b;
}
// four source vectors (8 channels at `c`) -> the four interpolated vectors of the block: o[0] = (ya, xa), o[1] = (ya, xb), o[2] = (yb, xa), o[3] = (yb, xb)
__device__ __forceinline__ void load2x8(const __nv_bfloat16* __restrict__ t, long pitch, long img0, int Wk, const Blk2x& b, int c, uint4 (&u)[4]) {
    u[0] = __ldg(reinterpret_cast<const uint4*>(t + (img0 + (long)b.r0 * Wk + b.c0) * pitch + c));
    u[1] = __ldg(reinterpret_cast<const uint4*>(t + (img0 + (long)b.r0 * Wk + b.c1) * pitch + c));
    u[2] = __ldg(reinterpret_cast<const uint4*>(t + (img0 + (long)b.r1 * Wk + b.c0) * pitch + c));
    u[3] = __ldg(reinterpret_cast<const uint4*>(t + (img0 + (long)b.r1 * Wk + b.c1) * pitch + c));
}
__device__ __forceinline__ void interp2x8(const uint4 (&u)[4], float (&o)[4][8]) {
    float s00[8], s01[8], s10[8], s11[8];
    unpack8(u[0], s00);
    unpack8(u[1], s01);
    unpack8(u[2], s10);
    unpack8(u[3], s11);
    const float2 w75 = make_float2(0.75f, 0.75f), w25 = make_float2(0.25f, 0.25f);
#pragma unroll
    for (int j = 0; j < 8; j += 2) {
        const float2 a = make_float2(s00[j], s00[j + 1]), bb = make_float2(s01[j], s01[j + 1]);
        const float2 cc = make_float2(s10[j], s10[j + 1]), d = make_float2(s11[j], s11[j + 1]);
        const float2 ta = ffma2(w75, a, fmul2(w25, bb)), tb = ffma2(w25, a, fmul2(w75, bb));    // row r0 at columns xa, xb
        const float2 ua = ffma2(w75, cc, fmul2(w25, d)), ub = ffma2(w25, cc, fmul2(w75, d));    // row r1 at columns xa, xb
        const float2 o0 = ffma2(w75, ta, fmul2(w25, ua)), o1 = ffma2(w75, tb, fmul2(w25, ub));  // output row ya
        const float2 o2 = ffma2(w25, ta, fmul2(w75, ua)), o3 = ffma2(w25, tb, fmul2(w75, ub));  // output row yb
        o[0][j] = o0.x; o[0][j + 1] = o0.y; o[1][j] = o1.x; o[1][j + 1] = o1.y;
        o[2][j] = o2.x; o[2][j + 1] = o2.y; o[3][j] = o3.x; o[3][j + 1] = o3.y;
    }
}
// one CTA per block row bi of an image: scores of output rows 2bi-1, 2bi staged in shared memory ([2][heads][W]), coalesced stores
__global__ void __launch_bounds__(256, 4)
cla_score2x_kernel(const __nv_bfloat16* __restrict__ q, long qpitch, const __nv_bfloat16* __restrict__ k, long kpitch,
                   int N, int H, int W, int Hk, int Wk, int Cq, int heads, float scale, float* __restrict__ s) {
    extern __shared__ float srow[];  // [2][heads][W]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const int n = blockIdx.x / (Hk + 1), bi = blockIdx.x - n * (Hk + 1);
    const int cph = Cq / heads, lph = cph >> 3;
    const long kimg = (long)n * Hk * Wk, qimg = (long)n * H * W;
    for (int bj = warp; bj <= Wk; bj += nwarp) {
        const Blk2x b = blk2x(bi, bj, Hk, Wk);
        const bool va = b.ya >= 0, vb = b.yb < H, ua = b.xa >= 0, ub = b.xb < W;
        for (int cbase = 0; cbase < Cq; cbase += 256) {
            const int c = cbase + lane * 8;
            float part[4] = {0.f, 0.f, 0.f, 0.f};
            if (c < Cq) {
                const int ys[4] = {b.ya, b.ya, b.yb, b.yb}, xs[4] = {b.xa, b.xb, b.xa, b.xb};
                const bool ok[4] = {va && ua, va && ub, vb && ua, vb && ub};
                uint4 ku[4], qu[4];   // all eight loads of the block are in flight before the first is used
                load2x8(k, kpitch, kimg, Wk, b, c, ku);
#pragma unroll
                for (int t = 0; t < 4; ++t)
                    qu[t] = ok[t] ? __ldg(reinterpret_cast<const uint4*>(q + (qimg + (long)ys[t] * W + xs[t]) * qpitch + c)) : make_uint4(0u, 0u, 0u, 0u);
                float kv[4][8];
                interp2x8(ku, kv);
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    float qv[8];
                    unpack8(qu[t], qv);
#pragma unroll
                    for (int j = 0; j < 8; ++j) part[t] += qv[j] * kv[t][j];   // (zeros where the output pixel does not exist)
                }
            }
            for (int o = lph >> 1; o > 0; o >>= 1) {
#pragma unroll
                for (int t = 0; t < 4; ++t) part[t] += __shfl_xor_sync(0xffffffffu, part[t], o);
            }
            if (c < Cq && (lane & (lph - 1)) == 0) {
                const int g = c / cph;
                if (va && ua) srow[(0 * heads + g) * W + b.xa] = part[0] * scale;
                if (va && ub) srow[(0 * heads + g) * W + b.xb] = part[1] * scale;
                if (vb && ua) srow[(1 * heads + g) * W + b.xa] = part[2] * scale;
                if (vb && ub) srow[(1 * heads + g) * W + b.xb] = part[3] * scale;
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * heads * W; i += blockDim.x) {
        const int r = i / (heads * W), rem = i - r * heads * W, g = rem / W, x = rem - g * W;
        const int y = 2 * bi - 1 + r;
        if (y >= 0 && y < H) s[(((long)n * heads + g) * H + y) * W + x] = srow[i];
    }
}
__global__ void __launch_bounds__(256)
cla_apply2x_kernel(const float* __restrict__ s, const float* __restrict__ st, const __nv_bfloat16* __restrict__ v, long vpitch,
                   int N, int H, int W, int Hk, int Wk, int Cv, int heads, float r2, __nv_bfloat16* __restrict__ o, long opitch) {
    extern __shared__ float arow[];  // [2][heads][W]: attention weights of output rows 2bi-1, 2bi
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const int n = blockIdx.x / (Hk + 1), bi = blockIdx.x - n * (Hk + 1);
    for (int i = threadIdx.x; i < 2 * heads * W; i += blockDim.x) {
        const int r = i / (heads * W), rem = i - r * heads * W, g = rem / W, x = rem - g * W;
        const int y = 2 * bi - 1 + r;
        float a = 0.f;
        if (y >= 0 && y < H) {
            const long ng = (long)n * heads + g;
            const float2 ms = *reinterpret_cast<const float2*>(st + (ng * W + x) * 2);
            a = r2 * expf(s[(ng * H + y) * W + x] - ms.x) * ms.y;
        }
        arow[i] = a;
    }
    __syncthreads();
    const int cph = Cv / heads;
    const long vimg = (long)n * Hk * Wk, oimg = (long)n * H * W;
    for (int bj = warp; bj <= Wk; bj += nwarp) {
        const Blk2x b = blk2x(bi, bj, Hk, Wk);
        const bool va = b.ya >= 0, vb = b.yb < H, ua = b.xa >= 0, ub = b.xb < W;
        for (int c = lane * 8; c < Cv; c += 256) {
            float ov[4][8];
            uint4 vu[4];
            load2x8(v, vpitch, vimg, Wk, b, c, vu);
            interp2x8(vu, ov);
            const int g = c / cph;
            const int ys[4] = {b.ya, b.ya, b.yb, b.yb}, xs[4] = {b.xa, b.xb, b.xa, b.xb};
            const bool ok[4] = {va && ua, va && ub, vb && ua, vb && ub};
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                if (ok[t]) {
                    const float w = arow[((t >> 1) * heads + g) * W + xs[t]];
#pragma unroll
                    for (int j = 0; j < 8; ++j) ov[t][j] *= w;
                    *reinterpret_cast<uint4*>(o + (oimg + (long)ys[t] * W + xs[t]) * opitch + c) = pack8(ov[t]);
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// LayerNorm over channels: one warp per token, VPL 16-byte vectors per lane (C = 256 * VPL, or less
// with idle lanes), TPI tokens in flight per warp so that 4 independent loads per lane are
// outstanding (the one-token version was latency-bound at 1.5 TB/s).
// ---------------------------------------------------------------------------------------------
template <int VPL, int TPI>
__global__ void __launch_bounds__(256, VPL <= 2 ? 4 : 2)
layernorm_kernel(const __nv_bfloat16* __restrict__ x, long xpitch, const float* __restrict__ gamma,
                 const float* __restrict__ beta, float eps, long ntok, int C, __nv_bfloat16* __restrict__ y, long ypitch) {
    const int lane = threadIdx.x & 31;
    const long wid = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5;
    const long nw = ((long)gridDim.x * blockDim.x) >> 5;
    const float invC = 1.0f / (float)C;
    for (long t0 = wid * TPI; t0 < ntok; t0 += nw * TPI) {
        uint4 u[TPI][VPL];
#pragma unroll
        for (int k = 0; k < TPI; ++k)
#pragma unroll
            for (int i = 0; i < VPL; ++i) {
                const int c = lane * 8 + i * 256;
                u[k][i] = (t0 + k < ntok && c < C) ? *reinterpret_cast<const uint4*>(x + (t0 + k) * xpitch + c) : make_uint4(0u, 0u, 0u, 0u);
            }
#pragma unroll
        for (int k = 0; k < TPI; ++k) {
            if (t0 + k >= ntok) break;
            float f[VPL][8];
            float s = 0.f;
#pragma unroll
            for (int i = 0; i < VPL; ++i) {
                unpack8(u[k][i], f[i]);
#pragma unroll
                for (int j = 0; j < 8; ++j) s += f[i][j];  // lanes beyond C hold zeros
            }
            const float mean = warp_sum(s) * invC;
            float v = 0.f;
#pragma unroll
            for (int i = 0; i < VPL; ++i) {
                if (lane * 8 + i * 256 < C) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) { const float d = f[i][j] - mean; v += d * d; }
                }
            }
            const float rstd = rsqrtf(warp_sum(v) * invC + eps);
#pragma unroll
            for (int i = 0; i < VPL; ++i) {
                const int c = lane * 8 + i * 256;
                if (c < C) {
                    const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + c)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + c + 4));
                    const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + c)), b1 = __ldg(reinterpret_cast<const float4*>(beta + c + 4));
                    float o[8];
                    o[0] = (f[i][0] - mean) * rstd * g0.x + b0.x; o[1] = (f[i][1] - mean) * rstd * g0.y + b0.y;
                    o[2] = (f[i][2] - mean) * rstd * g0.z + b0.z; o[3] = (f[i][3] - mean) * rstd * g0.w + b0.w;
                    o[4] = (f[i][4] - mean) * rstd * g1.x + b1.x; o[5] = (f[i][5] - mean) * rstd * g1.y + b1.y;
                    o[6] = (f[i][6] - mean) * rstd * g1.z + b1.z; o[7] = (f[i][7] - mean) * rstd * g1.w + b1.w;
                    *reinterpret_cast<uint4*>(y + (t0 + k) * ypitch + c) = pack8(o);
                }
            }
        }
    }
}

static int grid_for(long work_items, int per_block) {
    long g = (work_items + per_block - 1) / per_block;
    long cap = (long)num_sms() * 16;
    return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}
static bool view_ok_bf16(const skb_view* v) {
    return v && v->ptr && v->dtype == SKB_BF16 && v->c % 8 == 0 && v->pitch % 8 == 0 && v->c <= v->pitch && ((uintptr_t)v->ptr & 15) == 0 &&
           v->n > 0 && v->h > 0 && v->w > 0;
}

}  // namespace skb

using namespace skb;

extern "C" int skb_focus_nchw_f32(const float* img, int32_t n, int32_t h, int32_t w, const skb_view* y, void* stream) {
    int rc = check_device();
    if (rc != SKB_OK) return rc;
    SKB_REQUIRE(img && view_ok_bf16(y), SKB_ERR_ARG, "focus: bad arguments");
    SKB_REQUIRE(h % 2 == 0 && w % 2 == 0 && ((uintptr_t)img & 7) == 0, SKB_ERR_ARG, "focus: H, W must be even (got %dx%d)", h, w);
    SKB_REQUIRE(y->n == n && y->h == h / 2 && y->w == w / 2 && y->c >= 16, SKB_ERR_ARG, "focus: output view mismatch");
    const long total = (long)n * (h / 2) * (w / 2);
    focus_kernel<float><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(img, n, h, w, (__nv_bfloat16*)y->ptr, y->pitch, y->c);
    SKB_LAUNCH_CHECK();
    return SKB_OK;
}

extern "C" int skb_focus_nchw_u8(const uint8_t* img, int32_t n, int32_t h, int32_t w, const skb_view* y, void* stream) {
    int rc = check_device();
    if (rc != SKB_OK) return rc;
    SKB_REQUIRE(img && view_ok_bf16(y), SKB_ERR_ARG, "focus_u8: bad arguments");
    SKB_REQUIRE(h % 2 == 0 && w % 2 == 0 && ((uintptr_t)img & 1) == 0, SKB_ERR_ARG, "focus_u8: H, W must be even (got %dx%d)", h, w);
    SKB_REQUIRE(y->n == n && y->h == h / 2 && y->w == w / 2 && y->c >= 16, SKB_ERR_ARG, "focus_u8: output view mismatch");
    const long total = (long)n * (h / 2) * (w / 2);
    focus_kernel<uint8_t><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(img, n, h, w, (__nv_bfloat16*)y->ptr, y->pitch, y->c);
    SKB_LAUNCH_CHECK();
    return SKB_OK;
}

namespace skb {
int launch_focus_pad(const void* img, int img_dtype, int n, int h, int w, void* scratch, cudaStream_t st, const int* tiles,
                     int frame_h, int frame_w) {
    const int rows = n * (h / 2);  // one CTA per padded output row
    __nv_bfloat16* y = (__nv_bfloat16*)scratch;
    if (tiles) {
        if (img_dtype == SKB_F32) focus_pad_kernel<float, true><<<rows, 256, 0, st>>>((const float*)img, n, h, w, y, tiles, frame_h, frame_w);
        else focus_pad_kernel<uint8_t, true><<<rows, 256, 0, st>>>((const uint8_t*)img, n, h, w, y, tiles, frame_h, frame_w);
    } else {
        if (img_dtype == SKB_F32) {
            focus_pad_kernel<float, false><<<rows, 256, 0, st>>>((const float*)img, n, h, w, y, nullptr, 0, 0);
        } else if (w % 32 == 0 && ((uintptr_t)img & 15) == 0) {   // 16-byte loads: rows of a multiple of 16 bytes, 8 output pixels per thread
            const long items = (long)rows * (w / 16);
            focus_pad8_u8_kernel<<<grid_for(items, 256), 256, 0, st>>>((const uint8_t*)img, n, h, w, y);
        } else {
            focus_pad_kernel<uint8_t, false><<<rows, 256, 0, st>>>((const uint8_t*)img, n, h, w, y, nullptr, 0, 0);
        }
    }
    SKB_LAUNCH_CHECK();
    return SKB_OK;
}
}  // namespace skb

extern "C" int skb_letterbox_u8(const uint8_t* src, int32_t h0, int32_t w0, int32_t src_pitch, uint8_t* dst, int32_t H, int32_t W,
                                int32_t new_h, int32_t new_w, int32_t top, int32_t left, int32_t pad, void* stream) {
    int rc = check_device();
    if (rc != SKB_OK) return rc;
    SKB_REQUIRE(src && dst && h0 > 0 && w0 > 0 && H > 0 && W > 0 && new_h > 0 && new_w > 0, SKB_ERR_ARG, "letterbox: bad arguments");
    SKB_REQUIRE(src_pitch >= 3 * w0 && top >= 0 && left >= 0 && top + new_h <= H && left + new_w <= W && pad >= 0 && pad <= 255, SKB_ERR_ARG,
                "letterbox: resized image %dx%d at (%d,%d) does not fit %dx%d", new_h, new_w, top, left, H, W);
    const int resize = (new_h != h0 || new_w != w0) ? 1 : 0;
    const long total = (long)H * W;
    letterbox_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(src, h0, w0, src_pitch, dst, H, W, new_h, new_w, top, left, pad,
                                                                            (double)w0 / new_w, (double)h0 / new_h, resize);
    SKB_LAUNCH_CHECK();
    return SKB_OK;
}

extern "C" int skb_maxpool5_bf16(const skb_view* x, const skb_view* y, void* stream) {
    int rc = check_device();
    if (rc != SKB_OK) return rc;
    SKB_REQUIRE(view_ok_bf16(x) && view_ok_bf16(y), SKB_ERR_ARG, "maxpool5: bad view");
    SKB_REQUIRE(x->n == y->n && x->h == y->h && x->w == y->w && x->c == y->c, SKB_ERR_ARG, "maxpool5: shape mismatch");
    const int segs = x->h >= 32 ? 4 : (x->h >= 8 ? 2 : 1);
    const long total = (long)x->n * segs * x->w * (x->c / 8);  // one thread per (image, row segment, column, 8 channels)
    maxpool5_kernel<<<grid_for(total, 128), 128, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x->ptr, x->pitch, (__nv_bfloat16*)y->ptr,
                                                                             y->pitch, x->n, x->h, x->w, x->c / 8, segs);
    SKB_LAUNCH_CHECK();
    return SKB_OK;
}

// the three SPP pools of x (blocks.py:143-149) in one pass; all four views [n, h, w, c] with the same pixel pitch family
extern "C" int skb_spp_pools_bf16(const skb_view* x, const skb_view* y5, const skb_view* y9, const skb_view* y13, void* stream) {
    int rc = check_device();
    if (rc != SKB_OK) return rc;
    SKB_REQUIRE(view_ok_bf16(x) && view_ok_bf16(y5) && view_ok_bf16(y9) && view_ok_bf16(y13), SKB_ERR_ARG, "spp_pools: bad view");
    const skb_view* ys[3] = {y5, y9, y13};
    for (int i = 0; i < 3; ++i)
        SKB_REQUIRE(ys[i]->n == x->n && ys[i]->h == x->h && ys[i]->w == x->w && ys[i]->c == x->c && ys[i]->pitch == y5->pitch, SKB_ERR_ARG,
                    "spp_pools: output %d does not match the input map", i);
    SKB_REQUIRE(x->c % 16 == 0, SKB_ERR_UNSUPPORTED, "spp_pools: C=%d must be a multiple of 16", x->c);
    const size_t smem = (size_t)2 * x->h * x->w * 32;
    SKB_REQUIRE(smem <= 110 * 1024, SKB_ERR_UNSUPPORTED, "spp_pools: %dx%d map needs %zu bytes of shared memory (use skb_maxpool5_bf16 x 3)", x->h, x->w, smem);
    static PerDeviceOnce once;
    if (once.first()) SKB_CUDA(cudaFuncSetAttribute(spp_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024));
    const int groups = x->c / 16;
    spp_fused_kernel<<<x->n * groups, SPP_THREADS, smem, (cudaStream_t)stream>>>((const __nv_bfloat16*)x->ptr, x->pitch, (__nv_bfloat16*)y5->ptr,
                                                                         (__nv_bfloat16*)y9->ptr, (__nv_bfloat16*)y13->ptr, y5->pitch, x->h, x->w, groups);
    SKB_LAUNCH_CHECK();
    return SKB_OK;
}

static int cbam_slabs(int hw) { int s = (hw + 255) / 256; return s < 1 ? 1 : (s > 64 ? 64 : s); }

extern "C" size_t skb_cbam_workspace_bytes(int32_t n, int32_t h, int32_t w, int32_t c) {
    const size_t slabs = cbam_slabs(h * w);
    return sizeof(float) * ((size_t)2 * n * slabs * c + (size_t)n * c + (size_t)2 * n * h * w) + 64;
}

extern "C" int skb_cbam_bf16(const skb_view* x, const float* w0, const float* w1, int32_t reduced, const float* w7, const skb_view* y,
                             void* workspace, size_t workspace_bytes, void* stream) {
    int rc = check_device();
    if (rc != SKB_OK) return rc;
    SKB_REQUIRE(view_ok_bf16(x) && view_ok_bf16(y) && w0 && w1 && w7 && workspace, SKB_ERR_ARG, "cbam: bad arguments");
    SKB_REQUIRE(x->n == y->n && x->h == y->h && x->w == y->w && x->c == y->c, SKB_ERR_ARG, "cbam: shape mismatch");
    const int N = x->n, H = x->h, W = x->w, C = x->c, HW = H * W;
    SKB_REQUIRE(C / 8 <= 256 && reduced >= 1 && reduced <= 256, SKB_ERR_UNSUPPORTED, "cbam: C=%d reduced=%d", C, reduced);
    SKB_REQUIRE(workspace_bytes >= skb_cbam_workspace_bytes(N, H, W, C), SKB_ERR_WORKSPACE, "cbam: workspace too small");
    const int slabs = cbam_slabs(HW);
    float* psum = (float*)workspace;
    float* pmax = psum + (size_t)N * slabs * C;
    float* att = pmax + (size_t)N * slabs * C;
    float* stats = att + (size_t)N * C;
    stats = (float*)(((uintptr_t)stats + 15) & ~(uintptr_t)15);
    cudaStream_t st = (cudaStream_t)stream;
    const int C8 = C / 8;
    const int lanes = 256 / C8;
    cbam_pool_kernel<<<dim3(slabs, N), 256, sizeof(float) * 2 * lanes * C, st>>>((const __nv_bfloat16*)x->ptr, x->pitch, HW, C, slabs, psum, pmax);
    SKB_LAUNCH_CHECK();
    cbam_mlp_kernel<<<N, 1024, sizeof(float) * (2 * C + 2 * reduced), st>>>(psum, pmax, slabs, HW, C, reduced, w0, w1, att);  // latency-bound: one row / channel per thread
    SKB_LAUNCH_CHECK();
    const long npix = (long)N * HW;
    cbam_stats_kernel<<<grid_for(npix, 8), 256, 0, st>>>((const __nv_bfloat16*)x->ptr, x->pitch, att, npix, HW, C, stats);
    SKB_LAUNCH_CHECK();
    cbam_apply_kernel<<<grid_for(npix, 8), 256, 0, st>>>((const __nv_bfloat16*)x->ptr, x->pitch, att, stats, w7, N, H, W, C,
                                                        (__nv_bfloat16*)y->ptr, y->pitch);
    SKB_LAUNCH_CHECK();
    return SKB_OK;
}

extern "C" size_t skb_cla_workspace_bytes(int32_t n, int32_t h, int32_t w, int32_t heads) {
    return sizeof(float) * ((size_t)n * heads * h * w + (size_t)2 * n * heads * w) + 64;
}

extern "C" int skb_cla_core_bf16(const skb_view* q, const skb_view* k, const skb_view* v, const skb_view* o, int32_t heads, float scale,
                                 float r2, void* workspace, size_t workspace_bytes, void* stream) {
    int rc = check_device();
    if (rc != SKB_OK) return rc;
    SKB_REQUIRE(view_ok_bf16(q) && view_ok_bf16(k) && view_ok_bf16(v) && view_ok_bf16(o) && workspace, SKB_ERR_ARG, "cla: bad arguments");
    SKB_REQUIRE(k->c == q->c && v->n == k->n && v->h == k->h && v->w == k->w && k->n == q->n, SKB_ERR_ARG, "cla: q/k/v mismatch");
    SKB_REQUIRE(o->n == q->n && o->h == q->h && o->w == q->w && o->c == v->c, SKB_ERR_ARG, "cla: output mismatch");
    SKB_REQUIRE(heads >= 1 && heads <= 32 && q->c % heads == 0 && v->c % heads == 0 && (q->c / heads) % 8 == 0 && (v->c / heads) % 8 == 0 &&
                    q->c / heads <= 256 && 256 % (q->c / heads) == 0,
                SKB_ERR_UNSUPPORTED, "cla: heads=%d Cq=%d Cv=%d", heads, q->c, v->c);
    const int N = q->n, H = q->h, W = q->w;
    SKB_REQUIRE(workspace_bytes >= skb_cla_workspace_bytes(N, H, W, heads), SKB_ERR_WORKSPACE, "cla: workspace too small");
    float* s = (float*)workspace;
    float* st = s + (size_t)N * heads * H * W;
    st = (float*)(((uintptr_t)st + 15) & ~(uintptr_t)15);
    cudaStream_t cs = (cudaStream_t)stream;
    const long npix = (long)N * H * W;
    const size_t row_sh = sizeof(float) * (size_t)heads * W;
    SKB_REQUIRE(row_sh <= 48 * 1024, SKB_ERR_UNSUPPORTED, "cla: heads*W = %d too large for the row staging buffer", heads * W);
    static int cla2x = -1;  // tuning knob (not part of the ABI): SKB_CLA_2X=0 keeps the generic bilinear kernels
    if (cla2x < 0) { const char* e = getenv("SKB_CLA_2X"); cla2x = e ? atoi(e) : 1; }
    const bool exact2x = cla2x && H == 2 * k->h && W == 2 * k->w && 2 * row_sh <= 48 * 1024;
    if (exact2x) {
        cla_score2x_kernel<<<N * (k->h + 1), 256, 2 * row_sh, cs>>>((const __nv_bfloat16*)q->ptr, q->pitch, (const __nv_bfloat16*)k->ptr, k->pitch,
                                                                    N, H, W, k->h, k->w, q->c, heads, scale, s);
    } else {
        cla_score_kernel<<<N * H, 256, row_sh, cs>>>((const __nv_bfloat16*)q->ptr, q->pitch, (const __nv_bfloat16*)k->ptr, k->pitch, N, H, W,
                                                     k->h, k->w, q->c, heads, scale, s);
    }
    SKB_LAUNCH_CHECK();
    cla_colstat_kernel<<<N * heads * ((W + 31) / 32), 256, 0, cs>>>(s, N * heads, H, W, st);
    SKB_LAUNCH_CHECK();
    if (exact2x) {
        cla_apply2x_kernel<<<N * (k->h + 1), 256, 2 * row_sh, cs>>>(s, st, (const __nv_bfloat16*)v->ptr, v->pitch, N, H, W, v->h, v->w, v->c, heads, r2,
                                                                    (__nv_bfloat16*)o->ptr, o->pitch);
    } else {
        cla_apply_kernel<<<N * H, 256, row_sh, cs>>>(s, st, (const __nv_bfloat16*)v->ptr, v->pitch, N, H, W, v->h, v->w, v->c, heads, r2,
                                                     (__nv_bfloat16*)o->ptr, o->pitch);
    }
    SKB_LAUNCH_CHECK();
    return SKB_OK;
}

extern "C" int skb_layernorm_bf16(const skb_view* x, const float* gamma, const float* beta, float eps, const skb_view* y, void* stream) {
    int rc = check_device();
    if (rc != SKB_OK) return rc;
    SKB_REQUIRE(view_ok_bf16(x) && view_ok_bf16(y) && gamma && beta, SKB_ERR_ARG, "layernorm: bad arguments");
    SKB_REQUIRE(x->n == y->n && x->h == y->h && x->w == y->w && x->c == y->c, SKB_ERR_ARG, "layernorm: shape mismatch");
    SKB_REQUIRE(x->c <= 2048, SKB_ERR_UNSUPPORTED, "layernorm: C=%d > 2048", x->c);
    const long ntok = (long)x->n * x->h * x->w;
    const __nv_bfloat16* xp = (const __nv_bfloat16*)x->ptr;
    __nv_bfloat16* yp = (__nv_bfloat16*)y->ptr;
    cudaStream_t st = (cudaStream_t)stream;
    const int C = x->c;
    if (C <= 256)
        layernorm_kernel<1, 4><<<grid_for(ntok, 32), 256, 0, st>>>(xp, x->pitch, gamma, beta, eps, ntok, C, yp, y->pitch);
    else if (C <= 512)
        layernorm_kernel<2, 2><<<grid_for(ntok, 16), 256, 0, st>>>(xp, x->pitch, gamma, beta, eps, ntok, C, yp, y->pitch);
    else if (C <= 1024)
        layernorm_kernel<4, 1><<<grid_for(ntok, 8), 256, 0, st>>>(xp, x->pitch, gamma, beta, eps, ntok, C, yp, y->pitch);
    else
        layernorm_kernel<8, 1><<<grid_for(ntok, 8), 256, 0, st>>>(xp, x->pitch, gamma, beta, eps, ntok, C, yp, y->pitch);
    SKB_LAUNCH_CHECK();
    return SKB_OK;
}
