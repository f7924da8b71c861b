// NMS kernels.
//
// (1) skb_nms_f32: torchvision.ops.nms semantics (call site skyeye/utils/metrics.py:442), bit-exact
//     with the CPU op: stable descending radix sort, 64x64 IoU bitmask tiles built with explicit
//     round-to-nearest intrinsics (no FMA contraction, true division), sequential greedy reduce.
// (2) skb_nms_batched_f32: the whole wrapper non_max_suppression (metrics.py:361-457) for a batch
//     without host synchronisation.  Candidate rows are emitted as 64-bit sort keys
//     [image | ~score | slot] so one radix sort orders every image by (score desc, row asc) =
//     the reference's boolean-mask order + torchvision's stable sort.  Greedy suppression then
//     only ever compares a candidate with boxes that were KEPT before it, and the wrapper keeps at
//     most max_det of them (metrics.py:443-444), so each image needs <= n*max_det IoU tests instead of
//     n^2/2: one CTA per image walks the sorted candidates in chunks with the kept list in shared
//     memory (warp-ballot compaction, no n x n mask in HBM).
// Ordering is done in the kernels themselves (no library sort): the filter compacts the surviving keys per image, and the
// image's CTA orders them LAZILY in shared memory -- an exact radix SELECT of the next (up to) 8192 best keys, a bitonic sort
// of those, the greedy pass over them, and only if max_det boxes have not been kept yet the next 8192.  A detector's image
// needs one round; the worst case (30 000 candidates, nothing suppressed... all walked) needs four.
#include <float.h>

#include "common.cuh"

namespace skb {

__device__ __forceinline__ float box_area(float x1, float y1, float x2, float y2) { return __fmul_rn(__fsub_rn(x2, x1), __fsub_rn(y2, y1)); }
// iou(i, j) > thr exactly as torchvision's CPU kernel evaluates it
__device__ __forceinline__ bool iou_gt(float ax1, float ay1, float ax2, float ay2, float aarea, float bx1, float by1, float bx2, float by2,
                                       float barea, float thr) {
    const float xx1 = fmaxf(ax1, bx1), yy1 = fmaxf(ay1, by1);
    const float xx2 = fminf(ax2, bx2), yy2 = fminf(ay2, by2);
    const float w = fmaxf(0.0f, __fsub_rn(xx2, xx1)), h = fmaxf(0.0f, __fsub_rn(yy2, yy1));
    const float inter = __fmul_rn(w, h);
    const float ovr = __fdiv_rn(inter, __fsub_rn(__fadd_rn(aarea, barea), inter));
    return ovr > thr;  // NaN (0/0) never suppresses
}

__device__ __forceinline__ unsigned int desc_bits(float s) {
    unsigned int u = __float_as_uint(s);
    u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);  // ascending-orderable
    return ~u;                                       // descending
}

// =============================================================================================
// (0) CTA-wide ordering of unique 64-bit keys that live in global memory (512 threads)
// =============================================================================================
constexpr int NMS_THREADS = 512;
constexpr int SORT_KS = 8192;               // keys ordered per round in shared memory (64 KB)
constexpr int SEL_BITS = 11, SEL_BINS = 1 << SEL_BITS;
constexpr int SEL_SMALL = SEL_BINS / 2;     // a digit bin this small is resolved by rank counting (its keys alias the histogram)
static_assert(SEL_BINS == 4 * NMS_THREADS, "one thread owns four histogram bins");

struct alignas(16) SortScratch {
    unsigned int hist[SEL_BINS];  // aliased as unsigned long long small[SEL_SMALL]
    unsigned int warp_tot[NMS_THREADS / 32];
    int sel_bin, sel_cnt, n_small, n_out;
    unsigned int sel_below;
    unsigned long long result;
};

// f(key) for every key of g[0, n): eight independent coalesced 16-byte loads (two keys each) per thread are in flight before the
// first is used (the one-load-per-iteration form was latency-bound: ~700 cycles per key and thread, 0.4 ms for a 100 000-key
// image).  A segment that starts on an odd key is peeled by one key; an odd tail key is read on its own.
template <typename F>
__device__ __forceinline__ void cta_for_each_key(const unsigned long long* g, int n, F f) {
    constexpr int U = 8;
    if (n > 0 && (reinterpret_cast<uintptr_t>(g) & 8)) {  // segment start on an odd key: peel it
        if (threadIdx.x == 0) f(g[0]);
        ++g;
        --n;
    }
    const int n2 = n >> 1;
    const ulonglong2* g2 = reinterpret_cast<const ulonglong2*>(g);
    for (int base = 0; base < n2; base += NMS_THREADS * U) {
        ulonglong2 kbuf[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int i = base + u * NMS_THREADS + (int)threadIdx.x;
            kbuf[u] = i < n2 ? g2[i] : make_ulonglong2(0ULL, 0ULL);
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (base + u * NMS_THREADS + (int)threadIdx.x < n2) { f(kbuf[u].x); f(kbuf[u].y); }
    }
    if ((n & 1) && threadIdx.x == 0) f(g[n - 1]);
}

// The k-th smallest (k >= 1) of the keys in g[0, n) that are > lo (all keys if !have_lo).  Only the low `total_bits` bits
// differ between the keys of a segment, and their top `common_bits` bits are known to be equal as well (value `common`,
// right-aligned: sign and exponent of scores in (0, 1) -- without this the first digit would hit a handful of bins).  MSD
// radix select: one 11-bit digit per pass over the keys (a histogram in shared memory), until the digit bin that holds the
// answer has <= SEL_SMALL keys; those are ranked directly.  Keys are unique.
__device__ unsigned long long cta_select_kth(const unsigned long long* g, int n, bool have_lo, unsigned long long lo, int k,
                                             int total_bits, int common_bits, unsigned long long common, SortScratch& S) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned long long kmask = total_bits >= 64 ? ~0ULL : (1ULL << total_bits) - 1ULL;
    auto shr = [](unsigned long long v, int sh) { return sh >= 64 ? 0ULL : v >> sh; };
    unsigned long long prefix = common;  // resolved high bits of (key & kmask), right-aligned
    int resolved = common_bits;
    while (true) {
        const int nb = min(SEL_BITS, total_bits - resolved);
        const int shift = total_bits - resolved - nb;
        for (int i = tid; i < SEL_BINS; i += NMS_THREADS) S.hist[i] = 0u;
        __syncthreads();
        cta_for_each_key(g, n, [&](unsigned long long key) {
            const unsigned long long kk = key & kmask;
            if ((!have_lo || key > lo) && shr(kk, shift + nb) == prefix) atomicAdd(&S.hist[(unsigned int)(kk >> shift) & ((1u << nb) - 1u)], 1u);
        });
        __syncthreads();
        {   // the bin that holds the k-th key: thread t owns bins 4t .. 4t+3
            unsigned int c[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) c[q] = S.hist[4 * tid + q];
            const unsigned int mine = c[0] + c[1] + c[2] + c[3];
            unsigned int inc = mine;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned int v = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += v;
            }
            if (lane == 31) S.warp_tot[warp] = inc;
            __syncthreads();
            unsigned int base = 0;
            for (int w = 0; w < warp; ++w) base += S.warp_tot[w];
            unsigned int excl = base + inc - mine;
            if (excl < (unsigned int)k && (unsigned int)k <= excl + mine) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    if ((unsigned int)k <= excl + c[q]) { S.sel_bin = 4 * tid + q; S.sel_below = excl; S.sel_cnt = (int)c[q]; break; }
                    excl += c[q];
                }
            }
            __syncthreads();
        }
        const int cnt = S.sel_cnt;
        k -= (int)S.sel_below;
        prefix = (prefix << nb) | (unsigned long long)S.sel_bin;
        resolved += nb;
        if (resolved == total_bits) return (g[0] & ~kmask) | prefix;  // every bit resolved (the bin holds exactly the key)
        if (cnt <= SEL_SMALL) {
            unsigned long long* small = reinterpret_cast<unsigned long long*>(S.hist);
            __syncthreads();  // everyone has read sel_* and the histogram
            if (tid == 0) S.n_small = 0;
            __syncthreads();
            const int sh = total_bits - resolved;
            cta_for_each_key(g, n, [&](unsigned long long key) {
                if ((!have_lo || key > lo) && shr(key & kmask, sh) == prefix) small[atomicAdd(&S.n_small, 1)] = key;
            });
            __syncthreads();
            for (int t = tid; t < cnt; t += NMS_THREADS) {
                const unsigned long long mk = small[t];
                int r = 0;
                for (int j = 0; j < cnt; ++j) r += small[j] < mk ? 1 : 0;
                if (r == k - 1) S.result = mk;
            }
            __syncthreads();
            return S.result;
        }
        __syncthreads();  // the histogram is rebuilt by the next pass
    }
}

// dst[0, m) <- the keys of g[0, n) in (lo, hi] (order unspecified); returns m to every thread.  At most `cap` keys are stored:
// m > cap tells the caller that the range was too wide.
__device__ int cta_gather_range(const unsigned long long* g, int n, bool have_lo, unsigned long long lo, unsigned long long hi,
                                unsigned long long* dst, int cap, SortScratch& S) {
    if (threadIdx.x == 0) S.n_out = 0;
    __syncthreads();
    cta_for_each_key(g, n, [&](unsigned long long key) {
        if ((!have_lo || key > lo) && key <= hi) {
            const int pos = atomicAdd(&S.n_out, 1);
            if (pos < cap) dst[pos] = key;
        }
    });
    __syncthreads();
    return S.n_out;
}

// A cut `hi` such that (lo, hi] probably holds between 1 and SORT_KS keys, from a histogram of the first undetermined digit
// over a strided SAMPLE of the segment (shared-memory atomics cost ~2 cycles per key: a full histogram of a 100 000-key
// image was 0.12 ms).  The caller counts the range exactly while gathering it and falls back to cta_select_kth if the
// estimate was off (adversarial score distributions: everything inside one digit bin).
__device__ unsigned long long cta_sampled_cut(const unsigned long long* g, int n, bool have_lo, unsigned long long lo, int want,
                                              int total_bits, int common_bits, unsigned long long common, SortScratch& S) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned long long kmask = total_bits >= 64 ? ~0ULL : (1ULL << total_bits) - 1ULL;
    const int nb = min(SEL_BITS, total_bits - common_bits);
    const int shift = total_bits - common_bits - nb;
    const int stride = max(1, n >> 12);  // <= 8192 samples
    for (int i = tid; i < SEL_BINS; i += NMS_THREADS) S.hist[i] = 0u;
    __syncthreads();
    for (int j0 = tid; (long)j0 * stride < n; j0 += 4 * NMS_THREADS) {  // four sample loads in flight per thread
        unsigned long long kb[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const long idx = (long)(j0 + u * NMS_THREADS) * stride;
            kb[u] = idx < n ? g[idx] : 0ULL;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const unsigned long long key = kb[u];
            if ((long)(j0 + u * NMS_THREADS) * stride < n && (!have_lo || key > lo))
                atomicAdd(&S.hist[(unsigned int)((key & kmask) >> shift) & ((1u << nb) - 1u)], 1u);
        }
    }
    __syncthreads();
    const unsigned int target = (unsigned int)max(1, (want - want / 4) / stride);  // aim at 3/4 of the capacity
    unsigned int c[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) c[q] = S.hist[4 * tid + q];
    const unsigned int mine = c[0] + c[1] + c[2] + c[3];
    unsigned int inc = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned int v = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += v;
    }
    if (lane == 31) S.warp_tot[warp] = inc;
    if (tid == 0) S.sel_bin = SEL_BINS - 1;  // fewer samples than the target: take everything that is left
    __syncthreads();
    unsigned int base = 0;
    for (int w = 0; w < warp; ++w) base += S.warp_tot[w];
    unsigned int excl = base + inc - mine;
    if (excl < target && target <= excl + mine) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            if (target <= excl + c[q]) { S.sel_bin = 4 * tid + q; break; }
            excl += c[q];
        }
    }
    __syncthreads();
    const unsigned long long bin = (unsigned long long)S.sel_bin;
    __syncthreads();
    if (bin >= (1ULL << nb) - 1ULL) return ~0ULL;
    const unsigned long long low_ones = shift ? ((1ULL << shift) - 1ULL) : 0ULL;
    return (g[0] & ~kmask) | (((common << nb) | bin) << shift) | low_ones;  // upper edge of the digit bin
}

// ascending bitonic sort of a[0, m) in shared memory (padded to a power of two with all-ones keys, which sort last)
__device__ void cta_bitonic_sort(unsigned long long* a, int m) {
    int P = 2;
    while (P < m) P <<= 1;
    for (int i = m + threadIdx.x; i < P; i += NMS_THREADS) a[i] = ~0ULL;
    __syncthreads();
    for (int k = 2; k <= P; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int idx = threadIdx.x; idx < (P >> 1); idx += NMS_THREADS) {
                const int i = ((idx & ~(j - 1)) << 1) | (idx & (j - 1));
                const int l = i | j;
                const unsigned long long x = a[i], y = a[l];
                if ((x > y) == ((i & k) == 0)) { a[i] = y; a[l] = x; }
            }
            __syncthreads();
        }
}

// The next batch of a segment in ascending key order: the smallest m keys greater than `lo` (1 <= m <= SORT_KS, about `want` of
// them when more than `want` are left) are gathered into skeys and sorted.  Returns m (0 when nothing is left).
// `remaining` = keys of the segment that are > lo; want <= SORT_KS.
__device__ int cta_next_sorted_batch(const unsigned long long* g, int n, bool have_lo, unsigned long long lo, int remaining,
                                     int want, int total_bits, int common_bits, unsigned long long common, unsigned long long* skeys,
                                     SortScratch& S) {
    if (remaining <= 0 || want <= 0) return 0;
    int m;
    if (remaining <= want) {
        m = cta_gather_range(g, n, have_lo, lo, ~0ULL, skeys, SORT_KS, S);
    } else {
        // any non-empty prefix of the order that fits the buffer will do: cut at a sampled digit boundary, count exactly
        // while gathering, and only if that misses (empty or too many) pay for the exact selection of `want` keys
        const unsigned long long cut = cta_sampled_cut(g, n, have_lo, lo, want, total_bits, common_bits, common, S);
        m = cta_gather_range(g, n, have_lo, lo, cut, skeys, SORT_KS, S);
        if (m == 0 || m > SORT_KS) {
            const unsigned long long hi = cta_select_kth(g, n, have_lo, lo, want, total_bits, common_bits, common, S);
            m = cta_gather_range(g, n, have_lo, lo, hi, skeys, SORT_KS, S);
        }
    }
    cta_bitonic_sort(skeys, m);
    return m;
}

// =============================================================================================
// (1) generic bit-exact NMS
// =============================================================================================
// One CTA orders the boxes by (score descending, index ascending) = torchvision's stable descending sort: keys
// [~score | index] are written to scratch, then consumed batch by batch in ascending order; the batch's original indices and
// boxes go straight to `order` / `sboxes` (no separate iota / sort / gather launches).
__global__ void __launch_bounds__(NMS_THREADS)
nms_order_kernel(const float4* __restrict__ boxes, const float* __restrict__ scores, int n, unsigned long long* __restrict__ keys,
                 int* __restrict__ order, float4* __restrict__ sboxes) {
    extern __shared__ __align__(16) uint8_t sort_smem[];
    unsigned long long* skeys = reinterpret_cast<unsigned long long*>(sort_smem);
    SortScratch& S = *reinterpret_cast<SortScratch*>(sort_smem + (size_t)SORT_KS * 8);
    unsigned int v_or = 0u, v_and = ~0u;
    for (int i = threadIdx.x; i < n; i += NMS_THREADS) {
        const unsigned int d = desc_bits(scores[i]);
        v_or |= d;
        v_and &= d;
        keys[i] = ((unsigned long long)d << 32) | (unsigned int)i;
    }
    v_or = __reduce_or_sync(0xffffffffu, v_or);
    v_and = __reduce_and_sync(0xffffffffu, v_and);
    if ((threadIdx.x & 31) == 0) { S.hist[threadIdx.x >> 5] = v_or; S.hist[32 + (threadIdx.x >> 5)] = v_and; }
    __syncthreads();
    for (int w = 0; w < NMS_THREADS / 32; ++w) { v_or |= S.hist[w]; v_and &= S.hist[32 + w]; }
    __syncthreads();
    const int common_bits = __clz((int)(v_or ^ v_and));  // leading score bits every key shares (32 if all scores are equal)
    const unsigned long long common = common_bits ? (unsigned long long)(v_or >> (32 - common_bits)) : 0ULL;
    int consumed = 0;
    bool have_lo = false;
    unsigned long long lo = 0;
    while (consumed < n) {
        const int m = cta_next_sorted_batch(keys, n, have_lo, lo, n - consumed, SORT_KS, 64, common_bits, common, skeys, S);
        for (int i = threadIdx.x; i < m; i += NMS_THREADS) {
            const int idx = (int)(skeys[i] & 0xffffffffULL);
            order[consumed + i] = idx;
            sboxes[consumed + i] = boxes[idx];
        }
        lo = skeys[m - 1];
        have_lo = true;
        consumed += m;
        __syncthreads();
    }
}
// mask[i][cb] bit j: iou(sorted i, sorted cb*64 + j) > thr, only for column blocks >= row block
__global__ void nms_mask_kernel(const float4* __restrict__ sb, int n, float thr, int col_blocks, unsigned long long* __restrict__ mask) {
    const int rb = blockIdx.y, cb = blockIdx.x;
    if (cb < rb) return;
    __shared__ float4 cbox[64];
    const int cn = min(64, n - cb * 64);
    if (threadIdx.x < cn) cbox[threadIdx.x] = sb[cb * 64 + threadIdx.x];
    __syncthreads();
    const int i = rb * 64 + threadIdx.x;
    if (i >= n) return;
    const float4 a = sb[i];
    const float aarea = box_area(a.x, a.y, a.z, a.w);
    unsigned long long bits = 0;
    const int j0 = (rb == cb) ? threadIdx.x + 1 : 0;
    for (int j = j0; j < cn; ++j) {
        const float4 b = cbox[j];
        if (iou_gt(a.x, a.y, a.z, a.w, aarea, b.x, b.y, b.z, b.w, box_area(b.x, b.y, b.z, b.w), thr)) bits |= 1ULL << j;
    }
    mask[(size_t)i * col_blocks + cb] = bits;
}
// sequential greedy reduce over the mask (single CTA); writes kept ORIGINAL indices in score order
__global__ void nms_reduce_kernel(const unsigned long long* __restrict__ mask, const int* __restrict__ order, int n, int col_blocks,
                                  long long* __restrict__ keep, int* __restrict__ n_keep) {
    extern __shared__ unsigned long long remv[];
    for (int j = threadIdx.x; j < col_blocks; j += blockDim.x) remv[j] = 0ULL;
    __shared__ int cnt;
    if (threadIdx.x == 0) cnt = 0;
    __syncthreads();
    for (int i = 0; i < n; ++i) {
        const int nb = i >> 6, ib = i & 63;
        const bool alive = !((remv[nb] >> ib) & 1ULL);
        __syncthreads();  // everyone has read remv[nb] before it is updated
        if (alive) {
            if (threadIdx.x == 0) keep[cnt++] = (long long)order[i];
            const unsigned long long* row = mask + (size_t)i * col_blocks;
            for (int j = nb + threadIdx.x; j < col_blocks; j += blockDim.x) remv[j] |= row[j];
            __syncthreads();
        }
    }
    if (threadIdx.x == 0) *n_keep = cnt;
}

struct NmsWs {
    unsigned long long* keys;
    int* idx_out;
    float4* sboxes;
    unsigned long long* mask;
};
static size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }
static size_t nms_layout(int n, void* base, NmsWs* ws) {
    const int cbk = (n + 63) / 64;
    size_t off = 0;
    uint8_t* b = (uint8_t*)base;
    auto take = [&](size_t bytes) { size_t o = off; off += align256(bytes); return b ? (void*)(b + o) : nullptr; };
    void* k = take(sizeof(unsigned long long) * n);
    void* i1 = take(sizeof(int) * n);
    void* sb = take(sizeof(float4) * n);
    void* m = take(sizeof(unsigned long long) * (size_t)n * cbk);
    if (ws) { ws->keys = (unsigned long long*)k; ws->idx_out = (int*)i1; ws->sboxes = (float4*)sb; ws->mask = (unsigned long long*)m; }
    return off + 256;
}

// =============================================================================================
// (2) batched wrapper
// =============================================================================================
// sort key = [image | ~score (32 bits) | slot]: the slot field is sized per call for N * nc (nc = 80 at 1280^2 needs 23 bits),
// the image field takes what is left of the upper 32 bits
constexpr int KEY_MAX_SLOT_BITS = 27;
constexpr int MAX_NMS_BOXES = 30000;  // metrics.py:393
constexpr int NMS_MAX_KEEP = 1024;

struct FilterParams {
    const float* pred;
    int B, N, nc, no;
    float conf;
    int multi_label, compat;
    int slot_bits;
    int n_classes;
    float classes[32];
};

__device__ __forceinline__ bool class_pass(const FilterParams& p, float col5) {
    if (p.n_classes <= 0) return true;
    for (int i = 0; i < p.n_classes; ++i)
        if (col5 == p.classes[i]) return true;
    return false;
}

// One thread per (image, box).  Surviving rows are appended to their image's key segment keys[b * stride ..) (stride = the
// segment capacity N * keys-per-box) with warp-aggregated position atomics: the order inside a segment is arbitrary, the
// consumer orders the (unique) keys.  key = [~score (32 bits) | slot]: ascending key = (score descending, row ascending) = the
// reference's boolean-mask order + torchvision's stable sort.
__global__ void nms_filter_kernel(const FilterParams p, unsigned long long* __restrict__ keys, int* __restrict__ count, long stride,
                                  unsigned int* __restrict__ bits_or, unsigned int* __restrict__ bits_nor) {
    const long nbox = (long)p.B * p.N;
    const long nround = (nbox + blockDim.x - 1) / blockDim.x * blockDim.x;  // whole warps stay in the loop together
    const int lane = threadIdx.x & 31;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < nround; i += (long)gridDim.x * blockDim.x) {
        const bool inb = i < nbox;
        const int b = inb ? (int)(i / p.N) : -1;
        const int b0 = __shfl_sync(0xffffffffu, b, 0);
        const bool uni = __all_sync(0xffffffffu, b == b0);
        // whole block inside one image (all but the few blocks that straddle an image boundary): ONE position atomic per block
        const long i_first = i - threadIdx.x;
        const bool blk_uni = !p.multi_label && i_first + blockDim.x <= nbox && i_first / p.N == (i_first + blockDim.x - 1) / p.N;
        // called by the whole warp: lanes with ok append `key` to the segment of their image
        // (also accumulates, per image, the OR of the emitted score bits and of their complements: the bits all scores share)
        auto emit = [&](bool ok, unsigned long long key) {
            const unsigned int sb = (unsigned int)(key >> p.slot_bits);
            if (blk_uni) {  // (a contended same-address atomic WITH a return value per warp cost 0.09 ms per step)
                __shared__ int w_cnt[32];
                __shared__ unsigned int w_or_s[32], w_nor_s[32];
                __shared__ int blk_base;
                const unsigned int m = __ballot_sync(0xffffffffu, ok);
                const unsigned int w_or = __reduce_or_sync(0xffffffffu, ok ? sb : 0u), w_nor = __reduce_or_sync(0xffffffffu, ok ? ~sb : 0u);
                const int wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
                if (lane == 0) { w_cnt[wid] = __popc(m); w_or_s[wid] = w_or; w_nor_s[wid] = w_nor; }
                __syncthreads();
                if (threadIdx.x == 0) {
                    int tot = 0;
                    unsigned int bo = 0, bn = 0;
                    for (int w = 0; w < nw; ++w) { tot += w_cnt[w]; bo |= w_or_s[w]; bn |= w_nor_s[w]; }
                    blk_base = tot ? atomicAdd(count + b0, tot) : 0;
                    if (bo & ~__ldcg(bits_or + b0)) atomicOr(bits_or + b0, bo);
                    if (bn & ~__ldcg(bits_nor + b0)) atomicOr(bits_nor + b0, bn);
                }
                __syncthreads();
                if (ok) {
                    int off = blk_base;
                    for (int w = 0; w < wid; ++w) off += w_cnt[w];
                    keys[(long)b0 * stride + off + __popc(m & ((1u << lane) - 1u))] = key;
                }
                __syncthreads();  // the staging arrays are reused by the next grid-stride iteration
            } else if (uni) {
                const unsigned int m = __ballot_sync(0xffffffffu, ok);
                if (m) {
                    const unsigned int w_or = __reduce_or_sync(0xffffffffu, ok ? sb : 0u), w_nor = __reduce_or_sync(0xffffffffu, ok ? ~sb : 0u);
                    int base = 0;
                    if (lane == 0) {
                        base = atomicAdd(count + b0, __popc(m));
                        // the two masks saturate after a few warps: look before paying for the atomic (a stale read only costs a redundant one)
                        if (w_or & ~__ldcg(bits_or + b0)) atomicOr(bits_or + b0, w_or);
                        if (w_nor & ~__ldcg(bits_nor + b0)) atomicOr(bits_nor + b0, w_nor);
                    }
                    base = __shfl_sync(0xffffffffu, base, 0);
                    if (ok) keys[(long)b0 * stride + base + __popc(m & ((1u << lane) - 1u))] = key;
                }
            } else if (ok) {
                keys[(long)b * stride + atomicAdd(count + b, 1)] = key;
                atomicOr(bits_or + b, sb);
                atomicOr(bits_nor + b, ~sb);
            }
        };
        const int r = inb ? (int)(i % p.N) : 0;
        const float* x = p.pred + (inb ? i : 0) * p.no;
        const float obj = inb ? x[4] : 0.0f;
        const bool pass = inb && obj > p.conf;  // metrics.py:391,402
        auto key_of = [&](float score, int slot) { return ((unsigned long long)desc_bits(score) << p.slot_bits) | (unsigned long long)slot; };
        if (p.nc > 1 || (p.compat == 1 && p.nc == 1)) {
            if (p.multi_label) {  // metrics.py:407-410: one row per (box, class) above the threshold
                for (int j = 0; j < p.nc; ++j) {
                    bool ok = false;
                    unsigned long long k = 0;
                    if (pass) {
                        const float cp = x[5 + j];
                        const float conf = p.compat ? __fmul_rn(cp, obj) : cp;
                        if (conf > p.conf && class_pass(p, p.compat ? (float)j : cp)) { k = key_of(p.compat ? conf : obj, r * p.nc + j); ok = true; }
                    }
                    emit(ok, k);
                }
            } else {  // metrics.py:411-414: best class (first maximum)
                bool ok = false;
                unsigned long long k = 0;
                if (pass) {
                    float best = x[5];
                    int bj = 0;
                    if (p.compat) best = __fmul_rn(best, obj);
                    for (int j = 1; j < p.nc; ++j) {
                        float cp = x[5 + j];
                        if (p.compat) cp = __fmul_rn(cp, obj);
                        if (cp > best) { best = cp; bj = j; }
                    }
                    if (best > p.conf && class_pass(p, p.compat ? (float)bj : best)) { k = key_of(p.compat ? best : obj, r * p.nc + bj); ok = true; }
                }
                emit(ok, k);
            }
        } else {  // nc == 1 (metrics.py:415-419): row [cx,cy,w,h,obj,0]; nc == 0 in fixed mode
            const bool ok = pass && class_pass(p, 0.0f);
            emit(ok, ok ? key_of(obj, r) : 0ULL);
        }
    }
}

struct BatchedParams {
    const float* pred;
    int B, N, nc, no;
    float iou;
    int agnostic, multi_label, compat, max_det;
    int capacity;
    int slot_bits;
    // tiled inference: kept rows are shifted by the tile origin into frame coordinates and written, zero
    // padded, with the count in an extra trailing row, straight into the all_gather send buffer
    const int* tile_xy;  // [B][2] = (x0, y0) per image, or null
    int out_rows;        // rows per image in `out`: max_det, or max_det + 1 (row max_det = [count, 0, ...])
};

struct Cand {
    float x1, y1, x2, y2, area;
    float row[7];
};

__device__ __forceinline__ void load_cand(const BatchedParams& p, int b, unsigned long long key, Cand& c) {
    const int slot = (int)(key & ((1ULL << p.slot_bits) - 1));
    int r, j;
    const bool per_class = p.nc > 1 || (p.compat == 1 && p.nc == 1);
    if (per_class) { r = slot / p.nc; j = slot - r * p.nc; } else { r = slot; j = 0; }
    const float* x = p.pred + ((long)b * p.N + r) * p.no;
    const float cx = x[0], cy = x[1], w = x[2], h = x[3], obj = x[4];
    if (p.compat == 0) {
        const float col5 = p.nc > 1 ? x[5 + j] : 0.0f;
        c.row[0] = cx; c.row[1] = cy; c.row[2] = w; c.row[3] = h; c.row[4] = obj; c.row[5] = col5; c.row[6] = (float)j;
        // quirk X8: offset = column 5 (class PROBABILITY) * 4096, boxes = raw (cx,cy,w,h) + offset (metrics.py:435-436)
        const float off = p.agnostic ? 0.0f : __fmul_rn(col5, 4096.0f);
        c.x1 = __fadd_rn(cx, off); c.y1 = __fadd_rn(cy, off); c.x2 = __fadd_rn(w, off); c.y2 = __fadd_rn(h, off);
    } else {
        const float hw = __fdiv_rn(w, 2.0f), hh = __fdiv_rn(h, 2.0f);
        const float bx1 = __fsub_rn(cx, hw), by1 = __fsub_rn(cy, hh), bx2 = __fadd_rn(cx, hw), by2 = __fadd_rn(cy, hh);
        const float conf = per_class ? __fmul_rn(x[5 + j], obj) : obj;
        c.row[0] = bx1; c.row[1] = by1; c.row[2] = bx2; c.row[3] = by2; c.row[4] = conf; c.row[5] = (float)j; c.row[6] = 0.0f;
        const float off = p.agnostic ? 0.0f : __fmul_rn((float)j, 4096.0f);
        c.x1 = __fadd_rn(bx1, off); c.y1 = __fadd_rn(by1, off); c.x2 = __fadd_rn(bx2, off); c.y2 = __fadd_rn(by2, off);
    }
    c.area = box_area(c.x1, c.y1, c.x2, c.y2);
}

// Diagnostics (scripts/bench_nms_stress.py): when set, the COUNT variant of the kept-list kernel adds the number of IoU pair
// tests it performed (phase A: candidate vs kept box; phase B1: survivor vs earlier survivor of the chunk) to this counter.
static unsigned long long* g_nms_pair_counter = nullptr;

// grid = B images, 512 threads.  The (unordered) keys of image b are keys[b * stride, b * stride + count[b]); they are consumed
// in ascending order, up to SORT_KS per round (cta_next_sorted_batch), until max_det boxes are kept or 30 000 were walked.
template <bool COUNT>
__global__ void __launch_bounds__(NMS_THREADS)
nms_keptlist_kernel(const BatchedParams p, const unsigned long long* __restrict__ keys, long stride, const int* __restrict__ count,
                    const unsigned int* __restrict__ bits_or, const unsigned int* __restrict__ bits_nor, float* __restrict__ out,
                    int* __restrict__ out_count, unsigned long long* __restrict__ pair_counter) {
    __shared__ float kx1[NMS_MAX_KEEP], ky1[NMS_MAX_KEEP], kx2[NMS_MAX_KEEP], ky2[NMS_MAX_KEEP], kar[NMS_MAX_KEEP];
    __shared__ int surv[NMS_THREADS];   // chunk-local ids of phase-A survivors, in order
    __shared__ int warp_cnt[NMS_THREADS / 32];
    __shared__ int s_nsurv, s_kept;
    __shared__ float sx1[NMS_THREADS], sy1[NMS_THREADS], sx2[NMS_THREADS], sy2[NMS_THREADS], sar[NMS_THREADS];  // survivor boxes, compacted
    __shared__ float crow[NMS_THREADS][7];
    extern __shared__ __align__(16) unsigned int smask[];  // [NMS_THREADS][NMS_THREADS / 32] pairwise suppression bits of a chunk (32 KB)
    unsigned long long* skeys = reinterpret_cast<unsigned long long*>(smask + NMS_THREADS * (NMS_THREADS / 32));  // the round's sorted keys (64 KB)
    SortScratch& S = *reinterpret_cast<SortScratch*>(skeys + SORT_KS);

    const int b = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // frame-coordinate shift of this image's rows: columns (0, 1) of the reference rows [cx, cy, w, h, ...], columns
    // (0, 1, 2, 3) of the fixed rows [x1, y1, x2, y2, ...]
    const float shift_x = p.tile_xy ? (float)p.tile_xy[2 * b] : 0.0f, shift_y = p.tile_xy ? (float)p.tile_xy[2 * b + 1] : 0.0f;
    const int shift_cols = p.tile_xy ? (p.compat ? 4 : 2) : 0;
    float* out_b = out + (long)b * p.out_rows * 7;
    const unsigned long long* gkeys = keys + (long)b * stride;
    const int n_all = count[b];
    const int n_cap = n_all > MAX_NMS_BOXES ? MAX_NMS_BOXES : n_all;  // metrics.py:431-432 (ascending key order => top-30000 by score)
    if (tid == 0) s_kept = 0;
    __syncthreads();
    unsigned int n_pairs = 0;
    int consumed = 0;
    bool have_lo = false;
    unsigned long long lo = 0;
    const unsigned int v_or = bits_or[b], v_and = ~bits_nor[b];
    const int common_bits = __clz((int)(v_or ^ v_and));  // leading score bits all of this image's keys share
    const unsigned long long common = common_bits ? (unsigned long long)(v_or >> (32 - common_bits)) : 0ULL;

    // A detector's image is usually done inside its first few hundred candidates: the first round orders only ~4 * max_det keys
    // (a 1024-key bitonic sort instead of an 8192-key one), later rounds take full batches.
    const int want0 = min(SORT_KS, max(1024, 4 * p.max_det));
    while (consumed < n_cap && s_kept < p.max_det) {
    const int m_batch = cta_next_sorted_batch(gkeys, n_all, have_lo, lo, n_all - consumed, min(consumed == 0 ? want0 : SORT_KS, n_cap - consumed),
                                              32 + p.slot_bits, common_bits, common, skeys, S);
    const int n = min(m_batch, n_cap - consumed);  // the batch may run past the 30 000-candidate cap: the tail is not walked
    lo = skeys[n - 1];
    have_lo = true;
    consumed += n;
    for (int base = 0; base < n; base += NMS_THREADS) {
        const int kept0 = s_kept;
        if (kept0 >= p.max_det) break;
        const int i = base + tid;
        bool alive = i < n;
        Cand c;
        if (alive) {
            load_cand(p, b, skeys[i], c);
            // phase A: against everything kept in earlier chunks (parallel over candidates)
            // (a zero intersection gives an overlap of 0 or NaN, never > thr for thr >= 0: the division is skipped for disjoint
            // boxes, which is most of a dense scene's kept list -- 4.3 -> see profiles/r2c_nms_stress.md)
            for (int k = 0; k < kept0; ++k) {
                if (COUNT) ++n_pairs;
                const float bx1 = kx1[k], by1 = ky1[k], bx2 = kx2[k], by2 = ky2[k];
                if (fminf(bx2, c.x2) > fmaxf(bx1, c.x1) && fminf(by2, c.y2) > fmaxf(by1, c.y1) &&
                    iou_gt(bx1, by1, bx2, by2, kar[k], c.x1, c.y1, c.x2, c.y2, c.area, p.iou)) { alive = false; break; }
            }
        }
        // order-preserving compaction of survivors
        const unsigned int bal = __ballot_sync(0xffffffffu, alive);
        if (lane == 0) warp_cnt[warp] = __popc(bal);
        __syncthreads();
        int wbase = 0;
        for (int wj = 0; wj < warp; ++wj) wbase += warp_cnt[wj];
        if (alive) {
            const int pos = wbase + __popc(bal & ((1u << lane) - 1u));
            surv[pos] = tid;
            sx1[pos] = c.x1; sy1[pos] = c.y1; sx2[pos] = c.x2; sy2[pos] = c.y2; sar[pos] = c.area;
        }
        if (tid == NMS_THREADS - 1) s_nsurv = wbase + __popc(bal);
#pragma unroll
        for (int q = 0; q < 7; ++q) crow[tid][q] = c.row[q];
        __syncthreads();
        // phase B1: pairwise suppression masks among this chunk's survivors, in parallel.  Row s (survivor position s) has
        // bit j of word w set iff survivor 32 w + j < s overlaps s by more than the threshold (same arithmetic as everywhere:
        // a zero intersection gives 0 or NaN, never > thr, so it short-cuts the division).  The triangle of (word, row) tasks
        // -- word w pairs with rows 32 w .. ns - 1 -- is dealt out flat, one task = 32 column boxes (broadcast reads) against one
        // row box, independent loads: the row-per-thread form walked up to 511 dependent shared-memory chains of ~230 cycles
        // (117 000 cycles for a full chunk; this is ~3 000).
        {
            const int ns = s_nsurv;
            const int nw = (ns + 31) >> 5;
            const int n_tasks = nw * ns - 16 * nw * (nw - 1);   // sum over w of (ns - 32 w)
            int w = 0, start = 0;                                 // tasks of word w are [start, start + ns - 32 w)
            for (int q = tid; q < n_tasks; q += NMS_THREADS) {
                while (q >= start + ns - 32 * w) { start += ns - 32 * w; ++w; }
                const int s = 32 * w + (q - start);
                const float ax1 = sx1[s], ay1 = sy1[s], ax2 = sx2[s], ay2 = sy2[s], aar = sar[s];
                const int jn = min(32, s - 32 * w);               // columns of this word that precede row s
                if (COUNT) n_pairs += (unsigned int)jn;
                unsigned int m = 0;
#pragma unroll 8
                for (int j = 0; j < 32; ++j) {
                    const int u = 32 * w + j;
                    const float bx1 = sx1[u], by1 = sy1[u], bx2 = sx2[u], by2 = sy2[u];
                    const bool ov = (fminf(bx2, ax2) > fmaxf(bx1, ax1)) & (fminf(by2, ay2) > fmaxf(by1, ay1)) & (j < jn);
                    if (ov && iou_gt(bx1, by1, bx2, by2, sar[u], ax1, ay1, ax2, ay2, aar, p.iou)) m |= 1u << j;
                }
                smask[s * (NMS_THREADS / 32) + w] = m;
            }
        }
        __syncthreads();
        // phase B2: greedy resolution in order by one warp, 32 survivors (one mask row per lane) at a time.  Lane w holds word w
        // of the set kept within this chunk.  Suppression by survivors kept in EARLIER groups is tested by all lanes in
        // parallel; only the in-group order is sequential, and that runs on registers (one ballot per survivor) -- the
        // one-survivor-per-iteration form paid a shared-memory round trip per survivor (~30 000 cycles per full chunk).
        if (warp == 0) {
            const int ns = s_nsurv;
            int kept = kept0;
            unsigned int kw = 0;
            const int ngroups = (ns + 31) >> 5;
            for (int G = 0; G < ngroups && kept < p.max_det; ++G) {
                const int s = G * 32 + lane;
                const bool valid = s < ns;
                const unsigned int* mrow = smask + (valid ? s : 0) * (NMS_THREADS / 32);
                bool dead = !valid;
                for (int w = 0; w < G; ++w) {
                    const unsigned int kww = __shfl_sync(0xffffffffu, kw, w);
                    if (valid && (mrow[w] & kww) != 0u) dead = true;
                }
                const unsigned int diag = valid ? mrow[G] : 0u;  // earlier survivors of this group that overlap survivor s
                // in-group order: every lane resolves the whole group on registers from the 32 diagonal words (independent
                // shuffles) and the dead mask -- no vote per survivor on the dependent chain
                const unsigned int live = ~__ballot_sync(0xffffffffu, dead);
                unsigned int kg = 0;
                int room = p.max_det - kept;
#pragma unroll
                for (int t = 0; t < 32; ++t) {
                    const unsigned int dt = __shfl_sync(0xffffffffu, diag, t);
                    const bool keep_t = ((live >> t) & 1u) && (dt & kg) == 0u && room > 0;
                    if (keep_t) { kg |= 1u << t; --room; }
                }
                const int cnt = __popc(kg);
                if ((kg >> lane) & 1u) {  // the group's kept survivors append to the kept list and to the output, in order
                    const int pos = kept + __popc(kg & ((1u << lane) - 1u));
                    const int t = surv[s];
                    kx1[pos] = sx1[s]; ky1[pos] = sy1[s]; kx2[pos] = sx2[s]; ky2[pos] = sy2[s]; kar[pos] = sar[s];
#pragma unroll
                    for (int q = 0; q < 7; ++q) out_b[pos * 7 + q] = __fadd_rn(crow[t][q], q < shift_cols ? ((q & 1) ? shift_y : shift_x) : 0.0f);
                }
                if (lane == G) kw = kg;
                kept += cnt;
            }
            __syncwarp();
            if (lane == 0) s_kept = kept;
        }
        __syncthreads();
    }
    __syncthreads();  // s_kept and skeys[n - 1] have been read by everyone before the next round overwrites the batch
    }
    if (tid == 0) out_count[b] = s_kept;
    if (COUNT) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) n_pairs += __shfl_xor_sync(0xffffffffu, n_pairs, o);
        if (lane == 0 && pair_counter) atomicAdd(pair_counter, (unsigned long long)n_pairs);
    }
    if (p.out_rows > p.max_det) {  // gather-buffer form: zero the unused rows, append the count row
        const int kept = s_kept;
        for (int i = kept * 7 + tid; i < p.out_rows * 7; i += NMS_THREADS) out_b[i] = i == p.max_det * 7 ? (float)kept : 0.0f;
    }
}

struct BatchedWs {
    unsigned long long* keys;
    int* count;              // [B] candidates per image, then [B] OR of their score bits, [B] OR of the complements (one memset)
    unsigned int* bits_or;
    unsigned int* bits_nor;
};
static size_t batched_layout(int B, int N, int nc, int multi_label, void* base, BatchedWs* ws) {
    const size_t cap = (size_t)B * N * ((multi_label && nc > 1) ? nc : 1);
    size_t off = 0;
    uint8_t* b = (uint8_t*)base;
    auto take = [&](size_t bytes) { size_t o = off; off += align256(bytes); return b ? (void*)(b + o) : nullptr; };
    void* k0 = take(8 * cap);
    void* cnt = take(sizeof(int) * 3 * B);
    if (ws) { ws->keys = (unsigned long long*)k0; ws->count = (int*)cnt; ws->bits_or = (unsigned int*)cnt + B; ws->bits_nor = (unsigned int*)cnt + 2 * B; }
    return off + 256;
}

}  // namespace skb

using namespace skb;

extern "C" size_t skb_nms_workspace_bytes(int32_t n) { return n <= 0 ? 256 : nms_layout(n, nullptr, nullptr); }

extern "C" int skb_nms_f32(const float* boxes, const float* scores, int32_t n, float iou_thr, int64_t* keep, int32_t* n_keep, void* workspace,
                           size_t workspace_bytes, void* stream) {
    int rc = check_device();
    if (rc != SKB_OK) return rc;
    SKB_REQUIRE(n >= 0 && keep && n_keep, SKB_ERR_ARG, "nms: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    if (n == 0) {
        SKB_CUDA(cudaMemsetAsync(n_keep, 0, sizeof(int), st));
        return SKB_OK;
    }
    SKB_REQUIRE(boxes && scores && workspace && ((uintptr_t)boxes & 15) == 0, SKB_ERR_ARG, "nms: null or unaligned boxes");
    SKB_REQUIRE(workspace_bytes >= skb_nms_workspace_bytes(n), SKB_ERR_WORKSPACE, "nms: workspace %zu < %zu", workspace_bytes, skb_nms_workspace_bytes(n));
    NmsWs ws;
    nms_layout(n, workspace, &ws);
    const int cbk = (n + 63) / 64;
    constexpr int kSortBytes = SORT_KS * 8 + (int)sizeof(SortScratch);
    static PerDeviceOnce order_once;
    if (order_once.first()) SKB_CUDA(cudaFuncSetAttribute(nms_order_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSortBytes));
    nms_order_kernel<<<1, NMS_THREADS, kSortBytes, st>>>((const float4*)boxes, scores, n, ws.keys, ws.idx_out, ws.sboxes);
    SKB_LAUNCH_CHECK();
    nms_mask_kernel<<<dim3(cbk, cbk), 64, 0, st>>>(ws.sboxes, n, iou_thr, cbk, ws.mask);
    SKB_LAUNCH_CHECK();
    const size_t sh = sizeof(unsigned long long) * cbk;
    SKB_REQUIRE(sh <= 48 * 1024, SKB_ERR_UNSUPPORTED, "nms: n=%d too large for the single-CTA reduce", n);
    nms_reduce_kernel<<<1, 256, sh, st>>>(ws.mask, ws.idx_out, n, cbk, (long long*)keep, n_keep);
    SKB_LAUNCH_CHECK();
    return SKB_OK;
}

// Diagnostics: device pointer of an unsigned 64-bit counter that receives the IoU pair tests of every following batched NMS
// call (NULL switches the counting variant off again).
extern "C" int skb_debug_nms_pair_counter(unsigned long long* counter_dev) {
    g_nms_pair_counter = counter_dev;
    return SKB_OK;
}

extern "C" size_t skb_nms_batched_workspace_bytes(int32_t b, int32_t n, int32_t nc, int32_t multi_label) {
    if (b <= 0 || n <= 0) return 256;
    return batched_layout(b, n, nc, multi_label, nullptr, nullptr);
}

static int nms_batched_impl(const float* pred, int32_t b, int32_t n, int32_t nc, float conf_thr, float iou_thr,
                            const int32_t* classes_host, int32_t n_classes, int32_t agnostic, int32_t multi_label, int32_t max_det,
                            int32_t compat, float* out, int32_t* out_count, void* workspace, size_t workspace_bytes, void* stream,
                            const int32_t* tile_xy_dev, int32_t out_rows);

extern "C" int skb_nms_batched_f32(const float* pred, int32_t b, int32_t n, int32_t nc, float conf_thr, float iou_thr,
                                   const int32_t* classes_host, int32_t n_classes, int32_t agnostic, int32_t multi_label, int32_t max_det,
                                   int32_t compat, float* out, int32_t* out_count, void* workspace, size_t workspace_bytes, void* stream) {
    return nms_batched_impl(pred, b, n, nc, conf_thr, iou_thr, classes_host, n_classes, agnostic, multi_label, max_det, compat, out, out_count,
                            workspace, workspace_bytes, stream, nullptr, max_det);
}

extern "C" int skb_nms_batched_tiles_f32(const float* pred, int32_t b, int32_t n, int32_t nc, float conf_thr, float iou_thr, int32_t agnostic,
                                         int32_t multi_label, int32_t max_det, int32_t compat, const int32_t* tile_xy_dev, float* out_packed,
                                         int32_t* out_count, void* workspace, size_t workspace_bytes, void* stream) {
    SKB_REQUIRE(tile_xy_dev, SKB_ERR_ARG, "nms_batched_tiles: null tile origin table");
    return nms_batched_impl(pred, b, n, nc, conf_thr, iou_thr, nullptr, 0, agnostic, multi_label, max_det, compat, out_packed, out_count,
                            workspace, workspace_bytes, stream, tile_xy_dev, max_det + 1);
}

static int nms_batched_impl(const float* pred, int32_t b, int32_t n, int32_t nc, float conf_thr, float iou_thr,
                            const int32_t* classes_host, int32_t n_classes, int32_t agnostic, int32_t multi_label, int32_t max_det,
                            int32_t compat, float* out, int32_t* out_count, void* workspace, size_t workspace_bytes, void* stream,
                            const int32_t* tile_xy_dev, int32_t out_rows) {
    int rc = check_device();
    if (rc != SKB_OK) return rc;
    SKB_REQUIRE(pred && out && out_count && workspace && b > 0 && n > 0 && nc >= 0, SKB_ERR_ARG, "nms_batched: bad arguments");
    SKB_REQUIRE(compat == 0 || compat == 1, SKB_ERR_ARG, "nms_batched: compat must be 0 (reference) or 1 (fixed)");
    SKB_REQUIRE(iou_thr >= 0.0f && iou_thr <= 1.0f && conf_thr >= 0.0f && conf_thr <= 1.0f, SKB_ERR_ARG,
                "nms_batched: thresholds must lie in [0, 1] (metrics.py:386-387), got conf %g iou %g", (double)conf_thr, (double)iou_thr);
    SKB_REQUIRE(max_det >= 1 && max_det <= NMS_MAX_KEEP, SKB_ERR_UNSUPPORTED, "nms_batched: max_det=%d (supported: 1..%d)", max_det, NMS_MAX_KEEP);
    int slot_bits = 1;
    while ((1L << slot_bits) < (long)n * (nc > 1 ? nc : 1)) ++slot_bits;
    SKB_REQUIRE(slot_bits <= KEY_MAX_SLOT_BITS, SKB_ERR_UNSUPPORTED, "nms_batched: N*nc=%ld exceeds the %d-bit slot field of the sort key",
                (long)n * (nc > 1 ? nc : 1), KEY_MAX_SLOT_BITS);
    SKB_REQUIRE(n_classes <= 32, SKB_ERR_UNSUPPORTED, "nms_batched: at most 32 class filters");
    multi_label = (multi_label && nc > 1) ? 1 : 0;  // metrics.py:396
    SKB_REQUIRE(workspace_bytes >= skb_nms_batched_workspace_bytes(b, n, nc, multi_label), SKB_ERR_WORKSPACE, "nms_batched: workspace too small");
    BatchedWs ws;
    batched_layout(b, n, nc, multi_label, workspace, &ws);
    const size_t cap = (size_t)b * n * (multi_label ? nc : 1);
    cudaStream_t st = (cudaStream_t)stream;
    SKB_CUDA(cudaMemsetAsync(ws.count, 0, sizeof(int) * 3 * b, st));
    FilterParams fp;
    memset(&fp, 0, sizeof(fp));
    fp.pred = pred; fp.B = b; fp.N = n; fp.nc = nc; fp.no = nc + 5; fp.conf = conf_thr; fp.multi_label = multi_label; fp.compat = compat;
    fp.slot_bits = slot_bits;
    fp.n_classes = classes_host ? n_classes : 0;
    for (int i = 0; i < fp.n_classes; ++i) fp.classes[i] = (float)classes_host[i];
    const long nbox = (long)b * n;
    long g = (nbox + 255) / 256;
    const long gcap = (long)num_sms() * 16;
    const long stride = (long)n * (multi_label ? nc : 1);  // key segment per image
    nms_filter_kernel<<<(int)(g > gcap ? gcap : g), 256, 0, st>>>(fp, ws.keys, ws.count, stride, ws.bits_or, ws.bits_nor);
    SKB_LAUNCH_CHECK();
    BatchedParams bp;
    bp.pred = pred; bp.B = b; bp.N = n; bp.nc = nc; bp.no = nc + 5; bp.iou = iou_thr; bp.agnostic = agnostic; bp.multi_label = multi_label;
    bp.compat = compat; bp.max_det = max_det; bp.capacity = (int)cap; bp.slot_bits = slot_bits;
    bp.tile_xy = tile_xy_dev; bp.out_rows = out_rows;
    constexpr int kMaskBytes = NMS_THREADS * (NMS_THREADS / 32) * (int)sizeof(unsigned int) + SORT_KS * 8 + (int)sizeof(SortScratch);
    static PerDeviceOnce attr_once;
    if (attr_once.first()) {
        SKB_CUDA(cudaFuncSetAttribute(nms_keptlist_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaskBytes));
        SKB_CUDA(cudaFuncSetAttribute(nms_keptlist_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaskBytes));
    }
    if (g_nms_pair_counter)
        nms_keptlist_kernel<true><<<b, NMS_THREADS, kMaskBytes, st>>>(bp, ws.keys, stride, ws.count, ws.bits_or, ws.bits_nor, out, out_count, g_nms_pair_counter);
    else
        nms_keptlist_kernel<false><<<b, NMS_THREADS, kMaskBytes, st>>>(bp, ws.keys, stride, ws.count, ws.bits_or, ws.bits_nor, out, out_count, nullptr);
    SKB_LAUNCH_CHECK();
    return SKB_OK;
}


// =============================================================================================
// (3) cross-tile merge (BASELINE config 4; NOT IN REFERENCE): the all_gather'ed per-tile rows
//     [world][tiles_per_rank][max_det + 1][7] (rank r holds global tiles r, r + world, ...) are re-expressed as a prediction
//     tensor [frames][tiles_per_frame * max_det][5 + nc] so that the SAME wrapper kernels (metrics.py:361-457 semantics)
//     perform the per-frame merge NMS.  Reference rows [cx,cy,w,h,obj,cls_prob,cls_id] are copied with the class probability
//     at column 5 + cls_id (best-class selection recovers (cls_prob, cls_id) exactly); fixed rows [x1,y1,x2,y2,conf,cls] become
//     centre form with obj = conf and class probability 1 (conf * 1 is exact).  Rows beyond a tile's count stay zero
//     (objectness 0 -> dropped by the confidence filter).  One thread per output float: coalesced stores.
// =============================================================================================
namespace skb {
__global__ void tile_merge_pred_kernel(const float* __restrict__ gathered, int world, int tiles_per_rank, int n_tiles, int tiles_per_frame,
                                       int max_det, int nc, int compat, float* __restrict__ pred) {
    const int no = 5 + nc;
    const long total = (long)n_tiles * max_det * no;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int c = (int)(i % no);
        const long rj = i / no;
        const int j = (int)(rj % max_det);
        const int g = (int)(rj / max_det);  // global tile id = frame * tiles_per_frame + k: pred rows are already in (frame, k, j) order
        const float* src = gathered + ((long)(g % world) * tiles_per_rank + g / world) * (max_det + 1) * 7;
        float v = 0.0f;
        if (j < (int)src[max_det * 7]) {
            const float* r = src + j * 7;
            if (compat == 0) {
                if (c < 5) v = r[c];
                else if (nc == 1) v = 1.0f;
                else if (c - 5 == (int)r[6]) v = r[5];
            } else {
                if (c == 0) v = __fmul_rn(__fadd_rn(r[0], r[2]), 0.5f);
                else if (c == 1) v = __fmul_rn(__fadd_rn(r[1], r[3]), 0.5f);
                else if (c == 2) v = __fsub_rn(r[2], r[0]);
                else if (c == 3) v = __fsub_rn(r[3], r[1]);
                else if (c == 4) v = r[4];
                else if (c - 5 == (int)r[5]) v = 1.0f;
            }
        }
        pred[i] = v;
    }
}
}  // namespace skb

extern "C" int skb_tile_merge_pred_f32(const float* gathered, int32_t world, int32_t tiles_per_rank, int32_t n_frames, int32_t tiles_per_frame,
                                       int32_t max_det, int32_t nc, int32_t compat, float* pred, void* stream) {
    int rc = check_device();
    if (rc != SKB_OK) return rc;
    const long n_tiles = (long)n_frames * tiles_per_frame;
    SKB_REQUIRE(gathered && pred && world >= 1 && tiles_per_rank >= 1 && n_frames >= 1 && tiles_per_frame >= 1 && max_det >= 1 && nc >= 1 &&
                    (compat == 0 || compat == 1) && (long)world * tiles_per_rank >= n_tiles,
                SKB_ERR_ARG, "tile_merge_pred: bad arguments (world %d x %d tiles per rank < %ld tiles)", world, tiles_per_rank, n_tiles);
    const long total = n_tiles * max_det * (5 + nc);
    long g = (total + 255) / 256;
    const long gcap = (long)num_sms() * 16;
    tile_merge_pred_kernel<<<(int)(g > gcap ? gcap : g), 256, 0, (cudaStream_t)stream>>>(gathered, world, tiles_per_rank, (int)n_tiles,
                                                                                          tiles_per_frame, max_det, nc, compat, pred);
    SKB_LAUNCH_CHECK();
    return SKB_OK;
}
