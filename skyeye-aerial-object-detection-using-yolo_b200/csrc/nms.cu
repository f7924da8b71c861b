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
#include <cub/cub.cuh>

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

// =============================================================================================
// (1) generic bit-exact NMS
// =============================================================================================
__global__ void iota_kernel(int* idx, int n) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) idx[i] = i;
}
__global__ void gather_boxes_kernel(const float4* __restrict__ boxes, const int* __restrict__ order, int n, float4* __restrict__ out) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) out[i] = boxes[order[i]];
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
    float* keys_out;
    int* idx_in;
    int* idx_out;
    float4* sboxes;
    unsigned long long* mask;
    void* cub_tmp;
    size_t cub_bytes;
};
static size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }
static size_t nms_layout(int n, void* base, NmsWs* ws) {
    size_t cub_bytes = 0;
    cub::DeviceRadixSort::SortPairsDescending(nullptr, cub_bytes, (const float*)nullptr, (float*)nullptr, (const int*)nullptr, (int*)nullptr, n);
    const int cbk = (n + 63) / 64;
    size_t off = 0;
    uint8_t* b = (uint8_t*)base;
    auto take = [&](size_t bytes) { size_t o = off; off += align256(bytes); return b ? (void*)(b + o) : nullptr; };
    void* k = take(sizeof(float) * n);
    void* i0 = take(sizeof(int) * n);
    void* i1 = take(sizeof(int) * n);
    void* sb = take(sizeof(float4) * n);
    void* m = take(sizeof(unsigned long long) * (size_t)n * cbk);
    void* c = take(cub_bytes);
    if (ws) { ws->keys_out = (float*)k; ws->idx_in = (int*)i0; ws->idx_out = (int*)i1; ws->sboxes = (float4*)sb; ws->mask = (unsigned long long*)m; ws->cub_tmp = c; ws->cub_bytes = cub_bytes; }
    return off + 256;
}

// =============================================================================================
// (2) batched wrapper
// =============================================================================================
// sort key = [image | ~score (32 bits) | slot]: the slot field is sized per call for N * nc (nc = 80 at 1280^2 needs 23 bits),
// the image field takes what is left of the upper 32 bits
constexpr int KEY_MAX_SLOT_BITS = 27;
constexpr int MAX_NMS_BOXES = 30000;  // metrics.py:393
constexpr int NMS_THREADS = 512;
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

__device__ __forceinline__ unsigned int desc_bits(float s) {
    unsigned int u = __float_as_uint(s);
    u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);  // ascending-orderable
    return ~u;                                       // descending
}
__device__ __forceinline__ bool class_pass(const FilterParams& p, float col5) {
    if (p.n_classes <= 0) return true;
    for (int i = 0; i < p.n_classes; ++i)
        if (col5 == p.classes[i]) return true;
    return false;
}

// One thread per (image, box).  Every thread writes its own key slot(s): the sort key of a surviving
// row or the all-ones sentinel (sorts last), so no position atomics are needed; the per-image counts
// are warp-aggregated (one atomic per warp when the warp sits inside one image).
__global__ void nms_filter_kernel(const FilterParams p, unsigned long long* __restrict__ keys, int* __restrict__ count) {
    const long nbox = (long)p.B * p.N;
    const long nround = (nbox + blockDim.x - 1) / blockDim.x * blockDim.x;  // whole warps stay in the loop together
    const int per = p.multi_label ? p.nc : 1;
    const int lane = threadIdx.x & 31;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < nround; i += (long)gridDim.x * blockDim.x) {
        const bool inb = i < nbox;
        const int b = inb ? (int)(i / p.N) : -1;
        int emitted = 0;
        if (inb) {
            const int r = (int)(i % p.N);
            const float* x = p.pred + i * p.no;
            const float obj = x[4];
            unsigned long long* kout = keys + i * per;
            const bool pass = obj > p.conf;  // metrics.py:391,402
            auto key_of = [&](float score, int slot) {
                return ((unsigned long long)b << (32 + p.slot_bits)) | ((unsigned long long)desc_bits(score) << p.slot_bits) |
                       (unsigned long long)slot;
            };
            const unsigned long long none = ~0ULL;
            if (p.nc > 1 || (p.compat == 1 && p.nc == 1)) {
                if (p.multi_label) {  // metrics.py:407-410: one row per (box, class) above the threshold
                    for (int j = 0; j < p.nc; ++j) {
                        unsigned long long k = none;
                        if (pass) {
                            const float cp = x[5 + j];
                            const float conf = p.compat ? __fmul_rn(cp, obj) : cp;
                            if (conf > p.conf && class_pass(p, p.compat ? (float)j : cp)) { k = key_of(p.compat ? conf : obj, r * p.nc + j); ++emitted; }
                        }
                        kout[j] = k;
                    }
                } else {  // metrics.py:411-414: best class (first maximum)
                    unsigned long long k = none;
                    if (pass) {
                        float best = x[5];
                        int bj = 0;
                        if (p.compat) best = __fmul_rn(best, obj);
                        for (int j = 1; j < p.nc; ++j) {
                            float cp = x[5 + j];
                            if (p.compat) cp = __fmul_rn(cp, obj);
                            if (cp > best) { best = cp; bj = j; }
                        }
                        if (best > p.conf && class_pass(p, p.compat ? (float)bj : best)) { k = key_of(p.compat ? best : obj, r * p.nc + bj); ++emitted; }
                    }
                    kout[0] = k;
                }
            } else {  // nc == 1 (metrics.py:415-419): row [cx,cy,w,h,obj,0]; nc == 0 in fixed mode
                unsigned long long k = none;
                if (pass && class_pass(p, 0.0f)) { k = key_of(obj, r); ++emitted; }
                kout[0] = k;
            }
        }
        const int b0 = __shfl_sync(0xffffffffu, b, 0);
        if (__all_sync(0xffffffffu, b == b0)) {
            int tot = emitted;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
            if (lane == 0 && tot > 0 && b0 >= 0) atomicAdd(count + b0, tot);
        } else if (emitted > 0) {
            atomicAdd(count + b, emitted);
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

// grid = B images, 512 threads. Sorted keys of image b are the segment [off_b, off_b + cnt_b).
template <bool COUNT>
__global__ void __launch_bounds__(NMS_THREADS)
nms_keptlist_kernel(const BatchedParams p, const unsigned long long* __restrict__ keys, const int* __restrict__ count,
                    float* __restrict__ out, int* __restrict__ out_count, unsigned long long* __restrict__ pair_counter) {
    __shared__ float kx1[NMS_MAX_KEEP], ky1[NMS_MAX_KEEP], kx2[NMS_MAX_KEEP], ky2[NMS_MAX_KEEP], kar[NMS_MAX_KEEP];
    __shared__ int surv[NMS_THREADS];   // chunk-local ids of phase-A survivors, in order
    __shared__ int warp_cnt[NMS_THREADS / 32];
    __shared__ int s_nsurv, s_kept;
    __shared__ float cx1[NMS_THREADS], cy1[NMS_THREADS], cx2[NMS_THREADS], cy2[NMS_THREADS], car[NMS_THREADS];
    __shared__ float crow[NMS_THREADS][7];
    extern __shared__ unsigned int smask[];  // [NMS_THREADS][NMS_THREADS / 32] pairwise suppression bits of a chunk (32 KB)

    const int b = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // frame-coordinate shift of this image's rows: columns (0, 1) of the reference rows [cx, cy, w, h, ...], columns
    // (0, 1, 2, 3) of the fixed rows [x1, y1, x2, y2, ...]
    float shift = 0.0f;
    if (p.tile_xy && lane < (p.compat ? 4 : 2)) shift = (float)p.tile_xy[2 * b + (lane & 1)];
    float* out_b = out + (long)b * p.out_rows * 7;
    long off = 0;
    for (int i = 0; i < b; ++i) off += count[i];
    int n = count[b];
    if (n > MAX_NMS_BOXES) n = MAX_NMS_BOXES;  // metrics.py:431-432 (sorted order => top-30000 by score)
    if (tid == 0) s_kept = 0;
    __syncthreads();
    unsigned int n_pairs = 0;

    for (int base = 0; base < n; base += NMS_THREADS) {
        const int kept0 = s_kept;
        if (kept0 >= p.max_det) break;
        const int i = base + tid;
        bool alive = i < n;
        Cand c;
        if (alive) {
            load_cand(p, b, keys[off + i], c);
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
        }
        if (tid == NMS_THREADS - 1) s_nsurv = wbase + __popc(bal);
        cx1[tid] = c.x1; cy1[tid] = c.y1; cx2[tid] = c.x2; cy2[tid] = c.y2; car[tid] = c.area;
#pragma unroll
        for (int q = 0; q < 7; ++q) crow[tid][q] = c.row[q];
        __syncthreads();
        // phase B1: pairwise suppression masks among this chunk's survivors, in parallel.  Row s (survivor position s) has
        // bit j of word j / 32 set iff survivor j < s overlaps s by more than the threshold (same arithmetic as everywhere:
        // a zero intersection gives 0 or NaN, never > thr, so it short-cuts the division).
        {
            const int ns = s_nsurv;
            if (tid < ns) {
                const int t = surv[tid];
                const float ax1 = cx1[t], ay1 = cy1[t], ax2 = cx2[t], ay2 = cy2[t], aar = car[t];
                for (int w = 0; w <= (tid >> 5); ++w) {
                    const int j0 = w * 32, j1 = min(j0 + 32, tid);
                    unsigned int m = 0;
                    if (COUNT) n_pairs += (unsigned int)(j1 - j0);
                    for (int j = j0; j < j1; ++j) {
                        const int u = surv[j];
                        const float xx1 = fmaxf(cx1[u], ax1), yy1 = fmaxf(cy1[u], ay1);
                        const float xx2 = fminf(cx2[u], ax2), yy2 = fminf(cy2[u], ay2);
                        if (xx2 > xx1 && yy2 > yy1 &&
                            iou_gt(cx1[u], cy1[u], cx2[u], cy2[u], car[u], ax1, ay1, ax2, ay2, aar, p.iou))
                            m |= 1u << (j - j0);
                    }
                    smask[tid * (NMS_THREADS / 32) + w] = m;
                }
            }
        }
        __syncthreads();
        // phase B2: greedy resolution in order by one warp: lane w holds word w of the set kept within this chunk
        if (warp == 0) {
            const int ns = s_nsurv;
            int kept = kept0;
            unsigned int kw = 0;
            for (int s = 0; s < ns && kept < p.max_det; ++s) {
                const unsigned int m = lane <= (s >> 5) ? smask[s * (NMS_THREADS / 32) + lane] : 0u;
                if (!__any_sync(0xffffffffu, (m & kw) != 0u)) {
                    const int t = surv[s];
                    if (lane == 0) { kx1[kept] = cx1[t]; ky1[kept] = cy1[t]; kx2[kept] = cx2[t]; ky2[kept] = cy2[t]; kar[kept] = car[t]; }
                    if (lane < 7) out_b[kept * 7 + lane] = __fadd_rn(crow[t][lane], shift);
                    if (lane == (s >> 5)) kw |= 1u << (s & 31);
                    ++kept;
                }
            }
            __syncwarp();
            if (lane == 0) s_kept = kept;
        }
        __syncthreads();
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
    unsigned long long* keys_in;
    unsigned long long* keys_out;
    unsigned int* total;
    int* count;
    void* cub_tmp;
    size_t cub_bytes;
};
static size_t batched_layout(int B, int N, int nc, int multi_label, void* base, BatchedWs* ws) {
    const size_t cap = (size_t)B * N * ((multi_label && nc > 1) ? nc : 1);
    size_t cub_bytes = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, cub_bytes, (const unsigned long long*)nullptr, (unsigned long long*)nullptr, (long)cap);
    size_t off = 0;
    uint8_t* b = (uint8_t*)base;
    auto take = [&](size_t bytes) { size_t o = off; off += align256(bytes); return b ? (void*)(b + o) : nullptr; };
    void* k0 = take(8 * cap);
    void* k1 = take(8 * cap);
    void* cnt = take(sizeof(int) * (B + 1));
    void* c = take(cub_bytes);
    if (ws) { ws->keys_in = (unsigned long long*)k0; ws->keys_out = (unsigned long long*)k1; ws->count = (int*)cnt + 1; ws->total = (unsigned int*)cnt; ws->cub_tmp = c; ws->cub_bytes = cub_bytes; }
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
    iota_kernel<<<(n + 255) / 256, 256, 0, st>>>(ws.idx_in, n);
    SKB_LAUNCH_CHECK();
    SKB_CUDA(cub::DeviceRadixSort::SortPairsDescending(ws.cub_tmp, ws.cub_bytes, scores, ws.keys_out, ws.idx_in, ws.idx_out, n, 0, 32, st));
    gather_boxes_kernel<<<(n + 255) / 256, 256, 0, st>>>((const float4*)boxes, ws.idx_out, n, ws.sboxes);
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
    int slot_bits = 1, img_bits = 1;
    while ((1L << slot_bits) < (long)n * (nc > 1 ? nc : 1)) ++slot_bits;
    while ((1 << img_bits) < b) ++img_bits;
    SKB_REQUIRE(slot_bits <= KEY_MAX_SLOT_BITS && slot_bits + img_bits <= 32, SKB_ERR_UNSUPPORTED,
                "nms_batched: B=%d N*nc=%ld exceed the 64-bit sort key (image %d + slot %d bits > 32)", b, (long)n * (nc > 1 ? nc : 1),
                img_bits, slot_bits);
    SKB_REQUIRE(n_classes <= 32, SKB_ERR_UNSUPPORTED, "nms_batched: at most 32 class filters");
    multi_label = (multi_label && nc > 1) ? 1 : 0;  // metrics.py:396
    SKB_REQUIRE(workspace_bytes >= skb_nms_batched_workspace_bytes(b, n, nc, multi_label), SKB_ERR_WORKSPACE, "nms_batched: workspace too small");
    BatchedWs ws;
    batched_layout(b, n, nc, multi_label, workspace, &ws);
    const size_t cap = (size_t)b * n * (multi_label ? nc : 1);
    cudaStream_t st = (cudaStream_t)stream;
    SKB_CUDA(cudaMemsetAsync(ws.total, 0, sizeof(int) * (b + 1), st));
    FilterParams fp;
    memset(&fp, 0, sizeof(fp));
    fp.pred = pred; fp.B = b; fp.N = n; fp.nc = nc; fp.no = nc + 5; fp.conf = conf_thr; fp.multi_label = multi_label; fp.compat = compat;
    fp.slot_bits = slot_bits;
    fp.n_classes = classes_host ? n_classes : 0;
    for (int i = 0; i < fp.n_classes; ++i) fp.classes[i] = (float)classes_host[i];
    const long nbox = (long)b * n;
    long g = (nbox + 255) / 256;
    const long gcap = (long)num_sms() * 16;
    nms_filter_kernel<<<(int)(g > gcap ? gcap : g), 256, 0, st>>>(fp, ws.keys_in, ws.count);
    SKB_LAUNCH_CHECK();
    SKB_CUDA(cub::DeviceRadixSort::SortKeys(ws.cub_tmp, ws.cub_bytes, ws.keys_in, ws.keys_out, (long)cap, 0, 32 + slot_bits + img_bits, st));
    BatchedParams bp;
    bp.pred = pred; bp.B = b; bp.N = n; bp.nc = nc; bp.no = nc + 5; bp.iou = iou_thr; bp.agnostic = agnostic; bp.multi_label = multi_label;
    bp.compat = compat; bp.max_det = max_det; bp.capacity = (int)cap; bp.slot_bits = slot_bits;
    bp.tile_xy = tile_xy_dev; bp.out_rows = out_rows;
    constexpr int kMaskBytes = NMS_THREADS * (NMS_THREADS / 32) * (int)sizeof(unsigned int);
    static PerDeviceOnce attr_once;
    if (attr_once.first()) {
        SKB_CUDA(cudaFuncSetAttribute(nms_keptlist_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaskBytes));
        SKB_CUDA(cudaFuncSetAttribute(nms_keptlist_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaskBytes));
    }
    if (g_nms_pair_counter) nms_keptlist_kernel<true><<<b, NMS_THREADS, kMaskBytes, st>>>(bp, ws.keys_out, ws.count, out, out_count, g_nms_pair_counter);
    else nms_keptlist_kernel<false><<<b, NMS_THREADS, kMaskBytes, st>>>(bp, ws.keys_out, ws.count, out, out_count, nullptr);
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
