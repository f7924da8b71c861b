// Evaluation bookkeeping right behind the path: the correct-prediction matrix of validate()'s
// process_batch (skyeye/cli/validate.py:71-108, which calls box_iou, skyeye/utils/metrics.py:17-44) on the device, so the
// per-image matching of the validation loop needs no device->host round trip per IoU threshold.
//
// The reference, per threshold t: all (label, detection) pairs of equal class with IoU >= t, sorted by IoU descending; keep each
// detection's first pair (its best label), re-sort, keep each label's first pair (its best detection); those detections are
// correct at t.  A detection's best label and a label's best detection do not depend on t (raising t only removes pairs from the
// tail of the order), so:  correct[d, t] = (d is the highest-IoU detection among those whose best same-class label is l)
// and IoU(d, l) >= iouv[t].  One CTA per call: n detections x m labels is a few thousand pairs.
#include "common.cuh"

namespace skb {

__device__ __forceinline__ float pair_iou(const float* a, const float* b) {  // a, b = x1, y1, x2, y2; metrics.py:17-44, eps 1e-7
    const float w = fmaxf(__fsub_rn(fminf(a[2], b[2]), fmaxf(a[0], b[0])), 0.0f);
    const float h = fmaxf(__fsub_rn(fminf(a[3], b[3]), fmaxf(a[1], b[1])), 0.0f);
    const float inter = __fmul_rn(w, h);
    const float area_a = __fmul_rn(__fsub_rn(a[2], a[0]), __fsub_rn(a[3], a[1]));
    const float area_b = __fmul_rn(__fsub_rn(b[2], b[0]), __fsub_rn(b[3], b[1]));
    return __fdiv_rn(inter, __fadd_rn(__fsub_rn(__fadd_rn(area_a, area_b), inter), 1e-7f));
}

// labels [m][5] = (cls, x1, y1, x2, y2); dets [n][6] = (x1, y1, x2, y2, conf, cls); correct uint8 [n][T]
__global__ void __launch_bounds__(256)
match_detections_kernel(const float* __restrict__ labels, int m, const float* __restrict__ dets, int n, const float* __restrict__ iouv, int T,
                        uint8_t* __restrict__ correct, int* __restrict__ best_label, float* __restrict__ best_iou,
                        unsigned long long* __restrict__ label_best) {
    for (int l = threadIdx.x; l < m; l += blockDim.x) label_best[l] = 0ULL;
    for (int d = threadIdx.x; d < n; d += blockDim.x) {
        const float* dp = dets + (long)d * 6;
        const float box[4] = {dp[0], dp[1], dp[2], dp[3]};
        int bl = -1;
        float bi = -1.0f;
        for (int l = 0; l < m; ++l) {
            const float* lp = labels + (long)l * 5;
            if (lp[0] != dp[5]) continue;
            const float iou = pair_iou(lp + 1, box);
            if (iou > bi) { bi = iou; bl = l; }  // first maximum: the lowest label index on ties
        }
        best_label[d] = bl;
        best_iou[d] = bi;
    }
    __syncthreads();
    for (int d = threadIdx.x; d < n; d += blockDim.x) {
        const int bl = best_label[d];
        if (bl >= 0)  // highest IoU wins the label; ties go to the lowest detection index (IoU >= 0: its bit pattern orders like the value)
            atomicMax(label_best + bl, ((unsigned long long)__float_as_uint(best_iou[d]) << 32) | (unsigned int)(0x7fffffff - d));
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n * T; i += blockDim.x) {
        const int d = i / T, t = i - d * T;
        const int bl = best_label[d];
        bool ok = false;
        if (bl >= 0) {
            const int winner = 0x7fffffff - (int)(unsigned int)(label_best[bl] & 0xffffffffULL);
            ok = winner == d && best_iou[d] >= iouv[t];
        }
        correct[i] = ok ? 1 : 0;
    }
}

}  // namespace skb

using namespace skb;

extern "C" size_t skb_match_workspace_bytes(int32_t n_dets, int32_t n_labels) {
    return (size_t)(n_dets > 0 ? n_dets : 0) * 8 + (size_t)(n_labels > 0 ? n_labels : 0) * 8 + 512;
}

extern "C" int skb_match_detections_f32(const float* labels, int32_t n_labels, const float* dets, int32_t n_dets, const float* iouv, int32_t n_iou,
                                        uint8_t* correct, void* workspace, size_t workspace_bytes, void* stream) {
    int rc = check_device();
    if (rc != SKB_OK) return rc;
    SKB_REQUIRE(n_labels >= 0 && n_dets >= 0 && n_iou >= 1 && iouv && (n_dets == 0 || (dets && correct)) && (n_labels == 0 || labels), SKB_ERR_ARG,
                "match_detections: bad arguments");
    if (n_dets == 0) return SKB_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (n_labels == 0) {
        SKB_CUDA(cudaMemsetAsync(correct, 0, (size_t)n_dets * n_iou, st));
        return SKB_OK;
    }
    SKB_REQUIRE(workspace && workspace_bytes >= skb_match_workspace_bytes(n_dets, n_labels), SKB_ERR_WORKSPACE, "match_detections: workspace too small");
    uint8_t* w = (uint8_t*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    unsigned long long* label_best = (unsigned long long*)w;
    int* best_label = (int*)(w + (size_t)n_labels * 8);
    float* best_iou = (float*)(w + (size_t)n_labels * 8 + (size_t)n_dets * 4);
    match_detections_kernel<<<1, 256, 0, st>>>(labels, n_labels, dets, n_dets, iouv, n_iou, correct, best_label, best_iou, label_best);
    SKB_LAUNCH_CHECK();
    return SKB_OK;
}
