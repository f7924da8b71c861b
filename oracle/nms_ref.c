/* CPU oracle (TEST INFRASTRUCTURE): scalar restatement of greedy IoU NMS.
 *
 * The reference calls torchvision.ops.nms (skyeye/utils/metrics.py:442). torchvision is a
 * third-party dependency that is not vendored under /root/reference (requirements.txt:2 pins only
 * ">=0.8.1"; this image has torchvision 0.26.0).  Its published CPU algorithm, restated here:
 *   order = argsort(scores, descending, stable)
 *   area  = (x2 - x1) * (y2 - y1)                                   (fp32)
 *   for i in order: if suppressed[i] skip; keep i;
 *       for later j in order: w = max(0, min(x2) - max(x1)); h = max(0, min(y2) - max(y1));
 *           inter = w * h; iou = inter / (area_i + area_j - inter)  (fp32, true division, no FMA)
 *           suppressed[j] |= iou > thr                              (strict; NaN never suppresses)
 *   returns kept ORIGINAL indices in descending-score order (int64).
 * Pinned against torchvision.ops.nms itself in tests/test_oracle_nms.py (ties, duplicates,
 * zero-area boxes) and against the committed fixtures in tests/golden/nms_*.npz.
 *
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared nms_ref.c -o _build/libnms_ref.so
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { float s; int64_t i; } sk_t;

static void merge_sort_desc(sk_t* a, sk_t* tmp, int64_t n) {
    /* bottom-up stable merge sort, descending by score, ties keep ascending original index */
    for (int64_t w = 1; w < n; w *= 2) {
        for (int64_t lo = 0; lo < n; lo += 2 * w) {
            int64_t mid = lo + w < n ? lo + w : n, hi = lo + 2 * w < n ? lo + 2 * w : n;
            int64_t l = lo, r = mid, o = lo;
            while (l < mid && r < hi) tmp[o++] = (a[r].s > a[l].s) ? a[r++] : a[l++];
            while (l < mid) tmp[o++] = a[l++];
            while (r < hi) tmp[o++] = a[r++];
        }
        memcpy(a, tmp, (size_t)n * sizeof(sk_t));
    }
}

/* boxes: [n,4] (x1,y1,x2,y2) fp32; keep: capacity n. Returns number kept. */
int64_t skyeye_oracle_nms(const float* boxes, const float* scores, int64_t n, float thr, int64_t* keep) {
    if (n <= 0) return 0;
    sk_t* ord = (sk_t*)malloc((size_t)n * sizeof(sk_t));
    sk_t* tmp = (sk_t*)malloc((size_t)n * sizeof(sk_t));
    float* area = (float*)malloc((size_t)n * sizeof(float));
    unsigned char* sup = (unsigned char*)calloc((size_t)n, 1);
    for (int64_t i = 0; i < n; ++i) {
        ord[i].s = scores[i]; ord[i].i = i;
        area[i] = (boxes[4 * i + 2] - boxes[4 * i + 0]) * (boxes[4 * i + 3] - boxes[4 * i + 1]);
    }
    merge_sort_desc(ord, tmp, n);
    int64_t nk = 0;
    for (int64_t a = 0; a < n; ++a) {
        int64_t i = ord[a].i;
        if (sup[i]) continue;
        keep[nk++] = i;
        float ix1 = boxes[4 * i], iy1 = boxes[4 * i + 1], ix2 = boxes[4 * i + 2], iy2 = boxes[4 * i + 3];
        float iarea = area[i];
        for (int64_t b = a + 1; b < n; ++b) {
            int64_t j = ord[b].i;
            if (sup[j]) continue;
            float xx1 = ix1 > boxes[4 * j] ? ix1 : boxes[4 * j];
            float yy1 = iy1 > boxes[4 * j + 1] ? iy1 : boxes[4 * j + 1];
            float xx2 = ix2 < boxes[4 * j + 2] ? ix2 : boxes[4 * j + 2];
            float yy2 = iy2 < boxes[4 * j + 3] ? iy2 : boxes[4 * j + 3];
            float w = xx2 - xx1; if (!(w > 0.0f)) w = 0.0f;
            float h = yy2 - yy1; if (!(h > 0.0f)) h = 0.0f;
            float inter = w * h;
            float ovr = inter / (iarea + area[j] - inter);
            if (ovr > thr) sup[j] = 1;
        }
    }
    free(ord); free(tmp); free(area); free(sup);
    return nk;
}
