/* rpn_oracle.c -- plain-C restatement of the tf-rpn box hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Second, independent restatement next to oracle/rpn_oracle.py (tests cross-check the two and
 * both against tests/golden).  It is also the timed CPU baseline of bench.py ("port", OpenMP over
 * images).  PARITY UNPINNED against real TensorFlow -- see oracle/__init__.py.
 * Build: make -C oracle   (gcc -O2 -ffp-contract=off: no FMA contraction, IEEE float ops).
 *
 * Reference lines restated (relative to /root/reference):
 *   iou            utils/bbox_utils.py:126-150      encode   utils/bbox_utils.py:98-124
 *   decode         utils/bbox_utils.py:72-96        targets  utils/train_utils.py:84-144
 *   sampler        utils/train_utils.py:50-65 (counter RNG instead of tf.random)
 *   top_k/gather   predictor.py:58-60               nms      utils/bbox_utils.py:48-70 [TF-internal]
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define API __attribute__((visibility("default")))

static inline float area4(const float* b) { return (b[2] - b[0]) * (b[3] - b[1]); }

/* utils/bbox_utils.py:141-150 */
static inline float iou_ref(const float* b, float ba, const float* g, float ga) {
    float x_top = fmaxf(b[1], g[1]), y_top = fmaxf(b[0], g[0]);
    float x_bot = fminf(b[3], g[3]), y_bot = fminf(b[2], g[2]);
    float inter = fmaxf(x_bot - x_top, 0.0f) * fmaxf(y_bot - y_top, 0.0f);
    float uni = (ba + ga) - inter;
    return inter / uni;
}

API void oracle_iou_map(const float* boxes, int boxes_batched, const float* gt, int B, int N, int G, float* out) {
    for (int b = 0; b < B; ++b) {
        const float* bx = boxes + (boxes_batched ? (size_t)b * N * 4 : 0);
        for (int n = 0; n < N; ++n) {
            float ba = area4(bx + 4 * n);
            for (int g = 0; g < G; ++g) {
                const float* gb = gt + ((size_t)b * G + g) * 4;
                out[((size_t)b * N + n) * G + g] = iou_ref(bx + 4 * n, ba, gb, area4(gb));
            }
        }
    }
}

/* utils/bbox_utils.py:98-124 -> [dy, dx, dh, dw] */
static inline void encode_ref(const float* b, const float* g, float* d) {
    float bw = b[3] - b[1], bh = b[2] - b[0];
    float bcx = b[1] + 0.5f * bw, bcy = b[0] + 0.5f * bh;
    float gw = g[3] - g[1], gh = g[2] - g[0];
    float gcx = g[1] + 0.5f * gw, gcy = g[0] + 0.5f * gh;
    if (bw == 0.0f) bw = 1e-3f;
    if (bh == 0.0f) bh = 1e-3f;
    d[1] = (gw == 0.0f) ? 0.0f : (gcx - bcx) / bw;
    d[0] = (gh == 0.0f) ? 0.0f : (gcy - bcy) / bh;
    d[3] = (gw == 0.0f) ? 0.0f : logf(gw / bw);
    d[2] = (gh == 0.0f) ? 0.0f : logf(gh / bh);
}

/* utils/bbox_utils.py:72-96 */
static inline void decode_ref(const float* a, const float* d, float* o) {
    float aw = a[3] - a[1], ah = a[2] - a[0];
    float acx = a[1] + 0.5f * aw, acy = a[0] + 0.5f * ah;
    float w = expf(d[3]) * aw, h = expf(d[2]) * ah;
    float cx = (d[1] * aw) + acx, cy = (d[0] * ah) + acy;
    o[0] = cy - (0.5f * h);
    o[1] = cx - (0.5f * w);
    o[2] = h + o[0];
    o[3] = w + o[1];
}

API void oracle_decode(const float* anchors, int anchors_batched, const float* deltas, int B, int N, float* out) {
    for (size_t i = 0; i < (size_t)B * N; ++i)
        decode_ref(anchors + 4 * (anchors_batched ? i : i % N), deltas + 4 * i, out + 4 * i);
}

API void oracle_encode(const float* boxes, int boxes_batched, const float* gt, int B, int N, float* out) {
    for (size_t i = 0; i < (size_t)B * N; ++i)
        encode_ref(boxes + 4 * (boxes_batched ? i : i % N), gt + 4 * i, out + 4 * i);
}

/* Philox4x32-10, identical to oracle/rpn_oracle.py:philox4x32_10 */
static void philox(uint32_t c[4], uint32_t k0, uint32_t k1) {
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
        c[1] = (uint32_t)p1; c[3] = (uint32_t)p0; c[0] = n0; c[2] = n2;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}
static uint32_t sampling_key(uint32_t n, uint32_t img, uint64_t seed, uint64_t offset, int word) {
    uint32_t c[4] = {n, img, (uint32_t)offset, (uint32_t)(offset >> 32)};
    philox(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    return c[word];
}

typedef struct { uint32_t key; int32_t idx; } cand_t;
static int cmp_cand(const void* a, const void* b) {   /* key descending, index ascending */
    const cand_t *x = a, *y = b;
    if (x->key != y->key) return x->key > y->key ? -1 : 1;
    return (x->idx > y->idx) - (x->idx < y->idx);
}

/* utils/train_utils.py:50-65 semantics on one row: keep min(#True, quota) by (key desc, idx asc) */
static int select_row(const uint8_t* mask, int N, int quota, uint32_t img, uint64_t seed, uint64_t offset,
                      int word, uint8_t* out, cand_t* scratch) {
    int m = 0;
    memset(out, 0, (size_t)N);
    for (int n = 0; n < N; ++n) if (mask[n]) { scratch[m].idx = n; ++m; }
    if (quota <= 0 || m == 0) return 0;
    if (m <= quota) { for (int i = 0; i < m; ++i) out[scratch[i].idx] = 1; return m; }
    for (int i = 0; i < m; ++i) scratch[i].key = sampling_key((uint32_t)scratch[i].idx, img, seed, offset, word);
    qsort(scratch, (size_t)m, sizeof(cand_t), cmp_cand);
    for (int i = 0; i < quota; ++i) out[scratch[i].idx] = 1;
    return quota;
}

API void oracle_select_mask(const uint8_t* mask, const int32_t* select, int n_select, int B, int N, uint64_t seed,
                            uint64_t offset, int word, int image_offset, uint8_t* out) {
    cand_t* scratch = malloc(sizeof(cand_t) * (size_t)(N > 0 ? N : 1));
    for (int b = 0; b < B; ++b)
        select_row(mask + (size_t)b * N, N, select[n_select == 1 ? 0 : b], (uint32_t)(image_offset + b), seed, offset,
                   word, out + (size_t)b * N, scratch);
    free(scratch);
}

/* utils/train_utils.py:84-144.  labels (B,N) in {1,0,-1}; deltas (B,N,4).  dbg arrays may be NULL. */
API void oracle_rpn_targets(const float* anchors, const float* gt_boxes, const int32_t* gt_labels, int B, int N, int G,
                            float pos_thr, float neg_thr, int total_pos, int total_neg, const float* variances,
                            uint64_t seed, uint64_t offset, int image_offset, int threads, float* deltas,
                            float* labels, int32_t* argmax_row_out, int32_t* argmax_col_out, float* max_iou_out,
                            uint8_t* pos_pre_out, uint8_t* neg_pre_out) {
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#endif
    float* aarea = malloc(sizeof(float) * (size_t)N);
    for (int n = 0; n < N; ++n) aarea[n] = area4(anchors + 4 * n);
#pragma omp parallel for schedule(dynamic, 1)
    for (int b = 0; b < B; ++b) {
        const float* gt = gt_boxes + (size_t)b * G * 4;
        float* garea = malloc(sizeof(float) * (size_t)G);
        float* colmax = malloc(sizeof(float) * (size_t)G);
        int32_t* colarg = malloc(sizeof(int32_t) * (size_t)G);
        float* rowmax = malloc(sizeof(float) * (size_t)N);
        int32_t* rowarg = malloc(sizeof(int32_t) * (size_t)N);
        uint8_t* pos = malloc((size_t)N), *neg = malloc((size_t)N), *sel = malloc((size_t)N);
        cand_t* scratch = malloc(sizeof(cand_t) * (size_t)N);
        for (int g = 0; g < G; ++g) { garea[g] = area4(gt + 4 * g); colmax[g] = -INFINITY; colarg[g] = 0; }
        for (int n = 0; n < N; ++n) {                                   /* :106-112 */
            float best = -INFINITY; int arg = 0;
            for (int g = 0; g < G; ++g) {
                float v = iou_ref(anchors + 4 * n, aarea[n], gt + 4 * g, garea[g]);
                if (v != v) v = -INFINITY;                              /* NaN never wins '>' */
                if (v > best) { best = v; arg = g; }                    /* first maximal index */
                if (v > colmax[g]) { colmax[g] = v; colarg[g] = n; }
            }
            rowmax[n] = best; rowarg[n] = arg;
            pos[n] = best > pos_thr;                                    /* :114 */
        }
        for (int g = 0; g < G; ++g)                                     /* :116-122 */
            if (gt_labels[(size_t)b * G + g] != -1) pos[colarg[g]] = 1;
        if (pos_pre_out) memcpy(pos_pre_out + (size_t)b * N, pos, (size_t)N);
        int pos_count = select_row(pos, N, total_pos, (uint32_t)(image_offset + b), seed, offset, 0, sel, scratch);
        memcpy(pos, sel, (size_t)N);                                    /* :123 */
        for (int n = 0; n < N; ++n) neg[n] = (rowmax[n] < neg_thr) && !pos[n];   /* :128 */
        if (neg_pre_out) memcpy(neg_pre_out + (size_t)b * N, neg, (size_t)N);
        select_row(neg, N, (total_pos + total_neg) - pos_count, (uint32_t)(image_offset + b), seed, offset, 1, sel,
                   scratch);                                            /* :126,:129 */
        for (int n = 0; n < N; ++n) {
            size_t o = (size_t)b * N + n;
            labels[o] = (pos[n] ? 1.0f : -1.0f) + (sel[n] ? 1.0f : 0.0f);   /* :131-133 */
            float d[4] = {0.f, 0.f, 0.f, 0.f};
            if (pos[n]) {                                               /* :135-139 */
                encode_ref(anchors + 4 * n, gt + 4 * rowarg[n], d);
                for (int c = 0; c < 4; ++c) d[c] = d[c] / variances[c];
            }
            memcpy(deltas + 4 * o, d, sizeof(d));
            if (argmax_row_out) argmax_row_out[o] = rowarg[n];
            if (max_iou_out) max_iou_out[o] = rowmax[n];
        }
        if (argmax_col_out) memcpy(argmax_col_out + (size_t)b * G, colarg, sizeof(int32_t) * (size_t)G);
        free(garea); free(colmax); free(colarg); free(rowmax); free(rowarg); free(pos); free(neg); free(sel); free(scratch);
    }
    free(aarea);
}

/* tf.nn.top_k order [TF-internal]: score descending, equal scores -> lower index first */
typedef struct { float s; int32_t idx; } sc_t;
static int cmp_score(const void* a, const void* b) {
    const sc_t *x = a, *y = b;
    if (x->s != y->s) return x->s > y->s ? -1 : 1;
    return (x->idx > y->idx) - (x->idx < y->idx);
}

API void oracle_topk(const float* scores, int B, int N, int k, float* values, int32_t* indices) {
    sc_t* v = malloc(sizeof(sc_t) * (size_t)(N > 0 ? N : 1));
    for (int b = 0; b < B; ++b) {
        for (int n = 0; n < N; ++n) { v[n].s = scores[(size_t)b * N + n]; v[n].idx = n; }
        qsort(v, (size_t)N, sizeof(sc_t), cmp_score);
        for (int r = 0; r < k; ++r) { values[(size_t)b * k + r] = v[r].s; indices[(size_t)b * k + r] = v[r].idx; }
    }
    free(v);
}

/* IOU of TF's CombinedNonMaxSuppression kernel [TF-internal] */
static float nms_iou(const float* bi, const float* bj) {
    float ymin_i = fminf(bi[0], bi[2]), xmin_i = fminf(bi[1], bi[3]);
    float ymax_i = fmaxf(bi[0], bi[2]), xmax_i = fmaxf(bi[1], bi[3]);
    float ymin_j = fminf(bj[0], bj[2]), xmin_j = fminf(bj[1], bj[3]);
    float ymax_j = fmaxf(bj[0], bj[2]), xmax_j = fmaxf(bj[1], bj[3]);
    float area_i = (ymax_i - ymin_i) * (xmax_i - xmin_i);
    float area_j = (ymax_j - ymin_j) * (xmax_j - xmin_j);
    if (area_i <= 0 || area_j <= 0) return 0.0f;
    float iy = fmaxf(fminf(ymax_i, ymax_j) - fmaxf(ymin_i, ymin_j), 0.0f);
    float ix = fmaxf(fminf(xmax_i, xmax_j) - fmaxf(xmin_i, xmin_j), 0.0f);
    float inter = iy * ix;
    return inter / (area_i + area_j - inter);
}

/* greedy NMS over candidates already in score order; returns #kept; keep[] = positions in order[] */
static int nms_sorted(const float* boxes, const sc_t* order, int n_cand, int max_out, float thr, int32_t* keep) {
    int nk = 0;
    for (int c = 0; c < n_cand && nk < max_out; ++c) {
        const float* bc = boxes + 4 * (size_t)order[c].idx;
        int ok = 1;
        for (int j = nk - 1; j >= 0; --j)                     /* newest first, like TF */
            if (nms_iou(bc, boxes + 4 * (size_t)order[keep[j]].idx) > thr) { ok = 0; break; }
        if (ok) keep[nk++] = c;
    }
    return nk;
}

/* tf.image.combined_non_max_suppression, q = classes = 1 (utils/bbox_utils.py:48-70) */
API void oracle_nms(const float* boxes, const float* scores, int B, int K, int per_class, int total, float thr,
                    float score_thr, int pad_per_class, int clip_boxes, float* out_boxes, float* out_scores,
                    float* out_classes, int32_t* valid, int32_t* keep_idx) {
    int rows = pad_per_class ? (total < per_class ? total : per_class) : total;
    int max_out = per_class < rows ? per_class : rows;
    sc_t* v = malloc(sizeof(sc_t) * (size_t)(K > 0 ? K : 1));
    int32_t* keep = malloc(sizeof(int32_t) * (size_t)(max_out > 0 ? max_out : 1));
    for (int b = 0; b < B; ++b) {
        const float* bx = boxes + (size_t)b * K * 4;
        int m = 0;
        for (int n = 0; n < K; ++n)
            if (scores[(size_t)b * K + n] > score_thr) { v[m].s = scores[(size_t)b * K + n]; v[m].idx = n; ++m; }
        qsort(v, (size_t)m, sizeof(sc_t), cmp_score);
        int nk = nms_sorted(bx, v, m, max_out, thr, keep);
        for (int r = 0; r < rows; ++r) {
            size_t o = (size_t)b * rows + r;
            if (r < nk) {
                const float* s = bx + 4 * (size_t)v[keep[r]].idx;
                for (int c = 0; c < 4; ++c) out_boxes[4 * o + c] = clip_boxes ? fminf(fmaxf(s[c], 0.f), 1.f) : s[c];
                out_scores[o] = v[keep[r]].s;
                if (keep_idx) keep_idx[o] = v[keep[r]].idx;
            } else {
                memset(out_boxes + 4 * o, 0, 16);
                out_scores[o] = 0.f;
                if (keep_idx) keep_idx[o] = -1;
            }
            if (out_classes) out_classes[o] = 0.f;
        }
        valid[b] = nk;
    }
    free(v); free(keep);
}

/* composed proposal stage: predictor.py:52-60 -> clip -> NMS */
API void oracle_proposals(const float* rpn_reg, const float* rpn_cls, const float* anchors, int B, int N,
                          const float* variances, int pre_nms_topn, int post_nms_topn, float thr, int clip, int threads,
                          float* out_boxes, float* out_scores, int32_t* valid, int32_t* keep_idx) {
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#endif
    int k = pre_nms_topn < N ? pre_nms_topn : N;
#pragma omp parallel for schedule(dynamic, 1)
    for (int b = 0; b < B; ++b) {
        float* boxes = malloc(sizeof(float) * 4 * (size_t)N);
        sc_t* v = malloc(sizeof(sc_t) * (size_t)N);
        int32_t* keep = malloc(sizeof(int32_t) * (size_t)post_nms_topn);
        for (int n = 0; n < N; ++n) {
            float d[4];
            const float* r = rpn_reg + ((size_t)b * N + n) * 4;
            for (int c = 0; c < 4; ++c) d[c] = r[c] * variances[c];          /* predictor.py:55 */
            decode_ref(anchors + 4 * n, d, boxes + 4 * n);                   /* predictor.py:56 */
            if (clip) for (int c = 0; c < 4; ++c) boxes[4 * n + c] = fminf(fmaxf(boxes[4 * n + c], 0.f), 1.f);
            v[n].s = rpn_cls[(size_t)b * N + n]; v[n].idx = n;
        }
        qsort(v, (size_t)N, sizeof(sc_t), cmp_score);                        /* predictor.py:58 */
        int nk = nms_sorted(boxes, v, k, post_nms_topn, thr, keep);
        for (int r = 0; r < post_nms_topn; ++r) {
            size_t o = (size_t)b * post_nms_topn + r;
            if (r < nk) {
                const float* s = boxes + 4 * (size_t)v[keep[r]].idx;
                for (int c = 0; c < 4; ++c) out_boxes[4 * o + c] = fminf(fmaxf(s[c], 0.f), 1.f);
                out_scores[o] = v[keep[r]].s;
                if (keep_idx) keep_idx[o] = v[keep[r]].idx;
            } else {
                memset(out_boxes + 4 * o, 0, 16);
                out_scores[o] = 0.f;
                if (keep_idx) keep_idx[o] = -1;
            }
        }
        valid[b] = nk;
        free(boxes); free(v); free(keep);
    }
}

API int oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
