/* tfrpn.h -- C ABI of libtfrpn_cuda.so: the B200 (sm_100a) box hot path of FurkanOM/tf-rpn.
 *
 * The reference has no FFI: its hot path is plain Python over TensorFlow eager ops.  Each
 * entry point below replaces one reference function (cited file:line, relative to the
 * reference root) and is what a ctypes binding inside the reference's own modules would
 * call.  Conventions:
 *   - all tensors float32 (or int32 where stated), row-major, contiguous, 16-byte aligned;
 *     a box is [y1, x1, y2, x2], normalised to [0,1];
 *   - pointers are DEVICE pointers unless the name ends in _host;
 *   - every function returns 0 (TFRPN_OK) or a negative tfrpn_status; the message is in
 *     the thread-local tfrpn_last_error();
 *   - work is enqueued on the caller's stream (0 = legacy default stream); no hidden
 *     device synchronisation except in the *_host entry points, which return after the
 *     results are in the caller's host buffers;
 *   - the caller owns every buffer; the library owns only the per-handle workspace;
 *   - there is no CPU fallback: without a CUDA device every call fails with TFRPN_ERR_CUDA.
 */
#ifndef TFRPN_H_
#define TFRPN_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define TFRPN_API __attribute__((visibility("default")))
#else
#define TFRPN_API
#endif

#define TFRPN_VERSION 100 /* 0.1.0 */
#define TFRPN_MAX_BASE_ANCHORS 64

typedef enum {
    TFRPN_OK = 0,
    TFRPN_ERR_BAD_ARG = -1,     /* null pointer, bad shape, bad config value        */
    TFRPN_ERR_MISALIGNED = -2,  /* a box pointer is not 16-byte aligned              */
    TFRPN_ERR_CUDA = -3,        /* CUDA runtime error (text in tfrpn_last_error)     */
    TFRPN_ERR_WORKSPACE = -4,   /* workspace must grow but the stream is capturing   */
    TFRPN_ERR_UNSUPPORTED = -5  /* valid request outside what this build implements  */
} tfrpn_status;

typedef void* tfrpn_stream;            /* a cudaStream_t */
typedef struct tfrpn_ctx* tfrpn_handle; /* per-device workspace + staging; one thread at a time */

/* ---- hyper_params dict, utils/train_utils.py:5-38 --------------------------------- */
typedef struct {
    int32_t img_h, img_w;  /* img_size (reference: one int for both)                    */
    int32_t fm_h, fm_w;    /* feature_map_shape (reference: one int for both)           */
    int32_t n_scales, n_ratios;
    double scales[8];      /* anchor_scales, pixels                                      */
    double ratios[8];      /* anchor_ratios (Python floats = doubles)                    */
} tfrpn_anchor_cfg;

typedef struct {
    float pos_iou_threshold; /* 0.7f  utils/train_utils.py:114 */
    float neg_iou_threshold; /* 0.3f  utils/train_utils.py:128 */
    int32_t total_pos;       /* total_pos_bboxes */
    int32_t total_neg;       /* total_neg_bboxes */
    float variances[4];      /* [0.1,0.1,0.2,0.2]; deltas are DIVIDED by these (:139) */
    uint64_t seed;           /* counter-RNG key                                          */
    uint64_t offset;         /* counter-RNG offset (bump per step; resume = same value)  */
    int32_t image_offset;    /* global index of image 0 of this shard (multi-GPU)        */
    int32_t reserved;
} tfrpn_target_cfg;

/* optional intermediates of target assignment; any pointer may be NULL */
typedef struct {
    int32_t* argmax_row;  /* (B,N)  per-anchor best GT   utils/train_utils.py:108 */
    int32_t* argmax_col;  /* (B,G)  per-GT best anchor   utils/train_utils.py:110 */
    float* max_iou;       /* (B,N)                       utils/train_utils.py:112 */
    uint8_t* pos_pre;     /* (B,N)  positives before sampling   (:114-122)       */
    uint8_t* neg_pre;     /* (B,N)  negative candidates         (:128)           */
    int32_t* pos_count;   /* (B,)                                (:125)           */
    int32_t* neg_count;   /* (B,)                                                 */
} tfrpn_target_debug;

typedef struct {
    int32_t max_output_size_per_class;
    int32_t max_total_size;
    float iou_threshold;   /* suppress iff IoU > thr (strict)                */
    float score_threshold; /* candidates need score > thr; -INFINITY = all  */
    int32_t pad_per_class; /* TF flag; output rows = pad ? min(total, per_class) : total */
    int32_t clip_boxes;    /* clip OUTPUT boxes to [0,1]                     */
    int32_t pre_nms_topn;  /* > 0: only the top-n scores compete (tf.nn.top_k + gather of predictor.py:58-60
                              fused in front of the NMS, consumed lazily); 0 = all K boxes */
} tfrpn_nms_cfg;

typedef struct {
    float variances[4];   /* deltas are MULTIPLIED by these, predictor.py:55 */
    int32_t pre_nms_topn; /* k of tf.nn.top_k, predictor.py:58 (BASELINE: 6000) */
    int32_t post_nms_topn;/* test_nms_topn = 300, utils/train_utils.py:29   */
    float nms_iou_threshold; /* 0.7 */
    int32_t clip;         /* clip decoded boxes to [0,1] before NMS (north star) */
} tfrpn_proposal_cfg;

/* ---- library ---------------------------------------------------------------------- */
TFRPN_API int tfrpn_version(void);
TFRPN_API const char* tfrpn_last_error(void);
TFRPN_API int tfrpn_create(tfrpn_handle* out, int device /* -1 = current */);
TFRPN_API int tfrpn_destroy(tfrpn_handle h);
/* Pre-size the workspace (needed before CUDA-graph capture; otherwise it grows lazily). */
TFRPN_API int tfrpn_reserve(tfrpn_handle h, int B, int N, int G, int k);
TFRPN_API size_t tfrpn_workspace_bytes(int B, int N, int G, int k);
/* number of kernels launched by this library on the calling thread since process start */
TFRPN_API uint64_t tfrpn_launch_count(void);

/* ---- tracing (the reference has none): when enabled, every kernel launched through
 *      this handle is bracketed by CUDA events on its stream.  Not usable during graph capture. */
enum { TFRPN_K_IOU_ARGMAX = 0, TFRPN_K_LABEL_ENCODE = 1, TFRPN_K_SELECT_MASK = 2, TFRPN_K_PROPOSAL = 3,
       TFRPN_K_LOSS = 4, TFRPN_K_PROPOSAL_CLUSTER = 5, TFRPN_K_NMS_MASK = 6, TFRPN_K_NMS_SWEEP = 7, TFRPN_K_COUNT = 8 };
TFRPN_API int tfrpn_profile_enable(tfrpn_handle h, int on);
/* synchronises, then returns the summed device time and launch count of one kernel id and clears them */
TFRPN_API int tfrpn_profile_read(tfrpn_handle h, int kernel_id, double* total_ms, int* launches);
TFRPN_API const char* tfrpn_kernel_name(int kernel_id);

/* ---- anchors: utils/bbox_utils.py:3-21 and :23-46 ---------------------------------- */
TFRPN_API int tfrpn_base_anchors_host(const tfrpn_anchor_cfg* cfg, float* out_host /* (A,4) */);
TFRPN_API int tfrpn_anchors(const tfrpn_anchor_cfg* cfg, float* out /* (fm_h*fm_w*A,4) */, tfrpn_stream s);

/* ---- generate_iou_map: utils/bbox_utils.py:126-150 --------------------------------- */
TFRPN_API int tfrpn_iou_map(const float* boxes /* (N,4) or (B,N,4) */, int boxes_batched,
                  const float* gt_boxes /* (B,G,4) */, int B, int N, int G,
                  float* out /* (B,N,G) */, tfrpn_stream s);

/* Self-test: the IoU kernels divide inter / union with the fast path of div.rn.f32 minus its range check
 * when every box is in a range where that is exact ("nice" boxes, see common.cuh).  This compares that
 * sequence with __fdiv_rn on n_pairs pseudo-random operand pairs of its domain and writes the number of
 * results that differ in any bit (must be 0) to *mismatches_dev (device). */
TFRPN_API int tfrpn_selftest_division(uint64_t n_pairs, uint64_t seed, uint64_t* mismatches_dev, tfrpn_stream s);

/* ---- get_deltas_from_bboxes: utils/bbox_utils.py:98-124 (no variance scaling) ------ */
TFRPN_API int tfrpn_encode_deltas(const float* boxes /* (N,4) or (B,N,4) */, int boxes_batched,
                        const float* gt_boxes /* (B,N,4) */, int B, int N,
                        float* out /* (B,N,4) */, tfrpn_stream s);

/* ---- get_bboxes_from_deltas: utils/bbox_utils.py:72-96; optionally the caller's
 *      `deltas *= variances` (predictor.py:55) and the clip of the proposal pipeline --- */
TFRPN_API int tfrpn_decode(const float* anchors /* (N,4) or (B,N,4) */, int anchors_batched,
                 const float* deltas /* (B,N,4) */, const float* variances_host_or_null /* [4] */,
                 int clip, int B, int N, float* out /* (B,N,4) */, tfrpn_stream s);

/* The same with the anchors regenerated in registers from the hyper-parameters (utils/bbox_utils.py:23-46 fused
 * in front of :72-96): no anchor tensor is read.  N = fm_h * fm_w * n_scales * n_ratios; out is (B,N,4). */
TFRPN_API int tfrpn_decode_anchor_cfg(const tfrpn_anchor_cfg* acfg, const float* deltas /* (B,N,4) */,
                            const float* variances_host_or_null /* [4] */, int clip, int B,
                            float* out /* (B,N,4) */, tfrpn_stream s);

/* ---- normalize_bboxes / denormalize_bboxes: utils/bbox_utils.py:152-182 ------------ */
TFRPN_API int tfrpn_scale_boxes(const float* boxes, int64_t n_boxes, float height, float width,
                      int denormalize /* 0: divide; 1: multiply then round-half-even */,
                      float* out, tfrpn_stream s);

/* ---- calculate_rpn_actual_outputs: utils/train_utils.py:84-144 --------------------- */
TFRPN_API int tfrpn_rpn_targets(tfrpn_handle h, const float* anchors /* (N,4) */,
                      const float* gt_boxes /* (B,G,4) */, const int32_t* gt_labels /* (B,G) */,
                      int B, int N, int G, const tfrpn_target_cfg* cfg,
                      float* deltas /* (B,N,4) */, float* labels /* (B,N) == (B,F,F,A) */,
                      const tfrpn_target_debug* dbg_or_null, tfrpn_stream s);

/* Compact form of bbox_deltas, for callers that move the results over PCIe: it is exactly 0 outside
 * the <= total_pos sampled positives of each image (utils/train_utils.py:137), so only those rows travel,
 * as (anchor index, row) pairs; the rows of an image are in no particular order and unused index slots
 * are -1.  bbox_labels is returned as usual.  tfrpn_expand_targets_host rebuilds the dense tensor on the
 * host: with prev_pos_idx (the pos_idx of the step whose rows `deltas` still holds, prev_total_pos
 * columns) only those rows are cleared first, otherwise the whole array is zeroed. */
TFRPN_API int tfrpn_rpn_targets_compact(tfrpn_handle h, const float* anchors, const float* gt_boxes,
                              const int32_t* gt_labels, int B, int N, int G, const tfrpn_target_cfg* cfg,
                              float* labels /* (B,N) */, int32_t* pos_idx /* (B,total_pos) */,
                              float* pos_deltas /* (B,total_pos,4) */, tfrpn_stream s);
/* Fully sparse form: bbox_labels is -1 everywhere except the <= total_pos + total_neg sampled entries
 * (utils/train_utils.py:126-133), so it travels as codes 2 * anchor + label (label 1 or 0), in no particular order,
 * unused slots -1: ~1 KB per image instead of 4 * N bytes.  tfrpn_expand_labels_host rebuilds the dense tensor; with
 * prev_codes (the codes of the step whose labels `labels` still holds) only those entries are reset first. */
TFRPN_API int tfrpn_rpn_targets_sparse(tfrpn_handle h, const float* anchors, const float* gt_boxes,
                             const int32_t* gt_labels, int B, int N, int G, const tfrpn_target_cfg* cfg,
                             int32_t* label_codes /* (B,total_pos+total_neg) */, int32_t* pos_idx /* (B,total_pos) */,
                             float* pos_deltas /* (B,total_pos,4) */, tfrpn_stream s);
TFRPN_API int tfrpn_expand_labels_host(const int32_t* label_codes, int B, int N, int Q,
                             const int32_t* prev_codes_or_null, int prev_Q, float* labels /* (B,N) host */);
TFRPN_API int tfrpn_expand_targets_host(const int32_t* pos_idx, const float* pos_deltas, int B, int N, int total_pos,
                              const int32_t* prev_pos_idx_or_null, int prev_total_pos,
                              float* deltas /* (B,N,4) host */);

/* ---- randomly_select_xyz_mask: utils/train_utils.py:50-65 (counter RNG) ------------ */
TFRPN_API int tfrpn_select_mask(tfrpn_handle h, const uint8_t* mask /* (B,N) 0/1 */,
                      const int32_t* select /* (n_select,) device; n_select = 1 or B */,
                      int n_select, int B, int N, uint64_t seed, uint64_t offset,
                      int rng_stream /* 0 = positives word, 1 = negatives word */,
                      int image_offset, uint8_t* out /* (B,N) */, tfrpn_stream s);

/* ---- cls_loss + reg_loss: utils/train_utils.py:146-161 and :163-185 -------------------
 * The consumers of tfrpn_rpn_targets' outputs.  cls: BinaryCrossentropy (Keras, probabilities,
 * eps 1e-7) averaged over the entries with true label != -1 (NaN when there are none, as TF);
 * reg: Huber(delta) summed over the 4 coordinates of the rows whose true delta is not all-zero,
 * divided by max(1, #rows).  Either pair may be NULL to compute only the other loss.  Optional
 * gradients with respect to the predictions (what TF autograd derives from the same ops). */
typedef struct {
    float reg_loss;  /* utils/train_utils.py:185 */
    float cls_loss;  /* utils/train_utils.py:161 */
    int32_t n_pos;   /* rows with a non-zero true delta (:180-184) */
    int32_t n_cls;   /* entries with label != -1 (:156)            */
} tfrpn_loss_out;
TFRPN_API int tfrpn_rpn_losses(tfrpn_handle h, const float* true_deltas /* (B,N,4) or NULL */,
                     const float* pred_deltas /* (B,N,4) == (B,F,F,4A) */,
                     const float* true_labels /* (B,N) == (B,F,F,A) or NULL */,
                     const float* pred_scores /* (B,N) */, int B, int N, float huber_delta /* 1.0f */,
                     tfrpn_loss_out* out /* device, 16 bytes */,
                     float* grad_deltas_or_null /* (B,N,4) */, float* grad_scores_or_null /* (B,N) */,
                     tfrpn_stream s);

/* ---- tf.nn.top_k + tf.gather(batch_dims=1): predictor.py:58-60 --------------------- */
TFRPN_API int tfrpn_topk(tfrpn_handle h, const float* scores /* (B,N) */, int B, int N, int k,
               float* values /* (B,k) */, int32_t* indices /* (B,k) */,
               const float* boxes_or_null /* (N,4) or (B,N,4) */, int boxes_batched,
               float* gathered_or_null /* (B,k,4) */, tfrpn_stream s);

/* ---- the predictor loop body, predictor.py:52-60, in one launch: reshape, deltas *= variances,
 *      get_bboxes_from_deltas, tf.nn.top_k(rpn_labels, k), tf.gather(batch_dims=1).  Only the k
 *      selected rows are decoded.  clip = 0 reproduces the reference (it never clips). --------- */
TFRPN_API int tfrpn_predict_topk(tfrpn_handle h, const float* rpn_reg /* (B,N,4) == (B,F,F,4A) */,
                       const float* rpn_cls /* (B,N) == (B,F,F,A) */, const float* anchors /* (N,4) */,
                       int B, int N, int k, const float* variances_host /* [4] */, int clip,
                       float* out_boxes /* (B,k,4) */, float* out_scores /* (B,k) */,
                       int32_t* out_indices /* (B,k) */, tfrpn_stream s);

/* ---- GT-side preprocessing: utils/data_utils.py:54-68 (flip_horizontally's box transform) and
 *      :145-157 (padded_batch with boxes 0 / labels -1).  Ragged input: image b owns rows
 *      offsets[b] .. offsets[b+1] of the flat arrays; rows beyond G are dropped. ---------------- */
TFRPN_API int tfrpn_pad_gt(const float* flat_boxes /* (M,4) */, const int32_t* flat_labels /* (M,) */,
                 const int32_t* offsets /* (B+1,) */, const uint8_t* flip_or_null /* (B,) 0/1 */,
                 int B, int G, int label_add /* data_utils.py:20 uses +1 */,
                 float* out_boxes /* (B,G,4) */, int32_t* out_labels /* (B,G) */, tfrpn_stream s);

/* ---- non_max_suppression: utils/bbox_utils.py:48-70 (1 class, q = 1) ---------------
 * rows = cfg->pad_per_class ? min(max_total_size, per_class) : max_total_size          */
TFRPN_API int tfrpn_nms(tfrpn_handle h, const float* boxes /* (B,K,4) */, const float* scores /* (B,K) */,
              int B, int K, const tfrpn_nms_cfg* cfg,
              float* out_boxes /* (B,rows,4) */, float* out_scores /* (B,rows) */,
              float* out_classes /* (B,rows), zeros */, int32_t* valid /* (B,) */,
              int32_t* keep_idx_or_null /* (B,rows), -1 padded */, tfrpn_stream s);

/* ---- composed proposal stage: predictor.py:52-60 -> clip ->
 *      bbox_utils.py:48-70 with k = pre_nms_topn, 300 @ 0.7 ---------------------------- */
TFRPN_API int tfrpn_proposals(tfrpn_handle h, const float* rpn_reg /* (B,N,4) == (B,F,F,4A) */,
                    const float* rpn_cls /* (B,N) == (B,F,F,A) */, const float* anchors /* (N,4) */,
                    int B, int N, const tfrpn_proposal_cfg* cfg,
                    float* out_boxes /* (B,post,4) */, float* out_scores /* (B,post) */,
                    int32_t* valid /* (B,) */, int32_t* keep_idx_or_null /* (B,post) */,
                    tfrpn_stream s);

/* ... with the anchors regenerated in registers (generate_anchors, utils/bbox_utils.py:23-46, fused into the
 * decode of the candidates NMS examines).  N = fm_h * fm_w * n_scales * n_ratios, below the large-N prefilter's
 * threshold (40000); larger feature maps use tfrpn_anchors + tfrpn_proposals. */
TFRPN_API int tfrpn_proposals_anchor_cfg(tfrpn_handle h, const float* rpn_reg, const float* rpn_cls,
                               const tfrpn_anchor_cfg* acfg, int B, const tfrpn_proposal_cfg* cfg,
                               float* out_boxes, float* out_scores, int32_t* valid, int32_t* keep_idx_or_null,
                               tfrpn_stream s);

/* ---- host-buffer entry points (what a NumPy / tf.numpy() caller binds): pinned staging,
 *      H2D, kernels, D2H, stream sync -- all inside the call ---------------------------- */
TFRPN_API int tfrpn_rpn_targets_host(tfrpn_handle h, const float* anchors_dev /* (N,4) device */,
                           const float* gt_boxes_host, const int32_t* gt_labels_host,
                           int B, int N, int G, const tfrpn_target_cfg* cfg,
                           float* deltas_host, float* labels_host, tfrpn_stream s);
TFRPN_API int tfrpn_proposals_host(tfrpn_handle h, const float* rpn_reg_host, const float* rpn_cls_host,
                         const float* anchors_dev /* (N,4) device */, int B, int N,
                         const tfrpn_proposal_cfg* cfg, float* out_boxes_host,
                         float* out_scores_host, int32_t* valid_host, int32_t* keep_idx_host_or_null,
                         tfrpn_stream s);
/* One step from host buffers, both halves at once: the proposal half runs on an internal stream so its
 * large H2D overlaps the target half's large D2H (full-duplex PCIe).  Same results as the two calls. */
TFRPN_API int tfrpn_rpn_step_host(tfrpn_handle h, const float* anchors_dev,
                        const float* gt_boxes_host, const int32_t* gt_labels_host, int B, int N, int G,
                        const tfrpn_target_cfg* tcfg, float* deltas_host, float* labels_host,
                        const float* rpn_reg_host, const float* rpn_cls_host, const tfrpn_proposal_cfg* pcfg,
                        float* out_boxes_host, float* out_scores_host, int32_t* valid_host,
                        int32_t* keep_idx_host_or_null, tfrpn_stream s);
/* ---- pipelined host steps: the generator call site (utils/train_utils.py:67-82, driven by Keras
 *      from trainer.py:48-49,64-69) and the predictor loop (predictor.py:48-60).  A step is PCIe-bound
 *      (~11 MB each way at C2), so `depth` steps are kept in flight: the H2D of step i+1 runs under
 *      the D2H of step i on the pipeline's own streams.  submit() only enqueues and returns a ticket;
 *      wait() returns when that step's results are in the caller's host buffers (tickets complete in
 *      order; submitting into a slot that is still busy retires its old step first).  Either half may
 *      be skipped: gt_boxes_host == NULL -> no target assignment, rpn_reg_host == NULL -> no
 *      proposals.  Host buffers must stay valid until wait(); page-locked ones (tfrpn_host_alloc) are
 *      copied directly, others through the slot's pinned staging.  While steps are in flight the
 *      handle's workspace belongs to the pipeline: do not call tfrpn_rpn_targets on it concurrently. */
typedef struct tfrpn_pipe* tfrpn_pipeline;
TFRPN_API int tfrpn_pipeline_create(tfrpn_handle h, int depth /* 1..16 */, tfrpn_pipeline* out);
TFRPN_API int tfrpn_pipeline_submit(tfrpn_pipeline p, const float* anchors_dev /* (N,4) device */, int B, int N,
                          const float* gt_boxes_host, const int32_t* gt_labels_host, int G,
                          const tfrpn_target_cfg* tcfg, float* deltas_host, float* labels_host,
                          const float* rpn_reg_host, const float* rpn_cls_host, const tfrpn_proposal_cfg* pcfg,
                          float* out_boxes_host, float* out_scores_host, int32_t* valid_host,
                          int32_t* keep_idx_host_or_null, int64_t* ticket_out);
/* Zero-copy variant: borrow the next slot's page-locked step buffers, fill the inputs in place (the
 * data loader writes its padded batch / the head outputs straight into them), submit, and after
 * wait() read the results in place (do not write to the result arrays: the library keeps them
 * consistent incrementally from step to step).  Inputs are one contiguous block and results another, so a step
 * is exactly ONE H2D and ONE D2H copy -- the pattern that reaches the link's duplex rate.  The
 * pointers stay valid until the slot is acquired again (depth steps later).  tcfg / pcfg NULL skips
 * that half. */
typedef struct {
    float* gt_boxes;    /* (B,G,4)  in  */
    int32_t* gt_labels; /* (B,G)    in  */
    float* rpn_reg;     /* (B,N,4)  in  */
    float* rpn_cls;     /* (B,N)    in  */
    float* deltas;      /* (B,N,4)  out */
    float* labels;      /* (B,N)    out */
    float* out_boxes;   /* (B,post,4) out */
    float* out_scores;  /* (B,post) out */
    int32_t* valid;     /* (B,)     out */
    int32_t* keep_idx;  /* (B,post) out */
} tfrpn_step_buffers;
TFRPN_API int tfrpn_pipeline_acquire(tfrpn_pipeline p, int B, int N, int G, int post_nms_topn, tfrpn_step_buffers* out);
TFRPN_API int tfrpn_pipeline_submit_acquired(tfrpn_pipeline p, const float* anchors_dev,
                                   const tfrpn_target_cfg* tcfg_or_null, const tfrpn_proposal_cfg* pcfg_or_null,
                                   int64_t* ticket_out);
TFRPN_API int tfrpn_pipeline_wait(tfrpn_pipeline p, int64_t ticket);
/* Options (set while no step is in flight; steps in flight are retired first).
 * TFRPN_PIPE_OPT_STABLE_OUTPUTS = 1: the caller promises that every bbox_deltas array it passes to
 * tfrpn_pipeline_submit is written by this pipeline only (the usual ring of output arrays).  An array the pipeline has
 * filled before then still holds zeros plus the rows of that step, so only those rows are reset instead of zeroing
 * 16 * B * N bytes per step.  Off by default: without the promise a stale array would keep foreign rows. */
enum { TFRPN_PIPE_OPT_STABLE_OUTPUTS = 1 };
TFRPN_API int tfrpn_pipeline_set_option(tfrpn_pipeline p, int option, int value);
/* bytes the last submitted acquired step copies in each direction (bbox_deltas travels in compact form,
 * see tfrpn_rpn_targets_compact, and is expanded into the slot's dense host array by wait()) */
TFRPN_API int tfrpn_pipeline_last_copy_bytes(tfrpn_pipeline p, int64_t* h2d_bytes, int64_t* d2h_bytes);
/* Tracing (handles created with TFRPN_PIPE_TRACE=1 in the environment): device time stamps of step `ticket`, ms
 * since the pipeline was created -- [0,1] H2D begin/end, [2,3] target kernels, [4,5] proposal stage (rank launch,
 * host gather, NMS launch), [6,7] D2H + expansion; then host durations in ms: [8] the row gather, [9] the
 * expansion of the compact bbox_deltas.  Valid after wait(ticket) until the slot's next step is retired. */
TFRPN_API int tfrpn_pipeline_trace(tfrpn_pipeline p, int64_t ticket, float* ms10);
TFRPN_API int tfrpn_pipeline_drain(tfrpn_pipeline p); /* wait for every step in flight */
TFRPN_API int tfrpn_pipeline_destroy(tfrpn_pipeline p);
/* page-locked host memory for the caller's batches (so H2D/D2H run at full PCIe rate) */
TFRPN_API int tfrpn_host_alloc(void** out, size_t bytes);
TFRPN_API int tfrpn_host_free(void* p);

#ifdef __cplusplus
}
#endif
#endif /* TFRPN_H_ */
