// pipeline.cu -- host-buffer steps of the hot path, pipelined over PCIe.
//
// This is the call site the reference drives from Keras' generator thread (utils/train_utils.py:67-82,
// trainer.py:48-49,64-69): every step a padded host batch goes in and (bbox_deltas, bbox_labels) come
// out; predictor.py:48-60 does the same with the head outputs.  A step moves ~11 MB each way at C2
// while its kernels take ~0.1 ms, so the step is PCIe-bound and the job of this file is to keep BOTH
// directions of the link busy:
//
//   stream in   : H2D gt (tiny), then rpn_reg / rpn_cls chunk by chunk
//   stream tgt  : target kernels of chunk c            (needs gt)
//   stream prop : proposal kernel of chunk c           (needs its reg/cls chunk)
//   stream out  : D2H deltas/labels of chunk c as soon as its kernels are done, then the small
//                 proposal results
//
// All four streams are FIFO across steps, and each in-flight step owns a SLOT (device + pinned staging,
// events), so with depth >= 2 the H2D of step i+1 runs under the D2H of step i: the steady-state cost
// of a step is max(H2D, D2H, kernels) instead of their sum.  Results are bit-identical to the separate
// calls: images are independent and the counter RNG is keyed by the global image index.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace tfrpn {
size_t targets_workspace_bytes(int B, int N, int G);  // targets.cu
int launch_targets(tfrpn_handle h, const float* anchors, const float* gt_boxes, const int32_t* gt_labels, int B, int N,
                   int G, const tfrpn_target_cfg* cfg, float* deltas, float* labels, int32_t* pos_idx,
                   float* pos_deltas, const tfrpn_target_debug* dbg, tfrpn_stream s);
int proposals_enqueue(tfrpn_handle h, const float* rpn_reg, const float* rpn_cls, const float* anchors, int B, int N,
                      const tfrpn_proposal_cfg* cfg, float* out_boxes, float* out_scores, int32_t* valid,
                      int32_t* keep_idx_or_null, unsigned long long* rows_fetched_or_null, cudaStream_t st);   // proposals.cu
}

namespace {

constexpr int MAX_CHUNKS = 8;
constexpr int MAX_DEPTH = 8;
constexpr int COMPACT_MAX_POS = 256;   // compact result form is used when total_pos_bboxes <= this

size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

bool is_pinned(const void* p) {
    cudaPointerAttributes attr;
    bool r = cudaPointerGetAttributes(&attr, p) == cudaSuccess && attr.type == cudaMemoryTypeHost;
    cudaGetLastError();
    return r;
}

struct Slot {
    char* dev = nullptr; size_t dev_bytes = 0;
    char* pin = nullptr; size_t pin_bytes = 0;
    cudaEvent_t ev_gt = nullptr, ev_prop = nullptr, ev_done = nullptr;
    cudaEvent_t ev_in[MAX_CHUNKS] = {}, ev_tgt[MAX_CHUNKS] = {};
    long long ticket = -1;          // ticket in flight in this slot (-1 = free)
    cudaEvent_t tr[8] = {};         // TFRPN_PIPE_TRACE: timing events (h2d, targets, proposals, d2h: begin / end)
    float tr_ms[8] = {};            // ... of the last step retired from this slot, ms since the pipeline was created
    long long tr_ticket = -1;
    // compact results (acquired mode): bbox_deltas comes back as its <= total_pos non-zero rows per image and is
    // expanded into the slot's dense host array when the step is retired (the labels travel as they are)
    bool pulled = false;            // the step in flight pulls rpn_reg rows from the pinned block (row count at off_pc)
    size_t off_pc = 0;
    bool compact = false;           // the step in flight returns compact targets
    int cB = 0, cN = 0, cTP = 0;    // its shape
    size_t off_d = 0, off_ci = 0, off_cd = 0;
    bool dense_clean = false;       // pin + off_d holds zeros except the rows listed in prev_idx
    int pB = 0, pN = 0, pTP = 0;
    size_t p_off_d = 0;
    std::vector<int32_t> prev_idx;
    struct Copy { void* dst; const void* src; size_t bytes; } copies[8 + 2 * MAX_CHUNKS];
    int n_copies = 0;
    void defer(void* d, const void* s, size_t b) { copies[n_copies].dst = d; copies[n_copies].src = s; copies[n_copies].bytes = b; ++n_copies; }
};

}  // namespace

struct tfrpn_pipe {
    tfrpn_handle h = nullptr;
    int depth = 1;
    cudaStream_t s_in = nullptr, s_tgt = nullptr, s_prop = nullptr, s_out = nullptr;
    cudaEvent_t ev_after = nullptr;
    bool trace = false;             // TFRPN_PIPE_TRACE=1 (read when the handle was created)
    cudaEvent_t ev_base = nullptr;  // time origin of the trace
    Slot slots[MAX_DEPTH];
    long long next_ticket = 0;
    int acq_B = 0, acq_N = 0, acq_G = 0, acq_P = 0;   // shape of the slot handed out by the last acquire()
    bool acq_live = false;
    long long last_h2d = 0, last_d2h = 0;             // bytes copied by the last submitted step
    long long last_pulled = 0;                        // bytes of rpn_reg rows pulled by the kernels of the last RETIRED step
};

namespace tfrpn {

void pipe_destroy(tfrpn_pipe* p) {
    if (!p) return;
    DeviceGuard guard(p->h->device);
    for (cudaStream_t s : {p->s_in, p->s_tgt, p->s_prop, p->s_out})
        if (s) { cudaStreamSynchronize(s); cudaStreamDestroy(s); }
    if (p->ev_after) cudaEventDestroy(p->ev_after);
    if (p->ev_base) cudaEventDestroy(p->ev_base);
    for (int i = 0; i < p->depth; ++i) {
        Slot& s = p->slots[i];
        if (s.dev) cudaFree(s.dev);
        if (s.pin) cudaFreeHost(s.pin);
        for (cudaEvent_t e : {s.ev_gt, s.ev_prop, s.ev_done}) if (e) cudaEventDestroy(e);
        for (cudaEvent_t e : s.tr) if (e) cudaEventDestroy(e);
        for (int c = 0; c < MAX_CHUNKS; ++c) {
            if (s.ev_in[c]) cudaEventDestroy(s.ev_in[c]);
            if (s.ev_tgt[c]) cudaEventDestroy(s.ev_tgt[c]);
        }
    }
    cudaGetLastError();
    delete p;
}

static int pipe_create(tfrpn_handle h, int depth, tfrpn_pipe** out) {
    if (!h || !out) return fail(TFRPN_ERR_BAD_ARG, "pipeline_create: null pointer");
    if (depth < 1 || depth > MAX_DEPTH) return fail(TFRPN_ERR_BAD_ARG, "pipeline_create: depth must be 1..%d", MAX_DEPTH);
    TFRPN_ENTER(h);
    tfrpn_pipe* p = new tfrpn_pipe();
    p->h = h;
    p->depth = depth;
    cudaError_t e = cudaSuccess;
    for (cudaStream_t* s : {&p->s_in, &p->s_tgt, &p->s_prop, &p->s_out})
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(s, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->ev_after, cudaEventDisableTiming);
    for (int i = 0; i < depth && e == cudaSuccess; ++i) {
        Slot& s = p->slots[i];
        for (cudaEvent_t* ev : {&s.ev_gt, &s.ev_prop, &s.ev_done})
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(ev, cudaEventDisableTiming);
        for (int c = 0; c < MAX_CHUNKS && e == cudaSuccess; ++c) {
            e = cudaEventCreateWithFlags(&s.ev_in[c], cudaEventDisableTiming);
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s.ev_tgt[c], cudaEventDisableTiming);
        }
    }
    p->trace = h->opts.pipe_trace;
    if (p->trace && e == cudaSuccess) {
        e = cudaEventCreate(&p->ev_base);
        if (e == cudaSuccess) e = cudaEventRecord(p->ev_base, p->s_in);
        for (int i = 0; i < depth; ++i)
            for (cudaEvent_t& ev : p->slots[i].tr) if (e == cudaSuccess) e = cudaEventCreate(&ev);
    }
    if (e != cudaSuccess) { pipe_destroy(p); return cuda_fail(e, "pipeline_create"); }
    *out = p;
    return 0;
}

// block until the step in `slot` has landed in the caller's buffers
static int slot_finish(tfrpn_pipe* p, Slot& s) {
    if (s.ticket < 0) return 0;
    TFRPN_CHECK_CUDA(cudaEventSynchronize(s.ev_done));
    if (p->trace) {
        for (int i = 0; i < 8; ++i)
            if (cudaEventElapsedTime(&s.tr_ms[i], p->ev_base, s.tr[i]) != cudaSuccess) { s.tr_ms[i] = -1.0f; cudaGetLastError(); }
        s.tr_ticket = s.ticket;
    }
    for (int i = 0; i < s.n_copies; ++i) memcpy(s.copies[i].dst, s.copies[i].src, s.copies[i].bytes);
    s.n_copies = 0;
    if (s.pulled) {
        p->last_pulled = 16LL * (long long)*reinterpret_cast<const unsigned long long*>(s.pin + s.off_pc);
        s.pulled = false;
    }
    if (s.compact) {   // expand into the dense (B,N,4) / (B,N) host arrays of this slot
        const bool reuse = s.dense_clean && s.pB == s.cB && s.pN == s.cN && s.p_off_d == s.off_d;
        const int32_t* idx = reinterpret_cast<const int32_t*>(s.pin + s.off_ci);
        if (int rc = tfrpn_expand_targets_host(idx, reinterpret_cast<const float*>(s.pin + s.off_cd), s.cB, s.cN, s.cTP,
                                               reuse ? s.prev_idx.data() : nullptr, reuse ? s.pTP : 0,
                                               reinterpret_cast<float*>(s.pin + s.off_d))) return rc;
        s.prev_idx.assign(idx, idx + (size_t)s.cB * s.cTP);
        s.pB = s.cB; s.pN = s.cN; s.pTP = s.cTP; s.p_off_d = s.off_d;
        s.dense_clean = true;
        s.compact = false;
    }
    s.ticket = -1;
    return 0;
}

struct StepArgs {
    const float* anchors_dev;
    // target half (skipped when gt_boxes is null)
    const float* gt_boxes; const int32_t* gt_labels; int B, N, G; const tfrpn_target_cfg* tcfg;
    float* deltas; float* labels;
    // proposal half (skipped when rpn_reg is null)
    const float* rpn_reg; const float* rpn_cls; const tfrpn_proposal_cfg* pcfg;
    float* out_boxes; float* out_scores; int32_t* valid; int32_t* keep_idx;
};

// Staging layout of a slot; the device and the pinned block mirror each other.  Inputs are contiguous
// and results are contiguous, so a step whose host buffers ARE the slot's pinned block (acquired mode)
// is one H2D and one D2H copy -- on this link several copies per direction cost 30 % of the duplex
// rate (tools/src/pcie_pattern.cu: 309 us vs 231 us per C2 step).
struct Layout {
    size_t gt, gl, cls, small_end, reg, in_end; // inputs: the small ones first, the head's regression output last
    size_t d, l, ob, os, v, k, pc, dense_end;   // results (pc: rows of rpn_reg the proposal kernels pulled)
    size_t ci, cd, total;                       // compact bbox_deltas (row indices, rows)
};
static Layout make_layout(int B, int N, int G, int P) {
    Layout L;
    size_t o = 0;
    L.gt = o;  o += align256((size_t)B * G * 16);
    L.gl = o;  o += align256((size_t)B * G * 4);
    L.cls = o; o += align256((size_t)B * N * 4);
    L.small_end = o;
    L.reg = o; o += align256((size_t)B * N * 16);
    L.in_end = o;
    L.d = o;   o += align256((size_t)B * N * 16);
    L.l = o;   o += align256((size_t)B * N * 4);
    L.ob = o;  o += align256((size_t)B * P * 16);
    L.os = o;  o += align256((size_t)B * P * 4);
    L.v = o;   o += align256((size_t)B * 4);
    L.k = o;   o += align256((size_t)B * P * 4);
    L.pc = o;  o += 256;
    L.dense_end = o;
    L.ci = o;  o += align256((size_t)B * COMPACT_MAX_POS * 4);
    L.cd = o;  o += align256((size_t)B * COMPACT_MAX_POS * 16);
    L.total = o;
    return L;
}

static int slot_reserve(tfrpn_pipe* p, Slot& s, size_t total) {
    if (total <= s.dev_bytes && total <= s.pin_bytes) return 0;
    // growing frees the old buffers: nothing of this pipeline may still be using them
    for (int i = 0; i < p->depth; ++i) if (int rc = slot_finish(p, p->slots[i])) return rc;
    if (int rc = grow_buffer(&s.dev, &s.dev_bytes, total, p->s_out, false)) return rc;
    if (int rc = grow_buffer(&s.pin, &s.pin_bytes, total, p->s_out, true)) return rc;
    s.dense_clean = false;   // new host block: nothing is known about its contents
    return 0;
}

// `acquired`: the caller's buffers are the slot's own pinned block (tfrpn_pipeline_acquire).
static int pipe_submit(tfrpn_pipe* p, const StepArgs& a, bool acquired, bool order_after, cudaStream_t after,
                       long long* ticket_out) {
    const bool do_t = a.gt_boxes != nullptr, do_p = a.rpn_reg != nullptr;
    if (!do_t && !do_p) return fail(TFRPN_ERR_BAD_ARG, "pipeline_submit: neither half given");
    if (!a.anchors_dev) return fail(TFRPN_ERR_BAD_ARG, "pipeline_submit: anchors_dev is null");
    if (a.B <= 0 || a.N <= 0) return fail(TFRPN_ERR_BAD_ARG, "pipeline_submit: bad shape");
    if (do_t && (!a.gt_labels || !a.tcfg || !a.deltas || !a.labels || a.G <= 0))
        return fail(TFRPN_ERR_BAD_ARG, "pipeline_submit: target half needs gt_labels, cfg, deltas, labels and G >= 1");
    if (do_p && (!a.rpn_cls || !a.pcfg || !a.out_boxes || !a.out_scores || !a.valid || a.pcfg->post_nms_topn <= 0))
        return fail(TFRPN_ERR_BAD_ARG, "pipeline_submit: proposal half needs rpn_cls, cfg, out_boxes, out_scores, valid");
    tfrpn_handle h = p->h;
    TFRPN_ENTER(h);
    TFRPN_CHECK_ON_DEVICE(h, a.anchors_dev, "pipeline_submit: anchors_dev");
    const int B = a.B, N = a.N, G = a.G > 0 ? a.G : 1;
    const int P = acquired ? p->acq_P : (do_p ? a.pcfg->post_nms_topn : 1);
    Slot& s = p->slots[p->next_ticket % p->depth];
    if (!acquired) {
        if (int rc = slot_finish(p, s)) return rc;     // slot still busy with ticket - depth: retire it first
        if (int rc = slot_reserve(p, s, make_layout(B, N, G, P).total)) return rc;
    }
    const Layout L = make_layout(B, N, G, P);
    if (do_t) {
        char* ws;
        if (int rc = ensure_workspace(h, targets_workspace_bytes(B, N, G), p->s_tgt, &ws)) return rc;
    }
    char* d = s.dev;
    char* pin = s.pin;
    auto h2d = [&](size_t off, const void* src, size_t bytes) -> int {
        if (!is_pinned(src)) { memcpy(pin + off, src, bytes); src = pin + off; }
        TFRPN_CHECK_CUDA(cudaMemcpyAsync(d + off, src, bytes, cudaMemcpyHostToDevice, p->s_in));
        return 0;
    };
    auto d2h = [&](void* dst, size_t off, size_t bytes) -> int {
        if (is_pinned(dst)) {
            TFRPN_CHECK_CUDA(cudaMemcpyAsync(dst, d + off, bytes, cudaMemcpyDeviceToHost, p->s_out));
        } else {
            TFRPN_CHECK_CUDA(cudaMemcpyAsync(pin + off, d + off, bytes, cudaMemcpyDeviceToHost, p->s_out));
            s.defer(dst, pin + off, bytes);
        }
        return 0;
    };

    if (order_after) {   // everything of this step is ordered after the caller's prior work on `after`
        TFRPN_CHECK_CUDA(cudaEventRecord(p->ev_after, after));
        for (cudaStream_t st : {p->s_in, p->s_tgt, p->s_prop, p->s_out}) TFRPN_CHECK_CUDA(cudaStreamWaitEvent(st, p->ev_after, 0));
    }
    s.n_copies = 0;
    auto mark = [&](int i, cudaStream_t st) { if (p->trace) cudaEventRecord(s.tr[i], st); };
    if (p->trace) for (int i = 0; i < 8; ++i) cudaEventRecord(s.tr[i], p->s_in);   // halves that do not run read as 0-length
    mark(0, p->s_in);
    // Acquired slots return bbox_deltas in compact form (2.8 MB of results instead of 11.5 MB per C2 step: the
    // deltas are exactly zero outside the <= total_pos sampled positives, utils/train_utils.py:137);
    // slot_finish expands them into the slot's dense host array.
    const bool compact = acquired && do_t && !h->opts.pipe_dense && a.tcfg->total_pos <= COMPACT_MAX_POS;
    // The incremental expansion of slot_finish relies on the slot's dense deltas region still holding zeros plus
    // the previous step's rows.  Any step that is not a compact step of the SAME layout may write into that
    // region (dense results, or the inputs / results of another (B,N,G,P) layout): forget the invariant.
    if (!(compact && s.pB == B && s.pN == N && s.p_off_d == L.d)) s.dense_clean = false;
    // A synchronous step (depth 1) is chunked over images so that copies overlap its own kernels; with
    // several steps in flight the overlap comes from the neighbouring steps and fewer, larger copies win.
    int chunks = (p->depth == 1 && !acquired) ? (B >= 32 ? 4 : (B >= 8 ? 2 : 1)) : 1;
    if (h->opts.pipe_chunks > 0) chunks = h->opts.pipe_chunks;
    chunks = chunks < 1 ? 1 : (chunks > MAX_CHUNKS ? MAX_CHUNKS : chunks);
    // Acquired slots: rpn_reg stays in the slot's page-locked host block and the proposal kernels pull the rows of
    // the candidates they examine (<= ~1000 of N per image) over PCIe with their own loads -- 1 MB of 16-byte rows
    // instead of an 8.9 MB copy at C2.  TFRPN_PIPE_DENSE_IN=1 copies the whole tensor as before (A/B switch).
    const bool pull_reg = acquired && do_p && !h->opts.pipe_dense_in;
    if (acquired) {
        // one copy for all (copied) inputs of the halves that run
        chunks = 1;
        const size_t lo = do_t ? L.gt : L.cls, hi = do_p ? (pull_reg ? L.small_end : L.in_end) : L.cls;
        TFRPN_CHECK_CUDA(cudaMemcpyAsync(d + lo, pin + lo, hi - lo, cudaMemcpyHostToDevice, p->s_in));
        p->last_h2d = (long long)(hi - lo);
        mark(1, p->s_in);
        TFRPN_CHECK_CUDA(cudaEventRecord(s.ev_gt, p->s_in));
        if (do_t) TFRPN_CHECK_CUDA(cudaStreamWaitEvent(p->s_tgt, s.ev_gt, 0));
        if (do_p) TFRPN_CHECK_CUDA(cudaStreamWaitEvent(p->s_prop, s.ev_gt, 0));
        if (do_t) mark(2, p->s_tgt);
        if (do_p) mark(4, p->s_prop);
    } else if (do_t) {
        if (int rc = h2d(L.gt, a.gt_boxes, (size_t)B * G * 16)) return rc;
        if (int rc = h2d(L.gl, a.gt_labels, (size_t)B * G * 4)) return rc;
        TFRPN_CHECK_CUDA(cudaEventRecord(s.ev_gt, p->s_in));
        TFRPN_CHECK_CUDA(cudaStreamWaitEvent(p->s_tgt, s.ev_gt, 0));
    }
    for (int c = 0; c < chunks; ++c) {
        const int lo = (int)((long long)B * c / chunks), hi = (int)((long long)B * (c + 1) / chunks), nb = hi - lo;
        if (nb == 0) continue;
        if (do_p) {
            if (!acquired) {
                if (int rc = h2d(L.reg + (size_t)lo * N * 16, a.rpn_reg + (size_t)lo * N * 4, (size_t)nb * N * 16)) return rc;
                if (int rc = h2d(L.cls + (size_t)lo * N * 4, a.rpn_cls + (size_t)lo * N, (size_t)nb * N * 4)) return rc;
                TFRPN_CHECK_CUDA(cudaEventRecord(s.ev_in[c], p->s_in));
                TFRPN_CHECK_CUDA(cudaStreamWaitEvent(p->s_prop, s.ev_in[c], 0));
            }
            const char* reg_base = pull_reg ? pin : d;   // (pull_reg => one chunk)
            unsigned long long* pc = nullptr;
            if (pull_reg) {
                pc = reinterpret_cast<unsigned long long*>(d + L.pc);
                TFRPN_CHECK_CUDA(cudaMemsetAsync(pc, 0, 8, p->s_prop));
                s.pulled = true; s.off_pc = L.pc;
            }
            if (a.pcfg->pre_nms_topn <= 0) return fail(TFRPN_ERR_BAD_ARG, "pipeline_submit: pre_nms_topn must be > 0");
            if (int rc = proposals_enqueue(h, reinterpret_cast<const float*>(reg_base + L.reg) + (size_t)lo * N * 4,
                                           reinterpret_cast<const float*>(d + L.cls) + (size_t)lo * N, a.anchors_dev, nb, N,
                                           a.pcfg, reinterpret_cast<float*>(d + L.ob) + (size_t)lo * P * 4,
                                           reinterpret_cast<float*>(d + L.os) + (size_t)lo * P,
                                           reinterpret_cast<int32_t*>(d + L.v) + lo,
                                           reinterpret_cast<int32_t*>(d + L.k) + (size_t)lo * P, pc, p->s_prop)) return rc;
        }
        if (do_t) {
            tfrpn_target_cfg cc = *a.tcfg;
            cc.image_offset = a.tcfg->image_offset + lo;
            if (compact) {   // (acquired => one chunk)
                if (int rc = launch_targets(h, a.anchors_dev, reinterpret_cast<const float*>(d + L.gt),
                                            reinterpret_cast<const int32_t*>(d + L.gl), nb, N, G, &cc, nullptr,
                                            reinterpret_cast<float*>(d + L.l), reinterpret_cast<int32_t*>(d + L.ci),
                                            reinterpret_cast<float*>(d + L.cd), nullptr, p->s_tgt)) return rc;
            } else if (int rc = tfrpn_rpn_targets(h, a.anchors_dev, reinterpret_cast<const float*>(d + L.gt) + (size_t)lo * G * 4,
                                           reinterpret_cast<const int32_t*>(d + L.gl) + (size_t)lo * G, nb, N, G, &cc,
                                           reinterpret_cast<float*>(d + L.d) + (size_t)lo * N * 4,
                                           reinterpret_cast<float*>(d + L.l) + (size_t)lo * N, nullptr, p->s_tgt)) return rc;
            mark(3, p->s_tgt);
            TFRPN_CHECK_CUDA(cudaEventRecord(s.ev_tgt[c], p->s_tgt));
            TFRPN_CHECK_CUDA(cudaStreamWaitEvent(p->s_out, s.ev_tgt[c], 0));
            if (!acquired) {
                if (int rc = d2h(a.deltas + (size_t)lo * N * 4, L.d + (size_t)lo * N * 16, (size_t)nb * N * 16)) return rc;
                if (int rc = d2h(a.labels + (size_t)lo * N, L.l + (size_t)lo * N * 4, (size_t)nb * N * 4)) return rc;
            }
        }
    }
    if (do_p) {
        mark(5, p->s_prop);
        TFRPN_CHECK_CUDA(cudaEventRecord(s.ev_prop, p->s_prop));
        TFRPN_CHECK_CUDA(cudaStreamWaitEvent(p->s_out, s.ev_prop, 0));
        if (!acquired) {
            // the four small proposal results come back in ONE D2H copy through pinned staging
            TFRPN_CHECK_CUDA(cudaMemcpyAsync(pin + L.ob, d + L.ob, L.dense_end - L.ob, cudaMemcpyDeviceToHost, p->s_out));
            s.defer(a.out_boxes, pin + L.ob, (size_t)B * P * 16);
            s.defer(a.out_scores, pin + L.os, (size_t)B * P * 4);
            s.defer(a.valid, pin + L.v, (size_t)B * 4);
            if (a.keep_idx) s.defer(a.keep_idx, pin + L.k, (size_t)B * P * 4);
        }
    }
    if (acquired) {   // one copy for all results of the halves that ran
        size_t lo = do_t ? L.d : L.ob, hi = do_p ? L.dense_end : L.ob;
        if (compact) {   // labels, [proposal results,] row indices, rows: one contiguous range
            lo = L.l;
            hi = L.cd + (size_t)B * a.tcfg->total_pos * 16;
            s.compact = true; s.cB = B; s.cN = N; s.cTP = a.tcfg->total_pos;
            s.off_d = L.d; s.off_ci = L.ci; s.off_cd = L.cd;
        }
        mark(6, p->s_out);
        TFRPN_CHECK_CUDA(cudaMemcpyAsync(pin + lo, d + lo, hi - lo, cudaMemcpyDeviceToHost, p->s_out));
        p->last_d2h = (long long)(hi - lo);
    }
    mark(7, p->s_out);
    TFRPN_CHECK_CUDA(cudaEventRecord(s.ev_done, p->s_out));
    s.ticket = p->next_ticket++;
    if (ticket_out) *ticket_out = s.ticket;
    return 0;
}

static int pipe_wait(tfrpn_pipe* p, long long ticket) {
    if (ticket < 0 || ticket >= p->next_ticket) return fail(TFRPN_ERR_BAD_ARG, "pipeline_wait: unknown ticket %lld", ticket);
    Slot& s = p->slots[ticket % p->depth];
    if (s.ticket != ticket) return 0;     // already retired (waited for, or its slot was reused)
    return slot_finish(p, s);
}

}  // namespace tfrpn

using namespace tfrpn;

extern "C" int tfrpn_pipeline_create(tfrpn_handle h, int depth, tfrpn_pipeline* out) { return pipe_create(h, depth, out); }

extern "C" int tfrpn_pipeline_destroy(tfrpn_pipeline p) {
    pipe_destroy(p);
    return 0;
}

extern "C" int tfrpn_pipeline_submit(tfrpn_pipeline p, const float* anchors_dev, int B, int N,
                                     const float* gt_boxes_host, const int32_t* gt_labels_host, int G,
                                     const tfrpn_target_cfg* tcfg, float* deltas_host, float* labels_host,
                                     const float* rpn_reg_host, const float* rpn_cls_host, const tfrpn_proposal_cfg* pcfg,
                                     float* out_boxes_host, float* out_scores_host, int32_t* valid_host,
                                     int32_t* keep_idx_host_or_null, int64_t* ticket_out) {
    if (!p) return fail(TFRPN_ERR_BAD_ARG, "pipeline_submit: null pipeline");
    StepArgs a = {anchors_dev, gt_boxes_host, gt_labels_host, B, N, G, tcfg, deltas_host, labels_host,
                  rpn_reg_host, rpn_cls_host, pcfg, out_boxes_host, out_scores_host, valid_host, keep_idx_host_or_null};
    long long t = -1;
    if (p->acq_live) return fail(TFRPN_ERR_BAD_ARG, "pipeline_submit: a slot is acquired; submit it with tfrpn_pipeline_submit_acquired first");
    if (int rc = pipe_submit(p, a, false, false, nullptr, &t)) return rc;
    if (ticket_out) *ticket_out = t;
    return 0;
}

extern "C" int tfrpn_pipeline_acquire(tfrpn_pipeline p, int B, int N, int G, int post_nms_topn, tfrpn_step_buffers* out) {
    if (!p || !out) return fail(TFRPN_ERR_BAD_ARG, "pipeline_acquire: null pointer");
    if (B <= 0 || N <= 0 || G <= 0 || post_nms_topn <= 0) return fail(TFRPN_ERR_BAD_ARG, "pipeline_acquire: bad shape");
    TFRPN_ENTER(p->h);
    Slot& s = p->slots[p->next_ticket % p->depth];
    if (int rc = slot_finish(p, s)) return rc;
    const Layout L = make_layout(B, N, G, post_nms_topn);
    if (int rc = slot_reserve(p, s, L.total)) return rc;
    p->acq_B = B; p->acq_N = N; p->acq_G = G; p->acq_P = post_nms_topn; p->acq_live = true;
    char* pin = s.pin;
    out->gt_boxes = reinterpret_cast<float*>(pin + L.gt);
    out->gt_labels = reinterpret_cast<int32_t*>(pin + L.gl);
    out->rpn_reg = reinterpret_cast<float*>(pin + L.reg);
    out->rpn_cls = reinterpret_cast<float*>(pin + L.cls);
    out->deltas = reinterpret_cast<float*>(pin + L.d);
    out->labels = reinterpret_cast<float*>(pin + L.l);
    out->out_boxes = reinterpret_cast<float*>(pin + L.ob);
    out->out_scores = reinterpret_cast<float*>(pin + L.os);
    out->valid = reinterpret_cast<int32_t*>(pin + L.v);
    out->keep_idx = reinterpret_cast<int32_t*>(pin + L.k);
    return 0;
}

extern "C" int tfrpn_pipeline_submit_acquired(tfrpn_pipeline p, const float* anchors_dev, const tfrpn_target_cfg* tcfg_or_null,
                                              const tfrpn_proposal_cfg* pcfg_or_null, int64_t* ticket_out) {
    if (!p) return fail(TFRPN_ERR_BAD_ARG, "pipeline_submit_acquired: null pipeline");
    if (!p->acq_live) return fail(TFRPN_ERR_BAD_ARG, "pipeline_submit_acquired: no slot acquired");
    if (pcfg_or_null && pcfg_or_null->post_nms_topn != p->acq_P)
        return fail(TFRPN_ERR_BAD_ARG, "pipeline_submit_acquired: post_nms_topn %d differs from the acquired %d",
                    pcfg_or_null->post_nms_topn, p->acq_P);
    Slot& s = p->slots[p->next_ticket % p->depth];
    const Layout L = make_layout(p->acq_B, p->acq_N, p->acq_G, p->acq_P);
    char* pin = s.pin;
    StepArgs a = {};
    a.anchors_dev = anchors_dev; a.B = p->acq_B; a.N = p->acq_N; a.G = p->acq_G;
    if (tcfg_or_null) {
        a.gt_boxes = reinterpret_cast<float*>(pin + L.gt); a.gt_labels = reinterpret_cast<int32_t*>(pin + L.gl);
        a.tcfg = tcfg_or_null; a.deltas = reinterpret_cast<float*>(pin + L.d); a.labels = reinterpret_cast<float*>(pin + L.l);
    }
    if (pcfg_or_null) {
        a.rpn_reg = reinterpret_cast<float*>(pin + L.reg); a.rpn_cls = reinterpret_cast<float*>(pin + L.cls);
        a.pcfg = pcfg_or_null; a.out_boxes = reinterpret_cast<float*>(pin + L.ob);
        a.out_scores = reinterpret_cast<float*>(pin + L.os); a.valid = reinterpret_cast<int32_t*>(pin + L.v);
        a.keep_idx = reinterpret_cast<int32_t*>(pin + L.k);
    }
    long long t = -1;
    if (int rc = pipe_submit(p, a, true, false, nullptr, &t)) return rc;
    p->acq_live = false;
    if (ticket_out) *ticket_out = t;
    return 0;
}

extern "C" int tfrpn_pipeline_last_copy_bytes(tfrpn_pipeline p, int64_t* h2d_bytes, int64_t* d2h_bytes) {
    if (!p || !h2d_bytes || !d2h_bytes) return fail(TFRPN_ERR_BAD_ARG, "pipeline_last_copy_bytes: null pointer");
    *h2d_bytes = p->last_h2d + p->last_pulled;   // copy + the 16-byte rows the kernels loaded from the pinned block
    *d2h_bytes = p->last_d2h;
    return 0;
}

extern "C" int tfrpn_pipeline_trace(tfrpn_pipeline p, int64_t ticket, float* ms8) {
    if (!p || !ms8) return fail(TFRPN_ERR_BAD_ARG, "pipeline_trace: null pointer");
    if (!p->trace) return fail(TFRPN_ERR_UNSUPPORTED, "pipeline_trace: the handle was created without TFRPN_PIPE_TRACE=1");
    const Slot& s = p->slots[(ticket < 0 ? 0 : ticket) % p->depth];
    if (s.tr_ticket != ticket) return fail(TFRPN_ERR_BAD_ARG, "pipeline_trace: step %lld is not the last one retired from its slot", (long long)ticket);
    for (int i = 0; i < 8; ++i) ms8[i] = s.tr_ms[i];
    return 0;
}

extern "C" int tfrpn_pipeline_wait(tfrpn_pipeline p, int64_t ticket) {
    if (!p) return fail(TFRPN_ERR_BAD_ARG, "pipeline_wait: null pipeline");
    TFRPN_ENTER(p->h);
    return pipe_wait(p, ticket);
}

extern "C" int tfrpn_pipeline_drain(tfrpn_pipeline p) {
    if (!p) return fail(TFRPN_ERR_BAD_ARG, "pipeline_drain: null pipeline");
    TFRPN_ENTER(p->h);
    for (int i = 0; i < p->depth; ++i) if (int rc = slot_finish(p, p->slots[i])) return rc;
    return 0;
}

// One step from host buffers, synchronous: submit + wait on a depth-1 pipeline owned by the handle.
extern "C" int tfrpn_rpn_step_host(tfrpn_handle h, const float* anchors_dev, const float* gt_boxes_host,
                                   const int32_t* gt_labels_host, int B, int N, int G,
                                   const tfrpn_target_cfg* tcfg, float* deltas_host, float* labels_host,
                                   const float* rpn_reg_host, const float* rpn_cls_host,
                                   const tfrpn_proposal_cfg* pcfg, float* out_boxes_host, float* out_scores_host,
                                   int32_t* valid_host, int32_t* keep_idx_host_or_null, tfrpn_stream s) {
    if (!h) return fail(TFRPN_ERR_BAD_ARG, "rpn_step_host: null handle");
    if (!gt_boxes_host || !gt_labels_host || !deltas_host || !labels_host || !tcfg || !rpn_reg_host || !rpn_cls_host ||
        !pcfg || !out_boxes_host || !out_scores_host || !valid_host)
        return fail(TFRPN_ERR_BAD_ARG, "rpn_step_host: null pointer");
    if (B <= 0 || N <= 0 || G <= 0 || pcfg->post_nms_topn <= 0) return fail(TFRPN_ERR_BAD_ARG, "rpn_step_host: bad shape");
    TFRPN_ENTER(h);
    if (!h->step_pipe) if (int rc = pipe_create(h, 1, &h->step_pipe)) return rc;
    StepArgs a = {anchors_dev, gt_boxes_host, gt_labels_host, B, N, G, tcfg, deltas_host, labels_host,
                  rpn_reg_host, rpn_cls_host, pcfg, out_boxes_host, out_scores_host, valid_host, keep_idx_host_or_null};
    long long t = -1;
    if (int rc = pipe_submit(h->step_pipe, a, false, true, as_stream(s), &t)) return rc;
    return pipe_wait(h->step_pipe, t);
}
