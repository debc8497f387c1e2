// losses.cu -- the RPN losses that consume the target-assignment outputs:
//   cls_loss  utils/train_utils.py:146-161  BinaryCrossentropy over the entries with label != -1
//   reg_loss  utils/train_utils.py:163-185  Huber (delta 1) summed over the 4 coordinates of the
//             rows whose true delta is not all-zero, divided by max(1, #such rows)
// and, optionally, their gradients with respect to the predictions (what TF's autograd derives
// from the same ops), so that the pair can back a tf.custom_gradient.
//
//   L1 rpn_loss_partial_kernel  one pass over (labels, true deltas); the predictions are only
//                               read where a term exists (~3 % of the scores, ~1.5 % of the
//                               regression rows), so the kernel is HBM-bound on 20*B*N bytes.
//                               Per-element terms in float32 in the reference's op order;
//                               sums to float64 accuracy (float32 two-sum pairs inside a
//                               warp, doubles above), one partial per CTA; the last CTA
//                               sums the partials in a fixed order (deterministic) and divides.
//   L2 rpn_loss_grad_kernel     elementwise gradients, scaled by the counts L1 left on the device.
#include "common.cuh"

namespace tfrpn {

constexpr int LOSS_THREADS = 256;

// [TF-internal] Keras backend.binary_crossentropy(from_logits=False), TF 2.0.0:
//   p = clip_by_value(p, eps, 1 - eps); bce = t*log(p + eps); bce += (1 - t)*log(1 - p + eps); -bce
__device__ __forceinline__ float bce_term(float t, float p) {
    const float eps = 1e-7f;
    const float hi = __fsub_rn(1.0f, eps);
    p = fminf(fmaxf(p, eps), hi);
    float bce = __fmul_rn(t, logf(__fadd_rn(p, eps)));
    bce = __fadd_rn(bce, __fmul_rn(__fsub_rn(1.0f, t), logf(__fadd_rn(__fsub_rn(1.0f, p), eps))));
    return -bce;
}
// d bce / d p (clip_by_value passes the gradient where eps <= p <= 1 - eps)
__device__ __forceinline__ float bce_grad(float t, float p) {
    const float eps = 1e-7f;
    const float hi = __fsub_rn(1.0f, eps);
    if (!(p >= eps && p <= hi)) return 0.0f;
    const float a = __fdiv_rn(t, __fadd_rn(p, eps));
    const float b = __fdiv_rn(__fsub_rn(1.0f, t), __fadd_rn(__fsub_rn(1.0f, p), eps));
    return __fsub_rn(b, a);
}
// [TF-internal] keras huber_loss, TF 2.0.0 (elementwise; no mean over the last axis before 2.1):
//   e = pred - true; q = min(|e|, delta); lin = |e| - q; 0.5*q*q + delta*lin
__device__ __forceinline__ float huber_term(float t, float p, float delta) {
    const float ae = fabsf(__fsub_rn(p, t));
    const float q = fminf(ae, delta);
    const float lin = __fsub_rn(ae, q);
    return __fadd_rn(__fmul_rn(0.5f, __fmul_rn(q, q)), __fmul_rn(delta, lin));
}
__device__ __forceinline__ float huber_grad(float t, float p, float delta) {
    const float e = __fsub_rn(p, t);
    const float ae = fabsf(e);
    if (ae <= delta) return e;                       // minimum(x, y) routes the gradient to x when x <= y
    return e > 0.0f ? delta : -delta;
}

// per-CTA partial: the two sums as (hi, lo) float pairs and the two counts
struct LossPartial {
    float4 sums;   // reg.hi, reg.lo, cls.hi, cls.lo
    uint2 counts;  // n_pos, n_cls
    uint2 pad;
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// Non-tensor FP64 is slow on this part, so the hot path stays in float32: a thread adds its <= LOSS_EPT terms
// in float32, and the warp combines them as (hi, lo) pairs with the error-free two-sum, i.e. to float64-like
// accuracy without FP64 instructions.  Only one thread per CTA (and the last CTA) touches doubles.
__device__ __forceinline__ float2 pair_add(float2 a, float2 b) {
    const float s = __fadd_rn(a.x, b.x);
    const float bb = __fsub_rn(s, a.x);
    const float err = __fadd_rn(__fsub_rn(a.x, __fsub_rn(s, bb)), __fsub_rn(b.x, bb));
    return make_float2(s, __fadd_rn(err, __fadd_rn(a.y, b.y)));
}
__device__ __forceinline__ float2 warp_pair_sum(float v) {
    float2 p = make_float2(v, 0.0f);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        float2 q;
        q.x = __shfl_xor_sync(0xffffffffu, p.x, o);
        q.y = __shfl_xor_sync(0xffffffffu, p.y, o);
        p = pair_add(p, q);
    }
    return p;
}
__device__ __forceinline__ unsigned warp_sum(unsigned v) { return __reduce_add_sync(0xffffffffu, v); }

// One element (one anchor of one image) per thread and LOSS_EPT independent elements per thread, all of
// whose unconditional loads are issued before anything is consumed.  The last CTA to finish (a ticket in
// handle-owned memory, reset for the next call) sums the per-CTA partials in a fixed order, so the result
// does not depend on the order in which the CTAs ran.
constexpr int LOSS_EPT = 4;

__global__ void __launch_bounds__(LOSS_THREADS) rpn_loss_partial_kernel(
    const float4* __restrict__ true_deltas, const float4* __restrict__ pred_deltas,
    const float* __restrict__ true_labels, const float* __restrict__ pred_scores, long long total, float delta,
    LossPartial* __restrict__ partials, unsigned int* __restrict__ ticket, tfrpn_loss_out* __restrict__ out) {
    float reg = 0.0f, cls = 0.0f;
    unsigned npos = 0u, ncls = 0u;
    const long long base = (long long)blockIdx.x * (LOSS_THREADS * LOSS_EPT) + threadIdx.x;
    float tl[LOSS_EPT];
    float4 td[LOSS_EPT];
#pragma unroll
    for (int u = 0; u < LOSS_EPT; ++u) {
        const long long i = base + (long long)u * LOSS_THREADS;
        tl[u] = -1.0f;
        td[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < total) {
            if (true_labels) tl[u] = __ldg(true_labels + i);
            if (true_deltas) td[u] = ldg_f4_stream(true_deltas + i);
        }
    }
    // second wave of loads: the predictions, only where a term exists -- all issued before any is used
    // (a load inside each divergent branch would serialise LOSS_EPT memory latencies per warp)
    float ps[LOSS_EPT];
    float4 pp[LOSS_EPT];
    bool pos[LOSS_EPT];
#pragma unroll
    for (int u = 0; u < LOSS_EPT; ++u) {
        const long long i = base + (long long)u * LOSS_THREADS;
        const float4 t = td[u];
        pos[u] = t.x != 0.0f || t.y != 0.0f || t.z != 0.0f || t.w != 0.0f;      // train_utils.py:180
        ps[u] = 0.5f;
        pp[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (tl[u] != -1.0f) ps[u] = __ldg(pred_scores + i);                     // train_utils.py:156
        if (pos[u]) pp[u] = ldg_f4(pred_deltas + i);
    }
#pragma unroll
    for (int u = 0; u < LOSS_EPT; ++u) {
        if (tl[u] != -1.0f) {
            cls = __fadd_rn(cls, bce_term(tl[u], ps[u]));
            ++ncls;
        }
        if (pos[u]) {
            const float4 t = td[u], p = pp[u];
            float row = huber_term(t.x, p.x, delta);                      // reduce_sum(axis=-1), :178
            row = __fadd_rn(row, huber_term(t.y, p.y, delta));
            row = __fadd_rn(row, huber_term(t.z, p.z, delta));
            row = __fadd_rn(row, huber_term(t.w, p.w, delta));
            reg = __fadd_rn(reg, row);
            ++npos;
        }
    }
    __shared__ float2 s_reg[LOSS_THREADS / 32], s_cls[LOSS_THREADS / 32];
    __shared__ unsigned s_np[LOSS_THREADS / 32], s_nc[LOSS_THREADS / 32];
    __shared__ bool s_last;
    const float2 reg2 = warp_pair_sum(reg);
    const float2 cls2 = warp_pair_sum(cls);
    npos = warp_sum(npos);
    ncls = warp_sum(ncls);
    if (lane_id() == 0) {
        s_reg[warp_id()] = reg2; s_cls[warp_id()] = cls2; s_np[warp_id()] = npos; s_nc[warp_id()] = ncls;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float2 r = s_reg[0], c = s_cls[0];
        unsigned np = s_np[0], nc = s_nc[0];
        for (int w = 1; w < LOSS_THREADS / 32; ++w) {
            r = pair_add(r, s_reg[w]); c = pair_add(c, s_cls[w]); np += s_np[w]; nc += s_nc[w];
        }
        LossPartial p;
        p.sums = make_float4(r.x, r.y, c.x, c.y);
        p.counts = make_uint2(np, nc);
        p.pad = make_uint2(0u, 0u);
        partials[blockIdx.x] = p;
        __threadfence();
        s_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!s_last) return;
    // the last CTA: fixed-order sum of all partials (thread t takes t, t+256, ...; then fixed trees)
    __threadfence();
    float2 r2 = make_float2(0.f, 0.f), c2 = make_float2(0.f, 0.f);
    unsigned long long np2 = 0ull, nc2 = 0ull;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += LOSS_THREADS) {
        const float4 q = __ldcg(&partials[i].sums);
        const uint2 k = __ldcg(&partials[i].counts);
        r2 = pair_add(r2, make_float2(q.x, q.y));
        c2 = pair_add(c2, make_float2(q.z, q.w));
        np2 += k.x; nc2 += k.y;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        float2 q;
        q.x = __shfl_xor_sync(0xffffffffu, r2.x, o); q.y = __shfl_xor_sync(0xffffffffu, r2.y, o);
        r2 = pair_add(r2, q);
        q.x = __shfl_xor_sync(0xffffffffu, c2.x, o); q.y = __shfl_xor_sync(0xffffffffu, c2.y, o);
        c2 = pair_add(c2, q);
        np2 += __shfl_xor_sync(0xffffffffu, np2, o);
        nc2 += __shfl_xor_sync(0xffffffffu, nc2, o);
    }
    __shared__ float2 t_reg[LOSS_THREADS / 32], t_cls[LOSS_THREADS / 32];
    __shared__ unsigned long long t_np[LOSS_THREADS / 32], t_nc[LOSS_THREADS / 32];
    if (lane_id() == 0) { t_reg[warp_id()] = r2; t_cls[warp_id()] = c2; t_np[warp_id()] = np2; t_nc[warp_id()] = nc2; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double R = 0.0, Cs = 0.0;
        unsigned long long NP = 0ull, NC = 0ull;
        for (int w = 0; w < LOSS_THREADS / 32; ++w) {
            R += (double)t_reg[w].x + (double)t_reg[w].y;
            Cs += (double)t_cls[w].x + (double)t_cls[w].y;
            NP += t_np[w]; NC += t_nc[w];
        }
        tfrpn_loss_out r;
        // :183-185  loc_loss / max(1, #pos);   :159-161  mean over the gathered entries (0/0 -> NaN, as TF)
        r.reg_loss = true_deltas ? (float)(R / (double)(NP > 0ull ? NP : 1ull)) : 0.0f;
        r.cls_loss = true_labels ? (float)(Cs / (double)NC) : 0.0f;
        r.n_pos = (int32_t)NP;
        r.n_cls = (int32_t)NC;
        *out = r;
        *ticket = 0u;     // ready for the next call on this handle
    }
}

constexpr int GRAD_EPT = 4;
__global__ void __launch_bounds__(LOSS_THREADS) rpn_loss_grad_kernel(
    const float4* __restrict__ true_deltas, const float4* __restrict__ pred_deltas,
    const float* __restrict__ true_labels, const float* __restrict__ pred_scores, long long total, float delta,
    const tfrpn_loss_out* __restrict__ res, float4* __restrict__ grad_deltas, float* __restrict__ grad_scores) {
    const long long base = (long long)blockIdx.x * (LOSS_THREADS * GRAD_EPT) + threadIdx.x;
    float tl[GRAD_EPT];
    float4 td[GRAD_EPT];
#pragma unroll
    for (int u = 0; u < GRAD_EPT; ++u) {
        const long long i = base + (long long)u * LOSS_THREADS;
        tl[u] = -1.0f;
        td[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < total) {
            if (grad_scores) tl[u] = __ldg(true_labels + i);
            if (grad_deltas) td[u] = ldg_f4_stream(true_deltas + i);
        }
    }
    float ps[GRAD_EPT];
    float4 pp[GRAD_EPT];
    bool pos[GRAD_EPT];
#pragma unroll
    for (int u = 0; u < GRAD_EPT; ++u) {   // the predictions, only where a term exists, all in flight together
        const long long i = base + (long long)u * LOSS_THREADS;
        const float4 t = td[u];
        pos[u] = t.x != 0.0f || t.y != 0.0f || t.z != 0.0f || t.w != 0.0f;
        ps[u] = 0.5f;
        pp[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (tl[u] != -1.0f) ps[u] = __ldg(pred_scores + i);
        if (pos[u]) pp[u] = ldg_f4(pred_deltas + i);
    }
    const float inv_pos = __fdiv_rn(1.0f, (float)max(res->n_pos, 1));
    const float inv_cls = __fdiv_rn(1.0f, (float)res->n_cls);
#pragma unroll
    for (int u = 0; u < GRAD_EPT; ++u) {
        const long long i = base + (long long)u * LOSS_THREADS;
        if (i >= total) continue;
        if (grad_scores) {
            float g = 0.0f;
            if (tl[u] != -1.0f) g = __fmul_rn(bce_grad(tl[u], ps[u]), inv_cls);
            stg_f1_stream(grad_scores + i, g);
        }
        if (grad_deltas) {
            const float4 t = td[u], p = pp[u];
            float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
            if (pos[u])
                g = make_float4(__fmul_rn(huber_grad(t.x, p.x, delta), inv_pos), __fmul_rn(huber_grad(t.y, p.y, delta), inv_pos),
                                __fmul_rn(huber_grad(t.z, p.z, delta), inv_pos), __fmul_rn(huber_grad(t.w, p.w, delta), inv_pos));
            stg_f4_stream(grad_deltas + i, g);
        }
    }
}

size_t losses_workspace_bytes(long long total) {
    return (size_t)((total + LOSS_THREADS * LOSS_EPT - 1) / (LOSS_THREADS * LOSS_EPT) + 1) * sizeof(LossPartial) + 256;
}

}  // namespace tfrpn

using namespace tfrpn;

extern "C" int tfrpn_rpn_losses(tfrpn_handle h, const float* true_deltas, const float* pred_deltas,
                                const float* true_labels, const float* pred_scores, int B, int N, float huber_delta,
                                tfrpn_loss_out* out, float* grad_deltas_or_null, float* grad_scores_or_null,
                                tfrpn_stream s) {
    if (!h) return fail(TFRPN_ERR_BAD_ARG, "rpn_losses: null handle");
    if (!out) return fail(TFRPN_ERR_BAD_ARG, "rpn_losses: out is null");
    if ((true_deltas != nullptr) != (pred_deltas != nullptr) || (true_labels != nullptr) != (pred_scores != nullptr))
        return fail(TFRPN_ERR_BAD_ARG, "rpn_losses: true / predicted tensors must be given in pairs");
    if (!true_deltas && !true_labels) return fail(TFRPN_ERR_BAD_ARG, "rpn_losses: nothing to do");
    if ((grad_deltas_or_null && !true_deltas) || (grad_scores_or_null && !true_labels))
        return fail(TFRPN_ERR_BAD_ARG, "rpn_losses: gradient requested for a loss that is not computed");
    if (B < 0 || N < 0) return fail(TFRPN_ERR_BAD_ARG, "rpn_losses: negative shape");
    if (!(huber_delta > 0.0f)) return fail(TFRPN_ERR_BAD_ARG, "rpn_losses: huber_delta must be > 0");
    if (true_deltas && (!aligned16(true_deltas) || !aligned16(pred_deltas) || (grad_deltas_or_null && !aligned16(grad_deltas_or_null))))
        return fail(TFRPN_ERR_MISALIGNED, "rpn_losses: delta tensors must be 16-byte aligned");
    TFRPN_ENTER(h);
    TFRPN_CHECK_ON_DEVICE(h, out, "rpn_losses: out");
    cudaStream_t st = as_stream(s);
    const long long total = (long long)B * N;
    if (total > (1LL << 38)) return fail(TFRPN_ERR_UNSUPPORTED, "rpn_losses: tensor too large");
    char* ws = nullptr;
    if (int rc = ensure_workspace(h, losses_workspace_bytes(total), st, &ws)) return rc;
    LossPartial* partials = reinterpret_cast<LossPartial*>(ws);
    if (!h->ticket) return fail(TFRPN_ERR_CUDA, "rpn_losses: the handle has no device counter");
    const long long want = (total + LOSS_THREADS * LOSS_EPT - 1) / (LOSS_THREADS * LOSS_EPT);
    const int ctas = (int)(want < 1 ? 1 : want);
    const float4* td = reinterpret_cast<const float4*>(true_deltas);
    const float4* pd = reinterpret_cast<const float4*>(pred_deltas);
    prof_begin(h, TFRPN_K_LOSS, st);
    rpn_loss_partial_kernel<<<ctas, LOSS_THREADS, 0, st>>>(td, pd, true_labels, pred_scores, total, huber_delta, partials,
                                                           h->ticket, out);
    prof_end(h, st);
    TFRPN_AFTER_LAUNCH("rpn_loss_partial_kernel");
    if ((grad_deltas_or_null || grad_scores_or_null) && total > 0) {
        const int gctas = (int)((total + LOSS_THREADS * GRAD_EPT - 1) / (LOSS_THREADS * GRAD_EPT));
        rpn_loss_grad_kernel<<<gctas, LOSS_THREADS, 0, st>>>(td, pd, true_labels, pred_scores, total, huber_delta, out,
                                                            reinterpret_cast<float4*>(grad_deltas_or_null),
                                                            grad_scores_or_null);
        TFRPN_AFTER_LAUNCH("rpn_loss_grad_kernel");
    }
    return 0;
}
