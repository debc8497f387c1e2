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
//                               sums in float64, one partial per CTA.
//   L2 rpn_loss_final_kernel    fixed-order sum of the partials (deterministic), the divisions.
//   L3 rpn_loss_grad_kernel     elementwise gradients, scaled by the counts L2 left on the device.
#include "common.cuh"

namespace tfrpn {

constexpr int LOSS_THREADS = 256;
constexpr int LOSS_MAX_CTAS = 148 * 8;

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

struct LossPartial {
    double reg_sum, cls_sum;
    unsigned long long n_pos, n_cls;
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ unsigned warp_sum(unsigned v) { return __reduce_add_sync(0xffffffffu, v); }

__global__ void __launch_bounds__(LOSS_THREADS) rpn_loss_partial_kernel(
    const float4* __restrict__ true_deltas, const float4* __restrict__ pred_deltas,
    const float* __restrict__ true_labels, const float* __restrict__ pred_scores, long long total, float delta,
    LossPartial* __restrict__ partials) {
    double reg = 0.0, cls = 0.0;
    unsigned npos = 0u, ncls = 0u;
    const long long stride = (long long)gridDim.x * LOSS_THREADS;
    for (long long i = (long long)blockIdx.x * LOSS_THREADS + threadIdx.x; i < total; i += stride) {
        if (true_labels) {
            const float t = __ldg(true_labels + i);
            if (t != -1.0f) {                                        // train_utils.py:156
                cls += (double)bce_term(t, __ldg(pred_scores + i));
                ++ncls;
            }
        }
        if (true_deltas) {
            const float4 t = ldg_f4_stream(true_deltas + i);
            if (t.x != 0.0f || t.y != 0.0f || t.z != 0.0f || t.w != 0.0f) {   // train_utils.py:180
                const float4 p = ldg_f4(pred_deltas + i);
                float row = huber_term(t.x, p.x, delta);                      // reduce_sum(axis=-1), :178
                row = __fadd_rn(row, huber_term(t.y, p.y, delta));
                row = __fadd_rn(row, huber_term(t.z, p.z, delta));
                row = __fadd_rn(row, huber_term(t.w, p.w, delta));
                reg += (double)row;
                ++npos;
            }
        }
    }
    __shared__ double s_reg[LOSS_THREADS / 32], s_cls[LOSS_THREADS / 32];
    __shared__ unsigned s_np[LOSS_THREADS / 32], s_nc[LOSS_THREADS / 32];
    reg = warp_sum(reg);
    cls = warp_sum(cls);
    npos = warp_sum(npos);
    ncls = warp_sum(ncls);
    if (lane_id() == 0) {
        s_reg[warp_id()] = reg; s_cls[warp_id()] = cls; s_np[warp_id()] = npos; s_nc[warp_id()] = ncls;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        LossPartial p = {0.0, 0.0, 0ull, 0ull};
        for (int w = 0; w < LOSS_THREADS / 32; ++w) {
            p.reg_sum += s_reg[w]; p.cls_sum += s_cls[w]; p.n_pos += s_np[w]; p.n_cls += s_nc[w];
        }
        partials[blockIdx.x] = p;
    }
}

__global__ void __launch_bounds__(32) rpn_loss_final_kernel(const LossPartial* __restrict__ partials, int n,
                                                            int have_reg, int have_cls, tfrpn_loss_out* __restrict__ out) {
    // one warp, fixed order: lane l sums partials l, l+32, ... then a shuffle tree
    double reg = 0.0, cls = 0.0;
    unsigned long long npos = 0ull, ncls = 0ull;
    for (int i = lane_id(); i < n; i += 32) {
        const LossPartial p = partials[i];
        reg += p.reg_sum; cls += p.cls_sum; npos += p.n_pos; ncls += p.n_cls;
    }
    reg = warp_sum(reg);
    cls = warp_sum(cls);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        npos += __shfl_xor_sync(0xffffffffu, npos, o);
        ncls += __shfl_xor_sync(0xffffffffu, ncls, o);
    }
    if (lane_id() == 0) {
        tfrpn_loss_out r;
        // :183-185  loc_loss / max(1, #pos);   :159-161  mean over the gathered entries (0/0 -> NaN, as TF)
        r.reg_loss = have_reg ? (float)(reg / (double)(npos > 0ull ? npos : 1ull)) : 0.0f;
        r.cls_loss = have_cls ? (float)(cls / (double)ncls) : 0.0f;
        r.n_pos = (int32_t)npos;
        r.n_cls = (int32_t)ncls;
        *out = r;
    }
}

__global__ void __launch_bounds__(LOSS_THREADS) rpn_loss_grad_kernel(
    const float4* __restrict__ true_deltas, const float4* __restrict__ pred_deltas,
    const float* __restrict__ true_labels, const float* __restrict__ pred_scores, long long total, float delta,
    const tfrpn_loss_out* __restrict__ res, float4* __restrict__ grad_deltas, float* __restrict__ grad_scores) {
    const float inv_pos = __fdiv_rn(1.0f, (float)max(res->n_pos, 1));
    const float inv_cls = __fdiv_rn(1.0f, (float)res->n_cls);
    const long long stride = (long long)gridDim.x * LOSS_THREADS;
    for (long long i = (long long)blockIdx.x * LOSS_THREADS + threadIdx.x; i < total; i += stride) {
        if (grad_scores) {
            const float t = __ldg(true_labels + i);
            float g = 0.0f;
            if (t != -1.0f) g = __fmul_rn(bce_grad(t, __ldg(pred_scores + i)), inv_cls);
            stg_f1_stream(grad_scores + i, g);
        }
        if (grad_deltas) {
            const float4 t = ldg_f4_stream(true_deltas + i);
            float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
            if (t.x != 0.0f || t.y != 0.0f || t.z != 0.0f || t.w != 0.0f) {
                const float4 p = ldg_f4(pred_deltas + i);
                g = make_float4(__fmul_rn(huber_grad(t.x, p.x, delta), inv_pos), __fmul_rn(huber_grad(t.y, p.y, delta), inv_pos),
                                __fmul_rn(huber_grad(t.z, p.z, delta), inv_pos), __fmul_rn(huber_grad(t.w, p.w, delta), inv_pos));
            }
            stg_f4_stream(grad_deltas + i, g);
        }
    }
}

size_t losses_workspace_bytes() { return (size_t)LOSS_MAX_CTAS * sizeof(LossPartial) + 256; }

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
    cudaStream_t st = as_stream(s);
    const long long total = (long long)B * N;
    char* ws = nullptr;
    if (int rc = ensure_workspace(h, losses_workspace_bytes(), st, &ws)) return rc;
    LossPartial* partials = reinterpret_cast<LossPartial*>(ws);
    long long want = (total + LOSS_THREADS * 4 - 1) / (LOSS_THREADS * 4);   // ~4 elements per thread
    const int ctas = (int)(want < 1 ? 1 : (want > LOSS_MAX_CTAS ? LOSS_MAX_CTAS : want));
    const float4* td = reinterpret_cast<const float4*>(true_deltas);
    const float4* pd = reinterpret_cast<const float4*>(pred_deltas);
    prof_begin(h, TFRPN_K_LOSS, st);
    rpn_loss_partial_kernel<<<ctas, LOSS_THREADS, 0, st>>>(td, pd, true_labels, pred_scores, total, huber_delta, partials);
    prof_end(h, st);
    TFRPN_AFTER_LAUNCH("rpn_loss_partial_kernel");
    rpn_loss_final_kernel<<<1, 32, 0, st>>>(partials, ctas, true_deltas != nullptr, true_labels != nullptr, out);
    TFRPN_AFTER_LAUNCH("rpn_loss_final_kernel");
    if ((grad_deltas_or_null || grad_scores_or_null) && total > 0) {
        rpn_loss_grad_kernel<<<ctas, LOSS_THREADS, 0, st>>>(td, pd, true_labels, pred_scores, total, huber_delta, out,
                                                            reinterpret_cast<float4*>(grad_deltas_or_null),
                                                            grad_scores_or_null);
        TFRPN_AFTER_LAUNCH("rpn_loss_grad_kernel");
    }
    return 0;
}
