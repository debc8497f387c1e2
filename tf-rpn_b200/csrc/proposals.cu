// proposals.cu -- pre-NMS top-k, NMS and the fused proposal stage, one CTA per image.
//
// Replaces tf.nn.top_k + tf.gather (predictor.py:58-60), non_max_suppression ->
// tf.image.combined_non_max_suppression (utils/bbox_utils.py:48-70) and their composition
// (SURVEY.md 8a row P).  Phases inside the CTA (all in shared memory):
//   1 radix SELECT of the k-th largest score key (MSB-first, 8 bits/pass, early exit)
//   2 ORDERED compaction of the selected entries (ascending index, so equal scores stay in
//     index order: [TF-internal] top_k returns the lower index first)
//   3 stable LSD radix SORT by score descending (4 x 8-bit passes, match.any ranking)
//   4 top-k outputs, or chunked greedy NMS in score order: 128 candidates per round are tested
//     against the kept list, then against each other through 128-bit suppression masks that one
//     thread sweeps serially; stops at max_output_size like TF's loop does.
//   In the fused mode boxes are decoded (+clipped) on the fly for the candidates only.
#include <math.h>
#include <string.h>

#include "common.cuh"

namespace tfrpn {

constexpr int PR_THREADS = 1024;
constexpr int PR_WARPS = PR_THREADS / 32;
constexpr int NMS_CHUNK = 128;
constexpr int NMS_PARTS = PR_THREADS / NMS_CHUNK;  // 8

enum { MODE_TOPK = 0, MODE_NMS = 1, MODE_PROPOSALS = 2 };

struct PropParams {
    int mode;
    int N;     // entries per image
    int k;     // requested top-k (<= N)
    int kcap;  // shared-memory sort capacity (entries)
    int staged;
    const float* scores;  // (B,N)
    int use_sthr;
    float score_threshold;
    const float4* boxes;   // MODE_TOPK gather source / MODE_NMS input
    long long box_stride;  // elements between images (0 = shared (N,4))
    const float4* reg;     // MODE_PROPOSALS: (B,N,4) head regression output
    const float4* anchors; // MODE_PROPOSALS: (N,4)
    float4 var;
    int clip_decoded;
    float* values;   // MODE_TOPK (B,k)
    int* indices;    // MODE_TOPK (B,k)
    float4* gathered;
    int rows, max_out;
    int mo_pad;  // max_out rounded up to a multiple of 4 (keeps the carve 16-byte aligned)
    IouThreshold iou_thr;
    int clip_out;
    float4* out_boxes;
    float* out_scores;
    float* out_classes;
    int* valid;
    int* keep_idx;
};

struct PropShared {
    unsigned int wtot[PR_WARPS];
    unsigned int digit, remaining, bin_count;
    unsigned int count;
    int nk;
};

// sort key: descending score == ascending ~orderable(score)
__device__ __forceinline__ uint32_t score_key(float s, int use_sthr, float sthr) {
    if (use_sthr && !(s > sthr)) return 0u;
    return orderable(s);
}

// exclusive block scan of one unsigned per thread (1024 threads); returns exclusive prefix, total in *total
__device__ __forceinline__ unsigned int block_excl_scan(unsigned int v, PropShared* sh, unsigned int* total) {
    unsigned int incl = (unsigned int)warp_incl_scan((int)v);
    if (lane_id() == 31) sh->wtot[warp_id()] = incl;
    __syncthreads();
    if (warp_id() == 0) {
        unsigned int w = sh->wtot[lane_id()];
        unsigned int wi = (unsigned int)warp_incl_scan((int)w);
        sh->wtot[lane_id()] = wi - w;
        if (lane_id() == 31) sh->count = wi;
    }
    __syncthreads();
    unsigned int r = sh->wtot[warp_id()] + incl - v;
    *total = sh->count;
    __syncthreads();
    return r;
}

// TF CombinedNonMaxSuppression IOU on canonicalised boxes ([TF-internal]); c = (ymin,xmin,ymax,xmax).
// Returns IOU(i,j) > thr exactly as TF evaluates it: 0 if either area <= 0; otherwise
// RN(inter / ((ai + aj) - inter)) > thr, decided without the division (common.cuh: iou_exceeds).
__device__ __forceinline__ bool nms_suppresses(float4 ci, float ai, float4 cj, float aj, const IouThreshold& t) {
    if (ai > 0.0f && aj > 0.0f) {
        float iymin = fmaxf(ci.x, cj.x), ixmin = fmaxf(ci.y, cj.y);
        float iymax = fminf(ci.z, cj.z), ixmax = fminf(ci.w, cj.w);
        float inter = __fmul_rn(fmaxf(__fsub_rn(iymax, iymin), 0.0f), fmaxf(__fsub_rn(ixmax, ixmin), 0.0f));
        // inter == 0  =>  IOU == +0 exactly (union >= max(ai, aj) > 0)
        if (inter != 0.0f) return iou_exceeds(inter, __fsub_rn(__fadd_rn(ai, aj), inter), t);
    }
    return 0.0f > t.thr;
}

// physical index of counter (digit d, warp w): digit-major like the scan order, one pad word per
// 32 entries so that the 32 lanes of a warp (same w, different d) hit 32 different banks
__device__ __forceinline__ int cnt_index(uint32_t d, int w) { return (int)(d * 33u) + w; }
constexpr int CNT_WORDS = 256 * 33;

__global__ void __launch_bounds__(PR_THREADS, 1) proposal_kernel(PropParams p) {
    extern __shared__ float4 smem4[];
    __shared__ PropShared sh;
    const int tid = threadIdx.x, lane = lane_id(), warp = warp_id();
    const int b = blockIdx.x, N = p.N;
    const float* scores = p.scores + (long long)b * N;

    uint32_t* keyA = reinterpret_cast<uint32_t*>(smem4);
    uint32_t* idxA = keyA + p.kcap;
    uint32_t* keyB = idxA + p.kcap;
    uint32_t* idxB = keyB + p.kcap;
    unsigned int* cnt = idxB + p.kcap;                 // [CNT_WORDS]
    float4* kbox = reinterpret_cast<float4*>(cnt + CNT_WORDS);  // [max_out]
    float4* cbox = kbox + p.mo_pad;                   // [NMS_CHUNK]
    float* karea = reinterpret_cast<float*>(cbox + NMS_CHUNK);  // [max_out]
    float* carea = karea + p.mo_pad;                  // [NMS_CHUNK]
    unsigned int* alive = reinterpret_cast<unsigned int*>(carea + NMS_CHUNK);  // [NMS_CHUNK]
    int* slot = reinterpret_cast<int*>(alive + NMS_CHUNK);                     // [NMS_CHUNK]
    unsigned short* mask16 = reinterpret_cast<unsigned short*>(slot + NMS_CHUNK);  // [NMS_CHUNK][8]
    uint32_t* skeys = keyB;  // staged keys alias sort buffer B (free until the first sort pass)

    // ---- phase 0: stage keys, count entries above the score threshold ---------------------------
    unsigned int my_valid = 0;
    for (int i = tid; i < N; i += PR_THREADS) {
        uint32_t key = score_key(scores[i], p.use_sthr, p.score_threshold);
        if (p.staged) skeys[i] = key;
        my_valid += (key != 0u) ? 1u : 0u;
    }
    int M = N;
    if (p.use_sthr) {
        unsigned int total;
        block_excl_scan(my_valid, &sh, &total);
        M = (int)total;
    } else {
        __syncthreads();
    }
    const int K = min(p.k, M);  // entries that get sorted
    int nkept = 0;

    if (K > 0) {
        // ---- phase 1: radix select -- find (shift, P, need_eq): entry selected iff
        //      (key>>shift) > P, or == P and it is among the first need_eq such entries by index
        uint32_t prefix = 0u;
        unsigned int r = (unsigned int)K;
        int shift = 24;
        bool take_all = (K == N);
        if (!take_all) {
            for (int pass = 0; pass < 4; ++pass) {
                for (int i = tid; i < 256; i += PR_THREADS) cnt[i] = 0u;
                __syncthreads();
                for (int i = tid; i < N; i += PR_THREADS) {
                    uint32_t key = p.staged ? skeys[i] : score_key(scores[i], p.use_sthr, p.score_threshold);
                    if (pass == 0 || ((key ^ prefix) >> (shift + 8)) == 0u) atomicAdd(&cnt[(key >> shift) & 255u], 1u);
                }
                __syncthreads();
                if (tid < 32) {
                    const int top = 255 - 8 * lane;
                    unsigned int s = 0;
#pragma unroll
                    for (int d = 0; d < 8; ++d) s += cnt[top - d];
                    unsigned int incl = (unsigned int)warp_incl_scan((int)s);
                    unsigned int excl = incl - s;
                    if (excl < r && r <= incl) {
                        unsigned int acc = excl;
                        for (int d = 0; d < 8; ++d) {
                            unsigned int c = cnt[top - d];
                            if (acc + c >= r) {
                                sh.digit = (unsigned)(top - d);
                                sh.remaining = r - acc;
                                sh.bin_count = c;
                                break;
                            }
                            acc += c;
                        }
                    }
                }
                __syncthreads();
                prefix |= sh.digit << shift;
                r = sh.remaining;
                const bool done = (sh.bin_count == r);
                __syncthreads();
                if (done || pass == 3) break;
                shift -= 8;
            }
        }
        const uint32_t P = take_all ? 0u : (prefix >> shift);
        const unsigned int need_eq = take_all ? 0u : r;

        // ---- phase 2: ordered compaction into (keyA, idxA), ascending index ---------------------
        unsigned int run_gt = 0, run_eq = 0;
        for (int base = 0; base < N; base += PR_THREADS) {
            const int i = base + tid;
            uint32_t key = 0u;
            bool gt = false, eq = false;
            if (i < N) {
                key = p.staged ? skeys[i] : score_key(scores[i], p.use_sthr, p.score_threshold);
                if (take_all) gt = true;
                else {
                    uint32_t hs = key >> shift;
                    gt = hs > P;
                    eq = hs == P;
                }
            }
            unsigned int tot;
            unsigned int ex = block_excl_scan((gt ? 1u : 0u) | (eq ? 0x10000u : 0u), &sh, &tot);
            const unsigned int gt_before = run_gt + (ex & 0xFFFFu);
            const unsigned int eq_before = run_eq + (ex >> 16);
            if (gt || (eq && eq_before < need_eq)) {
                unsigned int pos = gt_before + min(eq_before, need_eq);
                keyA[pos] = ~key;
                idxA[pos] = (uint32_t)i;
            }
            run_gt += tot & 0xFFFFu;
            run_eq += tot >> 16;
            if (run_gt + min(run_eq, need_eq) >= (unsigned int)K) break;
        }
        __syncthreads();

        // ---- phase 3: stable LSD radix sort, 4 x 8 bits, (keyA,idxA) <-> (keyB,idxB) -------------
        // Each warp owns a contiguous segment; counters are per (digit, warp); an exclusive scan in
        // (digit, warp) order turns them into scatter bases; match.any ranks equal digits in a chunk.
        {
            const int seg = (((K + PR_WARPS - 1) / PR_WARPS) + 31) & ~31;
            const int start = min(warp * seg, K), end = min(start + seg, K);
            uint32_t *kin = keyA, *vin = idxA, *kout = keyB, *vout = idxB;
            for (int sft = 0; sft < 32; sft += 8) {
                for (int i = tid; i < CNT_WORDS; i += PR_THREADS) cnt[i] = 0u;
                __syncthreads();
                for (int e = start + lane; e < end; e += 32) atomicAdd(&cnt[cnt_index((kin[e] >> sft) & 255u, warp)], 1u);
                __syncthreads();
                {   // exclusive scan over the 8192 logical counters, 8 consecutive ones per thread
                    unsigned int loc[8], sum = 0;
#pragma unroll
                    for (int q = 0; q < 8; ++q) { const int L = tid * 8 + q; loc[q] = cnt[L + (L >> 5)]; sum += loc[q]; }
                    unsigned int tot;
                    unsigned int ex = block_excl_scan(sum, &sh, &tot);
#pragma unroll
                    for (int q = 0; q < 8; ++q) { const int L = tid * 8 + q; cnt[L + (L >> 5)] = ex; ex += loc[q]; }
                }
                __syncthreads();
                for (int e0 = start; e0 < end; e0 += 32) {
                    const int e = e0 + lane;
                    const bool v = e < end;
                    const uint32_t key = v ? kin[e] : 0u;
                    const uint32_t val = v ? vin[e] : 0u;
                    const uint32_t d = v ? ((key >> sft) & 255u) : (256u + lane);
                    const unsigned peers = __match_any_sync(0xffffffffu, d);
                    const int ci = cnt_index(d & 255u, warp);
                    unsigned int basepos = 0;
                    if (v) basepos = cnt[ci];
                    __syncwarp();
                    if (v) {
                        const unsigned int rank = __popc(peers & ((1u << lane) - 1u));
                        if (rank == 0) cnt[ci] = basepos + __popc(peers);
                        kout[basepos + rank] = key;
                        vout[basepos + rank] = val;
                    }
                    __syncwarp();
                }
                __syncthreads();
                uint32_t* t = kin; kin = kout; kout = t;
                t = vin; vin = vout; vout = t;
            }
            // 4 passes: result is back in (keyA, idxA)
        }

        // ---- phase 4a: top-k outputs (predictor.py:58-60) ------------------------------------------
        if (p.mode == MODE_TOPK) {
            for (int rnk = tid; rnk < K; rnk += PR_THREADS) {
                const uint32_t idx = idxA[rnk];
                p.values[(long long)b * p.k + rnk] = scores[idx];
                p.indices[(long long)b * p.k + rnk] = (int)idx;
                if (p.gathered) p.gathered[(long long)b * p.k + rnk] = ldg_f4(p.boxes + (long long)b * p.box_stride + idx);
            }
            return;
        }

        // ---- phase 4b: greedy NMS over the sorted candidates ------------------------------------
        // Rounds of NMS_CHUNK candidates in score order.  The next round's boxes are fetched (and, in
        // the fused mode, their deltas gathered) while the current round is being resolved.
        const IouThreshold thr = p.iou_thr;
        auto fetch = [&](int pos, float4& a, float4& d, uint32_t& idx) {
            if (pos + tid < K && tid < NMS_CHUNK) {
                idx = idxA[pos + tid];
                if (p.mode == MODE_PROPOSALS) {
                    d = ldg_f4(p.reg + (long long)b * N + idx);
                    a = ldg_f4(p.anchors + idx);
                } else {
                    a = ldg_f4(p.boxes + (long long)b * p.box_stride + idx);
                }
            }
        };
        float4 na = make_float4(0.f, 0.f, 0.f, 0.f), nd = na;
        uint32_t nidx = 0u;
        fetch(0, na, nd, nidx);
        for (int pos = 0; pos < K && nkept < p.max_out; pos += NMS_CHUNK) {
            const int C = min(NMS_CHUNK, K - pos);
            float4 raw = na;
            const uint32_t my_idx = nidx;
            if (tid < C) {
                if (p.mode == MODE_PROPOSALS) {
                    raw = decode_ref(na, mul4(nd, p.var));                               // predictor.py:55-56
                    if (p.clip_decoded) raw = clip01(raw);
                }
                float4 c = make_float4(fminf(raw.x, raw.z), fminf(raw.y, raw.w), fmaxf(raw.x, raw.z), fmaxf(raw.y, raw.w));
                cbox[tid] = c;
                carea[tid] = __fmul_rn(__fsub_rn(c.z, c.x), __fsub_rn(c.w, c.y));
                alive[tid] = 1u;
            }
            __syncthreads();
            fetch(pos + NMS_CHUNK, na, nd, nidx);   // in flight during the tests below
            {   // candidates vs kept list: thread = (candidate c, part), kept j strided by NMS_PARTS
                const int c = tid & (NMS_CHUNK - 1), part = tid >> 7;
                if (c < C) {
                    const float4 cb = cbox[c];
                    const float ca = carea[c];
                    bool dead = false;
                    for (int j = part; j < nkept && !dead; j += NMS_PARTS) dead = nms_suppresses(cb, ca, kbox[j], karea[j], thr);
                    if (dead) alive[c] = 0u;
                }
            }
            __syncthreads();
            {   // intra-chunk predecessor masks: thread = (row i, 16-column group w);
                // bit j set iff j < i, both alive, and j suppresses i
                const int i = tid >> 3, w = tid & 7;
                unsigned int bits = 0u;
                if (i < C && alive[i] && w * 16 < i) {
                    const float4 bi = cbox[i];
                    const float ai = carea[i];
                    const int jend = min(16, i - w * 16);
#pragma unroll 4
                    for (int jj = 0; jj < jend; ++jj) {
                        const int j = w * 16 + jj;
                        if (alive[j] && nms_suppresses(cbox[j], carea[j], bi, ai, thr)) bits |= 1u << jj;
                    }
                }
                mask16[i * 8 + w] = (unsigned short)bits;
            }
            __syncthreads();
            if (warp == 0) {
                // Resolve the chunk in parallel rounds (same result as the sequential greedy sweep):
                // an undecided candidate is REMOVED if a kept predecessor suppresses it, KEPT if no
                // undecided predecessor suppresses it, else stays undecided.  Lane l owns candidates
                // l, 32+l, 64+l, 96+l, so ballot word q is exactly bits [32q, 32q+32).
                const uint4* m4 = reinterpret_cast<const uint4*>(mask16);
                uint4 pr[4];
                unsigned U[4], Kp[4] = {0u, 0u, 0u, 0u};
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int c = q * 32 + lane;
                    pr[q] = m4[c];
                    U[q] = __ballot_sync(0xffffffffu, c < C && alive[c] != 0u);
                }
                while ((U[0] | U[1] | U[2] | U[3]) != 0u) {
                    unsigned nU[4], nK[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const bool und = (U[q] >> lane) & 1u;
                        const unsigned hitK = (pr[q].x & Kp[0]) | (pr[q].y & Kp[1]) | (pr[q].z & Kp[2]) | (pr[q].w & Kp[3]);
                        const unsigned hitU = (pr[q].x & U[0]) | (pr[q].y & U[1]) | (pr[q].z & U[2]) | (pr[q].w & U[3]);
                        const bool keep = und && hitK == 0u && hitU == 0u;
                        const bool stay = und && hitK == 0u && hitU != 0u;
                        nK[q] = __ballot_sync(0xffffffffu, keep);
                        nU[q] = __ballot_sync(0xffffffffu, stay);
                    }
#pragma unroll
                    for (int q = 0; q < 4; ++q) { Kp[q] |= nK[q]; U[q] = nU[q]; }
                }
                // kept candidates take consecutive output slots in score order, capped at max_out
                const int allowed = p.max_out - nkept;
                int before = 0, total = 0;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int rank = before + __popc(Kp[q] & ((1u << lane) - 1u));
                    const bool kept = ((Kp[q] >> lane) & 1u) && rank < allowed;
                    slot[q * 32 + lane] = kept ? nkept + rank : -1;
                    before += __popc(Kp[q]);
                }
                total = before;
                if (lane == 0) sh.nk = min(total, allowed);
            }
            __syncthreads();
            if (tid < C && slot[tid] >= 0) {
                const int s = slot[tid];
                kbox[s] = cbox[tid];
                karea[s] = carea[tid];
                const long long o = (long long)b * p.rows + s;
                p.out_boxes[o] = p.clip_out ? clip01(raw) : raw;
                p.out_scores[o] = scores[my_idx];
                if (p.keep_idx) p.keep_idx[o] = (int)my_idx;
            }
            nkept += sh.nk;
            __syncthreads();
        }
    }
    if (p.mode == MODE_TOPK) return;
    // zero padding (TF pads boxes/scores/classes with 0); keep_idx pads with -1
    for (int rnk = tid; rnk < p.rows; rnk += PR_THREADS) {
        const long long o = (long long)b * p.rows + rnk;
        if (rnk >= nkept) {
            p.out_boxes[o] = make_float4(0.f, 0.f, 0.f, 0.f);
            p.out_scores[o] = 0.0f;
            if (p.keep_idx) p.keep_idx[o] = -1;
        }
        if (p.out_classes) p.out_classes[o] = 0.0f;
    }
    if (tid == 0) p.valid[b] = nkept;
}

static IouThreshold make_threshold(float thr) {
    IouThreshold t;
    t.thr = thr;
    t.fast = (thr >= 1e-30f && thr <= 1e30f) ? 1 : 0;
    const float nxt = nextafterf(thr, INFINITY);
    t.mid = ((double)thr + (double)nxt) * 0.5;   // exact: 25 significant bits
    uint32_t nb;
    memcpy(&nb, &nxt, 4);
    t.tie_up = ((nb & 1u) == 0u) ? 1 : 0;        // a tie rounds to the even mantissa
    return t;
}

static size_t prop_smem_bytes(int kcap, int max_out) {
    max_out = (max_out + 3) & ~3;
    size_t s = (size_t)kcap * 16;                 // keyA, idxA, keyB, idxB
    s += (size_t)CNT_WORDS * 4;                   // cnt
    s += (size_t)(max_out + NMS_CHUNK) * (16 + 4);  // kbox+cbox, karea+carea
    s += (size_t)NMS_CHUNK * (4 + 4 + 16);        // alive, slot, mask16
    return s + 16;
}
constexpr size_t PROP_SMEM_LIMIT = 227 * 1024 - 1024;

// decide the shared-memory plan; returns 0 or an error
static int plan(PropParams& p, size_t* smem) {
    int kcap = (p.k + 3) & ~3;
    if (kcap < 4) kcap = 4;
    // staging the N score keys in sort buffer B needs N <= 2*kcap; grow kcap if that still fits
    int kcap_staged = max(kcap, ((p.N + 1) / 2 + 3) & ~3);
    if (prop_smem_bytes(kcap_staged, p.max_out) <= PROP_SMEM_LIMIT) {
        kcap = kcap_staged;
        p.staged = 1;
    } else {
        p.staged = 0;
    }
    p.kcap = kcap;
    *smem = prop_smem_bytes(kcap, p.max_out);
    if (*smem > PROP_SMEM_LIMIT)
        return fail(TFRPN_ERR_UNSUPPORTED, "top-k/NMS: k=%d with %d outputs needs %zu B of shared memory (> %zu); "
                    "k <= %d is supported by the in-SM sort", p.k, p.max_out, *smem, PROP_SMEM_LIMIT, TFRPN_MAX_SORT_K);
    return 0;
}

static int launch(tfrpn_handle h, PropParams& p, int B, cudaStream_t st) {
    size_t smem = 0;
    p.mo_pad = (p.max_out + 3) & ~3;
    if (int rc = plan(p, &smem)) return rc;
    static thread_local bool attr_set = false;
    if (!attr_set) {
        TFRPN_CHECK_CUDA(cudaFuncSetAttribute(proposal_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PROP_SMEM_LIMIT));
        attr_set = true;
    }
    prof_begin(h, TFRPN_K_PROPOSAL, st);
    proposal_kernel<<<B, PR_THREADS, smem, st>>>(p);
    prof_end(h, st);
    TFRPN_AFTER_LAUNCH("proposal_kernel");
    return 0;
}

}  // namespace tfrpn

using namespace tfrpn;

extern "C" int tfrpn_topk(tfrpn_handle h, const float* scores, int B, int N, int k, float* values, int32_t* indices,
                          const float* boxes_or_null, int boxes_batched, float* gathered_or_null, tfrpn_stream s) {
    if (!scores || !values || !indices) return fail(TFRPN_ERR_BAD_ARG, "topk: null pointer");
    if (B < 0 || N < 0 || k < 0) return fail(TFRPN_ERR_BAD_ARG, "topk: negative shape");
    if (k > N) return fail(TFRPN_ERR_BAD_ARG, "topk: k=%d > N=%d (tf.nn.top_k raises InvalidArgumentError too)", k, N);
    if ((gathered_or_null != nullptr) != (boxes_or_null != nullptr))
        return fail(TFRPN_ERR_BAD_ARG, "topk: boxes and gathered must be given together");
    if (boxes_or_null && (!aligned16(boxes_or_null) || !aligned16(gathered_or_null)))
        return fail(TFRPN_ERR_MISALIGNED, "topk: boxes must be 16-byte aligned");
    if (B == 0 || k == 0) return 0;
    PropParams p = {};
    p.mode = MODE_TOPK; p.N = N; p.k = k; p.scores = scores; p.use_sthr = 0;
    p.boxes = reinterpret_cast<const float4*>(boxes_or_null); p.box_stride = boxes_batched ? N : 0;
    p.values = values; p.indices = indices; p.gathered = reinterpret_cast<float4*>(gathered_or_null);
    p.max_out = 0; p.rows = 0;
    return launch(h, p, B, as_stream(s));
}

extern "C" int tfrpn_nms(tfrpn_handle h, const float* boxes, const float* scores, int B, int K, const tfrpn_nms_cfg* cfg,
                         float* out_boxes, float* out_scores, float* out_classes, int32_t* valid,
                         int32_t* keep_idx_or_null, tfrpn_stream s) {
    if (!boxes || !scores || !cfg || !out_boxes || !out_scores || !valid) return fail(TFRPN_ERR_BAD_ARG, "nms: null pointer");
    if (B < 0 || K < 0) return fail(TFRPN_ERR_BAD_ARG, "nms: negative shape");
    if (cfg->max_output_size_per_class <= 0 || cfg->max_total_size <= 0)
        return fail(TFRPN_ERR_BAD_ARG, "nms: max_output_size_per_class and max_total_size must be > 0");
    if (!aligned16(boxes) || !aligned16(out_boxes)) return fail(TFRPN_ERR_MISALIGNED, "nms: boxes must be 16-byte aligned");
    if (B == 0) return 0;
    PropParams p = {};
    p.mode = MODE_NMS; p.N = K; p.k = K; p.scores = scores;
    p.use_sthr = !(cfg->score_threshold == -INFINITY);
    p.score_threshold = cfg->score_threshold;
    p.boxes = reinterpret_cast<const float4*>(boxes); p.box_stride = K;
    p.rows = cfg->pad_per_class ? min(cfg->max_total_size, cfg->max_output_size_per_class) : cfg->max_total_size;
    p.max_out = min(cfg->max_output_size_per_class, p.rows);
    p.iou_thr = make_threshold(cfg->iou_threshold); p.clip_out = cfg->clip_boxes;
    p.out_boxes = reinterpret_cast<float4*>(out_boxes); p.out_scores = out_scores; p.out_classes = out_classes;
    p.valid = valid; p.keep_idx = keep_idx_or_null;
    return launch(h, p, B, as_stream(s));
}

extern "C" int tfrpn_proposals(tfrpn_handle h, const float* rpn_reg, const float* rpn_cls, const float* anchors, int B,
                               int N, const tfrpn_proposal_cfg* cfg, float* out_boxes, float* out_scores,
                               int32_t* valid, int32_t* keep_idx_or_null, tfrpn_stream s) {
    if (!rpn_reg || !rpn_cls || !anchors || !cfg || !out_boxes || !out_scores || !valid)
        return fail(TFRPN_ERR_BAD_ARG, "proposals: null pointer");
    if (B < 0 || N < 0) return fail(TFRPN_ERR_BAD_ARG, "proposals: negative shape");
    if (cfg->pre_nms_topn <= 0 || cfg->post_nms_topn <= 0) return fail(TFRPN_ERR_BAD_ARG, "proposals: topn must be > 0");
    if (!aligned16(rpn_reg) || !aligned16(anchors) || !aligned16(out_boxes))
        return fail(TFRPN_ERR_MISALIGNED, "proposals: rpn_reg / anchors / out_boxes must be 16-byte aligned");
    if (B == 0) return 0;
    PropParams p = {};
    p.mode = MODE_PROPOSALS; p.N = N; p.k = min(cfg->pre_nms_topn, N); p.scores = rpn_cls; p.use_sthr = 0;
    p.reg = reinterpret_cast<const float4*>(rpn_reg); p.anchors = reinterpret_cast<const float4*>(anchors);
    p.var = make_float4(cfg->variances[0], cfg->variances[1], cfg->variances[2], cfg->variances[3]);
    p.clip_decoded = cfg->clip;
    p.rows = cfg->post_nms_topn; p.max_out = cfg->post_nms_topn;
    p.iou_thr = make_threshold(cfg->nms_iou_threshold); p.clip_out = 1;  // combined NMS default clip_boxes=True
    p.out_boxes = reinterpret_cast<float4*>(out_boxes); p.out_scores = out_scores; p.out_classes = nullptr;
    p.valid = valid; p.keep_idx = keep_idx_or_null;
    return launch(h, p, B, as_stream(s));
}
