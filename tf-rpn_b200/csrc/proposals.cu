// proposals.cu -- pre-NMS top-k, NMS and the fused proposal stage, one CTA per image.
//
// Replaces tf.nn.top_k + tf.gather (predictor.py:58-60), non_max_suppression ->
// tf.image.combined_non_max_suppression (utils/bbox_utils.py:48-70) and their composition
// (SURVEY.md 8a row P).
//
// Order: every entry gets the unique 64-bit composite  (orderable(score) << 32) | ~index , so
// "larger composite" == "higher score, and among equal scores the LOWER index" -- the order of
// tf.nn.top_k and the order in which TF's NMS pops candidates ([TF-internal]; ties documented in
// the oracle).  NMS is lazy: it consumes candidates in BATCHES of 1024 ranks and usually stops
// (max_output_size kept) inside the first batch, so the other ~5000 of the pre-NMS top-6000 are
// never sorted, gathered or decoded.  Per batch, inside the CTA:
//   1 radix SELECT (MSB first, 8 bits/pass, early exit) of the composite at rank `hi`
//   2 unordered COMPACTION of the entries with rank in [lo, hi) into shared memory
//   3 BITONIC SORT of those <= 1024 composites, one per thread (shuffles below stride 32)
//   4 top-k outputs, or greedy NMS rounds of 128 candidates: test against the kept list, build
//     predecessor masks inside the round, resolve them in one warp with ballots.
//   In the fused mode boxes are decoded (+clipped) on the fly, for the examined candidates only.
#include <math.h>
#include <string.h>

#include "common.cuh"

namespace tfrpn {

constexpr int PR_THREADS = 1024;
constexpr int PR_WARPS = PR_THREADS / 32;
constexpr int BATCH = PR_THREADS;   // ranks sorted per batch: one composite per thread
constexpr int NMS_CHUNK = 128;
constexpr int NMS_PARTS = PR_THREADS / NMS_CHUNK;  // 8

enum { MODE_TOPK = 0, MODE_NMS = 1, MODE_PROPOSALS = 2 };

struct PropParams {
    int mode;
    int N;       // entries per image
    int k;       // requested top-k (<= N)
    int staged;  // score keys staged in shared memory (N * 4 bytes)
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
    unsigned int hist[256];
    unsigned int wtot[PR_WARPS];
    unsigned int digit, remaining, bin_count;
    unsigned int count;
    int nk;
};

// score -> sortable key; entries at or below the score threshold get key 0 (below every real key)
__device__ __forceinline__ uint32_t score_key(float s, int use_sthr, float sthr) {
    if (use_sthr && !(s > sthr)) return 0u;
    return orderable(s);
}
__device__ __forceinline__ unsigned long long make_comp(uint32_t key, int i) {
    return ((unsigned long long)key << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)i);
}

// block-wide sum of one unsigned per thread
__device__ __forceinline__ unsigned int block_sum(unsigned int v, PropShared* sh) {
    unsigned int incl = (unsigned int)warp_incl_scan((int)v);
    if (lane_id() == 31) sh->wtot[warp_id()] = incl;
    __syncthreads();
    if (warp_id() == 0) {
        unsigned int w = sh->wtot[lane_id()];
        unsigned int wi = (unsigned int)warp_incl_scan((int)w);
        if (lane_id() == 31) sh->count = wi;
    }
    __syncthreads();
    unsigned int r = sh->count;
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

// The same decision for SANITISED boxes and a normal positive threshold (t.fast): a box whose area is not
// > 0 has been replaced by FAR_BOX with area 0 when it was staged, so its intersection with anything is
// +0 and the area guards leave the loop; the two float pre-filters of iou_exceeds are evaluated without
// branches and only a quotient within 2^-12 of the threshold (or a degenerate pair) takes the exact path.
__device__ __forceinline__ bool nms_suppresses_fast(float4 ci, float ai, float4 cj, float aj, const IouThreshold& t) {
    const float iymin = fmaxf(ci.x, cj.x), ixmin = fmaxf(ci.y, cj.y);
    const float iymax = fminf(ci.z, cj.z), ixmax = fminf(ci.w, cj.w);
    const float inter = __fmul_rn(fmaxf(__fsub_rn(iymax, iymin), 0.0f), fmaxf(__fsub_rn(ixmax, ixmin), 0.0f));
    const float uni = __fsub_rn(__fadd_rn(ai, aj), inter);
    const bool sure = inter > __fmul_rn(t.hi_f, uni);
    const bool maybe = inter >= __fmul_rn(t.lo_f, uni);
    if (maybe && !sure) {
        if (inter == 0.0f) return false;   // degenerate pair (union 0): IOU is 0, and 0 > thr is false for thr > 0
        const double prod = __dmul_rn(t.mid, (double)uni);
        const double di = (double)inter;
        return di > prod || (di == prod && t.tie_up);
    }
    return sure;
}
#define TFRPN_FAR_BOX make_float4(3.0e38f, 3.0e38f, -3.0e38f, -3.0e38f)

// descending bitonic sort of PR_THREADS composites, one per thread; thread t ends with rank t.
// Strides < 32 use shuffles; strides >= 32 go through a double-buffered shared array (1 barrier each).
__device__ __forceinline__ unsigned long long bitonic_sort_desc(unsigned long long v, unsigned long long* buf) {
    const int t = threadIdx.x;
    int flip = 0;
    for (int k = 2; k <= PR_THREADS; k <<= 1) {
        const bool desc = (t & k) == 0;
        for (int j = k >> 1; j > 0; j >>= 1) {
            unsigned long long pv;
            if (j >= 32) {
                unsigned long long* bb = buf + flip * PR_THREADS;
                bb[t] = v;
                __syncthreads();
                pv = bb[t ^ j];
                flip ^= 1;
            } else {
                pv = __shfl_xor_sync(0xffffffffu, v, j);
            }
            const bool lower = (t & j) == 0;
            const bool keep_max = (lower == desc);
            v = keep_max ? max(v, pv) : min(v, pv);
        }
    }
    return v;
}

__global__ void __launch_bounds__(PR_THREADS, 1) proposal_kernel(PropParams p) {
    extern __shared__ float4 smem4[];
    __shared__ PropShared sh;
    const int tid = threadIdx.x, lane = lane_id(), warp = warp_id();
    const int b = blockIdx.x, N = p.N;
    const float* scores = p.scores + (long long)b * N;

    unsigned long long* sortbuf = reinterpret_cast<unsigned long long*>(smem4);   // [2 * BATCH]
    uint32_t* sidx = reinterpret_cast<uint32_t*>(sortbuf + 2 * BATCH);            // [BATCH] sorted indices
    float4* kbox = reinterpret_cast<float4*>(sidx + BATCH);                       // [mo_pad]
    float4* cbox = kbox + p.mo_pad;                                               // [NMS_CHUNK]
    float* karea = reinterpret_cast<float*>(cbox + NMS_CHUNK);                    // [mo_pad]
    float* carea = karea + p.mo_pad;                                              // [NMS_CHUNK]
    unsigned int* alive = reinterpret_cast<unsigned int*>(carea + NMS_CHUNK);     // [NMS_CHUNK]
    int* slot = reinterpret_cast<int*>(alive + NMS_CHUNK);                        // [NMS_CHUNK]
    unsigned int* mask32 = reinterpret_cast<unsigned int*>(slot + NMS_CHUNK);       // [NMS_CHUNK][4]
    uint32_t* skeys = mask32 + NMS_CHUNK * 4;                                     // [N] when staged

    // ---- phase 0: stage keys, count entries above the score threshold ---------------------------
    unsigned int my_valid = 0;
    for (int i = tid; i < N; i += PR_THREADS) {
        uint32_t key = score_key(scores[i], p.use_sthr, p.score_threshold);
        if (p.staged) skeys[i] = key;
        my_valid += (key != 0u) ? 1u : 0u;
    }
    int M = N;
    if (p.use_sthr) M = (int)block_sum(my_valid, &sh);
    else __syncthreads();
    const int K = min(p.k, M);  // ranks that may be consumed
    int nkept = 0;
    const IouThreshold thr = p.iou_thr;

    auto key_at = [&](int i) -> uint32_t {
        return p.staged ? skeys[i] : score_key(scores[i], p.use_sthr, p.score_threshold);
    };

    // rule "rank < r":  (comp >> shift) >= P   (the lo rule of the first batch selects nothing)
    int lo_shift = 0;
    unsigned long long lo_P = ~0ull;
    bool have_lo = false;

    for (int lo = 0; lo < K; lo += BATCH) {
        const int hi = min(lo + BATCH, K);
        // ---- 1. radix select of the composite at rank hi --------------------------------------
        int shift = 0;
        unsigned long long P = 0ull;          // hi == N: everything is selected
        if (hi < N) {
            unsigned long long prefix = 0ull;
            unsigned int r = (unsigned int)hi;
            shift = 56;
            for (int pass = 0; pass < 8; ++pass) {
                if (tid < 256) sh.hist[tid] = 0u;
                __syncthreads();
                const unsigned long long himask = pass == 0 ? 0ull : (~0ull << (shift + 8));
                for (int i = tid; i < N; i += PR_THREADS) {
                    const unsigned long long c = make_comp(key_at(i), i);
                    if (((c ^ prefix) & himask) == 0ull) atomicAdd(&sh.hist[(unsigned)(c >> shift) & 255u], 1u);
                }
                __syncthreads();
                if (tid < 32) {   // lane l owns digits [248-8l, 255-8l], scanned from the top
                    const int top = 255 - 8 * lane;
                    unsigned int s = 0;
#pragma unroll
                    for (int d = 0; d < 8; ++d) s += sh.hist[top - d];
                    const unsigned int incl = (unsigned int)warp_incl_scan((int)s);
                    const unsigned int excl = incl - s;
                    if (excl < r && r <= incl) {
                        unsigned int acc = excl;
                        for (int d = 0; d < 8; ++d) {
                            const unsigned int c = sh.hist[top - d];
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
                prefix |= (unsigned long long)sh.digit << shift;
                r = sh.remaining;
                const bool done = (sh.bin_count == r);   // the whole bin is taken: stop refining
                __syncthreads();
                if (done || pass == 7) break;
                shift -= 8;
            }
            P = prefix >> shift;
        }
        // ---- 2. compaction of ranks [lo, hi) (any order: composites are unique) -----------------
        if (tid == 0) sh.count = 0u;
        sortbuf[tid] = 0ull;                    // padding sorts last
        __syncthreads();
        for (int base = 0; base < N; base += PR_THREADS) {
            const int i = base + tid;
            bool take = false;
            unsigned long long c = 0ull;
            if (i < N) {
                c = make_comp(key_at(i), i);
                take = ((c >> shift) >= P) && !(have_lo && ((c >> lo_shift) >= lo_P));
            }
            const unsigned bal = __ballot_sync(0xffffffffu, take);
            if (bal != 0u) {
                unsigned int basepos = 0;
                if (lane == __ffs(bal) - 1) basepos = atomicAdd(&sh.count, (unsigned)__popc(bal));
                basepos = __shfl_sync(0xffffffffu, basepos, __ffs(bal) - 1);
                if (take) sortbuf[basepos + __popc(bal & ((1u << lane) - 1u))] = c;
            }
        }
        __syncthreads();
        // ---- 3. sort: thread t gets the composite of rank lo + t ---------------------------------
        unsigned long long mine = sortbuf[tid];
        __syncthreads();
        mine = bitonic_sort_desc(mine, sortbuf);
        const int nb = hi - lo;                 // entries in this batch
        const uint32_t my_i = 0xFFFFFFFFu - (uint32_t)(mine & 0xFFFFFFFFull);
        lo_shift = shift; lo_P = P; have_lo = true;

        // ---- 4a. top-k outputs (predictor.py:58-60) -------------------------------------------
        if (p.mode == MODE_TOPK) {
            if (tid < nb) {
                const long long o = (long long)b * p.k + lo + tid;
                p.values[o] = scores[my_i];
                p.indices[o] = (int)my_i;
                if (p.gathered) {
                    if (p.reg) {   // predictor.py:55-56 for the selected rows only: decode(anchor, delta * variances)
                        float4 bx = decode_ref(ldg_f4(p.anchors + my_i), mul4(ldg_f4(p.reg + (long long)b * N + my_i), p.var));
                        p.gathered[o] = p.clip_decoded ? clip01(bx) : bx;
                    } else {
                        p.gathered[o] = ldg_f4(p.boxes + (long long)b * p.box_stride + my_i);
                    }
                }
            }
            __syncthreads();   // sortbuf is rewritten by the next batch
            continue;
        }

        // ---- 4b. greedy NMS rounds over this batch ---------------------------------------------
        sidx[tid] = my_i;
        __syncthreads();
        auto fetch = [&](int pos, float4& a, float4& d, uint32_t& idx) {
            if (tid < NMS_CHUNK && pos + tid < nb) {
                idx = sidx[pos + tid];
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
        for (int pos = 0; pos < nb && nkept < p.max_out; pos += NMS_CHUNK) {
            const int C = min(NMS_CHUNK, nb - pos);
            float4 raw = na;
            const uint32_t my_idx = nidx;
            if (tid < C) {
                if (p.mode == MODE_PROPOSALS) {
                    raw = decode_ref(na, mul4(nd, p.var));                               // predictor.py:55-56
                    if (p.clip_decoded) raw = clip01(raw);
                }
                float4 c = make_float4(fminf(raw.x, raw.z), fminf(raw.y, raw.w), fmaxf(raw.x, raw.z), fmaxf(raw.y, raw.w));
                float ca = __fmul_rn(__fsub_rn(c.z, c.x), __fsub_rn(c.w, c.y));
                if (thr.fast && !(ca > 0.0f)) { c = TFRPN_FAR_BOX; ca = 0.0f; }   // see nms_suppresses_fast
                cbox[tid] = c;
                carea[tid] = ca;
                alive[tid] = 1u;
            }
            if (tid < NMS_CHUNK * 4) mask32[tid] = 0u;
            __syncthreads();
            fetch(pos + NMS_CHUNK, na, nd, nidx);   // in flight during the tests below
            {   // candidates vs kept list: thread = (candidate c, part), kept j strided by NMS_PARTS
                const int c = tid & (NMS_CHUNK - 1), part = tid >> 7;
                if (c < C) {
                    const float4 cb = cbox[c];
                    const float ca = carea[c];
                    bool dead = false;
                    if (thr.fast) {
                        int j = part;
                        for (; j + NMS_PARTS < nkept && !dead; j += 2 * NMS_PARTS) {
                            const bool d0 = nms_suppresses_fast(cb, ca, kbox[j], karea[j], thr);
                            const bool d1 = nms_suppresses_fast(cb, ca, kbox[j + NMS_PARTS], karea[j + NMS_PARTS], thr);
                            dead = d0 || d1;
                        }
                        if (!dead && j < nkept) dead = nms_suppresses_fast(cb, ca, kbox[j], karea[j], thr);
                    } else {
                        for (int j = part; j < nkept && !dead; j += NMS_PARTS) dead = nms_suppresses(cb, ca, kbox[j], karea[j], thr);
                    }
                    if (dead) alive[c] = 0u;
                }
            }
            __syncthreads();
            {   // intra-round predecessor masks: bit j of row i set iff j < i, both alive, j suppresses i.
                // The triangle is folded so that every 16-thread team gets the same number of pairs:
                // team r takes row iA = r + 1 (iA pairs) and row iB = C - 1 - r (iB pairs).
                const int r = tid >> 4, sub = tid & 15;
                const int iA = r + 1, iB = C - 1 - r;
                const int len = (iA < iB) ? iA + iB : (iA == iB ? iA : 0);
                for (int e = sub; e < len; e += 16) {
                    const int i = e < iA ? iA : iB;
                    const int j = e < iA ? e : e - iA;
                    if (alive[i] && alive[j] &&
                        (thr.fast ? nms_suppresses_fast(cbox[j], carea[j], cbox[i], carea[i], thr)
                                  : nms_suppresses(cbox[j], carea[j], cbox[i], carea[i], thr)))
                        atomicOr(&mask32[i * 4 + (j >> 5)], 1u << (j & 31));
                }
            }
            __syncthreads();
            if (warp == 0) {
                // Resolve the round in parallel sweeps (same result as the sequential greedy loop):
                // an undecided candidate is REMOVED if a kept predecessor suppresses it, KEPT if no
                // undecided predecessor suppresses it, else stays undecided.  Lane l owns candidates
                // l, 32+l, 64+l, 96+l, so ballot word q is exactly bits [32q, 32q+32).
                const uint4* m4 = reinterpret_cast<const uint4*>(mask32);
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
                        nK[q] = __ballot_sync(0xffffffffu, und && hitK == 0u && hitU == 0u);
                        nU[q] = __ballot_sync(0xffffffffu, und && hitK == 0u && hitU != 0u);
                    }
#pragma unroll
                    for (int q = 0; q < 4; ++q) { Kp[q] |= nK[q]; U[q] = nU[q]; }
                }
                // kept candidates take consecutive output slots in score order, capped at max_out
                const int allowed = p.max_out - nkept;
                int before = 0;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int rank = before + __popc(Kp[q] & ((1u << lane) - 1u));
                    const bool kept = ((Kp[q] >> lane) & 1u) && rank < allowed;
                    slot[q * 32 + lane] = kept ? nkept + rank : -1;
                    before += __popc(Kp[q]);
                }
                if (lane == 0) sh.nk = min(before, allowed);
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
        if (nkept >= p.max_out) break;
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
    t.lo_f = (float)((double)thr * (1.0 - 1.0 / 4096.0));
    t.hi_f = (float)((double)thr * (1.0 + 1.0 / 4096.0));
    uint32_t nb;
    memcpy(&nb, &nxt, 4);
    t.tie_up = ((nb & 1u) == 0u) ? 1 : 0;        // a tie rounds to the even mantissa
    return t;
}

static size_t prop_smem_bytes(int n_staged, int max_out) {
    max_out = (max_out + 3) & ~3;
    size_t s = (size_t)2 * BATCH * 8;               // sortbuf (double buffer)
    s += (size_t)BATCH * 4;                         // sidx
    s += (size_t)(max_out + NMS_CHUNK) * (16 + 4);  // kbox+cbox, karea+carea
    s += (size_t)NMS_CHUNK * (4 + 4 + 16);          // alive, slot, mask16
    s += (size_t)n_staged * 4;                      // staged score keys
    return s + 16;
}
constexpr size_t PROP_SMEM_LIMIT = 227 * 1024 - 4096;   // leaves room for the static PropShared

static int launch(tfrpn_handle h, PropParams& p, int B, cudaStream_t st) {
    p.mo_pad = (p.max_out + 3) & ~3;
    p.staged = prop_smem_bytes(p.N, p.max_out) <= PROP_SMEM_LIMIT ? 1 : 0;
    const size_t smem = prop_smem_bytes(p.staged ? p.N : 0, p.max_out);
    if (smem > PROP_SMEM_LIMIT)
        return fail(TFRPN_ERR_UNSUPPORTED, "NMS: %d output rows need %zu B of shared memory (> %zu)", p.max_out, smem,
                    PROP_SMEM_LIMIT);
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

extern "C" int tfrpn_predict_topk(tfrpn_handle h, const float* rpn_reg, const float* rpn_cls, const float* anchors, int B,
                                  int N, int k, const float* variances_host, int clip, float* out_boxes,
                                  float* out_scores, int32_t* out_indices, tfrpn_stream s) {
    if (!rpn_reg || !rpn_cls || !anchors || !variances_host || !out_boxes || !out_scores || !out_indices)
        return fail(TFRPN_ERR_BAD_ARG, "predict_topk: null pointer");
    if (B < 0 || N < 0 || k < 0) return fail(TFRPN_ERR_BAD_ARG, "predict_topk: negative shape");
    if (k > N) return fail(TFRPN_ERR_BAD_ARG, "predict_topk: k=%d > N=%d (tf.nn.top_k raises InvalidArgumentError too)", k, N);
    if (!aligned16(rpn_reg) || !aligned16(anchors) || !aligned16(out_boxes))
        return fail(TFRPN_ERR_MISALIGNED, "predict_topk: rpn_reg / anchors / out_boxes must be 16-byte aligned");
    if (B == 0 || k == 0) return 0;
    PropParams p = {};
    p.mode = MODE_TOPK; p.N = N; p.k = k; p.scores = rpn_cls; p.use_sthr = 0;
    p.reg = reinterpret_cast<const float4*>(rpn_reg); p.anchors = reinterpret_cast<const float4*>(anchors);
    p.var = make_float4(variances_host[0], variances_host[1], variances_host[2], variances_host[3]);
    p.clip_decoded = clip;
    p.values = out_scores; p.indices = out_indices; p.gathered = reinterpret_cast<float4*>(out_boxes);
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
