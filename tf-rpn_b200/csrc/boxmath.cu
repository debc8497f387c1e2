// boxmath.cu -- anchors, IoU map, delta encode / decode, (de)normalise.
// Replaces utils/bbox_utils.py:3-46, :72-96, :98-124, :126-150, :152-182 of the reference.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace tfrpn {

// ------------------------------------------------------------------------------------------------
// K0 anchors (utils/bbox_utils.py:23-46).  One thread per anchor; the grid coordinate is evaluated
// in float64 exactly like `tf.range(0,F)/F + stride/2` (int32 truediv -> f64) and rounded once.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) anchors_kernel(const __grid_constant__ AnchorGen gen, float4* __restrict__ out) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= gen.fm_h * gen.fm_w * gen.A) return;
    out[n] = anchor_at(gen, n);
}

static int check_anchor_cfg(const tfrpn_anchor_cfg* cfg) {
    if (!cfg) return fail(TFRPN_ERR_BAD_ARG, "anchor cfg is null");
    if (cfg->img_h <= 0 || cfg->img_w <= 0 || cfg->fm_h <= 0 || cfg->fm_w <= 0)
        return fail(TFRPN_ERR_BAD_ARG, "img_size / feature_map_shape must be positive");
    if (cfg->n_scales <= 0 || cfg->n_scales > 8 || cfg->n_ratios <= 0 || cfg->n_ratios > 8)
        return fail(TFRPN_ERR_BAD_ARG, "1..8 anchor scales and ratios supported");
    return 0;
}

// utils/bbox_utils.py:3-21 on the host, same dtype walk: the quotient scale^2/ratio is a double,
// rounded to f32, then f32 sqrt; h = w * f32(ratio); halves are exact.
static int base_anchors_host(const tfrpn_anchor_cfg* cfg, float* out) {
    int a = 0;
    for (int s = 0; s < cfg->n_scales; ++s) {
        double sw = cfg->scales[s] / (double)cfg->img_w;
        double sh = cfg->scales[s] / (double)cfg->img_h;
        for (int r = 0; r < cfg->n_ratios; ++r, ++a) {
            double ratio = cfg->ratios[r];
            float w = sqrtf((float)(pow(sw, 2.0) / ratio));
            float h = sqrtf((float)(pow(sh, 2.0) / ratio)) * (float)ratio;
            out[a * 4 + 0] = -h / 2.0f;
            out[a * 4 + 1] = -w / 2.0f;
            out[a * 4 + 2] = h / 2.0f;
            out[a * 4 + 3] = w / 2.0f;
        }
    }
    return 0;
}

int make_anchor_gen(const tfrpn_anchor_cfg* cfg, AnchorGen* out, long long* n_anchors) {
    if (int rc = check_anchor_cfg(cfg)) return rc;
    float tmp[TFRPN_MAX_BASE_ANCHORS * 4];
    base_anchors_host(cfg, tmp);
    const int A = cfg->n_scales * cfg->n_ratios;
    memset(out, 0, sizeof(*out));
    for (int a = 0; a < A; ++a) out->base[a] = make_float4(tmp[4 * a], tmp[4 * a + 1], tmp[4 * a + 2], tmp[4 * a + 3]);
    out->A = A; out->fm_h = cfg->fm_h; out->fm_w = cfg->fm_w;
    out->half_stride_y = (1.0 / cfg->fm_h) / 2; out->half_stride_x = (1.0 / cfg->fm_w) / 2;
    const long long N = (long long)cfg->fm_h * cfg->fm_w * A;
    if (N > (1LL << 30)) return fail(TFRPN_ERR_BAD_ARG, "too many anchors");
    if (n_anchors) *n_anchors = N;
    return 0;
}

// ------------------------------------------------------------------------------------------------
// K1 generate_iou_map (utils/bbox_utils.py:126-150), materialised (B,N,G).  HBM-write bound:
// 4*B*N*G bytes out.  One CTA owns IOU_TILE_N boxes of one image and walks its CONTIGUOUS output
// tile.  COLS variant (G <= 128): a thread keeps ONE GT box (column g = t % G) in registers for the
// whole tile and steps over rows, R = 256/G rows per iteration, so consecutive threads still write
// consecutive floats; per element that leaves one broadcast LDS.128 of the box row + ~14 ALU ops.
// Generic variant: any G, (n,g) advanced incrementally, both operands from shared memory.
// ------------------------------------------------------------------------------------------------
constexpr int IOU_THREADS = 256;
constexpr int IOU_TILE_N = 256;

template <bool COLS>
__global__ void __launch_bounds__(IOU_THREADS) iou_map_kernel(const float4* __restrict__ boxes,
                                                              long long box_batch_stride,
                                                              const float4* __restrict__ gt, int N, int G,
                                                              float* __restrict__ out) {
    extern __shared__ float4 smem4[];
    float4* sbox = smem4;                      // [IOU_TILE_N]
    float4* sgt = smem4 + IOU_TILE_N;          // [G]        (generic variant only)
    float* sbarea = reinterpret_cast<float*>(sgt + (COLS ? 0 : G));  // [IOU_TILE_N]
    float* sgarea = sbarea + IOU_TILE_N;       // [G]        (generic variant only)

    const int b = blockIdx.y;
    const int n0 = blockIdx.x * IOU_TILE_N;
    const int tn = min(IOU_TILE_N, N - n0);
    const float4* bx = boxes + (long long)b * box_batch_stride + n0;
    bool nice = true;
    for (int i = threadIdx.x; i < tn; i += IOU_THREADS) {
        float4 v = ldg_f4(bx + i);
        sbox[i] = v;
        sbarea[i] = box_area(v);
        nice = nice && nice_box(v);
    }
    const float4* gb = gt + (long long)b * G;
    float* o = out + ((long long)b * N + n0) * G;
    if (COLS) {
        const int R = IOU_THREADS / G;                 // rows per iteration
        const int r = threadIdx.x / G, g = threadIdx.x - r * G;
        const bool active = r < R;
        const float4 gbx = ldg_f4(gb + g);
        const float ga = box_area(gbx);
        // every box of the tile nice and every GT box of the image nice or degenerate (the zero padding)?
        // (the vote is also the barrier that publishes sbox / sbarea)
        nice = nice && nice_coords(gbx) && gbx.z >= gbx.x && gbx.w >= gbx.y;
        const bool all_nice = __syncthreads_and(nice) != 0;
        if (!active) return;
        // 32-bit element offsets from the CTA's base: one IMAD.WIDE (FMA pipe) per store address instead of
        // a 64-bit add pair on the ALU pipe, which is the pipe this kernel saturates first (6 FMNMX per element)
        unsigned e = threadIdx.x;                      // == r*G + g
        asm("" : "+l"(o));                             // keep the base in one register pair (no re-association)
        const unsigned step = (unsigned)(R * G);
        if (all_nice) {   // the division without its range check and the zero test: ~40 % fewer instructions
            // two rows per packed FADD2 / FMUL2 / FFMA2 (common.cuh: iou_nice2): fewer issue slots per element
            const f32x2 ga2 = pack2(ga, ga);
            int n = r;
#pragma unroll 2
            for (; n + R < tn; n += 2 * R, e += 2 * step) {
                float v0, v1;
                iou_nice2(sbox[n], sbox[n + R], pack2(sbarea[n], sbarea[n + R]), gbx, ga2, v0, v1);
                stg_f1_stream(o + e, v0);
                stg_f1_stream(o + (e + step), v1);
            }
            if (n < tn) stg_f1_stream(o + e, iou_nice(sbox[n], sbarea[n], gbx, ga));
        } else {
#pragma unroll 4
            for (int n = r; n < tn; n += R, e += step) stg_f1_stream(o + e, iou_ref(sbox[n], sbarea[n], gbx, ga));
        }
    } else {
        for (int g = threadIdx.x; g < G; g += IOU_THREADS) {
            float4 v = ldg_f4(gb + g);
            sgt[g] = v;
            sgarea[g] = box_area(v);
            nice = nice && nice_coords(v) && v.z >= v.x && v.w >= v.y;
        }
        const bool all_nice = __syncthreads_and(nice) != 0;
        const int total = tn * G;
        const int dn = IOU_THREADS / G, dg = IOU_THREADS - dn * G;
        int e = threadIdx.x;
        int n = e / G, g = e - n * G;
        if (all_nice) {
#pragma unroll 4
            for (; e < total; e += IOU_THREADS) {
                stg_f1_stream(o + e, iou_nice(sbox[n], sbarea[n], sgt[g], sgarea[g]));
                n += dn;
                g += dg;
                if (g >= G) { g -= G; n += 1; }
            }
        } else {
#pragma unroll 4
            for (; e < total; e += IOU_THREADS) {
                stg_f1_stream(o + e, iou_ref(sbox[n], sbarea[n], sgt[g], sgarea[g]));
                n += dn;
                g += dg;
                if (g >= G) { g -= G; n += 1; }
            }
        }
    }
}

// G a multiple of W = 2 or 4 (G <= W * 256): a thread owns W adjacent GT columns and walks the rows, so one
// staged box serves W map entries (one LDS.128 + one LDS per W entries instead of per entry), the quotients come
// out of packed divisions two at a time, and the W entries leave as one 8- or 16-byte streaming store (row
// starts are 4W-byte aligned when W divides G): 1/W of the load / store instructions of the one-column
// kernel.  IoU is symmetric in its two boxes down to the bits (fmax / fmin / fadd commute), so iou_nice2 is
// called with the roles of box and GT swapped.
template <int W, int TILE>
__global__ void __launch_bounds__(IOU_THREADS) iou_map_pairs_kernel(const float4* __restrict__ boxes,
                                                                    long long box_batch_stride,
                                                                    const float4* __restrict__ gt, int N, int G,
                                                                    float* __restrict__ out) {
    __shared__ float4 sbox[TILE];
    __shared__ float sbarea[TILE];
    const int b = blockIdx.y;
    const int n0 = blockIdx.x * TILE;
    const int tn = min(TILE, N - n0);
    const float4* bx = boxes + (long long)b * box_batch_stride + n0;
    bool nice = true;
    for (int i = threadIdx.x; i < tn; i += IOU_THREADS) {
        float4 v = ldg_f4(bx + i);
        sbox[i] = v;
        sbarea[i] = box_area(v);
        nice = nice && nice_box(v);
    }
    const int H = G / W;                               // column groups per row
    const int R = IOU_THREADS / H;                     // rows per iteration
    const int r = threadIdx.x / H, h = threadIdx.x - r * H;
    const bool active = r < R;
    const float4* gb = gt + (long long)b * G + (active ? W * h : 0);
    float4 g[W];
    float ga[W];
#pragma unroll
    for (int q = 0; q < W; ++q) {
        g[q] = ldg_f4(gb + q);
        ga[q] = box_area(g[q]);
        nice = nice && nice_coords(g[q]) && g[q].z >= g[q].x && g[q].w >= g[q].y;
    }
    const bool all_nice = __syncthreads_and(nice) != 0;   // also publishes sbox / sbarea
    if (!active) return;
    float* o = out + ((long long)b * N + n0) * G;
    asm("" : "+l"(o));                                 // keep the base in one register pair (IMAD.WIDE addressing)
    unsigned e = threadIdx.x;                          // == r*H + h, in units of W floats
    const unsigned step = (unsigned)(R * H);
    auto store = [&](unsigned at, const float (&v)[W]) {
        if constexpr (W == 4) stg_f4_stream(reinterpret_cast<float4*>(o) + at, make_float4(v[0], v[1], v[2], v[3]));
        else stg_f2_stream(reinterpret_cast<float2*>(o) + at, v[0], v[1]);
    };
    if (all_nice) {   // separate loops: the choice is uniform, keep it out of the loop body
#pragma unroll 2
        for (int n = r; n < tn; n += R, e += step) {
            const float4 bxn = sbox[n];
            const float ba = sbarea[n];
            const f32x2 ba2 = pack2(ba, ba);
            float v[W];
#pragma unroll
            for (int q = 0; q < W; q += 2) iou_nice2(g[q], g[q + 1], pack2(ga[q], ga[q + 1]), bxn, ba2, v[q], v[q + 1]);
            store(e, v);
        }
    } else {
#pragma unroll 1
        for (int n = r; n < tn; n += R, e += step) {
            const float4 bxn = sbox[n];
            const float ba = sbarea[n];
            float v[W];
#pragma unroll
            for (int q = 0; q < W; ++q) v[q] = iou_ref(bxn, ba, g[q], ga[q]);
            store(e, v);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Elementwise kernels: one float4 box per element.  grid = (ceil(N / (256*2)), B): blockIdx.y is the
// image, so the (N,4) broadcast operand is indexed without a modulo; each thread keeps two
// independent 128-bit loads per operand in flight.
// ------------------------------------------------------------------------------------------------
constexpr int EW_THREADS = 256;
constexpr int EW_PER_THREAD = 2;

// get_deltas_from_bboxes (utils/bbox_utils.py:98-124)
__global__ void __launch_bounds__(EW_THREADS) encode_kernel(const float4* __restrict__ boxes, int boxes_batched,
                                                            const float4* __restrict__ gt, int N,
                                                            float4* __restrict__ out) {
    const long long img = (long long)blockIdx.y * N;
    const float4* bx = boxes + (boxes_batched ? img : 0);
    const int n0 = blockIdx.x * (EW_THREADS * EW_PER_THREAD) + threadIdx.x;
#pragma unroll
    for (int u = 0; u < EW_PER_THREAD; ++u) {
        int n = n0 + u * EW_THREADS;
        if (n < N) {
            float4 b = boxes_batched ? ldg_f4_stream(bx + n) : ldg_f4(bx + n);
            float4 g = ldg_f4_stream(gt + img + n);
            stg_f4_stream(out + img + n, encode_ref(b, g));
        }
    }
}

// get_bboxes_from_deltas (utils/bbox_utils.py:72-96) with the caller-side `deltas *= variances`
// (predictor.py:55) and the proposal pipeline's clip fused in.  K3: 32*B*N algorithmic bytes.
template <bool SCALE, bool CLIP, int PT>
__global__ void __launch_bounds__(EW_THREADS) decode_kernel(const float4* __restrict__ anchors, int anchors_batched,
                                                            const float4* __restrict__ deltas, float4 var, int N,
                                                            float4* __restrict__ out) {
    constexpr int EW_PER_THREAD = PT;
    const long long img = (long long)blockIdx.y * N;
    const float4* an = anchors + (anchors_batched ? img : 0);
    const int n0 = blockIdx.x * (EW_THREADS * EW_PER_THREAD) + threadIdx.x;
    float4 d[EW_PER_THREAD], a[EW_PER_THREAD];
#pragma unroll
    for (int u = 0; u < EW_PER_THREAD; ++u) {
        int n = n0 + u * EW_THREADS;
        if (n < N) {
            d[u] = ldg_f4_stream(deltas + img + n);
            a[u] = anchors_batched ? ldg_f4_stream(an + n) : ldg_f4(an + n);
        }
    }
#pragma unroll
    for (int u = 0; u < EW_PER_THREAD; ++u) {
        int n = n0 + u * EW_THREADS;
        if (n < N) {
            float4 dd = SCALE ? mul4(d[u], var) : d[u];
            float4 o = decode_ref(a[u], dd);
            if (CLIP) o = clip01(o);
            stg_f4_stream(out + img + n, o);
        }
    }
}

// The north star's fused anchor-generate + decode + clip: the same kernel with the anchors regenerated in
// registers (anchor_at) instead of read -- 32 instead of 48 bytes of requests per box (the (N,4) tensor is shared
// by the batch and L2-resident, so the DRAM traffic is the same 32 B).
template <bool SCALE, bool CLIP, int PT>
__global__ void __launch_bounds__(EW_THREADS) decode_gen_kernel(const __grid_constant__ AnchorGen gen,
                                                                const float4* __restrict__ deltas, float4 var, int N,
                                                                float4* __restrict__ out) {
    const long long img = (long long)blockIdx.y * N;
    const int n0 = blockIdx.x * (EW_THREADS * PT) + threadIdx.x;
    float4 d[PT];
#pragma unroll
    for (int u = 0; u < PT; ++u) {
        const int n = n0 + u * EW_THREADS;
        if (n < N) d[u] = ldg_f4_stream(deltas + img + n);
    }
#pragma unroll
    for (int u = 0; u < PT; ++u) {
        const int n = n0 + u * EW_THREADS;
        if (n < N) {
            float4 o = decode_ref(anchor_at(gen, n), SCALE ? mul4(d[u], var) : d[u]);
            if (CLIP) o = clip01(o);
            stg_f4_stream(out + img + n, o);
        }
    }
}

// normalize_bboxes / denormalize_bboxes (utils/bbox_utils.py:152-182)
__global__ void __launch_bounds__(EW_THREADS) scale_boxes_kernel(const float4* __restrict__ in, long long total,
                                                                 float h, float w, int denorm,
                                                                 float4* __restrict__ out) {
    long long stride = (long long)gridDim.x * EW_THREADS;
    for (long long i = (long long)blockIdx.x * EW_THREADS + threadIdx.x; i < total; i += stride) {
        float4 b = ldg_f4_stream(in + i), o;
        if (denorm) {
            o = make_float4(rintf(__fmul_rn(b.x, h)), rintf(__fmul_rn(b.y, w)), rintf(__fmul_rn(b.z, h)),
                            rintf(__fmul_rn(b.w, w)));
        } else {
            o = make_float4(__fdiv_rn(b.x, h), __fdiv_rn(b.y, w), __fdiv_rn(b.z, h), __fdiv_rn(b.w, w));
        }
        stg_f4_stream(out + i, o);
    }
}

// GT-side preprocessing (utils/data_utils.py:54-68 flip, :145-157 padded batch): ragged boxes /
// labels (offsets[b] .. offsets[b+1]) -> (B,G,4) zero padded and (B,G) padded with -1.
__global__ void __launch_bounds__(EW_THREADS) pad_gt_kernel(const float4* __restrict__ flat_boxes,
                                                            const int* __restrict__ flat_labels,
                                                            const int* __restrict__ offsets,
                                                            const unsigned char* __restrict__ flip, int B, int G,
                                                            int label_add, float4* __restrict__ out_boxes,
                                                            int* __restrict__ out_labels) {
    const int i = blockIdx.x * EW_THREADS + threadIdx.x;
    if (i >= B * G) return;
    const int b = i / G, g = i - b * G;
    const int lo = __ldg(offsets + b), n = __ldg(offsets + b + 1) - lo;
    float4 bx = make_float4(0.f, 0.f, 0.f, 0.f);      // get_padding_values: boxes 0, labels -1
    int lab = -1;
    if (g < n) {
        bx = ldg_f4(flat_boxes + lo + g);
        lab = __ldg(flat_labels + lo + g) + label_add;
        if (flip && flip[b])                            // [y1, 1 - x2, y2, 1 - x1], data_utils.py:66-69
            bx = make_float4(bx.x, __fsub_rn(1.0f, bx.w), bx.z, __fsub_rn(1.0f, bx.y));
    }
    out_boxes[i] = bx;
    out_labels[i] = lab;
}

// Self-test of div_rn_inrange and its packed form div2_rn_inrange (common.cuh) against __fdiv_rn on pseudo-random operands of its whole domain:
// b = 2^eb * mb with eb in [-79, 19), a = 0, a = b, or 2^ea * ma with 2^-78 <= a <= b; every eighth pair
// uses extreme mantissas (all ones / all zeros / one bit) where reciprocal refinement is hardest.
__global__ void __launch_bounds__(256) selftest_division_kernel(unsigned long long n, unsigned long long seed,
                                                               unsigned long long* __restrict__ mismatches) {
    unsigned long long bad = 0;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const Philox4 r = philox4x32_10((uint32_t)i, (uint32_t)(i >> 32), 0x51f7u, 0u, (uint32_t)seed, (uint32_t)(seed >> 32));
        uint32_t mb = r.v[0] & 0x7fffffu, ma = r.v[1] & 0x7fffffu;
        if ((r.v[3] & 7u) == 0u) {
            const uint32_t pat[4] = {0x7fffffu, 0u, 1u, 0x400000u};
            mb = pat[(r.v[3] >> 3) & 3u];
            ma = pat[(r.v[3] >> 5) & 3u];
        }
        const int eb = -79 + (int)(r.v[2] % 98u);                              // [-79, 19)
        const float b = __uint_as_float(((uint32_t)(eb + 127) << 23) | mb);
        const int lo = -78, span = eb - lo + 1;                                // exponents of a: [-78, eb]
        float a;
        const uint32_t kind = (r.v[3] >> 8) & 15u;
        if (kind == 0u || span <= 0) a = 0.0f;
        else if (kind == 1u) a = b;
        else {
            const int ea = lo + (int)((r.v[2] >> 8) % (uint32_t)span);
            a = __uint_as_float(((uint32_t)(ea + 127) << 23) | ma);
            if (a > b) a = b;
        }
        const float want = __fdiv_rn(a, b), got = div_rn_inrange(a, b);
        bad += (__float_as_uint(want) != __float_as_uint(got)) ? 1ull : 0ull;
        // the packed form (FFMA2), this pair in one half and the pair (b, b) in the other
        float p0, p1;
        unpack2(div2_rn_inrange(pack2(a, b), pack2(-b, -b)), p0, p1);
        bad += (__float_as_uint(want) != __float_as_uint(p0) || p1 != 1.0f) ? 1ull : 0ull;
    }
    if (bad) atomicAdd(mismatches, bad);
}

int set_attributes_boxmath() {
    TFRPN_CHECK_CUDA(cudaFuncSetAttribute(iou_map_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    return 0;
}

static int ew_grid(long long total, int per_thread) {
    long long blocks = (total + (long long)EW_THREADS * per_thread - 1) / ((long long)EW_THREADS * per_thread);
    if (blocks < 1) blocks = 1;
    if (blocks > 148LL * 32) blocks = 148LL * 32;
    return (int)blocks;
}
static dim3 ew_grid2(int B, int N) {
    return dim3((N + EW_THREADS * EW_PER_THREAD - 1) / (EW_THREADS * EW_PER_THREAD), B);
}

}  // namespace tfrpn

using namespace tfrpn;

extern "C" int tfrpn_base_anchors_host(const tfrpn_anchor_cfg* cfg, float* out_host) {
    if (int rc = check_anchor_cfg(cfg)) return rc;
    if (!out_host) return fail(TFRPN_ERR_BAD_ARG, "out_host is null");
    return base_anchors_host(cfg, out_host);
}

extern "C" int tfrpn_anchors(const tfrpn_anchor_cfg* cfg, float* out, tfrpn_stream s) {
    if (int rc = check_anchor_cfg(cfg)) return rc;
    if (!out) return fail(TFRPN_ERR_BAD_ARG, "out is null");
    if (!aligned16(out)) return fail(TFRPN_ERR_MISALIGNED, "anchors output must be 16-byte aligned");
    TFRPN_ENTER_PTR(out, "anchors: out");
    AnchorGen gen;
    long long N = 0;
    if (int rc = make_anchor_gen(cfg, &gen, &N)) return rc;
    int blocks = (int)((N + 255) / 256);
    anchors_kernel<<<blocks, 256, 0, as_stream(s)>>>(gen, reinterpret_cast<float4*>(out));
    TFRPN_AFTER_LAUNCH("anchors_kernel");
    return 0;
}

extern "C" int tfrpn_iou_map(const float* boxes, int boxes_batched, const float* gt_boxes, int B, int N, int G,
                             float* out, tfrpn_stream s) {
    if (!boxes || !gt_boxes || !out) return fail(TFRPN_ERR_BAD_ARG, "iou_map: null pointer");
    if (B < 0 || N < 0 || G < 0) return fail(TFRPN_ERR_BAD_ARG, "iou_map: negative shape");
    if (B == 0 || N == 0 || G == 0) return 0;
    if (B > 65535) return fail(TFRPN_ERR_UNSUPPORTED, "iou_map: B > 65535");
    if (!aligned16(boxes) || !aligned16(gt_boxes)) return fail(TFRPN_ERR_MISALIGNED, "iou_map: boxes must be 16-byte aligned");
    const bool cols = G <= 128;
    size_t smem = (size_t)(IOU_TILE_N + (cols ? 0 : G)) * (sizeof(float4) + sizeof(float));
    if (smem > 200 * 1024) return fail(TFRPN_ERR_UNSUPPORTED, "iou_map: G=%d too large for shared memory", G);
    TFRPN_ENTER_PTR(out, "iou_map: out");
    dim3 grid((N + IOU_TILE_N - 1) / IOU_TILE_N, B);
    const float4* b4 = reinterpret_cast<const float4*>(boxes);
    const float4* g4 = reinterpret_cast<const float4*>(gt_boxes);
    const long long bstride = boxes_batched ? (long long)N : 0LL;
    static const int max_w = getenv("TFRPN_IOU_W") ? atoi(getenv("TFRPN_IOU_W")) : 4;   // A/B switch: 1, 2 or 4
    const uintptr_t oa = reinterpret_cast<uintptr_t>(out);
    // 512-row tiles amortise the per-CTA prologue (C3: 0.695 -> 0.725 of the HBM peak, C4: 0.886 -> 0.905) as long as
    // they still give every SM its 8 resident CTAs; below that (C2: 64 x 17 tiles) 256-row tiles are faster
    const dim3 grid2((N + 511) / 512, B);
    const bool big = (long long)grid2.x * B >= 8LL * 148;
    if (max_w >= 4 && (G & 3) == 0 && G <= 4 * IOU_THREADS && (oa & 15u) == 0) {
        if (big) iou_map_pairs_kernel<4, 512><<<grid2, IOU_THREADS, 0, as_stream(s)>>>(b4, bstride, g4, N, G, out);
        else iou_map_pairs_kernel<4, IOU_TILE_N><<<grid, IOU_THREADS, 0, as_stream(s)>>>(b4, bstride, g4, N, G, out);
    } else if (max_w >= 2 && (G & 1) == 0 && G <= 2 * IOU_THREADS && (oa & 7u) == 0) {
        if (big) iou_map_pairs_kernel<2, 512><<<grid2, IOU_THREADS, 0, as_stream(s)>>>(b4, bstride, g4, N, G, out);
        else iou_map_pairs_kernel<2, IOU_TILE_N><<<grid, IOU_THREADS, 0, as_stream(s)>>>(b4, bstride, g4, N, G, out);
    }
    else if (cols) iou_map_kernel<true><<<grid, IOU_THREADS, smem, as_stream(s)>>>(b4, bstride, g4, N, G, out);
    else iou_map_kernel<false><<<grid, IOU_THREADS, smem, as_stream(s)>>>(b4, bstride, g4, N, G, out);
    TFRPN_AFTER_LAUNCH("iou_map_kernel");
    return 0;
}

extern "C" int tfrpn_encode_deltas(const float* boxes, int boxes_batched, const float* gt_boxes, int B, int N,
                                   float* out, tfrpn_stream s) {
    if (!boxes || !gt_boxes || !out) return fail(TFRPN_ERR_BAD_ARG, "encode: null pointer");
    if (B < 0 || N < 0) return fail(TFRPN_ERR_BAD_ARG, "encode: negative shape");
    long long total = (long long)B * N;
    if (total == 0) return 0;
    if (!aligned16(boxes) || !aligned16(gt_boxes) || !aligned16(out))
        return fail(TFRPN_ERR_MISALIGNED, "encode: pointers must be 16-byte aligned");
    if (B > 65535) return fail(TFRPN_ERR_UNSUPPORTED, "encode: B > 65535");
    TFRPN_ENTER_PTR(out, "encode: out");
    encode_kernel<<<ew_grid2(B, N), EW_THREADS, 0, as_stream(s)>>>(
        reinterpret_cast<const float4*>(boxes), boxes_batched, reinterpret_cast<const float4*>(gt_boxes), N,
        reinterpret_cast<float4*>(out));
    TFRPN_AFTER_LAUNCH("encode_kernel");
    return 0;
}

extern "C" int tfrpn_decode(const float* anchors, int anchors_batched, const float* deltas,
                            const float* variances_host_or_null, int clip, int B, int N, float* out,
                            tfrpn_stream s) {
    if (!anchors || !deltas || !out) return fail(TFRPN_ERR_BAD_ARG, "decode: null pointer");
    if (B < 0 || N < 0) return fail(TFRPN_ERR_BAD_ARG, "decode: negative shape");
    long long total = (long long)B * N;
    if (total == 0) return 0;
    if (!aligned16(anchors) || !aligned16(deltas) || !aligned16(out))
        return fail(TFRPN_ERR_MISALIGNED, "decode: pointers must be 16-byte aligned");
    float4 var = make_float4(1.f, 1.f, 1.f, 1.f);
    const bool scale = variances_host_or_null != nullptr;
    if (scale) var = make_float4(variances_host_or_null[0], variances_host_or_null[1], variances_host_or_null[2],
                                 variances_host_or_null[3]);
    const float4* a4 = reinterpret_cast<const float4*>(anchors);
    const float4* d4 = reinterpret_cast<const float4*>(deltas);
    float4* o4 = reinterpret_cast<float4*>(out);
    if (B > 65535) return fail(TFRPN_ERR_UNSUPPORTED, "decode: B > 65535");
    TFRPN_ENTER_PTR(out, "decode: out");
    cudaStream_t st = as_stream(s);
    static const int pt = [] { const char* e = getenv("TFRPN_DECODE_PT"); const int v = e ? atoi(e) : 2; return (v == 1 || v == 4) ? v : 2; }();
#define TFRPN_DECODE_LAUNCH(PT)                                                                                      \
    do {                                                                                                             \
        dim3 grid((N + EW_THREADS * PT - 1) / (EW_THREADS * PT), B);                                                 \
        if (scale && clip) decode_kernel<true, true, PT><<<grid, EW_THREADS, 0, st>>>(a4, anchors_batched, d4, var, N, o4);   \
        else if (scale) decode_kernel<true, false, PT><<<grid, EW_THREADS, 0, st>>>(a4, anchors_batched, d4, var, N, o4);     \
        else if (clip) decode_kernel<false, true, PT><<<grid, EW_THREADS, 0, st>>>(a4, anchors_batched, d4, var, N, o4);      \
        else decode_kernel<false, false, PT><<<grid, EW_THREADS, 0, st>>>(a4, anchors_batched, d4, var, N, o4);               \
    } while (0)
    if (pt == 1) TFRPN_DECODE_LAUNCH(1); else if (pt == 4) TFRPN_DECODE_LAUNCH(4); else TFRPN_DECODE_LAUNCH(2);
#undef TFRPN_DECODE_LAUNCH
    TFRPN_AFTER_LAUNCH("decode_kernel");
    return 0;
}

extern "C" int tfrpn_decode_anchor_cfg(const tfrpn_anchor_cfg* acfg, const float* deltas, const float* variances_host_or_null,
                                       int clip, int B, float* out, tfrpn_stream s) {
    if (!deltas || !out) return fail(TFRPN_ERR_BAD_ARG, "decode_anchor_cfg: null pointer");
    if (B < 0) return fail(TFRPN_ERR_BAD_ARG, "decode_anchor_cfg: negative shape");
    AnchorGen gen;
    long long N = 0;
    if (int rc = make_anchor_gen(acfg, &gen, &N)) return rc;
    if (B == 0) return 0;
    if (!aligned16(deltas) || !aligned16(out)) return fail(TFRPN_ERR_MISALIGNED, "decode_anchor_cfg: pointers must be 16-byte aligned");
    if (B > 65535) return fail(TFRPN_ERR_UNSUPPORTED, "decode_anchor_cfg: B > 65535");
    float4 var = make_float4(1.f, 1.f, 1.f, 1.f);
    const bool scale = variances_host_or_null != nullptr;
    if (scale) var = make_float4(variances_host_or_null[0], variances_host_or_null[1], variances_host_or_null[2],
                                 variances_host_or_null[3]);
    TFRPN_ENTER_PTR(out, "decode_anchor_cfg: out");
    const float4* d4 = reinterpret_cast<const float4*>(deltas);
    float4* o4 = reinterpret_cast<float4*>(out);
    cudaStream_t st = as_stream(s);
    constexpr int PT = 2;
    const dim3 grid((unsigned)((N + EW_THREADS * PT - 1) / (EW_THREADS * PT)), B);
    if (scale && clip) decode_gen_kernel<true, true, PT><<<grid, EW_THREADS, 0, st>>>(gen, d4, var, (int)N, o4);
    else if (scale) decode_gen_kernel<true, false, PT><<<grid, EW_THREADS, 0, st>>>(gen, d4, var, (int)N, o4);
    else if (clip) decode_gen_kernel<false, true, PT><<<grid, EW_THREADS, 0, st>>>(gen, d4, var, (int)N, o4);
    else decode_gen_kernel<false, false, PT><<<grid, EW_THREADS, 0, st>>>(gen, d4, var, (int)N, o4);
    TFRPN_AFTER_LAUNCH("decode_gen_kernel");
    return 0;
}

extern "C" int tfrpn_scale_boxes(const float* boxes, int64_t n_boxes, float height, float width, int denormalize,
                                 float* out, tfrpn_stream s) {
    if (!boxes || !out) return fail(TFRPN_ERR_BAD_ARG, "scale_boxes: null pointer");
    if (n_boxes < 0) return fail(TFRPN_ERR_BAD_ARG, "scale_boxes: negative count");
    if (n_boxes == 0) return 0;
    if (!aligned16(boxes) || !aligned16(out)) return fail(TFRPN_ERR_MISALIGNED, "scale_boxes: 16-byte alignment");
    TFRPN_ENTER_PTR(out, "scale_boxes: out");
    scale_boxes_kernel<<<ew_grid(n_boxes, 1), EW_THREADS, 0, as_stream(s)>>>(
        reinterpret_cast<const float4*>(boxes), n_boxes, height, width, denormalize, reinterpret_cast<float4*>(out));
    TFRPN_AFTER_LAUNCH("scale_boxes_kernel");
    return 0;
}

extern "C" int tfrpn_pad_gt(const float* flat_boxes, const int32_t* flat_labels, const int32_t* offsets,
                            const uint8_t* flip_or_null, int B, int G, int label_add, float* out_boxes,
                            int32_t* out_labels, tfrpn_stream s) {
    if (!offsets || !out_boxes || !out_labels) return fail(TFRPN_ERR_BAD_ARG, "pad_gt: null pointer");
    if (B < 0 || G < 0) return fail(TFRPN_ERR_BAD_ARG, "pad_gt: negative shape");
    if ((long long)B * G == 0) return 0;
    if (!flat_boxes || !flat_labels) return fail(TFRPN_ERR_BAD_ARG, "pad_gt: null pointer");
    if ((long long)B * G > (1LL << 30)) return fail(TFRPN_ERR_UNSUPPORTED, "pad_gt: B*G too large");
    if (!aligned16(flat_boxes) || !aligned16(out_boxes)) return fail(TFRPN_ERR_MISALIGNED, "pad_gt: boxes must be 16-byte aligned");
    TFRPN_ENTER_PTR(out_boxes, "pad_gt: out_boxes");
    const int blocks = (B * G + EW_THREADS - 1) / EW_THREADS;
    pad_gt_kernel<<<blocks, EW_THREADS, 0, as_stream(s)>>>(reinterpret_cast<const float4*>(flat_boxes), flat_labels, offsets,
                                                           flip_or_null, B, G, label_add,
                                                           reinterpret_cast<float4*>(out_boxes), out_labels);
    TFRPN_AFTER_LAUNCH("pad_gt_kernel");
    return 0;
}

extern "C" int tfrpn_selftest_division(uint64_t n_pairs, uint64_t seed, uint64_t* mismatches_dev, tfrpn_stream s) {
    if (!mismatches_dev) return fail(TFRPN_ERR_BAD_ARG, "selftest_division: null pointer");
    TFRPN_ENTER_PTR(mismatches_dev, "selftest_division: mismatches_dev");
    TFRPN_CHECK_CUDA(cudaMemsetAsync(mismatches_dev, 0, sizeof(uint64_t), as_stream(s)));
    if (n_pairs == 0) return 0;
    selftest_division_kernel<<<148 * 8, 256, 0, as_stream(s)>>>(n_pairs, seed, reinterpret_cast<unsigned long long*>(mismatches_dev));
    TFRPN_AFTER_LAUNCH("selftest_division_kernel");
    return 0;
}
