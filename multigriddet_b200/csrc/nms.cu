// Greedy IoU / DIoU NMS, top-k and xyxy conversion for sm_100a.
//
// Replaces DIoUNMS / StandardNMS / ClusterNMS.apply_nms (reference
// multigriddet/postprocess/nms.py:83-231, 320-385: a Python while-loop doing one
// vectorised NumPy pass per kept box), _filter_boxes and _convert_to_xyxy
// (multigrid_decode.py:322-345, 397-422).
//
// One CTA per image, everything after the candidate gather stays in shared memory:
//   0. one thread per candidate rebuilds its box in float64 from the raw logits the
//      decode kernel recorded ((xy + cell) / grid, anchor * exp(wh), letterbox
//      correction with float32 constants: multigrid_decode.py:151-163, 205-228);
//   1. bitonic sort of (score desc, cell index asc) keys -- the deterministic
//      version of the reference's argsort(scores)[::-1];
//   2. the sorted list is walked in chunks of 64.  For each chunk
//        a. every (kept box, chunk member) pair is tested in parallel,
//        b. the 64x64 intra-chunk suppression bitmask is built in parallel,
//        c. one warp resolves the chunk's sequential greedy pass on the 64-bit
//           masks held in registers (ballot + shuffles; one iteration per box it
//           keeps, not per box it looks at);
//      the walk stops as soon as max_boxes boxes are kept (the reference keeps the
//      top max_boxes of the NMS output, which is sorted, :336-345);
//   3. kept boxes are written once: float64 xywh, int32 clipped/rounded xyxy,
//      score, class, cell index.
// The pair metric is the reference's float64 formula evaluated in its operation
// order (no FMA contraction), so `metric < threshold` decides identically.  HBM
// traffic is the candidate records only (48 B each, normally L2-resident).
#include <math.h>
#include "common.cuh"
#include "decode_math.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kChunk = 64;
constexpr int kSortSmem = 2048;   // (key, value) pairs sorted in shared memory

// nms.py:121-148 (IoU) and :189-231 (DIoU); true when b must be suppressed by a,
// i.e. when NOT (metric < threshold)  (nms.py:112,180 keep `metric < threshold`).
__device__ __forceinline__ bool suppresses(const BoxD& a, const BoxD& b, double thr, bool diou)
{
    const double ax2 = __dadd_rn(a.x, a.w), ay2 = __dadd_rn(a.y, a.h);
    const double bx2 = __dadd_rn(b.x, b.w), by2 = __dadd_rn(b.y, b.h);
    const double iw_raw = __dsub_rn(fmin(ax2, bx2), fmax(a.x, b.x));
    const double ih_raw = __dsub_rn(fmin(ay2, by2), fmax(a.y, b.y));
    // disjoint boxes: IoU is exactly 0 and DIoU <= 0, so with a positive threshold
    // the pair can never suppress -- skip the divisions for the common case
    if (thr > 0.0 && (iw_raw <= 0.0 || ih_raw <= 0.0)) return false;
    const double iw = fmax(0.0, iw_raw), ih = fmax(0.0, ih_raw);
    const double inter = __dmul_rn(iw, ih);
    const double uni = __dsub_rn(__dadd_rn(__dmul_rn(a.w, a.h), __dmul_rn(b.w, b.h)), inter);
    double m = __ddiv_rn(inter, __dadd_rn(uni, 1e-8));
    if (diou) {
        const double dxc = __dsub_rn(__dadd_rn(a.x, __ddiv_rn(a.w, 2.0)), __dadd_rn(b.x, __ddiv_rn(b.w, 2.0)));
        const double dyc = __dsub_rn(__dadd_rn(a.y, __ddiv_rn(a.h, 2.0)), __dadd_rn(b.y, __ddiv_rn(b.h, 2.0)));
        const double dist = __dadd_rn(__dmul_rn(dxc, dxc), __dmul_rn(dyc, dyc));
        const double ex = __dsub_rn(fmax(ax2, bx2), fmin(a.x, b.x));
        const double ey = __dsub_rn(fmax(ay2, by2), fmin(a.y, b.y));
        const double diag = __dadd_rn(__dmul_rn(ex, ex), __dmul_rn(ey, ey));
        m = __dsub_rn(m, __ddiv_rn(dist, __dadd_rn(diag, 1e-8)));
    }
    return !(m < thr);
}

// float32 box that contains the float64 box: [x1, y1, x2, y2] rounded outward
__device__ __forceinline__ float4 outer_box(const BoxD& b)
{
    return make_float4(__double2float_rd(b.x), __double2float_rd(b.y),
                       __double2float_ru(__dadd_rn(b.x, b.w)), __double2float_ru(__dadd_rn(b.y, b.h)));
}

// order-preserving map double -> u64, inverted so an ascending sort = descending score
__device__ __forceinline__ unsigned long long score_key(double s)
{
    unsigned long long u = (unsigned long long)__double_as_longlong(s);
    u = (u >> 63) ? ~u : (u | 0x8000000000000000ull);
    return ~u;
}

__device__ __forceinline__ double clipd(double v, double lo, double hi)
{
    return v < lo ? lo : (v > hi ? hi : v);
}

// Approximate pair metric on float32 outer boxes [x1, y1, x2, y2].
// Returns 1 (suppress), 0 (keep) or -1 (undecided: evaluate the exact formula).
// Error bound: every coordinate is within eps = 1.2e-7 * L of the float64 one (L = largest
// coordinate magnitude of the image's boxes).  With iw, ih >= tmin = 1e-3 * L:
//   |dIoU|        <= 6 eps (1/iw + 1/ih)              <= 1.5e-3
//   |d dist/diag| <= 8 eps / sqrt(diag), sqrt(diag) >= tmin  =>  <= 1e-3
// plus a few float32 roundings (1e-6): total < 3e-3, decided only beyond 1e-2.
__device__ __forceinline__ int approx_verdict(const float4& a, const float4& b, float thr, bool diou,
                                              float tmin)
{
    const float iw = fminf(a.z, b.z) - fmaxf(a.x, b.x);
    const float ih = fminf(a.w, b.w) - fmaxf(a.y, b.y);
    if (!(iw >= tmin && ih >= tmin)) return -1;
    const float inter = iw * ih;
    const float uni = ((a.z - a.x) * (a.w - a.y) + (b.z - b.x) * (b.w - b.y)) - inter;
    float m = __fdividef(inter, uni + 1e-8f);
    if (diou) {
        const float dxc = 0.5f * ((a.x + a.z) - (b.x + b.z));
        const float dyc = 0.5f * ((a.y + a.w) - (b.y + b.w));
        const float ex = fmaxf(a.z, b.z) - fminf(a.x, b.x);
        const float ey = fmaxf(a.w, b.w) - fminf(a.y, b.y);
        m -= __fdividef(dxc * dxc + dyc * dyc, (ex * ex + ey * ey) + 1e-8f);
    }
    if (m > thr + 1e-2f) return 1;
    if (m < thr - 1e-2f) return 0;
    return -1;                                   // also every NaN
}

// ---- output stores ----------------------------------------------------------------------
// Every detection leaves the kernels through these helpers: into the caller's tensors and,
// when those live in a detection exchange (mgd_exchange_*, api.cu), into the same tensors
// of every peer rank -- plain stores through the peers' IPC mappings, i.e. straight over
// NVLink, issued as each image's greedy pass produces its rows.  mirror_delta[r] is the byte
// distance from this rank's exchange buffer to its mapping of peer r's buffer, so one
// address computation serves every copy; the exchange guarantees 16-byte alignment, which
// lets the remote copies go out as 16-byte stores (fewer NVLink packets than the scalar
// stores the local tensors, whose alignment is the caller's business, get).  kMirror is a
// template parameter: the single-GPU kernels carry none of this.
template <bool kMirror, class T>
__device__ __forceinline__ void mirror_store(const NmsArgs& a, T* p, const T& v)
{
    if (kMirror) {
        #pragma unroll
        for (int r = 0; r < MGD_MAX_MIRRORS; ++r)
            if (r < a.n_mirrors)
                *reinterpret_cast<T*>(reinterpret_cast<char*>(p) + a.mirror_delta[r]) = v;
    }
}

// multigrid_decode.py:409-420 (_convert_to_xyxy): clip to the image, round half up
template <bool kMirror>
__device__ __forceinline__ void emit_det(const NmsArgs& a, size_t o, double W, double H, double x,
                                         double y, double w, double h, double score, int cls, int index)
{
    if (a.out_xywh) {
        double* p = a.out_xywh + o * 4;
        p[0] = x; p[1] = y; p[2] = w; p[3] = h;
        mirror_store<kMirror>(a, reinterpret_cast<double2*>(p), make_double2(x, y));
        mirror_store<kMirror>(a, reinterpret_cast<double2*>(p) + 1, make_double2(w, h));
    }
    if (a.out_xyxy) {
        int* p = a.out_xyxy + o * 4;
        int4 q;
        q.x = (int)floor(__dadd_rn(clipd(x, 0.0, W), 0.5));
        q.y = (int)floor(__dadd_rn(clipd(y, 0.0, H), 0.5));
        q.z = (int)floor(__dadd_rn(clipd(__dadd_rn(x, w), 0.0, W), 0.5));
        q.w = (int)floor(__dadd_rn(clipd(__dadd_rn(y, h), 0.0, H), 0.5));
        p[0] = q.x; p[1] = q.y; p[2] = q.z; p[3] = q.w;
        mirror_store<kMirror>(a, reinterpret_cast<int4*>(p), q);
    }
    if (a.out_scores) { a.out_scores[o] = score; mirror_store<kMirror>(a, a.out_scores + o, score); }
    if (a.out_classes) { a.out_classes[o] = cls; mirror_store<kMirror>(a, a.out_classes + o, cls); }
    if (a.out_index) { a.out_index[o] = index; mirror_store<kMirror>(a, a.out_index + o, index); }
}

// rows beyond the image's detections: 0 / -1
template <bool kMirror>
__device__ __forceinline__ void emit_pad(const NmsArgs& a, size_t o)
{
    if (a.out_xywh) {
        double* p = a.out_xywh + o * 4;
        for (int e = 0; e < 4; ++e) p[e] = 0.0;
        mirror_store<kMirror>(a, reinterpret_cast<double2*>(p), make_double2(0.0, 0.0));
        mirror_store<kMirror>(a, reinterpret_cast<double2*>(p) + 1, make_double2(0.0, 0.0));
    }
    if (a.out_xyxy) {
        int* p = a.out_xyxy + o * 4;
        for (int e = 0; e < 4; ++e) p[e] = 0;
        mirror_store<kMirror>(a, reinterpret_cast<int4*>(p), make_int4(0, 0, 0, 0));
    }
    if (a.out_scores) { a.out_scores[o] = 0.0; mirror_store<kMirror>(a, a.out_scores + o, 0.0); }
    if (a.out_classes) { a.out_classes[o] = -1; mirror_store<kMirror>(a, a.out_classes + o, -1); }
    if (a.out_index) { a.out_index[o] = -1; mirror_store<kMirror>(a, a.out_index + o, -1); }
}

template <bool kMirror>
__device__ __forceinline__ void emit_count(const NmsArgs& a, int b, int n)
{
    a.out_counts[b] = n;
    mirror_store<kMirror>(a, a.out_counts + b, n);
}

// Top-K window for images with more candidates than the shared-memory sort holds.  Greedy NMS
// visits candidates in descending score and stops at max_boxes kept, so a prefix of the
// order is usually all it ever looks at.  A 1 024-bin histogram of the float32 score bits
// (8 bins per octave) gives the lowest bin T such that the candidates in bins >= T number
// at most kSortSmem; exactly those are gathered (unsorted) into key / val.  Every selected
// candidate outranks every other one, so NMS on the selection is a prefix of the full
// result: if it ends with max_boxes kept (or everything was selected) it IS the result,
// otherwise the caller repeats with all candidates.  Returns the number selected (uniform).
__device__ int select_top_window(const Cand* cand, int M, unsigned long long* key,
                                 unsigned long long* val, int* hist, int* scan, int* s_n, int* s_bin)
{
    const int tid = threadIdx.x;
    for (int i = tid; i < 1024; i += kThreads) hist[i] = 0;
    if (tid == 0) { *s_n = 0; *s_bin = 1024; }
    __syncthreads();
    for (int i = tid; i < M; i += kThreads)
        atomicAdd(&hist[min(__float_as_uint(cand[i].score) >> 20, 1023u)], 1);
    __syncthreads();
    // suffix sums over threads: thread t owns bins [4t, 4t + 4)
    const int mine = hist[4 * tid] + hist[4 * tid + 1] + hist[4 * tid + 2] + hist[4 * tid + 3];
    scan[tid] = mine;
    __syncthreads();
    for (int off = 1; off < kThreads; off <<= 1) {
        const int v = scan[tid] + (tid + off < kThreads ? scan[tid + off] : 0);
        __syncthreads();
        scan[tid] = v;
        __syncthreads();
    }
    const int above = scan[tid] - mine;                  // candidates in bins of higher threads
    if (above <= kSortSmem && scan[tid] > kSortSmem) {   // the boundary lies inside my four bins
        int acc = above, bin = 4 * tid + 4;
        for (int q = 3; q >= 0; --q) {
            if (acc + hist[4 * tid + q] > kSortSmem) break;
            acc += hist[4 * tid + q];
            bin = 4 * tid + q;
        }
        *s_bin = bin;                                    // lowest bin that still fits
    }
    __syncthreads();
    const unsigned lo = (unsigned)*s_bin;
    __syncthreads();                                     // hist (aliases val) is dead from here
    for (int i = tid; i < M; i += kThreads) {
        const Cand cd = cand[i];
        if (min(__float_as_uint(cd.score) >> 20, 1023u) >= lo) {
            const int slot = atomicAdd(s_n, 1);
            key[slot] = score_key((double)cd.score);
            val[slot] = ((unsigned long long)(unsigned)cd.index << 32) | (unsigned)i;
        }
    }
    __syncthreads();
    return *s_n;
}

template <bool kMirror>
__global__ void __launch_bounds__(kThreads, 3)
nms_kernel(const __grid_constant__ NmsArgs a)
{
    __shared__ unsigned long long s_key[kSortSmem];
    __shared__ unsigned long long s_val[kSortSmem];
    __shared__ BoxD c_box[kChunk];
    __shared__ float4 c_out[kChunk];          // float32 outer box [x1, y1, x2, y2] (rounded outward)
    __shared__ int c_cls[kChunk];
    __shared__ int c_pos[kChunk];
    __shared__ int c_alive[kChunk];
    __shared__ unsigned long long c_mask[kChunk];
    __shared__ int c_new[kChunk];
    __shared__ int s_kept, s_new;
    __shared__ unsigned s_scale;              // bits of the largest coordinate magnitude (float)
    extern __shared__ __align__(16) unsigned char dyn[];
    const int b = blockIdx.x;
    unsigned char* kept_mem = a.kept_scratch ? a.kept_scratch + (size_t)b * a.kept_scratch_stride : dyn;
    BoxD* k_box = reinterpret_cast<BoxD*>(kept_mem);                // [max_boxes]
    float4* k_out = reinterpret_cast<float4*>(k_box + a.max_boxes); // [max_boxes]
    int* k_cls = reinterpret_cast<int*>(k_out + a.max_boxes);       // [max_boxes]

    const int tid = threadIdx.x;
    const int M = a.counts[b];
    if (a.skip_small && M <= a.skip_small) return;     // nms_warp_kernel handled this image
    const Cand* cand = a.cand ? a.cand + (size_t)b * a.cap : nullptr;
    const BoxD* boxes = a.cand ? a.boxes + (size_t)b * a.cap
                               : reinterpret_cast<const BoxD*>(a.in_boxes);
    const bool diou = a.use_diou != 0;

    __shared__ uint64_t s_tab[MGD_EXP2F_N];
    __shared__ int s_scan[kThreads];
    __shared__ int s_sel_n, s_sel_bin;
    if (cand) {
        if (tid < MGD_EXP2F_N) s_tab[tid] = mgd_exp2f_tab[tid];
        __syncthreads();
    }
    // decode mode with more candidates than the shared-memory sort holds: top-K window first
    bool window = cand != nullptr && M > kSortSmem;
  pass_again: ;
    // ---- 0. boxes + sort keys, one thread per candidate -----------------------------
    int n_in = M;                                   // candidates taking part in this pass
    if (window)
        n_in = select_top_window(cand, M, s_key, s_val, reinterpret_cast<int*>(s_val), s_scan,
                                 &s_sel_n, &s_sel_bin);
    int mpad = 2;
    while (mpad < n_in) mpad <<= 1;
    unsigned long long* key = s_key;
    unsigned long long* val = s_val;
    if (mpad > kSortSmem) {
        key = a.sort_scratch + (size_t)b * 2 * a.sort_scratch_stride;
        val = key + a.sort_scratch_stride;
    }
    if (!window) {
        for (int i = tid; i < n_in; i += kThreads) {
            const double sc = cand ? (double)cand[i].score : a.in_scores[i];
            const unsigned idx = cand ? (unsigned)cand[i].index : (unsigned)i;
            key[i] = score_key(sc);
            val[i] = ((unsigned long long)idx << 32) | (unsigned)i;
        }
    }
    for (int i = n_in + tid; i < mpad; i += kThreads) { key[i] = ~0ull; val[i] = ~0ull; }
    if (tid == 0) { s_kept = 0; s_new = 0; s_scale = 0u; }
    __syncthreads();
    {
        // boxes of the participating candidates (rebuilt from the raw logits in decode mode)
        // and the coordinate scale for the float32 level of the pair test
        const HeadGeom& g = a.g;
        const int ih = a.image_hw ? a.image_hw[2 * b] : a.in_h;
        const int iw = a.image_hw ? a.image_hw[2 * b + 1] : a.in_w;
        Letterbox lb;
        if (cand) lb = letterbox_consts(g.in_h, g.in_w, ih, iw);
        BoxD* out = cand ? a.boxes + (size_t)b * a.cap : nullptr;
        float sc = 0.f;
        for (int i = tid; i < n_in; i += kThreads) {
            const int pos = (int)(val[i] & 0xffffffffu);
            BoxD bx;
            if (cand) {
                const Cand cd = cand[pos];
                int layer = 0;
                while (layer + 1 < g.L && cd.index >= g.cell_off[layer + 1]) ++layer;
                const int cell = cd.index - g.cell_off[layer];
                const int rr = cell / g.gw[layer], cc = cell - rr * g.gw[layer];
                decode_axis_of(g, cd.t, layer, cd.anchor, rr, cc, &lb, 0, s_tab, bx.x, bx.w);
                decode_axis_of(g, cd.t, layer, cd.anchor, rr, cc, &lb, 1, s_tab, bx.y, bx.h);
                out[pos] = bx;
            } else {
                bx = boxes[pos];
            }
            const float4 o = outer_box(bx);
            sc = fmaxf(sc, fmaxf(fmaxf(fabsf(o.x), fabsf(o.y)), fmaxf(fabsf(o.z), fabsf(o.w))));
        }
        // non-negative floats order like their bit patterns; NaN (0x7fc00000) and inf sort on top
        atomicMax(&s_scale, __float_as_uint(sc));
    }
    // ---- 1. sort ---------------------------------------------------------------
    __syncthreads();
    for (int k = 2; k <= mpad; k <<= 1) {
        for (int jj = k >> 1; jj > 0; jj >>= 1) {
            for (int i = tid; i < mpad; i += kThreads) {
                const int p = i ^ jj;
                if (p > i) {
                    const unsigned long long ka = key[i], kb = key[p], va = val[i], vb = val[p];
                    const bool b_lt_a = kb < ka || (kb == ka && vb < va);
                    const bool up = (i & k) == 0;
                    if (b_lt_a == up) { key[i] = kb; key[p] = ka; val[i] = vb; val[p] = va; }
                }
            }
            __syncthreads();
        }
    }

    // ---- 2. chunked greedy pass ----------------------------------------------------
    // pair test in three levels like nms_warp_kernel: outer-box overlap, approximate float32
    // metric (decides when further than 1e-2 from the threshold), exact float64 formula
    const float scale_f = __uint_as_float(s_scale);
    const float tmin = scale_f < 3.0e38f ? 1e-3f * scale_f : __int_as_float(0x7f800000);
    const float thr_f = (float)a.thr;
    for (int c0 = 0; c0 < n_in; c0 += kChunk) {
        const int kept = s_kept;
        if (kept >= a.max_boxes) break;
        const int n = min(kChunk, n_in - c0);
        if (tid < kChunk) {
            c_mask[tid] = 0ull;
            c_alive[tid] = tid < n;
            if (tid < n) {
                const int pos = (int)(val[c0 + tid] & 0xffffffffu);
                const BoxD bx = boxes[pos];
                c_box[tid] = bx;
                c_out[tid] = outer_box(bx);
                c_cls[tid] = cand ? cand[pos].cls : (a.in_classes ? a.in_classes[pos] : 0);
                c_pos[tid] = pos;
            }
        }
        __syncthreads();
        // Pair tests: thread (member = tid % 64, group = tid / 64); the member's box stays
        // in registers, the other box is a shared-memory broadcast.  A float32 test on
        // outward-rounded boxes discards disjoint pairs (the vast majority) before any
        // float64 arithmetic: disjoint outer boxes imply disjoint boxes, for which the
        // metric is <= 0 < threshold.
        const int mem = tid & (kChunk - 1);
        const int grp = tid / kChunk;
        constexpr int kGroups = kThreads / kChunk;
        const bool pretest = a.thr > 0.0;
        // a. suppression by boxes kept in earlier chunks
        if (mem < n) {
            const BoxD cb = c_box[mem];
            const float4 co = c_out[mem];
            const int ccls = c_cls[mem];
            bool dead = false;
            for (int k = grp; k < kept && !dead; k += kGroups) {
                if (a.per_class && k_cls[k] != ccls) continue;
                const float4 ko = k_out[k];
                if (pretest && (ko.z <= co.x || co.z <= ko.x || ko.w <= co.y || co.w <= ko.y)) continue;
                const int v = approx_verdict(ko, co, thr_f, diou, tmin);
                dead = v < 0 ? suppresses(k_box[k], cb, a.thr, diou) : v != 0;
            }
            if (dead) c_alive[mem] = 0;
        }
        __syncthreads();
        // b. intra-chunk mask: bit j of c_mask[i] <=> i (earlier) suppresses j (later)
        if (mem < n && c_alive[mem]) {
            const BoxD cb = c_box[mem];
            const float4 co = c_out[mem];
            const int ccls = c_cls[mem];
            for (int i = grp; i < mem; i += kGroups) {
                if (!c_alive[i]) continue;
                if (a.per_class && c_cls[i] != ccls) continue;
                const float4 io = c_out[i];
                if (pretest && (io.z <= co.x || co.z <= io.x || io.w <= co.y || co.w <= io.y)) continue;
                const int v = approx_verdict(io, co, thr_f, diou, tmin);
                if (v < 0 ? suppresses(c_box[i], cb, a.thr, diou) : v != 0) atomicOr(&c_mask[i], 1ull << mem);
            }
        }
        __syncthreads();
        // c. sequential resolve on registers, one warp
        if (tid < 32) {
            const unsigned long long m_lo = c_mask[tid], m_hi = c_mask[tid + 32];
            const unsigned lo = __ballot_sync(0xffffffffu, c_alive[tid] != 0);
            const unsigned hi = __ballot_sync(0xffffffffu, c_alive[tid + 32] != 0);
            unsigned long long live = ((unsigned long long)hi << 32) | lo;
            int k = kept, nn = 0;
            while (live && k < a.max_boxes) {
                const int i = __ffsll((long long)live) - 1;
                const unsigned long long src = i < 32 ? m_lo : m_hi;
                const unsigned long long mi = __shfl_sync(0xffffffffu, src, i & 31);
                live &= ~(1ull << i);
                live &= ~mi;
                if (tid == 0) c_new[nn] = i;
                ++nn; ++k;
            }
            if (tid == 0) { s_new = nn; s_kept = k; }
        }
        __syncthreads();
        // 3. emit the boxes kept from this chunk
        if (tid < s_new) {
            const int i = c_new[tid];
            const int slot = kept + tid;
            k_box[slot] = c_box[i];
            k_out[slot] = c_out[i];
            k_cls[slot] = c_cls[i];
            const BoxD cd = c_box[i];
            const int pos = c_pos[i];
            const size_t o = (size_t)b * a.max_boxes + slot;
            const double W = (double)(a.image_hw ? a.image_hw[2 * b + 1] : a.in_w);
            const double H = (double)(a.image_hw ? a.image_hw[2 * b] : a.in_h);
            emit_det<kMirror>(a, o, W, H, cd.x, cd.y, cd.w, cd.h,
                              cand ? (double)cand[pos].score : a.in_scores[pos], c_cls[i],
                              cand ? cand[pos].index : pos);
        }
        __syncthreads();
    }

    // the window ran out before max_boxes were kept: the answer needs candidates outside it
    if (window && s_kept < a.max_boxes && n_in < M) {
        __syncthreads();
        window = false;
        goto pass_again;
    }

    // ---- padding + counts ----------------------------------------------------------
    const int kept = s_kept;
    for (int q = kept + tid; q < a.max_boxes; q += kThreads) {
        const size_t o = (size_t)b * a.max_boxes + q;
        emit_pad<kMirror>(a, o);
    }
    if (tid == 0) {
        emit_count<kMirror>(a, b, kept);
        if (a.stats) {
            atomicAdd(&a.stats[0], (unsigned long long)M);
            atomicAdd(&a.stats[1], (unsigned long long)kept);
        }
    }
}

// ---------------------------------------------------------------------------------
// Warp-per-image variant for the common case (decode mode, <= kWarpCap candidates,
// kept list small enough for shared memory).  Same algorithm as nms_kernel with
// chunks of 32, but everything is warp-synchronous: no CTA barrier, the intra-chunk
// mask is built with ballots and lives in registers, and several images share a CTA.
// Images it does not take (too many candidates) are left to nms_kernel.
//
// The kernel is bound by fixed-latency dependency stalls, so what matters is how many
// warps (= images) an SM holds and how few instructions a pair test costs:
//   * 8.2 KB of shared memory per warp (32-bit score keys, 16-bit positions, float32
//     outer boxes of the kept list; the float64 kept boxes stay in global memory / L1)
//     -> 6 CTAs x 4 warps per SM;
//   * a pair is tested in three levels: outer-box overlap in float32 (32 kept boxes per
//     unrolled sweep, bit positions are compile-time constants); an approximate float32
//     metric that decides whenever it is further than 1e-2 from the threshold (its error
//     is < 3e-3 once both intersection sides exceed 1e-3 of the coordinate scale, see
//     approx_verdict); and the exact float64 formula only for the few undecided pairs.
// ---------------------------------------------------------------------------------
constexpr int kWarpCapLarge = 1024;    // candidates per image handled by one warp
constexpr int kWarpsPerCtaW = 4;

__host__ __device__ inline int nms_kept_pad(int kept_cap) { return (kept_cap + 31) & ~31; }

__host__ __device__ inline size_t nms_warp_key_bytes(int cap, int kept_cap)
{
    // sort keys u32[cap]; dead once the sort is done, so the greedy pass keeps its state in
    // the same bytes: kept outer boxes float4[kpad] | chunk outer boxes float4[32]
    // | kept class int[kpad] | chunk class int[32] | kept position u16[kpad]
    const size_t kpad = (size_t)nms_kept_pad(kept_cap);
    const size_t overlay = kpad * 16 + 32 * 16 + kpad * 4 + 32 * 4 + kpad * 2;
    const size_t keys = (size_t)cap * 4 > overlay ? (size_t)cap * 4 : overlay;
    return (keys + 15) & ~(size_t)15;
}

__host__ __device__ inline size_t nms_warp_bytes(int cap, int kept_cap)
{
    // keys / greedy state | sorted position u16[cap]
    const size_t b = nms_warp_key_bytes(cap, kept_cap) + (size_t)cap * 2;
    return (b + 15) & ~(size_t)15;
}

// bit k set <=> kept box k0 + k may overlap the lane's box (and shares its class in
// per-class mode).  Slots beyond the kept count hold an empty box that never overlaps.
template <bool kPerClass>
__device__ __forceinline__ unsigned overlap_sweep(const float4* k_out, const int* k_cls,
                                                  const float4& co, int ccls, bool pretest)
{
    unsigned m = 0;
    #pragma unroll
    for (int k = 0; k < 32; ++k) {
        const float4 ko = k_out[k];
        bool hit = !(pretest && (ko.z <= co.x || co.z <= ko.x || ko.w <= co.y || co.w <= ko.y));
        if (kPerClass) hit = hit && k_cls[k] == ccls;
        m |= hit ? (1u << k) : 0u;
    }
    return m;
}

template <int kWarpCap, bool kMirror>
__global__ void __launch_bounds__(kWarpsPerCtaW * 32, 7)
nms_warp_kernel(const __grid_constant__ NmsArgs a, int kept_cap, int min_count)
{
    extern __shared__ __align__(16) unsigned char dyn[];
    __shared__ uint64_t s_tab[MGD_EXP2F_N];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x < MGD_EXP2F_N) s_tab[threadIdx.x] = mgd_exp2f_tab[threadIdx.x];
    __syncthreads();

    const int kpad = nms_kept_pad(kept_cap);
    unsigned char* mine = dyn + (size_t)warp * nms_warp_bytes(kWarpCap, kept_cap);
    unsigned* key = reinterpret_cast<unsigned*>(mine);
    float4* k_out = reinterpret_cast<float4*>(mine);         // (overlays the keys after the sort)
    float4* c_out = k_out + kpad;
    int* k_cls = reinterpret_cast<int*>(c_out + 32);
    int* c_cls = k_cls + kpad;
    unsigned short* k_pos = reinterpret_cast<unsigned short*>(c_cls + 32);
    unsigned short* pos_of = reinterpret_cast<unsigned short*>(mine + nms_warp_key_bytes(kWarpCap, kept_cap));

    const HeadGeom& g = a.g;
    const bool diou = a.use_diou != 0;
    const bool pretest = a.thr > 0.0;
    const float thr_f = (float)a.thr;
    const float inf = __int_as_float(0x7f800000);

    // Images are handed out dynamically (per-image cost varies by more than 10x), the
    // expensive ones first: phase 0 takes the images with more than `min_count`
    // candidates, phase 1 the rest, so light images fill the tail of the heavy ones.
    int phase = 0;
    for (;;) {
        int b = 0;
        if (lane == 0) b = atomicAdd(a.next_image + phase, 1);
        b = __shfl_sync(0xffffffffu, b, 0);
        if (b >= a.B) {
            if (phase == 1) break;
            phase = 1;
            continue;
        }
        const int M = a.counts[b];
        if (M > kWarpCap) continue;                      // nms_kernel takes this image
        if ((M > min_count) != (phase == 0)) continue;   // the other phase takes it
        const Cand* cand = a.cand + (size_t)b * a.cap;
        BoxD* boxes = a.boxes + (size_t)b * a.cap;
        const int ih = a.image_hw ? a.image_hw[2 * b] : a.in_h;
        const int iw = a.image_hw ? a.image_hw[2 * b + 1] : a.in_w;

        // ---- 0. boxes + keys, lane-strided ------------------------------------------
        int mpad = 64;                                   // (every lane takes part in every stage)
        while (mpad < M) mpad <<= 1;
        float scale = 0.f;                               // largest coordinate magnitude
        {
            const Letterbox lb = letterbox_consts(g.in_h, g.in_w, ih, iw);
            for (int i = lane; i < mpad; i += 32) {
                if (i < M) {
                    const Cand cd = cand[i];
                    int layer = 0;
                    while (layer + 1 < g.L && cd.index >= g.cell_off[layer + 1]) ++layer;
                    const int cell = cd.index - g.cell_off[layer];
                    const int rr = cell / g.gw[layer], cc = cell - rr * g.gw[layer];
                    BoxD bx;
                    decode_axis_of(g, cd.t, layer, cd.anchor, rr, cc, &lb, 0, s_tab, bx.x, bx.w);
                    decode_axis_of(g, cd.t, layer, cd.anchor, rr, cc, &lb, 1, s_tab, bx.y, bx.h);
                    boxes[i] = bx;
                    const float4 o = outer_box(bx);
                    scale = fmaxf(scale, fmaxf(fmaxf(fabsf(o.x), fabsf(o.y)), fmaxf(fabsf(o.z), fabsf(o.w))));
                    // scores are non-negative floats: their bit patterns order like the values.
                    // Real keys stay below 2^31, every padding key is unique and above them, so
                    // the tie branch of the sort only ever sees genuinely equal scores.
                    key[i] = 0x7fffffffu - (__float_as_uint(cd.score) & 0x7fffffffu);
                    pos_of[i] = (unsigned short)i;
                } else {
                    key[i] = 0x80000000u + (unsigned)i;
                    pos_of[i] = 0xffff;
                }
            }
        }
        #pragma unroll
        for (int o = 16; o > 0; o >>= 1) scale = fmaxf(scale, __shfl_xor_sync(0xffffffffu, scale, o));
        // NaN / infinite coordinates: no pair may be decided approximately
        const float tmin = scale < 3.0e38f ? 1e-3f * scale : inf;
        __syncwarp();
        // ---- 1. bitonic sort (score desc, cell index asc) ----------------------------
        // 32-bit score keys only; the deterministic tie rule (lower cell index first) is
        // applied afterwards to the runs of equal keys, which are rare.
        // Two compare-exchanges per lane and step, all loads issued before any store: the
        // pairs of one stage are disjoint, and a lone warp needs the ILP (mpad >= 64, so
        // both are always in range).
        for (int k = 2; k <= mpad; k <<= 1) {
            for (int jj = k >> 1; jj > 0; jj >>= 1) {
                for (int t = lane; t < (mpad >> 1); t += 64) {
                    // t-th compare-exchange of this stage: i has bit jj clear
                    int i[2], p[2];
                    unsigned ka[2], kb[2];
                    unsigned short pa[2], pb[2];
                    #pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        const int tt = t + 32 * u;
                        i[u] = ((tt & ~(jj - 1)) << 1) | (tt & (jj - 1));
                        p[u] = i[u] | jj;
                        ka[u] = key[i[u]]; kb[u] = key[p[u]];
                        pa[u] = pos_of[i[u]]; pb[u] = pos_of[p[u]];
                    }
                    #pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        const bool up = (i[u] & k) == 0;
                        if ((kb[u] < ka[u]) == up) {
                            key[i[u]] = kb[u]; key[p[u]] = ka[u];
                            pos_of[i[u]] = pb[u]; pos_of[p[u]] = pa[u];
                        }
                    }
                }
                __syncwarp();
            }
        }

        // equal scores (rare): order each run of equal keys by cell index, rank-sort per run
        {
            bool tie = false;
            for (int i = lane; i < M - 1; i += 32) tie = tie || key[i] == key[i + 1];
            if (__any_sync(0xffffffffu, tie)) {
                int s0 = 0;
                while (s0 < M) {                                  // warp-uniform walk over the runs
                    int e0 = s0 + 1;
                    while (e0 < M && key[e0] == key[s0]) ++e0;
                    if (e0 - s0 > 1) {
                        for (int j = s0 + lane; j < e0; j += 32) {
                            const unsigned short pj = pos_of[j];
                            const int ij = cand[pj].index;
                            int rank = 0;
                            for (int i = s0; i < e0; ++i) rank += cand[pos_of[i]].index < ij;
                            key[s0 + rank] = pj;                  // the run's keys are no longer needed
                        }
                        __syncwarp();
                        for (int j = s0 + lane; j < e0; j += 32) pos_of[j] = (unsigned short)key[j];
                        __syncwarp();
                    }
                    s0 = e0;
                }
            }
        }

        // ---- 2. chunked greedy pass, 32 candidates per chunk ---------------------------
        // (the keys are dead: their bytes now hold the kept list, initialised to empty boxes)
        for (int k = lane; k < kpad; k += 32) k_out[k] = make_float4(inf, inf, -inf, -inf);
        __syncwarp();
        int kept = 0;
        const double W = (double)iw, H = (double)ih;
        for (int c0 = 0; c0 < M && kept < a.max_boxes; c0 += 32) {
            const int n = min(32, M - c0);
            const bool have = lane < n;
            int pos = 0, ccls = 0;
            BoxD cb = {0.0, 0.0, 0.0, 0.0};
            float4 co = make_float4(0.f, 0.f, 0.f, 0.f);
            if (have) {
                pos = pos_of[c0 + lane];
                cb = boxes[pos];
                co = outer_box(cb);
                ccls = cand[pos].cls;
            }
            // a. suppression by boxes kept in earlier chunks, 32 kept boxes per sweep: a
            //    uniform float32 sweep records which kept boxes might overlap, then all lanes
            //    evaluate their next pending kept box together until suppressed or done.
            bool alive = have;
            for (int k0 = 0; k0 < kept; k0 += 32) {
                if (!__any_sync(0xffffffffu, alive)) break;
                unsigned pend = a.per_class ? overlap_sweep<true>(k_out + k0, k_cls + k0, co, ccls, pretest)
                                            : overlap_sweep<false>(k_out + k0, k_cls + k0, co, ccls, pretest);
                const int kn = kept - k0;
                if (kn < 32) pend &= (1u << kn) - 1u;
                if (!alive) pend = 0;
                while (__any_sync(0xffffffffu, pend != 0)) {
                    if (pend) {
                        const int k = k0 + __ffs((int)pend) - 1;
                        pend &= pend - 1;
                        int v = approx_verdict(k_out[k], co, thr_f, diou, tmin);
                        if (v < 0) v = suppresses(boxes[k_pos[k]], cb, a.thr, diou) ? 1 : 0;
                        if (v) pend = 0;
                        alive = alive && !v;
                    }
                }
            }
            // b. intra-chunk: lane j collects the set of EARLIER alive members that suppress
            //    it (same levels; only members still alive are visited, their outer boxes
            //    travel by shuffle)
            const unsigned live0 = __ballot_sync(0xffffffffu, alive);
            unsigned pend = 0;
            {
                unsigned rest = live0;
                while (rest) {                                    // warp-uniform
                    const int i = __ffs((int)rest) - 1;
                    rest &= rest - 1;
                    const float ox = __shfl_sync(0xffffffffu, co.x, i), oy = __shfl_sync(0xffffffffu, co.y, i);
                    const float oz = __shfl_sync(0xffffffffu, co.z, i), ow = __shfl_sync(0xffffffffu, co.w, i);
                    const int icls = __shfl_sync(0xffffffffu, ccls, i);
                    if (alive && lane > i && !(a.per_class && icls != ccls) &&
                        !(pretest && (oz <= co.x || co.z <= ox || ow <= co.y || co.w <= oy)))
                        pend |= 1u << i;
                }
            }
            c_out[lane] = co;
            __syncwarp();
            unsigned sup_by = 0;                                  // earlier members suppressing me
            while (__any_sync(0xffffffffu, pend != 0)) {
                if (pend) {
                    const int i = __ffs((int)pend) - 1;
                    pend &= pend - 1;
                    int v = approx_verdict(c_out[i], co, thr_f, diou, tmin);
                    if (v < 0) v = suppresses(boxes[pos_of[c0 + i]], cb, a.thr, diou) ? 1 : 0;
                    if (v) sup_by |= 1u << i;
                }
            }
            // c. sequential resolve: a kept member kills every later member it suppresses
            unsigned live = live0, keep_bits = 0;
            int k = kept;
            while (live && k < a.max_boxes) {
                const int i = __ffs((int)live) - 1;
                const unsigned killed = __ballot_sync(0xffffffffu, (sup_by >> i) & 1u);
                live &= ~(1u << i);
                live &= ~killed;
                keep_bits |= 1u << i;
                ++k;
            }
            // 3. emit
            if ((keep_bits >> lane) & 1u) {
                const int slot = kept + __popc(keep_bits & ((1u << lane) - 1u));
                k_out[slot] = co; k_cls[slot] = ccls; k_pos[slot] = (unsigned short)pos;
                const size_t o = (size_t)b * a.max_boxes + slot;
                emit_det<kMirror>(a, o, W, H, cb.x, cb.y, cb.w, cb.h, (double)cand[pos].score, ccls,
                                  cand[pos].index);
            }
            kept = k;
            __syncwarp();
        }
        // ---- padding + counts -------------------------------------------------------------
        for (int q = kept + lane; q < a.max_boxes; q += 32) {
            const size_t o = (size_t)b * a.max_boxes + q;
            emit_pad<kMirror>(a, o);
        }
        if (lane == 0) {
            emit_count<kMirror>(a, b, kept);
            if (a.stats) {
                atomicAdd(&a.stats[0], (unsigned long long)M);
                atomicAdd(&a.stats[1], (unsigned long long)kept);
            }
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------
// Gaussian SoftNMS (reference multigriddet/postprocess/nms.py:234-288), one CTA per
// image.  The visiting order is fixed by the ORIGINAL scores (the reference never
// re-sorts): candidate i multiplies the score of every later candidate by
// exp(-IoU^2 / sigma); a candidate whose decayed score is below the threshold when
// its turn comes is zeroed and does not decay anyone.  Survivors keep their decayed
// score and are returned in input order (ascending candidate index) -- unless there
// are more than max_boxes, then the top max_boxes by decayed score
// (multigrid_decode.py:336-345).  O(M^2) pair tests, M-1 CTA barriers.
// ---------------------------------------------------------------------------------
__device__ __forceinline__ double iou_xywh(const BoxD& a, const BoxD& b)
{
    const double iw = fmax(0.0, __dsub_rn(fmin(__dadd_rn(a.x, a.w), __dadd_rn(b.x, b.w)), fmax(a.x, b.x)));
    const double ih = fmax(0.0, __dsub_rn(fmin(__dadd_rn(a.y, a.h), __dadd_rn(b.y, b.h)), fmax(a.y, b.y)));
    const double inter = __dmul_rn(iw, ih);
    const double uni = __dsub_rn(__dadd_rn(__dmul_rn(a.w, a.h), __dmul_rn(b.w, b.h)), inter);
    return __ddiv_rn(inter, __dadd_rn(uni, 1e-8));
}

__device__ void cta_bitonic(unsigned long long* key, unsigned long long* val, int mpad, int tid)
{
    for (int k = 2; k <= mpad; k <<= 1) {
        for (int jj = k >> 1; jj > 0; jj >>= 1) {
            for (int i = tid; i < mpad; i += kThreads) {
                const int p = i ^ jj;
                if (p > i) {
                    const unsigned long long ka = key[i], kb = key[p], va = val[i], vb = val[p];
                    const bool b_lt_a = kb < ka || (kb == ka && vb < va);
                    const bool up = (i & k) == 0;
                    if (b_lt_a == up) { key[i] = kb; key[p] = ka; val[i] = vb; val[p] = va; }
                }
            }
            __syncthreads();
        }
    }
}

template <bool kMirror>
__global__ void __launch_bounds__(kThreads, 3)
soft_nms_kernel(const __grid_constant__ NmsArgs a)
{
    __shared__ unsigned long long s_key[kSortSmem];
    __shared__ unsigned long long s_val[kSortSmem];
    __shared__ uint64_t s_tab[MGD_EXP2F_N];
    __shared__ int s_count;
    const int b = blockIdx.x;
    const int tid = threadIdx.x;
    const int M = a.counts[b];
    const Cand* cand = a.cand ? a.cand + (size_t)b * a.cap : nullptr;
    const BoxD* boxes = a.cand ? a.boxes + (size_t)b * a.cap
                               : reinterpret_cast<const BoxD*>(a.in_boxes);
    double* soft = a.soft_scratch + (size_t)b * a.cap;

    int mpad = 2;
    while (mpad < M) mpad <<= 1;
    unsigned long long* key = s_key;
    unsigned long long* val = s_val;
    if (mpad > kSortSmem) {
        key = a.sort_scratch + (size_t)b * 2 * a.sort_scratch_stride;
        val = key + a.sort_scratch_stride;
    }
    if (tid < MGD_EXP2F_N) s_tab[tid] = mgd_exp2f_tab[tid];
    if (tid == 0) s_count = 0;
    __syncthreads();
    if (cand) {
        const HeadGeom& g = a.g;
        const int ih = a.image_hw ? a.image_hw[2 * b] : a.in_h;
        const int iw = a.image_hw ? a.image_hw[2 * b + 1] : a.in_w;
        const Letterbox lb = letterbox_consts(g.in_h, g.in_w, ih, iw);
        BoxD* out = a.boxes + (size_t)b * a.cap;
        for (int i = tid; i < M; i += kThreads) {
            const Cand cd = cand[i];
            int layer = 0;
            while (layer + 1 < g.L && cd.index >= g.cell_off[layer + 1]) ++layer;
            const int cell = cd.index - g.cell_off[layer];
            const int rr = cell / g.gw[layer], cc = cell - rr * g.gw[layer];
            BoxD bx;
            decode_axis_of(g, cd.t, layer, cd.anchor, rr, cc, &lb, 0, s_tab, bx.x, bx.w);
            decode_axis_of(g, cd.t, layer, cd.anchor, rr, cc, &lb, 1, s_tab, bx.y, bx.h);
            out[i] = bx;
        }
    }
    for (int i = tid; i < mpad; i += kThreads) {
        if (i < M) {
            const double sc = cand ? (double)cand[i].score : a.in_scores[i];
            const unsigned idx = cand ? (unsigned)cand[i].index : (unsigned)i;
            soft[i] = sc;
            key[i] = score_key(sc);
            val[i] = ((unsigned long long)idx << 32) | (unsigned)i;
        } else {
            key[i] = ~0ull; val[i] = ~0ull;
        }
    }
    __syncthreads();
    cta_bitonic(key, val, mpad, tid);

    // ---- sequential decay in the original score order ------------------------------------
    for (int i = 0; i < M; ++i) {
        const int cur = (int)(val[i] & 0xffffffffu);
        // every thread reads soft[cur] BEFORE the barrier and thread 0 zeroes it only after:
        // the branch below is then uniform for any threshold (also <= 0) and any score
        const double sc = soft[cur];
        __syncthreads();
        if (sc < a.soft_thr) {                             // nms.py:265-267
            if (tid == 0) soft[cur] = 0.0;
            continue;
        }
        const BoxD cb = boxes[cur];
        for (int j = i + 1 + tid; j < M; j += kThreads) {
            const int p = (int)(val[j] & 0xffffffffu);
            const double iou = iou_xywh(cb, boxes[p]);
            if (iou > 0.0) soft[p] = __dmul_rn(soft[p], exp(__ddiv_rn(-__dmul_rn(iou, iou), a.soft_sigma)));
        }
        __syncthreads();
    }
    __syncthreads();

    // ---- survivors: input order, or the top max_boxes by decayed score -------------------
    int mine = 0;
    for (int i = tid; i < M; i += kThreads) mine += soft[i] >= a.soft_thr;
    if (mine) atomicAdd(&s_count, mine);
    __syncthreads();
    const int K = s_count;
    const bool by_score = K > a.max_boxes;
    for (int i = tid; i < mpad; i += kThreads) {
        if (i < M && soft[i] >= a.soft_thr) {
            const unsigned idx = cand ? (unsigned)cand[i].index : (unsigned)i;
            key[i] = by_score ? score_key(soft[i]) : 0ull;
            val[i] = ((unsigned long long)idx << 32) | (unsigned)i;
        } else {
            key[i] = ~0ull; val[i] = ~0ull;
        }
    }
    __syncthreads();
    cta_bitonic(key, val, mpad, tid);
    const int n_out = min(K, a.max_boxes);
    const double W = (double)(a.image_hw ? a.image_hw[2 * b + 1] : a.in_w);
    const double H = (double)(a.image_hw ? a.image_hw[2 * b] : a.in_h);
    for (int q = tid; q < a.max_boxes; q += kThreads) {
        const size_t o = (size_t)b * a.max_boxes + q;
        if (q < n_out) {
            const int pos = (int)(val[q] & 0xffffffffu);
            const BoxD cd = boxes[pos];
            emit_det<kMirror>(a, o, W, H, cd.x, cd.y, cd.w, cd.h, soft[pos],
                              cand ? cand[pos].cls : (a.in_classes ? a.in_classes[pos] : 0),
                              cand ? cand[pos].index : pos);
        } else {
            emit_pad<kMirror>(a, o);
        }
    }
    if (tid == 0) {
        emit_count<kMirror>(a, b, n_out);
        if (a.stats) {
            atomicAdd(&a.stats[0], (unsigned long long)M);
            atomicAdd(&a.stats[1], (unsigned long long)n_out);
        }
    }
}

// ---------------------------------------------------------------------------------
// Weighted Boxes Fusion as the reference implements it (multigriddet/postprocess/
// wbf.py:98-218): per class, boxes sorted by score; every still unused box becomes a
// cluster leader and absorbs the later unused boxes of its class whose IoU WITH THE
// LEADER is >= iou_thr; a cluster is fused into the (score x weight)-weighted mean box
// and a fused confidence (mean / max / mean of score x weight).  One CTA per image:
// clustering is a leader loop with one barrier per leader, fusion is one thread per
// leader walking its members in the reference's order (deterministic, NumPy's float64
// summation order), output in (class asc, leader score desc) order, or the top
// max_boxes by fused score when there are more (multigrid_decode.py:336-345).
// ---------------------------------------------------------------------------------
__device__ __forceinline__ double wbf_iou(const BoxD& a, const BoxD& b)      // wbf.py:220-250
{
    const double x0 = fmax(a.x, b.x), y0 = fmax(a.y, b.y);
    const double x1 = fmin(__dadd_rn(a.x, a.w), __dadd_rn(b.x, b.w));
    const double y1 = fmin(__dadd_rn(a.y, a.h), __dadd_rn(b.y, b.h));
    if (x1 <= x0 || y1 <= y0) return 0.0;
    const double inter = __dmul_rn(__dsub_rn(x1, x0), __dsub_rn(y1, y0));
    const double uni = __dsub_rn(__dadd_rn(__dmul_rn(a.w, a.h), __dmul_rn(b.w, b.h)), inter);
    return uni > 0.0 ? __ddiv_rn(inter, uni) : 0.0;
}

// NumPy float64 add.reduce order (pairwise, 8 accumulators); get(k) returns the k-th term
template <typename F>
__device__ double np_sum_f64(F get, int first, int n)
{
    if (n < 8) {
        double res = 0.0;
        for (int i = 0; i < n; ++i) res = __dadd_rn(res, get(first + i));
        return res;
    }
    if (n <= 128) {
        double r[8];
        for (int j = 0; j < 8; ++j) r[j] = get(first + j);
        int i = 8;
        for (; i < n - (n & 7); i += 8)
            for (int j = 0; j < 8; ++j) r[j] = __dadd_rn(r[j], get(first + i + j));
        double res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                               __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
        for (; i < n; ++i) res = __dadd_rn(res, get(first + i));
        return res;
    }
    int n2 = n / 2;
    n2 -= n2 & 7;
    return __dadd_rn(np_sum_f64(get, first, n2), np_sum_f64(get, first + n2, n - n2));
}

template <bool kMirror>
__global__ void __launch_bounds__(kThreads)
wbf_kernel(const __grid_constant__ NmsArgs a)
{
    __shared__ uint64_t s_tab[MGD_EXP2F_N];
    __shared__ int s_count, s_clusters;
    const int b = blockIdx.x;
    const int tid = threadIdx.x;
    const int M = a.counts[b];
    const Cand* cand = a.cand ? a.cand + (size_t)b * a.cap : nullptr;
    const BoxD* boxes = a.cand ? a.boxes + (size_t)b * a.cap
                               : reinterpret_cast<const BoxD*>(a.in_boxes);
    // per-image scratch, all indexed by sorted position
    int mpad = 2;
    while (mpad < M) mpad <<= 1;
    unsigned long long* key = a.sort_scratch + (size_t)b * 2 * a.sort_scratch_stride;   // fused-score keys
    unsigned long long* val = key + a.sort_scratch_stride;
    int* order = a.wbf_ints + (size_t)b * 4 * a.cap;     // sorted position -> candidate position
    int* leader_of = order + a.cap;                      // sorted position -> leader's sorted position
    int* member = leader_of + a.cap;                     // cluster member lists (sorted positions)
    int* tmp = member + a.cap;
    double* fused = a.soft_scratch + (size_t)b * a.cap * 5;   // per cluster rank: x y w h score

    if (tid < MGD_EXP2F_N) s_tab[tid] = mgd_exp2f_tab[tid];
    if (tid == 0) { s_count = 0; s_clusters = 0; }
    __syncthreads();
    if (cand) {
        const HeadGeom& g = a.g;
        const int ih = a.image_hw ? a.image_hw[2 * b] : a.in_h;
        const int iw = a.image_hw ? a.image_hw[2 * b + 1] : a.in_w;
        const Letterbox lb = letterbox_consts(g.in_h, g.in_w, ih, iw);
        BoxD* out = a.boxes + (size_t)b * a.cap;
        for (int i = tid; i < M; i += kThreads) {
            const Cand cd = cand[i];
            int layer = 0;
            while (layer + 1 < g.L && cd.index >= g.cell_off[layer + 1]) ++layer;
            const int cell = cd.index - g.cell_off[layer];
            const int rr = cell / g.gw[layer], cc = cell - rr * g.gw[layer];
            BoxD bx;
            decode_axis_of(g, cd.t, layer, cd.anchor, rr, cc, &lb, 0, s_tab, bx.x, bx.w);
            decode_axis_of(g, cd.t, layer, cd.anchor, rr, cc, &lb, 1, s_tab, bx.y, bx.h);
            out[i] = bx;
        }
    }
    auto score_of = [&](int pos) { return cand ? (double)cand[pos].score : a.in_scores[pos]; };
    auto class_of = [&](int pos) { return cand ? cand[pos].cls : (a.in_classes ? a.in_classes[pos] : 0); };
    auto weight_of = [&](int pos) { return a.in_weights ? a.in_weights[pos] : 1.0; };
    auto index_of = [&](int pos) { return cand ? cand[pos].index : pos; };

    // ---- sort positions by (class asc, score desc, index asc); skipped boxes last -----------
    // bitonic network on `order` with a comparator that looks the three keys up
    for (int i = tid; i < mpad; i += kThreads) {
        const bool ok = i < M && score_of(i) >= a.soft_thr;          // wbf.py:74 skip_box_thr
        val[i] = ok ? (unsigned long long)i : ~0ull;
    }
    __syncthreads();
    auto before = [&](unsigned long long x, unsigned long long y) {   // strict order
        if (x == ~0ull || y == ~0ull) return y == ~0ull && x != ~0ull;
        const int px = (int)x, py = (int)y;
        const int cx = class_of(px), cy = class_of(py);
        if (cx != cy) return cx < cy;
        const double sx = score_of(px), sy = score_of(py);
        if (sx != sy) return sx > sy;
        return index_of(px) < index_of(py);
    };
    for (int k = 2; k <= mpad; k <<= 1) {
        for (int jj = k >> 1; jj > 0; jj >>= 1) {
            for (int i = tid; i < mpad; i += kThreads) {
                const int p = i ^ jj;
                if (p > i) {
                    const unsigned long long va = val[i], vb = val[p];
                    const bool up = (i & k) == 0;
                    if (before(vb, va) == up && va != vb) { val[i] = vb; val[p] = va; }
                }
            }
            __syncthreads();
        }
    }
    int n_valid = 0;
    for (int i = tid; i < M; i += kThreads) {
        const bool ok = val[i] != ~0ull;
        n_valid += ok;
        order[i] = ok ? (int)val[i] : -1;
        leader_of[i] = -1;
    }
    if (n_valid) atomicAdd(&s_count, n_valid);
    __syncthreads();
    const int V = s_count;                                     // boxes that take part
    __syncthreads();
    if (tid == 0) s_count = 0;
    __syncthreads();

    // ---- clustering: leader loop (wbf.py:159-183) -------------------------------------------
    for (int i = 0; i < V; ++i) {
        // leader_of[i] is only ever written in an earlier, non-skipped iteration, and each of
        // those ends in a barrier: the test is uniform.  Leaders keep -1 inside the loop (a
        // store here would race with slower warps still reading the flag) and are marked after.
        if (leader_of[i] >= 0) continue;
        const int pi = order[i];
        const BoxD lbx = boxes[pi];
        const int lcls = class_of(pi);
        for (int j = i + 1 + tid; j < V; j += kThreads) {
            if (leader_of[j] >= 0) continue;
            const int pj = order[j];
            if (class_of(pj) != lcls) continue;
            if (wbf_iou(lbx, boxes[pj]) >= a.thr) leader_of[j] = i;
        }
        __syncthreads();
    }
    for (int i = tid; i < V; i += kThreads)
        if (leader_of[i] < 0) leader_of[i] = i;
    __syncthreads();

    // ---- fusion: one thread per leader, members in sorted order (wbf.py:189-213) ---------------
    for (int i = tid; i < V; i += kThreads) {
        if (leader_of[i] != i) continue;
        int n = 0, rank = 0;
        for (int j = 0; j < i; ++j) rank += leader_of[j] == j;     // clusters before this one
        for (int j = i; j < V; ++j) n += leader_of[j] == i;
        const int base = atomicAdd(&s_count, n);
        int w = 0;
        for (int j = i; j < V && w < n; ++j)
            if (leader_of[j] == i) member[base + w++] = j;
        auto sw = [&](int k) { const int p = order[member[k]]; return __dmul_rn(score_of(p), weight_of(p)); };
        const double total = np_sum_f64(sw, base, n);
        auto tw = [&](int k) { return __ddiv_rn(sw(k), total); };
        const double scl = np_sum_f64(tw, base, n);                 // np.average: weights.sum()
        double acc[4] = {0.0, 0.0, 0.0, 0.0};
        for (int k = 0; k < n; ++k) {                               // axis-0 reduction: row by row
            const BoxD bx = boxes[order[member[base + k]]];
            const double t = tw(base + k);
            acc[0] = __dadd_rn(acc[0], __dmul_rn(bx.x, t)); acc[1] = __dadd_rn(acc[1], __dmul_rn(bx.y, t));
            acc[2] = __dadd_rn(acc[2], __dmul_rn(bx.w, t)); acc[3] = __dadd_rn(acc[3], __dmul_rn(bx.h, t));
        }
        double conf;
        if (a.wbf_conf_type == 1) {                                 // 'max'
            conf = score_of(order[member[base]]);
            for (int k = 1; k < n; ++k) conf = fmax(conf, score_of(order[member[base + k]]));
        } else if (a.wbf_conf_type == 2) {                          // mean(score * weight)
            conf = __ddiv_rn(np_sum_f64(sw, base, n), (double)n);
        } else {                                                    // 'avg'
            auto sc = [&](int k) { return score_of(order[member[k]]); };
            conf = __ddiv_rn(np_sum_f64(sc, base, n), (double)n);
        }
        double* f = fused + (size_t)rank * 5;
        f[0] = __ddiv_rn(acc[0], scl); f[1] = __ddiv_rn(acc[1], scl);
        f[2] = __ddiv_rn(acc[2], scl); f[3] = __ddiv_rn(acc[3], scl);
        f[4] = conf;
        tmp[rank] = i;                                              // cluster rank -> leader position
        atomicAdd(&s_clusters, 1);
    }
    __syncthreads();
    const int K = s_clusters;

    // ---- output order ---------------------------------------------------------------------------
    const bool by_score = K > a.max_boxes;
    int kpad = 2;
    while (kpad < K) kpad <<= 1;
    if (by_score) {
        for (int i = tid; i < kpad; i += kThreads) {
            key[i] = i < K ? score_key(fused[(size_t)i * 5 + 4]) : ~0ull;
            val[i] = i < K ? (unsigned long long)i : ~0ull;
        }
        __syncthreads();
        cta_bitonic(key, val, kpad, tid);
    }
    const int n_out = min(K, a.max_boxes);
    const double W = (double)(a.image_hw ? a.image_hw[2 * b + 1] : a.in_w);
    const double H = (double)(a.image_hw ? a.image_hw[2 * b] : a.in_h);
    for (int q = tid; q < a.max_boxes; q += kThreads) {
        const size_t o = (size_t)b * a.max_boxes + q;
        if (q < n_out) {
            const int rank = by_score ? (int)val[q] : q;
            const double* f = fused + (size_t)rank * 5;
            const int lead_pos = order[tmp[rank]];
            emit_det<kMirror>(a, o, W, H, f[0], f[1], f[2], f[3], f[4], class_of(lead_pos), index_of(lead_pos));
        } else {
            emit_pad<kMirror>(a, o);
        }
    }
    if (tid == 0) {
        emit_count<kMirror>(a, b, n_out);
        if (a.stats) {
            atomicAdd(&a.stats[0], (unsigned long long)M);
            atomicAdd(&a.stats[1], (unsigned long long)n_out);
        }
    }
}

__global__ void keep_from_index_kernel(const int* index, const int* counts, int max_keep,
                                       int* keep, int* n_keep)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int k = counts[0];
    if (i == 0) *n_keep = k;
    if (i < max_keep && i < k) keep[i] = index[i];
}

}  // namespace

int nms_smem_capacity() { return kSortSmem; }
size_t nms_kept_bytes(int max_boxes) { return (size_t)max_boxes * (sizeof(BoxD) + sizeof(float4) + sizeof(int)) + 16; }

cudaError_t launch_nms(const NmsArgs& a_in, int num_sms, cudaStream_t stream)
{
    NmsArgs a = a_in;
    cudaError_t err;
    prof_mark_begin(PROF_NMS, stream);
    const bool mirror = a.n_mirrors > 0;
    if (a.soft) {
        if (mirror) soft_nms_kernel<true><<<a.B, kThreads, 0, stream>>>(a);
        else        soft_nms_kernel<false><<<a.B, kThreads, 0, stream>>>(a);
        prof_mark_end(PROF_NMS, stream);
        return cudaGetLastError();
    }
    if (a.wbf) {
        if (mirror) wbf_kernel<true><<<a.B, kThreads, 0, stream>>>(a);
        else        wbf_kernel<false><<<a.B, kThreads, 0, stream>>>(a);
        prof_mark_end(PROF_NMS, stream);
        return cudaGetLastError();
    }
    // decode mode with a kept list that fits shared memory: the warp-per-image kernel takes
    // every image with <= 1024 candidates, nms_kernel the rest
    static int env_off = -1;
    if (env_off < 0) { const char* e = getenv("MGD_NMS_NO_WARP_KERNEL"); env_off = e ? atoi(e) : 0; }
    // A warp per image wins on throughput (4 096 images in one resident wave: 0.54 ms), a CTA
    // per image on latency (103 us against 206 us for one image).  Measured crossover at COCO
    // 608, ~770 candidates per image: 1 536 images 372 vs 389 us, 2 048 images 469 vs 413 us.
    // Small batches (the evaluator's one image per call) therefore take the CTA kernel.
    const char* e_min = getenv("MGD_NMS_WARP_MIN_IMAGES");      // (read per launch: tests force either kernel)
    const int warp_min = e_min ? atoi(e_min) : 1664;
    if (a.cand && nms_warp_bytes(kWarpCapLarge, a.max_boxes) <= 32 * 1024 && !env_off && a.B >= warp_min) {
        auto run = [&](auto kernel, int cap, int min_count) -> cudaError_t {
            static int env_pad = -1;
            if (env_pad < 0) { const char* e = getenv("MGD_NMS_SMEM_PAD"); env_pad = e ? atoi(e) : 0; }
            const size_t dyn = nms_warp_bytes(cap, a.max_boxes) * kWarpsPerCtaW + (size_t)env_pad;
            cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
            if (e != cudaSuccess) return e;
            int ctas_per_sm = (int)((220 * 1024) / (dyn + 1024));
            if (ctas_per_sm > 16) ctas_per_sm = 16;
            static int env_cap = -1;                             // measurements: resident NMS CTAs per SM
            if (env_cap < 0) { const char* e = getenv("MGD_NMS_WARP_CTAS_PER_SM"); env_cap = e ? atoi(e) : 0; }
            const int want_cap = env_cap > 0 ? env_cap : a.warp_ctas_per_sm;
            if (want_cap > 0 && ctas_per_sm > want_cap) ctas_per_sm = want_cap;
            if (ctas_per_sm < 1) ctas_per_sm = 1;
            long long grid = ((long long)a.B + kWarpsPerCtaW - 1) / kWarpsPerCtaW;
            const long long cap_grid = (long long)num_sms * ctas_per_sm;
            if (grid > cap_grid) grid = cap_grid;
            kernel<<<(unsigned)grid, kWarpsPerCtaW * 32, dyn, stream>>>(a, a.max_boxes, min_count);
            return cudaGetLastError();
        };
        err = mirror ? run(nms_warp_kernel<kWarpCapLarge, true>, kWarpCapLarge, kWarpCapLarge / 4)
                     : run(nms_warp_kernel<kWarpCapLarge, false>, kWarpCapLarge, kWarpCapLarge / 4);
        if (err != cudaSuccess) return err;
        prof_mark_end(PROF_NMS, stream);          // one span (= one counted launch) per kernel
        prof_mark_begin(PROF_NMS, stream);
        a.skip_small = kWarpCapLarge;
    }
    const size_t dyn = a.kept_scratch ? 16 : nms_kept_bytes(a.max_boxes);
    err = mirror ? cudaFuncSetAttribute(nms_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn)
                 : cudaFuncSetAttribute(nms_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
    if (err != cudaSuccess) return err;
    if (mirror) nms_kernel<true><<<a.B, kThreads, dyn, stream>>>(a);
    else        nms_kernel<false><<<a.B, kThreads, dyn, stream>>>(a);
    prof_mark_end(PROF_NMS, stream);
    return cudaGetLastError();
}

cudaError_t launch_keep_from_index(const int* index, const int* counts, int max_keep, int* keep,
                                   int* n_keep, cudaStream_t stream)
{
    const int blocks = max_keep > 0 ? (max_keep + 255) / 256 : 1;
    keep_from_index_kernel<<<blocks, 256, 0, stream>>>(index, counts, max_keep, keep, n_keep);
    return cudaGetLastError();
}
