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

__global__ void __launch_bounds__(kThreads)
nms_kernel(const __grid_constant__ NmsArgs a)
{
    __shared__ unsigned long long s_key[kSortSmem];
    __shared__ unsigned long long s_val[kSortSmem];
    __shared__ BoxD c_box[kChunk];
    __shared__ int c_cls[kChunk];
    __shared__ int c_pos[kChunk];
    __shared__ int c_alive[kChunk];
    __shared__ unsigned long long c_mask[kChunk];
    __shared__ int c_new[kChunk];
    __shared__ int s_kept, s_new;
    extern __shared__ __align__(16) unsigned char dyn[];
    const int b = blockIdx.x;
    unsigned char* kept_mem = a.kept_scratch ? a.kept_scratch + (size_t)b * a.kept_scratch_stride : dyn;
    BoxD* k_box = reinterpret_cast<BoxD*>(kept_mem);                // [max_boxes]
    int* k_cls = reinterpret_cast<int*>(k_box + a.max_boxes);       // [max_boxes]

    const int tid = threadIdx.x;
    const int M = a.counts[b];
    const Cand* cand = a.cand ? a.cand + (size_t)b * a.cap : nullptr;
    const BoxD* boxes = a.cand ? a.boxes + (size_t)b * a.cap
                               : reinterpret_cast<const BoxD*>(a.in_boxes);
    const bool diou = a.use_diou != 0;

    // ---- 0. boxes + sort keys, one thread per candidate -----------------------------
    int mpad = 2;
    while (mpad < M) mpad <<= 1;
    unsigned long long* key = s_key;
    unsigned long long* val = s_val;
    if (mpad > kSortSmem) {
        key = a.sort_scratch + (size_t)b * 2 * a.sort_scratch_stride;
        val = key + a.sort_scratch_stride;
    }
    if (cand) {
        __shared__ uint64_t s_tab[MGD_EXP2F_N];
        if (tid < MGD_EXP2F_N) s_tab[tid] = mgd_exp2f_tab[tid];
        __syncthreads();
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
            key[i] = score_key(sc);
            val[i] = ((unsigned long long)idx << 32) | (unsigned)i;
        } else {
            key[i] = ~0ull; val[i] = ~0ull;
        }
    }
    // ---- 1. sort ---------------------------------------------------------------
    if (tid == 0) { s_kept = 0; s_new = 0; }
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
    for (int c0 = 0; c0 < M; c0 += kChunk) {
        const int kept = s_kept;
        if (kept >= a.max_boxes) break;
        const int n = min(kChunk, M - c0);
        if (tid < kChunk) {
            c_mask[tid] = 0ull;
            c_alive[tid] = tid < n;
            if (tid < n) {
                const int pos = (int)(val[c0 + tid] & 0xffffffffu);
                c_box[tid] = boxes[pos];
                c_cls[tid] = cand ? cand[pos].cls : (a.in_classes ? a.in_classes[pos] : 0);
                c_pos[tid] = pos;
            }
        }
        __syncthreads();
        // a. suppression by boxes kept in earlier chunks
        for (int p = tid; p < n * kept; p += kThreads) {
            const int c = p % n, k = p / n;
            if (!c_alive[c]) continue;
            if (a.per_class && k_cls[k] != c_cls[c]) continue;
            if (suppresses(k_box[k], c_box[c], a.thr, diou)) c_alive[c] = 0;
        }
        __syncthreads();
        // b. intra-chunk mask: bit j of c_mask[i] <=> i (earlier) suppresses j (later)
        for (int p = tid; p < n * n; p += kThreads) {
            const int i = p / n, j = p % n;
            if (j <= i || !c_alive[i] || !c_alive[j]) continue;
            if (a.per_class && c_cls[i] != c_cls[j]) continue;
            if (suppresses(c_box[i], c_box[j], a.thr, diou)) atomicOr(&c_mask[i], 1ull << j);
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
            k_cls[slot] = c_cls[i];
            const BoxD cd = c_box[i];
            const int pos = c_pos[i];
            const size_t o = (size_t)b * a.max_boxes + slot;
            if (a.out_xywh) {
                a.out_xywh[o * 4 + 0] = cd.x; a.out_xywh[o * 4 + 1] = cd.y;
                a.out_xywh[o * 4 + 2] = cd.w; a.out_xywh[o * 4 + 3] = cd.h;
            }
            if (a.out_xyxy) {
                const double W = (double)(a.image_hw ? a.image_hw[2 * b + 1] : a.in_w);
                const double H = (double)(a.image_hw ? a.image_hw[2 * b] : a.in_h);
                a.out_xyxy[o * 4 + 0] = (int)floor(__dadd_rn(clipd(cd.x, 0.0, W), 0.5));
                a.out_xyxy[o * 4 + 1] = (int)floor(__dadd_rn(clipd(cd.y, 0.0, H), 0.5));
                a.out_xyxy[o * 4 + 2] = (int)floor(__dadd_rn(clipd(__dadd_rn(cd.x, cd.w), 0.0, W), 0.5));
                a.out_xyxy[o * 4 + 3] = (int)floor(__dadd_rn(clipd(__dadd_rn(cd.y, cd.h), 0.0, H), 0.5));
            }
            if (a.out_scores) a.out_scores[o] = cand ? (double)cand[pos].score : a.in_scores[pos];
            if (a.out_classes) a.out_classes[o] = c_cls[i];
            if (a.out_index) a.out_index[o] = cand ? cand[pos].index : pos;
        }
        __syncthreads();
    }

    // ---- padding + counts ----------------------------------------------------------
    const int kept = s_kept;
    for (int q = kept + tid; q < a.max_boxes; q += kThreads) {
        const size_t o = (size_t)b * a.max_boxes + q;
        if (a.out_xywh) for (int e = 0; e < 4; ++e) a.out_xywh[o * 4 + e] = 0.0;
        if (a.out_xyxy) for (int e = 0; e < 4; ++e) a.out_xyxy[o * 4 + e] = 0;
        if (a.out_scores) a.out_scores[o] = 0.0;
        if (a.out_classes) a.out_classes[o] = -1;
        if (a.out_index) a.out_index[o] = -1;
    }
    if (tid == 0) {
        a.out_counts[b] = kept;
        if (a.stats) {
            atomicAdd(&a.stats[0], (unsigned long long)M);
            atomicAdd(&a.stats[1], (unsigned long long)kept);
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
size_t nms_kept_bytes(int max_boxes) { return (size_t)max_boxes * (sizeof(BoxD) + sizeof(int)) + 16; }

cudaError_t launch_nms(const NmsArgs& a, int, cudaStream_t stream)
{
    const size_t dyn = a.kept_scratch ? 16 : nms_kept_bytes(a.max_boxes);
    cudaError_t err = cudaFuncSetAttribute(nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)dyn);
    if (err != cudaSuccess) return err;
    prof_mark_begin(PROF_NMS, stream);
    nms_kernel<<<a.B, kThreads, dyn, stream>>>(a);
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
