// Loss-side ignore mask for sm_100a.
//
// Replaces MultiGridLoss._compute_ignore_mask and _compute_iou_batch (reference
// multigriddet/losses/multigrid_loss.py:494-703, 445-492): per layer, every cell's predicted
// box for every anchor is compared with all ground-truth boxes of the image on that layer
// (the decoded positive cells of y_true); cells whose best IoU exceeds the threshold without
// being positive are ignored by the objectness / anchor losses.  All three outputs are
// stop_gradient / cast-from-bool in the reference, so this is a forward-only op.
//
// TensorFlow materialises a (G*G*A) x (positives) IoU matrix per image.  Here:
//   (mgd_encode_ignore_mask feeds both kernels from the encoder's owner table and box records
//   instead of a dense y_true: 4 bytes per cell instead of 352, and y_true need not exist.)
//   gt_gather_kernel    one CTA per (image, layer): decode the positive cells into corner
//       boxes, drop exact duplicates (the nine cells of one object decode to the same box
//       when the encoder's fractions are dyadic), write a compact list.  The maximum over the
//       list is unchanged by that, the pair count drops ~9x.
//   ignore_mask_kernel  a thread per cell, the image's list staged through shared memory in
//       chunks (broadcast reads); per anchor the running best pair is kept as a fraction
//       (inter, union + eps) compared by cross-multiplication, one division at the end.
// float32 arithmetic in the reference's operation order (-fmad=false).  The reference's
// quirks are kept: the grid offset of tensor position [row i, col j] is (x = i, y = j)
// (tf.meshgrid(..., indexing='ij'), :547) and anchors are multiplied by the stride (:570,599).
#include <math.h>
#include "common.cuh"

namespace {

constexpr int kGatherThreads = 256;
constexpr int kMaskThreads = 128;
constexpr int kGtChunk = 1024;            // ground-truth boxes staged in shared memory at a time

struct __align__(16) GtBox { float x1, y1, x2, y2; };

__global__ void __launch_bounds__(kGatherThreads)
gt_gather_kernel(const __grid_constant__ LossArgs a)
{
    __shared__ int s_n, s_kept;
    const HeadGeom& g = a.g;
    const int b = blockIdx.x, layer = blockIdx.y;
    const int gh = g.gh[layer], gw = g.gw[layer], D = g.D[layer], A = g.na[layer];
    const int cells = gh * gw;
    const float* yt = a.table ? nullptr : a.y_true[layer] + (size_t)b * cells * D;
    // scratch of this (image, layer): raw list, then the de-duplicated one
    GtBox* raw = reinterpret_cast<GtBox*>(a.gt_boxes) + ((size_t)b * g.cells + g.cell_off[layer]) * 2;
    float* raw_area = a.gt_area + ((size_t)b * g.cells + g.cell_off[layer]) * 2;
    GtBox* out = raw + cells;
    float* out_area = raw_area + cells;
    const float scale_w = __fdiv_rn((float)g.in_w, (float)gw), scale_h = __fdiv_rn((float)g.in_h, (float)gh);
    if (threadIdx.x == 0) { s_n = 0; s_kept = 0; }
    __syncthreads();
    const int* codes = a.table ? a.table + (size_t)a.B * g.cell_off[layer] + (size_t)b * cells : nullptr;
    for (int c = threadIdx.x; c < cells; c += kGatherThreads) {
        float row[4];
        int k = 0;
        if (codes) {
            // the y_true row the writer would store for this cell (encode.cu: encode_fill_kernel)
            const int code = codes[c];
            if (code < 0) continue;
            const BoxRec* rec = a.recs + (code >> 4);
            const int nb = code & 15;
            row[0] = (float)__dadd_rn((double)(1 - nb / 3), rec->fx);
            row[1] = (float)__dadd_rn((double)(1 - nb % 3), rec->fy);
            row[2] = rec->tw;
            row[3] = rec->th;
            k = rec->hot_anchor - 5;
        } else {
            const float* r = yt + (size_t)c * D;
            if (!(r[4] > 0.5f)) continue;                                 // :305 object_mask
            row[0] = r[0]; row[1] = r[1]; row[2] = r[2]; row[3] = r[3];
            float best = r[5];                                              // :562 argmax, first maximum
            for (int q = 1; q < A; ++q) if (r[5 + q] > best) { best = r[5 + q]; k = q; }
        }
        const int i = c / gw, j = c - i * gw;                              // tensor position [row i, col j]
        const int ga = g.anchor_first[layer] + k;
        const float cx = __fmul_rn(__fadd_rn(row[0], (float)i), scale_w);   // :558 with the 'ij' grid
        const float cy = __fmul_rn(__fadd_rn(row[1], (float)j), scale_h);
        const float w = __fmul_rn(__fmul_rn(expf(row[2]), g.anc32[ga][0]), scale_w);   // :570
        const float h = __fmul_rn(__fmul_rn(expf(row[3]), g.anc32[ga][1]), scale_h);
        const float hw = __fdiv_rn(w, 2.0f), hh = __fdiv_rn(h, 2.0f);
        const int pos = atomicAdd(&s_n, 1);
        raw[pos] = GtBox{__fsub_rn(cx, hw), __fsub_rn(cy, hh), __fadd_rn(cx, hw), __fadd_rn(cy, hh)};
        raw_area[pos] = __fmul_rn(w, h);
    }
    __syncthreads();
    const int n = s_n;
    // exact duplicates: keep the first occurrence (the maximum over the list is unchanged)
    for (int p = threadIdx.x; p < n; p += kGatherThreads) {
        const GtBox me = raw[p];
        const float ar = raw_area[p];
        bool dup = false;
        for (int q = 0; q < p && !dup; ++q) {
            const GtBox o = raw[q];
            dup = o.x1 == me.x1 && o.y1 == me.y1 && o.x2 == me.x2 && o.y2 == me.y2 && raw_area[q] == ar;
        }
        if (!dup) {
            const int pos = atomicAdd(&s_kept, 1);
            out[pos] = me;
            out_area[pos] = ar;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) a.gt_count[(size_t)b * g.L + layer] = s_kept;
}

__global__ void __launch_bounds__(kMaskThreads)
ignore_mask_kernel(const __grid_constant__ LossArgs a, int layer)
{
    __shared__ GtBox s_box[kGtChunk];
    __shared__ float s_area[kGtChunk];
    const HeadGeom& g = a.g;
    const int b = blockIdx.y;
    const int gh = g.gh[layer], gw = g.gw[layer], D = g.D[layer], A = g.na[layer];
    const int cells = gh * gw;
    const int c = blockIdx.x * kMaskThreads + threadIdx.x;
    const bool live = c < cells;
    const float scale_w = __fdiv_rn((float)g.in_w, (float)gw), scale_h = __fdiv_rn((float)g.in_h, (float)gh);
    const GtBox* gt = reinterpret_cast<const GtBox*>(a.gt_boxes) + ((size_t)b * g.cells + g.cell_off[layer]) * 2 + cells;
    const float* gt_area = a.gt_area + ((size_t)b * g.cells + g.cell_off[layer]) * 2 + cells;
    const int n_gt = a.gt_count[(size_t)b * g.L + layer];

    float cx = 0.f, cy = 0.f, ew = 0.f, eh = 0.f, obj = 0.f;
    int assigned = 0;
    if (live) {
        const float* yp = a.y_pred[layer] + ((size_t)b * cells + c) * D;
        const int i = c / gw, j = c - i * gw;
        const float ux = __fmul_rn(0.15f, yp[0]), uy = __fmul_rn(0.15f, yp[1]);
        const float ax = __fadd_rn(tanhf(ux), __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-ux))));   // :582
        const float ay = __fadd_rn(tanhf(uy), __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-uy))));
        cx = __fmul_rn(__fadd_rn(ax, (float)i), scale_w);                                       // :585
        cy = __fmul_rn(__fadd_rn(ay, (float)j), scale_h);
        ew = expf(yp[2]); eh = expf(yp[3]);
        if (a.table) {
            const int code = a.table[(size_t)a.B * g.cell_off[layer] + (size_t)b * cells + c];
            obj = code >= 0 ? 1.0f : 0.0f;
            assigned = code >= 0 ? a.recs[code >> 4].hot_anchor - 5 : 0;   // all-zero row: argmax 0
        } else {
            const float* yt = a.y_true[layer] + ((size_t)b * cells + c) * D;
            obj = yt[4] > 0.5f ? 1.0f : 0.0f;
            float best = yt[5];
            for (int q = 1; q < A; ++q) if (yt[5 + q] > best) { best = yt[5 + q]; assigned = q; }
        }
    }
    float iou_max = 0.f, iou_assigned = 0.f;
    for (int an = 0; an < A; ++an) {
        const int ga = g.anchor_first[layer] + an;
        const float w = __fmul_rn(__fmul_rn(ew, g.anc32[ga][0]), scale_w);                      // :599
        const float h = __fmul_rn(__fmul_rn(eh, g.anc32[ga][1]), scale_h);
        const float hw = __fdiv_rn(w, 2.0f), hh = __fdiv_rn(h, 2.0f);
        const float x1 = __fsub_rn(cx, hw), y1 = __fsub_rn(cy, hh), x2 = __fadd_rn(cx, hw), y2 = __fadd_rn(cy, hh);
        const float area = __fmul_rn(w, h);
        float bn = 0.f, bd = 1.f;                       // best pair so far as a fraction bn / bd
        for (int g0 = 0; g0 < n_gt; g0 += kGtChunk) {
            const int m = min(kGtChunk, n_gt - g0);
            __syncthreads();
            for (int t = threadIdx.x; t < m; t += kMaskThreads) { s_box[t] = gt[g0 + t]; s_area[t] = gt_area[g0 + t]; }
            __syncthreads();
            if (live) {
                #pragma unroll 4
                for (int t = 0; t < m; ++t) {
                    const GtBox o = s_box[t];
                    const float iw = fmaxf(__fsub_rn(fminf(x2, o.x2), fmaxf(x1, o.x1)), 0.f);   // :473-475
                    const float ih = fmaxf(__fsub_rn(fminf(y2, o.y2), fmaxf(y1, o.y1)), 0.f);
                    const float inter = __fmul_rn(iw, ih);
                    const float den = __fadd_rn(__fsub_rn(__fadd_rn(area, s_area[t]), inter), a.eps);   // :487-490
                    if (__fmul_rn(inter, bd) > __fmul_rn(bn, den)) { bn = inter; bd = den; }
                }
            }
        }
        const float iou = n_gt > 0 ? __fdiv_rn(bn, bd) : 0.f;                                    // :632 / :635
        iou_max = an == 0 ? iou : fmaxf(iou_max, iou);                                            // :664
        if (an == assigned) iou_assigned = iou;
    }
    if (live) {
        const size_t o = (size_t)b * cells + c;
        a.ignore[layer][o] = (iou_max > a.ignore_thresh && obj < 0.5f) ? 1.0f : 0.0f;           // :684-688
        a.assigned[layer][o] = __fmul_rn(iou_assigned, obj);                                       // :692-693
        a.max_iou[layer][o] = iou_max;
    }
}

}  // namespace

cudaError_t launch_ignore_mask(const LossArgs& a, cudaStream_t stream)
{
    const HeadGeom& g = a.g;
    if (a.B <= 0) return cudaSuccess;
    prof_mark_begin(PROF_OTHER, stream);
    gt_gather_kernel<<<dim3((unsigned)a.B, (unsigned)g.L), kGatherThreads, 0, stream>>>(a);
    cudaError_t err = cudaGetLastError();
    for (int l = 0; l < g.L && err == cudaSuccess; ++l) {
        const int cells = g.gh[l] * g.gw[l];
        ignore_mask_kernel<<<dim3((unsigned)((cells + kMaskThreads - 1) / kMaskThreads), (unsigned)a.B),
                             kMaskThreads, 0, stream>>>(a, l);
        err = cudaGetLastError();
    }
    prof_mark_end(PROF_OTHER, stream);
    return err;
}
