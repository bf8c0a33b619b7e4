// Box-side pre-step of the target encoder for sm_100a: the geometric
// transforms the reference applies to ground-truth boxes on the host before encoding, over
// a batch, so that the (B, N, 5) tensor the encoder consumes is produced on the device.
//
//   reshape_boxes_kernel   reshape_boxes (reference multigriddet/data/augmentation.py:112-164):
//       scale into the padded image, add the paste offset, optional horizontal / vertical
//       flip, clip to the target, drop boxes whose width or height is <= 1.
//   mosaic_merge_kernel    merge_mosaic_bboxes (:606-667): the boxes of four source images
//       are cut at (crop_x, crop_y) according to their quadrant, boxes left too small are
//       dropped, the survivors are concatenated in (quadrant, row) order and capped.
//
// One warp per output image; survivors keep their order (ballot + prefix popcount), rows
// beyond the count are zero.  float64 arithmetic in the reference's operation order
// (-fmad=false); int32 boxes are truncated toward zero where NumPy stores a float result
// into the caller's int32 array (the legacy loader's dtype, generators.py:2429).
#include <math.h>
#include "common.cuh"

namespace {

constexpr int kBoxWarps = 4;

template <typename T> __device__ __forceinline__ double box_store(double v);
template <> __device__ __forceinline__ double box_store<double>(double v) { return v; }
template <> __device__ __forceinline__ double box_store<int>(double v)
{
    // float64 -> int32 assignment: C cast, truncation toward zero (values are far from 2^31)
    return (double)(int)fmin(fmax(v, -2147483648.0), 2147483647.0);
}

template <typename T>
__global__ void __launch_bounds__(kBoxWarps * 32)
reshape_boxes_kernel(const __grid_constant__ BoxOpArgs a)
{
    const int lane = threadIdx.x & 31;
    const int b = blockIdx.x * kBoxWarps + (threadIdx.x >> 5);
    if (b >= a.B) return;
    const T* in = reinterpret_cast<const T*>(a.in) + (size_t)b * a.N * 5;
    T* out = reinterpret_cast<T*>(a.out) + (size_t)b * a.N * 5;
    float* out32 = a.out_f32 ? a.out_f32 + (size_t)b * a.N * 5 : nullptr;
    const int* p = a.params + (size_t)b * 10;
    const double src_w = p[0], src_h = p[1], tgt_w = p[2], tgt_h = p[3];
    const double pad_w = p[4], pad_h = p[5], dx = p[6], dy = p[7];
    const bool hflip = p[8] != 0, vflip = p[9] != 0;
    const int n = min(max(a.counts ? a.counts[b] : a.N, 0), a.N);
    int kept = 0;
    for (int i0 = 0; i0 < n; i0 += 32) {
        const int i = i0 + lane;
        bool keep = false;
        double x1 = 0, y1 = 0, x2 = 0, y2 = 0, cls = 0;
        if (i < n) {
            x1 = (double)in[i * 5 + 0]; y1 = (double)in[i * 5 + 1];
            x2 = (double)in[i * 5 + 2]; y2 = (double)in[i * 5 + 3]; cls = (double)in[i * 5 + 4];
            // :149-150  boxes*padding/src + d, stored back into the caller's dtype
            x1 = box_store<T>(__dadd_rn(__ddiv_rn(__dmul_rn(x1, pad_w), src_w), dx));
            x2 = box_store<T>(__dadd_rn(__ddiv_rn(__dmul_rn(x2, pad_w), src_w), dx));
            y1 = box_store<T>(__dadd_rn(__ddiv_rn(__dmul_rn(y1, pad_h), src_h), dy));
            y2 = box_store<T>(__dadd_rn(__ddiv_rn(__dmul_rn(y2, pad_h), src_h), dy));
            if (hflip) { const double t = __dsub_rn(tgt_w, x2); x2 = __dsub_rn(tgt_w, x1); x1 = t; }   // :153
            if (vflip) { const double t = __dsub_rn(tgt_h, y2); y2 = __dsub_rn(tgt_h, y1); y1 = t; }   // :156
            if (x1 < 0.0) x1 = 0.0;                                                                // :159-161
            if (y1 < 0.0) y1 = 0.0;
            if (x2 > tgt_w) x2 = tgt_w;
            if (y2 > tgt_h) y2 = tgt_h;
            keep = __dsub_rn(x2, x1) > 1.0 && __dsub_rn(y2, y1) > 1.0;                              // :164-166
        }
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        if (keep) {
            const int o = kept + __popc(m & ((1u << lane) - 1u));
            const double v[5] = {x1, y1, x2, y2, cls};
            #pragma unroll
            for (int e = 0; e < 5; ++e) {
                out[o * 5 + e] = (T)v[e];
                if (out32) out32[o * 5 + e] = (float)(T)v[e];
            }
        }
        kept += __popc(m);
    }
    for (int i = kept * 5 + lane; i < a.N * 5; i += 32) {
        out[i] = (T)0;
        if (out32) out32[i] = 0.f;
    }
    if (lane == 0 && a.out_counts) a.out_counts[b] = kept;
}

__global__ void __launch_bounds__(kBoxWarps * 32)
mosaic_merge_kernel(const __grid_constant__ BoxOpArgs a)
{
    const int lane = threadIdx.x & 31;
    const int b = blockIdx.x * kBoxWarps + (threadIdx.x >> 5);
    if (b >= a.B) return;
    const double* src = reinterpret_cast<const double*>(a.in);
    double* out = reinterpret_cast<double*>(a.out) + (size_t)b * a.N * 5;
    float* out32 = a.out_f32 ? a.out_f32 + (size_t)b * a.N * 5 : nullptr;
    const int* p = a.params + (size_t)b * 6;          // sample 0..3, crop_x, crop_y
    const double cx = p[4], cy = p[5];
    const double min_w = fmax(10.0, __dmul_rn((double)a.width, 0.01));     // :656
    const double min_h = fmax(10.0, __dmul_rn((double)a.height, 0.01));
    int kept = 0;
    for (int q = 0; q < 4 && kept < a.N; ++q) {
        const int s = p[q];
        const double* in = src + (size_t)s * a.N * 5;
        for (int i0 = 0; i0 < a.N && kept < a.N; i0 += 32) {
            const int i = i0 + lane;
            bool keep = false;
            double x1 = 0, y1 = 0, x2 = 0, y2 = 0, cls = 0;
            if (i < a.N && s >= 0 && s < a.n_src) {
                x1 = in[i * 5 + 0]; y1 = in[i * 5 + 1]; x2 = in[i * 5 + 2]; y2 = in[i * 5 + 3]; cls = in[i * 5 + 4];
                bool skip;
                const bool cut_y = y2 > cy && y1 < cy, cut_x = x2 > cx && x1 < cx;
                if (q == 0) {            // top-left (:624-630)
                    skip = y1 > cy || x1 > cx;
                    if (cut_y) y2 = cy;
                    if (cut_x) x2 = cx;
                } else if (q == 1) {     // bottom-left (:632-638)
                    skip = y2 < cy || x1 > cx;
                    if (cut_y) y1 = cy;
                    if (cut_x) x2 = cx;
                } else if (q == 2) {     // bottom-right (:640-646)
                    skip = y2 < cy || x2 < cx;
                    if (cut_y) y1 = cy;
                    if (cut_x) x1 = cx;
                } else {                 // top-right (:648-654)
                    skip = y1 > cy || x2 < cx;
                    if (cut_y) y2 = cy;
                    if (cut_x) x1 = cx;
                }
                keep = !skip && !(fabs(__dsub_rn(x2, x1)) < min_w || fabs(__dsub_rn(y2, y1)) < min_h);
            }
            const unsigned m = __ballot_sync(0xffffffffu, keep);
            const int o = kept + __popc(m & ((1u << lane) - 1u));
            if (keep && o < a.N) {                    // :662-663 cap at max_boxes
                const double v[5] = {x1, y1, x2, y2, cls};
                #pragma unroll
                for (int e = 0; e < 5; ++e) {
                    out[o * 5 + e] = v[e];
                    if (out32) out32[o * 5 + e] = (float)v[e];
                }
            }
            kept = min(kept + __popc(m), a.N);
        }
    }
    for (int i = kept * 5 + lane; i < a.N * 5; i += 32) {
        out[i] = 0.0;
        if (out32) out32[i] = 0.f;
    }
    if (lane == 0 && a.out_counts) a.out_counts[b] = kept;
}

// tf.data box pre-step (reference multigriddet/data/generators.py:1859-1916 letterbox / multi-scale
// transform, :227-256 horizontal flip, :1963-1976 padded_batch to max_boxes_per_image,
// :1983-2034 _expand_box_capacity): one warp per image, float32 arithmetic in the order of the
// TensorFlow ops (every op is an IEEE float32 multiply / divide / add; tf.cast(float -> int32)
// truncates).  Output rows beyond the image's count are zero, up to the expanded capacity.
//   params (B, 6) int32: src_h, src_w, scale_h, scale_w (multi-scale target shape, 0 = none),
//                        hflip, reserved
__global__ void __launch_bounds__(kBoxWarps * 32)
letterbox_boxes_kernel(const float* __restrict__ in, const int* __restrict__ counts,
                       const int* __restrict__ params, int B, int n_in, int n_keep, int capacity,
                       int input_h, int input_w, float* __restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const int b = blockIdx.x * kBoxWarps + (threadIdx.x >> 5);
    if (b >= B) return;
    const int* p = params + (size_t)b * 6;
    const float src_h = (float)p[0], src_w = (float)p[1];
    const float th = (float)input_h, tw = (float)input_w;
    float sx, sy, pad_left, pad_top;
    if (p[2] > 0 && p[3] > 0) {                                   // multi-scale branch, :1866-1897
        const float scale_h = __fdiv_rn((float)p[2], th), scale_w = __fdiv_rn((float)p[3], tw);
        const int scaled_h = (int)__fmul_rn(src_h, scale_h), scaled_w = (int)__fmul_rn(src_w, scale_w);
        const float fh = (float)scaled_h, fw = (float)scaled_w;
        const float ls0 = fminf(__fdiv_rn(tw, fw), __fdiv_rn(th, fh));       // tf_letterbox_resize :186
        const int new_w = (int)__fmul_rn(fw, ls0), new_h = (int)__fmul_rn(fh, ls0);
        const float ls = fminf(__fdiv_rn((float)new_w, fw), __fdiv_rn((float)new_h, fh));   // :1887-1890
        sx = __fmul_rn(scale_w, ls);
        sy = __fmul_rn(scale_h, ls);
        pad_left = (float)((input_w - new_w) / 2);                // // on non-negative ints
        pad_top = (float)((input_h - new_h) / 2);
    } else {                                                      // :1898-1916
        const float s0 = fminf(__fdiv_rn(tw, src_w), __fdiv_rn(th, src_h));
        const int new_w = (int)__fmul_rn(src_w, s0), new_h = (int)__fmul_rn(src_h, s0);
        const float sc = fminf(__fdiv_rn((float)new_w, src_w), __fdiv_rn((float)new_h, src_h));
        sx = sy = sc;
        pad_left = (float)((input_w - new_w) / 2);
        pad_top = (float)((input_h - new_h) / 2);
    }
    const bool flip = p[4] != 0;
    int n = counts ? counts[b] : n_in;
    n = min(max(n, 0), min(n_in, n_keep));
    const float* src = in + (size_t)b * n_in * 5;
    float* dst = out + (size_t)b * capacity * 5;
    for (int i = lane; i < capacity; i += 32) {
        float v[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
        if (i < n) {
            float x1 = __fadd_rn(__fmul_rn(src[i * 5 + 0], sx), pad_left);
            const float y1 = __fadd_rn(__fmul_rn(src[i * 5 + 1], sy), pad_top);
            float x2 = __fadd_rn(__fmul_rn(src[i * 5 + 2], sx), pad_left);
            const float y2 = __fadd_rn(__fmul_rn(src[i * 5 + 3], sy), pad_top);
            if (flip) { const float t = __fsub_rn(tw, x2); x2 = __fsub_rn(tw, x1); x1 = t; }   // :248-251
            v[0] = x1; v[1] = y1; v[2] = x2; v[3] = y2;
            v[4] = __fadd_rn(__fmul_rn(src[i * 5 + 4], 1.0f), 0.0f);
        }
        #pragma unroll
        for (int e = 0; e < 5; ++e) dst[i * 5 + e] = v[e];
    }
}

}  // namespace

cudaError_t launch_letterbox_boxes(const float* in, const int* counts, const int* params, int B,
                                   int n_in, int n_keep, int capacity, int input_h, int input_w,
                                   float* out, cudaStream_t stream)
{
    if (B <= 0) return cudaSuccess;
    const unsigned grid = (unsigned)((B + kBoxWarps - 1) / kBoxWarps);
    prof_mark_begin(PROF_OTHER, stream);
    letterbox_boxes_kernel<<<grid, kBoxWarps * 32, 0, stream>>>(in, counts, params, B, n_in, n_keep,
                                                                capacity, input_h, input_w, out);
    prof_mark_end(PROF_OTHER, stream);
    return cudaGetLastError();
}

cudaError_t launch_reshape_boxes(const BoxOpArgs& a, int boxes_i32, cudaStream_t stream)
{
    if (a.B <= 0) return cudaSuccess;
    const unsigned grid = (unsigned)((a.B + kBoxWarps - 1) / kBoxWarps);
    prof_mark_begin(PROF_OTHER, stream);
    if (boxes_i32) reshape_boxes_kernel<int><<<grid, kBoxWarps * 32, 0, stream>>>(a);
    else           reshape_boxes_kernel<double><<<grid, kBoxWarps * 32, 0, stream>>>(a);
    prof_mark_end(PROF_OTHER, stream);
    return cudaGetLastError();
}

cudaError_t launch_mosaic_merge(const BoxOpArgs& a, cudaStream_t stream)
{
    if (a.B <= 0) return cudaSuccess;
    const unsigned grid = (unsigned)((a.B + kBoxWarps - 1) / kBoxWarps);
    prof_mark_begin(PROF_OTHER, stream);
    mosaic_merge_kernel<<<grid, kBoxWarps * 32, 0, stream>>>(a);
    prof_mark_end(PROF_OTHER, stream);
    return cudaGetLastError();
}
