// Device-side arithmetic shared by the decode and NMS kernels: the reference's
// float32 probability / score path and float64 box path, restated operation by
// operation (reference multigriddet/postprocess/multigrid_decode.py:140-170,
// 185-235).  Compiled with -fmad=false; every rounding is explicit.
#pragma once
#include <limits.h>
#include <math.h>
#include "common.cuh"
#include "libm_emul.h"

// ---- NumPy float32 add.reduce order, one octet of lanes = the 8 accumulators --
// x: shared-memory array of n floats; j = lane within the octet (0..7); m = the
// octet's lane mask (octets of one warp may diverge from each other).
static __device__ float np_sum_octet(const float* x, int n, int j, unsigned m)
{
    if (n < 8) {
        float res = 0.f;
        for (int i = 0; i < n; ++i) res = __fadd_rn(res, x[i]);
        return res;
    }
    if (n <= 128) {
        float r = x[j];
        const int body = n - (n & 7);
        for (int i = 8 + j; i < body; i += 8) r = __fadd_rn(r, x[i]);
        r = __fadd_rn(r, __shfl_xor_sync(m, r, 1));   // (r0+r1) (r2+r3) ...
        r = __fadd_rn(r, __shfl_xor_sync(m, r, 2));   // ((r0+r1)+(r2+r3)) ...
        r = __fadd_rn(r, __shfl_xor_sync(m, r, 4));
        for (int i = body; i < n; ++i) r = __fadd_rn(r, x[i]);
        return r;
    }
    int n2 = n / 2;
    n2 -= n2 & 7;
    const float lo = np_sum_octet(x, n2, j, m);
    const float hi = np_sum_octet(x + n2, n - n2, j, m);
    return __fadd_rn(lo, hi);
}

// Max probability and its first index over x[0..n): softmax (scipy: exp(x-max)/sum)
// or element-wise expit.  Overwrites x with the exponentials / probabilities.
// Called by all eight lanes of an octet together.
static __device__ void octet_probs(float* x, int n, int j, unsigned m, bool use_softmax,
                            const uint64_t* tab, float& pmax, int& arg)
{
    if (use_softmax) {
        float mx = -INFINITY;
        for (int i = j; i < n; i += 8) mx = fmaxf(mx, x[i]);
        mx = fmaxf(mx, __shfl_xor_sync(m, mx, 1));
        mx = fmaxf(mx, __shfl_xor_sync(m, mx, 2));
        mx = fmaxf(mx, __shfl_xor_sync(m, mx, 4));
        // arguments are <= 0; below -104 expf is exactly 0, so clamping there keeps the
        // bits and lets the special-case branch of the emulation drop out
        int near_one = INT_MAX;                 // first own index whose exponential is ~1
        #pragma unroll 2
        for (int i = j; i < n; i += 8) {
            // (a select, not fmaxf: NaN -- from a NaN or +inf logit -- must reach the sum like
            //  it does in the reference, which then drops the row: NaN >= confidence is false)
            const float d = __fsub_rn(x[i], mx);
            const float e = mgd_expf_core(d < -104.0f ? -104.0f : d, tab);
            x[i] = e;
            if (e >= 0.99999f && near_one == INT_MAX) near_one = i;
        }
        __syncwarp(m);
        const float s = np_sum_octet(x, n, j, m);
        pmax = __fdiv_rn(1.0f, s);              // the maximum's exponential is exactly 1
        // argmax on the probabilities e/s like the reference: first index whose quotient
        // equals the maximum quotient (only exponentials within 1e-5 of 1 can tie)
        int first = INT_MAX;
        if (near_one != INT_MAX)
            for (int i = near_one; i < n; i += 8) {
                const float e = x[i];
                if (e >= 0.99999f && __fdiv_rn(e, s) == pmax) { first = i; break; }
            }
        first = min(first, __shfl_xor_sync(m, first, 1));
        first = min(first, __shfl_xor_sync(m, first, 2));
        first = min(first, __shfl_xor_sync(m, first, 4));
        arg = first;
    } else {
        float best = -INFINITY;
        int first = INT_MAX;
        for (int i = j; i < n; i += 8) {
            const float q = mgd_expitf_tab(x[i], tab);
            x[i] = q;
            if (q > best) { best = q; first = i; }
        }
        #pragma unroll
        for (int d = 1; d <= 4; d <<= 1) {
            const float ob = __shfl_xor_sync(m, best, d);
            const int oi = __shfl_xor_sync(m, first, d);
            if (ob > best || (ob == best && oi < first)) { best = ob; first = oi; }
        }
        pmax = best;
        arg = first;
    }
}

// e^x by one MUFU.EX2 (flush-to-zero: no range fix-up code).  Only for the conservative
// bounds of the filter levels: a flushed term makes a softmax denominator smaller, i.e. the
// bound larger.
__device__ __forceinline__ float fast_exp(float x)
{
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x * 1.4426950408889634f));
    return y;
}

__device__ __forceinline__ float fast_sigmoid(float x)
{
    return __fdividef(1.0f, 1.0f + fast_exp(-x));
}

// One coordinate pair of a cell's box in float64: multigrid_decode.py:151-163 then
// :219-228.  axis 0 -> (x_min, w), axis 1 -> (y_min, h); the two axes are independent,
// so two lanes evaluate them side by side.  t_xy / t_wh: the raw head outputs of that
// axis, cell: column (axis 0) or row (axis 1), grid / size: the divisors the reference
// uses for that axis (grid_size[axis], input_shape[axis]).
struct Letterbox { float off[2], sc[2], img[2]; };   // index 0: width axis, 1: height axis

__device__ __forceinline__ void decode_axis(float t_xy, float t_wh, int cell, int grid, int size,
                                            float anchor32, double anchor64, bool anchors_f64,
                                            const Letterbox* lb, int axis, const uint64_t* tab,
                                            double& lo, double& extent)
{
    const float u = __fmul_rn(0.15f, t_xy);
    // np.tanh float32: correctly rounded here (libm tanhf is within 2 ulp of this)
    const float act = __fadd_rn((float)tanh((double)u), mgd_expitf_tab(u, tab));
    double c = __ddiv_rn(__dadd_rn((double)act, (double)cell), (double)grid);       // :154-155
    const float e = mgd_expf_tab(t_wh, tab);
    double w;
    if (!anchors_f64)
        w = (double)(float)__ddiv_rn((double)__fmul_rn(anchor32, e), (double)size); // :163 f32 in place
    else
        w = __ddiv_rn(__dmul_rn(anchor64, (double)e), (double)size);
    if (lb) {
        c = __dmul_rn(__dsub_rn(c, (double)lb->off[axis]), (double)lb->sc[axis]);   // :219
        w = __dmul_rn(w, (double)lb->sc[axis]);                                     // :220
        c = __dsub_rn(c, __ddiv_rn(w, 2.0));                                        // :223
        c = __dmul_rn(c, (double)lb->img[axis]);                                    // :227-228
        w = __dmul_rn(w, (double)lb->img[axis]);
    }
    lo = c;
    extent = w;
}

// multigrid_decode.py:205-216 in float32
__device__ __forceinline__ Letterbox letterbox_consts(int in_h, int in_w, int ih_i, int iw_i)
{
    const float mh = (float)in_h, mw = (float)in_w, ih = (float)ih_i, iw = (float)iw_i;
    const float ratio = fminf(__fdiv_rn(mh, ih), __fdiv_rn(mw, iw));
    const float nh = rintf(__fmul_rn(ih, ratio)), nw = rintf(__fmul_rn(iw, ratio));
    Letterbox lb;
    lb.off[1] = __fdiv_rn(__fdiv_rn(__fsub_rn(mh, nh), 2.0f), mh);
    lb.off[0] = __fdiv_rn(__fdiv_rn(__fsub_rn(mw, nw), 2.0f), mw);
    lb.sc[1] = __fdiv_rn(mh, nh);
    lb.sc[0] = __fdiv_rn(mw, nw);
    lb.img[0] = iw;
    lb.img[1] = ih;
    return lb;
}

// axis 0 pairs the column with grid_h / input_h and axis 1 the row with grid_w /
// input_w, exactly like `box_xy /= grid_size` and `box_wh /= input_shape` (:155,163)
__device__ __forceinline__ void decode_axis_of(const HeadGeom& g, const float* x, int layer, int ga,
                                               int rr, int cc, const Letterbox* lb, int axis,
                                               const uint64_t* tab, double& lo, double& extent)
{
    decode_axis(x[axis], x[2 + axis], axis ? rr : cc, axis ? g.gw[layer] : g.gh[layer],
                axis ? g.in_w : g.in_h, g.anc32[ga][axis], g.anc64[ga][axis],
                g.anchors_f64 != 0, lb, axis, tab, lo, extent);
}

// n-th (0-based) set bit of m, or 32 if m has fewer (cheaper than __fns for n < 4)
__device__ __forceinline__ unsigned nth_set_bit(unsigned m, int n)
{
    #pragma unroll
    for (int q = 0; q < 3; ++q)
        if (q < n) m &= m - 1;
    return m ? (unsigned)(__ffs((int)m) - 1) : 32u;
}

