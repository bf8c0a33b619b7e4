// Bit-faithful float32 exp for the decode path.
//
// The reference computes scores with scipy.special.expit / softmax and np.exp in
// float32 (multigrid_decode.py:140-147,162,170).  On glibc hosts those reach
// glibc's expf (scipy always; NumPy whenever its AVX dispatch is off), whose
// algorithm is the public "exp2f-style" kernel from ARM Optimized Routines
// (glibc 2.28+, sysdeps/ieee754/flt-32/e_expf.c): reduce x*N/ln2 = k + r with a
// 32-entry table of 2^(i/32), evaluate a cubic in double, round once to float.
// Restating that algorithm here lets the CUDA kernel produce the SAME float32
// bits as the reference for every score, so the NMS order and the keep set can
// be compared exactly instead of "up to a few ulp".  The table is 2^(i/32) with
// the exponent bits pre-subtracted; tests/test_libm_emul.py checks this function
// against the host libm over a dense sweep of inputs.
#pragma once
#include <stdint.h>
#include <string.h>
#include <math.h>

#if defined(__CUDACC__)
#define MGD_HD __host__ __device__ __forceinline__
#else
#define MGD_HD static inline
#endif

#define MGD_EXP2F_N 32

#if defined(__CUDA_ARCH__)
__constant__
#else
static const
#endif
uint64_t mgd_exp2f_tab[MGD_EXP2F_N] = {
    0x3ff0000000000000ULL, 0x3fefd9b0d3158574ULL, 0x3fefb5586cf9890fULL, 0x3fef9301d0125b51ULL,
    0x3fef72b83c7d517bULL, 0x3fef54873168b9aaULL, 0x3fef387a6e756238ULL, 0x3fef1e9df51fdee1ULL,
    0x3fef06fe0a31b715ULL, 0x3feef1a7373aa9cbULL, 0x3feedea64c123422ULL, 0x3feece086061892dULL,
    0x3feebfdad5362a27ULL, 0x3feeb42b569d4f82ULL, 0x3feeab07dd485429ULL, 0x3feea47eb03a5585ULL,
    0x3feea09e667f3bcdULL, 0x3fee9f75e8ec5f74ULL, 0x3feea11473eb0187ULL, 0x3feea589994cce13ULL,
    0x3feeace5422aa0dbULL, 0x3feeb737b0cdc5e5ULL, 0x3feec49182a3f090ULL, 0x3feed503b23e255dULL,
    0x3feee89f995ad3adULL, 0x3feeff76f2fb5e47ULL, 0x3fef199bdd85529cULL, 0x3fef3720dcef9069ULL,
    0x3fef5818dcfba487ULL, 0x3fef7c97337b9b5fULL, 0x3fefa4afa2a490daULL, 0x3fefd0765b6e4540ULL,
};

MGD_HD double mgd_u64_as_double(uint64_t u)
{
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)u);
#else
    double d; memcpy(&d, &u, 8); return d;
#endif
}

MGD_HD uint64_t mgd_double_as_u64(double d)
{
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(d);
#else
    uint64_t u; memcpy(&u, &d, 8); return u;
#endif
}

// Core of the routine for finite |x| < ~104 (no overflow / NaN handling).
// `tab` lets device code pass a shared-memory copy of the table.
MGD_HD float mgd_expf_core(float x, const uint64_t *tab)
{
    const double inv_ln2_n = 0x1.71547652b82fep+0 * MGD_EXP2F_N;
    const double shift = 0x1.8p+52;
    const double c0 = 0x1.c6af84b912394p-5 / MGD_EXP2F_N / MGD_EXP2F_N / MGD_EXP2F_N;
    const double c1 = 0x1.ebfce50fac4f3p-3 / MGD_EXP2F_N / MGD_EXP2F_N;
    const double c2 = 0x1.62e42ff0c52d6p-1 / MGD_EXP2F_N;
    double z = inv_ln2_n * (double)x;
    double kd = z + shift;                       // round to nearest-even integer
    uint64_t ki = mgd_double_as_u64(kd);
    kd -= shift;
    // r is the exact residual: glibc's FMA build contracts `z - kd` with the product that
    // made z (gcc fuses a multiply into every add that consumes it, also when the product
    // has a second use), so r = fma(InvLn2N, x, -kd), not the difference of the rounded z
    double r = fma(inv_ln2_n, (double)x, -kd);
    uint64_t t = tab[ki % MGD_EXP2F_N];
    t += ki << (52 - 5);
    double s = mgd_u64_as_double(t);
    // glibc selects its FMA build of this routine on every x86-64 CPU with FMA3
    // (sysdeps/x86_64/fpu/multiarch/e_expf.c), where the compiler contracts the three
    // multiply-adds below and the residual above; fused here as well.  With these four
    // contractions the routine equals the host's expf on all 2^32 float inputs
    // (tests/test_libm_emul.py sweeps every one); with the residual left unfused it differs
    // for x = 0x1.04845ep+5 and x = -0x1.f8cbb2p+5.
    double q = fma(c0, r, c1);
    double r2 = r * r;
    double y = fma(c2, r, 1.0);
    y = fma(q, r2, y);
    y = y * s;
    return (float)y;
}

MGD_HD float mgd_expf_tab(float x, const uint64_t *tab)
{
    // |x| >= 88 or NaN: glibc's special-case block
    if (!(x < 88.0f && x > -88.0f)) {
        if (x != x) return x + x;
        if (x > 88.72283172607421875f) return __builtin_huge_valf();   // 0x1.62e42ep6f
        if (x < -103.972076416015625f) return 0.0f;                    // -0x1.9fe368p6f
    }
    return mgd_expf_core(x, tab);
}

MGD_HD float mgd_expf(float x) { return mgd_expf_tab(x, mgd_exp2f_tab); }

// scipy.special.expit for float32: 1 / (1 + expf(-x)), all in float
MGD_HD float mgd_expitf_tab(float x, const uint64_t *tab)
{
    return 1.0f / (1.0f + mgd_expf_tab(-x, tab));
}
