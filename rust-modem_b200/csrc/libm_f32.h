/*
 * libm_f32.h -- binary32 sin/cos (evaluated in IEEE binary64) and atan/atan2 for host AND device, bit-identical to
 * glibc 2.39.
 *
 * Third-party algorithms restated here (not code of the reference, which has none of this):
 *   - sincosf: the algorithm and constants of ARM "optimized-routines" math/sincosf.c, sincosf.h, sincosf_data.c, which
 *     glibc ships as sysdeps/ieee754/flt-32/s_sincosf.c.  Copyright (c) 2018-2019 Arm Limited; SPDX-License-Identifier:
 *     MIT (optimized-routines) / LGPL-2.1-or-later as distributed in glibc.  The MIT notice: "Permission is hereby
 *     granted, free of charge, to any person obtaining a copy of this software and associated documentation files ...
 *     THE SOFTWARE IS PROVIDED "AS IS", WITHOUT WARRANTY OF ANY KIND".
 *   - atanf / atan2f: the fdlibm routines s_atanf.c / e_atan2f.c (conversion to float by Ian Lance Taylor, Cygnus
 *     Support).  "Copyright (C) 1993 by Sun Microsystems, Inc. All rights reserved.  Developed at SunPro, a Sun
 *     Microsystems, Inc. business.  Permission to use, copy, modify, and distribute this software is freely granted,
 *     provided that this notice is preserved."
 * (Round 1 also carried glibc's logf for the Box-Muller stage; the AWGN extension is now defined as explicit binary32
 * operations -- oracle/modem_oracle.h -- and needs no library logarithm.)
 *
 * Why this exists: the reference computes its NCO with Rust `f32::sin_cos`
 * (/root/reference/src/modem/modulator.rs:46) and `f32::cos` / `f32::sin`
 * (/root/reference/src/modem/demodulator.rs:53-54), which lower to the platform
 * libm `sinf` / `cosf`.  CUDA's `sinf/cosf` are 1-2 ULP routines and are NOT
 * bit-identical to that.  B200 has a full-rate FP64 pipe (half the FP32 rate), so
 * we evaluate the same published algorithm the platform libm uses -- the
 * "optimized-routines" single-precision sincosf (range reduction by pi/2 in
 * binary64, two short binary64 polynomials, one final rounding to binary32) -- with
 * the same constants and the same operation order.  Every operation is a correctly
 * rounded IEEE binary64 op on both sides, so the result is bit-identical to glibc
 * 2.39 `sinf/cosf` (verified exhaustively on the host by tools/check_libm.c;
 * the same header is what the kernels compile).
 *
 * MG_LIBM_CONTRACT selects whether `a*b+c` is one fused op or two rounded ops.
 * x86-64 glibc dispatches (ifunc) to an FMA build of these routines on every CPU
 * that has FMA, so the default is 1.
 */
#ifndef MODEM_GPU_LIBM_F32_H
#define MODEM_GPU_LIBM_F32_H

#include <stdint.h>

#ifndef MG_LIBM_CONTRACT
#define MG_LIBM_CONTRACT 1
#endif

#if defined(__CUDACC__)
#define MG_HD __host__ __device__ __forceinline__
#else
#define MG_HD static inline
#endif

#if defined(__CUDA_ARCH__)
#define MG_DMUL(a, b) __dmul_rn((a), (b))
#define MG_DADD(a, b) __dadd_rn((a), (b))
#if MG_LIBM_CONTRACT
#define MG_DFMA(a, b, c) __fma_rn((a), (b), (c))
#else
#define MG_DFMA(a, b, c) __dadd_rn(__dmul_rn((a), (b)), (c))
#endif
#define MG_ASUINT(f) __float_as_uint(f)
#define MG_ASFLOAT(u) __uint_as_float(u)
#else
#include <string.h>
#include <math.h>
static inline double mg_host_mul(double a, double b) { volatile double r = a * b; return r; }
#define MG_DMUL(a, b) ((a) * (b))
#define MG_DADD(a, b) ((a) + (b))
#if MG_LIBM_CONTRACT
#define MG_DFMA(a, b, c) __builtin_fma((a), (b), (c))
#else
/* the volatile round-trip forbids the host compiler from fusing */
#define MG_DFMA(a, b, c) (mg_host_mul((a), (b)) + (c))
#endif
static inline uint32_t mg_asuint_host(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float mg_asfloat_host(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
#define MG_ASUINT(f) mg_asuint_host(f)
#define MG_ASFLOAT(u) mg_asfloat_host(u)
#endif

/* ---- constants (hex floats: exact) ------------------------------------------- */
#define MG_HPI_INV_2P24 0x1.45F306DC9C883p+23 /* 2/pi * 2^24 */
#define MG_HPI 0x1.921FB54442D18p0            /* pi/2 */
#define MG_C0 0x1p0
#define MG_C1 (-0x1.ffffffd0c621cp-2)
#define MG_C2 0x1.55553e1068f19p-5
#define MG_C3 (-0x1.6c087e89a359dp-10)
#define MG_C4 0x1.99343027bf8c3p-16
#define MG_S1 (-0x1.555545995a603p-3)
#define MG_S2 0x1.1107605230bc4p-7
#define MG_S3 (-0x1.994eb3774cf24p-13)

MG_HD uint32_t mg_abstop12(float x) { return (MG_ASUINT(x) >> 20) & 0x7ff; }

/*
 * Both polynomials on reduced argument x (|x| <= pi/4), x2 = x*x.
 * `neg` selects the sign-flipped cosine coefficient set (quadrants 2,3).
 */
MG_HD void mg_sincos_poly(double x, double x2, int neg, int n, float* sinp, float* cosp)
{
    const double c0 = neg ? -MG_C0 : MG_C0;
    const double c1 = neg ? -MG_C1 : MG_C1;
    const double c2k = neg ? -MG_C2 : MG_C2;
    const double c3 = neg ? -MG_C3 : MG_C3;
    const double c4 = neg ? -MG_C4 : MG_C4;

    double x4 = MG_DMUL(x2, x2);
    double x3 = MG_DMUL(x2, x);
    double c2 = MG_DFMA(x2, c4, c3);
    double s1 = MG_DFMA(x2, MG_S3, MG_S2);
    double c1v = MG_DFMA(x2, c1, c0);
    double x5 = MG_DMUL(x3, x2);
    double x6 = MG_DMUL(x4, x2);
    double s = MG_DFMA(x3, MG_S1, x);
    double c = MG_DFMA(x4, c2k, c1v);
    float sv = (float)MG_DFMA(x5, s1, s);
    float cv = (float)MG_DFMA(x6, c2, c);
    if (n & 1) {
        *cosp = sv;
        *sinp = cv;
    } else {
        *sinp = sv;
        *cosp = cv;
    }
}

/* 4/pi as overlapping 32-bit words, for |y| >= 120 (Payne-Hanek style). */
#define MG_INV_PIO4_INIT \
    {0xa2,       0xa2f9,     0xa2f983,   0xa2f9836e, 0xf9836e4e, 0x836e4e44, 0x6e4e4415, 0x4e441529, \
     0x441529fc, 0x1529fc27, 0x29fc2757, 0xfc2757d1, 0x2757d1f5, 0x57d1f534, 0xd1f534dd, 0xf534ddc0, \
     0x34ddc0db, 0xddc0db62, 0xc0db6295, 0xdb629599, 0x6295993c, 0x95993c43, 0x993c4390, 0x3c439041}
static const uint32_t mg_inv_pio4[24] = MG_INV_PIO4_INIT;
#if defined(__CUDACC__)
/* device copies live in global memory (L1-cached): the index differs per lane, which
 * would serialise a __constant__ access */
static __device__ const uint32_t mg_inv_pio4_dev[24] = MG_INV_PIO4_INIT;
#endif
#if defined(__CUDA_ARCH__)
#define MG_INV_PIO4 mg_inv_pio4_dev
#else
#define MG_INV_PIO4 mg_inv_pio4
#endif

MG_HD double mg_reduce_large(uint32_t xi, int* np)
{
    const uint32_t* arr = &MG_INV_PIO4[(xi >> 26) & 15];
    int shift = (xi >> 23) & 7;
    uint64_t n, res0, res1, res2;

    xi = (xi & 0xffffff) | 0x800000;
    xi <<= shift;

    res0 = (uint64_t)xi * arr[0];
    res1 = (uint64_t)xi * arr[4];
    res2 = (uint64_t)xi * arr[8];
    res0 = (res2 >> 32) | (res0 << 32);
    res0 += res1;

    n = (res0 + (1ULL << 61)) >> 62;
    res0 -= n << 62;
    double x = (double)(int64_t)res0;
    *np = (int)n;
    return MG_DMUL(x, 0x1.921FB54442D18p-62);
}

/* sinf and cosf of y in one go; same values as separate sinf(y), cosf(y). */
MG_HD void mg_sincosf(float y, float* sinp, float* cosp)
{
    double x = (double)y;
    int n;
    uint32_t top = mg_abstop12(y);

    if (top < 0x3f4) { /* |y| < pi/4  (abstop12(0x1.921FB6p-1f) == 0x3f4) */
        double x2 = MG_DMUL(x, x);
        if (top < 0x398) { /* |y| < 2^-12 */
            *sinp = y;
            *cosp = 1.0f;
            return;
        }
        mg_sincos_poly(x, x2, 0, 0, sinp, cosp);
    } else if (top < 0x42f) { /* |y| < 120 */
        double r = MG_DMUL(x, MG_HPI_INV_2P24);
        n = ((int32_t)r + 0x800000) >> 24;
        x = MG_DFMA(-(double)n, MG_HPI, x);
        double s = ((n & 3) == 1 || (n & 3) == 2) ? -1.0 : 1.0;
        mg_sincos_poly(MG_DMUL(x, s), MG_DMUL(x, x), (n & 2) != 0, n, sinp, cosp);
    } else if (top < 0x7f8) {
        uint32_t xi = MG_ASUINT(y);
        int sign = (int)(xi >> 31);
        x = mg_reduce_large(xi, &n);
        int q = (n + sign) & 3;
        double s = (q == 1 || q == 2) ? -1.0 : 1.0;
        mg_sincos_poly(MG_DMUL(x, s), MG_DMUL(x, x), (q & 2) != 0, n, sinp, cosp);
    } else {
        *sinp = *cosp = y - y;
    }
}

/* sinf / cosf for 0 <= y <= 7 without the magnitude dispatch: the pi/2 reduction branch of the routine above also
 * reproduces glibc for |y| < pi/4 (n = 0) and for |y| < 2^-12 (the polynomials round to y and 1.0f there) --
 * tools/check_libm.c sweeps every binary32 in [0, 7]: 0 mismatches.  Used where the argument is known to lie in
 * (0, 2 pi] (Box-Muller's angle): no branches, one copy of the polynomial in the instruction stream. */
MG_HD void mg_sincosf_0_7(float y, float* sinp, float* cosp)
{
    double x = (double)y;
    double r = MG_DMUL(x, MG_HPI_INV_2P24);
    int n = ((int32_t)r + 0x800000) >> 24;
    x = MG_DFMA(-(double)n, MG_HPI, x);
    double s = ((n & 3) == 1 || (n & 3) == 2) ? -1.0 : 1.0;
    mg_sincos_poly(MG_DMUL(x, s), MG_DMUL(x, x), (n & 2) != 0, n, sinp, cosp);
}

/* The same routine for any |y| < 120 (both signs): glibc's |y| < 120 branch is this very formula, and for |y| < pi/4 it gives
 * n = 0 and therefore the operands of the small-argument branch -- tools/check_libm.c sweeps every binary32 with |y| < 120
 * against this machine's sinf / cosf: 0 mismatches (with the signed-zero select below).  Used where the angle is phase(n) + a PLL offset: consecutive samples
 * of a warp land in different magnitude classes, so the dispatch of mg_sincosf made every warp run every branch. */
MG_HD void mg_sincosf_lt120(float y, float* sinp, float* cosp)
{
    mg_sincosf_0_7(y, sinp, cosp);
    if (y == 0.0f) *sinp = y; /* sinf(-0) = -0: the one input of the sweep the plain formula gets wrong (it returns +0) */
}
/* cos and sin of an NCO angle plus offset: branch-free for |y| < 120, the general routine beyond */
MG_HD void mg_sincosf_nco(float y, float* sinp, float* cosp)
{
    if (mg_abstop12(y) < 0x42f) mg_sincosf_lt120(y, sinp, cosp);
    else mg_sincosf(y, sinp, cosp);
}

/* ---- atanf / atan2f ---------------------------------------------------------
 * The reference's PLL takes `(x * carrier.conj()).arg()` (pll.rs:19), i.e. num::Complex::arg =
 * im.atan2(re) -> libm atan2f.  glibc 2.39 still ships the classic fdlibm binary32 routines for
 * these two (sysdeps/ieee754/flt-32/{s_atanf,e_atan2f}.c; the correctly-rounded replacements
 * arrived in 2.41): argument reduction to one of four intervals, an 11-term odd/even split
 * polynomial, every step a separately rounded binary32 operation (the generic x86-64 build has no
 * FMA contraction and these two have no ifunc variant).  Restated here with the published
 * constants; tools/check_libm.c compares against this machine's glibc over every binary32.
 */
#if defined(__CUDA_ARCH__)
#define MG_FMUL(a, b) __fmul_rn((a), (b))
#define MG_FADD(a, b) __fadd_rn((a), (b))
#define MG_FSUB(a, b) __fsub_rn((a), (b))
#define MG_FDIV(a, b) __fdiv_rn((a), (b))
#else
static inline float mg_host_fmul(float a, float b) { volatile float r = a * b; return r; }
#define MG_FMUL(a, b) mg_host_fmul((a), (b)) /* the volatile round-trip forbids host-side fusing */
#define MG_FADD(a, b) ((a) + (b))
#define MG_FSUB(a, b) ((a) - (b))
#define MG_FDIV(a, b) ((a) / (b))
#endif

MG_HD float mg_atanf(float x)
{
    const float hi[4] = {4.6364760399e-01f, 7.8539812565e-01f, 9.8279368877e-01f, 1.5707962513e+00f};
    const float lo[4] = {5.0121582440e-09f, 3.7748947079e-08f, 3.4473217170e-08f, 7.5497894159e-08f};
    const float aT0 = 3.3333334327e-01f, aT1 = -2.0000000298e-01f, aT2 = 1.4285714924e-01f, aT3 = -1.1111110449e-01f,
                aT4 = 9.0908870101e-02f, aT5 = -7.6918758452e-02f, aT6 = 6.6610731184e-02f, aT7 = -5.8335702866e-02f,
                aT8 = 4.9768779427e-02f, aT9 = -3.6531571299e-02f, aT10 = 1.6285819933e-02f;
    const uint32_t hx = MG_ASUINT(x), ix = hx & 0x7fffffffu;
    int id;
    if (ix >= 0x4c000000u) { /* |x| >= 2^25 */
        if (ix > 0x7f800000u) return MG_FADD(x, x);
        const float r = MG_FADD(hi[3], lo[3]);
        return (hx >> 31) ? -r : r;
    }
    if (ix < 0x3ee00000u) { /* |x| < 0.4375 */
        if (ix < 0x31000000u) return x; /* |x| < 2^-29 */
        id = -1;
    } else {
        x = MG_ASFLOAT(ix);
        if (ix < 0x3f980000u) { /* |x| < 1.1875 */
            if (ix < 0x3f300000u) { /* 7/16 <= |x| < 11/16 */
                id = 0;
                x = MG_FDIV(MG_FSUB(MG_FMUL(2.0f, x), 1.0f), MG_FADD(2.0f, x));
            } else {
                id = 1;
                x = MG_FDIV(MG_FSUB(x, 1.0f), MG_FADD(x, 1.0f));
            }
        } else if (ix < 0x401c0000u) { /* |x| < 2.4375 */
            id = 2;
            x = MG_FDIV(MG_FSUB(x, 1.5f), MG_FADD(1.0f, MG_FMUL(1.5f, x)));
        } else {
            id = 3;
            x = MG_FDIV(-1.0f, x);
        }
    }
    const float z = MG_FMUL(x, x);
    const float w = MG_FMUL(z, z);
    float s1 = MG_FADD(aT8, MG_FMUL(w, aT10));
    s1 = MG_FADD(aT6, MG_FMUL(w, s1));
    s1 = MG_FADD(aT4, MG_FMUL(w, s1));
    s1 = MG_FADD(aT2, MG_FMUL(w, s1));
    s1 = MG_FADD(aT0, MG_FMUL(w, s1));
    s1 = MG_FMUL(z, s1);
    float s2 = MG_FADD(aT7, MG_FMUL(w, aT9));
    s2 = MG_FADD(aT5, MG_FMUL(w, s2));
    s2 = MG_FADD(aT3, MG_FMUL(w, s2));
    s2 = MG_FADD(aT1, MG_FMUL(w, s2));
    s2 = MG_FMUL(w, s2);
    const float t = MG_FMUL(x, MG_FADD(s1, s2));
    if (id < 0) return MG_FSUB(x, t);
    const float r = MG_FSUB(hi[id], MG_FSUB(MG_FSUB(t, lo[id]), x));
    return (hx >> 31) ? -r : r;
}

MG_HD float mg_atan2f(float y, float x)
{
    const float tiny = 1.0e-30f, pi_o_4 = 7.8539818525e-01f, pi_o_2 = 1.5707963705e+00f, pi = 3.1415927410e+00f,
                pi_lo = -8.7422776573e-08f;
    const uint32_t hx = MG_ASUINT(x), hy = MG_ASUINT(y);
    const uint32_t ix = hx & 0x7fffffffu, iy = hy & 0x7fffffffu;
    if (ix > 0x7f800000u || iy > 0x7f800000u) return MG_FADD(x, y);
    if (hx == 0x3f800000u) return mg_atanf(y);
    const uint32_t m = (hy >> 31) | ((hx >> 30) & 2u); /* 2*sign(x) + sign(y) */
    if (iy == 0) {
        if (m < 2) return y;
        return m == 2 ? MG_FADD(pi, tiny) : MG_FSUB(-pi, tiny);
    }
    if (ix == 0) return (hy >> 31) ? MG_FSUB(-pi_o_2, tiny) : MG_FADD(pi_o_2, tiny);
    if (ix == 0x7f800000u) {
        if (iy == 0x7f800000u) {
            switch (m) {
            case 0: return MG_FADD(pi_o_4, tiny);
            case 1: return MG_FSUB(-pi_o_4, tiny);
            case 2: return MG_FADD(MG_FMUL(3.0f, pi_o_4), tiny);
            default: return MG_FSUB(MG_FMUL(-3.0f, pi_o_4), tiny);
            }
        }
        switch (m) {
        case 0: return 0.0f;
        case 1: return -0.0f;
        case 2: return MG_FADD(pi, tiny);
        default: return MG_FSUB(-pi, tiny);
        }
    }
    if (iy == 0x7f800000u) return (hy >> 31) ? MG_FSUB(-pi_o_2, tiny) : MG_FADD(pi_o_2, tiny);
    const int k = ((int)iy - (int)ix) >> 23;
    float z;
    if (k > 60) z = MG_FADD(pi_o_2, MG_FMUL(0.5f, pi_lo));
    else if ((hx >> 31) && k < -60) z = 0.0f;
    else z = mg_atanf(MG_ASFLOAT(MG_ASUINT(MG_FDIV(y, x)) & 0x7fffffffu));
    switch (m) {
    case 0: return z;
    case 1: return MG_ASFLOAT(MG_ASUINT(z) ^ 0x80000000u);
    case 2: return MG_FSUB(pi, MG_FSUB(z, pi_lo));
    default: return MG_FSUB(MG_FSUB(z, pi_lo), pi);
    }
}

#endif /* MODEM_GPU_LIBM_F32_H */
