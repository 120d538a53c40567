/*
 * q1_libm_sincos.cuh -- sin and cos of a double, bit for bit as the reference computes them.
 *
 * The reference's only inexact primitive is np.sin / np.cos on float64 (phys.py:58-59), which NumPy
 * forwards to the C library: glibc 2.39 sysdeps/ieee754/dbl-64/s_sin.c (__sin / __cos), an
 * un-vendored dependency.  glibc's functions are faithful (< 0.55 ulp), not correctly rounded, so
 * no other algorithm -- however accurate -- returns the same bits.  This header restates the
 * published algorithm:
 *
 *   |x| < 0.855469          do_sin(x, 0) / do_cos(x, 0)
 *   |x| < 2.426265          t = hp0 - |x|:  sin = do_cos(t, hp1),  cos = do_sin(t + hp1, tail)
 *   |x| < 105414350         x = n pi/2 + (a + da) by a four-constant Cody-Waite reduction
 *                           (reduce_sincos); quadrant n picks do_sin(a, da) / do_cos(a, da) and signs
 *   larger finite |x|       the same with (n, a, da) from __branred (branred.c): x 2^-600 split in two
 *                           26-bit pieces, each multiplied by 144 bits of 2/pi taken from a table
 *   do_sin / do_cos         x = xk + r with xk = k/128 from the 440-double table {sin, cos}(xk)
 *                           in two words each; short polynomials in r; angle-addition correction
 *   do_sin, |x| < 0.126     a degree-11 Taylor form instead of the table
 *
 * with the operation order AND the multiply-add fusions of the build every x86-64 machine with
 * FMA selects at run time (the `__sin_fma` / `__cos_fma` ifunc variants: the same C compiled with
 * -mfma -mavx2, where GCC contracts a*b+c).  The fusion pattern below was read off that build's
 * instruction stream; tests/test_libm_sincos.py compares this header, compiled for the host, with
 * the installed libm on 2*10^8 arguments (uniform, log-uniform, around every branch threshold and
 * around multiples of pi/2), and tests/test_cuda_parity.py does the same for the device build.
 *
 * The file compiles as CUDA (device, *_rn intrinsics: nothing the compiler may re-fuse) and as
 * plain C++ (host, for the CPU test: build with -ffp-contract=off).
 */
#pragma once

#include <cstdint>
#include <cstring>
#if !defined(__CUDACC__)
#include <cmath>
#endif

namespace q1libm {

#if defined(__CUDACC__)
#define Q1LIBM_FN __device__ __forceinline__
#define Q1LIBM_TABLE __device__ __align__(32) const double
Q1LIBM_FN double f_fma(double a, double b, double c) { return __fma_rn(a, b, c); }
Q1LIBM_FN double f_mul(double a, double b) { return __dmul_rn(a, b); }
Q1LIBM_FN double f_add(double a, double b) { return __dadd_rn(a, b); }
Q1LIBM_FN double f_sub(double a, double b) { return __dsub_rn(a, b); }
Q1LIBM_FN uint32_t f_hi(double a) { return (uint32_t)__double2hiint(a); }
Q1LIBM_FN uint32_t f_lo(double a) { return (uint32_t)__double2loint(a); }
Q1LIBM_FN double f_flip(double a, uint32_t sign_bit) /* a with its sign XORed by sign_bit */
{
    return __hiloint2double((int)(f_hi(a) ^ sign_bit), (int)f_lo(a));
}
Q1LIBM_FN double f_abs(double a) { return fabs(a); }
Q1LIBM_FN double f_from_hi(uint32_t hi) { return __hiloint2double((int)hi, 0); }
#else
#define Q1LIBM_FN static inline
#define Q1LIBM_TABLE static const double
Q1LIBM_FN double f_fma(double a, double b, double c) { return std::fma(a, b, c); }
Q1LIBM_FN double f_mul(double a, double b) { return a * b; }
Q1LIBM_FN double f_add(double a, double b) { return a + b; }
Q1LIBM_FN double f_sub(double a, double b) { return a - b; }
Q1LIBM_FN uint64_t f_bits(double a) { uint64_t u; std::memcpy(&u, &a, 8); return u; }
Q1LIBM_FN uint32_t f_hi(double a) { return (uint32_t)(f_bits(a) >> 32); }
Q1LIBM_FN uint32_t f_lo(double a) { return (uint32_t)f_bits(a); }
Q1LIBM_FN double f_flip(double a, uint32_t sign_bit)
{
    uint64_t u = f_bits(a) ^ ((uint64_t)sign_bit << 32);
    double r; std::memcpy(&r, &u, 8); return r;
}
Q1LIBM_FN double f_abs(double a) { return std::fabs(a); }
Q1LIBM_FN double f_from_hi(uint32_t hi)
{
    uint64_t u = (uint64_t)hi << 32;
    double r; std::memcpy(&r, &u, 8); return r;
}
#endif

/* glibc __sincostab: {sn, ssn, cs, ccs} for xk = k/128, k = 0..109 */
Q1LIBM_TABLE kTab[440] = {
#include "q1_libm_sincos_tab.inc"
};

/* glibc branred.h toverp[75]: the digits of 2/pi in base 2^24 */
Q1LIBM_TABLE kToverp[75] = {
#include "q1_libm_branred_tab.inc"
};

/* s_sin.c / usncs.h / trigo.h constants (values as stored in the library).  On the device they
 * sit in the constant bank, where the FP64 pipe reads them as instruction operands; as literals
 * each would cost two uniform moves per use. */
enum {
    C_BIG, C_TOINT, C_HPINV, C_MP1, C_MP2, C_PP3, C_PP4, C_HP0, C_HP1, C_SN3, C_SN5, C_CS2, C_CS4, C_CS6,
    C_S1, C_S2, C_S3, C_S4, C_S5, C_TAYLOR, C_COUNT
};
#if defined(__CUDACC__)
static __constant__ double kC[C_COUNT] = {
#else
static const double kC[C_COUNT] = {
#endif
    0x1.8p45,                    /* big: adding it rounds |x| to a multiple of 2^-7 */
    0x1.8p52,                    /* toint */
    0x1.45f306dc9c883p-1,        /* hpinv = 2/pi */
    0x1.921fb58000000p+0,        /* mp1, mp2, pp3, pp4: pi/2 in four pieces */
    -0x1.dde973c000000p-27,
    -0x1.cb3b398000000p-55,
    -0x1.d747f23e32ed7p-83,
    0x1.921fb54442d18p+0,        /* hp0, hp1: pi/2 high, low */
    0x1.1a62633145c07p-54,
    -0x1.5555555555515p-3,       /* sn3, sn5 */
    0x1.11110e829872fp-7,
    0.5,                         /* cs2, cs4, cs6 */
    -0x1.5555555555535p-5,
    0x1.6c16bedd9e239p-10,
    -0x1.5555555555555p-3,       /* s1 .. s5 of TAYLOR_SIN */
    0x1.1111111110ecep-7,
    -0x1.a01a019db08b8p-13,
    0x1.71de27b9a7ed9p-19,
    -0x1.addffc2fcdf59p-26,
    0.126                        /* do_sin switches to the Taylor form below this */
};

constexpr uint32_t kHiTiny = 0x3e500000u;    /* |x| < 2^-26 */
constexpr uint32_t kHiTable = 0x3feb6000u;   /* |x| < 0.855469: table directly */
constexpr uint32_t kHiFold = 0x400368fdu;    /* |x| < 2.426265: fold around pi/2 */
constexpr uint32_t kHiReduce = 0x419921FBu;  /* |x| < 105414350: reduce_sincos */

struct TabEntry { double sn, ssn, cs, ccs; };

/* table entry k (k <= 109 for every argument the paths below produce).  The table (3.5 KB) stays in
 * global memory and is read through L1; a copy in shared memory was measured 2 % slower. */
Q1LIBM_FN TabEntry lookup(uint32_t k)
{
    TabEntry e;
#if defined(__CUDACC__)
    /* one 32-byte entry = one sector = one 256-bit load (sm_100: LDG.E.256) */
    asm("ld.global.nc.v4.f64 {%0, %1, %2, %3}, [%4];"
        : "=d"(e.sn), "=d"(e.ssn), "=d"(e.cs), "=d"(e.ccs) : "l"(kTab + 4 * k));
#else
    e.sn = kTab[4 * k]; e.ssn = kTab[4 * k + 1]; e.cs = kTab[4 * k + 2]; e.ccs = kTab[4 * k + 3];
#endif
    return e;
}

/* One half of __branred (branred.c): x1 (a 26-bit piece of x 2^-600) times 2/pi, integer part
 * dropped, fraction as b + bb, integer part mod 4 kept in `sum`.  Plain multiplies and adds -- this
 * file of glibc is built without FMA contraction even for the FMA variants of sin / cos. */
Q1LIBM_FN void branred_half(double x1, double &b, double &bb, double &sum)
{
    const double big = 0x1.8p52, big1 = 0x1.8p54, tm24 = 0x1p-24;
    int k = (int)((f_hi(x1) >> 20) & 2047u);
    k = (k - 450) / 24;
    if (k < 0)
        k = 0;
    /* gor = 2^576 with its exponent lowered by 24 k */
    double gor = f_from_hi(0x63f00000u - ((uint32_t)(k * 24) << 20));
    double r[6];
#pragma unroll
    for (int i = 0; i < 6; i++) {
        r[i] = f_mul(f_mul(x1, kToverp[k + i]), gor);
        gor = f_mul(gor, tm24);
    }
    sum = 0.0;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const double s = f_sub(f_add(r[i], big), big);       /* round to nearest integer */
        sum = f_add(sum, s);
        r[i] = f_sub(r[i], s);
    }
    double t = 0.0;
#pragma unroll
    for (int i = 0; i < 6; i++)
        t = f_add(t, r[5 - i]);
    bb = f_add(f_add(f_add(f_add(f_add(f_sub(r[0], t), r[1]), r[2]), r[3]), r[4]), r[5]);
    double s = f_sub(f_add(t, big), big);
    sum = f_add(sum, s);
    t = f_sub(t, s);
    b = f_add(t, bb);
    bb = f_add(f_sub(t, b), bb);
    s = f_sub(f_add(sum, big1), big1);
    sum = f_sub(sum, s);
}

/* __branred (branred.c): x = n pi/2 + (a + aa) for |x| >= 105414350, n mod 4 returned */
Q1LIBM_FN uint32_t branred(double x, double &a, double &aa)
{
    const double split = 134217729.0;                         /* 2^27 + 1 */
    const double mp1 = 0x1.921fb58000000p+0, mp2 = -0x1.dde9740000000p-27;
    x = f_mul(x, 0x1p-600);
    double t = f_mul(x, split);                               /* split x into two numbers */
    const double x1 = f_sub(t, f_sub(t, x));
    const double x2 = f_sub(x, x1);
    double b1, bb1, sum1, b2, bb2, sum2;
    branred_half(x1, b1, bb1, sum1);
    branred_half(x2, b2, bb2, sum2);
    double sum = f_add(sum1, sum2);
    double b = f_add(b1, b2);
    double bb = f_abs(b1) > f_abs(b2) ? f_add(f_sub(b1, b), b2) : f_add(f_sub(b2, b), b1);
    if (b > 0.5) {
        b = f_sub(b, 1.0);
        sum = f_add(sum, 1.0);
    } else if (b < -0.5) {
        b = f_add(b, 1.0);
        sum = f_sub(sum, 1.0);
    }
    double s = f_add(b, f_add(f_add(bb, bb1), bb2));
    t = f_add(f_add(f_sub(b, s), bb), f_add(bb1, bb2));
    b = f_mul(s, split);
    const double t1 = f_sub(b, f_sub(b, s));
    const double t2 = f_sub(s, t1);
    b = f_mul(s, kC[C_HP0]);
    bb = f_add(f_add(f_add(f_sub(f_mul(t1, mp1), b), f_mul(t1, mp2)), f_mul(t2, mp1)),
               f_add(f_add(f_mul(t2, mp2), f_mul(s, kC[C_HP1])), f_mul(t, kC[C_HP0])));
    s = f_add(b, bb);
    t = f_add(f_sub(b, s), bb);
    a = s;
    aa = t;
    return (uint32_t)((int)sum) & 3u;
}

/* sin(x) and cos(x) as __sin / __cos return them, for every finite x (HUGE = false: only for
 * |x| < 105414350; the per-tick kernels use that -- carrying the __branred code, even out of line,
 * costs them 0.7 - 1.7 % -- and hand larger arguments, yaw past 6e9 degrees, to libdevice).  Returns false (outputs untouched) for larger, infinite or NaN arguments.
 *
 * In every range one do_sin and one do_cos evaluation serve both results; the ranges differ only
 * in the arguments handed to them and in which result goes where:
 *   direct:  S = do_sin(x, 0),     C = do_cos(x, 0);      sin = S, cos = C
 *   fold:    S = do_sin(af, daf),  C = do_cos(tf, hp1);   sin = copysign(C, x), cos = S
 *   reduce:  S = do_sin(b, db),    C = do_cos(b, db);     n odd swaps; sin negated when n & 2,
 *                                                         cos when (n + 1) & 2
 * The direct range is the reduction with xn forced to 0 (then y = t2 = b = x, db = 0, n = 0 come
 * out of the same operations exactly). */
template <bool HUGE = true>
Q1LIBM_FN bool sincos(double x, double &sin_out, double &cos_out)
{
    const uint32_t hx = f_hi(x);
    const uint32_t k = hx & 0x7fffffffu;
    const double ax = f_abs(x);
    const bool direct = k < kHiTable, fold = !direct & (k < kHiFold);
    double b, db;
    uint32_t n;
    if (!(k < kHiReduce)) {
        /* |x| >= 105414350: __branred, then the same do_sin / do_cos choice as the reduced range */
        if (!HUGE || k >= 0x7ff00000u)
            return false;
        n = branred(x, b, db);
    } else {
        /* reduce_sincos */
        double tq = f_fma(x, kC[C_HPINV], kC[C_TOINT]);
        tq = direct ? kC[C_TOINT] : tq;
        const double xn = f_sub(tq, kC[C_TOINT]);
        n = f_lo(tq) & 3u;
        const double y = f_fma(-xn, kC[C_MP2], f_fma(-xn, kC[C_MP1], x));
        const double t2 = f_fma(-xn, kC[C_PP3], y);
        const double db1 = f_fma(-kC[C_PP3], xn, f_sub(y, t2));
        b = f_fma(-xn, kC[C_PP4], t2);
        const double db2 = f_fma(-xn, kC[C_PP4], f_sub(t2, b));
        db = f_add(db1, db2);
    }

    /* fold around pi/2 */
    const double tf = f_sub(kC[C_HP0], ax);
    const double af = f_add(tf, kC[C_HP1]);
    const double daf = f_add(f_sub(tf, af), kC[C_HP1]);

    const double sa = fold ? af : b, sda = fold ? daf : db;      /* do_sin's (x, dx) */
    const double ca = fold ? tf : b, cda = fold ? kC[C_HP1] : db; /* do_cos's (x, dx) */

    /* ---- do_sin(sa, sda) ---- */
    const double asa = f_abs(sa);
    /* TAYLOR_SIN, taken for |sa| < 0.126 */
    const double xx0 = f_mul(sa, sa);
    double p0 = f_fma(kC[C_S5], xx0, kC[C_S4]);
    p0 = f_fma(p0, xx0, kC[C_S3]);
    p0 = f_fma(p0, xx0, kC[C_S2]);
    p0 = f_fma(p0, xx0, kC[C_S1]);
    const double w = f_fma(p0, sa, -f_mul(sda, kC[C_CS2]));
    const double taylor = f_add(sa, f_fma(xx0, w, sda));
    /* table form: `if (x <= 0) dx = -dx` -- the sign bit decides, since x = +-0 is a Taylor case */
    const double sd = f_flip(sda, f_hi(sa) & 0x80000000u);
    const double su = f_add(kC[C_BIG], asa);
    const double sr = f_sub(asa, f_sub(su, kC[C_BIG]));
    const uint32_t sk = f_lo(su);
    /* ---- do_cos(ca, cda) ----  `if (x < 0) dx = -dx`; for x = -0 the sign of dx cannot change the
     * result (k = 0: sn = ssn = ccs = 0, cs = 1, leaving 1 - c, even in r), so the sign bit decides */
    const double aca = f_abs(ca);
    const double cd = f_flip(cda, f_hi(ca) & 0x80000000u);
    const double cu = f_add(kC[C_BIG], aca);
    const double cr = f_add(f_sub(aca, f_sub(cu, kC[C_BIG])), cd);
    const uint32_t ck = f_lo(cu);
    /* (the two entries are the same one except in the fold range when |tf| and |af| round to
     * different nodes; two unconditional loads are cheaper than the test and the register copies) */
    const TabEntry tc = lookup(ck);
    const TabEntry ts = lookup(sk);

    const double sxx = f_mul(sr, sr);
    const double ss = f_add(sr, f_fma(f_mul(sr, sxx), f_fma(sxx, kC[C_SN5], kC[C_SN3]), sd));
    const double sq = f_fma(sxx, f_fma(sxx, kC[C_CS6], kC[C_CS4]), kC[C_CS2]);
    const double sc = f_fma(sr, sd, f_mul(sxx, sq));
    const double scor = f_fma(ss, ts.cs, f_fma(-sc, ts.sn, f_fma(ss, ts.ccs, ts.ssn)));
    const double sres = f_add(ts.sn, scor);
    /* copysign(res, x): res = sin(|x|) > 0.125 wherever the table form is the one selected */
    const double tabled = f_flip(sres, f_hi(sa) & 0x80000000u);
    const double vs = asa < kC[C_TAYLOR] ? taylor : tabled;

    const double cxx = f_mul(cr, cr);
    const double cs_ = f_fma(f_mul(cr, cxx), f_fma(cxx, kC[C_SN5], kC[C_SN3]), cr);
    const double cc = f_mul(cxx, f_fma(cxx, f_fma(cxx, kC[C_CS6], kC[C_CS4]), kC[C_CS2]));
    const double ccor = f_fma(-cs_, tc.sn, f_fma(-cc, tc.cs, f_fma(-cs_, tc.ssn, tc.ccs)));
    const double vc = f_add(tc.cs, ccor);

    /* which result is the sine, and the signs */
    const bool swap = fold | ((n & 1u) != 0u);
    const uint32_t sin_flip = fold ? (hx & 0x80000000u) : ((n & 2u) << 30);
    const uint32_t cos_flip = fold ? 0u : (((n + 1u) & 2u) << 30);
    const double sv = f_flip(swap ? vc : vs, sin_flip);
    sin_out = k < kHiTiny ? x : sv;          /* |x| < 2^-26: __sin returns x (keeps -0) */
    cos_out = f_flip(swap ? vs : vc, cos_flip);
    return true;
}

} // namespace q1libm
