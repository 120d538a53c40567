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
 * Arguments of 105414350 or more (yaw beyond 6*10^9 degrees) take the platform's own sincos.
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
#endif

/* glibc __sincostab: {sn, ssn, cs, ccs} for xk = k/128, k = 0..109 */
Q1LIBM_TABLE kTab[440] = {
#include "q1_libm_sincos_tab.inc"
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

/* sin(x) and cos(x) as __sin / __cos return them, for |x| < 105414350 (high word below
 * 0x419921FB).  Returns false (outputs untouched) for larger, infinite or NaN arguments.
 *
 * In every range one do_sin and one do_cos evaluation serve both results; the ranges differ only
 * in the arguments handed to them and in which result goes where:
 *   direct:  S = do_sin(x, 0),     C = do_cos(x, 0);      sin = S, cos = C
 *   fold:    S = do_sin(af, daf),  C = do_cos(tf, hp1);   sin = copysign(C, x), cos = S
 *   reduce:  S = do_sin(b, db),    C = do_cos(b, db);     n odd swaps; sin negated when n & 2,
 *                                                         cos when (n + 1) & 2
 * The direct range is the reduction with xn forced to 0 (then y = t2 = b = x, db = 0, n = 0 come
 * out of the same operations exactly). */
Q1LIBM_FN bool sincos(double x, double &sin_out, double &cos_out)
{
    const uint32_t hx = f_hi(x);
    const uint32_t k = hx & 0x7fffffffu;
    if (!(k < kHiReduce))
        return false;
    const double ax = f_abs(x);
    const bool direct = k < kHiTable, fold = !direct & (k < kHiFold);

    /* reduce_sincos */
    double tq = f_fma(x, kC[C_HPINV], kC[C_TOINT]);
    tq = direct ? kC[C_TOINT] : tq;
    const double xn = f_sub(tq, kC[C_TOINT]);
    const uint32_t n = f_lo(tq) & 3u;
    const double y = f_fma(-xn, kC[C_MP2], f_fma(-xn, kC[C_MP1], x));
    const double t2 = f_fma(-xn, kC[C_PP3], y);
    const double db1 = f_fma(-kC[C_PP3], xn, f_sub(y, t2));
    const double b = f_fma(-xn, kC[C_PP4], t2);
    const double db2 = f_fma(-xn, kC[C_PP4], f_sub(t2, b));
    const double db = f_add(db1, db2);

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
