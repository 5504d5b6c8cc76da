/*
 * oracle/orc_math.h -- TEST INFRASTRUCTURE.  exp / log of the Stage A oracle, in two modes:
 *   ORC_MATH_LIBM   (default)  the platform libm -- what the reference's std::exp / std::log resolve to
 *                              (SalmonMath.hpp:40-81, SalmonQuantify.cpp:599-857); the mode the pinning tests use to
 *                              show that the product's labels / bins are those a libm build produces;
 *   ORC_MATH_FDLIBM            a plain-C restatement of the published fdlibm algorithms (Sun Microsystems 1993,
 *                              e_exp.c / e_log.c, < 1 ulp), evaluated without FMA contraction (-ffp-contract=off): the
 *                              same ALGORITHM the product's include/sb_detmath.h implements for the GPU, so that
 *                              regression tests can demand bit-exact integers, labels and accumulators.
 * This file does not include anything of the product.
 */
#ifndef ORC_MATH_H
#define ORC_MATH_H
#include <math.h>
#include <stdint.h>
#include <string.h>

#define ORC_MATH_LIBM 0
#define ORC_MATH_FDLIBM 1

static inline uint64_t orc_d2u(double x) { uint64_t u; memcpy(&u, &x, 8); return u; }
static inline double orc_u2d(uint64_t u) { double x; memcpy(&x, &u, 8); return x; }

/* every operation is one IEEE-754 round-to-nearest double operation (the file is built with -ffp-contract=off) */
#define ORC_MUL(a, b) ((a) * (b))
#define ORC_ADD(a, b) ((a) + (b))
#define ORC_SUB(a, b) ((a) - (b))
#define ORC_DIV(a, b) ((a) / (b))

/* exp(x) for finite x; returns +inf above 709.78, 0 below -745.13 (subnormal results are
 * produced by a two-step scale, as fdlibm does). */
static double orc_fd_exp(double x) {
  const double ln2HI = 6.93147180369123816490e-01, ln2LO = 1.90821492927058770002e-10,
               invln2 = 1.44269504088896338700e+00;
  const double P1 = 1.66666666666666019037e-01, P2 = -2.77777777770155933842e-03,
               P3 = 6.61375632143793436117e-05, P4 = -1.65339022054652515390e-06,
               P5 = 4.13813679705723846039e-08;
  if (x != x) return x;
  if (x > 7.09782712893383973096e+02) return orc_u2d(0x7ff0000000000000ull);
  if (x < -7.45133219101941108420e+02) return 0.0;
  double hi = x, lo = 0.0;
  int k = 0;
  const double ax = x < 0 ? -x : x;
  if (ax > 0.34657359027997264 /* 0.5*ln2 */) {
    if (ax < 1.0397207708399179 /* 1.5*ln2 */) {
      if (x > 0) { hi = ORC_SUB(x, ln2HI); lo = ln2LO; k = 1; }
      else { hi = ORC_ADD(x, ln2HI); lo = -ln2LO; k = -1; }
    } else {
      k = (int)(ORC_ADD(ORC_MUL(invln2, x), (x < 0 ? -0.5 : 0.5)));
      const double t = (double)k;
      hi = ORC_SUB(x, ORC_MUL(t, ln2HI));
      lo = ORC_MUL(t, ln2LO);
    }
    x = ORC_SUB(hi, lo);
  } else if (ax < 3.7252902984619140625e-09 /* 2^-28 */) {
    return ORC_ADD(1.0, x);
  }
  const double t = ORC_MUL(x, x);
  double c = ORC_ADD(P4, ORC_MUL(t, P5));
  c = ORC_ADD(P3, ORC_MUL(t, c));
  c = ORC_ADD(P2, ORC_MUL(t, c));
  c = ORC_ADD(P1, ORC_MUL(t, c));
  c = ORC_SUB(x, ORC_MUL(t, c));
  double y;
  if (k == 0) return ORC_SUB(1.0, ORC_SUB(ORC_DIV(ORC_MUL(x, c), ORC_SUB(c, 2.0)), x));
  y = ORC_SUB(1.0, ORC_SUB(ORC_SUB(lo, ORC_DIV(ORC_MUL(x, c), ORC_SUB(2.0, c))), hi));
  if (k >= -1021) {
    return orc_u2d(orc_d2u(y) + ((uint64_t)(int64_t)k << 52));
  }
  /* subnormal result: scale in two steps */
  y = orc_u2d(orc_d2u(y) + ((uint64_t)(int64_t)(k + 1000) << 52));
  return ORC_MUL(y, 9.33263618503218878990e-302 /* 2^-1000 */);
}

/* log(x) for x > 0 (finite); callers guard x <= 0. */
static double orc_fd_log(double x) {
  const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10;
  const double Lg1 = 6.666666666666735130e-01, Lg2 = 3.999999999940941908e-01,
               Lg3 = 2.857142874366239149e-01, Lg4 = 2.222219843214978396e-01,
               Lg5 = 1.818357216161805012e-01, Lg6 = 1.531383769920937332e-01,
               Lg7 = 1.479819860511658591e-01;
  int k = 0;
  uint64_t u = orc_d2u(x);
  if ((u >> 52) == 0) {          /* subnormal: scale up by 2^54 */
    x = ORC_MUL(x, 18014398509481984.0);
    u = orc_d2u(x);
    k -= 54;
  }
  uint32_t hx = (uint32_t)(u >> 32);
  k += (int)(hx >> 20) - 1023;
  hx &= 0x000fffffu;
  const uint32_t i = (hx + 0x95f64u) & 0x100000u;
  u = ((uint64_t)(hx | (i ^ 0x3ff00000u)) << 32) | (u & 0xffffffffull);   /* normalise x or x/2 */
  k += (int)(i >> 20);
  x = orc_u2d(u);
  const double f = ORC_SUB(x, 1.0);
  const double dk = (double)k;
  const double s = ORC_DIV(f, ORC_ADD(2.0, f));
  const double z = ORC_MUL(s, s);
  const double w = ORC_MUL(z, z);
  double t1 = ORC_ADD(Lg4, ORC_MUL(w, Lg6));
  t1 = ORC_MUL(w, ORC_ADD(Lg2, ORC_MUL(w, t1)));
  double t2 = ORC_ADD(Lg5, ORC_MUL(w, Lg7));
  t2 = ORC_ADD(Lg3, ORC_MUL(w, t2));
  t2 = ORC_MUL(z, ORC_ADD(Lg1, ORC_MUL(w, t2)));
  const double R = ORC_ADD(t2, t1);
  const double hfsq = ORC_MUL(0.5, ORC_MUL(f, f));
  /* k*ln2_hi - ((hfsq - (s*(hfsq+R) + k*ln2_lo)) - f) */
  const double inner = ORC_ADD(ORC_MUL(s, ORC_ADD(hfsq, R)), ORC_MUL(dk, ln2_lo));
  return ORC_SUB(ORC_MUL(dk, ln2_hi), ORC_SUB(ORC_SUB(hfsq, inner), f));
}


extern int orc_math_mode_;
static inline double m_exp(double x) { return orc_math_mode_ == ORC_MATH_FDLIBM ? orc_fd_exp(x) : exp(x); }
static inline double m_log(double x) { return orc_math_mode_ == ORC_MATH_FDLIBM ? orc_fd_log(x) : log(x); }

#endif /* ORC_MATH_H */
