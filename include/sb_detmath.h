/*
 * sb_detmath.h -- bit-reproducible exp / log for the Stage A label arithmetic.
 *
 * processMiniBatch truncates normalised weights into range-factorisation bins
 * (src/quant/SalmonQuantify.cpp:845-853); equal-weight multi-mappers land exactly on bin
 * boundaries, so a 1-ulp difference between a host and a device libm would change labels.
 * Both the CUDA path and the CPU oracle therefore evaluate exp/log with the SAME sequence of
 * IEEE-754 double operations: the published fdlibm algorithms (e_exp.c / e_log.c, Sun
 * Microsystems 1993; < 1 ulp), written with explicit round-to-nearest mul/add so that no
 * compiler contracts them into FMAs (CUDA: __dmul_rn/__dadd_rn; host: build with
 * -ffp-contract=off).  This is a math primitive shared like a libm, not algorithm logic.
 */
#ifndef SB_DETMATH_H
#define SB_DETMATH_H

#include <stdint.h>
#include <string.h>

#ifdef __CUDACC__
#define SB_HD __host__ __device__ __forceinline__
#else
#define SB_HD static inline
#endif

#if defined(__CUDA_ARCH__)
#define SB_MUL(a, b) __dmul_rn((a), (b))
#define SB_ADD(a, b) __dadd_rn((a), (b))
#define SB_SUB(a, b) __dsub_rn((a), (b))
#define SB_DIV(a, b) __ddiv_rn((a), (b))
#else
#define SB_MUL(a, b) ((a) * (b))
#define SB_ADD(a, b) ((a) + (b))
#define SB_SUB(a, b) ((a) - (b))
#define SB_DIV(a, b) ((a) / (b))
#endif

SB_HD uint64_t sbm_d2u(double x) { uint64_t u; memcpy(&u, &x, 8); return u; }
SB_HD double sbm_u2d(uint64_t u) { double x; memcpy(&x, &u, 8); return x; }

/* exp(x) for finite x; returns +inf above 709.78, 0 below -745.13 (subnormal results are
 * produced by a two-step scale, as fdlibm does). */
SB_HD double sbm_det_exp(double x) {
  const double ln2HI = 6.93147180369123816490e-01, ln2LO = 1.90821492927058770002e-10,
               invln2 = 1.44269504088896338700e+00;
  const double P1 = 1.66666666666666019037e-01, P2 = -2.77777777770155933842e-03,
               P3 = 6.61375632143793436117e-05, P4 = -1.65339022054652515390e-06,
               P5 = 4.13813679705723846039e-08;
  if (x != x) return x;
  if (x > 7.09782712893383973096e+02) return sbm_u2d(0x7ff0000000000000ull);
  if (x < -7.45133219101941108420e+02) return 0.0;
  double hi = x, lo = 0.0;
  int k = 0;
  const double ax = x < 0 ? -x : x;
  if (ax > 0.34657359027997264 /* 0.5*ln2 */) {
    if (ax < 1.0397207708399179 /* 1.5*ln2 */) {
      if (x > 0) { hi = SB_SUB(x, ln2HI); lo = ln2LO; k = 1; }
      else { hi = SB_ADD(x, ln2HI); lo = -ln2LO; k = -1; }
    } else {
      k = (int)(SB_ADD(SB_MUL(invln2, x), (x < 0 ? -0.5 : 0.5)));
      const double t = (double)k;
      hi = SB_SUB(x, SB_MUL(t, ln2HI));
      lo = SB_MUL(t, ln2LO);
    }
    x = SB_SUB(hi, lo);
  } else if (ax < 3.7252902984619140625e-09 /* 2^-28 */) {
    return SB_ADD(1.0, x);
  }
  const double t = SB_MUL(x, x);
  double c = SB_ADD(P4, SB_MUL(t, P5));
  c = SB_ADD(P3, SB_MUL(t, c));
  c = SB_ADD(P2, SB_MUL(t, c));
  c = SB_ADD(P1, SB_MUL(t, c));
  c = SB_SUB(x, SB_MUL(t, c));
  double y;
  if (k == 0) return SB_SUB(1.0, SB_SUB(SB_DIV(SB_MUL(x, c), SB_SUB(c, 2.0)), x));
  y = SB_SUB(1.0, SB_SUB(SB_SUB(lo, SB_DIV(SB_MUL(x, c), SB_SUB(2.0, c))), hi));
  if (k >= -1021) {
    return sbm_u2d(sbm_d2u(y) + ((uint64_t)(int64_t)k << 52));
  }
  /* subnormal result: scale in two steps */
  y = sbm_u2d(sbm_d2u(y) + ((uint64_t)(int64_t)(k + 1000) << 52));
  return SB_MUL(y, 9.33263618503218878990e-302 /* 2^-1000 */);
}

/* log(x) for x > 0 (finite); callers guard x <= 0. */
SB_HD double sbm_det_log(double x) {
  const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10;
  const double Lg1 = 6.666666666666735130e-01, Lg2 = 3.999999999940941908e-01,
               Lg3 = 2.857142874366239149e-01, Lg4 = 2.222219843214978396e-01,
               Lg5 = 1.818357216161805012e-01, Lg6 = 1.531383769920937332e-01,
               Lg7 = 1.479819860511658591e-01;
  int k = 0;
  uint64_t u = sbm_d2u(x);
  if ((u >> 52) == 0) {          /* subnormal: scale up by 2^54 */
    x = SB_MUL(x, 18014398509481984.0);
    u = sbm_d2u(x);
    k -= 54;
  }
  uint32_t hx = (uint32_t)(u >> 32);
  k += (int)(hx >> 20) - 1023;
  hx &= 0x000fffffu;
  const uint32_t i = (hx + 0x95f64u) & 0x100000u;
  u = ((uint64_t)(hx | (i ^ 0x3ff00000u)) << 32) | (u & 0xffffffffull);   /* normalise x or x/2 */
  k += (int)(i >> 20);
  x = sbm_u2d(u);
  const double f = SB_SUB(x, 1.0);
  const double dk = (double)k;
  const double s = SB_DIV(f, SB_ADD(2.0, f));
  const double z = SB_MUL(s, s);
  const double w = SB_MUL(z, z);
  double t1 = SB_ADD(Lg4, SB_MUL(w, Lg6));
  t1 = SB_MUL(w, SB_ADD(Lg2, SB_MUL(w, t1)));
  double t2 = SB_ADD(Lg5, SB_MUL(w, Lg7));
  t2 = SB_ADD(Lg3, SB_MUL(w, t2));
  t2 = SB_MUL(z, SB_ADD(Lg1, SB_MUL(w, t2)));
  const double R = SB_ADD(t2, t1);
  const double hfsq = SB_MUL(0.5, SB_MUL(f, f));
  /* k*ln2_hi - ((hfsq - (s*(hfsq+R) + k*ln2_lo)) - f) */
  const double inner = SB_ADD(SB_MUL(s, SB_ADD(hfsq, R)), SB_MUL(dk, ln2_lo));
  return SB_SUB(SB_MUL(dk, ln2_hi), SB_SUB(SB_SUB(hfsq, inner), f));
}

#endif /* SB_DETMATH_H */
