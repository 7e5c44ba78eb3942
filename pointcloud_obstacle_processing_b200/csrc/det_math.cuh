// Deterministic double-precision elementary functions for the device side of the RANSAC
// stopping rule and the 3x3 eigen solve.  Written from the SPECIFICATION in
// oracle/det_math.hpp's header (shared in words): every operation is an explicit IEEE-754
// round-to-nearest intrinsic, so nothing is contracted into an FMA.
//
//   det_log(x):      x = m * 2^e, m in [sqrt(1/2), sqrt(2)); s = (m-1)/(m+1); z = s*s;
//                    P = Horner over 1/(2k+1), k = 14..0, in z;  e*LN2 + (2*s)*P
//   det_atan01(t):   k = (int)(4t + 0.5); c = k/4; u = (t-c)/(1+t*c);
//                    Q = Horner of (-1)^j/(2j+1), j = 10..0, in u*u;  ATAN_K[k] + u*Q
//   det_atan2_ypos:  octant reduction on det_atan01, y >= 0
//   det_sin/det_cos: alternating Taylor series to x^23 / x^22, Horner in x*x
//   det_asin(q):     NaN for NaN or |q| > 1; a = |q|; x = sqrt((1-a)*(1+a)); det_atan2_ypos(a, x), sign of q restored
//   det_tan(x):      det_sin(x) / det_cos(x)
#pragma once
#include "common.cuh"

namespace pcop {

// 1 / (2k + 1), k = 0..14: the correctly rounded quotients the specification's ddiv(1.0, 2k + 1) produces, evaluated
// by the compiler (IEEE double division) instead of fourteen dependent divisions per call
#define PCOP_INV_ODD_TABLE                                                                                            \
  {1.0 / 1.0,  1.0 / 3.0,  1.0 / 5.0,  1.0 / 7.0,  1.0 / 9.0,  1.0 / 11.0, 1.0 / 13.0, 1.0 / 15.0, 1.0 / 17.0, 1.0 / 19.0, \
   1.0 / 21.0, 1.0 / 23.0, 1.0 / 25.0, 1.0 / 27.0, 1.0 / 29.0}

__device__ inline double det_log(double x) {
  int e;
  double m = frexp(x, &e);  // exact; m in [0.5, 1)
  if (m < 0.70710678118654752440) {
    m = dmul(m, 2.0);
    e -= 1;
  }
  const double s = ddiv(dsub(m, 1.0), dadd(m, 1.0));
  const double z = dmul(s, s);
  const double INV_ODD[15] = PCOP_INV_ODD_TABLE;
  double p = INV_ODD[14];
#pragma unroll
  for (int k = 13; k >= 0; --k) p = dadd(dmul(p, z), INV_ODD[k]);
  return dadd(dmul((double)e, 0.69314718055994530942), dmul(dmul(2.0, s), p));
}

__device__ inline double det_atan01(double t) {
  const double ATAN_K[5] = {0.0, 0.24497866312686415417, 0.46364760900080611621, 0.64350110879328438680,
                            0.78539816339744830962};
  const int k = (int)dadd(dmul(4.0, t), 0.5);
  const double c = dmul((double)k, 0.25);
  const double u = ddiv(dsub(t, c), dadd(1.0, dmul(t, c)));
  const double z = dmul(u, u);
  const double INV_ODD[15] = PCOP_INV_ODD_TABLE;
  double q = INV_ODD[10];
#pragma unroll
  for (int j = 9; j >= 0; --j) q = dsub(INV_ODD[j], dmul(z, q));
  return dadd(ATAN_K[k], dmul(u, q));
}

__device__ inline double det_atan2_ypos(double y, double x) {
  const double PI = 3.14159265358979323846, PI_2 = 1.57079632679489661923;
  if (y == 0.0) return (x >= 0.0) ? 0.0 : PI;
  if (x == 0.0) return PI_2;
  const double ax = fabs(x);
  double a;
  if (y <= ax) a = det_atan01(ddiv(y, ax));
  else a = dsub(PI_2, det_atan01(ddiv(ax, y)));
  return (x > 0.0) ? a : dsub(PI, a);
}

__device__ inline double det_sin(double x) {
  const double F[11] = {1.0,
                        1.0 / 6.0,
                        1.0 / 120.0,
                        1.0 / 5040.0,
                        1.0 / 362880.0,
                        1.0 / 39916800.0,
                        1.0 / 6227020800.0,
                        1.0 / 1307674368000.0,
                        1.0 / 355687428096000.0,
                        1.0 / 121645100408832000.0,
                        1.0 / 51090942171709440000.0};
  const double z = dmul(x, x);
  double p = 1.0 / 25852016738884976640000.0;
  for (int j = 10; j >= 0; --j) p = dsub(F[j], dmul(z, p));
  return dmul(x, p);
}

__device__ inline double det_cos(double x) {
  const double F[11] = {1.0,
                        1.0 / 2.0,
                        1.0 / 24.0,
                        1.0 / 720.0,
                        1.0 / 40320.0,
                        1.0 / 3628800.0,
                        1.0 / 479001600.0,
                        1.0 / 87178291200.0,
                        1.0 / 20922789888000.0,
                        1.0 / 6402373705728000.0,
                        1.0 / 2432902008176640000.0};
  const double z = dmul(x, x);
  double p = 1.0 / 1124000727777607680000.0;
  for (int j = 10; j >= 0; --j) p = dsub(F[j], dmul(z, p));
  return p;
}

__device__ inline double det_asin(double q) {
  if (q != q || fabs(q) > 1.0) return __longlong_as_double(0x7ff8000000000000ll);
  const double a = fabs(q);
  const double x = __dsqrt_rn(dmul(dsub(1.0, a), dadd(1.0, a)));
  const double r = det_atan2_ypos(a, x);
  return (q < 0.0) ? -r : r;
}

__device__ inline double det_tan(double x) { return ddiv(det_sin(x), det_cos(x)); }

}  // namespace pcop
