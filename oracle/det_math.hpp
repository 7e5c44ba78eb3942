// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the shipped product path.
//
// Deterministic double-precision elementary functions used by the oracle's
// RANSAC stopping rule and 3x3 eigen solve.  PCL calls libm (log/pow/atan2f/
// cosf/sinf); libm and CUDA's device math differ in the last ulp, which would
// make "bit-exact inlier sets" depend on the math library.  Both the oracle
// (this file) and the CUDA library (csrc/det_math.cuh, written separately from
// the same *specification* below) therefore evaluate these functions with a
// fixed sequence of IEEE-754 +,-,*,/ operations in double, no FMA contraction
// (-ffp-contract=off here, -fmad=false on the device).  Results are within
// 1e-15 relative of libm; after rounding to float they equal the correctly
// rounded float result except in ~1e-8 of cases (tests/test_oracle_kat.py
// checks the deviation against libm).
//
// SPECIFICATION (shared in words, not in code, with csrc/det_math.cuh)
//   det_log(x), x>0 finite:   x = m*2^e with m in [sqrt(1/2), sqrt(2));
//       s=(m-1)/(m+1); z=s*s; P = Horner over 1/(2k+1), k=14..0, in z;
//       result = e*LN2 + 2*s*P            (LN2 = 0x1.62e42fefa39efp-1)
//   det_atan(t), 0<=t<=1:     k=(int)(4*t+0.5); c=k/4; u=(t-c)/(1+t*c);
//       Q = Horner of (-1)^j/(2j+1), j=10..0, in u*u; result = ATAN_K[k]+u*Q
//   det_atan2(y,x), y>=0:     standard octant reduction on det_atan
//   det_sin/det_cos(x), |x|<=pi/2 (callers use [0, pi/3]): Taylor to x^23 / x^22
//   det_asin(q):  NaN for NaN or |q|>1; a=|q|; x=sqrt((1-a)*(1+a)); r=det_atan2(a,x); sign of q restored
//   det_tan(x), |x|<=~pi/2:   det_sin(x)/det_cos(x)
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

namespace pcop_oracle {

inline double det_log(double x) {
  // frexp/ldexp are exact.
  int e;
  double m = std::frexp(x, &e);  // m in [0.5, 1)
  if (m < 0.70710678118654752440) { m = m * 2.0; e -= 1; }  // m in [sqrt(.5), sqrt(2))
  const double s = (m - 1.0) / (m + 1.0);
  const double z = s * s;
  double p = 1.0 / 29.0;
  for (int k = 13; k >= 0; --k) p = p * z + 1.0 / (double)(2 * k + 1);
  return (double)e * 0.69314718055994530942 + 2.0 * s * p;
}

inline double det_atan01(double t) {  // 0 <= t <= 1
  static const double ATAN_K[5] = {0.0, 0.24497866312686415417, 0.46364760900080611621,
                                   0.64350110879328438680, 0.78539816339744830962};
  const int k = (int)(4.0 * t + 0.5);
  const double c = (double)k * 0.25;
  const double u = (t - c) / (1.0 + t * c);
  const double z = u * u;
  double q = 1.0 / 21.0;
  // atan(u) = u*(1 - z*(1/3 - z*(1/5 - ...))): alternating series, Horner from the top
  for (int j = 9; j >= 0; --j) q = 1.0 / (double)(2 * j + 1) - z * q;
  return ATAN_K[k] + u * q;
}

// atan2 for y >= 0 (the only case eigen33 needs: y = sqrt(-q)).
inline double det_atan2_ypos(double y, double x) {
  const double PI = 3.14159265358979323846, PI_2 = 1.57079632679489661923;
  if (y == 0.0) return (x >= 0.0) ? 0.0 : PI;
  if (x == 0.0) return PI_2;
  const double ax = std::fabs(x);
  double a;
  if (y <= ax) a = det_atan01(y / ax);
  else a = PI_2 - det_atan01(ax / y);
  return (x > 0.0) ? a : (PI - a);
}

inline double det_sin(double x) {
  const double z = x * x;
  double p = 1.0 / 25852016738884976640000.0;  // 1/23!
  static const double INV_FACT[11] = {
      1.0, 1.0 / 6.0, 1.0 / 120.0, 1.0 / 5040.0, 1.0 / 362880.0, 1.0 / 39916800.0,
      1.0 / 6227020800.0, 1.0 / 1307674368000.0, 1.0 / 355687428096000.0,
      1.0 / 121645100408832000.0, 1.0 / 51090942171709440000.0};
  for (int j = 10; j >= 0; --j) p = INV_FACT[j] - z * p;
  return x * p;
}

inline double det_cos(double x) {
  const double z = x * x;
  double p = 1.0 / 1124000727777607680000.0;  // 1/22!
  static const double INV_FACT[11] = {
      1.0, 1.0 / 2.0, 1.0 / 24.0, 1.0 / 720.0, 1.0 / 40320.0, 1.0 / 3628800.0,
      1.0 / 479001600.0, 1.0 / 87178291200.0, 1.0 / 20922789888000.0,
      1.0 / 6402373705728000.0, 1.0 / 2432902008176640000.0};
  for (int j = 10; j >= 0; --j) p = INV_FACT[j] - z * p;
  return p;
}

// asin / tan for the shadow-length formula (od.cpp:543-545)
inline double det_asin(double q) {
  if (q != q || std::fabs(q) > 1.0) return std::nan("");
  const double a = std::fabs(q);
  const double x = std::sqrt((1.0 - a) * (1.0 + a));
  const double r = det_atan2_ypos(a, x);
  return (q < 0.0) ? -r : r;
}

inline double det_tan(double x) { return det_sin(x) / det_cos(x); }

}  // namespace pcop_oracle
