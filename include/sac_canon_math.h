// sac_canon_math.h -- the canonical fp64 elementary functions of the sac_b200 bitstream.
//
// Why this exists: a .sac stream is only decodable by a decoder that repeats the encoder's fp64 arithmetic bit for
// bit (the decoder re-runs the predictor, /root/reference src/libsac/libsac.cpp:144-199; SURVEY.md "fact 1"). The
// reference gets exp/pow/log from the C library (src/pred/ols.cpp:34, src/pred/blend.h:82, src/pred/rls.h:26,
// src/pred/bias.h:153, src/pred/ls.h:38-42), which differs between glibc and CUDA libdevice. Here every
// transcendental is spelled out in IEEE-754 +, *, / and fma only, so host (gcc -ffp-contract=off) and device
// (nvcc --fmad=false) produce identical bits. Accuracy: c_exp < 1 ulp, c_log < 1 ulp, c_pow a few ulp for
// |y*log x| < 32 (error grows linearly with |y*log x|), measured in tests/test_canon_math.py.
//
// Header-only, C++11, usable from .cu (SAC_HD = __host__ __device__) and from plain host C++.
#ifndef SAC_CANON_MATH_H
#define SAC_CANON_MATH_H

#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define SAC_HD __host__ __device__ __forceinline__
#else
#define SAC_HD inline
#endif

namespace sac_canon {

SAC_HD double c_fma(double a, double b, double c)
{
#if defined(__CUDA_ARCH__)
  return __fma_rn(a, b, c);
#else
  return __builtin_fma(a, b, c);
#endif
}

SAC_HD uint64_t c_bits(double x)
{
#if defined(__CUDA_ARCH__)
  return (uint64_t)__double_as_longlong(x);
#else
  uint64_t u; memcpy(&u, &x, 8); return u;
#endif
}
SAC_HD double c_frombits(uint64_t u)
{
#if defined(__CUDA_ARCH__)
  return __longlong_as_double((long long)u);
#else
  double x; memcpy(&x, &u, 8); return x;
#endif
}

SAC_HD double c_fabs(double x) { return c_frombits(c_bits(x) & 0x7fffffffffffffffULL); }

// truncation toward zero, exact for every finite double
SAC_HD double c_trunc(double x)
{
  const uint64_t u = c_bits(x);
  const int e = (int)((u >> 52) & 0x7ff) - 1023;
  if (e >= 52) return x;                                  // already integral (or inf/nan)
  if (e < 0) return c_frombits(u & 0x8000000000000000ULL); // |x|<1 -> +-0
  return c_frombits(u & ~((1ULL << (52 - e)) - 1));
}

// C round(): nearest integer, halfway cases away from zero (libsac.cpp:106, bias.h:129)
SAC_HD double c_round(double x)
{
  const double t = c_trunc(x);
  const double d = x - t;                                 // exact
  if (d >= 0.5) return t + 1.0;
  if (d <= -0.5) return t - 1.0;
  return t;
}

// x * 2^k for finite x, any k (two-step scaling keeps every factor a normal number)
SAC_HD double c_scalbn(double x, int k)
{
  if (k > 1023) { x *= c_frombits(0x7feULL << 52); k -= 1023; if (k > 1023) { x *= c_frombits(0x7feULL << 52); k -= 1023; if (k > 1023) k = 1023; } }
  else if (k < -1022) { x *= c_frombits(54ULL << 52); k += 969; if (k < -1022) { x *= c_frombits(54ULL << 52); k += 969; if (k < -1022) k = -1022; } }
  return x * c_frombits((uint64_t)(k + 1023) << 52);
}

// e^x. Range reduction x = k*ln2 + r, |r| <= ln2/2, degree-13 Taylor polynomial (remainder < 2^-57).
SAC_HD double c_exp(double x)
{
  if (x != x) return x;
  if (x > 709.782712893384) return c_frombits(0x7ffULL << 52);
  if (x < -745.2) return 0.0;
  const double kd = c_round(x * 0x1.71547652b82fep+0);
  double r = c_fma(-kd, 0x1.62e42fee00000p-1, x);
  r = c_fma(-kd, 0x1.a39ef35793c76p-33, r);
  double p = 0x1.6124613a86d09p-33;
  p = c_fma(p, r, 0x1.1eed8eff8d898p-29);
  p = c_fma(p, r, 0x1.ae64567f544e4p-26);
  p = c_fma(p, r, 0x1.27e4fb7789f5cp-22);
  p = c_fma(p, r, 0x1.71de3a556c734p-19);
  p = c_fma(p, r, 0x1.a01a01a01a01ap-16);
  p = c_fma(p, r, 0x1.a01a01a01a01ap-13);
  p = c_fma(p, r, 0x1.6c16c16c16c17p-10);
  p = c_fma(p, r, 0x1.1111111111111p-7);
  p = c_fma(p, r, 0x1.5555555555555p-5);
  p = c_fma(p, r, 0x1.5555555555555p-3);
  p = c_fma(p, r, 0.5);
  p = c_fma(p, r, 1.0);
  p = c_fma(p, r, 1.0);
  return c_scalbn(p, (int)kd);
}

// natural log as an unevaluated sum hi+lo (|lo| << |hi|); fdlibm-style atanh series on m in [sqrt(.5),sqrt(2))
SAC_HD void c_log2sum(double x, double &hi, double &lo)
{
  uint64_t u = c_bits(x);
  int e = 0;
  if ((u >> 52) == 0) { x *= 0x1p54; u = c_bits(x); e = -54; }   // subnormal
  e += (int)(u >> 52) - 1023;
  u = (u & 0x000fffffffffffffULL) | (1023ULL << 52);
  double m = c_frombits(u);
  if (m > 0x1.6a09e667f3bcdp+0) { m *= 0.5; e += 1; }
  const double f = m - 1.0;
  const double s = f / (2.0 + f);
  const double z = s * s;
  const double w = z * z;
  const double t1 = w * c_fma(w, c_fma(w, 0x1.39a09d078c69fp-3, 0x1.c71c51d8e78afp-3), 0x1.999999997fa04p-2);
  const double t2 = z * c_fma(w, c_fma(w, c_fma(w, 0x1.2f112df3e5244p-3, 0x1.7466496cb03dep-3), 0x1.2492494229359p-2), 0x1.5555555555593p-1);
  const double R = t2 + t1;
  const double hfsq = 0.5 * f * f;
  const double dk = (double)e;
  // log(m) = f - (hfsq - s*(hfsq+R)); log(x) = dk*ln2_hi + (log(m) + dk*ln2_lo)
  const double c = c_fma(s, hfsq + R, dk * 0x1.a39ef35793c76p-33);
  const double a = dk * 0x1.62e42fee00000p-1;                        // exact: ln2_hi has 21 trailing zero bits
  hi = a + f;
  lo = (c - hfsq) + ((a - hi) + f);                                  // fast two-sum: dk==0 or |a| > |f|
}

SAC_HD double c_log(double x)
{
  if (x != x || x < 0.0) return c_frombits(0x7ff8ULL << 48);
  if (x == 0.0) return c_frombits(0xfffULL << 52);
  if (c_bits(x) == (0x7ffULL << 52)) return x;
  double hi, lo;
  c_log2sum(x, hi, lo);
  return hi + lo;
}

// x^y for x > 0 (all call sites: ols.cpp:34, ls.h:39-41, vle.cpp:19,75). exp(y*log x) with the product carried
// as a double-double so the only amplified error is log's own rounding.
SAC_HD double c_pow(double x, double y)
{
  if (y == 0.0 || x == 1.0) return 1.0;
  if (x == 0.0) return y > 0.0 ? 0.0 : c_frombits(0x7ffULL << 52);
  double hi, lo;
  c_log2sum(x, hi, lo);
  const double l = hi + lo;
  const double le = lo - (l - hi);              // l + le == hi + lo (fast two-sum, |hi|>=|lo|)
  const double ph = y * l;
  const double pl = c_fma(y, l, -ph) + y * le;  // y*(l+le) = ph + pl
  const double e = c_exp(ph);
  return c_fma(e, pl, e);
}

} // namespace sac_canon
#endif
