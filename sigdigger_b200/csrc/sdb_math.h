// sdb_math.h -- SPEC.md section M: deterministic binary32 elementary functions, usable from host
// set-up code and from the chain kernels (compiled with -fmad=false / -ffp-contract=off).
// Fixed Cody-Waite reductions + fixed polynomials (classic Cephes single-precision coefficient sets),
// evaluated with one rounding per operator in the order written.
#pragma once
#include <math.h>
#include <string.h>
#ifdef __CUDACC__
#define SDB_HD static __host__ __device__ __forceinline__
#else
#define SDB_HD static inline
#endif

SDB_HD unsigned sdb_f2u(float f)
{
#ifdef __CUDA_ARCH__
  return __float_as_uint(f);
#else
  unsigned u; memcpy(&u, &f, 4); return u;
#endif
}
SDB_HD float sdb_u2f(unsigned u)
{
#ifdef __CUDA_ARCH__
  return __uint_as_float(u);
#else
  float f; memcpy(&f, &u, 4); return f;
#endif
}

// ------------------------------------------------------------------ SPEC M: deterministic math --
SDB_HD void d_sincosf(float x, float *s, float *c)
{
  float q = rintf(x * 0.636619772367581343f);
  float r = x - q * 1.5703125f;
  r = r - q * 4.837512969970703125e-4f;
  r = r - q * 7.54978995489188e-8f;
  float z = r * r;
  float sp = ((-1.9515295891e-4f * z + 8.3321608736e-3f) * z - 1.6666654611e-1f) * z * r + r;
  float cp = ((2.443315711809948e-5f * z - 1.388731625493765e-3f) * z + 4.166664568298827e-2f)
             * z * z - 0.5f * z + 1.0f;
  int n = (int) q & 3;
  float s0 = (n & 1) ? cp : sp;
  float c0 = (n & 1) ? sp : cp;
  *s = (n & 2) ? -s0 : s0;
  *c = ((n + 1) & 2) ? -c0 : c0;
}

// SPEC M.2 in select form: the three reduction ranges share one division (t/1 is exact, so the small range is
// unchanged) and every `if` of the statement is a select -- the same operations on the same operands, hence the same
// bits, without per-lane branches (32 chains of a warp rarely agree on the octant).
SDB_HD float d_atanf_pos(float t)
{
  const bool big = t > 2.414213562373095f, mid = !big && t > 0.4142135623730950f;
  const float y0 = big ? 1.5707963267948966f : (mid ? 0.7853981633974483f : 0.0f);
  const float num = big ? -1.0f : (mid ? t - 1.0f : t);
  const float den = big ? t : (mid ? t + 1.0f : 1.0f);
  t = num / den;
  float z = t * t;
  float y = (((8.05374449538e-2f * z - 1.38776856032e-1f) * z + 1.99777106478e-1f) * z
             - 3.33329491539e-1f) * z * t + t;
  return y0 + y;
}

SDB_HD float d_atan2f(float y, float x)
{
  const float ax = fabsf(x), ay = fabsf(y);
  float a = d_atanf_pos(ay / ax);
  a = ax == 0.0f ? 1.5707963267948966f : a;
  a = x < 0.0f ? 3.14159265358979323846f - a : a;
  a = y < 0.0f ? -a : a;
  return (ax == 0.0f && ay == 0.0f) ? 0.0f : a;
}

SDB_HD float d_log10f(float x)
{
  int e = 0;
  if (x < 1.17549435e-38f) { x = x * 8388608.0f; e = -23; }
  unsigned ix = sdb_f2u(x);
  e += (int) (ix >> 23) - 127;
  float m = sdb_u2f((ix & 0x007fffffu) | 0x3f800000u);
  if (m > 1.41421356f) { m = m * 0.5f; e += 1; }
  float f = m - 1.0f;
  float z = f * f;
  float y = ((((((((7.0376836292e-2f * f - 1.1514610310e-1f) * f + 1.1676998740e-1f) * f
              - 1.2420140846e-1f) * f + 1.4249322787e-1f) * f - 1.6668057665e-1f) * f
              + 2.0000714765e-1f) * f - 2.4999993993e-1f) * f + 3.3333331174e-1f) * f * z;
  y = y - 0.5f * z;
  float fe = (float) e;
  float r = y * 4.3429448190325176e-1f;
  r = r + f * 4.3429448190325176e-1f;
  r = r + fe * 3.0102999566398120e-1f;
  return r;
}

SDB_HD float d_exp10f(float x)
{
  if (x > 38.0f) x = 38.0f;
  if (x < -37.0f) x = -37.0f;
  float px = floorf(3.32192809488736234787f * x + 0.5f);
  int n = (int) px;
  x = x - px * 3.00781250000000000000e-1f;
  x = x - px * 2.48745663981195213739e-4f;
  float p = ((((2.063216740311022e-1f * x + 5.420251702225484e-1f) * x + 1.171292686296281f) * x
              + 2.034649854009453f) * x + 2.650948748208892f) * x + 2.302585167056758f;
  p = p * x + 1.0f;
  float sc = sdb_u2f((unsigned) (n + 127) << 23);
  return p * sc;
}

SDB_HD float d_db_to_mag(float db) { return d_exp10f(db * 0.05f); }
SDB_HD float d_cabsf(float re, float im) { return sqrtf(re * re + im * im); }

