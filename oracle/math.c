/*
 * math.c -- ORACLE (test infrastructure). Deterministic float32 elementary functions, SPEC.md section M.
 *
 * sigutils evaluates SU_SIN/SU_COS/SU_LOG/... through libm (macros listed in SURVEY.md Appendix B,
 * used e.g. at Tasks/QuadDemodTask.cpp:53 SU_C_ARG, Suscan/Messages/PSDMessage.cpp:36 SU_POWER_DB).
 * libm results are not reproducible across CPU/GPU, so SPEC section M replaces them by fixed
 * Cody-Waite reductions + fixed polynomials (classic Cephes single-precision coefficient sets)
 * evaluated with plain binary32 multiply/add in the order written here.  Max error vs. the exact
 * function is a few ulp (checked in tests/test_oracle_math.py against libm in double).
 *
 * Must be compiled with -ffp-contract=off.
 */
#include "sd_oracle.h"
#include <math.h>
#include <string.h>

static inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

/* M.1 sincos: q = rint(x * 2/pi); r = ((x - q*C1) - q*C2) - q*C3; degree-7/8 polynomials. */
void sdo_sincosf(float x, float *s, float *c)
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
  switch (n) {
    case 0:  *s =  sp; *c =  cp; break;
    case 1:  *s =  cp; *c = -sp; break;
    case 2:  *s = -sp; *c = -cp; break;
    default: *s = -cp; *c =  sp; break;
  }
}

/* M.2 atan on t >= 0 with the two Cephes range reductions. */
static float atanf_pos(float t)
{
  float y0;
  if (t > 2.414213562373095f) {
    y0 = 1.5707963267948966f;
    t = -1.0f / t;
  } else if (t > 0.4142135623730950f) {
    y0 = 0.7853981633974483f;
    t = (t - 1.0f) / (t + 1.0f);
  } else {
    y0 = 0.0f;
  }
  float z = t * t;
  float y = (((8.05374449538e-2f * z - 1.38776856032e-1f) * z + 1.99777106478e-1f) * z
             - 3.33329491539e-1f) * z * t + t;
  return y0 + y;
}

float sdo_atan2f(float y, float x)
{
  float ax = fabsf(x), ay = fabsf(y), a;
  if (ax == 0.0f && ay == 0.0f)
    return 0.0f;
  if (ax == 0.0f)
    a = 1.5707963267948966f;
  else
    a = atanf_pos(ay / ax);
  if (x < 0.0f)
    a = 3.14159265358979323846f - a;
  if (y < 0.0f)
    a = -a;
  return a;
}

/* M.3 log10 for x > 0. */
float sdo_log10f(float x)
{
  int e = 0;
  if (x < 1.17549435e-38f) { x = x * 8388608.0f; e = -23; }
  uint32_t ix = f2u(x);
  e += (int) (ix >> 23) - 127;
  float m = u2f((ix & 0x007fffffu) | 0x3f800000u);
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

/* M.4 10^x. */
float sdo_exp10f(float x)
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
  /* ldexp(p, n), n in [-126, 127] after the clamp above */
  float sc = u2f((uint32_t) (n + 127) << 23);
  return p * sc;
}

float sdo_cabsf(sdo_cpx z) { return sqrtf(z.re * z.re + z.im * z.im); }

float sdo_db_to_mag(float db) { return sdo_exp10f(db * 0.05f); }

/* M.6 SU_POWER_DB(p) = 10 log10(p + 1e-8) (floor -80 dB). */
float sdo_power_db(float p) { return 10.0f * sdo_log10f(p + 1e-8f); }
