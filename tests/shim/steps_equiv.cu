// Host-side equivalence check (no GPU needed; nvcc compiles SDB_HD functions for the host too): the select-form /
// compile-time variants the inspector kernel's recurrence warps run are bit-identical to the statement forms of
// sdb_chain_steps.h (which the shims call per sample and tests/test_shim_cpu.py holds to the oracle).
//   clock_step_sel vs clock_step, costas_step_t<K, A> vs costas_step, agc_level_sel vs the tracker inside agc_step,
//   d_atan2f (select form) vs the branching statement of SPEC M.2 restated here.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include "../../sigdigger_b200/csrc/sdb_chain_steps.h"

static unsigned long long rs = 88172645463325252ull;
static float frand() { rs ^= rs << 13; rs ^= rs >> 7; rs ^= rs << 17; return (float) ((rs >> 11) * (1.0 / 9007199254740992.0)); }
static bool same(float a, float b) { return !memcmp(&a, &b, 4); }

// SPEC M.2 as branching statements (the form sdb_math.h had before the select form)
static float ref_atan_pos(float t)
{
  float y0;
  if (t > 2.414213562373095f) { y0 = 1.5707963267948966f; t = -1.0f / t; }
  else if (t > 0.4142135623730950f) { y0 = 0.7853981633974483f; t = (t - 1.0f) / (t + 1.0f); }
  else y0 = 0.0f;
  float z = t * t;
  float y = (((8.05374449538e-2f * z - 1.38776856032e-1f) * z + 1.99777106478e-1f) * z - 3.33329491539e-1f) * z * t + t;
  return y0 + y;
}
static float ref_atan2(float y, float x)
{
  float ax = fabsf(x), ay = fabsf(y), a;
  if (ax == 0.0f && ay == 0.0f) return 0.0f;
  if (ax == 0.0f) a = 1.5707963267948966f;
  else a = ref_atan_pos(ay / ax);
  if (x < 0.0f) a = 3.14159265358979323846f - a;
  if (y < 0.0f) a = -a;
  return a;
}

template <int K, int A> static int costas_case(int n)
{
  CostasK k; CostasS a, b;
  k.kind = K; k.af_n = A; k.a = 6.2e-3f; k.b = 1.9e-5f;
  for (int i = 0; i < SDB_MAX_IIR; ++i) { k.af_b[i] = 0.1f + 0.05f * i; k.af_a[i] = i ? 0.07f * (i & 1 ? -1 : 1) : 1.0f; }
  memset(&a, 0, sizeof a); memset(&b, 0, sizeof b);
  int bad = 0;
  for (int i = 0; i < n; ++i) {
    float2 x = make_float2(2 * frand() - 1, 2 * frand() - 1);
    float2 u = costas_step(k, a, x), v = costas_step_t<K, A>(k, b, x);
    if (!same(u.x, v.x) || !same(u.y, v.y) || memcmp(&a, &b, sizeof a)) ++bad;
  }
  return bad;
}

int main()
{
  int bad = 0;
  // ---- atan2
  const float sp[] = { 0.0f, -0.0f, 1.0f, -1.0f, 1e-30f, -1e-30f, 1e30f, -1e30f, 0.4142135f, 0.4142136f, 2.4142134f, 2.4142137f };
  for (float y : sp) for (float x : sp) if (!same(d_atan2f(y, x), ref_atan2(y, x))) { ++bad; printf("atan2 %g %g\n", y, x); }
  for (int i = 0; i < 2000000; ++i) {
    float sc = ldexpf(1.0f, (int) (frand() * 40) - 20);
    float y = (2 * frand() - 1) * sc, x = (2 * frand() - 1) * (frand() < 0.5f ? sc : 1.0f);
    if (!same(d_atan2f(y, x), ref_atan2(y, x))) { if (++bad < 5) printf("atan2 %a %a\n", y, x); }
  }
  // ---- Gardner
  for (int trial = 0; trial < 200; ++trial) {
    ClockS a, b; memset(&a, 0, sizeof a);
    a.phi = 0.25f; a.bnor = trial < 4 ? (trial & 1 ? 1.0f : 0.0f) : 0.02f + 0.95f * frand();
    b = a;
    const float gain = 0.05f + frand(), alpha = 0.2f, beta = 1.2e-4f * (trial % 7 == 0 ? 500.0f : 1.0f);
    for (int i = 0; i < 20000; ++i) {
      float2 v = make_float2(2 * frand() - 1, 2 * frand() - 1), oa = make_float2(0, 0), ob = make_float2(0, 0);
      bool pa = clock_step(gain, alpha, beta, a, v, oa), pb = clock_step_sel(gain, alpha, beta, b, v, ob);
      if (pa != pb || (pa && (!same(oa.x, ob.x) || !same(oa.y, ob.y))) || memcmp(&a, &b, sizeof a)) { ++bad; break; }
    }
  }
  // ---- Costas
  bad += costas_case<1, 1>(50000) + costas_case<1, 3>(50000) + costas_case<2, 1>(50000) + costas_case<2, 3>(50000) +
         costas_case<3, 1>(50000) + costas_case<3, 3>(50000);
  // ---- AGC level tracker
  for (int trial = 0; trial < 60; ++trial) {
    AgcK k; k.knee = -100.0f; k.slope_m1 = -0.94f; k.fixed_gain = 1.0f;
    k.far_ = 0.3f * frand() + 0.01f; k.faf = 0.5f * k.far_; k.sar = 0.1f * k.far_; k.saf = 0.05f * k.far_;
    k.hang_max = (unsigned) (frand() * 20); k.dl_size = 1 + (unsigned) (frand() * 30); k.mh_size = 1 + (unsigned) (frand() * 40);
    AgcS a, b; a.fast = a.slow = a.peak = -160.0f; a.hang_n = a.dl_ptr = a.mh_ptr = 0; b = a;
    std::vector<float> dla(2 * k.dl_size, 0.0f), mha(k.mh_size, -160.0f), mhb(k.mh_size, -160.0f);
    for (int i = 0; i < 30000; ++i) {
      // magnitudes with repeats (so that peak == m_old happens with ties) and steps
      const float amp = (i / 700) % 3 == 0 ? 1.0f : 0.03f;
      float2 x = make_float2(amp * floorf(8 * frand()) / 8, amp * floorf(8 * frand()) / 8);
      float2 r = agc_step<1>(k, a, dla.data(), mha.data(), x);
      (void) r;
      const float m = 10.0f * d_log10f(x.x * x.x + x.y * x.y + 1e-16f);
      const float lvl = agc_level_sel<1>(k.far_, k.faf, k.sar, k.saf, k.hang_max, k.mh_size, b, mhb.data(), m);
      const float lvl_a = a.fast > a.slow ? a.fast : a.slow;
      if (!same(lvl, lvl_a) || !same(a.fast, b.fast) || !same(a.slow, b.slow) || !same(a.peak, b.peak) ||
          a.hang_n != b.hang_n || a.mh_ptr != b.mh_ptr || memcmp(mha.data(), mhb.data(), 4 * k.mh_size)) { ++bad; break; }
    }
  }
  printf("mismatches %d\n", bad);
  return bad != 0;
}
