// sdb_chain_steps.h -- the per-sample recurrences of the inspector chains and of the sigutils objects the
// reference's Tasks/ drive (SPEC.md sections N, I, A, C, G, D), written once for both sides:
//   * device: the role warps of k_inspectors and the Tasks kernels (chain_kernels.cu);
//   * host:   the per-sample entry points of the sigutils-named shim (sigutils_shim.cu: su_costas_feed,
//             su_pll_track, su_agc_feed, su_ncqo_read, su_clock_detector_feed ... are per-sample calls by ABI --
//             Tasks/CostasRecoveryTask.cpp:58-61 -- and cannot be kernel launches; their bulk twins run on the GPU).
// Every expression is a sequence of IEEE binary32 operations in source order (units are compiled with
// -fmad=false / -ffp-contract=off; the only fused operations are the explicit fmaf of SPEC I.1), so host and
// device results are bit-identical to each other and to the CPU oracle.
//
// Reference behaviour being replaced (all in suscan/sigutils, called from the reference at):
//   su_agc_feed            Tasks/AGCTask.cpp:70-73        su_costas_feed  Tasks/CostasRecoveryTask.cpp:58-61
//   su_pll_track           Tasks/PLLSyncTask.cpp:53-56    su_ncqo_read    Tasks/CarrierXlator.cpp:57-60
//   su_clock_detector_feed Tasks/WaveSampler.cpp:190-199  su_iir_filt_feed Tasks/WaveSampler.cpp:68-80
//   Decider                Default/GenericInspector/InspectorUI.cpp:836-846
#pragma once
#include "sdb_internal.h"
#include "sdb_math.h"
#define PI_F   3.14159265358979323846f
#define TWOPI_F 6.28318530717958647692f


SDB_HD float wrap_once(float phi)
{
  if (phi >= TWOPI_F) phi = phi - TWOPI_F;
  else if (phi < 0.0f) phi = phi + TWOPI_F;
  return phi;
}

// NCQO read: y = exp(i phi); phi <- wrap(phi + omega)   (SPEC N)
SDB_HD float2 ncqo_read(float &phi, float omega)
{
  float s, c;
  d_sincosf(phi, &s, &c);
  phi = wrap_once(phi + omega);
  return make_float2(c, s);
}

// ------------------------------------------------------------------ small IIR/FIR, shift registers --
// y[n] = sum_{i<N} b[i] x[n-i] - sum_{1<=i<N} a[i] y[n-i], single accumulator, ascending i, every term one fused
// multiply-add (SPEC I.1; explicit fma: the unit is compiled with -fmad=false).
// Lines are shift registers ([0] newest) so every index is a compile-time constant -> registers.
template <int N>
SDB_HD float2 iir_step(const float (&b)[SDB_MAX_IIR], const float (&a)[SDB_MAX_IIR],
                                                  float (&xr)[SDB_MAX_IIR], float (&xi)[SDB_MAX_IIR],
                                                  float (&yr)[SDB_MAX_IIR], float (&yi)[SDB_MAX_IIR], float2 in)
{
#pragma unroll
  for (int i = N - 1; i > 0; --i) { xr[i] = xr[i - 1]; xi[i] = xi[i - 1]; }
  xr[0] = in.x; xi[0] = in.y;
  float ar = 0.0f, ai = 0.0f;
#pragma unroll
  for (int i = 0; i < N; ++i) { ar = fmaf(b[i], xr[i], ar); ai = fmaf(b[i], xi[i], ai); }
  if (N > 1) {
#pragma unroll
    for (int i = 1; i < N; ++i) { ar = fmaf(-a[i], yr[i - 1], ar); ai = fmaf(-a[i], yi[i - 1], ai); }
#pragma unroll
    for (int i = N - 1; i > 0; --i) { yr[i] = yr[i - 1]; yi[i] = yi[i - 1]; }
    yr[0] = ar; yi[0] = ai;
  }
  return make_float2(ar, ai);
}

SDB_HD float2 iir_any(int n, const float (&b)[SDB_MAX_IIR], const float (&a)[SDB_MAX_IIR],
                                                 float (&xr)[SDB_MAX_IIR], float (&xi)[SDB_MAX_IIR],
                                                 float (&yr)[SDB_MAX_IIR], float (&yi)[SDB_MAX_IIR], float2 in)
{
  switch (n) {
    case 1:  return iir_step<1>(b, a, xr, xi, yr, yi, in);
    case 2:  return iir_step<2>(b, a, xr, xi, yr, yi, in);
    case 3:  return iir_step<3>(b, a, xr, xi, yr, yi, in);
    case 4:  return iir_step<4>(b, a, xr, xi, yr, yi, in);
    default: return iir_step<5>(b, a, xr, xi, yr, yi, in);
  }
}

SDB_HD float sgnf(float v) { return v < 0.0f ? -1.0f : (v > 0.0f ? 1.0f : 0.0f); }

// ------------------------------------------------------------------ per-block state, all in registers --
// AGC (SPEC A).  dl / mh are strided arrays (stride 32 floats: one column per lane).
struct AgcK {
  float knee, slope_m1, fixed_gain, far_, faf, sar, saf;
  unsigned hang_max, dl_size, mh_size;
};
struct AgcS { float fast, slow, peak; unsigned hang_n, dl_ptr, mh_ptr; };

// dl / mh: strided arrays (STRIDE floats between consecutive entries: 32 = one column per lane, 1 = a host object)
template <int STRIDE = 32>
SDB_HD float2 agc_step(const AgcK &k, AgcS &s, float *dl, float *mh, float2 x)
{
  float2 xd = make_float2(dl[(2 * s.dl_ptr) * STRIDE], dl[(2 * s.dl_ptr + 1) * STRIDE]);
  dl[(2 * s.dl_ptr) * STRIDE] = x.x; dl[(2 * s.dl_ptr + 1) * STRIDE] = x.y;
  if (++s.dl_ptr >= k.dl_size) s.dl_ptr = 0;
  float m = 10.0f * d_log10f(x.x * x.x + x.y * x.y + 1e-16f);
  float m_old = mh[s.mh_ptr * STRIDE];
  mh[s.mh_ptr * STRIDE] = m;
  if (++s.mh_ptr >= k.mh_size) s.mh_ptr = 0;
  if (m > s.peak) {
    s.peak = m;
  } else if (s.peak == m_old) {
    float pk = -160.0f;
    for (unsigned i = 0; i < k.mh_size; ++i) { float v = mh[i * STRIDE]; if (pk < v) pk = v; }
    s.peak = pk;
  }
  float d = s.peak - s.fast;
  if (d > 0.0f) s.fast = s.fast + k.far_ * d;
  else          s.fast = s.fast + k.faf * d;
  d = s.peak - s.slow;
  if (d > 0.0f) { s.slow = s.slow + k.sar * d; s.hang_n = 0; }
  else if (s.hang_n >= k.hang_max) s.slow = s.slow + k.saf * d;
  else ++s.hang_n;
  float lvl = s.fast > s.slow ? s.fast : s.slow;
  float g = lvl < k.knee ? k.fixed_gain : d_db_to_mag(lvl * k.slope_m1);
  g = g * 0.7f;
  xd.x = xd.x * g; xd.y = xd.y * g;
  return xd;
}

// The level tracker of agc_step (magnitude history, running peak, fast / slow levels) alone, with every per-lane
// `if` written as a select: the inspector kernel's track warp.  The chains of a warp disagree on these conditions at
// almost every sample, and a divergent branch costs the warp a reconvergence barrier plus an instruction-fetch bubble
// on either side (ncu: `no instruction` was the top stall of the recurrence warps).  Same operations on the same
// operands as agc_step, so the same bits.  Returns max(fast, slow).
template <int STRIDE = 32>
SDB_HD float agc_level_sel(float far_, float faf, float sar, float saf, unsigned hang_max, unsigned mh_size, AgcS &s,
                           float *mh, float m)
{
  const unsigned mp = s.mh_ptr;
  const float m_old = mh[mp * STRIDE];
  mh[mp * STRIDE] = m;
  s.mh_ptr = mp + 1 >= mh_size ? 0 : mp + 1;
  const bool rescan = !(m > s.peak) && s.peak == m_old;
  s.peak = m > s.peak ? m : s.peak;
  if (rescan) {                                          // the window's maximum left it: rare
    float pk = -160.0f;
    for (unsigned q = 0; q < mh_size; ++q) { float v = mh[q * STRIDE]; if (pk < v) pk = v; }
    s.peak = pk;
  }
  float d = s.peak - s.fast;
  s.fast = s.fast + (d > 0.0f ? far_ : faf) * d;
  d = s.peak - s.slow;
  const bool up = d > 0.0f, fall = !up && s.hang_n >= hang_max;
  const float slow_n = s.slow + (up ? sar : saf) * d;
  s.slow = (up || fall) ? slow_n : s.slow;
  s.hang_n = up ? 0u : (fall ? s.hang_n : s.hang_n + 1u);
  return s.fast > s.slow ? s.fast : s.slow;
}

struct CostasK { int kind, af_n; float a, b; float af_b[SDB_MAX_IIR], af_a[SDB_MAX_IIR]; };
struct CostasS { float phi, omega, lock, yre, yim; float xr[SDB_MAX_IIR], xi[SDB_MAX_IIR], yr[SDB_MAX_IIR], yi[SDB_MAX_IIR]; };

// KIND / AFN >= 0: loop kind and arm-filter length known at compile time (the inspector kernel's carrier warp, when
// every chain of the warp agrees); -1: read from k.  Same statements either way.
template <int KIND, int AFN>
SDB_HD float2 costas_step_t(const CostasK &k, CostasS &s, float2 x)
{
  const int kind = KIND >= 0 ? KIND : k.kind;
  float2 n = ncqo_read(s.phi, s.omega);
  float2 mixed = make_float2(x.x * n.x + x.y * n.y, x.y * n.x - x.x * n.y);
  float2 z = AFN >= 1 ? iir_step<(AFN >= 1 ? AFN : 1)>(k.af_b, k.af_a, s.xr, s.xi, s.yr, s.yi, mixed)
                      : iir_any(k.af_n, k.af_b, k.af_a, s.xr, s.xi, s.yr, s.yi, mixed);
  float e = 0.0f, lr, li;
  if (kind == 1) {
    e = -(z.x * z.y);
  } else if (kind == 2) {
    lr = sgnf(z.x); li = sgnf(z.y);
    e = lr * z.y - li * z.x;
  } else if (kind == 3) {
    lr = sgnf(z.x); li = sgnf(z.y);
    if (fabsf(z.x) >= fabsf(z.y)) e = lr * z.y - li * z.x * 0.41421356237309504f;
    else                          e = lr * z.y * 0.41421356237309504f - li * z.x;
  }
  s.lock = s.lock + k.a * (1.0f - e - s.lock);
  s.yre = s.yre + 1.0f * (z.x - s.yre);
  s.yim = s.yim + 1.0f * (z.y - s.yim);
  s.omega = s.omega + k.b * e;
  s.phi = wrap_once(s.phi + k.a * e);
  return make_float2(s.yre, s.yim);
}

SDB_HD float2 costas_step(const CostasK &k, CostasS &s, float2 x) { return costas_step_t<-1, -1>(k, s, x); }

SDB_HD float2 pll_step(float alpha, float beta, float &phi, float &omega, float2 x)
{
  float2 ref = ncqo_read(phi, omega);
  float2 mix = make_float2(x.x * ref.x + x.y * ref.y, x.y * ref.x - x.x * ref.y);
  float err = d_atan2f(x.y, x.x) - phi;
  if (err > PI_F) err = err - TWOPI_F;
  else if (err < -PI_F) err = err + TWOPI_F;
  omega = omega + alpha * err;
  phi = wrap_once(phi + beta * err);
  return mix;
}

// the same with the phase detector's angle atan2(x.im, x.re) supplied by the caller (it depends on the sample only)
SDB_HD float2 pll_step_ang(float alpha, float beta, float &phi, float &omega, float2 x, float ang)
{
  float2 ref = ncqo_read(phi, omega);
  float2 mix = make_float2(x.x * ref.x + x.y * ref.y, x.y * ref.x - x.x * ref.y);
  float err = ang - phi;
  if (err > PI_F) err = err - TWOPI_F;
  else if (err < -PI_F) err = err + TWOPI_F;
  omega = omega + alpha * err;
  phi = wrap_once(phi + beta * err);
  return mix;
}

struct ClockS { float phi, bnor, x0r, x0i, x1r, x1i, x2r, x2i, pr, pi; int half; };

SDB_HD bool clock_step(float gain, float alpha, float beta, ClockS &s, float2 v, float2 &out)
{
  bool produced = false;
  s.phi = s.phi + s.bnor;
  if (s.phi >= 0.5f) {
    float al = s.bnor * (s.phi - 0.5f);
    float om = 1.0f - al;
    float pr = om * v.x + al * s.pr;
    float pi = om * v.y + al * s.pi;
    s.half = !s.half;
    s.phi = s.phi - 0.5f;
    if (!s.half) {
      s.x2r = s.x0r; s.x2i = s.x0i;
      s.x0r = pr; s.x0i = pi;
      float dr = s.x0r - s.x2r;
      float di = s.x0i - s.x2i;
      float e = gain * (s.x1r * dr + s.x1i * di);
      s.phi = s.phi + alpha * e;
      float bn = s.bnor + beta * e;
      if (bn > 1.0f) bn = 1.0f;
      if (bn < 0.0f) bn = 0.0f;
      s.bnor = bn;
      out = make_float2(pr, pi);
      produced = true;
    } else {
      s.x1r = pr; s.x1i = pi;
    }
  }
  s.pr = v.x; s.pi = v.y;
  return produced;
}

// clock_step with every per-lane `if` written as a select (the inspector kernel's clock warp: its 32 chains cross
// the half-symbol mark at different samples, so a branching body is executed in full at almost every sample anyway,
// plus the reconvergence overhead).  Same operations on the same operands as clock_step: `phi - 0.5f` is evaluated
// once there too (for `al` and for the new phase), x0 - x2 after the shift is p - (old x0).
SDB_HD bool clock_step_sel(float gain, float alpha, float beta, ClockS &s, float2 v, float2 &out)
{
  const float phi1 = s.phi + s.bnor;
  const bool cross = phi1 >= 0.5f;
  const float ph = phi1 - 0.5f;
  const float al = s.bnor * ph, om = 1.0f - al;
  const float pr = om * v.x + al * s.pr, pi = om * v.y + al * s.pi;
  const bool full = cross && s.half != 0, halfc = cross && s.half == 0;
  const float dr = pr - s.x0r, di = pi - s.x0i;
  const float e = gain * (s.x1r * dr + s.x1i * di);
  const float phi_full = ph + alpha * e;
  float bn = s.bnor + beta * e;
  bn = bn > 1.0f ? 1.0f : bn;
  bn = bn < 0.0f ? 0.0f : bn;
  s.x2r = full ? s.x0r : s.x2r; s.x2i = full ? s.x0i : s.x2i;
  s.x0r = full ? pr : s.x0r;    s.x0i = full ? pi : s.x0i;
  s.x1r = halfc ? pr : s.x1r;   s.x1i = halfc ? pi : s.x1i;
  s.phi = cross ? (full ? phi_full : ph) : phi1;
  s.bnor = full ? bn : s.bnor;
  s.half = cross ? !s.half : s.half;
  s.pr = v.x; s.pi = v.y;
  out = make_float2(pr, pi);
  return full;
}

SDB_HD bool sampler_step(float period, float phase0, float &phase, float &pr, float &pi,
                                                    float2 v, float2 &out)
{
  bool sampled = false;
  if (period >= 1.0f) {
    phase = phase + 1.0f;
    if (phase >= period) phase = phase - period;
    float ph = phase - phase0;
    if (ph < 0.0f) ph = ph + period;
    float fl = floorf(ph);
    if (fl == 0.0f) {
      float al = ph - fl, om = 1.0f - al;
      out = make_float2(om * pr + al * v.x, om * pi + al * v.y);
      sampled = true;
    }
  }
  pr = v.x; pi = v.y;
  return sampled;
}

SDB_HD unsigned char decide(int mode, float dmin, float dh, int intervals, float2 x)
{
  float v = mode == 0 ? d_atan2f(x.y, x.x) : d_cabsf(x.x, x.y);
  float s = floorf((v - dmin) / dh * (float) intervals);
  int k = (int) s;
  if (!(s >= 0.0f)) k = 0;
  if (k > intervals - 1) k = intervals - 1;
  return (unsigned char) k;
}

