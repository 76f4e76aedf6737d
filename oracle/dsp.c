/*
 * dsp.c -- ORACLE (test infrastructure). Per-sample primitives: NCQO, IIR/FIR, tap design, AGC, PLL,
 * Costas loop, Gardner clock detector, manual sampler, decider, quadrature demodulator.
 * SPEC.md sections N, I, A, C, G, D.  Every function cites the reference call site it serves.
 *
 * Compile with -ffp-contract=off: each line below is one rounding per operator, in the order written.
 */
#include "sd_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <complex.h>

static inline sdo_cpx cmul(sdo_cpx a, sdo_cpx b)
{
  sdo_cpx r;
  r.re = a.re * b.re - a.im * b.im;
  r.im = a.re * b.im + a.im * b.re;
  return r;
}
/* a * conj(b) */
static inline sdo_cpx cmulc(sdo_cpx a, sdo_cpx b)
{
  sdo_cpx r;
  r.re = a.re * b.re + a.im * b.im;
  r.im = a.im * b.re - a.re * b.im;
  return r;
}

/* ------------------------------------------------------------------ N: NCQO ----------------------
 * su_ncqo_init(&ncqo, -relFreq); su_ncqo_set_phase(&ncqo, -phase); dst = src * su_ncqo_read(&ncqo)
 * (Tasks/CarrierXlator.cpp:36-37,57-60).  omega = pi * fnor; read() returns exp(i phi) for the
 * CURRENT phase and then advances it, keeping phi in [0, 2 pi).
 */
static inline float wrap_once(float phi)
{
  if (phi >= SDO_2PI_F) phi = phi - SDO_2PI_F;
  else if (phi < 0.0f)  phi = phi + SDO_2PI_F;
  return phi;
}

void sdo_ncqo_init(sdo_ncqo *o, float fnor) { o->phi = 0.0f; o->omega = SDO_PI_F * fnor; }

void sdo_ncqo_set_phase(sdo_ncqo *o, float phi)
{
  o->phi = phi - SDO_2PI_F * floorf(phi / SDO_2PI_F);
  if (o->phi >= SDO_2PI_F) o->phi = 0.0f;  /* rounding guard */
}

sdo_cpx sdo_ncqo_read(sdo_ncqo *o)
{
  sdo_cpx y;
  sdo_sincosf(o->phi, &y.im, &y.re);
  o->phi = wrap_once(o->phi + o->omega);
  return y;
}

void sdo_ncqo_inc_phase(sdo_ncqo *o, float d)   { o->phi = wrap_once(o->phi + d); }
void sdo_ncqo_inc_angfreq(sdo_ncqo *o, float d) { o->omega = o->omega + d; }

void sdo_carrier_xlate(const sdo_cpx *x, sdo_cpx *y, size_t n, sdo_ncqo *o)
{
  size_t i;
  for (i = 0; i < n; ++i) y[i] = cmul(x[i], sdo_ncqo_read(o));
}

/* ------------------------------------------------------------------ I: filters ------------------- */
int sdo_filt_init(sdo_filt *f, unsigned na, const float *a, unsigned nb, const float *b)
{
  memset(f, 0, sizeof(*f));
  if (nb < 1 || nb > SDO_FILT_MAX_TAPS) return -1;
  f->nb = nb; f->na = na;
  f->b = (float *) malloc(sizeof(float) * nb);
  f->x = (sdo_cpx *) calloc(nb, sizeof(sdo_cpx));
  if (na > 0) {
    f->a = (float *) malloc(sizeof(float) * na);
    f->y = (sdo_cpx *) calloc(na, sizeof(sdo_cpx));
    memcpy(f->a, a, sizeof(float) * na);
  }
  memcpy(f->b, b, sizeof(float) * nb);
  f->gain = 1.0f;
  return 0;
}

void sdo_filt_free(sdo_filt *f)
{
  free(f->a); free(f->b); free(f->x); free(f->y);
  memset(f, 0, sizeof(*f));
}

/* y[n] = sum_{i=0}^{nb-1} b[i] x[n-i] - sum_{i=1}^{na-1} a[i] y[n-i]   (a[0] == 1), single
 * accumulator, ascending i, feed-forward part first, every term one FUSED multiply-add
 * acc <- fma(+-coef, sample, acc) (SPEC I.1; su_iir_filt_feed, Tasks/WaveSampler.cpp:68-80 -- upstream's
 * own summation is a volk dot product / a compiler-contracted loop, i.e. unspecified, so the restatement
 * picks the form the GPU executes in one FFMA per term). */
sdo_cpx sdo_filt_feed(sdo_filt *f, sdo_cpx x)
{
  unsigned i, p;
  sdo_cpx acc = { 0.0f, 0.0f };
  f->x[f->xp] = x;
  p = f->xp;
  for (i = 0; i < f->nb; ++i) {
    acc.re = fmaf(f->b[i], f->x[p].re, acc.re);
    acc.im = fmaf(f->b[i], f->x[p].im, acc.im);
    p = p == 0 ? f->nb - 1 : p - 1;
  }
  f->xp = f->xp + 1 == f->nb ? 0 : f->xp + 1;
  if (f->na > 1) {
    /* y line holds y[n-1] at index yp, y[n-2] at yp-1 ... */
    p = f->yp;
    for (i = 1; i < f->na; ++i) {
      acc.re = fmaf(-f->a[i], f->y[p].re, acc.re);
      acc.im = fmaf(-f->a[i], f->y[p].im, acc.im);
      p = p == 0 ? f->na - 1 : p - 1;
    }
    f->yp = f->yp + 1 == f->na ? 0 : f->yp + 1;
    f->y[f->yp] = acc;
  }
  return acc;
}

/* su_iir_rrc_init(filt, n, T, beta): root-raised-cosine taps, Hamming-windowed (manual p.61),
 * unity DC gain scaling 1/T.  Evaluated in double, rounded to float once. */
void sdo_taps_rrc(float *h, unsigned n, float T, float beta)
{
  unsigned i;
  const double b = beta, Td = T;
  for (i = 0; i < n; ++i) {
    double t = ((double) i - (double) n / 2.0) / Td;
    double f = 4.0 * b * t;
    double dem = SDO_PI * t * (1.0 - f * f);
    double num = sin(SDO_PI * t * (1.0 - b)) + 4.0 * b * t * cos(SDO_PI * t * (1.0 + b));
    double v;
    if (fabs(t) < 1e-9)
      v = 1.0 - b + 4.0 * b / SDO_PI;
    else if (fabs(dem) < 1e-9)
      v = b / sqrt(2.0) * ((1.0 + 2.0 / SDO_PI) * sin(SDO_PI / (4.0 * b))
                           + (1.0 - 2.0 / SDO_PI) * cos(SDO_PI / (4.0 * b)));
    else
      v = num / dem;
    v /= Td;
    if (n > 1)
      v *= 0.54 - 0.46 * cos(2.0 * SDO_PI * (double) i / (double) (n - 1));
    h[i] = (float) v;
  }
}

/* su_taps_brickwall_lp_init: h[i] = fc sinc(fc (i - n/2)) x Hamming. */
void sdo_taps_brickwall_lp(float *h, unsigned n, float fc)
{
  unsigned i;
  for (i = 0; i < n; ++i) {
    double t = (double) i - (double) (n >> 1);
    double xx = SDO_PI * (double) fc * t;
    double v = fabs(xx) < 1e-12 ? (double) fc : (double) fc * sin(xx) / xx;
    if (n > 1)
      v *= 0.54 - 0.46 * cos(2.0 * SDO_PI * (double) i / (double) (n - 1));
    h[i] = (float) v;
  }
}

unsigned sdo_mf_span(float T)
{
  double s = ceil(6.0 * (double) T);   /* include/WaveSampler.h:29-30: 6 symbol periods, <= 1024 */
  if (s < 1.0) s = 1.0;
  if (s > 1024.0) s = 1024.0;
  return (unsigned) s;
}

/* Butterworth low-pass of given order, cut-off fc as a fraction of Nyquist, bilinear transform
 * (su_iir_bwlpf_init; the Costas arm filter of "order 3" is this with order 2,
 * Tasks/CostasRecoveryTask.cpp:41).  Double precision, rounded to float once. */
int sdo_butter_lp(unsigned order, float fc, float *b, float *a)
{
  double complex pz[16], pa[17], pb[17];
  double warped, gain;
  unsigned k, j;
  if (order < 1 || order > 16 || !(fc > 0.0f) || !(fc < 1.0f)) return -1;
  warped = 4.0 * tan(SDO_PI * (double) fc / 2.0);
  double complex kden = 1.0;
  for (k = 0; k < order; ++k) {
    double th = SDO_PI * (2.0 * k + order + 1.0) / (2.0 * order);
    double complex s = warped * (cos(th) + I * sin(th));
    pz[k] = (4.0 + s) / (4.0 - s);
    kden *= (4.0 - s);
  }
  gain = pow(warped, (double) order) * creal(1.0 / kden);
  /* expand prod (z - pz[k]) and (z + 1)^order, descending powers of z */
  for (k = 0; k <= order; ++k) { pa[k] = 0; pb[k] = 0; }
  pa[0] = 1.0; pb[0] = 1.0;
  for (k = 0; k < order; ++k) {
    for (j = k + 1; j >= 1; --j) {
      pa[j] = pa[j] - pz[k] * pa[j - 1];
      pb[j] = pb[j] + pb[j - 1];
    }
  }
  for (k = 0; k <= order; ++k) {
    a[k] = (float) creal(pa[k]);
    b[k] = (float) (gain * creal(pb[k]));
  }
  return 0;
}

/* ------------------------------------------------------------------ A: AGC -----------------------
 * Field names from Tasks/AGCTask.cpp:43-47; time-constant fractions Tasks/AGCTask.cpp:22-28. */
#define SDO_MIN_REF_DB (-160.0f)

void sdo_agc_params_default(sdo_agc_params *p)
{
  p->threshold = -100.0f; p->slope_factor = 6.0f;
  p->hang_max = 100; p->delay_line_size = 20; p->mag_history_size = 20;
  p->fast_rise_t = 2.0f; p->fast_fall_t = 4.0f; p->slow_rise_t = 20.0f; p->slow_fall_t = 40.0f;
}

/* frac_scale = 1 for the inspectors, 2 for SigDigger's AGCTask (Tasks/AGCTask.cpp:22 doubles). */
void sdo_agc_params_from_tau(sdo_agc_params *p, float tau, float frac_scale)
{
  const float rise = frac_scale * 3.9062e-1f;
  sdo_agc_params_default(p);
  p->fast_rise_t = tau * rise;
  p->fast_fall_t = tau * (2.0f * rise);
  p->slow_rise_t = tau * (10.0f * rise);
  p->slow_fall_t = tau * (10.0f * (2.0f * rise));
  p->hang_max = (unsigned) (tau * (rise * 5.0f));
  p->delay_line_size = (unsigned) (tau * (rise * 10.0f));
  p->mag_history_size = (unsigned) (tau * (rise * 10.0f));
  if (p->delay_line_size < 1) p->delay_line_size = 1;
  if (p->mag_history_size < 1) p->mag_history_size = 1;
  if (p->delay_line_size > SDO_AGC_MAX_HISTORY) p->delay_line_size = SDO_AGC_MAX_HISTORY;
  if (p->mag_history_size > SDO_AGC_MAX_HISTORY) p->mag_history_size = SDO_AGC_MAX_HISTORY;
}

static float alpha_of(float t) { return (float) (1.0 - exp(-1.0 / (double) t)); }

int sdo_agc_init(sdo_agc *a, const sdo_agc_params *p)
{
  unsigned i;
  memset(a, 0, sizeof(*a));
  if (p->delay_line_size < 1 || p->mag_history_size < 1) return -1;
  a->delay_line = (sdo_cpx *) calloc(p->delay_line_size, sizeof(sdo_cpx));
  a->mag_history = (float *) malloc(sizeof(float) * p->mag_history_size);
  if (!a->delay_line || !a->mag_history) return -1;
  for (i = 0; i < p->mag_history_size; ++i) a->mag_history[i] = SDO_MIN_REF_DB;
  a->delay_line_size = p->delay_line_size;
  a->mag_history_size = p->mag_history_size;
  a->knee = p->threshold;
  a->gain_slope = p->slope_factor * 1e-2f;
  a->fixed_gain = sdo_db_to_mag(a->knee * (a->gain_slope - 1.0f));
  a->hang_max = p->hang_max;
  a->fast_alpha_rise = alpha_of(p->fast_rise_t);
  a->fast_alpha_fall = alpha_of(p->fast_fall_t);
  a->slow_alpha_rise = alpha_of(p->slow_rise_t);
  a->slow_alpha_fall = alpha_of(p->slow_fall_t);
  a->fast_level = a->slow_level = a->peak = SDO_MIN_REF_DB;
  a->enabled = 1;
  return 0;
}

void sdo_agc_free(sdo_agc *a) { free(a->delay_line); free(a->mag_history); memset(a, 0, sizeof(*a)); }

sdo_cpx sdo_agc_feed(sdo_agc *a, sdo_cpx x)
{
  unsigned i;
  sdo_cpx xd = a->delay_line[a->delay_line_ptr];
  a->delay_line[a->delay_line_ptr] = x;
  if (++a->delay_line_ptr >= a->delay_line_size) a->delay_line_ptr = 0;

  if (a->enabled) {
    float m = 10.0f * sdo_log10f(x.re * x.re + x.im * x.im + 1e-16f);
    float m_old = a->mag_history[a->mag_history_ptr];
    float d, lvl, g;
    a->mag_history[a->mag_history_ptr] = m;
    if (++a->mag_history_ptr >= a->mag_history_size) a->mag_history_ptr = 0;

    if (m > a->peak) {
      a->peak = m;
    } else if (a->peak == m_old) {
      a->peak = SDO_MIN_REF_DB;
      for (i = 0; i < a->mag_history_size; ++i)
        if (a->peak < a->mag_history[i]) a->peak = a->mag_history[i];
    }

    d = a->peak - a->fast_level;
    if (d > 0.0f) a->fast_level = a->fast_level + a->fast_alpha_rise * d;
    else          a->fast_level = a->fast_level + a->fast_alpha_fall * d;

    d = a->peak - a->slow_level;
    if (d > 0.0f) {
      a->slow_level = a->slow_level + a->slow_alpha_rise * d;
      a->hang_n = 0;
    } else if (a->hang_n >= a->hang_max) {
      a->slow_level = a->slow_level + a->slow_alpha_fall * d;
    } else {
      ++a->hang_n;
    }

    lvl = a->fast_level > a->slow_level ? a->fast_level : a->slow_level;
    g = lvl < a->knee ? a->fixed_gain : sdo_db_to_mag(lvl * (a->gain_slope - 1.0f));
    g = g * 0.7f;
    xd.re = xd.re * g;
    xd.im = xd.im * g;
  }
  return xd;
}

/* ------------------------------------------------------------------ C: PLL / Costas -------------- */
void sdo_pll_init(sdo_pll *p, float fhint, float fc)
{
  float dinv;
  fc = SDO_PI_F * fc;
  dinv = 1.0f / (1.0f + 2.0f * 0.707f * fc + fc * fc);
  p->alpha = 4.0f * fc * fc * dinv;
  p->beta = 4.0f * 0.707f * fc * dinv;
  sdo_ncqo_init(&p->ncqo, fhint);
}

/* su_pll_track (Tasks/PLLSyncTask.cpp:53-56). */
sdo_cpx sdo_pll_track(sdo_pll *p, sdo_cpx x)
{
  sdo_cpx ref = sdo_ncqo_read(&p->ncqo);
  sdo_cpx mix = cmulc(x, ref);
  float err = sdo_atan2f(x.im, x.re) - p->ncqo.phi;
  if (err > SDO_PI_F) err = err - SDO_2PI_F;
  else if (err < -SDO_PI_F) err = err + SDO_2PI_F;
  sdo_ncqo_inc_angfreq(&p->ncqo, p->alpha * err);
  sdo_ncqo_inc_phase(&p->ncqo, p->beta * err);
  return mix;
}

/* su_costas_init(&costas, kind, fhint=0, arm_bw=1/tau, arm_order=3, loop_bw)
 * (Tasks/CostasRecoveryTask.cpp:36-41). */
int sdo_costas_init(sdo_costas *c, int kind, float fhint, float arm_bw, unsigned arm_order,
                    float loop_bw)
{
  float b[SDO_FILT_MAX_TAPS], a[17];
  memset(c, 0, sizeof(*c));
  c->kind = kind;
  c->a = SDO_PI_F * loop_bw;
  c->b = 0.5f * c->a * c->a;
  c->y_alpha = 1.0f;
  c->gain = 1.0f;
  sdo_ncqo_init(&c->ncqo, fhint);
  if (arm_order == 0) arm_order = 1;
  if (arm_order == 1) {
    b[0] = 1.0f;
    return sdo_filt_init(&c->af, 0, NULL, 1, b);
  }
  if (arm_order >= 20) {
    if (arm_order > SDO_FILT_MAX_TAPS) return -1;
    sdo_taps_brickwall_lp(b, arm_order, arm_bw);
    return sdo_filt_init(&c->af, 0, NULL, arm_order, b);
  }
  if (sdo_butter_lp(arm_order - 1, arm_bw, b, a) != 0) return -1;
  return sdo_filt_init(&c->af, arm_order, a, arm_order, b);
}

void sdo_costas_free(sdo_costas *c) { sdo_filt_free(&c->af); }

static inline float sgnf(float v) { return v < 0.0f ? -1.0f : (v > 0.0f ? 1.0f : 0.0f); }

/* su_costas_feed (Tasks/CostasRecoveryTask.cpp:58-61). */
sdo_cpx sdo_costas_feed(sdo_costas *c, sdo_cpx x)
{
  sdo_cpx s = sdo_ncqo_read(&c->ncqo);
  sdo_cpx z = sdo_filt_feed(&c->af, cmulc(x, s));
  float e = 0.0f, lr, li;
  z.re = c->gain * z.re; z.im = c->gain * z.im;
  c->z = z;
  switch (c->kind) {
    case SDO_COSTAS_BPSK:
      e = -(z.re * z.im);
      break;
    case SDO_COSTAS_QPSK:
      lr = sgnf(z.re); li = sgnf(z.im);
      e = lr * z.im - li * z.re;
      break;
    case SDO_COSTAS_8PSK:
      lr = sgnf(z.re); li = sgnf(z.im);
      if (fabsf(z.re) >= fabsf(z.im))
        e = lr * z.im - li * z.re * 0.41421356237309504f;
      else
        e = lr * z.im * 0.41421356237309504f - li * z.re;
      break;
    default:
      break;
  }
  c->lock = c->lock + c->a * (1.0f - e - c->lock);
  c->y.re = c->y.re + c->y_alpha * (z.re - c->y.re);
  c->y.im = c->y.im + c->y_alpha * (z.im - c->y.im);
  sdo_ncqo_inc_angfreq(&c->ncqo, c->b * e);
  sdo_ncqo_inc_phase(&c->ncqo, c->a * e);
  return c->y;
}

/* ------------------------------------------------------------------ G: clock recovery ------------
 * su_clock_detector_init(&cd, loopGain, bnor, bufsiz) / feed / read
 * (Tasks/WaveSampler.cpp:60-66,190-205). */
void sdo_clock_init(sdo_clock *c, float loop_gain, float bhint)
{
  memset(c, 0, sizeof(*c));
  c->alpha = 2e-1f;
  c->beta = 6e-4f * c->alpha;
  c->bnor = bhint;
  c->bmin = 0.0f;
  c->bmax = 1.0f;
  c->phi = 0.25f;
  c->gain = loop_gain;
}

int sdo_clock_feed(sdo_clock *c, sdo_cpx v, sdo_cpx *out)
{
  int produced = 0;
  c->phi = c->phi + c->bnor;
  if (c->phi >= 0.5f) {
    float al = c->bnor * (c->phi - 0.5f);
    float om = 1.0f - al;
    sdo_cpx p;
    p.re = om * v.re + al * c->prev.re;
    p.im = om * v.im + al * c->prev.im;
    c->halfcycle = !c->halfcycle;
    c->phi = c->phi - 0.5f;
    if (!c->halfcycle) {
      float dr, di, e, bn;
      c->x[2] = c->x[0];
      c->x[0] = p;
      dr = c->x[0].re - c->x[2].re;
      di = c->x[0].im - c->x[2].im;
      /* Re{conj(x1) * d} = x1.re*d.re + x1.im*d.im */
      e = c->gain * (c->x[1].re * dr + c->x[1].im * di);
      c->e = e;
      c->phi = c->phi + c->alpha * e;
      bn = c->bnor + c->beta * e;
      if (bn > c->bmax) bn = c->bmax;
      if (bn < c->bmin) bn = c->bmin;
      c->bnor = bn;
      *out = p;
      produced = 1;
    } else {
      c->x[1] = p;
    }
  }
  c->prev = v;
  return produced;
}

/* Manual sampler: clock.type = MANUAL, clock.baud, clock.phase
 * (Default/GenericInspector/InspectorCtl/ClockRecovery.cpp:59-93). */
void sdo_sampler_init(sdo_sampler *s, float bnor)
{
  memset(s, 0, sizeof(*s));
  s->bnor = bnor;
  s->period = bnor > 0.0f ? 1.0f / bnor : 0.0f;
}

void sdo_sampler_set_phase(sdo_sampler *s, float phase0_rel)
{
  s->phase0_rel = phase0_rel;
  s->phase0 = phase0_rel * s->period;
}

int sdo_sampler_feed(sdo_sampler *s, sdo_cpx v, sdo_cpx *out)
{
  int sampled = 0;
  if (s->period >= 1.0f) {
    float ph, fl;
    s->phase = s->phase + 1.0f;
    if (s->phase >= s->period) s->phase = s->phase - s->period;
    ph = s->phase - s->phase0;
    if (ph < 0.0f) ph = ph + s->period;
    fl = floorf(ph);
    if (fl == 0.0f) {
      float al = ph - fl, om = 1.0f - al;
      out->re = om * s->prev.re + al * v.re;
      out->im = om * s->prev.im + al * v.im;
      sampled = 1;
    }
  }
  s->prev = v;
  return sampled;
}

/* ------------------------------------------------------------------ E: CMA equaliser -------------
 * equalizer.type / equalizer.rate / equalizer.locked (Default/GenericInspector/InspectorCtl/
 * EqualizerControl.cpp:56-75); constant-modulus update on the symbol-rate stream, 10 complex weights. */
void sdo_equalizer_init(sdo_equalizer *e, float mu, int locked)
{
  memset(e, 0, sizeof(*e));
  e->mu = mu; e->locked = locked;
  e->w[0].re = 1.0f;
}

sdo_cpx sdo_equalizer_feed(sdo_equalizer *e, sdo_cpx x)
{
  int i;
  sdo_cpx y = { 0.0f, 0.0f };
  for (i = SDO_EQ_LEN - 1; i > 0; --i) e->x[i] = e->x[i - 1];
  e->x[0] = x;
  for (i = 0; i < SDO_EQ_LEN; ++i) {
    y.re = y.re + (e->w[i].re * e->x[i].re - e->w[i].im * e->x[i].im);
    y.im = y.im + (e->w[i].re * e->x[i].im + e->w[i].im * e->x[i].re);
  }
  if (!e->locked) {
    float y2 = y.re * y.re + y.im * y.im;
    float er = y.re * (y2 - 1.0f), ei = y.im * (y2 - 1.0f);
    for (i = 0; i < SDO_EQ_LEN; ++i) {
      /* w -= mu * conj(x) * err */
      float gr = e->x[i].re * er + e->x[i].im * ei;
      float gi = e->x[i].re * ei - e->x[i].im * er;
      e->w[i].re = e->w[i].re - e->mu * gr;
      e->w[i].im = e->w[i].im - e->mu * gi;
    }
  }
  return y;
}

/* ------------------------------------------------------------------ D: decider -------------------
 * ARGUMENT on [-pi, pi] for psk ("afc" prefix) and fsk, MODULUS on [0,1] for ask
 * (Default/GenericInspector/InspectorUI.cpp:228-253). */
void sdo_decider_init(sdo_decider *d, int mode, unsigned bps, float min, float max)
{
  d->mode = mode; d->bps = bps; d->intervals = 1u << bps; d->min = min; d->max = max;
  d->h = max - min;
}

void sdo_decider_decide(const sdo_decider *d, const sdo_cpx *x, uint8_t *sym, size_t n)
{
  size_t i;
  const float fi = (float) d->intervals;
  for (i = 0; i < n; ++i) {
    float v = d->mode == SDO_DECIDE_ARGUMENT ? sdo_atan2f(x[i].im, x[i].re) : sdo_cabsf(x[i]);
    float s = floorf((v - d->min) / d->h * fi);
    int k = (int) s;
    if (!(s >= 0.0f)) k = 0;
    if (k > (int) d->intervals - 1) k = (int) d->intervals - 1;
    sym[i] = (uint8_t) k;
  }
}

/* ------------------------------------------------------------------ quadrature demod -------------
 * dst[0] = 0; dst[p] = i * (1/pi) * arg(x[p] * conj(x[p-1]))  (Tasks/QuadDemodTask.cpp:44-60). */
void sdo_quad_demod(const sdo_cpx *x, sdo_cpx *y, size_t n, sdo_cpx *prev, int *primed)
{
  size_t i;
  const float k = (float) (1.0 / SDO_PI);
  for (i = 0; i < n; ++i) {
    if (!*primed) {
      y[i].re = 0.0f; y[i].im = 0.0f;
      *primed = 1;
    } else {
      sdo_cpx d = cmulc(x[i], *prev);
      y[i].re = 0.0f;
      y[i].im = k * sdo_atan2f(d.im, d.re);
    }
    *prev = x[i];
  }
}
