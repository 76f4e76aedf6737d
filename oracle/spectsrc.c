/* spectsrc.c -- ORACLE (test infrastructure): inspector spectrum sources and baud estimators, SPEC.md section U.
 *
 * What the reference shows: the GUI selects a source by index and toggles estimators by id
 * (Suscan/Analyzer.cpp:539-565 -> suscan_analyzer_inspector_set_spectrum_async / _estimator_cmd_async), receives
 * kind=SPECTRUM messages {spectsrc_id, spectrum_data[spectrum_size], samp_rate} that it converts to dB and
 * half-swaps itself (Default/GenericInspector/GenericInspector.cpp:232-250), and kind=ESTIMATOR messages
 * {estimator_id, value} (ibid. :262-264).  The class registries are looked up by name only
 * (Suscan/Messages/InspectorMessage.cpp:46,55): the inner definitions are upstream and NOT in the reference, so
 * the pre-transforms below are THIS project's definition (parity unpinned), except the fast autocorrelation,
 * which follows the in-repo GUI implementation Default/GenericInspector/FACTab.cpp:209-221
 * (FFT -> x conj(x) -> inverse FFT -> |.| of the first half).
 */
#include "sd_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

static sdo_cpx csq(sdo_cpx a)
{
  sdo_cpx r;
  const float m = a.re * a.im;
  r.re = a.re * a.re - a.im * a.im;
  r.im = m + m;
  return r;
}

/* U.2: pre-transform of sample x with predecessor p */
static sdo_cpx pretransform(int kind, sdo_cpx x, sdo_cpx p)
{
  sdo_cpx r = { 0.0f, 0.0f }, d;
  switch (kind) {
    case SDO_SPECTSRC_PSD:
    case SDO_SPECTSRC_FAC:
      return x;
    case SDO_SPECTSRC_CYCLO:
      r.re = x.re * p.re + x.im * p.im;
      r.im = x.im * p.re - x.re * p.im;
      return r;
    case SDO_SPECTSRC_FMSPECT:
      d.re = x.re * p.re + x.im * p.im;
      d.im = x.im * p.re - x.re * p.im;
      r.re = sdo_atan2f(d.im, d.re) * 0.318309886f;
      return r;
    case SDO_SPECTSRC_TIMEDIFF:
      r.re = x.re - p.re; r.im = x.im - p.im;
      return r;
    case SDO_SPECTSRC_ABSTIMEDIFF:
      d.re = x.re - p.re; d.im = x.im - p.im;
      r.re = sdo_cabsf(d);
      return r;
    case SDO_SPECTSRC_EXP_2: return csq(x);
    case SDO_SPECTSRC_EXP_4: return csq(csq(x));
    case SDO_SPECTSRC_EXP_8: return csq(csq(csq(x)));
    default: return r;
  }
}

unsigned sdo_spectsrc_out_size(int kind, unsigned ns) { return kind == SDO_SPECTSRC_FAC ? ns / 2 : ns; }

/* U.1-U.4.  c = the n_ch channel samples this feed produced for the channel.  Returns the number of floats
 * written to out (0 when n_ch < ns + 1: nothing is emitted for this feed). */
unsigned sdo_spectsrc_frame(int kind, unsigned ns, const sdo_cpx *c, size_t n_ch, float *out)
{
  sdo_cpx *s, *tmp, *tw;
  float *w;
  unsigned i;
  const float inv_n = 1.0f / (float) ns;
  if (ns < 64 || ns > 4096 || (ns & (ns - 1)) || n_ch < (size_t) ns + 1 || kind <= 0 || kind >= SDO_SPECTSRC_COUNT)
    return 0;
  s = (sdo_cpx *) malloc(sizeof(sdo_cpx) * ns);
  tmp = (sdo_cpx *) malloc(sizeof(sdo_cpx) * ns);
  w = (float *) malloc(sizeof(float) * ns);
  tw = sdo_spec_twiddles(ns);
  sdo_window_fill(w, ns, SDO_WINDOW_BLACKMANN_HARRIS);
  c += n_ch - ns;
  for (i = 0; i < ns; ++i) {
    s[i] = pretransform(kind, c[i], c[(long) i - 1]);
    if (kind != SDO_SPECTSRC_FAC) { s[i].re *= w[i]; s[i].im *= w[i]; }
  }
  sdo_spec_fft_stockham(s, ns, tw, -1, tmp);
  if (kind != SDO_SPECTSRC_FAC) {
    for (i = 0; i < ns; ++i) out[i] = fmaf(s[i].re, s[i].re, s[i].im * s[i].im) * inv_n;
  } else {
    for (i = 0; i < ns; ++i) { s[i].re = fmaf(s[i].re, s[i].re, s[i].im * s[i].im); s[i].im = 0.0f; }
    sdo_spec_fft_stockham(s, ns, tw, +1, tmp);
    for (i = 0; i < ns / 2; ++i) out[i] = sdo_cabsf(s[i]) * inv_n;
  }
  free(s); free(tmp); free(w); free(tw);
  return sdo_spectsrc_out_size(kind, ns);
}

/* U.5 baud estimators on the same frame.  Returns 1 and *baud when an estimate exists. */
int sdo_estimate_baud(int estimator, unsigned ns, float fs_ch, const sdo_cpx *c, size_t n_ch, float *baud)
{
  sdo_cpx *s, *tmp, *tw;
  float *w, *v;
  unsigned i, best = 0, from;
  int ok = 0;
  const float inv_n = 1.0f / (float) ns;
  if (ns < 64 || ns > 4096 || (ns & (ns - 1)) || n_ch < (size_t) ns + 1) return 0;
  s = (sdo_cpx *) malloc(sizeof(sdo_cpx) * ns);
  tmp = (sdo_cpx *) malloc(sizeof(sdo_cpx) * ns);
  w = (float *) malloc(sizeof(float) * ns);
  v = (float *) malloc(sizeof(float) * ns);
  tw = sdo_spec_twiddles(ns);
  sdo_window_fill(w, ns, SDO_WINDOW_BLACKMANN_HARRIS);
  c += n_ch - ns;
  if (estimator == SDO_ESTIMATOR_BAUD_NONLINEAR) {
    /* line at the symbol rate in the spectrum of |x[n] - x[n-1]|; bins below U5_KMIN hold the DC term's skirt */
    for (i = 0; i < ns; ++i) {
      s[i] = pretransform(SDO_SPECTSRC_ABSTIMEDIFF, c[i], c[(long) i - 1]);
      s[i].re *= w[i]; s[i].im *= w[i];
    }
    sdo_spec_fft_stockham(s, ns, tw, -1, tmp);
    for (i = 0; i < ns / 2; ++i) v[i] = fmaf(s[i].re, s[i].re, s[i].im * s[i].im) * inv_n;
    best = SDO_U5_KMIN;
    for (i = SDO_U5_KMIN; i < ns / 2; ++i) if (v[i] > v[best]) best = i;
    if (v[best] > 0.0f) { *baud = (float) best * fs_ch / (float) ns; ok = 1; }
  } else if (estimator == SDO_ESTIMATOR_BAUD_FAC) {
    /* autocorrelation of the mean-removed |x[n] - x[n-1]| (DC bin zeroed): first maximum after the first
     * zero crossing is one symbol period */
    for (i = 0; i < ns; ++i) s[i] = pretransform(SDO_SPECTSRC_ABSTIMEDIFF, c[i], c[(long) i - 1]);
    sdo_spec_fft_stockham(s, ns, tw, -1, tmp);
    for (i = 0; i < ns; ++i) { s[i].re = fmaf(s[i].re, s[i].re, s[i].im * s[i].im); s[i].im = 0.0f; }
    s[0].re = 0.0f;
    sdo_spec_fft_stockham(s, ns, tw, +1, tmp);
    for (i = 0; i < ns / 2; ++i) v[i] = s[i].re;
    from = ns / 2;
    for (i = 1; i < ns / 2; ++i) if (v[i] < 0.0f) { from = i; break; }
    if (from < ns / 2) {
      float thr;
      best = from;
      for (i = from; i < ns / 2; ++i) if (v[i] > v[best]) best = i;
      /* every multiple of the period peaks at about the same height: take the FIRST local maximum that
       * reaches half of the largest one */
      thr = 0.5f * v[best];
      if (v[best] > 0.0f)
        for (i = from; i + 1 < ns / 2; ++i)
          if (v[i] >= thr && v[i] >= v[i - 1] && v[i] >= v[i + 1]) { *baud = fs_ch / (float) i; ok = 1; break; }
    }
  }
  free(s); free(tmp); free(w); free(v); free(tw);
  return ok;
}
