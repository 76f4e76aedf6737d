/* tasks.c -- ORACLE (test infrastructure only; the product never links this).
 *
 * The GUI's offline TimeWindow tasks whose arithmetic is fully present in the reference, restated loop by
 * loop (SPEC.md section Y).  These are sequential CPU loops over one captured buffer; the restatement keeps
 * their evaluation order, their 4096-sample work() blocks where the blocks are visible in the results, and
 * their float / double mix.  Elementary functions are the SPEC M ones.
 *
 *   sdo_delayed_conj            Tasks/DelayedConjTask.cpp:58-100
 *   sdo_histogram_feed          Tasks/HistogramFeeder.cpp:35-87
 *   sdo_sample_manual           Tasks/WaveSampler.cpp:28-46 (set-up), 96-175 (loop)
 *   sdo_sample_zero_crossing    Tasks/WaveSampler.cpp:222-292
 *   sdo_carrier_detect          Tasks/CarrierDetector.cpp:49-147
 */
#include "sd_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define TASK_BLOCK 4096 /* SIGDIGGER_WAVESAMPLER_FEEDER_BLOCK_LENGTH, include/WaveSampler.h:28 */

/* y[p] = 0 for p < delay, else k * x[p] * conj(x[p - delay]), k = 1 / (|x[p - delay]| + 1e-3) (double) */
void sdo_delayed_conj(const sdo_cpx *x, sdo_cpx *y, size_t n, size_t delay)
{
  size_t p;
  for (p = 0; p < n; ++p) {
    if (p < delay) {
      y[p].re = 0.0f; y[p].im = 0.0f;
    } else {
      const sdo_cpx prev = x[p - delay];
      const float kinv = (float) (1.0 / ((double) sdo_cabsf(prev) + 1e-3));
      const float tr = kinv * x[p].re, ti = kinv * x[p].im;
      y[p].re = tr * prev.re + ti * prev.im;
      y[p].im = ti * prev.re - tr * prev.im;
    }
  }
}

/* decision-space values the histogram is fed with; returns how many (n, or n - 1 for FREQUENCY) */
size_t sdo_histogram_feed(const sdo_cpx *x, float *out, size_t n, int space)
{
  size_t p, q = 0;
  switch (space) {
    case SDO_SPACE_AMPLITUDE:
      for (p = 0; p < n; ++p) out[q++] = sdo_cabsf(x[p]);
      break;
    case SDO_SPACE_PHASE:
      for (p = 0; p < n; ++p) out[q++] = sdo_atan2f(x[p].im, x[p].re);
      break;
    default:
      for (p = 1; p < n; ++p) {
        const float dr = x[p].re * x[p - 1].re + x[p].im * x[p - 1].im;
        const float di = x[p].im * x[p - 1].re - x[p].re * x[p - 1].im;
        out[q++] = sdo_atan2f(di, dr);
      }
      break;
  }
  return q;
}

/* Manual sampler: box-car over every symbol period with fractional edge weights.  out: (long) symbol_count
 * values; the return value is that count. */
size_t sdo_sample_manual(const sdo_cpx *x, size_t n, int space, size_t symbol_sync, double symbol_count,
                         sdo_cpx *out)
{
  const double delta = (double) n / symbol_count;
  const double samp_offset = (double) symbol_sync / delta;
  const float delta_inv = 1.f / (float) delta;
  const long count = (long) symbol_count;
  sdo_cpx prev = { 0.0f, 0.0f };
  long p;
  for (p = 0; p < count; ++p) {
    const double start = ((double) p - samp_offset) * delta + (double) symbol_sync;
    const double end = start + delta;
    const long long i_start = (long long) floor(start), i_end = (long long) ceil(end);
    const float t_start = (float) (1 - (start - (double) i_start));
    const float t_end = (float) (1 - ((double) i_end - end));
    sdo_cpx avg = { 0.0f, 0.0f };
    long long i;
    for (i = i_start; i <= i_end; ++i) {
      sdo_cpx v = { 0.0f, 0.0f };
      if (i >= 0 && i < (long long) n) {
        if (i == i_start) { v.re = t_start * x[i].re; v.im = t_start * x[i].im; }
        else if (i == i_end) { v.re = t_end * x[i].re; v.im = t_end * x[i].im; }
        else v = x[i];
      }
      if (space == SDO_SPACE_AMPLITUDE) {
        avg.re += v.re * v.re + v.im * v.im;
        avg.im += v.im * v.re - v.re * v.im;
      } else {
        avg.re += v.re * prev.re + v.im * prev.im;
        avg.im += v.im * prev.re - v.re * prev.im;
      }
      prev = v;
    }
    if (space == SDO_SPACE_AMPLITUDE) { out[p].re = sqrtf(delta_inv * avg.re); out[p].im = 0.0f; }
    else { out[p].re = delta_inv * avg.re; out[p].im = delta_inv * avg.im; }
  }
  return count > 0 ? (size_t) count : 0;
}

/* Zero-crossing sampler.  Every work() call handles one 4096-sample block and starts from the object's
 * initial prevVar (-1) and prevSample (0) because the reference never stores them back; only lastZc
 * survives.  Inside the last block every sample counts as a crossing.  At most 4096 symbols per block. */
size_t sdo_sample_zero_crossing(const sdo_cpx *x, size_t n, int space, int amplitude, sdo_cpx threshold,
                                sdo_cpx zc_angle, float bnor, uint8_t *sym, size_t cap)
{
  size_t total = 0;
  long p = 0, last_zc = 0;
  float thres;
  if (amplitude) thres = threshold.re * threshold.re + threshold.im * threshold.im;
  else thres = threshold.re * zc_angle.re - threshold.im * zc_angle.im;
  while (p < (long) n) {
    long amount = (long) n - p, i = 0;
    sdo_cpx prev = { 0.0f, 0.0f };
    float var = 0.0f, prev_var = -1.0f;
    int last;
    if (amount > TASK_BLOCK) amount = TASK_BLOCK;
    last = p + amount >= (long) n;
    while (amount--) {
      const sdo_cpx d = x[p];
      switch (space) {
        case SDO_SPACE_AMPLITUDE:
          if (amplitude) var = d.re * d.re + d.im * d.im;
          else var = d.re * zc_angle.re - d.im * zc_angle.im;
          var -= thres;
          break;
        case SDO_SPACE_PHASE: {
          const float pr = d.re * zc_angle.re - d.im * zc_angle.im;
          const float pi = d.re * zc_angle.im + d.im * zc_angle.re;
          var = sdo_atan2f(pi, pr);
          break;
        }
        default: {
          /* arg(i * x * conj(prev)) */
          const float ir = -d.im, ii = d.re;
          const float pr = ir * prev.re + ii * prev.im;
          const float pi = ii * prev.re - ir * prev.im;
          var = sdo_atan2f(pi, pr);
          prev = d;
          break;
        }
      }
      if ((var > 0 || var < 0) || last) {
        if (var * prev_var < 0 || last) {
          const long samples = p - last_zc;
          long symbols = (long) round((double) ((float) samples * bnor));
          while (symbols-- > 0 && i < TASK_BLOCK) {
            if (total < cap) sym[total] = var > 0;
            ++total; ++i;
          }
          last_zc = p;
          prev_var = var;
        }
      }
      ++p;
    }
  }
  return total;
}

/* Carrier detector: Blackman-Harris over the n samples, zero-padded to a power of two, power spectrum,
 * strongest bin outside the DC notch, power-weighted circular centroid around it -> rad/sample in (-pi, pi].
 * SPEC Y.5: the transform is the SPEC PSD (|X|^2 / N); allocation >= 64; every bin of the centroid window
 * is weighted by its power (the reference leaves the bins inside the notch un-squared). */
float sdo_carrier_detect(const sdo_cpx *x, size_t n, double avg_rel_bw, double dc_notch_rel_bw)
{
  size_t alloc = 64, k;
  sdo_spec_plan plan;
  sdo_cpx *buf, *scr, acc = { 0.0f, 0.0f };
  float *w, *psd, max_val = 0.0f;
  int i, max_ndx = 0, bins, delta, start, skip;
  while (alloc < n) alloc <<= 1;
  if (dc_notch_rel_bw < 0.) dc_notch_rel_bw = 0.;
  if (dc_notch_rel_bw > 1.) dc_notch_rel_bw = 1.;
  if (sdo_spec_plan_init(&plan, (unsigned) alloc, 0)) return 0.0f;
  buf = calloc(alloc, sizeof *buf); scr = calloc(alloc, sizeof *scr);
  w = malloc((n ? n : 1) * sizeof *w); psd = malloc(alloc * sizeof *psd);
  sdo_window_fill(w, (unsigned) n, SDO_WINDOW_BLACKMANN_HARRIS);
  for (k = 0; k < n; ++k) { buf[k].re = x[k].re * w[k]; buf[k].im = x[k].im * w[k]; }
  sdo_psd_frame_spec(&plan, NULL, buf, psd, scr);
  bins = (int) ((double) alloc * avg_rel_bw) + 1;
  delta = (bins - 1) / 2;
  skip = (int) (.5 * dc_notch_rel_bw * (double) alloc);
  for (i = skip; i < (int) alloc - skip; ++i)
    if (psd[i] > max_val) { max_val = psd[i]; max_ndx = i; }
  start = max_ndx - delta;
  for (i = 0; i < bins; ++i) {
    int j = i + start;
    float nfreq, s, c;
    if (j < 0) j += (int) alloc;
    j %= (int) alloc;
    nfreq = 2.f * (float) j / (float) alloc;
    sdo_sincosf(3.14159265358979323846f * nfreq, &s, &c);
    acc.re += psd[j] * c;
    acc.im += psd[j] * s;
  }
  free(buf); free(scr); free(w); free(psd);
  sdo_spec_plan_free(&plan);
  return sdo_atan2f(acc.im, acc.re);
}

/* ------------------------------------------------------------------------------------------------
 * SNR estimator of the inspector's histogram (Misc/SNREstimator.cpp:30-169, include/SNREstimator.h): a
 * comb of 2^bps Gaussians of width sigma on the unit circle is fitted to the normalised histogram by
 * one gradient step per feed.  expf(t) is SPEC M's 10^(t log10 e).
 * ---------------------------------------------------------------------------------------------- */
static float snr_expf(float t) { return sdo_exp10f(t * 0.4342944819032518f); }

void sdo_snr_init(sdo_snr_estimator *e)
{
  memset(e, 0, sizeof *e);
  e->sigma = 1.f / 8.f;                       /* SNR_ESTIMATOR_DEFAULT_SIGMA */
  e->alpha = 1.f;                             /* SNR_ESTIMATOR_DEFAULT_ALPHA */
}

void sdo_snr_free(sdo_snr_estimator *e)
{
  free(e->gaussian); free(e->hi); free(e->htilde);
  memset(e, 0, sizeof *e);
}

/* setBps, SNREstimator.cpp:122-131 */
void sdo_snr_set_bps(sdo_snr_estimator *e, unsigned bps)
{
  if (e->bps != bps) {
    e->bps = bps;
    e->sigma = 1.f / 8.f;
    e->intervals = 1u << bps;
    e->hx = 1.f / e->length;
  }
}

/* recalculateModel, SNREstimator.cpp:30-78 */
static void snr_model(sdo_snr_estimator *e)
{
  unsigned i, j;
  float x, max = 0, intlen, start, sigma2 = e->sigma * e->sigma;
  if (!(e->length > 0 && e->intervals > 0)) return;
  for (i = 0; i < e->length; ++i) {
    x = i * e->hx;
    if (x >= .5f) x -= 1.f;
    e->gaussian[i] = snr_expf(-x * x / sigma2);
  }
  intlen = 1.f / e->intervals;
  start = .5f * intlen;
  for (i = 0; i < e->length; ++i) e->hi[i] = 0.f;
  for (j = 0; j < e->intervals; ++j) {
    float skip = start + j * intlen;
    float t = 1.f - (skip - floorf(skip));
    unsigned skipint = (unsigned) floorf(e->length * skip), i1, i2;
    for (i = 0; i < e->length; ++i) {
      i1 = (unsigned) (e->length + i - skipint) % e->length;
      i2 = (unsigned) (e->length + i1 - 1) % e->length;
      e->hi[i] += t * e->gaussian[i1];
      e->hi[i] += (1 - t) * e->gaussian[i2];
    }
  }
  for (i = 0; i < e->length; ++i) if (e->hi[i] > max) max = e->hi[i];
  if (max > 0.f) for (i = 0; i < e->length; ++i) e->hi[i] /= max;
}

/* feed + iterate, SNREstimator.cpp:80-120,133-158 */
void sdo_snr_feed(sdo_snr_estimator *e, const unsigned *history, unsigned n)
{
  unsigned i, j, max = 0;
  if (e->length != n) {
    e->length = n;
    e->gaussian = realloc(e->gaussian, n * sizeof(float));
    e->hi = realloc(e->hi, n * sizeof(float));
    e->htilde = realloc(e->htilde, n * sizeof(float));
    e->hx = 1.f / e->length;
  }
  for (i = 0; i < n; ++i) if (max < history[i]) max = history[i];
  if (max == 0) max = 1;
  for (i = 0; i < n; ++i) e->htilde[i] = (float) history[i] / max;
  if (e->length > 0 && e->intervals > 0) {
    float delta = 0, x, term, intlen, start, skip;
    float sigmainv = 1.f / e->sigma, sigma3inv = sigmainv * sigmainv * sigmainv;
    snr_model(e);
    intlen = 1.f / e->intervals;
    start = .5f * intlen;
    for (i = 0; i < e->length; ++i) {
      x = i * e->hx;
      if (x >= .5f) x -= 1.f;
      term = 0;
      for (j = 0; j < e->intervals; ++j) {
        skip = start + j * intlen;
        term += (x - skip) * (x - skip);
      }
      term *= (e->hi[i] - e->htilde[i]) / sigma3inv;
      delta += term;
    }
    e->delta = delta / e->length;
    e->sigma += -e->alpha * e->delta;
  }
}

float sdo_snr_get(const sdo_snr_estimator *e) { return 1.f / (e->intervals * e->sigma); }

/* ------------------------------------------------------------------ R: DC removal ------------------
 * suscan_analyzer_set_dc_remove (Suscan/Analyzer.cpp:229-236; toggled from Default/Source/SourceWidget.cpp).  The
 * estimator lives in suscan (absent); SPEC R fixes a block-wise single pole: every sample of the block has the
 * estimate c of the blocks before it subtracted (one binary32 subtraction per component), then
 * c <- c + alpha (m - c) with m the mean of the RAW block, summed in binary64 in two fixed levels (runs of 256
 * samples in index order, then the run sums in index order) and rounded to binary32 once. */
void sdo_dc_remove(float c[2], const sdo_cpx *x, sdo_cpx *y, size_t n, float alpha)
{
  const size_t runs = (n + 255) / 256;
  double tr = 0.0, ti = 0.0;
  size_t r, i;
  float mr, mi;
  for (r = 0; r < runs; ++r) {
    const size_t i0 = r * 256, i1 = i0 + 256 < n ? i0 + 256 : n;
    double ar = 0.0, ai = 0.0;
    for (i = i0; i < i1; ++i) { ar += (double) x[i].re; ai += (double) x[i].im; }
    tr += ar; ti += ai;
  }
  mr = (float) (tr / (double) n); mi = (float) (ti / (double) n);
  for (i = 0; i < n; ++i) { y[i].re = x[i].re - c[0]; y[i].im = x[i].im - c[1]; }
  c[0] = c[0] + alpha * (mr - c[0]);
  c[1] = c[1] + alpha * (mi - c[1]);
}
