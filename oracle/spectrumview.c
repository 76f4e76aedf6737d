/*
 * spectrumview.c -- ORACLE (test infrastructure). Panoramic-scanner PSD stitcher, SPEC.md section V.
 *
 * This is the one piece of the path whose arithmetic is fully present in the reference:
 * SpectrumView::{setRange, interpolate, feedLinearMode, feedHistogramMode, feed, reset} at
 * Panoramic/Scanner.cpp:46-293, constants at include/Scanner.h:26-32.  The statements below follow
 * that arithmetic (same operand types: double for frequencies / positions, float for PSD values)
 * so the results are comparable value by value.
 */
#include "sd_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

static unsigned next_pow2(double v)
{
  unsigned n = (unsigned) v, i = 1;
  while (i < n) i <<= 1;
  return i;
}

int sdo_sview_init(sdo_spectrum_view *v)
{
  memset(v, 0, sizeof(*v));
  v->psd = (float *) malloc(sizeof(float) * SDO_SCANNER_SPECTRUM_SIZE);
  v->psd_accum = (float *) malloc(sizeof(float) * SDO_SCANNER_SPECTRUM_SIZE);
  v->psd_count = (float *) malloc(sizeof(float) * SDO_SCANNER_SPECTRUM_SIZE);
  if (!v->psd || !v->psd_accum || !v->psd_count) return -1;
  v->spectrum_size = SDO_SCANNER_SPECTRUM_SIZE;
  v->fft_rel_bw = 0.5f;                 /* include/Scanner.h:70 default relBw */
  sdo_sview_reset(v);
  return 0;
}

void sdo_sview_free(sdo_spectrum_view *v)
{
  free(v->psd); free(v->psd_accum); free(v->psd_count);
  memset(v, 0, sizeof(*v));
}

/* Panoramic/Scanner.cpp:287-293 */
void sdo_sview_reset(sdo_spectrum_view *v)
{
  memset(v->psd, 0, sizeof(float) * SDO_SCANNER_SPECTRUM_SIZE);
  memset(v->psd_accum, 0, sizeof(float) * SDO_SCANNER_SPECTRUM_SIZE);
  memset(v->psd_count, 0, sizeof(float) * SDO_SCANNER_SPECTRUM_SIZE);
}

/* Panoramic/Scanner.cpp:41-54 */
void sdo_sview_set_range(sdo_spectrum_view *v, double fmin, double fmax)
{
  v->freq_min = fmin;
  v->freq_max = fmax;
  v->freq_range = fmax - fmin;
  v->spectrum_size = next_pow2(v->freq_range / SDO_SCANNER_FREQ_RESOLUTION);
  if (v->spectrum_size > SDO_SCANNER_SPECTRUM_SIZE)
    v->spectrum_size = SDO_SCANNER_SPECTRUM_SIZE;
  sdo_sview_reset(v);
}

/* Panoramic/Scanner.cpp:56-116: divide accumulators, forget old history (count > 5 -> 1),
 * fill empty runs linearly between their neighbours. */
void sdo_sview_interpolate(sdo_spectrum_view *v)
{
  unsigned i, j, count = 1, zero_pos = 0;
  int first = 1, in_gap = 0;
  float left = SDO_SCANNER_DEFAULT_BIN_VALUE, right, t;

  for (i = 0; i < v->spectrum_size; ++i) {
    const int empty = v->psd_count[i] <= .5f;
    if (!in_gap) {
      if (empty) {
        in_gap = 1; zero_pos = i; count = 1;
        first = i == 0;
        if (!first) left = v->psd[i - 1];
      } else {
        v->psd[i] = v->psd_accum[i] / v->psd_count[i];
        if (v->psd_count[i] > SDO_SCANNER_COUNT_MAX) {
          v->psd_count[i] = SDO_SCANNER_COUNT_RESET;
          v->psd_accum[i] = v->psd[i] * SDO_SCANNER_COUNT_RESET;
        }
      }
    } else if (empty) {
      ++count;
    } else {
      in_gap = 0;
      right = v->psd[i] = v->psd_accum[i] / v->psd_count[i];
      if (first) {
        for (j = 0; j < count; ++j) v->psd[j + zero_pos] = right;
      } else {
        for (j = 0; j < count; ++j) {
          t = (float) (j + .5f) / count;
          v->psd[j + zero_pos] = (1 - t) * left + t * right;
        }
      }
    }
  }
  if (in_gap)
    for (j = 0; j < count; ++j) v->psd[j + zero_pos] = left;
}

static int clampi(int x, int lo, int hi) { return x < lo ? lo : (x > hi ? hi : x); }

/* Panoramic/Scanner.cpp:118-185 */
void sdo_sview_feed_linear(sdo_spectrum_view *v, const float *psd, const float *count,
                           size_t psd_size, double fmin, double fmax, int adjust_sides)
{
  double inp_bw = fmax - fmin, bw, freq_skip, fft_count, bins, pos, delta, src_bin_w, dst_bin_w;
  int skip, j, k, i;

  skip = adjust_sides ? (int) (.5f * (1 - v->fft_rel_bw) * psd_size) : 0;
  freq_skip = (double) skip / psd_size * inp_bw;
  bw = inp_bw - 2 * freq_skip;

  fft_count = v->freq_range / bw;
  bins = v->spectrum_size / fft_count;
  src_bin_w = inp_bw / psd_size;
  dst_bin_w = v->freq_range / v->spectrum_size;
  delta = dst_bin_w / src_bin_w;

  pos = (freq_skip + fmin - v->freq_min) / v->freq_range;
  pos *= v->spectrum_size;

  j = pos > 0 ? (int) pos : 0;
  k = pos + bins < v->spectrum_size ? (int) (pos + bins) : (int) v->spectrum_size;

  while (j < k) {
    double freq_j = v->freq_min + dst_bin_w * j;
    double src_bin = (freq_j - fmin) / src_bin_w;
    int start_bin = (int) src_bin;
    int end_bin = (int) (src_bin + delta);
    float acc = 0, cnt = 0;
    start_bin = clampi(start_bin, 0, (int) psd_size - 1);
    end_bin = clampi(end_bin, start_bin + 1, (int) psd_size);
    for (i = start_bin; i < end_bin; ++i) {
      acc += psd[i];
      cnt += count != NULL ? count[i] : 1;
    }
    if (cnt > 0) {
      v->psd_accum[j] += acc / cnt;
      v->psd_count[j] += 1;
    }
    ++j;
  }
}

/* Panoramic/Scanner.cpp:187-237 */
void sdo_sview_feed_histogram(sdo_spectrum_view *v, const float *psd, size_t psd_size,
                              double fmin, double fmax)
{
  double rel_bw = (fmax - fmin) / v->freq_range;
  double f_start = (fmin - v->freq_min) / v->freq_range;
  double f_end = (fmax - v->freq_min) / v->freq_range;
  float t, inv = (float) (1. / psd_size), accum = 0;
  unsigned j;
  size_t i;

  f_start *= v->spectrum_size;
  f_end *= v->spectrum_size;
  rel_bw *= v->spectrum_size;

  j = (unsigned) f_start;
  if (f_start < 0) j = 0;
  if (j > v->spectrum_size - 1) j = v->spectrum_size - 1;

  for (i = 0; i < psd_size; ++i) accum += psd[i];
  accum *= inv;

  if (floor(f_start) != floor(f_end)) {
    t = (float) ((f_start - floor(f_start)) / rel_bw);
    v->psd_count[j] += 1 - t;
    v->psd_accum[j] += (1 - t) * accum;
    if (j + 1 < v->spectrum_size) {
      v->psd_count[j + 1] += t;
      v->psd_accum[j + 1] += t * accum;
    }
  } else {
    v->psd_count[j] += 1;
    v->psd_accum[j] += accum;
  }
}

/* Panoramic/Scanner.cpp:239-256 */
void sdo_sview_feed_range(sdo_spectrum_view *v, const float *psd, const float *count, size_t psd_size,
                          double fmin, double fmax, int adjust_sides)
{
  double fft_count = (fmax - fmin) / v->freq_range;
  if (fft_count * v->spectrum_size >= 2)
    sdo_sview_feed_linear(v, psd, count, psd_size, fmin, fmax, adjust_sides);
  else
    sdo_sview_feed_histogram(v, psd, psd_size, fmin, fmax);
  sdo_sview_interpolate(v);
}

/* Panoramic/Scanner.cpp:258-274 */
void sdo_sview_feed(sdo_spectrum_view *v, const float *psd, const float *count, size_t psd_size,
                    double center, int adjust_sides)
{
  sdo_sview_feed_range(v, psd, count, psd_size, center - v->fft_bandwidth / 2, center + v->fft_bandwidth / 2,
                       adjust_sides);
}

/* Panoramic/Scanner.cpp:276-286: the accumulators of the other view, weighted by its counts, sides untouched */
void sdo_sview_feed_view(sdo_spectrum_view *v, const sdo_spectrum_view *detail)
{
  sdo_sview_feed_range(v, detail->psd_accum, detail->psd_count, detail->spectrum_size, detail->freq_min,
                       detail->freq_max, 0);
}
