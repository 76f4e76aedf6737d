/*
 * fft.c -- ORACLE (test infrastructure). Power-of-two FFT, windows, main PSD.  SPEC.md sections F, W, P.
 *
 * The reference reaches FFTW3f through sigutils (SigDigger.pro:486 `PKGCONFIG += suscan fftw3f`;
 * SU_FFTW(_plan_dft_1d)/(_execute) at Tasks/CarrierDetector.cpp:58-75).  FFTW's operation order is
 * plan dependent, so any float32 FFT with correctly rounded twiddles is an equally valid restatement:
 * this one is the textbook iterative radix-2 decimation-in-time transform.
 */
#include "sd_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

int sdo_fft_plan_init(sdo_fft_plan *p, unsigned n)
{
  unsigned i, l = 0;
  memset(p, 0, sizeof(*p));
  if (n < 2 || (n & (n - 1)))
    return -1;
  while ((1u << l) < n) ++l;
  p->n = n; p->log2n = l;
  p->tw_re = (float *) malloc(sizeof(float) * (n / 2));
  p->tw_im = (float *) malloc(sizeof(float) * (n / 2));
  p->rev   = (unsigned *) malloc(sizeof(unsigned) * n);
  if (!p->tw_re || !p->tw_im || !p->rev) { sdo_fft_plan_free(p); return -1; }
  for (i = 0; i < n / 2; ++i) {
    double a = 2.0 * SDO_PI * (double) i / (double) n;
    p->tw_re[i] = (float) cos(a);
    p->tw_im[i] = (float) -sin(a);
  }
  for (i = 0; i < n; ++i) {
    unsigned r = 0, b;
    for (b = 0; b < l; ++b)
      if (i & (1u << b)) r |= 1u << (l - 1 - b);
    p->rev[i] = r;
  }
  return 0;
}

void sdo_fft_plan_free(sdo_fft_plan *p)
{
  free(p->tw_re); free(p->tw_im); free(p->rev);
  memset(p, 0, sizeof(*p));
}

void sdo_fft_exec(const sdo_fft_plan *p, const sdo_cpx *in, sdo_cpx *out, int sign)
{
  const unsigned n = p->n;
  unsigned i, len, k;
  if (in != out) {
    for (i = 0; i < n; ++i) out[p->rev[i]] = in[i];
  } else {
    for (i = 0; i < n; ++i) {
      unsigned r = p->rev[i];
      if (r > i) { sdo_cpx t = out[i]; out[i] = out[r]; out[r] = t; }
    }
  }
  for (len = 2; len <= n; len <<= 1) {
    const unsigned half = len >> 1, step = n / len;
    for (i = 0; i < n; i += len) {
      for (k = 0; k < half; ++k) {
        const float wr = p->tw_re[k * step];
        const float wi = sign < 0 ? p->tw_im[k * step] : -p->tw_im[k * step];
        sdo_cpx *a = out + i + k, *b = out + i + k + half;
        const float tr = wr * b->re - wi * b->im;
        const float ti = wr * b->im + wi * b->re;
        b->re = a->re - tr; b->im = a->im - ti;
        a->re = a->re + tr; a->im = a->im + ti;
      }
    }
  }
}

/* W: cosine-sum windows, symmetric form (denominator n-1), evaluated in double and rounded. */
void sdo_window_fill(float *w, unsigned n, int type)
{
  unsigned i;
  for (i = 0; i < n; ++i) {
    double x = n > 1 ? 2.0 * SDO_PI * (double) i / (double) (n - 1) : 0.0, v;
    switch (type) {
      case SDO_WINDOW_HAMMING:  v = 0.54 - 0.46 * cos(x); break;
      case SDO_WINDOW_HANN:     v = 0.5 - 0.5 * cos(x); break;
      case SDO_WINDOW_FLAT_TOP:
        v = 1.0 - 1.93 * cos(x) + 1.29 * cos(2 * x) - 0.388 * cos(3 * x) + 0.028 * cos(4 * x);
        break;
      case SDO_WINDOW_BLACKMANN_HARRIS:
        v = 0.35875 - 0.48829 * cos(x) + 0.14128 * cos(2 * x) - 0.01168 * cos(3 * x);
        break;
      default: v = 1.0; break;
    }
    w[i] = (float) v;
  }
}

/* P: psd[k] = |FFT(window .* x)[k]|^2 / N, linear power, DC at index 0 (un-shifted). */
void sdo_psd_frame(const sdo_fft_plan *p, const float *window, const sdo_cpx *x, float *psd,
                   sdo_cpx *scratch)
{
  const unsigned n = p->n;
  const float inv_n = 1.0f / (float) n;
  sdo_cpx *tmp = scratch, *X = scratch + n;
  unsigned i;
  for (i = 0; i < n; ++i) {
    float w = window ? window[i] : 1.0f;
    tmp[i].re = x[i].re * w;
    tmp[i].im = x[i].im * w;
  }
  sdo_fft_exec(p, tmp, X, -1);
  for (i = 0; i < n; ++i)
    psd[i] = (X[i].re * X[i].re + X[i].im * X[i].im) * inv_n;
}

/* same, with the SPEC transform (bit-identical to the CUDA path) */
void sdo_psd_frame_spec(const sdo_spec_plan *p, const float *window, const sdo_cpx *x, float *psd,
                        sdo_cpx *scratch)
{
  const unsigned n = p->N;
  const float inv_n = 1.0f / (float) n;
  unsigned i;
  sdo_spec_forward(p, x, window, scratch);
  for (i = 0; i < n; ++i)
    psd[i] = fmaf(scratch[i].re, scratch[i].re, scratch[i].im * scratch[i].im) * inv_n;
}

/* Suscan/Messages/PSDMessage.cpp:32-38 -- swap halves and convert to dB in one pass. */
void sdo_psd_shift_db(float *psd, unsigned n)
{
  unsigned i, half = n / 2;
  for (i = 0; i < half; ++i) {
    float tmp = psd[i + half];
    psd[i + half] = sdo_power_db(psd[i]);
    psd[i] = sdo_power_db(tmp);
  }
}

/* Misc/Averager.cpp:43-49 -- last += alpha * (x - last). */
void sdo_averager_feed(float *last, const float *frame, unsigned n, float alpha)
{
  unsigned i;
  if (alpha >= 1.0f) { memcpy(last, frame, sizeof(float) * n); return; }
  for (i = 0; i < n; ++i)
    last[i] += alpha * (frame[i] - last[i]);
}
