/*
 * fft_fast.c -- ORACLE (test infrastructure), speed leg of the CPU baseline only.
 *
 * The parity transforms (fft_spec.c) follow the fixed dataflows of SPEC F.2-F.4 so that the CUDA kernels can be
 * bit-identical to them; they are scalar and make no attempt to be quick, which made round 1's CPU arm a straw man
 * (1.79 ms per 65536-point frame against 0.62 ms for pocketfft).  This file is what a CPU
 * implementation that cares about speed would run instead: a Stockham autosort radix-4 transform on split
 * re / im arrays whose inner loops are unit-stride and vectorise (gcc -O3 -march=native), twiddles tabulated once
 * per size.  Results agree with the SPEC transform to float32 rounding (tests/test_oracle.py), not bit for bit, so
 * it is never used for parity -- sdo_set_fast_transforms(1) is called only by bench.py's CPU legs.
 * What it stands in for: the FFTW3f plans under sigutils (SigDigger.pro:486; wisdom at App/Loader.cpp:46).
 */
#include "sd_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

int sdo_fast_transforms = 0;
void sdo_set_fast_transforms(int on) { sdo_fast_transforms = on; }

typedef struct { unsigned n; float *wr, *wi; } fast_tab;
static fast_tab g_tabs[32];

static const fast_tab *tab_for(unsigned n)
{
  unsigned lg = 0, i;
  while ((1u << lg) < n) ++lg;
  if (g_tabs[lg].n == n) return &g_tabs[lg];
#pragma omp critical(sdo_fast_tab)
  {
    if (g_tabs[lg].n != n) {
      float *wr = (float *) malloc(sizeof(float) * n), *wi = (float *) malloc(sizeof(float) * n);
      for (i = 0; i < n; ++i) {
        double a = 2.0 * SDO_PI * (double) i / (double) n;
        wr[i] = (float) cos(a);
        wi[i] = (float) -sin(a);
      }
      g_tabs[lg].wr = wr; g_tabs[lg].wi = wi;
#pragma omp flush
      g_tabs[lg].n = n;
    }
  }
  return &g_tabs[lg];
}

/* one radix-4 Stockham stage: n = current transform length, s = stride (number of interleaved transforms) */
static void stage4(unsigned n, unsigned s, unsigned tstep, const fast_tab *t, int sign,
                   const float *restrict xr, const float *restrict xi, float *restrict yr, float *restrict yi)
{
  const unsigned m = n >> 2;
  unsigned p, q;
  for (p = 0; p < m; ++p) {
    const float w1r = t->wr[p * tstep], w1i = sign < 0 ? t->wi[p * tstep] : -t->wi[p * tstep];
    const float w2r = t->wr[2 * p * tstep], w2i = sign < 0 ? t->wi[2 * p * tstep] : -t->wi[2 * p * tstep];
    const float w3r = t->wr[3 * p * tstep], w3i = sign < 0 ? t->wi[3 * p * tstep] : -t->wi[3 * p * tstep];
    const float *ar = xr + (size_t) s * p, *ai = xi + (size_t) s * p;
    const float *br = ar + (size_t) s * m, *bi = ai + (size_t) s * m;
    const float *cr = br + (size_t) s * m, *ci = bi + (size_t) s * m;
    const float *dr = cr + (size_t) s * m, *di = ci + (size_t) s * m;
    float *y0r = yr + (size_t) s * 4 * p, *y0i = yi + (size_t) s * 4 * p;
    float *y1r = y0r + s, *y1i = y0i + s, *y2r = y1r + s, *y2i = y1i + s, *y3r = y2r + s, *y3i = y2i + s;
    for (q = 0; q < s; ++q) {
      const float apcr = ar[q] + cr[q], apci = ai[q] + ci[q], amcr = ar[q] - cr[q], amci = ai[q] - ci[q];
      const float bpdr = br[q] + dr[q], bpdi = bi[q] + di[q], bmdr = br[q] - dr[q], bmdi = bi[q] - di[q];
      /* forward: -i (b - d) = (bmdi, -bmdr); inverse: +i (b - d) = (-bmdi, bmdr) */
      const float jr = sign < 0 ? bmdi : -bmdi, ji = sign < 0 ? -bmdr : bmdr;
      const float t1r = amcr + jr, t1i = amci + ji, t2r = apcr - bpdr, t2i = apci - bpdi;
      const float t3r = amcr - jr, t3i = amci - ji;
      y0r[q] = apcr + bpdr; y0i[q] = apci + bpdi;
      y1r[q] = t1r * w1r - t1i * w1i; y1i[q] = t1r * w1i + t1i * w1r;
      y2r[q] = t2r * w2r - t2i * w2i; y2i[q] = t2r * w2i + t2i * w2r;
      y3r[q] = t3r * w3r - t3i * w3i; y3i[q] = t3r * w3i + t3i * w3r;
    }
  }
}

static void stage2(unsigned n, unsigned s, unsigned tstep, const fast_tab *t, int sign,
                   const float *restrict xr, const float *restrict xi, float *restrict yr, float *restrict yi)
{
  const unsigned m = n >> 1;
  unsigned p, q;
  for (p = 0; p < m; ++p) {
    const float wr = t->wr[p * tstep], wi = sign < 0 ? t->wi[p * tstep] : -t->wi[p * tstep];
    const float *ar = xr + (size_t) s * p, *ai = xi + (size_t) s * p;
    const float *br = ar + (size_t) s * m, *bi = ai + (size_t) s * m;
    float *y0r = yr + (size_t) s * 2 * p, *y0i = yi + (size_t) s * 2 * p, *y1r = y0r + s, *y1i = y0i + s;
    for (q = 0; q < s; ++q) {
      const float dr = ar[q] - br[q], di = ai[q] - bi[q];
      y0r[q] = ar[q] + br[q]; y0i[q] = ai[q] + bi[q];
      y1r[q] = dr * wr - di * wi; y1i[q] = dr * wi + di * wr;
    }
  }
}

/* out = DFT(in .* window), sign -1 forward / +1 inverse (unnormalised).  in == out allowed.  window may be NULL. */
void sdo_fast_fft(const sdo_cpx *in, const float *window, sdo_cpx *out, unsigned n, int sign)
{
  const fast_tab *t = tab_for(n);
  float *buf = (float *) malloc(sizeof(float) * 4 * (size_t) n);
  float *xr = buf, *xi = buf + n, *yr = buf + 2 * (size_t) n, *yi = buf + 3 * (size_t) n, *sw;
  unsigned i, len = n, s = 1, tstep = 1;
  if (window) for (i = 0; i < n; ++i) { xr[i] = in[i].re * window[i]; xi[i] = in[i].im * window[i]; }
  else        for (i = 0; i < n; ++i) { xr[i] = in[i].re; xi[i] = in[i].im; }
  while (len >= 4) {
    stage4(len, s, tstep, t, sign, xr, xi, yr, yi);
    sw = xr; xr = yr; yr = sw; sw = xi; xi = yi; yi = sw;
    len >>= 2; s <<= 2; tstep <<= 2;
  }
  if (len == 2) {
    stage2(len, s, tstep, t, sign, xr, xi, yr, yi);
    sw = xr; xr = yr; yr = sw; sw = xi; xi = yi; yi = sw;
  }
  for (i = 0; i < n; ++i) { out[i].re = xr[i]; out[i].im = xi[i]; }
  free(buf);
}
