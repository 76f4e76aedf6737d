/*
 * fft_spec.c -- ORACLE (test infrastructure). The "SPEC FFT": SPEC.md sections F.2-F.4.
 *
 * FFTW's operation order (what sigutils uses upstream, SU_FFTW at Tasks/CarrierDetector.cpp:58-75) is
 * plan dependent and unknowable here, so the project FIXES one dataflow per transform size and states it
 * in SPEC.md.  The CUDA kernels follow the same dataflow with FMA contraction off, which makes PSD bins,
 * channel samples and therefore soft symbols bit-identical end to end (the loop recurrences downstream
 * are discontinuous at a few thresholds, so anything short of bit-identical channel samples eventually
 * diverges; see tests/parity.py).  fft.c's textbook radix-2 transform stays as the independent
 * cross-check of this file (tests/test_oracle.py::test_spec_fft_vs_radix2_and_float64).
 *
 * Every complex product is (a.x*b.x - a.y*b.y, a.x*b.y + a.y*b.x), one rounding per operator.
 * Compile with -ffp-contract=off.
 */
#include "sd_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* SPEC F.1: twiddle products are one rounded product plus one fused multiply-add per component
 * (fmaf is exact-then-round on both sides; gcc -mfma makes it one instruction). */
static inline sdo_cpx cmul(sdo_cpx a, sdo_cpx b)
{
  sdo_cpx r;
  r.re = fmaf(a.re, b.re, -(a.im * b.im));
  r.im = fmaf(a.re, b.im, a.im * b.re);
  return r;
}
static inline sdo_cpx cadd(sdo_cpx a, sdo_cpx b) { sdo_cpx r = { a.re + b.re, a.im + b.im }; return r; }
static inline sdo_cpx csub(sdo_cpx a, sdo_cpx b) { sdo_cpx r = { a.re - b.re, a.im - b.im }; return r; }
static inline sdo_cpx mul_mi(sdo_cpx a) { sdo_cpx r = { a.im, -a.re }; return r; }

/* twiddle table W_n^i = (cos, -sin)(2 pi i / n), binary64 rounded once */
sdo_cpx *sdo_spec_twiddles(unsigned n)
{
  sdo_cpx *t = (sdo_cpx *) malloc(sizeof(sdo_cpx) * n);
  unsigned i;
  for (i = 0; i < n; ++i) {
    double a = 2.0 * SDO_PI * (double) i / (double) n;
    t[i].re = (float) cos(a);
    t[i].im = (float) -sin(a);
  }
  return t;
}

/* F.2: Stockham radix-4 stages (Ns = 1, 4, 16, ...) then one radix-2 stage if log2(M) is odd. In place. */
void sdo_spec_fft_stockham(sdo_cpx *s, unsigned M, const sdo_cpx *tw, int sign, sdo_cpx *tmp)
{
  unsigned logM = 0, logNs = 0, j;
  const unsigned q = M >> 2;
  if (sdo_fast_transforms && M >= 8) { sdo_fast_fft(s, NULL, s, M, sign); return; }   /* bench CPU legs only */
  while ((1u << logM) < M) ++logM;
  for (; logNs + 2 <= logM; logNs += 2) {
    const unsigned Ns = 1u << logNs;
    for (j = 0; j < q; ++j) {
      const unsigned k = j & (Ns - 1);
      sdo_cpx v0 = s[j], v1 = s[j + q], v2 = s[j + 2 * q], v3 = s[j + 3 * q], a, bb, c, dd, d;
      unsigned j0;
      if (logNs > 0) {
        const unsigned step = M >> (logNs + 2);
        sdo_cpx w1 = tw[k * step], w2 = tw[2 * k * step], w3 = tw[3 * k * step];
        if (sign > 0) { w1.im = -w1.im; w2.im = -w2.im; w3.im = -w3.im; }
        v1 = cmul(v1, w1); v2 = cmul(v2, w2); v3 = cmul(v3, w3);
      }
      a = cadd(v0, v2); bb = csub(v0, v2); c = cadd(v1, v3); dd = csub(v1, v3);
      if (sign < 0) { d.re = dd.im; d.im = -dd.re; } else { d.re = -dd.im; d.im = dd.re; }
      j0 = ((j - k) << 2) + k;
      tmp[j0] = cadd(a, c);
      tmp[j0 + Ns] = cadd(bb, d);
      tmp[j0 + 2 * Ns] = csub(a, c);
      tmp[j0 + 3 * Ns] = csub(bb, d);
    }
    memcpy(s, tmp, sizeof(sdo_cpx) * M);
  }
  if (logNs < logM) {
    const unsigned h = M >> 1;
    for (j = 0; j < h; ++j) {
      sdo_cpx w = tw[j], a = s[j], t;
      if (sign > 0) w.im = -w.im;
      t = cmul(s[j + h], w);
      s[j] = cadd(a, t);
      s[j + h] = csub(a, t);
    }
  }
}

/* ---- F.4 building block: forward 16-point DFT, result X[m + 4 q] left in v[4 m + q] ---- */
#define SC1 0.92387953251128675613f
#define SS1 0.38268343236508977173f
#define SR2 0.70710678118654752440f
#define REV16(p) ((((p) >> 2)) | (((p) & 3) << 2))

static inline void fft4(sdo_cpx *a, sdo_cpx *b, sdo_cpx *c, sdo_cpx *d)
{
  sdo_cpx s0 = cadd(*a, *c), d0 = csub(*a, *c), s1 = cadd(*b, *d), d1 = mul_mi(csub(*b, *d));
  *a = cadd(s0, s1); *b = cadd(d0, d1); *c = csub(s0, s1); *d = csub(d0, d1);
}

static void fft16(sdo_cpx *v)
{
  int i, m;
  const sdo_cpx w1 = { SC1, -SS1 }, w2 = { SR2, -SR2 }, w3 = { SS1, -SC1 }, w6 = { -SR2, -SR2 }, w9 = { -SC1, SS1 };
  for (i = 0; i < 4; ++i) fft4(&v[i], &v[i + 4], &v[i + 8], &v[i + 12]);
  v[5] = cmul(v[5], w1);  v[9] = cmul(v[9], w2);   v[13] = cmul(v[13], w3);
  v[6] = cmul(v[6], w2);  v[10] = mul_mi(v[10]);   v[14] = cmul(v[14], w6);
  v[7] = cmul(v[7], w3);  v[11] = cmul(v[11], w6); v[15] = cmul(v[15], w9);
  for (m = 0; m < 4; ++m) fft4(&v[4 * m], &v[4 * m + 1], &v[4 * m + 2], &v[4 * m + 3]);
}

/* 256-point forward transform of 256 values at stride `stride` of `in`, = 16 x 16 with both butterflies
 * as fft16 (F.4).  out[k], k = ka + 16 kb. */
static void fft256_16x16(const sdo_cpx *in, size_t stride, const float *win, size_t wstride,
                         const sdo_cpx *tw256, sdo_cpx *out)
{
  sdo_cpx y[16][16];   /* [ka][t] */
  int t, j, q, ka;
  for (t = 0; t < 16; ++t) {
    sdo_cpx v[16];
    for (j = 0; j < 16; ++j) {
      v[j] = in[(size_t) (t + 16 * j) * stride];
      if (win) { float w = win[(size_t) (t + 16 * j) * wstride]; v[j].re *= w; v[j].im *= w; }
    }
    fft16(v);
    for (q = 0; q < 16; ++q) {
      ka = REV16(q);
      y[ka][t] = ka ? cmul(v[q], tw256[t * ka]) : v[q];
    }
  }
  for (ka = 0; ka < 16; ++ka) {
    sdo_cpx v[16];
    for (t = 0; t < 16; ++t) v[t] = y[ka][t];
    fft16(v);
    for (q = 0; q < 16; ++q) out[ka + 16 * REV16(q)] = v[q];
  }
}

/* ---- F.5 building block: forward 8-point DFT = one radix-2 stage + two DFT4; X[kb] is left in v[4 (kb & 1) + (kb >> 1)] ---- */
#define REV8(p) (2 * ((p) & 3) + ((p) >> 2))
static void fft8(sdo_cpx *v)
{
  int i;
  const sdo_cpx w1 = { SR2, -SR2 }, w3 = { -SR2, -SR2 };
  for (i = 0; i < 4; ++i) {
    const sdo_cpx a = cadd(v[i], v[i + 4]), b = csub(v[i], v[i + 4]);
    v[i] = a; v[i + 4] = b;
  }
  v[5] = cmul(v[5], w1); v[6] = mul_mi(v[6]); v[7] = cmul(v[7], w3);
  fft4(&v[0], &v[1], &v[2], &v[3]);
  fft4(&v[4], &v[5], &v[6], &v[7]);
}

/* (16 T)-point forward transform of 16 T values at stride `stride` of `in`, T = 2, 4 or 8: n = t + T j -> fft16 over j ->
 * x W_(16T)^(t ka) -> T-point butterfly over t (F.5).  out[k], k = ka + 16 kb; X[kb] of the T-point butterfly sits at
 * position pos(kb): T = 8: 4 (kb & 1) + (kb >> 1) (fft8); T = 4: kb (DFT4); T = 2: kb. */
static void fft16T(const sdo_cpx *in, size_t stride, const float *win, size_t wstride, const sdo_cpx *twN1, int T,
                   sdo_cpx *out)
{
  sdo_cpx y[16][8];    /* [ka][t] */
  int t, j, q, ka;
  for (t = 0; t < T; ++t) {
    sdo_cpx v[16];
    for (j = 0; j < 16; ++j) {
      v[j] = in[(size_t) (t + T * j) * stride];
      if (win) { float w = win[(size_t) (t + T * j) * wstride]; v[j].re *= w; v[j].im *= w; }
    }
    fft16(v);
    for (q = 0; q < 16; ++q) {
      ka = REV16(q);
      y[ka][t] = ka ? cmul(v[q], twN1[t * ka]) : v[q];
    }
  }
  for (ka = 0; ka < 16; ++ka) {
    sdo_cpx v[8];
    for (t = 0; t < T; ++t) v[t] = y[ka][t];
    if (T == 8) {
      fft8(v);
      for (q = 0; q < 8; ++q) out[ka + 16 * REV8(q)] = v[q];
    } else if (T == 4) {
      fft4(&v[0], &v[1], &v[2], &v[3]);
      for (q = 0; q < 4; ++q) out[ka + 16 * q] = v[q];
    } else {
      out[ka] = cadd(v[0], v[1]);
      out[ka + 16] = csub(v[0], v[1]);
    }
  }
}

/* Plans: tables and scratch for one transform size.
 *   N == 65536            -> F.4 (256 x 256, fft16 butterflies, coarse x fine inter-pass twiddle)
 *   N == 32768 / 16384 / 8192 -> F.5 (N1 x 256, N1 = 16 T: columns = fft16 + T-point butterfly, rows as F.4)
 *   N <= 4096 && !four    -> F.2 single Stockham transform
 *   otherwise             -> F.3 four-step N1 x N2 with Stockham sub-transforms
 * `four` forces the four-step form (the channeliser's forward transform always uses it). */
int sdo_spec_plan_init(sdo_spec_plan *p, unsigned N, int four)
{
  unsigned l = 0;
  memset(p, 0, sizeof(*p));
  if (N < 4 || (N & (N - 1))) return -1;
  while ((1u << l) < N) ++l;
  p->N = N; p->four = four;
  if (N == 65536) {
    p->kind = 2; p->N1 = 256; p->N2 = 256;
    p->tw_a = sdo_spec_twiddles(256);
  } else if (N == 32768 || N == 16384 || N == 8192) {
    p->kind = 3; p->N1 = N / 256; p->N2 = 256;     /* F.5: N1 = 16 T, T = 8 / 4 / 2 */
    p->tw_a = sdo_spec_twiddles(p->N1);
    p->tw_b = sdo_spec_twiddles(256);
  } else if (N <= 4096 && !four) {
    p->kind = 0; p->N1 = N; p->N2 = 1;
  } else {
    p->kind = 1; p->N1 = 1u << (l / 2); p->N2 = N / p->N1;
    p->tw_a = sdo_spec_twiddles(p->N1);
    p->tw_b = sdo_spec_twiddles(p->N2);
  }
  p->tw_n = sdo_spec_twiddles(N);
  p->scr = (sdo_cpx *) malloc(sizeof(sdo_cpx) * N);
  p->buf = (sdo_cpx *) malloc(sizeof(sdo_cpx) * N);
  p->tmp = (sdo_cpx *) malloc(sizeof(sdo_cpx) * N);
  return 0;
}

void sdo_spec_plan_free(sdo_spec_plan *p)
{
  free(p->tw_a); free(p->tw_b); free(p->tw_n); free(p->scr); free(p->buf); free(p->tmp);
  memset(p, 0, sizeof(*p));
}

/* X = FFT_forward(window .* x) with the SPEC dataflow of this size. */
void sdo_spec_forward(const sdo_spec_plan *p, const sdo_cpx *x, const float *window, sdo_cpx *X)
{
  const unsigned N = p->N, N1 = p->N1, N2 = p->N2;
  unsigned n1, n2, k1, i;
  if (sdo_fast_transforms) { sdo_fast_fft(x, window, X, N, -1); return; }   /* bench CPU legs only (fft_fast.c) */
  if (p->kind == 2) {
    sdo_cpx col[256];
    for (n2 = 0; n2 < 256; ++n2) {
      fft256_16x16(x + n2, 256, window ? window + n2 : NULL, 256, p->tw_a, col);
      for (k1 = 0; k1 < 256; ++k1) {
        unsigned pw = n2 * k1;
        sdo_cpx tw = cmul(p->tw_a[pw >> 8], p->tw_n[pw & 255]);
        p->scr[(size_t) k1 * 256 + n2] = cmul(col[k1], tw);
      }
    }
    for (k1 = 0; k1 < 256; ++k1) {
      fft256_16x16(p->scr + (size_t) k1 * 256, 1, NULL, 0, p->tw_a, col);
      for (n2 = 0; n2 < 256; ++n2) X[k1 + 256 * n2] = col[n2];
    }
    return;
  }
  if (p->kind == 3) {
    /* F.5: 256 columns of N1 = 16 T, inter-pass twiddle W_N^p = W_N1^(p >> 8) x W_N^(p & 255), N1 rows of 256 */
    sdo_cpx col[256];
    const int T = (int) N1 / 16;
    for (n2 = 0; n2 < 256; ++n2) {
      fft16T(x + n2, 256, window ? window + n2 : NULL, 256, p->tw_a, T, col);
      for (k1 = 0; k1 < N1; ++k1) {
        unsigned pw = n2 * k1;
        sdo_cpx tw = cmul(p->tw_a[pw >> 8], p->tw_n[pw & 255]);
        p->scr[(size_t) k1 * 256 + n2] = cmul(col[k1], tw);
      }
    }
    for (k1 = 0; k1 < N1; ++k1) {
      fft256_16x16(p->scr + (size_t) k1 * 256, 1, NULL, 0, p->tw_b, col);
      for (n2 = 0; n2 < 256; ++n2) X[k1 + (size_t) N1 * n2] = col[n2];
    }
    return;
  }
  if (p->kind == 0) {
    for (i = 0; i < N; ++i) {
      X[i] = x[i];
      if (window) { X[i].re *= window[i]; X[i].im *= window[i]; }
    }
    sdo_spec_fft_stockham(X, N, p->tw_n, -1, p->tmp);
    return;
  }
  for (n2 = 0; n2 < N2; ++n2) {
    for (n1 = 0; n1 < N1; ++n1) {
      p->buf[n1] = x[(size_t) n1 * N2 + n2];
      if (window) { float w = window[(size_t) n1 * N2 + n2]; p->buf[n1].re *= w; p->buf[n1].im *= w; }
    }
    sdo_spec_fft_stockham(p->buf, N1, p->tw_a, -1, p->tmp);
    for (k1 = 0; k1 < N1; ++k1) p->scr[(size_t) k1 * N2 + n2] = cmul(p->buf[k1], p->tw_n[n2 * k1]);
  }
  for (k1 = 0; k1 < N1; ++k1) {
    memcpy(p->buf, p->scr + (size_t) k1 * N2, sizeof(sdo_cpx) * N2);
    sdo_spec_fft_stockham(p->buf, N2, p->tw_b, -1, p->tmp);
    for (n2 = 0; n2 < N2; ++n2) X[k1 + (size_t) N1 * n2] = p->buf[n2];
  }
}

/* In-place inverse (unnormalised) F.2 transform of p->N points: the channel IFFT. Plan must be kind 0
 * or built with four = 0; sizes above 4096 also use the single Stockham form here (SPEC S.4). */
void sdo_spec_inverse_stockham(const sdo_spec_plan *p, sdo_cpx *s)
{
  sdo_spec_fft_stockham(s, p->N, p->tw_n, +1, p->tmp);
}
