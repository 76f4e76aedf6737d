/*
 * specttuner.c -- ORACLE (test infrastructure). FFT filter-bank channeliser, SPEC.md section S.
 *
 * Restates su_specttuner as driven by the reference at Tasks/LPFTask.cpp:52-69 (params: f0, bw,
 * guard, privdata, on_data; `guard = 2 PI / bw` "ensures no decimation"), :83-87 (feed_bulk),
 * :104-107 (latency of half a window, flushed with zeros) and :28-42 (on_data contract: pointer valid
 * until the next feed, returns SUBOOL).  The manual (doc/SigDigger_User_Manual.pdf pp.31-32) fixes the
 * power-of-two decimation.  sigutils itself is absent from /root/reference; the inner equations are
 * this project's SPEC (parity unpinned).
 */
#include "sd_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

struct sdo_specttuner {
  unsigned     window_size, half_size, p;
  sdo_cpx     *window;   /* window_size input samples */
  sdo_cpx     *fft;      /* window_size spectrum */
  sdo_spec_plan plan;      /* forward transform: SPEC F.3 / F.4, always the four-step form */
  sdo_st_channel *channels;
  unsigned long hops;
};

/* S.2 geometry: even centre bin, power-of-two size holding guard*bw, width = bins copied. */
void sdo_st_channel_geometry(unsigned window_size, float f0, float bw, float guard,
                             unsigned *center, unsigned *size, unsigned *width)
{
  const double N = (double) window_size;
  double krel = (double) guard * (double) bw / (2.0 * SDO_PI);
  double c = 2.0 * floor((double) f0 / (4.0 * SDO_PI) * N + 0.5);
  double m = ceil(krel * N - 1e-3);
  unsigned msz, sz = 1, w;
  if (m < 4.0) m = 4.0;
  if (m > N) m = N;
  msz = (unsigned) m;
  while (sz < msz) sz <<= 1;
  if (sz > window_size) sz = window_size;
  w = (unsigned) ceil((double) msz / (double) guard - 1e-3);
  if (w > sz) w = sz;
  if (w < 2) w = 2;
  *center = ((unsigned) c) % window_size;
  *size = sz;
  *width = w;
}

/* S.3 shaping response: ideal brick wall over bins [-halfw, halfw) circularly convolved with the 7-tap
 * spectrum of the periodic Blackman-Harris window (== windowing the filter's impulse response in time).
 * h is indexed like the FFT: h[i] for i in [0, halfw) are the upper-sideband bins +i, h[ws - j] the
 * lower-sideband bins -j. */
void sdo_st_filter_response(unsigned window_size, unsigned halfw, float *h)
{
  static const double c[4] = { 0.35875, 0.48829 / 2.0, 0.14128 / 2.0, 0.01168 / 2.0 };
  unsigned i;
  int m;
  if (2 * halfw >= window_size) {
    for (i = 0; i < window_size; ++i) h[i] = 1.0f;
    return;
  }
  for (i = 0; i < window_size; ++i) {
    int k = i < window_size / 2 ? (int) i : (int) i - (int) window_size; /* signed bin */
    double acc = 0.0;
    for (m = -3; m <= 3; ++m) {
      int q = k - m;
      if (q >= -(int) halfw && q < (int) halfw) acc += c[m < 0 ? -m : m];
    }
    h[i] = (float) acc;
  }
}

sdo_specttuner *sdo_specttuner_new(unsigned window_size)
{
  sdo_specttuner *st = (sdo_specttuner *) calloc(1, sizeof(*st));
  if (!st) return NULL;
  if (sdo_spec_plan_init(&st->plan, window_size, 1) != 0) { free(st); return NULL; }
  st->window_size = window_size;
  st->half_size = window_size / 2;
  st->window = (sdo_cpx *) calloc(window_size, sizeof(sdo_cpx));
  st->fft = (sdo_cpx *) calloc(window_size, sizeof(sdo_cpx));
  return st;
}

static void channel_free(sdo_st_channel *ch)
{
  free(ch->h); free(ch->window); free(ch->fft); free(ch->ifft[0]); free(ch->ifft[1]); free(ch->out);
  free(ch->tw); free(ch->tmp);
  free(ch);
}

void sdo_specttuner_destroy(sdo_specttuner *st)
{
  sdo_st_channel *ch, *nx;
  if (!st) return;
  for (ch = st->channels; ch; ch = nx) { nx = ch->next; channel_free(ch); }
  sdo_spec_plan_free(&st->plan);
  free(st->window); free(st->fft); free(st);
}

unsigned sdo_specttuner_window_size(const sdo_specttuner *st) { return st->window_size; }

sdo_st_channel *sdo_specttuner_open_channel(sdo_specttuner *st, const sdo_st_channel_params *p)
{
  sdo_st_channel *ch, **tail;
  unsigned i;
  if (!(p->guard >= 1.0f) || !(p->bw > 0.0f) || p->bw > SDO_2PI_F * 1.0001f) return NULL;
  if (!(p->f0 >= 0.0f) || p->f0 >= SDO_2PI_F) return NULL;
  ch = (sdo_st_channel *) calloc(1, sizeof(*ch));
  if (!ch) return NULL;
  ch->params = *p;
  sdo_st_channel_geometry(st->window_size, p->f0, p->bw, p->guard, &ch->center, &ch->size, &ch->width);
  ch->halfw = ch->width >> 1;
  ch->halfsz = ch->size >> 1;
  ch->decimation = (float) st->window_size / (float) ch->size;
  ch->k = 1.0f / (float) st->window_size;
  ch->gain = 1.0f;
  ch->h = (float *) malloc(sizeof(float) * st->window_size);
  ch->window = (float *) malloc(sizeof(float) * ch->size);
  ch->fft = (sdo_cpx *) calloc(ch->size, sizeof(sdo_cpx));
  ch->ifft[0] = (sdo_cpx *) calloc(ch->size, sizeof(sdo_cpx));
  ch->ifft[1] = (sdo_cpx *) calloc(ch->size, sizeof(sdo_cpx));
  ch->out = (sdo_cpx *) calloc(ch->halfsz, sizeof(sdo_cpx));
  ch->tw = sdo_spec_twiddles(ch->size);
  ch->tmp = (sdo_cpx *) malloc(sizeof(sdo_cpx) * ch->size);
  sdo_st_filter_response(st->window_size, ch->halfw, ch->h);
  for (i = 0; i < st->window_size; ++i) ch->h[i] = ch->k * ch->h[i];
  for (i = 0; i < ch->size; ++i) {
    double s = sin(SDO_PI * (double) i / (double) ch->size);
    ch->window[i] = (float) (s * s);
  }
  if (p->precise) {
    /* residual offset from the even centre bin, scaled to the channel rate; fnor = omega / pi */
    double resid = (double) p->f0 - 2.0 * SDO_PI * (double) ch->center / (double) st->window_size;
    if (resid > SDO_PI) resid -= 2.0 * SDO_PI;
    sdo_ncqo_init(&ch->lo, (float) (resid * (double) ch->decimation / SDO_PI));
  }
  for (tail = &st->channels; *tail; tail = &(*tail)->next) ;
  *tail = ch;
  return ch;
}

static int feed_channel(sdo_specttuner *st, sdo_st_channel *ch)
{
  const unsigned ws = st->window_size, sz = ch->size, hw = ch->halfw, hs = ch->halfsz;
  unsigned i;
  sdo_cpx *curr, *prev;
  memset(ch->fft, 0, sizeof(sdo_cpx) * sz);
  /* upper sideband: bins center + i -> fft[i]; lower: bins center - j -> fft[sz - j] */
  for (i = 0; i < hw; ++i) {
    unsigned src = (ch->center + i) % ws;
    ch->fft[i].re = st->fft[src].re * ch->h[i];
    ch->fft[i].im = st->fft[src].im * ch->h[i];
  }
  for (i = 1; i <= hw; ++i) {
    unsigned src = (ch->center + ws - i) % ws;
    ch->fft[sz - i].re = st->fft[src].re * ch->h[ws - i];
    ch->fft[sz - i].im = st->fft[src].im * ch->h[ws - i];
  }
  curr = ch->ifft[ch->state];
  prev = ch->ifft[!ch->state] + hs;
  memcpy(curr, ch->fft, sizeof(sdo_cpx) * sz);
  sdo_spec_fft_stockham(curr, sz, ch->tw, +1, ch->tmp);   /* SPEC S.4: inverse F.2 */
  for (i = 0; i < hs; ++i) {
    float al = ch->window[i], be = ch->window[i + hs];
    sdo_cpx o;
    o.re = al * curr[i].re + be * prev[i].re;
    o.im = al * curr[i].im + be * prev[i].im;
    if (ch->params.precise) {
      sdo_cpx ph = sdo_ncqo_read(&ch->lo);   /* multiply by conj(lo) */
      sdo_cpx t;
      t.re = o.re * ph.re + o.im * ph.im;
      t.im = o.im * ph.re - o.re * ph.im;
      o = t;
    }
    ch->out[i] = o;
  }
  ch->state = !ch->state;
  if (ch->params.on_data)
    return ch->params.on_data(ch, ch->params.privdata, ch->out, hs);
  return 1;
}

/* S.1: first FFT after window_size samples, then one per half window (50 % overlap). */
int sdo_specttuner_feed_bulk(sdo_specttuner *st, const sdo_cpx *x, size_t n)
{
  while (n > 0) {
    size_t room = st->window_size - st->p, take = n < room ? n : room;
    sdo_st_channel *ch;
    memcpy(st->window + st->p, x, take * sizeof(sdo_cpx));
    st->p += (unsigned) take; x += take; n -= take;
    if (st->p == st->window_size) {
      sdo_spec_forward(&st->plan, st->window, NULL, st->fft);
      ++st->hops;
      for (ch = st->channels; ch; ch = ch->next)
        if (!feed_channel(st, ch)) return 0;
      memmove(st->window, st->window + st->half_size, st->half_size * sizeof(sdo_cpx));
      st->p = st->half_size;
    }
  }
  return 1;
}
