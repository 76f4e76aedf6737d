/*
 * analyzer.c -- ORACLE (test infrastructure). One whole pass of the analyzer hot path over a stream,
 * SPEC.md section Z: main PSD over every non-overlapping N-sample frame (coverage 1),
 * FFT channeliser with 50 % overlap, one inspector per channel, hard decision.
 *
 * Mirrors the structure of suscan's source-worker loop as the reference sees it: PSD messages
 * (Suscan/Messages/PSDMessage.cpp:26-39), inspector sample batches keyed by inspector id
 * (include/Suscan/Messages/SamplesMessage.h:33-59), decision on the GUI side
 * (Default/GenericInspector/InspectorUI.cpp:836-846).
 */
#include "sd_oracle.h"
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
  sdo_analyzer   *owner;
  unsigned        index;
  sdo_st_channel *ch;
  sdo_inspector  *insp;
  sdo_decider     decider;
  /* per-call sinks */
  sdo_cpx *chan_out; size_t chan_cap, chan_n;
  sdo_cpx *sym_out; uint8_t *hard_out; size_t sym_cap, sym_n;
  sdo_cpx *tmp; size_t tmp_cap;
} an_chan;

struct sdo_analyzer {
  sdo_an_params   p;
  sdo_an_channel *chdefs;
  sdo_spec_plan   psd_plan;
  float          *window;
  int             psd_window_none;
  sdo_cpx        *frame, *scratch;
  unsigned        frame_fill;
  sdo_specttuner *st;
  an_chan        *chans;
};

static int on_data(const sdo_st_channel *ch, void *priv, const sdo_cpx *data, size_t n)
{
  an_chan *c = (an_chan *) priv;
  size_t i, m;
  (void) ch;
  if (c->chan_out) {
    for (i = 0; i < n && c->chan_n + i < c->chan_cap; ++i) c->chan_out[c->chan_n + i] = data[i];
  }
  c->chan_n += n;
  if (c->tmp_cap < n) {
    c->tmp = (sdo_cpx *) realloc(c->tmp, n * sizeof(sdo_cpx));
    c->tmp_cap = n;
  }
  m = sdo_inspector_feed(c->insp, data, n, c->tmp, n);
  for (i = 0; i < m; ++i) {
    if (c->sym_n + i < c->sym_cap) {
      if (c->sym_out) c->sym_out[c->sym_n + i] = c->tmp[i];
      if (c->hard_out) sdo_decider_decide(&c->decider, &c->tmp[i], &c->hard_out[c->sym_n + i], 1);
    }
  }
  c->sym_n += m;
  return 1;
}

sdo_analyzer *sdo_analyzer_new(const sdo_an_params *p)
{
  sdo_analyzer *a = (sdo_analyzer *) calloc(1, sizeof(*a));
  unsigned k, ws;
  if (!a) return NULL;
  a->p = *p;
  a->chdefs = (sdo_an_channel *) malloc(sizeof(sdo_an_channel) * (p->n_channels ? p->n_channels : 1));
  memcpy(a->chdefs, p->channels, sizeof(sdo_an_channel) * p->n_channels);
  a->p.channels = a->chdefs;
  if (sdo_spec_plan_init(&a->psd_plan, p->psd_size, 0) != 0) { free(a->chdefs); free(a); return NULL; }
  a->window = (float *) malloc(sizeof(float) * p->psd_size);
  sdo_window_fill(a->window, p->psd_size, p->psd_window);
  a->psd_window_none = p->psd_window == SDO_WINDOW_NONE;
  a->frame = (sdo_cpx *) malloc(sizeof(sdo_cpx) * p->psd_size);
  a->scratch = (sdo_cpx *) malloc(sizeof(sdo_cpx) * 2 * p->psd_size);
  ws = p->st_window_size ? p->st_window_size : p->psd_size;
  a->st = sdo_specttuner_new(ws);
  a->chans = (an_chan *) calloc(p->n_channels ? p->n_channels : 1, sizeof(an_chan));
  for (k = 0; k < p->n_channels; ++k) {
    sdo_st_channel_params cp;
    sdo_insp_config ic = p->channels[k].insp;
    memset(&cp, 0, sizeof(cp));
    cp.f0 = p->channels[k].f0; cp.bw = p->channels[k].bw; cp.guard = p->channels[k].guard;
    cp.precise = p->channels[k].precise;
    cp.privdata = &a->chans[k]; cp.on_data = on_data;
    a->chans[k].owner = a; a->chans[k].index = k;
    a->chans[k].ch = sdo_specttuner_open_channel(a->st, &cp);
    if (!a->chans[k].ch) { sdo_analyzer_destroy(a); return NULL; }
    a->chans[k].insp = sdo_inspector_new(&ic);
    sdo_inspector_decider(&ic, &a->chans[k].decider);
  }
  return a;
}

void sdo_analyzer_destroy(sdo_analyzer *a)
{
  unsigned k;
  if (!a) return;
  for (k = 0; k < a->p.n_channels; ++k) {
    sdo_inspector_destroy(a->chans[k].insp);
    free(a->chans[k].tmp);
  }
  sdo_specttuner_destroy(a->st);
  sdo_spec_plan_free(&a->psd_plan);
  free(a->window); free(a->frame); free(a->scratch); free(a->chans); free(a->chdefs); free(a);
}

int sdo_analyzer_feed(sdo_analyzer *a, const sdo_cpx *x, size_t n,
                      float *psd_out, size_t n_frames_cap,
                      sdo_cpx **chan_out, size_t chan_cap,
                      sdo_cpx **sym_out, uint8_t **hard_out, size_t sym_cap,
                      sdo_an_counts *counts)
{
  const unsigned N = a->p.psd_size;
  size_t frames = 0, off = 0;
  unsigned k;
  for (k = 0; k < a->p.n_channels; ++k) {
    an_chan *c = &a->chans[k];
    c->chan_out = chan_out ? chan_out[k] : NULL; c->chan_cap = chan_cap; c->chan_n = 0;
    c->sym_out = sym_out ? sym_out[k] : NULL; c->hard_out = hard_out ? hard_out[k] : NULL;
    c->sym_cap = sym_cap; c->sym_n = 0;
  }
  /* main PSD: consecutive non-overlapping frames */
  while (off < n) {
    size_t take = N - a->frame_fill;
    if (take > n - off) take = n - off;
    memcpy(a->frame + a->frame_fill, x + off, take * sizeof(sdo_cpx));
    a->frame_fill += (unsigned) take; off += take;
    if (a->frame_fill == N) {
      if (psd_out && frames < n_frames_cap)
        sdo_psd_frame_spec(&a->psd_plan, a->psd_window_none ? NULL : a->window, a->frame, psd_out + frames * N,
                           a->scratch);
      ++frames;
      a->frame_fill = 0;
    }
  }
  /* channeliser + inspectors */
  if (a->p.n_channels > 0)
    if (!sdo_specttuner_feed_bulk(a->st, x, n)) return -1;
  if (counts) {
    counts->n_frames = frames;
    for (k = 0; k < a->p.n_channels; ++k) {
      if (counts->n_chan) counts->n_chan[k] = a->chans[k].chan_n;
      if (counts->n_sym) counts->n_sym[k] = a->chans[k].sym_n;
    }
  }
  return 0;
}

/* CPU baseline: S independent streams sharded over OpenMP threads, each a full analyzer pass.
 * Returns the wall seconds of the PROCESSING only: analyzers, plans and output buffers are created before the
 * clock starts (a running analyzer does not rebuild its plans; round 1 timed the set-up too, ~40 ms per stream
 * for 64 channels, which handicapped the CPU arm on short samples).  *checksum folds every output so the work
 * cannot be optimised away. */
typedef struct {
  sdo_analyzer *a; float *psd; sdo_cpx **sym; uint8_t **hard; size_t *nsym; uint64_t acc;
} baseline_stream;

double sdo_baseline_run(const sdo_an_params *p, const sdo_cpx *x, size_t n_streams, size_t n,
                        int n_threads, uint64_t *checksum)
{
  struct timespec t0, t1;
  uint64_t total = 0;
  long s;
  const size_t N = p->psd_size, nf = n / N + 1, sym_cap = n / 2 + 16, nch = p->n_channels ? p->n_channels : 1;
  baseline_stream *bs = (baseline_stream *) calloc(n_streams ? n_streams : 1, sizeof(*bs));
#ifdef _OPENMP
  if (n_threads > 0) omp_set_num_threads(n_threads);
#else
  (void) n_threads;
#endif
#pragma omp parallel for schedule(dynamic, 1)
  for (s = 0; s < (long) n_streams; ++s) {
    baseline_stream *b = &bs[s];
    unsigned k;
    b->a = sdo_analyzer_new(p);
    b->psd = (float *) malloc(sizeof(float) * N * nf);
    b->sym = (sdo_cpx **) calloc(nch, sizeof(*b->sym));
    b->hard = (uint8_t **) calloc(nch, sizeof(*b->hard));
    b->nsym = (size_t *) calloc(nch, sizeof(size_t));
    for (k = 0; k < p->n_channels; ++k) {
      b->sym[k] = (sdo_cpx *) malloc(sizeof(sdo_cpx) * sym_cap);
      b->hard[k] = (uint8_t *) malloc(sym_cap);
    }
  }
  clock_gettime(CLOCK_MONOTONIC, &t0);
#pragma omp parallel for schedule(dynamic, 1)
  for (s = 0; s < (long) n_streams; ++s) {
    baseline_stream *b = &bs[s];
    sdo_an_counts cnt;
    unsigned k;
    size_t i;
    uint64_t acc = 0;
    memset(&cnt, 0, sizeof(cnt));
    cnt.n_sym = b->nsym;
    sdo_analyzer_feed(b->a, x + (size_t) s * n, n, b->psd, nf, NULL, 0, b->sym, b->hard, sym_cap, &cnt);
    for (i = 0; i < cnt.n_frames * N; i += 97) { uint32_t u; memcpy(&u, &b->psd[i], 4); acc += u; }
    for (k = 0; k < p->n_channels; ++k)
      for (i = 0; i < b->nsym[k] && i < sym_cap; ++i) acc += b->hard[k][i];
    b->acc = acc;
  }
  clock_gettime(CLOCK_MONOTONIC, &t1);
  for (s = 0; s < (long) n_streams; ++s) {
    baseline_stream *b = &bs[s];
    unsigned k;
    total += b->acc;
    for (k = 0; k < p->n_channels; ++k) { free(b->sym[k]); free(b->hard[k]); }
    free(b->sym); free(b->hard); free(b->nsym); free(b->psd);
    sdo_analyzer_destroy(b->a);
  }
  free(bs);
  if (checksum) *checksum = total;
  return (double) (t1.tv_sec - t0.tv_sec) + 1e-9 * (double) (t1.tv_nsec - t0.tv_nsec);
}
