/*
 * inspector.c -- ORACLE (test infrastructure). Per-channel inspector chains psk / fsk / ask / audio /
 * raw, SPEC.md section X.
 *
 * The inspector implementations live in suscan (absent from /root/reference).  What the reference
 * pins down, and this file follows:
 *   - class strings "psk"/"fsk"/"ask" (Default/Inspection/InspToolWidget.cpp:932,938,944), "audio"
 *     (Default/Audio/AudioProcessor.cpp:153), "raw" (InspToolWidget.cpp:612);
 *   - the config vocabulary and enum values (Default/GenericInspector/InspectorCtl/AfcControl.cpp:54-83,
 *     GainControl.cpp:51-60, ToneControl.cpp:59-81, AskControl.cpp:53-77, MfControl.cpp:56-78,
 *     ClockRecovery.cpp:59-93; audio keys Default/Audio/AudioProcessor.cpp:257-265);
 *   - block order of each chain (doc/SigDigger_User_Manual.pdf pp.50-52): gain (AGC or manual) ->
 *     carrier stage -> SRRC matched filter -> manual sampler | Gardner -> x0.75 (-2.5 dB);
 *   - samples flow only while clock.running is true (manual p.62; ClockRecovery.cpp:90);
 *   - AGC time constants as fractions of the symbol period (Tasks/AGCTask.cpp:22-28 mirrors them,
 *     doubled), matched filter span of 6 symbols (include/WaveSampler.h:29-30), Costas arm order 3
 *     (Tasks/CostasRecoveryTask.cpp:41).
 */
#include "sd_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

struct sdo_inspector {
  sdo_insp_config cfg;
  float      T, bnor, gain2;
  int        have_agc, have_costas, have_mf, have_pll, have_lo;
  sdo_agc    agc;
  sdo_costas costas;
  sdo_pll    pll;
  sdo_ncqo   lo;
  sdo_filt   mf;
  sdo_clock  cd;
  sdo_sampler sampler;
  sdo_equalizer eq;
  sdo_cpx    prev;        /* fsk / fm discriminator memory */
  sdo_cpx    fsk_rot;
  /* audio */
  sdo_filt   alpf;
  int        have_alpf;
  float      dc, dc_alpha, sq_level, sq_alpha, sq_thr;
  double     rs_phase, rs_step;
  sdo_cpx    rs_prev;
};

void sdo_insp_config_default(sdo_insp_config *c, int insp_class, float fs)
{
  memset(c, 0, sizeof(*c));
  c->insp_class = insp_class;
  c->fs = fs;
  c->agc_enabled = 1;
  c->agc_gain_db = 0.0f;
  c->costas_order = insp_class == SDO_INSP_PSK ? 2 : 0;
  c->bits_per_symbol = insp_class == SDO_INSP_PSK ? 2 : 1;
  c->loop_bw = fs * 1e-3f;
  c->offset = 0.0f;
  c->fsk_phase = 0.0f;
  c->fsk_quad_demod = 0;
  c->ask_use_pll = 0;
  c->ask_channel = 0;
  c->mf_type = 0;
  c->mf_rolloff = 0.35f;
  c->clock_type = 1;
  c->baud = fs * 0.25f;
  c->clock_gain = 1.0f;
  c->clock_phase = 0.0f;
  c->clock_running = 1;
  c->audio_cutoff = 5000.0f;
  c->audio_volume = 1.0f;
  c->audio_sample_rate = 44100;
  c->audio_demod = SDO_AUDIO_FM;
  c->audio_squelch = 0;
  c->audio_squelch_level = 0.0f;
  c->agc_ts = 0.1f;
  c->eq_type = 0;
  c->eq_rate = 1e-3f;
  c->eq_locked = 0;
}

static float clampf(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }

sdo_inspector *sdo_inspector_new(const sdo_insp_config *c)
{
  sdo_inspector *s = (sdo_inspector *) calloc(1, sizeof(*s));
  sdo_agc_params ap;
  float taps[SDO_FILT_MAX_TAPS];
  if (!s) return NULL;
  s->cfg = *c;
  s->bnor = c->baud / c->fs;
  if (s->bnor > 1.0f) s->bnor = 1.0f;
  if (!(s->bnor > 0.0f)) s->bnor = 1e-6f;
  s->T = 1.0f / s->bnor;
  s->gain2 = 2.0f * sdo_db_to_mag(c->agc_gain_db);

  if (c->insp_class == SDO_INSP_RAW)
    return s;

  if (c->insp_class == SDO_INSP_AUDIO) {
    float b[17], a[17];
    float tau = c->agc_ts * c->fs;
    if (tau < 2.0f) tau = 2.0f;
    if (c->agc_enabled) {
      sdo_agc_params_from_tau(&ap, tau, 1.0f);
      sdo_agc_init(&s->agc, &ap);
      s->have_agc = 1;
    }
    /* audio low-pass: 4th-order Butterworth at audio.cutoff */
    if (sdo_butter_lp(4, clampf(2.0f * c->audio_cutoff / c->fs, 1e-4f, 0.95f), b, a) == 0) {
      sdo_filt_init(&s->alpf, 5, a, 5, b);
      s->have_alpf = 1;
    }
    s->dc_alpha = (float) (1.0 - exp(-1.0 / (0.05 * (double) c->fs)));
    s->sq_alpha = (float) (1.0 - exp(-1.0 / (0.01 * (double) c->fs)));
    s->sq_thr = c->audio_squelch_level;
    s->rs_step = (double) c->audio_sample_rate / (double) c->fs;
    s->rs_phase = 0.0;
    return s;
  }

  if (c->agc_enabled) {
    sdo_agc_params_from_tau(&ap, s->T, 1.0f);
    sdo_agc_init(&s->agc, &ap);
    s->have_agc = 1;
  }

  switch (c->insp_class) {
    case SDO_INSP_PSK:
      if (c->costas_order > 0) {
        sdo_costas_init(&s->costas, (int) c->costas_order, 0.0f,
                        clampf(2.0f * s->bnor, 1e-3f, 0.95f), 3, 2.0f * c->loop_bw / c->fs);
        s->have_costas = 1;
      } else {
        sdo_ncqo_init(&s->lo, 2.0f * c->offset / c->fs);
        s->have_lo = 1;
      }
      break;
    case SDO_INSP_FSK:
      sdo_sincosf(c->fsk_phase, &s->fsk_rot.im, &s->fsk_rot.re);
      break;
    case SDO_INSP_ASK:
      if (c->ask_use_pll) {
        sdo_pll_init(&s->pll, 0.0f, 2.0f * c->loop_bw / c->fs);
        s->have_pll = 1;
      } else {
        sdo_ncqo_init(&s->lo, 2.0f * c->offset / c->fs);
        s->have_lo = 1;
      }
      break;
    default:
      break;
  }

  if (c->mf_type == 1) {
    unsigned n = sdo_mf_span(s->T);
    sdo_taps_rrc(taps, n, s->T, c->mf_rolloff);
    sdo_filt_init(&s->mf, 0, NULL, n, taps);
    s->have_mf = 1;
  }

  sdo_equalizer_init(&s->eq, c->eq_rate, c->eq_locked);
  sdo_clock_init(&s->cd, c->clock_gain, s->bnor);
  sdo_sampler_init(&s->sampler, s->bnor);
  sdo_sampler_set_phase(&s->sampler, c->clock_phase);
  return s;
}

void sdo_inspector_destroy(sdo_inspector *s)
{
  if (!s) return;
  if (s->have_agc) sdo_agc_free(&s->agc);
  if (s->have_costas) sdo_costas_free(&s->costas);
  if (s->have_mf) sdo_filt_free(&s->mf);
  if (s->have_alpf) sdo_filt_free(&s->alpf);
  free(s);
}

static size_t feed_audio(sdo_inspector *s, const sdo_cpx *x, size_t n, sdo_cpx *out, size_t cap)
{
  size_t i, k = 0;
  const sdo_insp_config *c = &s->cfg;
  for (i = 0; i < n; ++i) {
    sdo_cpx y = x[i], d, o;
    float v = 0.0f, p;
    if (s->have_agc) y = sdo_agc_feed(&s->agc, y);
    /* squelch detector runs on the (gain-controlled) channel power */
    p = y.re * y.re + y.im * y.im;
    s->sq_level = s->sq_level + s->sq_alpha * (p - s->sq_level);
    switch (c->audio_demod) {
      case SDO_AUDIO_AM:
        v = sdo_cabsf(y);
        s->dc = s->dc + s->dc_alpha * (v - s->dc);
        v = v - s->dc;
        break;
      case SDO_AUDIO_FM:
        d.re = y.re * s->prev.re + y.im * s->prev.im;
        d.im = y.im * s->prev.re - y.re * s->prev.im;
        v = sdo_atan2f(d.im, d.re) * (float) (1.0 / SDO_PI);
        s->prev = y;
        break;
      case SDO_AUDIO_USB:
      case SDO_AUDIO_LSB:
        /* The GUI tunes the channel to carrier +- bw/2 and halves bw
         * (Default/Audio/AudioProcessor.cpp:200-228); the sideband is already centred, so the
         * audio is the real part after shifting back by -+ bw/2, done by the caller-set offset. */
        if (!s->have_lo) {
          float fo = c->audio_demod == SDO_AUDIO_USB ? c->offset : -c->offset;
          sdo_ncqo_init(&s->lo, 2.0f * fo / c->fs);
          s->have_lo = 1;
        }
        {
          sdo_cpx ph = sdo_ncqo_read(&s->lo);
          v = y.re * ph.re - y.im * ph.im;
        }
        break;
      default:
        v = 0.0f;
        break;
    }
    if (c->audio_squelch && !(s->sq_level > s->sq_thr)) v = 0.0f;
    o.re = v; o.im = 0.0f;
    if (s->have_alpf) o = sdo_filt_feed(&s->alpf, o);
    /* linear-interpolating resampler to audio.sample-rate */
    s->rs_phase += s->rs_step;
    if (s->rs_phase >= 1.0) {
      float al;
      s->rs_phase -= 1.0;
      al = (float) (s->rs_phase / s->rs_step);   /* fraction of a channel sample past the instant */
      if (al > 1.0f) al = 1.0f;
      if (k < cap) {
        out[k].re = c->audio_volume * ((1.0f - al) * o.re + al * s->rs_prev.re);
        out[k].im = 0.0f;
        ++k;
      }
    }
    s->rs_prev = o;
  }
  return k;
}

size_t sdo_inspector_feed(sdo_inspector *s, const sdo_cpx *x, size_t n, sdo_cpx *out, size_t cap)
{
  const sdo_insp_config *c = &s->cfg;
  size_t i, k = 0;

  if (c->insp_class == SDO_INSP_RAW) {
    size_t m = n < cap ? n : cap;
    memcpy(out, x, m * sizeof(sdo_cpx));
    return m;
  }
  if (c->insp_class == SDO_INSP_AUDIO)
    return feed_audio(s, x, n, out, cap);

  for (i = 0; i < n; ++i) {
    sdo_cpx y = x[i], o;
    int produced;

    /* manual carrier offset (afc.offset / ask.offset): x * conj(lo) */
    if (s->have_lo) {
      sdo_cpx ph = sdo_ncqo_read(&s->lo), t;
      t.re = y.re * ph.re + y.im * ph.im;
      t.im = y.im * ph.re - y.re * ph.im;
      y = t;
    }

    /* gain stage: 2 * agc(x) or 2 * 10^(g/20) * x */
    if (s->have_agc) {
      y = sdo_agc_feed(&s->agc, y);
      y.re = 2.0f * y.re; y.im = 2.0f * y.im;
    } else {
      y.re = s->gain2 * y.re; y.im = s->gain2 * y.im;
    }

    switch (c->insp_class) {
      case SDO_INSP_PSK:
        if (s->have_costas) y = sdo_costas_feed(&s->costas, y);
        break;
      case SDO_INSP_FSK: {
        sdo_cpx d;
        d.re = y.re * s->prev.re + y.im * s->prev.im;
        d.im = y.im * s->prev.re - y.re * s->prev.im;
        s->prev = y;
        if (c->fsk_quad_demod) {
          y.re = sdo_atan2f(d.im, d.re) * (float) (1.0 / SDO_PI);
          y.im = 0.0f;
        } else {
          y.re = d.re * s->fsk_rot.re - d.im * s->fsk_rot.im;
          y.im = d.re * s->fsk_rot.im + d.im * s->fsk_rot.re;
        }
        break;
      }
      case SDO_INSP_ASK:
        if (s->have_pll) y = sdo_pll_track(&s->pll, y);
        if (c->ask_channel == 0)      { y.re = sdo_cabsf(y); y.im = 0.0f; }
        else if (c->ask_channel == 1) { y.im = 0.0f; }
        else                          { y.re = y.im; y.im = 0.0f; }
        break;
      default:
        break;
    }

    if (s->have_mf) y = sdo_filt_feed(&s->mf, y);

    if (c->clock_type == 1) produced = sdo_clock_feed(&s->cd, y, &o);
    else                    produced = sdo_sampler_feed(&s->sampler, y, &o);

    /* CMA equaliser on the symbol stream (manual p.50: "... sampler / CR -> CMA (opt) -> -2.5 dB") */
    if (produced && c->eq_type == 1) o = sdo_equalizer_feed(&s->eq, o);

    if (produced && c->clock_running && k < cap) {
      out[k].re = 0.75f * o.re;
      out[k].im = 0.75f * o.im;
      ++k;
    }
  }
  return k;
}

void sdo_inspector_decider(const sdo_insp_config *c, sdo_decider *d)
{
  if (c->insp_class == SDO_INSP_ASK)
    sdo_decider_init(d, SDO_DECIDE_MODULUS, c->bits_per_symbol, 0.0f, 1.0f);
  else
    sdo_decider_init(d, SDO_DECIDE_ARGUMENT, c->bits_per_symbol, -SDO_PI_F, SDO_PI_F);
}
