/*
 * sd_oracle.h -- CPU ORACLE (TEST INFRASTRUCTURE ONLY) for the sigdigger-b200 hot path.
 *
 * PARITY UNPINNED: the reference repository (/root/reference, BatchDrake/SigDigger) contains only
 * the Qt GUI.  The arithmetic of this path lives in sigutils / suscan / SuWidgets, which are cloned
 * from `master` at build time (Scripts/dist-common.sh:330-334) and are NOT present here, and the
 * reference ships no tests, fixtures or golden vectors.  This oracle is a plain-C
 * restatement of the published algorithms of those libraries as pinned down by (a) the reference's
 * own call sites (cited per function), (b) the user manual's DSP-chain chapter and (c) oracle/SPEC.md.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * link or call this code.  The product (sigdigger_b200/) never does.
 *
 * Number formats follow the reference: SUFLOAT = float, SUCOMPLEX = interleaved float pair,
 * SUFREQ = double, SUSCOUNT = uint64 ("Number formats").
 *
 * Build flags that matter: -O2 -ffp-contract=off (no FMA contraction) so that every recurrence here
 * is a sequence of IEEE-754 binary32 + - * / sqrt in the order written.
 */
#ifndef SD_ORACLE_H
#define SD_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { float re, im; } sdo_cpx;

#define SDO_PI_F   3.14159265358979323846f
#define SDO_2PI_F  6.28318530717958647692f
#define SDO_PI     3.14159265358979323846

/* ------------------------------------------------------------------------------------------------
 * Deterministic float32 math (SPEC.md section M).  Same algorithm, same operation order as the
 * device implementation, so loop recurrences built on them are bit-identical on CPU and GPU.
 * ---------------------------------------------------------------------------------------------- */
void  sdo_sincosf(float x, float *s, float *c);
float sdo_atan2f(float y, float x);
float sdo_log10f(float x);          /* x > 0 */
float sdo_exp10f(float x);          /* 10^x, clamped to the finite float range */
float sdo_cabsf(sdo_cpx z);         /* sqrtf(re*re + im*im) */
float sdo_db_to_mag(float db);      /* 10^(db/20) */
float sdo_power_db(float p);        /* SU_POWER_DB: 10*log10(p + 1e-20)-like, see SPEC M.6 */

/* ------------------------------------------------------------------------------------------------
 * FFT (SPEC.md section F): power-of-two, radix-2 DIT, float32 data, twiddles rounded from double.
 * sign = -1 forward (exp(-i 2 pi k n / N)), +1 backward, both unnormalised (FFTW convention, which
 * is what sigutils drives through SU_FFTW(); seen at Tasks/CarrierDetector.cpp:58-75).
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  unsigned  n, log2n;
  float    *tw_re, *tw_im;   /* n/2 twiddles cos, -sin(2 pi k / n) */
  unsigned *rev;
} sdo_fft_plan;

int  sdo_fft_plan_init(sdo_fft_plan *p, unsigned n);
void sdo_fft_plan_free(sdo_fft_plan *p);
void sdo_fft_exec(const sdo_fft_plan *p, const sdo_cpx *in, sdo_cpx *out, int sign);

/* SPEC FFT (SPEC.md F.2-F.4, fft_spec.c): the fixed dataflows the CUDA kernels also follow, so that
 * transforms are bit-identical on both sides.  sdo_fft_exec above is the independent cross-check. */
typedef struct {
  unsigned N, N1, N2;
  int      four, kind;       /* kind 0: single Stockham, 1: four-step, 2: 65536 = 256 x 256 fft16 form */
  sdo_cpx *tw_a, *tw_b, *tw_n, *scr, *buf, *tmp;
} sdo_spec_plan;
/* SPEC R: block-wise DC removal of the source worker (tasks.c) */
void sdo_dc_remove(float c[2], const sdo_cpx *x, sdo_cpx *y, size_t n, float alpha);
/* speed leg of the CPU baseline (fft_fast.c): vectorisable Stockham transform, float32-rounding-equal to the SPEC
 * transforms, NOT bit-identical; enabled only by bench.py's CPU legs, never by a parity test */
extern int sdo_fast_transforms;
void sdo_set_fast_transforms(int on);
void sdo_fast_fft(const sdo_cpx *in, const float *window, sdo_cpx *out, unsigned n, int sign);
sdo_cpx *sdo_spec_twiddles(unsigned n);
void sdo_spec_fft_stockham(sdo_cpx *s, unsigned M, const sdo_cpx *tw, int sign, sdo_cpx *tmp);
int  sdo_spec_plan_init(sdo_spec_plan *p, unsigned N, int four);
void sdo_spec_plan_free(sdo_spec_plan *p);
void sdo_spec_forward(const sdo_spec_plan *p, const sdo_cpx *x, const float *window, sdo_cpx *X);
void sdo_spec_inverse_stockham(const sdo_spec_plan *p, sdo_cpx *s);
/* main PSD with the SPEC transform: psd[k] = (re^2 + im^2) * (1/N) */
void sdo_psd_frame_spec(const sdo_spec_plan *p, const float *window, const sdo_cpx *x, float *psd,
                        sdo_cpx *scratch /* N */);

/* ------------------------------------------------------------------------------------------------
 * Windows (SPEC.md section W; su_taps_apply_*_complex seen at Tasks/CarrierDetector.cpp:87-89,
 * enum order at include/Suscan/AnalyzerParams.h:37-43).
 * ---------------------------------------------------------------------------------------------- */
enum sdo_window { SDO_WINDOW_NONE = 0, SDO_WINDOW_HAMMING, SDO_WINDOW_HANN,
                  SDO_WINDOW_FLAT_TOP, SDO_WINDOW_BLACKMANN_HARRIS };
void sdo_window_fill(float *w, unsigned n, int type);

/* Main PSD (SPEC section P; fields consumed at Suscan/Messages/PSDMessage.cpp:26-39). */
void sdo_psd_frame(const sdo_fft_plan *p, const float *window, const sdo_cpx *x, float *psd,
                   sdo_cpx *scratch /* 2n */);
/* GUI-side post-processing: fft-shift + dB in place (Suscan/Messages/PSDMessage.cpp:32-38). */
void sdo_psd_shift_db(float *psd, unsigned n);
/* EMA in dB (Misc/Averager.cpp:43-49). */
void sdo_averager_feed(float *last, const float *frame, unsigned n, float alpha);

/* ------------------------------------------------------------------------------------------------
 * NCQO (SPEC section N; su_ncqo_init/set_phase/read seen at Tasks/CarrierXlator.cpp:36-37,57-60).
 * ---------------------------------------------------------------------------------------------- */
typedef struct { float phi, omega; } sdo_ncqo;
void    sdo_ncqo_init(sdo_ncqo *o, float fnor);
void    sdo_ncqo_set_phase(sdo_ncqo *o, float phi);
sdo_cpx sdo_ncqo_read(sdo_ncqo *o);
void    sdo_ncqo_inc_phase(sdo_ncqo *o, float d);
void    sdo_ncqo_inc_angfreq(sdo_ncqo *o, float d);

/* ------------------------------------------------------------------------------------------------
 * IIR / FIR filter (SPEC section I; su_iir_filt_feed / su_iir_rrc_init seen at
 * Tasks/WaveSampler.cpp:68-80, span constants include/WaveSampler.h:29-30).
 * ---------------------------------------------------------------------------------------------- */
#define SDO_FILT_MAX_TAPS 1024
typedef struct {
  unsigned nb, na;           /* number of b (x) and a (y) coefficients; a[0] == 1 */
  float   *b, *a;
  sdo_cpx *x, *y;            /* circular delay lines, length nb / na */
  unsigned xp, yp;
  float    gain;
} sdo_filt;
int     sdo_filt_init(sdo_filt *f, unsigned na, const float *a, unsigned nb, const float *b);
void    sdo_filt_free(sdo_filt *f);
sdo_cpx sdo_filt_feed(sdo_filt *f, sdo_cpx x);
/* Tap design, computed in double and rounded to float (host-side set-up, not on the hot path). */
void    sdo_taps_rrc(float *h, unsigned n, float T, float beta);
void    sdo_taps_brickwall_lp(float *h, unsigned n, float fc);
int     sdo_butter_lp(unsigned order, float fc, float *b, float *a); /* order+1 coefs each */
unsigned sdo_mf_span(float T);  /* ceil(6*T) clamped to [1, 1024] */

/* ------------------------------------------------------------------------------------------------
 * AGC (SPEC section A; struct su_agc_params fields seen at Tasks/AGCTask.cpp:43-47, fractions
 * Tasks/AGCTask.cpp:22-28).
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  float    threshold, slope_factor;
  unsigned hang_max, delay_line_size, mag_history_size;
  float    fast_rise_t, fast_fall_t, slow_rise_t, slow_fall_t;
} sdo_agc_params;
#define SDO_AGC_MAX_HISTORY 4096
typedef struct {
  int      enabled;
  float    knee, gain_slope, fixed_gain;
  unsigned hang_max, hang_n;
  float    fast_alpha_rise, fast_alpha_fall, slow_alpha_rise, slow_alpha_fall;
  float    fast_level, slow_level, peak;
  unsigned delay_line_size, delay_line_ptr, mag_history_size, mag_history_ptr;
  sdo_cpx *delay_line;
  float   *mag_history;
} sdo_agc;
void    sdo_agc_params_default(sdo_agc_params *p);
/* Inspector-style derivation of all AGC constants from the symbol period (SPEC A.3). */
void    sdo_agc_params_from_tau(sdo_agc_params *p, float tau, float frac_scale);
int     sdo_agc_init(sdo_agc *a, const sdo_agc_params *p);
void    sdo_agc_free(sdo_agc *a);
sdo_cpx sdo_agc_feed(sdo_agc *a, sdo_cpx x);

/* ------------------------------------------------------------------------------------------------
 * PLL and Costas loop (SPEC section C; su_pll_init/track seen at Tasks/PLLSyncTask.cpp:36,53-56;
 * su_costas_init/feed at Tasks/CostasRecoveryTask.cpp:36-41,58-61; kinds
 * include/CostasRecoveryTask.h:50).
 * ---------------------------------------------------------------------------------------------- */
typedef struct { float alpha, beta; sdo_ncqo ncqo; } sdo_pll;
void    sdo_pll_init(sdo_pll *p, float fhint, float fc);
sdo_cpx sdo_pll_track(sdo_pll *p, sdo_cpx x);

enum sdo_costas_kind { SDO_COSTAS_NONE = 0, SDO_COSTAS_BPSK, SDO_COSTAS_QPSK, SDO_COSTAS_8PSK };
typedef struct {
  int      kind;
  float    a, b, y_alpha, gain, lock;
  sdo_cpx  y, z;
  sdo_ncqo ncqo;
  sdo_filt af;
} sdo_costas;
int     sdo_costas_init(sdo_costas *c, int kind, float fhint, float arm_bw, unsigned arm_order,
                        float loop_bw);
void    sdo_costas_free(sdo_costas *c);
sdo_cpx sdo_costas_feed(sdo_costas *c, sdo_cpx x);

/* ------------------------------------------------------------------------------------------------
 * Clock recovery (SPEC section G; su_clock_detector_init/feed/read seen at
 * Tasks/WaveSampler.cpp:60-66,190-205) and manual sampler (keys at
 * Default/GenericInspector/InspectorCtl/ClockRecovery.cpp:59-93).
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  float   alpha, beta, bnor, bmin, bmax, phi, gain, e;
  int     halfcycle;
  sdo_cpx x[3], prev;
} sdo_clock;
void sdo_clock_init(sdo_clock *c, float loop_gain, float bhint);
/* returns 1 and writes *out when a symbol is produced */
int  sdo_clock_feed(sdo_clock *c, sdo_cpx v, sdo_cpx *out);

typedef struct { float bnor, period, phase, phase0_rel, phase0; sdo_cpx prev; } sdo_sampler;
void sdo_sampler_init(sdo_sampler *s, float bnor);
void sdo_sampler_set_phase(sdo_sampler *s, float phase0_rel);
int  sdo_sampler_feed(sdo_sampler *s, sdo_cpx v, sdo_cpx *out);

/* ------------------------------------------------------------------------------------------------
 * Decider (SPEC section D; set-up at Default/GenericInspector/InspectorUI.cpp:228-253, feed
 * :836-846, offline use Tasks/WaveSampler.cpp:316-317).
 * ---------------------------------------------------------------------------------------------- */
enum sdo_decision_mode { SDO_DECIDE_ARGUMENT = 0, SDO_DECIDE_MODULUS = 1 };
typedef struct { int mode; unsigned bps, intervals; float min, max, h; } sdo_decider;
void sdo_decider_init(sdo_decider *d, int mode, unsigned bps, float min, float max);
void sdo_decider_decide(const sdo_decider *d, const sdo_cpx *x, uint8_t *sym, size_t n);

/* ------------------------------------------------------------------------------------------------
 * Quadrature demodulator (in-repo: Tasks/QuadDemodTask.cpp:44-60) and carrier xlator
 * (Tasks/CarrierXlator.cpp:57-60).
 * ---------------------------------------------------------------------------------------------- */
void sdo_quad_demod(const sdo_cpx *x, sdo_cpx *y, size_t n, sdo_cpx *prev, int *primed);
void sdo_carrier_xlate(const sdo_cpx *x, sdo_cpx *y, size_t n, sdo_ncqo *o);

/* ------------------------------------------------------------------------------------------------
 * Spectral tuner = FFT filter-bank channeliser (SPEC section S; su_specttuner_new /
 * open_channel / feed_bulk / destroy + on_data contract seen at Tasks/LPFTask.cpp:28-42,52-69,
 * 83-87,104-107).
 * ---------------------------------------------------------------------------------------------- */
typedef struct sdo_specttuner sdo_specttuner;
typedef struct sdo_st_channel sdo_st_channel;
typedef int (*sdo_on_data_fn)(const sdo_st_channel *ch, void *priv, const sdo_cpx *data, size_t n);

typedef struct {
  float f0;      /* centre, rad/sample in [0, 2 pi) */
  float delta_f; /* unused by this path */
  float bw;      /* rad/sample */
  float guard;   /* >= 1 */
  int   precise;
  void *privdata;
  sdo_on_data_fn on_data;
} sdo_st_channel_params;

struct sdo_st_channel {
  sdo_st_channel_params params;
  unsigned center, size, width, halfw, halfsz;
  float    k, gain, decimation;
  sdo_ncqo lo;
  float   *h;        /* window_size real shaping response (SPEC S.3) */
  float   *window;   /* size cross-fade weights */
  sdo_cpx *fft;      /* size */
  sdo_cpx *ifft[2];  /* size each */
  sdo_cpx *out;      /* halfsz */
  int      state;
  sdo_cpx *tw, *tmp; /* W_size table and scratch of the inverse transform */
  struct sdo_st_channel *next;
};

sdo_specttuner *sdo_specttuner_new(unsigned window_size);
void            sdo_specttuner_destroy(sdo_specttuner *st);
sdo_st_channel *sdo_specttuner_open_channel(sdo_specttuner *st, const sdo_st_channel_params *p);
int             sdo_specttuner_feed_bulk(sdo_specttuner *st, const sdo_cpx *x, size_t n);
unsigned        sdo_specttuner_window_size(const sdo_specttuner *st);
/* channel geometry rule on its own (used by tests and by host code twins) */
void sdo_st_channel_geometry(unsigned window_size, float f0, float bw, float guard,
                             unsigned *center, unsigned *size, unsigned *width);
void sdo_st_filter_response(unsigned window_size, unsigned halfw, float *h);

/* ------------------------------------------------------------------------------------------------
 * Inspector chains (SPEC section X; class strings at Default/Inspection/InspToolWidget.cpp:932,938,
 * 944; config vocabulary Default/GenericInspector/InspectorCtl/ sources; block order from
 * doc/SigDigger_User_Manual.pdf pp.50-52).
 * ---------------------------------------------------------------------------------------------- */
enum sdo_insp_class { SDO_INSP_PSK = 0, SDO_INSP_FSK = 1, SDO_INSP_ASK = 2, SDO_INSP_AUDIO = 3,
                      SDO_INSP_RAW = 4 };
enum sdo_audio_demod { SDO_AUDIO_DISABLED = 0, SDO_AUDIO_AM, SDO_AUDIO_FM, SDO_AUDIO_USB,
                       SDO_AUDIO_LSB };

typedef struct {
  int      insp_class;
  float    fs;                 /* channel (equivalent) sample rate, Hz */
  /* gain */
  int      agc_enabled;        /* agc.enabled */
  float    agc_gain_db;        /* agc.gain [dB] */
  /* carrier */
  unsigned costas_order;       /* afc.costas-order 0 manual,1 BPSK,2 QPSK,3 8PSK */
  unsigned bits_per_symbol;    /* afc./fsk./ask.bits-per-symbol */
  float    loop_bw;            /* afc.loop-bw / ask.loop-bw [Hz] */
  float    offset;             /* afc.offset / ask.offset [Hz] */
  /* fsk */
  float    fsk_phase;          /* fsk.phase [rad] */
  int      fsk_quad_demod;     /* fsk.quad-demod */
  /* ask */
  int      ask_use_pll;        /* ask.use-pll */
  unsigned ask_channel;        /* ask.channel: 0 = |x|, 1 = I, 2 = Q (SPEC X.4) */
  /* matched filter */
  unsigned mf_type;            /* 0 bypass, 1 manual (RRC) */
  float    mf_rolloff;
  /* clock */
  unsigned clock_type;         /* 0 manual, 1 gardner */
  float    baud, clock_gain, clock_phase;
  int      clock_running;
  /* audio */
  float    audio_cutoff, audio_volume, audio_squelch_level, agc_ts;
  unsigned audio_sample_rate, audio_demod;
  int      audio_squelch;
  /* equalizer (Default/GenericInspector/InspectorCtl/EqualizerControl.cpp:56-75) */
  unsigned eq_type;            /* equalizer.type: 0 bypass, 1 CMA */
  float    eq_rate;            /* equalizer.rate (mu) */
  int      eq_locked;          /* equalizer.locked: weights frozen */
} sdo_insp_config;

/* CMA equaliser (SPEC section E) */
#define SDO_EQ_LEN 10
typedef struct { float mu; int locked; sdo_cpx w[SDO_EQ_LEN], x[SDO_EQ_LEN]; } sdo_equalizer;
void    sdo_equalizer_init(sdo_equalizer *e, float mu, int locked);
sdo_cpx sdo_equalizer_feed(sdo_equalizer *e, sdo_cpx x);

void sdo_insp_config_default(sdo_insp_config *c, int insp_class, float fs);

typedef struct sdo_inspector sdo_inspector;
sdo_inspector *sdo_inspector_new(const sdo_insp_config *c);
void           sdo_inspector_destroy(sdo_inspector *i);
/* feeds n channel-rate samples; appends produced samples to out (capacity cap); returns count */
size_t sdo_inspector_feed(sdo_inspector *i, const sdo_cpx *x, size_t n, sdo_cpx *out, size_t cap);
/* decider the GUI would attach to this inspector (Default/GenericInspector/InspectorUI.cpp:228-253) */
void   sdo_inspector_decider(const sdo_insp_config *c, sdo_decider *d);

/* ------------------------------------------------------------------------------------------------
 * Panoramic SpectrumView (in-repo, fully specified: Panoramic/Scanner.cpp:36-293,
 * constants include/Scanner.h:26-32).
 * ---------------------------------------------------------------------------------------------- */
#define SDO_SCANNER_SPECTRUM_SIZE 65536
#define SDO_SCANNER_FREQ_RESOLUTION 1e3
#define SDO_SCANNER_COUNT_MAX 5.f
#define SDO_SCANNER_COUNT_RESET 1.f
#define SDO_SCANNER_DEFAULT_BIN_VALUE (-200.0f)
typedef struct {
  double   freq_min, freq_max, freq_range, fft_bandwidth;
  float    fft_rel_bw;
  unsigned spectrum_size;
  float   *psd, *psd_accum, *psd_count;  /* SDO_SCANNER_SPECTRUM_SIZE each */
} sdo_spectrum_view;
int  sdo_sview_init(sdo_spectrum_view *v);
void sdo_sview_free(sdo_spectrum_view *v);
void sdo_sview_reset(sdo_spectrum_view *v);
void sdo_sview_set_range(sdo_spectrum_view *v, double fmin, double fmax);
void sdo_sview_feed_linear(sdo_spectrum_view *v, const float *psd, const float *count,
                           size_t psd_size, double fmin, double fmax, int adjust_sides);
void sdo_sview_feed_histogram(sdo_spectrum_view *v, const float *psd, size_t psd_size,
                              double fmin, double fmax);
void sdo_sview_interpolate(sdo_spectrum_view *v);
void sdo_sview_feed(sdo_spectrum_view *v, const float *psd, const float *count, size_t psd_size,
                    double center, int adjust_sides);
/* the explicit-range overload (Scanner.cpp:239-256) and feed(SpectrumView const &) (Scanner.cpp:276-286): the
 * zoom path of Scanner::setViewRange (Scanner.cpp:471-479) feeds the previous view into the new one */
void sdo_sview_feed_range(sdo_spectrum_view *v, const float *psd, const float *count, size_t psd_size,
                          double fmin, double fmax, int adjust_sides);
void sdo_sview_feed_view(sdo_spectrum_view *v, const sdo_spectrum_view *detail);

/* ------------------------------------------------------------------------------------------------
 * Whole analyzer pass over one stream (SPEC section Z): main PSD over every non-overlapping
 * frame + channeliser + inspectors + decision.  Used by parity tests and the CPU baseline.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  float    f0, bw, guard;      /* rad/sample */
  int      precise;
  sdo_insp_config insp;
} sdo_an_channel;

typedef struct {
  unsigned psd_size;           /* main FFT size N */
  int      psd_window;
  unsigned st_window_size;     /* channeliser window (0 = same as psd_size) */
  unsigned n_channels;
  const sdo_an_channel *channels;
} sdo_an_params;

typedef struct {
  size_t   n_frames;           /* PSD frames written */
  size_t  *n_chan;             /* per channel: channel-rate samples produced */
  size_t  *n_sym;              /* per channel: inspector output samples produced */
} sdo_an_counts;

typedef struct sdo_analyzer sdo_analyzer;
sdo_analyzer *sdo_analyzer_new(const sdo_an_params *p);
void          sdo_analyzer_destroy(sdo_analyzer *a);
/* psd_out: n_frames_cap * N floats (may be NULL); chan_out[k]/sym_out[k]/hard_out[k] per channel
 * buffers with capacities chan_cap / sym_cap (any may be NULL). State persists across calls. */
int sdo_analyzer_feed(sdo_analyzer *a, const sdo_cpx *x, size_t n,
                      float *psd_out, size_t n_frames_cap,
                      sdo_cpx **chan_out, size_t chan_cap,
                      sdo_cpx **sym_out, uint8_t **hard_out, size_t sym_cap,
                      sdo_an_counts *counts);

/* ---- K: channel detector on the main PSD (chdetect.c, SPEC.md section K) ---- */
typedef struct {
  unsigned bin_lo, bin_hi;     /* [lo, hi) in ascending-frequency (fft-shifted) bin order */
  float    s0, n0, snr;        /* peak averaged level, noise floor, s0 / n0 (linear power) */
} sdo_channel;
typedef struct {
  unsigned n, min_bins, last_total;
  float    alpha, gamma, snr, n0;
  int      primed, n0_primed;
  float   *avg, *tmp;
} sdo_chdet;
#define SDO_CHDET_MAXRAW 8192u
int      sdo_chdet_init(sdo_chdet *d, unsigned n, float alpha, float gamma, float snr, unsigned min_bins);
void     sdo_chdet_free(sdo_chdet *d);
unsigned sdo_chdet_feed(sdo_chdet *d, const float *psd, unsigned frames, sdo_channel *out, unsigned cap);

/* ---- U: inspector spectrum sources + baud estimators (spectsrc.c, SPEC.md section U) ---- */
enum sdo_spectsrc { SDO_SPECTSRC_NONE = 0, SDO_SPECTSRC_PSD, SDO_SPECTSRC_CYCLO, SDO_SPECTSRC_FMSPECT,
                    SDO_SPECTSRC_TIMEDIFF, SDO_SPECTSRC_ABSTIMEDIFF, SDO_SPECTSRC_EXP_2, SDO_SPECTSRC_EXP_4,
                    SDO_SPECTSRC_EXP_8, SDO_SPECTSRC_FAC, SDO_SPECTSRC_COUNT };
enum sdo_estimator { SDO_ESTIMATOR_BAUD_FAC = 0, SDO_ESTIMATOR_BAUD_NONLINEAR = 1, SDO_ESTIMATOR_COUNT };
#define SDO_U5_KMIN 8u
unsigned sdo_spectsrc_out_size(int kind, unsigned ns);
unsigned sdo_spectsrc_frame(int kind, unsigned ns, const sdo_cpx *c, size_t n_ch, float *out);
int      sdo_estimate_baud(int estimator, unsigned ns, float fs_ch, const sdo_cpx *c, size_t n_ch, float *baud);

/* ------------------------------------------------------------------------------------------------
 * Offline TimeWindow tasks that are fully specified in-repo (SPEC section Y; tasks.c).
 * ---------------------------------------------------------------------------------------------- */
enum sdo_space { SDO_SPACE_AMPLITUDE = 0, SDO_SPACE_PHASE = 1, SDO_SPACE_FREQUENCY = 2 }; /* SamplingProperties.h:27-31 */
void   sdo_delayed_conj(const sdo_cpx *x, sdo_cpx *y, size_t n, size_t delay);
size_t sdo_histogram_feed(const sdo_cpx *x, float *out, size_t n, int space);
size_t sdo_sample_manual(const sdo_cpx *x, size_t n, int space, size_t symbol_sync, double symbol_count,
                         sdo_cpx *out);
size_t sdo_sample_zero_crossing(const sdo_cpx *x, size_t n, int space, int amplitude, sdo_cpx threshold,
                                sdo_cpx zc_angle, float bnor, uint8_t *sym, size_t cap);
float  sdo_carrier_detect(const sdo_cpx *x, size_t n, double avg_rel_bw, double dc_notch_rel_bw);
/* SNR estimator on the decision-space histogram (Misc/SNREstimator.cpp:30-169; SPEC Y.7) */
typedef struct {
  float    sigma, alpha, hx, delta;
  unsigned bps, intervals, length;
  float   *gaussian, *hi, *htilde;
} sdo_snr_estimator;
void   sdo_snr_init(sdo_snr_estimator *e);
void   sdo_snr_free(sdo_snr_estimator *e);
void   sdo_snr_set_bps(sdo_snr_estimator *e, unsigned bps);
void   sdo_snr_feed(sdo_snr_estimator *e, const unsigned *history, unsigned n);
float  sdo_snr_get(const sdo_snr_estimator *e);

/* ------------------------------------------------------------------------------------------------
 * Analog-TV processor (SPEC section TV; tvproc.c).  Field names of sdo_tv_params are those of
 * sigutils_tv_processor_params as the reference fills them (Default/GenericInspector/TVProcessorTab.cpp:549-597).
 * ---------------------------------------------------------------------------------------------- */
#define SDO_TV_RING 4
#define SDO_TV_MAX_W 4096
#define SDO_TV_MAX_H 4096
typedef struct {
  int32_t  enable_sync, reverse, interlace, enable_agc;
  float    x_off;
  int32_t  dominance;
  uint32_t frame_lines;
  float    frame_spacing;
  int32_t  enable_comb, comb_reverse;
  float    hsync_len, vsync_len, line_len;
  uint32_t vsync_odd_trigger;
  float    t_tol, l_tol, g_tol;
  float    hsync_huge_err, hsync_max_err, hsync_min_err;
  float    hsync_len_tau, line_len_tau, agc_tau, hsync_fast_track_tau, hsync_slow_track_tau;
} sdo_tv_params;
typedef struct sdo_tv sdo_tv;
void     sdo_tv_params_pal(sdo_tv_params *p, float fs);
void     sdo_tv_params_ntsc(sdo_tv_params *p, float fs);
int      sdo_tv_params_valid(const sdo_tv_params *p);
sdo_tv  *sdo_tv_new(const sdo_tv_params *p);
void     sdo_tv_destroy(sdo_tv *t);
int      sdo_tv_set_params(sdo_tv *t, const sdo_tv_params *p);
void     sdo_tv_geometry(const sdo_tv *t, int *w, int *h);
int      sdo_tv_feed(sdo_tv *t, float x);
size_t   sdo_tv_feed_bulk(sdo_tv *t, const float *x, size_t n);
uint64_t sdo_tv_frames(const sdo_tv *t);
const float *sdo_tv_frame(const sdo_tv *t, uint64_t frame_no);
void     sdo_tv_estimates(const sdo_tv *t, float *line_len, float *hsync_len, float *gain);
void     sdo_tv_feed_transform(const sdo_cpx *x, size_t n, int mode, float k, float dc, float *out);

/* multi-threaded CPU baseline: S independent streams, each n samples, same params (OpenMP). */
double sdo_baseline_run(const sdo_an_params *p, const sdo_cpx *x, size_t n_streams, size_t n,
                        int n_threads, uint64_t *checksum);

#ifdef __cplusplus
}
#endif
#endif
