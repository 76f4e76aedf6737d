// ref_glue.cpp -- ORACLE SUPPORT (test infrastructure): C entry points around DSP code of the reference that is
// COMPILED FROM /root/reference where it lies (oracle/Makefile target `ref` -> oracle/_ref/libsdref.so; nothing of the
// reference is copied into this repository).  It pins the restatements of oracle/*.c to the reference itself for the
// arithmetic the reference does contain:
//   Panoramic/Scanner.cpp:36-293   SigDigger::SpectrumView (setRange, feed x3, feedLinearMode, feedHistogramMode,
//                                  interpolate, reset)            -> oracle/spectrumview.c
//   Tasks/QuadDemodTask.cpp        QuadDemodTask::work            -> sdo_quad_demod
//   Tasks/DelayedConjTask.cpp      DelayedConjTask::work          -> sdo_delayed_conj
//   Tasks/WaveSampler.cpp          sampleManual / sampleZeroCrossing / sampleGardner (the latter over this repo's
//                                  su_clock_detector shim)        -> sdo_sample_manual / _zero_crossing
//   Misc/Averager.cpp              Averager::feed                 -> sdo_averager_feed
//   Tasks/CostasRecoveryTask.cpp, PLLSyncTask.cpp, AGCTask.cpp, CarrierXlator.cpp, LPFTask.cpp   constructor + work()
//                                  loops over this repo's <sigutils/{pll,agc,ncqo,specttuner}.h> (the Tasks/ of the
//                                  north-star, drop-in)           -> sdo_costas / sdo_pll / sdo_agc / sdo_ncqo / specttuner
//   Tasks/HistogramFeeder.cpp      HistogramFeeder::work          -> sdo_histogram_feed
//   Tasks/CarrierDetector.cpp      CarrierDetector::work (Blackman-Harris taps of this repo's <sigutils/taps.h>, FFT
//                                  through the ref_shim/fftw3.h stand-in, notch, arg-max, circular centroid)
//                                                                  -> sdo_carrier_detect
//   Misc/SNREstimator.cpp          SNREstimator (setBps / setAlpha / setSigma / feed / getSigma / getSNR / getModel)
//                                                                  -> sdo_snr_*
//   Default/GenericInspector/TVProcessorWorker.cpp   the TV tab's worker (setParams / start / pushData / process /
//                                  work with its frame acknowledgement window) over this repo's <sigutils/tvproc.h>
//                                                                  -> the reference drives su_tv_processor_* unmodified
// The Qt base classes are the no-behaviour stubs of oracle/ref_shim/; what moc would generate (the signal bodies) and
// the CancellableTask plumbing are defined here.
#include <Scanner.h>
#include <QuadDemodTask.h>
#include <DelayedConjTask.h>
#include <WaveSampler.h>
#include <Averager.h>
#include <TVProcessorWorker.h>
#include <CostasRecoveryTask.h>
#include <PLLSyncTask.h>
#include <AGCTask.h>
#include <CarrierXlator.h>
#include <HistogramFeeder.h>
#include <LPFTask.h>
#include <SNREstimator.h>
#include <CarrierDetector.h>
#include <cmath>

// ---- Suscan::CancellableTask plumbing (Suscan/CancellableTask.cpp is Qt glue, not DSP)
Suscan::CancellableTask::CancellableTask(QObject *parent) : QObject(parent) { prog = 0; }
Suscan::CancellableTask::~CancellableTask(void) {}
void Suscan::CancellableTask::setProgress(qreal p) { prog = p; }
void Suscan::CancellableTask::setStatus(QString s) { status = s; }
void Suscan::CancellableTask::setDataSize(quint64 s) { dataSize = s; }
void Suscan::CancellableTask::done(void) {}
void Suscan::CancellableTask::cancelled(void) {}
void Suscan::CancellableTask::progress(qreal, QString) {}
void Suscan::CancellableTask::error(QString) {}

// ---- WaveSampler::data signal: collects what the task delivers
static thread_local std::vector<SUCOMPLEX> *g_ws_out = nullptr;
static thread_local std::vector<Symbol> *g_ws_sym = nullptr;
void SigDigger::WaveSampler::data(SigDigger::WaveSampleSet set)
{
  if (g_ws_out) g_ws_out->insert(g_ws_out->end(), set.block, set.block + set.len);
  if (g_ws_sym) g_ws_sym->insert(g_ws_sym->end(), set.symbols, set.symbols + set.len);
}

// ---- TVProcessorWorker: the signals moc would generate deliver to this collector, which answers as TVProcessorTab
// ---- does (onTVProcessorFrame, Default/GenericInspector/TVProcessorTab.cpp:657-663: acknowledgeFrame, then
// ---- tvProcessorDisposeFrame -> returnFrame)
struct RefTvSink { SigDigger::TVProcessorWorker *w; std::vector<std::vector<SUFLOAT>> frames; int width = 0, height = 0; bool failed = false; };
static thread_local RefTvSink *g_tv_sink = nullptr;
void SigDigger::TVProcessorWorker::frame(struct sigutils_tv_frame_buffer *f)
{
  RefTvSink *s = g_tv_sink;
  if (!s || !f) return;
  s->w->acknowledgeFrame();
  s->width = f->width; s->height = f->height;
  s->frames.emplace_back(f->buffer, f->buffer + (size_t) f->width * f->height);
  s->w->returnFrame(f);
}
void SigDigger::TVProcessorWorker::error(QString) { if (g_tv_sink) g_tv_sink->failed = true; }
void SigDigger::TVProcessorWorker::paramsChanged(sigutils_tv_processor_params) {}

// ---- HistogramFeeder::data signal
static thread_local std::vector<float> *g_hist_out = nullptr;
void SigDigger::HistogramFeeder::data(const float *d, unsigned int size)
{
  if (g_hist_out) g_hist_out->insert(g_hist_out->end(), d, d + size);
}

// ---- the FFTW stand-in of ref_shim/fftw3.h: radix-2, binary64 inside
struct ref_fftwf_plan_s { int n; fftwf_complex *in, *out; int sign; };
extern "C" void *fftwf_malloc(size_t n) { return malloc(n); }
extern "C" void fftwf_free(void *p) { free(p); }
extern "C" fftwf_plan fftwf_plan_dft_1d(int n, fftwf_complex *in, fftwf_complex *out, int sign, unsigned)
{
  if (n < 1 || (n & (n - 1))) return nullptr;
  ref_fftwf_plan_s *p = new ref_fftwf_plan_s{ n, in, out, sign };
  return p;
}
extern "C" void fftwf_destroy_plan(fftwf_plan p) { delete p; }
extern "C" void fftwf_execute(const fftwf_plan p)
{
  const int n = p->n;
  std::vector<double> re(n), im(n);
  for (int i = 0, j = 0; i < n; ++i) {            // bit reversal
    re[j] = p->in[i][0]; im[j] = p->in[i][1];
    int bit = n >> 1;
    for (; j & bit; bit >>= 1) j ^= bit;
    j ^= bit;
  }
  for (int len = 2; len <= n; len <<= 1) {
    const double ang = (p->sign < 0 ? -2.0 : 2.0) * M_PI / len;
    for (int i = 0; i < n; i += len)
      for (int k = 0; k < len / 2; ++k) {
        const double wr = cos(ang * k), wi = sin(ang * k);
        const double ur = re[i + k], ui = im[i + k];
        const double vr = re[i + k + len / 2] * wr - im[i + k + len / 2] * wi;
        const double vi = re[i + k + len / 2] * wi + im[i + k + len / 2] * wr;
        re[i + k] = ur + vr; im[i + k] = ui + vi;
        re[i + k + len / 2] = ur - vr; im[i + k + len / 2] = ui - vi;
      }
  }
  for (int i = 0; i < n; ++i) { p->out[i][0] = (float) re[i]; p->out[i][1] = (float) im[i]; }
}

extern "C" {

// ---- SpectrumView
void *ref_sview_new(void) { return new SigDigger::SpectrumView(); }
void ref_sview_free(void *v) { delete (SigDigger::SpectrumView *) v; }
void ref_sview_set_range(void *v, double fmin, double fmax, double fft_bandwidth, float rel_bw)
{
  SigDigger::SpectrumView *s = (SigDigger::SpectrumView *) v;
  s->setRange(fmin, fmax);
  s->fftBandwidth = fft_bandwidth; s->fftRelBw = rel_bw;
}
void ref_sview_feed(void *v, const float *psd, const float *count, unsigned long size, double center, int adjust_sides)
{ ((SigDigger::SpectrumView *) v)->feed(psd, count, size, center, adjust_sides != 0); }
void ref_sview_feed_range(void *v, const float *psd, const float *count, unsigned long size, double fmin, double fmax, int adjust)
{ ((SigDigger::SpectrumView *) v)->feed(psd, count, size, fmin, fmax, adjust != 0); }
void ref_sview_feed_view(void *v, const void *detail) { ((SigDigger::SpectrumView *) v)->feed(*(const SigDigger::SpectrumView *) detail); }
void ref_sview_interpolate(void *v) { ((SigDigger::SpectrumView *) v)->interpolate(); }
void ref_sview_reset(void *v) { ((SigDigger::SpectrumView *) v)->reset(); }
unsigned ref_sview_read(const void *v, float *psd, float *accum, float *count)
{
  const SigDigger::SpectrumView *s = (const SigDigger::SpectrumView *) v;
  const unsigned n = s->spectrumSize;
  memcpy(psd, s->psd, n * sizeof(float)); memcpy(accum, s->psdAccum, n * sizeof(float)); memcpy(count, s->psdCount, n * sizeof(float));
  return n;
}

// ---- Tasks
void ref_quad_demod(const SUCOMPLEX *data, SUCOMPLEX *dst, size_t n)
{
  QuadDemodTask t(data, dst, n);
  while (t.work()) ;
}
void ref_delayed_conj(const SUCOMPLEX *data, SUCOMPLEX *dst, size_t n, unsigned long delay)
{
  DelayedConjTask t(data, dst, n, delay);
  while (t.work()) ;
}
// sync: 0 MANUAL, 1 GARDNER, 2 ZERO_CROSSING; space: 0 AMPLITUDE, 1 PHASE, 2 FREQUENCY.  Returns the number of soft
// samples written to out (cap respected).
long ref_wave_sampler(const SUCOMPLEX *data, size_t n, int sync, int space, double fs, double rate, double loop_gain,
                      int amplitude, float thr_re, float thr_im, float zc_re, float zc_im, size_t symbol_sync,
                      double symbol_count, SUCOMPLEX *out, unsigned char *sym_out, size_t cap)
{
  SigDigger::SamplingProperties p;
  p.sync = (SigDigger::SamplingClockSync) sync; p.space = (SigDigger::SamplingSpace) space;
  p.fs = fs; p.loopGain = loop_gain; p.amplitude = amplitude != 0;
  p.threshold = SUCOMPLEX(thr_re, thr_im); p.zeroCrossingAngle = SUCOMPLEX(zc_re, zc_im);
  p.data = data; p.length = n; p.symbolSync = symbol_sync; p.symbolCount = symbol_count; p.rate = rate;
  Decider d;
  std::vector<SUCOMPLEX> got; std::vector<Symbol> syms;
  g_ws_out = &got; g_ws_sym = &syms;
  {
    SigDigger::WaveSampler ws(p, &d);
    while (ws.work()) ;
  }
  g_ws_out = nullptr; g_ws_sym = nullptr;
  const size_t m = got.size() < cap ? got.size() : cap;
  if (out) memcpy(out, got.data(), m * sizeof(SUCOMPLEX));
  if (sym_out) memcpy(sym_out, syms.data(), m);
  return (long) got.size();
}

// ---- Averager: frames [n_frames][size] in order, result = averager state after the last
void ref_averager(const float *frames, unsigned n_frames, unsigned size, float alpha, float *out)
{
  SigDigger::Averager a;
  a.setAlpha(alpha);
  for (unsigned f = 0; f < n_frames; ++f) a.feed(Suscan::PSDMessage(frames + (size_t) f * size, size));
  memcpy(out, a.get(), size * sizeof(float));
}


// same contract as tu_tv_worker of tests/shim/reference_tu.cpp: one pushData + process per block
long ref_tv_worker(const struct sigutils_tv_processor_params *params, const float *x, size_t n, size_t block, float *out,
                   size_t cap, int *width, int *height)
{
  SigDigger::TVProcessorWorker w;
  RefTvSink sink; sink.w = &w;
  g_tv_sink = &sink;
  w.setParams(*params);
  w.start();
  long result = -1;
  if (!sink.failed) {
    for (size_t p = 0; p < n; p += block) {
      std::vector<SUFLOAT> buf(x + p, x + p + (block < n - p ? block : n - p));
      w.pushData(buf);
      w.process();
    }
    *width = sink.width; *height = sink.height;
    const size_t px = (size_t) sink.width * sink.height;
    for (size_t f = 0; f < sink.frames.size() && f < cap; ++f) memcpy(out + f * px, sink.frames[f].data(), px * sizeof(float));
    result = (long) sink.frames.size();
  }
  w.stop();
  g_tv_sink = nullptr;
  return result;
}


// ---- the Tasks/ of the north-star, from the reference: constructor + work() until it returns false
int ref_task_costas(const SUCOMPLEX *data, SUCOMPLEX *dst, size_t n, float tau, float loopbw, int kind)
{
  try { CostasRecoveryTask t(data, dst, n, tau, loopbw, (enum sigutils_costas_kind) kind); while (t.work()) ; return 0; }
  catch (...) { return -1; }
}
int ref_task_pll(const SUCOMPLEX *data, SUCOMPLEX *dst, size_t n, float cutoff)
{
  try { PLLSyncTask t(data, dst, n, cutoff); while (t.work()) ; return 0; } catch (...) { return -1; }
}
int ref_task_agc(const SUCOMPLEX *data, SUCOMPLEX *dst, size_t n, float tau)
{
  try { AGCTask t(data, dst, n, tau); while (t.work()) ; return 0; } catch (...) { return -1; }
}
int ref_task_xlate(const SUCOMPLEX *data, SUCOMPLEX *dst, size_t n, float relFreq, float phase)
{
  try { SigDigger::CarrierXlator t(data, dst, n, relFreq, phase); while (t.work()) ; return 0; } catch (...) { return -1; }
}
int ref_task_lpf(const SUCOMPLEX *data, SUCOMPLEX *dst, size_t n, float bw)        // specttuner: needs the GPU
{
  try { LPFTask t(data, dst, n, bw); while (t.work()) ; return 0; } catch (...) { return -1; }
}
long ref_task_histogram(const SUCOMPLEX *data, size_t n, int space, float *out, size_t cap)
{
  SigDigger::SamplingProperties props;
  memset(&props, 0, sizeof(props));
  props.space = (SigDigger::SamplingSpace) space; props.data = data; props.length = n;
  std::vector<float> got;
  g_hist_out = &got;
  { SigDigger::HistogramFeeder t(props); while (t.work()) ; }
  g_hist_out = nullptr;
  for (size_t i = 0; i < got.size() && i < cap; ++i) out[i] = got[i];
  return (long) got.size();
}


// ---- Misc/SNREstimator.cpp: `feeds` successive histories of `length` bins each
int ref_snr_estimator(unsigned bps, float alpha, float sigma0, const unsigned *histories, unsigned length, unsigned feeds,
                      float *sigma_out, float *snr_out, float *model_out)
{
  SigDigger::SNREstimator e;
  e.setBps(bps);
  e.setAlpha(alpha);
  if (sigma0 > 0) e.setSigma(sigma0);
  for (unsigned f = 0; f < feeds; ++f) {
    std::vector<unsigned int> h(histories + (size_t) f * length, histories + (size_t) (f + 1) * length);
    e.feed(h);
    sigma_out[f] = e.getSigma(); snr_out[f] = e.getSNR();
  }
  const std::vector<float> &m = e.getModel();
  for (size_t i = 0; i < m.size() && i < length; ++i) model_out[i] = m[i];
  return (int) m.size();
}


// ---- Tasks/CarrierDetector.cpp: the four work() states, then the peak (rad / sample)
float ref_carrier_detect(const SUCOMPLEX *data, size_t n, double avgRelBw, double dcNotchRelBw)
{
  SigDigger::CarrierDetector d(data, n, avgRelBw, dcNotchRelBw);
  while (d.work()) ;
  return d.getPeak();
}

}  // extern "C"
