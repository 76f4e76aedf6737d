// ref_glue_suscan.cpp -- ORACLE SUPPORT (test infrastructure): C entry points around the reference's own C++ facade
// classes, COMPILED FROM /root/reference where they lie (oracle/Makefile target `ref` -> oracle/_ref/libsdref_suscan.so;
// nothing of the reference is copied):
//   Suscan/MQ.cpp, Suscan/Message.cpp, Suscan/Exception.cpp,
//   Suscan/Messages/{PSDMessage, SamplesMessage, StatusMessage, ChannelMessage, GenericMessage}.cpp
// built against THIS repository's include/analyzer/*.h + include/sigutils/*.h (the reference's <Suscan/...> headers come
// from /root/reference/include, Qt from the no-behaviour stubs of oracle/ref_shim/) and linked with libsuscan.so.
// What it shows: (1) SURVEY 8(b) -- these wrappers compile unmodified against the shim headers and run against the shim
// library (caller-owned suscan_mq, shared_ptr deleter -> suscan_analyzer_dispose_message, the struct fields they
// dereference); (2) SURVEY 8(a) row a17 -- the PSDMessage constructor's fft-shift + SU_POWER_DB pass
// (Suscan/Messages/PSDMessage.cpp:26-39) is the reference's own, and the oracle's / the kernels' dB epilogue is held to it.
// A separate library from libsdref.so: that one compiles Misc/Averager.cpp against a two-accessor stand-in of
// Suscan::PSDMessage, this one holds the real class.
#include <Suscan/MQ.h>
#include <Suscan/Message.h>
#include <Suscan/Messages/PSDMessage.h>
#include <Suscan/Messages/SamplesMessage.h>
#include <Suscan/Messages/StatusMessage.h>
#include <analyzer/analyzer.h>
#include <analyzer/mq.h>
#include <stdlib.h>
#include <string.h>

extern "C" {

// A PSD payload as libsuscan hands it over (heap struct + heap bins, linear power, DC first) through the reference's
// PSDMessage: out = what the GUI sees (halves swapped, dB).  The payload is released by the Message's deleter.
long ref_suscan_psd_message(const float *psd_lin, unsigned long n, double fc, unsigned samp_rate, float *out,
                            double *fc_out, unsigned *rate_out)
{
  struct suscan_analyzer_psd_msg *m = (struct suscan_analyzer_psd_msg *) calloc(1, sizeof(*m));
  if (!m) return -1;
  m->psd_size = n;
  m->psd_data = (SUFLOAT *) malloc(n * sizeof(SUFLOAT));
  memcpy(m->psd_data, psd_lin, n * sizeof(SUFLOAT));
  m->fc = (int64_t) fc; m->samp_rate = samp_rate; m->measured_samp_rate = (SUFLOAT) samp_rate;
  long size;
  {
    Suscan::PSDMessage msg(m);                   // PSDMessage.cpp:26-39: the swap + dB pass happens here
    Suscan::PSDMessage copy = msg;               // shared_ptr semantics of Suscan/Message.cpp:57-88
    size = (long) copy.size();
    memcpy(out, copy.get(), (size_t) size * sizeof(float));
    *fc_out = (double) copy.getFrequency(); *rate_out = copy.getSampleRate();
    if (copy.getType() != SUSCAN_ANALYZER_MESSAGE_TYPE_PSD) size = -2;
  }                                              // last owner gone -> suscan_analyzer_dispose_message(PSD, m)
  return size;
}

// caller-owned queue (Suscan/MQ.cpp:31-44) + a SAMPLES payload posted by the library side, read back by the wrapper
long ref_suscan_mq_samples(const SUCOMPLEX *samples, unsigned long n, unsigned inspector_id, SUCOMPLEX *out, unsigned *id_out)
{
  Suscan::MQ mq;
  // Suscan::Analyzer is a friend that passes &mq.mq to suscan_analyzer_new; the queue is the first member
  struct suscan_mq *raw = reinterpret_cast<struct suscan_mq *>(&mq);
  struct suscan_analyzer_sample_batch_msg *m = (struct suscan_analyzer_sample_batch_msg *) calloc(1, sizeof(*m));
  m->inspector_id = inspector_id; m->sample_count = n;
  m->samples = (SUCOMPLEX *) malloc(n * sizeof(SUCOMPLEX));
  memcpy(m->samples, samples, n * sizeof(SUCOMPLEX));
  if (!suscan_mq_write(raw, SUSCAN_ANALYZER_MESSAGE_TYPE_SAMPLES, m)) return -1;
  uint32_t type = 0;
  void *p = mq.read(type);
  if (type != SUSCAN_ANALYZER_MESSAGE_TYPE_SAMPLES || p != m) return -2;
  Suscan::SamplesMessage msg(static_cast<struct suscan_analyzer_sample_batch_msg *>(p));
  *id_out = msg.getInspectorId();
  memcpy(out, msg.getSamples(), msg.getCount() * sizeof(SUCOMPLEX));
  return (long) msg.getCount();
}

int ref_suscan_status_message(int code, const char *text, char *text_out, size_t cap)
{
  struct suscan_analyzer_status_msg *m = (struct suscan_analyzer_status_msg *) calloc(1, sizeof(*m));
  m->code = code; m->err_msg = text ? strdup(text) : nullptr;
  Suscan::StatusMessage msg(m);
  strncpy(text_out, msg.getMessage().toStdString().c_str(), cap - 1); text_out[cap - 1] = 0;
  return msg.getCode();
}

}  // extern "C"
