// analyzer.cu -- suscan-style asynchronous analyzer on top of the engine: a worker thread pulls IQ from a
// source, runs the engine block by block and posts heap-allocated messages on a FIFO that the caller
// drains with sdb_analyzer_read(); requests (open / set id / set config / close / set params) are queued
// and answered in order with messages carrying the caller's req_id.
//
// This is SURVEY.md section 8(a) row a18 + the command half of section 8(b), i.e. what the reference's
// Suscan::Analyzer wrapper drives:
//   suscan_analyzer_new / read / dispose_message / req_halt / destroy   Suscan/Analyzer.cpp:608, :111-115, :63-103, :321, :625-638
//   suscan_analyzer_open_ex_async                                      Suscan/Analyzer.cpp:459-484
//   suscan_analyzer_set_inspector_id_async / _config_async / close     Suscan/Analyzer.cpp:486-537
//   suscan_analyzer_set_params_async                                   Suscan/Analyzer.cpp:219-227
//   open handshake OPEN -> SET_ID                                      Suscan/AnalyzerRequestTracker.cpp:138-157
//   message payload fields                                             Suscan/Messages/PSDMessage.cpp:30-112,
//                                                                      include/Suscan/Messages/SamplesMessage.h:33-59,
//                                                                      Suscan/Messages/InspectorMessage.cpp:28-71
// Differences stated openly: the channel plan is rebuilt at a block boundary whenever an inspector is
// opened, closed or reconfigured, which restarts the loops of the inspectors that stay open (suscan keeps
// them running); one source per analyzer; no estimators / spectrum sources (section 8(f)).
#include "../../include/sigdigger_b200.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>
#include <sys/time.h>

namespace {

struct Cmd {
  enum Kind { OPEN, SET_ID, SET_CONFIG, CLOSE, SET_PARAMS, SET_SPECTRUM, ESTIMATOR_CMD, SET_IQ_REVERSE, SET_THROTTLE, SET_WATERMARK,
              SET_FREQ, SET_BANDWIDTH, SEEK, SET_HOP_RANGE, SET_REL_BW, SET_BUFFERING, SET_STRATEGY,
              SET_PARTITIONING, SET_HISTORY, REPLAY, SET_DC_REMOVE } kind;
  double value2 = 0;
  uint32_t req_id = 0;
  uint32_t aux_id = 0; int enabled = 0;       // spectrum source id / estimator id + enable flag
  double value = 0;                           // throttle rate
  int32_t handle = -1;
  int32_t parent = -1;                        // open_ex: handle of the inspector whose channel is the input
  uint32_t inspector_id = 0;
  std::string cls;
  sdb_sigutils_channel channel{};
  int precise = 0;
  sdb_inspector_config cfg{};
  sdb_analyzer_params params{};
};

struct Insp {
  int32_t handle;
  bool open;
  bool has_id;
  uint32_t inspector_id;
  int cls;
  sdb_sigutils_channel channel;
  int precise;
  sdb_inspector_config cfg;
  int engine_handle;      // handle inside the current engine, -1 if not mapped
  float fs, bandwidth, lo;
  int32_t parent;         // -1 = fed by the baseband; else sub-carrier inspector on that inspector's channel
  uint32_t size;          // channel IFFT size (samples per two hops) in the engine that feeds it
  uint32_t spectsrc_id;   // 0 = none (index 0 of the GUI's combo)
  uint32_t est_mask;
  uint32_t spect_size;    // frame size chosen at rebuild time (0 = too few channel samples per block)
};

int class_id(const std::string &c)
{
  if (c == "psk") return SDB_INSP_PSK;
  if (c == "fsk") return SDB_INSP_FSK;
  if (c == "ask") return SDB_INSP_ASK;
  if (c == "audio") return SDB_INSP_AUDIO;
  if (c == "raw") return SDB_INSP_RAW;
  return -1;
}

char *dupstr(const char *s)
{
  size_t n = strlen(s) + 1;
  char *p = (char *) malloc(n);
  memcpy(p, s, n);
  return p;
}

}  // namespace

struct sdb_analyzer {
  sdb_analyzer_params params;
  sdb_source_config src;
  size_t src_pos = 0;
  // output queue (suscan_mq): MPSC FIFO of (type, payload)
  std::mutex mq_m;
  std::condition_variable mq_cv;
  std::deque<std::pair<uint32_t, void *>> mq;
  // command queue
  std::mutex cmd_m;
  std::deque<Cmd> cmds;
  bool halt_req = false;
  std::thread worker;
  std::vector<Insp> insps;
  sdb_engine_t *eng = nullptr;
  // Sub-carrier inspection (GenericInspector.cpp:502-525: openInspectorTab(..., parent = this handle)): every
  // inspector with open children -- at any depth: a sub-carrier inspector can be a parent itself -- gets an engine
  // whose input is its channel stream and whose channeliser window equals its channel size, so that one parent hop
  // is exactly one child hop.  Children have inspector chains, spectrum sources and estimators like any other.
  struct Sub { int32_t parent; sdb_engine_t *eng; };
  std::vector<Sub> subs;
  void drop_subs() { for (auto &s : subs) if (s.eng) sdb_engine_destroy(s.eng); subs.clear(); }
  // the engine that runs inspector i's chain: the analyzer's own for a baseband inspector, its parent's
  // sub-carrier engine otherwise
  template <class I> sdb_engine_t *engine_of(const I &i) const
  {
    if (i.parent < 0) return eng;
    for (auto &sb : subs) if (sb.parent == i.parent) return sb.eng;
    return nullptr;
  }
  bool plan_dirty = true;
  size_t block = 0;
  std::atomic<double> measured_rate{ 0.0 };   // written by the worker, read by sdb_analyzer_get_measured_samp_rate
  uint64_t total_samples = 0;
  std::atomic<uint64_t> total_samples_pub{ 0 };   // copy for readers on other threads
  std::mutex insp_m;                              // guards insps against readers on other threads
  double psd_credit = 0;
  // source-side options of the worker loop (Suscan/Analyzer.cpp:117-135, 229-244; SourceWidget.cpp:1156-1184)
  bool iq_reverse = false, dc_remove = false;
  double throttle = 0;                                  // samples / s, 0 = as fast as the source delivers
  std::chrono::steady_clock::time_point throttle_t0; uint64_t throttle_s0 = 0;
  // history ring + replay (suscan_analyzer_set_history_size / _replay, Suscan/Analyzer.cpp:157-167;
  // Default/Source/SourceWidget.cpp:1070-1073, 1206, 1508)
  std::vector<sdb_complex> hist; size_t hist_start = 0, hist_fill = 0, replay_pos = 0;
  bool replaying = false, replay_wrapped = false; int looped = 0;
  struct BbFilter { sdb_baseband_filter_fn fn; void *priv; int64_t prio; };
  std::vector<BbFilter> bb_filters;

  void post(uint32_t type, void *payload)
  {
    std::lock_guard<std::mutex> l(mq_m);
    mq.emplace_back(type, payload);
    mq_cv.notify_one();
  }
  void post_status(uint32_t type, int code, const char *msg)
  {
    sdb_analyzer_status_msg *m = (sdb_analyzer_status_msg *) calloc(1, sizeof(*m));
    m->code = code;
    m->err_msg = msg ? dupstr(msg) : nullptr;
    post(type, m);
  }
  void post_inspector(int kind, const Cmd &c, const Insp *i)
  {
    sdb_analyzer_inspector_msg *m = (sdb_analyzer_inspector_msg *) calloc(1, sizeof(*m));
    m->kind = kind; m->req_id = c.req_id; m->handle = i ? i->handle : c.handle;
    m->inspector_id = i ? i->inspector_id : c.inspector_id;
    m->class_name = dupstr(i ? (const char *[]){ "psk", "fsk", "ask", "audio", "raw" }[i->cls] : c.cls.c_str());
    if (i) { m->channel = i->channel; m->config = i->cfg; m->fs = (float) src.samp_rate; m->equiv_fs = i->fs;
             m->bandwidth = i->bandwidth; m->lo = i->lo; m->spectsrc_id = i->spectsrc_id; }
    else m->channel = c.channel;
    m->spectsrc_count = SDB_SPECTSRC_COUNT - 1; m->estimator_count = SDB_ESTIMATOR_COUNT;
    if (kind == SDB_INSPECTOR_MSGKIND_ESTIMATOR) { m->estimator_id = c.aux_id; m->enabled = c.enabled; }
    post(SDB_ANALYZER_MESSAGE_TYPE_INSPECTOR, m);
  }
  double chan_credit = 0;

  // SAMPLES batches: suscan flushes an inspector's sample buffer once it holds `watermark` samples
  // (suscan_analyzer_set_inspector_watermark_async, Suscan/Analyzer.cpp:527-537; the audio path asks for half a
  // playback buffer, Default/Audio/AudioProcessor.cpp:745-747).  watermark 0 = one batch per block.
  struct Pending { std::vector<sdb_complex> soft; std::vector<uint8_t> hard; uint64_t watermark = 0; };
  std::map<int32_t, Pending> pending;
  void deliver(const Insp &i, const sdb_complex *s, const uint8_t *h, size_t n, bool flush)
  {
    Pending &p = pending[i.handle];
    p.soft.insert(p.soft.end(), s, s + n);
    p.hard.insert(p.hard.end(), h, h + n);
    if (p.soft.empty() || (!flush && p.soft.size() < p.watermark)) return;
    sdb_analyzer_sample_batch_msg *m = (sdb_analyzer_sample_batch_msg *) calloc(1, sizeof(*m));
    m->inspector_id = i.inspector_id; m->sample_count = (uint64_t) p.soft.size();
    m->samples = (sdb_complex *) malloc(p.soft.size() * sizeof(sdb_complex));
    m->symbols = (uint8_t *) malloc(p.hard.size());
    memcpy(m->samples, p.soft.data(), p.soft.size() * sizeof(sdb_complex));
    memcpy(m->symbols, p.hard.data(), p.hard.size());
    p.soft.clear(); p.hard.clear();
    post(SDB_ANALYZER_MESSAGE_TYPE_SAMPLES, m);
  }

  // (re)build the engine from the open inspectors at a block boundary.  The new engine takes over the running
  // state of the old one (sdb_engine_migrate_map): input history, DC estimate, and for every inspector that stays
  // open its cross-fade tail, LO phase and loop state -- opening, closing, retuning or reconfiguring ONE inspector
  // leaves the sample streams of the others untouched (suscan keeps them running too, Suscan/Analyzer.cpp:459-537).
  // Returns false on failure (status message posted).
  bool rebuild()
  {
    sdb_engine_t *old = eng; eng = nullptr;
    std::vector<Sub> old_subs; old_subs.swap(subs);
    std::vector<int> old_handle(insps.size(), -1);
    for (size_t q = 0; q < insps.size(); ++q) old_handle[q] = insps[q].engine_handle;
    auto drop_old = [&] { if (old) sdb_engine_destroy(old); for (auto &s : old_subs) if (s.eng) sdb_engine_destroy(s.eng); };
    sdb_engine_params ep;
    memset(&ep, 0, sizeof(ep));
    ep.n_streams = 1;
    ep.psd_size = (uint32_t) params.detector_params.window_size;
    ep.psd_window = params.detector_params.window;
    ep.st_window_size = 0;
    ep.max_feed = (uint32_t) block;
    ep.device = src.device;
    ep.flags = (iq_reverse ? SDB_FLAG_IQ_REVERSE : 0) | (dc_remove ? SDB_FLAG_DC_REMOVE : 0);
    ep.input_format = src.read ? SDB_FORMAT_FLOAT32 : src.input_format;
    eng = sdb_engine_new(&ep, src.samp_rate);
    if (!eng) { drop_old(); post_status(SDB_ANALYZER_MESSAGE_TYPE_SOURCE_INIT, SDB_ANALYZER_INIT_FAILURE, sdb_last_error()); return false; }
    std::vector<int32_t> old_of_new;
    for (auto &i : insps) {
      i.engine_handle = -1;
      if (!i.open || i.parent >= 0) continue;
      sdb_channel_params cp;
      double f = fmod(i.channel.fc / src.samp_rate + 1.0, 1.0);
      cp.f0 = (float) (2.0 * 3.14159265358979323846 * f);
      cp.bw = (float) (2.0 * 3.14159265358979323846 * (i.channel.f_hi - i.channel.f_lo) / src.samp_rate);
      cp.guard = 1.0f; cp.precise = i.precise;
      sdb_channel_info info;
      int h = sdb_engine_open_channel(eng, &cp, &info);
      if (h < 0) continue;
      i.engine_handle = h;
      old_of_new.push_back(old_handle[(size_t) i.handle]);
      i.size = info.size;
      i.fs = (float) (src.samp_rate / info.decimation);
      i.bandwidth = (float) (i.channel.f_hi - i.channel.f_lo);
      i.lo = (float) i.channel.fc;
      i.cfg.insp_class = i.cls;
      sdb_engine_set_inspector(eng, h, &i.cfg);
      // spectrum source / estimators: the largest frame (<= 4096) a steady-state block can fill (SPEC U.1)
      i.spect_size = 0;
      if (i.spectsrc_id || i.est_mask) {
        const size_t n_ch = block / (ep.psd_size / 2) * (info.size / 2);
        uint32_t ns = 4096;
        while (ns >= 64 && (size_t) ns + 1 > n_ch) ns >>= 1;
        if (ns >= 64) {
          i.spect_size = ns;
          sdb_engine_set_spectrum_source(eng, h, (int) i.spectsrc_id, ns);
          for (int e = 0; e < SDB_ESTIMATOR_COUNT; ++e)
            if (i.est_mask & (1u << e)) sdb_engine_set_estimator(eng, h, e, 1);
        }
      }
    }
    if (params.channel_update_int > 0)
      sdb_engine_set_channel_detector(eng, params.detector_params.alpha, params.detector_params.beta,
                                      params.detector_params.gamma,
                                      params.detector_params.snr > 0 ? params.detector_params.snr : 4.0f, 2);
    if (sdb_engine_commit(eng)) {
      drop_old();
      post_status(SDB_ANALYZER_MESSAGE_TYPE_SOURCE_INIT, SDB_ANALYZER_INIT_FAILURE, sdb_last_error());
      return false;
    }
    if (old && sdb_engine_same_geometry(eng, old))
      sdb_engine_migrate_map(eng, old, old_of_new.data(), old_of_new.size());
    // sub-carrier engines: one per inspector that has open children, parents before children (an inspector's handle
    // is larger than its parent's, so one pass in handle order visits every parent after it was planned itself).
    // Every level runs the same number of hops per block: a child's window is its parent's channel size.
    const size_t hops_per_block = block / (ep.psd_size / 2);
    for (size_t pi = 0; pi < insps.size(); ++pi) {
      Insp &p = insps[pi];
      if (!p.open || p.engine_handle < 0) continue;
      bool any = false;
      for (auto &c : insps) any = any || (c.open && c.parent == p.handle);
      if (!any || p.size < 16) continue;
      sdb_engine_params sp;
      memset(&sp, 0, sizeof(sp));
      sp.n_streams = 1; sp.psd_size = 0; sp.st_window_size = p.size;
      sp.max_feed = (uint32_t) (hops_per_block * (p.size / 2));
      sp.device = src.device; sp.input_format = SDB_FORMAT_FLOAT32;
      sdb_engine_t *se = sdb_engine_new(&sp, (double) p.fs);
      if (!se) continue;
      std::vector<int32_t> sub_old_of_new;
      for (auto &c : insps) {
        if (!c.open || c.parent != p.handle) continue;
        sdb_channel_params cp;
        cp.f0 = (float) (2.0 * 3.14159265358979323846 * fmod(c.channel.fc / (double) p.fs + 1.0, 1.0));
        cp.bw = (float) (2.0 * 3.14159265358979323846 * (c.channel.f_hi - c.channel.f_lo) / (double) p.fs);
        cp.guard = 1.0f; cp.precise = c.precise;
        sdb_channel_info info;
        int h = sdb_engine_open_channel(se, &cp, &info);
        if (h < 0) continue;
        c.engine_handle = h; c.size = info.size;
        sub_old_of_new.push_back(old_handle[(size_t) c.handle]);
        c.fs = (float) ((double) p.fs / info.decimation);
        c.bandwidth = (float) (c.channel.f_hi - c.channel.f_lo); c.lo = (float) c.channel.fc;
        c.cfg.insp_class = c.cls;
        sdb_engine_set_inspector(se, h, &c.cfg);
        // spectrum source / estimators of a sub-carrier inspector: as for a baseband one (SPEC U.1)
        c.spect_size = 0;
        if (c.spectsrc_id || c.est_mask) {
          const size_t n_ch = hops_per_block * (info.size / 2);
          uint32_t ns = 4096;
          while (ns >= 64 && (size_t) ns + 1 > n_ch) ns >>= 1;
          if (ns >= 64) {
            c.spect_size = ns;
            sdb_engine_set_spectrum_source(se, h, (int) c.spectsrc_id, ns);
            for (int e = 0; e < SDB_ESTIMATOR_COUNT; ++e)
              if (c.est_mask & (1u << e)) sdb_engine_set_estimator(se, h, e, 1);
          }
        }
      }
      if (sdb_engine_commit(se)) {
        sdb_engine_destroy(se);
        for (auto &c : insps) if (c.open && c.parent == p.handle) c.engine_handle = -1;
        continue;
      }
      for (auto &os : old_subs)
        if (os.parent == p.handle && os.eng && sdb_engine_same_geometry(se, os.eng))
          sdb_engine_migrate_map(se, os.eng, sub_old_of_new.data(), sub_old_of_new.size());
      subs.push_back(Sub{ p.handle, se });
    }
    drop_old();
    plan_dirty = false;
    return true;
  }

  void handle_cmds()
  {
    std::deque<Cmd> todo;
    { std::lock_guard<std::mutex> l(cmd_m); todo.swap(cmds); }
    std::lock_guard<std::mutex> il(insp_m);
    for (auto &c : todo) {
      switch (c.kind) {
        case Cmd::OPEN: {
          int cls = class_id(c.cls);
          double bw = c.channel.f_hi - c.channel.f_lo;
          if (cls < 0) { post_inspector(SDB_INSPECTOR_MSGKIND_WRONG_KIND, c, nullptr); break; }
          // the input of the new inspector: the baseband, or (sub-carrier inspection) the channel of `parent`
          double in_rate = src.samp_rate;
          uint32_t in_window = (uint32_t) params.detector_params.window_size;
          if (c.parent >= 0) {
            if (c.parent >= (int32_t) insps.size() || !insps[c.parent].open || insps[c.parent].size < 16) {
              Cmd w = c; w.handle = c.parent;
              post_inspector(SDB_INSPECTOR_MSGKIND_WRONG_HANDLE, w, nullptr); break;
            }
            in_rate = (double) insps[c.parent].fs; in_window = insps[c.parent].size;
          }
          if (!(bw > 0) || bw > in_rate || fabs(c.channel.fc) > in_rate / 2) {
            post_inspector(SDB_INSPECTOR_MSGKIND_INVALID_CHANNEL, c, nullptr); break;
          }
          Insp i;
          memset(&i, 0, sizeof(i));
          i.handle = (int32_t) insps.size(); i.open = true; i.has_id = false; i.cls = cls; i.channel = c.channel;
          i.precise = c.precise; i.engine_handle = -1; i.parent = c.parent;
          // geometry -> equivalent rate, needed for the default config the OPEN reply carries
          {
            sdb_engine_params ep; memset(&ep, 0, sizeof(ep));
            ep.n_streams = 1; ep.psd_size = c.parent >= 0 ? 0 : in_window; ep.st_window_size = in_window;
            ep.max_feed = c.parent >= 0 ? in_window : (uint32_t) block;
            ep.device = src.device;
            sdb_engine_t *probe = sdb_engine_new(&ep, in_rate);
            sdb_channel_params cp;
            cp.f0 = (float) (2.0 * 3.14159265358979323846 * fmod(c.channel.fc / in_rate + 1.0, 1.0));
            cp.bw = (float) (2.0 * 3.14159265358979323846 * bw / in_rate); cp.guard = 1.0f; cp.precise = c.precise;
            sdb_channel_info info; memset(&info, 0, sizeof(info));
            int h = probe ? sdb_engine_open_channel(probe, &cp, &info) : -1;
            if (probe) sdb_engine_destroy(probe);
            if (h < 0) { post_inspector(SDB_INSPECTOR_MSGKIND_INVALID_CHANNEL, c, nullptr); break; }
            i.fs = (float) (in_rate / info.decimation);
            i.size = info.size;
          }
          i.bandwidth = (float) bw; i.lo = (float) c.channel.fc;
          sdb_inspector_config_default(&i.cfg, cls, i.fs);
          i.cfg.clock_running = 0;       // "by default ... no samples are being delivered" (manual p.62)
          insps.push_back(i);
          plan_dirty = true;
          post_inspector(SDB_INSPECTOR_MSGKIND_OPEN, c, &insps.back());
          break;
        }
        case Cmd::SET_ID:
        case Cmd::SET_CONFIG:
        case Cmd::SET_SPECTRUM:
        case Cmd::ESTIMATOR_CMD:
        case Cmd::CLOSE: {
          if (c.handle < 0 || c.handle >= (int32_t) insps.size() || !insps[c.handle].open) {
            post_inspector(SDB_INSPECTOR_MSGKIND_WRONG_HANDLE, c, nullptr); break;
          }
          Insp &i = insps[c.handle];
          if (c.kind == Cmd::SET_SPECTRUM) {
            if (c.aux_id >= (uint32_t) SDB_SPECTSRC_COUNT) { post_inspector(SDB_INSPECTOR_MSGKIND_WRONG_OBJECT, c, nullptr); break; }
            i.spectsrc_id = c.aux_id; plan_dirty = true;
            post_inspector(SDB_INSPECTOR_MSGKIND_SPECTRUM, c, &i);     // acknowledgement: no spectrum_data yet
          } else if (c.kind == Cmd::ESTIMATOR_CMD) {
            if (c.aux_id >= (uint32_t) SDB_ESTIMATOR_COUNT) { post_inspector(SDB_INSPECTOR_MSGKIND_WRONG_OBJECT, c, nullptr); break; }
            if (c.enabled) i.est_mask |= 1u << c.aux_id; else i.est_mask &= ~(1u << c.aux_id);
            plan_dirty = true;
            post_inspector(SDB_INSPECTOR_MSGKIND_ESTIMATOR, c, &i);    // acknowledgement: value follows per block
          } else if (c.kind == Cmd::SET_ID) {
            i.inspector_id = c.inspector_id; i.has_id = true;
            post_inspector(SDB_INSPECTOR_MSGKIND_SET_ID, c, &i);
          } else if (c.kind == Cmd::SET_CONFIG) {
            if (c.cfg.insp_class != i.cls) { post_inspector(SDB_INSPECTOR_MSGKIND_WRONG_KIND, c, nullptr); break; }
            i.cfg = c.cfg; plan_dirty = true;
            post_inspector(SDB_INSPECTOR_MSGKIND_SET_CONFIG, c, &i);
          } else {
            i.open = false; plan_dirty = true;
            // sub-carrier inspectors go with it, to any depth (handles grow with the order of opening: one pass)
            for (auto &ch : insps) if (ch.open && ch.parent >= 0 && !insps[ch.parent].open) ch.open = false;
            post_inspector(SDB_INSPECTOR_MSGKIND_CLOSE, c, &i);
          }
          break;
        }
        case Cmd::SET_WATERMARK:
          if (c.handle < 0 || c.handle >= (int32_t) insps.size() || !insps[c.handle].open) {
            post_inspector(SDB_INSPECTOR_MSGKIND_WRONG_HANDLE, c, nullptr); break;
          }
          pending[c.handle].watermark = (uint64_t) c.value;
          break;
        case Cmd::SET_HISTORY:
          hist.assign((size_t) c.value, sdb_complex{ 0.0f, 0.0f });
          hist_start = hist_fill = replay_pos = 0; replaying = false; looped = 0;
          break;
        case Cmd::REPLAY:
          replaying = c.enabled != 0 && hist_fill > 0;
          replay_pos = 0; looped = 0; replay_wrapped = false;
          break;
        case Cmd::SET_HOP_RANGE:
          if (c.value <= c.value2) { hop_min = c.value; hop_max = c.value2; hop_index = 0; }
          break;
        case Cmd::SET_REL_BW:
          if (c.value > 0) rel_bw = (float) (c.value > 1 ? 1 : c.value);
          break;
        case Cmd::SET_BUFFERING: buffering = (uint64_t) c.value; break;
        case Cmd::SET_STRATEGY: strategy = c.enabled; hop_index = 0; break;
        case Cmd::SET_PARTITIONING: partitioning = c.enabled; break;
        case Cmd::SEEK:       // suscan_analyzer_seek (Suscan/Analyzer.cpp:151-155): in-memory captures only
          if (!src.read && src.data && c.value >= 0) {
            const double p = c.value * src.samp_rate;
            src_pos = p >= (double) src.length ? src.length : (size_t) p;
          }
          break;
        case Cmd::SET_FREQ:
        case Cmd::SET_BANDWIDTH: {
          // "overridable" setters (Suscan/Analyzer.cpp:509-526): no reply; the latest value wins.  The channel is
          // re-planned at the next block boundary (see the a18 limits in DESIGN.md).
          if (c.handle < 0 || c.handle >= (int32_t) insps.size() || !insps[c.handle].open) break;
          Insp &i = insps[c.handle];
          const double rate = i.parent >= 0 ? (double) insps[i.parent].fs : src.samp_rate;
          if (c.kind == Cmd::SET_FREQ) {
            if (fabs(c.value) > rate / 2) break;
            i.channel.fc = c.value; i.lo = (float) c.value;
          } else {
            if (!(c.value > 0) || c.value > rate) break;
            i.channel.f_lo = -0.5 * c.value; i.channel.f_hi = 0.5 * c.value; i.channel.bw = (float) c.value;
            i.bandwidth = (float) c.value;
          }
          plan_dirty = true;
          break;
        }
        case Cmd::SET_IQ_REVERSE:
          if (iq_reverse != (c.enabled != 0)) { iq_reverse = c.enabled != 0; plan_dirty = true; }
          break;
        case Cmd::SET_DC_REMOVE:
          if (dc_remove != (c.enabled != 0)) { dc_remove = c.enabled != 0; plan_dirty = true; }
          break;
        case Cmd::SET_THROTTLE:
          throttle = c.value > 0 ? c.value : 0;
          throttle_t0 = std::chrono::steady_clock::now(); throttle_s0 = total_samples;   // pace from here on
          break;
        case Cmd::SET_PARAMS: {
          params = c.params; plan_dirty = true;
          hop_min = params.min_freq; hop_max = params.max_freq; hop_index = 0;
          sdb_analyzer_params *m = (sdb_analyzer_params *) malloc(sizeof(*m));
          *m = params;
          post(SDB_ANALYZER_MESSAGE_TYPE_PARAMS, m);
          break;
        }
      }
    }
  }

  long source_read(sdb_complex *dst, size_t n)
  {
    if (src.read) return src.read(src.priv, dst, n);
    if (!src.data) return -1;
    // in-memory capture in its native sample format (bps bytes per IQ pair); dst is a byte buffer of n * 8
    const size_t bps = src.input_format == SDB_FORMAT_FLOAT32 ? 8 : src.input_format == SDB_FORMAT_SIGNED16 ? 4 : 2;
    const unsigned char *base = reinterpret_cast<const unsigned char *>(src.data);
    unsigned char *out = reinterpret_cast<unsigned char *>(dst);
    size_t got = 0;
    while (got < n) {
      if (src_pos >= src.length) { if (!src.loop) break; src_pos = 0; }
      size_t take = std::min(n - got, src.length - src_pos);
      memcpy(out + got * bps, base + src_pos * bps, take * bps);
      got += take; src_pos += take;
    }
    return (long) got;
  }

  // ---------------------------------------------------------------------------------------------------------
  // SUSCAN_ANALYZER_MODE_WIDE_SPECTRUM: the panoramic scanner's analyzer (Panoramic/Scanner.cpp:296-372, 419-523):
  // retune, drop `buffering` samples (tuner round trip), take one window, post its PSD with the hop's centre; the GUI
  // stitches the hops (SpectrumView).  Hops are independent, so HB of them go through the engine as HB streams.
  // Hop plan (own definition, the upstream one is not in the reference): step = rel_bw fs; PROGRESSIVE: centres
  // min + step/2 + i step while the hop still starts below max, then wrap; STOCHASTIC: uniform in [min, max] from a
  // 32-bit LCG, snapped to the progressive grid when the partitioning is DISCRETE.  min == max: one centre.
  // ---------------------------------------------------------------------------------------------------------
  double hop_min = 0, hop_max = 0; float rel_bw = 0.5f; uint64_t buffering = 0;
  int strategy = SDB_SWEEP_STRATEGY_PROGRESSIVE, partitioning = SDB_SPECTRUM_PARTITIONING_DISCRETE;
  uint64_t hop_index = 0; uint32_t lcg = 0x5167D166u;
  double next_hop()
  {
    const double step = (double) rel_bw * src.samp_rate;
    if (!(hop_max > hop_min) || !(step > 0)) return hop_min;
    const uint64_t n_grid = (uint64_t) ceil((hop_max - hop_min) / step);
    if (strategy == SDB_SWEEP_STRATEGY_PROGRESSIVE) {
      const uint64_t i = hop_index++ % (n_grid ? n_grid : 1);
      return hop_min + 0.5 * step + (double) i * step;
    }
    lcg = lcg * 1664525u + 1013904223u;
    const double u = (double) (lcg >> 8) / 16777216.0;
    if (partitioning == SDB_SPECTRUM_PARTITIONING_DISCRETE) {
      uint64_t i = (uint64_t) (u * (double) n_grid);
      if (i >= n_grid) i = n_grid - 1;
      return hop_min + 0.5 * step + (double) i * step;
    }
    return hop_min + u * (hop_max - hop_min);
  }

  void run_wide()
  {
    const size_t HB = 16;
    uint32_t exit_type = SDB_WORKER_MSG_TYPE_HALT;
    std::vector<sdb_complex> buf, drop;
    std::vector<float> psd;
    std::vector<double> fcs(HB);
    hop_min = params.min_freq; hop_max = params.max_freq;
    for (;;) {
      { std::lock_guard<std::mutex> l(cmd_m); if (halt_req) break; }
      handle_cmds();
      const size_t N = (size_t) params.detector_params.window_size;
      if (plan_dirty) {
        if (eng) { sdb_engine_destroy(eng); eng = nullptr; }
        sdb_engine_params ep;
        memset(&ep, 0, sizeof(ep));
        ep.n_streams = (uint32_t) HB; ep.psd_size = (uint32_t) N; ep.psd_window = params.detector_params.window;
        ep.max_feed = (uint32_t) N; ep.device = src.device;
        ep.flags = iq_reverse ? SDB_FLAG_IQ_REVERSE : 0;
        ep.input_format = src.read ? SDB_FORMAT_FLOAT32 : src.input_format;   // in-memory captures keep their format
        eng = sdb_engine_new(&ep, src.samp_rate);
        if (!eng || sdb_engine_commit(eng)) {
          post_status(SDB_ANALYZER_MESSAGE_TYPE_SOURCE_INIT, SDB_ANALYZER_INIT_FAILURE, sdb_last_error());
          exit_type = SDB_ANALYZER_MESSAGE_TYPE_READ_ERROR; break;
        }
        plan_dirty = false;
      }
      // hops are packed back to back in the source's native sample format: bps bytes per IQ pair
      const size_t bps = src.read || src.input_format == SDB_FORMAT_FLOAT32 ? 8
                         : src.input_format == SDB_FORMAT_SIGNED16 ? 4 : 2;
      unsigned char *bytes = nullptr;
      buf.resize(HB * N);
      bytes = reinterpret_cast<unsigned char *>(buf.data());
      size_t got_hops = 0;
      bool eos = false, err = false;
      for (size_t h = 0; h < HB && !eos && !err; ++h) {
        fcs[h] = next_hop();
        if (src.set_frequency && src.set_frequency(src.priv, fcs[h]) != 0) { err = true; break; }
        for (uint64_t left = buffering; left > 0 && !eos && !err;) {          // the tuner's round trip
          const size_t take = (size_t) std::min<uint64_t>(left, 65536);
          drop.resize(take);
          const long g = source_read(drop.data(), take);
          if (g < 0) err = true; else if ((size_t) g < take) eos = true;
          left -= take;
        }
        if (eos || err) break;
        const long g = source_read(reinterpret_cast<sdb_complex *>(bytes + h * N * bps), N);
        if (g < 0) err = true; else if ((size_t) g < N) eos = true; else ++got_hops;
      }
      if (err) { exit_type = SDB_ANALYZER_MESSAGE_TYPE_READ_ERROR; break; }
      if (got_hops) {
        for (size_t h = got_hops; h < HB; ++h)
          memset(bytes + h * N * bps, src.input_format == SDB_FORMAT_UNSIGNED8 && !src.read ? 0x80 : 0, N * bps);
        if (sdb_engine_feed_host(eng, buf.data(), N, N) || sdb_engine_sync(eng)) {
          exit_type = SDB_ANALYZER_MESSAGE_TYPE_READ_ERROR; break;
        }
        psd.resize(HB * N);
        if (sdb_engine_read_psd(eng, psd.data(), psd.size())) { exit_type = SDB_ANALYZER_MESSAGE_TYPE_READ_ERROR; break; }
        for (size_t h = 0; h < got_hops; ++h) {
          sdb_analyzer_psd_msg *m = (sdb_analyzer_psd_msg *) calloc(1, sizeof(*m));
          m->fc = (int64_t) llround(fcs[h]); m->samp_rate = (float) src.samp_rate;
          m->measured_samp_rate = (float) measured_rate.load();
          gettimeofday(&m->rt_time, nullptr);
          m->psd_size = N;
          m->psd_data = (float *) malloc(N * sizeof(float));
          memcpy(m->psd_data, &psd[h * N], N * sizeof(float));
          post(SDB_ANALYZER_MESSAGE_TYPE_PSD, m);
        }
        total_samples += got_hops * (N + buffering);
      }
      if (eos) { exit_type = SDB_ANALYZER_MESSAGE_TYPE_EOS; break; }
    }
    if (eng) { sdb_engine_destroy(eng); eng = nullptr; }
    post_status(exit_type, 0, exit_type == SDB_ANALYZER_MESSAGE_TYPE_READ_ERROR ? sdb_last_error() : nullptr);
  }

  void run()
  {
    std::vector<sdb_complex> buf(block);
    std::vector<float> psd;
    std::vector<sdb_complex> soft;
    std::vector<uint8_t> hard;
    if (params.mode == SDB_ANALYZER_MODE_WIDE_SPECTRUM) {
      sdb_source_info *si = (sdb_source_info *) calloc(1, sizeof(*si));
      si->source_samp_rate = (uint64_t) src.samp_rate; si->effective_samp_rate = (uint64_t) src.samp_rate;
      si->frequency = src.freq; si->seekable = 0;
      post(SDB_ANALYZER_MESSAGE_TYPE_SOURCE_INFO, si);
      run_wide();
      return;
    }
    {
      sdb_source_info *si = (sdb_source_info *) calloc(1, sizeof(*si));
      si->source_samp_rate = (uint64_t) src.samp_rate; si->effective_samp_rate = (uint64_t) src.samp_rate;
      si->measured_samp_rate = 0; si->frequency = src.freq; si->seekable = src.read ? 0 : 1;
      post(SDB_ANALYZER_MESSAGE_TYPE_SOURCE_INFO, si);
    }
    auto t0 = std::chrono::steady_clock::now();
    uint32_t exit_type = SDB_WORKER_MSG_TYPE_HALT;
    for (;;) {
      { std::lock_guard<std::mutex> l(cmd_m); if (halt_req) break; }
      handle_cmds();
      const size_t N = (size_t) params.detector_params.window_size;
      if (block % N) block = std::max<size_t>(N, (block / N) * N);     // after a PARAMS change
      if (buf.size() != block) buf.resize(block);
      if (plan_dirty && !rebuild()) { exit_type = SDB_ANALYZER_MESSAGE_TYPE_READ_ERROR; break; }
      long got;
      const bool f32_src = src.read || src.input_format == SDB_FORMAT_FLOAT32;
      if (replaying && hist_fill > 0) {
        // suscan_analyzer_replay: the source is paused and the history ring is played back, oldest sample first
        for (size_t i = 0; i < block; ++i) {
          if (replay_pos == 0 && replay_wrapped) looped = 1;         // playing the ring for the second time
          buf[i] = hist[(hist_start + replay_pos) % hist.size()];
          if (++replay_pos >= hist_fill) { replay_pos = 0; replay_wrapped = true; }
        }
        got = (long) block;
      } else {
        got = source_read(buf.data(), block);
        if (got == (long) block && f32_src && !hist.empty()) {       // suscan_analyzer_set_history_size: keep the tail
          for (size_t i = 0; i < block; ++i) {
            hist[(hist_start + hist_fill) % hist.size()] = buf[i];
            if (hist_fill < hist.size()) ++hist_fill; else hist_start = (hist_start + 1) % hist.size();
          }
        }
      }
      if (got < 0) { exit_type = SDB_ANALYZER_MESSAGE_TYPE_READ_ERROR; break; }
      if ((size_t) got < block) { exit_type = SDB_ANALYZER_MESSAGE_TYPE_EOS; break; }   // partial tail blocks are dropped
      // baseband filters see (and may rewrite) every float32 block before the analyzer does, in registration order
      if (src.read || src.input_format == SDB_FORMAT_FLOAT32) {
        std::vector<BbFilter> fl;
        { std::lock_guard<std::mutex> l(cmd_m); fl = bb_filters; }
        for (auto &f : fl) f.fn(f.priv, this, buf.data(), (uint64_t) block, total_samples);
      }
      if (sdb_engine_feed_host(eng, buf.data(), block, block) || sdb_engine_sync(eng)) {
        exit_type = SDB_ANALYZER_MESSAGE_TYPE_READ_ERROR; break;
      }
      total_samples += block;
      total_samples_pub.store(total_samples);
      if (throttle > 0) {                                // suscan_analyzer_set_throttle_async: pace to `throttle` samples/s
        const double due = (double) (total_samples - throttle_s0) / throttle;
        const double now = std::chrono::duration<double>(std::chrono::steady_clock::now() - throttle_t0).count();
        if (due > now) std::this_thread::sleep_for(std::chrono::duration<double>(due - now));
      }
      double el = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      measured_rate.store(el > 0 ? (double) total_samples / el : 0);
      // PSD messages at psd_update_int cadence of SIGNAL time (coverage = N / (fs * interval))
      const size_t frames = block / N;
      psd.resize(frames * N);
      if (sdb_engine_read_psd(eng, psd.data(), psd.size()) == 0) {
        const double per_frame = (double) N / src.samp_rate;
        for (size_t f = 0; f < frames; ++f) {
          psd_credit += per_frame;
          if (psd_credit + 1e-12 >= params.psd_update_int) {
            psd_credit = params.psd_update_int > 0 ? fmod(psd_credit, params.psd_update_int) : 0;
            sdb_analyzer_psd_msg *m = (sdb_analyzer_psd_msg *) calloc(1, sizeof(*m));
            m->fc = (int64_t) src.freq; m->samp_rate = (float) src.samp_rate; m->measured_samp_rate = (float) measured_rate.load();
            gettimeofday(&m->rt_time, nullptr);
            double ts = (double) (total_samples - block + f * N) / src.samp_rate;
            m->timestamp.tv_sec = (time_t) ts; m->timestamp.tv_usec = (suseconds_t) ((ts - floor(ts)) * 1e6);
            m->looped = looped; m->history_size = (uint64_t) hist_fill;
            m->psd_size = N;
            m->psd_data = (float *) malloc(N * sizeof(float));
            memcpy(m->psd_data, &psd[f * N], N * sizeof(float));
            post(SDB_ANALYZER_MESSAGE_TYPE_PSD, m);
          }
        }
      }
      // sample batches, keyed by the caller-chosen inspector_id
      // sub-carrier inspectors: the parent's channel samples of this block through the child engine.  `subs` is in
      // parent-before-child order (rebuild), so a nested parent's engine has been fed by the time it is read.
      for (auto &sb : subs) {
        const Insp &p = insps[sb.parent];
        sdb_engine_t *host = engine_of(p);
        if (!p.open || p.engine_handle < 0 || !host) continue;
        std::vector<sdb_complex> ch(block);
        long nch = sdb_engine_read_channel(host, 0, p.engine_handle, ch.data(), ch.size());
        if (nch <= 0) continue;
        if (sdb_engine_feed_host(sb.eng, ch.data(), (size_t) nch, (size_t) nch) || sdb_engine_sync(sb.eng)) continue;
        for (auto &i : insps) {
          if (!i.open || i.parent != p.handle || !i.has_id || i.engine_handle < 0) continue;
          soft.resize((size_t) nch); hard.resize((size_t) nch);
          long n = sdb_engine_read_symbols(sb.eng, 0, i.engine_handle, soft.data(), hard.data(), (size_t) nch);
          if (n <= 0) continue;
          deliver(i, soft.data(), hard.data(), (size_t) n, false);
        }
      }
      for (auto &i : insps) {
        if (!i.open || !i.has_id || i.engine_handle < 0 || i.parent >= 0) continue;
        size_t cap = block;
        soft.resize(cap); hard.resize(cap);
        long n = sdb_engine_read_symbols(eng, 0, i.engine_handle, soft.data(), hard.data(), cap);
        if (n <= 0) continue;
        deliver(i, soft.data(), hard.data(), (size_t) n, false);
      }
      // kind=SPECTRUM / kind=ESTIMATOR inspector messages (one per block that filled a frame)
      for (auto &i : insps) {
        sdb_engine_t *ie = engine_of(i);
        if (!i.open || i.engine_handle < 0 || !i.spect_size || !ie) continue;
        if (i.spectsrc_id) {
          std::vector<float> sp(i.spect_size);
          uint32_t emitted = 0;
          if (sdb_engine_read_spectrum(ie, i.engine_handle, sp.data(), &emitted) == 0 && emitted) {
            sdb_analyzer_inspector_msg *m = (sdb_analyzer_inspector_msg *) calloc(1, sizeof(*m));
            m->kind = SDB_INSPECTOR_MSGKIND_SPECTRUM; m->handle = i.handle; m->inspector_id = i.inspector_id;
            m->class_name = dupstr((const char *[]){ "psk", "fsk", "ask", "audio", "raw" }[i.cls]);
            m->spectsrc_id = i.spectsrc_id; m->spectrum_size = emitted; m->samp_rate = (uint64_t) i.fs;
            m->fs = (float) src.samp_rate; m->equiv_fs = i.fs;
            m->spectrum_data = (float *) malloc(emitted * sizeof(float));
            memcpy(m->spectrum_data, sp.data(), emitted * sizeof(float));
            post(SDB_ANALYZER_MESSAGE_TYPE_INSPECTOR, m);
          }
        }
        for (int e = 0; e < SDB_ESTIMATOR_COUNT; ++e) {
          if (!(i.est_mask & (1u << e))) continue;
          float v = 0; int32_t ok = 0;
          if (sdb_engine_read_estimate(ie, i.engine_handle, e, &v, &ok) == 0 && ok) {
            sdb_analyzer_inspector_msg *m = (sdb_analyzer_inspector_msg *) calloc(1, sizeof(*m));
            m->kind = SDB_INSPECTOR_MSGKIND_ESTIMATOR; m->handle = i.handle; m->inspector_id = i.inspector_id;
            m->class_name = dupstr((const char *[]){ "psk", "fsk", "ask", "audio", "raw" }[i.cls]);
            m->estimator_id = (uint32_t) e; m->enabled = 1; m->value = v;
            post(SDB_ANALYZER_MESSAGE_TYPE_INSPECTOR, m);
          }
        }
      }
      // MESSAGE_TYPE_CHANNEL at channel_update_int cadence of signal time
      if (params.channel_update_int > 0) {
        chan_credit += (double) block / src.samp_rate;
        if (chan_credit + 1e-12 >= params.channel_update_int) {
          chan_credit = fmod(chan_credit, params.channel_update_int);
          std::vector<sdb_detected_channel> ch(256);
          uint32_t total = 0;
          long nc = sdb_engine_read_channels(eng, 0, src.freq, ch.data(), ch.size(), &total);
          if (nc >= 0) {
            sdb_analyzer_channel_msg *m = (sdb_analyzer_channel_msg *) calloc(1, sizeof(*m));
            m->channel_count = (uint32_t) nc;
            m->channel_list = (sdb_detected_channel *) malloc(std::max<long>(1, nc) * sizeof(sdb_detected_channel));
            memcpy(m->channel_list, ch.data(), (size_t) nc * sizeof(sdb_detected_channel));
            post(SDB_ANALYZER_MESSAGE_TYPE_CHANNEL, m);
          }
        }
      }
    }
    for (auto &i : insps)                                // what the watermark still holds back goes out with the end
      if (i.has_id && pending.count(i.handle)) deliver(i, nullptr, nullptr, 0, true);
    drop_subs();
    if (eng) { sdb_engine_destroy(eng); eng = nullptr; }
    post_status(exit_type, 0, exit_type == SDB_ANALYZER_MESSAGE_TYPE_READ_ERROR ? sdb_last_error() : nullptr);
  }
};

extern "C" sdb_analyzer_t *sdb_analyzer_new(const sdb_analyzer_params *params, const sdb_source_config *src)
{
  if (!params || !src) return nullptr;
  if (sdb_device_count() <= 0) return nullptr;   // sdb_last_error() is set by the engine on first use; no CPU fallback
  const uint64_t N = params->detector_params.window_size;
  if (N < 16 || (N & (N - 1)) || !(src->samp_rate > 0) || (!src->read && !src->data)) return nullptr;
  if (src->input_format < SDB_FORMAT_FLOAT32 || src->input_format > SDB_FORMAT_SIGNED16) return nullptr;
  sdb_analyzer *a = new sdb_analyzer();
  a->params = *params;
  a->src = *src;
  size_t blk = src->read_size ? src->read_size : (size_t) N * 8;
  blk = std::max<size_t>(N, (blk / N) * N);
  a->block = blk;
  a->worker = std::thread([a] { a->run(); });
  return a;
}

extern "C" void *sdb_analyzer_read(sdb_analyzer_t *a, uint32_t *type)
{
  if (!a) return nullptr;
  std::unique_lock<std::mutex> l(a->mq_m);
  a->mq_cv.wait(l, [a] { return !a->mq.empty(); });
  auto m = a->mq.front();
  a->mq.pop_front();
  if (type) *type = m.first;
  return m.second;
}

extern "C" void *sdb_analyzer_read_timeout(sdb_analyzer_t *a, uint32_t *type, unsigned timeout_ms)
{
  if (!a) return nullptr;
  std::unique_lock<std::mutex> l(a->mq_m);
  if (!a->mq_cv.wait_for(l, std::chrono::milliseconds(timeout_ms), [a] { return !a->mq.empty(); })) {
    if (type) *type = 0xfffffffeu;
    return nullptr;
  }
  auto m = a->mq.front();
  a->mq.pop_front();
  if (type) *type = m.first;
  return m.second;
}

extern "C" void sdb_analyzer_dispose_message(uint32_t type, void *ptr)
{
  if (!ptr) return;
  switch (type) {
    case SDB_ANALYZER_MESSAGE_TYPE_PSD:
      free(((sdb_analyzer_psd_msg *) ptr)->psd_data); break;
    case SDB_ANALYZER_MESSAGE_TYPE_SAMPLES:
      free(((sdb_analyzer_sample_batch_msg *) ptr)->samples);
      free(((sdb_analyzer_sample_batch_msg *) ptr)->symbols); break;
    case SDB_ANALYZER_MESSAGE_TYPE_INSPECTOR:
      free(((sdb_analyzer_inspector_msg *) ptr)->class_name);
      free(((sdb_analyzer_inspector_msg *) ptr)->spectrum_data); break;
    case SDB_ANALYZER_MESSAGE_TYPE_CHANNEL:
      free(((sdb_analyzer_channel_msg *) ptr)->channel_list); break;
    case SDB_ANALYZER_MESSAGE_TYPE_EOS:
    case SDB_ANALYZER_MESSAGE_TYPE_READ_ERROR:
    case SDB_ANALYZER_MESSAGE_TYPE_SOURCE_INIT:
    case SDB_WORKER_MSG_TYPE_HALT:
      free(((sdb_analyzer_status_msg *) ptr)->err_msg); break;
    default: break;
  }
  free(ptr);
}

extern "C" void sdb_analyzer_req_halt(sdb_analyzer_t *a)
{
  if (!a) return;
  std::lock_guard<std::mutex> l(a->cmd_m);
  a->halt_req = true;
}

extern "C" void sdb_analyzer_destroy(sdb_analyzer_t *a)
{
  if (!a) return;
  sdb_analyzer_req_halt(a);
  if (a->worker.joinable()) a->worker.join();
  for (auto &m : a->mq) sdb_analyzer_dispose_message(m.first, m.second);
  delete a;
}

static int push_cmd(sdb_analyzer_t *a, Cmd &&c)
{
  if (!a) return -1;
  std::lock_guard<std::mutex> l(a->cmd_m);
  a->cmds.push_back(std::move(c));
  return 0;
}

extern "C" int sdb_analyzer_open_ex_async(sdb_analyzer_t *a, const char *class_name, const sdb_sigutils_channel *ch,
                                          int precise, int32_t parent, uint32_t req_id)
{
  if (!class_name || !ch) return -1;
  if (parent < -1) return -1;
  // parent >= 0: sub-carrier inspection (Default/GenericInspector/GenericInspector.cpp:502-525); the channel is
  // given relative to the parent's channel centre, at the parent's equivalent sample rate
  Cmd c; c.kind = Cmd::OPEN; c.req_id = req_id; c.cls = class_name; c.channel = *ch; c.precise = precise;
  c.parent = parent;
  return push_cmd(a, std::move(c));
}
extern "C" int sdb_analyzer_set_inspector_id_async(sdb_analyzer_t *a, int32_t handle, uint32_t inspector_id, uint32_t req_id)
{
  Cmd c; c.kind = Cmd::SET_ID; c.req_id = req_id; c.handle = handle; c.inspector_id = inspector_id;
  return push_cmd(a, std::move(c));
}
extern "C" int sdb_analyzer_set_inspector_config_async(sdb_analyzer_t *a, int32_t handle, const sdb_inspector_config *cfg,
                                                       uint32_t req_id)
{
  if (!cfg) return -1;
  Cmd c; c.kind = Cmd::SET_CONFIG; c.req_id = req_id; c.handle = handle; c.cfg = *cfg;
  return push_cmd(a, std::move(c));
}
extern "C" int sdb_analyzer_close_async(sdb_analyzer_t *a, int32_t handle, uint32_t req_id)
{
  Cmd c; c.kind = Cmd::CLOSE; c.req_id = req_id; c.handle = handle;
  return push_cmd(a, std::move(c));
}
extern "C" int sdb_analyzer_set_inspector_watermark_async(sdb_analyzer_t *a, int32_t handle, uint64_t watermark,
                                                          uint32_t req_id)
{
  Cmd c; c.kind = Cmd::SET_WATERMARK; c.req_id = req_id; c.handle = handle; c.value = (double) watermark;
  return push_cmd(a, std::move(c));
}
extern "C" int sdb_analyzer_set_history_size(sdb_analyzer_t *a, uint64_t samples)
{
  if (samples > (1ull << 31)) return -1;
  Cmd c; c.kind = Cmd::SET_HISTORY; c.value = (double) samples;
  return push_cmd(a, std::move(c));
}
extern "C" int sdb_analyzer_replay(sdb_analyzer_t *a, int enabled)
{
  Cmd c; c.kind = Cmd::REPLAY; c.enabled = enabled;
  return push_cmd(a, std::move(c));
}
extern "C" int sdb_analyzer_set_hop_range(sdb_analyzer_t *a, double min_freq, double max_freq)
{
  if (min_freq > max_freq) return -1;
  Cmd c; c.kind = Cmd::SET_HOP_RANGE; c.value = min_freq; c.value2 = max_freq;
  return push_cmd(a, std::move(c));
}
extern "C" int sdb_analyzer_set_rel_bandwidth(sdb_analyzer_t *a, float rel_bw)
{
  Cmd c; c.kind = Cmd::SET_REL_BW; c.value = rel_bw;
  return push_cmd(a, std::move(c));
}
extern "C" int sdb_analyzer_set_buffering_size(sdb_analyzer_t *a, uint64_t samples)
{
  Cmd c; c.kind = Cmd::SET_BUFFERING; c.value = (double) samples;
  return push_cmd(a, std::move(c));
}
extern "C" int sdb_analyzer_set_sweep_strategy(sdb_analyzer_t *a, int strategy)
{
  if (strategy != SDB_SWEEP_STRATEGY_STOCHASTIC && strategy != SDB_SWEEP_STRATEGY_PROGRESSIVE) return -1;
  Cmd c; c.kind = Cmd::SET_STRATEGY; c.enabled = strategy;
  return push_cmd(a, std::move(c));
}
extern "C" int sdb_analyzer_set_spectrum_partitioning(sdb_analyzer_t *a, int partitioning)
{
  if (partitioning != SDB_SPECTRUM_PARTITIONING_DISCRETE && partitioning != SDB_SPECTRUM_PARTITIONING_CONTINUOUS) return -1;
  Cmd c; c.kind = Cmd::SET_PARTITIONING; c.enabled = partitioning;
  return push_cmd(a, std::move(c));
}
extern "C" int sdb_analyzer_seek(sdb_analyzer_t *a, const struct timeval *pos)
{
  if (!a || !pos || a->src.read || !a->src.data) return -1;      // seekable sources only (source_info.seekable)
  Cmd c; c.kind = Cmd::SEEK; c.value = (double) pos->tv_sec + 1e-6 * (double) pos->tv_usec;
  return push_cmd(a, std::move(c));
}
extern "C" int sdb_analyzer_set_inspector_freq_overridable(sdb_analyzer_t *a, int32_t handle, double freq)
{
  Cmd c; c.kind = Cmd::SET_FREQ; c.handle = handle; c.value = freq;
  return push_cmd(a, std::move(c));
}
extern "C" int sdb_analyzer_set_inspector_bandwidth_overridable(sdb_analyzer_t *a, int32_t handle, double bw)
{
  Cmd c; c.kind = Cmd::SET_BANDWIDTH; c.handle = handle; c.value = bw;
  return push_cmd(a, std::move(c));
}
extern "C" int sdb_analyzer_set_dc_remove(sdb_analyzer_t *a, int enabled)
{
  Cmd c; c.kind = Cmd::SET_DC_REMOVE; c.enabled = enabled;
  return push_cmd(a, std::move(c));
}
extern "C" int sdb_analyzer_set_iq_reverse(sdb_analyzer_t *a, int enabled)
{
  Cmd c; c.kind = Cmd::SET_IQ_REVERSE; c.enabled = enabled;
  return push_cmd(a, std::move(c));
}
extern "C" int sdb_analyzer_set_throttle_async(sdb_analyzer_t *a, uint64_t samp_rate, uint32_t req_id)
{
  Cmd c; c.kind = Cmd::SET_THROTTLE; c.req_id = req_id; c.value = (double) samp_rate;
  return push_cmd(a, std::move(c));
}
// suscan_analyzer_register_baseband_filter_with_prio (Suscan/Analyzer.cpp:137-143): filters run in ascending priority
// value, registration order among equals; the plain registration uses priority 0
extern "C" int sdb_analyzer_register_baseband_filter_prio(sdb_analyzer_t *a, sdb_baseband_filter_fn fn, void *priv,
                                                          int64_t prio)
{
  if (!a || !fn) return -1;
  std::lock_guard<std::mutex> l(a->cmd_m);
  auto it = a->bb_filters.begin();
  while (it != a->bb_filters.end() && it->prio <= prio) ++it;
  a->bb_filters.insert(it, { fn, priv, prio });
  return 0;
}
extern "C" int sdb_analyzer_register_baseband_filter(sdb_analyzer_t *a, sdb_baseband_filter_fn fn, void *priv)
{
  return sdb_analyzer_register_baseband_filter_prio(a, fn, priv, 0);
}
// signal time of the source in seconds since its start (suscan_analyzer_get_source_time, Suscan/Analyzer.cpp:145-149)
extern "C" double sdb_analyzer_get_source_time(const sdb_analyzer_t *a)
{
  return a ? (double) a->total_samples_pub.load() / a->src.samp_rate : 0.0;
}
// current configuration of an open inspector (what a GET_CONFIG reply would carry); -1 if the handle is not open
extern "C" int sdb_analyzer_get_inspector_config(sdb_analyzer_t *a, int32_t handle, sdb_inspector_config *cfg)
{
  if (!a || !cfg) return -1;
  std::lock_guard<std::mutex> l(a->insp_m);
  if (handle < 0 || handle >= (int32_t) a->insps.size() || !a->insps[handle].open) return -1;
  *cfg = a->insps[handle].cfg;
  return 0;
}
extern "C" int sdb_analyzer_inspector_set_spectrum_async(sdb_analyzer_t *a, int32_t handle, uint32_t spectsrc_id,
                                                         uint32_t req_id)
{
  Cmd c; c.kind = Cmd::SET_SPECTRUM; c.req_id = req_id; c.handle = handle; c.aux_id = spectsrc_id;
  return push_cmd(a, std::move(c));
}
extern "C" int sdb_analyzer_inspector_estimator_cmd_async(sdb_analyzer_t *a, int32_t handle, uint32_t estimator_id,
                                                          int enabled, uint32_t req_id)
{
  Cmd c; c.kind = Cmd::ESTIMATOR_CMD; c.req_id = req_id; c.handle = handle; c.aux_id = estimator_id; c.enabled = enabled;
  return push_cmd(a, std::move(c));
}
extern "C" int sdb_analyzer_set_params_async(sdb_analyzer_t *a, const sdb_analyzer_params *p, uint32_t req_id)
{
  if (!p) return -1;
  const uint64_t N = p->detector_params.window_size;
  if (N < 16 || (N & (N - 1))) return -1;
  Cmd c; c.kind = Cmd::SET_PARAMS; c.req_id = req_id; c.params = *p;
  return push_cmd(a, std::move(c));
}
extern "C" uint64_t sdb_analyzer_get_samp_rate(const sdb_analyzer_t *a) { return a ? (uint64_t) a->src.samp_rate : 0; }
extern "C" float sdb_analyzer_get_measured_samp_rate(const sdb_analyzer_t *a) { return a ? (float) a->measured_rate.load() : 0; }
