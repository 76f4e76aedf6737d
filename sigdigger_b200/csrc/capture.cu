// capture.cu -- capture-file front end of the path: container parsing only; the sample
// conversion itself happens inside the first load of the transforms (SPEC Q, sdb_iq.h).  Host code, no kernels.
//
// Reference behaviour: the file source offers AUTO / raw float32 / u8 / s8 / s16 / WAV / SigMF
// (Default/SourceConfig/FileSourcePage.cpp:68-104) and guesses sample rate, frequency and start time from the file
// name (FileSourcePage.cpp:107-140; the GUI's own recordings are named
// "sigdigger_%Y%m%d_%H%M%SZ_<samp_rate>_<freq>_float32_iq.raw", Default/Source/SourceWidget.cpp:1092-1100).
// The parsers themselves live in suscan (not in the reference): RIFF/WAVE and SigMF are public formats and are
// read here from their specifications -- two-channel PCM8 (unsigned) / PCM16 / IEEE float32 WAV with I on the left
// channel, SigMF "core:datatype" in {cf32_le, ci16_le, ci8, cu8} with "core:sample_rate" and the first capture's
// "core:frequency".
#include "../../include/sigdigger_b200.h"
#include <fcntl.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>
#include <string>

struct sdb_capture {
  int fd = -1;
  void *map = nullptr; size_t map_len = 0;
  sdb_capture_info info{};
};

static thread_local std::string g_cap_err;
extern "C" const char *sdb_capture_last_error(void) { return g_cap_err.c_str(); }
static sdb_capture_t *cap_fail(sdb_capture *c, const char *msg)
{
  g_cap_err = msg;
  if (c) sdb_capture_close(c);
  return nullptr;
}

static size_t fmt_bytes(int f) { return f == SDB_FORMAT_FLOAT32 ? 8 : f == SDB_FORMAT_SIGNED16 ? 4 : 2; }

static bool ends_with(const std::string &s, const char *suf)
{
  const size_t n = strlen(suf);
  return s.size() >= n && strcasecmp(s.c_str() + s.size() - n, suf) == 0;
}

// "sigdigger_20240131_235959Z_2000000_433920000_float32_iq.raw"
static void guess_from_name(const std::string &path, sdb_capture_info *info)
{
  const size_t slash = path.find_last_of('/');
  const std::string base = slash == std::string::npos ? path : path.substr(slash + 1);
  int Y, M, D, h, m, s;
  long long rate;
  double freq;
  char fmt[32];
  if (sscanf(base.c_str(), "sigdigger_%4d%2d%2d_%2d%2d%2dZ_%lld_%lf_%31[a-z0-9]_iq", &Y, &M, &D, &h, &m, &s, &rate, &freq,
             fmt) == 9) {
    struct tm tm;
    memset(&tm, 0, sizeof(tm));
    tm.tm_year = Y - 1900; tm.tm_mon = M - 1; tm.tm_mday = D; tm.tm_hour = h; tm.tm_min = m; tm.tm_sec = s;
    info->start_time = (int64_t) timegm(&tm);
    info->samp_rate = (double) rate; info->frequency = freq;
    info->guessed |= SDB_CAPTURE_GUESS_START_TIME | SDB_CAPTURE_GUESS_SAMP_RATE | SDB_CAPTURE_GUESS_FREQ;
    int f = -1;
    if (!strcmp(fmt, "float32")) f = SDB_FORMAT_FLOAT32;
    else if (!strcmp(fmt, "unsigned8")) f = SDB_FORMAT_UNSIGNED8;
    else if (!strcmp(fmt, "signed8")) f = SDB_FORMAT_SIGNED8;
    else if (!strcmp(fmt, "signed16")) f = SDB_FORMAT_SIGNED16;
    if (f >= 0) { info->sample_format = f; info->guessed |= SDB_CAPTURE_GUESS_FORMAT; }
  }
}

static uint32_t rd32(const unsigned char *p) { return p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32_t) p[3] << 24); }
static uint16_t rd16(const unsigned char *p) { return (uint16_t) (p[0] | (p[1] << 8)); }

static bool parse_wav(const unsigned char *d, size_t n, sdb_capture_info *info)
{
  if (n < 12 || memcmp(d, "RIFF", 4) || memcmp(d + 8, "WAVE", 4)) { g_cap_err = "not a RIFF/WAVE file"; return false; }
  size_t off = 12;
  bool have_fmt = false;
  unsigned tag = 0, ch = 0, bits = 0, rate = 0;
  while (off + 8 <= n) {
    const uint32_t len = rd32(d + off + 4);
    if (!memcmp(d + off, "fmt ", 4) && off + 8 + 16 <= n) {
      tag = rd16(d + off + 8); ch = rd16(d + off + 10); rate = rd32(d + off + 12); bits = rd16(d + off + 22);
      if (tag == 0xfffe && len >= 40 && off + 8 + 26 <= n) tag = rd16(d + off + 8 + 24);    // WAVE_FORMAT_EXTENSIBLE
      have_fmt = true;
    } else if (!memcmp(d + off, "data", 4)) {
      if (!have_fmt) { g_cap_err = "WAV: data chunk before fmt chunk"; return false; }
      if (ch != 2) { g_cap_err = "WAV: IQ captures need two channels (I left, Q right)"; return false; }
      int f;
      if (tag == 1 && bits == 8) f = SDB_FORMAT_UNSIGNED8;
      else if (tag == 1 && bits == 16) f = SDB_FORMAT_SIGNED16;
      else if (tag == 3 && bits == 32) f = SDB_FORMAT_FLOAT32;
      else { g_cap_err = "WAV: unsupported sample type (PCM8, PCM16 and float32 are read)"; return false; }
      size_t avail = n - (off + 8);
      size_t bytes = len == 0xffffffffu || len > avail ? avail : len;      // streaming writers leave the length open
      info->container = SDB_CONTAINER_WAV; info->sample_format = f; info->samp_rate = (double) rate;
      info->data_offset = off + 8; info->n_samples = bytes / fmt_bytes(f);
      return true;
    }
    off += 8 + (size_t) len + (len & 1u);
  }
  g_cap_err = "WAV: no data chunk";
  return false;
}

// minimal JSON field extraction: the value of the first `"key"` found at or after `from`
static bool json_value(const std::string &js, const char *key, size_t from, std::string *out)
{
  const std::string k = std::string("\"") + key + "\"";
  size_t p = js.find(k, from);
  if (p == std::string::npos) return false;
  p = js.find(':', p + k.size());
  if (p == std::string::npos) return false;
  ++p;
  while (p < js.size() && (js[p] == ' ' || js[p] == '\t' || js[p] == '\n' || js[p] == '\r')) ++p;
  if (p >= js.size()) return false;
  if (js[p] == '"') {
    const size_t e = js.find('"', p + 1);
    if (e == std::string::npos) return false;
    *out = js.substr(p + 1, e - p - 1);
  } else {
    size_t e = p;
    while (e < js.size() && js[e] != ',' && js[e] != '}' && js[e] != ']' && js[e] != '\n') ++e;
    *out = js.substr(p, e - p);
  }
  return true;
}

static bool parse_sigmf_meta(const std::string &meta_path, sdb_capture_info *info)
{
  FILE *f = fopen(meta_path.c_str(), "rb");
  if (!f) { g_cap_err = "SigMF: cannot open " + meta_path; return false; }
  std::string js;
  char buf[4096];
  size_t r;
  while ((r = fread(buf, 1, sizeof(buf), f)) > 0) js.append(buf, r);
  fclose(f);
  std::string v;
  if (!json_value(js, "core:datatype", 0, &v)) { g_cap_err = "SigMF: no core:datatype"; return false; }
  int fm;
  if (v == "cf32_le" || v == "cf32") fm = SDB_FORMAT_FLOAT32;
  else if (v == "ci16_le" || v == "ci16") fm = SDB_FORMAT_SIGNED16;
  else if (v == "ci8") fm = SDB_FORMAT_SIGNED8;
  else if (v == "cu8") fm = SDB_FORMAT_UNSIGNED8;
  else { g_cap_err = "SigMF: unsupported core:datatype " + v; return false; }
  info->container = SDB_CONTAINER_SIGMF; info->sample_format = fm;
  if (json_value(js, "core:sample_rate", 0, &v)) info->samp_rate = atof(v.c_str());
  const size_t caps = js.find("\"captures\"");
  if (caps != std::string::npos && json_value(js, "core:frequency", caps, &v)) info->frequency = atof(v.c_str());
  return true;
}

extern "C" sdb_capture_t *sdb_capture_open(const char *path, int32_t container, int32_t sample_format,
                                            sdb_capture_info *out)
{
  if (!path || !out) return cap_fail(nullptr, "null argument");
  sdb_capture *c = new sdb_capture();
  sdb_capture_info &info = c->info;
  memset(&info, 0, sizeof(info));
  info.sample_format = sample_format >= 0 ? sample_format : SDB_FORMAT_FLOAT32;
  std::string p(path), data_path(path);
  if (container < 0) {                                   // SUSCAN_SOURCE_FORMAT_AUTO: by extension
    if (ends_with(p, ".wav")) container = SDB_CONTAINER_WAV;
    else if (ends_with(p, ".sigmf-meta") || ends_with(p, ".sigmf-data") || ends_with(p, ".sigmf")) container = SDB_CONTAINER_SIGMF;
    else container = SDB_CONTAINER_RAW;
  }
  if (container == SDB_CONTAINER_SIGMF) {
    std::string stem = p;
    for (const char *suf : { ".sigmf-meta", ".sigmf-data", ".sigmf" })
      if (ends_with(stem, suf)) { stem.resize(stem.size() - strlen(suf)); break; }
    if (!parse_sigmf_meta(stem + ".sigmf-meta", &info)) { delete c; return nullptr; }
    data_path = stem + ".sigmf-data";
  }
  c->fd = open(data_path.c_str(), O_RDONLY);
  if (c->fd < 0) return cap_fail(c, ("cannot open " + data_path).c_str());
  struct stat st;
  if (fstat(c->fd, &st) != 0 || st.st_size <= 0) return cap_fail(c, "empty or unreadable capture");
  c->map_len = (size_t) st.st_size;
  c->map = mmap(nullptr, c->map_len, PROT_READ, MAP_PRIVATE, c->fd, 0);
  if (c->map == MAP_FAILED) { c->map = nullptr; return cap_fail(c, "mmap failed"); }
  if (container == SDB_CONTAINER_WAV) {
    if (!parse_wav((const unsigned char *) c->map, c->map_len, &info)) { sdb_capture_close(c); return nullptr; }
  } else if (container == SDB_CONTAINER_SIGMF) {
    info.data_offset = 0; info.n_samples = c->map_len / fmt_bytes(info.sample_format);
  } else {
    info.container = SDB_CONTAINER_RAW;
    guess_from_name(p, &info);
    if (sample_format >= 0) info.sample_format = sample_format;      // an explicit format wins over the name
    if (info.sample_format < SDB_FORMAT_FLOAT32 || info.sample_format > SDB_FORMAT_SIGNED16)
      return cap_fail(c, "unknown sample format");
    info.data_offset = 0; info.n_samples = c->map_len / fmt_bytes(info.sample_format);
  }
  if (info.n_samples == 0) return cap_fail(c, "capture holds no complete sample");
  *out = info;
  return c;
}

extern "C" const void *sdb_capture_data(const sdb_capture_t *c)
{
  return c && c->map ? (const unsigned char *) c->map + c->info.data_offset : nullptr;
}

extern "C" void sdb_capture_close(sdb_capture_t *c)
{
  if (!c) return;
  if (c->map) munmap(c->map, c->map_len);
  if (c->fd >= 0) close(c->fd);
  delete c;
}

// ------------------------------------------------------------------------------------------------
// Recording side (SURVEY.md 8(f) rank 2 "source decode + recording").  Host code, no kernels.
//
// Reference behaviour:
//  * baseband capture: the GUI registers a baseband filter whose body writes the analyzer's complex float32
//    blocks to a file named "sigdigger_%Y%m%d_%H%M%SZ_<rate>_<freq>_float32_iq.raw"
//    (Default/Source/SourceWidget.cpp:1078-1100,1156-1171; Misc/FileDataSaver.cpp:68-82: plain write()).
//  * audio: mono 16-bit PCM WAV of Re{x}, "audio-<MOD>-<freq>-<rate>-<NNNN>.wav", first free index
//    (Audio/AudioFileSaver.cpp:57-107,131-152).
//  * inspector recording: one of five data variables (Default/GenericInspector/InspectorUI.cpp:860-930).
// Beyond the reference (its writers are float32-only) the recorder also stores u8 / s8 / s16 and the WAV /
// SigMF containers the capture reader above understands, with the inverse of the SPEC Q scaling, so that a
// recording can be played back through sdb_capture_open() without an external tool.
// ------------------------------------------------------------------------------------------------
#include <errno.h>
#include <math.h>
#include <vector>
#include "sdb_math.h"

struct sdb_recorder {
  int fd = -1;
  int container = SDB_CONTAINER_RAW, format = SDB_FORMAT_FLOAT32;
  bool audio = false;                       // mono PCM16 of the real part
  double samp_rate = 0, frequency = 0;
  int64_t start_time = 0;
  uint64_t samples = 0, data_bytes = 0;
  std::string path, meta_path;
  std::vector<unsigned char> tmp;
};

static const char *fmt_token(int f)
{
  return f == SDB_FORMAT_UNSIGNED8 ? "unsigned8" : f == SDB_FORMAT_SIGNED8 ? "signed8"
       : f == SDB_FORMAT_SIGNED16 ? "signed16" : "float32";
}

extern "C" int sdb_capture_file_name(char *dst, size_t cap, int64_t utc_seconds, int32_t sample_format,
                                     double samp_rate, double frequency)
{
  if (!dst || cap == 0) { g_cap_err = "null argument"; return -1; }
  time_t t = (time_t) utc_seconds;
  struct tm tm;
  char datetime[24];
  gmtime_r(&t, &tm);
  strftime(datetime, sizeof(datetime), "%Y%m%d_%H%M%SZ", &tm);
  const int n = snprintf(dst, cap, "sigdigger_%s_%d_%.0lf_%s_iq.raw", datetime, (int) samp_rate, frequency,
                         fmt_token(sample_format));
  if (n < 0 || (size_t) n >= cap) { g_cap_err = "file name buffer too small"; return -1; }
  return n;
}

static const char *demod_token(int d)
{
  switch (d) {                                // enum AudioDemod, include/SigDiggerHelpers.h:39-45
    case 0: return "AM"; case 1: return "FM"; case 2: return "USB"; case 3: return "LSB"; default: return "RAW";
  }
}

static void wr16(unsigned char *p, unsigned v) { p[0] = v & 0xff; p[1] = (v >> 8) & 0xff; }
static void wr32(unsigned char *p, uint32_t v) { wr16(p, v & 0xffff); wr16(p + 2, v >> 16); }

// 44-byte canonical header; sizes are patched again on close
static bool write_wav_header(sdb_recorder *r)
{
  unsigned char h[44];
  const unsigned ch = r->audio ? 1 : 2;
  const unsigned bits = r->audio ? 16 : r->format == SDB_FORMAT_FLOAT32 ? 32 : r->format == SDB_FORMAT_SIGNED16 ? 16 : 8;
  const unsigned tag = (!r->audio && r->format == SDB_FORMAT_FLOAT32) ? 3 : 1;
  const uint32_t rate = (uint32_t) (r->samp_rate + 0.5);
  const uint32_t data = r->data_bytes > 0xffffffffull - 36 ? 0xffffffffu - 36 : (uint32_t) r->data_bytes;
  memcpy(h, "RIFF", 4); wr32(h + 4, 36 + data); memcpy(h + 8, "WAVEfmt ", 8); wr32(h + 16, 16);
  wr16(h + 20, tag); wr16(h + 22, ch); wr32(h + 24, rate); wr32(h + 28, rate * ch * (bits / 8));
  wr16(h + 32, ch * (bits / 8)); wr16(h + 34, bits); memcpy(h + 36, "data", 4); wr32(h + 40, data);
  return pwrite(r->fd, h, sizeof(h), 0) == (ssize_t) sizeof(h);
}

static bool write_sigmf_meta(const sdb_recorder *r)
{
  FILE *f = fopen(r->meta_path.c_str(), "w");
  if (!f) return false;
  const char *dt = r->format == SDB_FORMAT_UNSIGNED8 ? "cu8" : r->format == SDB_FORMAT_SIGNED8 ? "ci8"
                 : r->format == SDB_FORMAT_SIGNED16 ? "ci16_le" : "cf32_le";
  time_t t = (time_t) r->start_time;
  struct tm tm;
  char iso[32];
  gmtime_r(&t, &tm);
  strftime(iso, sizeof(iso), "%Y-%m-%dT%H:%M:%SZ", &tm);
  fprintf(f, "{\n  \"global\": {\n    \"core:datatype\": \"%s\",\n    \"core:sample_rate\": %.17g,\n"
             "    \"core:version\": \"1.0.0\",\n    \"core:recorder\": \"sigdigger_b200\"\n  },\n"
             "  \"captures\": [\n    { \"core:sample_start\": 0, \"core:frequency\": %.17g, \"core:datetime\": \"%s\" }\n  ],\n"
             "  \"annotations\": []\n}\n", dt, r->samp_rate, r->frequency, iso);
  return fclose(f) == 0;
}

static sdb_recorder_t *rec_fail(sdb_recorder *r, const std::string &msg)
{
  g_cap_err = msg;
  if (r) { if (r->fd >= 0) close(r->fd); delete r; }
  return nullptr;
}

extern "C" sdb_recorder_t *sdb_recorder_open(const char *path, int32_t auto_name, const sdb_recorder_params *p)
{
  if (!path || !p) return rec_fail(nullptr, "null argument");
  if (p->sample_format < SDB_FORMAT_FLOAT32 || p->sample_format > SDB_FORMAT_SIGNED16)
    return rec_fail(nullptr, "unknown sample format");
  if (p->container < SDB_CONTAINER_RAW || p->container > SDB_CONTAINER_SIGMF) return rec_fail(nullptr, "unknown container");
  if (p->container == SDB_CONTAINER_WAV && p->sample_format == SDB_FORMAT_SIGNED8)
    return rec_fail(nullptr, "WAV has no signed 8-bit PCM");
  if (!(p->samp_rate > 0)) return rec_fail(nullptr, "sample rate must be positive");
  sdb_recorder *r = new sdb_recorder();
  r->container = p->container; r->format = p->sample_format; r->samp_rate = p->samp_rate; r->frequency = p->frequency;
  r->start_time = p->start_time ? p->start_time : (int64_t) time(nullptr);
  r->path = path;
  if (auto_name) {                              // `path` is the directory (DataSaverUI's record path)
    char name[128];
    if (sdb_capture_file_name(name, sizeof(name), r->start_time, r->format, r->samp_rate, r->frequency) < 0)
      return rec_fail(r, "file name buffer too small");
    std::string base(name);
    if (r->container == SDB_CONTAINER_WAV) base.replace(base.size() - 4, 4, ".wav");
    else if (r->container == SDB_CONTAINER_SIGMF) base.resize(base.size() - 4);
    r->path = std::string(path) + "/" + base;
  }
  if (r->container == SDB_CONTAINER_SIGMF) {
    std::string stem = r->path;
    for (const char *suf : { ".sigmf-meta", ".sigmf-data", ".sigmf" })
      if (ends_with(stem, suf)) { stem.resize(stem.size() - strlen(suf)); break; }
    r->meta_path = stem + ".sigmf-meta";
    r->path = stem + ".sigmf-data";
  }
  r->fd = open(r->path.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0600);   // creat(path, 0600), SourceWidget.cpp:1102
  if (r->fd < 0) return rec_fail(r, "cannot create " + r->path + ": " + strerror(errno));
  if (r->container == SDB_CONTAINER_WAV) {
    if (!write_wav_header(r) || lseek(r->fd, 44, SEEK_SET) != 44) return rec_fail(r, "cannot write the WAV header");
  } else if (r->container == SDB_CONTAINER_SIGMF && !write_sigmf_meta(r)) {
    return rec_fail(r, "cannot write " + r->meta_path);
  }
  return r;
}

extern "C" sdb_recorder_t *sdb_audio_recorder_open(const char *dir, int32_t demod, double frequency, uint32_t samp_rate)
{
  if (!dir || samp_rate == 0) return rec_fail(nullptr, "invalid argument");
  sdb_recorder *r = new sdb_recorder();
  r->audio = true; r->container = SDB_CONTAINER_WAV; r->format = SDB_FORMAT_SIGNED16;
  r->samp_rate = samp_rate; r->frequency = frequency; r->start_time = (int64_t) time(nullptr);
  char name[128];
  unsigned index = 1;
  do {
    snprintf(name, sizeof(name), "audio-%s-%.0lf-%d-%04u.wav", demod_token(demod), frequency, (int) samp_rate, index++);
    r->path = std::string(dir) + "/" + name;
  } while (access(r->path.c_str(), F_OK) != -1 && index < 10000);
  r->fd = open(r->path.c_str(), O_WRONLY | O_CREAT | O_EXCL, 0644);
  if (r->fd < 0) return rec_fail(r, "Save file " + r->path + " failed: " + strerror(errno));
  if (!write_wav_header(r) || lseek(r->fd, 44, SEEK_SET) != 44) return rec_fail(r, "cannot write the WAV header");
  return r;
}

extern "C" const char *sdb_recorder_path(const sdb_recorder_t *r) { return r ? r->path.c_str() : nullptr; }
extern "C" uint64_t sdb_recorder_samples(const sdb_recorder_t *r) { return r ? r->samples : 0; }

static bool write_all(sdb_recorder *r, const void *data, size_t len)
{
  const unsigned char *p = (const unsigned char *) data;
  while (len) {
    const ssize_t w = write(r->fd, p, len);
    if (w < 1) { g_cap_err = std::string("write() failed: ") + strerror(errno); return false; }
    p += w; len -= (size_t) w;
  }
  return true;
}

static inline long sat(float v, long lo, long hi, long nan_code = 0)
{
  if (!(v == v)) return nan_code;                         // NaN -> the code of 0.0
  const float q = nearbyintf(v);
  return q < (float) lo ? lo : q > (float) hi ? hi : (long) q;
}

// x: n complex float32 samples (the baseband-filter hook's `samples`, or a SAMPLES message)
extern "C" long sdb_recorder_write(sdb_recorder_t *r, const sdb_complex *x, size_t n)
{
  if (!r || (!x && n)) { g_cap_err = "null argument"; return -1; }
  if (r->fd < 0) return 0;
  const float *f = (const float *) x;
  size_t bytes;
  const void *src;
  if (r->audio) {
    r->tmp.resize(n * 2);
    for (size_t i = 0; i < n; ++i) wr16(&r->tmp[2 * i], (unsigned) (sat(f[2 * i] * 32767.0f, -32768, 32767) & 0xffff));
    src = r->tmp.data(); bytes = n * 2;
  } else if (r->format == SDB_FORMAT_FLOAT32) {
    src = x; bytes = n * 8;
  } else if (r->format == SDB_FORMAT_SIGNED16) {
    r->tmp.resize(n * 4);
    for (size_t i = 0; i < 2 * n; ++i) wr16(&r->tmp[2 * i], (unsigned) (sat(f[i] * 32768.0f, -32768, 32767) & 0xffff));
    src = r->tmp.data(); bytes = n * 4;
  } else {
    r->tmp.resize(n * 2);
    if (r->format == SDB_FORMAT_UNSIGNED8)
      for (size_t i = 0; i < 2 * n; ++i) r->tmp[i] = (unsigned char) sat(f[i] * 128.0f + 128.0f, 0, 255, 128);
    else
      for (size_t i = 0; i < 2 * n; ++i) r->tmp[i] = (unsigned char) (sat(f[i] * 128.0f, -128, 127) & 0xff);
    src = r->tmp.data(); bytes = n * 2;
  }
  if (!write_all(r, src, bytes)) return -1;
  r->samples += n; r->data_bytes += bytes;
  return (long) n;
}

extern "C" int sdb_recorder_close(sdb_recorder_t *r)
{
  if (!r) return 0;
  bool ok = true;
  if (r->fd >= 0) {
    if (r->container == SDB_CONTAINER_WAV) ok = write_wav_header(r);
    ok = (close(r->fd) == 0) && ok;
  }
  if (!ok) g_cap_err = "closing " + r->path + " failed";
  delete r;
  return ok ? 0 : -1;
}

// Inspector recording / forwarding formats (Default/GenericInspector/InspectorUI.cpp:860-930): what the data
// saver receives for each "data variable".  dst must hold n floats, n complex or n bytes; returns bytes.
extern "C" long sdb_inspector_forward(int32_t data_var, int32_t decision_mode, const sdb_complex *soft,
                                      const uint8_t *hard, size_t n, void *dst)
{
  if (!dst || (n && !soft && data_var != SDB_DATAVAR_SYMBOLS)) { g_cap_err = "null argument"; return -1; }
  const float *s = (const float *) soft;
  float *o = (float *) dst;
  switch (data_var) {
    case SDB_DATAVAR_DECISION_SPACE:
      if (decision_mode == 1) for (size_t i = 0; i < n; ++i) o[i] = d_cabsf(s[2 * i], s[2 * i + 1]);
      else for (size_t i = 0; i < n; ++i)                               // arg(i x) / pi
        o[i] = d_atan2f(s[2 * i], -s[2 * i + 1]) / 3.14159265358979323846f;
      return (long) (n * sizeof(float));
    case SDB_DATAVAR_SOFT_BITS:
      memcpy(dst, soft, n * 8);
      return (long) (n * 8);
    case SDB_DATAVAR_SOFT_BITS_I:
      for (size_t i = 0; i < n; ++i) o[i] = s[2 * i];
      return (long) (n * sizeof(float));
    case SDB_DATAVAR_SOFT_BITS_Q:
      for (size_t i = 0; i < n; ++i) o[i] = s[2 * i + 1];
      return (long) (n * sizeof(float));
    case SDB_DATAVAR_SYMBOLS:
      if (!hard && n) { g_cap_err = "no decision available"; return -1; }
      memcpy(dst, hard, n);
      return (long) n;
    default:
      g_cap_err = "unknown data variable";
      return -1;
  }
}

// ------------------------------------------------------------------------------------------------
// Host-side rules of the audio inspector's caller (Default/Audio/AudioProcessor.cpp): which channel is
// opened, where the LO and the bandwidth really go for the side-band modes, and what is written into the
// audio.* configuration keys.  Pure parameter arithmetic in binary64, as in the reference.
// ------------------------------------------------------------------------------------------------
extern "C" int sdb_audio_plan_make(double analyzer_samp_rate, uint32_t requested_rate, int32_t demod, double lo,
                                   double bw, sdb_audio_plan *out)
{
  if (!out || !(analyzer_samp_rate > 0) || demod < 0 || demod > 4) { g_cap_err = "invalid argument"; return -1; }
  memset(out, 0, sizeof(*out));
  // openAudio(), AudioProcessor.cpp:118-128: m_maxAudioBw = min(fs / 2, 2e5); the rate is floored to it
  const double max_bw = analyzer_samp_rate / 2 < 2e5 ? analyzer_samp_rate / 2 : 2e5;
  uint32_t rate = requested_rate;
  if ((double) rate > max_bw) rate = (uint32_t) floor(max_bw);
  if (rate < 1) { g_cap_err = "Audio device does not support the current sample rate"; return -1; }
  // calcTrueBandwidth(), :200-214
  double tbw = bw;
  if (demod == 2 || demod == 3) tbw *= .5;
  if (tbw > max_bw) tbw = max_bw; else if (tbw < 1) tbw = 1;
  // calcTrueLoFreq(), :216-228
  double delta = 0;
  if (demod == 2) delta += .5 * tbw; else if (demod == 3) delta -= .5 * tbw;
  out->max_audio_bw = max_bw; out->sample_rate = rate; out->true_bw = tbw; out->true_lo = lo + delta;
  // the channel of requestOpen("audio", ch), :143-151
  out->ch_bw = max_bw; out->ch_ft = 0; out->ch_fc = out->true_lo;
  out->ch_f_lo = -.5 * max_bw; out->ch_f_hi = +.5 * max_bw;
  if (out->ch_fc > max_bw || out->ch_fc < -max_bw) out->ch_fc = 0;
  return 0;
}

// setParams(), AudioProcessor.cpp:250-269: volume is handled at UI level (1), the demodulator goes +1 on the wire
extern "C" int sdb_audio_plan_config(const sdb_audio_plan *plan, int32_t demod, float cutoff, int32_t squelch,
                                     float squelch_level, int32_t agc, float agc_ts, sdb_inspector_config *cfg)
{
  if (!plan || !cfg || demod < 0 || demod > 4) { g_cap_err = "invalid argument"; return -1; }
  cfg->audio_cutoff = cutoff;
  cfg->audio_volume = 1.f;
  cfg->audio_sample_rate = plan->sample_rate;
  cfg->audio_demod = (uint32_t) demod + 1;
  cfg->audio_squelch = squelch ? 1 : 0;
  cfg->audio_squelch_level = squelch_level;
  cfg->agc_enabled = agc ? 1 : 0;
  cfg->agc_ts = agc_ts;
  return 0;
}
