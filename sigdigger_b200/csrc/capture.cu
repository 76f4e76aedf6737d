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
