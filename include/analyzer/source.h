/* analyzer/source.h -- source configuration object (shim): what the reference's Suscan::Source::Config wraps
 * (Suscan/Source.cpp:26-659).  Types served: "file" (raw float32 / u8 / s8 / s16, WAV, SigMF through
 * sdb_capture_open: Default/SourceConfig/FileSourcePage.cpp:80-104) and "tonegen" (Default/SourceConfig/
 * ToneGenSourcePage.cpp:81-87); an in-memory buffer or a read callback can be attached for tests and embedders.
 * Device back-ends (SoapySDR, stdin, remote) are outside the hot path and report an init failure. */
#ifndef _SUSCAN_SOURCE_H
#define _SUSCAN_SOURCE_H
#include <sigutils/types.h>
#include <sys/time.h>
#ifdef __cplusplus
extern "C" {
#endif

enum suscan_source_format {
  SUSCAN_SOURCE_FORMAT_AUTO = 0,
  SUSCAN_SOURCE_FORMAT_RAW_FLOAT32,
  SUSCAN_SOURCE_FORMAT_RAW_UNSIGNED8,
  SUSCAN_SOURCE_FORMAT_RAW_SIGNED16,
  SUSCAN_SOURCE_FORMAT_RAW_SIGNED8,
  SUSCAN_SOURCE_FORMAT_WAV,
  SUSCAN_SOURCE_FORMAT_SIGMF
};

#define SUSCAN_SOURCE_DEFAULT_NAME      "Default source"
#define SUSCAN_SOURCE_DEFAULT_FREQ      433920000
#define SUSCAN_SOURCE_DEFAULT_SAMP_RATE 1000000
#define SUSCAN_SOURCE_DEFAULT_BANDWIDTH SUSCAN_SOURCE_DEFAULT_SAMP_RATE

struct suscan_source_config;
typedef struct suscan_source_config suscan_source_config_t;

suscan_source_config_t *suscan_source_config_new(const char *type, enum suscan_source_format format);
suscan_source_config_t *suscan_source_config_clone(const suscan_source_config_t *config);
void        suscan_source_config_destroy(suscan_source_config_t *config);
SUBOOL      suscan_source_config_set_type_format(suscan_source_config_t *c, const char *type, enum suscan_source_format f);
const char *suscan_source_config_get_type(const suscan_source_config_t *c);
enum suscan_source_format suscan_source_config_get_format(const suscan_source_config_t *c);
SUBOOL      suscan_source_config_set_label(suscan_source_config_t *c, const char *label);
const char *suscan_source_config_get_label(const suscan_source_config_t *c);
SUBOOL      suscan_source_config_set_path(suscan_source_config_t *c, const char *path);
const char *suscan_source_config_get_path(const suscan_source_config_t *c);
void        suscan_source_config_set_freq(suscan_source_config_t *c, SUFREQ freq);
SUFREQ      suscan_source_config_get_freq(const suscan_source_config_t *c);
void        suscan_source_config_set_lnb_freq(suscan_source_config_t *c, SUFREQ freq);
SUFREQ      suscan_source_config_get_lnb_freq(const suscan_source_config_t *c);
void        suscan_source_config_set_samp_rate(suscan_source_config_t *c, unsigned int samp_rate);
unsigned int suscan_source_config_get_samp_rate(const suscan_source_config_t *c);
void        suscan_source_config_set_average(suscan_source_config_t *c, unsigned int average);
unsigned int suscan_source_config_get_average(const suscan_source_config_t *c);
void        suscan_source_config_set_bandwidth(suscan_source_config_t *c, SUFLOAT bw);
SUFLOAT     suscan_source_config_get_bandwidth(const suscan_source_config_t *c);
void        suscan_source_config_set_loop(suscan_source_config_t *c, SUBOOL loop);
SUBOOL      suscan_source_config_get_loop(const suscan_source_config_t *c);
void        suscan_source_config_set_dc_remove(suscan_source_config_t *c, SUBOOL dc_remove);
SUBOOL      suscan_source_config_get_dc_remove(const suscan_source_config_t *c);
void        suscan_source_config_set_iq_balance(suscan_source_config_t *c, SUBOOL iq_balance);
SUBOOL      suscan_source_config_get_iq_balance(const suscan_source_config_t *c);
void        suscan_source_config_set_ppm(suscan_source_config_t *c, SUFLOAT ppm);
SUFLOAT     suscan_source_config_get_ppm(const suscan_source_config_t *c);
void        suscan_source_config_set_start_time(suscan_source_config_t *c, struct timeval tv);
void        suscan_source_config_get_start_time(const suscan_source_config_t *c, struct timeval *tv);
SUBOOL      suscan_source_config_set_param(suscan_source_config_t *c, const char *key, const char *val);
const char *suscan_source_config_get_param(const suscan_source_config_t *c, const char *key);
void        suscan_source_config_clear_params(suscan_source_config_t *c);
SUBOOL      suscan_source_config_is_seekable(const suscan_source_config_t *c);
SUBOOL      suscan_source_config_is_real_time(const suscan_source_config_t *c);
SUBOOL      suscan_source_config_file_is_valid(const suscan_source_config_t *c);
/* embedders / tests: an in-memory capture (format = the config's raw format, length in IQ pairs; not copied) or a
 * read callback delivering complex float32 (returns samples read, 0 = end of stream, < 0 = error) */
SUBOOL      suscan_source_config_set_memory(suscan_source_config_t *c, const void *data, SUSCOUNT length);
typedef SUSDIFF (*suscan_source_read_fn)(void *priv, SUCOMPLEX *dst, SUSCOUNT max);
SUBOOL      suscan_source_config_set_read_callback(suscan_source_config_t *c, suscan_source_read_fn fn, void *priv);
/* GPU the analyzer of this source runs on (default 0) and samples per worker-loop block (0: 8 PSD windows) */
void        suscan_source_config_set_gpu(suscan_source_config_t *c, int device);
void        suscan_source_config_set_read_size(suscan_source_config_t *c, SUSCOUNT samples);

#ifdef __cplusplus
}
#endif
#endif
