/* analyzer/source/info.h -- struct suscan_source_info, the payload of SUSCAN_ANALYZER_MESSAGE_TYPE_SOURCE_INFO (shim).
 * Fields read by the reference: include/Suscan/Analyzer.h:113-254, Suscan/Messages/SourceInfoMessage.cpp:25-47. */
#ifndef _ANALYZER_SOURCE_INFO_H
#define _ANALYZER_SOURCE_INFO_H
#include <sigutils/types.h>
#include <sys/time.h>
#ifdef __cplusplus
extern "C" {
#endif

struct suscan_source_gain_info { char *name; SUFLOAT min, max, step, value; };
struct suscan_source_info {
  uint64_t permissions;
  SUSCOUNT source_samp_rate, effective_samp_rate;
  SUFLOAT  measured_samp_rate;
  SUFREQ   frequency, freq_min, freq_max, lnb;
  SUFLOAT  bandwidth, ppm;
  char    *antenna;
  SUBOOL   dc_remove, iq_reverse, agc, seekable, replay;
  SUSCOUNT history_length;
  struct timeval source_time, source_start, source_end;
  struct suscan_source_gain_info **gain_list; unsigned int gain_count;
  char   **antenna_list; unsigned int antenna_count;
};
void   suscan_source_info_init(struct suscan_source_info *info);
SUBOOL suscan_source_info_init_copy(struct suscan_source_info *dst, const struct suscan_source_info *src);
void   suscan_source_info_finalize(struct suscan_source_info *info);

#ifdef __cplusplus
}
#endif
#endif
