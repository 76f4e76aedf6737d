/* sigutils/softtune.h -- struct sigutils_channel, the channel descriptor of the detector and of open_ex (shim).
 * Call sites: include/Suscan/Channel.h:23-32 (the wrapper's Channel is built from it), Suscan/Analyzer.cpp:417-424
 * (fc, ft, f_lo, f_hi, bw filled for suscan_analyzer_open_ex_async), Suscan/Messages/ChannelMessage.cpp:25-70. */
#ifndef _SIGUTILS_SOFTTUNE_H
#define _SIGUTILS_SOFTTUNE_H
#include <sigutils/types.h>
#ifdef __cplusplus
extern "C" {
#endif

struct sigutils_channel {
  SUFREQ  fc, f_lo, f_hi;
  SUFLOAT bw, snr, S0, N0;
  SUFREQ  ft;
  uint32_t age, present;
};
#define sigutils_channel_INITIALIZER { 0, 0, 0, 0, 0, 0, 0, 0, 0, 0 }

#ifdef __cplusplus
}
#endif
#endif
