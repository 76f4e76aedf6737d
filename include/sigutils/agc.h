/* sigutils/agc.h -- look-ahead peak-tracking AGC (shim).  Call sites: Tasks/AGCTask.cpp:41-53 (`struct su_agc_params
 * agc_params = su_agc_params_INITIALIZER`, five time constants, su_agc_init), :70-73 (su_agc_feed); member of the
 * caller (include/AGCTask.h:39).  SPEC A. */
#ifndef _SIGUTILS_AGC_H
#define _SIGUTILS_AGC_H
#include <sigutils/types.h>
#ifdef __cplusplus
extern "C" {
#endif

struct su_agc_params {
  SUFLOAT threshold;          /* knee, dB */
  SUFLOAT slope_factor;       /* percent */
  unsigned int hang_max;
  unsigned int delay_line_size;
  unsigned int mag_history_size;
  SUFLOAT fast_rise_t, fast_fall_t, slow_rise_t, slow_fall_t;   /* samples */
};
#define su_agc_params_INITIALIZER { -100, 6, 100, 20, 20, 2, 4, 20, 40 }

struct sigutils_agc {
  SUBOOL  enabled;
  /* configuration (sdb AgcK) */
  SUFLOAT knee, slope_m1, fixed_gain, fast_alpha_rise, fast_alpha_fall, slow_alpha_rise, slow_alpha_fall;
  unsigned int hang_max, delay_line_size, mag_history_size;
  /* state (sdb AgcS) */
  SUFLOAT fast_level, slow_level, peak;
  unsigned int hang_n, delay_line_ptr, mag_history_ptr;
  SUFLOAT *delay_line;        /* 2 x delay_line_size floats (re, im interleaved), heap, owned */
  SUFLOAT *mag_history;       /* mag_history_size floats, heap, owned */
};
typedef struct sigutils_agc su_agc_t;
#define su_agc_INITIALIZER { 0 }

SUBOOL    su_agc_init(su_agc_t *agc, const struct su_agc_params *params);
SUCOMPLEX su_agc_feed(su_agc_t *agc, SUCOMPLEX x);
void      su_agc_finalize(su_agc_t *agc);

#ifdef __cplusplus
}
#endif
#endif
