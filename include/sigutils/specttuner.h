/* sigutils/specttuner.h -- FFT filter-bank channeliser (shim over the sdb engine).  Call sites:
 * Tasks/LPFTask.cpp:52-69 (params / channel params INITIALIZERs, f0, bw = pi bw_norm, guard = 2 pi / bw, privdata,
 * on_data), :83-87 (su_specttuner_feed_bulk), :104-107 (flush one zero at a time), :119-123 (destroy closes the
 * channels); callback contract :28-42 (the data pointer stays valid until the next feed; return SU_FALSE aborts the
 * feed).  SPEC S.  The forward transform, the bin gather and the per-channel inverse transforms run on the GPU
 * (sdb_engine_*); feed_bulk buffers input up to the next half window and delivers the channel samples of every
 * completed hop, in channel-open order, from the calling thread. */
#ifndef _SIGUTILS_SPECTTUNER_H
#define _SIGUTILS_SPECTTUNER_H
#include <sigutils/types.h>
#include <sigutils/sampling.h>   /* SU_NORM2ANG_FREQ and friends: Tasks/LPFTask.cpp:64 uses them through this header */
#ifdef __cplusplus
extern "C" {
#endif

struct sigutils_specttuner_params {
  SUSCOUNT window_size;
  SUBOOL   early_windowing;
};
#define sigutils_specttuner_params_INITIALIZER { 4096, SU_TRUE }

struct sigutils_specttuner_channel;
typedef struct sigutils_specttuner_channel su_specttuner_channel_t;
typedef SUBOOL (*su_specttuner_on_data_fn)(const su_specttuner_channel_t *channel, void *privdata,
                                           const SUCOMPLEX *data, SUSCOUNT size);

struct sigutils_specttuner_channel_params {
  SUFLOAT f0;       /* centre, rad / sample, [0, 2 pi) */
  SUFLOAT delta_f;
  SUFLOAT bw;       /* rad / sample */
  SUFLOAT guard;    /* >= 1: the channel rate is at least guard x bw */
  SUBOOL  precise;  /* per-sample LO for the residual sub-bin offset */
  void   *privdata;
  su_specttuner_on_data_fn on_data;
};
#define sigutils_specttuner_channel_params_INITIALIZER { 0, 0, 0, 1, SU_FALSE, NULL, NULL }

struct sigutils_specttuner_channel {
  struct sigutils_specttuner_channel_params params;
  int          index;      /* engine handle */
  SUFLOAT      k;          /* 1 / window_size */
  SUFLOAT      decimation;
  unsigned int center, size, width, halfw, halfsz;
};

struct sigutils_specttuner;
typedef struct sigutils_specttuner su_specttuner_t;

su_specttuner_t         *su_specttuner_new(const struct sigutils_specttuner_params *params);
su_specttuner_channel_t *su_specttuner_open_channel(su_specttuner_t *st,
                                                    const struct sigutils_specttuner_channel_params *params);
SUBOOL                   su_specttuner_close_channel(su_specttuner_t *st, su_specttuner_channel_t *channel);
SUBOOL                   su_specttuner_feed_bulk(su_specttuner_t *st, const SUCOMPLEX *buf, SUSCOUNT size);
SUSCOUNT                 su_specttuner_get_channel_count(const su_specttuner_t *st);
void                     su_specttuner_destroy(su_specttuner_t *st);
SUFLOAT                  su_specttuner_channel_get_decimation(const su_specttuner_channel_t *channel);

#ifdef __cplusplus
}
#endif
#endif
