/* sigutils/tvproc.h -- analog-TV processor (shim).  Call sites: Default/GenericInspector/TVProcessorWorker.cpp:120-151
 * (`if (su_tv_processor_feed(processor, *samples++))` per sample, `emit frame(su_tv_processor_take_frame(processor))`),
 * :175 (destroy), :204 (new; NULL = invalid parameters), :215-218 (return_frame / su_tv_frame_buffer_destroy), :222
 * (set_params); parameter block filled at Default/GenericInspector/TVProcessorTab.cpp:549-597, presets :629,633;
 * frame consumed as {width, height, buffer} by the TV display (TVProcessorTab.cpp:657-663).
 * su_tv_processor_feed is a per-sample call by ABI: it runs, on the caller's thread, the step the GPU kernel runs
 * (sdb_tv_steps.h, SPEC TV) -- bit-identical to sdb_tv_processor_feed over the same samples. */
#ifndef _SIGUTILS_TVPROC_H
#define _SIGUTILS_TVPROC_H
#include <sigutils/types.h>
#ifdef __cplusplus
extern "C" {
#endif

struct sigutils_tv_processor_params {
  SUBOOL   enable_sync, reverse, interlace, enable_agc;
  SUFLOAT  x_off;
  SUBOOL   dominance;
  SUSCOUNT frame_lines;
  SUFLOAT  frame_spacing;
  SUBOOL   enable_comb, comb_reverse;
  SUFLOAT  hsync_len, vsync_len, line_len;
  SUSCOUNT vsync_odd_trigger;
  SUFLOAT  t_tol, l_tol, g_tol;
  SUFLOAT  hsync_huge_err, hsync_max_err, hsync_min_err;
  SUFLOAT  hsync_len_tau, line_len_tau, agc_tau, hsync_fast_track_tau, hsync_slow_track_tau;
};

struct sigutils_tv_frame_buffer {
  int width, height;
  SUFLOAT *buffer;                              /* [height][width], 0 = black ... 1 = white */
  struct sigutils_tv_frame_buffer *next;        /* free-pool link */
};

typedef struct sigutils_tv_processor su_tv_processor_t;

void   su_tv_processor_params_pal(struct sigutils_tv_processor_params *p, SUFLOAT samp_rate);
void   su_tv_processor_params_ntsc(struct sigutils_tv_processor_params *p, SUFLOAT samp_rate);
su_tv_processor_t *su_tv_processor_new(const struct sigutils_tv_processor_params *p);
SUBOOL su_tv_processor_set_params(su_tv_processor_t *t, const struct sigutils_tv_processor_params *p);
SUBOOL su_tv_processor_feed(su_tv_processor_t *t, SUFLOAT x);          /* SU_TRUE: a frame was completed */
struct sigutils_tv_frame_buffer *su_tv_processor_take_frame(su_tv_processor_t *t);
void   su_tv_processor_return_frame(su_tv_processor_t *t, struct sigutils_tv_frame_buffer *f);
void   su_tv_frame_buffer_destroy(struct sigutils_tv_frame_buffer *f);
void   su_tv_processor_destroy(su_tv_processor_t *t);

#ifdef __cplusplus
}
#endif
#endif
