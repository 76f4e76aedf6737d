/* sigutils/clock.h -- Gardner clock detector and interpolating sampler (shim).  Call sites:
 * Tasks/WaveSampler.cpp:60-66 (su_clock_detector_init(&cd, loopGain, bnor, bufsiz) != -1), :190-199 (feed),
 * :202-205 (count = su_clock_detector_read(&cd, block, max)); member of the caller (include/WaveSampler.h:44).
 * SPEC G: phase accumulator at two samples per symbol, linear interpolation, error Re{conj(x_mid)(x_now - x_prev)}. */
#ifndef _SIGUTILS_CLOCK_H
#define _SIGUTILS_CLOCK_H
#include <sigutils/types.h>
#ifdef __cplusplus
extern "C" {
#endif

struct sigutils_clock_detector {
  /* configuration */
  SUFLOAT gain, alpha, beta, bmin, bmax;
  /* state (sdb ClockS) */
  SUFLOAT phi, bnor, x0r, x0i, x1r, x1i, x2r, x2i, pr, pi;
  int     half;
  /* output stream drained by su_clock_detector_read */
  SUCOMPLEX *buf; SUSCOUNT buf_size, buf_avail;
};
typedef struct sigutils_clock_detector su_clock_detector_t;
#define su_clock_detector_INITIALIZER { 0 }

/* returns 0 on success, -1 on failure (the reference tests `!= -1`) */
int       su_clock_detector_init(su_clock_detector_t *cd, SUFLOAT loop_gain, SUFLOAT bhint, SUSCOUNT bufsiz);
void      su_clock_detector_set_baud(su_clock_detector_t *cd, SUFLOAT bnor);
SUBOOL    su_clock_detector_set_bnor_limits(su_clock_detector_t *cd, SUFLOAT lo, SUFLOAT hi);
void      su_clock_detector_feed(su_clock_detector_t *cd, SUCOMPLEX x);
SUSDIFF   su_clock_detector_read(su_clock_detector_t *cd, SUCOMPLEX *buf, SUSCOUNT size);
void      su_clock_detector_finalize(su_clock_detector_t *cd);

struct sigutils_sampler {
  SUFLOAT bnor, period, phase, phase0, phase0_rel;
  SUCOMPLEX prev;
};
typedef struct sigutils_sampler su_sampler_t;
#define su_sampler_INITIALIZER { 0 }
SUBOOL su_sampler_init(su_sampler_t *s, SUFLOAT bnor);
SUBOOL su_sampler_set_rate(su_sampler_t *s, SUFLOAT bnor);
void   su_sampler_set_phase(su_sampler_t *s, SUFLOAT phase_rel);
SUBOOL su_sampler_feed(su_sampler_t *s, SUCOMPLEX *sample);   /* in: new sample; out (when SU_TRUE): the symbol */
void   su_sampler_finalize(su_sampler_t *s);

#ifdef __cplusplus
}
#endif
#endif
