/* sigutils/iir.h -- IIR / FIR filter object (shim).  Call sites: Tasks/WaveSampler.cpp:68-80 (su_iir_rrc_init(&mf,
 * span, tau, rolloff), su_iir_filt_feed, su_iir_filt_finalize), member of the caller (include/WaveSampler.h:50).
 * SPEC I.1: y[n] = sum b_i x[n-i] - sum_{i>=1} a_i y[n-i], one accumulator, ascending i, fused terms; I.2: RRC taps,
 * Hamming-windowed, unity DC gain. */
#ifndef _SIGUTILS_IIR_H
#define _SIGUTILS_IIR_H
#include <sigutils/types.h>
#ifdef __cplusplus
extern "C" {
#endif

struct sigutils_iir_filt {
  unsigned int x_size, y_size;   /* feed-forward taps, feedback coefficients (0: FIR) */
  unsigned int x_ptr, y_ptr;
  SUFLOAT   *a, *b;              /* heap, owned */
  SUCOMPLEX *x, *y;              /* delay lines, heap, owned */
  SUCOMPLEX  curr_y;
  SUFLOAT    gain;
};
typedef struct sigutils_iir_filt su_iir_filt_t;
#define su_iir_filt_INITIALIZER { 0, 0, 0, 0, NULL, NULL, NULL, NULL, 0, 1 }

SUBOOL    su_iir_filt_init(su_iir_filt_t *filt, SUSCOUNT y_size, const SUFLOAT *a, SUSCOUNT x_size, const SUFLOAT *b);
SUBOOL    su_iir_rrc_init(su_iir_filt_t *filt, SUSCOUNT n, SUFLOAT T, SUFLOAT beta);
SUBOOL    su_iir_bwlpf_init(su_iir_filt_t *filt, SUSCOUNT order, SUFLOAT fc);
SUBOOL    su_iir_brickwall_lp_init(su_iir_filt_t *filt, SUSCOUNT n, SUFLOAT fc);
SUCOMPLEX su_iir_filt_feed(su_iir_filt_t *filt, SUCOMPLEX x);
void      su_iir_filt_feed_bulk(su_iir_filt_t *filt, const SUCOMPLEX *x, SUCOMPLEX *y, SUSCOUNT len);
SUCOMPLEX su_iir_filt_get(const su_iir_filt_t *filt);
void      su_iir_filt_reset(su_iir_filt_t *filt);
void      su_iir_filt_set_gain(su_iir_filt_t *filt, SUFLOAT gain);
void      su_iir_filt_finalize(su_iir_filt_t *filt);

#ifdef __cplusplus
}
#endif
#endif
