/* sigutils/ncqo.h -- numerically controlled quadrature oscillator (shim).  Call sites: Tasks/CarrierXlator.cpp:36-37
 * (su_ncqo_init(&ncqo, -relFreq); su_ncqo_set_phase(&ncqo, -phase)) and :57-60 (dst = src * su_ncqo_read(&ncqo));
 * the object is a member of the caller (include/CarrierXlator.h:39).  SPEC N: omega = pi fnor, read() returns
 * exp(i phi) for the current phase and then advances it, phi kept in [0, 2 pi).  The per-sample calls compute on the
 * host with the arithmetic of the GPU kernels (sdb_chain_steps.h); su_ncqo_read_bulk / sdb_task_carrier_xlate are the
 * device paths. */
#ifndef _SIGUTILS_NCQO_H
#define _SIGUTILS_NCQO_H
#include <sigutils/types.h>
#ifdef __cplusplus
extern "C" {
#endif

struct sigutils_ncqo {
  SUFLOAT phi;     /* current phase, [0, 2 pi) */
  SUFLOAT omega;   /* rad / sample */
  SUFLOAT fnor;    /* the normalised frequency it was built from */
};
typedef struct sigutils_ncqo su_ncqo_t;
#define su_ncqo_INITIALIZER { 0, 0, 0 }

void      su_ncqo_init(su_ncqo_t *ncqo, SUFLOAT fnor);
void      su_ncqo_set_phase(su_ncqo_t *ncqo, SUFLOAT phi);
SUFLOAT   su_ncqo_get_phase(const su_ncqo_t *ncqo);
void      su_ncqo_inc_phase(su_ncqo_t *ncqo, SUFLOAT delta);
void      su_ncqo_set_freq(su_ncqo_t *ncqo, SUFLOAT fnor);
void      su_ncqo_set_angfreq(su_ncqo_t *ncqo, SUFLOAT omega);
void      su_ncqo_inc_angfreq(su_ncqo_t *ncqo, SUFLOAT delta);
SUFLOAT   su_ncqo_get_freq(const su_ncqo_t *ncqo);
SUFLOAT   su_ncqo_get_angfreq(const su_ncqo_t *ncqo);
SUCOMPLEX su_ncqo_read(su_ncqo_t *ncqo);
/* dst[i] = src[i] * read(): the body of CarrierXlator::work as one device pass (n samples) */
SUBOOL    su_ncqo_mix_bulk(su_ncqo_t *ncqo, const SUCOMPLEX *src, SUCOMPLEX *dst, SUSCOUNT n);

#ifdef __cplusplus
}
#endif
#endif
