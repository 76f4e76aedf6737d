/* sigutils/pll.h -- PLL and Costas loop objects (shim).  Call sites: Tasks/PLLSyncTask.cpp:36 (su_pll_init(&pll, 0,
 * bw)), :53-56 (su_pll_track); Tasks/CostasRecoveryTask.cpp:41 (su_costas_init(&costas, kind, 0, bw, 3, loopbw)),
 * :58-61 (destination[p] = su_costas_feed(&costas, origin[p])).  Both objects are members of the caller
 * (include/PLLSyncTask.h:39, include/CostasRecoveryTask.h:39: `su_costas_t costas = su_costas_INITIALIZER`), so
 * their layout is part of this header.  SPEC C: arm filter "order 3" = 2-pole Butterworth at arm_bw, error
 * -I Q | sgn cross products | 8PSK variant, loop gains a = pi loop_bw, b = a^2 / 2.
 * Per-sample calls compute on the host with the arithmetic of the GPU kernels; the _bulk calls are device passes
 * (they restart from the object's state and write it back). */
#ifndef _SIGUTILS_PLL_H
#define _SIGUTILS_PLL_H
#include <sigutils/types.h>
#include <sigutils/ncqo.h>
#ifdef __cplusplus
extern "C" {
#endif

#define SU_PLL_MAX_IIR 5

struct sigutils_pll {
  SUFLOAT alpha, beta;   /* loop gains */
  SUFLOAT phi, omega;    /* NCO phase and angular frequency */
  SUFLOAT lock;
  SUCOMPLEX a;           /* last output */
};
typedef struct sigutils_pll su_pll_t;
#define su_pll_INITIALIZER { 0, 0, 0, 0, 0, 0 }

SUBOOL    su_pll_init(su_pll_t *pll, SUFLOAT fhint, SUFLOAT fc);
SUCOMPLEX su_pll_track(su_pll_t *pll, SUCOMPLEX x);
SUBOOL    su_pll_track_bulk(su_pll_t *pll, const SUCOMPLEX *x, SUCOMPLEX *y, SUSCOUNT n);
void      su_pll_finalize(su_pll_t *pll);

enum sigutils_costas_kind {
  SU_COSTAS_KIND_NONE = 0,
  SU_COSTAS_KIND_BPSK,
  SU_COSTAS_KIND_QPSK,
  SU_COSTAS_KIND_8PSK
};

struct sigutils_costas {
  /* configuration (sdb CostasK) */
  int     kind, af_n;
  SUFLOAT a, b;
  SUFLOAT af_b[SU_PLL_MAX_IIR], af_a[SU_PLL_MAX_IIR];
  /* state (sdb CostasS) */
  SUFLOAT phi, omega, lock, y_re, y_im;
  SUFLOAT xr[SU_PLL_MAX_IIR], xi[SU_PLL_MAX_IIR], yr[SU_PLL_MAX_IIR], yi[SU_PLL_MAX_IIR];
  /* what callers read after a feed */
  SUCOMPLEX y;
};
typedef struct sigutils_costas su_costas_t;
#define su_costas_INITIALIZER { 0 }

SUBOOL    su_costas_init(su_costas_t *costas, enum sigutils_costas_kind kind, SUFLOAT fhint, SUFLOAT arm_bw,
                         unsigned int arm_order, SUFLOAT loop_bw);
SUCOMPLEX su_costas_feed(su_costas_t *costas, SUCOMPLEX x);
SUBOOL    su_costas_feed_bulk(su_costas_t *costas, const SUCOMPLEX *x, SUCOMPLEX *y, SUSCOUNT n);
void      su_costas_set_kind(su_costas_t *costas, enum sigutils_costas_kind kind);
void      su_costas_set_loop_bw(su_costas_t *costas, SUFLOAT loop_bw);
void      su_costas_finalize(su_costas_t *costas);

#ifdef __cplusplus
}
#endif
#endif
