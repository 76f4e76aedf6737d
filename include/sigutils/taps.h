/* sigutils/taps.h -- window and filter taps (shim).  Call site: Tasks/CarrierDetector.cpp:87-89
 * (su_taps_apply_blackmann_harris_complex(buffer, size)); window enum include/Suscan/AnalyzerParams.h:37-43.
 * SPEC W: symmetric cosine-sum windows (denominator size - 1), evaluated in binary64 and rounded once. */
#ifndef _SIGUTILS_TAPS_H
#define _SIGUTILS_TAPS_H
#include <sigutils/types.h>
#ifdef __cplusplus
extern "C" {
#endif
void su_taps_apply_hamming(SUFLOAT *h, SUSCOUNT size);
void su_taps_apply_hann(SUFLOAT *h, SUSCOUNT size);
void su_taps_apply_flat_top(SUFLOAT *h, SUSCOUNT size);
void su_taps_apply_blackmann_harris(SUFLOAT *h, SUSCOUNT size);
void su_taps_apply_hamming_complex(SUCOMPLEX *h, SUSCOUNT size);
void su_taps_apply_hann_complex(SUCOMPLEX *h, SUSCOUNT size);
void su_taps_apply_flat_top_complex(SUCOMPLEX *h, SUSCOUNT size);
void su_taps_apply_blackmann_harris_complex(SUCOMPLEX *h, SUSCOUNT size);
void su_taps_rrc_init(SUFLOAT *h, SUFLOAT T, SUFLOAT beta, SUSCOUNT size);
void su_taps_brickwall_lp_init(SUFLOAT *h, SUFLOAT fc, SUSCOUNT size);
#ifdef __cplusplus
}
#endif
#endif
