// ref_shim/fftw3.h -- ORACLE SUPPORT (test infrastructure): the five FFTW calls Tasks/CarrierDetector.cpp makes
// (fftwf_malloc / _free / _plan_dft_1d / _execute / _destroy_plan), so that the reference's carrier detector compiles
// and runs in oracle/_ref without FFTW (absent from this image; the GUI links the real fftw3f itself,
// SigDigger.pro:486 -- FFTW is not part of the suscan / sigutils boundary).  The transform behind them is a plain
// radix-2 FFT evaluated in binary64 (ref_glue.cpp): accurate to float rounding, any FFTW plan would agree to that.
#ifndef REF_SHIM_FFTW3_H
#define REF_SHIM_FFTW3_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef float fftwf_complex[2];
typedef struct ref_fftwf_plan_s *fftwf_plan;
#define FFTW_FORWARD  (-1)
#define FFTW_BACKWARD (+1)
#define FFTW_ESTIMATE (1U << 6)
void      *fftwf_malloc(size_t n);
void       fftwf_free(void *p);
fftwf_plan fftwf_plan_dft_1d(int n, fftwf_complex *in, fftwf_complex *out, int sign, unsigned flags);
void       fftwf_execute(const fftwf_plan p);
void       fftwf_destroy_plan(fftwf_plan p);
#ifdef __cplusplus
}
#endif
#endif
