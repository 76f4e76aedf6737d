/*
 * sigutils/types.h -- drop-in header of the sigdigger_b200 shim (libsigutils.so).
 *
 * The reference includes <sigutils/types.h> everywhere (include/QuadDemodTask.h:23, include/Suscan/Compat.h, ...)
 * and uses the scalar types and macros below; the sigutils sources themselves are absent from /root/reference
 * (SURVEY.md H1-H2), so the names and the meaning are fixed by the callers and the values by SURVEY.md Appendix A.
 * Everything that computes is declared here and implemented in sigdigger_b200/csrc/sigutils_shim.cu on top of the
 * C-ABI of include/sigdigger_b200.h.
 */
#ifndef _SIGUTILS_TYPES_H
#define _SIGUTILS_TYPES_H

#include <stdint.h>
#include <stddef.h>
#include <math.h>

#ifdef __cplusplus
#  include <complex>
typedef std::complex<float> SUCOMPLEX;
#  define SU_C_REAL(c) ((c).real())
#  define SU_C_IMAG(c) ((c).imag())
#  define SU_C_ABS(c)  (std::abs(c))
#  define SU_C_ARG(c)  (std::arg(c))
#  define SU_C_CONJ(c) (std::conj(c))
#  define SU_C_EXP(c)  (std::exp(c))
#  define SU_I         (SUCOMPLEX(0.0f, 1.0f))
#else
#  include <complex.h>
typedef float _Complex SUCOMPLEX;
#  define SU_C_REAL(c) crealf(c)
#  define SU_C_IMAG(c) cimagf(c)
#  define SU_C_ABS(c)  cabsf(c)
#  define SU_C_ARG(c)  cargf(c)
#  define SU_C_CONJ(c) conjf(c)
#  define SU_C_EXP(c)  cexpf(c)
#  define SU_I         _Complex_I
#endif

typedef float    SUFLOAT;
typedef double   SUDOUBLE;
typedef double   SUFREQ;
typedef uint64_t SUSCOUNT;
typedef int64_t  SUSDIFF;
typedef int      SUBOOL;
typedef int32_t  SUHANDLE;
typedef uint32_t SUBITS;

/* sigutils/types.h carries these for its callers (SU_ATTEMPT, include/Suscan/Compat.h:28-36) */
#ifndef STRINGIFY
#  define _STRINGIFY(x) #x
#  define STRINGIFY(x) _STRINGIFY(x)
#endif

#define SU_TRUE  1
#define SU_FALSE 0

#ifndef PI
#  define PI 3.14159265358979323846
#endif

#define SU_ADDSFX(x)   x##f
#define SU_ASFLOAT(x)  ((SUFLOAT) (x))
#define SU_SQRT(x)     sqrtf(x)
#define SU_POW(x, y)   powf(x, y)
#define SU_LOG(x)      log10f(x)
#define SU_LN(x)       logf(x)
#define SU_EXP(x)      expf(x)
#define SU_FLOOR(x)    floorf(x)
#define SU_CEIL(x)     ceilf(x)
#define SU_ABS(x)      fabsf(x)
#define SU_COS(x)      cosf(x)
#define SU_SIN(x)      sinf(x)
#define SU_MIN(a, b)   ((a) < (b) ? (a) : (b))
#define SU_MAX(a, b)   ((a) > (b) ? (a) : (b))
#define SU_RAD2DEG(r)  ((r) * (180 / PI))
#define SU_DEG2RAD(d)  ((d) * (PI / 180))

/* decibels: "power" = 10 log10, "magnitude" = 20 log10 (Suscan/Messages/PSDMessage.cpp:32-38 applies SU_POWER_DB
 * to the PSD bins; Default/GenericInspector/GenericInspector.cpp:232-254 adds 1e-20 by hand) */
#define SUFLOAT_MIN_REF_MAG  1e-8f                     /* SPEC M.6: the floor the PSD kernels' dB epilogue uses too */
#define SU_POWER_DB_RAW(p)  (10 * SU_LOG(p))
#define SU_POWER_DB(p)      SU_POWER_DB_RAW((p) + SUFLOAT_MIN_REF_MAG)
#define SU_DB_RAW(m)        (20 * SU_LOG(m))
#define SU_DB(m)            SU_DB_RAW((m) + 1e-10f)
#define SU_POWER_MAG_RAW(d) SU_POW(10, (d) * .1f)
#define SU_POWER_MAG(d)     SU_POWER_MAG_RAW(d)
#define SU_MAG_RAW(d)       SU_POW(10, (d) * .05f)
#define SU_MAG(d)           SU_MAG_RAW(d)

/* single-pole low-pass: alpha from a time constant in samples, y += alpha (x - y) */
#define SU_SPLPF_ALPHA(tau)     (1.f - SU_EXP(-1.f / (tau)))
#define SU_SPLPF_FEED(y, x, a)  (y) += (a) * ((x) - (y))

/* the FFTW prefix the GUI-side tasks paste (Tasks/CarrierDetector.cpp:58-75); FFTW itself is not part of this shim:
 * the GUI links fftw3f on its own (SigDigger.pro:486), and like upstream's types.h this header pulls <fftw3.h> in when
 * the build has it */
#define SU_FFTW(method) fftwf##method
#if defined(__has_include)
#  if __has_include(<fftw3.h>)
#    include <fftw3.h>
#  endif
#endif

#ifdef __cplusplus
extern "C" {
#endif
/* App/Loader.cpp:46: FFTW wisdom for the sizes sigutils plans; nothing to plan here (twiddle tables are built with
 * the engine), always SU_TRUE */
SUBOOL su_lib_gen_wisdom(void);
#ifdef __cplusplus
}
#endif

#endif /* _SIGUTILS_TYPES_H */
