/* sigutils/sampling.h -- normalised frequency conventions (shim): a normalised frequency
 * is a fraction of Nyquist (fnor = 2 f / fs), an angular one is radians per sample (omega = pi fnor), a normalised
 * baud rate is symbols per sample.  Used at Tasks/WaveSampler.cpp:48-51, Components/TimeWindow.cpp:1984-1989,
 * Tasks/LPFTask.cpp:64. */
#ifndef _SIGUTILS_SAMPLING_H
#define _SIGUTILS_SAMPLING_H
#include <sigutils/types.h>

#define SU_ABS2NORM_FREQ(fs, f)     (2 * (SUFLOAT) (f) / (SUFLOAT) (fs))
#define SU_NORM2ABS_FREQ(fs, fnor)  ((SUFLOAT) (fs) * (SUFLOAT) (fnor) / 2.f)
#define SU_NORM2ANG_FREQ(fnor)      ((SUFLOAT) PI * (fnor))
#define SU_ANG2NORM_FREQ(omega)     ((omega) / (SUFLOAT) PI)
#define SU_ABS2NORM_BAUD(fs, baud)  ((SUFLOAT) (baud) / (SUFLOAT) (fs))
#define SU_NORM2ABS_BAUD(fs, bnor)  ((SUFLOAT) (fs) * (SUFLOAT) (bnor))
#define SU_T2N(fs, t)               ((unsigned int) floorf((t) * (SUFLOAT) (fs)))
#define SU_T2N_FLOAT(fs, t)         ((t) * (SUFLOAT) (fs))

#endif
