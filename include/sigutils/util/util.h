/* sigutils/util/util.h -- the few helper macros the reference's wrappers expect next to the types */
#ifndef _SIGUTILS_UTIL_H
#define _SIGUTILS_UTIL_H
#include <sigutils/types.h>
#include <stdlib.h>
#define SU_DISPOSE(type, ptr) do { if ((ptr) != NULL) { type##_destroy(ptr); (ptr) = NULL; } } while (0)
#define SU_TRYCATCH(expr, action) do { if (!(expr)) { action; } } while (0)
#endif
