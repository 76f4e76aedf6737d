/* sigutils/version.h -- API version of the shim (main.cpp:158-168 prints it) */
#ifndef _SIGUTILS_VERSION_H
#define _SIGUTILS_VERSION_H
#define SIGUTILS_VERSION_MAJOR 0
#define SIGUTILS_VERSION_MINOR 3
#define SIGUTILS_VERSION_PATCH 0
#define SU_VER(a, b, c) (((a) << 16) | ((b) << 8) | (c))
#define SIGUTILS_VERSION SU_VER(SIGUTILS_VERSION_MAJOR, SIGUTILS_VERSION_MINOR, SIGUTILS_VERSION_PATCH)
#ifdef __cplusplus
extern "C" {
#endif
unsigned int sigutils_abi_version(void);
const char *sigutils_api_version(void);
const char *sigutils_pkgversion(void);
#ifdef __cplusplus
}
#endif
#endif
