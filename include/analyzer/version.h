/* analyzer/version.h -- API version of the suscan-named shim (main.cpp:158-168) */
#ifndef _SUSCAN_VERSION_H
#define _SUSCAN_VERSION_H
#define SUSCAN_VERSION_MAJOR 0
#define SUSCAN_VERSION_MINOR 3
#define SUSCAN_VERSION_PATCH 0
#ifdef __cplusplus
extern "C" {
#endif
unsigned int suscan_abi_version(void);
const char *suscan_api_version(void);
const char *suscan_pkgversion(void);
#ifdef __cplusplus
}
#endif
#endif
