/* analyzer/inspector/params.h -- the inspector enums live in analyzer/msg.h of this shim */
#ifndef _SUSCAN_INSPECTOR_PARAMS_H
#define _SUSCAN_INSPECTOR_PARAMS_H
#include <analyzer/msg.h>
#endif
