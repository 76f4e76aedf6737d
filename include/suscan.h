/* suscan.h -- umbrella header of the suscan-named shim.  Included by Suscan/Library.cpp:24 (lifecycle:
 * suscan_sigutils_init, suscan_init_{sources,estimators,spectsrcs,inspectors}, Suscan/Library.cpp:97-200) and by
 * Default/GenericInspector/InspectorUI.cpp:43.  Everything it brings in is declared in the headers below; the parts of
 * upstream's suscan.h that serve the configuration database, device discovery and plug-ins are out of scope
 *  and absent. */
#ifndef _SUSCAN_H
#define _SUSCAN_H
#include <sigutils/types.h>
#include <sigutils/version.h>
#include <analyzer/version.h>
#include <analyzer/mq.h>
#include <analyzer/msg.h>
#include <analyzer/source.h>
#include <analyzer/inspector/params.h>
#include <analyzer/analyzer.h>
#endif
