// ref_shim: SuWidgets' Decider is absent; WaveSampler only hands it the block it produced (Tasks/WaveSampler.cpp:317).
// The glue compares the soft samples of `block`, so decide() does nothing here.
#ifndef REF_SHIM_DECIDER_H
#define REF_SHIM_DECIDER_H
#include <sigutils/types.h>
#include <cstdint>
#include <cstddef>
typedef uint8_t Symbol;
class Decider { public: void decide(const SUCOMPLEX *, Symbol *, size_t) const {} };
#endif
