// ref_shim: the reference's Suscan/Library.h drags in the whole suscan wrapper; the DSP units only need the types
#ifndef REF_SHIM_LIBRARY_H
#define REF_SHIM_LIBRARY_H
#include <sigutils/types.h>
#include <stdexcept>
#include <string>
namespace Suscan {
#ifndef REF_SHIM_EXCEPTION
#define REF_SHIM_EXCEPTION
  class Exception : public std::runtime_error { public: Exception(std::string const &w) : std::runtime_error(w) {} };
#endif
}
#define SU_ATTEMPT(expr) do { if (!(expr)) throw std::runtime_error(#expr); } while (0)
#endif
