// ref_shim: the two accessors Misc/Averager.cpp uses of Suscan::PSDMessage (Suscan/Messages/PSDMessage.cpp:41-60)
#ifndef REF_SHIM_PSDMESSAGE_H
#define REF_SHIM_PSDMESSAGE_H
#include <sigutils/types.h>
#include <stdexcept>
#include <string>
#include <cstring>
namespace Suscan {
#ifndef REF_SHIM_EXCEPTION
#define REF_SHIM_EXCEPTION
  class Exception : public std::runtime_error { public: Exception(std::string const &w) : std::runtime_error(w) {} };
#endif
  class PSDMessage {
    const SUFLOAT *data; unsigned int n;
  public:
    PSDMessage(const SUFLOAT *d, unsigned int n_) : data(d), n(n_) {}
    unsigned int size(void) const { return n; }
    const SUFLOAT *get(void) const { return data; }
  };
}
#endif
