// ref_shim: declarations Scanner.h needs to parse; nothing of the analyzer wrapper is compiled (only SpectrumView is)
#ifndef REF_SHIM_ANALYZER_H
#define REF_SHIM_ANALYZER_H
#include <QObject>
#include <sigutils/types.h>
namespace Suscan {
  class PSDMessage;
  class Source { public: class Config {}; };
  class Analyzer {
  public:
    enum SweepStrategy { STOCHASTIC, PROGRESSIVE };
    enum SpectrumPartitioning { DISCRETE, CONTINUOUS };
  };
}
#endif
