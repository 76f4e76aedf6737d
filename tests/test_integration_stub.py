"""CPU: the reference-side binding INTEGRATION.md shows (section 1, a back-end TU on the native API) is compiled
against include/sigdigger_b200.h, with two-line stand-ins for the reference's own types it mentions
(Suscan::AnalyzerParams, Channel, Handle, SU_ATTEMPT): the documented calls match the header's signatures."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

MOCKS = r'''
#include <cmath>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <complex>
typedef std::complex<float> SUCOMPLEX;
#define SU_ATTEMPT(expr) do { if (!(expr)) throw std::runtime_error(#expr); } while (0)
struct su_detector_params_mock { unsigned window_size; int window; };
struct suscan_analyzer_params_mock { su_detector_params_mock detector_params; };
namespace Suscan {
  typedef int32_t Handle;
  struct Channel { double fc, fLow, fHigh; };
  struct AnalyzerParams { suscan_analyzer_params_mock c; suscan_analyzer_params_mock &getCParams() { return c; } };
  static inline int classId(std::string const &cls) { return cls == "psk" ? SDB_INSP_PSK : SDB_INSP_RAW; }
}
'''


def test_integration_stub_compiles(tmp_path):
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    m = re.search(r"```cpp\n(// Suscan/B200Backend\.cpp.*?)```", text, flags=re.S)
    assert m, "the binding stub of INTEGRATION.md section 1 is gone"
    code = m.group(1).replace("#include <Suscan/Analyzer.h>", MOCKS)
    src = tmp_path / "B200Backend.cpp"
    src.write_text(code + "\nint main() { return 0; }\n")
    r = subprocess.run(["g++", "-std=c++17", "-Wall", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), str(src)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
